/* oracle/sr_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's seed-and-vote read-mapping step: what
 * `do_voting` (reference src/core.c:3049-3323) computes for one chunk of reads against one
 * index block, up to and including the bigtable update done by
 * `process_voting_junction_PE_topK` (src/core-junction.c:2199-2530).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * The product (subread_b200/csrc, libsubread_cuda.so) never includes or links it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks this restatement against
 * bigtables dumped from the unmodified reference (oracle/_ref/ref_dump, see oracle/ref_hooks.c)
 * and tests/test_oracle_golden.py against committed fixtures made by the same tool
 * (tests/golden/make_golden.py).
 */
#ifndef SR_ORACLE_H
#define SR_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same byte layout as the reference's mapping_result_t (src/core.h:350-370), 68 bytes. */
typedef struct {
	uint32_t selected_position;
	int16_t result_flags;
	int16_t read_length;
	int16_t selected_votes;
	int16_t used_subreads_in_vote;
	uint8_t noninformative_subreads_in_vote;
	int8_t indels_in_confident_coverage;
	int8_t is_fully_covered;
	uint8_t _pad0;
	int16_t selected_indel_record[22];
	uint16_t confident_coverage_start;
	uint16_t confident_coverage_end;
	int16_t subread_quality;
	uint16_t _pad1;
} or_mapping_result_t;

/* Same byte layout as subjunc_result_t (src/core.h:397-410), 16 bytes. */
typedef struct {
	int16_t split_point;
	int16_t minor_votes;
	int8_t double_indel_offset;
	int8_t indel_at_junction;
	int8_t small_side_increasing_coordinate;
	int8_t large_side_increasing_coordinate;
	uint32_t minor_position;
	uint16_t minor_coverage_start;
	uint16_t minor_coverage_end;
} or_subjunc_result_t;

/* In-memory index block: the three on-disk files of one block. */
typedef struct {
	/* <prefix>.NN.b.tab  -- gehash_load, src/sorted-hashtable.c:1390-1568 */
	int32_t index_gap;
	int32_t padding;
	uint32_t buckets_number;
	uint64_t n_items;
	uint32_t *bucket_start; /* buckets_number+1 */
	int16_t *keys;          /* n_items, ascending inside a bucket */
	uint32_t *vals;         /* n_items */
	/* <prefix>.NN.b.array -- gvindex_load, src/gene-value-index.c:190-228 */
	uint32_t start_point, length, start_base_offset, values_bytes;
	uint8_t *values;
	/* <prefix>.reads      -- load_offsets, src/gene-algorithms.c:1293 */
	int32_t total_offsets;
	uint32_t *read_offsets; /* cumulative END offsets, one per chromosome */
} or_index_t;

/* The scalars do_voting / topK read from configuration_t (src/core.h:128-253). */
typedef struct {
	int32_t total_subreads;                  /* 10 subread-align, 14 subjunc */
	int32_t top_scores;                      /* 3 */
	int32_t max_vote_simples;                /* 3 SE, 64 PE (core.c:4075-4084) */
	int32_t max_vote_combinations;           /* 3 */
	int32_t max_vote_number_cutoff;          /* 2 */
	int32_t multi_best_reads;                /* 3 */
	int32_t minimum_subread_for_first_read;  /* 3 (subjunc 1) */
	int32_t minimum_subread_for_second_read; /* 1 */
	int32_t max_indel_length;                /* 5; voting uses min(16, this) core.c:3059 */
	int32_t minimum_pair_distance;           /* 50 */
	int32_t maximum_pair_distance;           /* 600 */
	int32_t is_first_read_reversed;          /* 0 */
	int32_t is_second_read_reversed;         /* 1 */
	int32_t do_breakpoint_detection;         /* subjunc: 1 */
	int32_t do_big_margin_filtering_for_junctions;
	int32_t big_margin_record_size;          /* 9 */
	int32_t maximum_intron_length;           /* 500000 */
	int32_t limited_tree_scan;               /* 0 */
	int32_t check_donor_at_junctions;        /* 1 */
	int32_t more_accurate_fusions;           /* 0 unless fusion detection (core-interface-subjunc.c:671) */
	int32_t max_insertion_at_junctions;      /* 0; other values are not supported */
	int32_t prefer_donor_receptor_junctions; /* 1 */
	int32_t reserved[2];
} or_params_t;

/* Exact algorithmic-byte accounting gathered while voting. */
typedef struct {
	uint64_t lookups;        /* gehash_go_X calls */
	uint64_t probes;         /* i16 keys inspected (binary search + run scan) */
	uint64_t hits;           /* u32 positions consumed */
	uint64_t read_bytes;     /* read characters in */
	uint64_t verify_bytes;   /* genome bytes a K<=3 verification touches */
	uint64_t result_bytes;   /* mapping_result_t bytes out */
} or_stats_t;

void or_default_params(or_params_t *p, int is_paired, int is_subjunc);

int or_index_load(or_index_t *idx, const char *tab_path, const char *array_path, const char *reads_path);
void or_index_free(or_index_t *idx);

/* 16 characters -> key, genekey2int (src/input-files.c:1232-1251). */
uint32_t or_genekey2int(const char *s);
/* in-place reverse complement, reverse_read (src/input-files.c:1113-1188, base space). */
void or_reverse_read(char *s, int len);
/* match_chro (src/gene-value-index.c:856-959, base space). */
int or_match_chro(const char *read, const or_index_t *idx, uint32_t pos, int test_len, int is_negative_strand);
/* subread schedule (src/core.c:3117-3130,3169-3171): fills offsets[], returns number of lookups. */
int or_subread_schedule(int rlen, int gap, int total_subreads, int *applied_subreads, int *offsets, int max_out);

/* One chunk through the voting step against one index block.
 * Reads are raw (as in the FASTQ, trimmed); read r of fragment f is text+off[f*mates+m], len[...].
 * out_align is read-modify-write, laid out [(frag*mates+mate)*multi_best_reads + k]
 * (src/core-bigtable.c:161-166).  out_verify (optional) gets, per result slot, the number of
 * read bases that match the genome at the selected position (match_chro of the whole read in the
 * slot's orientation) or 0xffff for an unused slot.
 * Returns 0, or non-zero on invalid arguments. */
int or_vote_chunk(const or_index_t *idx, const or_params_t *p, const char *text, const uint64_t *off,
                  const uint16_t *len, int64_t n_frags, int is_paired, or_mapping_result_t *out_align,
                  or_subjunc_result_t *out_junc, uint16_t *out_big_margin, uint16_t *out_verify,
                  or_stats_t *stats);

/* Multi-threaded driver over disjoint fragment ranges (for the cpu_baseline leg only). */
int or_vote_chunk_mt(const or_index_t *idx, const or_params_t *p, const char *text, const uint64_t *off,
                     const uint16_t *len, int64_t n_frags, int is_paired, or_mapping_result_t *out_align,
                     or_subjunc_result_t *out_junc, uint16_t *out_big_margin, uint16_t *out_verify,
                     or_stats_t *stats, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
