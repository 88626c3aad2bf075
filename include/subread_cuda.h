/* include/subread_cuda.h -- C ABI of libsubread_cuda.so
 *
 * B200-native (sm_100a) replacement for the seed-and-vote step of subread-align / subjunc.
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference's src/).  Plain pointers and sizes only; no C++ or torch types.  All functions
 * return 0 on success and a non-zero code on failure (the reference's own convention,
 * core.c:906-914); the message is available from srcuda_last_error().  Nothing here ever
 * calls exit() and there is no CPU fallback: without a usable GPU every compute entry point
 * fails with SRCUDA_ERR_NO_DEVICE.
 *
 * Threading model: the reference runs do_voting on up to 40 pthreads that share one input
 * mutex (core.c:3404-3456).  Here one host thread drives one handle (= one GPU); several
 * handles (one per GPU) may be driven concurrently from different threads or processes on
 * disjoint fragment ranges of the same chunk.
 */
#ifndef SUBREAD_CUDA_H
#define SUBREAD_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRCUDA_OK 0
#define SRCUDA_ERR_NO_DEVICE 1
#define SRCUDA_ERR_CUDA 2
#define SRCUDA_ERR_IO 3
#define SRCUDA_ERR_FORMAT 4
#define SRCUDA_ERR_ARG 5
#define SRCUDA_ERR_UNSUPPORTED 6
#define SRCUDA_ERR_NOMEM 7
#define SRCUDA_ERR_NO_INDEX 8

typedef struct srcuda_handle srcuda_handle;

/* Byte-identical to the reference's mapping_result_t (core.h:350-370): 68 bytes.  This is the
 * bigtable element (core-bigtable.c:84-125); out_align buffers can be the reference's own
 * bigtable_cache_malloc_ptr. */
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
} srcuda_mapping_result;

/* Byte-identical to subjunc_result_t (core.h:397-410): 16 bytes. */
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
} srcuda_subjunc_result;

/* The scalars do_voting (core.c:3049-3235) and process_voting_junction_PE_topK
 * (core-junction.c:2199-2530) read from configuration_t (core.h:128-253), same names. */
typedef struct {
	int32_t total_subreads;                  /* -n; 10 subread-align, 14 subjunc */
	int32_t top_scores;                      /* 3 */
	int32_t max_vote_simples;                /* 3 SE / 64 PE (core.c:4075-4084) */
	int32_t max_vote_combinations;           /* 3 */
	int32_t max_vote_number_cutoff;          /* 2 */
	int32_t multi_best_reads;                /* 3 */
	int32_t minimum_subread_for_first_read;  /* -m; 3 */
	int32_t minimum_subread_for_second_read; /* -p; 1 */
	int32_t max_indel_length;                /* -I; 5.  Voting uses min(16, this), core.c:3059 */
	int32_t minimum_pair_distance;           /* -d; 50 */
	int32_t maximum_pair_distance;           /* -D; 600 */
	int32_t is_first_read_reversed;          /* -S; 0 */
	int32_t is_second_read_reversed;         /* -S; 1 */
	int32_t do_breakpoint_detection;         /* subjunc */
	int32_t do_big_margin_filtering_for_junctions;
	int32_t big_margin_record_size;          /* 9 */
	int32_t maximum_intron_length;           /* 500000 */
	int32_t limited_tree_scan;
	int32_t check_donor_at_junctions;        /* 1 */
	int32_t more_accurate_fusions;           /* 0 unless fusion detection (core-interface-subjunc.c:671) */
	int32_t max_insertion_at_junctions;      /* 0; other values are not supported */
	int32_t prefer_donor_receptor_junctions; /* 1 */
	int32_t reserved[2];
} srcuda_params;

/* What the host needs back from the loaded block to compute low/high border exactly as
 * core.c:3081-3082 does, and to size its loops. */
typedef struct {
	int32_t index_gap;            /* gehash_t.index_gap  (subread.h:300-310) */
	int32_t padding;              /* gehash_t.padding */
	uint32_t buckets_number;
	uint64_t n_items;
	uint32_t start_point;         /* gene_value_index_t (subread.h:313-322) */
	uint32_t length;
	uint32_t start_base_offset;
	uint32_t values_bytes;
	int32_t total_offsets;        /* gene_offset_t.total_offsets (subread.h:395-401) */
	uint64_t device_bytes;        /* HBM held by this block */
	uint32_t accelerators;        /* lookup accelerators derived at load time: 1 Bloom words, 2 bucket records,
	                                 4 direct filter (dense tables) */
	uint32_t _pad;
} srcuda_index_desc;

/* Device-side timing and counters of the last srcuda_vote_chunk* call (CUDA events on the
 * library's own stream). */
typedef struct {
	float ms_total;          /* first H2D to last D2H (or kernels only for the device variant) */
	float ms_kernel_vote;    /* k_map_simple / k_map_simple_pe + k_map_warp (pack+lookup+vote+topK(+verify)), summed over batches */
	float ms_kernel_overflow;/* exact slow path for fragments whose vote table left shared memory */
	uint64_t n_overflow;     /* fragments that took the slow path */
	uint64_t h2d_bytes;
	uint64_t d2h_bytes;
	uint32_t n_launches;     /* kernels launched */
	uint32_t n_batches;
	float ms_kernel_simple;  /* part of ms_kernel_vote spent in the one-thread-per-fragment tier (k_map_simple, k_map_simple_pe) */
	uint32_t _pad;
	uint64_t n_simple_punted;/* fragments that tier handed on to the warp kernel */
	uint64_t n_gather_requests; /* random index fetches the one-thread-per-fragment tier issued that miss L2 by construction (bucket
	                             * records, direct-filter words, direct_first entries, positions): one DRAM line each, the numerator
	                             * of the line-rate ("HBM-gather") roofline.  The Bloom words (L2-resident) are not counted. */
	uint64_t n_simple_indel;    /* single-end fragments whose result came from that tier's one-indel closed form */
} srcuda_timing;

/* ---- life cycle ------------------------------------------------------------------------ */

/* Number of visible CUDA devices (0 when there is none); never fails. */
int srcuda_device_count(void);

/* Bind a handle to one GPU.  Replaces nothing in the reference (it has no device). */
int srcuda_open(int device, srcuda_handle **out);
void srcuda_close(srcuda_handle *h);

/* Thread-safe: message of the last failure on this handle (h may be NULL for open errors).
 * Replaces the SUBREADprintf diagnostics (subread.h:143). */
const char *srcuda_last_error(const srcuda_handle *h);

/* ---- index ----------------------------------------------------------------------------- */

/* Load one index block from the reference's on-disk files and keep it resident in HBM.
 * Replaces gehash_load (sorted-hashtable.c:1390), gvindex_load (gene-value-index.c:190) and
 * load_offsets (gene-algorithms.c:1293) for block `block_no` of `index_prefix`
 * (<prefix>.NN.b.tab, <prefix>.NN.b.array, <prefix>.reads). */
int srcuda_index_load(srcuda_handle *h, const char *index_prefix, int block_no);

/* Same, from structures the caller already holds in host memory (what a reference-side
 * binding would pass from gehash_t / gene_value_index_t / gene_offset_t).  bucket_start has
 * buckets_number+1 entries; keys/vals are the buckets' items back to back. */
int srcuda_index_upload(srcuda_handle *h, int32_t index_gap, int32_t padding, uint32_t buckets_number,
                        uint64_t n_items, const uint32_t *bucket_start, const int16_t *keys, const uint32_t *vals,
                        uint32_t start_point, uint32_t length, const uint8_t *values, uint32_t values_bytes,
                        const uint32_t *read_offsets, int32_t total_offsets);

/* Copy the resident block of `src` to the GPU of `dst` with a peer copy over NVLink
 * (no host round trip).  Both handles live in this process. */
int srcuda_index_replicate(srcuda_handle *dst, const srcuda_handle *src);

/* The same replication when every GPU is driven by its own process (one process per GPU, as
 * bench.py launches them): the owner exports CUDA IPC handles of its resident block into an
 * opaque blob, the other processes import it -- the block is copied device to device over
 * NVLink and the mapping is closed again.  The exporter must keep the block loaded until every
 * importer returned.  SRCUDA_IPC_BLOB_BYTES bytes are enough for the blob. */
#define SRCUDA_IPC_BLOB_BYTES 512
int srcuda_index_export_ipc(const srcuda_handle *h, void *blob, size_t blob_bytes);
int srcuda_index_import_ipc(srcuda_handle *h, const void *blob, size_t blob_bytes);

int srcuda_index_describe(const srcuda_handle *h, srcuda_index_desc *out);
void srcuda_index_unload(srcuda_handle *h);

/* ---- pinned host memory (optional, recommended for the chunk buffers) ------------------- */
void *srcuda_host_alloc(size_t bytes);
void srcuda_host_free(void *p);

/* ---- FASTQ batches (host side, no device work) ------------------------------------------ */

#define SRCUDA_FASTQ_MORE 0 /* the buffer ran out inside or between records: refill from buf + consumed and call again */
#define SRCUDA_FASTQ_FULL 1 /* max_reads records were indexed */
#define SRCUDA_FASTQ_END 2  /* end of the input (at_eof), or a read that is empty after trimming: the reference ends its
                             * input there (fetch_next_read_pair, core.c:1168-1208) */
typedef struct {
	int64_t n_reads;       /* records indexed by this call */
	uint64_t consumed;     /* bytes of buf those records span; the next call starts at buf + consumed */
	int32_t stop;          /* SRCUDA_FASTQ_* */
	uint32_t max_len;      /* longest read indexed (max_read_len of srcuda_vote_chunk_device) */
	uint64_t error_offset; /* with SRCUDA_ERR_FORMAT: offset of the byte the reference would report */
} srcuda_fastq_batch;

/* Indexes every complete record of a buffer of plain FASTQ text: the FASTQ branch of geinput_next_read_trim
 * (input-files.c:982-1093; read_line_noempty :209, SKIP_LINE :650, trim_read_inner :716) for a whole batch at once
 * instead of one call per read under the input mutex (core.c:1130).  Nothing is copied: read i of the call gets
 * off[i*stride] = base_offset + its offset in buf and len[i*stride] (after trimming), so that with text = the
 * caller's chunk buffer, base_offset = buf - text and stride = 1 the arrays are srcuda_vote_chunk's (single-end);
 * stride = 2 interleaves the mates of two files (R2 written through off + 1, len + 1) for srcuda_reads_pack or a
 * device-side gather.  name_off/name_len (cut at the first blank, at most 199 characters) and qual_off/qual_len are
 * optional (NULL).  at_eof = 0: a record the buffer cuts short is left for the next call; at_eof != 0: the end of
 * the buffer ends the last line, as EOF does.  max_reads bounds the arrays.  A record that does not start with '@'
 * or lacks its '+' line returns SRCUDA_ERR_FORMAT (the reference prints "a format issue" and ends the input, :1003,
 * :1041); the records in front of it are valid in `out`.  gzip input is inflated by the caller (seek-zlib.c). */
int srcuda_fastq_split(const char *buf, uint64_t n_bytes, int at_eof, int trim_5, int trim_3, int64_t max_reads,
                       uint64_t base_offset, int64_t stride, uint64_t *off, uint16_t *len, uint64_t *name_off,
                       uint16_t *name_len, uint64_t *qual_off, uint16_t *qual_len, srcuda_fastq_batch *out);

/* The same result from n_threads host threads: the buffer is cut at guessed record starts, every piece is parsed as
 * an input of its own, and the pieces are accepted only if each one ends exactly where the next begins -- otherwise
 * (blank-line oddities, broken records, a trimmed-empty read) the single-threaded loop above decides, so the outcome
 * never depends on n_threads. */
int srcuda_fastq_split_mt(const char *buf, uint64_t n_bytes, int at_eof, int trim_5, int trim_3, int64_t max_reads,
                          uint64_t base_offset, int64_t stride, uint64_t *off, uint16_t *len, uint64_t *name_off,
                          uint16_t *name_len, uint64_t *qual_off, uint16_t *qual_len, int n_threads, srcuda_fastq_batch *out);

/* Gathers indexed reads into the dense chunk layout of srcuda_vote_chunk: fragment f's mates back to back
 * (read r = f*mates + m), off[r] non-decreasing from 0, *text_bytes = bytes written.  buf2 = NULL: single-end.
 * The input half of fetch_next_read_pair (core.c:1121-1175) for a batch; mate strands stay as they are in the files
 * (the voting step applies is_first/second_read_reversed itself). */
int srcuda_reads_pack(const char *buf1, const uint64_t *off1, const uint16_t *len1, const char *buf2,
                      const uint64_t *off2, const uint16_t *len2, int64_t n_frags, char *text, uint64_t text_capacity,
                      uint64_t *off, uint16_t *len, uint64_t *text_bytes);
/* The same gather from n_threads host threads (sizes first, then disjoint stretches of text). */
int srcuda_reads_pack_mt(const char *buf1, const uint64_t *off1, const uint16_t *len1, const char *buf2,
                         const uint64_t *off2, const uint16_t *len2, int64_t n_frags, char *text, uint64_t text_capacity,
                         uint64_t *off, uint16_t *len, int n_threads, uint64_t *text_bytes);

/* ---- chunk reader (host side, no device work) ------------------------------------------- */

typedef struct srcuda_reader srcuda_reader;

/* Opens one (path2 = NULL) or two FASTQ files, plain or gzip: geinput_open (input-files.c:455) for the FASTQ types.
 * The files are read in windows of window_bytes (0: 16 MB), each indexed by srcuda_fastq_split_mt with n_threads.
 * *out is set even on failure (srcuda_reader_error says why; close it). */
int srcuda_reader_open(const char *path1, const char *path2, uint64_t window_bytes, int n_threads, int trim_5, int trim_3,
                       srcuda_reader **out);
/* The next batch of at most max_frags fragments, gathered into the caller's buffers in the dense chunk layout of
 * srcuda_vote_chunk (read r = f*mates + m, off[r] from 0): the read loop of fetch_next_read_pair (core.c:1121-1215) up
 * to reads_per_chunk, for a batch.  off and len need max_frags * mates entries.  *n_frags = 0: end of input.  Two files
 * that do not hold the same number of reads are an error where the reference prints one (core.c:1177-1182); a record
 * without '@' or '+' is SRCUDA_ERR_FORMAT after the records in front of it have been handed out. */
int srcuda_reader_next(srcuda_reader *r, int64_t max_frags, char *text, uint64_t text_capacity, uint64_t *off, uint16_t *len,
                       int64_t *n_frags, uint64_t *text_bytes);
/* The same batch with the names (cut at the first blank, at most 199 characters, not terminated) and the quality lines (trimmed like
 * the reads; not terminated) of its reads gathered as well: read k's name is names[name_off[k] .. + name_len[k]), its qualities
 * quals[qual_off[k] .. + qual_len[k]).  What a caller needs to stand in for fetch_next_read_pair (core.c:1121) in every pass over a
 * chunk, not only the voting step.  Capacities too small for the batch are SRCUDA_ERR_ARG. */
int srcuda_reader_next_full(srcuda_reader *r, int64_t max_frags, char *text, uint64_t text_capacity, uint64_t *off, uint16_t *len,
                            char *names, uint64_t names_capacity, uint64_t *name_off, uint16_t *name_len, char *quals,
                            uint64_t quals_capacity, uint64_t *qual_off, uint16_t *qual_len, int64_t *n_frags, uint64_t *text_bytes);
/* Passes over the next n_frags fragments without gathering them (*skipped = how many there were): a process that maps
 * one slice of every chunk (one GPU of several) skips the slices of the others. */
int srcuda_reader_skip(srcuda_reader *r, int64_t n_frags, int64_t *skipped);
/* File offsets (in inflated bytes) behind the last fragment handed out: where a reader of its own -- the reference's gene_input_t
 * under a binding -- has to continue (geinput_seek, input-files.c:699: a plain FASTQ file's position is its byte offset).  pos2 may
 * be NULL for single-end input. */
int srcuda_reader_tell(const srcuda_reader *r, uint64_t *pos1, uint64_t *pos2);
const char *srcuda_reader_error(const srcuda_reader *r);
void srcuda_reader_close(srcuda_reader *r);

/* ---- the voting step -------------------------------------------------------------------- */

/* One chunk of reads against the resident index block: everything do_voting (core.c:3049)
 * does per read up to and including the bigtable update, i.e. genekey2int / reverse_read
 * (input-files.c:1232,1113), gehash_go_X (sorted-hashtable.c:937) and
 * process_voting_junction_PE_topK (core-junction.c:2199).  find_new_indels /
 * find_new_junctions stay with the caller.
 *
 *   text, off, len  reads as they are in the input file after trimming; read r of fragment f
 *                   is mate m = r % mates of fragment f = r / mates (mates = 1 + is_paired),
 *                   characters text[off[r]] .. text[off[r]+len[r]-1].  Mate strands are
 *                   handled inside (is_first/second_read_reversed), like fetch_next_read_pair
 *                   (core.c:1187-1199).
 *   out_align       read-modify-write, [(f*mates+m)*multi_best_reads + k], the layout of
 *                   core-bigtable.c:161-166: successive index blocks merge into it.  Pass
 *                   out_align_is_zero != 0 when the buffer is known to be all zero (first
 *                   block of a chunk) to skip its upload.
 *   out_junc, out_big_margin   subjunc only (do_breakpoint_detection); may be NULL otherwise.
 *   out_verify      optional, one uint16 per out_align slot: number of read bases equal to
 *                   the genome at selected_position in the slot's orientation (match_chro,
 *                   gene-value-index.c:856), 0xffff for an unused slot.
 */
int srcuda_vote_chunk(srcuda_handle *h, const srcuda_params *params, const char *text, const uint64_t *off,
                      const uint16_t *len, int64_t n_frags, int is_paired, int out_align_is_zero,
                      srcuda_mapping_result *out_align, srcuda_subjunc_result *out_junc, uint16_t *out_big_margin,
                      uint16_t *out_verify);

/* Same computation with every buffer already resident on the handle's GPU (device pointers);
 * no host<->device traffic.  d_text_bytes is the size of the buffer at d_text: reads whose
 * (off, len) leave it are reported as an error instead of being dereferenced.  With out_align_is_zero != 0 the bigtable slice is not read (it is
 * treated as all zero, the state of a chunk's first index block) and every slot is written.
 * Used by bench.py for the HBM-resident figure. */
int srcuda_vote_chunk_device(srcuda_handle *h, const srcuda_params *params, const char *d_text, uint64_t d_text_bytes,
                             const uint64_t *d_off, const uint16_t *d_len, int64_t n_frags, int is_paired, int max_read_len, int out_align_is_zero,
                             srcuda_mapping_result *d_out_align, srcuda_subjunc_result *d_out_junc,
                             uint16_t *d_out_big_margin, uint16_t *d_out_verify);

/* One chunk over several GPUs: the chunk-level driver the reference's run_maybe_threads(STEP_VOTING) is (core.c:3404-3456:
 * one chunk handed to N workers that write disjoint slots of one bigtable).  The fragment range [0, n_frags) is cut into
 * n_handles contiguous slices whose sizes differ by at most one (slice k = [k*base + min(k, extra), ...), SURVEY section
 * 8e); one host thread per handle runs srcuda_vote_chunk on its slice, reading the caller's text/off/len in place and
 * writing its own part of out_align (and out_junc, out_big_margin, out_verify).  No collective and no device-to-device
 * traffic on the data path.  Every handle must hold the same index block (srcuda_index_load on one of them and
 * srcuda_index_replicate to the others, or srcuda_index_load_all).  Returns the first failure; its message is on
 * handles[0].  Per-GPU timings stay on the handles (srcuda_get_timing). */
int srcuda_vote_chunk_multi(srcuda_handle *const *handles, int n_handles, const srcuda_params *params, const char *text,
                            const uint64_t *off, const uint16_t *len, int64_t n_frags, int is_paired, int out_align_is_zero,
                            srcuda_mapping_result *out_align, srcuda_subjunc_result *out_junc, uint16_t *out_big_margin,
                            uint16_t *out_verify);

/* Load one index block on handles[0] (srcuda_index_load) and replicate it to the other handles' GPUs with peer copies over
 * NVLink (srcuda_index_replicate), all replicas at once: "the index is loaded once and replicated to every GPU". */
int srcuda_index_load_all(srcuda_handle *const *handles, int n_handles, const char *index_prefix, int block_no);

int srcuda_get_timing(const srcuda_handle *h, srcuda_timing *out);

/* Debug/validation switch: route every fragment through the exact slow path (one thread per
 * fragment, full 30x24 table) instead of the fast tiers.  Results must not change.
 * Environment switches read at srcuda_open, all for A/B runs and cross-checks (defaults in brackets):
 *   SRCUDA_SIMPLE_TIER [1]     thread-per-fragment tier (single- and paired-end)
 *   SRCUDA_SIMPLE_TIER_PE [1]  its paired-end form only
 *   SRCUDA_TIER_JUNC [1]       subjunc chunks in that tier
 *   SRCUDA_TIER_INDEL [auto]   its one-indel closed form: 0 never, 1 always, unset: by the last call's counters
 *   SRCUDA_PHASE [1]           the phase table of gapped indexes
 *   SRCUDA_WARP_JUNC [1]       subjunc fragments in the warp kernel (0: scalar path)
 *   SRCUDA_ACCEL [auto]        lookup accelerators: 0 none, 1 Bloom words, 2 bucket records, 3 both, 4 direct filter
 *   SRCUDA_PE_CW3 [1]          paired-end tier on dense tables: three-word vote clusters for reads of at most 160 bases
 *   SRCUDA_PE_POOL [auto]      ... its cluster pool per strand pass (both mates); SRCUDA_PE_PAD_SMEM: extra shared memory per block
 *   SRCUDA_COMPACT_D2H [1]     host-buffer entry point, first index block: results cross PCIe compacted and are scattered into the
 *                              caller's table by SRCUDA_EXPAND_THREADS [host threads / GPUs] host threads
 *   SRCUDA_BATCH_FRAGS [auto]  fragments per batch of a chunk (default: at least eight batches, 256 Ki to 1 Mi reads each)
 *   SRCUDA_SCALAR_BLOCKS, SRCUDA_L2_PERSIST, SRCUDA_L2_FETCH   tuning knobs */
int srcuda_set_force_slow_path(srcuda_handle *h, int on);

/* Measured random 32-byte-sector gather bandwidth over `table_bytes` of HBM (GB/s): the
 * "HBM-gather roofline" denominator of SURVEY.md section 8(d). */
int srcuda_measure_gather_bandwidth(srcuda_handle *h, uint64_t table_bytes, int iters, double *gbps_out);

/* The lookup's memory roofline for a mix of accesses, measured without any compute: units of work per second when every unit
 * makes n_small random 2-byte probes of a table of small_bytes that stays in L2 (the Bloom words' load and cache policy) and
 * n_big random 32-byte fetches from a table of big_bytes (the bucket records' load: 64 bytes per DRAM miss).  With the counts
 * of the running workload (lookups and gather requests per fragment) this is the fragment rate the memory system allows: the
 * denominator of bench.py's gather fraction. */
int srcuda_measure_lookup_roofline(srcuda_handle *h, uint64_t small_bytes, int n_small, uint64_t big_bytes, int n_big, int iters,
                                   double *units_per_s);

#ifdef __cplusplus
}
#endif
#endif
