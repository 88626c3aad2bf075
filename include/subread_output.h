/* include/subread_output.h -- C ABI of the output side of libsubread_cuda.so
 *
 * Host-side replacement for the reference's SAM / BAM writer: write_single_fragment (core.c:1888-2170), calc_flags / calc_tlen
 * (core.c:1635-1800), calc_edit_dist (core.c:1325) and the BGZF writer of sambam-file.c (SamBam_writer_create :1097 ...
 * SamBam_writer_close :1310).  Paths are relative to the reference's src/.  Plain pointers and sizes only.
 *
 * What is kept: the bytes.  A BAM file written here equals, byte for byte and compressed blocks included, the file the
 * reference writes through its single stream (thread_no < 0: -T 1, or any -T with --keepReadOrder), because records are
 * encoded the same way, blocks are cut where the reference cuts them (a block ends behind the first committable record that
 * takes it past 55 000 bytes, :1793) and every block is deflated by itself with the reference's zlib settings (level 1, raw
 * deflate, :1220).  SAM text equals the reference's fprintf line (core.c:2156).
 * What is different: blocks are independent, so they are deflated by a pool of threads and written in order -- the
 * reference's ordered output compresses under its output lock, one block at a time (core.c:1857-1866), and its parallel
 * output (one buffer per thread, core.c:1851) gives up record order.  Here order and parallel compression come together.
 *
 * Not built: coordinate-sorted output and the .bai index (sort_reads_by_coord, sambam-file.c:1833-2460), colour space.
 * All functions return 0 (SRCUDA_OK) or one of the SRCUDA_ERR_* codes of subread_cuda.h; srcuda_out_error() has the text.
 * One writer is driven by one thread at a time.
 */
#ifndef SUBREAD_OUTPUT_H
#define SUBREAD_OUTPUT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct srcuda_writer srcuda_writer;

#define SRCUDA_OUT_SAM 0
#define SRCUDA_OUT_BAM 1

/* One alignment record, the argument list of SamBam_writer_add_read (sambam-file.c:1704) as a struct.  `pos` and `next_pos`
 * are 1-based as in SAM; `additional` are the optional columns as tab-separated text ("HI:i:1\tNH:i:1\tNM:i:0"), may be
 * NULL or empty; a leading tab is tolerated as the reference tolerates it (:1487). */
typedef struct {
	const char *name;
	uint32_t flags;
	const char *chro;      /* "*" when unmapped */
	uint32_t pos;
	int32_t mapping_quality;
	const char *cigar;     /* "*" when unmapped */
	const char *next_chro; /* "=", "*" or a name */
	uint32_t next_pos;
	int64_t tlen;
	int32_t read_len;
	const char *seq;       /* read_len bases */
	const char *qual;      /* read_len characters, phred + 33 */
	const char *additional;
} srcuda_sam_record;

/* Open an output.  path NULL = stdout.  threads = deflate threads of a BAM output (>= 1).  no_compression != 0 stores the
 * blocks (the reference's fastest_compression of temporary files, sambam-file.h:42).
 * Replaces SamBam_writer_create (sambam-file.c:1097) / f_subr_open(.., "wb") of the SAM output (core.c:4147). */
int srcuda_out_open(const char *path, int format, int threads, int no_compression, srcuda_writer **out);

/* One header line without its newline.  add_chro != 0: an "@SQ" line also enters the chromosome table.
 * Replaces SamBam_writer_add_header (sambam-file.c:1357) and the header fprintf's of core.c:3875-3887. */
int srcuda_out_add_header(srcuda_writer *w, const char *line, int add_chro);

/* A chromosome of the index, in index order (BAM refID = order of insertion).  add_header != 0 also writes its "@SQ" line.
 * Replaces SamBam_writer_add_chromosome (sambam-file.c:1401).  A name given twice is SRCUDA_ERR_ARG (the reference would
 * write a BAM whose records point past its reference list). */
int srcuda_out_add_chromosome(srcuda_writer *w, const char *name, uint32_t length, int add_header);

/* Header complete: BAM writes the magic, the text (blocks cut behind the first newline past 55 000 bytes, :1248-1275) and
 * the reference list.  Replaces SamBam_writer_finish_header (sambam-file.c:1595).  Implied by the first record and by close. */
int srcuda_out_finish_header(srcuda_writer *w);

/* Append one record.  committable = 0 keeps the record in one block with the next one (the reference keeps the two mates of
 * a pair together this way, core.c:2147-2150).  Replaces SamBam_writer_add_read (sambam-file.c:1704) and the record fprintf
 * (core.c:2156-2164). */
int srcuda_out_add_read(srcuda_writer *w, const srcuda_sam_record *rec, int committable);

/* n records at once; records_per_fragment = 1 (every record committable) or 2 (every second one).  Same bytes as n calls. */
int srcuda_out_add_reads(srcuda_writer *w, const srcuda_sam_record *recs, int64_t n, int records_per_fragment);

/* A buffer of complete SAM alignment lines ('\n'-terminated, no header lines): each is parsed like
 * SamBam_writer_add_read_line (sambam-file.c:1668: eleven mandatory fields, the rest are optional columns) and appended.
 * A line with fewer than eleven fields is SRCUDA_ERR_FORMAT. */
int srcuda_out_add_sam_lines(srcuda_writer *w, const char *text, size_t bytes, int records_per_fragment);

/* ---- fragment level: write_single_fragment (core.c:1888) ------------------------------------------------------------- */

/* The scalars of global_context_t the record builder reads. */
typedef struct {
	int32_t is_paired_end;            /* input_reads.is_paired_end_reads */
	int32_t minimum_pair_distance;    /* config.minimum_pair_distance, -d */
	int32_t maximum_pair_distance;    /* config.maximum_pair_distance, -D */
	int32_t is_first_read_reversed;   /* config.is_first_read_reversed, -S */
	int32_t is_second_read_reversed;  /* config.is_second_read_reversed */
	int32_t _pad;
	const char *read_group_id;        /* config.read_group_id, "" or NULL = none; at most 300 characters */
} srcuda_sam_config;

/* One mate as realignment (row f1) hands it over: the fields of subread_output_tmp_t (core.c:1252-1274) that
 * write_single_fragment reads, plus the read itself. */
typedef struct {
	int32_t present;          /* rec != NULL (a realignment result exists for this mate) */
	int32_t ok;               /* is_R1_OK / is_R2_OK (convert_read_to_tmp succeeded) */
	int32_t chro_id;          /* number of rec->chro in the writer's chromosome table; the reference compares the pointers */
	int32_t offset;           /* rec->offset (an int there too), 1-based position in the chromosome */
	int32_t strand;           /* rec->strand */
	int32_t mapping_quality;  /* rec->mapping_quality */
	int32_t mismatched_bases; /* raw_r->final_mismatched_bases; NM = this + the I and D lengths of the CIGAR (core.c:1325) */
	int32_t read_len;
	const char *cigar;        /* rec->cigar */
	const char *additional;   /* rec->additional_information before NM is appended ("" or "\tSB:i:1..."), may be NULL */
	char *name;               /* cut at the first '/' in place (remove_backslash, subread.h:251) */
	char *text;               /* reverse-complemented in place when the record says so (core.c:1947-1964) */
	char *qual;               /* reversed in place likewise; "" = no qualities: filled with 'I' (needs read_len + 1 bytes) */
} srcuda_sam_mate;

/* Counters write_single_fragment / calc_flags keep (thread_context_t, core.h:565-580); added to, never cleared. */
typedef struct {
	uint64_t all_mapped_reads;
	uint64_t all_correct_PE_reads;
	uint64_t not_properly_pairs_different_chro;
	uint64_t not_properly_different_strands;
	uint64_t not_properly_pairs_TLEN_wrong;
	uint64_t not_properly_pairs_wrong_arrangement;
} srcuda_sam_counters;

/* SAM flag of one mate.  Replaces calc_flags (core.c:1635-1716). */
int srcuda_sam_calc_flags(const srcuda_sam_config *cfg, const srcuda_sam_mate *m1, const srcuda_sam_mate *m2, int is_second_read,
                          int current_location, int tlen, srcuda_sam_counters *counters);

/* Template length of two mapped mates on one chromosome.  Replaces calc_tlen (core.c:1719-1790). */
int srcuda_sam_calc_tlen(int32_t offset1, const char *cigar1, int read_len_1, int32_t offset2, const char *cigar2, int read_len_2);

/* One fragment at one of its reported locations -> one (single-end) or two records, appended to the writer in the
 * reference's form: flags, TLEN and its sign, mate columns ("=" for one chromosome), HI / NH / RG / the record's own
 * columns / NM, reads of reverse-strand records reverse-complemented.  m2 is NULL for single-end input.
 * Replaces write_single_fragment (core.c:1888-2170) for base-space reads. */
int srcuda_out_add_fragment(srcuda_writer *w, const srcuda_sam_config *cfg, srcuda_sam_mate *m1, srcuda_sam_mate *m2,
                            int all_locations, int current_location, srcuda_sam_counters *counters);

/* The same in two steps, for callers that build records on several threads (the reference's workers do, core.c:2436):
 * srcuda_out_format_fragment does everything srcuda_out_add_fragment does up to the encoded bytes (SAM text or BAM records of
 * the one or two mates) and may run on any number of threads at once -- it only reads the writer, whose header must be
 * finished; srcuda_out_append hands such bytes to the stream, one caller at a time, in the order the file shall have.
 * *bytes receives the encoded size; SRCUDA_ERR_NOMEM when it exceeds `capacity` (nothing is copied then).  Failures of
 * srcuda_out_format_fragment leave their message in srcuda_out_error(NULL) of the calling thread. */
int srcuda_out_format_fragment(const srcuda_writer *w, const srcuda_sam_config *cfg, srcuda_sam_mate *m1, srcuda_sam_mate *m2,
                               int all_locations, int current_location, srcuda_sam_counters *counters, void *buffer, size_t capacity,
                               size_t *bytes);
int srcuda_out_append(srcuda_writer *w, const void *encoded, size_t bytes, int committable);

/* Everything appended so far is compressed and handed to the file. */
int srcuda_out_flush(srcuda_writer *w);

/* Flush, write the BAM end-of-file block, close.  bytes_written (may be NULL) receives the file size.  The writer is freed
 * whatever the result.  Replaces SamBam_writer_close (sambam-file.c:1310) / fclose of the SAM output (core.c:4276). */
int srcuda_out_close(srcuda_writer *w, uint64_t *bytes_written);

/* Message of the last failure (w may be NULL: failures of srcuda_out_open). */
const char *srcuda_out_error(const srcuda_writer *w);

#ifdef __cplusplus
}
#endif
#endif
