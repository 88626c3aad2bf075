/* integration/output_native.c -- the reference-side binding of the output half of libsubread_cuda.so.
 *
 * What a Subread maintainer adds to let the library write the SAM / BAM file: replacements for write_sam_headers
 * (core.c:3843), write_single_fragment (core.c:1888), the SamBam_writer_* functions core.c calls (sambam-file.c:1097-1797)
 * and the closing step of destroy_global_context (core.c:4270-4292), written against the reference's own headers and calling
 * include/subread_output.h.  Realignment (convert_read_to_tmp, write_realignments_for_fragment) stays the reference's: it hands
 * every fragment to write_single_fragment exactly as before.
 *
 * Built only as a test vehicle (oracle/Makefile: _ref/subread-align-out, _ref/subjunc-out, and linked into the -cuda
 * binaries): the reference's CLI main + this file against libsubread_ref.so, whose cross-function calls go through the PLT,
 * so that the definitions here take the place of the originals without touching a reference source line.  The native path
 * is taken when SRCUDA_NATIVE_OUTPUT=1 is in the environment; otherwise every function forwards to the original (dlsym
 * RTLD_NEXT), and for what the library does not write (coordinate-sorted BAM, colour-space reads) it forwards too.
 * tests/test_output_program.py: SAM and BAM byte-identical to the stock programs' (apart from @PG).
 *
 * Order: SAM output and BAM output with --keepReadOrder are written in input order, as in the reference -- but no worker waits
 * for its turn (hand_over_in_order below; the reference's relay is add_buffered_fragment, core.c:1855-1881); plain multi-thread
 * BAM output is written in the order the workers arrive -- the reference's own order there is per-thread and changes from run
 * to run.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "subread.h"
#include "core.h"
#include "input-files.h"
#include "gene-algorithms.h"
#include "sambam-file.h"

#include "../include/subread_cuda.h"
#include "../include/subread_output.h"

/* subread_output_tmp_t is private to core.c (1252-1274); a binding compiled into core.c would see it.  Same members, same
 * order, sizes from core.h's constants; the byte-for-byte SAM comparison of tests/test_output_program.py is what pins it. */
typedef struct {
	char current_cigar_decompress[CORE_MAX_CIGAR_STR_LEN + 1];
	char cigar[CORE_MAX_CIGAR_STR_LEN];
	unsigned short chimeric_sections;
	unsigned int out_poses[12];
	short out_lens[12];
	char out_cigars[12][60];
	char out_strands[12];
	char additional_information[CORE_ADDITIONAL_INFO_LENGTH + 1];
	mapping_result_t *raw_result;
	unsigned int linear_position;
	short soft_clipping_movements;
	char *chro;
	int offset;
	int strand;
	int is_first_section_jumpped;
	int mapping_quality;
	int is_NM_appied;
} output_tmp_t;

typedef struct {
	int decided, native;
	srcuda_writer *writer;
	pthread_mutex_t lock;
} output_binding_t;

static output_binding_t g_out = {0, 0, NULL, PTHREAD_MUTEX_INITIALIZER};

static int native_output(void)
{
	if (!g_out.decided) {
		const char *e = getenv("SRCUDA_NATIVE_OUTPUT");
		g_out.native = e && atoi(e) != 0;
		g_out.decided = 1;
	}
	return g_out.native;
}

static void *original(const char *name)
{
	void *f = dlsym(RTLD_NEXT, name);
	if (!f) {
		SUBREADprintf("ERROR: the reference's %s is not reachable\n", name);
		exit(1);
	}
	return f;
}

static int deflate_threads(int asked)
{
	const char *e = getenv("SRCUDA_OUT_THREADS");
	if (e && atoi(e) > 0) return atoi(e);
	long n = sysconf(_SC_NPROCESSORS_ONLN);
	if (n > 16) n = 16;
	return asked > n ? asked : (int)(n > 0 ? n : 1);
}

static void output_failed(const char *what)
{
	SUBREADprintf("ERROR: %s: %s\n", what, srcuda_out_error(g_out.writer));
}

/* ---- SamBam_writer_*: the calls core.c makes (core.c:4142,4163, 3850-3870, 1851-1864, 2147-2150, 3015-3016, 4285) ------ */

int SamBam_writer_create(SamBam_Writer *writer, char *BAM_fname, int threads, int sort_reads_by_coord, int is_temp_BAM_file, char *tmpfname)
{
	if (native_output() && !sort_reads_by_coord && !g_out.writer) {
		memset(writer, 0, sizeof(SamBam_Writer));
		writer->threads = threads;
		if (srcuda_out_open(BAM_fname, SRCUDA_OUT_BAM, deflate_threads(threads), is_temp_BAM_file, &g_out.writer)) {
			SUBREADprintf("ERROR: %s\n", srcuda_out_error(NULL));
			return -1;
		}
		return 0;
	}
	if (native_output()) g_out.native = 0; /* a sorted BAM (or a second writer): all of it is the reference's */
	return ((int (*)(SamBam_Writer *, char *, int, int, int, char *))original("SamBam_writer_create"))(writer, BAM_fname, threads,
	                                                                                                 sort_reads_by_coord, is_temp_BAM_file, tmpfname);
}

int SamBam_writer_add_header(SamBam_Writer *writer, char *header_text, int add_chro)
{
	if (native_output() && g_out.writer) return srcuda_out_add_header(g_out.writer, header_text, add_chro) ? -1 : 0;
	return ((int (*)(SamBam_Writer *, char *, int))original("SamBam_writer_add_header"))(writer, header_text, add_chro);
}

int SamBam_writer_add_chromosome(SamBam_Writer *writer, char *chro_name, unsigned int chro_length, int add_header)
{
	if (native_output() && g_out.writer) return srcuda_out_add_chromosome(g_out.writer, chro_name, chro_length, add_header) ? -1 : 0;
	return ((int (*)(SamBam_Writer *, char *, unsigned int, int))original("SamBam_writer_add_chromosome"))(writer, chro_name, chro_length, add_header);
}

void SamBam_writer_finish_header(SamBam_Writer *writer)
{
	if (native_output() && g_out.writer) {
		if (srcuda_out_finish_header(g_out.writer)) writer->is_internal_error = 1;
		writer->writer_state = 10;
		return;
	}
	((void (*)(SamBam_Writer *))original("SamBam_writer_finish_header"))(writer);
}

int SamBam_writer_add_read(SamBam_Writer *writer, int thread_no, char *read_name, unsigned int flags, char *chro_name, unsigned int chro_position,
                           int mapping_quality, char *cigar, char *next_chro_name, unsigned int next_chro_position, int temp_len, int read_len,
                           char *read_text, char *qual_text, char *additional_columns, int committable)
{
	if (native_output() && g_out.writer) {
		srcuda_sam_record r = {read_name, flags, chro_name, chro_position, mapping_quality, cigar, next_chro_name, next_chro_position,
		                       temp_len, read_len, read_text, qual_text, additional_columns};
		pthread_mutex_lock(&g_out.lock);
		int rc = srcuda_out_add_read(g_out.writer, &r, committable);
		pthread_mutex_unlock(&g_out.lock);
		if (rc) writer->is_internal_error = 1;
		return rc ? 1 : 0;
	}
	return ((int (*)(SamBam_Writer *, int, char *, unsigned int, char *, unsigned int, int, char *, char *, unsigned int, int, int, char *, char *,
	                 char *, int))original("SamBam_writer_add_read"))(writer, thread_no, read_name, flags, chro_name, chro_position, mapping_quality,
	                                                                 cigar, next_chro_name, next_chro_position, temp_len, read_len, read_text,
	                                                                 qual_text, additional_columns, committable);
}

void SamBam_writer_finalise_thread(SamBam_Writer *writer, int thread_id)
{
	if (native_output() && g_out.writer) return; /* one stream: nothing is parked per thread */
	((void (*)(SamBam_Writer *, int))original("SamBam_writer_finalise_thread"))(writer, thread_id);
}

void SamBam_writer_finalise_one_thread(SamBam_Writer *writer)
{
	if (native_output() && g_out.writer) return;
	((void (*)(SamBam_Writer *))original("SamBam_writer_finalise_one_thread"))(writer);
}

int SamBam_writer_close(SamBam_Writer *writer)
{
	if (native_output() && g_out.writer) {
		srcuda_writer *w = g_out.writer;
		g_out.writer = NULL;
		if (srcuda_out_close(w, NULL)) {
			SUBREADprintf("ERROR: %s\n", srcuda_out_error(NULL));
			writer->is_internal_error = 1;
		}
		return 0;
	}
	return ((int (*)(SamBam_Writer *))original("SamBam_writer_close"))(writer);
}

/* ---- write_sam_headers (core.c:3843-3890) ---------------------------------------------------------------------------- */

void write_sam_headers(global_context_t *context)
{
	if (!native_output() || context->config.sort_reads_by_coordinates || (context->config.is_BAM_output && !g_out.writer)) {
		((void (*)(global_context_t *))original("write_sam_headers"))(context);
		return;
	}
	if (!context->config.is_BAM_output) {
		/* the reference has opened (and so emptied) the file already, core.c:4147; its own stream stays unused and is
		 * closed empty in destroy_global_context */
		char *path = (context->config.report_sam_file && context->config.output_prefix[0]) ? context->config.output_prefix : NULL;
		if (srcuda_out_open(path, SRCUDA_OUT_SAM, 1, 0, &g_out.writer)) {
			SUBREADprintf("ERROR: %s\n", srcuda_out_error(NULL));
			context->output_sam_is_full = 1;
			g_out.native = 0;
			return;
		}
	}
	srcuda_writer *w = g_out.writer;
	SUBREADprintf("Output: %s records are written by libsubread_cuda (srcuda_out_*)\n", context->config.is_BAM_output ? "BAM" : "SAM");
	char *line = malloc(15000);
	int rc = srcuda_out_add_header(w, "@HD\tVN:1.0\tSO:unsorted", 0);
	unsigned int last_offset = 0;
	for (int k = 0; k < context->chromosome_table.total_offsets && !rc; k++) {
		unsigned int length = context->chromosome_table.read_offsets[k] - last_offset + 16 - 2 * context->chromosome_table.padding;
		rc = srcuda_out_add_chromosome(w, context->chromosome_table.read_names + (size_t)k * MAX_CHROMOSOME_NAME_LEN, length, 1);
		last_offset = context->chromosome_table.read_offsets[k];
	}
	if (!rc && context->config.read_group_id[0]) {
		snprintf(line, 10000, "@RG\tID:%s%s", context->config.read_group_id, context->config.read_group_txt);
		rc = srcuda_out_add_header(w, line, 0);
	}
	if (!rc) {
		snprintf(line, 9999 + 4900, "@PG\tID:subread\tPN:subread\tVN:%s\tCL:%s", SUBREAD_VERSION, context->rebuilt_command_line);
		rc = srcuda_out_add_header(w, line, 0);
	}
	if (!rc) rc = srcuda_out_finish_header(w);
	free(line);
	if (context->config.is_BAM_output) context->output_bam_writer->writer_state = 10;
	if (rc) {
		output_failed("cannot write the header");
		if (context->config.is_BAM_output) context->output_bam_writer->is_internal_error = 1;
		else context->output_sam_is_full = 1;
	}
}

/* ---- input order without waiting ---------------------------------------------------------------------------------------
 * The reference keeps SAM output and --keepReadOrder BAM output in input order by making every worker WAIT, with the finished
 * records in its hands, until the fragment in front has been written (add_buffered_fragment, core.c:1855-1881: a 2 us sleep per
 * look, one fragment per turn) -- the realignment step runs at the speed of that relay.  Here a worker never waits: a fragment
 * that is next in line goes to the stream at once, together with every parked fragment behind it; any other fragment is parked
 * (its encoded bytes, one piece per reported location so that the blocks are cut where the reference cuts them) and its worker
 * goes on with the next read.  last_written_fragment_number is the reference's own counter (reset for every step, core.c:3385). */
typedef struct {
	long long pair;
	char *bytes; /* pieces: u32 length + payload, one per location */
	size_t used, room;
	int complete;
} parked_t;

static struct {
	pthread_mutex_t mu;
	parked_t *ring;
	size_t cap;
} g_order = {PTHREAD_MUTEX_INITIALIZER, NULL, 0};

static __thread parked_t t_mine; /* the fragment this worker is collecting (a multi-mapped read arrives location by location) */

static int write_pieces(const char *bytes, size_t used)
{
	int rc = 0;
	for (size_t at = 0; at + 4 <= used;) {
		unsigned int n;
		memcpy(&n, bytes + at, 4);
		if (n && srcuda_out_append(g_out.writer, bytes + at + 4, n, 1)) rc = 1;
		at += 4 + n;
	}
	return rc;
}

static int hand_over_in_order(global_context_t *gc, long long pair_number, const char *encoded, size_t bytes, int last_location)
{
	parked_t *m = &t_mine;
	if (m->used + 4 + bytes > m->room) {
		m->room = 2 * (m->used + 4 + bytes) + 4096;
		m->bytes = realloc(m->bytes, m->room);
		if (!m->bytes) { m->room = m->used = 0; return 1; }
	}
	unsigned int n = (unsigned int)bytes;
	memcpy(m->bytes + m->used, &n, 4);
	memcpy(m->bytes + m->used + 4, encoded, bytes);
	m->used += 4 + bytes;
	if (!last_location) return 0;
	int rc = 0;
	pthread_mutex_lock(&g_order.mu);
	long long next = (long long)gc->last_written_fragment_number + 1;
	if (pair_number == next) {
		rc |= write_pieces(m->bytes, m->used);
		next++;
		while (g_order.cap) { /* everybody who was waiting right behind */
			parked_t *p = &g_order.ring[(size_t)next % g_order.cap];
			if (!p->complete || p->pair != next) break;
			rc |= write_pieces(p->bytes, p->used);
			free(p->bytes);
			memset(p, 0, sizeof *p);
			next++;
		}
		gc->last_written_fragment_number = next - 1;
		m->used = 0;
	} else {
		size_t ahead = (size_t)(pair_number - next);
		if (ahead >= g_order.cap) { /* a ring that holds every fragment between the next one and this one */
			size_t cap = g_order.cap ? g_order.cap : 1024;
			while (cap <= ahead) cap *= 2;
			parked_t *ring = calloc(cap, sizeof(parked_t));
			if (!ring) { pthread_mutex_unlock(&g_order.mu); return 1; }
			for (size_t i = 0; i < g_order.cap; i++)
				if (g_order.ring[i].complete) ring[(size_t)g_order.ring[i].pair % cap] = g_order.ring[i];
			free(g_order.ring);
			g_order.ring = ring;
			g_order.cap = cap;
		}
		parked_t *p = &g_order.ring[(size_t)pair_number % g_order.cap];
		*p = *m; /* the worker's buffer moves into the ring; it starts a new one */
		p->pair = pair_number;
		p->complete = 1;
		memset(m, 0, sizeof *m);
	}
	pthread_mutex_unlock(&g_order.mu);
	return rc;
}

/* ---- write_single_fragment (core.c:1888-2170) ------------------------------------------------------------------------- */

static void fill_mate(global_context_t *g, srcuda_sam_mate *m, output_tmp_t *rec, realignment_result_t *raw, int ok, char *name, char *text, char *qual,
                      int read_len)
{
	memset(m, 0, sizeof *m);
	m->present = rec != NULL;
	m->ok = ok;
	m->chro_id = -1;
	m->read_len = read_len;
	m->name = name;
	m->text = text;
	m->qual = qual;
	if (!rec) return;
	char *names = g->chromosome_table.read_names;
	if (rec->chro >= names && rec->chro < names + (size_t)g->chromosome_table.total_offsets * MAX_CHROMOSOME_NAME_LEN)
		m->chro_id = (int)((rec->chro - names) / MAX_CHROMOSOME_NAME_LEN);
	else if (rec->chro)
		m->chro_id = -2 - (int)(((size_t)rec->chro >> 3) & 0xffffff); /* not a table entry: equal pointers still compare equal */
	m->offset = rec->offset;
	m->strand = rec->strand;
	m->mapping_quality = rec->mapping_quality;
	m->cigar = rec->cigar;
	m->additional = rec->additional_information;
	m->mismatched_bases = raw ? raw->final_mismatched_bases : 0;
}

void write_single_fragment(global_context_t *global_context, thread_context_t *thread_context, output_tmp_t *rec1, realignment_result_t *raw_r1,
                           output_tmp_t *rec2, realignment_result_t *raw_r2, int all_locations, int current_location, char *read_name_1,
                           char *read_name_2, int read_len_1, int read_len_2, char *read_text_1, char *read_text_2, char *qual_text_1,
                           char *qual_text_2, subread_read_number_t pair_number, int non_informative_subread_r1, int non_informative_subread_r2,
                           int is_R1_OK, int is_R2_OK)
{
	if (!native_output() || !g_out.writer || global_context->config.space_type != GENE_SPACE_BASE || global_context->config.convert_color_to_base) {
		((void (*)(global_context_t *, thread_context_t *, output_tmp_t *, realignment_result_t *, output_tmp_t *, realignment_result_t *, int, int,
		           char *, char *, int, int, char *, char *, char *, char *, subread_read_number_t, int, int, int, int))original("write_single_fragment"))(
		    global_context, thread_context, rec1, raw_r1, rec2, raw_r2, all_locations, current_location, read_name_1, read_name_2, read_len_1, read_len_2,
		    read_text_1, read_text_2, qual_text_1, qual_text_2, pair_number, non_informative_subread_r1, non_informative_subread_r2, is_R1_OK, is_R2_OK);
		return;
	}
	int paired = global_context->input_reads.is_paired_end_reads;
	srcuda_sam_config cfg = {paired, global_context->config.minimum_pair_distance, global_context->config.maximum_pair_distance,
	                         global_context->config.is_first_read_reversed, global_context->config.is_second_read_reversed, 0,
	                         global_context->config.read_group_id};
	srcuda_sam_mate m1, m2;
	fill_mate(global_context, &m1, rec1, raw_r1, is_R1_OK, read_name_1, read_text_1, qual_text_1, read_len_1);
	if (paired) fill_mate(global_context, &m2, rec2, raw_r2, is_R2_OK, read_name_2, read_text_2, qual_text_2, read_len_2);
	srcuda_sam_counters c;
	memset(&c, 0, sizeof c);

	/* the records are built here, on the worker's own thread; only handing the bytes over takes a turn */
	static __thread char *encoded;
	static __thread size_t encoded_room;
	size_t bytes = 0, room = 4096 + 2 * strlen(global_context->config.read_group_id);
	room += 4 * (size_t)read_len_1 + strlen(read_name_1) + (is_R1_OK ? strlen(rec1->cigar) + strlen(rec1->additional_information) : 0);
	if (paired) room += 4 * (size_t)read_len_2 + strlen(read_name_2) + (is_R2_OK ? strlen(rec2->cigar) + strlen(rec2->additional_information) : 0);
	if (room > encoded_room) {
		free(encoded);
		encoded = malloc(encoded_room = 2 * room);
	}
	int rc = encoded ? srcuda_out_format_fragment(g_out.writer, &cfg, &m1, paired ? &m2 : NULL, all_locations, current_location, &c, encoded,
	                                              encoded_room, &bytes)
	                 : SRCUDA_ERR_NOMEM;
	if (rc) {
		SUBREADprintf("ERROR: cannot build a record: %s\n", srcuda_out_error(NULL));
		bytes = 0; /* the turn below is still taken, or the workers behind would wait for ever */
	}
	int in_order = thread_context && !(global_context->config.is_BAM_output && !global_context->config.is_input_read_order_required);
	if (in_order && global_context->config.ignore_unmapped_reads) {
		/* fragments may be missing from the sequence: keep the original's turn-taking (add_buffered_fragment, core.c:1855-1881) */
		for (;;) {
			int done = 0;
			subread_lock_occupy(&global_context->output_lock);
			if (global_context->last_written_fragment_number == pair_number - 1) {
				if (bytes && srcuda_out_append(g_out.writer, encoded, bytes, 1)) rc = SRCUDA_ERR_IO;
				if (all_locations <= current_location + 1) global_context->last_written_fragment_number = pair_number;
				done = 1;
			}
			subread_lock_release(&global_context->output_lock);
			if (done) break;
			usleep(2);
		}
	} else if (in_order) {
		if (hand_over_in_order(global_context, pair_number, encoded, bytes, all_locations <= current_location + 1)) rc = SRCUDA_ERR_IO;
	} else if (bytes) {
		pthread_mutex_lock(&g_out.lock);
		if (srcuda_out_append(g_out.writer, encoded, bytes, 1)) rc = SRCUDA_ERR_IO;
		pthread_mutex_unlock(&g_out.lock);
	}
	if (rc) {
		output_failed("cannot write a record");
		if (global_context->config.is_BAM_output) global_context->output_bam_writer->is_internal_error = 1;
		else global_context->output_sam_is_full = 1;
	}
	if (thread_context) {
		thread_context->all_mapped_reads += c.all_mapped_reads;
		thread_context->all_correct_PE_reads += c.all_correct_PE_reads;
		thread_context->not_properly_pairs_different_chro += c.not_properly_pairs_different_chro;
		thread_context->not_properly_different_strands += c.not_properly_different_strands;
		thread_context->not_properly_pairs_TLEN_wrong += c.not_properly_pairs_TLEN_wrong;
		thread_context->not_properly_pairs_wrong_arrangement += c.not_properly_pairs_wrong_arrangement;
	} else {
		global_context->all_mapped_reads += c.all_mapped_reads;
		global_context->all_correct_PE_reads += c.all_correct_PE_reads;
		global_context->not_properly_pairs_different_chro += c.not_properly_pairs_different_chro;
		global_context->not_properly_different_strands += c.not_properly_different_strands;
		global_context->not_properly_pairs_TLEN_wrong += c.not_properly_pairs_TLEN_wrong;
		global_context->not_properly_pairs_wrong_arrangement += c.not_properly_pairs_wrong_arrangement;
	}
	/* the tail of the original (core.c:2167-2168): the bigtable position moves behind the soft clip */
	if (is_R1_OK) rec1->raw_result->selected_position += rec1->soft_clipping_movements;
	if (is_R2_OK) rec2->raw_result->selected_position += rec2->soft_clipping_movements;
}

/* ---- destroy_global_context (core.c:4258): the SAM stream of the library is closed before the reference tidies up ------ */

int destroy_global_context(global_context_t *context)
{
	if (getenv("SRCUDA_PHASE_TIMES")) /* the reference's own clocks (core.c:3552-3640), for the whole-program breakdown in DESIGN.md */
		fprintf(stderr, "PHASES load_index=%.2f voting=%.2f before_realign=%.2f realign=%.2f\n", context->timecost_load_index, context->timecost_voting,
		        context->timecost_before_realign, context->timecost_for_realign);
	if (native_output() && g_out.writer && !context->config.is_BAM_output) {
		srcuda_writer *w = g_out.writer;
		g_out.writer = NULL;
		if (srcuda_out_close(w, NULL)) {
			SUBREADprintf("ERROR: %s\n", srcuda_out_error(NULL));
			context->output_sam_is_full = 1;
		}
	}
	return ((int (*)(global_context_t *))original("destroy_global_context"))(context);
}
