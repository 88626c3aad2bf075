/* integration/do_voting_cuda.c -- the reference-side binding of libsubread_cuda.so.
 *
 * This is what a Subread maintainer adds to run STEP_VOTING on a B200: a replacement for
 * `do_voting` (core.c:3049) and for `gehash_load` (sorted-hashtable.c:1390) written against the
 * reference's own headers, calling the C ABI of include/subread_cuda.h.  Everything else -- CLI,
 * FASTQ/BAM input, chunk loop (read_chunk_circles, core.c:3539), indel/junction event tables,
 * realignment, SAM/BAM output -- is the reference's unmodified code.
 *
 * In this repository it is built only as a demonstration/test vehicle (oracle/Makefile, target
 * _ref/subread-align-cuda): the reference's CLI main + this file, linked against the -fPIC build
 * of the reference (libsubread_ref.so).  Because the reference's cross-unit calls go through the
 * PLT, defining do_voting / gehash_load here interposes on them without touching a reference
 * source line.  A maintainer would instead compile this file into the tree and drop the two
 * originals.  tests/test_integration_sam.py checks that the SAM output is byte-identical to the
 * stock subread-align (minus @PG).
 *
 * Threading: run_maybe_threads (core.c:3379) starts config.all_threads workers that all enter
 * do_voting.  Worker 0 reads the whole chunk -- through the library's chunk reader (srcuda_reader_*) when
 * the input is plain FASTQ, else through the reference's own fetch_next_read_pair -- and
 * makes ONE srcuda_vote_chunk_multi call -- the chunk spread over the GPUs of SRCUDA_DEVICES -- that
 * writes straight into the bigtable (bigtable_cache_malloc_ptr has exactly the out_align layout,
 * bigtable_cache_malloc_ptr_junc the out_junc one); then all workers share the per-read tail of
 * do_voting (find_new_indels / find_new_junctions, core.c:3241-3290).  subjunc (do_breakpoint_detection)
 * is bound as well, for reads up to 160 bases.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include "subread.h"
#include "core.h"
#include "core-indel.h"
#include "core-junction.h"
#include "input-files.h"
#include "sorted-hashtable.h"

#include "../include/subread_cuda.h"

/* Everything the binding keeps between calls, in one place (the reference itself runs one global_context_t per process).
 * SRCUDA_DEVICES=0,1,... lists the GPUs to use (default: SRCUDA_DEVICE or device 0): the index block is loaded once and
 * replicated to the others over NVLink (srcuda_index_load_all), every chunk is cut into one contiguous slice per GPU
 * (srcuda_vote_chunk_multi). */
#define BINDING_MAX_GPUS 16
typedef struct {
	srcuda_handle *handles[BINDING_MAX_GPUS];
	int n_handles;
	int opened;
} binding_gpus_t;

static binding_gpus_t g_gpus;

static binding_gpus_t *cuda_handles(void)
{
	binding_gpus_t *b = &g_gpus;
	if (b->opened) return b->n_handles ? b : NULL;
	b->opened = 1;
	int devs[BINDING_MAX_GPUS], n = 0;
	const char *list = getenv("SRCUDA_DEVICES"), *one = getenv("SRCUDA_DEVICE");
	if (list && *list) {
		const char *p = list;
		while (*p && n < BINDING_MAX_GPUS) {
			char *end;
			long d = strtol(p, &end, 10);
			if (end == p) break;
			devs[n++] = (int)d;
			p = *end == ',' ? end + 1 : end;
		}
	}
	if (!n) devs[n++] = one ? atoi(one) : 0;
	for (int k = 0; k < n; k++) {
		if (srcuda_open(devs[k], &b->handles[b->n_handles])) {
			SUBREADprintf("ERROR: cannot open CUDA device %d: %s\n", devs[k], srcuda_last_error(NULL));
			for (int j = 0; j < b->n_handles; j++) srcuda_close(b->handles[j]);
			b->n_handles = 0;
			return NULL;
		}
		b->n_handles++;
	}
	return b;
}

/* ---- gehash_load: the block goes to HBM instead of host memory -------------------------------- */
int gehash_load(gehash_t *the_table, const char fname[])
{
	memset(the_table, 0, sizeof(gehash_t));
	size_t n = strlen(fname);
	/* "<prefix>.NN.b.tab" */
	if (n < 9 || strcmp(fname + n - 6, ".b.tab") != 0) {
		SUBREADprintf("ERROR: unexpected index block name '%s'\n", fname);
		return 1;
	}
	char prefix[MAX_FILE_NAME_LENGTH + 40];
	strncpy(prefix, fname, sizeof(prefix) - 1);
	prefix[n - 9] = 0;
	int block_no = atoi(fname + n - 8);
	binding_gpus_t *g = cuda_handles();
	if (!g) return 1;
	srcuda_handle *h = g->handles[0];
	if (srcuda_index_load_all(g->handles, g->n_handles, prefix, block_no)) {
		SUBREADprintf("ERROR: %s\n", srcuda_last_error(h));
		return 1;
	}
	srcuda_index_desc d;
	srcuda_index_describe(h, &d);
	the_table->version_number = SUBINDEX_VER2;
	the_table->index_gap = d.index_gap;
	the_table->padding = d.padding;
	the_table->current_items = d.n_items;
	the_table->buckets_number = 0; /* nothing on the host: gehash_destory has nothing to free */
	the_table->buckets = NULL;
	return 0;
}

/* ---- chunk buffers --------------------------------------------------------------------------- */
typedef struct {
	char *text, *qual, *names;
	uint64_t *off;      /* per read, into text (and qual) */
	uint64_t *name_off;
	uint16_t *len;
	size_t text_cap, text_used, names_cap, names_used, reads_cap, n_reads;
	int raw_orientation;  /* the reads are as they are in the files (chunk reader): the mate reversal of fetch_next_read_pair is still to come */
} chunk_buf_t;

static chunk_buf_t g_chunk;

static void chunk_push(chunk_buf_t *c, const char *name, const char *text, const char *qual, int len)
{
	if (c->n_reads + 1 >= c->reads_cap) {
		c->reads_cap = c->reads_cap ? c->reads_cap * 2 : (1u << 20);
		c->off = realloc(c->off, sizeof(uint64_t) * (c->reads_cap + 1));
		c->name_off = realloc(c->name_off, sizeof(uint64_t) * (c->reads_cap + 1));
		c->len = realloc(c->len, sizeof(uint16_t) * c->reads_cap);
	}
	if (c->text_used + (size_t)len + 2 > c->text_cap) {
		c->text_cap = c->text_cap ? c->text_cap * 2 : (256u << 20);
		c->text = realloc(c->text, c->text_cap);
		c->qual = realloc(c->qual, c->text_cap);
	}
	size_t nl = strlen(name) + 1;
	if (c->names_used + nl > c->names_cap) {
		c->names_cap = c->names_cap ? c->names_cap * 2 : (64u << 20);
		c->names = realloc(c->names, c->names_cap);
	}
	c->off[c->n_reads] = c->text_used;
	c->len[c->n_reads] = (uint16_t)len;
	memcpy(c->text + c->text_used, text, (size_t)len);
	c->text[c->text_used + len] = 0;
	if (qual && qual[0]) memcpy(c->qual + c->text_used, qual, (size_t)len); else memset(c->qual + c->text_used, 0, (size_t)len);
	c->qual[c->text_used + len] = 0;
	c->text_used += (size_t)len + 1;
	c->name_off[c->n_reads] = c->names_used;
	memcpy(c->names + c->names_used, name, nl);
	c->names_used += nl;
	c->n_reads++;
}

/* ---- the chunk through the library's own reader (plain FASTQ) ---------------------------------- */
/* fetch_next_read_pair hands out one read per call under the input lock (0.8 M reads/s); srcuda_reader_* indexes and gathers a
 * whole chunk from several host threads.  Used when both inputs are plain FASTQ files and the mode needs neither names nor
 * qualities in the voting step (subread-align: find_new_indels reads neither).  The reference's own gene_input_t is then moved
 * on to where the library's reader stopped (a plain file's position is its byte offset, geinput_seek, input-files.c:699), so
 * that locate_read_files / go_chunk_start / go_chunk_nextchunk (core.c:3482-3537) and the later steps, which re-read the chunk
 * with the reference's reader, see exactly the state fetch_next_read_pair would have left.  SRCUDA_BINDING_READER=0 turns it off. */
typedef struct {
	srcuda_reader *rd;
	int state;            /* 0 not tried, 1 in use, -1 not usable */
} binding_reader_t;
static binding_reader_t g_reader;

static int reader_usable(global_context_t *gc)
{
	const char *e = getenv("SRCUDA_BINDING_READER");
	if (e && e[0] == '0') return 0;
	if (gc->config.do_breakpoint_detection || gc->config.do_fusion_detection || gc->config.do_long_del_detection) return 0;   /* find_new_junctions reads qualities */
	if (gc->input_reads.first_read_file.file_type != GENE_INPUT_FASTQ) return 0;
	if (gc->input_reads.is_paired_end_reads && gc->input_reads.second_read_file.file_type != GENE_INPUT_FASTQ) return 0;
	return 1;
}

/* Returns 1 when the chunk was read (c filled, *n_frags set), 0 when the reference's reader has to do it, -1 on error. */
static int read_chunk_with_library(global_context_t *gc, chunk_buf_t *c, subread_read_number_t *n_frags)
{
	binding_reader_t *b = &g_reader;
	const int is_paired = gc->input_reads.is_paired_end_reads, mates = 1 + is_paired;
	gene_input_t *g1 = &gc->input_reads.first_read_file, *g2 = is_paired ? &gc->input_reads.second_read_file : NULL;
	if (b->state < 0 || !reader_usable(gc)) return 0;
	if (b->state == 0) {
		if (ftello((FILE *)g1->input_fp) != 0 || (g2 && ftello((FILE *)g2->input_fp) != 0)) { b->state = -1; return 0; }   /* not at the start: leave it */
		if (srcuda_reader_open(gc->config.first_read_file, is_paired ? gc->config.second_read_file : NULL, 0, gc->config.all_threads,
		                       gc->config.read_trim_5, gc->config.read_trim_3, &b->rd)) {
			if (b->rd) srcuda_reader_close(b->rd);
			b->rd = NULL;
			b->state = -1;
			return 0;
		}
		b->state = 1;
	}
	/* both readers must stand at the same place */
	uint64_t p1 = 0, p2 = 0;
	srcuda_reader_tell(b->rd, &p1, &p2);
	if ((uint64_t)ftello((FILE *)g1->input_fp) != p1 || (g2 && (uint64_t)ftello((FILE *)g2->input_fp) != p2)) {
		SUBREADprintf("ERROR: the chunk reader and the input files are out of step (%llu vs %lld).\n", (unsigned long long)p1, (long long)ftello((FILE *)g1->input_fp));
		return -1;
	}
	const int64_t want = (int64_t)gc->config.reads_per_chunk;
	struct stat st1, st2;
	uint64_t cap = 64;
	if (stat(gc->config.first_read_file, &st1) == 0) cap += (uint64_t)st1.st_size - p1;
	if (is_paired && stat(gc->config.second_read_file, &st2) == 0) cap += (uint64_t)st2.st_size - p2;
	const uint64_t worst = (uint64_t)want * mates * (MAX_READ_LENGTH + 1);
	if (cap > worst) cap = worst;                /* (the reads of a chunk are part of the bytes left in the files) */
	if (c->text_cap < cap) { free(c->text); free(c->qual); c->text = malloc(cap); c->qual = NULL; c->text_cap = cap; }
	if (c->reads_cap < (size_t)want * mates + 1) {
		c->reads_cap = (size_t)want * mates + 1;
		c->off = realloc(c->off, sizeof(uint64_t) * (c->reads_cap + 1));
		c->name_off = realloc(c->name_off, sizeof(uint64_t) * (c->reads_cap + 1));
		c->len = realloc(c->len, sizeof(uint16_t) * c->reads_cap);
	}
	if (!c->text || !c->off || !c->len) { SUBREADprintf("ERROR: out of memory for the read chunk.\n"); return -1; }
	int64_t got = 0;
	uint64_t bytes = 0;
	if (srcuda_reader_next(b->rd, want, c->text, c->text_cap, c->off, c->len, &got, &bytes)) {
		SUBREADprintf("ERROR: %s\n", srcuda_reader_error(b->rd));
		return -1;
	}
	c->n_reads = (size_t)got * mates;
	c->text_used = bytes;
	c->raw_orientation = 1;
	*n_frags = (subread_read_number_t)got;
	/* the reference's files move on to the end of the chunk, its read counter too (fetch_next_read_pair, core.c:1175) */
	srcuda_reader_tell(b->rd, &p1, &p2);
	subread_lock_occupy(&gc->input_reads.input_lock);
	fseeko((FILE *)g1->input_fp, (off_t)p1, SEEK_SET);
	if (g2) fseeko((FILE *)g2->input_fp, (off_t)p2, SEEK_SET);
	gc->running_processed_reads_in_chunk = got;
	subread_lock_release(&gc->input_reads.input_lock);
	return 1;
}

/* ---- worker rendezvous ------------------------------------------------------------------------ */
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_cv = PTHREAD_COND_INITIALIZER;
static int g_vote_done = 0, g_vote_rc = 0, g_departed = 0;
static subread_read_number_t g_n_frags = 0;

static int unsupported_mode(global_context_t *gc)
{
	return gc->config.do_fusion_detection || gc->config.do_long_del_detection ||
	       gc->config.is_methylation_reads || gc->config.space_type != GENE_SPACE_BASE || gc->config.scRNA_input_mode ||
	       gc->config.SAM_extra_columns;
}

/* Worker 0: read the chunk with the reference's reader and vote it on the GPU. */
static int vote_chunk_on_gpu(global_context_t *gc, thread_context_t *tc)
{
	gene_input_t *ginp1 = &gc->input_reads.first_read_file, *ginp2 = NULL;
	int is_paired = gc->input_reads.is_paired_end_reads;
	if (is_paired) ginp2 = &gc->input_reads.second_read_file;
	char *t1 = malloc(MAX_READ_LENGTH + 1), *t2 = malloc(MAX_READ_LENGTH + 1), *q1 = malloc(MAX_READ_LENGTH + 1), *q2 = malloc(MAX_READ_LENGTH + 1);
	char n1[MAX_READ_NAME_LEN + 1], n2[MAX_READ_NAME_LEN + 1];
	chunk_buf_t *c = &g_chunk;
	int by_library = 0;
	if (gc->current_index_block_number > 0 && c->raw_orientation && g_reader.state == 1) {
		/* a later index block of the same chunk: the reads are still here (the reference rewound its files to re-read them) */
		by_library = 1;
		subread_lock_occupy(&gc->input_reads.input_lock);
		gc->running_processed_reads_in_chunk = g_n_frags;
		subread_lock_release(&gc->input_reads.input_lock);
	} else {
		c->n_reads = 0; c->text_used = 0; c->names_used = 0; c->raw_orientation = 0;
		subread_read_number_t got = 0;
		by_library = read_chunk_with_library(gc, c, &got);
		if (by_library < 0) { free(t1); free(t2); free(q1); free(q2); return 1; }
		if (by_library) g_n_frags = got;
	}
	if (!by_library) {
		if (!c->qual && c->text_cap) c->qual = malloc(c->text_cap);      /* (the library's reader left no quality buffer) */
		for (;;) {
			int l1 = 0, l2 = 0;
			subread_read_number_t no = 0;
			fetch_next_read_pair(gc, tc, ginp1, ginp2, &l1, &l2, n1, n2, t1, t2, q1, q2, 1, &no);
			if (no < 0) break;
			chunk_push(c, n1, t1, q1, l1);
			if (is_paired) chunk_push(c, n2, t2, q2, l2);
		}
		g_n_frags = (subread_read_number_t)(c->n_reads / (1 + is_paired));
	}
	free(t1); free(t2); free(q1); free(q2);
	if (!g_n_frags) return 0;

	srcuda_params p;
	memset(&p, 0, sizeof(p));
	p.total_subreads = gc->config.total_subreads;
	p.top_scores = gc->config.top_scores;
	p.max_vote_simples = gc->config.max_vote_simples;
	p.max_vote_combinations = gc->config.max_vote_combinations;
	p.max_vote_number_cutoff = gc->config.max_vote_number_cutoff;
	p.multi_best_reads = gc->config.multi_best_reads;
	p.minimum_subread_for_first_read = gc->config.minimum_subread_for_first_read;
	p.minimum_subread_for_second_read = gc->config.minimum_subread_for_second_read;
	p.max_indel_length = gc->config.max_indel_length;
	p.minimum_pair_distance = gc->config.minimum_pair_distance;
	p.maximum_pair_distance = gc->config.maximum_pair_distance;
	p.is_first_read_reversed = 0;  /* fetch_next_read_pair already oriented both mates (core.c:1187-1199) */
	p.is_second_read_reversed = 0;
	if (c->raw_orientation) {      /* ... the library's reader did not: the voting step does it itself */
		p.is_first_read_reversed = gc->config.is_first_read_reversed;
		p.is_second_read_reversed = gc->config.is_second_read_reversed;
	}
	/* subjunc: the junction half of copy_vote_to_alignment_res (core-junction.c:1073-1334) runs on the GPU too */
	p.do_breakpoint_detection = gc->config.do_breakpoint_detection;
	p.do_big_margin_filtering_for_junctions = gc->config.do_big_margin_filtering_for_junctions;
	p.big_margin_record_size = gc->config.big_margin_record_size;
	p.maximum_intron_length = (int32_t)gc->config.maximum_intron_length;
	p.limited_tree_scan = gc->config.limited_tree_scan;
	p.check_donor_at_junctions = gc->config.check_donor_at_junctions;
	p.more_accurate_fusions = gc->config.more_accurate_fusions && gc->config.do_fusion_detection;
	p.max_insertion_at_junctions = gc->config.max_insertion_at_junctions;
	p.prefer_donor_receptor_junctions = gc->config.prefer_donor_receptor_junctions;
	binding_gpus_t *g = cuda_handles();
	if (!g) return 1;
	srcuda_subjunc_result *junc = NULL;
	uint16_t *bm = NULL;
	const int first_block = gc->current_index_block_number == 0;
	if (p.do_breakpoint_detection) {
		/* reads beyond 160 bases also feed core_fragile_junction_voting (core.c:3138), which walks the HOST hash table this binding
		 * never loads: say so instead of mapping them differently */
		for (size_t r = 0; r < c->n_reads; r++)
			if (c->len[r] > EXON_LONG_READ_LENGTH) {
				SUBREADprintf("ERROR: subjunc reads longer than %d bases need core_fragile_junction_voting on the host index; not supported by this binding.\n", EXON_LONG_READ_LENGTH);
				return 1;
			}
		junc = (srcuda_subjunc_result *)gc->bigtable_cache_malloc_ptr_junc;
		/* the big-margin records live inside bigtable_cached_result_t, one per read: gathered into the dense [read][27] layout of the
		 * C ABI and scattered back */
		bm = malloc(sizeof(uint16_t) * 27 * c->n_reads);
		for (size_t r = 0; r < c->n_reads; r++) memcpy(bm + 27 * r, gc->bigtable_cache[r].big_margin_data, sizeof(uint16_t) * 27);
	}
	int rc = srcuda_vote_chunk_multi(g->handles, g->n_handles, &p, c->text, c->off, c->len, (int64_t)g_n_frags, is_paired, first_block,
	                                 (srcuda_mapping_result *)gc->bigtable_cache_malloc_ptr, junc, bm, NULL);
	if (rc) SUBREADprintf("ERROR: srcuda_vote_chunk_multi: %s\n", srcuda_last_error(g->handles[0]));
	if (bm) {
		if (!rc) for (size_t r = 0; r < c->n_reads; r++) memcpy(gc->bigtable_cache[r].big_margin_data, bm + 27 * r, sizeof(uint16_t) * 27);
		free(bm);
	}
	return rc;
}

/* The per-read tail of do_voting on the last index block (core.c:3236-3290), for fragments [f0, f1). */
static void discover_events(global_context_t *gc, thread_context_t *tc, subread_read_number_t f0, subread_read_number_t f1)
{
	int mates = 1 + gc->input_reads.is_paired_end_reads;
	int need_junction_step = gc->config.do_breakpoint_detection || gc->config.do_fusion_detection || gc->config.do_long_del_detection;
	char *text = malloc(MAX_READ_LENGTH + 1), *qual = malloc(MAX_READ_LENGTH + 1);
	chunk_buf_t *c = &g_chunk;
	for (subread_read_number_t f = f0; f < f1; f++)
		for (int m = 0; m < mates; m++) {
			size_t r = (size_t)f * mates + m;
			int rlen = c->len[r];
			memcpy(text, c->text + c->off[r], (size_t)rlen);
			text[rlen] = 0;
			if (c->raw_orientation) {
				/* as fetch_next_read_pair would have handed it over (core.c:1187-1199); no qualities, no name on this path */
				if (m ? gc->config.is_second_read_reversed : gc->config.is_first_read_reversed) reverse_read(text, rlen, gc->config.space_type);
				memset(qual, 0, (size_t)rlen + 1);
			} else
				memcpy(qual, c->qual + c->off[r], (size_t)rlen + 1);
			const char *rname = c->raw_orientation ? "" : c->names + c->name_off[r];
			/* after its two strand passes do_voting holds the read reversed (core.c:3229-3234) and the
			 * quality string as fetched; read_?_reversed starts at 1 (core.c:3237-3238) */
			reverse_read(text, rlen, gc->config.space_type);
			int has_reversed = 1;
			char *current_qual = qual;
			for (int best = 0; best < gc->config.multi_best_reads; best++) {
				mapping_result_t *cur = _global_retrieve_alignment_ptr(gc, f, m, best);
				if (cur->selected_votes < 1) continue;
				int should = (cur->result_flags & CORE_IS_NEGATIVE_STRAND) ? 1 : 0;
				if (should != has_reversed) {
					has_reversed = !has_reversed;
					reverse_read(text, rlen, gc->config.space_type);
					reverse_quality(current_qual, rlen);
				}
				gene_value_index_t *saved = tc ? tc->current_value_index : gc->current_value_index;
				locate_current_value_index(gc, tc, cur, rlen);
				if (!has_better_mapping(gc, tc, f, m, best))
					find_new_indels(gc, tc, f, (char *)rname, text, current_qual, rlen, m, best);
				if (need_junction_step)
					find_new_junctions(gc, tc, f, (char *)rname, text, current_qual, rlen, m, best);
				if (tc) tc->current_value_index = saved; else gc->current_value_index = saved;
			}
		}
	free(text);
	free(qual);
}

/* TEST KNOB (both oracle/ref_hooks.c and integration/do_voting_cuda.c carry it, so that the stock reference and the bound one can be
 * run with the same small chunks): SUBREAD_TEST_READS_PER_CHUNK=<n> replaces config.reads_per_chunk (20 M reads divided by
 * multi_best_reads and mates, core-indel.c:4483, core.c:4092) right before the bigtable is sized from it (core-bigtable.c:84). */
int init_bigtable_results(global_context_t *global_context, int is_rewinding)
{
	static int (*real_fn)(global_context_t *, int) = NULL;
	if (!real_fn) real_fn = (int (*)(global_context_t *, int))dlsym(RTLD_NEXT, "init_bigtable_results");
	const char *n = getenv("SUBREAD_TEST_READS_PER_CHUNK");
	if (n && atoll(n) > 0) global_context->config.reads_per_chunk = (unsigned long long)atoll(n);
	return real_fn(global_context, is_rewinding);
}

int do_voting(global_context_t *gc, thread_context_t *tc)
{
	if (unsupported_mode(gc)) {
		/* modes the GPU library refuses (fusion, long deletions, methylation, colour space ...): refused loudly -- the host hash
		 * table is not loaded by this binding, so there is no CPU do_voting to fall back to */
		static int told = 0;
		if (!told) { told = 1; SUBREADprintf("ERROR: this mode is not supported by libsubread_cuda (fusion / long-deletion detection, methylation, colour space, scRNA, --extraColumns).\n"); }
		return -1;
	}
	int n_threads = tc ? gc->config.all_threads : 1;
	int me = tc ? tc->thread_id : 0;
	if (tc) tc->current_value_index = gc->current_value_index;

	if (me == 0) {
		int rc = vote_chunk_on_gpu(gc, tc);
		pthread_mutex_lock(&g_mu);
		g_vote_rc = rc;
		g_vote_done = 1;
		pthread_cond_broadcast(&g_cv);
		pthread_mutex_unlock(&g_mu);
	} else {
		pthread_mutex_lock(&g_mu);
		while (!g_vote_done) pthread_cond_wait(&g_cv, &g_mu);
		pthread_mutex_unlock(&g_mu);
	}
	int rc = g_vote_rc;
	if (!rc && gc->is_final_voting_run && g_n_frags > 0) {
		subread_read_number_t f0 = g_n_frags * me / n_threads, f1 = g_n_frags * (me + 1) / n_threads;
		discover_events(gc, tc, f0, f1);
	}
	pthread_mutex_lock(&g_mu);
	if (++g_departed == n_threads) { g_departed = 0; g_vote_done = 0; }  /* every worker is joined before the next call */
	pthread_mutex_unlock(&g_mu);
	return rc;
}
