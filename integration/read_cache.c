/* integration/read_cache.c -- the reference-side binding of the library's FASTQ batches for the passes AFTER
 * voting: a replacement for fetch_next_read_pair (core.c:1121-1215).
 *
 * The reference reads every chunk of reads two or three times -- once per index block for voting, once for realignment
 * (do_iteration_two), once more with a third iteration -- and every time one read at a time, character by character through
 * stdio, under input_lock (core.c:1130).  In a process with threads every fgetc takes the FILE lock, so this serial section
 * delivers about 100 k reads/s whatever -T is, and once voting runs on the GPU it IS the realignment step's wall-clock
 * (build container, 600 k x 150 bp reads, -T 8: 7.0 s with it, 1.4 s without).
 *
 * Here the chunk's bytes are read once, when the first pass after voting asks for a read (the chunk's file range is known
 * then: current_circle_start/end_position, core.c:3482-3495), indexed by srcuda_fastq_split_mt on all threads, and every
 * fetch is a copy out of that buffer: which read comes next is one atomic step on running_processed_reads_in_chunk.  Names,
 * trimming, mate orientation (is_first/second_read_reversed) and the end-of-chunk answer are the original's.
 *
 * Memory: the chunk's bytes (a full chunk of 7 M reads of 150 bases is 2.4 GB per file) plus 30 bytes per read, freed when the next
 * chunk is loaded.
 *
 * Used in base space when SRCUDA_READ_CACHE=1.  Plain FASTQ input goes the way described above.  Every other input (gzip
 * FASTQ -- the seekable inflater re-inflates the chunk under the same lock --, FASTA, SAM / BAM) cannot be indexed from a byte
 * range, so its reads are KEPT as the chunk's first voting pass fetches them through the original (one allocation per fragment:
 * names, bases and qualities of its mates as the original returned them, mates already oriented) and the later passes are
 * served from that record.  locate_read_files (core.c:3482; called when a chunk starts and when its first voting pass is over)
 * tells the two phases apart.  Colour space, scRNA input, and any plain-FASTQ chunk whose bytes the splitter does not account
 * for exactly (a damaged record, a read that is empty after trimming, a quality line of another length than its read) go
 * through the original.  Linked into oracle/_ref/subread-align-cuda,
 * subjunc-cuda, subread-align-out and subjunc-out; tests/test_output_program.py compares SAM files with and without it.
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

#include "../include/subread_cuda.h"

typedef struct {
	char *raw;
	uint64_t bytes;
	uint64_t *off, *name_off, *qual_off;
	uint16_t *len, *name_len, *qual_len;
} cached_file_t;

typedef struct {
	pthread_mutex_t mu;
	int enabled; /* -1 unknown, 0 off, 1 on */
	int state;   /* of the chunk that starts at `key`: 0 not tried, 1 usable, -1 the original reads it */
	long long key;
	int64_t n_frags;
	cached_file_t f[2];
	int told;
} read_cache_t;

static read_cache_t g_cache = {PTHREAD_MUTEX_INITIALIZER, -1, 0, -1, 0, {{0}, {0}}, 0};

typedef int (*fetch_fn)(global_context_t *, thread_context_t *, gene_input_t *, gene_input_t *, int *, int *, char *, char *, char *, char *, char *,
                        char *, int, subread_read_number_t *);

static void drop_file(cached_file_t *f)
{
	free(f->raw); free(f->off); free(f->name_off); free(f->qual_off); free(f->len); free(f->name_len); free(f->qual_len);
	memset(f, 0, sizeof *f);
}

/* The bytes [from, to) of one read file, indexed.  1 = every read of the chunk is accounted for. */
static int load_file(global_context_t *gc, gene_input_t *in, long long from, long long to, int64_t n_reads, cached_file_t *f)
{
	if (to <= from || n_reads <= 0) return 0;
	f->bytes = (uint64_t)(to - from);
	f->raw = malloc(f->bytes + 1);
	f->off = malloc(sizeof(uint64_t) * n_reads);
	f->name_off = malloc(sizeof(uint64_t) * n_reads);
	f->qual_off = malloc(sizeof(uint64_t) * n_reads);
	f->len = malloc(sizeof(uint16_t) * n_reads);
	f->name_len = malloc(sizeof(uint16_t) * n_reads);
	f->qual_len = malloc(sizeof(uint16_t) * n_reads);
	if (!f->raw || !f->off || !f->name_off || !f->qual_off || !f->len || !f->name_len || !f->qual_len) return 0;
	int fd = fileno((FILE *)in->input_fp);
	uint64_t got = 0;
	while (got < f->bytes) {
		ssize_t n = pread(fd, f->raw + got, f->bytes - got, (off_t)(from + (long long)got));
		if (n <= 0) return 0;
		got += (uint64_t)n;
	}
	srcuda_fastq_batch b;
	int threads = gc->config.all_threads > 0 ? gc->config.all_threads : 1;
	if (srcuda_fastq_split_mt(f->raw, f->bytes, 1, gc->config.read_trim_5, gc->config.read_trim_3, n_reads, 0, 1, f->off, f->len, f->name_off,
	                          f->name_len, f->qual_off, f->qual_len, threads, &b))
		return 0;
	if (b.n_reads != n_reads || b.consumed != f->bytes) return 0; /* the chunk is not what the reference read: leave it to the original */
	for (int64_t i = 0; i < n_reads; i++)
		if (f->qual_len[i] != f->len[i] || f->len[i] == 0 || f->name_len[i] >= MAX_READ_NAME_LEN) return 0;
	return 1;
}

static int cache_enabled(void)
{
	if (g_cache.enabled < 0) {
		const char *e = getenv("SRCUDA_READ_CACHE");
		g_cache.enabled = e && atoi(e) != 0;
	}
	return g_cache.enabled;
}

/* 1 when the current chunk can be served from memory (loading it on the first call). */
static int chunk_ready(global_context_t *gc, gene_input_t *g1, gene_input_t *g2)
{
	if (!cache_enabled() || gc->config.space_type != GENE_SPACE_BASE || gc->config.scRNA_input_mode) return 0;
	if (g1->file_type != GENE_INPUT_FASTQ || (g2 && g2->file_type != GENE_INPUT_FASTQ)) return 0;
	long long s1 = (long long)gc->current_circle_start_position_file1.simple_file_position;
	long long e1 = (long long)gc->current_circle_end_position_file1.simple_file_position;
	if (e1 <= s1 || gc->processed_reads_in_chunk <= 0) return 0; /* the chunk's first voting pass: its end is not known yet */
	int st = __atomic_load_n(&g_cache.state, __ATOMIC_ACQUIRE);
	if (st && __atomic_load_n(&g_cache.key, __ATOMIC_ACQUIRE) == s1) return st > 0;
	pthread_mutex_lock(&g_cache.mu);
	if (!(g_cache.key == s1 && g_cache.state)) {
		__atomic_store_n(&g_cache.state, 0, __ATOMIC_RELEASE); /* nobody trusts the buffers while they change */
		drop_file(&g_cache.f[0]);
		drop_file(&g_cache.f[1]);
		int64_t n = (int64_t)gc->processed_reads_in_chunk;
		int ok = load_file(gc, g1, s1, e1, n, &g_cache.f[0]);
		if (ok && g2)
			ok = load_file(gc, g2, (long long)gc->current_circle_start_position_file2.simple_file_position,
			               (long long)gc->current_circle_end_position_file2.simple_file_position, n, &g_cache.f[1]);
		if (!ok) {
			drop_file(&g_cache.f[0]);
			drop_file(&g_cache.f[1]);
		} else if (!g_cache.told) {
			g_cache.told = 1;
			SUBREADprintf("Input: the passes after voting read their chunks from memory (libsubread_cuda, srcuda_fastq_split_mt)\n");
		}
		g_cache.n_frags = n;
		__atomic_store_n(&g_cache.key, s1, __ATOMIC_RELEASE);
		__atomic_store_n(&g_cache.state, ok ? 1 : -1, __ATOMIC_RELEASE);
	}
	st = g_cache.state;
	pthread_mutex_unlock(&g_cache.mu);
	return st > 0;
}

static void copy_read(const cached_file_t *f, int64_t k, char *name, char *text, char *qual, int *len)
{
	memcpy(name, f->raw + f->name_off[k], f->name_len[k]);
	name[f->name_len[k]] = 0;
	memcpy(text, f->raw + f->off[k], f->len[k]);
	text[f->len[k]] = 0;
	if (qual) {
		memcpy(qual, f->raw + f->qual_off[k], f->qual_len[k]);
		qual[f->qual_len[k]] = 0;
	}
	*len = f->len[k];
}

/* ---- inputs that are not plain FASTQ: the first voting pass is recorded ------------------------------------------------ */
typedef struct {
	char *blob; /* name1 \0 text1 \0 qual1 \0 [name2 \0 text2 \0 qual2 \0] */
	int len1, len2;
} kept_fragment_t;

static struct {
	int phase; /* 0 nothing, 1 the chunk's first voting pass is running (keep what it fetches), 2 later passes */
	int usable;
	kept_fragment_t *frag;
	long long cap, kept;
	int without_quality;
} g_kept;

static int keeps_reads(global_context_t *gc, gene_input_t *g1)
{
	return cache_enabled() && gc->config.space_type == GENE_SPACE_BASE && !gc->config.scRNA_input_mode && g1->file_type != GENE_INPUT_FASTQ &&
	       g1->file_type != GENE_INPUT_BCL;
}

static void drop_kept(void)
{
	for (long long i = 0; i < g_kept.cap && g_kept.frag; i++) free(g_kept.frag[i].blob);
	free(g_kept.frag);
	g_kept.frag = NULL;
	g_kept.cap = g_kept.kept = 0;
	g_kept.usable = g_kept.without_quality = 0;
}

void locate_read_files(global_context_t *global_context, int type)
{
	static void (*original)(global_context_t *, int);
	if (!original) original = (void (*)(global_context_t *, int))dlsym(RTLD_NEXT, "locate_read_files");
	original(global_context, type);
	if (!keeps_reads(global_context, &global_context->input_reads.first_read_file)) return;
	if (type == SEEK_SET) { /* a chunk starts (core.c:3566) */
		drop_kept();
		g_kept.cap = (long long)global_context->config.reads_per_chunk;
		g_kept.frag = calloc((size_t)g_kept.cap, sizeof(kept_fragment_t));
		g_kept.phase = g_kept.frag ? 1 : 0;
	} else if (g_kept.phase == 1) { /* its first voting pass is over (core.c:3598) */
		g_kept.phase = 2;
		g_kept.usable = g_kept.kept > 0 && g_kept.kept == (long long)global_context->running_processed_reads_in_chunk && !g_kept.without_quality;
		if (g_kept.usable && !g_cache.told) {
			g_cache.told = 1;
			SUBREADprintf("Input: the passes after voting read their chunks from memory (libsubread_cuda, reads kept from the first pass)\n");
		}
		if (!g_kept.usable) drop_kept();
	}
}

static void keep_fragment(subread_read_number_t k, int paired, const char *n1, const char *t1, const char *q1, int l1, const char *n2, const char *t2,
                          const char *q2, int l2)
{
	if (k < 0 || k >= g_kept.cap) return;
	if (!q1 || (paired && !q2)) {
		__atomic_store_n(&g_kept.without_quality, 1, __ATOMIC_RELAXED);
		return;
	}
	size_t a = strlen(n1) + 1, b = strlen(t1) + 1, c = strlen(q1) + 1, d = 0, e = 0, f = 0;
	if (paired) { d = strlen(n2) + 1; e = strlen(t2) + 1; f = strlen(q2) + 1; }
	char *blob = malloc(a + b + c + d + e + f);
	if (!blob) return;
	memcpy(blob, n1, a); memcpy(blob + a, t1, b); memcpy(blob + a + b, q1, c);
	if (paired) { memcpy(blob + a + b + c, n2, d); memcpy(blob + a + b + c + d, t2, e); memcpy(blob + a + b + c + d + e, q2, f); }
	kept_fragment_t *kf = &g_kept.frag[k];
	kf->blob = blob;
	kf->len1 = l1;
	kf->len2 = l2;
	__atomic_fetch_add(&g_kept.kept, 1, __ATOMIC_RELAXED);
}

static const char *take(const char *p, char *to)
{
	size_t n = strlen(p) + 1;
	if (to) memcpy(to, p, n);
	return p + n;
}

int fetch_next_read_pair(global_context_t *global_context, thread_context_t *thread_context, gene_input_t *ginp1, gene_input_t *ginp2, int *read_len_1,
                         int *read_len_2, char *read_name_1, char *read_name_2, char *read_text_1, char *read_text_2, char *qual_text_1,
                         char *qual_text_2, int remove_color_head, subread_read_number_t *read_no_in_chunk)
{
	static fetch_fn original;
	if (!original) original = (fetch_fn)dlsym(RTLD_NEXT, "fetch_next_read_pair");
	if (keeps_reads(global_context, ginp1)) {
		if (g_kept.phase == 2 && g_kept.usable) { /* a later pass over a kept chunk */
			subread_read_number_t k = __atomic_load_n(&global_context->running_processed_reads_in_chunk, __ATOMIC_RELAXED);
			for (;;) {
				if (k >= g_kept.kept) {
					*read_no_in_chunk = -1;
					return 1;
				}
				if (__atomic_compare_exchange_n(&global_context->running_processed_reads_in_chunk, &k, k + 1, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) break;
			}
			const kept_fragment_t *kf = &g_kept.frag[k];
			const char *p = take(take(take(kf->blob, read_name_1), read_text_1), qual_text_1);
			if (ginp2) take(take(take(p, read_name_2), read_text_2), qual_text_2);
			*read_no_in_chunk = k;
			*read_len_1 = kf->len1;
			if (ginp2) *read_len_2 = kf->len2;
			return 0;
		}
		int rc = original(global_context, thread_context, ginp1, ginp2, read_len_1, read_len_2, read_name_1, read_name_2, read_text_1, read_text_2,
		                  qual_text_1, qual_text_2, remove_color_head, read_no_in_chunk);
		if (g_kept.phase == 1 && rc == 0 && *read_no_in_chunk >= 0)
			keep_fragment(*read_no_in_chunk, ginp2 != NULL, read_name_1, read_text_1, qual_text_1, *read_len_1, read_name_2, read_text_2, qual_text_2,
			              ginp2 ? *read_len_2 : 0);
		return rc;
	}
	if (!chunk_ready(global_context, ginp1, ginp2))
		return original(global_context, thread_context, ginp1, ginp2, read_len_1, read_len_2, read_name_1, read_name_2, read_text_1, read_text_2,
		                qual_text_1, qual_text_2, remove_color_head, read_no_in_chunk);
	/* which read: the counter the original advances under input_lock (core.c:1131,1175), advanced in one atomic step; every pass
	 * starts it from zero (go_chunk_start, core.c:3528) */
	subread_read_number_t k = __atomic_load_n(&global_context->running_processed_reads_in_chunk, __ATOMIC_RELAXED);
	for (;;) {
		if (k >= g_cache.n_frags) {
			*read_no_in_chunk = -1;
			return 1;
		}
		if (__atomic_compare_exchange_n(&global_context->running_processed_reads_in_chunk, &k, k + 1, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) break;
	}
	int rl1 = 0, rl2 = 0;
	copy_read(&g_cache.f[0], k, read_name_1, read_text_1, qual_text_1, &rl1);
	if (ginp2) copy_read(&g_cache.f[1], k, read_name_2, read_text_2, qual_text_2, &rl2);
	if (global_context->config.is_first_read_reversed) { /* core.c:1187-1199 */
		reverse_read(read_text_1, rl1, global_context->config.space_type);
		if (qual_text_1) reverse_quality(qual_text_1, rl1);
	}
	if (ginp2 && global_context->config.is_second_read_reversed) {
		reverse_read(read_text_2, rl2, global_context->config.space_type);
		if (qual_text_2) reverse_quality(qual_text_2, rl2);
	}
	*read_no_in_chunk = k;
	*read_len_1 = rl1;
	if (ginp2) *read_len_2 = rl2;
	return 0;
}
