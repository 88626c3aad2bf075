/* oracle/ref_hooks.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Linked (together with the reference's own CLI main) against oracle/_ref/libsubread_ref.so,
 * which is the UNMODIFIED reference compiled -fPIC.  Because calls between the reference's
 * translation units go through the PLT, an executable that defines a function of the same
 * name interposes on it.  We use that to look at the state of the real `do_voting`
 * (core.c:3049) without touching a line of the reference:
 *
 *   anti_supporting_read_scan (core-indel.c:268) is the first thing read_chunk_circles
 *   (core.c:3629) calls after every index block of a chunk went through STEP_VOTING.
 *   Our interposer dumps the chunk's bigtable (core-bigtable.c:84-125) and then chains to the
 *   real function, so the run continues to the reference's normal SAM/BAM output.
 *
 *   process_voting_junction (core-junction.c:2607) is what do_voting calls once per strand
 *   with the finished gene_vote_t tables.  Optional dump of every non-empty vote slot.
 *
 * Environment:
 *   SRDUMP_BIGTABLE=<path>  binary dump, appended per chunk:
 *        header  { u32 magic 'SRBT', u32 n_frags, u32 mates, u32 multi_best, u32 has_junc, u32 bm_size }
 *        n_frags*mates*multi_best * mapping_result_t (68 B each)
 *        if has_junc: n_frags*mates*multi_best * subjunc_result_t (16 B) and
 *                     n_frags*mates * 27 * u16 big-margin record
 *   SRDUMP_VOTES=<path>     text dump of vote tables (debugging aid)
 *   SRDUMP_TIMING=<path>    text, one line per chunk: the reference's OWN accumulated clocks (core.c:3553-3640: miltime()
 *                           around gehash_load/gvindex_load and around run_maybe_threads(STEP_VOTING)) as they stand when
 *                           the chunk's voting is over -- "chunk_end frags=<n> voting_s=<timecost_voting>
 *                           load_index_s=<timecost_load_index>".  bench.py's reference arm reads the voting step's time
 *                           from it; nothing of the reference is changed or re-timed.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "subread.h"
#include "core.h"

int anti_supporting_read_scan(global_context_t *global_context)
{
	static int (*real_fn)(global_context_t *) = NULL;
	if (!real_fn) real_fn = (int (*)(global_context_t *))dlsym(RTLD_NEXT, "anti_supporting_read_scan");

	const char *tpath = getenv("SRDUMP_TIMING");
	if (tpath) {
		FILE *tf = fopen(tpath, "a");
		if (tf) {
			fprintf(tf, "chunk_end frags=%llu voting_s=%.6f load_index_s=%.6f\n", (unsigned long long)global_context->processed_reads_in_chunk,
			        global_context->timecost_voting, global_context->timecost_load_index);
			fclose(tf);
		}
	}
	const char *path = getenv("SRDUMP_BIGTABLE");
	if (path) {
		FILE *fp = fopen(path, "ab");
		if (!fp) { perror("SRDUMP_BIGTABLE"); exit(97); }
		unsigned int n_frags = (unsigned int)global_context->processed_reads_in_chunk;
		unsigned int mates = 1 + global_context->input_reads.is_paired_end_reads;
		unsigned int mb = global_context->config.multi_best_reads;
		unsigned int has_junc = global_context->config.do_breakpoint_detection ? 1 : 0;
		unsigned int hdr[6] = {0x54425253u, n_frags, mates, mb, has_junc,
		                       (unsigned int)global_context->config.big_margin_record_size};
		fwrite(hdr, 4, 6, fp);
		fwrite(global_context->bigtable_cache_malloc_ptr, sizeof(mapping_result_t), (size_t)n_frags * mates * mb, fp);
		if (has_junc) {
			fwrite(global_context->bigtable_cache_malloc_ptr_junc, sizeof(subjunc_result_t), (size_t)n_frags * mates * mb, fp);
			size_t x;
			for (x = 0; x < (size_t)n_frags * mates; x++)
				fwrite(global_context->bigtable_cache[x].big_margin_data, sizeof(unsigned short), 27, fp);
		}
		fclose(fp);
	}
	return real_fn(global_context);
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

static void dump_vote(FILE *fp, gene_vote_t *v, int mate)
{
	int i, j, k;
	for (i = 0; i < GENE_VOTE_TABLE_SIZE; i++)
		for (j = 0; j < v->items[i]; j++) {
			fprintf(fp, "  m%d [%d][%d] pos=%u votes=%d mask=%d cov=%d-%d last=%d toli=%d cur=%d marked=%d rec=",
			        mate, i, j, v->pos[i][j], v->votes[i][j], v->masks[i][j], v->coverage_start[i][j],
			        v->coverage_end[i][j], v->last_subread_cluster[i][j], v->toli[i][j],
			        v->current_indel_cursor[i][j], v->marked_shift_indel[i][j]);
			for (k = 0; k < MAX_INDEL_TOLERANCE * 3 && v->indel_recorder[i][j][k]; k += 3)
				fprintf(fp, "%d,%d,%d;", v->indel_recorder[i][j][k], v->indel_recorder[i][j][k + 1], v->indel_recorder[i][j][k + 2]);
			fputc('\n', fp);
		}
}

int process_voting_junction(global_context_t *global_context, thread_context_t *thread_context, subread_read_number_t pair_number,
                            gene_vote_t *vote_1, gene_vote_t *vote_2, char *read_name_1, char *read_name_2, char *read_text_1,
                            char *read_text_2, int read_len_1, int read_len_2, int is_negative_strand,
                            gene_vote_number_t v1_all_subreads, gene_vote_number_t v2_all_subreads)
{
	static int (*real_fn)(global_context_t *, thread_context_t *, subread_read_number_t, gene_vote_t *, gene_vote_t *, char *, char *,
	                      char *, char *, int, int, int, gene_vote_number_t, gene_vote_number_t) = NULL;
	static FILE *vfp = NULL;
	static int vfp_tried = 0;
	if (!real_fn) real_fn = dlsym(RTLD_NEXT, "process_voting_junction");
	if (!vfp_tried) {
		const char *path = getenv("SRDUMP_VOTES");
		vfp_tried = 1;
		if (path) vfp = fopen(path, "w");
	}
	if (vfp) {
		fprintf(vfp, "FRAG %lld strand=%d max1=%d\n", (long long)pair_number, is_negative_strand, vote_1->max_vote);
		dump_vote(vfp, vote_1, 0);
		if (read_text_2) dump_vote(vfp, vote_2, 1);
		fflush(vfp);
	}
	return real_fn(global_context, thread_context, pair_number, vote_1, vote_2, read_name_1, read_name_2, read_text_1, read_text_2,
	               read_len_1, read_len_2, is_negative_strand, v1_all_subreads, v2_all_subreads);
}
