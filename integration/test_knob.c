/* integration/test_knob.c -- TEST KNOB for the CPU-only bound binaries (oracle/_ref/subread-align-out, subjunc-out):
 * SUBREAD_TEST_READS_PER_CHUNK=<n> replaces config.reads_per_chunk right before the bigtable is sized from it
 * (core-bigtable.c:84), so that a test can walk the chunk loop several times with a few thousand reads.  oracle/ref_hooks.c
 * and integration/do_voting_cuda.c carry the same knob for the stock and the GPU-bound binaries. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdlib.h>

#include "subread.h"
#include "core.h"

int init_bigtable_results(global_context_t *global_context, int is_rewinding)
{
	static int (*real_fn)(global_context_t *, int) = NULL;
	if (!real_fn) real_fn = (int (*)(global_context_t *, int))dlsym(RTLD_NEXT, "init_bigtable_results");
	const char *n = getenv("SUBREAD_TEST_READS_PER_CHUNK");
	if (n && atoll(n) > 0) global_context->config.reads_per_chunk = (unsigned long long)atoll(n);
	return real_fn(global_context, is_rewinding);
}
