// sr_config.h -- validation of srcuda_params and their digestion into the kernel's VoteConfig.
// Host-only (shared by the API and the CPU-side tests).
#pragma once
#include "../../include/subread_cuda.h"
#include "sr_common.cuh"
#include <cstdio>

namespace sr {

// Returns SRCUDA_OK or an error code with a message in err.  The limits are the compile-time
// capacities of the kernels; every configuration reachable from the reference's CLIs in DNA mode
// fits (multi_best_reads is fixed to 3 by load_global_context, core.c:4075-4084; top_scores is 3,
// core-indel.c:4415; max_vote_simples / combinations grow to at most 16 with -B).
inline int make_vote_config(const srcuda_params &p, int is_paired, int index_gap, uint32_t start_base_offset,
                            uint32_t length, VoteConfig *c, char *err, size_t errlen)
{
#define SR_REQUIRE(cond, code, ...) do { if (!(cond)) { snprintf(err, errlen, __VA_ARGS__); return code; } } while (0)
	SR_REQUIRE(index_gap >= 1 && index_gap <= 3, SRCUDA_ERR_UNSUPPORTED, "index gap %d not in 1..3", index_gap);
	SR_REQUIRE(p.total_subreads >= 2 && p.total_subreads <= 63, SRCUDA_ERR_ARG, "total_subreads %d not in 2..63", p.total_subreads);
	SR_REQUIRE(p.top_scores >= 1 && p.top_scores <= kMaxTop, SRCUDA_ERR_UNSUPPORTED, "top_scores %d not in 1..%d", p.top_scores, kMaxTop);
	SR_REQUIRE(p.multi_best_reads >= 1 && p.multi_best_reads <= kMaxBest, SRCUDA_ERR_UNSUPPORTED, "multi_best_reads %d not in 1..%d", p.multi_best_reads, kMaxBest);
	SR_REQUIRE(p.max_vote_simples >= 1 && p.max_vote_simples <= kMaxSimples, SRCUDA_ERR_UNSUPPORTED, "max_vote_simples %d not in 1..%d", p.max_vote_simples, kMaxSimples);
	SR_REQUIRE(p.max_vote_combinations >= 1 && p.max_vote_combinations <= kMaxCombos, SRCUDA_ERR_UNSUPPORTED, "max_vote_combinations %d not in 1..%d", p.max_vote_combinations, kMaxCombos);
	SR_REQUIRE(p.max_vote_number_cutoff >= 0, SRCUDA_ERR_ARG, "max_vote_number_cutoff %d < 0", p.max_vote_number_cutoff);
	SR_REQUIRE(p.max_indel_length >= 0, SRCUDA_ERR_ARG, "max_indel_length %d < 0", p.max_indel_length);
	SR_REQUIRE(!p.do_breakpoint_detection || p.max_insertion_at_junctions == 0, SRCUDA_ERR_UNSUPPORTED, "max_insertion_at_junctions %d != 0 is not supported", p.max_insertion_at_junctions);
	SR_REQUIRE(!p.do_breakpoint_detection || (p.big_margin_record_size >= 0 && p.big_margin_record_size <= kBigMarginU16), SRCUDA_ERR_ARG, "big_margin_record_size %d not in 0..27", p.big_margin_record_size);
#undef SR_REQUIRE
	c->total_subreads = p.total_subreads;
	c->top_scores = p.top_scores;
	c->max_vote_simples = p.max_vote_simples;
	c->max_vote_combinations = p.max_vote_combinations;
	c->max_vote_number_cutoff = p.max_vote_number_cutoff;
	c->multi_best_reads = p.multi_best_reads;
	c->min_first = p.minimum_subread_for_first_read;
	c->min_second = p.minimum_subread_for_second_read;
	c->indel_tolerance = p.max_indel_length < 16 ? p.max_indel_length : 16;  // core.c:3059
	c->min_pair_dist = p.minimum_pair_distance;
	c->max_pair_dist = p.maximum_pair_distance;
	c->first_reversed = p.is_first_read_reversed ? 1 : 0;
	c->second_reversed = p.is_second_read_reversed ? 1 : 0;
	c->is_paired = is_paired ? 1 : 0;
	c->low_border = start_base_offset;            // core.c:3081
	c->high_border = start_base_offset + length;  // core.c:3082
	c->do_junctions = p.do_breakpoint_detection ? 1 : 0;
	c->big_margin_filtering = p.do_big_margin_filtering_for_junctions ? 1 : 0;
	c->big_margin_record_size = p.big_margin_record_size;
	c->maximum_intron_length = p.maximum_intron_length;
	c->check_donor = p.check_donor_at_junctions ? 1 : 0;
	c->mismatch_allowed = p.more_accurate_fusions ? 0 : 1;  // mismatch_in_between_allowd, core-junction.c:3735
	c->prefer_donor = p.prefer_donor_receptor_junctions ? 1 : 0;
	return SRCUDA_OK;
}

}  // namespace sr
