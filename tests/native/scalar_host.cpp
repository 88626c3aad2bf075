// tests/native/scalar_host.cpp -- TEST ONLY: compiles the product's exact slow path
// (subread_b200/csrc/sr_scalar.cuh) as plain host C++ so that a CPU-only box can check it
// against the oracle.  The shipped library never runs this code on the host.
#define SR_HOST_ONLY 1
#include "../../subread_b200/csrc/sr_scalar.cuh"
#include "../../subread_b200/csrc/sr_config.h"
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" int sr_scalar_host_vote_chunk(int32_t index_gap, int32_t padding, uint32_t buckets_number,
                                         const uint32_t *bucket_start, const int16_t *keys, const uint32_t *vals,
                                         uint32_t start_point, uint32_t length, const uint8_t *values, uint32_t values_bytes,
                                         const uint32_t *read_offsets, int32_t total_offsets, const srcuda_params *p,
                                         const uint8_t *text, const uint64_t *off, const uint16_t *len, int64_t n_frags,
                                         int is_paired, srcuda_mapping_result *out_align, srcuda_subjunc_result *out_junc, uint16_t *out_big_margin,
                                         uint16_t *out_verify)
{
	sr::DevIndex ix{};
	ix.buckets_number = buckets_number; ix.index_gap = index_gap; ix.padding = padding;
	ix.start_point = start_point; ix.length = length; ix.start_base_offset = start_point - start_point % 4;
	ix.values_bytes = values_bytes; ix.total_offsets = total_offsets;
	ix.bucket_start = bucket_start; ix.keys = keys; ix.vals = vals; ix.chr_offsets = read_offsets;
	std::vector<uint32_t> genome((values_bytes + 3) / 4 + 8, 0u);
	memcpy(genome.data(), values, values_bytes);
	ix.genome = genome.data();
	sr::VoteConfig cfg;
	char err[256];
	int rc = sr::make_vote_config(*p, is_paired, ix.index_gap, ix.start_base_offset, ix.length, &cfg, err, sizeof(err));
	if (rc) return rc;
	sr::ScalarScratch *sc = (sr::ScalarScratch *)malloc(sizeof(sr::ScalarScratch));
	const int mates = 1 + is_paired, mb = cfg.multi_best_reads;
	for (int64_t f = 0; f < n_frags; f++) {
		const uint8_t *r0 = text + off[f * mates], *r1 = is_paired ? text + off[f * mates + 1] : nullptr;
		int l0 = len[f * mates], l1 = is_paired ? len[f * mates + 1] : 0;
		sr::scalar_vote_fragment(ix, cfg, r0, l0, r1, l1, (sr::ResultWords *)(out_align + (size_t)f * mates * mb),
		                         out_junc ? (sr::JuncWords *)(out_junc + (size_t)f * mates * mb) : nullptr,
		                         out_big_margin ? out_big_margin + (size_t)f * mates * sr::kBigMarginU16 : nullptr,
		                         out_verify ? out_verify + (size_t)f * mates * mb : nullptr, sc);
	}
	free(sc);
	return 0;
}
