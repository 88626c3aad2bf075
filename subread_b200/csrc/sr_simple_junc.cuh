// sr_simple_junc.cuh -- subjunc (do_breakpoint_detection) in the thread-per-fragment tier: the junction half of
// copy_vote_to_alignment_res (core-junction.c:1073-1334) on tier 0's cluster lists.
//
// A separable pass leaves one cluster per vote-table slot with a closed form (sr_simple.cuh), and a cluster carries all the
// minor-half search reads from a slot: position, votes, coverage start / end (offsets of its first / last subread) and an
// indel record without indels.  So the reference's search is replayed per thread:
//   * on the first level's walk over the table every slot goes through insert_big_margin_record (:789-813, :2281);
//   * every slot emitted into the bigtable gets its subjunc_result_t: the other slots are screened in the table's visiting
//     order (test_junction_minor :889, is_better_inner :975) and donor_score (:3675-3832, sc_donor_score_t in sr_scalar.cuh)
//     scans the split point on the packed read;
//   * old results carry their subjunc_result_t along, slots left over get minor_votes = 0 (:2524).
// Reads with characters other than A/C/G/T never reach this point (the tier queues them), so comparing 2-bit codes with the
// genome equals match_chro's character compare.
#pragma once
#include "sr_scalar.cuh"
#include "sr_simple.cuh"

namespace sr {

constexpr int kTierJuncWordsPerThread = 2 * kMaxBest * 4 + 6;   // subjunc results so far + being built, big-margin record (12 u16)
SR_HD int simple_junc_smem_bytes(int max_len, int clusters) { return simple_smem_bytes(max_len, clusters) + kSimpleThreads * kTierJuncWordsPerThread * 4; }

#if defined(__CUDACC__) && !defined(SR_HOST_ONLY)

// The read of the running pass as 2-bit codes in the block's word-interleaved packed array (stride kSimpleThreads).
// rc_len > 0: only the forward array exists (paired-end tier) and the pass runs on the reverse complement of a read of
// that length.
struct TierText {
	const uint32_t *pk;
	int rc_len;
	__device__ __forceinline__ int code(int i) const
	{
		if (rc_len == 0) return (int)((pk[(i >> 4) * kSimpleThreads] >> (30 - 2 * (i & 15))) & 3u);
		const int j = rc_len - 1 - i;
		return 3 - (int)((pk[(j >> 4) * kSimpleThreads] >> (30 - 2 * (j & 15))) & 3u);
	}
	__device__ __forceinline__ int match(const DevIndex &ix, int off, uint32_t pos, int len) const
	{
		if ((uint32_t)(pos + (uint32_t)len) >= ix.length + ix.start_point) return 0;
		if (pos > 0xffff0000u) return 0;
		uint32_t q = pos - ix.start_base_offset;
		if ((q >> 2) >= ix.values_bytes) return 0;
		if (len <= 17 && off >= 0 && (rc_len == 0 || off + len <= rc_len)) {
			// Packed: the window of the genome (2 bits per base, base 0 lowest) against the window of the read in the same layout,
			// XOR, fold the pairs, count.  34 bits at a bit offset of at most 30 fit two words on either side.
			const uint32_t gw = q >> 4;
			const unsigned long long g = (((unsigned long long)ix.genome[gw + 1] << 32) | ix.genome[gw]) >> (2 * (q & 15));
			unsigned long long r;
			if (rc_len == 0) {
				// forward strand: bases off .. off + len - 1, first base highest -> reverse the order of the pairs
				const int w = off >> 4;
				const unsigned long long v = (((unsigned long long)pk[w * kSimpleThreads] << 32) | pk[(w + 1) * kSimpleThreads]) << (2 * (off & 15));
				unsigned long long x = __brevll(v >> (64 - 2 * len)) >> (64 - 2 * len);
				r = ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
			} else {
				// reverse strand: base i is the complement of forward base rc_len - 1 - i, so the forward window [f0, f0 + len) read
				// with its first base highest already has reverse-strand base off + i at bits 2 i
				const int f0 = rc_len - off - len, w = f0 >> 4;
				const unsigned long long v = (((unsigned long long)pk[w * kSimpleThreads] << 32) | pk[(w + 1) * kSimpleThreads]) << (2 * (f0 & 15));
				r = ~(v >> (64 - 2 * len));
			}
			unsigned long long d = r ^ g;
			d = (d | (d >> 1)) & 0x5555555555555555ull & ((1ull << (2 * len)) - 1ull);
			return len - __popcll(d);
		}
		int ret = 0;
		for (int i = 0; i < len; i++, q++) ret += (int)((ix.genome[q >> 4] >> (2 * (q & 15))) & 3u) == code(off + i);
		return ret;
	}
};

// Coverage of a cluster: offsets of its first and last subread (+ 16).
template <int GAP>
__device__ __forceinline__ void tier_cluster_cov(uint32_t w1, int step, int *cs, int *ce)
{
	int s;
	*cs = simple_offset<GAP>((int)(w1 & 255u), step, &s);
	*ce = simple_offset<GAP>((int)((w1 >> 8) & 255u), step, &s) + 16;
}

// The minor-half search for the emitted cluster `self` (sc_minor_half / warp_minor_half on cluster lists).  jr: the slot's
// subjunc_result_t being built, four words at stride kSimpleThreads.  Returns the junction bits of result_flags (0, 1 or 3).
template <int GAP>
__device__ __noinline__ uint32_t tier_minor_half(const DevIndex &ix, const VoteConfig &cfg, const SimpleClusters &C, int self, int step, int rlen,
                                                 const TierText &text, uint32_t *jr)
{
	const int sel_votes = (int)C.w(self, 2);
	const uint32_t pm = C.w(self, 0);
	int cs_m, ce_m;
	tier_cluster_cov<GAP>(C.w(self, 1), step, &cs_m, &ce_m);
	uint32_t flags = 0;
	// Which of the other slots can take part at all: every screen that does not look at the running best (test_junction_minor :889,
	// the overlap and distance limits) in one pass over the list, all lanes together -- on a human-sized index nearly every slot is
	// a one-vote chance hit gigabases away.  The ordered walk below (quadratic in the slots it visits) then sees one or two.
	unsigned done = ~0u;
#pragma unroll 1
	for (int c = 0; c < C.n; c++) {
		if (c == self || !(C.w(c, 1) >> 31)) continue;
		const int mv = (int)C.w(c, 2);
		const uint32_t pn = C.w(c, 0);
		const long long dist = (long long)pm - (long long)pn;
		int cs_n, ce_n;
		tier_cluster_cov<GAP>(C.w(c, 1), step, &cs_n, &ce_n);
		bool in = sel_votes >= mv && sc_abs((int)dist) <= cfg.maximum_intron_length && cs_m != cs_n && ce_m != ce_n;
		if (cs_m > cs_n) { if (pm < pn) in = false; } else { if (pm > pn) in = false; }
		const int overlapped = cs_m > cs_n ? ce_n - cs_m : ce_m - cs_n;
		if (overlapped > 14 || sc_abs((int)dist) < 6) in = false;
		if (in) done &= ~(1u << c);
	}
#pragma unroll 1
	for (;;) {                                                     // the slots left, in (row, rank in row) order
		int s = -1, best_key = 0x7fffffff;
		if (done == ~0u) break;
#pragma unroll 1
		for (int c = 0; c < C.n; c++)
			if (!((done >> c) & 1u)) {
				const int key = C.order(c);
				if (key < best_key) { best_key = key; s = c; }
			}
		if (s < 0) break;
		done |= 1u << s;
		const int mv = (int)C.w(s, 2);
		const uint32_t pn = C.w(s, 0);
		const long long dist = (long long)pm - (long long)pn;
		int cs_n, ce_n;
		tier_cluster_cov<GAP>(C.w(s, 1), step, &cs_n, &ce_n);
		const uint32_t j0 = jr[0], j2 = jr[2 * kSimpleThreads], j3 = jr[3 * kSimpleThreads];
		const int16_t jr_votes = (int16_t)(j0 >> 16);
		const int jcov = (int)(uint16_t)(j3 >> 16) - (int)(uint16_t)(j3 & 0xffffu);
		uint32_t oi = pm - j2; if (oi > 0x7fffffffu) oi = (0xffffffffu - oi) + 1;
		uint32_t ni = pm - pn; if (ni > 0x7fffffffu) ni = (0xffffffffu - ni) + 1;
		const int mcov = ce_n - cs_n;
		const bool better = mv > jr_votes || (mv == jr_votes && mcov > jcov) || (mv == jr_votes && mcov == jcov && (int)ni < (int)oi);
		if (!better) continue;
		const int overlapped = cs_m > cs_n ? ce_n - cs_m : ce_m - cs_n;
		if (overlapped > 14) continue;
		if (sc_abs((int)dist) < 6) continue;
		const int guess_start = cs_m > cs_n ? ce_n - 8 : ce_m - 8;
		const int guess_end = cs_m < cs_n ? cs_n + 8 : cs_m + 8;
		const int normally_arranged = 1 != (cs_m > cs_n) + (pm > pn);
		// (a separable cluster's indel record holds no indel: major_indels = minor_indel_offset = left_indel = 0)
		int split = 0, gt_ag = 0, donor_found = 0;
		const int replace = sc_donor_score_t(ix, cfg, pm < pn ? pm : pn, pm > pn ? pm : pn, 0, 0, normally_arranged, guess_start > 0 ? guess_start : 0,
		                                     guess_end < rlen ? guess_end : rlen, text, rlen, &split, &gt_ag, &donor_found);
		if (!replace) continue;
		jr[0] = (uint32_t)(uint16_t)split | ((uint32_t)(uint16_t)mv << 16);
		jr[kSimpleThreads] = ((uint32_t)(uint8_t)(!normally_arranged) << 16) | ((uint32_t)(uint8_t)normally_arranged << 24);
		jr[2 * kSimpleThreads] = pn;
		jr[3 * kSimpleThreads] = (uint32_t)(uint16_t)cs_n | ((uint32_t)(uint16_t)ce_n << 16);
		flags &= ~0x3u;
		if (!donor_found || gt_ag > 2) flags |= 3u;
		else flags = gt_ag ? (flags | 1u) : (flags & ~1u);
		flags &= ~4u;
	}
	return flags;
}

// simple_merge (sr_simple.cuh) with the junction half: J / JT = the fragment's subjunc_result_t slots so far / being built
// ([k][4] words at stride kSimpleThreads), bm = its big-margin record.  The junction bits of result_flags ride in bits 12-14
// of SimpleResult::meta.
template <int GAP>
__device__ __forceinline__ void simple_merge_junc(const DevIndex &ix, const VoteConfig &cfg, const SimpleClusters &C, int strand, int applied,
                                                  int step, SimpleResult (&res)[kMaxBest], uint32_t *J, uint32_t *JT, uint16_t *bm,
                                                  const TierText &text, int L)
{
	const int mb = cfg.multi_best_reads, ts = cfg.top_scores;
	int max_vote = 0;
#pragma unroll 1
	for (int c = 0; c < C.n; c++) max_vote = max(max_vote, (int)C.w(c, 2));
	if (max_vote < cfg.min_first) {                      // core.c:3217-3225
		if (res[0].votes < 1) {
			if ((res[0].meta & 0xffu) == 0) { res[0].meta = 2u; res[0].used = 0; }
			if (res[0].used < applied) res[0].used = applied;
		}
		return;
	}
	int t0 = 0, t1 = 0, t2 = 0;
#pragma unroll 1
	for (int c = 0; c < C.n + kMaxBest; c++) {
		int v;
		if (c < C.n) v = (int)C.w(c, 2);
		else {
			const int k = c - C.n;
			v = k >= mb ? 0 : (k == 0 ? res[0].votes : (k == 1 ? res[1].votes : res[2].votes));
		}
		if (v > 0) simple_update_top(ts, t0, t1, t2, v);
	}
	const int top_last = ts == 1 ? t0 : (ts == 2 ? t1 : t2);
	SimpleResult tmp[kMaxBest];
#pragma unroll
	for (int k = 0; k < kMaxBest; k++) tmp[k].pos = 0;
	int n = 0, cursor = 0;
	const int max_simples = cfg.max_vote_simples;
	auto exists = [&](uint32_t pos) {
		bool e = false;
#pragma unroll
		for (int k = 0; k < kMaxBest; k++) e = e || (k < cursor && tmp[k].pos == pos);
		return e;
	};
	auto put = [&](const SimpleResult &r) {
#pragma unroll
		for (int k = 0; k < kMaxBest; k++)
			if (k == cursor) tmp[k] = r;
		cursor++;
	};
#pragma unroll 1
	for (int lv = 0; lv < ts; lv++) {
		if (n >= max_simples) break;
		const int want = lv == 0 ? t0 : (lv == 1 ? t1 : t2);
		if (want < 1 || t0 - want > cfg.max_vote_number_cutoff) break;
		unsigned done = 0;
#pragma unroll 1
		for (;;) {       // the table's slots in (row, rank) order; on the first level all of them are visited (big-margin record, :2281)
			int best = -1, best_key = 0x7fffffff;
#pragma unroll 1
			for (int c = 0; c < C.n; c++)
				if (!((done >> c) & 1u) && (C.w(c, 1) >> 31) && (lv == 0 || (int)C.w(c, 2) == want)) {
					const int key = C.order(c);
					if (key < best_key) { best_key = key; best = c; }
				}
			if (best < 0 || n >= max_simples) break;
			done |= 1u << best;
			const int vv = (int)C.w(best, 2);
			int cs, ce;
			tier_cluster_cov<GAP>(C.w(best, 1), step, &cs, &ce);
			if (lv == 0 && cfg.big_margin_filtering && vv >= top_last) sc_big_margin_insert(cfg, bm, vv, cs, ce, L, strand != 0);
			if (vv == want && vv >= cfg.min_second) {
				n++;
				const uint32_t pos = C.w(best, 0);
				if (cursor < mb && want >= cfg.min_first && !exists(pos)) {
					const uint32_t w1 = C.w(best, 1);
					int s_first, s_last;
					simple_offset<GAP>((int)(w1 & 255u), step, &s_first);
					simple_offset<GAP>((int)((w1 >> 8) & 255u), step, &s_last);
					uint32_t *jr = JT + cursor * 4 * kSimpleThreads;
					jr[0] = jr[kSimpleThreads] = jr[2 * kSimpleThreads] = jr[3 * kSimpleThreads] = 0u;
					const uint32_t jflags = tier_minor_half<GAP>(ix, cfg, C, best, step, L, text, jr);
					SimpleResult r;
					r.pos = pos; r.votes = want; r.used = applied;
					r.meta = 1u | ((uint32_t)strand << 8) | (jflags << 12) | ((uint32_t)(s_first + 1) << 16) | ((uint32_t)(s_last + 1) << 24);
					r.cov = (uint32_t)cs | ((uint32_t)ce << 16);
					put(r);
				}
			}
		}
#pragma unroll
		for (int k = 0; k < kMaxBest; k++) {
			if (k < mb && n < max_simples && res[k].votes == want) {
				n++;
				if (cursor < mb && want >= cfg.min_first && !exists(res[k].pos)) {
#pragma unroll
					for (int i = 0; i < 4; i++) JT[(cursor * 4 + i) * kSimpleThreads] = J[(k * 4 + i) * kSimpleThreads];
					put(res[k]);
				}
			}
		}
	}
#pragma unroll
	for (int k = 0; k < kMaxBest; k++) {                 // :2505-2527
		if (k < mb) {
			if (k < cursor) {
				res[k] = tmp[k];
#pragma unroll
				for (int i = 0; i < 4; i++) J[(k * 4 + i) * kSimpleThreads] = JT[(k * 4 + i) * kSimpleThreads];
			} else {
				res[k].votes = 0;
				J[(k * 4) * kSimpleThreads] &= 0x0000ffffu;                       // minor_votes = 0
			}
		}
	}
}

#endif  // __CUDACC__

}  // namespace sr
