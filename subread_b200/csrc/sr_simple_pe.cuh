// sr_simple_pe.cuh -- tier 0 for paired-end fragments: one THREAD per fragment (two mates), one block per 128 fragments.
//
// The voting of a mate in a strand pass is tier 0's own (sr_simple.cuh: separable passes have closed-form vote
// clusters), run four times per fragment -- do_voting's order: for each strand pass, mate 1 then mate 2 into their own
// tables (core.c:3106-3198), then process_voting_junction_PE_topK (core-junction.c:2199-2530) on the two tables and on
// what the bigtable already holds for the fragment.  That function is replayed per thread on the cluster lists:
//   * per mate the top_scores largest distinct vote levels over table and old results (update_top_three, :908) and the
//     candidate list in the reference's order: level by level, table slots in (row, rank in row) order, then the old
//     results of that level (:2251-2316);
//   * every pair (i, j) of candidates is weighed (test_PE_and_same_chro, core.c:4819: 1300 proper pair, 1000 same
//     chromosome, 800 otherwise) and the max_vote_combinations best kept by (score desc, visiting order asc), put
//     through the reference's unstable selection sort and consumed from the tail (:2321-2372, core.c:4716-4749, :2405);
//   * without any pair the mates fall back to their own candidates of at least minimum_subread_for_first_read votes;
//   * emitted positions are de-duplicated per mate, unused slots get selected_votes = 0 (:2505-2527).
// A fragment whose pass is not separable, whose candidate list passes kPeCand entries, with a mate shorter than 16 bases
// or other than A/C/G/T characters is queued for the warp kernel untouched, as in the single-end tier.  The chunk's
// bigtable slice must be all zero (first index block); verification requests skip the tier.
// Shared memory per thread is what bounds the resident warps here, so only the forward strand of each mate is kept packed
// (reverse-strand keys are derived from it) and the two mates of a strand pass share one pool of vote clusters.
#pragma once
#include "sr_simple.cuh"
#include "sr_simple_junc.cuh"

#ifndef SR_PE_PACK_ALL
#define SR_PE_PACK_ALL 1       // the 32 lanes of a warp pack the 32 staged reads at once (instead of 16 lanes packing two each).  A/B switch
#endif

namespace sr {

constexpr int kPeCand = 8;        // candidates per mate
constexpr int kPeCandJunc = 16;   // ... in subjunc mode, where minimum_subread_for_first_read = 1 makes every one-vote slot of a
                                  // dense table's wrong-strand pass a candidate
constexpr int kPePoolDenseJunc = 32;
constexpr int kPeResWords = 3;    // SimpleResult in shared memory: pos | votes (8 bits) + coverage start, end (12 bits each) | meta (+ used subreads)
constexpr int kPePool = 12;       // vote clusters of both mates of one strand pass, shared (the second mate takes what
constexpr int kPePoolDense = 26;  // the first one left); dense tables: a handful of one-vote random hits per mate
constexpr int kPePoolDense3 = 24; // ... with three-word clusters (reads <= kLongRead): 448 bytes of shared memory per thread, four blocks per SM

SR_HD int simple_pe_warp_bytes(int max_len, int pool, int cw = kSimpleClusterWords)
{
	const int ml = max_len < kSimpleMaxLen ? max_len : kSimpleMaxLen;
	int area = 32 * (((ml + 3) & ~3) + 4) + 32;                               // text of 32 reads (half of the warp's 64)
	const int clus = simple_cluster_bytes(pool, cw);                          // the cluster pool reuses the staging area
	if (area < clus) area = clus;
	area = (area + 15) & ~15;
	return area + 2 * kMaxBest * kPeResWords * 32 * 4;                         // + what the bigtable holds so far
}
SR_HD int simple_pe_area_bytes(int max_len, int pool, int cw = kSimpleClusterWords) { return simple_pe_warp_bytes(max_len, pool, cw) - 2 * kMaxBest * kPeResWords * 32 * 4; }
// Only the forward strand of each mate is kept packed (reverse-strand keys are derived, simple_key_at).
SR_HD int simple_pe_smem_bytes(int max_len, int pool, int cw = kSimpleClusterWords)
{
	return 2 * simple_words(max_len) * kSimpleThreads * 4 + (kSimpleThreads / 32) * simple_pe_warp_bytes(max_len, pool, cw);
}

// subjunc: per thread the subjunc_result_t slots of both mates so far (24 words), those of the mate being emitted (12) and the
// two big-margin records (2 x 12 u16)
constexpr int kPeJuncWordsPerThread = 2 * kMaxBest * 4 + kMaxBest * 4 + 12;
SR_HD int simple_pe_junc_smem_bytes(int max_len, int pool) { return simple_pe_smem_bytes(max_len, pool) + kSimpleThreads * kPeJuncWordsPerThread * 4; }

#if defined(__CUDACC__) && !defined(SR_HOST_ONLY)

// The fragment's result slots so far, [mate][k] x 4 words, interleaved over the warp's lanes.  used_subreads_in_vote
// (<= 63) rides in bits 2-7 of the meta word, whose kind field needs two bits.
struct PeResults {
	uint32_t *base;
	__device__ __forceinline__ uint32_t &w(int m, int k, int i) const { return base[((m * kMaxBest + k) * kPeResWords + i) * 32]; }
	__device__ __forceinline__ int votes(int m, int k) const { return (int)(w(m, k, 1) & 255u); }
	__device__ __forceinline__ void clear_votes(int m, int k) const { w(m, k, 1) &= ~255u; }     // (a record that lost its place keeps its other bytes)
	__device__ __forceinline__ SimpleResult get(int m, int k) const
	{
		SimpleResult r;
		const uint32_t vc = w(m, k, 1), me = w(m, k, 2);
		r.pos = w(m, k, 0); r.votes = (int)(vc & 255u); r.used = (int)((me >> 2) & 63u); r.meta = me & ~0xfcu;
		r.cov = ((vc >> 8) & 0xfffu) | ((vc >> 20) << 16);
		return r;
	}
	__device__ __forceinline__ void put(int m, int k, const SimpleResult &r) const
	{
		w(m, k, 0) = r.pos; w(m, k, 1) = ((uint32_t)r.votes & 255u) | ((r.cov & 0xfffu) << 8) | ((r.cov >> 16) << 20);
		w(m, k, 2) = (r.meta & ~0xfcu) | ((uint32_t)r.used << 2);
	}
	// what the packed form can hold (the read-modify-write form loads records of earlier index blocks)
	static __device__ __forceinline__ bool fits(const SimpleResult &r) { return (unsigned)r.votes <= 255u && (r.cov & 0xffffu) < 4096u && (r.cov >> 16) < 4096u; }
};

// process_voting_junction_PE_topK for one strand pass of a paired-end fragment.  Returns false when a candidate list does
// not fit (the fragment is then queued for the warp kernel).  simple_pass has left every cluster's vote count in word 2.
// JUNC (subjunc): J = both mates' subjunc_result_t slots ([m][k][4] words at stride kSimpleThreads), JT = the three slots of the
// mate being emitted, bm = the two big-margin records (12 u16 each), text0 / text1 = the mates' reads of this pass.
template <int GAP, bool JUNC = false, class CL = SimpleClusters>
__device__ __forceinline__ bool simple_pe_merge(const DevIndex &ix, const VoteConfig &cfg, const CL &C0, const CL &C1,
                                                const PeResults &R, int strand, int applied0, int applied1, int step0, int step1, int len0, int len1,
                                                uint32_t *J = nullptr, uint32_t *JT = nullptr, uint16_t *bm = nullptr,
                                                const TierText *text0 = nullptr, const TierText *text1 = nullptr)
{
	const int mb = cfg.multi_best_reads, ts = cfg.top_scores;
	// Shortcut for the pass on the wrong strand of a fresh fragment (a handful of one-vote random hits per mate on a large
	// index): when no table slot and no old result of either mate reaches minimum_subread_for_first_read, every pair fails
	// the first test of :2333 and the fallback of :2470 emits nothing, so all slots just end with selected_votes = 0.
	{
		int top = 0;
#pragma unroll 1
		for (int m = 0; m < 2; m++) {
			const CL &C = m ? C1 : C0;
#pragma unroll 1
			for (int c = 0; c < C.n; c++) top = max(top, (int)C.w(c, 2));
#pragma unroll 1
			for (int k = 0; k < mb; k++) top = max(top, R.votes(m, k));
		}
		if (top < cfg.min_first && (!JUNC || top < 1)) {       // (subjunc walks the table for its big-margin records whenever a slot has votes)
#pragma unroll 1
			for (int m = 0; m < 2; m++)
#pragma unroll 1
				for (int k = 0; k < mb; k++) {
					R.clear_votes(m, k);
					if (JUNC) J[((m * kMaxBest + k) * 4) * kSimpleThreads] &= 0x0000ffffu;
				}
			return true;
		}
	}
	constexpr int KC = JUNC ? kPeCandJunc : kPeCand;
	unsigned long long cl[2] = {0ull, 0ull}, ch[2] = {0ull, 0ull};   // candidate sources, 8 bits each: cluster index, or 32 + k for an old result
	int nc[2] = {0, 0};
#pragma unroll 1
	for (int m = 0; m < 2; m++) {
		const CL &C = m ? C1 : C0;
		int t0 = 0, t1 = 0, t2 = 0;
#pragma unroll 1
		for (int c = 0; c < C.n + kMaxBest; c++) {
			int v;
			if (c < C.n) v = (int)C.w(c, 2);
			else v = c - C.n < mb ? R.votes(m, c - C.n) : 0;
			if (v > 0) simple_update_top(ts, t0, t1, t2, v);
		}
		int n = 0;
		unsigned long long list = 0ull, list_hi = 0ull;
		// subjunc walks every slot of the table in (row, rank in row) order, once per level and again per emitted slot; picking
		// the next one by a scan over the list is quadratic (a sixth of the kernel's instructions on a dense table with its
		// sixteen one-vote slots), so the visiting order is laid down once: word 3 of list position p = order key << 5 | index
		// of the p-th slot to visit (insertion sort in place; the keys of a pass are distinct).
		int n_sorted = 0;
		if constexpr (JUNC) {
#pragma unroll 1
			for (int c = 0; c < C.n; c++)
				if (C.w(c, 1) >> 31) {
					const uint32_t e = ((uint32_t)C.order(c) << 5) | (uint32_t)c;
					int p = n_sorted;
#pragma unroll 1
					for (; p > 0; p--) {
						const uint32_t before = C.w(p - 1, 3);
						if (before < e) break;
						C.w(p, 3) = before;
					}
					C.w(p, 3) = e;
					n_sorted++;
				}
		}
#pragma unroll 1
		for (int lv = 0; lv < ts; lv++) {
			const int want = lv == 0 ? t0 : (lv == 1 ? t1 : t2);
			if (want < 1 || t0 - want > cfg.max_vote_number_cutoff) break;
			const bool walk_all = JUNC && lv == 0 && cfg.big_margin_filtering;   // insert_big_margin_record visits every slot (:2281)
			if (want >= cfg.min_second || walk_all) {
				const int top_last = ts == 1 ? t0 : (ts == 2 ? t1 : t2);
				unsigned done = 0;
				int p_sorted = 0;
#pragma unroll 1
				for (;;) {                                   // table slots of this level in (row, rank in row) order
					int best = -1;
					if constexpr (JUNC) {
#pragma unroll 1
						for (; p_sorted < n_sorted && best < 0; p_sorted++) {
							const int c = (int)(C.w(p_sorted, 3) & 31u);
							if (walk_all || (int)C.w(c, 2) == want) best = c;
						}
					} else {
						int best_key = 0x7fffffff;
#pragma unroll 1
						for (int c = 0; c < C.n; c++)
							if (!((done >> c) & 1u) && (int)C.w(c, 2) == want) {
								const int key = C.order(c);
								if (key < best_key) { best_key = key; best = c; }
							}
					}
					if (best < 0) break;
					done |= 1u << best;
					const int vv = (int)C.w(best, 2);
					if (walk_all && vv >= top_last) {
						int cs, ce;
						tier_cluster_cov<GAP>(C.w(best, 1), m ? step1 : step0, &cs, &ce);
						sc_big_margin_insert(cfg, bm + 12 * m, vv, cs, ce, m ? len1 : len0, strand != 0);
					}
					if (vv != want || want < cfg.min_second) continue;
					if (n == KC) return false;
					if (n < 8) list |= (unsigned long long)best << (8 * n);
					else list_hi |= (unsigned long long)best << (8 * (n - 8));
					n++;
				}
			}
#pragma unroll 1
			for (int k = 0; k < mb; k++)
				if (R.votes(m, k) == want) {
					if (n == KC) return false;
					if (n < 8) list |= (unsigned long long)(32 + k) << (8 * n);
					else list_hi |= (unsigned long long)(32 + k) << (8 * (n - 8));
					n++;
				}
		}
		cl[m] = list;
		ch[m] = list_hi;
		nc[m] = n;
	}
	const int n0 = nc[0], n1 = nc[1];
	auto cand_src = [&](int m, int i) { return (int)(((i < 8 ? cl[m] : ch[m]) >> (8 * (i & 7))) & 0xffull); };
	auto cand_pos = [&](int m, int src) { return src < 32 ? (m ? C1 : C0).w(src, 0) : R.w(m, src - 32, 0); };
	auto cand_votes = [&](int m, int src) { return src < 32 ? (int)(m ? C1 : C0).w(src, 2) : R.votes(m, src - 32); };

	// ---- pairs (:2321-2372): the best max_vote_combinations by (score desc, visiting order asc) ----
	uint32_t comb0 = 0, comb1 = 0, comb2 = 0;
#pragma unroll 1
	for (int i = 0; i < n0; i++) {
		const int si = cand_src(0, i), v1 = cand_votes(0, si);
		int re1 = 0, chr1 = 0, lp1 = 0;
		bool located = false;
#pragma unroll 1
		for (int j = 0; j < n1; j++) {
			const int sj = cand_src(1, j), v2 = cand_votes(1, sj);
			if (max(v1, v2) < cfg.min_first) continue;
			if (!located) { re1 = locate_position(ix.chr_offsets, ix.total_offsets, ix.padding, cand_pos(0, si), &chr1, &lp1, ix.chr_hint); located = true; }
			int chr2, lp2, is_pe;
			int re2;
			// The second mate's candidates are located once per strand pass: the chromosome found by the binary search is
			// remembered in the cluster's spare word 3 (bit 31 set, re in bits 16-17, chromosome in bits 0-15).
			uint32_t memo = 0u;
			if constexpr (CL::kWords > 3 && !JUNC) memo = sj < 32 ? C1.w(sj, 3) : 0u;   // (three-word clusters: no spare word; subjunc: word 3 holds the visiting order; the hint table made the lookup cheap)
			if (memo >> 31) {
				re2 = (int)((memo >> 16) & 3u);
				if (re2 == 2) re2 = -1;
				chr2 = (int)(memo & 0xffffu);
				lp2 = 0;
				if (re2 == 0) {
					const uint32_t linear = cand_pos(1, sj);
					lp2 = (int)(chr2 == 0 ? linear : linear - ix.chr_offsets[chr2 - 1]) - ix.padding;
				}
			} else {
				re2 = locate_position(ix.chr_offsets, ix.total_offsets, ix.padding, cand_pos(1, sj), &chr2, &lp2, ix.chr_hint);
				if constexpr (CL::kWords > 3 && !JUNC)
					if (sj < 32 && (re2 != 0 || chr2 < 0x10000)) C1.w(sj, 3) = 0x80000000u | ((uint32_t)(re2 < 0 ? 2 : re2) << 16) | (uint32_t)(re2 == 0 ? chr2 : 0);
			}
			const int weight = pair_weight(re1, chr1, lp1, re2, chr2, lp2, len0, len1, cfg.min_pair_dist, cfg.max_pair_dist, &is_pe);
			if (!is_pe && min(v1, v2) < cfg.min_first) continue;
			const uint32_t key = (uint32_t)((v1 + v2) * weight) * 4096u + (uint32_t)(4095 - (i * n1 + j));
			if (key > comb0) { comb2 = comb1; comb1 = comb0; comb0 = key; }
			else if (key > comb1) { comb2 = comb1; comb1 = key; }
			else if (key > comb2) comb2 = key;
		}
	}
	if (cfg.max_vote_combinations < 3) comb2 = 0;
	if (cfg.max_vote_combinations < 2) comb1 = 0;
	const int n_comb = (comb0 != 0) + (comb1 != 0) + (comb2 != 0);
	// merge_sort -> basic_sort_run (core.c:4716-4749): unstable selection sort, ascending score
	{
		auto score = [](uint32_t k) { return (int)(k >> 12); };
		if (n_comb == 2) {
			if (score(comb0) - score(comb1) > 0) { const uint32_t t = comb0; comb0 = comb1; comb1 = t; }
		} else if (n_comb == 3) {
			// i = 0: mn over j = 1, 2
			int mn = 0;
			if (score(comb0) - score(comb1) > 0) mn = 1;
			if (score(mn == 0 ? comb0 : comb1) - score(comb2) > 0) mn = 2;
			if (mn == 1) { const uint32_t t = comb0; comb0 = comb1; comb1 = t; }
			else if (mn == 2) { const uint32_t t = comb0; comb0 = comb2; comb2 = t; }
			// i = 1: j = 2
			if (score(comb1) - score(comb2) > 0) { const uint32_t t = comb1; comb1 = comb2; comb2 = t; }
		}
	}

	// ---- emission (:2386-2527), one mate at a time: the new slots are composed in registers before any is stored ----
#pragma unroll 1
	for (int m = 0; m < 2; m++) {
		const CL &C = m ? C1 : C0;
		const int applied = m ? applied1 : applied0, step = m ? step1 : step0;
		int cursor = 0;
		int esrc0 = 0, esrc1 = 0, esrc2 = 0;
		uint32_t epos0 = 0, epos1 = 0, epos2 = 0;
		auto try_emit = [&](int src) {
			const uint32_t pos = cand_pos(m, src);
			if ((cursor > 0 && epos0 == pos) || (cursor > 1 && epos1 == pos) || (cursor > 2 && epos2 == pos)) return;
			if (cursor == 0) { epos0 = pos; esrc0 = src; }
			else if (cursor == 1) { epos1 = pos; esrc1 = src; }
			else { epos2 = pos; esrc2 = src; }
			cursor++;
		};
		if (n_comb > 0) {
#pragma unroll 1
			for (int c = n_comb - 1; c >= 0 && cursor < mb; c--) {
				const uint32_t key = c == 0 ? comb0 : (c == 1 ? comb1 : comb2);
				const int q = 4095 - (int)(key & 4095u);
				const int i = q / n1, j = q - i * n1;
				try_emit(cand_src(m, m ? j : i));
			}
		} else if (n0 > 0 || n1 > 0) {
#pragma unroll 1
			for (int s = 0; s < nc[m] && cursor < mb; s++) {
				const int src = cand_src(m, s);
				if (cand_votes(m, src) < cfg.min_first) continue;
				try_emit(src);
			}
		}
		SimpleResult nr[kMaxBest];
#pragma unroll
		for (int k = 0; k < kMaxBest; k++) {
			nr[k].pos = 0; nr[k].votes = 0; nr[k].used = 0; nr[k].meta = 0; nr[k].cov = 0;
			if (k < cursor) {
				const int src = k == 0 ? esrc0 : (k == 1 ? esrc1 : esrc2);
				if (src < 32) {
					const uint32_t w1 = C.w(src, 1);
					int s_first, s_last;
					const int o_first = simple_offset<GAP>((int)(w1 & 255u), step, &s_first);
					const int o_last = simple_offset<GAP>((int)((w1 >> 8) & 255u), step, &s_last);
					nr[k].pos = C.w(src, 0); nr[k].votes = (int)C.w(src, 2); nr[k].used = applied;
					nr[k].meta = 1u | ((uint32_t)strand << 8) | ((uint32_t)(s_first + 1) << 16) | ((uint32_t)(s_last + 1) << 24);
					nr[k].cov = (uint32_t)o_first | ((uint32_t)(o_last + 16) << 16);
					if constexpr (JUNC) {                    // the slot's subjunc_result_t: minor half / donor search
						uint32_t *jr = JT + k * 4 * kSimpleThreads;
						jr[0] = jr[kSimpleThreads] = jr[2 * kSimpleThreads] = jr[3 * kSimpleThreads] = 0u;
						nr[k].meta |= tier_minor_half<GAP>(ix, cfg, C, src, step, m ? len1 : len0, m ? *text1 : *text0, jr) << 12;
					}
				} else {
					nr[k] = R.get(m, src - 32);
					if (JUNC) {
#pragma unroll
						for (int i = 0; i < 4; i++) JT[(k * 4 + i) * kSimpleThreads] = J[((m * kMaxBest + src - 32) * 4 + i) * kSimpleThreads];
					}
				}
			}
		}
#pragma unroll
		for (int k = 0; k < kMaxBest; k++) {
			if (k < mb) {
				if (k < cursor) {
					R.put(m, k, nr[k]);
					if (JUNC) {
#pragma unroll
						for (int i = 0; i < 4; i++) J[((m * kMaxBest + k) * 4 + i) * kSimpleThreads] = JT[(k * 4 + i) * kSimpleThreads];
					}
				} else {
					R.clear_votes(m, k);
					if (JUNC) J[((m * kMaxBest + k) * 4) * kSimpleThreads] &= 0x0000ffffu;   // minor_votes = 0
				}
			}
		}
	}
	return true;
}

// k_map_simple_pe: thread t of block b maps paired-end fragment 128 b + t (reads 2f and 2f + 1 of the chunk) of a chunk
// whose bigtable slice is all zero.  Dynamic shared memory: simple_pe_smem_bytes(max_len, clusters).
template <int GAP, int MODE, bool JUNC = false, bool RMW = false, int CW = kSimpleClusterWords>
__global__ void __launch_bounds__(kSimpleThreads, (JUNC || (MODE == 4 && CW > 3)) ? 3 : 4)   // (dense tables, four-word clusters: the pool in shared memory allows three blocks anyway)
k_map_simple_pe(const DevIndex ix, const VoteConfig cfg, const uint8_t *text, uint64_t text_origin, uint64_t text_bytes,
                const uint64_t *off, const uint16_t *len, int64_t n_frags, int max_len, int pool, ResultWords *out, uint32_t *queue,
                uint32_t *queue_count, uint32_t *bad_reads, uint32_t *gather_count, JuncWords *jout = nullptr, uint16_t *bmout = nullptr)
{
	extern __shared__ uint4 smem4[];
	constexpr unsigned kFullMask = 0xffffffffu;
	const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
	const int n_words = simple_words(max_len);
	static_assert(CW == kSimpleClusterWords || (MODE == 4 && !JUNC), "three-word clusters: dense phase pass without the junction half only");
	using CL = SimpleClustersT<CW>;
	const int warp_cap = simple_pe_warp_bytes(max_len, pool, CW), area_cap = simple_pe_area_bytes(max_len, pool, CW);
	uint32_t *pk = reinterpret_cast<uint32_t *>(smem4) + t;                                     // [mate][n_words][128], forward strand
	uint8_t *stage = reinterpret_cast<uint8_t *>(reinterpret_cast<uint32_t *>(smem4) + 2 * n_words * kSimpleThreads) + wid * warp_cap;
	const int64_t fw0 = (int64_t)blockIdx.x * kSimpleThreads + 32 * wid, f = fw0 + lane;
	const int mb = cfg.multi_best_reads;
	if (fw0 >= n_frags) return;

	// ---- 1. stage and pack: the warp's 64 reads in two halves of 32 (fragments fw0 + 16 h .. + 15) ----
	bool active = f < n_frags, punt = false;
	int L0 = 0, L1 = 0;
#pragma unroll 1
	for (int h = 0; h < 2; h++) {
		const int64_t rr = 2 * (fw0 + 16 * h) + lane;                // the read this lane stages
		bool have = rr < 2 * n_frags, bad = false;
		uint64_t o = 0;
		int L = 0;
		if (have) {
			o = off[rr] - text_origin;
			L = len[rr];
			if (o > text_bytes || (uint64_t)L > text_bytes - o || L > kMaxReadLen) { atomicAdd(bad_reads, 1u); bad = true; }
		}
		const bool too_long = L > kSimpleMaxLen || L > max_len;
		const bool want = have && !bad && !too_long && L >= 16;
		unsigned long long lo = want ? o : ~0ull, hi = want ? o + (uint64_t)L : 0ull;
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) {
			lo = min(lo, __shfl_xor_sync(kFullMask, lo, d));
			hi = max(hi, __shfl_xor_sync(kFullMask, hi, d));
		}
		unsigned so = 0;
		bool staged = want;
		if (hi > lo) {
			const uint8_t *g0 = text + lo;
			const unsigned head = (unsigned)(reinterpret_cast<uintptr_t>(g0) & 15u);
			if (hi - lo + head <= (unsigned long long)area_cap) {
				const uint4 *g4 = reinterpret_cast<const uint4 *>(g0 - head);
				const int n4 = (int)((hi - lo + head + 15) >> 4);
				uint4 *s4 = reinterpret_cast<uint4 *>(stage);
#pragma unroll 2
				for (int i = lane; i < n4; i += 32) s4[i] = __ldcs(g4 + i);
				so = (unsigned)(o - lo) + head;
			} else {
				const int slot = ((area_cap - 16) / 32) & ~3;
				if (want) {
					const uint8_t *g = text + o;
					const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(g) & 3u);
					const int nw4 = (int)((mis + (unsigned)L + 3u) >> 2);
					if (nw4 * 4 > slot) staged = false;
					else {
						const uint32_t *gw = reinterpret_cast<const uint32_t *>(g - mis);
						uint32_t *sw = reinterpret_cast<uint32_t *>(stage + lane * slot);
#pragma unroll 1
						for (int i = 0; i < nw4; i++) sw[i] = __ldg(gw + i);
						so = (unsigned)(lane * slot) + mis;
					}
				}
			}
		}
		__syncwarp();
#if SR_PE_PACK_ALL
		// every lane packs the read it staged -- all 32 at once -- into the column of the fragment's owner (thread 16 h + lane / 2 of
		// the warp, mate lane & 1), who then only picks up lengths and flags
		bool unpacked = false;
		if (staged) unpacked = !simple_pack<false>(stage, so, L, reinterpret_cast<uint32_t *>(smem4) + 32 * wid + 16 * h + (lane >> 1) + (lane & 1) * n_words * kSimpleThreads, n_words);
		const int src0 = 2 * (lane & 15);
		const int l0 = __shfl_sync(kFullMask, L, src0), l1 = __shfl_sync(kFullMask, L, src0 + 1);
		const int ok0 = __shfl_sync(kFullMask, (int)(staged && !unpacked), src0), ok1 = __shfl_sync(kFullMask, (int)(staged && !unpacked), src0 + 1);
		const int bad0 = __shfl_sync(kFullMask, (int)bad, src0), bad1 = __shfl_sync(kFullMask, (int)bad, src0 + 1);
		if ((lane >> 4) == h && active) {
			L0 = l0; L1 = l1;
			if (bad0 || bad1) active = false;
			else if (!ok0 || !ok1) punt = true;                       // a mate shorter than 16 bases, longer than the tier takes, odd characters, ...
		}
#else
		// the owner of fragment fw0 + lane (lanes 16 h .. 16 h + 15) picks up its two reads from lanes 2 (lane & 15), + 1
		const int src0 = 2 * (lane & 15);
		const unsigned so0 = __shfl_sync(kFullMask, so, src0), so1 = __shfl_sync(kFullMask, so, src0 + 1);
		const int l0 = __shfl_sync(kFullMask, L, src0), l1 = __shfl_sync(kFullMask, L, src0 + 1);
		const int ok0 = __shfl_sync(kFullMask, (int)staged, src0), ok1 = __shfl_sync(kFullMask, (int)staged, src0 + 1);
		const int bad0 = __shfl_sync(kFullMask, (int)bad, src0), bad1 = __shfl_sync(kFullMask, (int)bad, src0 + 1);
		if ((lane >> 4) == h && active) {
			L0 = l0; L1 = l1;
			if (bad0 || bad1) active = false;
			else if (!ok0 || !ok1) punt = true;                       // a mate shorter than 16 bases, longer than the tier takes, ...
			else {
				if (!simple_pack<false>(stage, so0, l0, pk, n_words)) punt = true;
				if (!simple_pack<false>(stage, so1, l1, pk + n_words * kSimpleThreads, n_words)) punt = true;
			}
		}
#endif
		__syncwarp();
	}

	// ---- 2. four passes (strand x mate), pair scoring after each strand ----
	PeResults R;
	R.base = reinterpret_cast<uint32_t *>(stage + area_cap) + lane;
#pragma unroll 1
	for (int i = 0; i < 2 * kMaxBest * kPeResWords; i++) R.base[i * 32] = 0u;
	if (RMW && active && !punt) {
		// later index block: the slots of the blocks before, if they have the plain form this tier writes itself
		const ResultWords *src = out + (size_t)f * 2 * mb;
#pragma unroll 1
		for (int m = 0; m < 2; m++)
#pragma unroll 1
			for (int k = 0; k < mb; k++) {
				const ResultWords &o = src[m * mb + k];
				const uint32_t w1 = o.w[1], w2 = o.w[2];
				SimpleResult r;
				r.pos = 0; r.meta = 0; r.cov = 0;
				r.votes = (int)(int16_t)(w2 & 0xffffu);
				r.used = (int)(w2 >> 16) & 63;
				if (r.votes >= 1) {
					uint32_t rest = o.w[3] | o.w[16] | (w1 & ~(uint32_t)kCoreNegStrand) | ((w2 >> 16) & ~63u);
#pragma unroll 1
					for (int w = 5; w < 15; w++) rest |= o.w[w];
					const uint32_t w4 = o.w[4], first = w4 & 0xffffu, last = w4 >> 16;
					if (rest != 0 || first < 1 || first > 63 || last < 1 || last > 63) punt = true;
					r.pos = o.w[0];
					r.meta = 1u | ((w1 & (uint32_t)kCoreNegStrand) ? 0x100u : 0u) | (first << 16) | (last << 24);
					r.cov = o.w[15];
				} else if ((w2 >> 16) & ~63u) punt = true;        // (a used_subreads value the packed slot cannot hold)
				if (r.votes < 0 || !PeResults::fits(r)) { punt = true; r.votes = 0; r.cov = 0; }
				R.put(m, k, r);
			}
	}
	uint32_t *jblock = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(smem4) + simple_pe_smem_bytes(max_len, pool, CW));
	uint32_t *J = jblock + t, *JT = J + 2 * kMaxBest * 4 * kSimpleThreads;
	uint16_t *bm = reinterpret_cast<uint16_t *>(jblock + 3 * kMaxBest * 4 * kSimpleThreads + 12 * t);
	if (JUNC) {
#pragma unroll 1
		for (int i = 0; i < 2 * kMaxBest * 4; i++) J[i * kSimpleThreads] = 0u;
#pragma unroll 1
		for (int i = 0; i < 24; i++) bm[i] = 0;
	}
	CL C0, C1;
	C0.base = C1.base = reinterpret_cast<uint32_t *>(stage) + lane;
	C0.n = C1.n = 0;
	C0.cap = C1.cap = pool;
	uint32_t n_gather = 0;
	if (active && !punt) {
		int step0 = 65536, step1 = 65536;
		const int applied0 = subread_schedule(L0, GAP, cfg.total_subreads, &step0);
		const int applied1 = subread_schedule(L1, GAP, cfg.total_subreads, &step1);
		const int rev0 = cfg.first_reversed ? 1 : 0, rev1 = cfg.second_reversed ? 1 : 0;   // reversed at fetch, core.c:1187-1199
#pragma unroll 1
		for (int it = 0; it < 4 && !punt; it++) {                     // one copy of the pass in the instruction cache
			const int strand = it >> 1, m = it & 1;
			const bool rc = (strand ^ (m ? rev1 : rev0)) != 0;
			const uint32_t *pass = pk + m * n_words * kSimpleThreads;
			if (m == 1) {                                             // the second mate's clusters follow the first one's
				C1.base = C0.base + C0.n * CW * 32;
				C1.cap = pool - C0.n;
			}
			bool pass_ok;
			if constexpr (MODE == 4) pass_ok = simple_pass_phase_dense<true, CL>(ix, cfg, pass, m ? L1 : L0, m ? applied1 : applied0, m ? step1 : step0, m ? C1 : C0, rc ? (m ? L1 : L0) : 0, n_gather);
			else pass_ok = simple_pass<GAP, MODE, true>(ix, cfg, pass, m ? L1 : L0, m ? applied1 : applied0, m ? step1 : step0, m ? C1 : C0, rc ? (m ? L1 : L0) : 0, n_gather);
			if (!pass_ok) punt = true;
			else {
				if constexpr (MODE == 4 && !JUNC) {
					// Dense table: most of a mate's clusters are one-vote chance hits.  process_voting_junction_PE_topK only looks at the
					// slots of a mate whose votes lie within max_vote_number_cutoff of that mate's best (table or earlier results,
					// core-junction.c:2251-2316) -- the others neither become candidates nor move a vote level that could -- so they are
					// dropped here, before the second mate takes what is left of the pool and before the merge walks the lists.
					CL &C = m ? C1 : C0;
					int best = 0;
#pragma unroll 1
					for (int c = 0; c < C.n; c++) best = max(best, (int)C.w(c, 2));
#pragma unroll 1
					for (int k = 0; k < mb; k++) best = max(best, R.votes(m, k));
					const int least = max(1, best - cfg.max_vote_number_cutoff);
					int kept = 0;
#pragma unroll 1
					for (int c = 0; c < C.n; c++)
						if ((int)C.w(c, 2) >= least) {
							if (kept != c) {
#pragma unroll
								for (int i = 0; i < CW; i++) C.w(kept, i) = C.w(c, i);
							}
							kept++;
						}
					C.n = kept;
				}
				if (m == 1) {
					if constexpr (JUNC) {
						const TierText tx0{pk, (strand ^ rev0) ? L0 : 0}, tx1{pk + n_words * kSimpleThreads, (strand ^ rev1) ? L1 : 0};
						if (!simple_pe_merge<GAP, true>(ix, cfg, C0, C1, R, strand, applied0, applied1, step0, step1, L0, L1, J, JT, bm, &tx0, &tx1)) punt = true;
					} else if (!simple_pe_merge<GAP, false, CL>(ix, cfg, C0, C1, R, strand, applied0, applied1, step0, step1, L0, L1)) punt = true;
				}
			}
		}
	}
	n_gather = __reduce_add_sync(kFullMask, n_gather);
	if (lane == 0 && n_gather) atomicAdd(gather_count, n_gather);
	__syncwarp();
	const unsigned pm = __ballot_sync(kFullMask, active && punt);
	if (pm) {
		uint32_t qb = 0;
		if (lane == __ffs(pm) - 1) qb = atomicAdd(queue_count, (uint32_t)__popc(pm));
		qb = __shfl_sync(kFullMask, qb, __ffs(pm) - 1);
		if (active && punt) queue[qb + __popc(pm & ((1u << lane) - 1u))] = (uint32_t)f;
	}

	// ---- 3. the warp's bigtable slice: zero-fill, then the few non-zero words on top (see k_map_simple) ----
	if (RMW) {      // later index block: complete records for the emitted slots, the votes / used word for the others
		if (!active || punt) return;
#pragma unroll 1
		for (int m = 0; m < 2; m++)
#pragma unroll 1
			for (int k = 0; k < mb; k++) {
				const SimpleResult r = R.get(m, k);
				ResultWords *d = out + ((size_t)f * 2 + m) * mb + k;
				const uint32_t w2 = (uint32_t)(uint16_t)r.votes | ((uint32_t)(uint16_t)r.used << 16);
				if ((r.meta & 0xffu) == 1u) {
#pragma unroll
					for (int w = 0; w < kResultWords; w++) {
						uint32_t v = 0;
						if (w == 0) v = r.pos;
						else if (w == 1) v = ((r.meta >> 8) & 1u) ? (uint32_t)kCoreNegStrand : 0u;
						else if (w == 2) v = w2;
						else if (w == 4) v = ((r.meta >> 16) & 0xffu) | (((r.meta >> 24) & 0xffu) << 16);
						else if (w == 15) v = r.cov;
						d->w[w] = v;
					}
				} else d->w[2] = w2;
			}
		return;
	}
	{
		const int64_t nf = min((int64_t)32, n_frags - fw0);
		uint32_t *dst = reinterpret_cast<uint32_t *>(out + (size_t)fw0 * 2 * mb);
		const int total = (int)nf * 2 * mb * kResultWords;
		if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
			uint4 *d4 = reinterpret_cast<uint4 *>(dst);
			const int n4 = total >> 2;
#pragma unroll 2
			for (int i = lane; i < n4; i += 32) __stcs(d4 + i, make_uint4(0u, 0u, 0u, 0u));
#pragma unroll 1
			for (int i = (n4 << 2) + lane; i < total; i += 32) __stcs(dst + i, 0u);
		} else {
#pragma unroll 1
			for (int i = lane; i < total; i += 32) __stcs(dst + i, 0u);
		}
	}
	if (JUNC) {     // the same for the warp's subjunc_result_t slots and big-margin records
		const int64_t nf = min((int64_t)32, n_frags - fw0);
		uint32_t *jd = reinterpret_cast<uint32_t *>(jout + (size_t)fw0 * 2 * mb);
		const int jtotal = (int)nf * 2 * mb * 4;
#pragma unroll 1
		for (int i = lane; i < jtotal; i += 32) jd[i] = 0u;
		uint16_t *bd = bmout + (size_t)fw0 * 2 * kBigMarginU16;
		const int btotal = (int)nf * 2 * kBigMarginU16;
#pragma unroll 1
		for (int i = lane; i < btotal; i += 32) bd[i] = 0;
	}
	__syncwarp();
	if (!active || punt) return;
	if (JUNC) {
#pragma unroll 1
		for (int m = 0; m < 2; m++) {
#pragma unroll 1
			for (int k = 0; k < mb; k++)
#pragma unroll
				for (int i = 0; i < 4; i++) {
					const uint32_t w = J[((m * kMaxBest + k) * 4 + i) * kSimpleThreads];
					if (w) jout[((size_t)f * 2 + m) * mb + k].w[i] = w;
				}
#pragma unroll 1
			for (int i = 0; i < 12; i++)
				if (bm[12 * m + i]) bmout[((size_t)f * 2 + m) * kBigMarginU16 + i] = bm[12 * m + i];
		}
	}
#pragma unroll 1
	for (int m = 0; m < 2; m++)
#pragma unroll 1
		for (int k = 0; k < mb; k++) {
			const SimpleResult r = R.get(m, k);
			ResultWords *d = out + ((size_t)f * 2 + m) * mb + k;
			if ((r.meta & 0xffu) == 1u) {
				__stcs(&d->w[0], r.pos);
				__stcs(&d->w[1], (((r.meta >> 8) & 1u) ? (uint32_t)kCoreNegStrand : 0u) | (JUNC ? (r.meta >> 12) & 7u : 0u));
				__stcs(&d->w[2], (uint32_t)(uint16_t)r.votes | ((uint32_t)(uint16_t)r.used << 16));
				__stcs(&d->w[4], ((r.meta >> 16) & 0xffu) | (((r.meta >> 24) & 0xffu) << 16));
				__stcs(&d->w[15], r.cov);
			}
		}
}

#endif  // __CUDACC__

}  // namespace sr
