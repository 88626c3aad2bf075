// sr_warp.cuh -- the fused seed-and-vote kernel body: one warp maps one fragment.
//
//   pack      ASCII -> 2-bit codes through a 256-entry shared-memory table (genekey2int / reverse_read
//             semantics, input-files.c:1232,1113); the reverse-complement pass is derived from the
//             forward words with bit operations (__brev + pair swap), never from the characters
//   lookup    lane l owns lookup l of the subread schedule (core.c:3117-3171): bucket gather, the
//             reference's own (imin+imax)/2 search over the sorted int16 keys and the equal-key run
//             (sorted-hashtable.c:947-981,1109-1119) -- up to 32 gathers in flight per warp
//   vote      the compact vote table lives in shared memory, slot s owned by lane s; a hit is
//             broadcast and matched by all lanes at once.  When every hit of a batch of lookups
//             agrees on one position the whole batch is applied at once (ballot / popc / reduce_max),
//             which is provably what the sequential replay would leave; otherwise hits are replayed
//             one by one and the first *accepting* slot in the reference's probe order is elected
//             with __reduce_min_sync (gehash_go_X, sorted-hashtable.c:984-1107)
//   top-K     process_voting_junction_PE_topK (core-junction.c:2199-2530) with warp reductions;
//             pair combinations are scored lane-parallel
//   verify    XOR + __popc of packed read vs packed genome (match_chro, gene-value-index.c:856)
//   junction  (subjunc, JUNC) the minor-half / donor search of copy_vote_to_alignment_res on the shared-memory table
//             and the big-margin records (warp_minor_half; shared with the scalar path through sc_donor_score_t)
//
// A table that needs a 33rd slot sets `overflow`; the fragment is then finished by the exact
// scalar path (sr_scalar.cuh) and nothing of it is written here.
#pragma once
#include "sr_common.cuh"
#include "sr_scalar.cuh"   // the junction search (donor score, big-margin record) is shared with the scalar path

namespace sr {

constexpr int kWarpSlots = 32;
constexpr int kRecStride = 22;   // int16 per slot; 11 words -> odd stride, conflict-free
constexpr int kMaxCand = 40;     // <= 32 slots + 3 old results per mate (SE/PE alike)
constexpr unsigned kFull = 0xffffffffu;

struct WarpTable {  // one (mate) vote table, slot s <-> lane s
	uint32_t pos[kWarpSlots];
	int16_t votes[kWarpSlots];
	int16_t last_cluster[kWarpSlots];
	int16_t cov_start[kWarpSlots];
	int16_t cov_end[kWarpSlots];
	int8_t cursor[kWarpSlots];
	int8_t toli[kWarpSlots];
	uint8_t row[kWarpSlots];
	uint8_t j[kWarpSlots];
	int16_t rec[kWarpSlots][kRecStride];
};

template <int MATES>
struct WarpShared {
	WarpTable tab[MATES];
	ResultWords res[MATES][kMaxBest];
	ResultWords tmp[MATES][kMaxBest];
	uint32_t shift_locs[kWarpSlots];
	uint32_t cand_pos[MATES][kMaxCand];
	int32_t cand_lpos[MATES][kMaxCand];   // locate_position results (PE only)
	int16_t cand_chr[MATES][kMaxCand];
	int16_t cand_votes[MATES][kMaxCand];
	uint8_t cand_src[MATES][kMaxCand];    // slot lane, or 32+k for old result k
	int8_t cand_re[MATES][kMaxCand];
	uint32_t comb[kMaxCombos];            // packed (score, r1, r2) after selection
	// subjunc (do_breakpoint_detection) only
	JuncWords jres[MATES][kMaxBest];      // subjunc_result_t of the fragment's slots so far
	JuncWords jtmp[MATES][kMaxBest];
	uint16_t bigm[MATES][kBigMarginU16];  // big-margin record of each mate
	uint8_t order[MATES][kWarpSlots];     // the table's slots in the reference's (row, rank in row) visiting order
};

// Per-mate packed-read area in shared memory, in words: [pad][X: n_words][pad] [A0][A1][A2], see pack_read.
__host__ __device__ constexpr int packed_words_per_mate(int n_words) { return 4 * n_words + 2; }

#if defined(__CUDACC__) && !defined(SR_HOST_ONLY)

// 256-entry character table: bits 0-1 base2int(c) on the sign-extended char (subread.h:238), bit 2 set
// when reverse_read's table (input-files.c:1111) maps c to a base (A, C, G, T, U) rather than to 'N'.
__device__ __forceinline__ void fill_char_table(uint8_t *lut, int tid, int nthreads)
{
	for (int c = tid; c < 256; c += nthreads) {
		int sc = (int)(int8_t)c;
		lut[c] = (uint8_t)(base_code(sc) | (is_table_base(sc) ? 4 : 0));
	}
}

// 16-mer starting at base `o` of an MSB-first packed array (16 bases per word).
__device__ __forceinline__ uint32_t extract_key(const uint32_t *packed, int o)
{
	int w = o >> 4, s = (o & 15) * 2;
	return __funnelshift_l(packed[w + 1], packed[w], s);
}

__device__ __forceinline__ uint32_t reverse_bases(uint32_t g)
{
	uint32_t r = __brev(g);
	return ((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u);
}

// ---- pack ------------------------------------------------------------------------------------
// Builds, for one mate, three MSB-first packed arrays of n_words words each:
//   A0  pass 0 (the read as fetched: reversed at fetch when the mate's strand flag says so)
//   A1  pass 1 (after do_voting's reverse_read, core.c:3229-3234)
//   A2  pass 1 reversed once more = what find_new_indels and the verification see for a
//       positive-strand hit (non-bases have become N, code 3)
// With F = base2int codes of the raw characters and V = 11 per base where the character is a base:
//   raw forward            F
//   reverse_read(raw)      REV(~F | ~V)         (complement 3-f for bases, N=3 otherwise)
//   reverse_read twice     F | ~V
// Bases beyond the read are NUL in the reference's buffer, i.e. code 2.
__device__ __forceinline__ void pack_read(const uint8_t *raw, int L, bool fetched_reversed, uint32_t *area, int n_words,
                                          const uint8_t *lut, int lane)
{
	uint32_t *X = area + 1;                 // X[-1] and X[n_words] are zero pads
	uint32_t *A0 = area + n_words + 2, *A1 = A0 + n_words, *A2 = A1 + n_words;
	if (lane == 0) { area[0] = 0; area[n_words + 1] = 0; }
	const uintptr_t p0 = (uintptr_t)raw, pend = p0 + (uintptr_t)L;
	for (int k = lane; k < n_words; k += 32) {
		uint32_t f = 0xAAAAAAAAu, v = 0;
		const int nvalid = min(16, max(0, L - 16 * k));
		if (nvalid > 0) {
			const uintptr_t p = p0 + 16u * (uintptr_t)k;
			const uintptr_t a = p & ~(uintptr_t)3;
			const unsigned mis = (unsigned)(p & 3) * 8u;
			uint32_t w[5];
#pragma unroll
			for (int q = 0; q < 5; q++) w[q] = (a + 4u * q < pend) ? __ldg((const uint32_t *)(a + 4u * q)) : 0u;
			f = 0; v = 0;
#pragma unroll
			for (int q = 0; q < 4; q++) {
				uint32_t c4 = __funnelshift_r(w[q], w[q + 1], mis);
#pragma unroll
				for (int b = 0; b < 4; b++) {
					const int t = 4 * q + b;
					uint32_t ch = (c4 >> (8 * b)) & 0xffu;
					if (t >= nvalid) ch = 0;
					const uint32_t e = lut[ch];
					f |= (e & 3u) << (30 - 2 * t);
					v |= ((e >> 2) * 3u) << (30 - 2 * t);
				}
			}
		}
		const uint32_t keep = nvalid >= 16 ? 0xffffffffu : (nvalid <= 0 ? 0u : ~(0xffffffffu >> (2 * nvalid)));
		const uint32_t d = ((f | ~v) & keep) | (0xAAAAAAAAu & ~keep);   // reversed twice
		X[k] = ~f | ~v;
		if (!fetched_reversed) { A0[k] = f; A2[k] = d; } else { A1[k] = d; }
	}
	__syncwarp();
	for (int k = lane; k < n_words; k += 32) {
		const int nvalid = min(16, max(0, L - 16 * k));
		uint32_t r = 0xAAAAAAAAu;
		if (nvalid > 0) {
			const int s = L - 16 - 16 * k;              // first source base of this word (may be < 0)
			const int w = s >> 4, sh = (s & 15) * 2;
			r = reverse_bases(__funnelshift_l(X[w + 1], X[w], sh));
			const uint32_t keep = nvalid >= 16 ? 0xffffffffu : ~(0xffffffffu >> (2 * nvalid));
			r = (r & keep) | (0xAAAAAAAAu & ~keep);
		}
		if (!fetched_reversed) A1[k] = r; else { A0[k] = r; A2[k] = r; }
	}
	__syncwarp();
}

// ---- one hit through the vote table ----------------------------------------------------------
struct HitEnv {
	int tol, ii_end, round, lane;
	uint32_t low_border, high_border;
};

// Lane-private slot registers of the table being built (the rest of the slot is in WarpTable).
struct SlotRegs {
	uint32_t pos;
	int row, j, marked;
};

struct TableState {
	int nslots, max_vote, shift_n;
	bool overflow;
};

// The reference's update of one matching slot (sorted-hashtable.c:1020-1070), executed by the lane
// that owns the slot.  Returns acc | mk << 1 | new_votes << 2.
__device__ __forceinline__ int slot_accept(WarpTable &T, SlotRegs &S, uint32_t *shift_locs, int shift_n, const HitEnv &E, int d,
                                           int offset, int sub_p1)
{
	const int lane = E.lane;
	int acc = 0, mk = 0, nv = 0;
	int toli = T.toli[lane];
	if (E.round == 0 && toli > 0 && d == 0 && !S.marked) {
		S.marked = 1;
		shift_locs[shift_n] = S.pos;
		mk = 1;
	}
	int last = T.last_cluster[lane];
	if (sub_p1 == last && toli > 0) {  // roll back the previous indel section (:1027-1039)
		int move = toli >= 3 ? T.rec[lane][toli - 1] : 0;
		int nd = move;
		move -= T.rec[lane][toli + 2];
		nd -= d;
		if (abs(move) > abs(nd)) {
			toli -= 3;
			T.toli[lane] = (int8_t)toli;
			last--;
			T.last_cluster[lane] = (int16_t)last;
			T.votes[lane]--;
		}
	}
	if (!(sub_p1 <= last)) {
		nv = T.votes[lane] + 1;
		T.votes[lane] = (int16_t)nv;
		if (offset + 16 > T.cov_end[lane]) T.cov_end[lane] = (int16_t)(offset + 16);
		if (d == T.cursor[lane]) {
			T.rec[lane][toli + 1] = (int16_t)sub_p1;
		} else {
			toli += 3;
			if (toli < kMaxSections * 3) {
				T.toli[lane] = (int8_t)toli;
				T.rec[lane][toli] = (int16_t)sub_p1;
				T.rec[lane][toli + 1] = (int16_t)sub_p1;
				T.rec[lane][toli + 2] = (int16_t)d;
				if (toli < kMaxSections * 3 - 3) T.rec[lane][toli + 3] = 0;
			}
			T.cursor[lane] = (int8_t)d;
		}
		T.last_cluster[lane] = (int16_t)sub_p1;
		acc = 1;
	}
	return acc | (mk << 1) | (nv << 2);
}

// Open a new slot in the home row of kv (sorted-hashtable.c:1074-1106).  Warp-uniform.
__device__ __forceinline__ void slot_create(WarpTable &T, SlotRegs &S, const uint32_t *shift_locs, TableState &st, const HitEnv &E,
                                            uint32_t kv, int offset, int sub_p1)
{
	const int lane = E.lane;
	if (!(kv >= E.low_border && kv <= E.high_border)) return;
	const uint32_t home = vote_row(kv);
	const int in_home = __popc(__ballot_sync(kFull, lane < st.nslots && (uint32_t)S.row == home));
	if (in_home >= kVoteSpace) return;
	if (st.nslots == kWarpSlots) { st.overflow = true; return; }
	int mk = 0;
	if (E.round > 0) {
		bool near = lane < st.shift_n && kv >= shift_locs[lane] - (uint32_t)E.tol && kv <= shift_locs[lane] + (uint32_t)E.tol;
		mk = __any_sync(kFull, near) ? 1 : 0;
	}
	if (lane == st.nslots) {
		S.pos = kv; S.row = (int)home; S.j = in_home; S.marked = mk;
		T.votes[lane] = 1;
		T.toli[lane] = 0;
		T.rec[lane][0] = (int16_t)sub_p1; T.rec[lane][1] = (int16_t)sub_p1; T.rec[lane][2] = 0; T.rec[lane][3] = 0;
		T.cursor[lane] = 0;
		T.cov_start[lane] = (int16_t)offset;
		T.cov_end[lane] = (int16_t)(offset + 16);
		T.last_cluster[lane] = (int16_t)sub_p1;
	}
	st.nslots++;
	if (st.max_vote == 0) st.max_vote = 1;
}

// Position of my slot's row in the probe order iix = 0,+5,-5,+10,-10,... (sorted-hashtable.c:1007);
// 99 when the row is not probed.  The closed form holds whenever kv +- ii_end does not wrap.
__device__ __forceinline__ int probe_rank(uint32_t kv, int my_row, int ii_end)
{
	if (kv >= 32u && kv <= 0xffffffdfu) {
		int r0 = (int)((kv / kSeg) % kVoteRows);
		int delta = my_row - r0;
		if (delta < 0) delta += kVoteRows;
		int side = ii_end / kSeg;
		if (delta == 0) return 0;
		if (delta <= side) return 2 * delta - 1;
		if (delta >= kVoteRows - side) return 2 * (kVoteRows - delta);
		return 99;
	}
	int rank = 99, p = 0;
	for (int iix = 0; iix <= ii_end; iix = iix > 0 ? -iix : (-iix + kSeg), p++) {
		uint32_t r = vote_row(kv + (uint32_t)iix);
		if (rank == 99 && (uint32_t)my_row == r) rank = p;
	}
	return rank;
}

#ifndef SR_VOTE_HIT_ATTR
#define SR_VOTE_HIT_ATTR __forceinline__
#endif
__device__ SR_VOTE_HIT_ATTR void vote_hit(WarpTable &T, SlotRegs &S, uint32_t *shift_locs, TableState &st, const HitEnv &E, uint32_t kv,
                                         int offset, int sub_p1)
{
	const int lane = E.lane;
	const bool valid = lane < st.nslots;
	const int d = (int)(kv - S.pos);
	const int t = (E.round > 0 && S.marked) ? 0 : E.tol;
	bool cand = valid && d >= -t && d <= t;
	unsigned cm = __ballot_sync(kFull, cand);
	const bool edge = !(kv >= 32u && kv <= 0xffffffdfu);
	if (cm != 0 && ((cm & (cm - 1)) != 0 || edge)) {
		// two or more slots within tolerance (or u32 wrap-around): order them as the reference probes them
		const int rank = probe_rank(kv, S.row, E.ii_end);
		cand = cand && rank != 99;
		const int key = rank * 32 + S.j;
		cm = __ballot_sync(kFull, cand);
		while (cm) {
			const int best = __reduce_min_sync(kFull, cand ? key : 0x7fffffff);
			const bool is_w = cand && key == best;
			int res = 0;
			if (is_w) res = slot_accept(T, S, shift_locs, st.shift_n, E, d, offset, sub_p1);
			const int w = __ffs(__ballot_sync(kFull, is_w)) - 1;
			res = __shfl_sync(kFull, res, w);
			if (res & 2) st.shift_n++;
			if (res & 1) {
				if (st.max_vote < (res >> 2)) st.max_vote = res >> 2;
				return;
			}
			cand = cand && !is_w;
			cm = __ballot_sync(kFull, cand);
		}
	} else if (cm != 0) {
		// exactly one slot within tolerance: it is necessarily in a probed row (|d| <= tol <= ii_end)
		const int w = __ffs(cm) - 1;
		int res = 0;
		if (lane == w) res = slot_accept(T, S, shift_locs, st.shift_n, E, d, offset, sub_p1);
		res = __shfl_sync(kFull, res, w);
		if (res & 2) st.shift_n++;
		if (res & 1) {
			if (st.max_vote < (res >> 2)) st.max_vote = res >> 2;
			return;
		}
	}
	slot_create(T, S, shift_locs, st, E, kv, offset, sub_p1);
}

// A batch of hits (one per lane in `hm`, lanes in lookup order) that all point at the same kv and come
// from strictly increasing subreads.  If no slot is within tolerance, or exactly one is and it sits
// at distance 0 with indel cursor 0 and an older subread, the sequential replay of the batch is: (create
// the slot with the first hit,) then every further hit takes the simple branch of :1041-1052 --
// votes++, coverage_end = max, recorder[toli+1] = last_subread_cluster = its subread.  That end state
// is written here in one step.  Returns false (and changes nothing) when the batch does not qualify.
__device__ __forceinline__ bool vote_batch_uniform(WarpTable &T, SlotRegs &S, uint32_t *shift_locs, TableState &st, const HitEnv &E,
                                                   unsigned hm, uint32_t kv, int my_offset, int my_sub_p1)
{
	const int lane = E.lane;
	if (!(kv >= 32u && kv <= 0xffffffdfu)) return false;
	const bool valid = lane < st.nslots;
	const int d = (int)(kv - S.pos);
	const int t = (E.round > 0 && S.marked) ? 0 : E.tol;
	const unsigned cm = __ballot_sync(kFull, valid && d >= -t && d <= t);
	const int first = __ffs(hm) - 1, n_hits = __popc(hm);
	const bool mine = (hm >> lane) & 1u;
	const int sub_first = __shfl_sync(kFull, my_sub_p1, first);
	const int off_first = __shfl_sync(kFull, my_offset, first);
	const int off_max = __reduce_max_sync(kFull, mine ? my_offset : 0);
	const int sub_last = __reduce_max_sync(kFull, mine ? my_sub_p1 : 0);
	if (cm == 0) {
		const int before = st.nslots;
		slot_create(T, S, shift_locs, st, E, kv, off_first, sub_first);
		if (st.overflow) return true;
		if (st.nslots == before) return true;   // outside the block or row full: every hit of the batch is dropped alike
		if (n_hits > 1) {
			if (lane == before) {
				T.votes[lane] = (int16_t)n_hits;
				T.cov_end[lane] = (int16_t)(off_max + 16);
				T.rec[lane][1] = (int16_t)sub_last;
				T.last_cluster[lane] = (int16_t)sub_last;
			}
			if (st.max_vote < n_hits) st.max_vote = n_hits;
		}
		return true;
	}
	if (cm & (cm - 1)) return false;
	const int w = __ffs(cm) - 1;
	int ok = 0, res = 0;
	if (lane == w) ok = (d == 0 && T.cursor[lane] == 0 && (int)T.last_cluster[lane] < sub_first) ? 1 : 0;
	ok = __shfl_sync(kFull, ok, w);
	if (!ok) return false;
	if (lane == w) {
		const int toli = T.toli[lane];
		int mk = 0;
		if (E.round == 0 && toli > 0 && !S.marked) {
			S.marked = 1;
			shift_locs[st.shift_n] = S.pos;
			mk = 1;
		}
		const int nv = T.votes[lane] + n_hits;
		T.votes[lane] = (int16_t)nv;
		if (off_max + 16 > T.cov_end[lane]) T.cov_end[lane] = (int16_t)(off_max + 16);
		T.rec[lane][toli + 1] = (int16_t)sub_last;
		T.last_cluster[lane] = (int16_t)sub_last;
		res = mk | (nv << 1);
	}
	res = __shfl_sync(kFull, res, w);
	if (res & 1) st.shift_n++;
	if (st.max_vote < (res >> 1)) st.max_vote = res >> 1;
	return true;
}

// ---- lookup + vote for one mate and one strand pass --------------------------------------------
// Returns applied_subreads.  Leaves the table in T (pos/row/j flushed from registers).
template <int GAP>
__device__ __forceinline__ int vote_mate(const DevIndex &ix, const VoteConfig &cfg, const uint32_t *packed /* this pass */, int L,
                                         int applied, int step, WarpTable &T, uint32_t *shift_locs, int lane, int &nslots_out,
                                         int &max_vote_out, bool &overflow)
{
	nslots_out = 0;
	max_vote_out = 0;
	if (L < 16) return 0;  // core.c:3116 (the reference leaves a stale table here; we define it empty)
	const int n_lk = applied * GAP;
	HitEnv E;
	E.tol = cfg.indel_tolerance;
	E.ii_end = kSeg;
	if (E.tol > 5) E.ii_end = (E.tol % kSeg) ? (E.tol - E.tol % kSeg + kSeg) : E.tol;
	E.lane = lane;
	E.low_border = cfg.low_border;
	E.high_border = cfg.high_border - (uint32_t)L;
	const uint32_t nb = ix.buckets_number;
	const uint64_t nb_magic = ix.buckets_magic;

	TableState st;
	st.shift_n = 0;
	st.overflow = false;
	SlotRegs S;
#pragma unroll 1
	for (int round = 0; round < 2; round++) {  // core.c:3153-3186
		E.round = round;
		st.nslots = 0; st.max_vote = 0;
		S.pos = 0; S.row = 255; S.j = 0; S.marked = 0;
		for (int base = 0; base < n_lk; base += 32) {
			const int l = base + lane;
			const int my_sub_p1 = l / GAP + 1;
			int o = 0, h = 0, land = 0;
			uint32_t run_abs = 0, firstval = 0;
			if (l < n_lk) {
				o = lookup_offset(l, GAP, step);
				const uint32_t key = extract_key(packed, o);
				const uint32_t q = (uint32_t)__umul64hi((uint64_t)key, nb_magic);   // key / nb, exact (sr_api.cu)
				const uint32_t b = key - q * nb;
				const int k16 = (int)(int16_t)q;
				bool maybe = true;
				if (ix.bloom) { const uint32_t m = bloom_bits(k16); maybe = ((uint32_t)__ldg(ix.bloom + b) & m) == m; }
				if (ix.direct) {
					// dense table: the direct filter says absent / one position / several (sr_common.cuh)
					uint32_t u;
					const uint32_t bits = direct_bits(ix, key, &u);
					maybe = bits == 3u;                                      // several positions: the bucket path below
					if (bits == 1u) {
						uint32_t again;
						h = 1;
						firstval = ld_direct_first(ix, direct_rank(ld_direct_unit(ix, u), key, u, &again)).x;
					}
				}
				if (maybe) {
					int items = -1;
					if (ix.records) {
						// one 32-byte sector holds the whole bucket when it has <= 5 items
						const Record8 rec = ld_record(reinterpret_cast<const uint4 *>(ix.records), b);
						const uint4 r0 = make_uint4(rec.w[0], rec.w[1], rec.w[2], rec.w[3]), r1 = make_uint4(rec.w[4], rec.w[5], rec.w[6], rec.w[7]);
						const int n = (int)(r0.x & 0xffffu);
						if (n <= kInlineItems) {
							items = 0;
							const int k0 = (int)(int16_t)(r0.x >> 16), k1 = (int)(int16_t)(r0.y & 0xffffu), k2 = (int)(int16_t)(r0.y >> 16),
							          k3 = (int)(int16_t)(r0.z & 0xffffu), k4 = (int)(int16_t)(r0.z >> 16);
							// keys are sorted: the equal-key run is [lt, le)
							const int lt = (n > 0 && k0 < k16) + (n > 1 && k1 < k16) + (n > 2 && k2 < k16) + (n > 3 && k3 < k16) + (n > 4 && k4 < k16);
							const int le = (n > 0 && k0 <= k16) + (n > 1 && k1 <= k16) + (n > 2 && k2 <= k16) + (n > 3 && k3 <= k16) + (n > 4 && k4 <= k16);
							if (le > lt) {
								// where the reference's (imin+imax)/2 search lands inside the run (sorted-hashtable.c:960-981)
								int imin = 0, imax = n - 1, at;
								for (;;) {
									at = (imin + imax) >> 1;
									if (at >= le) imax = at - 1;
									else if (at < lt) imin = at + 1;
									else break;
								}
								h = le - lt;
								land = at - lt;
								firstval = at == 0 ? r0.w : (at == 1 ? r1.x : (at == 2 ? r1.y : (at == 3 ? r1.z : r1.w)));
								if (h > 1) run_abs = ld_gather(ix.bucket_start + b) + (uint32_t)lt;
							}
						}
					}
					if (items < 0) {
						const uint32_t lo = ld_gather(ix.bucket_start + b);
						items = (int)(ld_gather(ix.bucket_start + b + 1) - lo);
						// (a record of a bucket with more than 5 items only carries its size)
						if (items > 0) {
							const int16_t *kp = ix.keys + lo;
							int imin = 0, imax = items - 1, at;
							bool found = false;
							for (;;) {  // sorted-hashtable.c:960-981
								at = (imin + imax) >> 1;
								int kk = (int)__ldg(kp + at);
								if (kk > k16) imax = at - 1;
								else if (kk < k16) imin = at + 1;
								else { found = true; break; }
								if (imax < imin) break;
							}
							if (found) {
								int e = at + 1, s = at;
								while (e < items && (int)__ldg(kp + e) == k16) e++;
								while (s > 0 && (int)__ldg(kp + s - 1) == k16) s--;
								h = e - s;
								land = at - s;
								run_abs = lo + (uint32_t)s;
								firstval = ld_gather(ix.vals + lo + at);
							}
						}
					}
				}
			}
			unsigned fm = __ballot_sync(kFull, h > 0);
			if (fm == 0) continue;
			// fast path: every hit is a single value, all agree on one position, subreads strictly increase
			{
				const uint32_t kv = firstval - (uint32_t)o;
				const int first = __ffs(fm) - 1;
				const uint32_t kv0 = __shfl_sync(kFull, kv, first);
				// two hit lanes of the same subread (GAP lookups each) must be replayed one by one
				bool clash = h > 1 || (h > 0 && kv != kv0);
				if (GAP > 1) {
					const unsigned same_sub = __match_any_sync(kFull, h > 0 ? my_sub_p1 : 0x10000 + lane);
					clash = clash || (same_sub & (same_sub - 1)) != 0;
				}
				if (!__any_sync(kFull, clash)) {
					if (vote_batch_uniform(T, S, shift_locs, st, E, fm, kv0, o, my_sub_p1)) {
						if (st.overflow) { overflow = true; return applied; }
						continue;
					}
				}
			}
			// replay the hits of this batch of lookups in lookup order, each run in visit order
			while (fm) {
				const int src = __ffs(fm) - 1;
				fm &= fm - 1;
				const int o_s = __shfl_sync(kFull, o, src);
				const int h_s = __shfl_sync(kFull, h, src);
				const int sub_p1 = __shfl_sync(kFull, my_sub_p1, src);
				const int land_s = __shfl_sync(kFull, land, src);
				const uint32_t run_s = __shfl_sync(kFull, run_abs, src);
				const uint32_t first_s = __shfl_sync(kFull, firstval, src);
				const int nfwd = h_s - land_s;  // landing .. end of run, then landing-1 .. start (:1109-1119)
				for (int c = 0; c < h_s; c += 32) {
					const int tt = c + lane;
					uint32_t v = first_s;
					if (h_s > 1 && tt < h_s) {
						int idx = tt < nfwd ? land_s + tt : land_s - 1 - (tt - nfwd);
						v = ld_gather(ix.vals + run_s + idx);
					}
					const int cnt = min(32, h_s - c);
					for (int u = 0; u < cnt; u++) {
						const uint32_t vu = __shfl_sync(kFull, v, u);
						vote_hit(T, S, shift_locs, st, E, vu - (uint32_t)o_s, o_s, sub_p1);
						if (st.overflow) { overflow = true; return applied; }
					}
				}
			}
		}
		__syncwarp();
		if (st.shift_n == 0) break;
	}
	T.pos[lane] = S.pos;
	T.row[lane] = (uint8_t)S.row;
	T.j[lane] = (uint8_t)S.j;
	__syncwarp();
	nslots_out = st.nslots;
	max_vote_out = st.max_vote;
	return applied;
}

// ---- subjunc: minor half / donor search on the shared-memory table -----------------------------
// The read of one strand pass as 2-bit codes, MSB first, 16 bases per word (pack_read's A0 / A1).  For a read made of
// A, C, G, T only the codes compare with the genome exactly like the characters do in match_chro's forward branch.
struct PackedText {
	const uint32_t *packed;
	__device__ __forceinline__ int match(const DevIndex &ix, int off, uint32_t pos, int len) const
	{
		if ((uint32_t)(pos + (uint32_t)len) >= ix.length + ix.start_point) return 0;
		if (pos > 0xffff0000u) return 0;
		uint32_t q = pos - ix.start_base_offset;
		if ((q >> 2) >= ix.values_bytes) return 0;
		int ret = 0;
		for (int i = 0; i < len; i++, q++) {
			const int g = (int)((ix.genome[q >> 4] >> (2 * (q & 15))) & 3u);
			const int b = off + i;
			const int c = (int)((packed[b >> 4] >> (30 - 2 * (b & 15))) & 3u);
			ret += g == c;
		}
		return ret;
	}
};

// The do_breakpoint_detection half of copy_vote_to_alignment_res (core-junction.c:1073-1334) for the emitted slot `self`
// of the table, run by ONE lane: the other slots are screened in the reference's visiting order (test_junction_minor :889,
// is_better_inner :975) and sc_donor_score_t picks the split.  Same arithmetic as sc_minor_half (sr_scalar.cuh).
__device__ __noinline__ void warp_minor_half(const DevIndex &ix, const VoteConfig &cfg, ResultWords *r, JuncWords *jr, const WarpTable &T,
                                             const uint8_t *order, int n_slots, int self, int rlen, const uint32_t *packed)
{
	const int sel_votes = (int16_t)(r->w[2] & 0xffffu);
	const uint32_t pm = T.pos[self];
	const int cs_m = T.cov_start[self], ce_m = T.cov_end[self];
	for (int q = 0; q < n_slots; q++) {
		const int s = order[q];
		if (s == self) continue;
		const int mv = T.votes[s];
		if (sel_votes < mv) continue;
		const uint32_t pn = T.pos[s];
		const long long dist = (long long)pm - (long long)pn;
		const int cs_n = T.cov_start[s], ce_n = T.cov_end[s];
		if (sc_abs((int)dist) > cfg.maximum_intron_length) continue;
		if (cs_m == cs_n) continue;
		if (ce_m == ce_n) continue;
		if (cs_m > cs_n) { if (pm < pn) continue; } else { if (pm > pn) continue; }
		const int16_t jr_votes = (int16_t)(jr->w[0] >> 16);
		const int jcov = (int)(uint16_t)(jr->w[3] >> 16) - (int)(uint16_t)(jr->w[3] & 0xffffu);
		uint32_t oi = pm - jr->w[2]; if (oi > 0x7fffffffu) oi = (0xffffffffu - oi) + 1;
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
		int left_indel = 0, major_indels = 0, minor_indel_offset = 0;
		if (rlen > kLongRead) {
			for (int k = 0; k < kMaxSections; k++) { if (!T.rec[self][k * 3]) break; major_indels += T.rec[self][k * 3 + 2]; }
			for (int k = 0; k < kMaxSections; k++) { if (!T.rec[s][k * 3]) break; minor_indel_offset += T.rec[s][k * 3 + 2]; }
			left_indel = pm < pn ? major_indels : minor_indel_offset;
		}
		int split = 0, gt_ag = 0, donor_found = 0;
		const int replace = sc_donor_score_t(ix, cfg, pm < pn ? pm : pn, pm > pn ? pm : pn, left_indel, 0, normally_arranged,
		                                     guess_start > 0 ? guess_start : 0, guess_end < rlen ? guess_end : rlen, PackedText{packed}, rlen,
		                                     &split, &gt_ag, &donor_found);
		if (!replace) continue;
		jr->w[0] = (uint32_t)(uint16_t)split | ((uint32_t)(uint16_t)mv << 16);
		jr->w[1] = (uint32_t)(uint8_t)((minor_indel_offset & 0xf) | ((major_indels & 0xf) << 4)) | (0u << 8) |
		           ((uint32_t)(uint8_t)(!normally_arranged) << 16) | ((uint32_t)(uint8_t)normally_arranged << 24);
		jr->w[2] = pn;
		jr->w[3] = (uint32_t)(uint16_t)cs_n | ((uint32_t)(uint16_t)ce_n << 16);
		uint32_t flags = r->w[1] & 0xffffu;
		flags &= ~0x3u;
		if (!donor_found || gt_ag > 2) flags |= 3u;
		else flags = gt_ag ? (flags | 1u) : (flags & ~1u);
		flags &= ~4u;
		r->w[1] = (r->w[1] & 0xffff0000u) | flags;
	}
}

// ---- vote table -> bigtable (top-K) ------------------------------------------------------------

// Build the mapping_result_t image of slot s into dst, word `lane` by lane `lane`
// (copy_vote_to_alignment_res core-junction.c:1058-1072 + indel_recorder_copy).
__device__ __forceinline__ void emit_slot(ResultWords &dst, const WarpTable &T, int s, int used_subreads, uint32_t strand_flag, int lane)
{
	const int16_t *rec = T.rec[s];
	bool sec = lane < kMaxSections && rec[3 * lane] != 0;
	unsigned m = __ballot_sync(kFull, sec) & 0x7fu;
	int z = __ffs(~m) - 1;  // first empty section (7 when all are in use)
	int ncopy = 3 * z;
	if (lane < kResultWords) {
		uint32_t w = 0;
		if (lane == 0) w = T.pos[s];
		else if (lane == 1) w = strand_flag;
		else if (lane == 2) w = (uint32_t)(uint16_t)T.votes[s] | ((uint32_t)(uint16_t)used_subreads << 16);
		else if (lane == 3) w = ncopy > 0 ? ((uint32_t)(uint8_t)(int8_t)rec[ncopy - 1]) << 8 : 0u;
		else if (lane < 15) {
			int a = 2 * (lane - 4), b = a + 1;
			uint32_t lo16 = a < ncopy ? (uint32_t)(uint16_t)rec[a] : 0u;
			uint32_t hi16 = b < ncopy ? (uint32_t)(uint16_t)rec[b] : 0u;
			w = lo16 | (hi16 << 16);
		} else if (lane == 15) w = (uint32_t)(uint16_t)T.cov_start[s] | ((uint32_t)(uint16_t)T.cov_end[s] << 16);
		dst.w[lane] = w;
	}
}

template <int MATES, bool JUNC>
__device__ __forceinline__ void emit_candidate(const DevIndex &ix, const VoteConfig &cfg, WarpShared<MATES> &ws, int m, int s, int &cursor,
                                               int used_subreads, uint32_t strand_flag, int n_slots, int rlen, const uint32_t *packed, int lane)
{
	const uint32_t pos = ws.cand_pos[m][s];
	bool exist = false;
#pragma unroll 1
	for (int k = 0; k < cursor; k++)
		if (ws.tmp[m][k].w[0] == pos) exist = true;
	if (exist) return;
	const int src = ws.cand_src[m][s];
	if (JUNC && lane < 4) ws.jtmp[m][cursor].w[lane] = src < 32 ? 0u : ws.jres[m][src - 32].w[lane];
	if (src < 32) {
		emit_slot(ws.tmp[m][cursor], ws.tab[m], src, used_subreads, strand_flag, lane);
		if (JUNC) {   // the minor half of a junction read: a serial search over the other slots (core-junction.c:1073-1334)
			__syncwarp();
			if (lane == 0) warp_minor_half(ix, cfg, &ws.tmp[m][cursor], &ws.jtmp[m][cursor], ws.tab[m], ws.order[m], n_slots, src, rlen, packed);
		}
	} else if (lane < kResultWords) ws.tmp[m][cursor].w[lane] = ws.res[m][src - 32].w[lane];
	cursor++;
	__syncwarp();
}

// process_voting_junction_PE_topK, core-junction.c:2199-2530 (DNA mode).
template <int MATES, bool JUNC>
__device__ __forceinline__ void topk_warp(const DevIndex &ix, const VoteConfig &cfg, WarpShared<MATES> &ws, int strand,
                                          const int nslots[2], const int rlen[2], const int all_subreads[2], const uint32_t *const packed[2],
                                          int lane)
{
	const int mb = cfg.multi_best_reads;
	int n_cand[2] = {0, 0};

	for (int m = 0; m < MATES; m++) {
		const WarpTable &T = ws.tab[m];
		const bool valid = lane < nslots[m];
		const int v = valid ? (int)T.votes[lane] : 0;
		int ov = lane < mb ? (int)result_votes(ws.res[m][lane]) : 0;
		if (ov < 0) ov = 0;
		// update_top_three (:908): the top_scores largest distinct positive values
		int top[kMaxTop];
		int bound = 0x7fffffff;
		for (int k = 0; k < kMaxTop; k++) {
			int x = max(v < bound ? v : 0, ov < bound ? ov : 0);
			top[k] = k < cfg.top_scores ? __reduce_max_sync(kFull, x) : 0;
			bound = top[k] > 0 ? top[k] : 0;  // once 0, nothing more qualifies
			if (bound == 0) { for (int k2 = k + 1; k2 < kMaxTop; k2++) top[k2] = 0; break; }
		}
		const int key = valid ? (int)T.row[lane] * 32 + (int)T.j[lane] : 0x7fffffff;
		if (JUNC) {
			// the slots in the reference's visiting order (row, rank in row), for the big-margin record and the minor-half search
			int rank = 0;
#pragma unroll 1
			for (int u = 0; u < nslots[m]; u++) rank += __shfl_sync(kFull, key, u) < key ? 1 : 0;   // (the valid slots are lanes 0 .. nslots - 1)
			if (valid) ws.order[m][rank] = (uint8_t)lane;
			__syncwarp();
			// insert_big_margin_record on the first level's walk over the table (:2281), which stops with the candidate list
			if (cfg.big_margin_filtering && lane == 0 && top[0] >= 1) {
				int seen = 0;
				for (int q = 0; q < nslots[m] && seen < cfg.max_vote_simples; q++) {
					const int sl = ws.order[m][q], vv = T.votes[sl];
					if (vv >= top[cfg.top_scores - 1]) sc_big_margin_insert(cfg, ws.bigm[m], vv, T.cov_start[sl], T.cov_end[sl], rlen[m], strand != 0);
					if (vv == top[0] && vv >= cfg.min_second) seen++;
				}
			}
			__syncwarp();
		}
		// candidate ("simple") list, :2256-2316
		int n = 0;
		for (int lv = 0; lv < cfg.top_scores; lv++) {
			if (n >= cfg.max_vote_simples) break;
			const int want = top[lv];
			if (want < 1 || top[0] - want > cfg.max_vote_number_cutoff) break;
			bool q = valid && v == want && v >= cfg.min_second;
			unsigned qm = __ballot_sync(kFull, q);
			while (qm && n < cfg.max_vote_simples) {  // (row, j) ascending = the reference's i, j loops
				int best = __reduce_min_sync(kFull, q ? key : 0x7fffffff);
				if (q && key == best) {
					ws.cand_pos[m][n] = T.pos[lane];
					ws.cand_votes[m][n] = (int16_t)want;
					ws.cand_src[m][n] = (uint8_t)lane;
					q = false;
				}
				n++;
				qm = __ballot_sync(kFull, q);
			}
			for (int k = 0; k < mb; k++) {
				if (n >= cfg.max_vote_simples) break;
				if ((int)result_votes(ws.res[m][k]) == want) {
					if (lane == 0) {
						ws.cand_pos[m][n] = ws.res[m][k].w[0];
						ws.cand_votes[m][n] = (int16_t)want;
						ws.cand_src[m][n] = (uint8_t)(32 + k);
					}
					n++;
				}
			}
		}
		n_cand[m] = n;
	}
	__syncwarp();

	int n_comb = 0;
	if (MATES == 2) {  // :2321-2372
		const int n0 = n_cand[0], n1 = n_cand[1];
		for (int c = lane; c < n0 + n1; c += 32) {
			int m = c >= n0 ? 1 : 0, i = c - (m ? n0 : 0), chr, lp;
			int re = locate_position(ix.chr_offsets, ix.total_offsets, ix.padding, ws.cand_pos[m][i], &chr, &lp, ix.chr_hint);
			ws.cand_re[m][i] = (int8_t)re;
			ws.cand_chr[m][i] = (int16_t)chr;
			ws.cand_lpos[m][i] = lp;
		}
		__syncwarp();
		// The reference keeps the max_vote_combinations best pairs ordered by (score desc, visiting
		// order asc) through a stable insertion; that is a strict total order, so we select them one
		// by one with a packed key  score * 4096 + (4095 - q),  q = i * n1 + j.
		const int total = n0 * n1;
		uint32_t prev = 0xffffffffu;
		for (int r = 0; r < cfg.max_vote_combinations; r++) {
			uint32_t best = 0;
			for (int q = lane; q < total; q += 32) {
				int i = q / n1, j = q - i * n1;
				int v1 = ws.cand_votes[0][i], v2 = ws.cand_votes[1][j];
				if (max(v1, v2) < cfg.min_first) continue;
				int is_pe;
				int weight = pair_weight(ws.cand_re[0][i], ws.cand_chr[0][i], ws.cand_lpos[0][i], ws.cand_re[1][j], ws.cand_chr[1][j],
				                         ws.cand_lpos[1][j], rlen[0], rlen[1], cfg.min_pair_dist, cfg.max_pair_dist, &is_pe);
				if (!is_pe && min(v1, v2) < cfg.min_first) continue;
				uint32_t k = (uint32_t)((v1 + v2) * weight) * 4096u + (uint32_t)(4095 - q);
				if (k < prev && k > best) best = k;
			}
			best = __reduce_max_sync(kFull, best);
			if (best == 0) break;
			if (lane == 0) ws.comb[r] = best;
			prev = best;
			n_comb++;
		}
		__syncwarp();
		if (n_comb > 1 && lane == 0) {
			// merge_sort -> basic_sort_run (core.c:4716-4749): unstable selection sort, ascending score
			for (int i = 0; i < n_comb - 1; i++) {
				int mn = i;
				for (int j = i + 1; j < n_comb; j++)
					if ((int)(ws.comb[mn] >> 12) - (int)(ws.comb[j] >> 12) > 0) mn = j;
				if (i != mn) { uint32_t t = ws.comb[i]; ws.comb[i] = ws.comb[mn]; ws.comb[mn] = t; }
			}
		}
		__syncwarp();
	}

	// emission into zeroed temporaries (:2386-2495)
	for (int w = lane; w < MATES * kMaxBest * kResultWords; w += 32) (&ws.tmp[0][0].w[0])[w] = 0;
	__syncwarp();
	int cursor[2] = {0, 0};
	const uint32_t strand_flag = strand ? (uint32_t)kCoreNegStrand : 0u;
	if (n_comb > 0) {
		const int n1 = n_cand[1];
		for (int m = 0; m < MATES; m++)
			for (int c = n_comb - 1; c >= 0 && cursor[m] < mb; c--) {
				int q = 4095 - (int)(ws.comb[c] & 4095u);
				int i = q / n1, j = q - i * n1;
				emit_candidate<MATES, JUNC>(ix, cfg, ws, m, m ? j : i, cursor[m], all_subreads[m], strand_flag, nslots[m], rlen[m], packed[m], lane);
			}
	} else {
		if (lane == 0) {
			if (n_cand[0] == 0) ws.res[0][0].w[3] &= ~0xffu;
			if (MATES == 2 && n_cand[1] == 0) ws.res[MATES - 1][0].w[3] &= ~0xffu;
		}
		__syncwarp();
		if (n_cand[0] > 0 || n_cand[1] > 0)
			for (int m = 0; m < MATES; m++)
				for (int s = 0; s < n_cand[m] && cursor[m] < mb; s++) {
					if (ws.cand_votes[m][s] < cfg.min_first) continue;
					emit_candidate<MATES, JUNC>(ix, cfg, ws, m, s, cursor[m], all_subreads[m], strand_flag, nslots[m], rlen[m], packed[m], lane);
				}
	}
	__syncwarp();
#pragma unroll 1
	for (int m = 0; m < MATES; m++)  // :2505-2527
#pragma unroll 1
		for (int k = 0; k < mb; k++) {
			if (k < cursor[m]) {
				if (lane < kResultWords) ws.res[m][k].w[lane] = ws.tmp[m][k].w[lane];
				if (JUNC && lane < 4) ws.jres[m][k].w[lane] = ws.jtmp[m][k].w[lane];
			} else if (lane == 0) {
				ws.res[m][k].w[2] &= 0xffff0000u;
				if (JUNC) ws.jres[m][k].w[0] &= 0x0000ffffu;      // minor_votes = 0 (core-junction.c:2524)
			}
		}
	__syncwarp();
}

// ---- verification ------------------------------------------------------------------------------
// Bases of the read (MSB-first packed codes, L bases) equal to the genome at `pos`: match_chro's
// forward branch (gene-value-index.c:892-940).  Warp-wide; every lane returns the count.
__device__ __forceinline__ int match_packed(const DevIndex &ix, const uint32_t *packed, int L, uint32_t pos, int lane)
{
	if ((uint32_t)(pos + (uint32_t)L) >= ix.length + ix.start_point) return 0;
	if (pos > 0xffff0000u) return 0;
	const uint32_t q0 = pos - ix.start_base_offset;
	if ((q0 >> 2) >= ix.values_bytes) return 0;
	int cnt = 0;
	const int n_words = (L + 15) >> 4;
	for (int k = lane; k < n_words; k += 32) {
		const uint32_t r = reverse_bases(packed[k]);   // base order LSB-first like the genome array
		const uint32_t q = q0 + 16u * (uint32_t)k;
		const uint32_t gw = q >> 4, sh = (q & 15u) * 2u;
		const uint32_t g = __funnelshift_r(__ldg(ix.genome + gw), __ldg(ix.genome + gw + 1), sh);
		uint32_t x = r ^ g;
		x = (x | (x >> 1)) & 0x55555555u;
		const int nbases = min(16, L - 16 * k);
		if (nbases < 16) x &= (1u << (2 * nbases)) - 1u;
		cnt += nbases - __popc(x);
	}
	return __reduce_add_sync(kFull, cnt);
}

// ---- one fragment ------------------------------------------------------------------------------
// area: MATES * packed_words_per_mate(n_words) words of per-warp shared memory.
// Returns false when the fragment overflowed the shared-memory table (nothing was written).
template <int MATES, int GAP, bool VERIFY, bool JUNC>
__device__ __forceinline__ bool map_fragment(const DevIndex &ix, const VoteConfig &cfg, WarpShared<MATES> &ws, uint32_t *area,
                                             int n_words, const uint8_t *lut, const uint8_t *raw0, int len0, const uint8_t *raw1,
                                             int len1, ResultWords *out /* global, this fragment */, bool out_is_zero,
                                             uint16_t *verify /* global or null */, JuncWords *jout, uint16_t *bmout, int lane)
{
	const int mb = cfg.multi_best_reads;
	const int rlen[2] = {len0, len1};
	const uint8_t *raw[2] = {raw0, raw1};
	const int per_mate = packed_words_per_mate(n_words);
	uint32_t *res_flat = &ws.res[0][0].w[0];

	// old results (read-modify-write): ws.res is [MATES][kMaxBest] slots, the bigtable [MATES][mb]
	if (mb == kMaxBest) {
		for (int w = lane; w < MATES * kMaxBest * kResultWords; w += 32) res_flat[w] = out_is_zero ? 0u : (&out[0].w[0])[w];
	} else {
#pragma unroll 1
		for (int m = 0; m < MATES; m++)
#pragma unroll 1
			for (int k = 0; k < mb; k++)
				if (lane < kResultWords) ws.res[m][k].w[lane] = out_is_zero ? 0u : out[m * mb + k].w[lane];
	}
	for (int m = 0; m < MATES; m++)
		pack_read(raw[m], rlen[m], m == 0 ? cfg.first_reversed != 0 : cfg.second_reversed != 0, area + m * per_mate, n_words, lut, lane);
	if (JUNC) {
		// The donor search compares read and genome base by base; on the packed codes that equals match_chro's character
		// compare unless a raw character other than A/C/G/T has code 2 (A2 = "reversed twice" turns those into 3): such
		// fragments take the scalar path with its character buffer.
		__syncwarp();
		bool odd = false;
		for (int m = 0; m < MATES; m++) {
			const uint32_t *A0 = area + m * per_mate + n_words + 2, *A2 = A0 + 2 * n_words;
			for (int k = lane; k < n_words; k += 32) odd = odd || A0[k] != A2[k];
		}
		if (__any_sync(kFull, odd)) return false;
		for (int w = lane; w < MATES * kMaxBest * 4; w += 32) {
			const int slot = w >> 2, m = slot / kMaxBest, k = slot - m * kMaxBest;
			(&ws.jres[0][0].w[0])[w] = (out_is_zero || k >= mb) ? 0u : jout[m * mb + k].w[w & 3];
		}
		for (int w = lane; w < MATES * kBigMarginU16; w += 32) (&ws.bigm[0][0])[w] = out_is_zero ? (uint16_t)0 : bmout[w];
	}

	int all_subreads[2] = {0, 0};
	bool overflow = false;
	// the schedule depends on the read length only: computed once per mate, reused by both strand passes
	int sched_step[2] = {0, 0}, sched_applied[2] = {0, 0};
	for (int m = 0; m < MATES; m++)
		if (rlen[m] >= 16) sched_applied[m] = subread_schedule(rlen[m], GAP, cfg.total_subreads, &sched_step[m]);
#pragma unroll 1
	for (int strand = 0; strand < 2; strand++) {
		int nslots[2] = {0, 0}, max_vote[2] = {0, 0};
		int applied = 0;
#pragma unroll 1
		for (int m = 0; m < MATES; m++) {
			const uint32_t *pass = area + m * per_mate + n_words + 2 + strand * n_words;
			applied = vote_mate<GAP>(ix, cfg, pass, rlen[m], sched_applied[m], sched_step[m], ws.tab[m], ws.shift_locs, lane, nslots[m],
			                         max_vote[m], overflow);
			if (rlen[m] >= 16) all_subreads[m] = applied;
			if (overflow) return false;
		}
		if (MATES == 2 || max_vote[0] >= cfg.min_first) {
			const uint32_t *const packed[2] = {area + n_words + 2 + strand * n_words, area + per_mate + n_words + 2 + strand * n_words};
			topk_warp<MATES, JUNC>(ix, cfg, ws, strand, nslots, rlen, all_subreads, packed, lane);
		} else if (lane == 0 && result_votes(ws.res[0][0]) < 1) {  // core.c:3217-3225
			ws.res[0][0].w[3] &= ~0xffu;
			int16_t used = (int16_t)(ws.res[0][0].w[2] >> 16);
			if (used < applied) ws.res[0][0].w[2] = (ws.res[0][0].w[2] & 0xffffu) | ((uint32_t)(uint16_t)applied << 16);
		}
		__syncwarp();
	}

	if (mb == kMaxBest) {
		for (int w = lane; w < MATES * kMaxBest * kResultWords; w += 32) (&out[0].w[0])[w] = res_flat[w];
	} else {
#pragma unroll 1
		for (int m = 0; m < MATES; m++)
#pragma unroll 1
			for (int k = 0; k < mb; k++)
				if (lane < kResultWords) out[m * mb + k].w[lane] = ws.res[m][k].w[lane];
	}
	if (JUNC) {
		__syncwarp();
		if (mb == kMaxBest) {
			for (int w = lane; w < MATES * kMaxBest * 4; w += 32) (&jout[0].w[0])[w] = (&ws.jres[0][0].w[0])[w];
		} else {
#pragma unroll 1
			for (int m = 0; m < MATES; m++)
#pragma unroll 1
				for (int k = 0; k < mb; k++)
					if (lane < 4) jout[m * mb + k].w[lane] = ws.jres[m][k].w[lane];
		}
		for (int w = lane; w < MATES * kBigMarginU16; w += 32) bmout[w] = (&ws.bigm[0][0])[w];
	}
	if (VERIFY && verify) {
#pragma unroll 1
		for (int m = 0; m < MATES; m++)
#pragma unroll 1
			for (int k = 0; k < mb; k++) {
				const ResultWords &r = ws.res[m][k];
				int v = 0xffff;
				if (result_votes(r) >= 1) {
					bool neg = (r.w[1] & kCoreNegStrand) != 0;
					const uint32_t *arr = area + m * per_mate + n_words + 2 + (neg ? 1 : 2) * n_words;
					v = match_packed(ix, arr, rlen[m], r.w[0], lane);
				}
				if (lane == 0) verify[m * mb + k] = (uint16_t)v;
			}
	}
	return true;
}

#endif  // __CUDACC__

}  // namespace sr
