// sr_simple.cuh -- tier 0 of the voting step: one THREAD per single-end fragment.
//
// The warp kernel (sr_warp.cuh) spends most of its instructions on warp-uniform work; for the
// common fragment that work is tiny.  A strand pass is "separable" when
//   * every lookup returns at most one position,
//   * the distinct values kv = position - offset that occur are pairwise more than the indel
//     tolerance apart, and
//   * the hits of one kv come from distinct subreads.
// Then no hit is ever within tolerance of another slot, so gehash_go_X (sorted-hashtable.c:937-1123)
// builds exactly one vote slot per distinct kv -- created by its first hit (or never, when kv lies
// outside the block's border, :1074), every later hit taking the plain branch of :1041-1052 -- and
// each slot has a closed form:
//     votes = #hits, coverage = [offset(first hit), offset(last hit) + 16),
//     indel recorder = {subread(first)+1, subread(last)+1, 0},    row = (kv/5)%30, j = creation rank in its row,
// with no shift-indel round (toli stays 0).  Spurious single hits -- the usual companions of the true
// cluster on a large index -- are just extra one-vote slots.  process_voting_junction_PE_topK
// (core-junction.c:2199-2530) is then replayed on that small slot list and on the results of the first
// strand pass (the chunk's bigtable slice must be all zero, i.e. the first index block).
//
// Everything else -- a lookup with several positions, kv values within tolerance of each other
// (indels), two phases of one subread hitting one kv, more than kSimpleClusters distinct kv,
// characters that are not A/C/G/T, reads longer than kSimpleMaxLen -- is PUNTED untouched to the
// warp kernel through a queue, so the tiered path returns the reference's result for every input.
// Paired-end and subjunc chunks, chunks merged into a non-zero bigtable, and verification requests
// skip this tier altogether.
#pragma once
#include "sr_common.cuh"

namespace sr {

constexpr int kSimpleMaxLen = 256;                       // longer reads go to the warp kernel
constexpr int kSimpleWords = kSimpleMaxLen / 16 + 1;     // packed words per strand array
constexpr int kSimpleClusters = 8;                       // distinct kv per strand pass

#if defined(__CUDACC__) && !defined(SR_HOST_ONLY)

struct SimpleClusters {     // the vote table of a separable pass, in creation order
	int n;
	uint32_t kv[kSimpleClusters];
	int16_t votes[kSimpleClusters];        // 0: kv outside the border, no slot (its hits are dropped)
	int16_t first_off[kSimpleClusters], last_off[kSimpleClusters];
	uint8_t first_sub[kSimpleClusters], last_sub[kSimpleClusters];   // subread + 1
	uint8_t row[kSimpleClusters], j[kSimpleClusters];
};

struct SimpleResult {       // one bigtable slot in compact form
	uint8_t kind;           // 0 untouched (all zero), 1 a record, 2 only used_subreads_in_vote is set
	uint8_t strand, first_sub, last_sub;
	int16_t votes;          // selected_votes as stored (0 for a record that lost its place: stale bytes stay)
	int16_t used;
	uint32_t pos;
	int16_t cov_start, cov_end;
};

// One bucket lookup by one thread: number of positions with key16 == k16 in bucket b, and the
// position when there is exactly one.  Same accelerators as the warp kernel.
__device__ __forceinline__ int simple_lookup(const DevIndex &ix, uint32_t b, int k16, uint32_t *value)
{
	if (ix.bloom && !((__ldg(ix.bloom + b) >> (k16 & 15)) & 1)) return 0;
	if (ix.records) {
		const uint4 *rp = reinterpret_cast<const uint4 *>(ix.records) + 2 * (size_t)b;
		const uint4 r0 = __ldg(rp);
		const int n = (int)(r0.x & 0xffffu);
		if (n <= kInlineItems) {
			if (n == 0) return 0;
			const int k0 = (int)(int16_t)(r0.x >> 16), k1 = (int)(int16_t)(r0.y & 0xffffu), k2 = (int)(int16_t)(r0.y >> 16),
			          k3 = (int)(int16_t)(r0.z & 0xffffu), k4 = (int)(int16_t)(r0.z >> 16);
			const int e0 = k0 == k16, e1 = n > 1 && k1 == k16, e2 = n > 2 && k2 == k16, e3 = n > 3 && k3 == k16, e4 = n > 4 && k4 == k16;
			const int h = e0 + e1 + e2 + e3 + e4;
			if (h == 1) {
				if (e0) *value = r0.w;
				else {
					const uint4 r1 = __ldg(rp + 1);
					*value = e1 ? r1.x : (e2 ? r1.y : (e3 ? r1.z : r1.w));
				}
			}
			return h;
		}
	}
	const uint32_t lo = __ldg(ix.bucket_start + b);
	const int items = (int)(__ldg(ix.bucket_start + b + 1) - lo);
	if (items <= 0) return 0;
	const int16_t *kp = ix.keys + lo;
	int imin = 0, imax = items - 1, at;
	for (;;) {
		at = (imin + imax) >> 1;
		const int kk = (int)__ldg(kp + at);
		if (kk > k16) imax = at - 1;
		else if (kk < k16) imin = at + 1;
		else break;
		if (imax < imin) return 0;
	}
	if ((at + 1 < items && (int)__ldg(kp + at + 1) == k16) || (at > 0 && (int)__ldg(kp + at - 1) == k16)) return 2;  // a run: not simple
	*value = __ldg(ix.vals + lo + at);
	return 1;
}

// Pack the thread's read: forward codes fw[] (MSB first, 16 bases per word, code 2 beyond the read like
// the reference's NUL) and reverse-complement codes rc[].  Returns false when a character is not one
// of A, C, G, T (then base2int / reverse_read quirks apply and the fragment is punted).
__device__ __forceinline__ bool simple_pack(const uint8_t *raw, int L, uint32_t *fw, uint32_t *rc)
{
	const int n_words = (L + 15) / 16 + 1;
	const uintptr_t p0 = (uintptr_t)raw, a0 = p0 & ~(uintptr_t)3, pend = p0 + (uintptr_t)L;
	const unsigned mis = (unsigned)(p0 & 3) * 8u;
	bool ok = true;
	uint32_t wprev = __ldg((const uint32_t *)a0);
	for (int k = 0; k < n_words; k++) {
		uint32_t f = 0;
		const int nvalid = min(16, max(0, L - 16 * k));
		if (nvalid == 0) { fw[k] = 0xAAAAAAAAu; continue; }
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const uintptr_t an = a0 + 16u * (uintptr_t)k + 4u * (q + 1);
			const uint32_t wnext = an < pend ? __ldg((const uint32_t *)an) : 0u;
			const uint32_t c4 = __funnelshift_r(wprev, wnext, mis);
			wprev = wnext;
			// A 0x41, C 0x43, G 0x47, T 0x54: code bit0 = bit 2, code bit1 = bit 1 ^ bit 2  (A0 G1 C2 T3)
			const uint32_t b2 = (c4 >> 2) & 0x01010101u, b1 = ((c4 >> 1) ^ (c4 >> 2)) & 0x01010101u;
			const uint32_t code4 = b2 | (b1 << 1);
			// validate by rebuilding the characters from the codes: "AGCT"[code]
			const uint32_t sel = (code4 & 3u) | ((code4 >> 4) & 0x30u) | ((code4 >> 8) & 0x300u) | ((code4 >> 12) & 0x3000u);
			const uint32_t rebuilt = __byte_perm(0x54434741u, 0u, sel);
			const int nb = nvalid - 4 * q;  // characters of this group inside the read
			uint32_t diff = rebuilt ^ c4;
			if (nb < 4) diff = nb <= 0 ? 0u : (diff & (0xffffffffu >> (8 * (4 - nb))));
			ok = ok && diff == 0;
			f |= ((code4 * 0x40100401u) >> 24) << (24 - 8 * q);
		}
		const uint32_t keep = nvalid >= 16 ? 0xffffffffu : ~(0xffffffffu >> (2 * nvalid));
		fw[k] = (f & keep) | (0xAAAAAAAAu & ~keep);
	}
	// reverse complement of an all-base read: rc base i = 3 - f base (L-1-i)
	for (int k = 0; k < n_words; k++) {
		const int nvalid = min(16, max(0, L - 16 * k));
		uint32_t r = 0xAAAAAAAAu;
		if (nvalid > 0) {
			const int s = L - 16 - 16 * k;
			const int w = s >> 4, sh = (s & 15) * 2;
			const uint32_t hi = w >= 0 ? ~fw[w] : 0u, lo = ~fw[w + 1];
			uint32_t g = __funnelshift_l(lo, hi, sh);
			g = __brev(g);
			g = ((g & 0x55555555u) << 1) | ((g >> 1) & 0x55555555u);
			const uint32_t keep = nvalid >= 16 ? 0xffffffffu : ~(0xffffffffu >> (2 * nvalid));
			r = (g & keep) | (0xAAAAAAAAu & ~keep);
		}
		rc[k] = r;
	}
	return ok;
}

// One strand pass.  Returns false when the pass is not separable.
template <int GAP>
__device__ __forceinline__ bool simple_pass(const DevIndex &ix, const VoteConfig &cfg, const uint32_t *packed, int L, int applied, int step,
                                            SimpleClusters &C)
{
	const uint32_t nb = ix.buckets_number;
	const uint64_t magic = ix.buckets_magic;
	const int n_lk = applied * GAP, tol = cfg.indel_tolerance;
	const uint32_t high = cfg.high_border - (uint32_t)L;
	C.n = 0;
	for (int l = 0; l < n_lk; l++) {
		const int o = lookup_offset(l, GAP, step);
		const int w = o >> 4, s = (o & 15) * 2;
		const uint32_t key = __funnelshift_l(packed[w + 1], packed[w], s);
		const uint32_t q = (uint32_t)__umul64hi((uint64_t)key, magic);
		const uint32_t b = key - q * nb;
		uint32_t value = 0;
		const int h = simple_lookup(ix, b, (int)(int16_t)q, &value);
		if (h > 1) return false;
		if (h == 1) {
			const uint32_t kv = value - (uint32_t)o;
			if (!(kv >= 32u && kv <= 0xffffffdfu)) return false;   // u32 wrap-around of the probe rows
			const int sub_p1 = l / GAP + 1;
			int found = -1;
			for (int c = 0; c < C.n; c++) {
				const int d = (int)(kv - C.kv[c]);
				if (d == 0) found = c;
				else if (d >= -tol && d <= tol) return false;          // within tolerance of another slot: an indel vote
			}
			if (found >= 0) {
				if (C.votes[found] == 0) continue;                      // no slot for this kv: dropped again
				if (sub_p1 == (int)C.last_sub[found]) return false;     // same subread twice: the reference opens a duplicate slot
				C.votes[found]++;
				C.last_off[found] = (int16_t)o;
				C.last_sub[found] = (uint8_t)sub_p1;
			} else {
				if (C.n == kSimpleClusters) return false;
				const int c = C.n++;
				C.kv[c] = kv;
				C.votes[c] = 0;
				if (kv >= cfg.low_border && kv <= high) {               // :1074
					const int row = (int)vote_row(kv);
					int jj = 0;
					for (int e = 0; e < c; e++) jj += (C.votes[e] > 0 && (int)C.row[e] == row) ? 1 : 0;
					C.votes[c] = 1;
					C.first_off[c] = C.last_off[c] = (int16_t)o;
					C.first_sub[c] = C.last_sub[c] = (uint8_t)sub_p1;
					C.row[c] = (uint8_t)row;
					C.j[c] = (uint8_t)jj;
				}
			}
		}
	}
	return true;
}

__device__ __forceinline__ void simple_update_top(int top_scores, int *top, int v)   // update_top_three core-junction.c:908
{
	if (v > top[top_scores - 1])
		for (int x = 0; x < top_scores; x++) {
			if (v > top[x]) {
				for (int y = top_scores - 1; y > x; y--) top[y] = top[y - 1];
				top[x] = v;
				break;
			} else if (v == top[x]) break;
		}
}

// process_voting_junction_PE_topK for one single-end pass over a separable table (core-junction.c:2199-2530,
// the SE branch :2436-2495) and the hand-over logic of do_voting (core.c:3212-3226).
__device__ __forceinline__ void simple_merge(const VoteConfig &cfg, const SimpleClusters &C, int strand, int applied, SimpleResult *res)
{
	const int mb = cfg.multi_best_reads;
	int max_vote = 0;
	for (int c = 0; c < C.n; c++) max_vote = max(max_vote, (int)C.votes[c]);
	if (max_vote < cfg.min_first) {                      // core.c:3217-3225
		if (res[0].votes < 1) {
			if (res[0].kind == 0) { res[0].kind = 2; res[0].used = 0; }
			if (res[0].used < applied) res[0].used = (int16_t)applied;
		}
		return;
	}
	int top[kMaxTop] = {0, 0, 0};
	for (int c = 0; c < C.n; c++)
		if (C.votes[c] > 0) simple_update_top(cfg.top_scores, top, C.votes[c]);
	for (int k = 0; k < mb; k++)
		if (res[k].votes > 0) simple_update_top(cfg.top_scores, top, res[k].votes);

	// candidate list: code < 16 = cluster index, 16 + k = bigtable slot k
	int list[kMaxSimples > 16 ? 16 : kMaxSimples], n = 0;
	const int max_simples = min(cfg.max_vote_simples, 16);   // <= 8 clusters + 3 slots can ever qualify
	for (int lv = 0; lv < cfg.top_scores; lv++) {
		if (n >= max_simples) break;
		const int want = top[lv];
		if (want < 1 || top[0] - want > cfg.max_vote_number_cutoff) break;
		unsigned done = 0;
		for (;;) {                                       // slots of this level in (row, j) order = the reference's i, j loops
			int best = -1, best_key = 0x7fffffff;
			for (int c = 0; c < C.n; c++)
				if (!((done >> c) & 1u) && C.votes[c] == want) {
					const int key = (int)C.row[c] * 32 + (int)C.j[c];
					if (key < best_key) { best_key = key; best = c; }
				}
			if (best < 0 || n >= max_simples) break;
			done |= 1u << best;
			if (want >= cfg.min_second) list[n++] = best;
		}
		for (int k = 0; k < mb; k++) {
			if (n >= max_simples) break;
			if (res[k].votes == want) list[n++] = 16 + k;
		}
	}
	SimpleResult tmp[kMaxBest];
	int cursor = 0;
	for (int s = 0; s < n && cursor < mb; s++) {
		const int code = list[s];
		const int votes = code < 16 ? (int)C.votes[code] : (int)res[code - 16].votes;
		if (votes < cfg.min_first) continue;
		const uint32_t pos = code < 16 ? C.kv[code] : res[code - 16].pos;
		bool exist = false;
		for (int e = 0; e < cursor; e++) exist = exist || tmp[e].pos == pos;
		if (exist) continue;
		if (code < 16) {
			SimpleResult &t = tmp[cursor];
			t.kind = 1; t.strand = (uint8_t)strand; t.first_sub = C.first_sub[code]; t.last_sub = C.last_sub[code];
			t.votes = C.votes[code]; t.used = (int16_t)applied; t.pos = pos;
			t.cov_start = C.first_off[code]; t.cov_end = (int16_t)(C.last_off[code] + 16);
		} else tmp[cursor] = res[code - 16];
		cursor++;
	}
	for (int k = 0; k < mb; k++) {                       // :2505-2527
		if (k < cursor) res[k] = tmp[k];
		else res[k].votes = 0;
	}
}

// k_map_simple: thread f maps single-end fragment f of a chunk whose bigtable slice is all zero.
// Writes the fragment's result slots, or queues the fragment for the warp kernel.
template <int GAP>
__global__ void __launch_bounds__(128)
k_map_simple(const DevIndex ix, const VoteConfig cfg, const uint8_t *text, uint64_t text_origin, uint64_t text_bytes,
             const uint64_t *off, const uint16_t *len, int64_t n_frags, ResultWords *out, uint32_t *queue, uint32_t *queue_count,
             uint32_t *bad_reads)
{
	const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	bool punt = false, active = f < n_frags;
	SimpleResult res[kMaxBest];
#pragma unroll
	for (int k = 0; k < kMaxBest; k++) { res[k].kind = 0; res[k].votes = 0; res[k].used = 0; res[k].pos = 0; }
	int applied = 0, L = 0;
	if (active) {
		const uint64_t o0 = off[f] - text_origin;
		L = len[f];
		if (o0 + (uint64_t)L > text_bytes || L > kMaxReadLen) {
			atomicAdd(bad_reads, 1u);
			active = false;
		} else if (L > kSimpleMaxLen) {
			punt = true;
		} else if (L >= 16) {
			uint32_t fw[kSimpleWords], rc[kSimpleWords];
			if (!simple_pack(text + o0, L, fw, rc)) punt = true;
			else {
				int step;
				applied = subread_schedule(L, GAP, cfg.total_subreads, &step);
				SimpleClusters C;
				if (!simple_pass<GAP>(ix, cfg, fw, L, applied, step, C)) punt = true;
				else {
					simple_merge(cfg, C, 0, applied, res);
					if (!simple_pass<GAP>(ix, cfg, rc, L, applied, step, C)) punt = true;
					else simple_merge(cfg, C, 1, applied, res);
				}
			}
		} else {
			SimpleClusters C;
			C.n = 0;
			simple_merge(cfg, C, 0, 0, res);   // a read shorter than 16 bases: two empty passes
			simple_merge(cfg, C, 1, 0, res);
		}
	}
	// queue the punted fragments (one atomic per warp)
	const unsigned pm = __ballot_sync(0xffffffffu, active && punt);
	if (pm) {
		uint32_t base = 0;
		if (lane == __ffs(pm) - 1) base = atomicAdd(queue_count, (uint32_t)__popc(pm));
		base = __shfl_sync(0xffffffffu, base, __ffs(pm) - 1);
		if (active && punt) queue[base + __popc(pm & ((1u << lane) - 1u))] = (uint32_t)f;
	}
	if (!active || punt) return;

	const int mb = cfg.multi_best_reads;
	ResultWords *dst = out + (size_t)f * mb;
	for (int k = 0; k < mb; k++) {   // copy_vote_to_alignment_res image of each slot (core-junction.c:1058-1072)
		uint32_t w[kResultWords];
#pragma unroll
		for (int i = 0; i < kResultWords; i++) w[i] = 0;
		const SimpleResult &r = res[k];
		if (r.kind == 1) {
			w[0] = r.pos;
			w[1] = r.strand ? (uint32_t)kCoreNegStrand : 0u;
			w[2] = (uint32_t)(uint16_t)r.votes | ((uint32_t)(uint16_t)r.used << 16);
			w[4] = (uint32_t)r.first_sub | ((uint32_t)r.last_sub << 16);
			w[15] = (uint32_t)(uint16_t)r.cov_start | ((uint32_t)(uint16_t)r.cov_end << 16);
		} else if (r.kind == 2) {
			w[2] = (uint32_t)(uint16_t)r.used << 16;
		}
#pragma unroll
		for (int i = 0; i < kResultWords; i++) dst[k].w[i] = w[i];
	}
}

#endif  // __CUDACC__

}  // namespace sr
