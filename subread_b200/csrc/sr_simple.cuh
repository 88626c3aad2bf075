// sr_simple.cuh -- tier 0 of the voting step: one THREAD per single-end fragment.
//
// The warp kernel (sr_warp.cuh) spends most of its instructions on warp-uniform work; for the
// common fragment that work is tiny.  A fragment is "simple" for a strand pass when every lookup
// returns at most one position, all returned positions agree on one kv = position - offset, and the
// hits come from distinct subreads.  Then gehash_go_X (sorted-hashtable.c:937-1123) builds exactly
// one vote slot -- created by the first hit, every later hit taking the plain branch of :1041-1052 --
// whose fields have a closed form:
//     votes = #hits, coverage = [offset(first hit), offset(last hit) + 16),
//     indel recorder = {subread(first)+1, subread(last)+1, 0},
// no shift-indel round (toli stays 0), or no slot at all when kv falls outside the block's border.
// With an all-zero bigtable slice (first index block) process_voting_junction_PE_topK
// (core-junction.c:2199-2530) then only ever merges at most one earlier slot (pass 0) with at
// most one new slot (pass 1), which is enumerated below exactly as the reference orders it.
//
// Everything else -- a lookup with several positions, disagreeing positions (indels, repeats),
// two phases of one subread hitting, characters that are not A/C/G/T, reads longer than
// kSimpleMaxLen -- is PUNTED untouched to the warp kernel through a queue, so the result of the
// tiered path is the reference's for every input.  Paired-end and subjunc chunks, chunks merged
// into a non-zero bigtable, and verification requests skip this tier altogether.
#pragma once
#include "sr_common.cuh"

namespace sr {

constexpr int kSimpleMaxLen = 256;                       // longer reads go to the warp kernel
constexpr int kSimpleWords = kSimpleMaxLen / 16 + 1;     // packed words per strand array

#if defined(__CUDACC__) && !defined(SR_HOST_ONLY)

struct SimpleSlot {     // the one vote slot a simple strand pass can produce
	uint32_t pos;
	int votes;          // 0 = no slot
	int cov_start, cov_end, first_sub_p1, last_sub_p1;
};

// mapping_result_t image of a simple slot (copy_vote_to_alignment_res core-junction.c:1058-1072).
__device__ __forceinline__ void simple_slot_words(const SimpleSlot &s, int strand, int used_subreads, uint32_t *w)
{
#pragma unroll
	for (int k = 0; k < kResultWords; k++) w[k] = 0;
	w[0] = s.pos;
	w[1] = strand ? (uint32_t)kCoreNegStrand : 0u;
	w[2] = (uint32_t)(uint16_t)s.votes | ((uint32_t)(uint16_t)used_subreads << 16);
	w[4] = (uint32_t)(uint16_t)s.first_sub_p1 | ((uint32_t)(uint16_t)s.last_sub_p1 << 16);
	w[15] = (uint32_t)(uint16_t)s.cov_start | ((uint32_t)(uint16_t)s.cov_end << 16);
}

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

// One strand pass.  Returns false when the pass is not simple.
template <int GAP>
__device__ __forceinline__ bool simple_pass(const DevIndex &ix, const VoteConfig &cfg, const uint32_t *packed, int L, int applied, int step,
                                            SimpleSlot *slot)
{
	const uint32_t nb = ix.buckets_number;
	const uint64_t magic = ix.buckets_magic;
	const int n_lk = applied * GAP;
	uint32_t kv0 = 0;
	int nh = 0, first_off = 0, last_off = 0, first_sub = 0, last_sub = 0;
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
			const int sub_p1 = l / GAP + 1;
			if (nh == 0) { kv0 = kv; first_off = o; first_sub = sub_p1; }
			else if (kv != kv0 || sub_p1 == last_sub) return false;
			nh++;
			last_off = o;
			last_sub = sub_p1;
		}
	}
	slot->votes = 0;
	if (nh > 0 && kv0 >= cfg.low_border && kv0 <= cfg.high_border - (uint32_t)L) {  // :1074, otherwise every hit is dropped
		slot->pos = kv0;
		slot->votes = nh;
		slot->cov_start = first_off;
		slot->cov_end = last_off + 16;
		slot->first_sub_p1 = first_sub;
		slot->last_sub_p1 = last_sub;
	}
	return true;
}

// k_map_simple: thread f maps single-end fragment f of a chunk whose bigtable slice is all zero.
// Writes the fragment's 3 result slots, or queues the fragment for the warp kernel.
template <int GAP>
__global__ void __launch_bounds__(128)
k_map_simple(const DevIndex ix, const VoteConfig cfg, const uint8_t *text, uint64_t text_origin, uint64_t text_bytes,
             const uint64_t *off, const uint16_t *len, int64_t n_frags, ResultWords *out, uint32_t *queue, uint32_t *queue_count,
             uint32_t *bad_reads)
{
	const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	bool punt = false, active = f < n_frags;
	uint32_t fw[kSimpleWords], rc[kSimpleWords];
	SimpleSlot s0, s1;
	s0.votes = s1.votes = 0;
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
			if (!simple_pack(text + o0, L, fw, rc)) punt = true;
			else {
				int step;
				applied = subread_schedule(L, GAP, cfg.total_subreads, &step);
				if (!simple_pass<GAP>(ix, cfg, fw, L, applied, step, &s0) || !simple_pass<GAP>(ix, cfg, rc, L, applied, step, &s1)) punt = true;
			}
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

	// do_voting's per-strand hand-over (core.c:3212-3226) and the top-K merge for <= 1 old + <= 1 new slot.
	// src: 0 = untouched zero slot, 1 = pass-0 slot, 2 = pass-1 slot, 3 = "unmapped" slot 0 that only
	// carries used_subreads_in_vote (core.c:3217-3225).
	const int mb = cfg.multi_best_reads;
	int src0 = 3, src1 = 0;
	const bool has0 = s0.votes >= cfg.min_first;      // pass 0: max_vote >= min_first -> topK with an empty bigtable
	if (has0) src0 = 1;
	if (s1.votes >= cfg.min_first) {                  // pass 1: merge the new slot with pass 0's result
		// vote levels (update_top_three): distinct values of {new, old} descending; a level more than
		// max_vote_number_cutoff below the best ends the list.  Inside a level the vote-table slot comes
		// before bigtable entries (:2262-2314); emission keeps list order and drops a repeated position.
		bool take_new = true, take_old = has0, new_first = true;
		if (has0) {
			if (s1.votes >= s0.votes) take_old = s1.votes - s0.votes <= cfg.max_vote_number_cutoff;
			else { new_first = false; take_new = s0.votes - s1.votes <= cfg.max_vote_number_cutoff; }
			if (take_new && take_old && s0.pos == s1.pos) { if (new_first) take_old = false; else take_new = false; }
		}
		if (take_new && take_old) {
			if (mb >= 2) { src0 = new_first ? 2 : 1; src1 = new_first ? 1 : 2; }
			else src0 = new_first ? 2 : 1;
		} else if (take_new) src0 = 2;
		// else: the pass-0 slot alone stays (the new one is cut off or repeats its position)
	}
	ResultWords *dst = out + (size_t)f * mb;
	uint32_t w[kResultWords];
	for (int k = 0; k < mb; k++) {
		const int src = k == 0 ? src0 : (k == 1 ? src1 : 0);
		if (src == 1) simple_slot_words(s0, 0, applied, w);
		else if (src == 2) simple_slot_words(s1, 1, applied, w);
		else {
#pragma unroll
			for (int i = 0; i < kResultWords; i++) w[i] = 0;
			if (src == 3) w[2] = (uint32_t)(uint16_t)applied << 16;
		}
#pragma unroll
		for (int i = 0; i < kResultWords; i++) dst[k].w[i] = w[i];
	}
}

#endif  // __CUDACC__

}  // namespace sr
