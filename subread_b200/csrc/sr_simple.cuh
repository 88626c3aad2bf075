// sr_simple.cuh -- tier 0 of the voting step: one THREAD per single-end fragment, one block per
// 128 consecutive fragments.
//
// The warp kernel (sr_warp.cuh) spends most of its instructions on warp-uniform work; for the
// common fragment that work is tiny.  A strand pass is "separable" when
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
// cluster on a large index -- are just extra one-vote slots, and a lookup that returns several
// positions contributes them in the reference's visit order (:1109-1119).
// process_voting_junction_PE_topK (core-junction.c:2199-2530) is then replayed on that small slot list
// and on the results of the first strand pass (the chunk's bigtable slice must be all zero, i.e. the
// first index block).
//
// Everything else -- kv values within tolerance of each other (indels), two phases of one subread
// hitting one kv, more than kSimpleClusters distinct kv, characters that are not A/C/G/T, reads longer
// than kSimpleMaxLen -- is PUNTED untouched to the warp kernel through a queue, so the tiered path
// returns the reference's result for every input.  Paired-end fragments have their own form of this tier
// (sr_simple_pe.cuh), subjunc chunks add the junction half (sr_simple_junc.cuh); chunks merged into a non-zero
// bigtable (later index blocks) and verification requests skip the tier altogether.
//
// Data movement is organised per block so that the per-thread work stays in registers and shared memory:
//   1. the block's read text (contiguous in the chunk) is staged into shared memory with coalesced
//      16-byte loads; every thread packs its own read from there (2 bits per base, both strands) into a
//      word-interleaved shared array, so later key extraction never touches local memory;
//   2. per strand pass, all Bloom words of a batch of 32 lookups are requested before any is tested
//      (memory-level parallelism, no branch between the loads) and the bucket records of the survivors
//      are prefetched into L2;
//   3. only the survivors (the true hits plus ~12 % false positives) run the divergent vote code;
//   4. the block's slice of the bigtable is zero-filled with coalesced stores and the handful of
//      non-zero words per fragment is scattered on top.
#pragma once
#include "sr_common.cuh"

#ifndef SR_SIMPLE_PREFETCH
#define SR_SIMPLE_PREFETCH 0   // prefetch.global.L2 of the survivors' records from the filter loop (A/B switch)
#endif
#ifndef SR_SIMPLE_MIN_BLOCKS
#define SR_SIMPLE_MIN_BLOCKS 6  // resident blocks per SM the register allocation aims at: 84 registers, no spills (7 and 8 measured slower)
#endif
#ifndef SR_SIMPLE_HIT_INLINE
#define SR_SIMPLE_HIT_INLINE __forceinline__   // the general-path hit: inlined (a real call costs spills around it; A/B switch)
#endif
#ifndef SR_SIMPLE_REORDER
#define SR_SIMPLE_REORDER 0    // sparse table: every lane takes its fuller strand first.  A/B switch, off: 17 % fewer warp instructions but
                               // 1.39 instead of 1.55 G reads/s -- the kernel runs at the memory system's request rate, profiles/r02_tier0_experiments.md
#endif
#ifndef SR_SIMPLE_DEFER_LIMIT
#define SR_SIMPLE_DEFER_LIMIT 3    // lookups of one batch a lane may take through the general path before the fragment is queued for the warp kernel
                                   // (A/B on B200: mixed workload 419 -> 446 M reads/s, clean 1 502 -> 1 533 M)
#endif
#ifndef SR_DENSE_PIPE
#define SR_DENSE_PIPE 1        // dense phase units (MODE 4): a unit is applied one iteration after it was read (its value loads overlap the unit before).  A/B switch
#endif
#ifndef SR_DENSE_PROBE3
#define SR_DENSE_PROBE3 1      // ... the first three clusters tried before the walk over the list.  A/B switch
#endif
#ifndef SR_SIMPLE_SMEM_OUT
#define SR_SIMPLE_SMEM_OUT 0   // compose the warp's bigtable slice in shared memory and store whole sectors (A/B switch)
#endif

namespace sr {

constexpr int kSimpleThreads = 128;                      // fragments (threads) per block
constexpr int kSimpleMaxLen = 256;                       // longer reads go to the warp kernel
constexpr int kSimpleClusters = 9;                       // distinct kv of both strand passes together (sparse table: the two lists share the pool)
constexpr int kSimpleClustersDense = 16;                 // ... on a dense table, where a quarter of all 16-mers has some position
constexpr int kSimpleClusterWords = 4;                   // shared-memory words per cluster

// Shared-memory plan of one block (host and device agree through these functions): the packed reads of the 128
// threads, then one private area per warp that first stages the warp's text and later holds its vote clusters.
SR_HD int simple_words(int max_len) { return ((max_len < kSimpleMaxLen ? max_len : kSimpleMaxLen) + 15) / 16 + 1; }
SR_HD int simple_cluster_bytes(int clusters, int words = kSimpleClusterWords) { return clusters * words * 32 * 4; }
SR_HD int simple_warp_bytes(int max_len, int clusters = kSimpleClusters)
{
	const int ml = max_len < kSimpleMaxLen ? max_len : kSimpleMaxLen;
	int stage = 32 * (((ml + 3) & ~3) + 4) + 32;                             // text of 32 reads (+ alignment slack)
	const int clus = simple_cluster_bytes(clusters) + 3 * 32 * 4;            // the vote clusters reuse the staging area (+ three words per thread: indel sections, simple_primary_hit)
	if (stage < clus) stage = clus;
	const int outb = 16 * kMaxBest * kResultWords * 4;                       // ... and later half a warp's bigtable slots
	if (stage < outb) stage = outb;
	return (stage + 15) & ~15;
}
SR_HD int simple_smem_bytes(int max_len, int clusters = kSimpleClusters)
{
	return 2 * simple_words(max_len) * kSimpleThreads * 4 + (kSimpleThreads / 32) * simple_warp_bytes(max_len, clusters);
}

#if defined(__CUDACC__) && !defined(SR_HOST_ONLY)

// One bigtable slot in compact form (registers).
struct SimpleResult {
	uint32_t pos;
	int votes;          // selected_votes as stored (0 for a record that lost its place: its other bytes stay)
	int used;
	uint32_t meta;      // kind | strand << 8 | first_sub << 16 | last_sub << 24; kind 0 untouched, 1 record, 2 only `used` is set
	uint32_t cov;       // confident_coverage_start | end << 16
	                    // meta bit 9: the slot's indel record has a second section, kept per strand pass by the kernel (`ind` below: first_sub + 1 |
	                    // (last_sub + 1) << 8 | (indel & 0xff) << 16 | 1 << 31); meta's last_sub is then the last subread of the FIRST section
};

// One vote cluster = the slot gehash_go_X builds for one kv of a separable pass.  Nothing in it depends on the
// order in which its hits are applied:
//   w0  kv
//   w1  first_l[0:8] last_l[8:16] first_u[16:23] row[24:29] has_slot[31]
//       first_l / last_l: lookup index (subread * gap + phase) of the cluster's first / last hit in reference order --
//       offset and subread number follow from it; first_u: visit index of the first hit inside its lookup's run
//       (only ever compared between clusters created by the same lookup)
//   w2, w3  bit s set = subread s voted; votes = popcount
// The reference appends a new slot at the end of its row (row = (kv/5)%30), so the rank j of a slot inside its row
// is its rank by creation time = by (first_l, first_u); candidates are visited in (row, j) order, i.e. in the order of
// the key row << 15 | first_l << 7 | first_u.  has_slot == 0: kv lies outside the border (:1074), its hits are dropped.
struct SimpleCluster {
	uint32_t kv, w1;
	unsigned long long subs;
	__device__ __forceinline__ int votes() const { return (w1 >> 31) ? __popcll(subs) : 0; }
	__device__ __forceinline__ int first_l() const { return (int)(w1 & 255u); }
	__device__ __forceinline__ int last_l() const { return (int)((w1 >> 8) & 255u); }
	__device__ __forceinline__ int order() const { return (int)(((w1 >> 24) & 31u) << 15 | (w1 & 255u) << 7 | ((w1 >> 16) & 127u)); }
};

// The clusters of the running pass in shared memory, CW words each, interleaved over the warp's lanes.  CW = 4 everywhere but in
// the paired-end tier on dense tables with reads of at most kLongRead bases (at most 32 subreads, so one word of subread bits
// does): there shared memory per thread bounds the resident warps, and three words per cluster make room for a fourth block.
template <int CW>
struct SimpleClustersT {
	static constexpr int kWords = CW;
	uint32_t *base;     // &clusters[0][0][lane] in the warp's private area
	int n, cap;
	__device__ __forceinline__ uint32_t &w(int c, int i) const { return base[(c * CW + i) * 32]; }
	__device__ __forceinline__ int votes(int c) const { return (w(c, 1) >> 31) ? __popc(w(c, 2)) + (CW > 3 ? __popc(w(c, CW - 1)) : 0) : 0; }
	__device__ __forceinline__ int order(int c) const
	{
		const uint32_t w1 = w(c, 1);
		return (int)(((w1 >> 24) & 31u) << 15 | (w1 & 255u) << 7 | ((w1 >> 16) & 127u));
	}
	__device__ __forceinline__ SimpleCluster get(int c) const
	{
		SimpleCluster r;
		r.kv = w(c, 0); r.w1 = w(c, 1);
		r.subs = (unsigned long long)w(c, 2);
		if (CW > 3) r.subs |= (unsigned long long)w(c, CW - 1) << 32;
		return r;
	}
	__device__ __forceinline__ void put(int c, const SimpleCluster &r) const
	{
		w(c, 0) = r.kv; w(c, 1) = r.w1; w(c, 2) = (uint32_t)r.subs;
		if (CW > 3) w(c, CW - 1) = (uint32_t)(r.subs >> 32);
	}
};
using SimpleClusters = SimpleClustersT<kSimpleClusterWords>;

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Cache policy by role.  The Bloom words (32 MB, hit at random 60 times per read) are the only data worth keeping:
// they are loaded evict-last; the bucket records (one random sector out of 0.5 GB, never reused) and the streamed
// text / bigtable are evict-first, so they do not push the filter out of L2.
__device__ __forceinline__ uint64_t policy_evict_last()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ __forceinline__ uint32_t ld_bloom(const bloom_t *p, uint64_t pol)
{
#if SR_BLOOM_BITS == 8
	uint32_t v;
	asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
	return v & 0xffu;
#else
	uint16_t v;
	asm volatile("ld.global.nc.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
	return v;
#endif
}

// A fresh cluster for the hit (kv, lookup l, visit index u) of subread `sub`.
__device__ __forceinline__ SimpleCluster simple_new_cluster(uint32_t kv, int l, int u, int sub, uint32_t low, uint32_t high)
{
	SimpleCluster c;
	c.kv = kv;
	c.w1 = 0;
	c.subs = 0;
	if (kv >= low && kv <= high) {                                 // :1074
		c.w1 = (uint32_t)l | ((uint32_t)l << 8) | ((uint32_t)u << 16) | (vote_row(kv) << 24) | 0x80000000u;
		c.subs = 1ull << sub;
	}
	return c;
}

// One hit on the general path (clusters in shared memory).  Returns false when the pass is not separable.
// (Rare path: a real call keeps the hot loop small.)
template <class CL>
__device__ SR_SIMPLE_HIT_INLINE bool simple_hit(CL &C, uint32_t kv, int l, int u, int sub, int tol, uint32_t low, uint32_t high, int *found_out = nullptr)
{
	if (!(kv >= 32u && kv <= 0xffffffdfu)) return false;           // u32 wrap-around of the probe rows
	int found = -1;
	bool near = false;
#pragma unroll 1
	for (int c = 0; c < C.n; c++) {                                // (no branch inside: the warp stays together)
		const int d = (int)(kv - C.w(c, 0));
		found = d == 0 ? c : found;
		near = near || (d != 0 && (unsigned)(d + tol) <= 2u * (unsigned)tol);
	}
	if (near) return false;                                        // within tolerance of another slot: an indel vote
	if (found_out) *found_out = found;
	if (found >= 0) {
		SimpleCluster c = C.get(found);
		if (!(c.w1 >> 31)) return true;                            // no slot for this kv: dropped again
		const unsigned long long bit = 1ull << sub;
		if (c.subs & bit) return false;                            // one subread twice: the reference opens a duplicate slot
		c.subs |= bit;
		if (l < c.first_l()) c.w1 = (c.w1 & ~0x007f00ffu) | (uint32_t)l | ((uint32_t)u << 16);
		if (l > c.last_l()) c.w1 = (c.w1 & ~0x0000ff00u) | ((uint32_t)l << 8);
		C.put(found, c);
		return true;
	}
	if (C.n >= C.cap) return false;
	C.put(C.n++, simple_new_cluster(kv, l, u, sub, low, high));
	return true;
}

// Pack the thread's read from the staging area: forward codes (MSB first, 16 bases per word, code 2 beyond
// the read like the reference's NUL) into pk[k * 128], reverse-complement codes into pk[(n_words + k) * 128].
// Returns false when a character is not one of A, C, G, T (then base2int / reverse_read quirks apply and the
// fragment is punted) -- except, with ALLOW_N, the letter N, the one odd character real reads are full of: genekey2int
// reads it as code 3 (base2int, subread.h:238: anything from 'H' up that is not 'G'), and reverse_read (input-files.c:1171-1186)
// leaves it an 'N', so it is code 3 on the reverse strand as well, where a T would have become an A.  A read with Ns gets its
// reverse-strand words character by character (the rare path); everything downstream only sees codes.
template <bool WITH_RC = true, bool ALLOW_N = false>
__device__ __forceinline__ bool simple_pack(const uint8_t *stage, unsigned so, int L, uint32_t *pk, int n_words)
{
	const int nw = (L + 15) / 16 + 1;
	const uint32_t *sw = reinterpret_cast<const uint32_t *>(stage + (so & ~3u));
	const unsigned mis = (so & 3u) * 8u;
	const int last_word = (int)(((so & 3u) + (unsigned)L + 3u) / 4u);    // words [0, last_word) hold characters
	uint32_t bad = 0, any_n = 0;
	uint32_t wprev = sw[0];
	for (int k = 0; k < nw; k++) {
		uint32_t f = 0;
		const int nvalid = min(16, max(0, L - 16 * k));
		if (nvalid == 0) { pk[k * kSimpleThreads] = 0xAAAAAAAAu; continue; }
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const int j = 4 * k + q + 1;
			const uint32_t wnext = j < last_word ? sw[j] : 0u;
			const uint32_t c4 = __funnelshift_r(wprev, wnext, mis);
			wprev = wnext;
			// A 0x41, C 0x43, G 0x47, T 0x54: code bit0 = bit 2, code bit1 = bit 1 ^ bit 2  (A0 G1 C2 T3)
			const uint32_t b2 = (c4 >> 2) & 0x01010101u, b1 = ((c4 >> 1) ^ (c4 >> 2)) & 0x01010101u;
			uint32_t code4 = b2 | (b1 << 1);
			// validate by rebuilding the characters from the codes: "AGCT"[code]
			const uint32_t sel = (code4 & 3u) | ((code4 >> 4) & 0x30u) | ((code4 >> 8) & 0x300u) | ((code4 >> 12) & 0x3000u);
			const uint32_t rebuilt = __byte_perm(0x54434741u, 0u, sel);
			const int nb = nvalid - 4 * q;  // characters of this group inside the read
			uint32_t diff = rebuilt ^ c4;
			if (nb < 4) diff = nb <= 0 ? 0u : (diff & (0xffffffffu >> (8 * (4 - nb))));
			if (ALLOW_N && diff) {                                       // (rare) forgive the Ns: code 3
				uint32_t isn = __vcmpeq4(c4, 0x4E4E4E4Eu);
				if (nb < 4) isn = nb <= 0 ? 0u : (isn & (0xffffffffu >> (8 * (4 - nb))));
				code4 |= isn & 0x03030303u;
				diff &= ~isn;
				any_n |= isn;
			}
			bad |= diff;
			f |= ((code4 * 0x40100401u) >> 24) << (24 - 8 * q);
		}
		const uint32_t keep = nvalid >= 16 ? 0xffffffffu : ~(0xffffffffu >> (2 * nvalid));
		pk[k * kSimpleThreads] = (f & keep) | (0xAAAAAAAAu & ~keep);
	}
	if (!WITH_RC) return bad == 0;                                 // the caller derives reverse-strand keys from this array
	uint32_t *rc = pk + n_words * kSimpleThreads;
	if (ALLOW_N && any_n && bad == 0) {
		// reverse complement with Ns: character by character, rc base i = reverse_read(char L-1-i)
		const uint8_t *ch = stage + so;
		for (int k = 0; k < nw; k++) {
			const int nvalid = min(16, max(0, L - 16 * k));
			uint32_t r = 0xAAAAAAAAu;
			if (nvalid > 0) {
				r = 0;
#pragma unroll 4
				for (int j = 0; j < 16; j++) {
					const int i = 16 * k + j;
					uint32_t code = 2u;                                  // beyond the read
					if (i < L) {
						const int c = (int)ch[L - 1 - i];
						code = c == 'N' ? 3u : (uint32_t)(3 - base_code(c));
					}
					r |= code << (30 - 2 * j);
				}
			}
			rc[k * kSimpleThreads] = r;
		}
		return true;
	}
	// reverse complement of an all-base read: rc base i = 3 - f base (L-1-i)
	for (int k = 0; k < nw; k++) {
		const int nvalid = min(16, max(0, L - 16 * k));
		uint32_t r = 0xAAAAAAAAu;
		if (nvalid > 0) {
			const int s = L - 16 - 16 * k;
			const int w = s >> 4, sh = (s & 15) * 2;
			const uint32_t hi = w >= 0 ? ~pk[w * kSimpleThreads] : 0u, lo = ~pk[(w + 1) * kSimpleThreads];
			uint32_t g = __funnelshift_l(lo, hi, sh);
			g = __brev(g);
			g = ((g & 0x55555555u) << 1) | ((g >> 1) & 0x55555555u);
			const uint32_t keep = nvalid >= 16 ? 0xffffffffu : ~(0xffffffffu >> (2 * nvalid));
			r = (g & keep) | (0xAAAAAAAAu & ~keep);
		}
		rc[k * kSimpleThreads] = r;
	}
	return bad == 0;
}

// Offset of the first lookup of subread s (phase 0); lookup s * GAP + x sits at this offset + x (core.c:3169-3171).
template <int GAP>
__device__ __forceinline__ int simple_sub_offset(int s, int step)
{
	unsigned o = ((unsigned)step * (unsigned)s) >> 16;        // step * s < 2^31: s <= 62, step < 2^24
	if (GAP > 1) o -= o % (unsigned)GAP;
	return (int)o;
}
template <int GAP>
__device__ __forceinline__ int simple_offset(int l, int step, int *sub)
{
	const unsigned s = (unsigned)l / (unsigned)GAP, x = (unsigned)l - s * (unsigned)GAP;
	*sub = (int)s;
	return simple_sub_offset<GAP>((int)s, step) + (int)x;
}

__device__ __forceinline__ uint32_t simple_key(const uint32_t *pk, int o)
{
	const int w = o >> 4, sh = (o & 15) * 2;
	return __funnelshift_l(pk[(w + 1) * kSimpleThreads], pk[w * kSimpleThreads], sh);
}
// The same for a pass that may run on the reverse strand of a read of which only the forward array is kept (rc_len > 0 =
// the read's length): the 16-mer at offset o of the reverse complement is the reverse complement of the forward 16-mer at
// offset L - 16 - o, followed by the reference's padding where it passes the end of the read.
template <bool RCKEYS>
__device__ __forceinline__ uint32_t simple_key_at(const uint32_t *pk, int o, int rc_len)
{
	if (!RCKEYS || rc_len == 0) return simple_key(pk, o);
	const int fo = rc_len - 16 - o;
	uint32_t g = __brev(~simple_key(pk, max(fo, 0)));
	g = ((g & 0x55555555u) << 1) | ((g >> 1) & 0x55555555u);
	// A read of 16 or 17 bases is still looked up at offsets 0..2 (core.c:3130: one subread): those keys run past its
	// end, where the reference's buffer holds NUL = code 2
	if (fo < 0) g = (g << (-2 * fo)) | (0xAAAAAAAAu >> (32 + 2 * fo));
	return g;
}

// What one survivor's bucket record says about key16 = k16.
//   kind 0: the key is absent   1: exactly one position (*value)   2: anything else -- several positions, or a bucket
//   whose record cannot answer -- which the general path sorts out.
__device__ __forceinline__ int simple_probe(const DevIndex &ix, const Record8 &r, int k16, uint32_t *value, uint32_t &n_gather)
{
	const int n = (int)(r.w[0] & 0xffffu);
	// The key16 fields of a record are sorted and the unused ones hold kRecordNoKey (0x7fff, above every key: records are only
	// built for tables whose key16 stays below it), so the equal-key run [lt, le) follows from two packed compares per
	// word: with a, b < 0x8000 bit 15 of (a | 0x8000) - b is clear exactly when a < b, and no borrow crosses the halves.
	const uint32_t M = 0x80008000u, kk = (uint32_t)k16 * 0x10001u, kk1 = kk + 0x10001u;
	if (n <= kInlineItems) {
		const uint32_t a0 = (r.w[0] | 0x0000ffffu) & 0xffff7fffu, a1 = r.w[1], a2 = r.w[2];   // (the count's half becomes a no-key)
		const int lt = __popc(~((a0 | M) - kk) & M) + __popc(~((a1 | M) - kk) & M) + __popc(~((a2 | M) - kk) & M);
		const int le = __popc(~((a0 | M) - kk1) & M) + __popc(~((a1 | M) - kk1) & M) + __popc(~((a2 | M) - kk1) & M);
		*value = lt == 0 ? r.w[3] : (lt == 1 ? r.w[4] : (lt == 2 ? r.w[5] : (lt == 3 ? r.w[6] : r.w[7])));
		return min(le - lt, 2);
	}
	if (n <= kRecordKeys) {                                        // 6..12 items: the record lists the keys, the positions stay in vals[]
		int lt = 0, le = 0;
#pragma unroll
		for (int w = 2; w < 8; w++) {
			lt += __popc(~((r.w[w] | M) - kk) & M);
			le += __popc(~((r.w[w] | M) - kk1) & M);
		}
		if (le - lt == 1) { *value = ld_gather(ix.vals + r.w[1] + lt); n_gather++; }
		return min(le - lt, 2);
	}
	return 2;
}

// The general path of a sparse table's pass: the lookups of `defer` (bits = lookups base + i), in schedule order, every
// position in the reference's visit order, through the cluster list.  Returns false when the pass is not separable.
template <int GAP, bool RCKEYS>
__device__ __forceinline__ bool simple_general(const DevIndex &ix, const VoteConfig &cfg, const uint32_t *pk, int step, int rc_len, int base,
                                               unsigned defer, SimpleClusters &C, uint32_t low, uint32_t high, uint32_t &n_gather)
{
	const uint32_t nb = ix.buckets_number;
	const uint64_t magic = ix.buckets_magic;
	const int tol = cfg.indel_tolerance;
	const uint4 *records = reinterpret_cast<const uint4 *>(ix.records);
	while (defer) {
		const int i = __ffs(defer) - 1;
		defer &= defer - 1;
		const int l = base + i;
		int sub;
		const int o = simple_offset<GAP>(l, step, &sub);
		const uint32_t key = simple_key_at<RCKEYS>(pk, o, rc_len);
		const uint32_t q = (uint32_t)__umul64hi((uint64_t)key, magic);
		const uint32_t b = key - q * nb;
		const int k16 = (int)(int16_t)q;
		bool in_arrays = true;
		if (records) {
			const Record8 r = ld_record(records, b);
			n_gather++;
			const int n = (int)(r.w[0] & 0xffffu);
			if (n <= kInlineItems) {
				in_arrays = false;
				const int k0 = (int)(int16_t)(r.w[0] >> 16), k1 = (int)(int16_t)(r.w[1] & 0xffffu), k2 = (int)(int16_t)(r.w[1] >> 16),
				          k3 = (int)(int16_t)(r.w[2] & 0xffffu), k4 = (int)(int16_t)(r.w[2] >> 16);
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
					const int h = le - lt, nfwd = le - at;       // landing .. end of run, then landing-1 .. start (:1109-1119)
					for (int u = 0; u < h; u++) {
						const int idx = u < nfwd ? at + u : at - 1 - (u - nfwd);
						const uint32_t v = idx == 0 ? r.w[3] : (idx == 1 ? r.w[4] : (idx == 2 ? r.w[5] : (idx == 3 ? r.w[6] : r.w[7])));
						if (!simple_hit(C, v - (uint32_t)o, l, u, sub, tol, low, high)) return false;
					}
				}
			}
		}
		if (in_arrays) {                                         // a bucket with more than 5 items, or no records at all
			const uint32_t lo = ld_gather(ix.bucket_start + b);
			const int items = (int)(ld_gather(ix.bucket_start + b + 1) - lo);
			n_gather++;
			if (items <= 0) continue;
			const int16_t *kp = ix.keys + lo;
			int imin = 0, imax = items - 1, at;
			bool found = false;
			n_gather++;
			for (;;) {                                           // sorted-hashtable.c:960-981
				at = (imin + imax) >> 1;
				const int kk = (int)__ldg(kp + at);
				if (kk > k16) imax = at - 1;
				else if (kk < k16) imin = at + 1;
				else { found = true; break; }
				if (imax < imin) break;
			}
			if (!found) continue;
			int e = at + 1, st = at;
			while (e < items && (int)__ldg(kp + e) == k16) e++;
			while (st > 0 && (int)__ldg(kp + st - 1) == k16) st--;
			const int h = e - st, nfwd = e - at;
			if (h > C.cap) return false;                         // cannot stay separable
			for (int u = 0; u < h; u++) {
				const int idx = u < nfwd ? at + u : at - 1 - (u - nfwd);
				const uint32_t v = ld_gather(ix.vals + lo + idx);
				n_gather++;
				if (!simple_hit(C, v - (uint32_t)o, l, u, sub, tol, low, high)) return false;
			}
		}
	}
	return true;
}

// One single-position hit against the primary cluster P (registers).  Returns true when the hit is not P's business and has to
// go the general way.  With `p2` (single-end tier) a read with ONE indel stays here too: its hits fall on two offsets, kv and
// kv + d (|d| within the tolerance), those on kv first.  gehash_go_X puts them all into the slot the first hit opened
// (sorted-hashtable.c:1007-1072): a hit at distance d from the slot's position with d != current_indel_cursor opens a new
// section of the indel record {subread + 1, subread + 1, d}, later hits at that distance extend it.  In closed form, for hits in
// increasing subread order: votes = number of subreads, record = {first1 + 1, last1 + 1, 0, first2 + 1, last2 + 1, d}, coverage
// from the first hit to the last.  Anything else -- a hit back on the first offset (the reference then marks a shift indel and
// votes a second round, :1022, core.c:3153-3186), a third offset, a subread voting twice (roll-back, :1027), a slot that was
// never created -- makes the pass non-separable (ok = false) and the fragment is queued.
// The second section's state is one word of shared memory per thread (`sec2`, null = not in this tier), touched only by the
// reads that have one, so that the common path carries no register for it:
//   first lookup of section 2 | last lookup of section 2 << 8 | (d & 0xff) << 16 | 1 << 31;  0 = no second section
// P's own last lookup stays that of the FIRST section; simple_indel_close checks after the pass that the two did not interleave.
__device__ __forceinline__ bool simple_primary_hit(SimpleCluster &P, bool &have_p, SimpleClusters &C, uint32_t *sec2, bool &ok, uint32_t kv, int l,
                                                   int sub, int tol, uint32_t low, uint32_t high)
{
	if (!have_p && C.n == 0) {                                       // the first position of the pass founds P
		if (!(kv >= 32u && kv <= 0xffffffdfu)) ok = false;
		P = simple_new_cluster(kv, l, 0, sub, low, high);
		have_p = true;
		C.n = 1;
		return false;
	}
	if (!have_p) return true;
	const int d = (int)(kv - P.kv);
	const unsigned long long bit = 1ull << sub;
	if (d == 0) {                                                    // the plain branch :1041-1052
		if (P.w1 >> 31) {
			if (P.subs & bit) ok = false;
			P.subs |= bit;
			P.w1 = (P.w1 & ~0x0000ff00u) | ((uint32_t)l << 8);       // lookups come in increasing order here
		}
		return false;
	}
	if (sec2 && d >= -tol && d <= tol) {
		const uint32_t s2 = *sec2;
		if (!(P.w1 >> 31) || (P.subs & bit) || (s2 && (int)(int8_t)(s2 >> 16) != d)) { ok = false; return false; }
		*sec2 = s2 ? ((s2 & ~0x0000ff00u) | ((uint32_t)l << 8)) : (0x80000000u | (uint32_t)l | ((uint32_t)l << 8) | ((uint32_t)(d & 0xff) << 16));
		P.subs |= bit;
		return false;
	}
	return true;
}

// After the pass, for a primary cluster with a second section (sec2 != 0): the closed form is the reference's only if every hit of
// the first section came before the second began (else :1022 marks a shift indel) and no other cluster could have taken the
// section's hits.  Moves P's last lookup on to the section's (coverage end) and returns the merge's word: last subread of section
// 1 + 1 | (first subread of section 2 + 1) << 8 | (d & 0xff) << 16 | 1 << 31 -- or 0 when the pass is not separable after all.
__device__ __forceinline__ uint32_t simple_indel_close(SimpleClusters &C, uint32_t s2, int tol, int gap)
{
	const uint32_t kv = C.w(0, 0), w1 = C.w(0, 1);
	const uint32_t last1 = (w1 >> 8) & 255u, first2 = s2 & 255u, last2 = (s2 >> 8) & 255u;
	if (first2 <= last1) return 0u;
#pragma unroll 1
	for (int c = 1; c < C.n; c++) {
		const int d = (int)(C.w(c, 0) - kv);
		if (d >= -2 * tol && d <= 2 * tol) return 0u;
	}
	C.w(0, 1) = (w1 & ~0x0000ff00u) | (last2 << 8);
	return (last1 / (uint32_t)gap + 1u) | ((first2 / (uint32_t)gap + 1u) << 8) | (s2 & 0x80ff0000u);
}

// The filter of one batch of a strand pass: lookups base .. base + n_sub * GAP - 1 (whole subreads s_base .. s_base + n_sub - 1)
// of the read packed at pk.  Returns the survivors as a bit mask (bit i = lookup base + i).  All Bloom words (sparse table) or
// direct-filter words (dense table) of a group of six lookups are requested before any is tested -- memory-level
// parallelism, no branch between the loads.  Without a filter every lookup survives.
template <int GAP, int MODE, bool RCKEYS>
__device__ __forceinline__ unsigned simple_filter(const DevIndex &ix, const uint32_t *pk, int s_base, int n_sub, int step, int rc_len,
                                                  uint32_t &n_gather)
{
	constexpr bool DIRECT = MODE == 2;
	constexpr int kGroup = GAP == 3 ? 2 : (GAP == 2 ? 3 : 6);     // subreads per group of 6 lookups in flight
	const int cnt = n_sub * GAP;
	unsigned mask = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
	if (DIRECT) {
		// dense table: the key's two bits in the direct filter, one 64-byte fetch per lookup
		mask = 0;
#pragma unroll 1
		for (int g = 0; g < n_sub; g += kGroup) {
			uint32_t sh[kGroup * GAP];
			const uint32_t *wp[kGroup * GAP];
#pragma unroll
			for (int ss = 0; ss < kGroup; ss++) {
				const int ob = simple_sub_offset<GAP>(s_base + min(g + ss, n_sub - 1), step);
#pragma unroll
				for (int x = 0; x < GAP; x++) {
					const uint32_t key = simple_key_at<RCKEYS>(pk, ob + x, rc_len);
					const uint32_t u = direct_unit(key), e = key - u * kDirectKeys;
					wp[ss * GAP + x] = ix.direct + 4 * (size_t)u + 1 + (e >> 4);
					sh[ss * GAP + x] = 2 * (e & 15u);
				}
			}
			uint32_t word[kGroup * GAP];
#pragma unroll
			for (int u = 0; u < kGroup * GAP; u++) word[u] = ld_gather(wp[u]);
#pragma unroll
			for (int u = 0; u < kGroup * GAP; u++) {
				const bool in = g + u / GAP < n_sub && ((word[u] >> sh[u]) & 1u);
				mask |= (in ? 1u : 0u) << ((g * GAP + u) & 31);
			}
		}
		n_gather += (uint32_t)cnt;
	} else if (ix.bloom) {
		const uint32_t nb = ix.buckets_number;
		const uint64_t magic = ix.buckets_magic;
		const uint64_t pol = policy_evict_last();
		mask = 0;
#pragma unroll 1
		for (int g = 0; g < n_sub; g += kGroup) {
			uint32_t bw[kGroup * GAP], bq[kGroup * GAP];
#pragma unroll
			for (int ss = 0; ss < kGroup; ss++) {
				const int ob = simple_sub_offset<GAP>(s_base + min(g + ss, n_sub - 1), step);
#pragma unroll
				for (int x = 0; x < GAP; x++) {
					const uint32_t key = simple_key_at<RCKEYS>(pk, ob + x, rc_len);
					const uint32_t q = (uint32_t)__umul64hi((uint64_t)key, magic);
					bq[ss * GAP + x] = q;
					bw[ss * GAP + x] = key - q * nb;
				}
			}
			uint32_t word[kGroup * GAP];
#pragma unroll
			for (int u = 0; u < kGroup * GAP; u++) word[u] = ld_bloom(ix.bloom + bw[u], pol);
#pragma unroll
			for (int u = 0; u < kGroup * GAP; u++) {
				const uint32_t m = bloom_bits((int)bq[u]);
				const bool in = g + u / GAP < n_sub && (word[u] & m) == m;
				mask |= (in ? 1u : 0u) << ((g * GAP + u) & 31);
			}
		}
	}
	return mask;
}

// One strand pass.  pk = the pass's packed words of this thread (stride 128).  Leaves the clusters in C (shared
// memory).  Returns false when the pass is not separable.
//
// Per batch of up to 32 lookups: (1) filter (simple_filter; the caller may hand in the first batch's mask, `first_mask`,
// when it has run that filter already to decide which strand to take first); (2) a uniform loop over the survivors reads each
// bucket record (the next one is in flight while the current one is applied) and takes the hit if it is a single position
// that belongs to the "primary" cluster P, which lives in registers (the first such hit creates P); (3) whatever else turned
// up -- a position away from P, a lookup with several positions, a bucket of more than twelve items -- is replayed on the
// general path, lookup by lookup in schedule order and position by position in the reference's visit order.  Applying P's
// hits before the others is exact because a cluster does not depend on the order of application.
// MODE: 0 sparse table (Bloom word + bucket record), 2 dense table (direct filter)
__device__ __noinline__ bool simple_hit_call(SimpleClusters &C, uint32_t kv, int l, int u, int sub, int tol, uint32_t low, uint32_t high)
{
	return simple_hit(C, kv, l, u, sub, tol, low, high);
}

__device__ __forceinline__ uint32_t ld_phase_any(const uint32_t *p, uint64_t pol)
{
	uint32_t v;
	asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
	return v;
}

// One strand pass over the phase table of a gapped index (MODE 3; gap 3, sparse table; sr_common.cuh): the three lookups of a
// subread share one 32-byte unit.  Per batch of ten subreads: (1) the units' occupancy bits are requested for all ten (32 MB
// bitmap, L2-resident) -- a subread whose unit is empty has no hit in any phase; (2) a uniform loop over the subreads left
// fetches each unit (the next one in flight while this one is applied), reads the two bits of each of the three keys and takes
// the position of a key with a single one straight from the unit -- into the primary cluster P as in the other modes; (3) a
// key with several positions, or beyond the unit's five values, or not of P's kv, goes through the general path by its lookup
// number exactly as before.  The lookups are applied in schedule order (subread by subread, phase 0, 1, 2).
template <bool RCKEYS>
__device__ __forceinline__ bool simple_pass_phase(const DevIndex &ix, const VoteConfig &cfg, const uint32_t *pk, int L, int applied, int step,
                                                  SimpleClusters &C, int rc_len, uint32_t &n_gather, uint32_t *sec2)
{
	constexpr int GAP = 3, kSubs = 10;                             // ten subreads = thirty lookups per batch (the general path's bit mask)
	const uint32_t low = cfg.low_border, high = cfg.high_border - (uint32_t)L;
	const uint4 *units = reinterpret_cast<const uint4 *>(ix.phase_units);
	const uint64_t pol = policy_evict_last();
	C.n = 0;
	if (C.cap < 1) return false;
	SimpleCluster P;
	P.kv = 0; P.w1 = 0; P.subs = 0;
	bool have_p = false, ok = true;
	if (sec2) *sec2 = 0u;                                                       // one-indel reads in closed form (simple_primary_hit)
	for (int s_base = 0; s_base < applied; s_base += kSubs) {
		const int n_sub = min(kSubs, applied - s_base), base = s_base * GAP;
		// ---- (1) which units hold anything ----
		unsigned mask = 0;
#pragma unroll 1
		for (int g = 0; g < n_sub; g += 5) {
			uint32_t c[5], w[5];
#pragma unroll
			for (int u = 0; u < 5; u++) {
				const int ob = simple_sub_offset<GAP>(s_base + min(g + u, n_sub - 1), step);
				c[u] = simple_key_at<RCKEYS>(pk, ob, rc_len) & kPhaseMask;
			}
#pragma unroll
			for (int u = 0; u < 5; u++) w[u] = ld_phase_any(ix.phase_any + (c[u] >> 5), pol);
#pragma unroll
			for (int u = 0; u < 5; u++) mask |= ((g + u < n_sub && ((w[u] >> (c[u] & 31u)) & 1u)) ? 1u : 0u) << (g + u);
		}
		// ---- (2) the subreads left: one unit each ----
		unsigned defer = 0;
		Record8 rn;
		int si_next = -1;
		uint32_t k0_next = 0;
		auto issue = [&]() {
			si_next = -1;
			if (!mask) return;
			si_next = __ffs(mask) - 1;
			mask &= mask - 1;
			k0_next = simple_key_at<RCKEYS>(pk, simple_sub_offset<GAP>(s_base + si_next, step), rc_len);
			rn = ld_record(units, k0_next & kPhaseMask);
			n_gather++;
		};
		issue();
		while (si_next >= 0) {
			const Record8 r = rn;
			const int si = si_next, sub = s_base + si;
			const uint32_t k0 = k0_next;
			issue();
			const int ob = simple_sub_offset<GAP>(sub, step);            // (the other two keys only now: fewer registers across the fetch)
			const uint32_t k1 = simple_key_at<RCKEYS>(pk, ob + 1, rc_len), k2 = simple_key_at<RCKEYS>(pk, ob + 2, rc_len);
			const uint32_t c = k0 & kPhaseMask;
			if (((k1 >> 2) & kPhaseMask) != c || (k2 >> 4) != c) {   // (the three keys do not overlap as a read's do: padding of a tiny read's reverse strand)
				defer |= 7u << (3 * si);
				continue;
			}
			const uint32_t p0 = r.w[0] & 0x55555555u, p1 = r.w[1] & 0x55555555u;
#pragma unroll
			for (int x = 0; x < 3; x++) {
				const uint32_t key = x == 0 ? k0 : (x == 1 ? k1 : k2), v = phase_variant(key, x);
				const uint32_t wx = x == 0 ? r.w[0] : (x == 1 ? r.w[1] : r.w[2]);
				const uint32_t two = (wx >> (2 * v)) & 3u;
				if (!(two & 1u)) continue;                               // the table does not hold this key
				const int rank = __popc(wx & 0x55555555u & ((1u << (2 * v)) - 1u)) + (x > 0 ? __popc(p0) : 0) + (x > 1 ? __popc(p1) : 0);
				const int li = 3 * si + x, l = base + li;
				if ((two & 2u) || rank >= kPhaseValues) { defer |= 1u << li; continue; }
				const uint32_t value = rank == 0 ? r.w[3] : (rank == 1 ? r.w[4] : (rank == 2 ? r.w[5] : (rank == 3 ? r.w[6] : r.w[7])));
				const uint32_t kv = value - (uint32_t)(ob + x);
				if (simple_primary_hit(P, have_p, C, sec2, ok, kv, l, sub, cfg.indel_tolerance, low, high)) defer |= 1u << li;
			}
		}
		if (!ok) return false;
		// ---- (3) everything else, on the general path ----
		if (defer) {
			if (__popc(defer) > SR_SIMPLE_DEFER_LIMIT) return false;
			if (have_p) C.put(0, P);
			const unsigned long long before = P.subs;
			if (!simple_general<GAP, RCKEYS>(ix, cfg, pk, step, rc_len, base, defer, C, low, high, n_gather)) return false;
			if (have_p) P = C.get(0);
			if (sec2 && *sec2 && P.subs != before) return false;         // (a late hit on the first offset of a read with an indel)
		}
	}
	if (have_p) C.put(0, P);
	if (sec2 && *sec2) {
		const uint32_t m = simple_indel_close(C, *sec2, cfg.indel_tolerance, GAP);
		if (!m) return false;
		*sec2 = m;
	}
#pragma unroll 1
	for (int c = 0; c < C.n; c++) {
		C.w(c, 2) = (uint32_t)C.votes(c);
		C.w(c, 3) = 0u;
	}
	return true;
}

// The same over the 128-byte units of a dense table (MODE 4; sr_common.cuh).  No unit is empty, so every subread's unit is
// fetched: one uniform loop over the subreads of a batch, software-pipelined in three stages -- the next unit's first 32 bytes
// (the 2-bit codes and four values) in flight; this unit read (which keys it holds, the loads of the values beyond its first
// four under way); the unit before applied to the vote clusters.  Single positions go through the cluster list at once, the
// keys with several positions (their value is the rank of their direct_first entry) wait in registers: their entries are asked
// for as they turn up, read behind the loop, and their runs are walked in one flat loop, the next position's load under way.
// What round 2 measured on the human-sized paired-end workload (profiles/r02_human_pe_tier.md): 306 -> 481 M reads/s from the
// sieves and the convergent hit(), the two-step unit, the entries behind the loop, the pipelined runs and the compacted hits;
// prefetch.global.L2 of the units further down (3 or 6 ahead, or the next pass's first two) cost 7 %; 64 bytes of a unit in
// registers changed nothing.
template <bool RCKEYS, class CL>
__device__ __forceinline__ bool simple_pass_phase_dense(const DevIndex &ix, const VoteConfig &cfg, const uint32_t *pk, int L, int applied, int step,
                                                        CL &C_out, int rc_len, uint32_t &n_gather)
{
	constexpr int GAP = 3, kSubs = 10;
	CL C = C_out;         // (the caller hands over one of two lists picked at run time: a copy keeps base / n / cap in plain registers)
	const int tol = cfg.indel_tolerance;
	const uint32_t low = cfg.low_border, high = cfg.high_border - (uint32_t)L;
	const uint4 *units4 = reinterpret_cast<const uint4 *>(ix.phase_dense);
	C.n = 0;
	if (C.cap < 1) return false;
	bool ok = true;
	// A pass over a dense table meets a handful of one-vote chance hits besides the true cluster, and looking every hit up in the
	// cluster list was half of the kernel's instructions.  Three 32-bit sieves stand in front of the list.  occ_a / occ_b: one bit
	// per 16 consecutive kv, at two alignments 8 apart -- two kv within the tolerance (<= 8) of each other share a bin in one of
	// them; occ_f: one bit per 512.  A cluster near the hit shows in (a or b) AND in f; a hit that does not is neither equal to nor
	// within the tolerance of any cluster and opens a new one unseen.  The cluster the last hit was found equal to (the true
	// cluster, as a rule) is tried first -- two clusters of the list are never within tolerance of each other, so a hit that
	// equals one needs no look at the others.  These two common cases share ONE body that touches only the words it changes, so
	// the lanes of a warp that hold either run it together; the walk over the list is left to the hits that do have a neighbour.
	uint32_t occ_a = 0, occ_b = 0, occ_f = 0;
	uint32_t last_kv = 0;
	int last_c = -1;
	auto hit = [&](uint32_t kv, int l, int u, int sub) -> bool {
		bool same = last_c >= 0 && kv == last_kv;
		const uint32_t ba = 1u << ((kv >> 4) & 31u), bb = 1u << (((kv + 8u) >> 4) & 31u);
		const uint32_t bf = (1u << (((kv - 8u) >> 9) & 31u)) | (1u << (((kv + 8u) >> 9) & 31u));
		const bool clear = (tol <= 8) & ((((occ_a & ba) | (occ_b & bb)) == 0u) | ((occ_f & bf) == 0u));
#if SR_DENSE_PROBE3
		// The second hit of the true cluster finds no `last` yet and the sieves set: as a rule that cluster was opened by the read's
		// first subread, so the first three of the list are tried before the walk.
		if (!same && !clear && last_c < 0) {
			const int np = min(C.n, 3);
			int j = -1;
			if (np > 0 && C.w(0, 0) == kv) j = 0;
			else if (np > 1 && C.w(1, 0) == kv) j = 1;
			else if (np > 2 && C.w(2, 0) == kv) j = 2;
			if (j >= 0) { last_c = j; last_kv = kv; same = true; }
		}
#endif
		// (bitwise on purpose: with && / || the compiler sends the `same` lanes into the body ahead of the others and the body runs twice)
		const bool fresh = (!same) & clear;
		if (same | fresh) {
			if (fresh && (!(kv >= 32u && kv <= 0xffffffdfu) || C.n >= C.cap)) return false;   // u32 wrap-around of the probe rows; list full
			const int idx = same ? last_c : C.n;
			uint32_t *cw = C.base + idx * (CL::kWords * 32), *cs = cw + (2 + (sub >> 5)) * 32;   // the cluster's words; the one with the subread's bit (three-word clusters: sub < 32)
			const uint32_t bit = 1u << (sub & 31);
			uint32_t w1 = (kv >= low && kv <= high) ? ((uint32_t)l | ((uint32_t)l << 8) | ((uint32_t)u << 16) | 0x80000000u) : 0u;   // :1074
			uint32_t sw = 0u;
			if (same) { w1 = cw[32]; sw = *cs; }
			if (w1 >> 31) {
				if (sw & bit) return false;                            // one subread twice: the reference opens a duplicate slot
				sw |= bit;
				if ((uint32_t)l < (w1 & 255u)) w1 = (w1 & ~0x007f00ffu) | (uint32_t)l | ((uint32_t)u << 16);
				if ((uint32_t)l > ((w1 >> 8) & 255u)) w1 = (w1 & ~0x0000ff00u) | ((uint32_t)l << 8);
			}
			if (fresh) {
				cw[0] = kv;
				if (CL::kWords > 3) cw[(3 - (sub >> 5)) * 32] = 0u;
				C.n++;
				occ_a |= ba;
				occ_b |= bb;
				occ_f |= 1u << ((kv >> 9) & 31u);
			}
			cw[32] = w1;
			*cs = sw;
			return true;
		}
		int found = -1;
		const int n0 = C.n;
		if (!simple_hit(C, kv, l, u, sub, tol, low, high, &found)) return false;   // (inlined: as a real call the pass ran 25 % slower)
		if (C.n > n0) { occ_a |= ba; occ_b |= bb; occ_f |= 1u << ((kv >> 9) & 31u); }
		if (found >= 0) { last_kv = kv; last_c = found; }
		return true;
	};
	for (int s_base = 0; s_base < applied && ok; s_base += kSubs) {
		const int n_sub = min(kSubs, applied - s_base), base = s_base * GAP;
		uint32_t ra0 = 0, ra1 = 0, ra2 = 0, ra3 = 0;     // runs of keys with several positions waiting to be applied:
		uint32_t me0 = 0, me1 = 0, me2 = 0, me3 = 0;     // first item in vals[] (until read: the rank of the entry);  h | landing << 8 | lookup << 16
		int ne = 0;
		unsigned slow = 0;                               // lookups the unit could not answer (value beyond its 28, a fifth run): the direct filter does
		Record8 rn;
		uint32_t k0_next = 0;
		auto issue = [&](int si) {
			k0_next = simple_key_at<RCKEYS>(pk, simple_sub_offset<GAP>(s_base + min(si, n_sub - 1), step), rc_len);
			rn = ld_record(units4, 4u * (k0_next & kPhaseMask));         // (ld_record counts in 32-byte records: a unit is four)
			n_gather++;
		};
		// stage 2: what the unit says about its three keys.  st: two bits per role (0 nothing, 1 one position, 2 several); the
		// values the unit holds past its first four are loaded here and used a stage later
		auto look = [&](int si, uint32_t (&val)[3], int &st) {
			const Record8 r = rn;
			const uint32_t k0 = k0_next;
			if (si + 1 < n_sub) issue(si + 1);
			const int ob = simple_sub_offset<GAP>(s_base + si, step);    // (the other two keys only now: fewer registers across the fetch)
			const uint32_t k1 = simple_key_at<RCKEYS>(pk, ob + 1, rc_len), k2 = simple_key_at<RCKEYS>(pk, ob + 2, rc_len);
			const uint32_t c = k0 & kPhaseMask;
			st = 0;
			val[0] = val[1] = val[2] = 0u;
			if (((k1 >> 2) & kPhaseMask) != c || (k2 >> 4) != c) { slow |= 7u << (3 * si); return; }
			const uint32_t p0 = r.w[0] & 0x55555555u, p1 = r.w[1] & 0x55555555u;
#pragma unroll
			for (int x = 0; x < 3; x++) {
				const uint32_t key = x == 0 ? k0 : (x == 1 ? k1 : k2), v = phase_variant(key, x);
				const uint32_t wx = x == 0 ? r.w[0] : (x == 1 ? r.w[1] : r.w[2]);
				const uint32_t two = (wx >> (2 * v)) & 3u;
				if (!(two & 1u)) continue;
				const int rank = __popc(wx & 0x55555555u & ((1u << (2 * v)) - 1u)) + (x > 0 ? __popc(p0) : 0) + (x > 1 ? __popc(p1) : 0);
				if (rank >= kPhaseDenseValues) { slow |= 1u << (3 * si + x); continue; }
				st |= (1 + (int)((two >> 1) & 1u)) << (2 * x);
				if (rank < 4) val[x] = rank == 0 ? r.w[4] : (rank == 1 ? r.w[5] : (rank == 2 ? r.w[6] : r.w[7]));
				else {
					val[x] = ld_gather(ix.phase_dense + (size_t)kPhaseDenseWords * c + 4 + rank);       // same unit: as a rule the 64 bytes already fetched
					if (rank >= 12) n_gather++;                                                    // (the unit's second half: another fetch)
				}
			}
		};
		// stage 3: the entries of keys with several positions asked for (read behind the loop); the unit's single positions side
		// by side (role order) in front of ONE copy of hit(), run as often as the fullest lane needs -- 2.4 times per unit on
		// average instead of once per role
		auto apply = [&](int si, const uint32_t (&val)[3], int st) {
			const int sub = s_base + si, ob = simple_sub_offset<GAP>(sub, step);
			// keys with several positions (one unit in four holds one): ONE copy of the stash code, looped over the roles that have
			// one -- the entry is asked for here and read behind the subread loop
			int multi = ((st >> 1) & 1 & ~st) | ((st >> 2) & 2 & ~(st >> 1)) | ((st >> 3) & 4 & ~(st >> 2));   // bit x: role x holds code 2
#pragma unroll 1
			while (multi) {
				const int x = __ffs(multi) - 1;
				multi &= multi - 1;
				const uint32_t v = x == 0 ? val[0] : (x == 1 ? val[1] : val[2]);
				prefetch_l2(ix.direct_first + v);
				n_gather++;
				const int li = 3 * si + x;
				const uint32_t me = (uint32_t)(base + li) << 16;
				if (ne == 0) { ra0 = v; me0 = me; ne = 1; }
				else if (ne == 1) { ra1 = v; me1 = me; ne = 2; }
				else if (ne == 2) { ra2 = v; me2 = me; ne = 3; }
				else if (ne == 3) { ra3 = v; me3 = me; ne = 4; }
				else slow |= 1u << li;
			}
			uint32_t h0 = 0, h1 = 0, h2 = 0;
			int hx = 0, nh = 0;                                          // roles, two bits each
#pragma unroll
			for (int x = 0; x < 3; x++)
				if (((st >> (2 * x)) & 3) == 1) {
					const uint32_t kvx = val[x] - (uint32_t)(ob + x);
					if (nh == 0) h0 = kvx; else if (nh == 1) h1 = kvx; else h2 = kvx;
					hx |= x << (2 * nh);
					nh++;
				}
#pragma unroll 1
			for (int j = 0; j < nh; j++) {
				const uint32_t kvj = j == 0 ? h0 : (j == 1 ? h1 : h2);
				if (!hit(kvj, base + 3 * si + ((hx >> (2 * j)) & 3), 0, sub)) { ok = false; break; }
			}
		};
		issue(0);
#if SR_DENSE_PIPE
		uint32_t pv[3] = {0u, 0u, 0u};
		int pst = 0;
#pragma unroll 1
		for (int si = 0; si < n_sub; si++) {
			uint32_t cv[3];
			int cst;
			look(si, cv, cst);
			if (pst) apply(si - 1, pv, pst);
			pv[0] = cv[0]; pv[1] = cv[1]; pv[2] = cv[2];
			pst = cst;
		}
		if (pst) apply(n_sub - 1, pv, pst);
#else
#pragma unroll 1
		for (int si = 0; si < n_sub; si++) {
			uint32_t cv[3];
			int cst;
			look(si, cv, cst);
			if (cst) apply(si, cv, cst);
		}
#endif
		// the entries asked for in the loop (ra = their rank so far): all four loads under way, then run start, length and landing
		// index in place of the rank, and the runs' positions asked for in turn
		if (ok && ne) {
			uint2 e0 = make_uint2(0u, 0u), e1 = e0, e2 = e0, e3 = e0;
			e0 = ld_direct_first(ix, ra0);
			if (ne > 1) e1 = ld_direct_first(ix, ra1);
			if (ne > 2) e2 = ld_direct_first(ix, ra2);
			if (ne > 3) e3 = ld_direct_first(ix, ra3);
			uint32_t longest = e0.y & 255u;
			ra0 = e0.x; me0 = (e0.y & 0xffffu) | (me0 & 0xffff0000u);
			prefetch_l2(ix.vals + ra0); prefetch_l2(ix.vals + ra0 + (e0.y & 255u) - 1u);
			if (ne > 1) { ra1 = e1.x; me1 = (e1.y & 0xffffu) | (me1 & 0xffff0000u); longest = max(longest, e1.y & 255u); prefetch_l2(ix.vals + ra1); prefetch_l2(ix.vals + ra1 + (e1.y & 255u) - 1u); }
			if (ne > 2) { ra2 = e2.x; me2 = (e2.y & 0xffffu) | (me2 & 0xffff0000u); longest = max(longest, e2.y & 255u); prefetch_l2(ix.vals + ra2); prefetch_l2(ix.vals + ra2 + (e2.y & 255u) - 1u); }
			if (ne > 3) { ra3 = e3.x; me3 = (e3.y & 0xffffu) | (me3 & 0xffff0000u); longest = max(longest, e3.y & 255u); prefetch_l2(ix.vals + ra3); prefetch_l2(ix.vals + ra3 + (e3.y & 255u) - 1u); }
			if (longest > (uint32_t)C.cap) ok = false;                // cannot stay separable
		}
		if (!ok) return false;
		// the keys with several positions: one flat loop over all their positions in the reference's visit order; then, one by
		// one, what the units left open (through the direct filter, as MODE 2 would)
		while ((ne || slow) && ok) {
#pragma unroll 1
			while (slow && ne < 4 && ok) {
				const int i = __ffs(slow) - 1;
				slow &= slow - 1;
				const int l = base + i;
				int sub;
				const int o = simple_offset<GAP>(l, step, &sub);
				const uint32_t key = simple_key_at<RCKEYS>(pk, o, rc_len);
				const uint32_t u = direct_unit(key);
				uint32_t bits;
				const uint32_t rank = direct_rank(ld_direct_unit(ix, u), key, u, &bits);
				n_gather++;
				if (!bits) continue;
				const uint2 entry = ld_direct_first(ix, rank);
				n_gather++;
				if (bits == 1u) { if (!hit(entry.x - (uint32_t)o, l, 0, sub)) ok = false; continue; }
				const uint32_t me = (entry.y & 0xffffu) | ((uint32_t)l << 16);
				if ((entry.y & 255u) > (uint32_t)C.cap) { ok = false; break; }
				if (ne == 0) { ra0 = entry.x; me0 = me; } else if (ne == 1) { ra1 = entry.x; me1 = me; } else if (ne == 2) { ra2 = entry.x; me2 = me; } else { ra3 = entry.x; me3 = me; }
				ne++;
			}
			if (!ok) break;
			// positions t = 0 .. total-1 of all runs in a row; the load of position t + 1 under way while t is applied
			const int h0 = ne > 0 ? (int)(me0 & 255u) : 0, h1 = ne > 1 ? (int)(me1 & 255u) : 0, h2 = ne > 2 ? (int)(me2 & 255u) : 0, h3 = ne > 3 ? (int)(me3 & 255u) : 0;
			const int total = h0 + h1 + h2 + h3;
			uint32_t me_t = 0;
			int u_t = 0;
			auto fetch = [&](int t) -> uint32_t {
				uint32_t ra;
				if (t < h0) { ra = ra0; me_t = me0; u_t = t; }
				else if (t < h0 + h1) { ra = ra1; me_t = me1; u_t = t - h0; }
				else if (t < h0 + h1 + h2) { ra = ra2; me_t = me2; u_t = t - h0 - h1; }
				else { ra = ra3; me_t = me3; u_t = t - h0 - h1 - h2; }
				const int h = (int)(me_t & 255u), land = (int)((me_t >> 8) & 255u);
				const int nfwd = h - land;                               // landing .. end of run, then landing-1 .. start (:1109-1119)
				const int idx = u_t < nfwd ? land + u_t : land - 1 - (u_t - nfwd);
				return ld_gather(ix.vals + ra + idx);
			};
			uint32_t v_next = total > 0 ? fetch(0) : 0u;
			n_gather += (uint32_t)total;
#pragma unroll 1
			for (int t = 0; t < total; t++) {
				const uint32_t v = v_next, me = me_t;
				const int u = u_t;
				if (t + 1 < total) v_next = fetch(t + 1);
				const int l = (int)(me >> 16);
				int sub;
				const int o = simple_offset<GAP>(l, step, &sub);
				if (!hit(v - (uint32_t)o, l, u, sub)) { ok = false; break; }
			}
			ne = 0;
		}
	}
	if (!ok) return false;
#pragma unroll 1
	for (int c = 0; c < C.n; c++) {
		C.w(c, 2) = (uint32_t)C.votes(c);
		if (CL::kWords > 3) C.w(c, 3) = 0u;
		C.w(c, 1) |= vote_row(C.w(c, 0)) << 24;                    // (hit() leaves the row field to this loop: once per cluster, all lanes together)
	}
	C_out.n = C.n;
	return true;
}

template <int GAP, int MODE, bool RCKEYS = false>
__device__ __forceinline__ bool simple_pass(const DevIndex &ix, const VoteConfig &cfg, const uint32_t *pk, int L, int applied, int step,
                                            SimpleClusters &C, int rc_len, uint32_t &n_gather, bool have_first_mask = false, unsigned first_mask = 0u,
                                            uint32_t *sec2 = nullptr /* shared-memory word of the thread: one-indel reads stay in the tier; on return the merge's word */)
{
	if (MODE == 3) return simple_pass_phase<RCKEYS>(ix, cfg, pk, L, applied, step, C, rc_len, n_gather, sec2);
	if (MODE == 4) return simple_pass_phase_dense<RCKEYS, SimpleClusters>(ix, cfg, pk, L, applied, step, C, rc_len, n_gather);
	constexpr bool DIRECT = MODE == 2;
	constexpr int kSubsPerBatch = 32 / GAP;                        // whole subreads per batch of <= 32 lookups
	const uint32_t nb = ix.buckets_number;
	const uint64_t magic = ix.buckets_magic;
	const int tol = cfg.indel_tolerance;
	const uint32_t low = cfg.low_border, high = cfg.high_border - (uint32_t)L;
	const uint4 *records = reinterpret_cast<const uint4 *>(ix.records);
	C.n = 0;
	if (C.cap < 1) return false;      // (paired-end tier: the first mate took the whole cluster pool)
	SimpleCluster P;                  // cluster 0 while have_p
	P.kv = 0; P.w1 = 0; P.subs = 0;
	bool have_p = false, ok = true;
	if (sec2) *sec2 = 0u;                                                       // one-indel reads in closed form (simple_primary_hit; not on dense tables)
	for (int s_base = 0; s_base < applied; s_base += kSubsPerBatch) {
		const int n_sub = min(kSubsPerBatch, applied - s_base), base = s_base * GAP;
		// ---- (1) filter ----
		unsigned mask = (have_first_mask && s_base == 0) ? first_mask : simple_filter<GAP, MODE, RCKEYS>(ix, pk, s_base, n_sub, step, rc_len, n_gather);
		// ---- (2) survivors: single positions of the primary cluster ----
		// One step ahead of the survivor being applied, the load of the next one is already under way.
		unsigned defer = mask;
		uint32_t ra0 = 0, ra1 = 0, ra2 = 0, ra3 = 0;     // dense table: runs of keys with several positions waiting to be applied:
		uint32_t me0 = 0, me1 = 0, me2 = 0, me3 = 0;     // first item in vals[];  h | landing << 8 | lookup << 16
		int ne = 0;
		if (DIRECT || records) {
			defer = 0;
			Record8 rn;
			int i_next = -1, o_next = 0, sub_next = 0;
			uint32_t key_next = 0, q_next = 0;
			auto issue = [&]() {
				i_next = -1;
				if (!mask) return;
				i_next = __ffs(mask) - 1;
				mask &= mask - 1;
				o_next = simple_offset<GAP>(base + i_next, step, &sub_next);
				key_next = simple_key_at<RCKEYS>(pk, o_next, rc_len);
				if (DIRECT) {
					const uint4 un = ld_direct_unit(ix, direct_unit(key_next));   // an L2 hit: the filter pass fetched this line
					rn.w[0] = un.x; rn.w[1] = un.y; rn.w[2] = un.z; rn.w[3] = un.w;
				} else {
					q_next = (uint32_t)__umul64hi((uint64_t)key_next, magic);
					rn = ld_record(records, key_next - q_next * nb);
					n_gather++;
				}
			};
			issue();
			while (i_next >= 0) {
				const Record8 r = rn;
				const int i = i_next, o = o_next, sub = sub_next, l = base + i;
				const uint32_t key = key_next, q = q_next;
				issue();
				int kind;
				uint32_t kv;
				uint2 entry = make_uint2(0u, 0u);
				if (DIRECT) {
					uint32_t bits;
					const uint32_t rank = direct_rank(make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]), key, direct_unit(key), &bits);
					kind = bits == 1u ? 1 : (bits ? 2 : 0);
					if (kind) { entry = ld_direct_first(ix, rank); n_gather++; }
					kv = entry.x - (uint32_t)o;
				} else {
					uint32_t value = 0;
					kind = simple_probe(ix, r, (int)(int16_t)q, &value, n_gather);
					kv = value - (uint32_t)o;
				}
				if (DIRECT) {
					// dense table: every fourth random 16-mer has a position, so a pass sees a handful of one-vote clusters
					// next to the true one and a register-resident primary cluster would diverge the warp (the first hit
					// is often a random one).  Every single position goes through the cluster list at once -- clusters do
					// not depend on the order of application -- and the warp stays converged in this loop.
					if (kind == 1) { if (!simple_hit(C, kv, l, 0, sub, tol, low, high)) ok = false; }
					else if (kind != 0) {
						// several positions: the entry holds the run (first item, length, landing index); up to four wait in
						// registers for the flat loop below, more go the long way
						const uint32_t me = (entry.y & 0xffffu) | ((uint32_t)l << 16);
						if ((entry.y & 255u) > (uint32_t)C.cap) ok = false;       // cannot stay separable
						else if (ne == 0) { ra0 = entry.x; me0 = me; ne = 1; }
						else if (ne == 1) { ra1 = entry.x; me1 = me; ne = 2; }
						else if (ne == 2) { ra2 = entry.x; me2 = me; ne = 3; }
						else if (ne == 3) { ra3 = entry.x; me3 = me; ne = 4; }
						else defer |= 1u << i;
					}
				} else if (kind == 1) {
					if (simple_primary_hit(P, have_p, C, DIRECT ? nullptr : sec2, ok, kv, l, sub, tol, low, high)) defer |= 1u << i;
				} else if (kind != 0)
					defer |= 1u << i;
			}
			if (!ok) return false;
		}
		// ---- (3) everything else, on the general path ----
		if (DIRECT) {
			// Dense table: what is left are the keys with several positions (about one lookup in twenty), their runs already
			// in registers: one flat loop over all their positions in the reference's visit order, so that the warp stays
			// converged while the hits are applied.
			while ((ne || defer) && ok) {
#pragma unroll 1
				while (defer && ne < 4) {                                // (rare: a fifth key with several positions in one batch)
					const int i = __ffs(defer) - 1;
					defer &= defer - 1;
					const int l = base + i;
					int sub;
					const int o = simple_offset<GAP>(l, step, &sub);
					const uint32_t key = simple_key_at<RCKEYS>(pk, o, rc_len);
					const uint32_t u = direct_unit(key);
					uint32_t bits;
					const uint2 entry = ld_direct_first(ix, direct_rank(ld_direct_unit(ix, u), key, u, &bits));
					n_gather++;
					const uint32_t me = (entry.y & 0xffffu) | ((uint32_t)l << 16);
					if (ne == 0) { ra0 = entry.x; me0 = me; } else if (ne == 1) { ra1 = entry.x; me1 = me; } else if (ne == 2) { ra2 = entry.x; me2 = me; } else { ra3 = entry.x; me3 = me; }
					ne++;
				}
				int e = 0, u = 0;
#pragma unroll 1
				while (e < ne) {
					const uint32_t ra = e == 0 ? ra0 : (e == 1 ? ra1 : (e == 2 ? ra2 : ra3)), me = e == 0 ? me0 : (e == 1 ? me1 : (e == 2 ? me2 : me3));
					const int h = (int)(me & 255u), land = (int)((me >> 8) & 255u), l = (int)(me >> 16);
					if (u >= h) { e++; u = 0; continue; }
					const int nfwd = h - land;                           // landing .. end of run, then landing-1 .. start (:1109-1119)
					const int idx = u < nfwd ? land + u : land - 1 - (u - nfwd);
					const uint32_t v = ld_gather(ix.vals + ra + idx);
					n_gather++;
					int sub;
					const int o = simple_offset<GAP>(l, step, &sub);
					if (!simple_hit(C, v - (uint32_t)o, l, u, sub, tol, low, high)) { ok = false; break; }
					u++;
				}
				ne = 0;
			}
			if (!ok) return false;
		} else if (defer) {
			// A lane on the general path holds up its warp: beyond a few lookups the fragment is better off in the warp kernel
			// (an indel splits the hits over two kv, a repeat answers every lookup with several positions)
			if (__popc(defer) > SR_SIMPLE_DEFER_LIMIT) return false;
			if (have_p) C.put(0, P);
			const unsigned long long before = P.subs;
			if (!simple_general<GAP, RCKEYS>(ix, cfg, pk, step, rc_len, base, defer, C, low, high, n_gather)) return false;
			if (have_p) P = C.get(0);
			if (!DIRECT && sec2 && *sec2 && P.subs != before) return false;   // (a late hit on the first offset of a read with an indel)
		}
	}
	if (have_p) C.put(0, P);
	if (!DIRECT && sec2 && *sec2) {
		const uint32_t m = simple_indel_close(C, *sec2, tol, GAP);
		if (!m) return false;
		*sec2 = m;
	}
#pragma unroll 1
	for (int c = 0; c < C.n; c++) {                                   // the subread sets have served: word 2 becomes the vote count,
		C.w(c, 2) = (uint32_t)C.votes(c);                             // word 3 is free for the merge (paired-end tier: chromosome memo)
		C.w(c, 3) = 0u;
	}
	return true;
}

// update_top_three (core-junction.c:908) on three registers; ts = top_scores.
__device__ __forceinline__ void simple_update_top(int ts, int &t0, int &t1, int &t2, int v)
{
	const int last = ts == 1 ? t0 : (ts == 2 ? t1 : t2);
	if (v <= last) return;
	if (v > t0) {
		if (ts > 2) t2 = t1;
		if (ts > 1) t1 = t0;
		t0 = v;
	} else if (v == t0) {
	} else if (ts > 1 && v > t1) {
		if (ts > 2) t2 = t1;
		t1 = v;
	} else if (ts > 1 && v == t1) {
	} else if (ts > 2 && v > t2) t2 = v;
}

// process_voting_junction_PE_topK for one single-end pass over a separable table (core-junction.c:2199-2530,
// the SE branch :2436-2495) and the hand-over logic of do_voting (core.c:3212-3226).  res[] is fully unrolled so
// that it stays in registers; the loops over clusters are kept rolled (small code matters more than their speed).
// simple_pass has left every cluster's vote count in its word 2.
template <int GAP>
__device__ __forceinline__ void simple_merge(const VoteConfig &cfg, const SimpleClusters &C, int strand, int applied, int step, SimpleResult (&res)[kMaxBest],
                                             uint32_t *p_ind = nullptr /* in: cluster 0's second indel section (simple_primary_hit), 0 = none; out: the
                                                                          slot's `ind` word if cluster 0 was emitted with it, else 0 */)
{
	const uint32_t p_in = p_ind ? *p_ind : 0u;
	if (p_ind) *p_ind = 0u;
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
	for (int c = 0; c < C.n + kMaxBest; c++) {           // the table's slots, then the slots the bigtable already holds
		int v;
		if (c < C.n) v = (int)C.w(c, 2);
		else {
			const int k = c - C.n;
			v = k >= mb ? 0 : (k == 0 ? res[0].votes : (k == 1 ? res[1].votes : res[2].votes));
		}
		if (v > 0) simple_update_top(ts, t0, t1, t2, v);
	}

	// The reference first lists the candidates level by level (:2251-2330) and then consumes the list in order
	// (:2436-2495); consuming each candidate as soon as it is listed is the same thing.
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
		for (;;) {                                       // slots of this level in (row, j) order = the reference's i, j loops
			int best = -1, best_key = 0x7fffffff;
#pragma unroll 1
			for (int c = 0; c < C.n; c++)
				if (!((done >> c) & 1u) && (int)C.w(c, 2) == want) {
					const int key = C.order(c);
					if (key < best_key) { best_key = key; best = c; }
				}
			if (best < 0 || n >= max_simples) break;
			done |= 1u << best;
			if (want >= cfg.min_second) {
				n++;
				const uint32_t pos = C.w(best, 0);
				if (cursor < mb && want >= cfg.min_first && !exists(pos)) {
					const uint32_t w1 = C.w(best, 1);
					int s_first, s_last;
					const int o_first = simple_offset<GAP>((int)(w1 & 255u), step, &s_first);
					const int o_last = simple_offset<GAP>((int)((w1 >> 8) & 255u), step, &s_last);
					SimpleResult r;
					r.pos = pos; r.votes = want; r.used = applied;
					r.meta = 1u | ((uint32_t)strand << 8) | ((uint32_t)(s_first + 1) << 16) | ((uint32_t)(s_last + 1) << 24);
					r.cov = (uint32_t)o_first | ((uint32_t)(o_last + 16) << 16);
					if (best == 0 && p_in) {                             // {first1 + 1, last1 + 1, 0, first2 + 1, last2 + 1, d}
						r.meta = (r.meta & 0x00ffffffu) | ((p_in & 255u) << 24) | 0x200u;
						*p_ind = ((p_in >> 8) & 255u) | ((uint32_t)(s_last + 1) << 8) | (p_in & 0x80ff0000u);
					}
					put(r);
				}
			}
		}
#pragma unroll
		for (int k = 0; k < kMaxBest; k++) {
			if (k < mb && n < max_simples && res[k].votes == want) {
				n++;
				if (cursor < mb && want >= cfg.min_first && !exists(res[k].pos)) put(res[k]);
			}
		}
	}
#pragma unroll
	for (int k = 0; k < kMaxBest; k++) {                 // :2505-2527
		if (k < mb) {
			if (k < cursor) res[k] = tmp[k];
			else res[k].votes = 0;
		}
	}
}

#endif  // __CUDACC__
}  // namespace sr
#include "sr_simple_junc.cuh"   // the junction half for subjunc chunks (needs the definitions above)
namespace sr {
#if defined(__CUDACC__) && !defined(SR_HOST_ONLY)

// k_map_simple: thread t of block b maps single-end fragment 128 b + t of a chunk whose bigtable slice is all
// zero.  Writes the fragment's result slots, or queues the fragment for the warp kernel (its slots are then zero).
// Every warp works on its own 32 fragments and its own slice of shared memory: there is no block-wide barrier.
// Dynamic shared memory: simple_smem_bytes(max_len, clusters).
// JUNC (subjunc): also the fragment's subjunc_result_t slots (jout) and big-margin record (bmout), both all zero on entry;
// dynamic shared memory simple_junc_smem_bytes(max_len, clusters).
// RMW (second and later index blocks): the slice holds the results of the blocks before; a fragment whose used slots all
// have the plain form this tier produces itself (no indel, no junction bits) merges into them, any other is queued.
// INDEL: reads with one indel stay in the tier in closed form (simple_primary_hit; sparse tables, not subjunc).  It costs the
// error-free chunk 13 % (register pressure in the lookup loop), so the host launches this form only while the chunks it sees
// actually hold such reads (srcuda_handle::tier_indel, from the last call's counters); the bytes are the same either way.
// indel_count: fragments whose result came from that closed form.
template <int GAP, int MODE, bool JUNC = false, bool RMW = false, bool INDEL = false>
__global__ void __launch_bounds__(kSimpleThreads, JUNC ? 4 : ((MODE == 2 || MODE == 4) ? 5 : SR_SIMPLE_MIN_BLOCKS))   // (dense tables: shared memory allows five blocks anyway)
k_map_simple(const DevIndex ix, const VoteConfig cfg, const uint8_t *text, uint64_t text_origin, uint64_t text_bytes,
             const uint64_t *off, const uint16_t *len, int64_t n_frags, int max_len, int clusters, ResultWords *out, uint32_t *queue,
             uint32_t *queue_count, uint32_t *bad_reads, uint32_t *gather_count, JuncWords *jout = nullptr, uint16_t *bmout = nullptr,
             uint32_t *indel_count = nullptr)
{
	static_assert(!INDEL || (!JUNC && (MODE == 0 || MODE == 3)), "the one-indel closed form lives on the primary cluster of a sparse table");
	extern __shared__ uint4 smem4[];
	constexpr unsigned kFullMask = 0xffffffffu;
	const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
	const int n_words = simple_words(max_len);
	const int warp_cap = simple_warp_bytes(max_len, clusters);
	uint32_t *pk = reinterpret_cast<uint32_t *>(smem4) + t;                                     // [2][n_words][128]
	uint8_t *stage = reinterpret_cast<uint8_t *>(reinterpret_cast<uint32_t *>(smem4) + 2 * n_words * kSimpleThreads) + wid * warp_cap;
	const int64_t fw0 = (int64_t)blockIdx.x * kSimpleThreads + 32 * wid, f = fw0 + lane;
	const int mb = cfg.multi_best_reads;
	if (fw0 >= n_frags) return;

	// ---- 1. stage the text ----
	bool active = f < n_frags, punt = false;
	uint64_t o0 = 0;
	int L = 0;
	if (active) {
		o0 = off[f] - text_origin;
		L = len[f];
		if (o0 > text_bytes || (uint64_t)L > text_bytes - o0 || L > kMaxReadLen) {
			atomicAdd(bad_reads, 1u);
			active = false;
		} else if (L > kSimpleMaxLen || L > max_len) punt = true;
	}
	const bool want = active && !punt && L >= 16;
	unsigned long long lo = want ? o0 : ~0ull, hi = want ? o0 + (uint64_t)L : 0ull;
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		lo = min(lo, __shfl_xor_sync(kFullMask, lo, d));
		hi = max(hi, __shfl_xor_sync(kFullMask, hi, d));
	}
	unsigned so = 0;
	if (hi > lo) {
		const uint8_t *g0 = text + lo;
		const unsigned head = (unsigned)(reinterpret_cast<uintptr_t>(g0) & 15u);
		if (hi - lo + head <= (unsigned long long)warp_cap) {
			// the usual layout: the warp's reads lie back to back
			const uint4 *g4 = reinterpret_cast<const uint4 *>(g0 - head);
			const int n4 = (int)((hi - lo + head + 15) >> 4);
			uint4 *s4 = reinterpret_cast<uint4 *>(stage);
#pragma unroll 2
			for (int i = lane; i < n4; i += 32) s4[i] = __ldcs(g4 + i);
			so = (unsigned)(o0 - lo) + head;
		} else {
			// scattered reads: every thread copies its own into a private slot
			const int slot = ((warp_cap - 16) / 32) & ~3;
			if (want) {
				const uint8_t *g = text + o0;
				const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(g) & 3u);
				const int nw4 = (int)((mis + (unsigned)L + 3u) >> 2);
				if (nw4 * 4 > slot) punt = true;
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

	// ---- 1b. pack both strands ----
	const bool run = active && !punt && L >= 16;
	if (run && !simple_pack<true, !JUNC>(stage, so, L, pk, n_words)) punt = true;      // (subjunc compares 2-bit codes with the genome: no Ns there)
	__syncwarp();   // the staging area becomes the cluster store

	// ---- 2./3. two strand passes ----
	SimpleResult res[kMaxBest];
#pragma unroll
	for (int k = 0; k < kMaxBest; k++) { res[k].pos = 0; res[k].votes = 0; res[k].used = 0; res[k].meta = 0; res[k].cov = 0; }
	uint32_t old_w3 = 0;                     // RMW: word 3 of slot 0, for the single-end fall-through (core.c:3220)
	if (RMW && active && !punt) {
		const ResultWords *src = out + (size_t)f * mb;
#pragma unroll
		for (int k = 0; k < kMaxBest; k++) {
			if (k >= mb) break;
			const uint32_t w1 = src[k].w[1], w2 = src[k].w[2], w3 = src[k].w[3];
			res[k].votes = (int)(int16_t)(w2 & 0xffffu);
			res[k].used = (int)(int16_t)(w2 >> 16);
			if (k == 0) { old_w3 = w3; res[0].meta = 2u; }
			if (res[k].votes >= 1) {
				uint32_t rest = w3 | src[k].w[16] | (w1 & ~(uint32_t)kCoreNegStrand);
#pragma unroll
				for (int w = 5; w < 15; w++) rest |= src[k].w[w];
				const uint32_t w4 = src[k].w[4], first = w4 & 0xffffu, last = w4 >> 16;
				if (rest != 0 || first < 1 || first > 63 || last < 1 || last > 63) punt = true;
				res[k].pos = src[k].w[0];
				res[k].meta = 1u | ((w1 & (uint32_t)kCoreNegStrand) ? 0x100u : 0u) | (first << 16) | (last << 24);
				res[k].cov = src[k].w[15];
			}
		}
	}
	SimpleClusters C;
	C.base = reinterpret_cast<uint32_t *>(stage) + lane;
	C.n = 0;
	C.cap = clusters;
	// subjunc: the fragment's subjunc_result_t slots so far / being built and its big-margin record, behind the block's other areas
	uint32_t *jblock = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(smem4) + simple_smem_bytes(max_len, clusters));
	uint32_t *J = jblock + t, *JT = J + kMaxBest * 4 * kSimpleThreads;
	uint16_t *bm = reinterpret_cast<uint16_t *>(jblock + 2 * kMaxBest * 4 * kSimpleThreads + 6 * t);
	if (JUNC) {
#pragma unroll
		for (int i = 0; i < kMaxBest * 4; i++) J[i * kSimpleThreads] = 0u;
#pragma unroll
		for (int i = 0; i < 12; i++) bm[i] = 0;
	}
	uint32_t n_gather = 0;
	// per strand pass: second section of the indel record of the slot whose meta bit 9 is set -- in shared memory behind the clusters
	// (word 0: the running pass's, simple_primary_hit; words 1, 2: what the merges of strand 0 / 1 emitted)
	uint32_t *const indw = reinterpret_cast<uint32_t *>(stage + simple_cluster_bytes(clusters)) + lane;
	if (INDEL) { indw[32] = 0u; indw[64] = 0u; }
	if (active && !punt) {
		// a read shorter than 16 bases goes through two empty passes (core.c:3116)
		int step = 65536;
		const int applied = L >= 16 ? subread_schedule(L, GAP, cfg.total_subreads, &step) : 0;
		// Which strand first?  The result does not depend on it -- the two passes fill separate cluster lists and only the merges
		// (process_voting_junction_PE_topK, below) have an order -- but the warp's time does: the survivors loop of a pass runs about
		// eleven times on the read's own strand and once or twice on the other one, and the lanes of a warp hold reads of both
		// strands, so with a fixed order half of the lanes idle in either pass (17 of 32 threads active in round 1).  On a sparse
		// table both strands' first batch is filtered up front and every lane takes its fuller strand first: the long loops of all
		// lanes run side by side, then the short ones.  The lists of both passes then live at the same time and share the
		// cluster pool (the second list starts where the first one ended).
		constexpr bool kReorder = MODE == 0 && SR_SIMPLE_REORDER;
		unsigned m0 = 0, m1 = 0;
		int first = 0;
		if (kReorder && applied > 0) {
			const int n_sub0 = min(32 / GAP, applied);
			m0 = simple_filter<GAP, MODE, false>(ix, pk, 0, n_sub0, step, 0, n_gather);
			m1 = simple_filter<GAP, MODE, false>(ix, pk + n_words * kSimpleThreads, 0, n_sub0, step, 0, n_gather);
			first = __popc(m1) > __popc(m0) ? 1 : 0;
		}
		SimpleClusters C0 = C, C1 = C;                                // by strand
#pragma unroll 1
		for (int k = 0; k < 2 && !punt; k++) {                       // one copy of the pass in the instruction cache
			const int strand = first ^ k;
			const uint32_t *pass = pk + strand * n_words * kSimpleThreads;
			SimpleClusters Ck = C;
			if (kReorder && k == 1) {                                 // behind the first pass's list
				const int used = first ? C1.n : C0.n;
				Ck.base = C.base + used * kSimpleClusterWords * 32;
				Ck.cap = C.cap - used;
			}
			// a read with one indel stays in the tier (not in subjunc mode, not with the strand reordering): its second section's word
			uint32_t *const sec2 = (!INDEL || kReorder) ? nullptr : indw;
			if (!simple_pass<GAP, MODE>(ix, cfg, pass, L, applied, step, Ck, 0, n_gather, kReorder && applied > 0, strand ? m1 : m0, sec2)) punt = true;
			else if (!kReorder) {                                     // one list, merged at once
				if (JUNC) simple_merge_junc<GAP>(ix, cfg, Ck, strand, applied, step, res, J, JT, bm, TierText{pass, 0}, L);
				else {
					uint32_t p_ind = sec2 ? *sec2 : 0u;
					simple_merge<GAP>(cfg, Ck, strand, applied, step, res, &p_ind);
					if (sec2) indw[32 + 32 * strand] = p_ind;
				}
			}
			if (strand == 0) C0 = Ck; else C1 = Ck;
		}
		if (kReorder && !punt) {
#pragma unroll 1
			for (int strand = 0; strand < 2; strand++) {
				const SimpleClusters &Cs = strand ? C1 : C0;
				SimpleClusters Cm;
				Cm.base = Cs.base; Cm.n = Cs.n; Cm.cap = Cs.cap;
				const uint32_t *pass = pk + strand * n_words * kSimpleThreads;
				if (JUNC) simple_merge_junc<GAP>(ix, cfg, Cm, strand, applied, step, res, J, JT, bm, TierText{pass, 0}, L);
				else simple_merge<GAP>(cfg, Cm, strand, applied, step, res);
			}
		}
	}
	n_gather = __reduce_add_sync(kFullMask, n_gather);
	if (lane == 0 && n_gather) atomicAdd(gather_count, n_gather);
	if (INDEL) {
		const unsigned im = __ballot_sync(kFullMask, active && !punt && (indw[32] | indw[64]) != 0u);
		if (lane == 0 && im && indel_count) atomicAdd(indel_count, (uint32_t)__popc(im));
	}
	__syncwarp();
	// queue the punted fragments (one atomic per warp)
	const unsigned pm = __ballot_sync(kFullMask, active && punt);
	if (pm) {
		uint32_t qb = 0;
		if (lane == __ffs(pm) - 1) qb = atomicAdd(queue_count, (uint32_t)__popc(pm));
		qb = __shfl_sync(kFullMask, qb, __ffs(pm) - 1);
		if (active && punt) queue[qb + __popc(pm & ((1u << lane) - 1u))] = (uint32_t)f;
	}

	// ---- 4. the warp's bigtable slice (copy_vote_to_alignment_res image of each slot, core-junction.c:1058-1072) ----
	if (RMW) {
		// Later index block: nothing is cleared.  An emitted slot is a complete record (the reference copies a zeroed temporary,
		// :2516); any other slot only gets its selected_votes / used_subreads word, whatever else it holds stays (:2517).
		if (!active || punt) return;
		ResultWords *dst = out + (size_t)f * mb;
#pragma unroll
		for (int k = 0; k < kMaxBest; k++) {
			if (k >= mb) break;
			const SimpleResult &r = res[k];
			const uint32_t kind = r.meta & 0xffu, w2 = (uint32_t)(uint16_t)r.votes | ((uint32_t)(uint16_t)r.used << 16);
			if (kind == 1) {
#pragma unroll
				for (int w = 0; w < kResultWords; w++) {
					uint32_t v = 0;
					if (w == 0) v = r.pos;
					else if (w == 1) v = ((r.meta >> 8) & 1u) ? (uint32_t)kCoreNegStrand : 0u;
					else if (w == 2) v = w2;
					else if (w == 4) v = ((r.meta >> 16) & 0xffu) | (((r.meta >> 24) & 0xffu) << 16);
					else if (w == 15) v = r.cov;
					else if (INDEL && (r.meta & 0x200u)) {
						const uint32_t ind = indw[((r.meta >> 8) & 1u) ? 64 : 32];
						const uint32_t d16 = (uint32_t)(uint16_t)(int16_t)(int8_t)(ind >> 16);
						if (w == 3) v = (d16 & 0xffu) << 8;
						else if (w == 5) v = (ind & 0xffu) << 16;
						else if (w == 6) v = ((ind >> 8) & 0xffu) | (d16 << 16);
					}
					dst[k].w[w] = v;
				}
			} else {
				dst[k].w[2] = w2;
				if (k == 0 && kind == 2) dst[0].w[3] = old_w3 & ~0xffu;
			}
		}
		return;
	}
#if SR_SIMPLE_SMEM_OUT
	// Composed in the warp's shared area (the clusters are dead now) and stored as whole 16-byte pieces, half a warp
	// at a time: no partial-sector writes, so nothing has to be read back from DRAM to merge them.
	{
		const int64_t nf = min((int64_t)32, n_frags - fw0);
		uint32_t *sw = reinterpret_cast<uint32_t *>(stage);
		uint32_t *gdst = reinterpret_cast<uint32_t *>(out + (size_t)fw0 * mb);
		const int words_per_frag = mb * kResultWords;
		const bool vec = (reinterpret_cast<uintptr_t>(gdst) & 15u) == 0 && ((16 * words_per_frag) & 3) == 0;
#pragma unroll 1
		for (int half = 0; half < 2; half++) {
			const int first = 16 * half, cnt = (int)min((int64_t)16, nf - first);
			if (cnt <= 0) break;
			const int total = cnt * words_per_frag;
			__syncwarp();
			for (int i = lane; i < total; i += 32) sw[i] = 0u;
			__syncwarp();
			if (active && !punt && lane >= first && lane < first + 16) {
				uint32_t *d = sw + (lane - first) * words_per_frag;
#pragma unroll
				for (int k = 0; k < kMaxBest; k++) {
					if (k >= mb) break;
					const SimpleResult &r = res[k];
					const uint32_t kind = r.meta & 0xffu;
					uint32_t *dk = d + k * kResultWords;
					if (kind == 1) {
						dk[0] = r.pos;
						dk[1] = ((r.meta >> 8) & 1u) ? (uint32_t)kCoreNegStrand : 0u;
						dk[2] = (uint32_t)(uint16_t)r.votes | ((uint32_t)(uint16_t)r.used << 16);
						dk[4] = ((r.meta >> 16) & 0xffu) | (((r.meta >> 24) & 0xffu) << 16);
						dk[15] = r.cov;
					} else if (kind == 2)
						dk[2] = (uint32_t)(uint16_t)r.used << 16;
				}
			}
			__syncwarp();
			uint32_t *g = gdst + first * words_per_frag;
			if (vec) {
				const uint4 *s4 = reinterpret_cast<const uint4 *>(sw);
				uint4 *g4 = reinterpret_cast<uint4 *>(g);
				const int n4 = total >> 2;
				for (int i = lane; i < n4; i += 32) __stcs(g4 + i, s4[i]);
				for (int i = (n4 << 2) + lane; i < total; i += 32) __stcs(g + i, sw[i]);
			} else
				for (int i = lane; i < total; i += 32) __stcs(g + i, sw[i]);
		}
	}
#else
	// zero-fill the warp's slice with coalesced stores, then scatter the few non-zero words on top while the lines are
	// still in L2 (done early, the zeroed lines are written back before the scatter and every 4-byte store has to read
	// its sector back from DRAM)
	{
		const int64_t nf = min((int64_t)32, n_frags - fw0);
		uint32_t *dst = reinterpret_cast<uint32_t *>(out + (size_t)fw0 * mb);
		const int total = (int)nf * mb * kResultWords;
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
		uint32_t *jd = reinterpret_cast<uint32_t *>(jout + (size_t)fw0 * mb);
		const int jtotal = (int)nf * mb * 4;
#pragma unroll 1
		for (int i = lane; i < jtotal; i += 32) jd[i] = 0u;
		uint16_t *bd = bmout + (size_t)fw0 * kBigMarginU16;
		const int btotal = (int)nf * kBigMarginU16;
#pragma unroll 1
		for (int i = lane; i < btotal; i += 32) bd[i] = 0;
	}
	__syncwarp();   // orders the scattered stores after the zero-fill
	if (!active || punt) return;
	ResultWords *dst = out + (size_t)f * mb;
#pragma unroll
	for (int k = 0; k < kMaxBest; k++) {
		if (k >= mb) break;
		const SimpleResult &r = res[k];
		const uint32_t kind = r.meta & 0xffu;
		if (kind == 1) {
			__stcs(&dst[k].w[0], r.pos);
			__stcs(&dst[k].w[1], (((r.meta >> 8) & 1u) ? (uint32_t)kCoreNegStrand : 0u) | (JUNC ? (r.meta >> 12) & 7u : 0u));
			__stcs(&dst[k].w[2], (uint32_t)(uint16_t)r.votes | ((uint32_t)(uint16_t)r.used << 16));
			__stcs(&dst[k].w[4], ((r.meta >> 16) & 0xffu) | (((r.meta >> 24) & 0xffu) << 16));
			__stcs(&dst[k].w[15], r.cov);
			if (INDEL && (r.meta & 0x200u)) {                        // second section of the indel record; indels_in_confident_coverage = its indel
				const uint32_t ind = indw[((r.meta >> 8) & 1u) ? 64 : 32];
				const uint32_t d16 = (uint32_t)(uint16_t)(int16_t)(int8_t)(ind >> 16);
				__stcs(&dst[k].w[3], (d16 & 0xffu) << 8);
				__stcs(&dst[k].w[5], (ind & 0xffu) << 16);
				__stcs(&dst[k].w[6], ((ind >> 8) & 0xffu) | (d16 << 16));
			}
		} else if (kind == 2) {
			__stcs(&dst[k].w[2], (uint32_t)(uint16_t)r.used << 16);
		}
	}
	if (JUNC) {
#pragma unroll 1
		for (int k = 0; k < mb; k++)
#pragma unroll
			for (int i = 0; i < 4; i++) {
				const uint32_t w = J[(k * 4 + i) * kSimpleThreads];
				if (w) jout[(size_t)f * mb + k].w[i] = w;
			}
#pragma unroll 1
		for (int i = 0; i < 12; i++)
			if (bm[i]) bmout[(size_t)f * kBigMarginU16 + i] = bm[i];
	}
#endif
}

#endif  // __CUDACC__

}  // namespace sr
