// sr_api.cu -- libsubread_cuda.so: kernels' launchers and the C ABI of include/subread_cuda.h.
//
// Data layout in HBM (one index block, resident for the life of the handle):
//   bucket_start u32[buckets+1] | keys i16[items] | vals u32[items] | genome 2-bit (u32 words) |
//   chr_offsets u32[chromosomes]
// Per chunk batch (kSets buffer sets on as many streams so that H2D, kernels and D2H overlap):
//   text u8[...] | off u64[reads] | len u16[reads] | out 68 B x reads x 3 | verify u16 x reads x 3 |
//   overflow queue u32[frags] + counter
#include "../../include/subread_cuda.h"
#include "sr_config.h"
#include "sr_index.h"
#include "sr_scalar.cuh"
#include "sr_simple.cuh"
#include "sr_simple_pe.cuh"
#include "sr_warp.cuh"

#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace sr {

constexpr int kWarpsPerBlock = 8;
#ifndef SR_MIN_BLOCKS
#define SR_MIN_BLOCKS 5   // resident blocks (x 8 warps) per SM the register allocation is planned for: the kernel is latency-bound, 40 warps beat 24 (A/B in profiles/)
#endif
#ifndef SR_JUNC_MIN_BLOCKS
#define SR_JUNC_MIN_BLOCKS 4   // the warp kernel with subjunc's junction search: 64 registers (A/B against 3 and 5)
#endif
constexpr int kScalarThreads = 32;
constexpr int kScalarBlocksPerSM = 1;        // overflow queue of the warp kernel: a handful of fragments
constexpr int kScalarBlocksPerSMAll = 16;    // every fragment takes the scalar path (subjunc, forced): fill the SMs

// ------------------------------------------------------------------------------------ kernels

struct ChunkArgs {
	const uint8_t *text;       // device
	uint64_t text_origin;      // off[] values are relative to this
	uint64_t text_bytes;       // bytes available at `text`
	const uint64_t *off;       // device, per read of this launch
	const uint16_t *len;       // device
	int64_t n_frags;
	ResultWords *out;          // device, [frag][mate][multi_best]
	uint16_t *verify;          // device or null
	uint32_t *queue;           // overflow queue
	uint32_t *queue_count;
	uint32_t *bad_reads;       // counter of reads whose (off,len) fall outside the text buffer
	int out_zero;              // the bigtable slice is known to be all zero: do not read it
	const uint32_t *in_list;   // fragments the warp kernel has to map (punted by k_map_simple), null = all
	const uint32_t *in_count;
	JuncWords *junc;           // subjunc only: [frag][mate][multi_best] subjunc_result_t, else null
	uint16_t *big_margin;      // subjunc only: [frag][mate][27]
};

template <int MATES, int GAP, bool VERIFY, bool JUNC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, JUNC ? SR_JUNC_MIN_BLOCKS : SR_MIN_BLOCKS)
k_map_warp(const DevIndex ix, const VoteConfig cfg, const ChunkArgs a, int n_words, int warp_bytes)
{
	extern __shared__ __align__(16) uint8_t smem[];
	uint8_t *lut = smem;  // 256-entry character table shared by the block
	fill_char_table(lut, threadIdx.x, blockDim.x);
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint8_t *mine = smem + 256 + (size_t)warp * warp_bytes;
	WarpShared<MATES> &ws = *reinterpret_cast<WarpShared<MATES> *>(mine);
	uint32_t *area = reinterpret_cast<uint32_t *>(mine + ((sizeof(WarpShared<MATES>) + 15) & ~size_t(15)));
	const int mb = cfg.multi_best_reads;

	const int64_t n_todo = a.in_list ? (int64_t)*a.in_count : a.n_frags;
	for (int64_t it = (int64_t)blockIdx.x * kWarpsPerBlock + warp; it < n_todo; it += (int64_t)gridDim.x * kWarpsPerBlock) {
		const int64_t f = a.in_list ? (int64_t)a.in_list[it] : it;
		const int64_t r0 = f * MATES;
		const uint64_t o0 = a.off[r0] - a.text_origin;
		const int l0 = a.len[r0];
		uint64_t o1 = 0;
		int l1 = 0;
		// (overflow-safe: an offset below text_origin wraps to a huge o and must be refused, not added to)
		bool bad = o0 > a.text_bytes || (uint64_t)l0 > a.text_bytes - o0 || l0 > kMaxReadLen;
		if (MATES == 2) {
			o1 = a.off[r0 + 1] - a.text_origin;
			l1 = a.len[r0 + 1];
			bad = bad || o1 > a.text_bytes || (uint64_t)l1 > a.text_bytes - o1 || l1 > kMaxReadLen;
		}
		if (bad) {
			if (lane == 0) atomicAdd(a.bad_reads, 1u);
			continue;
		}
		bool ok = map_fragment<MATES, GAP, VERIFY, JUNC>(ix, cfg, ws, area, n_words, lut, a.text + o0, l0, a.text + o1, l1,
		                                         a.out + (size_t)f * MATES * mb, a.out_zero != 0,
		                                         a.verify ? a.verify + (size_t)f * MATES * mb : nullptr,
		                                         JUNC ? a.junc + (size_t)f * MATES * mb : nullptr,
		                                         JUNC ? a.big_margin + (size_t)f * MATES * kBigMarginU16 : nullptr, lane);
		if (!ok && lane == 0) a.queue[atomicAdd(a.queue_count, 1u)] = (uint32_t)f;
		__syncwarp();
	}
}

// Exact slow path: one thread per queued fragment (or per fragment when use_queue == 0).
__global__ void __launch_bounds__(kScalarThreads)
k_map_scalar(const DevIndex ix, const VoteConfig cfg, const ChunkArgs a, int use_queue, ScalarScratch *scratch)
{
	const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
	const int mates = 1 + cfg.is_paired, mb = cfg.multi_best_reads;
	const int64_t n = use_queue ? (int64_t)*a.queue_count : a.n_frags;
	for (int64_t i = tid; i < n; i += nthreads) {
		const int64_t f = use_queue ? (int64_t)a.queue[i] : i;
		const int64_t r0 = f * mates;
		const uint64_t o0 = a.off[r0] - a.text_origin;
		const int l0 = a.len[r0];
		uint64_t o1 = 0;
		int l1 = 0;
		bool bad = o0 > a.text_bytes || (uint64_t)l0 > a.text_bytes - o0 || l0 > kMaxReadLen;
		if (mates == 2) {
			o1 = a.off[r0 + 1] - a.text_origin;
			l1 = a.len[r0 + 1];
			bad = bad || o1 > a.text_bytes || (uint64_t)l1 > a.text_bytes - o1 || l1 > kMaxReadLen;
		}
		if (bad) {
			if (!use_queue) atomicAdd(a.bad_reads, 1u);
			continue;
		}
		JuncWords *jr = a.junc ? a.junc + (size_t)f * mates * mb : nullptr;
		uint16_t *bm = a.big_margin ? a.big_margin + (size_t)f * mates * kBigMarginU16 : nullptr;
		if (a.out_zero) {
			ResultWords *r = a.out + (size_t)f * mates * mb;
			for (int k = 0; k < mates * mb; k++)
				for (int w = 0; w < kResultWords; w++) r[k].w[w] = 0;
			if (jr) for (int k = 0; k < mates * mb; k++) for (int w = 0; w < 4; w++) jr[k].w[w] = 0;
			if (bm) for (int k = 0; k < mates * kBigMarginU16; k++) bm[k] = 0;
		}
		scalar_vote_fragment(ix, cfg, a.text + o0, l0, a.text + o1, l1, a.out + (size_t)f * mates * mb, jr, bm,
		                     a.verify ? a.verify + (size_t)f * mates * mb : nullptr, scratch + tid);
	}
}

// Derives the two lookup accelerators from the resident SoA index, one thread per bucket:
// a 16-bit Bloom word (bloom_bits of every key16) and the 32-byte inline record (sr_common.cuh).
__global__ void k_build_records(const uint32_t *bucket_start, const int16_t *keys, const uint32_t *vals, uint32_t nb, uint4 *records,
                                sr::bloom_t *bloom)
{
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
		const uint32_t lo = bucket_start[b], n = bucket_start[b + 1] - lo;
		if (bloom) {
			uint32_t bl = 0;
			for (uint32_t i = 0; i < n; i++) bl |= sr::bloom_bits((int)keys[lo + i]);
			bloom[b] = (sr::bloom_t)bl;
		}
		if (records) {
			uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
			w[0] = n > 0xffffu ? 0xffffu : n;
			if (n <= (uint32_t)kInlineItems) {
				uint32_t k[kInlineItems] = {kRecordNoKey, kRecordNoKey, kRecordNoKey, kRecordNoKey, kRecordNoKey}, v[kInlineItems] = {0, 0, 0, 0, 0};
				for (uint32_t i = 0; i < n; i++) { k[i] = (uint16_t)keys[lo + i]; v[i] = vals[lo + i]; }
				w[0] |= k[0] << 16; w[1] = k[1] | (k[2] << 16); w[2] = k[3] | (k[4] << 16);
				w[3] = v[0]; w[4] = v[1]; w[5] = v[2]; w[6] = v[3]; w[7] = v[4];
			} else if (n <= (uint32_t)sr::kRecordKeys) {
				w[1] = lo;
				for (uint32_t i = 0; i < (uint32_t)sr::kRecordKeys; i++)       // unused slots: the no-key, above every key (packed compares in simple_probe)
					w[2 + (i >> 1)] |= (i < n ? (uint32_t)(uint16_t)keys[lo + i] : kRecordNoKey) << ((i & 1) * 16);
			}
			records[2 * (size_t)b] = make_uint4(w[0], w[1], w[2], w[3]);
			records[2 * (size_t)b + 1] = make_uint4(w[4], w[5], w[6], w[7]);
		}
	}
}

// The phase table of a gapped index (sr_common.cuh).  (1) one thread per bucket marks every distinct key of its bucket in the
// three units it belongs to (one per role) and in the units' occupancy bitmap; (2) once all marks are in, the keys with a
// single position write it into their units' value slots, by rank among the unit's present keys.
__global__ void k_phase_mark(const uint32_t *bucket_start, const int16_t *keys, uint32_t nb, uint32_t *units, uint32_t *any)
{
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
		const uint32_t lo = bucket_start[b], n = bucket_start[b + 1] - lo;
		for (uint32_t i = 0; i < n;) {
			const int16_t k = keys[lo + i];
			uint32_t j = i + 1;
			while (j < n && keys[lo + j] == k) j++;
			const uint64_t key64 = (uint64_t)(uint16_t)k * nb + b;
			if (key64 <= 0xffffffffull) {                                      // (a key no 16-mer can produce is never looked up)
				const uint32_t key = (uint32_t)key64;
#pragma unroll
				for (int role = 0; role < 3; role++) {
					const uint32_t c = phase_unit(key, role), v = phase_variant(key, role);
					atomicOr(units + 8 * (size_t)c + role, (j - i > 1 ? 3u : 1u) << (2 * v));
					atomicOr(any + (c >> 5), 1u << (c & 31u));
				}
			}
			i = j;
		}
	}
}
__global__ void k_phase_fill(const uint32_t *bucket_start, const int16_t *keys, const uint32_t *vals, uint32_t nb, uint32_t *units)
{
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
		const uint32_t lo = bucket_start[b], n = bucket_start[b + 1] - lo;
		for (uint32_t i = 0; i < n;) {
			const int16_t k = keys[lo + i];
			uint32_t j = i + 1;
			while (j < n && keys[lo + j] == k) j++;
			const uint64_t key64 = (uint64_t)(uint16_t)k * nb + b;
			if (key64 <= 0xffffffffull && j - i == 1) {
				const uint32_t key = (uint32_t)key64;
#pragma unroll
				for (int role = 0; role < 3; role++) {
					const uint32_t c = phase_unit(key, role), v = phase_variant(key, role);
					uint32_t *u = units + 8 * (size_t)c;
					uint32_t rank = __popc(u[role] & 0x55555555u & ((1u << (2 * v)) - 1u));
					if (role > 0) rank += __popc(u[0] & 0x55555555u);
					if (role > 1) rank += __popc(u[1] & 0x55555555u);
					if (rank < (uint32_t)kPhaseValues) u[3 + rank] = vals[lo + i];
				}
			}
			i = j;
		}
	}
}

// The same for a dense table (128-byte units, no bitmap; the direct filter must exist: a key with several positions stores the
// rank of its direct_first entry).
__global__ void k_phase_mark_dense(const uint32_t *bucket_start, const int16_t *keys, uint32_t nb, uint32_t *units)
{
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
		const uint32_t lo = bucket_start[b], n = bucket_start[b + 1] - lo;
		for (uint32_t i = 0; i < n;) {
			const int16_t k = keys[lo + i];
			uint32_t j = i + 1;
			while (j < n && keys[lo + j] == k) j++;
			const uint64_t key64 = (uint64_t)(uint16_t)k * nb + b;
			if (key64 <= 0xffffffffull) {
				const uint32_t key = (uint32_t)key64;
#pragma unroll
				for (int role = 0; role < 3; role++)
					atomicOr(units + (size_t)kPhaseDenseWords * phase_unit(key, role) + role, (j - i > 1 ? 3u : 1u) << (2 * phase_variant(key, role)));
			}
			i = j;
		}
	}
}
__global__ void k_phase_fill_dense(const uint32_t *bucket_start, const int16_t *keys, const uint32_t *vals, uint32_t nb, const uint32_t *direct, uint32_t *units)
{
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
		const uint32_t lo = bucket_start[b], n = bucket_start[b + 1] - lo;
		for (uint32_t i = 0; i < n;) {
			const int16_t k = keys[lo + i];
			uint32_t j = i + 1;
			while (j < n && keys[lo + j] == k) j++;
			const uint64_t key64 = (uint64_t)(uint16_t)k * nb + b;
			if (key64 <= 0xffffffffull) {
				const uint32_t key = (uint32_t)key64;
				uint32_t value = vals[lo + i];
				if (j - i > 1) {                                             // several positions: where direct_first describes the run
					const uint32_t du = direct_unit(key);
					uint32_t bits;
					value = direct_rank(make_uint4(direct[4 * (size_t)du], direct[4 * (size_t)du + 1], direct[4 * (size_t)du + 2], direct[4 * (size_t)du + 3]), key, du, &bits);
				}
#pragma unroll
				for (int role = 0; role < 3; role++) {
					const uint32_t v = phase_variant(key, role);
					uint32_t *u = units + (size_t)kPhaseDenseWords * phase_unit(key, role);
					uint32_t rank = __popc(u[role] & 0x55555555u & ((1u << (2 * v)) - 1u));
					if (role > 0) rank += __popc(u[0] & 0x55555555u);
					if (role > 1) rank += __popc(u[1] & 0x55555555u);
					if (rank < (uint32_t)kPhaseDenseValues) u[4 + rank] = value;
				}
			}
			i = j;
		}
	}
}

// Direct filter of a dense table (sr_common.cuh), built in four steps.  (1) one thread per bucket marks every distinct
// key of its bucket: key = key16 * buckets + bucket, bit 0 present, bit 1 more than one position.
__global__ void k_direct_mark(const uint32_t *bucket_start, const int16_t *keys, uint32_t nb, uint32_t *words)
{
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
		const uint32_t lo = bucket_start[b], n = bucket_start[b + 1] - lo;
		for (uint32_t i = 0; i < n;) {
			const int16_t k = keys[lo + i];
			uint32_t j = i + 1;
			while (j < n && keys[lo + j] == k) j++;
			const uint64_t key64 = (uint64_t)(uint16_t)k * nb + b;
			if (key64 <= 0xffffffffull) {                                      // (a key no 16-mer can produce is never looked up)
				const uint32_t key = (uint32_t)key64, u = direct_unit(key), e = key - u * kDirectKeys;
				atomicOr(words + 4 * (size_t)u + 1 + (e >> 4), (j - i > 1 ? 3u : 1u) << (2 * (e & 15u)));
			}
			i = j;
		}
	}
}
// (2) present keys per unit; (3) after the exclusive scan the rank base goes into word 0 of every unit
__global__ void k_direct_count(const uint4 *units, uint32_t n_units, uint32_t *counts)
{
	for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) {
		const uint4 v = units[u];
		counts[u] = __popc(v.y & 0x55555555u) + __popc(v.z & 0x55555555u) + __popc(v.w & 0x55555555u);
	}
}
__global__ void k_direct_base(uint4 *units, uint32_t n_units, const uint32_t *bases)
{
	for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) units[u].x = bases[u];
}
// (4) one entry per distinct key, by rank: {position, 1}, or for a key with several positions {first item of the run,
// run length | landing index << 8} -- the landing index of the reference's (imin + imax) / 2 search over the bucket
// (sorted-hashtable.c:960-981), from which its visit order follows (:1109-1119)
__global__ void k_direct_fill(const uint32_t *bucket_start, const int16_t *keys, const uint32_t *vals, uint32_t nb, const uint32_t *units, uint64_t *first)
{
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
		const uint32_t lo = bucket_start[b], n = bucket_start[b + 1] - lo;
		for (uint32_t i = 0; i < n;) {
			const int16_t k = keys[lo + i];
			uint32_t j = i + 1;
			while (j < n && keys[lo + j] == k) j++;
			const uint64_t key64 = (uint64_t)(uint16_t)k * nb + b;
			if (key64 <= 0xffffffffull) {
				const uint32_t key = (uint32_t)key64, u = direct_unit(key);
				uint32_t bits;
				const uint4 unit = make_uint4(units[4 * (size_t)u], units[4 * (size_t)u + 1], units[4 * (size_t)u + 2], units[4 * (size_t)u + 3]);
				const uint32_t rank = direct_rank(unit, key, u, &bits);
				if (j - i == 1) first[rank] = (uint64_t)vals[lo + i] | (1ull << 32);
				else {
					int imin = 0, imax = (int)n - 1, at;
					for (;;) {
						at = (imin + imax) >> 1;
						if (at >= (int)j) imax = at - 1;
						else if (at < (int)i) imin = at + 1;
						else break;
					}
					const uint32_t h = j - i, land = (uint32_t)at - i;
					// runs the 8-bit fields cannot describe are marked h = 255, which no vote cluster list takes: bucket path
					const uint32_t meta = (h > 254u || land > 254u) ? 255u : (h | (land << 8));
					first[rank] = (uint64_t)(lo + i) | ((uint64_t)meta << 32);
				}
			}
			i = j;
		}
	}
}

// The bigtable slice of a batch, compacted for the trip over PCIe: of the 3 x 68 bytes per read a uniquely mapped read uses five
// words of slot 0, and the host buffer of a chunk's first index block is all zero already (core-bigtable.c:96), so only the slots
// that hold anything have to travel -- 24 bytes {slot, w0, w1, w2, w4, w15} for a slot of the plain form (one section in the indel
// record, nothing else), 72 bytes {slot, 17 words} for any other.  The order of the records does not matter (the host scatters them
// by slot).  counters[5] / [6]: records written; counters[7]: a list was too short (the batch then takes the plain copy).
__global__ void k_compact_results(const ResultWords *out, uint32_t n_slots, uint32_t *simple, uint32_t cap_simple, uint32_t *full, uint32_t cap_full,
                                  uint32_t *counters)
{
	const unsigned lane = threadIdx.x & 31u;
	for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n_slots; base += gridDim.x * blockDim.x) {
		const uint32_t slot = base + lane;
		uint32_t w[kResultWords];
		uint32_t any = 0, rest = 0;
		if (slot < n_slots) {
#pragma unroll
			for (int i = 0; i < kResultWords; i++) {
				w[i] = out[slot].w[i];
				any |= w[i];
				if (i == 3 || (i >= 5 && i <= 14) || i == 16) rest |= w[i];
			}
		}
		const unsigned ms = __ballot_sync(0xffffffffu, any != 0 && rest == 0), mf = __ballot_sync(0xffffffffu, any != 0 && rest != 0);
		uint32_t bs = 0, bf = 0;
		if (lane == 0) {
			if (ms) bs = atomicAdd(counters + 5, (uint32_t)__popc(ms));
			if (mf) bf = atomicAdd(counters + 6, (uint32_t)__popc(mf));
		}
		bs = __shfl_sync(0xffffffffu, bs, 0);
		bf = __shfl_sync(0xffffffffu, bf, 0);
		if (any != 0 && rest == 0) {
			const uint32_t at = bs + __popc(ms & ((1u << lane) - 1u));
			if (at < cap_simple) {
				uint32_t *r = simple + 6 * (size_t)at;
				r[0] = slot; r[1] = w[0]; r[2] = w[1]; r[3] = w[2]; r[4] = w[4]; r[5] = w[15];
			} else counters[7] = 1u;
		} else if (any != 0) {
			const uint32_t at = bf + __popc(mf & ((1u << lane) - 1u));
			if (at < cap_full) {
				uint32_t *r = full + 18 * (size_t)at;
				r[0] = slot;
#pragma unroll
				for (int i = 0; i < kResultWords; i++) r[1 + i] = w[i];
			} else counters[7] = 1u;
		}
	}
}

// chr_hint[b] = the first chromosome whose cumulative end offset lies above b << kChrHintShift (total when there is none):
// where locate_position may start its scan for any position of that 64 Kbp block.
__global__ void k_chr_hint(const uint32_t *offs, int total, uint32_t *hint, uint32_t n_blocks)
{
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += gridDim.x * blockDim.x) {
		const uint32_t x = b << kChrHintShift;
		int lo = 0, hi = total;                          // first n with offs[n] > x
		while (lo < hi) {
			const int mid = (lo + hi) >> 1;
			if (offs[mid] > x) hi = mid; else lo = mid + 1;
		}
		hint[b] = (uint32_t)lo;
	}
}

// Random 32-byte-sector gather: the "HBM-gather roofline" probe.
__global__ void k_gather_probe(const uint4 *table, uint64_t n_sectors, int per_thread, uint32_t seed, uint32_t *sink)
{
	uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + seed;
	uint32_t acc = 0;
	for (int i = 0; i < per_thread; i += 8) {
		uint4 v[8];
#pragma unroll
		for (int u = 0; u < 8; u++) {
			x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 29;  // cheap mixer
			uint64_t s = x % n_sectors;
			const Record8 r = ld_record(table, (uint32_t)s);   // the lookup's own load: 64 bytes per DRAM miss
			v[u] = make_uint4(r.w[0], r.w[1], r.w[6], r.w[7]);
		}
#pragma unroll
		for (int u = 0; u < 8; u++) acc ^= v[u].x ^ v[u].w;
	}
	if (acc == 0x12345678u) *sink = acc;
}

// The lookup's memory roofline for a given mix (no compute): per unit of work n_small random 2-byte probes of a small table
// that stays in L2 (the Bloom words: same load, same evict-last policy) and n_big random 32-byte records of a big table
// (the bucket records / direct-filter words / positions: the lookup's own 64-byte-fetch load), six and four in flight.
__global__ void k_mix_probe(const bloom_t *small_t, uint64_t small_n, const uint4 *big_t, uint64_t big_n, int n_small, int n_big, int units_per_thread,
                            uint32_t seed, uint32_t *sink)
{
	const uint64_t pol = policy_evict_last();
	uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + seed;
	uint32_t acc = 0;
	for (int r = 0; r < units_per_thread; r++) {
		for (int i = 0; i < n_small; i += 6) {
			uint32_t v[6];
#pragma unroll
			for (int u = 0; u < 6; u++) {
				x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 29;
				v[u] = ld_bloom(small_t + x % small_n, pol);
			}
#pragma unroll
			for (int u = 0; u < 6; u++) acc += v[u];
		}
		for (int i = 0; i < n_big; i += 4) {
			uint32_t w[4];
#pragma unroll
			for (int u = 0; u < 4; u++) {
				x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 29;
				const Record8 rec = ld_record(big_t, (uint32_t)(x % big_n));
				w[u] = rec.w[0] ^ rec.w[7];
			}
#pragma unroll
			for (int u = 0; u < 4; u++) acc += w[u];
		}
	}
	if (acc == 0x12345678u) *sink = acc;
}

}  // namespace sr

// ------------------------------------------------------------------------------------ handle

namespace {

std::mutex g_err_mutex;
std::string g_open_error;

struct DeviceBuf {
	void *p = nullptr;
	size_t cap = 0;
};

struct BatchSet {  // one of two buffer sets; used by one stream
	cudaStream_t stream = nullptr;
	DeviceBuf text, off, len, out, verify, queue, queue0, junc, big_margin;
	uint32_t *counters = nullptr;       // device: [0] queue_count (-> scalar), [1] bad_reads, [2] queue0_count (simple -> warp)
	uint32_t *h_counters = nullptr;     // pinned host mirror, 8 per batch slot
	size_t h_counters_bytes = 0;
	DeviceBuf c_simple, c_full;         // compacted results of the batch (k_compact_results)
	uint32_t *h_simple = nullptr, *h_full = nullptr;   // ... their pinned landing buffers
	size_t h_simple_cap = 0, h_full_cap = 0;           // in records
	cudaEvent_t ev_counts = nullptr, ev_records = nullptr;
	sr::ScalarScratch *scratch = nullptr;
	size_t scratch_threads = 0;
};

}  // namespace

// Buffer sets / streams a chunk's batches rotate over: each stream runs H2D -> kernels -> D2H in order, so
// with k sets up to k batches are in different stages at once (PCIe up, SMs, PCIe down all busy).
constexpr int kSets = 4;

// Host threads that scatter compacted result records into the caller's table (srcuda_vote_chunk): started once per handle and kept --
// starting fifteen threads per batch cost as much as the scatter of a small batch.  run(n_parts, fn) executes fn(0 .. n_parts-1) on
// the workers and the calling thread and returns when all parts are done.
struct ScatterPool {
	std::vector<std::thread> workers;
	std::mutex mu;
	std::condition_variable cv_job, cv_done;
	std::function<void(int)> job;
	int n_parts = 0, next_part = 0, running = 0;
	uint64_t generation = 0;
	bool quit = false;
	void worker()
	{
		uint64_t seen = 0;
		std::unique_lock<std::mutex> lk(mu);
		for (;;) {
			cv_job.wait(lk, [&] { return quit || generation != seen; });
			if (quit) return;
			seen = generation;
			while (next_part < n_parts) {
				const int t = next_part++;
				running++;
				lk.unlock();
				job(t);
				lk.lock();
				running--;
			}
			if (running == 0) cv_done.notify_all();
		}
	}
	void ensure(int n_workers)
	{
		while ((int)workers.size() < n_workers) {
			try { workers.emplace_back([this] { worker(); }); } catch (...) { break; }   // fewer workers: the caller takes the parts left
		}
	}
	void run(int parts, const std::function<void(int)> &fn)
	{
		std::unique_lock<std::mutex> lk(mu);
		job = fn; n_parts = parts; next_part = 0;
		generation++;
		cv_job.notify_all();
		while (next_part < n_parts) {            // the caller works too
			const int t = next_part++;
			running++;
			lk.unlock();
			fn(t);
			lk.lock();
			running--;
		}
		cv_done.wait(lk, [&] { return running == 0 && next_part >= n_parts; });
	}
	~ScatterPool()
	{
		{
			std::lock_guard<std::mutex> g(mu);
			quit = true;
		}
		cv_job.notify_all();
		for (auto &t : workers) t.join();
	}
};

struct srcuda_handle {
	int device = -1;
	int sm_count = 0;
	// The kernels' dynamic shared memory limit is a property of the function on the device, not of a handle: every handle raises it to
	// the device's opt-in maximum (never to the size of its own launch -- a second handle with shorter reads or a smaller pool would
	// lower it under the first one's feet); the bytes a launch asks for still decide how many blocks fit an SM.
	int smem_optin = 227 * 1024;
	mutable std::mutex mu;
	std::string err;
	bool has_index = false;
	sr::DevIndex ix{};
	srcuda_index_desc desc{};
	void *d_bucket_start = nullptr, *d_keys = nullptr, *d_vals = nullptr, *d_genome = nullptr, *d_chr = nullptr, *d_chr_hint = nullptr;
	void *d_records = nullptr, *d_bloom = nullptr, *d_direct = nullptr, *d_direct_first = nullptr, *d_phase_any = nullptr, *d_phase_units = nullptr, *d_phase_dense = nullptr;
	bool use_phase = true;         // SRCUDA_PHASE=0: no phase table for gapped sparse indexes (A/B and cross-checks)
	// The single-end tier's one-indel closed form costs an error-free chunk 13 %, so it is launched only while the chunks hold such
	// reads: after every call, from that call's counters (fragments queued + fragments it settled) -- the results are the same either way.
	// Host-buffer entry point, first index block of a chunk (the caller's bigtable is all zero): the results cross PCIe compacted
	// (k_compact_results) and expand_threads host threads scatter them into the caller's table.  On the bench box (16 host threads)
	// that scatter runs at 380 M records/s, the DMA engine writes the full table at 250-270 M reads/s (profiles/r02_host_expand.txt).
	bool compact_d2h = true;       // SRCUDA_COMPACT_D2H=0: always the plain copy
	int expand_threads = 0;        // SRCUDA_EXPAND_THREADS; default: host threads / visible GPUs
	ScatterPool *scatter = nullptr;
	int tier_indel_mode = -1;      // SRCUDA_TIER_INDEL: 0 never, 1 always, unset: by the last call
	bool tier_indel = false;
	int simple_indel_smem_set[4][5] = {}, simple_rmw_indel_smem_set[4][5] = {};
	int accel_mode = -1;  // -1 auto, 0 off, 1 bloom only, 2 records only, 3 both, 4 direct filter (SRCUDA_ACCEL)
	bool l2_persist = true;        // keep the Bloom words resident in L2 (SRCUDA_L2_PERSIST=0 turns it off)
	size_t l2_persist_max = 0, l2_window_max = 0;
	const void *l2_policy_for = nullptr;   // Bloom array the streams' access-policy window currently covers
	size_t genome_words = 0;
	BatchSet set[kSets];
	std::vector<cudaEvent_t> events;
	cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_join[kSets] = {};
	srcuda_timing timing{};
	bool force_slow = false;
	bool use_simple_tier = true;   // SRCUDA_SIMPLE_TIER=0 disables tier 0 (A/B and cross-checks)
	bool use_simple_tier_pe = true;  // SRCUDA_SIMPLE_TIER_PE=0 disables its paired-end form only
	int pe_pool_dense = 0;           // SRCUDA_PE_POOL: vote clusters of both mates of a strand pass on a dense table (A/B: shared memory per thread bounds the resident warps); 0 = kPePoolDense
	bool use_warp_junc = true;       // SRCUDA_WARP_JUNC=0: subjunc fragments all take the scalar path
	bool use_tier_junc = true;       // SRCUDA_TIER_JUNC=0: subjunc chunks skip the thread-per-fragment tier
	int simple_junc_smem_set[4][5] = {}, simple_pe_junc_smem_set[4][5] = {}, simple_rmw_smem_set[4][5] = {}, simple_pe_rmw_smem_set[4][5] = {};
	bool use_tier_rmw = true;        // SRCUDA_TIER_RMW=0: chunks merged into a non-zero bigtable (later index blocks) skip the thread-per-fragment tier
	int simple_pe_smem_set[4][5] = {};
	int simple_smem_set[4][5] = {};  // dynamic shared memory already granted to k_map_simple<GAP, MODE>
	int64_t batch_frags = 1 << 20;
	bool batch_frags_set = false;    // SRCUDA_BATCH_FRAGS given: taken as is
	bool pe_cw3 = true;              // SRCUDA_PE_CW3=0: always four-word clusters in the paired-end tier
	int simple_pe_cw3_smem_set[2] = {0, 0};
	int pe_pad_smem = 0;             // SRCUDA_PE_PAD_SMEM: extra dynamic shared memory per block of the paired-end tier (occupancy experiments)
	uint32_t *d_sink = nullptr;
};

namespace {

int fail(srcuda_handle *h, int code, const char *fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (h) {
		std::lock_guard<std::mutex> g(h->mu);
		h->err = buf;
	} else {
		std::lock_guard<std::mutex> g(g_err_mutex);
		g_open_error = buf;
	}
	return code;
}

#define CU(h, call)                                                                                        \
	do {                                                                                                   \
		cudaError_t e_ = (call);                                                                           \
		if (e_ != cudaSuccess) return fail(h, SRCUDA_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
	} while (0)

int ensure(srcuda_handle *h, DeviceBuf &b, size_t bytes)
{
	if (bytes <= b.cap) return 0;
	if (b.p) {
		CU(h, cudaDeviceSynchronize());
		CU(h, cudaFree(b.p));
		b.p = nullptr;
		b.cap = 0;
	}
	size_t want = bytes + bytes / 8 + 256;
	CU(h, cudaMalloc(&b.p, want));
	b.cap = want;
	return 0;
}

void free_index(srcuda_handle *h)
{
	cudaFree(h->d_bucket_start); cudaFree(h->d_keys); cudaFree(h->d_vals); cudaFree(h->d_genome); cudaFree(h->d_chr); cudaFree(h->d_chr_hint);
	h->d_chr_hint = nullptr;
	cudaFree(h->d_records); cudaFree(h->d_bloom); cudaFree(h->d_direct); cudaFree(h->d_direct_first); cudaFree(h->d_phase_any); cudaFree(h->d_phase_units); cudaFree(h->d_phase_dense);
	h->d_direct = h->d_direct_first = h->d_phase_any = h->d_phase_units = h->d_phase_dense = nullptr;
	h->d_bucket_start = h->d_keys = h->d_vals = h->d_genome = h->d_chr = h->d_records = h->d_bloom = nullptr;
	h->has_index = false;
	h->l2_policy_for = nullptr;
	memset(&h->ix, 0, sizeof(h->ix));
	memset(&h->desc, 0, sizeof(h->desc));
}

// The direct filter of a dense table (sr_common.cuh): mark, count, scan, fill.
int build_direct(srcuda_handle *h)
{
	const uint32_t nb = h->ix.buckets_number, nu = sr::kDirectUnits;
	const int grid = h->sm_count * 8;
	CU(h, cudaMalloc(&h->d_direct, 16 * (size_t)nu));
	CU(h, cudaMemset(h->d_direct, 0, 16 * (size_t)nu));
	sr::k_direct_mark<<<grid, 256>>>(h->ix.bucket_start, h->ix.keys, nb, (uint32_t *)h->d_direct);
	CU(h, cudaGetLastError());
	uint32_t *counts = nullptr;
	void *temp = nullptr;
	size_t temp_bytes = 0;
	CU(h, cudaMalloc((void **)&counts, sizeof(uint32_t) * (size_t)nu));
	sr::k_direct_count<<<grid, 256>>>((const uint4 *)h->d_direct, nu, counts);
	CU(h, cudaGetLastError());
	uint32_t last_count = 0, last_base = 0;
	CU(h, cudaMemcpy(&last_count, counts + (nu - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
	CU(h, cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, counts, counts, (int)nu));
	CU(h, cudaMalloc(&temp, temp_bytes));
	CU(h, cub::DeviceScan::ExclusiveSum(temp, temp_bytes, counts, counts, (int)nu));
	CU(h, cudaMemcpy(&last_base, counts + (nu - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost));
	sr::k_direct_base<<<grid, 256>>>((uint4 *)h->d_direct, nu, counts);
	CU(h, cudaGetLastError());
	CU(h, cudaDeviceSynchronize());
	cudaFree(temp);
	cudaFree(counts);
	const uint64_t n_distinct = (uint64_t)last_base + last_count;
	CU(h, cudaMalloc(&h->d_direct_first, sizeof(uint64_t) * (size_t)(n_distinct + 8)));
	CU(h, cudaMemset(h->d_direct_first, 0, sizeof(uint64_t) * (size_t)(n_distinct + 8)));
	sr::k_direct_fill<<<grid, 256>>>(h->ix.bucket_start, h->ix.keys, h->ix.vals, nb, (const uint32_t *)h->d_direct, (uint64_t *)h->d_direct_first);
	CU(h, cudaGetLastError());
	CU(h, cudaDeviceSynchronize());
	h->ix.direct = (const uint32_t *)h->d_direct;
	h->ix.direct_first = (const uint64_t *)h->d_direct_first;
	h->desc.device_bytes += 16ull * nu + 8ull * (n_distinct + 8);
	h->desc.accelerators = 4;
	return 0;
}

// Build the lookup accelerators for the resident block.  Sparse tables (mean bucket <= 4 items: the one-block
// index of a 100 Mbp genome) get the per-bucket Bloom word, which stays in L2, and the 32-byte bucket record; dense
// tables (a human-sized index, ~64 items per bucket) get the direct filter.  SRCUDA_ACCEL overrides (0 = none).
int build_accelerators(srcuda_handle *h)
{
	const uint32_t nb = h->ix.buckets_number;
	const double mean = (double)h->desc.n_items / (double)nb;
	int mode = h->accel_mode;
	if (mode < 0) mode = mean <= 4.0 ? 3 : 4;
	if (nb < sr::kRecordMinBuckets) mode &= ~2;      // the records' packed key compares need key16 < 0x7fff
	h->ix.bloom = nullptr;
	h->ix.records = nullptr;
	h->ix.direct = nullptr;
	h->ix.direct_first = nullptr;
	h->ix.phase_any = nullptr;
	h->ix.phase_units = nullptr;
	h->ix.phase_dense = nullptr;
	h->desc.accelerators = 0;
	{   // where locate_position starts (sr_common.cuh): one entry per 64 Kbp of the 32-bit coordinate space, 256 KB
		const uint32_t n_blocks = 1u << (32 - sr::kChrHintShift);
		h->ix.chr_hint = nullptr;
		cudaFree(h->d_chr_hint);
		h->d_chr_hint = nullptr;
		CU(h, cudaMalloc(&h->d_chr_hint, sizeof(uint32_t) * (size_t)n_blocks));
		sr::k_chr_hint<<<(n_blocks + 255) / 256, 256>>>(h->ix.chr_offsets, h->ix.total_offsets, (uint32_t *)h->d_chr_hint, n_blocks);
		CU(h, cudaGetLastError());
		h->ix.chr_hint = (const uint32_t *)h->d_chr_hint;
	}
	if (mode & 4) {
		int rc = build_direct(h);
		if (rc) return rc;
		// gapped index, dense table: the 128-byte phase units on top (the thread-per-fragment tiers use them; the warp kernel and
		// the keys with several positions keep the direct filter)
		if (h->ix.index_gap == 3 && h->use_phase) {
			const size_t unit_bytes = 4ull * sr::kPhaseDenseWords * sr::kPhaseUnits;
			size_t free_b = 0, total_b = 0;
			cudaMemGetInfo(&free_b, &total_b);
			if (free_b > unit_bytes + (24ull << 30)) {                       // leave room for the chunk buffers
				CU(h, cudaMalloc(&h->d_phase_dense, unit_bytes));
				CU(h, cudaMemset(h->d_phase_dense, 0, unit_bytes));
				sr::k_phase_mark_dense<<<h->sm_count * 8, 256>>>(h->ix.bucket_start, h->ix.keys, nb, (uint32_t *)h->d_phase_dense);
				CU(h, cudaGetLastError());
				sr::k_phase_fill_dense<<<h->sm_count * 8, 256>>>(h->ix.bucket_start, h->ix.keys, h->ix.vals, nb, h->ix.direct, (uint32_t *)h->d_phase_dense);
				CU(h, cudaGetLastError());
				CU(h, cudaDeviceSynchronize());
				h->ix.phase_dense = (const uint32_t *)h->d_phase_dense;
				h->desc.device_bytes += unit_bytes;
				h->desc.accelerators |= 16;
			}
		}
		return 0;
	}
	if (mode & 1) CU(h, cudaMalloc(&h->d_bloom, sizeof(sr::bloom_t) * (size_t)nb));
	if (mode & 2) CU(h, cudaMalloc(&h->d_records, 32 * (size_t)nb));
	if (mode) {
		sr::k_build_records<<<h->sm_count * 8, 256>>>(h->ix.bucket_start, h->ix.keys, h->ix.vals, nb, (uint4 *)h->d_records, (sr::bloom_t *)h->d_bloom);
		CU(h, cudaGetLastError());
		CU(h, cudaDeviceSynchronize());
	}
	h->ix.bloom = (const sr::bloom_t *)h->d_bloom;
	h->ix.records = h->d_records;
	h->desc.device_bytes += (mode & 1 ? (uint64_t)sizeof(sr::bloom_t) * nb : 0) + (mode & 2 ? 32ull * nb : 0);
	h->desc.accelerators = (uint32_t)(mode & 3);
	// gapped index, sparse table: the phase table on top (the thread-per-fragment tiers use it, the warp kernel keeps the
	// Bloom words and bucket records)
	if ((mode & 3) == 3 && h->ix.index_gap == 3 && h->use_phase) {
		const size_t unit_bytes = 32ull * sr::kPhaseUnits, any_bytes = sr::kPhaseUnits / 8;
		size_t free_b = 0, total_b = 0;
		cudaMemGetInfo(&free_b, &total_b);
		if (free_b > unit_bytes + any_bytes + (8ull << 30)) {              // leave room for the chunk buffers
			CU(h, cudaMalloc(&h->d_phase_units, unit_bytes));
			CU(h, cudaMalloc(&h->d_phase_any, any_bytes));
			CU(h, cudaMemset(h->d_phase_units, 0, unit_bytes));
			CU(h, cudaMemset(h->d_phase_any, 0, any_bytes));
			sr::k_phase_mark<<<h->sm_count * 8, 256>>>(h->ix.bucket_start, h->ix.keys, nb, (uint32_t *)h->d_phase_units, (uint32_t *)h->d_phase_any);
			CU(h, cudaGetLastError());
			sr::k_phase_fill<<<h->sm_count * 8, 256>>>(h->ix.bucket_start, h->ix.keys, h->ix.vals, nb, (uint32_t *)h->d_phase_units);
			CU(h, cudaGetLastError());
			CU(h, cudaDeviceSynchronize());
			h->ix.phase_any = (const uint32_t *)h->d_phase_any;
			h->ix.phase_units = h->d_phase_units;
			h->desc.device_bytes += unit_bytes + any_bytes;
			h->desc.accelerators |= 16;
		}
	}
	return 0;
}

int upload_index(srcuda_handle *h, int32_t index_gap, int32_t padding, uint32_t nb, uint64_t n_items, const uint32_t *bucket_start,
                 const int16_t *keys, const uint32_t *vals, uint32_t start_point, uint32_t length, const uint8_t *values,
                 uint32_t values_bytes, const uint32_t *read_offsets, int32_t total_offsets)
{
	if (!bucket_start || !keys || !vals || !values || !read_offsets || nb < 1 || total_offsets < 1)
		return fail(h, SRCUDA_ERR_ARG, "srcuda_index_upload: null array or empty table");
	if (index_gap < 1 || index_gap > 3) return fail(h, SRCUDA_ERR_UNSUPPORTED, "index gap %d not in 1..3", index_gap);
	if (nb < 2) return fail(h, SRCUDA_ERR_UNSUPPORTED, "an index needs at least 2 buckets");
	if (bucket_start[0] != 0 || bucket_start[nb] != (uint32_t)n_items) return fail(h, SRCUDA_ERR_ARG, "bucket_start does not span [0, n_items]");
	for (uint32_t b = 0; b < nb; b++)  // the kernels trust this table: refuse a corrupt one here, not with a device fault
		if (bucket_start[b] > bucket_start[b + 1]) return fail(h, SRCUDA_ERR_FORMAT, "bucket_start is not non-decreasing at bucket %u", b);
	if ((uint64_t)values_bytes < (((uint64_t)length + start_point - (start_point - start_point % 4)) >> 2))
		return fail(h, SRCUDA_ERR_FORMAT, "genome array shorter than its declared length");
	for (int32_t c = 1; c < total_offsets; c++)
		if (read_offsets[c] < read_offsets[c - 1]) return fail(h, SRCUDA_ERR_FORMAT, "chromosome offsets are not increasing");
	CU(h, cudaSetDevice(h->device));
	free_index(h);
	size_t gw = ((size_t)values_bytes + 3) / 4 + 8;  // +8 words so funnel shifts may read one word ahead
	CU(h, cudaMalloc(&h->d_bucket_start, sizeof(uint32_t) * ((size_t)nb + 1)));
	CU(h, cudaMalloc(&h->d_keys, sizeof(int16_t) * (size_t)n_items + 64));
	CU(h, cudaMalloc(&h->d_vals, sizeof(uint32_t) * (size_t)n_items + 64));
	CU(h, cudaMalloc(&h->d_genome, sizeof(uint32_t) * gw));
	CU(h, cudaMalloc(&h->d_chr, sizeof(uint32_t) * (size_t)total_offsets));
	CU(h, cudaMemcpy(h->d_bucket_start, bucket_start, sizeof(uint32_t) * ((size_t)nb + 1), cudaMemcpyHostToDevice));
	CU(h, cudaMemcpy(h->d_keys, keys, sizeof(int16_t) * (size_t)n_items, cudaMemcpyHostToDevice));
	CU(h, cudaMemcpy(h->d_vals, vals, sizeof(uint32_t) * (size_t)n_items, cudaMemcpyHostToDevice));
	CU(h, cudaMemset(h->d_genome, 0, sizeof(uint32_t) * gw));
	CU(h, cudaMemcpy(h->d_genome, values, values_bytes, cudaMemcpyHostToDevice));
	CU(h, cudaMemcpy(h->d_chr, read_offsets, sizeof(uint32_t) * (size_t)total_offsets, cudaMemcpyHostToDevice));
	h->genome_words = gw;
	sr::DevIndex &ix = h->ix;
	ix.buckets_number = nb; ix.buckets_magic = ~0ull / nb + 1; ix.index_gap = index_gap; ix.padding = padding;
	ix.start_point = start_point; ix.length = length; ix.start_base_offset = start_point - start_point % 4;
	ix.values_bytes = values_bytes; ix.total_offsets = total_offsets;
	ix.bucket_start = (const uint32_t *)h->d_bucket_start; ix.keys = (const int16_t *)h->d_keys;
	ix.vals = (const uint32_t *)h->d_vals; ix.genome = (const uint32_t *)h->d_genome; ix.chr_offsets = (const uint32_t *)h->d_chr;
	srcuda_index_desc &d = h->desc;
	d.index_gap = index_gap; d.padding = padding; d.buckets_number = nb; d.n_items = n_items;
	d.start_point = start_point; d.length = length; d.start_base_offset = ix.start_base_offset; d.values_bytes = values_bytes;
	d.total_offsets = total_offsets;
	d.device_bytes = sizeof(uint32_t) * ((uint64_t)nb + 1) + 6 * n_items + sizeof(uint32_t) * gw + sizeof(uint32_t) * (uint64_t)total_offsets;
	int rc = build_accelerators(h);
	if (rc) return rc;
	h->has_index = true;
	return SRCUDA_OK;
}

struct LaunchPlan {
	int n_words, warp_bytes, smem_bytes, grid;
};

template <int MATES, int GAP, bool VERIFY, bool JUNC>
int plan_warp_launch(srcuda_handle *h, int max_len, LaunchPlan *p)
{
	p->n_words = (max_len + 15) / 16 + 1;
	size_t wb = ((sizeof(sr::WarpShared<MATES>) + 15) & ~size_t(15)) + (size_t)MATES * sr::packed_words_per_mate(p->n_words) * 4;
	p->warp_bytes = (int)((wb + 15) & ~size_t(15));
	p->smem_bytes = 256 + p->warp_bytes * sr::kWarpsPerBlock;
	if (p->smem_bytes > 227 * 1024) return fail(h, SRCUDA_ERR_UNSUPPORTED, "shared memory plan %d B exceeds 227 KB", p->smem_bytes);
	CU(h, cudaFuncSetAttribute(sr::k_map_warp<MATES, GAP, VERIFY, JUNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
	int per_sm = 0;
	CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sr::k_map_warp<MATES, GAP, VERIFY, JUNC>, sr::kWarpsPerBlock * 32, p->smem_bytes));
	if (per_sm < 1) return fail(h, SRCUDA_ERR_CUDA, "warp kernel does not fit on an SM");
	p->grid = h->sm_count * per_sm;  // persistent grid: a whole number of resident waves
	return 0;
}

template <int MATES, int GAP, bool VERIFY, bool JUNC>
int launch_warp_v(srcuda_handle *h, cudaStream_t s, const sr::VoteConfig &cfg, const sr::ChunkArgs &a, int max_len)
{
	LaunchPlan p;
	int rc = plan_warp_launch<MATES, GAP, VERIFY, JUNC>(h, max_len, &p);
	if (rc) return rc;
	int64_t blocks_needed = (a.n_frags + sr::kWarpsPerBlock - 1) / sr::kWarpsPerBlock;
	int grid = (int)std::min<int64_t>(p.grid, std::max<int64_t>(blocks_needed, 1));
	sr::k_map_warp<MATES, GAP, VERIFY, JUNC><<<grid, sr::kWarpsPerBlock * 32, p.smem_bytes, s>>>(h->ix, cfg, a, p.n_words, p.warp_bytes);
	CU(h, cudaGetLastError());
	return 0;
}

template <int MATES, int GAP>
int launch_warp(srcuda_handle *h, cudaStream_t s, const sr::VoteConfig &cfg, const sr::ChunkArgs &a, int max_len)
{
	if (cfg.do_junctions) return launch_warp_v<MATES, GAP, false, true>(h, s, cfg, a, max_len);     // (never with verification)
	return a.verify ? launch_warp_v<MATES, GAP, true, false>(h, s, cfg, a, max_len) : launch_warp_v<MATES, GAP, false, false>(h, s, cfg, a, max_len);
}

template <int MATES>
int launch_warp_gap(srcuda_handle *h, cudaStream_t s, const sr::VoteConfig &cfg, const sr::ChunkArgs &a, int max_len)
{
	switch (h->ix.index_gap) {
	case 1: return launch_warp<MATES, 1>(h, s, cfg, a, max_len);
	case 2: return launch_warp<MATES, 2>(h, s, cfg, a, max_len);
	case 3: return launch_warp<MATES, 3>(h, s, cfg, a, max_len);
	}
	return fail(h, SRCUDA_ERR_UNSUPPORTED, "index gap %d not in 1..3", h->ix.index_gap);
}

// Enqueue the kernels of one launch on `s`.  ev[0..3] (optional) bracket warp and scalar kernels.
int launch_vote(srcuda_handle *h, cudaStream_t s, const sr::VoteConfig &cfg, const sr::ChunkArgs &a, int max_len, BatchSet &bs,
                cudaEvent_t *ev)
{
	// subjunc's minor-half / donor search (K_junc) runs in the warp kernel (sr_warp.cuh, warp_minor_half); with a
	// verification request, or SRCUDA_WARP_JUNC=0, every fragment takes the exact scalar path instead
	const bool scalar_only = h->force_slow || (cfg.do_junctions && (!h->use_warp_junc || a.verify));
	int scalar_blocks = scalar_only ? sr::kScalarBlocksPerSMAll : sr::kScalarBlocksPerSM;
	const char *sb = getenv("SRCUDA_SCALAR_BLOCKS");
	if (sb && atoi(sb) > 0 && atoi(sb) <= 32) scalar_blocks = atoi(sb);
	const int scalar_grid = h->sm_count * scalar_blocks;
	const size_t scratch_threads = (size_t)scalar_grid * sr::kScalarThreads;
	if (bs.scratch_threads < scratch_threads) {       // one private table area per resident thread
		if (bs.scratch) { CU(h, cudaDeviceSynchronize()); CU(h, cudaFree(bs.scratch)); bs.scratch = nullptr; bs.scratch_threads = 0; }
		CU(h, cudaMalloc((void **)&bs.scratch, sizeof(sr::ScalarScratch) * scratch_threads));
		bs.scratch_threads = scratch_threads;
	}
	if (ev) CU(h, cudaEventRecord(ev[0], s));
	sr::ChunkArgs aw = a;
	// tier 0: one thread per single-end fragment of a chunk whose bigtable slice is zero (sr_simple.cuh)
	const bool tier0 = !scalar_only && (!cfg.do_junctions || (h->use_tier_junc && cfg.big_margin_record_size <= 12)) && h->use_simple_tier &&
	                   !cfg.is_paired && (a.out_zero || (h->use_tier_rmw && !cfg.do_junctions)) && !a.verify && bs.queue0.p &&
	                   cfg.min_second <= cfg.min_first && cfg.top_scores >= 2 && cfg.max_vote_simples >= 2;
	if (tier0) {
		const int threads = sr::kSimpleThreads;
		const int64_t blocks = (a.n_frags + threads - 1) / threads;
		const int clusters = h->ix.direct ? sr::kSimpleClustersDense : sr::kSimpleClusters;
		const int smem = sr::simple_smem_bytes(max_len, clusters);
		uint32_t *q0 = (uint32_t *)bs.queue0.p;
#define SR_LAUNCH_SIMPLE_U(G, U)                                                                                                     \
	do {                                                                                                                             \
		if (smem > h->simple_smem_set[G][U]) {                                                                                       \
			CU(h, cudaFuncSetAttribute(sr::k_map_simple<G, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));                   \
			h->simple_smem_set[G][U] = smem;                                                                                         \
		}                                                                                                                            \
		sr::k_map_simple<G, U><<<(unsigned)blocks, threads, smem, s>>>(h->ix, cfg, a.text, a.text_origin, a.text_bytes, a.off,    \
		                                                              a.len, a.n_frags, max_len, clusters, a.out, q0,              \
		                                                              bs.counters + 2, a.bad_reads, bs.counters + 3);                               \
	} while (0)
#define SR_LAUNCH_SIMPLE_R(G, U)                                                                                                     \
	do {                                                                                                                             \
		if (smem > h->simple_rmw_smem_set[G][U]) {                                                                                   \
			CU(h, (cudaFuncSetAttribute(sr::k_map_simple<G, U, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin)));    \
			h->simple_rmw_smem_set[G][U] = smem;                                                                                     \
		}                                                                                                                            \
		sr::k_map_simple<G, U, false, true><<<(unsigned)blocks, threads, smem, s>>>(h->ix, cfg, a.text, a.text_origin,              \
		                                                                           a.text_bytes, a.off, a.len, a.n_frags, max_len,   \
		                                                                           clusters, a.out, q0, bs.counters + 2,            \
		                                                                           a.bad_reads, bs.counters + 3);                                    \
	} while (0)
#define SR_LAUNCH_SIMPLE_UI(G, U)                                                                                                    \
	do {                                                                                                                             \
		if (smem > h->simple_indel_smem_set[G][U]) {                                                                                 \
			CU(h, (cudaFuncSetAttribute(sr::k_map_simple<G, U, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin))); \
			h->simple_indel_smem_set[G][U] = smem;                                                                                   \
		}                                                                                                                            \
		sr::k_map_simple<G, U, false, false, true><<<(unsigned)blocks, threads, smem, s>>>(h->ix, cfg, a.text, a.text_origin,        \
		                                                                            a.text_bytes, a.off, a.len, a.n_frags, max_len,   \
		                                                                            clusters, a.out, q0, bs.counters + 2, a.bad_reads,\
		                                                                            bs.counters + 3, nullptr, nullptr, bs.counters + 4); \
	} while (0)
#define SR_LAUNCH_SIMPLE_RI(G, U)                                                                                                    \
	do {                                                                                                                             \
		if (smem > h->simple_rmw_indel_smem_set[G][U]) {                                                                             \
			CU(h, (cudaFuncSetAttribute(sr::k_map_simple<G, U, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin))); \
			h->simple_rmw_indel_smem_set[G][U] = smem;                                                                               \
		}                                                                                                                            \
		sr::k_map_simple<G, U, false, true, true><<<(unsigned)blocks, threads, smem, s>>>(h->ix, cfg, a.text, a.text_origin,         \
		                                                                           a.text_bytes, a.off, a.len, a.n_frags, max_len,    \
		                                                                           clusters, a.out, q0, bs.counters + 2, a.bad_reads, \
		                                                                           bs.counters + 3, nullptr, nullptr, bs.counters + 4); \
	} while (0)
#define SR_LAUNCH_SIMPLE_J(G, U)                                                                                                     \
	do {                                                                                                                             \
		const int jsmem = sr::simple_junc_smem_bytes(max_len, clusters);                                                              \
		if (jsmem > h->simple_junc_smem_set[G][U]) {                                                                                 \
			CU(h, (cudaFuncSetAttribute(sr::k_map_simple<G, U, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin)));          \
			h->simple_junc_smem_set[G][U] = jsmem;                                                                                   \
		}                                                                                                                            \
		sr::k_map_simple<G, U, true><<<(unsigned)blocks, threads, jsmem, s>>>(h->ix, cfg, a.text, a.text_origin, a.text_bytes,    \
		                                                                     a.off, a.len, a.n_frags, max_len, clusters, a.out, q0, \
		                                                                     bs.counters + 2, a.bad_reads, bs.counters + 3, a.junc, a.big_margin);  \
	} while (0)
#define SR_LAUNCH_SIMPLE(G)                                                                                                          \
	do {                                                                                                                             \
		if (cfg.do_junctions) {                                                                                                      \
			if (h->ix.direct) SR_LAUNCH_SIMPLE_J(G, 2);                                                                              \
			else SR_LAUNCH_SIMPLE_J(G, 0);                                                                                           \
		} else if (!a.out_zero) {                                                                                                    \
			if (h->ix.direct) SR_LAUNCH_SIMPLE_R(G, 2);                                                                              \
			else if (h->tier_indel) SR_LAUNCH_SIMPLE_RI(G, 0);                                                                       \
			else SR_LAUNCH_SIMPLE_R(G, 0);                                                                                           \
		} else if (h->ix.direct) SR_LAUNCH_SIMPLE_U(G, 2);                                                                           \
		else if (h->tier_indel) SR_LAUNCH_SIMPLE_UI(G, 0);                                                                           \
		else SR_LAUNCH_SIMPLE_U(G, 0);                                                                                               \
	} while (0)
		switch (h->ix.index_gap) {
		case 1: SR_LAUNCH_SIMPLE(1); break;
		case 2: SR_LAUNCH_SIMPLE(2); break;
		default:
			if (h->ix.phase_units) {                     // gapped index with a phase table: mode 3 (sparse table), mode 4 (dense)
				if (cfg.do_junctions) SR_LAUNCH_SIMPLE_J(3, 3);
				else if (!a.out_zero) { if (h->tier_indel) SR_LAUNCH_SIMPLE_RI(3, 3); else SR_LAUNCH_SIMPLE_R(3, 3); }
				else if (h->tier_indel) SR_LAUNCH_SIMPLE_UI(3, 3);
				else SR_LAUNCH_SIMPLE_U(3, 3);
			} else if (h->ix.phase_dense) {
				if (cfg.do_junctions) SR_LAUNCH_SIMPLE_J(3, 4);
				else if (!a.out_zero) SR_LAUNCH_SIMPLE_R(3, 4);
				else SR_LAUNCH_SIMPLE_U(3, 4);
			} else SR_LAUNCH_SIMPLE(3);
			break;
		}
#undef SR_LAUNCH_SIMPLE_U
#undef SR_LAUNCH_SIMPLE_UI
#undef SR_LAUNCH_SIMPLE_RI
#undef SR_LAUNCH_SIMPLE_R
#undef SR_LAUNCH_SIMPLE_J
#undef SR_LAUNCH_SIMPLE
		CU(h, cudaGetLastError());
		h->timing.n_launches++;
		aw.in_list = q0;
		aw.in_count = bs.counters + 2;
	}
	// ... and its paired-end form: one thread per fragment, pair scoring per thread (sr_simple_pe.cuh)
	const bool tier0_pe = !scalar_only && (!cfg.do_junctions || (h->use_tier_junc && cfg.big_margin_record_size <= 12)) && h->use_simple_tier &&
	                      h->use_simple_tier_pe && cfg.is_paired && (a.out_zero || (h->use_tier_rmw && !cfg.do_junctions)) && !a.verify && bs.queue0.p &&
	                      cfg.max_vote_combinations <= 3 && cfg.max_vote_simples >= 24 && cfg.multi_best_reads <= sr::kMaxBest;
	if (tier0_pe) {
		const int threads = sr::kSimpleThreads;
		const int64_t blocks = (a.n_frags + threads - 1) / threads;
		// dense phase units, reads of at most kLongRead bases (at most 32 subreads): three-word clusters, a fourth block per SM
		const bool cw3 = h->pe_cw3 && h->ix.phase_dense && h->ix.index_gap == 3 && !cfg.do_junctions && max_len <= sr::kLongRead && cfg.total_subreads <= 32;
		const int clusters = h->ix.direct ? (cfg.do_junctions ? sr::kPePoolDenseJunc : (h->pe_pool_dense ? h->pe_pool_dense : (cw3 ? sr::kPePoolDense3 : sr::kPePoolDense))) : sr::kPePool;     // one pool for both mates
		const int smem = sr::simple_pe_smem_bytes(max_len, clusters, cw3 ? 3 : sr::kSimpleClusterWords) + h->pe_pad_smem;
		uint32_t *q0 = (uint32_t *)bs.queue0.p;
#define SR_LAUNCH_PE_U(G, U)                                                                                                         \
	do {                                                                                                                             \
		if (smem > h->simple_pe_smem_set[G][U]) {                                                                                    \
			CU(h, cudaFuncSetAttribute(sr::k_map_simple_pe<G, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));                \
			h->simple_pe_smem_set[G][U] = smem;                                                                                      \
		}                                                                                                                            \
		sr::k_map_simple_pe<G, U><<<(unsigned)blocks, threads, smem, s>>>(h->ix, cfg, a.text, a.text_origin, a.text_bytes, a.off, \
		                                                                 a.len, a.n_frags, max_len, clusters, a.out, q0,           \
		                                                                 bs.counters + 2, a.bad_reads, bs.counters + 3);                            \
	} while (0)
#define SR_LAUNCH_PE_R(G, U)                                                                                                         \
	do {                                                                                                                             \
		if (smem > h->simple_pe_rmw_smem_set[G][U]) {                                                                                \
			CU(h, (cudaFuncSetAttribute(sr::k_map_simple_pe<G, U, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin))); \
			h->simple_pe_rmw_smem_set[G][U] = smem;                                                                                  \
		}                                                                                                                            \
		sr::k_map_simple_pe<G, U, false, true><<<(unsigned)blocks, threads, smem, s>>>(h->ix, cfg, a.text, a.text_origin,           \
		                                                                              a.text_bytes, a.off, a.len, a.n_frags,        \
		                                                                              max_len, clusters, a.out, q0,                 \
		                                                                              bs.counters + 2, a.bad_reads, bs.counters + 3);                \
	} while (0)
#define SR_LAUNCH_PE_3(RMW_, SET_)                                                                                                   \
	do {                                                                                                                             \
		if (smem > h->simple_pe_cw3_smem_set[SET_]) {                                                                                \
			CU(h, (cudaFuncSetAttribute(sr::k_map_simple_pe<3, 4, false, RMW_, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin))); \
			h->simple_pe_cw3_smem_set[SET_] = smem;                                                                                  \
		}                                                                                                                            \
		sr::k_map_simple_pe<3, 4, false, RMW_, 3><<<(unsigned)blocks, threads, smem, s>>>(h->ix, cfg, a.text, a.text_origin,         \
		                                                                              a.text_bytes, a.off, a.len, a.n_frags,        \
		                                                                              max_len, clusters, a.out, q0,                 \
		                                                                              bs.counters + 2, a.bad_reads, bs.counters + 3);                \
	} while (0)
#define SR_LAUNCH_PE_J(G, U)                                                                                                         \
	do {                                                                                                                             \
		if (jsmem > h->simple_pe_junc_smem_set[G][U]) {                                                                              \
			CU(h, (cudaFuncSetAttribute(sr::k_map_simple_pe<G, U, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin)));       \
			h->simple_pe_junc_smem_set[G][U] = jsmem;                                                                                \
		}                                                                                                                            \
		sr::k_map_simple_pe<G, U, true><<<(unsigned)blocks, threads, jsmem, s>>>(h->ix, cfg, a.text, a.text_origin, a.text_bytes,  \
		                                                                        a.off, a.len, a.n_frags, max_len, clusters, a.out,  \
		                                                                        q0, bs.counters + 2, a.bad_reads, bs.counters + 3, a.junc,           \
		                                                                        a.big_margin);                                      \
	} while (0)
#define SR_LAUNCH_PE(G)                                                                                                              \
	do {                                                                                                                             \
		if (cfg.do_junctions) {                                                                                                      \
			if (h->ix.direct) SR_LAUNCH_PE_J(G, 2);                                                                                  \
			else SR_LAUNCH_PE_J(G, 0);                                                                                               \
		} else if (!a.out_zero) {                                                                                                    \
			if (h->ix.direct) SR_LAUNCH_PE_R(G, 2);                                                                                  \
			else SR_LAUNCH_PE_R(G, 0);                                                                                               \
		} else if (h->ix.direct) SR_LAUNCH_PE_U(G, 2);                                                                               \
		else SR_LAUNCH_PE_U(G, 0);                                                                                                   \
	} while (0)
		const int jsmem = sr::simple_pe_junc_smem_bytes(max_len, clusters);
		if ((cfg.do_junctions ? jsmem : smem) <= 227 * 1024) {
			switch (h->ix.index_gap) {
			case 1: SR_LAUNCH_PE(1); break;
			case 2: SR_LAUNCH_PE(2); break;
			default:
				if (h->ix.phase_units) {
					if (cfg.do_junctions) SR_LAUNCH_PE_J(3, 3);
					else if (!a.out_zero) SR_LAUNCH_PE_R(3, 3);
					else SR_LAUNCH_PE_U(3, 3);
				} else if (h->ix.phase_dense) {
					if (cfg.do_junctions) SR_LAUNCH_PE_J(3, 4);
					else if (cw3) { if (!a.out_zero) SR_LAUNCH_PE_3(true, 1); else SR_LAUNCH_PE_3(false, 0); }
					else if (!a.out_zero) SR_LAUNCH_PE_R(3, 4);
					else SR_LAUNCH_PE_U(3, 4);
				} else SR_LAUNCH_PE(3);
				break;
			}
			CU(h, cudaGetLastError());
			h->timing.n_launches++;
			aw.in_list = q0;
			aw.in_count = bs.counters + 2;
		}
#undef SR_LAUNCH_PE_U
#undef SR_LAUNCH_PE_3
#undef SR_LAUNCH_PE_R
#undef SR_LAUNCH_PE_J
#undef SR_LAUNCH_PE
	}
	if (ev) CU(h, cudaEventRecord(ev[3], s));
	if (!scalar_only) {
		int rc = cfg.is_paired ? launch_warp_gap<2>(h, s, cfg, aw, max_len) : launch_warp_gap<1>(h, s, cfg, aw, max_len);
		if (rc) return rc;
		h->timing.n_launches++;
	}
	if (ev) CU(h, cudaEventRecord(ev[1], s));
	sr::k_map_scalar<<<scalar_grid, sr::kScalarThreads, 0, s>>>(h->ix, cfg, a, scalar_only ? 0 : 1, bs.scratch);
	CU(h, cudaGetLastError());
	h->timing.n_launches++;
	if (ev) CU(h, cudaEventRecord(ev[2], s));
	return 0;
}

// Best effort: the Bloom words (2 bytes per bucket, 32 MB for the one-block gapped table) are read 60 times per
// read at random, everything else streams through; a persisting access-policy window keeps them in L2.
void apply_l2_policy(srcuda_handle *h)
{
	// the table probed at random for every lookup: the phase table's occupancy bitmap when there is one, else the Bloom words
	const void *hot = h->d_phase_any ? h->d_phase_any : h->d_bloom;
	if (!h->l2_persist || !hot || h->l2_policy_for == hot || !h->l2_persist_max || !h->l2_window_max) return;
	const size_t bytes = h->d_phase_any ? (size_t)(sr::kPhaseUnits / 8) : sizeof(sr::bloom_t) * (size_t)h->ix.buckets_number;
	const size_t carve = std::min(bytes, h->l2_persist_max), window = std::min(bytes, h->l2_window_max);
	if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
	cudaStreamAttrValue v;
	memset(&v, 0, sizeof(v));
	v.accessPolicyWindow.base_ptr = const_cast<void *>(hot);
	v.accessPolicyWindow.num_bytes = window;
	v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)window);
	v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
	v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
	for (int i = 0; i < kSets; i++)
		if (h->set[i].stream && cudaStreamSetAttribute(h->set[i].stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
	h->l2_policy_for = hot;
}

int init_sets(srcuda_handle *h)
{
	for (int i = 0; i < kSets; i++) {
		BatchSet &b = h->set[i];
		if (b.stream) continue;
		CU(h, cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
		CU(h, cudaMalloc((void **)&b.counters, 8 * sizeof(uint32_t)));
	}
	if (!h->ev_begin) {
		CU(h, cudaEventCreate(&h->ev_begin));
		CU(h, cudaEventCreate(&h->ev_end));
		for (int i = 0; i < kSets; i++) CU(h, cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
	}
	apply_l2_policy(h);
	return 0;
}

// Which form of the single-end tier the next call launches (see srcuda_handle::tier_indel): the one-indel closed form once more
// than 3 % of a call's fragments were either queued for the warp kernel or settled by that form, back to the lean form below 1.5 %.
void update_tier_policy(srcuda_handle *h, const sr::VoteConfig &cfg, int64_t n_frags)
{
	if (h->tier_indel_mode >= 0 || cfg.is_paired || cfg.do_junctions || n_frags < 1000 || h->timing.ms_kernel_simple <= 0) return;
	const double f = (double)(h->timing.n_simple_punted + h->timing.n_simple_indel) / (double)n_frags;
	if (f > 0.03) h->tier_indel = true;
	else if (f < 0.015) h->tier_indel = false;
}

}  // namespace

// ------------------------------------------------------------------------------------ C ABI

extern "C" {

int srcuda_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int srcuda_open(int device, srcuda_handle **out)
{
	if (!out) return fail(nullptr, SRCUDA_ERR_ARG, "srcuda_open: out is NULL");
	*out = nullptr;
	int n = srcuda_device_count();
	if (n < 1) return fail(nullptr, SRCUDA_ERR_NO_DEVICE, "no CUDA device is visible (there is no CPU fallback)");
	if (device < 0 || device >= n) return fail(nullptr, SRCUDA_ERR_ARG, "device %d out of range 0..%d", device, n - 1);
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, SRCUDA_ERR_CUDA, "cudaGetDeviceProperties failed");
	if (prop.major < 10) return fail(nullptr, SRCUDA_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
	srcuda_handle *h = new srcuda_handle();
	h->device = device;
	h->sm_count = prop.multiProcessorCount;
	if (prop.sharedMemPerBlockOptin > 0) h->smem_optin = (int)prop.sharedMemPerBlockOptin;
	h->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
	h->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
	const char *lp = getenv("SRCUDA_L2_PERSIST");
	if (lp && lp[0] == '0') h->l2_persist = false;
	if (cudaSetDevice(device) != cudaSuccess) { delete h; return fail(nullptr, SRCUDA_ERR_CUDA, "cudaSetDevice(%d) failed", device); }
	const char *bf = getenv("SRCUDA_BATCH_FRAGS");
	if (bf && atoll(bf) > 0) { h->batch_frags = atoll(bf); h->batch_frags_set = true; }
	const char *c3 = getenv("SRCUDA_PE_CW3");
	if (c3 && c3[0] == '0') h->pe_cw3 = false;
	const char *ps = getenv("SRCUDA_PE_PAD_SMEM");
	if (ps && atoi(ps) > 0) h->pe_pad_smem = atoi(ps);
	const char *st = getenv("SRCUDA_SIMPLE_TIER");
	if (st && st[0] == '0') h->use_simple_tier = false;
	const char *ph = getenv("SRCUDA_PHASE");
	if (ph && ph[0] == '0') h->use_phase = false;
	const char *cd = getenv("SRCUDA_COMPACT_D2H");
	if (cd && cd[0] == '0') h->compact_d2h = false;
	const bool compact_forced = cd && cd[0] == '1';
	const char *et = getenv("SRCUDA_EXPAND_THREADS");
	h->expand_threads = et && atoi(et) > 0 ? atoi(et) : (int)std::max(1u, std::thread::hardware_concurrency() / (unsigned)std::max(1, n));
	if (h->expand_threads > 64) h->expand_threads = 64;
	// measured on the bench box: with 16 scatter threads the compacted copy wins (cfg2_se end to end 239 vs 221-229 M reads/s), with 8 it
	// loses (211) -- the scatter then is slower than the DMA engine writing the whole table -- so a process that shares the host with
	// seven others (one per GPU, 4 threads each) keeps the plain copy unless SRCUDA_COMPACT_D2H=1 asks for it
	if (h->expand_threads < 12 && !compact_forced) h->compact_d2h = false;
	const char *ti = getenv("SRCUDA_TIER_INDEL");
	if (ti && (ti[0] == '0' || ti[0] == '1')) { h->tier_indel_mode = ti[0] - '0'; h->tier_indel = ti[0] == '1'; }
	const char *tr = getenv("SRCUDA_TIER_RMW");
	if (tr && tr[0] == '0') h->use_tier_rmw = false;
	const char *tj = getenv("SRCUDA_TIER_JUNC");
	if (tj && tj[0] == '0') h->use_tier_junc = false;
	const char *wj = getenv("SRCUDA_WARP_JUNC");
	if (wj && wj[0] == '0') h->use_warp_junc = false;
	const char *stp = getenv("SRCUDA_SIMPLE_TIER_PE");
	if (stp && stp[0] == '0') h->use_simple_tier_pe = false;
	const char *pp = getenv("SRCUDA_PE_POOL");
	if (pp && atoi(pp) >= 4 && atoi(pp) <= 32) h->pe_pool_dense = atoi(pp);
	const char *ac = getenv("SRCUDA_ACCEL");  // 0 off, 1 bloom, 2 records, 3 both, 4 direct filter; default auto
	if (ac && ac[0] >= '0' && ac[0] <= '4') h->accel_mode = ac[0] - '0';
	const char *fg = getenv("SRCUDA_L2_FETCH");  // L2 fetch granularity hint in bytes: 32, 64 or 128
	// The hot loads are random 32-byte sectors (Bloom words, bucket records): fetching the neighbour sector too only
	// doubles the DRAM traffic.  Default 32; the streaming loads are full lines either way.
	cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, fg && atoi(fg) > 0 ? (size_t)atoi(fg) : (size_t)32);
	cudaGetLastError();
	*out = h;
	return SRCUDA_OK;
}

void srcuda_close(srcuda_handle *h)
{
	if (!h) return;
	delete h->scatter;
	h->scatter = nullptr;
	cudaSetDevice(h->device);
	cudaDeviceSynchronize();
	free_index(h);
	for (int i = 0; i < kSets; i++) {
		BatchSet &b = h->set[i];
		for (DeviceBuf *d : {&b.text, &b.off, &b.len, &b.out, &b.verify, &b.queue, &b.queue0, &b.junc, &b.big_margin}) cudaFree(d->p);
		cudaFree(b.counters);
		cudaFree(b.scratch);
		if (b.h_counters) cudaFreeHost(b.h_counters);
		cudaFree(b.c_simple.p); cudaFree(b.c_full.p);
		if (b.h_simple) cudaFreeHost(b.h_simple);
		if (b.h_full) cudaFreeHost(b.h_full);
		if (b.ev_counts) cudaEventDestroy(b.ev_counts);
		if (b.ev_records) cudaEventDestroy(b.ev_records);
		if (b.stream) cudaStreamDestroy(b.stream);
	}
	for (cudaEvent_t e : h->events) cudaEventDestroy(e);
	if (h->ev_begin) cudaEventDestroy(h->ev_begin);
	if (h->ev_end) cudaEventDestroy(h->ev_end);
	for (int i = 0; i < kSets; i++) if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
	cudaFree(h->d_sink);
	delete h;
}

const char *srcuda_last_error(const srcuda_handle *h)
{
	if (!h) {
		std::lock_guard<std::mutex> g(g_err_mutex);
		static thread_local std::string copy;
		copy = g_open_error;
		return copy.c_str();
	}
	std::lock_guard<std::mutex> g(h->mu);
	static thread_local std::string copy;
	copy = h->err;
	return copy.c_str();
}

int srcuda_index_load(srcuda_handle *h, const char *index_prefix, int block_no)
{
	if (!h || !index_prefix) return fail(h, SRCUDA_ERR_ARG, "srcuda_index_load: null argument");
	CU(h, cudaSetDevice(h->device));
	sr::HostIndex hi;
	std::string err;
	int rc = sr::load_index_block(index_prefix, block_no, &hi, &err);
	if (rc) return fail(h, rc, "%s", err.c_str());
	return upload_index(h, hi.index_gap, hi.padding, hi.buckets_number, hi.n_items, hi.bucket_start, hi.keys, hi.vals, hi.start_point,
	                    hi.length, hi.values.data(), hi.values_bytes, hi.read_offsets.data(), (int32_t)hi.read_offsets.size());
}

int srcuda_index_upload(srcuda_handle *h, int32_t index_gap, int32_t padding, uint32_t buckets_number, uint64_t n_items,
                        const uint32_t *bucket_start, const int16_t *keys, const uint32_t *vals, uint32_t start_point, uint32_t length,
                        const uint8_t *values, uint32_t values_bytes, const uint32_t *read_offsets, int32_t total_offsets)
{
	if (!h) return fail(h, SRCUDA_ERR_ARG, "srcuda_index_upload: null handle");
	return upload_index(h, index_gap, padding, buckets_number, n_items, bucket_start, keys, vals, start_point, length, values,
	                    values_bytes, read_offsets, total_offsets);
}

int srcuda_index_replicate(srcuda_handle *dst, const srcuda_handle *src)
{
	if (!dst || !src) return fail(dst, SRCUDA_ERR_ARG, "srcuda_index_replicate: null handle");
	if (!src->has_index) return fail(dst, SRCUDA_ERR_NO_INDEX, "source handle holds no index");
	CU(dst, cudaSetDevice(dst->device));
	free_index(dst);
	const srcuda_index_desc &d = src->desc;
	size_t nb1 = sizeof(uint32_t) * ((size_t)d.buckets_number + 1), nk = sizeof(int16_t) * (size_t)d.n_items,
	       nv = sizeof(uint32_t) * (size_t)d.n_items, ng = sizeof(uint32_t) * src->genome_words,
	       nc = sizeof(uint32_t) * (size_t)d.total_offsets;
	int can = 0;
	if (dst->device != src->device) {
		cudaDeviceCanAccessPeer(&can, dst->device, src->device);
		if (can) {
			cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
			cudaGetLastError();
		}
	}
	CU(dst, cudaMalloc(&dst->d_bucket_start, nb1));
	CU(dst, cudaMalloc(&dst->d_keys, nk + 64));
	CU(dst, cudaMalloc(&dst->d_vals, nv + 64));
	CU(dst, cudaMalloc(&dst->d_genome, ng));
	CU(dst, cudaMalloc(&dst->d_chr, nc));
	// cudaMemcpyPeer goes over NVLink when peer access is possible and stages through the host otherwise
	CU(dst, cudaMemcpyPeer(dst->d_bucket_start, dst->device, src->d_bucket_start, src->device, nb1));
	CU(dst, cudaMemcpyPeer(dst->d_keys, dst->device, src->d_keys, src->device, nk));
	CU(dst, cudaMemcpyPeer(dst->d_vals, dst->device, src->d_vals, src->device, nv));
	CU(dst, cudaMemcpyPeer(dst->d_genome, dst->device, src->d_genome, src->device, ng));
	CU(dst, cudaMemcpyPeer(dst->d_chr, dst->device, src->d_chr, src->device, nc));
	dst->genome_words = src->genome_words;
	dst->ix = src->ix;
	dst->ix.bucket_start = (const uint32_t *)dst->d_bucket_start; dst->ix.keys = (const int16_t *)dst->d_keys;
	dst->ix.vals = (const uint32_t *)dst->d_vals; dst->ix.genome = (const uint32_t *)dst->d_genome;
	dst->ix.chr_offsets = (const uint32_t *)dst->d_chr;
	dst->desc = src->desc;
	dst->desc.device_bytes = sizeof(uint32_t) * ((uint64_t)d.buckets_number + 1) + 6 * d.n_items + ng + nc;
	int rc = build_accelerators(dst);
	if (rc) return rc;
	dst->has_index = true;
	return SRCUDA_OK;
}

namespace {
struct IpcBlob {
	uint32_t magic;
	int32_t src_device;
	srcuda_index_desc desc;
	uint64_t genome_words;
	cudaIpcMemHandle_t mem[5];
};
static_assert(sizeof(IpcBlob) <= SRCUDA_IPC_BLOB_BYTES, "IPC blob too large");
}  // namespace

int srcuda_index_export_ipc(const srcuda_handle *h, void *blob, size_t blob_bytes)
{
	srcuda_handle *hm = const_cast<srcuda_handle *>(h);
	if (!h || !blob || blob_bytes < sizeof(IpcBlob)) return fail(hm, SRCUDA_ERR_ARG, "srcuda_index_export_ipc: bad argument");
	if (!h->has_index) return fail(hm, SRCUDA_ERR_NO_INDEX, "no index block is loaded");
	CU(hm, cudaSetDevice(h->device));
	IpcBlob b;
	memset(&b, 0, sizeof(b));
	b.magic = 0x53524950u;
	b.src_device = h->device;
	b.desc = h->desc;
	b.genome_words = h->genome_words;
	void *ptrs[5] = {h->d_bucket_start, h->d_keys, h->d_vals, h->d_genome, h->d_chr};
	for (int i = 0; i < 5; i++) CU(hm, cudaIpcGetMemHandle(&b.mem[i], ptrs[i]));
	memset(blob, 0, blob_bytes);
	memcpy(blob, &b, sizeof(b));
	return SRCUDA_OK;
}

int srcuda_index_import_ipc(srcuda_handle *h, const void *blob, size_t blob_bytes)
{
	if (!h || !blob || blob_bytes < sizeof(IpcBlob)) return fail(h, SRCUDA_ERR_ARG, "srcuda_index_import_ipc: bad argument");
	IpcBlob b;
	memcpy(&b, blob, sizeof(b));
	if (b.magic != 0x53524950u) return fail(h, SRCUDA_ERR_ARG, "srcuda_index_import_ipc: not an index blob");
	CU(h, cudaSetDevice(h->device));
	free_index(h);
	const srcuda_index_desc &d = b.desc;
	size_t bytes[5] = {sizeof(uint32_t) * ((size_t)d.buckets_number + 1), sizeof(int16_t) * (size_t)d.n_items,
	                   sizeof(uint32_t) * (size_t)d.n_items, sizeof(uint32_t) * (size_t)b.genome_words,
	                   sizeof(uint32_t) * (size_t)d.total_offsets};
	void **dst[5] = {&h->d_bucket_start, &h->d_keys, &h->d_vals, &h->d_genome, &h->d_chr};
	for (int i = 0; i < 5; i++) {
		void *remote = nullptr;
		CU(h, cudaIpcOpenMemHandle(&remote, b.mem[i], cudaIpcMemLazyEnablePeerAccess));
		cudaError_t e = cudaMalloc(dst[i], bytes[i] + 64);
		if (e == cudaSuccess) e = cudaMemcpy(*dst[i], remote, bytes[i], cudaMemcpyDefault);  // peer copy over NVLink
		cudaIpcCloseMemHandle(remote);
		if (e != cudaSuccess) return fail(h, SRCUDA_ERR_CUDA, "peer copy of index array %d failed: %s", i, cudaGetErrorString(e));
	}
	h->genome_words = (size_t)b.genome_words;
	sr::DevIndex &ix = h->ix;
	ix.buckets_number = d.buckets_number; ix.buckets_magic = ~0ull / d.buckets_number + 1; ix.index_gap = d.index_gap; ix.padding = d.padding;
	ix.start_point = d.start_point; ix.length = d.length; ix.start_base_offset = d.start_base_offset;
	ix.values_bytes = d.values_bytes; ix.total_offsets = d.total_offsets;
	ix.bucket_start = (const uint32_t *)h->d_bucket_start; ix.keys = (const int16_t *)h->d_keys;
	ix.vals = (const uint32_t *)h->d_vals; ix.genome = (const uint32_t *)h->d_genome; ix.chr_offsets = (const uint32_t *)h->d_chr;
	h->desc = d;
	h->desc.device_bytes = bytes[0] + bytes[1] + bytes[2] + bytes[3] + bytes[4];
	int rc = build_accelerators(h);
	if (rc) return rc;
	h->has_index = true;
	return SRCUDA_OK;
}

int srcuda_index_describe(const srcuda_handle *h, srcuda_index_desc *out)
{
	if (!h || !out) return SRCUDA_ERR_ARG;
	if (!h->has_index) return SRCUDA_ERR_NO_INDEX;
	*out = h->desc;
	return SRCUDA_OK;
}

void srcuda_index_unload(srcuda_handle *h)
{
	if (!h) return;
	cudaSetDevice(h->device);
	cudaDeviceSynchronize();
	free_index(h);
}

void *srcuda_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}

void srcuda_host_free(void *p)
{
	if (p) cudaFreeHost(p);
}

int srcuda_set_force_slow_path(srcuda_handle *h, int on)
{
	if (!h) return SRCUDA_ERR_ARG;
	h->force_slow = on != 0;
	return SRCUDA_OK;
}

int srcuda_get_timing(const srcuda_handle *h, srcuda_timing *out)
{
	if (!h || !out) return SRCUDA_ERR_ARG;
	*out = h->timing;
	return SRCUDA_OK;
}

int srcuda_vote_chunk_device(srcuda_handle *h, const srcuda_params *params, const char *d_text, uint64_t d_text_bytes,
                             const uint64_t *d_off, const uint16_t *d_len, int64_t n_frags, int is_paired, int max_read_len, int out_align_is_zero,
                             srcuda_mapping_result *d_out_align, srcuda_subjunc_result *d_out_junc, uint16_t *d_out_big_margin,
                             uint16_t *d_out_verify)
{
	if (!h || !params || !d_text || !d_off || !d_len || !d_out_align) return fail(h, SRCUDA_ERR_ARG, "srcuda_vote_chunk_device: null argument");
	if (params->do_breakpoint_detection && (!d_out_junc || !d_out_big_margin)) return fail(h, SRCUDA_ERR_ARG, "do_breakpoint_detection needs out_junc and out_big_margin");
	if (!h->has_index) return fail(h, SRCUDA_ERR_NO_INDEX, "no index block is loaded");
	if (n_frags < 0 || n_frags > 0x7fffffffLL) return fail(h, SRCUDA_ERR_ARG, "n_frags %lld out of range", (long long)n_frags);
	if (max_read_len < 1 || max_read_len > sr::kMaxReadLen) return fail(h, SRCUDA_ERR_ARG, "max_read_len %d not in 1..%d", max_read_len, sr::kMaxReadLen);
	CU(h, cudaSetDevice(h->device));
	sr::VoteConfig cfg;
	char msg[256];
	int rc = sr::make_vote_config(*params, is_paired, h->ix.index_gap, h->ix.start_base_offset, h->ix.length, &cfg, msg, sizeof(msg));
	if (rc) return fail(h, rc, "%s", msg);
	rc = init_sets(h);
	if (rc) return rc;
	memset(&h->timing, 0, sizeof(h->timing));
	if (n_frags == 0) return SRCUDA_OK;
	BatchSet &bs = h->set[0];
	rc = ensure(h, bs.queue, sizeof(uint32_t) * (size_t)n_frags);
	if (rc) return rc;
	rc = ensure(h, bs.queue0, sizeof(uint32_t) * (size_t)n_frags);
	if (rc) return rc;
	while (h->events.size() < 4) { cudaEvent_t e; CU(h, cudaEventCreate(&e)); h->events.push_back(e); }
	CU(h, cudaMemsetAsync(bs.counters, 0, 8 * sizeof(uint32_t), bs.stream));
	sr::ChunkArgs a;
	a.text = (const uint8_t *)d_text; a.text_origin = 0; a.text_bytes = d_text_bytes;
	a.off = d_off; a.len = d_len; a.n_frags = n_frags;
	a.out = (sr::ResultWords *)d_out_align; a.verify = d_out_verify;
	a.queue = (uint32_t *)bs.queue.p; a.queue_count = bs.counters; a.bad_reads = bs.counters + 1;
	a.out_zero = out_align_is_zero ? 1 : 0;
	a.in_list = nullptr; a.in_count = nullptr;
	a.junc = cfg.do_junctions ? (sr::JuncWords *)d_out_junc : nullptr;
	a.big_margin = cfg.do_junctions ? d_out_big_margin : nullptr;
	rc = launch_vote(h, bs.stream, cfg, a, max_read_len, bs, h->events.data());
	if (rc) return rc;
	uint32_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	CU(h, cudaMemcpyAsync(counters, bs.counters, sizeof(counters), cudaMemcpyDeviceToHost, bs.stream));
	CU(h, cudaStreamSynchronize(bs.stream));
	float ms = 0;
	CU(h, cudaEventElapsedTime(&ms, h->events[0], h->events[1]));
	h->timing.ms_kernel_vote = ms;
	CU(h, cudaEventElapsedTime(&ms, h->events[0], h->events[3]));
	h->timing.ms_kernel_simple = ms;
	h->timing.n_simple_punted = counters[2];
	h->timing.n_gather_requests = counters[3];
	h->timing.n_simple_indel = counters[4];
	update_tier_policy(h, cfg, n_frags);
	CU(h, cudaEventElapsedTime(&ms, h->events[1], h->events[2]));
	h->timing.ms_kernel_overflow = ms;
	CU(h, cudaEventElapsedTime(&ms, h->events[0], h->events[2]));
	h->timing.ms_total = ms;
	h->timing.n_overflow = (h->force_slow || (cfg.do_junctions && (!h->use_warp_junc || d_out_verify))) ? (uint64_t)n_frags : counters[0];
	h->timing.n_batches = 1;
	if (counters[1]) return fail(h, SRCUDA_ERR_ARG, "%u reads have (off,len) outside the text buffer or len > %d", counters[1], sr::kMaxReadLen);
	return SRCUDA_OK;
}

int srcuda_vote_chunk(srcuda_handle *h, const srcuda_params *params, const char *text, const uint64_t *off, const uint16_t *len,
                      int64_t n_frags, int is_paired, int out_align_is_zero, srcuda_mapping_result *out_align,
                      srcuda_subjunc_result *out_junc, uint16_t *out_big_margin, uint16_t *out_verify)
{
	if (!h || !params || !text || !off || !len || !out_align) return fail(h, SRCUDA_ERR_ARG, "srcuda_vote_chunk: null argument");
	if (params->do_breakpoint_detection && (!out_junc || !out_big_margin)) return fail(h, SRCUDA_ERR_ARG, "do_breakpoint_detection needs out_junc and out_big_margin");
	if (!h->has_index) return fail(h, SRCUDA_ERR_NO_INDEX, "no index block is loaded");
	if (n_frags < 0) return fail(h, SRCUDA_ERR_ARG, "n_frags < 0");
	CU(h, cudaSetDevice(h->device));
	sr::VoteConfig cfg;
	char msg[256];
	int rc = sr::make_vote_config(*params, is_paired, h->ix.index_gap, h->ix.start_base_offset, h->ix.length, &cfg, msg, sizeof(msg));
	if (rc) return fail(h, rc, "%s", msg);
	rc = init_sets(h);
	if (rc) return rc;
	memset(&h->timing, 0, sizeof(h->timing));
	if (n_frags == 0) return SRCUDA_OK;

	const int mates = 1 + (is_paired ? 1 : 0), mb = cfg.multi_best_reads;
	// Batches: up to batch_frags fragments (SRCUDA_BATCH_FRAGS; default 1 Mi reads), but at least eight per chunk while they stay above
	// 256 Ki reads -- the first copy up and the last results down are not overlapped with anything, so a chunk of 3.5 M pairs in four
	// batches spent a third of its time filling and draining (human_pe end to end: 177 -> 232 M reads/s with eight)
	int64_t batch = std::max<int64_t>(1, std::min<int64_t>(h->batch_frags, 0x7fffffffLL));
	if (!h->batch_frags_set) {
		batch = std::max<int64_t>(1, batch / mates);
		const int64_t eighth = ((n_frags + 7) / 8 + 65535) / 65536 * 65536, floor_frags = std::max<int64_t>(1, (256 << 10) / mates);
		batch = std::min(batch, std::max(eighth, floor_frags));
	}
	const int64_t n_batches = (n_frags + batch - 1) / batch;
	while ((int64_t)h->events.size() < 4 * n_batches) { cudaEvent_t e; CU(h, cudaEventCreate(&e)); h->events.push_back(e); }
	for (int i = 0; i < kSets; i++) {
		BatchSet &b = h->set[i];
		size_t need = sizeof(uint32_t) * 8 * (size_t)(n_batches / kSets + 2);
		if (need > b.h_counters_bytes) {                 // (kept between calls: pinning memory costs more than a chunk's kernels)
			if (b.h_counters) { cudaFreeHost(b.h_counters); b.h_counters = nullptr; b.h_counters_bytes = 0; }
			CU(h, cudaHostAlloc((void **)&b.h_counters, need, cudaHostAllocDefault));
			b.h_counters_bytes = need;
		}
		memset(b.h_counters, 0, need);
	}

	// The results cross PCIe compacted when the caller's table is all zero (see srcuda_handle::compact_d2h): a helper thread follows the
	// batches -- it waits for a batch's record counts, enqueues the copy of exactly that many records behind the batch's kernels,
	// and scatters them into the caller's table from expand_threads threads while the GPU is busy with the batches behind.
	const bool compact = h->compact_d2h && out_align_is_zero && !cfg.do_junctions && !out_verify && n_frags >= 256;
	if (compact) {
		const size_t nr_max = (size_t)std::min<int64_t>(batch, n_frags) * mates;
		const size_t cap_s = nr_max + nr_max / 4 + 1024, cap_f = nr_max / 8 + 1024;
		for (int i = 0; i < kSets && i < n_batches; i++) {
			BatchSet &b = h->set[i];
			if ((rc = ensure(h, b.c_simple, 24 * cap_s))) return rc;
			if ((rc = ensure(h, b.c_full, 72 * cap_f))) return rc;
			if (b.h_simple_cap < cap_s) {
				if (b.h_simple) cudaFreeHost(b.h_simple);
				b.h_simple = nullptr; b.h_simple_cap = 0;
				CU(h, cudaHostAlloc((void **)&b.h_simple, 24 * cap_s, cudaHostAllocDefault));
				b.h_simple_cap = cap_s;
			}
			if (b.h_full_cap < cap_f) {
				if (b.h_full) cudaFreeHost(b.h_full);
				b.h_full = nullptr; b.h_full_cap = 0;
				CU(h, cudaHostAlloc((void **)&b.h_full, 72 * cap_f, cudaHostAllocDefault));
				b.h_full_cap = cap_f;
			}
			if (!b.ev_counts) CU(h, cudaEventCreateWithFlags(&b.ev_counts, cudaEventDisableTiming));
			if (!b.ev_records) CU(h, cudaEventCreateWithFlags(&b.ev_records, cudaEventDisableTiming));
		}
	}
	struct Follow {
		std::mutex mu;
		std::condition_variable cv;
		int64_t submitted = -1, enqueued = -1;   // last batch the main thread enqueued / whose record copies the helper enqueued
		bool stop = false;
		int rc = SRCUDA_OK;
		std::string err;
		uint64_t d2h = 0;
	} follow;
	auto expand = [&](const uint32_t *simple, uint32_t n_simple, const uint32_t *full, uint32_t n_full, srcuda_mapping_result *dst) {
		const int T = std::max(1, std::min<int>(h->expand_threads, (int)((n_simple + n_full) / 4096) + 1));
		auto part = [&](int t) {
			for (uint32_t i = (uint32_t)((uint64_t)n_simple * t / T), e = (uint32_t)((uint64_t)n_simple * (t + 1) / T); i < e; i++) {
				const uint32_t *r = simple + 6 * (size_t)i;
				uint32_t *d = reinterpret_cast<uint32_t *>(dst + r[0]);
				d[0] = r[1]; d[1] = r[2]; d[2] = r[3]; d[4] = r[4]; d[15] = r[5];
			}
			for (uint32_t i = (uint32_t)((uint64_t)n_full * t / T), e = (uint32_t)((uint64_t)n_full * (t + 1) / T); i < e; i++) {
				const uint32_t *r = full + 18 * (size_t)i;
				memcpy(dst + r[0], r + 1, sizeof(srcuda_mapping_result));
			}
		};
		if (T == 1) { part(0); return; }
		if (!h->scatter) h->scatter = new ScatterPool();
		h->scatter->ensure(h->expand_threads - 1);
		h->scatter->run(T, part);
	};
	auto follower = [&]() {
		cudaSetDevice(h->device);
		for (int64_t bi = 0; bi < n_batches; bi++) {
			{
				std::unique_lock<std::mutex> lk(follow.mu);
				follow.cv.wait(lk, [&] { return follow.submitted >= bi || follow.stop; });
				if (follow.submitted < bi) return;           // the main thread gave up
			}
			BatchSet &bs = h->set[bi % kSets];
			const int64_t f0 = bi * batch, nf = std::min(n_frags, f0 + batch) - f0;
			const size_t r0 = (size_t)f0 * mates, n_slots = (size_t)nf * mates * mb;
			cudaError_t e = cudaEventSynchronize(bs.ev_counts);
			const uint32_t *c = bs.h_counters + 8 * (bi / kSets);
			const uint32_t n_simple = c[5], n_full = c[6];
			const bool plain = c[7] != 0 || n_simple > bs.h_simple_cap || n_full > bs.h_full_cap;
			if (e == cudaSuccess) {
				if (plain) e = cudaMemcpyAsync(out_align + r0 * mb, bs.out.p, sizeof(srcuda_mapping_result) * n_slots, cudaMemcpyDeviceToHost, bs.stream);
				else {
					if (n_simple) e = cudaMemcpyAsync(bs.h_simple, bs.c_simple.p, 24 * (size_t)n_simple, cudaMemcpyDeviceToHost, bs.stream);
					if (e == cudaSuccess && n_full) e = cudaMemcpyAsync(bs.h_full, bs.c_full.p, 72 * (size_t)n_full, cudaMemcpyDeviceToHost, bs.stream);
				}
			}
			if (e == cudaSuccess) e = cudaEventRecord(bs.ev_records, bs.stream);
			{
				std::lock_guard<std::mutex> g(follow.mu);
				follow.enqueued = bi;                         // the batch's device buffers may be overwritten by the one kSets further on
				if (e != cudaSuccess) { follow.rc = SRCUDA_ERR_CUDA; follow.err = cudaGetErrorString(e); follow.stop = true; }
			}
			follow.cv.notify_all();
			if (e != cudaSuccess) return;
			e = cudaEventSynchronize(bs.ev_records);
			if (e != cudaSuccess) {
				std::lock_guard<std::mutex> g(follow.mu);
				follow.rc = SRCUDA_ERR_CUDA; follow.err = cudaGetErrorString(e); follow.stop = true;
				follow.cv.notify_all();
				return;
			}
			if (plain) follow.d2h += sizeof(srcuda_mapping_result) * n_slots;
			else {
				follow.d2h += 24 * (uint64_t)n_simple + 72 * (uint64_t)n_full;
				expand(bs.h_simple, n_simple, bs.h_full, n_full, out_align + r0 * mb);
			}
		}
	};
	std::thread follow_thread;
	if (compact) {
		try { follow_thread = std::thread(follower); } catch (...) { return fail(h, SRCUDA_ERR_NOMEM, "cannot start the result thread"); }
	}
	auto stop_follower = [&]() {
		if (!follow_thread.joinable()) return;
		{
			std::lock_guard<std::mutex> g(follow.mu);
			follow.stop = true;
		}
		follow.cv.notify_all();
		follow_thread.join();
	};

	// A batch is checked right before it is enqueued (so that checking batch k + 1 overlaps the copies and kernels of batch k: the
	// scan of 10 M reads is 12 ms of host time); once a batch is in flight the caller's buffers are the source and target of
	// asynchronous copies, so a refusal further down drains the streams before it returns.
	auto check_batch = [&](int64_t r0, int64_t r1, int *max_len_out) -> int {
		int max_len = 1;
		uint64_t lo = off[r0], end = 0;
		for (int64_t r = r0; r < r1; r++) {
			max_len = std::max<int>(max_len, len[r]);
			lo = std::min(lo, off[r]);
			end = std::max(end, off[r] + len[r]);
		}
		if (max_len > sr::kMaxReadLen) return fail(h, SRCUDA_ERR_ARG, "a read is longer than %d", sr::kMaxReadLen);
		// the batch's reads must lie in text[off[first] .. off[last] + len[last])
		if (lo < off[r0] || end > off[r1 - 1] + len[r1 - 1]) return fail(h, SRCUDA_ERR_ARG, "read offsets must be non-decreasing");
		*max_len_out = max_len;
		return SRCUDA_OK;
	};
	// any failure from here on waits for the copies already enqueued before it returns
	auto drain = [&](int rc_in) {
		stop_follower();
		for (int i = 0; i < kSets; i++) if (h->set[i].stream) cudaStreamSynchronize(h->set[i].stream);
		cudaGetLastError();
		return rc_in;
	};
#define CUD(h, call)                                                                                                      \
	do {                                                                                                                  \
		cudaError_t e_ = (call);                                                                                          \
		if (e_ != cudaSuccess) return drain(fail(h, SRCUDA_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)));     \
	} while (0)

	CUD(h, cudaEventRecord(h->ev_begin, h->set[0].stream));
	for (int i = 1; i < kSets; i++) CUD(h, cudaStreamWaitEvent(h->set[i].stream, h->ev_begin, 0));
	for (int64_t bi = 0; bi < n_batches; bi++) {
		BatchSet &bs = h->set[bi % kSets];
		cudaStream_t s = bs.stream;
		const int64_t f0 = bi * batch, f1 = std::min(n_frags, f0 + batch), nf = f1 - f0;
		const int64_t r0 = f0 * mates, nr = nf * mates;
		int max_len = 1;
		if ((rc = check_batch(r0, r0 + nr, &max_len))) return drain(rc);
		if (compact && bi >= kSets) {                    // the batch that used this set before must have its record copies enqueued
			std::unique_lock<std::mutex> lk(follow.mu);
			follow.cv.wait(lk, [&] { return follow.enqueued >= bi - kSets || follow.stop; });
			if (follow.stop) { lk.unlock(); return drain(fail(h, follow.rc ? follow.rc : SRCUDA_ERR_CUDA, "copying the results failed: %s", follow.err.c_str())); }
		}
		const uint64_t origin = off[r0];
		const uint64_t text_bytes = off[r0 + nr - 1] + len[r0 + nr - 1] - origin;
		const size_t n_slots = (size_t)nr * mb;
		if ((rc = ensure(h, bs.text, text_bytes + 16))) return drain(rc);
		if ((rc = ensure(h, bs.off, sizeof(uint64_t) * (size_t)nr))) return drain(rc);
		if ((rc = ensure(h, bs.len, sizeof(uint16_t) * (size_t)nr))) return drain(rc);
		if ((rc = ensure(h, bs.out, sizeof(srcuda_mapping_result) * n_slots))) return drain(rc);
		if ((rc = ensure(h, bs.queue, sizeof(uint32_t) * (size_t)nf))) return drain(rc);
		if ((rc = ensure(h, bs.queue0, sizeof(uint32_t) * (size_t)nf))) return drain(rc);
		if (out_verify && (rc = ensure(h, bs.verify, sizeof(uint16_t) * n_slots))) return drain(rc);
		const size_t bm_bytes = sizeof(uint16_t) * sr::kBigMarginU16 * (size_t)nr;
		if (cfg.do_junctions) {
			if ((rc = ensure(h, bs.junc, sizeof(srcuda_subjunc_result) * n_slots))) return drain(rc);
			if ((rc = ensure(h, bs.big_margin, bm_bytes))) return drain(rc);
		}

		CUD(h, cudaMemcpyAsync(bs.text.p, text + origin, text_bytes, cudaMemcpyHostToDevice, s));
		CUD(h, cudaMemcpyAsync(bs.off.p, off + r0, sizeof(uint64_t) * (size_t)nr, cudaMemcpyHostToDevice, s));
		CUD(h, cudaMemcpyAsync(bs.len.p, len + r0, sizeof(uint16_t) * (size_t)nr, cudaMemcpyHostToDevice, s));
		h->timing.h2d_bytes += text_bytes + 10 * (uint64_t)nr;
		if (!out_align_is_zero) {
			CUD(h, cudaMemcpyAsync(bs.out.p, out_align + (size_t)r0 * mb, sizeof(srcuda_mapping_result) * n_slots, cudaMemcpyHostToDevice, s));
			h->timing.h2d_bytes += sizeof(srcuda_mapping_result) * n_slots;
			if (cfg.do_junctions) {
				CUD(h, cudaMemcpyAsync(bs.junc.p, out_junc + (size_t)r0 * mb, sizeof(srcuda_subjunc_result) * n_slots, cudaMemcpyHostToDevice, s));
				CUD(h, cudaMemcpyAsync(bs.big_margin.p, out_big_margin + (size_t)r0 * sr::kBigMarginU16, bm_bytes, cudaMemcpyHostToDevice, s));
				h->timing.h2d_bytes += sizeof(srcuda_subjunc_result) * n_slots + bm_bytes;
			}
		}
		CUD(h, cudaMemsetAsync(bs.counters, 0, 8 * sizeof(uint32_t), s));
		sr::ChunkArgs a;
		a.text = (const uint8_t *)bs.text.p; a.text_origin = origin; a.text_bytes = text_bytes;
		a.off = (const uint64_t *)bs.off.p; a.len = (const uint16_t *)bs.len.p; a.n_frags = nf;
		a.out = (sr::ResultWords *)bs.out.p; a.verify = out_verify ? (uint16_t *)bs.verify.p : nullptr;
		a.queue = (uint32_t *)bs.queue.p; a.queue_count = bs.counters; a.bad_reads = bs.counters + 1;
		a.out_zero = out_align_is_zero ? 1 : 0;
		a.in_list = nullptr; a.in_count = nullptr;
		a.junc = cfg.do_junctions ? (sr::JuncWords *)bs.junc.p : nullptr;
		a.big_margin = cfg.do_junctions ? (uint16_t *)bs.big_margin.p : nullptr;
		if ((rc = launch_vote(h, s, cfg, a, max_len, bs, &h->events[4 * bi]))) return drain(rc);
		if (compact) {
			sr::k_compact_results<<<h->sm_count * 8, 256, 0, s>>>((const sr::ResultWords *)bs.out.p, (uint32_t)n_slots, (uint32_t *)bs.c_simple.p,
			                                                      (uint32_t)std::min<size_t>(bs.h_simple_cap, bs.c_simple.cap / 24), (uint32_t *)bs.c_full.p,
			                                                      (uint32_t)std::min<size_t>(bs.h_full_cap, bs.c_full.cap / 72), bs.counters);
			CUD(h, cudaGetLastError());
			h->timing.n_launches++;
		} else {
			CUD(h, cudaMemcpyAsync(out_align + (size_t)r0 * mb, bs.out.p, sizeof(srcuda_mapping_result) * n_slots, cudaMemcpyDeviceToHost, s));
			h->timing.d2h_bytes += sizeof(srcuda_mapping_result) * n_slots;
		}
		if (out_verify) {
			CUD(h, cudaMemcpyAsync(out_verify + (size_t)r0 * mb, bs.verify.p, sizeof(uint16_t) * n_slots, cudaMemcpyDeviceToHost, s));
			h->timing.d2h_bytes += sizeof(uint16_t) * n_slots;
		}
		if (cfg.do_junctions) {
			CUD(h, cudaMemcpyAsync(out_junc + (size_t)r0 * mb, bs.junc.p, sizeof(srcuda_subjunc_result) * n_slots, cudaMemcpyDeviceToHost, s));
			CUD(h, cudaMemcpyAsync(out_big_margin + (size_t)r0 * sr::kBigMarginU16, bs.big_margin.p, bm_bytes, cudaMemcpyDeviceToHost, s));
			h->timing.d2h_bytes += sizeof(srcuda_subjunc_result) * n_slots + bm_bytes;
		}
		CUD(h, cudaMemcpyAsync(bs.h_counters + 8 * (bi / kSets), bs.counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
		if (compact) {
			CUD(h, cudaEventRecord(bs.ev_counts, s));
			{
				std::lock_guard<std::mutex> g(follow.mu);
				follow.submitted = bi;
			}
			follow.cv.notify_all();
		}
	}
	if (compact) {                                       // the helper enqueues the last copies and scatters the last records
		follow_thread.join();
		if (follow.rc) return drain(fail(h, follow.rc, "copying the results failed: %s", follow.err.c_str()));
		h->timing.d2h_bytes += follow.d2h;
	}
	for (int i = 1; i < kSets; i++) {  // join every stream into stream 0, then stamp the end there
		CUD(h, cudaEventRecord(h->ev_join[i], h->set[i].stream));
		CUD(h, cudaStreamWaitEvent(h->set[0].stream, h->ev_join[i], 0));
	}
	CUD(h, cudaEventRecord(h->ev_end, h->set[0].stream));
	for (int i = 0; i < kSets; i++) CUD(h, cudaStreamSynchronize(h->set[i].stream));
#undef CUD

	float ms = 0;
	CU(h, cudaEventElapsedTime(&ms, h->ev_begin, h->ev_end));
	h->timing.ms_total = ms;
	uint64_t bad = 0;
	for (int64_t bi = 0; bi < n_batches; bi++) {
		CU(h, cudaEventElapsedTime(&ms, h->events[4 * bi], h->events[4 * bi + 1]));
		h->timing.ms_kernel_vote += ms;
		CU(h, cudaEventElapsedTime(&ms, h->events[4 * bi], h->events[4 * bi + 3]));
		h->timing.ms_kernel_simple += ms;
		CU(h, cudaEventElapsedTime(&ms, h->events[4 * bi + 1], h->events[4 * bi + 2]));
		h->timing.ms_kernel_overflow += ms;
		const uint32_t *c = h->set[bi % kSets].h_counters + 8 * (bi / kSets);
		h->timing.n_simple_punted += c[2];
		h->timing.n_gather_requests += c[3];
		h->timing.n_simple_indel += c[4];
		h->timing.n_overflow += (h->force_slow || (cfg.do_junctions && (!h->use_warp_junc || out_verify))) ? (uint64_t)(std::min(n_frags, (bi + 1) * batch) - bi * batch) : c[0];
		bad += c[1];
	}
	h->timing.n_batches = (uint32_t)n_batches;
	update_tier_policy(h, cfg, n_frags);
	if (bad) return fail(h, SRCUDA_ERR_ARG, "%llu reads have (off,len) outside the text buffer", (unsigned long long)bad);
	return SRCUDA_OK;
}

int srcuda_vote_chunk_multi(srcuda_handle *const *handles, int n_handles, const srcuda_params *params, const char *text,
                            const uint64_t *off, const uint16_t *len, int64_t n_frags, int is_paired, int out_align_is_zero,
                            srcuda_mapping_result *out_align, srcuda_subjunc_result *out_junc, uint16_t *out_big_margin,
                            uint16_t *out_verify)
{
	if (!handles || n_handles < 1 || !handles[0]) return fail(nullptr, SRCUDA_ERR_ARG, "srcuda_vote_chunk_multi: no handle");
	srcuda_handle *h0 = handles[0];
	for (int k = 0; k < n_handles; k++) {
		if (!handles[k]) return fail(h0, SRCUDA_ERR_ARG, "srcuda_vote_chunk_multi: handle %d is NULL", k);
		for (int j = 0; j < k; j++)
			if (handles[j] == handles[k]) return fail(h0, SRCUDA_ERR_ARG, "srcuda_vote_chunk_multi: handle %d is listed twice", k);
		if (!handles[k]->has_index) return fail(h0, SRCUDA_ERR_NO_INDEX, "handle %d holds no index block", k);
		const srcuda_index_desc &a = h0->desc, &b = handles[k]->desc;
		if (a.n_items != b.n_items || a.buckets_number != b.buckets_number || a.length != b.length || a.start_point != b.start_point ||
		    a.index_gap != b.index_gap)
			return fail(h0, SRCUDA_ERR_ARG, "handle %d holds a different index block than handle 0", k);
	}
	if (!params || !text || !off || !len || !out_align || n_frags < 0) return fail(h0, SRCUDA_ERR_ARG, "srcuda_vote_chunk_multi: null argument");
	if (n_handles == 1)
		return srcuda_vote_chunk(h0, params, text, off, len, n_frags, is_paired, out_align_is_zero, out_align, out_junc, out_big_margin, out_verify);
	const int mates = is_paired ? 2 : 1;
	const int mb = params->multi_best_reads;
	if (mb < 1 || mb > sr::kMaxBest) return fail(h0, SRCUDA_ERR_ARG, "multi_best_reads %d not in 1..%d", mb, sr::kMaxBest);
	const int64_t base = n_frags / n_handles, extra = n_frags % n_handles;
	std::vector<int> rcs((size_t)n_handles, SRCUDA_OK);
	auto work = [&](int k) {
		const int64_t f0 = k * base + std::min<int64_t>(k, extra), nf = base + (k < extra ? 1 : 0);
		const int64_t r0 = f0 * mates;
		rcs[(size_t)k] = srcuda_vote_chunk(handles[k], params, text, off + r0, len + r0, nf, is_paired, out_align_is_zero, out_align + (size_t)r0 * mb,
		                                   out_junc ? out_junc + (size_t)r0 * mb : nullptr,
		                                   out_big_margin ? out_big_margin + (size_t)r0 * sr::kBigMarginU16 : nullptr,
		                                   out_verify ? out_verify + (size_t)r0 * mb : nullptr);
	};
	{
		std::vector<std::thread> th;
		th.reserve((size_t)n_handles);
		for (int k = 1; k < n_handles; k++) {
			bool started = false;
			try { th.emplace_back(work, k); started = true; } catch (...) {}
			if (!started) work(k);                       // no thread to be had: the slice is mapped from this one
		}
		work(0);
		for (auto &t : th) t.join();
	}
	for (int k = 0; k < n_handles; k++)
		if (rcs[(size_t)k]) {
			if (k > 0) return fail(h0, rcs[(size_t)k], "GPU %d (handle %d): %s", handles[k]->device, k, srcuda_last_error(handles[k]));
			return rcs[0];
		}
	return SRCUDA_OK;
}

int srcuda_index_load_all(srcuda_handle *const *handles, int n_handles, const char *index_prefix, int block_no)
{
	if (!handles || n_handles < 1 || !handles[0]) return fail(nullptr, SRCUDA_ERR_ARG, "srcuda_index_load_all: no handle");
	srcuda_handle *h0 = handles[0];
	for (int k = 1; k < n_handles; k++)
		if (!handles[k]) return fail(h0, SRCUDA_ERR_ARG, "srcuda_index_load_all: handle %d is NULL", k);
	int rc = srcuda_index_load(h0, index_prefix, block_no);
	if (rc) return rc;
	std::vector<int> rcs((size_t)n_handles, SRCUDA_OK);
	{
		std::vector<std::thread> th;
		for (int k = 1; k < n_handles; k++) {
			auto one = [&, k] { rcs[(size_t)k] = srcuda_index_replicate(handles[k], h0); };
			bool started = false;
			try { th.emplace_back(one); started = true; } catch (...) {}
			if (!started) one();
		}
		for (auto &t : th) t.join();
	}
	for (int k = 1; k < n_handles; k++)
		if (rcs[(size_t)k]) return fail(h0, rcs[(size_t)k], "replicating the index to GPU %d failed: %s", handles[k]->device, srcuda_last_error(handles[k]));
	return SRCUDA_OK;
}

int srcuda_measure_gather_bandwidth(srcuda_handle *h, uint64_t table_bytes, int iters, double *gbps_out)
{
	if (!h || !gbps_out || table_bytes < (1u << 20)) return fail(h, SRCUDA_ERR_ARG, "srcuda_measure_gather_bandwidth: bad argument");
	CU(h, cudaSetDevice(h->device));
	void *table = nullptr;
	CU(h, cudaMalloc(&table, table_bytes));
	CU(h, cudaMemset(table, 1, table_bytes));
	if (!h->d_sink) CU(h, cudaMalloc((void **)&h->d_sink, 4));
	const int threads = 256, blocks = h->sm_count * 8, per_thread = 256;
	const uint64_t n_sectors = table_bytes / 32;
	cudaEvent_t e0, e1;
	CU(h, cudaEventCreate(&e0));
	CU(h, cudaEventCreate(&e1));
	double best = 0;
	for (int it = 0; it < std::max(1, iters) + 1; it++) {
		CU(h, cudaEventRecord(e0, 0));
		sr::k_gather_probe<<<blocks, threads>>>((const uint4 *)table, n_sectors, per_thread, 1234u + it, h->d_sink);
		CU(h, cudaEventRecord(e1, 0));
		CU(h, cudaEventSynchronize(e1));
		float ms = 0;
		CU(h, cudaEventElapsedTime(&ms, e0, e1));
		double gbps = 32.0 * blocks * threads * per_thread / (ms * 1e-3) / 1e9;
		if (it > 0 && gbps > best) best = gbps;  // first iteration is the warm-up
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(table);
	*gbps_out = best;
	return SRCUDA_OK;
}

int srcuda_measure_lookup_roofline(srcuda_handle *h, uint64_t small_bytes, int n_small, uint64_t big_bytes, int n_big, int iters, double *units_per_s)
{
	if (!h || !units_per_s || n_small < 0 || n_big < 0 || n_small + n_big < 1 || (n_small && small_bytes < 4096) || (n_big && big_bytes < (1u << 20)))
		return fail(h, SRCUDA_ERR_ARG, "srcuda_measure_lookup_roofline: bad argument");
	CU(h, cudaSetDevice(h->device));
	void *small_t = nullptr, *big_t = nullptr;
	if (n_small) { CU(h, cudaMalloc(&small_t, small_bytes)); CU(h, cudaMemset(small_t, 1, small_bytes)); }
	if (n_big) {
		const uint64_t capped = std::min<uint64_t>(big_bytes, 0xffffffffull * 32ull);        // records are addressed by a 32-bit number
		big_bytes = capped;
		cudaError_t e = cudaMalloc(&big_t, big_bytes);
		if (e != cudaSuccess) { cudaFree(small_t); return fail(h, SRCUDA_ERR_NOMEM, "cudaMalloc of the %.1f GB probe table failed", big_bytes / 1e9); }
		CU(h, cudaMemset(big_t, 1, big_bytes));
	}
	if (!h->d_sink) CU(h, cudaMalloc((void **)&h->d_sink, 4));
	const int threads = 256, blocks = h->sm_count * 8, per_thread = std::max(1, 1024 / (n_small + n_big));
	cudaEvent_t e0, e1;
	CU(h, cudaEventCreate(&e0));
	CU(h, cudaEventCreate(&e1));
	double best = 0;
	for (int it = 0; it < std::max(1, iters) + 1; it++) {
		CU(h, cudaEventRecord(e0, 0));
		sr::k_mix_probe<<<blocks, threads>>>((const sr::bloom_t *)small_t, small_bytes / sizeof(sr::bloom_t), (const uint4 *)big_t, big_bytes / 32, n_small, n_big,
		                                     per_thread, 99u + it, h->d_sink);
		CU(h, cudaEventRecord(e1, 0));
		CU(h, cudaEventSynchronize(e1));
		float ms = 0;
		CU(h, cudaEventElapsedTime(&ms, e0, e1));
		const double rate = (double)blocks * threads * per_thread / (ms * 1e-3);
		if (it > 0 && rate > best) best = rate;          // first iteration is the warm-up
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(small_t);
	cudaFree(big_t);
	*units_per_s = best;
	return SRCUDA_OK;
}

}  // extern "C"
