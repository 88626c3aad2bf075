// tools/micro/sector_fetch2.cu -- second round: is there any load flavour on B200 that fetches less than a full
// 128-byte line from DRAM for one random 32-byte record?  (Design input; not product code.)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 29; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 32; return x; }

template <int FLAVOUR>
__global__ void k(uint4 *big, uint64_t n_records, int per_thread, uint32_t *sink)
{
	__shared__ __align__(128) uint32_t sm[256 * 8];
	__shared__ __align__(8) uint64_t bar;
	uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 777;
	uint32_t acc = 0;
	if (FLAVOUR == 8 && threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
	}
	__syncthreads();
	uint32_t phase = 0;
	for (int i = 0; i < per_thread; i++) {
		x = mix(x);
		uint4 *p = big + 2 * (x % n_records);
		uint32_t w0 = 0;
		if (FLAVOUR == 0) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(w0) : "l"(p));
		if (FLAVOUR == 1) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(w0) : "l"(p));
		if (FLAVOUR == 2) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(w0) : "l"(p));
		if (FLAVOUR == 3) asm volatile("ld.global.L2::64B.u32 %0, [%1];" : "=r"(w0) : "l"(p));
		if (FLAVOUR == 4) asm volatile("ld.global.L2::128B.u32 %0, [%1];" : "=r"(w0) : "l"(p));
		if (FLAVOUR == 5) asm volatile("ld.global.L2::256B.u32 %0, [%1];" : "=r"(w0) : "l"(p));
		if (FLAVOUR == 6) w0 = atomicAdd(reinterpret_cast<uint32_t *>(p), 0u);
		if (FLAVOUR == 7) {
			const uint32_t d = (uint32_t)__cvta_generic_to_shared(&sm[threadIdx.x * 8]);
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(p));
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16), "l"(p + 1));
			asm volatile("cp.async.commit_group;");
			asm volatile("cp.async.wait_group 0;");
			w0 = sm[threadIdx.x * 8];
		}
		if (FLAVOUR == 8) {
			// one elected thread per warp issues 32 bulk copies of 32 bytes (addresses exchanged by shuffle)
			const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
			__syncthreads();
			if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(b), "r"(256u * 32u));
			__syncthreads();
			const uint32_t d = (uint32_t)__cvta_generic_to_shared(&sm[threadIdx.x * 8]);
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" ::"r"(d), "l"(p), "r"(b) : "memory");
			uint32_t done = 0;
			while (!done) asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared.b64 q, [%1], %2; selp.u32 %0, 1, 0, q; }" : "=r"(done) : "r"(b), "r"(phase));
			phase ^= 1;
			w0 = sm[threadIdx.x * 8];
		}
		if (FLAVOUR == 9) asm volatile("ld.global.L1::no_allocate.L2::evict_first.u32 %0, [%1];" : "=r"(w0) : "l"(p));   // should fail to assemble? kept last
		acc += w0;
	}
	if (acc == 0x12345678u) *sink = acc;
}

int main(int argc, char **argv)
{
	const uint64_t big_bytes = 512ull << 20;
	uint4 *big; cudaMalloc(&big, big_bytes); cudaMemset(big, 0, big_bytes);
	uint32_t *sink; cudaMalloc(&sink, 4);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int blocks = 148 * 8, threads = 256, per_thread = 64;
	const char *names[] = {"ld.cv", "ld.volatile", "ld.relaxed.gpu", "ld.L2::64B", "ld.L2::128B", "ld.L2::256B", "atomicAdd 0", "cp.async.cg 16x2", "cp.async.bulk 32"};
	for (int f = 0; f < 9; f++) {
		float best = 1e9;
		for (int it = 0; it < 3; it++) {
			cudaEventRecord(e0);
			switch (f) {
			case 0: k<0><<<blocks, threads>>>(big, big_bytes / 32, per_thread, sink); break;
			case 1: k<1><<<blocks, threads>>>(big, big_bytes / 32, per_thread, sink); break;
			case 2: k<2><<<blocks, threads>>>(big, big_bytes / 32, per_thread, sink); break;
			case 3: k<3><<<blocks, threads>>>(big, big_bytes / 32, per_thread, sink); break;
			case 4: k<4><<<blocks, threads>>>(big, big_bytes / 32, per_thread, sink); break;
			case 5: k<5><<<blocks, threads>>>(big, big_bytes / 32, per_thread, sink); break;
			case 6: k<6><<<blocks, threads>>>(big, big_bytes / 32, per_thread, sink); break;
			case 7: k<7><<<blocks, threads>>>(big, big_bytes / 32, per_thread, sink); break;
			case 8: k<8><<<blocks, threads>>>(big, big_bytes / 32, per_thread, sink); break;
			}
			cudaEventRecord(e1); cudaEventSynchronize(e1);
			float ms; cudaEventElapsedTime(&ms, e0, e1);
			if (ms < best) best = ms;
		}
		printf("flavour %d %-18s %.3f ms  %.1f G records/s  (%s)\n", f, names[f], best, (double)blocks * threads * per_thread / best / 1e6, cudaGetErrorString(cudaGetLastError()));
	}
	return cudaDeviceSynchronize() != cudaSuccess;
}
