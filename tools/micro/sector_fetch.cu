// tools/micro/sector_fetch.cu -- what does one random 32-byte record cost in L2 / DRAM sectors, per load flavour?
// (Design input for the bucket-record loads of the lookup kernel; not product code.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sector_fetch sector_fetch.cu
//   ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ./sector_fetch [granularity]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 29; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 32; return x; }

template <int FLAVOUR>
__global__ void k(const uint4 *big, uint64_t n_records, int per_thread, uint32_t *sink)
{
	uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 777;
	uint32_t acc = 0;
	for (int i = 0; i < per_thread; i++) {
		x = mix(x);
		const uint4 *p = big + 2 * (x % n_records);
		uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		if (FLAVOUR == 0) asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(w[0]) : "l"(p));
		if (FLAVOUR == 1)
			asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
		if (FLAVOUR == 2) { const uint4 a = __ldg(p), b = __ldg(p + 1); w[0] = a.x ^ a.w; w[1] = b.x ^ b.w; }
		if (FLAVOUR == 3) { const uint4 a = __ldcs(p), b = __ldcs(p + 1); w[0] = a.x ^ a.w; w[1] = b.x ^ b.w; }
		if (FLAVOUR == 4)
			asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
		if (FLAVOUR == 5)
			asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
		if (FLAVOUR == 6) {
			uint64_t pol;
			asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
			asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p), "l"(pol));
			asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p + 1), "l"(pol));
		}
		if (FLAVOUR == 7) {
			asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
			asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p + 1));
		}
		acc += w[0] ^ w[1] ^ w[2] ^ w[3] ^ w[4] ^ w[5] ^ w[6] ^ w[7];
	}
	if (acc == 0x12345678u) *sink = acc;
}

int main(int argc, char **argv)
{
	if (argc > 1) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1])); }
	size_t gran = 0; cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
	const uint64_t big_bytes = 512ull << 20;
	uint4 *big; cudaMalloc(&big, big_bytes); cudaMemset(big, 1, big_bytes);
	uint32_t *sink; cudaMalloc(&sink, 4);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int blocks = 148 * 8, threads = 256, per_thread = 64;
	const char *names[] = {"ld.cs.u32", "ld.L1na.L2ef.v8", "ldg uint4 x2", "ldcs uint4 x2", "ld.v8", "ld.L1na.v8", "ld.L1na.hint-ef.v4 x2", "ld.nc.L1na.v4 x2"};
	printf("L2 fetch granularity limit %zu; %llu records requested per launch\n", gran, (unsigned long long)blocks * threads * per_thread);
	for (int f = 0; f < 8; f++) {
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
			}
			cudaEventRecord(e1); cudaEventSynchronize(e1);
			float ms; cudaEventElapsedTime(&ms, e0, e1);
			if (ms < best) best = ms;
		}
		printf("flavour %d %-20s %.3f ms  %.1f G records/s\n", f, names[f], best, (double)blocks * threads * per_thread / best / 1e6);
	}
	return cudaDeviceSynchronize() != cudaSuccess;
}
