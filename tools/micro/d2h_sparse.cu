// tools/micro/d2h_sparse.cu -- how to bring a mostly-zero bigtable (3 x 68-byte slots per read, usually only slot 0 used)
// back to pinned host memory: full copy vs 2-D strided copy vs a kernel that stores only the non-zero slots through the
// mapped host pointer.  (Design input for srcuda_vote_chunk's D2H leg; not product code.)
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

constexpr int kW = 17;   // words per slot

__global__ void k_fill(uint32_t *d, size_t n_reads)
{
	for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < n_reads; r += (size_t)gridDim.x * blockDim.x) {
		uint32_t *s = d + r * 3 * kW;
		for (int w = 0; w < 3 * kW; w++) s[w] = 0;
		s[0] = (uint32_t)r * 2654435761u | 1u; s[1] = (r & 1) ? 8u : 0u; s[2] = 10u | (10u << 16); s[4] = 1u | (10u << 16); s[15] = 134u << 16;
		if (r % 50 == 0) { s[kW] = 77u; s[kW + 2] = 3u; }
	}
}

// warp per 32 slots: lane l tests slot l, then the 17 word rows are stored coalesced for the non-zero slots only
__global__ void k_scatter(const uint32_t *__restrict__ d, uint32_t *__restrict__ h, size_t n_slots)
{
	const int lane = threadIdx.x & 31;
	const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
	for (size_t base = warp * 32; base < n_slots; base += n_warps * 32) {
		const size_t s = base + lane;
		uint32_t any = 0;
		if (s < n_slots) {
			const uint32_t *p = d + s * kW;
#pragma unroll
			for (int w = 0; w < kW; w++) any |= p[w];
		}
		const unsigned nz = __ballot_sync(0xffffffffu, any != 0);
		if (!nz) continue;
		const size_t w0 = base * kW;
		const int n_words = (int)(min((size_t)32, n_slots - base) * kW);
		for (int i = lane; i < n_words; i += 32)
			if ((nz >> (i / kW)) & 1u) h[w0 + i] = d[w0 + i];
	}
}

int main()
{
	const size_t n_reads = 10000000, n_slots = 3 * n_reads, bytes = n_slots * kW * 4;
	uint32_t *d, *h;
	cudaMalloc(&d, bytes);
	cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
	memset(h, 0, bytes);
	k_fill<<<148 * 8, 256>>>(d, n_reads);
	cudaDeviceSynchronize();
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaStream_t s; cudaStreamCreate(&s);
	auto time = [&](const char *name, auto fn, double payload) {
		float best = 1e9;
		for (int it = 0; it < 3; it++) {
			cudaEventRecord(e0, s); fn(); cudaEventRecord(e1, s); cudaEventSynchronize(e1);
			float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
		}
		printf("%-44s %8.2f ms  %6.1f GB/s payload  (%s)\n", name, best, payload / best / 1e6, cudaGetErrorString(cudaGetLastError()));
	};
	time("full contiguous copy (2.04 GB)", [&] { cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s); }, (double)bytes);
	time("2-D copy, slot 0 only (68 of 204 B rows)", [&] { cudaMemcpy2DAsync(h, 204, d, 204, 68, n_reads, cudaMemcpyDeviceToHost, s); }, 68.0 * n_reads);
	uint32_t *hd = nullptr;
	cudaHostGetDevicePointer(&hd, h, 0);
	for (int blocks : {16, 32, 74, 148, 296, 592})
		for (int threads : {256}) {
			char nm[96]; snprintf(nm, sizeof nm, "kernel scatter of non-zero slots, %d x %d", blocks, threads);
			memset(h, 0, 1 << 20);
			time(nm, [&] { k_scatter<<<blocks, threads, 0, s>>>(d, hd, n_slots); }, 68.0 * n_reads * 1.02);
		}
	// verify the last variant
	uint32_t *chk = (uint32_t *)malloc(bytes);
	cudaMemcpy(chk, d, bytes, cudaMemcpyDeviceToHost);
	printf("scatter result %s\n", memcmp(chk, h, bytes) == 0 ? "matches" : "DIFFERS");
	// both directions at once: H2D of 1.6 GB on another stream while the scatter runs
	uint8_t *hin, *din; cudaHostAlloc(&hin, 1600000000, cudaHostAllocDefault); cudaMalloc(&din, 1600000000);
	cudaStream_t s2; cudaStreamCreate(&s2);
	for (int mode = 0; mode < 2; mode++) {
		cudaDeviceSynchronize();
		cudaEventRecord(e0, s);
		cudaMemcpyAsync(din, hin, 1600000000, cudaMemcpyHostToDevice, s2);
		if (mode == 0) cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s);
		else k_scatter<<<148, 256, 0, s>>>(d, hd, n_slots);
		cudaStreamSynchronize(s2); cudaStreamSynchronize(s);
		cudaEventRecord(e1, s); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		printf("H2D 1.6 GB concurrent with %s: %.2f ms\n", mode == 0 ? "full D2H copy" : "kernel scatter", ms);
	}
	return 0;
}
