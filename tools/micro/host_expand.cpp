// tools/micro/host_expand.cpp -- would a compacted device->host copy of the bigtable pay?  (Design input, not product code.)
// The voting step's result is the reference's bigtable: 3 x 68 bytes per read, of which a uniquely mapped read uses 5 words of slot 0.
// A compacted copy (24 bytes per used slot) would cut the PCIe bytes by 8x -- but the caller's buffer must still end up in the
// reference's layout, so host threads would have to scatter the records into it.  This measures that scatter alone: N records of
// 24 bytes into an N x 204-byte table, T threads, with and without zeroing the table first.  Compare with the DMA engine writing
// the full 2.04 GB of 10 M reads in 37-40 ms (250-270 M reads/s, profiles/r01_d2h_sparse.txt).
//   g++ -O2 -pthread -o host_expand host_expand.cpp && ./host_expand 10000000 16 0
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
struct Rec { uint32_t slot, pos, flags, votes_used, subs, cov; };
int main(int argc, char **argv)
{
	const size_t n = argc > 1 ? atoll(argv[1]) : 10000000;
	const int T = argc > 2 ? atoi(argv[2]) : 8;
	const int zero_first = argc > 3 ? atoi(argv[3]) : 0;
	std::vector<Rec> rec(n);
	for (size_t i = 0; i < n; i++) rec[i] = {(uint32_t)(3 * i), (uint32_t)i * 7u, 8u, 10u | (10u << 16), 1u | (10u << 16), 0u | (150u << 16)};
	uint8_t *big = (uint8_t *)aligned_alloc(4096, n * 204);
	memset(big, 0, n * 204);
	for (int rep = 0; rep < 4; rep++) {
		auto t0 = std::chrono::steady_clock::now();
		std::vector<std::thread> th;
		for (int t = 0; t < T; t++)
			th.emplace_back([&, t] {
				const size_t lo = n * t / T, hi = n * (t + 1) / T;
				if (zero_first) memset(big + lo * 204, 0, (hi - lo) * 204);
				for (size_t i = lo; i < hi; i++) {
					const Rec &r = rec[i];
					uint32_t *d = (uint32_t *)(big + (size_t)r.slot * 68);
					d[0] = r.pos; d[1] = r.flags; d[2] = r.votes_used; d[4] = r.subs; d[15] = r.cov;
				}
			});
		for (auto &x : th) x.join();
		double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		printf("T=%d zero_first=%d: %.1f ms -> %.0f M records/s\n", T, zero_first, s * 1e3, n / s / 1e6);
	}
	return 0;
}
