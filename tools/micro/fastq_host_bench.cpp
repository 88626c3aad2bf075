// tools/micro/fastq_host_bench.cpp -- host-side rates of the input side (row f3), no GPU needed:
//   fastq_host_bench FILE [window_MB]        srcuda_fastq_split_mt over the whole file in memory, then the chunk reader
//   fastq_host_bench --make FILE N_READS     writes N_READS records of 150 bp ("@read%07d")
// Build:  g++ -O2 -std=c++17 -o tools/micro/fastq_host_bench tools/micro/fastq_host_bench.cpp -I. \
//             -Lsubread_b200 -lsubread_cuda -Wl,-rpath,$PWD/subread_b200 -pthread
// The reference's reader on the same file: oracle/ref_reader_bench.c.  Results: profiles/r01_fastq_host_container.txt.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "include/subread_cuda.h"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
	if (argc >= 4 && !strcmp(argv[1], "--make")) {
		FILE *f = fopen(argv[2], "wb");
		if (!f) return 1;
		const std::string seq(150, 'A'), qual(150, 'H');
		for (long i = 0; i < atol(argv[3]); i++) fprintf(f, "@read%07ld\n%s\n+\n%s\n", i, seq.c_str(), qual.c_str());
		fclose(f);
		return 0;
	}
	if (argc < 2) return 1;
	FILE *f = fopen(argv[1], "rb");
	if (!f) return 1;
	fseek(f, 0, SEEK_END);
	const size_t n = (size_t)ftell(f);
	fseek(f, 0, SEEK_SET);
	std::vector<char> buf(n);
	if (fread(buf.data(), 1, n, f) != n) return 1;
	fclose(f);
	const int64_t cap = std::max<int64_t>((int64_t)(n / 32 + 1024), 262144);   // also the reader's batch below
	std::vector<uint64_t> off((size_t)cap);
	std::vector<uint16_t> len((size_t)cap);
	const bool gz = n > 2 && (unsigned char)buf[0] == 31 && (unsigned char)buf[1] == 139;      // gzip: the reader only
	for (int T : {1, 2, 4, 8, 16}) {
		if (gz) break;
		double best = 1e9;
		srcuda_fastq_batch b;
		for (int rep = 0; rep < 7; rep++) {
			const double t0 = now();
			if (srcuda_fastq_split_mt(buf.data(), n, 1, 0, 0, cap, 0, 1, off.data(), len.data(), 0, 0, 0, 0, T, &b)) return 2;
			best = std::min(best, now() - t0);
		}
		printf("split   T=%-2d %9ld reads %7.1f M reads/s %6.2f GB/s\n", T, (long)b.n_reads, b.n_reads / best / 1e6, n / best / 1e9);
	}
	const uint64_t window = (uint64_t)(argc > 2 ? atoi(argv[2]) : 16) << 20;
	const int64_t batch = 262144;
	std::vector<char> text((size_t)batch * 1210);
	for (int T : {1, 2, 4, 8}) {
		double best = 1e9;
		long total = 0;
		for (int rep = 0; rep < 4; rep++) {
			const double t0 = now();
			srcuda_reader *r = nullptr;
			if (srcuda_reader_open(argv[1], nullptr, window, T, 0, 0, &r)) return 3;
			int64_t k = 0;
			total = 0;
			do {
				uint64_t used;
				if (srcuda_reader_next(r, batch, text.data(), text.size(), off.data(), len.data(), &k, &used)) return 4;
				total += k;
			} while (k);
			srcuda_reader_close(r);
			best = std::min(best, now() - t0);
		}
		printf("reader  T=%-2d %9ld reads %7.1f M reads/s (window %llu MB, open to end of file)\n", T, total, total / best / 1e6,
		       (unsigned long long)(window >> 20));
	}
	return 0;
}
