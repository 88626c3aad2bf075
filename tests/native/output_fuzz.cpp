// tests/native/output_fuzz.cpp -- the SAM / BAM writer (sr_output.cpp) under AddressSanitizer / UBSan or ThreadSanitizer:
// valid SAM lines with random damage (bytes replaced, removed, inserted; columns dropped; optional columns of every type
// cut anywhere) through srcuda_out_add_sam_lines, and fragments through srcuda_out_format_fragment / srcuda_out_append from
// several threads at once.  A damaged line may be refused (SRCUDA_ERR_FORMAT / _ARG); nothing may crash, read or write out of
// bounds, or race.  Built and run by tests/test_output_writer.py::test_sanitizer_fuzz.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>
#include "../../include/subread_cuda.h"
#include "../../include/subread_output.h"

static srcuda_writer *open_writer(int format, int threads) {
	srcuda_writer *w = nullptr;
	if (srcuda_out_open("/dev/null", format, threads, 0, &w)) abort();
	srcuda_out_add_header(w, "@HD\tVN:1.0", 0);
	srcuda_out_add_chromosome(w, "chr1", 1000000, 1);
	srcuda_out_add_chromosome(w, "chr2", 2000000, 1);
	return w;
}

int main() {
	std::mt19937_64 rng(4242);
	const char *columns[] = {"NM:i:3", "HI:i:1", "XS:A:+", "RG:Z:group", "XF:f:1.5", "XB:B:i,1,2,3,-4", "XG:B:f,0.5,2", "XH:H:1AE3", "XZ:Z:", "XQ:q:1", "X"};
	const char junk[] = "\t\t:,*=ACGTNIMD0123456789\n\x01\x7f";
	long refused = 0, taken = 0;
	for (int format = 0; format < 2; format++) {
		srcuda_writer *w = open_writer(format, 3);
		for (int it = 0; it < 20000; it++) {
			int L = 1 + (int)(rng() % 180);
			std::string seq, qual;
			for (int k = 0; k < L; k++) {
				seq += "ACGTN"[rng() % 5];
				qual += (char)(33 + rng() % 60);
			}
			std::string line = "read" + std::to_string(it) + "\t" + std::to_string(rng() % 4096) + "\t" + (rng() % 5 ? "chr1" : "*") + "\t" +
			                   std::to_string(rng() % 999999) + "\t" + std::to_string(rng() % 70) + "\t" + std::to_string(L) + "M\t" +
			                   (rng() % 2 ? "=" : "chr2") + "\t" + std::to_string(rng() % 999999) + "\t" + std::to_string((int)(rng() % 2000) - 1000) + "\t" + seq +
			                   "\t" + qual;
			int ncol = (int)(rng() % 6);
			for (int c = 0; c < ncol; c++) line += std::string("\t") + columns[rng() % (sizeof columns / sizeof *columns)];
			int muts = (int)(rng() % 4);
			for (int m = 0; m < muts && !line.empty(); m++) {
				size_t at = rng() % line.size();
				switch (rng() % 4) {
				case 0: line[at] = junk[rng() % (sizeof junk - 1)]; break;
				case 1: line.erase(at, 1 + rng() % 8); break;
				case 2: line.insert(at, 1, junk[rng() % (sizeof junk - 1)]); break;
				default: line.resize(at); break;
				}
			}
			for (char &c : line) if (c == '\n' || c == 0) c = ' ';
			line += "\n";
			// exact-size heap copy so that the sanitizer sees any read past the end
			char *buf = (char *)malloc(line.size());
			memcpy(buf, line.data(), line.size());
			int rc = srcuda_out_add_sam_lines(w, buf, line.size(), 1);
			free(buf);
			if (rc) {
				refused++;
				if (rc != SRCUDA_ERR_FORMAT && rc != SRCUDA_ERR_ARG) { fprintf(stderr, "unexpected code %d: %s\n", rc, srcuda_out_error(w)); return 1; }
				srcuda_out_close(w, nullptr);
				w = open_writer(format, 3);
			} else {
				taken++;
			}
		}
		if (srcuda_out_close(w, nullptr)) return 2;
	}
	// fragments built on four threads at once, appended under a lock-free hand-over (one appender at a time by construction here)
	for (int format = 0; format < 2; format++) {
		srcuda_writer *w = open_writer(format, 4);
		if (srcuda_out_finish_header(w)) return 3;
		srcuda_sam_config cfg{1, 50, 600, 0, 1, 0, "grp"};
		const int per_thread = 3000, n_threads = 4;
		std::vector<std::vector<std::string>> out(n_threads);
		std::vector<std::thread> pool;
		for (int t = 0; t < n_threads; t++)
			pool.emplace_back([&, t]() {
				std::mt19937_64 r(77 + t);
				char buf[8192];
				for (int i = 0; i < per_thread; i++) {
					int L = 20 + (int)(r() % 130);
					std::string t1, t2, q1, q2, n1 = "frag" + std::to_string(t) + "_" + std::to_string(i) + "/1", n2 = n1;
					for (int k = 0; k < L; k++) { t1 += "ACGTNacgu"[r() % 9]; t2 += "ACGT"[r() % 4]; q1 += (char)(33 + r() % 40); q2 += (char)(33 + r() % 40); }
					if (r() % 20 == 0) { q1.clear(); q2.clear(); q1.resize(L + 1, 0); q2.resize(L + 1, 0); }
					std::string c1 = std::to_string(L) + "M", c2 = "5S" + std::to_string(L - 5) + "M";
					srcuda_sam_mate m1{1, (int)(r() % 4 != 0), (int)(r() % 2), (int)(r() % 900000), (int)(r() % 2), 30, (int)(r() % 4), L, c1.c_str(), "\tXS:A:+", &n1[0], &t1[0], &q1[0]};
					srcuda_sam_mate m2{(int)(r() % 8 != 0), (int)(r() % 4 != 0), (int)(r() % 2), (int)(r() % 900000), (int)(r() % 2), 30, 0, L, c2.c_str(), nullptr, &n2[0], &t2[0], &q2[0]};
					size_t n = 0;
					int rc = srcuda_out_format_fragment(w, &cfg, &m1, &m2, 1 + (int)(r() % 2), 0, nullptr, buf, sizeof buf, &n);
					if (rc) { fprintf(stderr, "format_fragment: %d %s\n", rc, srcuda_out_error(nullptr)); abort(); }
					out[t].emplace_back(buf, n);
				}
			});
		for (auto &th : pool) th.join();
		for (int i = 0; i < per_thread; i++)
			for (int t = 0; t < n_threads; t++)
				if (srcuda_out_append(w, out[t][i].data(), out[t][i].size(), 1)) return 4;
		if (srcuda_out_close(w, nullptr)) return 5;
	}
	printf("lines taken %ld, refused %ld\n", taken, refused);
	return taken > 10000 && refused > 100 ? 0 : 6;
}
