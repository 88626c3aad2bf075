// tests/native/fastq_fuzz.cpp -- srcuda_fastq_split_mt under AddressSanitizer / UBSan: generated FASTQ with random damage
// (bytes replaced, removed, inserted), random trimming, max_reads and at_eof, in exact-size heap buffers.  One thread and
// several threads must report the same records, verdict and error offset, and everything reported must lie inside the
// buffer.  Built and run by tests/test_fastq_split.py::test_sanitizer_fuzz.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "../../include/subread_cuda.h"
int main() {
	std::mt19937_64 rng(12345);
	const char alphabet[] = "ACGTN@+\n\n\n \tIH#r1/";
	long cases = 0, mt_cases = 0;
	for (int it = 0; it < 3000; it++) {
		std::string s;
		int nrec = rng() % 3000;
		bool big = it % 10 == 0;
		if (!big) nrec %= 40;
		for (int i = 0; i < nrec; i++) {
			int L = 1 + rng() % (big ? 200 : 60);
			s += "@r" + std::to_string(i) + (rng() % 4 ? "" : " x y") + "\n";
			for (int k = 0; k < L; k++) s += "ACGTN"[rng() % 5];
			s += (rng() % 50 ? "\n+\n" : "\n\n+r\n");
			for (int k = 0; k < L; k++) s += "IH#@+5"[rng() % 6];
			s += "\n";
		}
		int muts = rng() % 4;
		for (int m = 0; m < muts && !s.empty(); m++) {
			size_t p = rng() % s.size();
			switch (rng() % 3) {
			case 0: s[p] = alphabet[rng() % (sizeof(alphabet) - 1)]; break;
			case 1: s.erase(p, 1 + rng() % 5); break;
			default: s.insert(p, 1, alphabet[rng() % (sizeof(alphabet) - 1)]); break;
			}
		}
		// exact-size heap copy so that ASan sees any over-read
		char *buf = (char *)malloc(s.size() ? s.size() : 1);
		memcpy(buf, s.data(), s.size());
		size_t cap = s.size() / 6 + 2;
		int t5 = rng() % 3 ? 0 : rng() % 8, t3 = rng() % 3 ? 0 : rng() % 8;
		int64_t maxr = rng() % 3 ? (int64_t)cap : (int64_t)(rng() % (cap + 1));
		int at_eof = rng() % 4 != 0;
		std::vector<uint64_t> o1(cap), o2(cap), n1(cap), n2(cap), q1(cap), q2(cap);
		std::vector<uint16_t> l1(cap), l2(cap), nl1(cap), nl2(cap), ql1(cap), ql2(cap);
		srcuda_fastq_batch b1, b2;
		int r1 = srcuda_fastq_split_mt(buf, s.size(), at_eof, t5, t3, maxr, 5, 1, o1.data(), l1.data(), n1.data(), nl1.data(), q1.data(), ql1.data(), 1, &b1);
		int T = 2 + rng() % 7;
		int r2 = srcuda_fastq_split_mt(buf, s.size(), at_eof, t5, t3, maxr, 5, 1, o2.data(), l2.data(), n2.data(), nl2.data(), q2.data(), ql2.data(), T, &b2);
		bool same = r1 == r2 && b1.n_reads == b2.n_reads && b1.consumed == b2.consumed && b1.stop == b2.stop && b1.max_len == b2.max_len &&
		            b1.error_offset == b2.error_offset;
		for (int64_t i = 0; same && i < b1.n_reads; i++)
			same = o1[i] == o2[i] && l1[i] == l2[i] && n1[i] == n2[i] && nl1[i] == nl2[i] && q1[i] == q2[i] && ql1[i] == ql2[i];
		for (int64_t i = 0; same && i < b1.n_reads; i++)    // everything reported lies inside the buffer
			same = o1[i] - 5 + l1[i] <= s.size() && n1[i] - 5 + nl1[i] <= s.size() && q1[i] - 5 + ql1[i] <= s.size() && b1.consumed <= s.size();
		if (!same) { printf("MISMATCH at case %d (T=%d, size %zu, rc %d/%d, n %ld/%ld)\n", it, T, s.size(), r1, r2, (long)b1.n_reads, (long)b2.n_reads); return 1; }
		cases++;
		if (s.size() >= 2 * 65536) mt_cases++;
		free(buf);
	}
	printf("ok: %ld cases, %ld large enough for the threaded path\n", cases, mt_cases);
	return 0;
}
