/* oracle/ref_reader_bench.c -- the reference's own FASTQ reader (geinput_open + geinput_next_read_trim,
 * input-files.c:455,982-1093) timed on one thread, for the comparison in DESIGN.md's "FASTQ batches".  Test
 * infrastructure: links the unmodified reference (oracle/_ref/libsubread_ref.so), not the product.
 *   gcc -O2 -o oracle/_ref/ref_reader_bench oracle/ref_reader_bench.c -Loracle/_ref -lsubread_ref \
 *       -Wl,-rpath,$PWD/oracle/_ref -lz -lm -lpthread
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

int geinput_open(const char *, void *);
int geinput_next_read_trim(void *, char *, char *, char *, short, short, int *);

int main(int argc, char **argv)
{
	void *g = calloc(1, 1 << 20); /* gene_input_t, generously */
	char name[4096], seq[4096], qual[4096];
	int sec = 0;
	long n = 0;
	struct timespec a, b;
	if (argc < 2 || geinput_open(argv[1], g)) return 1;
	clock_gettime(CLOCK_MONOTONIC, &a);
	while (geinput_next_read_trim(g, name, seq, qual, 0, 0, &sec) > 0) n++;
	clock_gettime(CLOCK_MONOTONIC, &b);
	double dt = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
	printf("reference reader: %ld reads in %.3f s = %.2f M reads/s\n", n, dt, n / dt / 1e6);
	return 0;
}
