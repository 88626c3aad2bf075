/* oracle/ref_writer_bench.c -- the reference's BAM writer (SamBam_writer_add_read_line through the single stream,
 * sambam-file.c:1668,1193) and the library's (srcuda_out_add_sam_lines) timed on the same SAM records, for the figures in
 * DESIGN.md's "Output side".  Test infrastructure: links the unmodified reference (oracle/_ref/libsubread_ref.so) next to the
 * product library; nothing in the product links this.
 *   gcc -O2 -o oracle/_ref/ref_writer_bench oracle/ref_writer_bench.c -Iinclude -Loracle/_ref -lsubread_ref \
 *       -Lsubread_b200 -lsubread_cuda -Wl,-rpath,$PWD/oracle/_ref -Wl,-rpath,$PWD/subread_b200 -lz -lm -lpthread
 *   usage: ref_writer_bench records.sam repeat out_dir threads...
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "subread_output.h"

int SamBam_writer_create(void *, char *, int, int, int, char *);
int SamBam_writer_add_header(void *, char *, int);
void SamBam_writer_finish_header(void *);
int SamBam_writer_add_read_line(void *, int, char *, int);
int SamBam_writer_close(void *);

static double now(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int argc, char **argv)
{
	if (argc < 5) return 1;
	FILE *f = fopen(argv[1], "rb");
	if (!f) return 1;
	fseek(f, 0, SEEK_END);
	long size = ftell(f);
	fseek(f, 0, SEEK_SET);
	char *all = malloc(size + 1);
	if (fread(all, 1, size, f) != (size_t)size) return 1;
	all[size] = 0;
	fclose(f);
	long head = 0; /* header lines first */
	while (all[head] == '@') head = strchr(all + head, '\n') - all + 1;
	long body = size - head, n_lines = 0;
	for (long i = head; i < size; i++) n_lines += all[i] == '\n';
	int repeat = atoi(argv[2]);
	char path[1024];

	/* the reference, single stream */
	void *w = calloc(1, 1 << 20);
	snprintf(path, sizeof path, "%s/ref.bam", argv[3]);
	SamBam_writer_create(w, path, 1, 0, 0, "");
	for (long at = 0; at < head;) {
		char *nl = strchr(all + at, '\n');
		char line[20000];
		memcpy(line, all + at, nl - (all + at));
		line[nl - (all + at)] = 0;
		SamBam_writer_add_header(w, line, 1);
		at = nl - all + 1;
	}
	SamBam_writer_finish_header(w);
	double t0 = now();
	char *line = malloc(100000);
	for (int r = 0; r < repeat; r++)
		for (long at = head; at < size;) {
			char *nl = strchr(all + at, '\n');
			memcpy(line, all + at, nl - (all + at) + 1);
			line[nl - (all + at) + 1] = 0;
			SamBam_writer_add_read_line(w, -1, line, 1);
			at = nl - all + 1;
		}
	SamBam_writer_close(w);
	double dt = now() - t0;
	printf("reference writer (one stream): %ld records, %.1f MB of SAM in %.3f s = %.4f M records/s\n", n_lines * repeat, body * (double)repeat / 1e6, dt,
	       n_lines * repeat / dt / 1e6);

	for (int a = 4; a < argc; a++) {
		int threads = atoi(argv[a]);
		srcuda_writer *nw;
		snprintf(path, sizeof path, "%s/native_%d.bam", argv[3], threads);
		if (srcuda_out_open(path, SRCUDA_OUT_BAM, threads, 0, &nw)) return 2;
		for (long at = 0; at < head;) {
			char *nl = strchr(all + at, '\n');
			char hl[20000];
			memcpy(hl, all + at, nl - (all + at));
			hl[nl - (all + at)] = 0;
			srcuda_out_add_header(nw, hl, 1);
			at = nl - all + 1;
		}
		t0 = now();
		for (int r = 0; r < repeat; r++)
			if (srcuda_out_add_sam_lines(nw, all + head, body, 1)) return 3;
		if (srcuda_out_close(nw, NULL)) return 4;
		dt = now() - t0;
		printf("library writer, %2d deflate threads: %.3f s = %.4f M records/s\n", threads, dt, n_lines * repeat / dt / 1e6);
	}
	return 0;
}
