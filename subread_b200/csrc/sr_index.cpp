// sr_index.cpp -- see sr_index.h.  The .tab file is walked through mmap in one pass: per bucket
// {i32 n, i32 space, i16 keys[n], u32 vals[n]} (the u32 values are only 2-byte aligned when n is
// odd, so they are memcpy'd).
#include "sr_index.h"
#include "../../include/subread_cuda.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace sr {

void HostIndex::free_arrays()
{
	void (*rel)(void *) = release ? release : free;
	if (bucket_start) rel(bucket_start);
	if (keys) rel(keys);
	if (vals) rel(vals);
	bucket_start = nullptr; keys = nullptr; vals = nullptr;
}

HostIndex::~HostIndex() { free_arrays(); }

namespace {

struct Mapped {
	const uint8_t *p = nullptr;
	size_t n = 0;
	int fd = -1;
	~Mapped()
	{
		if (p) munmap((void *)p, n);
		if (fd >= 0) close(fd);
	}
	int open_file(const std::string &path, std::string *err)
	{
		fd = ::open(path.c_str(), O_RDONLY);
		if (fd < 0) { *err = "cannot open " + path; return SRCUDA_ERR_IO; }
		struct stat st;
		if (fstat(fd, &st) || st.st_size <= 0) { *err = "cannot stat " + path; return SRCUDA_ERR_IO; }
		n = (size_t)st.st_size;
		void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
		if (m == MAP_FAILED) { *err = "cannot mmap " + path; return SRCUDA_ERR_IO; }
		p = (const uint8_t *)m;
		madvise(m, n, MADV_SEQUENTIAL);
		return 0;
	}
};

template <typename T>
bool take(const uint8_t *&cur, const uint8_t *end, T *out)
{
	if ((size_t)(end - cur) < sizeof(T)) return false;
	memcpy(out, cur, sizeof(T));
	cur += sizeof(T);
	return true;
}

int load_tab(const std::string &path, HostIndex *ix, std::string *err)
{
	Mapped f;
	int rc = f.open_file(path, err);
	if (rc) return rc;
	const uint8_t *cur = f.p, *end = f.p + f.n;
	if (f.n < 8 || memcmp(cur, "2subindx", 8)) { *err = path + ": not a version-2 subread index (magic)"; return SRCUDA_ERR_FORMAT; }
	cur += 8;
	ix->index_gap = 0;
	ix->padding = 0;
	for (;;) {  // option cells, sorted-hashtable.c:1436-1466
		int16_t okey, olen, v;
		if (!take(cur, end, &okey)) { *err = path + ": truncated header"; return SRCUDA_ERR_FORMAT; }
		if (!okey) break;
		if (!take(cur, end, &olen)) { *err = path + ": truncated header"; return SRCUDA_ERR_FORMAT; }
		if (okey == 0x0101 || okey == 0x0102) {
			if (!take(cur, end, &v)) { *err = path + ": truncated header"; return SRCUDA_ERR_FORMAT; }
			if (okey == 0x0101) ix->index_gap = v; else ix->padding = v;
		} else {
			if ((size_t)(end - cur) < (size_t)olen) { *err = path + ": truncated header"; return SRCUDA_ERR_FORMAT; }
			cur += olen;
		}
	}
	if (ix->index_gap < 1) { *err = path + ": index gap missing"; return SRCUDA_ERR_FORMAT; }
	int64_t items;
	int32_t nb;
	if (!take(cur, end, &items) || !take(cur, end, &nb) || items < 1 || items > 0xffffffffLL || nb < 1) {
		*err = path + ": unrecognizable item/bucket counts";
		return SRCUDA_ERR_FORMAT;
	}
	ix->n_items = (uint64_t)items;
	ix->buckets_number = (uint32_t)nb;
	void *(*al)(size_t) = ix->alloc ? ix->alloc : malloc;
	ix->bucket_start = (uint32_t *)al(sizeof(uint32_t) * ((size_t)nb + 1));
	ix->keys = (int16_t *)al(sizeof(int16_t) * (size_t)items + 16);
	ix->vals = (uint32_t *)al(sizeof(uint32_t) * (size_t)items + 16);
	if (!ix->bucket_start || !ix->keys || !ix->vals) { *err = "out of host memory for the index"; return SRCUDA_ERR_NOMEM; }
	uint64_t at = 0;
	for (int32_t b = 0; b < nb; b++) {
		int32_t n, space;
		if (!take(cur, end, &n) || !take(cur, end, &space) || n < 0) { *err = path + ": truncated bucket table"; return SRCUDA_ERR_FORMAT; }
		ix->bucket_start[b] = (uint32_t)at;
		size_t need = (size_t)n * 6;
		if ((size_t)(end - cur) < need || at + (uint64_t)n > ix->n_items) { *err = path + ": index is incomplete"; return SRCUDA_ERR_FORMAT; }
		memcpy(ix->keys + at, cur, (size_t)n * 2);
		memcpy(ix->vals + at, cur + (size_t)n * 2, (size_t)n * 4);
		cur += need;
		at += (uint64_t)n;
	}
	ix->bucket_start[nb] = (uint32_t)at;
	if (at != ix->n_items) { *err = path + ": item count mismatch"; return SRCUDA_ERR_FORMAT; }
	return 0;
}

int load_array(const std::string &path, HostIndex *ix, std::string *err)
{
	FILE *fp = fopen(path.c_str(), "rb");
	if (!fp) { *err = "cannot open " + path; return SRCUDA_ERR_IO; }
	if (fread(&ix->start_point, 4, 1, fp) != 1 || fread(&ix->length, 4, 1, fp) != 1) {
		fclose(fp);
		*err = path + ": the array index is incomplete";
		return SRCUDA_ERR_FORMAT;
	}
	ix->start_base_offset = ix->start_point - ix->start_point % 4;  // gene-value-index.c:208
	uint32_t useful = (ix->length + ix->start_point - ix->start_base_offset) >> 2;
	ix->values_bytes = useful + 1;
	ix->values.assign((size_t)ix->values_bytes, 0);
	size_t got = fread(ix->values.data(), 1, ix->values_bytes, fp);
	fclose(fp);
	if (got < useful) { *err = path + ": the array index is incomplete"; return SRCUDA_ERR_FORMAT; }
	return 0;
}

int load_reads(const std::string &path, HostIndex *ix, std::string *err)
{
	FILE *fp = fopen(path.c_str(), "r");
	if (!fp) { *err = "cannot open " + path; return SRCUDA_ERR_IO; }
	char line[4096];
	ix->read_offsets.clear();
	while (fgets(line, sizeof(line), fp)) {
		if (strlen(line) < 2) continue;
		char *tab = strchr(line, '\t');
		if (!tab) continue;
		*tab = 0;
		ix->read_offsets.push_back((uint32_t)atoll(line));
	}
	fclose(fp);
	if (ix->read_offsets.empty()) { *err = path + ": no chromosome offsets"; return SRCUDA_ERR_FORMAT; }
	return 0;
}

}  // namespace

int load_index_block(const std::string &prefix, int block_no, HostIndex *out, std::string *err)
{
	char suffix[32];
	snprintf(suffix, sizeof(suffix), ".%02d.b.tab", block_no);
	int rc = load_tab(prefix + suffix, out, err);
	if (rc) return rc;
	snprintf(suffix, sizeof(suffix), ".%02d.b.array", block_no);
	rc = load_array(prefix + suffix, out, err);
	if (rc) return rc;
	return load_reads(prefix + ".reads", out, err);
}

}  // namespace sr
