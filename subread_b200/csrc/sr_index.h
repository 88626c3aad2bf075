// sr_index.h -- host-side loader of the reference's on-disk index block
// (<prefix>.NN.b.tab / .array and <prefix>.reads) into the SoA form the
// kernels use.  Replaces gehash_load (sorted-hashtable.c:1390-1568), gvindex_load
// (gene-value-index.c:190-228) and the numeric half of load_offsets (gene-algorithms.c:1293).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sr {

struct HostIndex {
	int32_t index_gap = 0, padding = 0;
	uint32_t buckets_number = 0;
	uint64_t n_items = 0;
	// The three big arrays may live in pinned memory (alloc/free supplied by the caller).
	uint32_t *bucket_start = nullptr;  // buckets_number + 1
	int16_t *keys = nullptr;
	uint32_t *vals = nullptr;
	uint32_t start_point = 0, length = 0, start_base_offset = 0, values_bytes = 0;
	std::vector<uint8_t> values;
	std::vector<uint32_t> read_offsets;
	void *(*alloc)(size_t) = nullptr;
	void (*release)(void *) = nullptr;
	~HostIndex();
	void free_arrays();
};

// Returns 0 or an SRCUDA_ERR_* code, with a message in err.
int load_index_block(const std::string &prefix, int block_no, HostIndex *out, std::string *err);

}  // namespace sr
