// sr_common.cuh -- types and small helpers shared by the warp kernel, the exact scalar path and
// the host API.  Everything marked SR_HD compiles both as device code (nvcc) and as plain host
// C++ (g++ with -DSR_HOST_ONLY), the latter only so that tests can run the scalar path on a CPU
// box; the shipped library never executes it on the host.
//
// Reference citations are relative to the reference's src/ directory.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(SR_HOST_ONLY)
#define SR_HD __host__ __device__ __forceinline__
#define SR_D __device__ __forceinline__
#else
#define SR_HD inline
#define SR_D inline
#endif

namespace sr {

constexpr int kVoteRows = 30;        // GENE_VOTE_TABLE_SIZE   subread.h:217
constexpr int kVoteSpace = 24;       // GENE_VOTE_SPACE        subread.h:216
constexpr int kMaxSections = 7;      // MAX_INDEL_SECTIONS     subread.h:88
constexpr int kRecLen = 21;          // MAX_INDEL_TOLERANCE*3  subread.h:231
constexpr int kSeg = 5;              // INDEL_SEGMENT_SIZE     sorted-hashtable.c:344
constexpr int kCoreNegStrand = 8;    // CORE_IS_NEGATIVE_STRAND core.h:258
constexpr int kLongRead = 160;       // EXON_LONG_READ_LENGTH  subread.h:83
constexpr int kMaxReadLen = 1210;    // MAX_READ_LENGTH        subread.h:73
constexpr int kMaxLookups = 63 * 3;  // <= 63 subreads (core.c:3122-3126) x gap <= 3
constexpr int kMaxBest = 3;          // multi_best_reads (core.c:4075-4084)
constexpr int kMaxTop = 3;           // top_scores (core-indel.c:4415)
constexpr int kMaxSimples = 64;      // max_vote_simples PE (core.c:4077)
constexpr int kMaxCombos = 16;       // max_vote_combinations, raised by -B up to 16
constexpr int kResultWords = 17;     // sizeof(mapping_result_t) / 4

// Width of the per-bucket Bloom word (A/B switch): 16 bits = 32 MB for the one-block gapped table, 8 bits = 16 MB.
#ifndef SR_BLOOM_BITS
#define SR_BLOOM_BITS 16
#endif
#if SR_BLOOM_BITS == 8
typedef uint8_t bloom_t;
#else
typedef uint16_t bloom_t;
#endif

// One index block resident in HBM (SoA form of gehash_t + gene_value_index_t + gene_offset_t).
struct DevIndex {
	uint32_t buckets_number;
	uint64_t buckets_magic;        // floor(2^64 / buckets_number) + 1: key / buckets = umul64hi(key, magic), exact for 32-bit keys
	int32_t index_gap;
	int32_t padding;
	uint32_t start_point, length, start_base_offset, values_bytes;
	int32_t total_offsets;
	const uint32_t *bucket_start;  // buckets_number + 1
	const int16_t *keys;           // n_items, ascending inside a bucket (signed compare)
	const uint32_t *vals;          // n_items
	const uint32_t *genome;        // 2-bit bases, LSB first, as u32 words (padded)
	const uint32_t *chr_offsets;   // total_offsets cumulative END offsets
	const uint32_t *chr_hint;      // derived on device: [b] = first chromosome whose end offset lies above b << kChrHintShift (locate_position starts there); null = off
	// Lookup accelerators derived on the device at load time (sr_api.cu, k_build_records); both are
	// exact: a filter without false negatives and a re-packed copy of small buckets.
	const bloom_t *bloom;          // per bucket: bloom_bits(key16) set for every key16 in the bucket; null = off
	const void *records;           // per bucket 32 B (two uint4): see below; null = off
	// Dense tables (a human-sized index holds ~64 items per bucket): the direct filter, see below; null = off
	// Gapped index (gap 3), sparse table: the phase table, see below; null = off
	const uint32_t *phase_any;     // bit c set: unit c holds at least one key
	const void *phase_units;       // 2^28 units of 32 bytes
	const uint32_t *phase_dense;   // dense table: 2^28 units of 128 bytes; null = off
	const uint32_t *direct;        // one 16-byte unit (four words) per kDirectKeys consecutive 32-bit keys
	const uint64_t *direct_first;  // by rank among the present keys: {position, 1} for a key with one position, else
	                               // {first item of its run in vals[], h | landing << 8} (low, high word): run length and the index inside the
	                               // run where the reference's binary search lands (visit order, sorted-hashtable.c:1109-1119)
};

// The direct filter answers a lookup of a dense table with one 64-byte DRAM fetch when the key is absent and two when it
// has a single position, instead of bucket header + binary search over the bucket's keys + position (three dependent
// fetches).  bucket = key % buckets and key16 = key / buckets are a bijection of the 32-bit key space, so the table's
// keys can be addressed by the key itself.  Unit u covers keys [48 u, 48 u + 48):
//   w0      number of distinct keys present in all lower units (rank base)
//   w1..w3  two bits per key, key k at bit 2 (k % 16) of word 1 + (k % 48) / 16:  bit 0 = the table holds the key,
//           bit 1 = with more than one position (those lookups take the bucket path: the reference's visit order
//           starts at the landing index of its own binary search, sorted-hashtable.c:960-981, 1109-1119).
// 2^32 / 48 units = 1.43 GB whatever the table; direct_first adds 8 bytes per distinct key.
constexpr uint32_t kDirectKeys = 48;
constexpr uint32_t kDirectUnits = 89478486u;   // ceil(2^32 / 48)
SR_HD uint32_t direct_unit(uint32_t key) { return (uint32_t)(((uint64_t)key * 0xAAAAAAABull) >> 37); }   // key / 48

// The phase table of a gapped index (gap 3).  do_voting looks every subread up at three consecutive offsets o, o+1, o+2
// (core.c:3169-3171: the index holds every third position, so one of the three phases lines up).  The three 16-mers share 14
// bases, c = bases o+2 .. o+15 (28 bits): K0 = xx|c, K1 = x|c|y, K2 = c|yy.  For a fixed c there are 16 + 16 + 16 = 48 such
// keys, and unit c of the table says for each of them whether the index holds it -- so ONE 32-byte fetch answers all three
// lookups of a subread, instead of three filter probes and a bucket record per survivor.  Unit c (eight words):
//   w0, w1, w2   role 0, 1, 2: two bits per variant v (bit 0 = the table holds the key, bit 1 = with several positions);
//                role 0: v = K0 >> 28;  role 1: v = (K1 >> 30) << 2 | (K1 & 3);  role 2: v = K2 & 15
//   w3 .. w7     the positions of the first five keys present, in (role, v) order (a key with several positions, or one
//                beyond the fifth, is looked up in its bucket: the reference's visit order needs the bucket anyway)
// Every index key sits in three units (once per role).  2^28 units = 8.6 GB whatever the index; phase_any (32 MB, kept in L2)
// tells which units are not empty, so that only those are fetched: on a 100 Mbp genome two thirds of the foreign subreads
// stop there.
// Dense tables (a human-sized index: a quarter of all 16-mers present, twelve keys per unit) use units of 128 bytes: w0-w2 as
// above, w3 spare, w4 .. w31 the values of the first 28 keys present in (role, v) order -- the position of a key with a single
// one, and for a key with several the rank of its entry in direct_first (run start, length, landing index).  There is no
// occupancy bitmap (no unit is empty): one 64-byte fetch per subread answers three lookups and, two times out of three, holds
// the position too.  2^28 units = 34 GB.
constexpr int kPhaseDenseValues = 28;
constexpr int kPhaseDenseWords = 32;
constexpr uint32_t kPhaseMask = 0x0fffffffu;
constexpr uint32_t kPhaseUnits = 1u << 28;
constexpr int kPhaseValues = 5;
SR_HD uint32_t phase_unit(uint32_t key, int role) { return role == 0 ? (key & kPhaseMask) : (role == 1 ? ((key >> 2) & kPhaseMask) : (key >> 4)); }
SR_HD uint32_t phase_variant(uint32_t key, int role) { return role == 0 ? (key >> 28) : (role == 1 ? (((key >> 30) << 2) | (key & 3u)) : (key & 15u)); }

// 32-byte bucket record (one sector), n = items in the bucket (saturated at 0xffff):
//   n <= kInlineItems:  bytes 0-1 n   bytes 2-11 key16[5]   bytes 12-31 value[5]        (the whole bucket)
//   n <= kRecordKeys:   bytes 0-1 n   bytes 4-7 bucket_start   bytes 8-31 key16[12]      (positions stay in vals[])
//   larger:             only n is valid; the lookup falls back to bucket_start / keys / vals.
constexpr int kInlineItems = 5;
constexpr int kRecordKeys = 12;
constexpr uint32_t kRecordNoKey = 0x7fffu;   // key16 field of an unused slot of an inline record: above every key16 of a table
                                             // with at least kRecordMinBuckets buckets (key16 = key / buckets <= 0x7ffe)
constexpr uint32_t kRecordMinBuckets = 131076u;

// The two bits a key16 sets in its bucket's 16-bit Bloom word (a 2-hash blocked Bloom filter: with two items per
// bucket a quarter of the bits are set and a foreign key passes about one time in twenty).
SR_HD uint32_t bloom_bits(int k16)
{
	const uint32_t q = (uint32_t)k16;
#if SR_BLOOM_BITS == 8
	return (1u << (q & 7u)) | (1u << ((q >> 3) & 7u));
#else
	return (1u << (q & 15u)) | (1u << ((q >> 4) & 15u));
#endif
}

// The voting scalars, validated and pre-digested on the host (see srcuda_params).
struct VoteConfig {
	int32_t total_subreads, top_scores, max_vote_simples, max_vote_combinations, max_vote_number_cutoff;
	int32_t multi_best_reads, min_first, min_second, indel_tolerance;  // indel_tolerance = min(16, max_indel_length)
	int32_t min_pair_dist, max_pair_dist, first_reversed, second_reversed;
	int32_t is_paired;
	uint32_t low_border, high_border;  // core.c:3081-3082
	// subjunc (do_breakpoint_detection): minor-half / donor-site search, core-junction.c:1073-1334, 3675-3832
	int32_t do_junctions, big_margin_filtering, big_margin_record_size, maximum_intron_length;
	int32_t check_donor, mismatch_allowed, prefer_donor;
};

// subjunc_result_t (core.h:397-410) as four little-endian words:
//  w0 split_point,minor_votes | w1 double_indel_offset,indel_at_junction,small_inc,large_inc | w2 minor_position
//  w3 minor_coverage_start,minor_coverage_end
#if defined(__CUDACC__) && !defined(SR_HOST_ONLY)
// Gather loads.  Measured on B200 (tools/micro/sector_fetch*.cu, profiles/): a load that misses L2 pulls the whole
// 128-byte line from DRAM (3.6 sectors, 107 B per random 32-byte record; cudaLimitMaxL2FetchGranularity changes
// nothing), unless the instruction carries the .L2::64B prefetch size -- then two sectors (58 B).  Every random
// single-sector access of the lookup (bucket record, bucket header, positions) therefore asks for 64 B.
#ifndef SR_GATHER_64B
#define SR_GATHER_64B 1
#endif
#if SR_GATHER_64B
#define SR_L2PF ".L2::64B"
#else
#define SR_L2PF ""
#endif
__device__ __forceinline__ uint32_t ld_gather(const uint32_t *p)
{
	uint32_t v;
	asm volatile("ld.global.nc" SR_L2PF ".u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
// Direct filter, step 1: the word that holds the key's two bits (one 64-byte fetch).  Returns the bits.
__device__ __forceinline__ uint32_t direct_bits(const DevIndex &ix, uint32_t key, uint32_t *unit_out)
{
	const uint32_t u = direct_unit(key), e = key - u * kDirectKeys;
	*unit_out = u;
	const uint32_t w = ld_gather(ix.direct + 4 * (size_t)u + 1 + (e >> 4));
	return (w >> (2 * (e & 15u))) & 3u;
}
// Step 2 for a present key: the whole unit (an L2 hit after step 1) -> rank among the present keys.
__device__ __forceinline__ uint4 ld_direct_unit(const DevIndex &ix, uint32_t u)
{
	uint4 v;
	asm volatile("ld.global.nc" SR_L2PF ".v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ix.direct + 4 * (size_t)u));
	return v;
}
__device__ __forceinline__ uint2 ld_direct_first(const DevIndex &ix, uint32_t rank)
{
	uint2 v;
	asm volatile("ld.global.nc" SR_L2PF ".v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(ix.direct_first + rank));
	return v;
}
__device__ __forceinline__ uint32_t direct_rank(const uint4 &unit, uint32_t key, uint32_t u, uint32_t *bits_out)
{
	const uint32_t e = key - u * kDirectKeys, wi = e >> 4, sh = 2 * (e & 15u);
	const uint32_t w = wi == 0 ? unit.y : (wi == 1 ? unit.z : unit.w);
	*bits_out = (w >> sh) & 3u;
	uint32_t r = unit.x + __popc(w & 0x55555555u & ((1u << sh) - 1u));
	if (wi > 0) r += __popc(unit.y & 0x55555555u);
	if (wi > 1) r += __popc(unit.z & 0x55555555u);
	return r;
}

// One 32-byte bucket record in a single 256-bit load (sm_100), streaming: not kept in L1, first to leave L2.
struct __align__(32) Record8 { uint32_t w[8]; };
__device__ __forceinline__ Record8 ld_record(const uint4 *records, uint32_t b)
{
	Record8 r;
	asm volatile("ld.global.L1::no_allocate.L2::evict_first" SR_L2PF ".v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
	             : "l"(records + 2 * (size_t)b));
	return r;
}
#endif

struct JuncWords { uint32_t w[4]; };
constexpr int kBigMarginU16 = 27;    // bigtable_cached_result_t.big_margin_data, core.h:453
constexpr int kConfirmWindow = 17;   // JUNCTION_CONFIRM_WINDOW core-junction.c:3569

// mapping_result_t as 17 little-endian words:
//  w0 pos | w1 flags,read_length | w2 votes,used | w3 noninf,indels,full,pad | w4..w14 record[22]
//  w15 cov_start,cov_end | w16 subread_quality,pad
struct ResultWords { uint32_t w[kResultWords]; };

SR_HD int16_t result_votes(const ResultWords &r) { return (int16_t)(r.w[2] & 0xffff); }

// base2int (subread.h:238) on a *signed* char, as genekey2int passes it (input-files.c:1241).
SR_HD int base_code(int c /* sign-extended char */) { return c < 'G' ? (c == 'A' ? 0 : 2) : (c == 'G' ? 1 : 3); }

// Characters the reverse-complement table (input-files.c:1111) maps to a base rather than 'N'.
SR_HD bool is_table_base(int c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'U'; }

// 2-bit code of reverse_read(c) (input-files.c:1171-1186): A<->T, C<->G, U->A, others -> 'N' -> 3.
SR_HD int rc_code(int c) { return is_table_base(c) ? 3 - base_code(c) : 3; }

// 2-bit code of reverse_read(reverse_read(c)): bases keep their code ('U' becomes 'T', code 3
// either way), everything else has become 'N' -> 3.
SR_HD int rcrc_code(int c) { return is_table_base(c) ? base_code(c) : 3; }

// _index_vote_tol, sorted-hashtable.c:347
SR_HD uint32_t vote_row(uint32_t kv) { return (kv / kSeg) % kVoteRows; }

// Subread schedule of do_voting (core.c:3117-3130).  Returns applied_subreads and the 16.16 step.
SR_HD int subread_schedule(int rlen, int gap, int total_subreads, int *step_out)
{
	int cr15 = (rlen - 15 - gap) * 65536;
	int step;
	if (rlen <= kLongRead) {
		step = cr15 / (total_subreads - 1);
		if (step < gap * 65536) step = gap * 65536;
	} else {
		step = 6 * 65536;
		if (cr15 / step > 62) step = cr15 / 62;
	}
	*step_out = step;
	return 1 + cr15 / step;
}

// Offset of lookup l = subread_no * gap + xk1 (core.c:3169-3171).
SR_HD int lookup_offset(int l, int gap, int step)
{
	int s = l / gap, x = l - s * gap;
	int o = (step * s) >> 16;
	if (gap > 1) o -= o % gap - x;
	return o;
}

// locate_gene_position_max with NULL cuts and rl 0 (gene-algorithms.c:441-510): 0 when `linear`
// lies inside a chromosome body; then *chr is the chromosome number and *pos the 0-based offset.
// `hint` (DevIndex::chr_hint, optional): the scan may start at any chromosome not above the first one whose end offset exceeds
// `linear`; hint[linear >> kChrHintShift] is such a start, one dependent load instead of the thirteen of a binary search over the
// 6 200 contigs of a human-sized index (9 % of the paired-end tier's stall samples went there).
constexpr int kChrHintShift = 16;
SR_HD int locate_position(const uint32_t *offs, int total, int padding, uint32_t linear, int *chr, int *pos, const uint32_t *hint = nullptr)
{
	int n, lo = 0, hi = total;
	*chr = -1;
	*pos = -1;
	if (hint) n = (int)hint[linear >> kChrHintShift];
	else
	for (;;) {
		if (hi <= lo + 1) { n = lo - 2 > 0 ? lo - 2 : 0; break; }
		int mid = (lo + hi) / 2;
		if (offs[mid] > linear) hi = mid; else lo = mid + 1;
	}
	for (; n < total; n++) {
		uint32_t end = offs[n];
		if (end > linear) {
			int p = (int)(n == 0 ? linear : linear - offs[n - 1]);
			*pos = p;
			if (linear > end + 15 - (uint32_t)padding) return 1;
			if (p < padding) return 1;
			*pos = p - padding;
			*chr = n;
			return 0;
		}
	}
	return -1;
}

// test_PE_and_same_chro (core.c:4819-4845) on pre-located mates.  Returns the pair weight of
// core-junction.c:2337-2349 without annotation: 1300 proper pair, 1000 same chromosome, 800 else;
// *is_pe tells whether the pair is within [min,max] distance.
SR_HD int pair_weight(int re1, int chr1, int pos1, int re2, int chr2, int pos2, int rlen1, int rlen2, int min_dist,
                      int max_dist, int *is_pe)
{
	*is_pe = 0;
	if (re1 == 0 && re2 == 0 && chr1 == chr2) {
		long long t = (long long)pos1 - pos2;
		if (t < 0) t = -t;
		t += pos1 > pos2 ? rlen1 : rlen2;
		uint32_t tl = (uint32_t)t;
		if (tl >= (uint32_t)min_dist && tl <= (uint32_t)max_dist) { *is_pe = 1; return 1300; }
		return 1000;
	}
	return 800;
}

}  // namespace sr
