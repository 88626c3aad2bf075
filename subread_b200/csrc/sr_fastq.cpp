// sr_fastq.cpp -- FASTQ batches for the voting step: a whole buffer of FASTQ text is
// indexed in one pass (memchr over lines) into the (off, len) arrays srcuda_vote_chunk takes, instead of one
// geinput_next_read_trim call per read under the input mutex (reference core.c:1130, input-files.c:982-1093).
// The read characters are not copied: offsets point into the caller's buffer, which can be the pinned chunk buffer.
//
// Record grammar = the reference's plain-FASTQ branch, byte for byte (tests/test_fastq_split.py holds it against
// geinput_next_read_trim of the unmodified reference):
//   * '\n' characters in front of a record are skipped; the next character must be '@' (input-files.c:993-1001);
//   * name, sequence and quality are read with read_line_noempty (input-files.c:209-262): leading '\n' skipped, the
//     line ends at '\n' or at the end of the input, at most max-1 characters are kept (199 for the name, 1209 for
//     sequence and quality; the rest of an over-long line is dropped), '\r' is an ordinary character;
//   * the name is cut at the first ' ' or '\t' behind its first character (input-files.c:1021-1030);
//   * between sequence and quality: '\n' skipped, then a '+' and the rest of that line (SKIP_LINE, :1038-1059);
//   * trimming is trim_read_inner (input-files.c:716-747);
//   * a read that is empty after trimming ends the input, as it does in fetch_next_read_pair (core.c:1168-1208).
// Bytes >= 0x80 are outside the contract (the reference reads them as EOF marks through a plain char).
#include <algorithm>
#include <cstring>

#include "../../include/subread_cuda.h"

namespace {

constexpr int kMaxReadLength = 1210;   // MAX_READ_LENGTH, subread.h:73
constexpr int kMaxReadNameLen = 200;   // MAX_READ_NAME_LEN, subread.h:74

struct Line {
	const char *begin;
	uint32_t kept;      // characters the reference keeps (line length clipped to max-1)
	bool terminated;    // a '\n' ended it (false: the buffer did)
};

// read_line_noempty (must_upper = 0): p is advanced past the terminating '\n'.
inline Line line_noempty(const char *&p, const char *e, int max_len)
{
	while (p < e && *p == '\n') p++;
	Line l;
	l.begin = p;
	const char *q = p < e ? static_cast<const char *>(memchr(p, '\n', (size_t)(e - p))) : nullptr;
	const char *end = q ? q : e;
	l.kept = (uint32_t)std::min<size_t>((size_t)(end - p), (size_t)(max_len - 1));
	l.terminated = q != nullptr;
	p = q ? q + 1 : e;
	return l;
}

}  // namespace

extern "C" int srcuda_fastq_split(const char *buf, uint64_t n_bytes, int at_eof, int trim_5, int trim_3, int64_t max_reads,
                                  uint64_t base_offset, int64_t stride, uint64_t *off, uint16_t *len, uint64_t *name_off,
                                  uint16_t *name_len, uint64_t *qual_off, uint16_t *qual_len, srcuda_fastq_batch *out)
{
	if (!out) return SRCUDA_ERR_ARG;
	memset(out, 0, sizeof(*out));
	if ((!buf && n_bytes) || !off || !len || max_reads < 0 || stride < 1 || trim_5 < 0 || trim_3 < 0) return SRCUDA_ERR_ARG;
	const char *const b = buf, *const e = buf + n_bytes;
	const char *p = b;
	int64_t n = 0;
	uint32_t longest = 0;
	int stop = SRCUDA_FASTQ_MORE, rc = SRCUDA_OK;
	while (true) {
		if (n >= max_reads) { stop = SRCUDA_FASTQ_FULL; break; }
		const char *rec = p;
		while (rec < e && *rec == '\n') rec++;
		if (rec == e) {               // nothing but line ends left
			p = rec;
			stop = at_eof ? SRCUDA_FASTQ_END : SRCUDA_FASTQ_MORE;
			break;
		}
		if (*rec != '@') { out->error_offset = (uint64_t)(rec - b); rc = SRCUDA_ERR_FORMAT; break; }
		const char *c = rec + 1;
		const Line name = line_noempty(c, e, kMaxReadNameLen);
		if (!name.terminated && !at_eof) break;
		const Line seq = line_noempty(c, e, kMaxReadLength);
		if (!seq.terminated && !at_eof) break;
		while (c < e && *c == '\n') c++;
		if (c == e) {
			if (!at_eof) break;
			out->error_offset = n_bytes;   // the reference meets EOF where it wants '+': "a format issue" (:1041)
			rc = SRCUDA_ERR_FORMAT;
			break;
		}
		if (*c != '+') { out->error_offset = (uint64_t)(c - b); rc = SRCUDA_ERR_FORMAT; break; }
		const char *plus_end = static_cast<const char *>(memchr(c, '\n', (size_t)(e - c)));
		if (!plus_end && !at_eof) break;
		c = plus_end ? plus_end + 1 : e;
		const Line qual = line_noempty(c, e, kMaxReadLength);
		if (!qual.terminated && !at_eof) break;

		// trim_read_inner
		int rlen = (int)seq.kept, start = 0, qlen = (int)qual.kept;
		if (trim_5 || trim_3) {
			if (rlen > trim_5 && rlen - trim_5 > trim_3) {
				start = trim_5;
				rlen = rlen - trim_5 - trim_3;
				qlen = std::max(0, std::min(qlen - trim_5, rlen));
			} else rlen = 0;
		}
		if (rlen <= 0) {              // "normally finished" for the reference: the record is not consumed
			p = rec;
			stop = SRCUDA_FASTQ_END;
			break;
		}
		uint32_t nl = name.kept;
		for (uint32_t i = 1; i < name.kept; i++)
			if (name.begin[i] == ' ' || name.begin[i] == '\t') { nl = i; break; }
		const int64_t at = n * stride;
		off[at] = base_offset + (uint64_t)(seq.begin - b) + (uint64_t)start;
		len[at] = (uint16_t)rlen;
		if (name_off) name_off[at] = base_offset + (uint64_t)(name.begin - b);
		if (name_len) name_len[at] = (uint16_t)nl;
		if (qual_off) qual_off[at] = base_offset + (uint64_t)(qual.begin - b) + (uint64_t)start;
		if (qual_len) qual_len[at] = (uint16_t)qlen;
		longest = std::max<uint32_t>(longest, (uint32_t)rlen);
		n++;
		p = c;
	}
	out->n_reads = n;
	out->consumed = (uint64_t)(p - b);
	out->stop = rc ? SRCUDA_FASTQ_END : stop;
	out->max_len = longest;
	return rc;
}

extern "C" int srcuda_reads_pack(const char *buf1, const uint64_t *off1, const uint16_t *len1, const char *buf2,
                                 const uint64_t *off2, const uint16_t *len2, int64_t n_frags, char *text, uint64_t text_capacity,
                                 uint64_t *off, uint16_t *len, uint64_t *text_bytes)
{
	if (n_frags < 0 || !off || !len || (n_frags && (!buf1 || !off1 || !len1 || !text))) return SRCUDA_ERR_ARG;
	if (buf2 && (!off2 || !len2)) return SRCUDA_ERR_ARG;
	uint64_t w = 0;
	int64_t r = 0;
	for (int64_t f = 0; f < n_frags; f++) {
		for (int m = 0; m < (buf2 ? 2 : 1); m++, r++) {
			const char *src = m ? buf2 + off2[f] : buf1 + off1[f];
			const uint16_t l = m ? len2[f] : len1[f];
			if (w + l > text_capacity) return SRCUDA_ERR_ARG;
			memcpy(text + w, src, l);
			off[r] = w;
			len[r] = l;
			w += l;
		}
	}
	if (text_bytes) *text_bytes = w;
	return SRCUDA_OK;
}
