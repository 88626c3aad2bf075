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
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

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

struct Record {
	uint64_t seq, name, qual;   // offsets from the start of the parsed range
	uint16_t len, name_len, qual_len;
};

// What a piece of the multi-threaded split keeps until the pieces are known to fit together: its records in the
// caller's own format, so that accepting them is one block copy per array.
struct Columns {
	std::vector<uint64_t> off, name_off, qual_off;
	std::vector<uint16_t> len, name_len, qual_len;
	size_t n = 0;
	void room(size_t want, bool names, bool quals)
	{
		if (off.size() >= want && (!names || name_off.size() >= want) && (!quals || qual_off.size() >= want)) return;
		want += want / 2 + 1024;
		off.resize(want);
		len.resize(want);
		if (names) { name_off.resize(want); name_len.resize(want); }
		if (quals) { qual_off.resize(want); qual_len.resize(want); }
	}
};

struct ParseEnd {
	int rc = SRCUDA_OK, stop = SRCUDA_FASTQ_MORE;
	uint64_t consumed = 0, error_offset = 0;
};

// The record loop of geinput_next_read_trim over [b, e).  emit(Record) is called for every record the reference
// would hand out, in order.
template <class Emit>
ParseEnd parse_range(const char *b, const char *e, bool at_eof, int trim_5, int trim_3, int64_t max_reads, Emit emit)
{
	ParseEnd r;
	const char *p = b;
	int64_t n = 0;
	while (true) {
		if (n >= max_reads) { r.stop = SRCUDA_FASTQ_FULL; break; }
		const char *rec = p;
		while (rec < e && *rec == '\n') rec++;
		if (rec == e) {               // nothing but line ends left
			p = rec;
			r.stop = at_eof ? SRCUDA_FASTQ_END : SRCUDA_FASTQ_MORE;
			break;
		}
		if (*rec != '@') { r.error_offset = (uint64_t)(rec - b); r.rc = SRCUDA_ERR_FORMAT; break; }
		const char *c = rec + 1;
		const Line name = line_noempty(c, e, kMaxReadNameLen);
		if (!name.terminated && !at_eof) break;
		const Line seq = line_noempty(c, e, kMaxReadLength);
		if (!seq.terminated && !at_eof) break;
		while (c < e && *c == '\n') c++;
		if (c == e) {
			if (!at_eof) break;
			r.error_offset = (uint64_t)(e - b);   // the reference meets EOF where it wants '+': "a format issue" (:1041)
			r.rc = SRCUDA_ERR_FORMAT;
			break;
		}
		if (*c != '+') { r.error_offset = (uint64_t)(c - b); r.rc = SRCUDA_ERR_FORMAT; break; }
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
			r.stop = SRCUDA_FASTQ_END;
			break;
		}
		uint32_t nl = name.kept;
		for (uint32_t i = 1; i < name.kept; i++)
			if (name.begin[i] == ' ' || name.begin[i] == '\t') { nl = i; break; }
		Record R;
		R.seq = (uint64_t)(seq.begin - b) + (uint64_t)start;
		R.name = (uint64_t)(name.begin - b);
		R.qual = (uint64_t)(qual.begin - b) + (uint64_t)std::min<uint32_t>((uint32_t)start, qual.kept);   // never past its line
		R.len = (uint16_t)rlen;
		R.name_len = (uint16_t)nl;
		R.qual_len = (uint16_t)qlen;
		emit(R);
		n++;
		p = c;
	}
	r.consumed = (uint64_t)(p - b);
	if (r.rc) r.stop = SRCUDA_FASTQ_END;
	return r;
}

// The pieces' arrays of the threaded split are kept between calls (a streaming caller indexes window after window of the same
// size): fresh memory costs a page fault per 4 KB, and page faults of many threads serialise in the kernel.
// srcuda_fastq_split_mt(NULL, 0, ...) -- an empty buffer -- gives the memory back.
class ScratchPool {
public:
	Columns take()
	{
		std::lock_guard<std::mutex> g(mu_);
		if (free_.empty()) return {};
		Columns v = std::move(free_.back());
		free_.pop_back();
		v.n = 0;
		return v;
	}
	void give(Columns &&v)
	{
		std::lock_guard<std::mutex> g(mu_);
		if (free_.size() < 64) free_.push_back(std::move(v));
	}
	void drop()
	{
		std::lock_guard<std::mutex> g(mu_);
		free_.clear();
		free_.shrink_to_fit();
	}

private:
	std::mutex mu_;
	std::vector<Columns> free_;
};

ScratchPool &scratch_pool()
{
	static ScratchPool pool;
	return pool;
}

// fn(t) for t = 0 .. T-1, t = 0 on the calling thread.  Nothing escapes: a task that throws (out of memory) or a thread
// that cannot be started makes the call return false after every started thread has been joined (the library never
// terminates the caller's process); tasks whose thread could not be started run on the calling thread.
template <class Fn>
bool run_tasks(int T, Fn fn)
{
	std::atomic<bool> ok{true};
	auto guarded = [&](int t) {
		try { fn(t); } catch (...) { ok = false; }
	};
	std::vector<std::thread> th;
	try { th.reserve((size_t)T); } catch (...) { ok = false; }
	for (int t = 1; t < T; t++) {
		bool started = false;
		if (th.capacity() > th.size()) {
			try { th.emplace_back(guarded, t); started = true; } catch (...) {}
		}
		if (!started) guarded(t);
	}
	guarded(0);
	for (auto &x : th) x.join();
	return ok;
}

struct Sink {
	uint64_t base;       // added to every offset
	int64_t stride;
	uint64_t *off;
	uint16_t *len;
	uint64_t *name_off;
	uint16_t *name_len;
	uint64_t *qual_off;
	uint16_t *qual_len;
	inline void put(int64_t i, const Record &R, uint64_t shift) const
	{
		const int64_t at = i * stride;
		off[at] = base + shift + R.seq;
		len[at] = R.len;
		if (name_off) name_off[at] = base + shift + R.name;
		if (name_len) name_len[at] = R.name_len;
		if (qual_off) qual_off[at] = base + shift + R.qual;
		if (qual_len) qual_len[at] = R.qual_len;
	}
};

int split_sequential(const char *buf, uint64_t n_bytes, bool at_eof, int trim_5, int trim_3, int64_t max_reads, const Sink &sink,
                     srcuda_fastq_batch *out)
{
	int64_t n = 0;
	uint32_t longest = 0;
	const ParseEnd r = parse_range(buf, buf + n_bytes, at_eof, trim_5, trim_3, max_reads, [&](const Record &R) {
		sink.put(n++, R, 0);
		longest = std::max<uint32_t>(longest, R.len);
	});
	out->n_reads = n;
	out->consumed = r.consumed;
	out->stop = r.stop;
	out->max_len = longest;
	out->error_offset = r.error_offset;
	return r.rc;
}

// A position >= from where a record very probably starts: a line that begins with '@' whose second next line begins
// with '+'.  (A quality line may begin with '@' too, but then the second next line is a sequence.)  Only a guess: the
// caller checks that the piece in front ends exactly here.
const char *guess_record_start(const char *b, const char *from, const char *e)
{
	const char *p = from;
	for (int tries = 0; tries < 64; tries++) {
		const char *nl = p < e ? static_cast<const char *>(memchr(p, '\n', (size_t)(e - p))) : nullptr;
		if (!nl || nl + 1 >= e) return nullptr;
		p = nl + 1;
		if (*p != '@') continue;
		const char *l2 = static_cast<const char *>(memchr(p, '\n', (size_t)(e - p)));
		if (!l2 || l2 + 1 >= e) return nullptr;
		const char *l3 = static_cast<const char *>(memchr(l2 + 1, '\n', (size_t)(e - l2 - 1)));
		if (!l3 || l3 + 1 >= e) return nullptr;
		if (l3[1] == '+') return p;
	}
	return nullptr;
}

int split_threaded(const char *buf, uint64_t n_bytes, bool at_eof, int trim_5, int trim_3, int64_t max_reads, const Sink &sink, int T,
                   srcuda_fastq_batch *out);

}  // namespace

extern "C" int srcuda_fastq_split(const char *buf, uint64_t n_bytes, int at_eof, int trim_5, int trim_3, int64_t max_reads,
                                  uint64_t base_offset, int64_t stride, uint64_t *off, uint16_t *len, uint64_t *name_off,
                                  uint16_t *name_len, uint64_t *qual_off, uint16_t *qual_len, srcuda_fastq_batch *out)
{
	return srcuda_fastq_split_mt(buf, n_bytes, at_eof, trim_5, trim_3, max_reads, base_offset, stride, off, len, name_off, name_len,
	                             qual_off, qual_len, 1, out);
}

extern "C" int srcuda_fastq_split_mt(const char *buf, uint64_t n_bytes, int at_eof, int trim_5, int trim_3, int64_t max_reads,
                                     uint64_t base_offset, int64_t stride, uint64_t *off, uint16_t *len, uint64_t *name_off,
                                     uint16_t *name_len, uint64_t *qual_off, uint16_t *qual_len, int n_threads,
                                     srcuda_fastq_batch *out)
{
	if (!out) return SRCUDA_ERR_ARG;
	memset(out, 0, sizeof(*out));
	if ((!buf && n_bytes) || !off || !len || max_reads < 0 || stride < 1 || trim_5 < 0 || trim_3 < 0) return SRCUDA_ERR_ARG;
	const Sink sink{base_offset, stride, off, len, name_off, name_len, qual_off, qual_len};
	if (n_bytes == 0) scratch_pool().drop();
	const uint64_t kMinPiece = 1u << 16;
	int T = (int)std::min<uint64_t>((uint64_t)std::max(1, n_threads), n_bytes / kMinPiece);
	if (T <= 1 || max_reads == 0) return split_sequential(buf, n_bytes, at_eof != 0, trim_5, trim_3, max_reads, sink, out);
	try {
		return split_threaded(buf, n_bytes, at_eof != 0, trim_5, trim_3, max_reads, sink, T, out);
	} catch (...) {      // no memory for the pieces' records: the plain loop needs none
		memset(out, 0, sizeof(*out));
		return split_sequential(buf, n_bytes, at_eof != 0, trim_5, trim_3, max_reads, sink, out);
	}
}

namespace {

int split_threaded(const char *buf, uint64_t n_bytes, bool at_eof_b, int trim_5, int trim_3, int64_t max_reads, const Sink &sink, int T,
                   srcuda_fastq_batch *out)
{
	const int at_eof = at_eof_b ? 1 : 0;
	// Speculative cut points; piece t = [cut[t], cut[t+1]).
	const char *const b = buf, *const e = buf + n_bytes;
	std::vector<const char *> cut;
	cut.push_back(b);
	for (int t = 1; t < T; t++) {
		const char *g = guess_record_start(b, std::max(cut.back(), b + n_bytes / (uint64_t)T * (uint64_t)t), e);
		if (g && g > cut.back()) cut.push_back(g);
	}
	cut.push_back(e);
	T = (int)cut.size() - 1;
	if (T <= 1) return split_sequential(buf, n_bytes, at_eof != 0, trim_5, trim_3, max_reads, sink, out);

	// Every piece is parsed as an input of its own.  The pieces together are the sequential parse if and only if every
	// piece but the last one is consumed entirely without an error and without an early end (induction over the pieces:
	// piece 0 starts where the sequential parse starts, and a parse that consumes its piece hands over at cut[t+1]).
	std::vector<Columns> recs((size_t)T);
	for (auto &v : recs) v = scratch_pool().take();
	struct GiveBack {
		std::vector<Columns> &r;
		~GiveBack() { for (auto &v : r) if (!v.off.empty()) scratch_pool().give(std::move(v)); }
	} give_back{recs};
	const bool names = sink.name_off || sink.name_len, quals = sink.qual_off || sink.qual_len;
	std::vector<ParseEnd> ends((size_t)T);
	std::vector<uint32_t> longest((size_t)T, 0);
	auto work = [&](int t) {
		const bool last = t == T - 1;
		Columns &C = recs[(size_t)t];
		C.room((size_t)(cut[t + 1] - cut[t]) / 256 + 16, names, quals);
		const uint64_t shift = sink.base + (uint64_t)(cut[t] - b);
		uint32_t m = 0;
		// a piece that is not the last one is parsed like a buffer with more to come (at_eof = false): it is only accepted when
		// its last record ends with its own '\n' inside the piece -- with at_eof a piece that stops right behind a '+' line
		// would pass with an empty quality line, which the sequential parse never sees
		ends[(size_t)t] = parse_range(cut[t], cut[t + 1], last ? at_eof != 0 : false, trim_5, trim_3, INT64_MAX, [&](const Record &R) {
			if (C.n == C.off.size()) C.room(C.n + 1, names, quals);
			const size_t i = C.n++;
			C.off[i] = shift + R.seq;
			C.len[i] = R.len;
			if (names) { C.name_off[i] = shift + R.name; C.name_len[i] = R.name_len; }
			if (quals) { C.qual_off[i] = shift + R.qual; C.qual_len[i] = R.qual_len; }
			m = std::max<uint32_t>(m, R.len);
		});
		longest[(size_t)t] = m;
	};
	const bool ran = run_tasks(T, work);
	int64_t total = 0;
	bool ok = ran;
	for (int t = 0; t < T; t++) {
		if (t < T - 1 && (ends[(size_t)t].rc || ends[(size_t)t].stop != SRCUDA_FASTQ_MORE || ends[(size_t)t].consumed != (uint64_t)(cut[t + 1] - cut[t]))) ok = false;
		total += (int64_t)recs[(size_t)t].n;
	}
	if (!ok || ends[(size_t)T - 1].rc) {  // odd input: the plain loop decides
#ifdef SR_FASTQ_TRACE
		fprintf(stderr, "fastq_split_mt: fallback\n");
#endif
		return split_sequential(buf, n_bytes, at_eof != 0, trim_5, trim_3, max_reads, sink, out);
	}
#ifdef SR_FASTQ_TRACE
	fprintf(stderr, "fastq_split_mt: %d pieces accepted\n", T);
#endif

	std::vector<int64_t> first((size_t)T + 1, 0);
	for (int t = 0; t < T; t++) first[(size_t)t + 1] = first[(size_t)t] + (int64_t)recs[(size_t)t].n;
	const int64_t keep = std::min(total, max_reads);
	auto place = [&](auto *dst, const auto &src, int64_t at, int64_t n) {      // dst[(at + i) * stride] = src[i]
		if (!dst || n <= 0) return;
		if (sink.stride == 1) memcpy(dst + at, src.data(), (size_t)n * sizeof(*dst));
		else for (int64_t i = 0; i < n; i++) dst[(at + i) * sink.stride] = src[(size_t)i];
	};
	auto copy = [&](int t) {
		const Columns &C = recs[(size_t)t];
		const int64_t at = first[(size_t)t], n = std::min<int64_t>((int64_t)C.n, keep - at);
		place(sink.off, C.off, at, n);
		place(sink.len, C.len, at, n);
		place(sink.name_off, C.name_off, at, n);
		place(sink.name_len, C.name_len, at, n);
		place(sink.qual_off, C.qual_off, at, n);
		place(sink.qual_len, C.qual_len, at, n);
		if (n < (int64_t)C.n) {          // cut short by max_reads: the longest read among the kept ones
			uint32_t m = 0;
			for (int64_t i = 0; i < n; i++) m = std::max<uint32_t>(m, C.len[(size_t)i]);
			longest[(size_t)t] = m;
		}
	};
	if (!run_tasks(T, copy)) return split_sequential(buf, n_bytes, at_eof != 0, trim_5, trim_3, max_reads, sink, out);
	out->n_reads = keep;
	out->max_len = *std::max_element(longest.begin(), longest.end());
	if (total >= max_reads) {            // the arrays are full: stop behind record keep-1, as the plain loop does
		int t = 0;
		while (first[(size_t)t + 1] < keep) t++;
		const ParseEnd cutoff = parse_range(cut[t], cut[t + 1], t == T - 1 ? at_eof != 0 : false, trim_5, trim_3, keep - first[(size_t)t],
		                                    [](const Record &) {});      // where record keep-1 ends
		out->consumed = (uint64_t)(cut[t] - b) + cutoff.consumed;
		out->stop = SRCUDA_FASTQ_FULL;
		return SRCUDA_OK;
	}
	const ParseEnd &le = ends[(size_t)T - 1];
	out->consumed = (uint64_t)(cut[(size_t)T - 1] - b) + le.consumed;
	out->stop = le.stop;
	return SRCUDA_OK;
}

}  // namespace

extern "C" int srcuda_reads_pack(const char *buf1, const uint64_t *off1, const uint16_t *len1, const char *buf2,
                                 const uint64_t *off2, const uint16_t *len2, int64_t n_frags, char *text, uint64_t text_capacity,
                                 uint64_t *off, uint16_t *len, uint64_t *text_bytes)
{
	return srcuda_reads_pack_mt(buf1, off1, len1, buf2, off2, len2, n_frags, text, text_capacity, off, len, 1, text_bytes);
}

extern "C" int srcuda_reads_pack_mt(const char *buf1, const uint64_t *off1, const uint16_t *len1, const char *buf2,
                                    const uint64_t *off2, const uint16_t *len2, int64_t n_frags, char *text, uint64_t text_capacity,
                                    uint64_t *off, uint16_t *len, int n_threads, uint64_t *text_bytes)
{
	if (n_frags < 0 || !off || !len || (n_frags && (!buf1 || !off1 || !len1 || !text))) return SRCUDA_ERR_ARG;
	if (buf2 && (!off2 || !len2)) return SRCUDA_ERR_ARG;
	const int T = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, n_frags / 4096));
	const int mates = buf2 ? 2 : 1;
	// fragments [lo(t), lo(t+1)) per thread: sizes first, then every thread copies to its own stretch of text
	auto lo = [&](int t) { return n_frags / T * t + std::min<int64_t>(t, n_frags % T); };
	std::vector<uint64_t> start((size_t)T + 1, 0);
	auto measure = [&](int t) {
		uint64_t w = 0;
		for (int64_t f = lo(t); f < lo(t + 1); f++) w += (uint64_t)len1[f] + (buf2 ? (uint64_t)len2[f] : 0u);
		start[(size_t)t + 1] = w;
	};
	auto copy = [&](int t) {
		uint64_t w = start[(size_t)t];
		int64_t r = lo(t) * mates;
		for (int64_t f = lo(t); f < lo(t + 1); f++) {
			for (int m = 0; m < mates; m++, r++) {
				const char *src = m ? buf2 + off2[f] : buf1 + off1[f];
				const uint16_t l = m ? len2[f] : len1[f];
				memcpy(text + w, src, l);
				off[r] = w;
				len[r] = l;
				w += l;
			}
		}
	};
	if (!run_tasks(T, measure)) return SRCUDA_ERR_NOMEM;
	for (int t = 0; t < T; t++) start[(size_t)t + 1] += start[(size_t)t];
	if (start[(size_t)T] > text_capacity) return SRCUDA_ERR_ARG;
	if (!run_tasks(T, copy)) return SRCUDA_ERR_NOMEM;
	if (text_bytes) *text_bytes = start[(size_t)T];
	return SRCUDA_OK;
}
