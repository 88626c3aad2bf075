// sr_reader.cpp -- chunk reader in front of the voting step: one or two FASTQ files, plain or
// gzip, handed out as batches of fragments in the dense chunk layout srcuda_vote_chunk takes.  It stands where the
// reference has geinput_open + fetch_next_read_pair under the input mutex (input-files.c:455, core.c:1121-1215): the file
// is read in windows, every window is indexed at once by srcuda_fastq_split_mt and gathered by srcuda_reads_pack_mt
// (sr_fastq.cpp), and a batch ends after max_frags fragments the way a chunk ends after reads_per_chunk (core.c:1132).
// The record grammar is sr_fastq.cpp's (= the reference's plain-FASTQ branch); gzip input is inflated by zlib in front of
// the same grammar (the reference's own gzip branch reads lines with seekgz_gets and does not skip blank lines: the two
// agree on every file without blank lines).  Host code only.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/subread_cuda.h"

namespace {

// One block of file bytes read ahead by a thread of its own while the caller works on the block before.
struct Prefetch {
	std::thread th;
	std::mutex mu;
	std::condition_variable cv;
	std::vector<char> block;        // the finished block (valid while ready)
	size_t n = 0;
	bool ready = false, eof = false, stop = false;
	std::string err;                // not empty: reading failed
};

struct Stream {
	std::unique_ptr<Prefetch> pf;   // null: bytes are read on the caller's thread
	gzFile fp = nullptr;            // gzip input
	int fd = -1;                    // plain input: read() straight into the window
	uint64_t size = 0, read_pos = 0; // regular plain file: its size and the offset of the next unread byte (size 0: a pipe)
	std::vector<char> win;          // window of file bytes
	size_t have = 0;                // valid bytes in win
	size_t consumed = 0;            // bytes of win behind the last indexed record
	bool eof = false;               // the file has no more bytes
	bool ended = false;             // the grammar ended the input (end of file, or a read trimmed to nothing)
	uint64_t file_pos = 0;          // file offset (inflated bytes) of win[0]
	std::vector<uint64_t> off;      // indexed, not yet handed out: records [cursor, n)
	std::vector<uint64_t> name_off; // ... and where their names begin (srcuda_reader_tell: a record starts one byte earlier, at its '@')
	std::vector<uint64_t> qual_off; // ... and their quality lines (srcuda_reader_next_full)
	std::vector<uint16_t> len, name_len, qual_len;
	int64_t n = 0, cursor = 0;
	int64_t pending() const { return n - cursor; }
};

}  // namespace

struct srcuda_reader {
	Stream s[2];
	int files = 1;
	int threads = 1, trim_5 = 0, trim_3 = 0;
	size_t window = 0;
	std::string error;
	std::mutex error_mu;
	bool error_fresh = false;       // an error was recorded during the current call
	int64_t frags_out = 0;
	int deferred_rc = 0;            // a refill failed after fragments had been gathered: they were handed out, the error comes with the next call
	// srcuda_reader_next_full: where the names and quality lines of the call in progress go (null otherwise)
	struct Extra {
		char *names, *quals;
		uint64_t names_cap, quals_cap, names_used, quals_used;
		uint64_t *name_off, *qual_off;
		uint16_t *name_len, *qual_len;
	} *extra = nullptr;
};

namespace {

int fail(srcuda_reader *r, int rc, const char *fmt, ...)
{
	char msg[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(msg, sizeof(msg), fmt, ap);
	va_end(ap);
	std::lock_guard<std::mutex> g(r->error_mu);     // the two files of a pair are refilled side by side
	if (r->error_fresh) r->error += "; ";
	else r->error.clear();
	r->error += msg;
	r->error_fresh = true;
	return rc;
}

// Up to `want` bytes of the file into dst; fewer only at the end of the file (*eof).  false: err says why.
bool read_block(Stream &st, int threads, char *dst, size_t want, size_t *n_out, bool *eof, std::string &err)
{
	size_t have = 0;
	*eof = false;
	while (have < want && !*eof) {
		const size_t ask = std::min<size_t>(want - have, 1u << 30);
		long got;
		if (st.fp) {
			got = gzread(st.fp, dst + have, (unsigned)ask);
			if (got <= 0) {
				// gzread answers a stream that was cut short with 0 and Z_BUF_ERROR ("unexpected end of file"), not with -1:
				// only errnum tells a damaged file from a clean end
				int errnum = 0;
				const char *m = gzerror(st.fp, &errnum);
				if (got < 0 || (errnum != Z_OK && errnum != Z_STREAM_END)) {
					err = m && *m ? m : "gzip read error";
					return false;
				}
			}
		} else if (st.size && threads > 1 && std::min<uint64_t>(ask, st.size - st.read_pos) >= (8u << 20)) {
			// a regular file: the threads copy disjoint slices out of the page cache at once
			const uint64_t total = std::min<uint64_t>(ask, st.size - st.read_pos);
			const int T = (int)std::min<uint64_t>((uint64_t)threads, total >> 20);
			std::vector<long> res((size_t)T, 0);
			auto slice = [&](int t) {
				const uint64_t lo = total / T * t, hi = t == T - 1 ? total : total / T * (t + 1);
				uint64_t at = lo;
				while (at < hi) {
					const ssize_t k = pread(st.fd, dst + have + at, hi - at, (off_t)(st.read_pos + at));
					if (k < 0 && errno == EINTR) continue;
					if (k <= 0) { res[(size_t)t] = k < 0 ? -(long)errno : -1000000L; return; }   // error, or the file shrank
					at += (uint64_t)k;
				}
			};
			{
				std::vector<std::thread> th;
				th.reserve((size_t)T);
				for (int t = 1; t < T; t++) {
					bool started = false;
					try { th.emplace_back(slice, t); started = true; } catch (...) {}
					if (!started) slice(t);
				}
				slice(0);
				for (auto &x : th) x.join();
			}
			for (long v : res)
				if (v) { err = v == -1000000L ? "the file changed size" : strerror((int)-v); return false; }
			got = (long)total;
			st.read_pos += total;
		} else {
			do got = st.size ? (long)pread(st.fd, dst + have, ask, (off_t)st.read_pos) : (long)read(st.fd, dst + have, ask);
			while (got < 0 && errno == EINTR);
			if (got < 0) { err = strerror(errno); return false; }
			st.read_pos += (uint64_t)got;
		}
		if (got == 0) *eof = true;
		have += (size_t)got;
	}
	*n_out = have;
	return true;
}

// The read-ahead thread: block after block of `bytes` bytes, one finished block waiting at most.
void prefetch_loop(Stream *st, int threads, size_t bytes)
{
	Prefetch &pf = *st->pf;
	std::vector<char> work;
	while (true) {
		size_t n = 0;
		bool eof = false;
		std::string err;
		try {
			if (work.size() != bytes) work.resize(bytes);
			if (!read_block(*st, threads, work.data(), bytes, &n, &eof, err) && err.empty()) err = "read error";
		} catch (...) {
			err = "out of memory";
		}
		std::unique_lock<std::mutex> lk(pf.mu);
		pf.cv.wait(lk, [&] { return !pf.ready || pf.stop; });
		if (pf.stop) return;
		pf.block.swap(work);
		pf.n = n;
		pf.eof = eof;
		pf.err = err;
		pf.ready = true;
		pf.cv.notify_all();
		if (eof || !err.empty()) return;
	}
}

// More bytes behind win[have): one block (from the read-ahead thread when there is one).  Sets st.eof at the end of the file.
bool acquire(Stream &st, int threads, std::string &err)
{
	if (!st.pf) {
		if (st.have == st.win.size()) return true;      // the window is full: the caller grows it when it has to
		size_t n = 0;
		bool eof = false;
		if (!read_block(st, threads, st.win.data() + st.have, st.win.size() - st.have, &n, &eof, err)) return false;
		st.have += n;
		st.eof = eof;
		return true;
	}
	Prefetch &pf = *st.pf;
	std::unique_lock<std::mutex> lk(pf.mu);
	pf.cv.wait(lk, [&] { return pf.ready; });
	if (!pf.err.empty()) { err = pf.err; return false; }
	if (st.win.size() < st.have + pf.n) st.win.resize(st.have + pf.n);
	memcpy(st.win.data() + st.have, pf.block.data(), pf.n);
	st.have += pf.n;
	st.eof = pf.eof;
	pf.ready = false;
	pf.cv.notify_all();
	return true;
}

// Indexes the next window of one file: drops the bytes already handed out, tops the window up, splits it.
int refill(srcuda_reader *r, int which)
{
	Stream &st = r->s[which];
	if (st.ended) return SRCUDA_OK;
	if (st.consumed) {
		memmove(st.win.data(), st.win.data() + st.consumed, st.have - st.consumed);
		st.have -= st.consumed;
		st.file_pos += st.consumed;
		st.consumed = 0;
	}
	bool starved = false;
	while (true) {
		if (!st.eof && (st.have < r->window || starved)) {   // a window's worth of bytes, more only when they hold no record
			std::string err;
			if (!acquire(st, r->threads, err)) return fail(r, SRCUDA_ERR_IO, "read error in input file %d: %s", which + 1, err.c_str());
		}
		// room for records of 32 bytes and more; shorter ones just take another round (SRCUDA_FASTQ_FULL leaves the rest
		// of the window for the next refill)
		const size_t cap = st.have / 32 + 1024;
		if (st.off.size() < cap) {
			st.off.resize(cap); st.len.resize(cap); st.name_off.resize(cap); st.name_len.resize(cap); st.qual_off.resize(cap); st.qual_len.resize(cap);
		}
		srcuda_fastq_batch b;
		const int rc = srcuda_fastq_split_mt(st.win.data(), st.have, st.eof ? 1 : 0, r->trim_5, r->trim_3, (int64_t)cap, 0, 1, st.off.data(),
		                                     st.len.data(), st.name_off.data(), st.name_len.data(), st.qual_off.data(), st.qual_len.data(), r->threads, &b);
		if (rc == SRCUDA_ERR_FORMAT && b.n_reads == 0)
			return fail(r, rc, "input file %d: a format issue at byte %llu (a record must start with '@' and have a '+' line)", which + 1,
			            (unsigned long long)(st.file_pos + b.error_offset));
		if (rc && rc != SRCUDA_ERR_FORMAT) return fail(r, rc, "srcuda_fastq_split_mt failed (%d)", rc);
		st.n = b.n_reads;
		st.cursor = 0;
		st.consumed = (size_t)b.consumed;
		if (rc == SRCUDA_ERR_FORMAT) return SRCUDA_OK;   // the records in front of the issue go out first; the next refill reports it
		if (b.stop == SRCUDA_FASTQ_END) st.ended = true;
		if (b.n_reads > 0 || st.ended) return SRCUDA_OK;
		// not one complete record yet: a record longer than the window -- take more bytes
		starved = true;
		if (st.have >= ((size_t)1 << 32)) return fail(r, SRCUDA_ERR_FORMAT, "input file %d: a record longer than 4 GB", which + 1);
		if (!st.pf && st.have == st.win.size()) st.win.resize(st.win.size() * 2);
	}
}

}  // namespace

extern "C" int srcuda_reader_open(const char *path1, const char *path2, uint64_t window_bytes, int n_threads, int trim_5, int trim_3,
                                  srcuda_reader **out)
{
	if (!out) return SRCUDA_ERR_ARG;
	*out = nullptr;
	if (!path1 || trim_5 < 0 || trim_3 < 0) return SRCUDA_ERR_ARG;
	srcuda_reader *r = new (std::nothrow) srcuda_reader;
	if (!r) return SRCUDA_ERR_NOMEM;
	*out = r;                       // handed out even on failure so that srcuda_reader_error can say why
	r->files = path2 ? 2 : 1;
	r->threads = std::max(1, n_threads);
	r->trim_5 = trim_5;
	r->trim_3 = trim_3;
	// default 16 MB: the window stays in the last-level cache between read, index and gather
	r->window = (size_t)std::max<uint64_t>(window_bytes ? window_bytes : (16u << 20), 4096);
	for (int i = 0; i < r->files; i++) {
		const char *p = i ? path2 : path1;
		const int fd = open(p, O_RDONLY);
		if (fd < 0) return fail(r, SRCUDA_ERR_IO, "cannot open '%s': %s", p, strerror(errno));
		unsigned char magic[2] = {0, 0};
		const bool gz = pread(fd, magic, 2, 0) == 2 && magic[0] == 31 && magic[1] == 139;   // as geinput_open tells them apart (:473)
		if (gz) {
			r->s[i].fp = gzdopen(fd, "rb");
			if (!r->s[i].fp) { close(fd); return fail(r, SRCUDA_ERR_IO, "cannot open '%s' as gzip", p); }
			gzbuffer(r->s[i].fp, 1u << 20);
		} else {
			r->s[i].fd = fd;
			struct stat sb;
			if (fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode)) r->s[i].size = (uint64_t)sb.st_size;
#ifdef POSIX_FADV_SEQUENTIAL
			posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
		}
		try {
			r->s[i].win.resize(r->window);
		} catch (...) {
			return fail(r, SRCUDA_ERR_NOMEM, "no memory for a window of %zu bytes", r->window);
		}
		// gzip: a thread of its own inflates the next block while this one is indexed, gathered and mapped (inflate is the
		// serial part).  Plain files are copied out of the page cache by all threads at once instead; reading them ahead
		// through a second buffer was measured slower (7 vs 14 M reads/s on the build container).
		if (r->threads > 1 && r->s[i].fp) {
			try {
				r->s[i].pf.reset(new Prefetch);
				r->s[i].pf->th = std::thread(prefetch_loop, &r->s[i], r->threads, r->window);
			} catch (...) {
				r->s[i].pf.reset();     // no thread: bytes are read on the caller's thread
			}
		}
	}
	return SRCUDA_OK;
}

extern "C" int srcuda_reader_next(srcuda_reader *r, int64_t max_frags, char *text, uint64_t text_capacity, uint64_t *off, uint16_t *len,
                                  int64_t *n_frags, uint64_t *text_bytes)
{
	const bool skip = !text && !off && !len;       // srcuda_reader_skip: the fragments are indexed and passed over
	if (!r || !n_frags || max_frags < 0 || (max_frags && !skip && (!text || !off || !len))) return SRCUDA_ERR_ARG;
	*n_frags = 0;
	if (text_bytes) *text_bytes = 0;
	if (r->deferred_rc) {               // the error the call before ran into after it had gathered its fragments
		const int rc = r->deferred_rc;
		r->deferred_rc = 0;
		return rc;
	}
	r->error_fresh = false;
	try {
		int64_t done = 0;
		uint64_t written = 0;
		const int mates = r->files;
		// a failure with fragments already gathered (their cursors have moved on): hand those out now, fail the next call
		auto stop = [&](int rc) {
			if (done == 0) return rc;
			r->deferred_rc = rc;
			*n_frags = done;
			if (text_bytes) *text_bytes = written;
			r->frags_out += done;
			return (int)SRCUDA_OK;
		};
		while (done < max_frags) {
			const bool need0 = r->s[0].pending() == 0, need1 = r->files == 2 && r->s[1].pending() == 0;
			if (need0 && need1) {               // both files at once: two inflates (or two page-cache copies) side by side
				int rc1 = SRCUDA_OK;
				bool threaded = true;
				std::thread other;
				try {
					other = std::thread([&] {
						try { rc1 = refill(r, 1); } catch (...) { rc1 = fail(r, SRCUDA_ERR_NOMEM, "out of memory"); }
					});
				} catch (...) { threaded = false; }
				int rc0;
				try { rc0 = refill(r, 0); } catch (...) { rc0 = fail(r, SRCUDA_ERR_NOMEM, "out of memory"); }
				if (threaded) other.join();
				else rc1 = refill(r, 1);
				if (rc0 || rc1) return stop(rc0 ? rc0 : rc1);
			} else if (need0 || need1) {
				const int rc = refill(r, need0 ? 0 : 1);
				if (rc) return stop(rc);
			}
			int64_t take = max_frags - done;
			for (int i = 0; i < r->files; i++) take = std::min(take, r->s[i].pending());
			if (take == 0) {
				// the end: both files must have run out together (core.c:1177-1182)
				if (r->files == 2 && (r->s[0].pending() > 0) != (r->s[1].pending() > 0))
					return stop(fail(r, SRCUDA_ERR_FORMAT, "two input files have different amounts of reads"));
				break;
			}
			if (skip) {
				done += take;
				for (int i = 0; i < r->files; i++) r->s[i].cursor += take;
				continue;
			}
			const Stream &a = r->s[0], &b = r->s[1];
			uint64_t w = 0;
			const int rc = srcuda_reads_pack_mt(a.win.data(), a.off.data() + a.cursor, a.len.data() + a.cursor, mates == 2 ? b.win.data() : nullptr,
			                                    mates == 2 ? b.off.data() + b.cursor : nullptr, mates == 2 ? b.len.data() + b.cursor : nullptr, take,
			                                    text + written, text_capacity - written, off + done * mates, len + done * mates, r->threads, &w);
			if (rc == SRCUDA_ERR_ARG) return fail(r, rc, "the text buffer (%llu bytes) is too small for %lld fragments", (unsigned long long)text_capacity, (long long)max_frags);
			if (rc) return fail(r, rc, "srcuda_reads_pack_mt failed (%d)", rc);
			if (written)
				for (int64_t k = done * mates; k < (done + take) * mates; k++) off[k] += written;
			if (r->extra) {              // names and quality lines of the same fragments, in the same order (read = fragment * mates + mate)
				srcuda_reader::Extra &x = *r->extra;
				for (int64_t f = 0; f < take; f++)
					for (int m = 0; m < mates; m++) {
						const Stream &st = r->s[m];
						const size_t i = (size_t)(st.cursor + f), k = (size_t)((done + f) * mates + m);
						if (x.names_used + st.name_len[i] > x.names_cap || x.quals_used + st.qual_len[i] > x.quals_cap)
							return fail(r, SRCUDA_ERR_ARG, "the name or quality buffer is too small for %lld fragments", (long long)max_frags);
						memcpy(x.names + x.names_used, st.win.data() + st.name_off[i], st.name_len[i]);
						x.name_off[k] = x.names_used;
						x.name_len[k] = st.name_len[i];
						x.names_used += st.name_len[i];
						memcpy(x.quals + x.quals_used, st.win.data() + st.qual_off[i], st.qual_len[i]);
						x.qual_off[k] = x.quals_used;
						x.qual_len[k] = st.qual_len[i];
						x.quals_used += st.qual_len[i];
					}
			}
			written += w;
			done += take;
			for (int i = 0; i < r->files; i++) r->s[i].cursor += take;
		}
		*n_frags = done;
		if (text_bytes) *text_bytes = written;
		r->frags_out += done;
		return SRCUDA_OK;
	} catch (...) {
		return fail(r, SRCUDA_ERR_NOMEM, "out of memory");
	}
}

extern "C" int srcuda_reader_next_full(srcuda_reader *r, int64_t max_frags, char *text, uint64_t text_capacity, uint64_t *off, uint16_t *len,
                                       char *names, uint64_t names_capacity, uint64_t *name_off, uint16_t *name_len, char *quals,
                                       uint64_t quals_capacity, uint64_t *qual_off, uint16_t *qual_len, int64_t *n_frags, uint64_t *text_bytes)
{
	if (!r || !names || !name_off || !name_len || !quals || !qual_off || !qual_len || !text || !off || !len) return SRCUDA_ERR_ARG;
	srcuda_reader::Extra x{names, quals, names_capacity, quals_capacity, 0, 0, name_off, qual_off, name_len, qual_len};
	r->extra = &x;
	const int rc = srcuda_reader_next(r, max_frags, text, text_capacity, off, len, n_frags, text_bytes);
	r->extra = nullptr;
	return rc;
}

extern "C" int srcuda_reader_skip(srcuda_reader *r, int64_t n_frags, int64_t *skipped)
{
	int64_t n = 0;
	const int rc = srcuda_reader_next(r, n_frags, nullptr, 0, nullptr, nullptr, &n, nullptr);
	if (skipped) *skipped = n;
	return rc;
}

extern "C" int srcuda_reader_tell(const srcuda_reader *r, uint64_t *pos1, uint64_t *pos2)
{
	if (!r || !pos1) return SRCUDA_ERR_ARG;
	uint64_t *out[2] = {pos1, pos2};
	for (int i = 0; i < r->files; i++) {
		const Stream &st = r->s[i];
		// behind the last record handed out: where the next one begins (its '@'), or the end of what has been indexed
		const uint64_t in_win = st.cursor < st.n ? st.name_off[(size_t)st.cursor] - 1 : (uint64_t)st.consumed;
		if (out[i]) *out[i] = st.file_pos + in_win;
	}
	return SRCUDA_OK;
}

extern "C" const char *srcuda_reader_error(const srcuda_reader *r) { return r ? r->error.c_str() : "null reader"; }

extern "C" void srcuda_reader_close(srcuda_reader *r)
{
	if (!r) return;
	for (int i = 0; i < 2; i++) {
		if (r->s[i].pf && r->s[i].pf->th.joinable()) {
			{
				std::lock_guard<std::mutex> g(r->s[i].pf->mu);
				r->s[i].pf->stop = true;
			}
			r->s[i].pf->cv.notify_all();
			r->s[i].pf->th.join();
		}
	}
	for (int i = 0; i < 2; i++)
		if (r->s[i].fp) gzclose(r->s[i].fp);
		else if (r->s[i].fd >= 0) close(r->s[i].fd);
	delete r;
}
