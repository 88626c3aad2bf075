// sr_output.cpp -- SAM / BAM output side of libsubread_cuda.so (include/subread_output.h).
//
// Host C++ only.  The reference builds one record at a time into a 70 000-byte buffer and deflates the buffer whenever it
// passes 55 000 bytes, inline, under the output lock (sambam-file.c:1193-1246, :1793).  A block depends on nothing but its
// own bytes, so here record encoding runs ahead into a list of sealed blocks and the blocks are deflated by a team of
// threads and written in order: the file is the reference's single-stream file byte for byte (same cuts, same zlib
// settings), produced at the speed of its unordered multi-stream output.
#include "../../include/subread_cuda.h"
#include "../../include/subread_output.h"

#include <zlib.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

constexpr size_t kBlockCut = 55000;     // sambam-file.c:1793 / :1253: a block is sealed once it is past this
constexpr size_t kDeflateRoom = 70000;  // sambam-file.c:1209: room the reference gives deflate
constexpr int kMaxCigarOps = 96;        // FC_MAX_CIGAR_SECTIONS, sambam-file.c:1603
constexpr size_t kSamBuffer = 4u << 20;

static thread_local std::string g_open_error;

struct Chromosome {
	std::string name;
	uint32_t length;
};

struct Deflated {
	std::vector<unsigned char> bytes;  // the whole BGZF block: header, payload, CRC, size
	bool ok = true;
};

struct srcuda_writer {
	FILE *fp = nullptr;
	bool own_fp = false;
	int format = SRCUDA_OUT_SAM;
	int threads = 1;
	int level = Z_BEST_SPEED;
	bool header_done = false;
	bool header_done_writing = false;  // the header blocks are out: blocks from here on hold records
	bool failed = false;
	std::string error;
	std::string header_text;  // BAM: kept until finish_header
	std::vector<Chromosome> chromosomes;
	std::unordered_map<std::string, int> chromosome_id;
	std::string block;                 // the block under construction (BAM) or the text buffer (SAM)
	// complete blocks on their way through the deflate team, in file order
	struct Job {
		std::string plain;
		Deflated out;
		bool done = false;
	};
	std::deque<std::unique_ptr<Job>> jobs;
	uint64_t front_seq = 0, claim_seq = 0;  // sequence number of jobs.front() / of the next job a worker takes
	std::mutex mu;
	std::condition_variable cv_work, cv_done;
	std::vector<std::thread> team;
	bool stop = false;
	uint64_t bytes_written = 0;

	int fail(int code, const char *fmt, ...) {
		char buf[512];
		va_list ap;
		va_start(ap, fmt);
		vsnprintf(buf, sizeof buf, fmt, ap);
		va_end(ap);
		error = buf;
		failed = true;
		return code;
	}
};

namespace {

void put32(std::string &b, uint32_t v) {
	char c[4] = {char(v), char(v >> 8), char(v >> 16), char(v >> 24)};
	b.append(c, 4);
}

// One BGZF block around `plain` (header bytes of SamBam_writer_chunk_header, sambam-file.c:1155-1185).
void deflate_block(const std::string &plain, int level, Deflated &out) {
	out.bytes.resize(18 + kDeflateRoom + 8);
	z_stream zs;
	memset(&zs, 0, sizeof zs);
	if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
		out.ok = false;
		return;
	}
	zs.next_in = (Bytef *)plain.data();
	zs.avail_in = (uInt)plain.size();
	zs.next_out = out.bytes.data() + 18;
	zs.avail_out = kDeflateRoom;
	int rc = deflate(&zs, Z_FINISH);
	size_t csize = kDeflateRoom - zs.avail_out;
	deflateEnd(&zs);
	if (rc != Z_STREAM_END || csize + 25 > 65535) {  // the reference would write a block no reader can take
		out.ok = false;
		return;
	}
	unsigned char *h = out.bytes.data();
	const unsigned char head[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 66, 67, 2, 0};
	memcpy(h, head, 16);
	uint32_t bsize = (uint32_t)csize + 19 + 6;
	h[16] = bsize & 0xff;
	h[17] = (bsize >> 8) & 0xff;
	uint32_t crc = (uint32_t)crc32(crc32(0, nullptr, 0), (const Bytef *)plain.data(), (uInt)plain.size());
	uint32_t isize = (uint32_t)plain.size();
	unsigned char *t = h + 18 + csize;
	for (int i = 0; i < 4; i++) t[i] = (crc >> (8 * i)) & 0xff;
	for (int i = 0; i < 4; i++) t[4 + i] = (isize >> (8 * i)) & 0xff;
	out.bytes.resize(18 + csize + 8);
}

int write_bytes(srcuda_writer *w, const void *p, size_t n) {
	if (n && fwrite(p, 1, n, w->fp) != n) return w->fail(SRCUDA_ERR_IO, "short write: no space left in the output directory?");
	w->bytes_written += n;
	return SRCUDA_OK;
}

// The team: a worker takes the oldest block nobody has taken, deflates it and marks it done.
void team_worker(srcuda_writer *w) {
	std::unique_lock<std::mutex> lk(w->mu);
	for (;;) {
		w->cv_work.wait(lk, [&] { return w->stop || w->claim_seq < w->front_seq + w->jobs.size(); });
		if (w->claim_seq >= w->front_seq + w->jobs.size()) return;  // stop, and nothing left to take
		srcuda_writer::Job *job = w->jobs[(size_t)(w->claim_seq - w->front_seq)].get();
		w->claim_seq++;
		lk.unlock();
		deflate_block(job->plain, w->level, job->out);
		lk.lock();
		job->done = true;
		w->cv_done.notify_all();
	}
}

void start_team(srcuda_writer *w) {
	if (w->threads < 2 || !w->team.empty()) return;
	for (int t = 0; t < w->threads; t++) {
		try {
			w->team.emplace_back(team_worker, w);
		} catch (...) {
			break;  // fewer workers; with none the caller's thread deflates (harvest)
		}
	}
}

void stop_team(srcuda_writer *w) {
	{
		std::lock_guard<std::mutex> lk(w->mu);
		w->stop = true;
	}
	w->cv_work.notify_all();
	for (auto &t : w->team) t.join();
	w->team.clear();
}

// Write the blocks at the head of the line that are done; wait for more while more than `leave` blocks are in the line.
int harvest(srcuda_writer *w, size_t leave) {
	std::unique_lock<std::mutex> lk(w->mu);
	while (!w->jobs.empty() && (w->jobs.front()->done || w->jobs.size() > leave)) {
		srcuda_writer::Job *front = w->jobs.front().get();
		if (!front->done) {
			if (w->team.empty() || w->claim_seq == w->front_seq) {  // nobody has it: deflate it here
				if (w->claim_seq == w->front_seq) w->claim_seq++;
				lk.unlock();
				deflate_block(front->plain, w->level, front->out);
				lk.lock();
				front->done = true;
			} else {
				w->cv_done.wait(lk);
			}
			continue;
		}
		std::unique_ptr<srcuda_writer::Job> job = std::move(w->jobs.front());
		w->jobs.pop_front();
		w->front_seq++;
		lk.unlock();
		if (!job->out.ok) return w->fail(SRCUDA_ERR_FORMAT, "a block of %zu bytes does not fit a BGZF block", job->plain.size());
		int rc = write_bytes(w, job->out.bytes.data(), job->out.bytes.size());
		if (rc) return rc;
		lk.lock();
	}
	return SRCUDA_OK;
}

int drain(srcuda_writer *w) { return harvest(w, 0); }

int seal_block(srcuda_writer *w) {
	if (w->block.empty()) return SRCUDA_OK;
	std::unique_ptr<srcuda_writer::Job> job(new srcuda_writer::Job());
	job->plain.swap(w->block);
	w->block.reserve(kDeflateRoom);
	// The team starts with the first block of RECORDS, not with the header: a process that has ever started a thread pays for
	// locked stdio from then on, and the reference's index loader (millions of small fread calls, sorted-hashtable.c:1390)
	// runs between the header and the first record -- 2 s of 5 on the build container.
	if (w->header_done_writing) start_team(w);
	{
		std::lock_guard<std::mutex> lk(w->mu);
		w->jobs.push_back(std::move(job));
	}
	w->cv_work.notify_one();
	return harvest(w, w->team.empty() ? 0 : (size_t)w->threads * 8);
}

// Header text and reference list of a BAM file, cut into blocks the way SamBam_writer_write_header does (sambam-file.c:1248-1303):
// a text block ends at the first newline more than 55 000 bytes behind its start (or at the last byte), the first one carries
// the magic and the text length; the reference list is cut behind the entry that takes a block past 55 000 bytes.
int bam_write_header(srcuda_writer *w) {
	const std::string &t = w->header_text;
	size_t start = 0;
	for (size_t p = 0; p < t.size(); p++) {
		if ((p - start > kBlockCut || p + 1 >= t.size()) && t[p] == '\n') {
			if (start == 0) {
				w->block.append("BAM\1", 4);
				put32(w->block, (uint32_t)t.size());
			}
			w->block.append(t, start, p - start + 1);
			int rc = seal_block(w);
			if (rc) return rc;
			start = p + 1;
		}
	}
	size_t n = w->chromosomes.size();
	put32(w->block, (uint32_t)n);
	for (size_t i = 0; i < n; i++) {
		const Chromosome &c = w->chromosomes[i];
		put32(w->block, (uint32_t)c.name.size() + 1);
		w->block.append(c.name.c_str(), c.name.size() + 1);
		put32(w->block, c.length);
		if (i + 1 == n || w->block.size() > kBlockCut) {
			int rc = seal_block(w);
			if (rc) return rc;
		}
	}
	// (no chromosomes: the count stays in the buffer and leaves with the first records, as in the reference, :1281-1301)
	std::string().swap(w->header_text);
	return SRCUDA_OK;
}

int finish_header(srcuda_writer *w) {
	if (w->header_done) return SRCUDA_OK;
	w->header_done = true;
	int rc = w->format == SRCUDA_OUT_BAM ? bam_write_header(w) : SRCUDA_OK;
	w->header_done_writing = true;
	return rc;
}

// CIGAR text -> BAM operations (SamBam_compress_cigar, sambam-file.c:1426-1458): operation codes from "MIDNSHP=X", anything
// else lands on 8; reference bases covered = M + N + D (that is what the bin is computed from there).
int pack_cigar(const char *cigar, uint32_t *ops, int *covered) {
	*covered = 0;
	if (cigar[0] == '*') return 0;
	static const char codes[] = "MIDNSHP=X";
	int n = 0, cover = 0;
	uint32_t len = 0;
	for (const char *p = cigar; *p; p++) {
		unsigned char c = (unsigned char)*p;
		if (c >= '0' && c <= '9') {
			len = len * 10 + (c - '0');
			continue;
		}
		if (c == 'M' || c == 'N' || c == 'D') cover += (int)len;
		int code = 0;
		while (code < 8 && codes[code] != (char)c) code++;
		ops[n++] = (len << 4) | (uint32_t)code;
		len = 0;
		if (n >= kMaxCigarOps) break;
	}
	*covered = cover;
	return n;
}

// UCSC binning on ints as the reference computes it (SamBam_reg2bin, sambam-file.c:1584-1593), unmapped reads included (beg = -1).
int reg2bin(int beg, int end) {
	--end;
	if (beg >> 14 == end >> 14) return 4681 + (beg >> 14);
	if (beg >> 17 == end >> 17) return 585 + (beg >> 17);
	if (beg >> 20 == end >> 20) return 73 + (beg >> 20);
	if (beg >> 23 == end >> 23) return 9 + (beg >> 23);
	if (beg >> 26 == end >> 26) return 1 + (beg >> 26);
	return 0;
}

struct NibbleTable {
	unsigned char v[256];
	NibbleTable() {
		memset(v, 15, sizeof v);  // anything that is not a code letter is N (SamBam_read2bin, :1460-1476)
		const char *letters = "=ACMGRSVTWYHKDB";
		for (int i = 0; i < 15; i++) v[(unsigned char)letters[i]] = (unsigned char)i;
	}
};
const NibbleTable kNibble;

// Optional columns, text -> BAM (SamBam_compress_additional, sambam-file.c:1478-1582), bytes as the reference's: a column
// starts at offset 0 or behind a tab; 'i' is a 32-bit int, 'f' is stored as 0.0 (the reference converts into the wrong
// variable), 'Z' / 'H' are copied up to the 780-byte mark, 'A' one character, 'B' arrays of 32-bit ints ('i') or floats
// (every other cell type); columns of another type are passed over; nothing is started past 750 bytes.
int pack_additional(const char *a, unsigned char *bin) {
	int len = (int)strlen(a), at = 0, used = 0;
	while (at < len) {
		if (!(at == 0 || a[at] == '\t')) {
			at++;
			continue;
		}
		if (a[at] == '\t') at++;
		if (at + 3 >= len + 1) {  // a tag cut short: the reference reads the terminator and finds no type
			if (at == 0) return -1;
			continue;
		}
		unsigned char tag0 = (unsigned char)a[at], tag1 = (unsigned char)a[at + 1];
		char type = a[at + 3];
		if (type == 'i' || type == 'f') {
			int digits = 0;
			while (at + 5 + digits < len && a[at + 5 + digits] != '\t') digits++;
			int32_t v = (type == 'i' && at + 5 <= len) ? (int32_t)atoi(a + at + 5) : 0;
			bin[used] = tag0;
			bin[used + 1] = tag1;
			bin[used + 2] = (unsigned char)type;
			memcpy(bin + used + 3, &v, 4);
			used += 7;
			at += 5 + digits;
		} else if (type == 'Z' || type == 'H') {
			bin[used] = tag0;
			bin[used + 1] = tag1;
			bin[used + 2] = (unsigned char)type;
			used += 3;
			at += 5;
			int n = 0;
			while (at + n < len && a[at + n] != '\t') {
				bin[used + n] = (unsigned char)a[at + n];
				n++;
				if (used + n > 780) break;
			}
			bin[used + n] = 0;
			used += n + 1;
			at += n;
		} else if (type == 'A') {
			bin[used] = tag0;
			bin[used + 1] = tag1;
			bin[used + 2] = 'A';
			bin[used + 3] = (at + 5 <= len) ? (unsigned char)a[at + 5] : 0;
			used += 4;
			at += 6;
		} else if (type == 'B') {
			char cell_type = (at + 5 <= len) ? a[at + 5] : 0;
			bin[used] = tag0;
			bin[used + 1] = tag1;
			bin[used + 2] = 'B';
			bin[used + 3] = (unsigned char)cell_type;
			int count_at = used + 4;
			int32_t count = 0;
			used += 8;
			at += 7;
			int cell = at;
			for (;;) {
				char c = (at < len) ? a[at] : 0;
				if (c == ',' || c == '\t' || c == 0) {
					if (at - cell < 29 && at >= cell) {
						char text[30];
						memcpy(text, a + cell, (size_t)(at - cell));
						text[at - cell] = 0;
						int32_t iv = 0;
						float fv = 0;
						if (cell_type == 'i') iv = (int32_t)atoi(text);
						else fv = (float)atof(text);
						if (used < 780) {
							memcpy(bin + used, cell_type == 'i' ? (void *)&iv : (void *)&fv, 4);
							used += 4;
							count++;
						}
					}
					cell = at + 1;
				}
				if (c == '\t' || c == 0) break;
				at++;
			}
			memcpy(bin + count_at, &count, 4);
		} else if (at == 0) {
			return -1;  // the reference never leaves this column (its cursor does not move)
		}
		if (used > 750) break;
	}
	return used;
}

// The encoders below only read the writer (format, chromosome table): several threads may run them at once.  Their failures
// are returned as text.
int encode_error(std::string &err, int code, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	err = buf;
	return code;
}

int lookup_chromosome(const srcuda_writer *w, const char *name) {
	auto it = w->chromosome_id.find(name);
	return it == w->chromosome_id.end() ? -1 : it->second;
}

// One record in BAM form at the end of `b` (SamBam_writer_add_read, sambam-file.c:1704-1797).
int bam_encode(const srcuda_writer *w, const srcuda_sam_record *r, std::string &b, std::string &err) {
	uint32_t ops[kMaxCigarOps];
	int covered = 0;
	int n_ops = pack_cigar(r->cigar, ops, &covered);
	size_t name_len = strlen(r->name) + 1;
	unsigned char extra[1100];
	int extra_len = 0;
	if (r->additional) {
		if (strlen(r->additional) > 100000) return encode_error(err, SRCUDA_ERR_ARG, "optional columns of %s are too long", r->name);
		extra_len = pack_additional(r->additional, extra);
		if (extra_len < 0) return encode_error(err, SRCUDA_ERR_FORMAT, "optional columns of %s: unknown type in '%.20s'", r->name, r->additional);
	}
	int read_len = r->read_len;
	if (read_len < 0 || strlen(r->seq) < (size_t)read_len || strlen(r->qual) < (size_t)read_len)
		return encode_error(err, SRCUDA_ERR_ARG, "record %s: sequence or qualities shorter than read_len %d", r->name, read_len);
	uint32_t record_len = 32 + (uint32_t)name_len + 4 * (uint32_t)n_ops + (uint32_t)(read_len + 1) / 2 + (uint32_t)read_len + (uint32_t)extra_len;
	put32(b, record_len);
	int bin = reg2bin((int)(r->pos - 1), (int)(r->pos - 1) + covered);
	int ref_id = lookup_chromosome(w, r->chro);
	int next_ref = -1;
	if (r->next_chro[0] == '=') next_ref = ref_id;
	else if (r->next_chro[0] != '*') next_ref = lookup_chromosome(w, r->next_chro);
	put32(b, (uint32_t)ref_id);
	put32(b, r->pos - 1);
	put32(b, ((uint32_t)bin << 16) | ((uint32_t)r->mapping_quality << 8) | (uint32_t)name_len);
	put32(b, (r->flags << 16) | (uint32_t)n_ops);
	put32(b, (uint32_t)read_len);
	put32(b, (uint32_t)next_ref);
	put32(b, r->next_pos - 1);
	put32(b, (uint32_t)(int32_t)r->tlen);
	b.append(r->name, name_len);
	for (int i = 0; i < n_ops; i++) put32(b, ops[i]);
	size_t at = b.size();
	b.resize(at + (size_t)(read_len + 1) / 2 + (size_t)read_len);
	unsigned char *p = (unsigned char *)&b[at];
	const unsigned char *s = (const unsigned char *)r->seq;
	for (int i = 0; i + 1 < read_len; i += 2) p[i >> 1] = (unsigned char)((kNibble.v[s[i]] << 4) | kNibble.v[s[i + 1]]);
	if (read_len & 1) p[read_len >> 1] = (unsigned char)(kNibble.v[s[read_len - 1]] << 4);
	p += (read_len + 1) / 2;
	for (int i = 0; i < read_len; i++) p[i] = (unsigned char)(r->qual[i] - 33);
	b.append((const char *)extra, (size_t)extra_len);
	return SRCUDA_OK;
}

void append_number(std::string &b, long long v) {
	char t[24];
	int n = snprintf(t, sizeof t, "%lld", v);
	b.append(t, (size_t)n);
}

// The record line of core.c:2156-2164: "%s\t%d\t%s\t%u\t%d\t%s\t%s\t%u\t%lld\t%s\t%s%s%s\n".
void sam_encode(const srcuda_sam_record *r, std::string &b) {
	b.append(r->name).push_back('\t');
	append_number(b, (int)r->flags);
	b.push_back('\t');
	b.append(r->chro).push_back('\t');
	append_number(b, r->pos);
	b.push_back('\t');
	append_number(b, r->mapping_quality);
	b.push_back('\t');
	b.append(r->cigar).push_back('\t');
	b.append(r->next_chro).push_back('\t');
	append_number(b, r->next_pos);
	b.push_back('\t');
	append_number(b, r->tlen);
	b.push_back('\t');
	b.append(r->seq).push_back('\t');
	b.append(r->qual);
	if (r->additional && r->additional[0]) {
		b.push_back('\t');
		b.append(r->additional);
	}
	b.push_back('\n');
}

int sam_flush(srcuda_writer *w) {
	int rc = write_bytes(w, w->block.data(), w->block.size());
	w->block.clear();
	return rc;
}

int encode_record(const srcuda_writer *w, const srcuda_sam_record *r, std::string &b, std::string &err) {
	if (!r->name || !r->chro || !r->cigar || !r->next_chro || !r->seq || !r->qual) return encode_error(err, SRCUDA_ERR_ARG, "record with a NULL column");
	if (w->format == SRCUDA_OUT_SAM) {
		sam_encode(r, b);
		return SRCUDA_OK;
	}
	return bam_encode(w, r, b, err);
}

// Behind every committable record: a BAM block past the mark is sealed, a full text buffer is written.
int after_records(srcuda_writer *w, int committable) {
	if (w->format == SRCUDA_OUT_SAM) return w->block.size() > kSamBuffer ? sam_flush(w) : SRCUDA_OK;
	if (w->block.size() > kBlockCut && committable) return seal_block(w);
	return SRCUDA_OK;
}

int add_read(srcuda_writer *w, const srcuda_sam_record *r, int committable) {
	if (w->failed) return SRCUDA_ERR_IO;
	if (!w->header_done) {
		int rc = finish_header(w);
		if (rc) return rc;
	}
	std::string err;
	size_t before = w->block.size();
	int rc = encode_record(w, r, w->block, err);
	if (rc) {
		w->block.resize(before);
		return w->fail(rc, "%s", err.c_str());
	}
	return after_records(w, committable);
}

// ---- fragment level ---------------------------------------------------------------------------------------------------

constexpr int kFlagPaired = 0x1, kFlagProper = 0x2, kFlagUnmapped = 0x4, kFlagMateUnmapped = 0x8, kFlagReverse = 0x10,
              kFlagMateReverse = 0x20, kFlagFirst = 0x40, kFlagSecond = 0x80, kFlagSecondary = 0x100;  // subread.h:45-54

struct ComplementTable {
	unsigned char v[256];
	ComplementTable() {  // __converting_char_table, input-files.c:1111: A<->T, C<->G, U->A, all else N
		memset(v, 'N', sizeof v);
		v['A'] = 'T';
		v['T'] = 'A';
		v['C'] = 'G';
		v['G'] = 'C';
		v['U'] = 'A';
	}
};
const ComplementTable kComplement;

void reverse_complement(char *s, int n) {
	unsigned char *u = (unsigned char *)s;
	for (int i = 0, j = n - 1; i < j; i++, j--) {
		unsigned char a = kComplement.v[u[i]], b = kComplement.v[u[j]];
		u[i] = b;
		u[j] = a;
	}
	if (n & 1) u[n / 2] = kComplement.v[u[n / 2]];
}

void reverse_text(char *s, int n) {
	if (!s || !s[0]) return;
	for (int i = 0, j = n - 1; i < j; i++, j--) std::swap(s[i], s[j]);
}

int indel_bases(const char *cigar) {  // calc_edit_dist, core.c:1325-1348
	int total = 0;
	unsigned int len = 0;
	for (const char *p = cigar; *p; p++) {
		if (*p >= '0' && *p <= '9') {
			len = len * 10 + (unsigned)(*p - '0');
			continue;
		}
		if (*p == 'I' || *p == 'D') total += (int)len;
		len = 0;
	}
	return total;
}

}  // namespace

extern "C" {

int srcuda_out_open(const char *path, int format, int threads, int no_compression, srcuda_writer **out) {
	if (!out || (format != SRCUDA_OUT_SAM && format != SRCUDA_OUT_BAM) || threads < 1) {
		g_open_error = "srcuda_out_open: bad argument";
		return SRCUDA_ERR_ARG;
	}
	*out = nullptr;
	srcuda_writer *w = new (std::nothrow) srcuda_writer();
	if (!w) {
		g_open_error = "out of memory";
		return SRCUDA_ERR_NOMEM;
	}
	if (path) {
		w->fp = fopen(path, "wb");
		w->own_fp = true;
		if (!w->fp) {
			g_open_error = std::string("cannot create ") + path;
			delete w;
			return SRCUDA_ERR_IO;
		}
		setvbuf(w->fp, nullptr, _IOFBF, 1 << 20);
	} else {
		w->fp = stdout;
	}
	w->format = format;
	w->threads = threads > 64 ? 64 : threads;
	w->level = no_compression ? Z_NO_COMPRESSION : Z_BEST_SPEED;
	w->block.reserve(kDeflateRoom);
	*out = w;
	return SRCUDA_OK;
}

int srcuda_out_add_chromosome(srcuda_writer *w, const char *name, uint32_t length, int add_header) {
	if (!w || !name) return SRCUDA_ERR_ARG;
	if (w->header_done) return w->fail(SRCUDA_ERR_ARG, "the header is closed");
	if (w->chromosome_id.count(name)) return w->fail(SRCUDA_ERR_ARG, "chromosome %s given twice", name);
	int id = (int)w->chromosomes.size();
	w->chromosomes.push_back({name, length});
	w->chromosome_id.emplace(name, id);
	if (add_header) {
		char line[1000];
		snprintf(line, 999, "@SQ\tSN:%s\tLN:%u", name, length);
		return srcuda_out_add_header(w, line, 0);
	}
	return SRCUDA_OK;
}

int srcuda_out_add_header(srcuda_writer *w, const char *line, int add_chro) {
	if (!w || !line) return SRCUDA_ERR_ARG;
	if (w->header_done) return w->fail(SRCUDA_ERR_ARG, "the header is closed");
	if (w->format == SRCUDA_OUT_SAM) {
		w->block.append(line).push_back('\n');
	} else {
		w->header_text.append(line).push_back('\n');
	}
	if (add_chro && strncmp(line, "@SQ", 3) == 0) {  // SN: / LN: fields of the line (sambam-file.c:1375-1393)
		std::string name;
		long length = -1;
		bool have_name = false;
		const char *p = line;
		while (*p) {
			const char *e = strchr(p, '\t');
			size_t n = e ? (size_t)(e - p) : strlen(p);
			if (n >= 3 && memcmp(p, "SN:", 3) == 0) {
				name.assign(p + 3, n - 3);
				have_name = true;
			} else if (n >= 3 && memcmp(p, "LN:", 3) == 0) {
				length = atoi(std::string(p + 3, n - 3).c_str());
			}
			if (!e) break;
			p = e + 1;
		}
		if (have_name && length > 0) return srcuda_out_add_chromosome(w, name.c_str(), (uint32_t)length, 0);
	}
	return SRCUDA_OK;
}

int srcuda_out_finish_header(srcuda_writer *w) {
	if (!w) return SRCUDA_ERR_ARG;
	return finish_header(w);
}

int srcuda_out_add_read(srcuda_writer *w, const srcuda_sam_record *rec, int committable) {
	if (!w || !rec) return SRCUDA_ERR_ARG;
	return add_read(w, rec, committable);
}

int srcuda_out_add_reads(srcuda_writer *w, const srcuda_sam_record *recs, int64_t n, int records_per_fragment) {
	if (!w || (n > 0 && !recs) || n < 0 || (records_per_fragment != 1 && records_per_fragment != 2)) return SRCUDA_ERR_ARG;
	for (int64_t i = 0; i < n; i++) {
		int rc = add_read(w, recs + i, records_per_fragment == 1 || (i & 1));
		if (rc) return rc;
	}
	return SRCUDA_OK;
}

int srcuda_out_add_sam_lines(srcuda_writer *w, const char *text, size_t bytes, int records_per_fragment) {
	if (!w || (bytes && !text) || (records_per_fragment != 1 && records_per_fragment != 2)) return SRCUDA_ERR_ARG;
	std::string line;
	size_t at = 0;
	int64_t n = 0;
	while (at < bytes) {
		const char *nl = (const char *)memchr(text + at, '\n', bytes - at);
		if (!nl) return w->fail(SRCUDA_ERR_FORMAT, "the last SAM line has no newline");
		line.assign(text + at, (size_t)(nl - (text + at)));
		at = (size_t)(nl - text) + 1;
		// eleven mandatory fields cut in place; the rest of the line are the optional columns (sambam-file.c:1668-1702)
		char *f[12] = {nullptr};
		char *p = &line[0];
		int nf = 0;
		f[nf++] = p;
		while (nf < 12 && (p = strchr(p, '\t')) != nullptr) {
			*p++ = 0;
			f[nf++] = p;
		}
		if (nf < 11 || !f[10][0]) return w->fail(SRCUDA_ERR_FORMAT, "bad read format: %.60s", line.c_str());
		srcuda_sam_record r;
		r.name = f[0];
		r.flags = (uint32_t)atoi(f[1]);
		r.chro = f[2];
		r.pos = (uint32_t)atoi(f[3]);
		r.mapping_quality = atoi(f[4]);
		r.cigar = f[5];
		r.next_chro = f[6];
		r.next_pos = (uint32_t)atoi(f[7]);
		r.tlen = atoi(f[8]);
		r.seq = f[9];
		r.qual = f[10];
		r.read_len = (int32_t)strlen(f[9]);
		r.additional = (nf == 12 && f[11][0]) ? f[11] : nullptr;
		int rc = add_read(w, &r, records_per_fragment == 1 || (n & 1));
		if (rc) return rc;
		n++;
	}
	return SRCUDA_OK;
}

int srcuda_sam_calc_flags(const srcuda_sam_config *cfg, const srcuda_sam_mate *m1, const srcuda_sam_mate *m2, int is_second_read,
                          int current_location, int tlen, srcuda_sam_counters *counters) {
	if (!cfg || !m1) return -1;
	srcuda_sam_counters scratch;
	if (!counters) counters = &scratch;
	int flag = 0;
	bool have1 = m1->present != 0, have2 = m2 && m2->present;
	if (cfg->is_paired_end) {
		if (!m2) return -1;
		const srcuda_sam_mate *me = is_second_read ? m2 : m1, *mate = is_second_read ? m1 : m2;
		flag = kFlagPaired | (is_second_read ? kFlagSecond : kFlagFirst);
		if (!me->ok) flag |= kFlagUnmapped;
		else if (me->strand + is_second_read == 1) flag |= kFlagReverse;
		if (!mate->ok) flag |= kFlagMateUnmapped;
		else if (mate->strand + is_second_read != 1) flag |= kFlagMateReverse;
		if (have1 && have2) {
			// properly paired: one chromosome, TLEN inside [-d, -D], the mates on one strand of the fragment (the second mate's
			// strand is stored flipped) and in the order the library type asks for (core.c:1661-1690)
			bool proper = false, wrong_order = false;
			bool same_chro = m1->chro_id == m2->chro_id;
			bool tlen_ok = tlen >= cfg->minimum_pair_distance && tlen <= cfg->maximum_pair_distance;
			if (same_chro && tlen_ok && me->strand == mate->strand) {
				bool mate_right = mate->offset > me->offset, mate_left = mate->offset < me->offset, same = mate->offset == me->offset;
				if (cfg->is_first_read_reversed && !cfg->is_second_read_reversed) {
					if (me->strand == 0) {
						if ((is_second_read + (int)mate_right == 1) || same) proper = true;
						else wrong_order = true;
					}
				} else if (me->strand) {
					if ((is_second_read + (int)mate_left == 1) || same) proper = true;
					else wrong_order = true;
				} else {
					if ((is_second_read + (int)mate_right == 1) || same) proper = true;
					else wrong_order = true;
				}
			}
			if (proper) flag |= kFlagProper;
			else if (is_second_read) {
				if (!same_chro) counters->not_properly_pairs_different_chro++;
				else if (me->strand != mate->strand) counters->not_properly_different_strands++;
				else if (!tlen_ok) counters->not_properly_pairs_TLEN_wrong++;
				else if (wrong_order) counters->not_properly_pairs_wrong_arrangement++;
			}
		}
	} else {
		if (!m1->ok) flag |= kFlagUnmapped;
		else if (m1->strand) flag |= kFlagReverse;
	}
	if (current_location > 0 && ((have1 && !is_second_read) || (have2 && is_second_read))) flag |= kFlagSecondary;
	return flag;
}

int srcuda_sam_calc_tlen(int32_t offset1, const char *cigar1, int read_len_1, int32_t offset2, const char *cigar2, int read_len_2) {
	uint32_t head1 = (uint32_t)offset1, head2 = (uint32_t)offset2;
	if (head1 == head2) return read_len_1 > read_len_2 ? read_len_1 : read_len_2;
	// walk the CIGAR of the mate that starts first until one of its matched sections reaches the other mate's start
	bool second_first = head2 < head1;
	const char *cigar = second_first ? cigar2 : cigar1;
	uint32_t left = second_first ? head2 : head1, right = second_first ? head1 : head2;
	int len_right = second_first ? read_len_1 : read_len_2, len_left = second_first ? read_len_2 : read_len_1;
	uint32_t len = 0, chro_at = left, section_end = 0, read_at = 0;
	int tlen = -1;
	for (const char *p = cigar ? cigar : ""; *p; p++) {
		char c = *p;
		if (c <= 0) break;
		if (c >= '0' && c <= '9') {
			len = len * 10 + (uint32_t)(c - '0');
			continue;
		}
		if (c == 'M' || c == 'S') {
			chro_at += len;
			read_at += len;
			section_end = chro_at;
		}
		if (c == 'N' || c == 'B' || c == 'b' || c == 'n' || c == 'I' || c == 'D' || p[1] == 0) {
			if (c == 'N' || c == 'D') chro_at += len;
			if (section_end >= right) {
				tlen = (int)(read_at + right - section_end + (uint32_t)len_right);
				break;
			}
		}
		if (c == 'I') read_at += len;
		if (c == 'B' || c == 'b' || c == 'n') break;
		len = 0;
	}
	if (tlen < 0) tlen = (int)(right - section_end + (uint32_t)len_right + (uint32_t)len_left);
	return tlen;
}

}  // extern "C"

namespace {

// write_single_fragment (core.c:1888-2170) up to the encoded records, appended to `dst`.  Reads the writer only.
int build_fragment(const srcuda_writer *w, const srcuda_sam_config *cfg, srcuda_sam_mate *m1, srcuda_sam_mate *m2, int all_locations,
                   int current_location, srcuda_sam_counters *counters, std::string &dst, std::string &err) {
	bool paired = cfg->is_paired_end != 0;
	if (paired && !m2) return encode_error(err, SRCUDA_ERR_ARG, "paired-end output needs both mates");
	if (!paired) m2 = nullptr;
	srcuda_sam_counters scratch;
	memset(&scratch, 0, sizeof scratch);
	if (!counters) counters = &scratch;
	srcuda_sam_mate *mates[2] = {m1, m2};
	int n_mates = paired ? 2 : 1;
	for (int i = 0; i < n_mates; i++) {
		srcuda_sam_mate *m = mates[i];
		if (!m->name || !m->text || !m->qual || m->read_len < 0) return encode_error(err, SRCUDA_ERR_ARG, "mate %d: NULL read columns", i + 1);
		if (m->ok && (!m->cigar || m->chro_id < 0 || (size_t)m->chro_id >= w->chromosomes.size()))
			return encode_error(err, SRCUDA_ERR_ARG, "mate %d: mapped without a CIGAR or a chromosome of the table", i + 1);
	}
	const char *rg = cfg->read_group_id ? cfg->read_group_id : "";
	if (strlen(rg) > 300) return encode_error(err, SRCUDA_ERR_ARG, "read group id longer than 300 characters");
	bool ok1 = m1->ok != 0, ok2 = m2 && m2->ok;

	int tlen = 0;
	if (ok1 && ok2 && m1->chro_id == m2->chro_id) tlen = srcuda_sam_calc_tlen(m1->offset, m1->cigar, m1->read_len, m2->offset, m2->cigar, m2->read_len);
	int flags[2];
	flags[0] = srcuda_sam_calc_flags(cfg, m1, m2, 0, current_location, tlen, counters);
	flags[1] = -1;
	if (paired) {
		flags[1] = srcuda_sam_calc_flags(cfg, m1, m2, 1, current_location, tlen, counters);
		if (current_location == 0 && (flags[1] & kFlagProper)) counters->all_correct_PE_reads++;
	}
	// a read is printed on the strand of its record; a mate without a record: first as read, second reversed (core.c:1791-1803)
	for (int i = 0; i < n_mates; i++) {
		srcuda_sam_mate *m = mates[i];
		int reverse = m->present ? m->strand : i;
		if (reverse) {
			reverse_complement(m->text, m->read_len);
			reverse_text(m->qual, m->read_len);
		}
		if (char *slash = strchr(m->name, '/')) *slash = 0;
	}
	if (!m1->qual[0]) {  // no qualities in the input: 'I' for every base of both mates (core.c:2011-2018)
		for (int i = 0; i < n_mates; i++) {
			size_t n = strlen(mates[i]->text);
			memset(mates[i]->qual, 'I', n);
			mates[i]->qual[n] = 0;
		}
	}
	std::string columns[2];
	for (int i = 0; i < n_mates; i++) {
		srcuda_sam_mate *m = mates[i];
		std::string &c = columns[i];
		if (ok1 || ok2) {  // HI and NH unless both ends are unmapped
			c += "HI:i:";
			append_number(c, current_location + 1);
			c += "\tNH:i:";
			append_number(c, all_locations);
		}
		if (rg[0]) c.append("\tRG:Z:").append(rg);
		if (m->ok) {
			if (m->additional) c += m->additional;
			c += "\tNM:i:";
			append_number(c, (short)(m->mismatched_bases + indel_bases(m->cigar)));
		}
	}
	long long tlens[2] = {tlen, tlen};
	if (ok1 && ok2) {  // the leftmost mate carries the positive length; a tie goes by the first mate's strand (core.c:2084-2092)
		if (m1->offset > m2->offset) tlens[0] = -tlens[0];
		else if (m2->offset > m1->offset) tlens[1] = -tlens[1];
		else if (m1->strand) tlens[0] = -tlens[0];
		else tlens[1] = -tlens[1];
	}
	if (current_location == 0 && (paired ? (ok1 || ok2) : ok1)) counters->all_mapped_reads++;

	srcuda_sam_record rec[2];
	for (int i = 0; i < n_mates; i++) {
		srcuda_sam_mate *m = mates[i];
		srcuda_sam_record &r = rec[i];
		r.name = m->name;
		r.flags = (uint32_t)flags[i];
		r.chro = m->ok ? w->chromosomes[(size_t)m->chro_id].name.c_str() : "*";
		r.pos = m->ok ? (uint32_t)(m->offset > 1 ? m->offset : 1) : 0;
		r.mapping_quality = m->ok ? m->mapping_quality : 0;
		r.cigar = m->ok ? m->cigar : "*";
		r.tlen = tlens[i];
		r.read_len = m->read_len;
		r.seq = m->text;
		r.qual = m->qual;
		r.additional = columns[i].c_str();
	}
	bool one_chromosome = ok1 && ok2 && m1->chro_id == m2->chro_id;
	rec[0].next_chro = one_chromosome ? "=" : (paired ? rec[1].chro : "*");
	rec[0].next_pos = paired ? rec[1].pos : 0;
	if (paired) {
		rec[1].next_chro = one_chromosome ? "=" : rec[0].chro;
		rec[1].next_pos = rec[0].pos;
	}
	size_t before = dst.size();
	int rc = encode_record(w, &rec[0], dst, err);
	if (rc == SRCUDA_OK && paired) rc = encode_record(w, &rec[1], dst, err);
	if (rc) dst.resize(before);
	return rc;
}

}  // namespace

extern "C" {

int srcuda_out_add_fragment(srcuda_writer *w, const srcuda_sam_config *cfg, srcuda_sam_mate *m1, srcuda_sam_mate *m2,
                            int all_locations, int current_location, srcuda_sam_counters *counters) {
	if (!w || !cfg || !m1) return SRCUDA_ERR_ARG;
	if (w->failed) return SRCUDA_ERR_IO;
	if (!w->header_done) {
		int rc = finish_header(w);
		if (rc) return rc;
	}
	std::string err;
	int rc = build_fragment(w, cfg, m1, m2, all_locations, current_location, counters, w->block, err);
	if (rc) return w->fail(rc, "%s", err.c_str());
	return after_records(w, 1);
}

int srcuda_out_format_fragment(const srcuda_writer *w, const srcuda_sam_config *cfg, srcuda_sam_mate *m1, srcuda_sam_mate *m2,
                               int all_locations, int current_location, srcuda_sam_counters *counters, void *buffer, size_t capacity,
                               size_t *bytes) {
	if (!w || !cfg || !m1 || !bytes || (capacity && !buffer)) return SRCUDA_ERR_ARG;
	*bytes = 0;
	if (!w->header_done) {
		g_open_error = "srcuda_out_format_fragment: the header is still open (srcuda_out_finish_header first)";
		return SRCUDA_ERR_ARG;
	}
	static thread_local std::string scratch;
	scratch.clear();
	std::string err;
	int rc = build_fragment(w, cfg, m1, m2, all_locations, current_location, counters, scratch, err);
	if (rc) {
		g_open_error = err;
		return rc;
	}
	*bytes = scratch.size();
	if (scratch.size() > capacity) {
		g_open_error = "srcuda_out_format_fragment: the buffer is too small";
		return SRCUDA_ERR_NOMEM;
	}
	memcpy(buffer, scratch.data(), scratch.size());
	return SRCUDA_OK;
}

int srcuda_out_append(srcuda_writer *w, const void *encoded, size_t bytes, int committable) {
	if (!w || (bytes && !encoded)) return SRCUDA_ERR_ARG;
	if (w->failed) return SRCUDA_ERR_IO;
	if (!w->header_done) {
		int rc = finish_header(w);
		if (rc) return rc;
	}
	w->block.append((const char *)encoded, bytes);
	return after_records(w, committable);
}

int srcuda_out_flush(srcuda_writer *w) {
	if (!w) return SRCUDA_ERR_ARG;
	if (w->failed) return SRCUDA_ERR_IO;
	int rc = finish_header(w);
	if (rc) return rc;
	if (w->format == SRCUDA_OUT_SAM) rc = sam_flush(w);
	else {
		rc = seal_block(w);
		if (rc == SRCUDA_OK) rc = drain(w);
	}
	if (rc == SRCUDA_OK && fflush(w->fp) != 0) rc = w->fail(SRCUDA_ERR_IO, "flush failed");
	return rc;
}

int srcuda_out_close(srcuda_writer *w, uint64_t *bytes_written) {
	if (!w) return SRCUDA_ERR_ARG;
	int rc = w->failed ? SRCUDA_ERR_IO : srcuda_out_flush(w);
	if (rc == SRCUDA_OK && w->format == SRCUDA_OUT_BAM) {  // the empty block every BAM file ends with (sambam-file.c:1203-1209)
		static const unsigned char eof_block[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
		                                            0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
		rc = write_bytes(w, eof_block, sizeof eof_block);
	}
	if (w->own_fp) {
		if (fclose(w->fp) != 0 && rc == SRCUDA_OK) rc = SRCUDA_ERR_IO;
	} else {
		fflush(w->fp);
	}
	stop_team(w);
	if (bytes_written) *bytes_written = w->bytes_written;
	if (rc != SRCUDA_OK) g_open_error = w->error.empty() ? "write failed" : w->error;
	delete w;
	return rc;
}

const char *srcuda_out_error(const srcuda_writer *w) { return w ? w->error.c_str() : g_open_error.c_str(); }

}  // extern "C"
