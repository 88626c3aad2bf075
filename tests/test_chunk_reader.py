"""Chunk reader: srcuda_reader_* (subread_b200/csrc/sr_reader.cpp) hands out the fragments
geinput_next_read_trim / fetch_next_read_pair hand out, batch by batch, from plain and gzip files, for any window size
and thread count.  Host code only, CPU only."""
import gzip
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN
from subread_b200 import load_fastq
from subread_b200 import binding
from test_fastq_split import ref, reference_records  # noqa: F401  (the unmodified reference's reader)


def read_all(path1, path2=None, batch=1000, **kw):
    rd = binding.ChunkReader(path1, path2, **kw)
    reads, sizes = [], []
    while True:
        text, off, ln, n = rd.next(batch)
        if n == 0:
            break
        assert n <= batch and off[0] == 0 and off.shape[0] == n * rd.mates + 1
        assert np.array_equal(np.diff(off.astype(np.int64)), ln.astype(np.int64))      # dense: reads back to back
        reads += [text[int(o):int(o) + int(k)].tobytes() for o, k in zip(off[:-1], ln)]
        sizes.append(n)
    rd.close()
    return reads, sizes


def chunk_reads(chunk):
    return [chunk.text[int(chunk.off[i]):int(chunk.off[i + 1])].tobytes() for i in range(chunk.off.shape[0] - 1)]


@pytest.mark.parametrize("window,threads,batch", [(0, 1, 1000), (4096, 1, 7), (5000, 3, 333), (1 << 20, 4, 100000)])
def test_golden_files_plain_and_gzip(window, threads, batch, tmp_path):
    for r1, r2 in (("se_R1.fastq.gz", None), ("pe_R1.fastq.gz", "pe_R2.fastq.gz")):
        p1, p2 = os.path.join(GOLDEN, r1), os.path.join(GOLDEN, r2) if r2 else None
        expect = chunk_reads(load_fastq(p1, p2))
        got, sizes = read_all(p1, p2, batch, window_bytes=window, threads=threads)                  # gzip
        assert got == expect
        mates = 2 if r2 else 1
        assert all(s == batch for s in sizes[:-1]) and sum(sizes) * mates == len(expect)           # full batches but the last
        plain = []
        for p in (p1, p2):
            if p:
                q = str(tmp_path / (os.path.basename(p)[:-3]))
                with gzip.open(p, "rb") as a, open(q, "wb") as b:
                    b.write(a.read())
                plain.append(q)
        got, _ = read_all(plain[0], plain[1] if r2 else None, batch, window_bytes=window, threads=threads)
        assert got == expect


def test_reader_against_the_reference_reader(ref, tmp_path):  # noqa: F811
    rnd = random.Random(4)
    recs = []
    for i in range(5000):
        n = rnd.randrange(1, 260)
        recs.append(b"@r%d comment\n%s\n+\n%s\n" % (i, bytes(rnd.choice(b"ACGTNacgt") for _ in range(n)), bytes(rnd.choice(b"IH#@+5") for _ in range(n))))
    recs.insert(2500, b"@long\n" + b"ACGT" * 400 + b"\n+\n" + b"I" * 1600 + b"\n")       # 1600 bases: cut to 1209
    path = str(tmp_path / "x.fq")
    with open(path, "wb") as fh:
        fh.write(b"\n".join(recs))                       # blank line between records, none at the end
    for t5, t3 in ((0, 0), (3, 2)):
        exp = [e[1] for e in reference_records(ref, path, t5, t3)]
        for window, threads in ((0, 1), (4096, 2), (100000, 4)):
            got, _ = read_all(path, None, 999, window_bytes=window, threads=threads, trim_5=t5, trim_3=t3)
            assert got == exp, (t5, t3, window, threads)


def test_reader_errors(tmp_path):
    with pytest.raises(binding.SubreadCudaError, match="cannot open"):
        binding.ChunkReader(str(tmp_path / "missing.fq"))
    a, b = str(tmp_path / "a.fq"), str(tmp_path / "b.fq")
    with open(a, "wb") as fh:
        fh.write(b"@1\nACGT\n+\nIIII\n@2\nACGT\n+\nIIII\n")
    with open(b, "wb") as fh:
        fh.write(b"@1\nACGT\n+\nIIII\n")
    rd = binding.ChunkReader(a, b)
    with pytest.raises(binding.SubreadCudaError, match="different amounts of reads"):
        while rd.next(10)[3]:
            pass
    rd.close()
    with open(b, "wb") as fh:
        fh.write(b"@1\nACGT\n+\nIIII\n@2\nACGT\nIIII\n@3\nAC\n+\nII\n")
    rd = binding.ChunkReader(b)
    text, off, ln, n = rd.next(1)
    assert n == 1 and text.tobytes() == b"ACGT"
    with pytest.raises(binding.SubreadCudaError, match="format issue at byte 23"):
        rd.next(10)
    rd.close()
    empty = str(tmp_path / "empty.fq")
    open(empty, "wb").close()
    rd = binding.ChunkReader(empty)
    assert rd.next(5)[3] == 0
    rd.close()
    # a buffer that cannot hold the batch is an argument error, not an overrun
    rd = binding.ChunkReader(a)
    with pytest.raises(binding.SubreadCudaError, match="too small"):
        rd.next(2, max_read_len=1)
    rd.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_ranks_reading_their_slices_cover_the_file(world):
    """Read sharding: every rank opens the same files, and per chunk gathers only its own contiguous
    slice (subread_b200.sharding.shard_range) and skips the others'; together the ranks see every fragment once, in order."""
    from subread_b200.sharding import shard_range
    p1, p2 = os.path.join(GOLDEN, "pe_R1.fastq.gz"), os.path.join(GOLDEN, "pe_R2.fastq.gz")
    whole = load_fastq(p1, p2)
    expect = chunk_reads(whole)
    chunk_frags = 300                                   # stands for reads_per_chunk
    per_rank = []
    for rank in range(world):
        rd = binding.ChunkReader(p1, p2, window_bytes=20000, threads=2)
        mine, f0 = [], 0
        while f0 < whole.n_frags:
            n_chunk = min(chunk_frags, whole.n_frags - f0)
            lo, hi = shard_range(n_chunk, world, rank)
            assert rd.skip(lo) == lo
            text, off, ln, n = rd.next(hi - lo)
            assert n == hi - lo
            mine.append((f0 + lo, [text[int(o):int(o) + int(k)].tobytes() for o, k in zip(off[:-1], ln)]))
            assert rd.skip(n_chunk - hi) == n_chunk - hi
            f0 += n_chunk
        assert rd.skip(10) == 0 and rd.next(10)[3] == 0        # the end
        rd.close()
        per_rank.append(mine)
    got = [None] * len(expect)
    for mine in per_rank:
        for first, reads in mine:
            for i, rtext in enumerate(reads):
                assert got[2 * first + i] is None
                got[2 * first + i] = rtext
    assert got == expect


@pytest.mark.timeout(120)
@pytest.mark.parametrize("threads", [1, 4])
def test_damaged_gzip_is_an_error_not_a_hang(threads, tmp_path):
    """A gzip file cut short or corrupted in the middle: the reader (and its read-ahead thread) reports a read error or the
    records that were intact followed by a format issue -- it never hangs and never hands out garbage silently."""
    rnd = random.Random(8)
    recs = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rnd.choice(b"ACGT") for _ in range(100)), b"I" * 100) for i in range(20000))
    blob = gzip.compress(recs, compresslevel=1)
    cut = str(tmp_path / "cut.fq.gz")
    with open(cut, "wb") as fh:
        fh.write(blob[:len(blob) // 2])
    bad = bytearray(blob)
    for k in range(len(bad) // 2, len(bad) // 2 + 64):
        bad[k] ^= 0x5A
    broken = str(tmp_path / "broken.fq.gz")
    with open(broken, "wb") as fh:
        fh.write(bytes(bad))
    for path in (cut, broken):
        rd = binding.ChunkReader(path, threads=threads, window_bytes=1 << 20)
        got = 0
        with pytest.raises(binding.SubreadCudaError):
            while True:
                n = rd.next(5000)[3]
                if n == 0:
                    break
                got += n
        assert got < 20000
        rd.close()
    # closing a reader whose read-ahead thread still holds an unread block
    rd = binding.ChunkReader(os.path.join(GOLDEN, "pe_R1.fastq.gz"), os.path.join(GOLDEN, "pe_R2.fastq.gz"), threads=threads)
    assert rd.next(10)[3] == 10
    rd.close()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("threads", [1, 4])
@pytest.mark.parametrize("where", ["record_boundary", "inside_quality"])
def test_gzip_cut_at_a_flush_point_is_an_error(threads, where, tmp_path):
    """A gzip stream that simply stops -- at a Z_FULL_FLUSH point, so that inflate sees no damage -- on a record boundary or
    inside a quality line.  gzread answers that with 0 and Z_BUF_ERROR, not -1: the reader must report it instead of taking it
    for the end of the input, and the records in front of the cut are handed out before the error (none is dropped)."""
    import zlib
    rnd = random.Random(9)
    rec = [b"@r%d\n%s\n+\n%s\n" % (i, bytes(rnd.choice(b"ACGT") for _ in range(100)), b"I" * 100) for i in range(20000)]
    head = b"".join(rec[:10000])
    if where == "inside_quality":
        head += rec[10000][:-40]                        # 61 of the 100 quality characters, no newline
    co = zlib.compressobj(1, zlib.DEFLATED, 31)         # gzip container
    blob = co.compress(head) + co.flush(zlib.Z_FULL_FLUSH)
    path = str(tmp_path / "flushcut.fq.gz")
    with open(path, "wb") as fh:
        fh.write(blob)                                  # no trailer: the file ends at the flush point
    rd = binding.ChunkReader(path, threads=threads, window_bytes=1 << 20)
    got = 0
    with pytest.raises(binding.SubreadCudaError):
        while True:
            n = rd.next(3000)[3]
            if n == 0:
                break
            got += n
    rd.close()
    assert got <= 10000
    # the intact file for comparison: all 20000, no error
    full = str(tmp_path / "full.fq.gz")
    with open(full, "wb") as fh:
        fh.write(gzip.compress(b"".join(rec), compresslevel=1))
    reads, _ = read_all(full, batch=3000, threads=threads, window_bytes=1 << 20)
    assert len(reads) == 20000


@pytest.mark.parametrize("window,threads", [(4096, 1), (1 << 16, 3)])
def test_tell_points_behind_the_last_fragment_handed_out(window, threads, tmp_path):
    """srcuda_reader_tell: after every batch the offsets are where a second reader -- the reference's gene_input_t under the
    binding -- has to continue: parsing the files from there yields exactly the fragments not handed out yet."""
    rnd = random.Random(12)
    recs1, recs2 = [], []
    for i in range(3000):
        n = 20 + (i * 13) % 140
        recs1.append(b"@a%d\n%s\n+\n%s\n%s" % (i, bytes(rnd.choice(b"ACGT") for _ in range(n)), b"I" * n, b"\n" if i % 17 == 0 else b""))
        recs2.append(b"@b%d/2 x\n%s\n+b%d\n%s\n" % (i, bytes(rnd.choice(b"ACGT") for _ in range(n + 3)), i, b"H" * (n + 3)))
    p1, p2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    d1, d2 = b"".join(recs1), b"".join(recs2)
    open(p1, "wb").write(d1)
    open(p2, "wb").write(d2)
    rd = binding.ChunkReader(p1, p2, window_bytes=window, threads=threads)
    assert rd.tell() == (0, 0)
    done = 0
    for batch in (1, 7, 500, 123, 1000, 5000):
        n = rd.next(batch)[3]
        done += n
        a, b = rd.tell()
        for data, pos, tag in ((d1, a, b"@a"), (d2, b, b"@b")):
            rest = data[pos:]
            if done < 3000:
                assert rest.startswith(tag + str(done).encode()), (batch, done, rest[:20])
            else:
                assert rest.strip(b"\n") == b""
            left = binding.fastq_split(np.frombuffer(rest, dtype=np.uint8))[2].n_reads if rest else 0
            assert left == 3000 - done
    assert done == 3000
    rd.close()


def test_next_full_hands_out_names_and_qualities(tmp_path):
    """srcuda_reader_next_full: the batch of srcuda_reader_next plus the names (cut at the first blank) and the quality lines
    (trimmed like the reads) of its reads, for a paired gzip input read in several batches with small windows."""
    import ctypes
    import gzip
    import random
    from subread_b200 import binding
    rnd = random.Random(5)
    recs = [[], []]
    for m in range(2):
        for i in range(5000):
            n = rnd.randrange(30, 160)
            recs[m].append((b"frag%d/%d some comment" % (i, m + 1) if i % 3 else b"frag%d" % i, bytes(rnd.choice(b"ACGTN") for _ in range(n)),
                            bytes(33 + rnd.randrange(0, 41) for _ in range(n))))
        with gzip.open(str(tmp_path / f"r{m + 1}.fq.gz"), "wb") as fh:
            for name, seq, qual in recs[m]:
                fh.write(b"@" + name + b"\n" + seq + b"\n+\n" + qual + b"\n")
    L = binding.load_library()
    rd = ctypes.c_void_p()
    trim5, trim3 = 2, 3
    assert L.srcuda_reader_open(str(tmp_path / "r1.fq.gz").encode(), str(tmp_path / "r2.fq.gz").encode(), 1 << 16, 3, trim5, trim3, ctypes.byref(rd)) == 0
    seen = 0
    while True:
        want = 1700
        text = ctypes.create_string_buffer(want * 2 * 200)
        names = ctypes.create_string_buffer(want * 2 * 200)
        quals = ctypes.create_string_buffer(want * 2 * 200)
        off, noff, qoff = ((ctypes.c_uint64 * (want * 2))() for _ in range(3))
        ln, nlen, qlen = ((ctypes.c_uint16 * (want * 2))() for _ in range(3))
        got, used = ctypes.c_int64(), ctypes.c_uint64()
        rc = L.srcuda_reader_next_full(rd, want, text, len(text), off, ln, names, len(names), noff, nlen, quals, len(quals), qoff, qlen,
                                       ctypes.byref(got), ctypes.byref(used))
        assert rc == 0, L.srcuda_reader_error(rd)
        if got.value == 0:
            break
        for f in range(got.value):
            for m in range(2):
                k = 2 * f + m
                name, seq, qual = recs[m][seen + f]
                assert names.raw[noff[k]:noff[k] + nlen[k]] == name.split(b" ")[0]
                assert text.raw[off[k]:off[k] + ln[k]] == seq[trim5:len(seq) - trim3]
                assert quals.raw[qoff[k]:qoff[k] + qlen[k]] == qual[trim5:len(qual) - trim3]
        seen += got.value
    assert seen == 5000
    L.srcuda_reader_close(rd)
