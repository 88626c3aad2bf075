"""FASTQ batches: srcuda_fastq_split / srcuda_reads_pack (subread_b200/csrc/sr_fastq.cpp)
against the UNMODIFIED reference's geinput_open + geinput_next_read_trim (input-files.c:455,982-1093) from
oracle/_ref/libsubread_ref.so, on well-formed and on awkward files, with trimming, and fed in pieces.  Host code
only: no device work, CPU only."""
import ctypes
import gzip
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, REF_DIR, have_ref
from subread_b200 import load_fastq, load_fastq_indexed
from subread_b200 import binding


def split_whole(data: bytes, t5=0, t3=0):
    buf = np.frombuffer(data, dtype=np.uint8)
    off, ln, b, ex = binding.fastq_split(buf, True, t5, t3, want_names=True)
    recs = []
    for i in range(off.shape[0]):
        take = lambda o, n: data[int(o):int(o) + int(n)]  # noqa: E731
        recs.append((take(ex["name_off"][i], ex["name_len"][i]), take(off[i], ln[i]), take(ex["qual_off"][i], ex["qual_len"][i])))
    return recs, b


def split_in_pieces(data: bytes, piece: int, t5=0, t3=0, max_reads=None):
    """Feeds the input as a streaming caller does: a window that grows by `piece` bytes, consumed bytes dropped."""
    recs, pos, end = [], 0, 0
    while True:
        end = min(len(data), max(end, pos) + piece)
        at_eof = end == len(data)
        window = data[pos:end]
        buf = np.frombuffer(window, dtype=np.uint8)
        off, ln, b, _ = binding.fastq_split(buf, at_eof, t5, t3, max_reads=max_reads)
        recs += [window[int(o):int(o) + int(n)] for o, n in zip(off, ln)]
        pos += int(b.consumed)
        if b.stop == binding.FASTQ_END or (at_eof and b.stop == binding.FASTQ_MORE):
            return recs
        assert b.stop in (binding.FASTQ_MORE, binding.FASTQ_FULL)


@pytest.fixture(scope="module")
def ref():
    if not have_ref():
        pytest.skip("oracle/_ref (reference binaries) not built")
    L = ctypes.CDLL(os.path.join(REF_DIR, "libsubread_ref.so"))
    L.geinput_open.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    L.geinput_next_read_trim.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_short,
                                         ctypes.c_short, ctypes.POINTER(ctypes.c_int)]
    L.geinput_close.argtypes = [ctypes.c_void_p]
    return L


def reference_records(ref, path, t5=0, t3=0):
    """Every (name, read, quality) geinput_next_read_trim returns until it reports a length <= 0, where
    fetch_next_read_pair ends the input (core.c:1168)."""
    ginp = ctypes.create_string_buffer(1 << 20)        # gene_input_t, generously
    assert ref.geinput_open(path.encode(), ginp) == 0
    name, seq, qual = (ctypes.create_string_buffer(4096) for _ in range(3))
    sec = ctypes.c_int(0)
    out = []
    while True:
        n = ref.geinput_next_read_trim(ginp, name, seq, qual, t5, t3, ctypes.byref(sec))
        if n <= 0:
            break
        assert len(seq.value) == n
        out.append((name.value, seq.value, qual.value))
    ref.geinput_close(ginp)
    return out


def awkward_files():
    rnd = random.Random(77)
    dna = lambda n: bytes(rnd.choice(b"ACGTNacgt.") for _ in range(n))  # noqa: E731
    q = lambda n: bytes(rnd.randrange(33, 74) for _ in range(n))  # noqa: E731
    plain = b"".join(b"@r%d extra words\n%s\n+\n%s\n" % (i, dna(50 + i % 30), q(50 + i % 30)) for i in range(200))
    yield "plain", plain
    yield "no_final_newline", plain[:-1]
    yield "blank_lines", b"\n\n" + plain.replace(b"\n+\n", b"\n\n+\n\n", 40).replace(b"\n@r7", b"\n\n\n@r7") + b"\n\n\n"
    yield "crlf", plain.replace(b"\n", b"\r\n")
    yield "tab_names_and_plus_repeat", b"".join(b"@n%d\tx y\n%s\n+n%d comment\n%s\n" % (i, dna(40), i, q(40)) for i in range(50))
    yield "name_starts_with_blank", b"@ lead\nACGT\n+\nIIII\n@\tx\nAC\n+\nII\n"
    yield "long_name_and_long_read", (b"@" + b"N" * 450 + b"\n" + dna(1500) + b"\n+\n" + q(1500) + b"\n@b\n" + dna(1209) + b"\n+\n" + q(1209)
                                      + b"\n@c\n" + dna(1210) + b"\n+\n" + q(1210) + b"\n")
    yield "short_quality", b"@a\nACGTACGTAC\n+\nIII\n@b\nACGTAC\n+\nIIIIIIIII\n"
    yield "quality_starts_with_at", b"@a\nACGT\n+\n@III\n@b\nGGCC\n+\n@@@@\n"
    yield "single_record", b"@only\nACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIII"
    yield "mixed_lengths", b"".join(b"@m%d\n%s\n+\n%s\n" % (i, dna(n), q(n)) for i, n in enumerate([1, 2, 15, 16, 17, 100, 300, 1, 1000, 5]))


@pytest.mark.parametrize("trim", [(0, 0), (3, 0), (0, 4), (5, 7), (16, 1)], ids=lambda t: f"trim{t[0]}_{t[1]}")
def test_split_matches_geinput_next_read_trim(ref, trim, tmp_path):
    t5, t3 = trim
    for name, data in awkward_files():
        path = str(tmp_path / f"{name}.fq")
        with open(path, "wb") as fh:
            fh.write(data)
        exp = reference_records(ref, path, t5, t3)
        got, b = split_whole(data, t5, t3)
        assert len(got) == len(exp), f"{name}: {len(got)} records, the reference reads {len(exp)}"
        for i, (g, e) in enumerate(zip(got, exp)):
            assert g[0] == e[0], f"{name}: name of record {i}"
            assert g[1] == e[1], f"{name}: read {i}"
            if name != "short_quality" or (t5, t3) == (0, 0):      # a quality line shorter than trim_5 leaves the reference's buffer stale
                assert g[2] == e[2], f"{name}: quality of record {i}"
        assert b.stop == binding.FASTQ_END
        assert b.max_len == max((len(e[1]) for e in exp), default=0)


def test_broken_records_stop_where_the_reference_stops(ref, tmp_path):
    cases = {
        "no_at": (b"@a\nACGT\n+\nIIII\nACGT\n+\nIIII\n", 1),
        "no_plus": (b"@a\nACGT\n+\nIIII\n@b\nACGT\nIIII\n@c\nAC\n+\nII\n", 1),
        "cut_after_read": (b"@a\nACGT\n+\nIIII\n@b\nACGT\n", 1),
        "fasta_like": (b"@a\nACGT\n+\nIIII\n>b\nACGT\n", 1),
    }
    for name, (data, n_good) in cases.items():
        path = str(tmp_path / f"{name}.fq")
        with open(path, "wb") as fh:
            fh.write(data)
        assert len(reference_records(ref, path)) == n_good, name
        buf = np.frombuffer(data, dtype=np.uint8)
        with pytest.raises(binding.SubreadCudaError, match=f"{n_good} records before it"):
            binding.fastq_split(buf)
    # a record cut short at the very end: the reference returns what it has (an empty quality line)
    data = b"@a\nACGT\n+\nIIII\n@b\nACGT\n+\n"
    path = str(tmp_path / "cut_after_plus.fq")
    with open(path, "wb") as fh:
        fh.write(data)
    exp = reference_records(ref, path)
    got, _ = split_whole(data)
    assert [g[1] for g in got] == [e[1] for e in exp] == [b"ACGT", b"ACGT"]


@pytest.mark.parametrize("piece", [1, 7, 64, 1000, 1 << 20])
def test_pieces_give_the_records_of_the_whole(piece):
    for name, data in awkward_files():
        if piece < 64 and len(data) > 6000 and name in ("plain", "no_final_newline", "blank_lines", "crlf"):
            data = data[:data.rfind(b"\n@r", 0, 6000) + 1]      # whole records only
        whole = [r[1] for r in split_whole(data)[0]]
        assert split_in_pieces(data, piece) == whole, f"{name}, pieces of {piece}"
        assert split_in_pieces(data, max(piece, 64), max_reads=3) == whole, f"{name}, at most 3 records per call"


def test_empty_and_degenerate_inputs():
    for data in (b"", b"\n", b"\n\n\n"):
        recs, b = split_whole(data)
        assert recs == [] and b.stop == binding.FASTQ_END and b.consumed == len(data)
    buf = np.frombuffer(b"@a\nACGT\n+\nII", dtype=np.uint8)
    off, ln, b, _ = binding.fastq_split(buf, at_eof=False)
    assert off.shape[0] == 0 and b.stop == binding.FASTQ_MORE and b.consumed == 0
    # trimming that empties a read ends the input there, as fetch_next_read_pair does (core.c:1168)
    recs, b = split_whole(b"@a\nACGTACGT\n+\nIIIIIIII\n@b\nACG\n+\nIII\n@c\nACGTACGT\n+\nIIIIIIII\n", 2, 1)
    assert [r[1] for r in recs] == [b"GTACG"] and b.stop == binding.FASTQ_END


def test_indexed_loader_equals_the_python_loader():
    """load_fastq_indexed (split + pack) builds the chunk load_fastq builds, on the golden read files (SE, PE, subjunc)."""
    for r1, r2 in (("se_R1.fastq.gz", None), ("pe_R1.fastq.gz", "pe_R2.fastq.gz"), ("sj_pe_R1.fastq.gz", "sj_pe_R2.fastq.gz")):
        p1, p2 = os.path.join(GOLDEN, r1), os.path.join(GOLDEN, r2) if r2 else None
        a, b = load_fastq(p1, p2), load_fastq_indexed(p1, p2)
        assert a.n_frags == b.n_frags and a.is_paired == b.is_paired and a.names == b.names
        assert np.array_equal(a.text, b.text) and np.array_equal(a.off, b.off) and np.array_equal(a.len, b.len)


def test_zero_copy_offsets_point_into_the_file_bytes():
    """Single-end: with base_offset the (off, len) arrays address the raw FASTQ bytes themselves (non-decreasing, as
    srcuda_vote_chunk wants them), so the chunk buffer can be the file buffer."""
    with gzip.open(os.path.join(GOLDEN, "se_R1.fastq.gz"), "rb") as fh:
        data = fh.read()
    buf = np.frombuffer(data, dtype=np.uint8)
    off, ln, b, _ = binding.fastq_split(buf, base_offset=4096)
    ref_chunk = load_fastq(os.path.join(GOLDEN, "se_R1.fastq.gz"))
    assert off.shape[0] == ref_chunk.n_frags and (np.diff(off.astype(np.int64)) > 0).all()
    for i in (0, 1, ref_chunk.n_frags - 1):
        s = int(off[i]) - 4096
        assert data[s:s + int(ln[i])] == ref_chunk.text[int(ref_chunk.off[i]):int(ref_chunk.off[i + 1])].tobytes()


def big_files():
    """Inputs large enough for srcuda_fastq_split_mt to cut (>= 64 KB per piece), tidy and untidy."""
    rnd = random.Random(5)
    dna = lambda n: bytes(rnd.choice(b"ACGT") for _ in range(n))  # noqa: E731
    recs = []
    for i in range(6000):
        n = 30 + (i * 7) % 120
        q = bytes(rnd.choice(b"@+IH#5") for _ in range(n))          # quality lines that look like name and '+' lines
        recs.append(b"@q%d/1 x\n%s\n+\n%s\n" % (i, dna(n), q))
    tidy = b"".join(recs)
    yield "tidy", tidy, None
    yield "no_final_newline", tidy[:-1], None
    yield "blank_lines_everywhere", b"\n".join(recs) + b"\n\n", None
    yield "crlf", tidy.replace(b"\n", b"\r\n"), None
    yield "two_line_gap_inside_records", tidy.replace(b"\n+\n", b"\n\n+\n"), None
    broken = recs[:4000] + [b"ACGT\n"] + recs[4000:]
    yield "broken_in_the_middle", b"".join(broken), 4000
    yield "short_read_ends_input_when_trimmed", b"".join(recs[:3000] + [b"@tiny\nACG\n+\nIII\n"] + recs[3000:]), None
    # every quality line looks like a name line and every sequence line like a '+' line: a guessed cut lands on a quality
    # line, and a piece that ends right behind a '+' line must not pass as complete (fixed records of 100 bases so that
    # any cut guess that is off is off consistently)
    yield "qualities_like_names_sequences_like_plus", b"".join(b"@n%05d\n+%s\n+\n@%s\n" % (i, dna(100), b"I" * 100) for i in range(20000)), None


@pytest.mark.parametrize("threads", [2, 3, 8])
def test_threads_never_change_the_outcome(threads):
    for name, data, bad_at in big_files():
        buf = np.frombuffer(data, dtype=np.uint8)
        for t5, t3, cap in ((0, 0, None), (2, 3, None), (0, 0, 2500), (0, 0, 6000), (4, 0, 4001)):
            def run(T):
                try:
                    off, ln, b, ex = binding.fastq_split(buf, True, t5, t3, max_reads=cap, want_names=True, threads=T)
                except binding.SubreadCudaError as err:
                    return str(err)
                return (off.tobytes(), ln.tobytes(), b.n_reads, b.consumed, b.stop, b.max_len) + tuple(v.tobytes() for v in ex.values())
            one, many = run(1), run(threads)
            assert one == many, f"{name}: trim {t5}/{t3}, max_reads {cap}"
            if bad_at is not None and (cap is None or cap > bad_at):
                assert isinstance(one, str) and f"{bad_at} records before it" in one
        # streaming: a window that is not at the end of the input
        off1, ln1, b1, _ = binding.fastq_split(buf[:len(data) // 2], False, threads=1)
        offT, lnT, bT, _ = binding.fastq_split(buf[:len(data) // 2], False, threads=threads)
        assert np.array_equal(off1, offT) and np.array_equal(ln1, lnT) and (b1.consumed, b1.stop) == (bT.consumed, bT.stop), name


def test_threaded_pack_equals_the_plain_one():
    rnd = np.random.default_rng(3)
    n = 50_000
    ln1 = rnd.integers(0, 300, n).astype(np.uint16)
    ln2 = rnd.integers(0, 300, n).astype(np.uint16)
    buf1 = rnd.integers(65, 91, int(ln1.sum()) + 1000, dtype=np.uint8)
    buf2 = rnd.integers(65, 91, int(ln2.sum()) + 1000, dtype=np.uint8)
    off1 = np.concatenate(([0], np.cumsum(ln1[:-1], dtype=np.uint64))).astype(np.uint64) + np.uint64(7)
    off2 = np.concatenate(([0], np.cumsum(ln2[:-1], dtype=np.uint64))).astype(np.uint64)
    for pe in (False, True):
        second = (buf2, off2, ln2) if pe else ()
        text, off, ln = binding.reads_pack(buf1, off1, ln1, *second)
        for T in (2, 5, 16):
            again = binding.reads_pack(buf1, off1, ln1, *second, threads=T)
            assert all(np.array_equal(x, y) for x, y in zip((text, off, ln), again)), f"paired={pe}, {T} threads"
        mates = 2 if pe else 1
        for f in (0, 123, n - 1):
            r = f * mates
            assert text[int(off[r]):int(off[r]) + int(ln[r])].tobytes() == buf1[int(off1[f]):int(off1[f]) + int(ln1[f])].tobytes()
            if pe:
                assert text[int(off[r + 1]):int(off[r + 1]) + int(ln[r + 1])].tobytes() == buf2[int(off2[f]):int(off2[f]) + int(ln2[f])].tobytes()


def test_sanitizer_fuzz(tmp_path):
    """tests/native/fastq_fuzz.cpp: 3000 damaged inputs through the splitter compiled with -fsanitize=address,undefined;
    one thread and several must agree and nothing may be read or reported outside the buffer."""
    import subprocess
    from conftest import ROOT
    exe = str(tmp_path / "fastq_fuzz")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", exe,
           os.path.join(ROOT, "tests", "native", "fastq_fuzz.cpp"), os.path.join(ROOT, "subread_b200", "csrc", "sr_fastq.cpp"), "-pthread"]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0:
        pytest.skip("no sanitizer runtime for g++ here: " + built.stderr[-200:])
    run = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"), timeout=300)
    assert run.returncode == 0 and run.stdout.startswith("ok:"), run.stdout[-400:] + run.stderr[-2000:]


def test_random_damage_against_the_reference_reader(ref, tmp_path):
    """Generated FASTQ with random damage: the splitter hands out exactly the records geinput_next_read_trim hands out
    before it stops, and reports a format issue exactly when the reference stops early on one."""
    rnd = random.Random(2026)
    alphabet = b"ACGTN@+\n\n \tIH#r1/"
    stopped_early = 0
    for case in range(250):
        recs = []
        for i in range(rnd.randrange(1, 25)):
            n = rnd.randrange(1, 70)
            recs.append(b"@r%d%s\n%s\n+\n%s\n" % (i, b" x" if rnd.random() < 0.3 else b"", bytes(rnd.choice(b"ACGTN") for _ in range(n)),
                                                   bytes(rnd.choice(b"IH#@+5") for _ in range(n))))
        data = bytearray(b"".join(recs))
        for _ in range(rnd.randrange(0, 3)):
            p = rnd.randrange(len(data))
            kind = rnd.randrange(3)
            if kind == 0:
                data[p] = rnd.choice(alphabet)
            elif kind == 1:
                del data[p:p + rnd.randrange(1, 5)]
            else:
                data.insert(p, rnd.choice(alphabet))
        data = bytes(data)
        if not data.strip(b"\n").startswith(b"@"):
            continue                                      # geinput_open would not call it FASTQ
        path = str(tmp_path / "damaged.fq")
        with open(path, "wb") as fh:
            fh.write(data)
        t5, t3 = (rnd.randrange(0, 4), rnd.randrange(0, 4)) if case % 3 == 0 else (0, 0)
        exp = reference_records(ref, path, t5, t3)
        buf = np.frombuffer(data, dtype=np.uint8)
        off, ln, b, ex, rc = binding.fastq_split(buf, True, t5, t3, want_names=True, partial=True)
        got = [(data[int(a):int(a) + int(n)], data[int(o):int(o) + int(m)]) for a, n, o, m in zip(ex["name_off"], ex["name_len"], off, ln)]
        assert got == [(e[0], e[1]) for e in exp], f"case {case}: {data!r}"
        stopped_early += rc != 0
    assert stopped_early > 20      # the damage did hit record structure often enough
