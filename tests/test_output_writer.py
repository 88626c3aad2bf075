"""Row f2 (SAM / BAM writer, include/subread_output.h) against the UNMODIFIED reference's writer.

Record level: the same records go through the reference's SamBam_writer_* (libsubread_ref.so, single stream = thread_no -1,
sambam-file.c:1097-1797) and through srcuda_out_*; the two BAM files must be equal byte for byte, compressed blocks included,
whatever the number of deflate threads.  CPU only; skipped where oracle/_ref has not been built."""
import ctypes
import gzip
import os
import random
import zlib

import pytest

from conftest import REF_DIR, have_ref
from subread_b200 import binding
from subread_b200.binding import OUT_BAM, OUT_SAM, OutputWriter, SamRecord

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref (reference binaries) not built")


@pytest.fixture(scope="module")
def ref():
    L = ctypes.CDLL(os.path.join(REF_DIR, "libsubread_ref.so"))
    cp, vp, ci, cu = ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_uint
    L.SamBam_writer_create.argtypes = [vp, cp, ci, ci, ci, cp]
    L.SamBam_writer_add_header.argtypes = [vp, cp, ci]
    L.SamBam_writer_add_chromosome.argtypes = [vp, cp, cu, ci]
    L.SamBam_writer_finish_header.argtypes = [vp]
    L.SamBam_writer_add_read.argtypes = [vp, ci, cp, cu, cp, cu, ci, cp, cp, cu, ci, ci, cp, cp, cp, ci]
    L.SamBam_writer_close.argtypes = [vp]
    return L


def bgzf_payload(path):
    """Concatenated uncompressed payload of a BGZF file and the list of its block sizes."""
    raw = open(path, "rb").read()
    out, sizes, at = [], [], 0
    while at < len(raw):
        assert raw[at:at + 4] == b"\x1f\x8b\x08\x04"
        bsize = int.from_bytes(raw[at + 16:at + 18], "little") + 1
        data = zlib.decompress(raw[at + 18:at + bsize - 8], -15)
        assert zlib.crc32(data) == int.from_bytes(raw[at + bsize - 8:at + bsize - 4], "little")
        out.append(data)
        sizes.append(len(data))
        at += bsize
    return b"".join(out), sizes


def random_cigar(rng, read_len):
    if rng.random() < 0.5:
        return f"{read_len}M"
    left = rng.randrange(1, read_len - 1)
    kind = rng.choice("IDNS")
    if kind == "S":
        return f"{left}S{read_len - left}M"
    if kind == "I":
        ins = rng.randrange(1, min(6, read_len - left))
        return f"{left}M{ins}I{read_len - left - ins}M"
    return f"{left}M{rng.randrange(1, 5000)}{kind}{read_len - left}M"


def random_columns(rng):
    cols = [f"HI:i:{rng.randrange(1, 4)}", f"NH:i:{rng.randrange(1, 4)}"]
    if rng.random() < 0.3:
        cols.append("RG:Z:" + "".join(rng.choice("abcXYZ019._-") for _ in range(rng.randrange(1, 12))))
    if rng.random() < 0.3:
        cols.append("XS:A:" + rng.choice("+-"))
    if rng.random() < 0.2:
        cols.append(f"XF:f:{rng.random() * 100:.3f}")
    if rng.random() < 0.2:
        cols.append("XB:B:i," + ",".join(str(rng.randrange(-500, 500)) for _ in range(rng.randrange(1, 6))))
    if rng.random() < 0.1:
        cols.append("XG:B:f," + ",".join(f"{rng.random():.2f}" for _ in range(rng.randrange(1, 4))))
    if rng.random() < 0.05:
        cols.append("XL:Z:" + "Q" * rng.randrange(700, 900))          # runs into the reference's 780 / 750 byte marks
    cols.append(f"NM:i:{rng.randrange(0, 9)}")
    text = "\t".join(cols)
    if rng.random() < 0.1:
        text = "\t" + text                                              # a leading tab, as write_single_fragment can produce
    return text


def make_records(seed, n, chromosomes, paired):
    rng = random.Random(seed)
    recs = []
    for i in range(n):
        L = rng.choice([36, 75, 100, 101, 150, 151, 250])
        name = f"read{seed}_{i}" + ("x" * rng.randrange(0, 30))
        seq = "".join(rng.choice("ACGTACGTACGTN" + ("RYKMacgtn" if rng.random() < 0.05 else "")) for _ in range(L))
        qual = "".join(chr(33 + rng.randrange(0, 42)) for _ in range(L))
        mapped = rng.random() < 0.9
        if mapped:
            chro, clen = rng.choice(chromosomes)
            pos = rng.randrange(1, clen)
            cigar = random_cigar(rng, L)
            cols = random_columns(rng)
        else:
            chro, pos, cigar, cols = "*", 0, "*", ("" if rng.random() < 0.5 else "RG:Z:grp")
        flags = rng.randrange(0, 1 << 11)
        if paired:
            nxt = rng.choice(["=", "*", rng.choice(chromosomes)[0]])
            npos = rng.randrange(0, 1 << 28)
            tlen = rng.randrange(-100000, 100000)
        else:
            nxt, npos, tlen = "*", 0, 0
        if rng.random() < 0.02:
            chro = "not_in_the_table"                                    # refID -1 in both writers
        recs.append(dict(name=name, flags=flags, chro=chro, pos=pos, mapq=rng.randrange(0, 60), cigar=cigar, next_chro=nxt, next_pos=npos,
                         tlen=tlen, read_len=L, seq=seq, qual=qual, additional=cols))
    return recs


def write_with_reference(ref, path, header_lines, chromosomes, recs, paired):
    w = ctypes.create_string_buffer(1 << 20)                            # SamBam_Writer (sambam-file.h:110-149), a few KB
    assert ref.SamBam_writer_create(w, os.fsencode(path), 1, 0, 0, b"") == 0
    for line in header_lines:
        ref.SamBam_writer_add_header(w, ctypes.create_string_buffer(line.encode()), 0)   # the reference tokenises the line in place
    for name, length in chromosomes:
        ref.SamBam_writer_add_chromosome(w, name.encode(), length, 1)
    ref.SamBam_writer_finish_header(w)
    for i, r in enumerate(recs):
        committable = (not paired) or (i & 1)
        ref.SamBam_writer_add_read(w, -1, r["name"].encode(), r["flags"], r["chro"].encode(), r["pos"], r["mapq"], r["cigar"].encode(),
                                   r["next_chro"].encode(), r["next_pos"], r["tlen"], r["read_len"], r["seq"].encode(), r["qual"].encode(),
                                   r["additional"].encode(), int(committable))
    ref.SamBam_writer_close(w)


def to_struct(r):
    return SamRecord(r["name"].encode(), r["flags"], r["chro"].encode(), r["pos"], r["mapq"], r["cigar"].encode(), r["next_chro"].encode(),
                     r["next_pos"], r["tlen"], r["read_len"], r["seq"].encode(), r["qual"].encode(), r["additional"].encode())


def write_native(path, fmt, header_lines, chromosomes, recs, paired, threads, batch=False):
    w = OutputWriter(path, fmt, threads)
    for line in header_lines:
        w.add_header(line)
    for name, length in chromosomes:
        w.add_chromosome(name, length, True)
    w.finish_header()
    if batch:
        w.add_reads([to_struct(r) for r in recs], 2 if paired else 1)
    else:
        for i, r in enumerate(recs):
            w.add_read(to_struct(r), (not paired) or bool(i & 1))
    return w.close()


CHROMS = [("chr1", 248956422), ("chr2", 242193529), ("chrX", 156040895), ("scaffold_with_a_long_name.1", 5000), ("chrM", 16569)]


@pytest.mark.parametrize("paired", [False, True])
@pytest.mark.parametrize("threads", [1, 4])
def test_bam_bytes_equal_reference_writer(ref, tmp_path, paired, threads):
    recs = make_records(11 + paired, 6000, CHROMS, paired)
    header = ["@HD\tVN:1.0\tSO:unsorted"]
    tail = ["@PG\tID:subread\tPN:subread\tVN:2.0.6\tCL:\"subread-align\" \"-i\" \"idx\""]
    a, b = str(tmp_path / "ref.bam"), str(tmp_path / "native.bam")
    # the reference writes @HD, the @SQ lines, then @PG (core.c:3850-3870)
    w = ctypes.create_string_buffer(1 << 20)
    assert ref.SamBam_writer_create(w, os.fsencode(a), 1, 0, 0, b"") == 0
    ref.SamBam_writer_add_header(w, ctypes.create_string_buffer(header[0].encode()), 0)
    for name, length in CHROMS:
        ref.SamBam_writer_add_chromosome(w, name.encode(), length, 1)
    ref.SamBam_writer_add_header(w, ctypes.create_string_buffer(tail[0].encode()), 0)
    ref.SamBam_writer_finish_header(w)
    for i, r in enumerate(recs):
        ref.SamBam_writer_add_read(w, -1, r["name"].encode(), r["flags"], r["chro"].encode(), r["pos"], r["mapq"], r["cigar"].encode(),
                                   r["next_chro"].encode(), r["next_pos"], r["tlen"], r["read_len"], r["seq"].encode(), r["qual"].encode(),
                                   r["additional"].encode(), int((not paired) or (i & 1)))
    ref.SamBam_writer_close(w)

    nw = OutputWriter(b, OUT_BAM, threads)
    nw.add_header(header[0])
    for name, length in CHROMS:
        nw.add_chromosome(name, length, True)
    nw.add_header(tail[0])
    nw.finish_header()
    half = len(recs) // 2
    for i, r in enumerate(recs[:half]):
        nw.add_read(to_struct(r), (not paired) or bool(i & 1))
    nw.add_reads([to_struct(r) for r in recs[half:]], 2 if paired else 1)
    size = nw.close()
    assert size == os.path.getsize(b)
    ra, rb = open(a, "rb").read(), open(b, "rb").read()
    if ra != rb:
        pa, sa = bgzf_payload(a)
        pb, sb = bgzf_payload(b)
        assert sa == sb, "block cuts differ"
        assert pa == pb, "record bytes differ"
    assert ra == rb
    payload, sizes = bgzf_payload(b)
    assert payload[:4] == b"BAM\x01" and sizes[-1] == 0 and len(sizes) > 10


def test_long_header_is_cut_like_the_reference(ref, tmp_path):
    """6 200 contigs (BASELINE configs[2]): header text and reference list both span several blocks."""
    chroms = [(f"contig_{i:05d}_of_the_synthetic_genome", 500000 + i) for i in range(6200)]
    recs = make_records(5, 300, chroms, False)
    a, b = str(tmp_path / "ref.bam"), str(tmp_path / "native.bam")
    write_with_reference(ref, a, ["@HD\tVN:1.0\tSO:unsorted"], chroms, recs, False)
    write_native(b, OUT_BAM, ["@HD\tVN:1.0\tSO:unsorted"], chroms, recs, False, 3, batch=True)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert len(bgzf_payload(b)[1]) > 8


def test_header_lines_with_sq_fill_the_chromosome_table(ref, tmp_path):
    """add_header(.., add_chro=1) takes SN / LN from an @SQ line (sambam-file.c:1375-1393)."""
    a, b = str(tmp_path / "ref.bam"), str(tmp_path / "native.bam")
    lines = ["@HD\tVN:1.0", "@SQ\tSN:alpha\tLN:1000", "@SQ\tLN:2000\tSN:beta\tM5:abc", "@CO\tcomment"]
    recs = make_records(9, 50, [("alpha", 1000), ("beta", 2000)], True)
    w = ctypes.create_string_buffer(1 << 20)
    assert ref.SamBam_writer_create(w, os.fsencode(a), 1, 0, 0, b"") == 0
    for line in lines:
        ref.SamBam_writer_add_header(w, ctypes.create_string_buffer(line.encode()), 1)
    ref.SamBam_writer_finish_header(w)
    for i, r in enumerate(recs):
        ref.SamBam_writer_add_read(w, -1, r["name"].encode(), r["flags"], r["chro"].encode(), r["pos"], r["mapq"], r["cigar"].encode(),
                                   r["next_chro"].encode(), r["next_pos"], r["tlen"], r["read_len"], r["seq"].encode(), r["qual"].encode(),
                                   r["additional"].encode(), i & 1)
    ref.SamBam_writer_close(w)
    nw = OutputWriter(b, OUT_BAM, 2)
    for line in lines:
        nw.add_header(line, add_chro=True)
    nw.add_reads([to_struct(r) for r in recs], 2)
    nw.close()
    assert open(a, "rb").read() == open(b, "rb").read()


def test_sam_lines_to_bam_and_back(ref, tmp_path):
    """SAM text written by the native writer, fed back through add_sam_lines, gives the BAM of the records themselves
    (SamBam_writer_add_read_line, sambam-file.c:1668) -- and the text is the reference's record line (core.c:2156)."""
    recs = [r for r in make_records(21, 3000, CHROMS, True) if not r["additional"].startswith("\t")]
    recs = recs[: len(recs) & ~1]
    sam, bam1, bam2 = str(tmp_path / "o.sam"), str(tmp_path / "direct.bam"), str(tmp_path / "via_text.bam")
    write_native(sam, OUT_SAM, ["@HD\tVN:1.0\tSO:unsorted"], CHROMS, recs, True, 1)
    text = open(sam, "rb").read()
    lines = text.split(b"\n")
    n_header = 1 + len(CHROMS)
    assert lines[0] == b"@HD\tVN:1.0\tSO:unsorted" and lines[1] == b"@SQ\tSN:chr1\tLN:248956422"
    r = recs[0]
    expect = "%s\t%d\t%s\t%u\t%d\t%s\t%s\t%u\t%d\t%s\t%s%s%s" % (r["name"], r["flags"], r["chro"], r["pos"], r["mapq"], r["cigar"], r["next_chro"],
                                                            r["next_pos"], r["tlen"], r["seq"], r["qual"], "\t" if r["additional"] else "",
                                                            r["additional"])
    assert lines[n_header].decode() == expect
    write_native(bam1, OUT_BAM, ["@HD\tVN:1.0\tSO:unsorted"], CHROMS, recs, True, 2)
    w = OutputWriter(bam2, OUT_BAM, 2)
    w.add_header("@HD\tVN:1.0\tSO:unsorted")
    for name, length in CHROMS:
        w.add_chromosome(name, length, True)
    body = b"\n".join(lines[n_header:])
    w.add_sam_lines(body, 2)
    w.close()
    assert open(bam1, "rb").read() == open(bam2, "rb").read()
    a = str(tmp_path / "ref.bam")
    write_with_reference(ref, a, ["@HD\tVN:1.0\tSO:unsorted"], CHROMS, recs, True)
    assert open(a, "rb").read() == open(bam1, "rb").read()


def test_no_compression_blocks(tmp_path):
    """fastest_compression of the reference's temporary files: stored blocks (sambam-file.h:42), still valid BGZF."""
    recs = make_records(3, 500, CHROMS, False)
    a, b = str(tmp_path / "stored.bam"), str(tmp_path / "deflated.bam")
    w = OutputWriter(a, OUT_BAM, 2, no_compression=True)
    w.add_header("@HD\tVN:1.0")
    for name, length in CHROMS:
        w.add_chromosome(name, length)
    w.add_reads([to_struct(r) for r in recs], 1)
    w.close()
    write_native(b, OUT_BAM, ["@HD\tVN:1.0"], CHROMS, recs, False, 1)
    assert bgzf_payload(a)[0] == bgzf_payload(b)[0]
    assert os.path.getsize(a) > os.path.getsize(b)
    assert gzip.open(a, "rb").read() == bgzf_payload(b)[0]           # any gzip reader takes the file


def test_error_paths(tmp_path):
    L = binding.load_library()
    h = ctypes.c_void_p()
    assert L.srcuda_out_open(os.fsencode(str(tmp_path / "no_such_dir" / "x.bam")), OUT_BAM, 1, 0, ctypes.byref(h)) == 3   # SRCUDA_ERR_IO
    assert b"cannot create" in L.srcuda_out_error(None)
    assert L.srcuda_out_open(None, 7, 1, 0, ctypes.byref(h)) == 5
    assert L.srcuda_out_add_read(None, None, 1) == 5
    assert L.srcuda_out_close(None, None) == 5
    w = OutputWriter(str(tmp_path / "e.bam"), OUT_BAM, 1)
    w.add_chromosome("a", 10)
    with pytest.raises(binding.SubreadCudaError, match="twice"):
        w.add_chromosome("a", 10)
    w = OutputWriter(str(tmp_path / "f.bam"), OUT_BAM, 1)
    with pytest.raises(binding.SubreadCudaError, match="bad read format"):
        w.add_sam_lines(b"name\t0\tchr1\t1\n", 1)


def _mate(ok, chro_id, offset, strand, cigar, name, text, qual, additional=b"", mm=0, mapq=40):
    bufs = [ctypes.create_string_buffer(x, len(x) + 8) for x in (name, text, qual)]
    m = binding.SamMate(1, int(ok), chro_id, offset, strand, mapq, mm, len(text), cigar, additional,
                        ctypes.cast(bufs[0], ctypes.c_void_p), ctypes.cast(bufs[1], ctypes.c_void_p), ctypes.cast(bufs[2], ctypes.c_void_p))
    m._keep = bufs
    return m


def test_fragment_records_and_two_step_form(tmp_path):
    """srcuda_out_add_fragment against hand-checked records (flags, TLEN sign, mate columns, NM = mismatches + indel bases, reverse
    strand printed reverse-complemented), and srcuda_out_format_fragment + srcuda_out_append give the same file."""
    cfg = binding.SamConfig(1, 50, 600, 0, 1, 0, b"")
    seq1, seq2 = b"ACGTTGCANN" * 5, b"TTGACCAGTA" * 5
    q1, q2 = bytes(range(40, 90)), bytes(range(89, 39, -1))
    paths = [str(tmp_path / "one.sam"), str(tmp_path / "two.sam")]
    L = binding.load_library()
    for two_step, path in enumerate(paths):
        w = OutputWriter(path, OUT_SAM, 1)
        w.add_chromosome("chrA", 100000)
        w.add_chromosome("chrB", 50000)
        w.finish_header()
        cases = [
            # a proper pair: mate 1 forward at 1000, mate 2 (stored strand 0 = reverse of the fragment) at 1200 with a deletion
            (_mate(1, 0, 1000, 0, b"50M", b"frag1/1", seq1, q1, b"\tSB:i:7", 2), _mate(1, 0, 1200, 0, b"20M3D30M", b"frag1/2", seq2, q2, b"", 1)),
            # mate 2 unmapped
            (_mate(1, 1, 77, 1, b"10S40M", b"frag2", seq1, q1), _mate(0, -1, 0, 0, None, b"frag2", seq2, q2)),
            # different chromosomes
            (_mate(1, 0, 5, 0, b"50M", b"frag3", seq1, q1), _mate(1, 1, 9, 1, b"25M2I23M", b"frag3", seq2, q2, b"", 3)),
            # both unmapped
            (_mate(0, -1, 0, 0, None, b"frag4", seq1, q1), _mate(0, -1, 0, 0, None, b"frag4", seq2, q2)),
        ]
        counters = binding.SamCounters()
        for m1, m2 in cases:
            if two_step:
                buf = ctypes.create_string_buffer(8192)
                n = ctypes.c_size_t()
                assert L.srcuda_out_format_fragment(w.h, ctypes.byref(cfg), ctypes.byref(m1), ctypes.byref(m2), 1, 0, ctypes.byref(counters), buf, 8192,
                                                    ctypes.byref(n)) == 0
                assert L.srcuda_out_append(w.h, buf, n.value, 1) == 0
            else:
                w.add_fragment(cfg, m1, m2, 1, 0, counters)
        w.close()
        assert counters.all_mapped_reads == 3 and counters.all_correct_PE_reads == 1 and counters.not_properly_pairs_different_chro == 2   # frag2 (a record without a position counts, as in the reference) and frag3
    a, b = open(paths[0], "rb").read(), open(paths[1], "rb").read()
    assert a == b
    rows = [l.split("\t") for l in a.decode().splitlines() if not l.startswith("@")]
    assert len(rows) == 8
    r1, r2 = rows[0], rows[1]
    assert r1[:9] == ["frag1", str(0x1 | 0x2 | 0x20 | 0x40), "chrA", "1000", "40", "50M", "=", "1200", "250"]   # calc_tlen walks the left mate only: 1200 - 1050 + 50 + 50
    assert r2[:9] == ["frag1", str(0x1 | 0x2 | 0x10 | 0x80), "chrA", "1200", "40", "20M3D30M", "=", "1000", "-250"]
    assert r1[9] == seq1.decode() and r1[11:] == ["HI:i:1", "NH:i:1", "SB:i:7", "NM:i:2"]
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    # a second mate on stored strand 0 is printed as it came in (the input step has turned it already), flag 0x10 notwithstanding
    assert r2[9] == seq2.decode() and r2[10] == q2.decode() and r2[11:] == ["HI:i:1", "NH:i:1", "NM:i:4"]
    assert rows[2][9] == seq1.translate(comp)[::-1].decode() and rows[2][10] == q1[::-1].decode()      # strand 1: reverse-complemented
    assert rows[2][1] == str(0x1 | 0x8 | 0x10 | 0x40) and rows[2][6:9] == ["*", "0", "0"] and rows[2][2:4] == ["chrB", "77"]
    assert rows[3][1] == str(0x1 | 0x4 | 0x20 | 0x80) and rows[3][2:9] == ["*", "0", "0", "*", "chrB", "77", "0"] and rows[3][11:] == ["HI:i:1", "NH:i:1"]
    assert rows[4][6:9] == ["chrB", "9", "0"] and rows[5][6:9] == ["chrA", "5", "0"] and rows[5][11:] == ["HI:i:1", "NH:i:1", "NM:i:5"]
    assert len(rows[6]) == 11 and rows[6][1] == str(0x1 | 0x4 | 0x8 | 0x40) and len(rows[7]) == 11
    # calc_tlen on its own: equal starts -> the longer read; a gap between the mates -> both reads plus the gap
    assert L.srcuda_sam_calc_tlen(100, b"50M", 50, 100, b"75M", 75) == 75
    assert L.srcuda_sam_calc_tlen(100, b"50M", 50, 400, b"50M", 50) == 350
    assert L.srcuda_sam_calc_tlen(400, b"50M", 50, 100, b"20M200N30M", 50) == 150   # the fallback counts from the last matched section, intron left out (core.c:1783)


def test_optional_columns_follow_the_reference_converter(ref, tmp_path):
    """The text -> BAM rules of the optional columns against SamBam_compress_additional (sambam-file.c:1478-1582) itself, on
    2 000 random column sets: every type, empty values, long strings around the 750 / 780 byte marks, arrays with odd cells."""
    ref.SamBam_compress_additional.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    rng = random.Random(99)
    alphabet = "abcXYZ019.:_-+,"

    def column():
        tag = rng.choice("XYZN") + rng.choice("ABCMHIS019")
        kind = rng.choice("iiZZAfBBH")
        if kind == "i":
            return f"{tag}:i:{rng.choice(['', '-'])}{rng.randrange(0, 10 ** rng.randrange(1, 10))}"
        if kind == "f":
            return f"{tag}:f:{rng.random() * 1000:.4f}"
        if kind in "ZH":
            n = rng.choice([0, 1, 5, 40, 300, 760, 800])
            return f"{tag}:{kind}:" + "".join(rng.choice(alphabet.replace(",", "")) for _ in range(n))
        if kind == "A":
            return f"{tag}:A:{rng.choice('+-.xyz')}"
        cell = rng.choice("iifcsS")
        cells = [str(rng.randrange(-70000, 70000)) if cell != "f" else f"{rng.random() * 9:.3f}" for _ in range(rng.randrange(1, 60))]
        if rng.random() < 0.1:
            cells.append("9" * 35)                                           # a cell longer than the reference's 29-character buffer
        return f"{tag}:B:{cell}," + ",".join(cells)

    texts = []
    for _ in range(2000):
        t = "\t".join(column() for _ in range(rng.randrange(1, 7)))
        texts.append(("\t" if rng.random() < 0.15 else "") + t)
    path = str(tmp_path / "cols.bam")
    w = OutputWriter(path, OUT_BAM, 2)
    w.add_header("@HD\tVN:1.0")
    w.add_chromosome("c", 1000)
    for i, t in enumerate(texts):
        w.add_read(SamRecord(b"r", 0, b"c", 5, 1, b"4M", b"*", 0, 0, 4, b"ACGT", b"IIII", t.encode()), True)
    w.close()
    payload, _ = bgzf_payload(path)
    l_text = int.from_bytes(payload[4:8], "little")
    at = 8 + l_text + 4 + 4 + 2 + 4                                        # magic, text, n_ref, one reference entry "c\0"
    for i, t in enumerate(texts):
        n = int.from_bytes(payload[at:at + 4], "little")
        aux = payload[at + 4 + 32 + 2 + 4 + 2 + 4:at + 4 + n]                # fixed part, name "r\0", one CIGAR op, 2 + 4 bytes of read
        out = ctypes.create_string_buffer(4000)
        m = ref.SamBam_compress_additional(t.encode(), out)
        assert aux == out.raw[:m], (i, t[:80])
        at += 4 + n
    assert at == len(payload)


@pytest.mark.parametrize("sanitizer", ["address,undefined", "thread"])
def test_sanitizer_fuzz(sanitizer, tmp_path):
    """tests/native/output_fuzz.cpp: 40 000 damaged SAM lines through srcuda_out_add_sam_lines and fragments formatted on four
    threads at once, with the writer compiled under AddressSanitizer + UBSan and under ThreadSanitizer: damaged lines may be
    refused, nothing may crash, touch memory outside its buffers, or race."""
    import subprocess
    from conftest import ROOT
    exe = str(tmp_path / "output_fuzz")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=" + sanitizer, "-fno-sanitize-recover=all", "-fno-omit-frame-pointer", "-o", exe,
           os.path.join(ROOT, "tests", "native", "output_fuzz.cpp"), os.path.join(ROOT, "subread_b200", "csrc", "sr_output.cpp"), "-lz", "-pthread"]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0:
        pytest.skip("no sanitizer runtime for g++ here: " + built.stderr[-200:])
    run = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"), timeout=600)
    assert run.returncode == 0 and run.stdout.startswith("lines taken"), run.stdout[-400:] + run.stderr[-2000:]
