"""Row f2 inside the real program: the reference's own CLI, voting and realignment with the SAM / BAM file written by the
library (integration/output_native.c -> srcuda_out_*, include/subread_output.h).

`subread-align-out` / `subjunc-out` are the reference's mains linked with the binding; SRCUDA_NATIVE_OUTPUT=1 takes the
library's writer, =0 forwards every call to the reference's own code.  Run twice with the same arguments (same argv[0], same
relative output name in two directories) the two files must be equal byte for byte, @PG line and compressed BAM blocks
included, wherever the reference's output is deterministic: SAM, BAM at -T 1, BAM with --keepReadOrder.  Plain multi-thread BAM
(per-thread order in the reference) is compared as a set of records.  CPU only: voting stays the reference's here."""
import os
import subprocess
import zlib

import pytest

from conftest import REF_DIR

FIX = os.path.join(REF_DIR, "fixtures")
ALIGN_OUT = os.path.join(REF_DIR, "subread-align-out")
SUBJUNC_OUT = os.path.join(REF_DIR, "subjunc-out")
STOCK = os.path.join(REF_DIR, "bin", "subread-align")
STOCK_SUBJUNC = os.path.join(REF_DIR, "bin", "subjunc")
BUILDINDEX = os.path.join(REF_DIR, "bin", "subread-buildindex")

pytestmark = pytest.mark.skipif(not (os.path.exists(ALIGN_OUT) and os.path.exists(SUBJUNC_OUT) and os.path.exists(os.path.join(FIX, "chr901.fa"))),
                                reason="oracle/_ref binaries or fixtures not built")


@pytest.fixture(scope="module")
def index(tmp_path_factory):
    d = tmp_path_factory.mktemp("idx")
    prefix = str(d / "chr901")
    # -M 3: still one block for this genome (393 216 items per block), but a table of 262 147 buckets instead of 11 million --
    # the reference spends two seconds of every run loading the empty buckets of the default size
    subprocess.run([BUILDINDEX, "-M", "3", "-o", prefix, os.path.join(FIX, "chr901.fa")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert not os.path.exists(prefix + ".01.b.tab")
    return prefix


def run(exe, args, cwd, native, out_name):
    os.makedirs(cwd, exist_ok=True)
    env = dict(os.environ, SRCUDA_NATIVE_OUTPUT="1" if native else "0", SRCUDA_OUT_THREADS="3")
    p = subprocess.run([exe] + args + ["-o", out_name], cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert p.returncode == 0, p.stdout.decode()[-2000:]
    assert (b"written by libsubread_cuda" in p.stdout) == bool(native), "which writer ran?"
    return open(os.path.join(cwd, out_name), "rb").read()


def bam_payload(raw):
    out, at = [], 0
    while at < len(raw):
        bsize = int.from_bytes(raw[at + 16:at + 18], "little") + 1
        out.append(zlib.decompress(raw[at + 18:at + bsize - 8], -15))
        at += bsize
    return b"".join(out)


def bam_records(raw):
    """(header text, sorted list of alignment records) of a BAM file."""
    p = bam_payload(raw)
    assert p[:4] == b"BAM\x01"
    l_text = int.from_bytes(p[4:8], "little")
    text = p[8:8 + l_text]
    at = 8 + l_text
    n_ref = int.from_bytes(p[at:at + 4], "little")
    at += 4
    for _ in range(n_ref):
        l_name = int.from_bytes(p[at:at + 4], "little")
        at += 4 + l_name + 4
    recs = []
    while at < len(p):
        n = int.from_bytes(p[at:at + 4], "little")
        recs.append(p[at:at + 4 + n])
        at += 4 + n
    return text, recs


def no_pg(text):
    return b"".join(l for l in text.splitlines(True) if not l.startswith(b"@PG"))


PE = ["-r", os.path.join(FIX, "test-err-mut-r1.fq.gz"), "-R", os.path.join(FIX, "test-err-mut-r2.fq.gz")]


@pytest.fixture(scope="module")
def pe_cut(tmp_path_factory):
    """The first 6 000 pairs of the fixture as plain FASTQ (keeps the CPU suite short; the whole fixture runs in the first case)."""
    import gzip
    d = tmp_path_factory.mktemp("pe_cut")
    args = []
    for flag, mate in (("-r", "r1"), ("-R", "r2")):
        lines = gzip.open(os.path.join(FIX, f"test-err-mut-{mate}.fq.gz"), "rb").read().split(b"\n")[:24000]
        path = str(d / f"{mate}.fq")
        open(path, "wb").write(b"\n".join(lines) + b"\n")
        args += [flag, path]
    return args


@pytest.mark.parametrize("tag,args,out", [
    ("sam_T4", ["-t", "1", "-T", "4", "--SAMoutput"], "o.sam"),
    ("bam_T1", ["-t", "1", "-T", "1"], "o.bam"),
    ("bam_T4_ordered", ["-t", "1", "-T", "4", "--keepReadOrder"], "o.bam"),
    ("sam_T3_multimapping_rg", ["-t", "1", "-T", "3", "--SAMoutput", "--multiMapping", "-B", "3", "--rg-id", "grp1", "--rg", "SM:sample", "-d", "100", "-D", "300"], "o.sam"),
    ("bam_T2_ordered_ff", ["-t", "1", "-T", "2", "--keepReadOrder", "-S", "ff", "--multiMapping", "-B", "2"], "o.bam"),
])
def test_paired_end_fixture_files_identical(index, pe_cut, tmp_path, tag, args, out):
    """test/subread-align/data/test-err-mut-r{1,2}.fq.gz (20 000 pairs with errors and indels) through the library's writer."""
    reads = PE if tag == "sam_T4" else pe_cut
    a = run(ALIGN_OUT, args + ["-i", index] + reads, str(tmp_path / "ref"), 0, out)
    b = run(ALIGN_OUT, args + ["-i", index] + reads, str(tmp_path / "native"), 1, out)
    assert len(a) > 100000
    if a != b and out.endswith(".bam"):
        ta, ra = bam_records(a)
        tb, rb = bam_records(b)
        assert ta == tb, "header text differs"
        assert ra == rb, "records differ"
    assert a == b
    if out.endswith(".sam"):   # ... and the stock binary (another argv[0], so another @PG) agrees on everything else
        c = run(STOCK, args + ["-i", index] + reads, str(tmp_path / "stock"), 0, out)
        assert no_pg(c) == no_pg(b)


def test_unordered_bam_holds_the_same_records(index, pe_cut, tmp_path):
    """-T 4 BAM without --keepReadOrder: the reference's order is per thread and changes between runs; same records, same header."""
    args = ["-t", "1", "-T", "4", "-i", index] + pe_cut
    ta, ra = bam_records(run(ALIGN_OUT, args, str(tmp_path / "ref"), 0, "o.bam"))
    tb, rb = bam_records(run(ALIGN_OUT, args, str(tmp_path / "native"), 1, "o.bam"))
    assert ta == tb and len(ra) == len(rb) == 12000
    assert sorted(ra) == sorted(rb)
    # mates stay next to each other in the library's stream (one block never splits a pair)
    names = [r[36:36 + r[12]] for r in rb]
    assert all(names[i] == names[i + 1] for i in range(0, len(names), 2))


@pytest.mark.parametrize("out,extra", [("o.sam", ["--SAMoutput"]), ("o.bam", ["--keepReadOrder"])])
def test_subjunc_fixture_files_identical(index, tmp_path, out, extra):
    """test/subjunc/data/junction-reads-{A,B}.fq through subjunc: spliced CIGARs (N sections), XS:A columns, TLEN across introns."""
    args = ["-T", "4", "-i", index, "-r", os.path.join(FIX, "junction-reads-A.fq"), "-R", os.path.join(FIX, "junction-reads-B.fq")] + extra
    a = run(SUBJUNC_OUT, args, str(tmp_path / "ref"), 0, out)
    b = run(SUBJUNC_OUT, args, str(tmp_path / "native"), 1, out)
    assert len(a) > 100000 and a == b
    if out.endswith(".sam"):
        assert b"N" in b.split(b"\n")[-2000].split(b"\t")[5] or any(b"N" in l.split(b"\t")[5] for l in b.split(b"\n")[-3000:-1])
        c = run(STOCK_SUBJUNC, args, str(tmp_path / "stock"), 0, out)
        assert no_pg(c) == no_pg(b)


def test_single_end_reads_with_and_without_qualities(index, tmp_path):
    """Single-end input: 8 000 genRandomReads reads as FASTQ, and the same reads as FASTA (no qualities: the writer fills
    'I', core.c:2011-2018), SAM and ordered BAM."""
    gen = os.path.join(REF_DIR, "bin", "genRandomReads")
    expr = str(tmp_path / "expr.txt")
    open(expr, "w").write("GeneID\tTPM\nchr901\t1000000\n")
    p = subprocess.run([gen, "--transcriptFasta", os.path.join(FIX, "chr901.fa"), "--expressionLevels", expr, "--outputPrefix", str(tmp_path / "sim"),
                        "--totalReads", "8000", "--readLen", "100", "--randSeed", "7"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    fq = str(tmp_path / "sim_R1.fastq.gz")
    if p.returncode != 0 or not os.path.exists(fq):
        pytest.skip("genRandomReads did not produce reads here")
    import gzip
    lines = gzip.open(fq, "rb").read().split(b"\n")
    fa = str(tmp_path / "sim.fa")
    with open(fa, "wb") as fh:
        for i in range(0, len(lines) - 3, 4):
            fh.write(b">" + lines[i][1:] + b"\n" + lines[i + 1] + b"\n")
    for reads, tag, out, extra in ((fq, "fastq", "o.bam", ["--keepReadOrder"]), (fa, "fasta", "o.sam", ["--SAMoutput"])):
        args = ["-t", "1", "-T", "4", "-i", index, "-r", reads] + extra
        a = run(ALIGN_OUT, args, str(tmp_path / f"ref_{tag}"), 0, out)
        b = run(ALIGN_OUT, args, str(tmp_path / f"native_{tag}"), 1, out)
        assert len(a) > 100000 and a == b, (tag, out)
        if tag == "fasta":
            assert b.split(b"\n")[-2].split(b"\t")[10] == b"I" * 100


def run_cached(exe, args, cwd, cache, out_name, extra_env=None):
    """One run with the later passes read from memory (integration/read_cache.c, SRCUDA_READ_CACHE=1) or through the original."""
    os.makedirs(cwd, exist_ok=True)
    env = dict(os.environ, SRCUDA_NATIVE_OUTPUT="0", SRCUDA_READ_CACHE="1" if cache else "0", **(extra_env or {}))
    p = subprocess.run([exe] + args + ["-o", out_name], cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert p.returncode == 0, p.stdout.decode()[-2000:]
    assert (b"read their chunks from memory" in p.stdout) == bool(cache), "which reader ran?"
    return open(os.path.join(cwd, out_name), "rb").read()


@pytest.mark.parametrize("tag,args,env", [
    ("pe", ["-t", "1", "-T", "4", "--SAMoutput"], {}),
    ("pe_three_chunks", ["-t", "1", "-T", "3", "--SAMoutput"], {"SUBREAD_TEST_READS_PER_CHUNK": "2500"}),
    ("pe_trimmed_ff", ["-t", "1", "-T", "4", "--SAMoutput", "--trim5", "3", "--trim3", "2", "-S", "ff"], {}),
    ("pe_bam_ordered", ["-t", "1", "-T", "4", "--keepReadOrder"], {"SUBREAD_TEST_READS_PER_CHUNK": "4000"}),
])
def test_later_passes_read_from_memory(index, pe_cut, tmp_path, tag, args, env):
    """fetch_next_read_pair served from the chunk's bytes in memory (row f3 for the passes after voting): the files are the same,
    @PG line included, as with the reference's read-at-a-time reader -- over several chunks (a test knob sizes them), with trimming
    and with both mates on the forward strand."""
    out = "o.bam" if "bam" in tag else "o.sam"
    a = run_cached(ALIGN_OUT, args + ["-i", index] + pe_cut, str(tmp_path / "original"), 0, out, env)
    b = run_cached(ALIGN_OUT, args + ["-i", index] + pe_cut, str(tmp_path / "cached"), 1, out, env)
    assert len(a) > 100000 and a == b


def test_later_passes_read_from_memory_subjunc_and_single_end(index, tmp_path):
    """The same for subjunc on the reference's junction fixture (plain FASTQ, paired) and for a single-end run of its first file."""
    reads = ["-r", os.path.join(FIX, "junction-reads-A.fq"), "-R", os.path.join(FIX, "junction-reads-B.fq")]
    a = run_cached(SUBJUNC_OUT, ["-T", "4", "-i", index, "--SAMoutput"] + reads, str(tmp_path / "original"), 0, "o.sam")
    b = run_cached(SUBJUNC_OUT, ["-T", "4", "-i", index, "--SAMoutput"] + reads, str(tmp_path / "cached"), 1, "o.sam")
    assert len(a) > 100000 and a == b
    a = run_cached(ALIGN_OUT, ["-t", "1", "-T", "4", "-i", index, "--SAMoutput"] + reads[:2], str(tmp_path / "original_se"), 0, "o.sam")
    b = run_cached(ALIGN_OUT, ["-t", "1", "-T", "4", "-i", index, "--SAMoutput"] + reads[:2], str(tmp_path / "cached_se"), 1, "o.sam")
    assert len(a) > 100000 and a == b


@pytest.mark.parametrize("tag,args,env", [
    ("gz_pe", ["-t", "1", "-T", "4", "--SAMoutput"], {}),
    ("gz_pe_chunks_multimapping", ["-t", "1", "-T", "3", "--SAMoutput", "--multiMapping", "-B", "2", "-S", "ff"], {"SUBREAD_TEST_READS_PER_CHUNK": "7000"}),
])
def test_gzip_input_is_kept_from_the_first_pass(index, tmp_path, tag, args, env):
    """gzip FASTQ cannot be indexed from a byte range: with SRCUDA_READ_CACHE=1 the fragments are kept as the chunk's first voting
    pass fetches them and the later passes are served from that record (the reference's test-err-mut fixture, 20 000 pairs, also in
    three chunks).  Same file as with the original reader."""
    a = run_cached(ALIGN_OUT, args + ["-i", index] + PE, str(tmp_path / "original"), 0, "o.sam", env)
    b = run_cached(ALIGN_OUT, args + ["-i", index] + PE, str(tmp_path / "cached"), 1, "o.sam", env)
    assert len(a) > 1000000 and a == b


def test_fasta_input_is_kept_from_the_first_pass(index, tmp_path):
    """FASTA input (no qualities) through the same record."""
    import gzip
    lines = gzip.open(os.path.join(FIX, "test-err-mut-r1.fq.gz"), "rb").read().split(b"\n")[:16000]
    fa = str(tmp_path / "r.fa")
    with open(fa, "wb") as fh:
        for i in range(0, len(lines) - 3, 4):
            fh.write(b">" + lines[i][1:] + b"\n" + lines[i + 1] + b"\n")
    args = ["-t", "1", "-T", "4", "-i", index, "-r", fa, "--SAMoutput"]
    a = run_cached(ALIGN_OUT, args, str(tmp_path / "original"), 0, "o.sam")
    b = run_cached(ALIGN_OUT, args, str(tmp_path / "cached"), 1, "o.sam")
    assert len(a) > 100000 and a == b
