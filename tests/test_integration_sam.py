"""SAM-level parity through the reference's own host code: `subread-align-cuda` is the reference's CLI,
chunk loop, event tables, realignment and SAM writer with STEP_VOTING bound to libsubread_cuda.so
(integration/do_voting_cuda.c).  Its SAM output must be byte-identical to the stock subread-align's,
apart from the @PG header line (north_star: "identical SAM records apart from the @PG header")."""
import os
import subprocess
import sys

import pytest

from conftest import REF_DIR, ROOT

pytestmark = pytest.mark.gpu

STOCK = os.path.join(REF_DIR, "bin", "subread-align")
CUDA = os.path.join(REF_DIR, "subread-align-cuda")
BUILDINDEX = os.path.join(REF_DIR, "bin", "subread-buildindex")


def sam_body(path):
    with open(path, "rb") as fh:
        return b"".join(l for l in fh if not l.startswith(b"@PG"))


@pytest.mark.skipif(not (os.path.exists(STOCK) and os.path.exists(CUDA)), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("kind,paired,threads", [("uniform", False, 1), ("uniform", True, 4), ("repeats", False, 4), ("repeats", True, 2)])
def test_sam_identical_to_stock_subread_align(kind, paired, threads, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_data
    fa = str(tmp_path / "g.fa")
    gen_data.write_genome(fa, 4, 400000, seed=31, repeats=500 if kind == "repeats" else 0)
    prefix = str(tmp_path / "idx")
    # the index both binaries read is the reference builder's own (a defect of this repository's builder must not cancel out)
    subprocess.run([BUILDINDEX, "-B", "-o", prefix, fa], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    gen_data.write_reads(fa, str(tmp_path / "r"), 20000, 120, 9, paired, 0.01, 0.15, True, 0.002)
    reads = ["-r", str(tmp_path / "r_R1.fastq")] + (["-R", str(tmp_path / "r_R2.fastq")] if paired else [])
    outs = []
    for exe, tag in ((STOCK, "stock"), (CUDA, "cuda")):
        out = str(tmp_path / f"{tag}.sam")
        subprocess.run([exe, "-t", "1", "-T", str(threads), "-i", prefix] + reads + ["--SAMoutput", "-o", out], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        outs.append(sam_body(out))
    assert len(outs[0]) > 1000
    assert outs[0] == outs[1]


SUBJUNC = os.path.join(REF_DIR, "bin", "subjunc")
SUBJUNC_CUDA = os.path.join(REF_DIR, "subjunc-cuda")
FIX = os.path.join(REF_DIR, "fixtures")


@pytest.mark.skipif(not (os.path.exists(SUBJUNC) and os.path.exists(SUBJUNC_CUDA) and os.path.exists(os.path.join(FIX, "chr901.fa"))),
                    reason="oracle/_ref binaries or fixtures not built")
@pytest.mark.parametrize("devices", ["0", "0,0"], ids=["one_handle", "two_handles"])
def test_subjunc_sam_identical_to_stock_subjunc(devices, tmp_path):
    """subjunc through the binding (do_breakpoint_detection: junction tables and big-margin records from the GPU, find_new_junctions
    and everything after it the reference's own host code) on the reference's junction fixture, the command of
    test/subjunc/subjunc-test.sh: SAM and the junction table (.junction.bed) byte-identical to the stock subjunc's.  With
    SRCUDA_DEVICES=0,0 the chunk is split over two handles (srcuda_vote_chunk_multi behind the reference's CLI)."""
    prefix = str(tmp_path / "small1")
    subprocess.run([BUILDINDEX, "-o", prefix, os.path.join(FIX, "chr901.fa")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    outs = []
    for exe, tag in ((SUBJUNC, "stock"), (SUBJUNC_CUDA, "cuda")):
        out = str(tmp_path / f"{tag}.sam")
        subprocess.run([exe, "-T", "4", "-i", prefix, "-r", os.path.join(FIX, "junction-reads-A.fq"), "-R", os.path.join(FIX, "junction-reads-B.fq"),
                        "--SAMoutput", "-o", out], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=dict(os.environ, SRCUDA_DEVICES=devices))
        bed = out + ".junction.bed"
        outs.append((sam_body(out), open(bed, "rb").read() if os.path.exists(bed) else None))
    assert len(outs[0][0]) > 100000
    assert outs[0][0] == outs[1][0], "SAM records differ"
    assert outs[0][1] is not None and outs[0][1] == outs[1][1], "junction tables differ"
    assert outs[0][1].count(b"\n") > 10


REF_DUMP = os.path.join(REF_DIR, "ref_dump")


@pytest.mark.skipif(not (os.path.exists(REF_DUMP) and os.path.exists(CUDA)), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("paired,gz,reader", [(False, False, "1"), (True, False, "1"), (True, True, "1"), (True, False, "0")],
                         ids=["se_plain_library_reader", "pe_plain_library_reader", "pe_gzip_reference_reader", "pe_plain_reference_reader"])
def test_sam_identical_over_several_chunks(paired, gz, reader, tmp_path):
    """The chunk loop (read_chunk_circles, core.c:3539): 20 000 fragments in chunks of 6 000 (a test knob both binaries carry, since
    the real chunk is 7 M reads), so that the bound binary goes through chunk start / chunk end / next chunk several times.  With
    plain FASTQ input it reads every chunk through the library's own reader (srcuda_reader_*) and moves the reference's files on
    to where that reader stopped; with gzip input (or SRCUDA_BINDING_READER=0) through fetch_next_read_pair.  SAM identical to the
    unmodified reference's either way."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_data
    fa = str(tmp_path / "g.fa")
    gen_data.write_genome(fa, 4, 400000, seed=77, repeats=100)
    prefix = str(tmp_path / "idx")
    subprocess.run([BUILDINDEX, "-B", "-o", prefix, fa], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    gen_data.write_reads(fa, str(tmp_path / "r"), 20000, 110, 19, paired, 0.01, 0.15, True, 0.002, gz=gz)
    ext = ".fastq.gz" if gz else ".fastq"
    reads = ["-r", str(tmp_path / ("r_R1" + ext))] + (["-R", str(tmp_path / ("r_R2" + ext))] if paired else [])
    env = dict(os.environ, SUBREAD_TEST_READS_PER_CHUNK="6000", SRCUDA_BINDING_READER=reader)
    outs = []
    for exe, tag in ((REF_DUMP, "stock"), (CUDA, "cuda")):
        out = str(tmp_path / f"{tag}.sam")
        subprocess.run([exe, "-t", "1", "-T", "3", "-i", prefix] + reads + ["--SAMoutput", "-o", out], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
        outs.append(sam_body(out))
    assert outs[0].count(b"\n") >= 20000 * (2 if paired else 1)
    assert outs[0] == outs[1]


@pytest.mark.skipif(not (os.path.exists(STOCK) and os.path.exists(CUDA) and os.path.exists(os.path.join(FIX, "chr901.fa"))),
                    reason="oracle/_ref binaries or fixtures not built")
@pytest.mark.parametrize("out_name,extra", [("o.sam", ["--SAMoutput"]), ("o.bam", ["--keepReadOrder"])])
def test_gpu_voting_with_the_librarys_writer(out_name, extra, tmp_path):
    """Both halves of the library behind the reference's CLI: STEP_VOTING on the GPU (integration/do_voting_cuda.c) and the SAM /
    BAM file written by srcuda_out_* (integration/output_native.c, SRCUDA_NATIVE_OUTPUT=1), on the reference's paired-end fixture
    (test-err-mut).  Alignment records identical to the stock subread-align's: the SAM text minus @PG, the BAM records one by one."""
    from test_output_program import bam_records, no_pg
    prefix = str(tmp_path / "chr901")
    subprocess.run([BUILDINDEX, "-o", prefix, os.path.join(FIX, "chr901.fa")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    reads = ["-r", os.path.join(FIX, "test-err-mut-r1.fq.gz"), "-R", os.path.join(FIX, "test-err-mut-r2.fq.gz")]
    outs = []
    for exe, tag, native in ((STOCK, "stock", "0"), (CUDA, "cuda", "1")):
        out = str(tmp_path / f"{tag}_{out_name}")
        p = subprocess.run([exe, "-t", "1", "-T", "4", "-i", prefix] + reads + extra + ["-o", out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           env=dict(os.environ, SRCUDA_NATIVE_OUTPUT=native))
        assert p.returncode == 0, p.stdout.decode()[-2000:]
        assert (b"written by libsubread_cuda" in p.stdout) == (native == "1")
        outs.append(open(out, "rb").read())
    if out_name.endswith(".sam"):
        assert no_pg(outs[0]) == no_pg(outs[1]) and len(outs[0]) > 1000000
    else:
        (ta, ra), (tb, rb) = bam_records(outs[0]), bam_records(outs[1])
        assert no_pg(ta) == no_pg(tb)
        assert ra == rb and len(ra) > 40000


@pytest.mark.skipif(not (os.path.exists(REF_DUMP) and os.path.exists(CUDA)), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("paired", [False, True], ids=["se", "pe"])
def test_sam_identical_with_later_passes_read_from_memory(paired, tmp_path):
    """The GPU-bound binary with integration/read_cache.c switched on (SRCUDA_READ_CACHE=1): the voting pass reads each chunk
    through the library's reader, the realignment pass copies its reads out of the chunk's bytes in memory instead of parsing the
    file a character at a time under input_lock.  Several chunks; SAM identical to the unmodified reference's."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_data
    fa = str(tmp_path / "g.fa")
    gen_data.write_genome(fa, 4, 400000, seed=78, repeats=100)
    prefix = str(tmp_path / "idx")
    subprocess.run([BUILDINDEX, "-B", "-o", prefix, fa], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    gen_data.write_reads(fa, str(tmp_path / "r"), 20000, 110, 21, paired, 0.01, 0.15, True, 0.002)
    reads = ["-r", str(tmp_path / "r_R1.fastq")] + (["-R", str(tmp_path / "r_R2.fastq")] if paired else [])
    outs = []
    for exe, tag, cache in ((REF_DUMP, "stock", "0"), (CUDA, "cuda", "1")):
        out = str(tmp_path / f"{tag}.sam")
        p = subprocess.run([exe, "-t", "1", "-T", "3", "-i", prefix] + reads + ["--SAMoutput", "-o", out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           env=dict(os.environ, SUBREAD_TEST_READS_PER_CHUNK="6000", SRCUDA_READ_CACHE=cache))
        assert p.returncode == 0, p.stdout.decode()[-2000:]
        assert (b"read their chunks from memory" in p.stdout) == (cache == "1")
        outs.append(sam_body(out))
    assert outs[0].count(b"\n") >= 20000 * (2 if paired else 1)
    assert outs[0] == outs[1]
