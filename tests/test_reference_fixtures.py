"""The reference's OWN fixtures through the voting step (SURVEY section 8c/8d cfg 1, BASELINE configs[0]):

  * test/chr901.fa indexed by the unmodified `subread-buildindex` -- default (gapped) and `-B -F -M100`, the form
    test/subread-align/subread-align-test.sh uses -- with 100 000 x 100 bp single-end reads from the reference's own
    `genRandomReads --randSeed 42 --truthInReadNames` (the survey's probe run: first read R000000000:chr901:600508);
  * test/subread-align/data/test-err-mut-r{1,2}.fq.gz (paired-end, sequencing errors and mutations, gzip input);
  * test/subjunc/data/junction-reads-{A,B}.fq through `subjunc` (junction half of the voting step).

The expected bigtable is dumped by the real do_voting (oracle/_ref/ref_dump, ref_dump_subjunc: the reference's CLI with
oracle/ref_hooks.c reading the table right after STEP_VOTING).  CPU leg: the oracle restatement must equal it.  GPU leg
(-m gpu): libsubread_cuda.so through the C ABI must equal it byte for byte on every used slot -- with the index read from the
files `subread-buildindex` wrote, nothing of this repository's own index builder involved.

The data files are the reference's; oracle/Makefile copies them next to the reference binaries (oracle/_ref/fixtures,
git-ignored, shipped to the GPU box with oracle/_ref).  Nothing here reads /root/reference at run time."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REF_DIR, describe_mismatch, have_ref, junc_tables_equal
from subread_b200 import VoteParams, load_fastq, used_slots_equal

FIX = os.path.join(REF_DIR, "fixtures")
BIN = os.path.join(REF_DIR, "bin")
pytestmark = pytest.mark.skipif(not (have_ref() and os.path.exists(os.path.join(FIX, "chr901.fa"))),
                                reason="oracle/_ref (reference binaries + fixtures) not built")

# index flavour -> subread-buildindex arguments (subread-align-test.sh:7, subjunc-test.sh:3)
INDEXES = {"default": [], "full_one_block": ["-B", "-F", "-M100"]}
# read set -> (tool, mates, subjunc)
READSETS = {"chr901_100k_se": ("ref_dump", 1, False), "err_mut_pe": ("ref_dump", 2, False), "junction_pe": ("ref_dump_subjunc", 2, True)}
CASES = [("default", "chr901_100k_se"), ("full_one_block", "chr901_100k_se"), ("full_one_block", "err_mut_pe"), ("default", "err_mut_pe"),
         ("default", "junction_pe")]


def _run(cmd, **kw):
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, **kw)


@pytest.fixture(scope="module")
def work(tmp_path_factory):
    """Indexes, read files and reference bigtables, made once per test module (about 40 s of CPU)."""
    d = tmp_path_factory.mktemp("ref_fixtures")
    out = {"dir": str(d), "index": {}, "reads": {}, "expect": {}}
    for name, args in INDEXES.items():
        prefix = str(d / f"chr901_{name}")
        _run([os.path.join(BIN, "subread-buildindex"), "-o", prefix] + args + [os.path.join(FIX, "chr901.fa")])
        out["index"][name] = (prefix, len([f for f in os.listdir(d) if f.startswith(f"chr901_{name}.") and f.endswith(".b.tab")]))
    with open(d / "expr.txt", "w") as fh:
        fh.write("TranscriptID\tTPM\nchr901\t1000000\n")
    _run([os.path.join(BIN, "genRandomReads"), "--transcriptFasta", os.path.join(FIX, "chr901.fa"), "--expressionLevels", str(d / "expr.txt"),
          "--totalReads", "100000", "--readLen", "100", "--randSeed", "42", "--truthInReadNames", "--outputPrefix", str(d / "r100k")])
    out["reads"]["chr901_100k_se"] = (str(d / "r100k_R1.fastq.gz"), None)
    out["reads"]["err_mut_pe"] = (os.path.join(FIX, "test-err-mut-r1.fq.gz"), os.path.join(FIX, "test-err-mut-r2.fq.gz"))
    out["reads"]["junction_pe"] = (os.path.join(FIX, "junction-reads-A.fq"), os.path.join(FIX, "junction-reads-B.fq"))
    return out


def expected(work, oracle_mod, index, reads):
    """The reference's bigtable for (index, read set); cached."""
    key = (index, reads)
    if key not in work["expect"]:
        tool, _, _ = READSETS[reads]
        prefix, _ = work["index"][index]
        r1, r2 = work["reads"][reads]
        bt = os.path.join(work["dir"], f"bt_{index}_{reads}.bin")
        cmd = [os.path.join(REF_DIR, tool), "-T", "4", "-i", prefix, "-r", r1, "--SAMoutput", "-o", os.path.join(work["dir"], "o.sam")]
        if tool == "ref_dump":
            cmd += ["-t", "1"]
        if r2:
            cmd += ["-R", r2]
        _run(cmd, env=dict(os.environ, SRDUMP_BIGTABLE=bt))
        dump = oracle_mod.read_bigtable_dump(bt)
        assert len(dump) == 1                    # one chunk
        work["expect"][key] = dump[0]
    return work["expect"][key]


def check(got, js, exp, subjunc, what):
    ok, bad = used_slots_equal(got, exp["align"], ignore_gapped_flag=True)
    assert ok, what + ": " + describe_mismatch(got, exp["align"], bad)
    if subjunc:
        assert (exp["junc"]["minor_votes"] > 0).sum() > 100, "the fixture must exercise the junction search"
        assert junc_tables_equal(got, js["junc"], js["big_margin"], exp["junc"], np.ascontiguousarray(exp["big_margin"][:, :, :9]).reshape(-1, 9)), what


@pytest.mark.parametrize("index,reads", CASES)
def test_oracle_equals_the_reference_on_its_own_fixtures(index, reads, work, oracle_mod):
    _, mates, subjunc = READSETS[reads]
    exp = expected(work, oracle_mod, index, reads)
    prefix, n_blocks = work["index"][index]
    chunk = load_fastq(*work["reads"][reads])
    assert chunk.n_frags == exp["align"].shape[0] and chunk.mates == mates
    params = VoteParams.defaults(mates == 2, subjunc)
    got, js = None, {}
    for b in range(n_blocks):
        idx = oracle_mod.OracleIndex(prefix, b)
        got, _, _ = oracle_mod.vote_chunk(idx, params, chunk, out_align=got, junc_state=js if subjunc else None, threads=8)
        idx.close()
    mapped = float((got[:, :, 0]["selected_votes"] > 0).mean())
    assert mapped > 0.5, f"{index}/{reads}: only {mapped:.2f} of the reads have a candidate"
    check(got, js, exp, subjunc, f"oracle, {index}/{reads}")


@pytest.mark.gpu
@pytest.mark.parametrize("index,reads", CASES)
def test_cuda_equals_the_reference_on_its_own_fixtures(index, reads, work, oracle_mod):
    """The product path: index files written by subread-buildindex -> srcuda_index_load, reads -> srcuda_vote_chunk, every
    block of the index merged read-modify-write; the default tiers and the warp kernel alone must both give the
    reference's table."""
    from subread_b200.binding import Mapper
    _, mates, subjunc = READSETS[reads]
    exp = expected(work, oracle_mod, index, reads)
    prefix, n_blocks = work["index"][index]
    chunk = load_fastq(*work["reads"][reads])
    params = VoteParams.defaults(mates == 2, subjunc)
    for env in ({}, {"SRCUDA_SIMPLE_TIER": "0"}, {"SRCUDA_TIER_INDEL": "1"}):      # default tiers; warp kernel alone; the tier's one-indel form forced on
        os.environ.update(env)
        try:
            m = Mapper(0)
        finally:
            for k in env:
                del os.environ[k]
        got, js = None, {}
        punted = 0
        for b in range(n_blocks):
            m.index_load(prefix, b)
            got, _ = m.vote_chunk(chunk, params, out_align=None if got is None else got.reshape(-1), junc_state=js if subjunc else None)
            punted += m.timing().n_simple_punted
        m.close()
        check(got, js, exp, subjunc, f"cuda {env or 'default tiers'}, {index}/{reads}")
        if not env and reads == "chr901_100k_se":
            # error-free reads mostly stay in tier 0; chr901 is a real chromosome, and its tandem repeats and low-complexity
            # stretches make a quarter of the passes non-separable (24 % queued for the warp kernel, measured)
            assert punted < chunk.n_frags // 2, f"tier 0 queued {punted} of {chunk.n_frags}"
