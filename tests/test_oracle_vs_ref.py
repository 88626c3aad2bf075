"""Pins the oracle against the UNMODIFIED reference built from /root/reference (oracle/_ref):
function level through libsubread_ref.so, and voting-step level through ref_dump, which is the
reference's own subread-align with its real do_voting (oracle/ref_hooks.c only reads the bigtable).
Skipped where oracle/_ref has not been built.  CPU only, about 40 s."""
import ctypes
import os
import random
import subprocess
import sys

import numpy as np
import pytest

from conftest import REF_DIR, ROOT, describe_mismatch, have_ref
from subread_b200 import VoteParams, load_fastq, used_slots_equal

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref (reference binaries) not built")


@pytest.fixture(scope="module")
def ref():
    L = ctypes.CDLL(os.path.join(REF_DIR, "libsubread_ref.so"))
    L.genekey2int.restype = ctypes.c_int
    L.match_chro.restype = ctypes.c_int
    return L


def test_reverse_read_table_every_byte(ref, oracle_mod):
    L = oracle_mod.lib()
    for c in range(1, 128):
        for n in (1, 2, 5):
            a = ctypes.create_string_buffer(bytes([c]) * n + b"\0")
            b = ctypes.create_string_buffer(bytes([c]) * n + b"\0")
            ref.reverse_read(a, n, 1)
            L.or_reverse_read(b, n)
            assert a.raw == b.raw, (c, n)


def test_genekey2int_random_ascii(ref, oracle_mod):
    L = oracle_mod.lib()
    rnd = random.Random(3)
    for _ in range(5000):
        s = bytes(rnd.randrange(1, 128) for _ in range(16))
        assert (ref.genekey2int(s, 1) & 0xFFFFFFFF) == L.or_genekey2int(s)


def test_subread_schedule_matches_do_voting_arithmetic(oracle_mod):
    """Closed forms of SURVEY.md section 8: S = 10 for 45..160 bp, step floor below, 6-bp grid above."""
    L = oracle_mod.lib()
    applied = ctypes.c_int()
    offs = (ctypes.c_int * 256)()
    for rlen, gap, total, want in ((100, 3, 10, 10), (150, 3, 10, 10), (250, 3, 10, 39), (100, 1, 10, 10), (30, 3, 10, 5), (16, 3, 10, 1),
                                   (1210, 3, 10, 63), (100, 3, 14, 14)):
        n = L.or_subread_schedule(rlen, gap, total, ctypes.byref(applied), offs, 256)
        assert applied.value == want, (rlen, gap, total, applied.value)
        assert n == want * gap
        assert all(0 <= offs[i] <= max(0, rlen - 16) + gap for i in range(n))


def test_match_chro_on_reference_index(ref, oracle_mod, golden_index_dir):
    """or_match_chro == match_chro for forward and reverse-complement windows on the tiny genome."""
    class GV(ctypes.Structure):  # gene_value_index_t, subread.h:313-322
        _fields_ = [("memory_block_size", ctypes.c_uint), ("start_base_offset", ctypes.c_uint), ("start_point", ctypes.c_uint),
                    ("length", ctypes.c_uint), ("values", ctypes.c_void_p), ("values_bytes", ctypes.c_uint),
                    ("appendix1", ctypes.c_void_p), ("appendix2", ctypes.c_void_p)]
    gv = GV()
    assert ref.gvindex_load(ctypes.byref(gv), os.path.join(golden_index_dir, "tiny.00.b.array").encode()) == 0
    idx = oracle_mod.OracleIndex(os.path.join(golden_index_dir, "tiny"))
    L = oracle_mod.lib()
    rnd = random.Random(9)
    for _ in range(3000):
        n = rnd.randrange(1, 120)
        s = bytes(rnd.choice(b"ACGTNacgtU.-B") for _ in range(n))
        pos = rnd.randrange(1000, 60000)
        for neg in (0, 1):
            assert ref.match_chro(s, ctypes.byref(gv), ctypes.c_uint(pos), n, neg, 1) == L.or_match_chro(s, ctypes.byref(idx.c), pos, n, neg)
    for pos in (0, 1, 1209, gv.length - 50, gv.length - 1, gv.length, 0xFFFF0001, 0xFFFFFFF0):
        assert ref.match_chro(b"ACGT" * 10, ctypes.byref(gv), ctypes.c_uint(pos), 40, 0, 1) == \
            L.or_match_chro(b"ACGT" * 10, ctypes.byref(idx.c), pos, 40, 0)


def test_voting_step_against_real_do_voting_multiblock(oracle_mod, tmp_path):
    """A 10-block index (-M 1 on a 1.2 Mbp repeat-rich genome): exercises the read-modify-write merge
    across index blocks exactly as read_chunk_circles (core.c:3567-3613) drives it, SE and PE."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_data
    fa = str(tmp_path / "g.fa")
    gen_data.write_genome(fa, 3, 400000, seed=5, repeats=150)
    prefix = str(tmp_path / "idx")
    subprocess.check_call([os.path.join(REF_DIR, "bin", "subread-buildindex"), "-o", prefix, "-M", "1", fa],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    n_blocks = len([f for f in os.listdir(tmp_path) if f.endswith(".b.tab")])
    assert n_blocks > 1
    gen_data.write_reads(fa, str(tmp_path / "se"), 4000, 100, 1, False, 0.01, 0.2, True, 0.002)
    gen_data.write_reads(fa, str(tmp_path / "pe"), 3000, 120, 2, True, 0.01, 0.2, True, 0.002)
    for r1, r2 in ((str(tmp_path / "se_R1.fastq"), None), (str(tmp_path / "pe_R1.fastq"), str(tmp_path / "pe_R2.fastq"))):
        bt = str(tmp_path / "bt.bin")
        if os.path.exists(bt):
            os.remove(bt)
        cmd = [os.path.join(REF_DIR, "ref_dump"), "-t", "1", "-T", "2", "-i", prefix, "-r", r1, "--SAMoutput", "-o", str(tmp_path / "o.sam")]
        if r2:
            cmd += ["-R", r2]
        subprocess.check_call(cmd, env=dict(os.environ, SRDUMP_BIGTABLE=bt), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ref_align = oracle_mod.read_bigtable_dump(bt)[0]["align"]
        chunk = load_fastq(r1, r2)
        params = VoteParams.defaults(r2 is not None)
        got = None
        for b in range(n_blocks):
            idx = oracle_mod.OracleIndex(prefix, b)
            got, _, _ = oracle_mod.vote_chunk(idx, params, chunk, out_align=got)
            idx.close()
        ok, bad = used_slots_equal(got, ref_align, ignore_gapped_flag=True)
        assert ok, describe_mismatch(got, ref_align, bad)


@pytest.mark.parametrize("blocks", ["one_block", "split_index"])
def test_subjunc_voting_step_against_real_subjunc(blocks, oracle_mod, tmp_path):
    """subjunc (do_breakpoint_detection) beyond the golden fixture: a 1.2 Mbp genome with planted GT..AG / CT..AC
    introns, spliced SE reads of mixed lengths (some > 160 bp) and spliced pairs, through the reference's own subjunc
    (ref_dump_subjunc: real do_voting + copy_vote_to_alignment_res' junction half, core-junction.c:1073-1334) — alignment
    table, subjunc_result_t table and big-margin records against the oracle; also with an index split into blocks
    (-M 1), where every table is merged read-modify-write across blocks (core.c:3567-3613)."""
    from conftest import junc_tables_equal
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_data
    fa, faS = str(tmp_path / "g.fa"), str(tmp_path / "gS.fa")
    gen_data.write_genome(fa, 3, 400000, seed=11, repeats=40 if blocks == "split_index" else 0)
    gen_data.write_spliced(fa, faS, str(tmp_path / "se"), 3000, 180, 21, False, mixed_len=True)   # same seed: same introns
    gen_data.write_spliced(fa, faS, str(tmp_path / "pe"), 2000, 100, 21, True)
    prefix = str(tmp_path / "idx")
    cmd = [os.path.join(REF_DIR, "bin", "subread-buildindex"), "-o", prefix]
    cmd += ["-M", "1"] if blocks == "split_index" else ["-B"]
    subprocess.check_call(cmd + [faS], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    n_blocks = len([f for f in os.listdir(tmp_path) if f.endswith(".b.tab")])
    assert (n_blocks > 1) == (blocks == "split_index")
    for r1, r2 in ((str(tmp_path / "se_R1.fastq"), None), (str(tmp_path / "pe_R1.fastq"), str(tmp_path / "pe_R2.fastq"))):
        bt = str(tmp_path / "bt.bin")
        if os.path.exists(bt):
            os.remove(bt)
        cmd = [os.path.join(REF_DIR, "ref_dump_subjunc"), "-T", "1", "-i", prefix, "-r", r1, "--SAMoutput", "-o", str(tmp_path / "o.sam")]
        if r2:
            cmd += ["-R", r2]
        subprocess.check_call(cmd, env=dict(os.environ, SRDUMP_BIGTABLE=bt), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dump = oracle_mod.read_bigtable_dump(bt)
        assert len(dump) == 1
        exp, exp_junc, exp_bm = dump[0]["align"], dump[0]["junc"], dump[0]["big_margin"]
        assert (exp_junc["minor_votes"] > 0).sum() > 100, "the input must exercise the junction search"
        chunk = load_fastq(r1, r2)
        params = VoteParams.defaults(r2 is not None, True)
        got, js = None, {}
        for b in range(n_blocks):
            idx = oracle_mod.OracleIndex(prefix, b)
            got, _, _ = oracle_mod.vote_chunk(idx, params, chunk, out_align=got, junc_state=js)
            idx.close()
        ok, bad = used_slots_equal(got, exp, ignore_gapped_flag=True)
        assert ok, describe_mismatch(got, exp, bad)
        assert junc_tables_equal(got, js["junc"], js["big_margin"], exp_junc, np.ascontiguousarray(exp_bm[:, :, :9]).reshape(-1, 9))
