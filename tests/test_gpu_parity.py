"""Parity tests proper: the CUDA path (through the C ABI) against the oracle and the reference's
golden bigtables.  Bit-exact bar: every byte of every used slot (and, against the golden files,
every byte of the table)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, REF_DIR, ROOT, describe_mismatch, have_ref, junc_tables_equal, load_case, load_case_junc
from subread_b200 import MAPPING_RESULT, VoteParams, load_fastq, used_slots_equal

pytestmark = pytest.mark.gpu

CASE_NAMES = ["se_gapped", "pe_gapped", "se_full", "pe_full", "se_I16", "pe_n14_ff"]


@pytest.fixture(scope="module")
def mapper():
    from subread_b200.binding import Mapper
    m = Mapper(0)
    yield m
    m.close()


@pytest.mark.parametrize("slow", [False, True], ids=["warp", "scalar"])
@pytest.mark.parametrize("name", CASE_NAMES)
def test_golden_bigtables(name, slow, mapper, golden_cases, golden_index_dir, oracle_mod):
    case = next(c for c in golden_cases if c["name"] == name)
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    mapper.index_load(prefix, 0)
    mapper.set_force_slow_path(slow)
    try:
        got, ver = mapper.vote_chunk(chunk, params, want_verify=True)
    finally:
        mapper.set_force_slow_path(False)
    ok, bad = used_slots_equal(got, exp)
    assert ok, describe_mismatch(got, exp, bad)
    assert got.tobytes() == exp.tobytes()
    idx = oracle_mod.OracleIndex(prefix)
    _, over, _ = oracle_mod.vote_chunk(idx, params, chunk, want_verify=True)
    assert (ver == over).all(), f"verify differs at {np.flatnonzero(ver != over)[:5]}"
    t = mapper.timing()
    assert t.n_launches >= 1
    if slow:
        assert t.n_overflow == chunk.n_frags


def test_upload_from_arrays_equals_load_from_files(mapper, golden_cases, golden_index_dir, oracle_mod):
    case = next(c for c in golden_cases if c["name"] == "pe_gapped")
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    oidx = oracle_mod.OracleIndex(prefix)
    a = oidx.arrays()
    mapper.index_upload(a["index_gap"], a["padding"], a["bucket_start"], a["keys"], a["vals"], a["start_point"], a["length"],
                        a["values"], a["read_offsets"])
    d = mapper.describe()
    assert (d.index_gap, d.buckets_number, d.n_items) == (a["index_gap"], a["buckets_number"], a["n_items"])
    got, _ = mapper.vote_chunk(chunk, params)
    assert got.tobytes() == exp.tobytes()


def test_many_small_batches_and_ragged_input(golden_cases, golden_index_dir, oracle_mod):
    """Double-buffered batching must not change anything; also covers empty chunks and reads < 16 bp."""
    from subread_b200 import pack_reads
    from subread_b200.binding import Mapper
    os.environ["SRCUDA_BATCH_FRAGS"] = "97"
    try:
        m = Mapper(0)
    finally:
        del os.environ["SRCUDA_BATCH_FRAGS"]
    case = next(c for c in golden_cases if c["name"] == "pe_gapped")
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    m.index_load(prefix, 0)
    got, _ = m.vote_chunk(chunk, params)
    assert got.tobytes() == exp.tobytes()
    assert m.timing().n_batches == (chunk.n_frags + 96) // 97
    # empty chunk
    empty = pack_reads([], None)
    got, _ = m.vote_chunk(empty, VoteParams.defaults(False))
    assert got.shape[0] == 0
    # ragged: very short reads (skipped by the reference, core.c:3116), 16..20 bp, and a maximum-length read
    rng = np.random.default_rng(5)
    import gzip
    genome = b"".join(l.strip() for l in gzip.open(os.path.join(GOLDEN, "tiny.fa.gz")) if not l.startswith(b">"))
    seqs = [b"ACGT", b"A" * 15, genome[100:116], genome[200:219], genome[300:320], genome[1000:2210], genome[5000:5100], b"N" * 50]
    ragged = pack_reads(seqs, None)
    p = VoteParams.defaults(False)
    got, ver = m.vote_chunk(ragged, p, want_verify=True)
    idx = oracle_mod.OracleIndex(prefix)
    ora, over, _ = oracle_mod.vote_chunk(idx, p, ragged, want_verify=True)
    assert got.tobytes() == ora.tobytes()
    assert (ver == over).all()
    m.close()


@pytest.mark.parametrize("compact", ["1", "0"], ids=["compacted", "plain"])
def test_result_copy_forms_agree(compact, golden_cases, golden_index_dir):
    """The host-buffer entry point brings a zero table's results down either whole or compacted (k_compact_results: 24-byte records of
    plain slots, 72-byte records of the others -- the golden cases hold reads with indels -- scattered by host threads behind a
    follower thread).  Forced either way, over several batches, with few scatter threads: the golden bigtables, byte for byte;
    the compacted form moves fewer bytes."""
    from subread_b200.binding import Mapper
    os.environ.update(SRCUDA_BATCH_FRAGS="300", SRCUDA_COMPACT_D2H=compact, SRCUDA_EXPAND_THREADS="3")
    try:
        m = Mapper(0)
    finally:
        for k in ("SRCUDA_BATCH_FRAGS", "SRCUDA_COMPACT_D2H", "SRCUDA_EXPAND_THREADS"):
            del os.environ[k]
    for name in ("se_gapped", "pe_gapped"):
        case = next(c for c in golden_cases if c["name"] == name)
        prefix, chunk, params, exp = load_case(case, golden_index_dir)
        m.index_load(prefix, 0)
        got, _ = m.vote_chunk(chunk, params)
        assert got.tobytes() == exp.tobytes(), name
        t = m.timing()
        whole = chunk.n_frags * chunk.mates * params.multi_best_reads * 68
        assert t.n_batches == (chunk.n_frags + 299) // 300
        assert (0 < t.d2h_bytes < whole) if compact == "1" else t.d2h_bytes >= whole, (name, t.d2h_bytes, whole)
    m.close()


def test_read_modify_write_merges_like_the_oracle(mapper, golden_cases, golden_index_dir, oracle_mod):
    """Second pass over the same block with the first pass's table as input (multi-block merge path)."""
    for name in ("se_gapped", "pe_gapped"):
        case = next(c for c in golden_cases if c["name"] == name)
        prefix, chunk, params, exp = load_case(case, golden_index_dir)
        mapper.index_load(prefix, 0)
        idx = oracle_mod.OracleIndex(prefix)
        twice_ora, _, _ = oracle_mod.vote_chunk(idx, params, chunk, out_align=exp.copy())
        twice_gpu, _ = mapper.vote_chunk(chunk, params, out_align=exp.copy().reshape(-1))
        ok, bad = used_slots_equal(twice_gpu, twice_ora)
        assert ok, describe_mismatch(twice_gpu, twice_ora, bad)


def test_device_resident_entry_point(mapper, golden_cases, golden_index_dir):
    import torch
    case = next(c for c in golden_cases if c["name"] == "se_gapped")
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    mapper.index_load(prefix, 0)
    dev = torch.device("cuda:0")
    d_text = torch.from_numpy(chunk.text).to(dev)
    d_off = torch.from_numpy(chunk.off.astype(np.int64)).to(dev)
    d_len = torch.from_numpy(chunk.len.astype(np.int16)).to(dev)
    n = chunk.n_frags * chunk.mates * params.multi_best_reads
    d_out = torch.zeros(n * 68, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    mapper.vote_chunk_device(params, d_text.data_ptr(), d_text.numel(), d_off.data_ptr(), d_len.data_ptr(), chunk.n_frags, False,
                             int(chunk.len.max()), d_out.data_ptr())
    got = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=MAPPING_RESULT)
    assert got.tobytes() == exp.tobytes()
    assert mapper.timing().ms_kernel_vote > 0
    # out_is_zero: the slice is not read, so stale contents must not matter
    d_out.fill_(0xAB)
    torch.cuda.synchronize()
    mapper.vote_chunk_device(params, d_text.data_ptr(), d_text.numel(), d_off.data_ptr(), d_len.data_ptr(), chunk.n_frags, False,
                             int(chunk.len.max()), d_out.data_ptr(), out_is_zero=True)
    got = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=MAPPING_RESULT)
    assert got.tobytes() == exp.tobytes()
    # a read that leaves the text buffer is reported, never dereferenced
    from subread_b200.binding import SubreadCudaError
    d_off2 = d_off.clone()
    d_off2[3] = d_text.numel() - 5
    d_off2[7] = 1 << 62
    torch.cuda.synchronize()
    with pytest.raises(SubreadCudaError, match="outside the text buffer"):
        mapper.vote_chunk_device(params, d_text.data_ptr(), d_text.numel(), d_off2.data_ptr(), d_len.data_ptr(), chunk.n_frags, False,
                                 int(chunk.len.max()), d_out.data_ptr(), out_is_zero=True)


def test_errors_are_loud(mapper, golden_cases, golden_index_dir):
    from subread_b200.binding import Mapper, SubreadCudaError
    case = next(c for c in golden_cases if c["name"] == "se_gapped")
    prefix, chunk, params, _ = load_case(case, golden_index_dir)
    fresh = Mapper(0)
    with pytest.raises(SubreadCudaError, match="no index"):
        fresh.vote_chunk(chunk, params)
    with pytest.raises(SubreadCudaError):
        fresh.index_load("/nonexistent/prefix", 0)
    fresh.close()
    mapper.index_load(prefix, 0)
    bad = VoteParams.defaults(False)
    bad.multi_best_reads = 7
    with pytest.raises(SubreadCudaError, match="multi_best_reads"):
        mapper.vote_chunk(chunk, bad)
    sub = VoteParams.defaults(False, is_subjunc=True)
    sub.max_insertion_at_junctions = 3
    with pytest.raises(SubreadCudaError, match="max_insertion_at_junctions"):
        mapper.vote_chunk(chunk, sub, junc_state={})


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref (reference binaries) not built")
@pytest.mark.parametrize("kind", ["uniform", "repeats"])
def test_larger_genome_against_oracle(kind, mapper, oracle_mod, tmp_path):
    """2 Mbp genome indexed by the reference's own builder (gapped, one block), 30k SE + 15k PE reads
    with errors and indels.  The repeat-rich variant drives tables past 32 slots, i.e. through the
    overflow queue and the exact slow path."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_data
    fa = str(tmp_path / "g.fa")
    gen_data.write_genome(fa, 4, 500000, seed=77, repeats=600 if kind == "repeats" else 0)
    prefix = str(tmp_path / "idx")
    subprocess.check_call([os.path.join(REF_DIR, "bin", "subread-buildindex"), "-o", prefix, "-B", fa],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    gen_data.write_reads(fa, str(tmp_path / "se"), 30000, 150, 1, False, 0.01, 0.1, False, 0.001)
    gen_data.write_reads(fa, str(tmp_path / "pe"), 15000, 100, 2, True, 0.01, 0.1, True, 0.001)
    mapper.index_load(prefix, 0)
    idx = oracle_mod.OracleIndex(prefix)
    total_overflow = 0
    for r1, r2 in ((str(tmp_path / "se_R1.fastq"), None), (str(tmp_path / "pe_R1.fastq"), str(tmp_path / "pe_R2.fastq"))):
        chunk = load_fastq(r1, r2)
        params = VoteParams.defaults(r2 is not None)
        got, ver = mapper.vote_chunk(chunk, params, want_verify=True)
        total_overflow += mapper.timing().n_overflow
        ora, over, _ = oracle_mod.vote_chunk(idx, params, chunk, want_verify=True, threads=8)
        ok, bad = used_slots_equal(got, ora)
        assert ok, describe_mismatch(got, ora, bad)
        assert got.tobytes() == ora.tobytes()
        assert (ver == over).all()
    if kind == "repeats":
        assert total_overflow > 0, "the repeat-rich case is meant to exercise the overflow path"


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref (reference binaries) not built")
def test_multi_block_index_merges_across_blocks(mapper, oracle_mod, tmp_path):
    """read_chunk_circles (core.c:3567-3613) votes every chunk against each index block in turn and the
    results merge in the bigtable; here the host loop reloads the block on the GPU and carries out_align."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_data
    fa = str(tmp_path / "g.fa")
    gen_data.write_genome(fa, 3, 400000, seed=5, repeats=150)
    prefix = str(tmp_path / "idx")
    subprocess.check_call([os.path.join(REF_DIR, "bin", "subread-buildindex"), "-o", prefix, "-M", "1", fa],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    n_blocks = len([f for f in os.listdir(tmp_path) if f.endswith(".b.tab")])
    assert n_blocks > 1
    gen_data.write_reads(fa, str(tmp_path / "pe"), 5000, 120, 2, True, 0.01, 0.2, True, 0.002)
    gen_data.write_reads(fa, str(tmp_path / "se"), 8000, 150, 3, False, 0.005, 0.05, False, 0.001)
    for paired in (True, False):
        chunk = load_fastq(str(tmp_path / ("pe_R1.fastq" if paired else "se_R1.fastq")), str(tmp_path / "pe_R2.fastq") if paired else None)
        params = VoteParams.defaults(paired)
        got, ora = None, None
        for b in range(n_blocks):
            mapper.index_load(prefix, b)
            got, _ = mapper.vote_chunk(chunk, params, out_align=None if got is None else got.reshape(-1))
            idx = oracle_mod.OracleIndex(prefix, b)
            ora, _, _ = oracle_mod.vote_chunk(idx, params, chunk, out_align=ora)
            idx.close()
            assert got.tobytes() == ora.tobytes(), f"block {b} ({'PE' if paired else 'SE'})"


def test_index_builder_output_loads_and_maps(mapper, oracle_mod, tmp_path):
    """An index written by subread_b200.index_build (GPU sort) is usable by both sides and gives the
    golden result when built from the golden genome with the golden bucket count."""
    from subread_b200.index_build import build_index, calculate_buckets_by_size
    import json
    nb = calculate_buckets_by_size(131072, 3)
    prefix = str(tmp_path / "mine")
    build_index(os.path.join(GOLDEN, "tiny.fa.gz"), prefix, gapped=True, buckets=nb, use_gpu=True)
    cases = json.load(open(os.path.join(GOLDEN, "cases.json")))
    case = next(c for c in cases if c["name"] == "pe_gapped")
    _, chunk, params, exp = load_case(case, str(tmp_path))  # index dir unused
    mapper.index_load(prefix, 0)
    got, _ = mapper.vote_chunk(chunk, params)
    assert got.tobytes() == exp.tobytes()


@pytest.mark.parametrize("name", ["sj_se", "sj_pe"])
def test_subjunc_junction_voting(name, mapper, golden_cases, golden_index_dir, oracle_mod):
    """K_junc: subjunc's minor-half / donor-site search and big-margin records, through the
    C ABI, against the reference's golden tables and the oracle; then a second pass (read-modify-write)."""
    case = next(c for c in golden_cases if c["name"] == name)
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    exp_junc, exp_bm = load_case_junc(case, chunk, params)
    mapper.index_load(prefix, 0)
    js = {}
    got, _ = mapper.vote_chunk(chunk, params, junc_state=js)
    ok, bad = used_slots_equal(got, exp)
    assert ok, describe_mismatch(got, exp, bad)
    assert junc_tables_equal(got, js["junc"], js["big_margin"], exp_junc, exp_bm)
    idx = oracle_mod.OracleIndex(prefix)
    jo = {}
    ora, _, _ = oracle_mod.vote_chunk(idx, params, chunk, junc_state=jo)
    assert got.tobytes() == ora.tobytes()
    assert js["junc"].tobytes() == jo["junc"].tobytes()
    assert (js["big_margin"] == jo["big_margin"]).all()
    got2, _ = mapper.vote_chunk(chunk, params, out_align=got.copy().reshape(-1), junc_state=js)
    ora2, _, _ = oracle_mod.vote_chunk(idx, params, chunk, out_align=ora.copy(), junc_state=jo)
    assert got2.tobytes() == ora2.tobytes()
    assert js["junc"].tobytes() == jo["junc"].tobytes()
    assert (js["big_margin"] == jo["big_margin"]).all()


@pytest.mark.parametrize("kind", ["uniform", "repeats", "full", "dense", "dense_repeats", "dense_gapped", "dense_gapped_repeats"])
def test_tier0_single_end_chunks_against_oracle(kind, mapper, oracle_mod, tmp_path):
    """Tier 0 (k_map_simple, one thread per fragment) only runs for single-end chunks on a zero bigtable without
    verification -- the shape of the bench.  2 Mbp genome (uniform / repeat-rich / full index); error-free reads,
    reads with errors, indels, Ns and mixed lengths up to 300 bp (beyond the tier's 256), lower-case and IUPAC
    characters, read text with gaps between the reads.  Bit-exact against the oracle, with and without the tier."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_data
    from subread_b200 import pack_reads
    from subread_b200.binding import Mapper
    from subread_b200.index_build import build_index
    from subread_b200.reads import ReadChunk
    fa = str(tmp_path / "g.fa")
    gen_data.write_genome(fa, 4, 500000, seed=91, repeats=600 if kind.endswith("repeats") else 0)
    prefix = str(tmp_path / "idx")
    own = None
    if "gapped" in kind:
        # a gapped index with the dense-table accelerators forced onto it (SRCUDA_ACCEL=4): direct filter + the 128-byte phase
        # units, the thread-per-fragment tier's mode 4 -- what a human-sized gapped index gets by itself
        build_index(fa, prefix, gapped=True, use_gpu=True)
        os.environ["SRCUDA_ACCEL"] = "4"
        try:
            own = mapper = Mapper(0)
        finally:
            del os.environ["SRCUDA_ACCEL"]
    elif kind.startswith("dense"):
        # 2 M items in 262 147 buckets (7.6 per bucket, the shape of a human-sized table): the direct filter is chosen
        build_index(fa, prefix, gapped=False, buckets=262147, use_gpu=True)
    else:
        build_index(fa, prefix, gapped=(kind != "full"), use_gpu=True)
    mapper.index_load(prefix, 0)
    acc = mapper.describe().accelerators
    assert acc & 7 == (4 if kind.startswith("dense") else 3), "direct filter on dense tables only"
    assert bool(acc & 16) == (kind in ("uniform", "repeats", "dense_gapped", "dense_gapped_repeats")), "phase table on gapped indexes only"
    idx = oracle_mod.OracleIndex(prefix)
    params = VoteParams.defaults(False)
    gen_data.write_reads(fa, str(tmp_path / "clean"), 40000, 150, 3, False, 0.0, 0.0, False, 0.0)
    gen_data.write_reads(fa, str(tmp_path / "noisy"), 30000, 150, 4, False, 0.01, 0.1, True, 0.002)
    gen_data.write_reads(fa, str(tmp_path / "short"), 10000, 50, 5, False, 0.005, 0.0, False, 0.0)
    os.environ["SRCUDA_SIMPLE_TIER"] = "0"
    try:
        plain = Mapper(0)
    finally:
        del os.environ["SRCUDA_SIMPLE_TIER"]
    plain.index_load(prefix, 0)
    chunks = {n: load_fastq(str(tmp_path / f"{n}_R1.fastq")) for n in ("clean", "noisy", "short")}
    # mixed lengths 16..300, odd characters, and a scattered layout (reads far apart, unordered offsets)
    rng = np.random.default_rng(8)
    _, seqs = gen_data.read_fasta(fa)
    g = seqs[1]
    odd = []
    for i in range(3000):
        L = int(rng.integers(16, 301))
        s = int(rng.integers(0, g.shape[0] - L))
        r = bytearray(g[s:s + L].tobytes())
        if i % 7 == 0:
            r[int(rng.integers(0, L))] = ord("acgtnRYU."[int(rng.integers(0, 9))])
        if i % 11 == 0:
            r = bytearray(gen_data.revcomp(np.frombuffer(bytes(r), dtype=np.uint8)).tobytes())
        odd.append(bytes(r))
    odd += [b"ACGT", b"A" * 15, b"", g[100:116].tobytes(), g[5000:6210].tobytes()]
    chunks["odd"] = pack_reads(odd, None)
    c = chunks["clean"]
    n = 5000
    text = np.full(n * 700 + 64, ord("#"), dtype=np.uint8)
    off = np.zeros(n + 1, dtype=np.uint64)
    for i in range(n):
        o = 3 + i * 700 + int(rng.integers(0, 500))
        off[i] = o
        text[o:o + 150] = c.text[int(c.off[i]):int(c.off[i]) + 150]
    off[n] = text.shape[0]
    chunks["scattered"] = ReadChunk(text, off, np.full(n, 150, np.uint16), n, False, [])
    for name, chunk in chunks.items():
        ora, _, _ = oracle_mod.vote_chunk(idx, params, chunk, threads=8)
        got, _ = mapper.vote_chunk(chunk, params)
        t = mapper.timing()
        ok, bad = used_slots_equal(got, ora)
        assert ok, name + ": " + describe_mismatch(got, ora, bad)
        assert got.tobytes() == ora.tobytes(), name
        ref, _ = plain.vote_chunk(chunk, params)
        assert ref.tobytes() == ora.tobytes(), name + " (tier 0 off)"
        assert plain.timing().n_simple_punted == 0
        # a second pass over the same table, as a later index block would make it (read-modify-write form of the tier)
        ora2, _, _ = oracle_mod.vote_chunk(idx, params, chunk, out_align=ora.copy(), threads=8)
        got2, _ = mapper.vote_chunk(chunk, params, out_align=got.copy().reshape(-1))
        t2 = mapper.timing()
        assert got2.tobytes() == ora2.tobytes(), name + " (second pass)"
        ref2, _ = plain.vote_chunk(chunk, params, out_align=ref.copy().reshape(-1))
        assert ref2.tobytes() == ora2.tobytes(), name + " (second pass, tier 0 off)"
        if name in ("clean", "scattered") and not kind.endswith("repeats"):
            assert t2.ms_kernel_simple > 0 and t2.n_simple_punted < chunk.n_frags // 10, f"{name}: second pass queued {t2.n_simple_punted}"
        if name in ("clean", "scattered") and not kind.endswith("repeats"):
            assert t.n_simple_punted < chunk.n_frags // 20, f"{name}: tier 0 punted {t.n_simple_punted} of {chunk.n_frags}"
    plain.close()
    if own is not None:
        own.close()


@pytest.mark.parametrize("force", [False, True], ids=["auto", "forced_on_sparse"])
def test_direct_filter_warp_kernel_pe_and_se(force, oracle_mod, tmp_path):
    """The direct filter of dense tables in the warp kernel (paired-end, verification, reads with errors and indels) and
    with SRCUDA_ACCEL=4 forced onto a sparse gapped table; a repeat-rich genome puts many keys on the several-positions
    path and tables into the overflow queue.  Bit-exact against the oracle."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_data
    from subread_b200.binding import Mapper
    from subread_b200.index_build import build_index
    fa = str(tmp_path / "g.fa")
    gen_data.write_genome(fa, 4, 500000, seed=123, repeats=500)
    prefix = str(tmp_path / "idx")
    if force:
        build_index(fa, prefix, gapped=True, use_gpu=True)
        os.environ["SRCUDA_ACCEL"] = "4"
    else:
        build_index(fa, prefix, gapped=False, buckets=262147, use_gpu=True)
    try:
        m = Mapper(0)
    finally:
        os.environ.pop("SRCUDA_ACCEL", None)
    m.index_load(prefix, 0)
    assert m.describe().accelerators & 7 == 4
    idx = oracle_mod.OracleIndex(prefix)
    gen_data.write_reads(fa, str(tmp_path / "se"), 20000, 150, 1, False, 0.01, 0.1, False, 0.001)
    gen_data.write_reads(fa, str(tmp_path / "pe"), 15000, 100, 2, True, 0.01, 0.1, True, 0.001)
    for r1, r2 in ((str(tmp_path / "se_R1.fastq"), None), (str(tmp_path / "pe_R1.fastq"), str(tmp_path / "pe_R2.fastq"))):
        chunk = load_fastq(r1, r2)
        params = VoteParams.defaults(r2 is not None)
        for verify in (True, False):        # without verification single-end chunks start in tier 0
            got, ver = m.vote_chunk(chunk, params, want_verify=verify)
            ora, over, _ = oracle_mod.vote_chunk(idx, params, chunk, want_verify=verify, threads=8)
            ok, bad = used_slots_equal(got, ora)
            assert ok, describe_mismatch(got, ora, bad)
            assert got.tobytes() == ora.tobytes()
            if verify:
                assert (ver == over).all()
    m.close()


@pytest.mark.parametrize("kind", ["uniform", "repeats", "dense", "dense_gapped"])
def test_tier0_paired_end_chunks_against_oracle(kind, oracle_mod, tmp_path):
    """The paired-end form of tier 0 (k_map_simple_pe: one thread per fragment, pair scoring per thread) on a zero
    bigtable: clean, noisy (errors, indels, Ns, mixed lengths) and short pairs, pairs with odd mates (shorter than 16
    bases, empty, lower case / IUPAC, longer than the tier takes), mates from different contigs and improper distances,
    non-default pair distance / strand options.  Bit-exact against the oracle, and identical with the tier switched off."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_data
    from subread_b200 import pack_reads
    from subread_b200.binding import Mapper
    from subread_b200.index_build import build_index
    fa = str(tmp_path / "g.fa")
    gen_data.write_genome(fa, 4, 500000, seed=191, repeats=600 if kind == "repeats" else 0)
    prefix = str(tmp_path / "idx")
    if kind == "dense":
        build_index(fa, prefix, gapped=False, buckets=262147, use_gpu=True)
    else:
        build_index(fa, prefix, gapped=True, use_gpu=True)
    if kind == "dense_gapped":
        os.environ["SRCUDA_ACCEL"] = "4"                  # dense accelerators on a gapped index: the 128-byte phase units (mode 4)
    try:
        tier = Mapper(0)
    finally:
        os.environ.pop("SRCUDA_ACCEL", None)
    tier.index_load(prefix, 0)
    os.environ["SRCUDA_SIMPLE_TIER_PE"] = "0"
    try:
        plain = Mapper(0)
    finally:
        del os.environ["SRCUDA_SIMPLE_TIER_PE"]
    plain.index_load(prefix, 0)
    alt = None
    if kind == "dense_gapped":
        # the other forms of the same tier: four-word clusters for every read length (the default takes three-word clusters for
        # reads of at most 160 bases), a small cluster pool (more fragments queued), results copied down uncompacted
        os.environ.update(SRCUDA_ACCEL="4", SRCUDA_PE_CW3="0", SRCUDA_PE_POOL="12", SRCUDA_COMPACT_D2H="0")
        try:
            alt = Mapper(0)
        finally:
            for k in ("SRCUDA_ACCEL", "SRCUDA_PE_CW3", "SRCUDA_PE_POOL", "SRCUDA_COMPACT_D2H"):
                os.environ.pop(k, None)
        alt.index_load(prefix, 0)
    idx = oracle_mod.OracleIndex(prefix)
    gen_data.write_reads(fa, str(tmp_path / "clean"), 30000, 150, 13, True, 0.0, 0.0, False, 0.0)
    gen_data.write_reads(fa, str(tmp_path / "noisy"), 20000, 100, 14, True, 0.01, 0.1, True, 0.002)
    gen_data.write_reads(fa, str(tmp_path / "short"), 6000, 50, 15, True, 0.005, 0.0, False, 0.0)
    chunks = {n: load_fastq(str(tmp_path / f"{n}_R1.fastq"), str(tmp_path / f"{n}_R2.fastq")) for n in ("clean", "noisy", "short")}
    rng = np.random.default_rng(18)
    _, seqs = gen_data.read_fasta(fa)
    m1, m2 = [], []
    for i in range(4000):
        c = int(rng.integers(0, len(seqs)))
        g = seqs[c]
        La, Lb = int(rng.integers(14, 301)), int(rng.integers(14, 301))
        s = int(rng.integers(0, g.shape[0] - 2000))
        ins = int(rng.integers(max(La, Lb), 900))                      # proper and improper distances (-D 600)
        a = bytearray(g[s:s + La].tobytes())
        g2 = seqs[int(rng.integers(0, len(seqs)))] if i % 13 == 0 else g   # mates on different contigs
        b = bytearray(gen_data.revcomp(g2[s + ins - Lb:s + ins]).tobytes())
        if i % 7 == 0:
            a[int(rng.integers(0, La))] = ord("acgtnRYU."[int(rng.integers(0, 9))])
        if i % 9 == 0:
            a, b = b, a                                                    # the pair on the other strand
        if i % 17 == 0:
            b = bytearray(gen_data.revcomp(np.frombuffer(bytes(b), dtype=np.uint8)).tobytes())   # FF orientation
        m1.append(bytes(a)); m2.append(bytes(b))
    m1 += [b"ACGT", b"", seqs[0][100:250].tobytes(), seqs[0][5000:6210].tobytes()]
    m2 += [seqs[0][300:450].tobytes(), seqs[1][10:160].tobytes(), b"A" * 15, gen_data.revcomp(seqs[0][6300:6450]).tobytes()]
    chunks["odd"] = pack_reads(m1, m2)
    variants = [VoteParams.defaults(True)]
    p2 = VoteParams.defaults(True)
    p2.minimum_pair_distance, p2.maximum_pair_distance, p2.is_second_read_reversed = 100, 300, 0
    p2.minimum_subread_for_first_read, p2.minimum_subread_for_second_read = 2, 2
    variants.append(p2)
    for name, chunk in chunks.items():
        for vi, params in enumerate(variants):
            ora, _, _ = oracle_mod.vote_chunk(idx, params, chunk, threads=8)
            got, _ = tier.vote_chunk(chunk, params)
            t = tier.timing()
            ok, bad = used_slots_equal(got, ora)
            assert ok, f"{name}/{vi}: " + describe_mismatch(got, ora, bad)
            assert got.tobytes() == ora.tobytes(), f"{name}/{vi}"
            ref, _ = plain.vote_chunk(chunk, params)
            assert ref.tobytes() == ora.tobytes(), f"{name}/{vi} (tier off)"
            assert plain.timing().n_simple_punted == 0
            if alt is not None:
                other, _ = alt.vote_chunk(chunk, params)
                assert other.tobytes() == ora.tobytes(), f"{name}/{vi} (four-word clusters, small pool, plain copy)"
            ora2, _, _ = oracle_mod.vote_chunk(idx, params, chunk, out_align=ora.copy(), threads=8)
            got2, _ = tier.vote_chunk(chunk, params, out_align=got.copy().reshape(-1))
            assert got2.tobytes() == ora2.tobytes(), f"{name}/{vi} (second pass)"
            if name == "clean" and kind != "repeats" and vi == 0:
                assert tier.timing().n_simple_punted < chunk.n_frags // 5, "second pass: read-modify-write tier"
            if name == "clean" and kind != "repeats" and vi == 0:
                assert t.n_simple_punted < chunk.n_frags // 10, f"tier 0 queued {t.n_simple_punted} of {chunk.n_frags} clean pairs"
    tier.close()
    plain.close()
    if alt is not None:
        alt.close()


@pytest.mark.parametrize("paired", [False, True], ids=["se", "pe"])
def test_subjunc_spliced_reads_warp_and_scalar_paths(paired, oracle_mod, tmp_path):
    """subjunc on a 2 Mbp genome with planted GT..AG / CT..AC introns and 20 k reads from the spliced transcripts (errors,
    mixed lengths, reads longer than 160 bp): the thread-per-fragment tier's and the warp kernel's junction search
    (tier_minor_half, warp_minor_half, big-margin records) and the scalar path (SRCUDA_WARP_JUNC=0) against the oracle -- bigtable, subjunc_result_t and big-margin tables byte for byte;
    then a second pass over the same tables (read-modify-write)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_data
    from subread_b200.binding import Mapper
    from subread_b200.index_build import build_index
    fa, faS = str(tmp_path / "g.fa"), str(tmp_path / "gS.fa")
    gen_data.write_genome(fa, 4, 500000, seed=77, repeats=100)
    gen_data.write_spliced(fa, faS, str(tmp_path / "sj"), 14000, 100, 5, paired, mixed_len=True)
    gen_data.write_spliced(fa, faS, str(tmp_path / "sjl"), 6000, 200, 5, paired)       # same seed: the same introns
    prefix = str(tmp_path / "idx")
    build_index(faS, prefix, gapped=True, use_gpu=True)
    idx = oracle_mod.OracleIndex(prefix)
    params = VoteParams.defaults(paired, True)
    mappers = {"warp": Mapper(0)}
    os.environ["SRCUDA_WARP_JUNC"] = "0"
    try:
        mappers["scalar"] = Mapper(0)
    finally:
        del os.environ["SRCUDA_WARP_JUNC"]
    n_junc = 0
    for name in ("sj", "sjl"):
        chunk = load_fastq(str(tmp_path / f"{name}_R1.fastq"), str(tmp_path / f"{name}_R2.fastq") if paired else None)
        jo = {}
        ora, _, _ = oracle_mod.vote_chunk(idx, params, chunk, junc_state=jo, threads=8)
        n_junc += int((jo["junc"]["minor_votes"] > 0).sum())
        for which, m in mappers.items():
            m.index_load(prefix, 0)
            js = {}
            got, _ = m.vote_chunk(chunk, params, junc_state=js)
            t = m.timing()
            assert got.tobytes() == ora.tobytes(), f"{name}/{which}"
            assert js["junc"].tobytes() == jo["junc"].tobytes(), f"{name}/{which}"
            assert (js["big_margin"] == jo["big_margin"]).all(), f"{name}/{which}"
            assert (t.n_overflow == chunk.n_frags) == (which == "scalar"), f"{name}/{which}: {t.n_overflow} of {chunk.n_frags} on the scalar path"
            if which == "warp" and name == "sj":    # ... after the thread-per-fragment tier took the separable fragments
                assert 0 < t.n_simple_punted < 0.7 * chunk.n_frags, f"tier queued {t.n_simple_punted} of {chunk.n_frags}"
            jo2 = {k: v.copy() for k, v in jo.items()}
            got2, _ = m.vote_chunk(chunk, params, out_align=got.copy().reshape(-1), junc_state=js)
            ora2, _, _ = oracle_mod.vote_chunk(idx, params, chunk, out_align=ora.copy(), junc_state=jo2, threads=8)
            assert got2.tobytes() == ora2.tobytes() and js["junc"].tobytes() == jo2["junc"].tobytes(), f"{name}/{which} second pass"
            assert (js["big_margin"] == jo2["big_margin"]).all()
    assert n_junc > 1000, "the spliced reads are meant to exercise the junction search"
    for m in mappers.values():
        m.close()


@pytest.mark.gpu
def test_mapping_straight_from_fastq_bytes(mapper, golden_cases, golden_index_dir):
    """Row f3: the chunk buffer is the FASTQ file's own bytes -- srcuda_fastq_split indexes the records in place
    (names, '+' lines and qualities stay between the reads) and srcuda_vote_chunk maps them from there; the bigtable
    is the golden one of the same reads."""
    import gzip
    from conftest import GOLDEN
    from subread_b200 import ReadChunk
    from subread_b200.binding import fastq_split
    case = next(c for c in golden_cases if c["name"] == "se_gapped")
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    with gzip.open(os.path.join(GOLDEN, case["r1"]), "rb") as fh:
        raw = np.frombuffer(fh.read(), dtype=np.uint8)
    off, ln, b, _ = fastq_split(raw)
    assert b.n_reads == chunk.n_frags and np.array_equal(ln, chunk.len)
    in_place = ReadChunk(raw, np.append(off, np.uint64(raw.shape[0])), ln, chunk.n_frags, False, [])
    mapper.index_load(prefix, 0)
    got, _ = mapper.vote_chunk(in_place, params)
    assert got.tobytes() == exp.tobytes()
    got, ver = mapper.vote_chunk(in_place, params, want_verify=True)      # the warp kernel reads the same buffer
    assert got.tobytes() == exp.tobytes()
