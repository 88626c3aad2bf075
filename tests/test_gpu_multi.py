"""One chunk over several GPUs: srcuda_vote_chunk_multi cuts the chunk into contiguous fragment slices, one
host thread per handle, disjoint slices of ONE host bigtable; the index is loaded once and replicated with peer copies
(srcuda_index_replicate / srcuda_index_load_all).  The bytes must equal the single-handle run and the oracle's.  With one
visible GPU the handles share it (two handles = two sets of streams and two copies of the index); with more GPUs every handle
gets its own -- `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` is the two-GPU form."""
import ctypes
import os

import numpy as np
import pytest

from conftest import load_case, load_case_junc, junc_tables_equal
from subread_b200 import VoteParams, load_fastq, used_slots_equal
from subread_b200 import binding
from subread_b200.sharding import shard_range

pytestmark = pytest.mark.gpu


def devices(n):
    have = binding.load_library().srcuda_device_count()
    return [k % max(have, 1) for k in range(n)]


@pytest.mark.parametrize("n_handles", [2, 3])
@pytest.mark.parametrize("case_name", ["se_gapped", "pe_gapped", "sj_pe"])
def test_chunk_over_several_handles_equals_one_handle(case_name, n_handles, golden_cases, golden_index_dir, oracle_mod):
    case = next(c for c in golden_cases if c["name"] == case_name)
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    subjunc = bool(case.get("subjunc"))
    one = binding.Mapper(0)
    one.index_load(prefix, 0)
    js1 = {} if subjunc else None
    ref, _ = one.vote_chunk(chunk, params, junc_state=js1)
    mm = binding.MultiMapper(devices(n_handles))
    mm.index_load(prefix, 0)                      # srcuda_index_load_all: load on the first handle, peer copies to the others
    d0, dk = mm.mappers[0].describe(), mm.mappers[-1].describe()
    assert (d0.n_items, d0.buckets_number, d0.length, d0.accelerators) == (dk.n_items, dk.buckets_number, dk.length, dk.accelerators)
    jsm = {} if subjunc else None
    got, _ = mm.vote_chunk(chunk, params, junc_state=jsm)
    assert got.tobytes() == ref.tobytes(), "the sliced chunk differs from the single-handle run"
    assert used_slots_equal(got, exp)[0], "... and from the reference's golden bigtable"
    if subjunc:
        ej, eb = load_case_junc(case, chunk, params)
        assert junc_tables_equal(got, jsm["junc"], jsm["big_margin"], ej, eb)
        assert jsm["junc"].tobytes() == js1["junc"].tobytes() and jsm["big_margin"].tobytes() == js1["big_margin"].tobytes()
    # every handle mapped its slice, and only its slice
    ts = mm.timings()
    for k, t in enumerate(ts):
        f0, f1 = shard_range(chunk.n_frags, n_handles, k)
        n_reads = (f1 - f0) * chunk.mates
        assert t.h2d_bytes >= int(chunk.len[f0 * chunk.mates:f1 * chunk.mates].sum()) + 10 * n_reads      # its reads, offsets and lengths went up
        assert 0 < t.d2h_bytes <= n_reads * params.multi_best_reads * 68 * (2 if subjunc else 1) + 64 * n_reads   # ... its results came down (compacted on a zero table)
        assert t.n_batches >= 1
    # a second pass over the same table (later index block): read-modify-write through the slices
    again, _ = mm.vote_chunk(chunk, params, out_align=got.copy().reshape(-1), junc_state=jsm)
    ref2, _ = one.vote_chunk(chunk, params, out_align=ref.copy().reshape(-1), junc_state=js1)
    assert again.tobytes() == ref2.tobytes()
    # replicate by hand as well (srcuda_index_replicate handle to handle)
    mm.index_replicate()
    third, _ = mm.vote_chunk(chunk, params, junc_state={} if subjunc else None)
    assert third.tobytes() == ref.tobytes()
    mm.close()
    one.close()


def test_multi_driver_argument_errors(golden_cases, golden_index_dir):
    case = next(c for c in golden_cases if c["name"] == "se_gapped")
    prefix, chunk, params, _ = load_case(case, golden_index_dir)
    mm = binding.MultiMapper(devices(2))
    with pytest.raises(binding.SubreadCudaError):      # no index on any handle
        mm.vote_chunk(chunk, params)
    mm.mappers[0].index_load(prefix, 0)
    with pytest.raises(binding.SubreadCudaError):      # ... on the first one only
        mm.vote_chunk(chunk, params)
    lib = mm.lib
    arr = (ctypes.c_void_p * 2)(mm.mappers[0].h, mm.mappers[0].h)
    out = np.zeros(chunk.n_frags * 3, dtype=binding.MAPPING_RESULT)
    rc = lib.srcuda_vote_chunk_multi(arr, 2, ctypes.byref(params), chunk.text.ctypes.data_as(ctypes.c_void_p), chunk.off.ctypes.data_as(ctypes.c_void_p),
                                     chunk.len.ctypes.data_as(ctypes.c_void_p), chunk.n_frags, 0, 1, out.ctypes.data_as(ctypes.c_void_p), None, None, None)
    assert rc == 5                                     # SRCUDA_ERR_ARG: one handle listed twice
    # fewer fragments than handles: empty slices are fine
    mm.index_replicate()
    tiny = load_fastq(os.path.join(os.path.dirname(__file__), "golden", case["r1"]), limit=1)
    one = binding.Mapper(0)
    one.index_load(prefix, 0)
    a, _ = mm.vote_chunk(tiny, params)
    b, _ = one.vote_chunk(tiny, params)
    assert a.tobytes() == b.tobytes()
    mm.close()
    one.close()
