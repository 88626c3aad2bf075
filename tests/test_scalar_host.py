"""The product's exact slow path (subread_b200/csrc/sr_scalar.cuh), compiled as host C++ for this
test only, against the oracle and the golden bigtables.  CPU only."""
import ctypes

import numpy as np
import pytest

from conftest import describe_mismatch, junc_tables_equal, load_case, load_case_junc
from subread_b200 import MAPPING_RESULT, used_slots_equal


def run_scalar(lib, arrays, params, chunk, out=None, verify=False, junc_state=None):
    from subread_b200 import SUBJUNC_RESULT
    mb = params.multi_best_reads
    n = chunk.n_frags * chunk.mates * mb
    out = np.zeros(n, dtype=MAPPING_RESULT) if out is None else out.reshape(-1)
    ver = np.zeros(n, dtype=np.uint16) if verify else None
    junc = bm = None
    if params.do_breakpoint_detection:
        if "junc" not in junc_state:
            junc_state["junc"] = np.zeros(n, dtype=SUBJUNC_RESULT)
            junc_state["big_margin"] = np.zeros(chunk.n_frags * chunk.mates * 27, dtype=np.uint16)
        junc, bm = junc_state["junc"], junc_state["big_margin"]
    vp = lambda x: x.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    a = arrays
    rc = lib.sr_scalar_host_vote_chunk(
        ctypes.c_int32(a["index_gap"]), ctypes.c_int32(a["padding"]), ctypes.c_uint32(a["buckets_number"]),
        vp(a["bucket_start"]), vp(a["keys"]), vp(a["vals"]), ctypes.c_uint32(a["start_point"]), ctypes.c_uint32(a["length"]),
        vp(a["values"]), ctypes.c_uint32(a["values"].shape[0]), vp(a["read_offsets"]),
        ctypes.c_int32(a["read_offsets"].shape[0]), ctypes.byref(params), vp(chunk.text), vp(chunk.off), vp(chunk.len),
        ctypes.c_int64(chunk.n_frags), ctypes.c_int(chunk.is_paired), vp(out), vp(junc) if junc is not None else None,
        vp(bm) if bm is not None else None, vp(ver) if verify else None)
    assert rc == 0, rc
    return out.reshape(chunk.n_frags, chunk.mates, mb), ver


@pytest.mark.parametrize("name", ["se_gapped", "pe_gapped", "se_full", "pe_full", "se_I16", "pe_n14_ff"])
def test_scalar_path_matches_golden_and_oracle(name, golden_cases, golden_index_dir, oracle_mod, scalar_host_lib):
    case = next(c for c in golden_cases if c["name"] == name)
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    idx = oracle_mod.OracleIndex(prefix)
    got, ver = run_scalar(scalar_host_lib, idx.arrays(), params, chunk, verify=True)
    ok, bad = used_slots_equal(got, exp)
    assert ok, describe_mismatch(got, exp, bad)
    assert got.tobytes() == exp.tobytes()
    _, over, _ = oracle_mod.vote_chunk(idx, params, chunk, want_verify=True)
    assert (ver == over).all()


def test_scalar_rejects_unsupported_params(golden_cases, golden_index_dir, oracle_mod, scalar_host_lib):
    case = next(c for c in golden_cases if c["name"] == "se_gapped")
    prefix, chunk, params, _ = load_case(case, golden_index_dir)
    idx = oracle_mod.OracleIndex(prefix)
    params.multi_best_reads = 9
    with pytest.raises(AssertionError):
        run_scalar(scalar_host_lib, idx.arrays(), params, chunk)


@pytest.mark.parametrize("name", ["sj_se", "sj_pe"])
def test_scalar_path_subjunc(name, golden_cases, golden_index_dir, oracle_mod, scalar_host_lib):
    case = next(c for c in golden_cases if c["name"] == name)
    prefix, chunk, params, exp = load_case(case, golden_index_dir)
    exp_junc, exp_bm = load_case_junc(case, chunk, params)
    idx = oracle_mod.OracleIndex(prefix)
    js = {}
    got, _ = run_scalar(scalar_host_lib, idx.arrays(), params, chunk, junc_state=js)
    ok, bad = used_slots_equal(got, exp)
    assert ok, describe_mismatch(got, exp, bad)
    assert junc_tables_equal(got, js["junc"], js["big_margin"], exp_junc, exp_bm)
    jo = {}
    ora, _, _ = oracle_mod.vote_chunk(idx, params, chunk, junc_state=jo)
    assert got.tobytes() == ora.tobytes() and js["junc"].tobytes() == jo["junc"].tobytes()
