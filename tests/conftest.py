import gzip
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "ref_dump")) and os.path.exists(os.path.join(REF_DIR, "libsubread_ref.so"))


@pytest.fixture(scope="session")
def golden_index_dir(tmp_path_factory):
    """Unpacks the committed golden index blocks (gz) into a session temp dir; returns that dir."""
    d = tmp_path_factory.mktemp("golden_idx")
    for name in ("tiny", "tinyF", "tinyS"):
        for ext in (".00.b.tab", ".00.b.array"):
            with gzip.open(os.path.join(GOLDEN, name + ext + ".gz"), "rb") as a, open(d / (name + ext), "wb") as b:
                shutil.copyfileobj(a, b)
        shutil.copy(os.path.join(GOLDEN, name + ".reads"), d / (name + ".reads"))
    return str(d)


@pytest.fixture(scope="session")
def golden_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as fh:
        return json.load(fh)


def load_case(case, golden_index_dir):
    """-> (index_prefix, ReadChunk, VoteParams, expected align array); subjunc cases: see load_case_junc"""
    from subread_b200 import MAPPING_RESULT, VoteParams, load_fastq
    chunk = load_fastq(os.path.join(GOLDEN, case["r1"]), os.path.join(GOLDEN, case["r2"]) if case["r2"] else None)
    p = VoteParams.defaults(case["r2"] is not None, bool(case.get("subjunc")))
    for k, v in case["params"].items():
        setattr(p, k, v)
    n = chunk.n_frags * chunk.mates * p.multi_best_reads
    with gzip.open(os.path.join(GOLDEN, f"expect_{case['name']}.bin.gz"), "rb") as fh:
        exp = np.frombuffer(fh.read(n * 68), dtype=MAPPING_RESULT).reshape(chunk.n_frags, chunk.mates, p.multi_best_reads)
    return os.path.join(golden_index_dir, case["index"]), chunk, p, exp


def load_case_junc(case, chunk, params):
    """Expected subjunc_result_t table and big-margin records (9 live u16 per read) of a subjunc golden case."""
    from subread_b200 import SUBJUNC_RESULT
    n = chunk.n_frags * chunk.mates * params.multi_best_reads
    with gzip.open(os.path.join(GOLDEN, f"expect_{case['name']}.bin.gz"), "rb") as fh:
        fh.read(n * 68)
        junc = np.frombuffer(fh.read(n * 16), dtype=SUBJUNC_RESULT)
        bm = np.frombuffer(fh.read(chunk.n_frags * chunk.mates * 18), dtype="<u2").reshape(-1, 9)
    return junc, bm


def junc_tables_equal(align, junc, bm, exp_junc, exp_bm):
    """subjunc parity rule: for used alignment slots the whole subjunc_result_t must match when it holds a
    junction (minor_votes > 0) and minor_votes itself must match otherwise (core-junction.c:2524 only
    clears minor_votes of unused slots); the big-margin records must match entirely."""
    used = align.reshape(-1)["selected_votes"] > 0
    j, e = junc.reshape(-1), exp_junc.reshape(-1)
    ok = bool((j["minor_votes"][used] == e["minor_votes"][used]).all())
    has = used & (e["minor_votes"] > 0)
    ok = ok and bool((j.view(np.uint8).reshape(-1, 16)[has] == e.view(np.uint8).reshape(-1, 16)[has]).all())
    ok = ok and bool((bm.reshape(-1, 27)[:, :9] == exp_bm).all())
    return ok


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def scalar_host_lib():
    """The product's exact slow path compiled as host C++ (tests/native/scalar_host.cpp)."""
    import ctypes
    src = os.path.join(ROOT, "tests", "native", "scalar_host.cpp")
    so = os.path.join(ROOT, "tests", "native", "libscalar_host.so")
    deps = [src] + [os.path.join(ROOT, "subread_b200", "csrc", f) for f in ("sr_scalar.cuh", "sr_common.cuh", "sr_config.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    return ctypes.CDLL(so)


def describe_mismatch(got, exp, bad):
    per = got.shape[1] * got.shape[2]
    f = bad // per
    return f"first mismatch at flat slot {bad} (fragment {f}):\n got {got[f]}\n exp {exp[f]}"
