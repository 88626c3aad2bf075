"""C-ABI surface: the shared library loads, exports every symbol include/subread_cuda.h declares,
the POD layouts have the reference's sizes, and -- on a box without a GPU -- the compute entry
points fail loudly instead of falling back to a CPU path.  No compute calls here."""
import ctypes
import os
import re

import pytest

from subread_b200 import MAPPING_RESULT, SUBJUNC_RESULT, VoteParams
from subread_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(binding.LIB_PATH):
        binding.build_library()
    return binding.load_library()


def declared_functions():
    src = open(os.path.join(ROOT, "include", "subread_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(srcuda_[a-z_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(binding.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"libsubread_cuda.so does not export {name}"


def test_pod_sizes_match_reference_structs():
    assert MAPPING_RESULT.itemsize == 68       # mapping_result_t, reference src/core.h:350-370
    assert SUBJUNC_RESULT.itemsize == 16       # subjunc_result_t, reference src/core.h:397-410
    assert ctypes.sizeof(VoteParams) == 24 * 4
    assert ctypes.sizeof(binding.IndexDesc) == 64
    assert MAPPING_RESULT.fields["selected_votes"][1] == 8
    assert MAPPING_RESULT.fields["selected_indel_record"][1] == 16
    assert MAPPING_RESULT.fields["confident_coverage_start"][1] == 60


def test_no_cpu_fallback_without_gpu(lib):
    n = lib.srcuda_device_count()
    if n > 0:
        pytest.skip("a GPU is visible; the loud-failure path is exercised on CPU boxes")
    h = ctypes.c_void_p()
    rc = lib.srcuda_open(0, ctypes.byref(h))
    assert rc == 1  # SRCUDA_ERR_NO_DEVICE
    assert b"no CUDA device" in lib.srcuda_last_error(None)
    with pytest.raises(binding.SubreadCudaError):
        binding.Mapper(0)


def test_product_never_references_the_oracle():
    """The shipped sources must not include, link or import anything under oracle/."""
    pkg = os.path.join(ROOT, "subread_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in text and "from oracle" not in text and "import oracle" not in text, f
