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
    src = "".join(open(os.path.join(ROOT, "include", f)).read() for f in sorted(os.listdir(os.path.join(ROOT, "include"))) if f.endswith(".h"))
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


def test_null_arguments_are_argument_errors(lib):
    """Every entry point checks its pointers before touching a device: a NULL handle (or out pointer) is
    SRCUDA_ERR_ARG (5) or a no-op, never a crash -- the error behaviour include/subread_cuda.h documents.
    No device work is started, so this runs without a GPU."""
    ERR_ARG = 5
    null = ctypes.c_void_p(None)
    assert lib.srcuda_open(0, None) == ERR_ARG
    assert b"out is NULL" in lib.srcuda_last_error(None)
    assert lib.srcuda_index_load(null, b"/nonexistent/prefix", 0) == ERR_ARG
    assert lib.srcuda_index_replicate(null, null) == ERR_ARG
    blob = ctypes.create_string_buffer(4096)
    assert lib.srcuda_index_export_ipc(null, blob, ctypes.c_size_t(4096)) == ERR_ARG
    assert lib.srcuda_index_import_ipc(null, blob, ctypes.c_size_t(4096)) == ERR_ARG
    desc = binding.IndexDesc()
    assert lib.srcuda_index_describe(null, ctypes.byref(desc)) == ERR_ARG
    assert lib.srcuda_set_force_slow_path(null, 1) == ERR_ARG
    assert lib.srcuda_get_timing(null, None) == ERR_ARG
    params = VoteParams.defaults(False)
    assert lib.srcuda_vote_chunk(null, ctypes.byref(params), None, None, None, ctypes.c_int64(0), 0, 1, None, None, None, None) == ERR_ARG
    gb = ctypes.c_double()
    assert lib.srcuda_measure_gather_bandwidth(null, ctypes.c_uint64(1 << 22), 1, ctypes.byref(gb)) == ERR_ARG
    n = ctypes.c_int64()
    assert lib.srcuda_reader_open(None, None, 0, 1, 0, 0, None) == ERR_ARG
    assert lib.srcuda_reader_next(null, 1, None, 0, None, None, ctypes.byref(n), None) == ERR_ARG
    assert lib.srcuda_reader_skip(null, 1, ctypes.byref(n)) == ERR_ARG
    assert lib.srcuda_reader_error(null) == b"null reader"
    lib.srcuda_reader_close(null)
    batch = binding.FastqBatch()
    assert lib.srcuda_fastq_split(None, 10, 1, 0, 0, 1, 0, 1, None, None, None, None, None, None, ctypes.byref(batch)) == ERR_ARG
    assert lib.srcuda_reads_pack(None, None, None, None, None, None, 1, None, 0, None, None, None) == ERR_ARG
    lib.srcuda_index_unload(null)      # no-ops
    lib.srcuda_close(null)
    lib.srcuda_host_free(None)
