"""ctypes binding of libsubread_cuda.so (C ABI: include/subread_cuda.h).

This is the Python mirror of the drop-in boundary: a ``Mapper`` owns one ``srcuda_handle``
(= one GPU), holds one index block resident in HBM and runs the voting step of
``do_voting`` (reference src/core.c:3049) over chunks of reads.  There is no CPU fallback:
if the library is not built, or no B200 is visible, construction raises.
"""
import ctypes
import os
import subprocess

import numpy as np

from .layout import MAPPING_RESULT, VoteParams

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRCUDA_LIB") or os.path.join(_PKG, "libsubread_cuda.so")  # SRCUDA_LIB: A/B builds while tuning
_LIB = None

EXPORTS = [
    "srcuda_device_count", "srcuda_open", "srcuda_close", "srcuda_last_error", "srcuda_index_load",
    "srcuda_index_upload", "srcuda_index_replicate", "srcuda_index_export_ipc", "srcuda_index_import_ipc",
    "srcuda_index_describe", "srcuda_index_unload",
    "srcuda_host_alloc", "srcuda_host_free", "srcuda_vote_chunk", "srcuda_vote_chunk_device", "srcuda_vote_chunk_multi", "srcuda_index_load_all",
    "srcuda_get_timing", "srcuda_set_force_slow_path", "srcuda_measure_gather_bandwidth", "srcuda_measure_lookup_roofline",
    "srcuda_fastq_split", "srcuda_fastq_split_mt", "srcuda_reads_pack", "srcuda_reads_pack_mt",
    "srcuda_reader_open", "srcuda_reader_next", "srcuda_reader_next_full", "srcuda_reader_skip", "srcuda_reader_tell", "srcuda_reader_error", "srcuda_reader_close",
    # include/subread_output.h (row f2)
    "srcuda_out_open", "srcuda_out_add_header", "srcuda_out_add_chromosome", "srcuda_out_finish_header", "srcuda_out_add_read",
    "srcuda_out_add_reads", "srcuda_out_add_sam_lines", "srcuda_out_add_fragment", "srcuda_out_format_fragment", "srcuda_out_append", "srcuda_out_flush", "srcuda_out_close",
    "srcuda_out_error", "srcuda_sam_calc_flags", "srcuda_sam_calc_tlen",
]


class IndexDesc(ctypes.Structure):
    _fields_ = [
        ("index_gap", ctypes.c_int32), ("padding", ctypes.c_int32), ("buckets_number", ctypes.c_uint32),
        ("n_items", ctypes.c_uint64), ("start_point", ctypes.c_uint32), ("length", ctypes.c_uint32),
        ("start_base_offset", ctypes.c_uint32), ("values_bytes", ctypes.c_uint32),
        ("total_offsets", ctypes.c_int32), ("device_bytes", ctypes.c_uint64),
        ("accelerators", ctypes.c_uint32), ("_pad", ctypes.c_uint32),
    ]


class FastqBatch(ctypes.Structure):
    """srcuda_fastq_batch (include/subread_cuda.h)."""
    _fields_ = [("n_reads", ctypes.c_int64), ("consumed", ctypes.c_uint64), ("stop", ctypes.c_int32),
                ("max_len", ctypes.c_uint32), ("error_offset", ctypes.c_uint64)]


FASTQ_MORE, FASTQ_FULL, FASTQ_END = 0, 1, 2


class SamRecord(ctypes.Structure):
    """srcuda_sam_record (include/subread_output.h)."""
    _fields_ = [("name", ctypes.c_char_p), ("flags", ctypes.c_uint32), ("chro", ctypes.c_char_p), ("pos", ctypes.c_uint32),
                ("mapping_quality", ctypes.c_int32), ("cigar", ctypes.c_char_p), ("next_chro", ctypes.c_char_p),
                ("next_pos", ctypes.c_uint32), ("tlen", ctypes.c_int64), ("read_len", ctypes.c_int32), ("seq", ctypes.c_char_p),
                ("qual", ctypes.c_char_p), ("additional", ctypes.c_char_p)]


class SamConfig(ctypes.Structure):
    """srcuda_sam_config."""
    _fields_ = [("is_paired_end", ctypes.c_int32), ("minimum_pair_distance", ctypes.c_int32), ("maximum_pair_distance", ctypes.c_int32),
                ("is_first_read_reversed", ctypes.c_int32), ("is_second_read_reversed", ctypes.c_int32), ("_pad", ctypes.c_int32),
                ("read_group_id", ctypes.c_char_p)]


class SamMate(ctypes.Structure):
    """srcuda_sam_mate."""
    _fields_ = [("present", ctypes.c_int32), ("ok", ctypes.c_int32), ("chro_id", ctypes.c_int32), ("offset", ctypes.c_int32),
                ("strand", ctypes.c_int32), ("mapping_quality", ctypes.c_int32), ("mismatched_bases", ctypes.c_int32),
                ("read_len", ctypes.c_int32), ("cigar", ctypes.c_char_p), ("additional", ctypes.c_char_p),
                ("name", ctypes.c_void_p), ("text", ctypes.c_void_p), ("qual", ctypes.c_void_p)]


class SamCounters(ctypes.Structure):
    """srcuda_sam_counters."""
    _fields_ = [(n, ctypes.c_uint64) for n in ("all_mapped_reads", "all_correct_PE_reads", "not_properly_pairs_different_chro",
                                               "not_properly_different_strands", "not_properly_pairs_TLEN_wrong",
                                               "not_properly_pairs_wrong_arrangement")]


class Timing(ctypes.Structure):
    _fields_ = [
        ("ms_total", ctypes.c_float), ("ms_kernel_vote", ctypes.c_float), ("ms_kernel_overflow", ctypes.c_float),
        ("n_overflow", ctypes.c_uint64), ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64),
        ("n_launches", ctypes.c_uint32), ("n_batches", ctypes.c_uint32),
        ("ms_kernel_simple", ctypes.c_float), ("_pad", ctypes.c_uint32), ("n_simple_punted", ctypes.c_uint64),
        ("n_gather_requests", ctypes.c_uint64), ("n_simple_indel", ctypes.c_uint64),
    ]


def build_library(force: bool = False) -> str:
    """Compile libsubread_cuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_PKG, "csrc")
    if force:
        subprocess.check_call(["make", "-C", src, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", src], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, u32, u64, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int64
    L.srcuda_device_count.restype = ctypes.c_int
    L.srcuda_open.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.srcuda_close.argtypes = [vp]
    L.srcuda_close.restype = None
    L.srcuda_last_error.argtypes = [vp]
    L.srcuda_last_error.restype = ctypes.c_char_p
    L.srcuda_index_load.argtypes = [vp, ctypes.c_char_p, ctypes.c_int]
    L.srcuda_index_upload.argtypes = [vp, i32, i32, u32, u64, vp, vp, vp, u32, u32, vp, u32, vp, i32]
    L.srcuda_index_replicate.argtypes = [vp, vp]
    L.srcuda_index_export_ipc.argtypes = [vp, vp, ctypes.c_size_t]
    L.srcuda_index_import_ipc.argtypes = [vp, vp, ctypes.c_size_t]
    L.srcuda_index_describe.argtypes = [vp, ctypes.POINTER(IndexDesc)]
    L.srcuda_index_unload.argtypes = [vp]
    L.srcuda_index_unload.restype = None
    L.srcuda_host_alloc.argtypes = [ctypes.c_size_t]
    L.srcuda_host_alloc.restype = vp
    L.srcuda_host_free.argtypes = [vp]
    L.srcuda_host_free.restype = None
    L.srcuda_vote_chunk.argtypes = [vp, ctypes.POINTER(VoteParams), vp, vp, vp, i64, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    L.srcuda_vote_chunk_device.argtypes = [vp, ctypes.POINTER(VoteParams), vp, u64, vp, vp, i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    L.srcuda_vote_chunk_multi.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.POINTER(VoteParams), vp, vp, vp, i64, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    L.srcuda_index_load_all.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    L.srcuda_get_timing.argtypes = [vp, ctypes.POINTER(Timing)]
    L.srcuda_set_force_slow_path.argtypes = [vp, ctypes.c_int]
    L.srcuda_measure_gather_bandwidth.argtypes = [vp, u64, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    L.srcuda_measure_lookup_roofline.argtypes = [vp, u64, ctypes.c_int, u64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    L.srcuda_fastq_split.argtypes = [vp, u64, ctypes.c_int, ctypes.c_int, ctypes.c_int, i64, u64, i64, vp, vp, vp, vp, vp, vp,
                                     ctypes.POINTER(FastqBatch)]
    L.srcuda_fastq_split_mt.argtypes = [vp, u64, ctypes.c_int, ctypes.c_int, ctypes.c_int, i64, u64, i64, vp, vp, vp, vp, vp, vp,
                                        ctypes.c_int, ctypes.POINTER(FastqBatch)]
    L.srcuda_reads_pack.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp, u64, vp, vp, ctypes.POINTER(u64)]
    L.srcuda_reads_pack_mt.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp, u64, vp, vp, ctypes.c_int, ctypes.POINTER(u64)]
    L.srcuda_reader_open.argtypes = [ctypes.c_char_p, ctypes.c_char_p, u64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
    L.srcuda_reader_next.argtypes = [vp, i64, vp, u64, vp, vp, ctypes.POINTER(i64), ctypes.POINTER(u64)]
    L.srcuda_reader_next_full.argtypes = [vp, i64, vp, u64, vp, vp, vp, u64, vp, vp, vp, u64, vp, vp, ctypes.POINTER(i64), ctypes.POINTER(u64)]
    L.srcuda_reader_skip.argtypes = [vp, i64, ctypes.POINTER(i64)]
    L.srcuda_reader_tell.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.srcuda_reader_error.argtypes = [vp]
    L.srcuda_reader_error.restype = ctypes.c_char_p
    L.srcuda_reader_close.argtypes = [vp]
    L.srcuda_reader_close.restype = None
    cp = ctypes.c_char_p
    L.srcuda_out_open.argtypes = [cp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
    L.srcuda_out_add_header.argtypes = [vp, cp, ctypes.c_int]
    L.srcuda_out_add_chromosome.argtypes = [vp, cp, u32, ctypes.c_int]
    L.srcuda_out_finish_header.argtypes = [vp]
    L.srcuda_out_add_read.argtypes = [vp, ctypes.POINTER(SamRecord), ctypes.c_int]
    L.srcuda_out_add_reads.argtypes = [vp, ctypes.POINTER(SamRecord), i64, ctypes.c_int]
    L.srcuda_out_add_sam_lines.argtypes = [vp, cp, ctypes.c_size_t, ctypes.c_int]
    L.srcuda_out_add_fragment.argtypes = [vp, ctypes.POINTER(SamConfig), ctypes.POINTER(SamMate), ctypes.POINTER(SamMate),
                                          ctypes.c_int, ctypes.c_int, ctypes.POINTER(SamCounters)]
    L.srcuda_out_format_fragment.argtypes = [vp, ctypes.POINTER(SamConfig), ctypes.POINTER(SamMate), ctypes.POINTER(SamMate), ctypes.c_int,
                                             ctypes.c_int, ctypes.POINTER(SamCounters), vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    L.srcuda_out_append.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_int]
    L.srcuda_out_flush.argtypes = [vp]
    L.srcuda_out_close.argtypes = [vp, ctypes.POINTER(u64)]
    L.srcuda_out_error.argtypes = [vp]
    L.srcuda_out_error.restype = cp
    L.srcuda_sam_calc_flags.argtypes = [ctypes.POINTER(SamConfig), ctypes.POINTER(SamMate), ctypes.POINTER(SamMate), ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.POINTER(SamCounters)]
    L.srcuda_sam_calc_tlen.argtypes = [i32, cp, ctypes.c_int, i32, cp, ctypes.c_int]
    _LIB = L
    return L


class SubreadCudaError(RuntimeError):
    pass


class PinnedArray:
    """numpy view over cudaHostAlloc'ed memory (srcuda_host_alloc)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(np.atleast_1d(shape))
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        self._lib = load_library()
        self.ptr = self._lib.srcuda_host_alloc(max(n, 1))
        if not self.ptr:
            raise SubreadCudaError("srcuda_host_alloc failed")
        buf = (ctypes.c_uint8 * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.srcuda_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Mapper:
    """One GPU, one resident index block."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        rc = self.lib.srcuda_open(device, ctypes.byref(self.h))
        if rc:
            raise SubreadCudaError(f"srcuda_open({device}) -> {rc}: {self.lib.srcuda_last_error(None).decode()}")
        self.device = device

    def _check(self, rc, what):
        if rc:
            raise SubreadCudaError(f"{what} -> {rc}: {self.lib.srcuda_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.lib.srcuda_close(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- index -------------------------------------------------------------------------------
    def index_load(self, prefix: str, block: int = 0):
        self._check(self.lib.srcuda_index_load(self.h, str(prefix).encode(), block), "srcuda_index_load")

    def index_upload(self, index_gap, padding, bucket_start, keys, vals, start_point, length, values, read_offsets):
        bucket_start = np.ascontiguousarray(bucket_start, dtype=np.uint32)
        keys = np.ascontiguousarray(keys, dtype=np.int16)
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        values = np.ascontiguousarray(values, dtype=np.uint8)
        read_offsets = np.ascontiguousarray(read_offsets, dtype=np.uint32)
        self._check(self.lib.srcuda_index_upload(
            self.h, int(index_gap), int(padding), bucket_start.shape[0] - 1, keys.shape[0], _ptr(bucket_start), _ptr(keys),
            _ptr(vals), int(start_point), int(length), _ptr(values), values.shape[0], _ptr(read_offsets),
            read_offsets.shape[0]), "srcuda_index_upload")

    def index_replicate_from(self, other: "Mapper"):
        self._check(self.lib.srcuda_index_replicate(self.h, other.h), "srcuda_index_replicate")

    IPC_BLOB_BYTES = 512

    def index_export_ipc(self) -> bytes:
        buf = ctypes.create_string_buffer(self.IPC_BLOB_BYTES)
        self._check(self.lib.srcuda_index_export_ipc(self.h, buf, self.IPC_BLOB_BYTES), "srcuda_index_export_ipc")
        return buf.raw

    def index_import_ipc(self, blob: bytes):
        buf = ctypes.create_string_buffer(blob, self.IPC_BLOB_BYTES)
        self._check(self.lib.srcuda_index_import_ipc(self.h, buf, self.IPC_BLOB_BYTES), "srcuda_index_import_ipc")

    def describe(self) -> IndexDesc:
        d = IndexDesc()
        self._check(self.lib.srcuda_index_describe(self.h, ctypes.byref(d)), "srcuda_index_describe")
        return d

    # -- voting ------------------------------------------------------------------------------
    def vote_chunk(self, chunk, params: VoteParams, out_align=None, want_verify=False, junc_state=None):
        """Host-buffer entry point (copies inside).  Returns (align[n_frags, mates, multi_best], verify|None).
        subjunc (params.do_breakpoint_detection): pass junc_state = {} -- it receives 'junc' and 'big_margin',
        which are read-modify-write across index blocks together with out_align."""
        from .layout import SUBJUNC_RESULT
        mb = params.multi_best_reads
        n = chunk.n_frags * chunk.mates * mb
        is_zero = out_align is None
        junc = bm = None
        if params.do_breakpoint_detection:
            if junc_state is None:
                junc_state = {}
            if "junc" not in junc_state:
                junc_state["junc"] = np.zeros(n, dtype=SUBJUNC_RESULT)
                junc_state["big_margin"] = np.zeros(chunk.n_frags * chunk.mates * 27, dtype=np.uint16)
            junc, bm = junc_state["junc"], junc_state["big_margin"]
        if out_align is None:
            out_align = np.zeros(n, dtype=MAPPING_RESULT)
        out_align = out_align.reshape(-1)
        assert out_align.flags["C_CONTIGUOUS"] and out_align.shape[0] == n and out_align.dtype == MAPPING_RESULT
        verify = np.zeros(n, dtype=np.uint16) if want_verify else None
        text = np.ascontiguousarray(chunk.text, dtype=np.uint8)
        off = np.ascontiguousarray(chunk.off, dtype=np.uint64)
        ln = np.ascontiguousarray(chunk.len, dtype=np.uint16)
        if text.shape[0] == 0:
            text = np.zeros(1, np.uint8)
        rc = self.lib.srcuda_vote_chunk(self.h, ctypes.byref(params), _ptr(text), _ptr(off), _ptr(ln), chunk.n_frags,
                                        int(chunk.is_paired), int(is_zero), _ptr(out_align), _ptr(junc), _ptr(bm), _ptr(verify))
        self._check(rc, "srcuda_vote_chunk")
        return out_align.reshape(chunk.n_frags, chunk.mates, mb), verify

    def vote_chunk_raw(self, params, text_ptr, off_ptr, len_ptr, n_frags, is_paired, out_is_zero, out_ptr, verify_ptr=None):
        """Pointer-level call (host pointers, e.g. pinned buffers) used by bench.py's e2e timing."""
        rc = self.lib.srcuda_vote_chunk(self.h, ctypes.byref(params), text_ptr, off_ptr, len_ptr, n_frags, int(is_paired),
                                        int(out_is_zero), out_ptr, None, None, verify_ptr)
        self._check(rc, "srcuda_vote_chunk")

    def vote_chunk_device(self, params, d_text, d_text_bytes, d_off, d_len, n_frags, is_paired, max_read_len, d_out, d_verify=None,
                          out_is_zero=False, d_junc=None, d_big_margin=None):
        """All pointers are device addresses on this mapper's GPU (e.g. torch tensors' data_ptr()); the caller makes
        sure the buffers are complete before the call (the library launches on its own streams)."""
        rc = self.lib.srcuda_vote_chunk_device(self.h, ctypes.byref(params), d_text, int(d_text_bytes), d_off, d_len, n_frags, int(is_paired),
                                               int(max_read_len), int(out_is_zero), d_out, d_junc, d_big_margin, d_verify)
        self._check(rc, "srcuda_vote_chunk_device")

    def timing(self) -> Timing:
        t = Timing()
        self.lib.srcuda_get_timing(self.h, ctypes.byref(t))
        return t

    def set_force_slow_path(self, on: bool):
        self.lib.srcuda_set_force_slow_path(self.h, int(bool(on)))

    def lookup_roofline(self, small_bytes: int, n_small: int, big_bytes: int, n_big: int, iters: int = 3) -> float:
        """Units of work per second the memory system allows for n_small L2-resident probes + n_big DRAM gathers per unit."""
        out = ctypes.c_double()
        self._check(self.lib.srcuda_measure_lookup_roofline(self.h, int(small_bytes), int(n_small), int(big_bytes), int(n_big), iters, ctypes.byref(out)),
                    "srcuda_measure_lookup_roofline")
        return out.value

    def gather_bandwidth(self, table_bytes: int, iters: int = 3) -> float:
        out = ctypes.c_double()
        self._check(self.lib.srcuda_measure_gather_bandwidth(self.h, int(table_bytes), iters, ctypes.byref(out)),
                    "srcuda_measure_gather_bandwidth")
        return out.value


class MultiMapper:
    """Several GPUs behind one chunk call (srcuda_vote_chunk_multi): one handle per entry of ``devices`` (a device may be
    listed twice -- two handles with their own streams and their own copy of the index -- which is how the driver is
    exercised on a one-GPU box), the index loaded once and replicated to the others over NVLink, every chunk cut into
    contiguous fragment slices written into one host bigtable (run_maybe_threads, core.c:3404-3456)."""

    def __init__(self, devices):
        self.mappers = [Mapper(d) for d in devices]
        self.lib = self.mappers[0].lib
        self._arr = (ctypes.c_void_p * len(self.mappers))(*[m.h for m in self.mappers])

    def close(self):
        for m in self.mappers:
            m.close()
        self.mappers = []

    def index_load(self, prefix: str, block: int = 0):
        rc = self.lib.srcuda_index_load_all(self._arr, len(self.mappers), str(prefix).encode(), block)
        self.mappers[0]._check(rc, "srcuda_index_load_all")

    def index_replicate(self):
        """The block resident on the first handle -> every other handle (peer copies)."""
        for m in self.mappers[1:]:
            m.index_replicate_from(self.mappers[0])

    def vote_chunk(self, chunk, params: VoteParams, out_align=None, want_verify=False, junc_state=None):
        """Same contract as Mapper.vote_chunk, the chunk spread over all handles."""
        from .layout import SUBJUNC_RESULT
        mb = params.multi_best_reads
        n = chunk.n_frags * chunk.mates * mb
        is_zero = out_align is None
        junc = bm = None
        if params.do_breakpoint_detection:
            if junc_state is None:
                junc_state = {}
            if "junc" not in junc_state:
                junc_state["junc"] = np.zeros(n, dtype=SUBJUNC_RESULT)
                junc_state["big_margin"] = np.zeros(chunk.n_frags * chunk.mates * 27, dtype=np.uint16)
            junc, bm = junc_state["junc"], junc_state["big_margin"]
        if out_align is None:
            out_align = np.zeros(n, dtype=MAPPING_RESULT)
        out_align = out_align.reshape(-1)
        assert out_align.flags["C_CONTIGUOUS"] and out_align.shape[0] == n and out_align.dtype == MAPPING_RESULT
        verify = np.zeros(n, dtype=np.uint16) if want_verify else None
        text = np.ascontiguousarray(chunk.text, dtype=np.uint8)
        off = np.ascontiguousarray(chunk.off, dtype=np.uint64)
        ln = np.ascontiguousarray(chunk.len, dtype=np.uint16)
        if text.shape[0] == 0:
            text = np.zeros(1, np.uint8)
        rc = self.lib.srcuda_vote_chunk_multi(self._arr, len(self.mappers), ctypes.byref(params), _ptr(text), _ptr(off), _ptr(ln), chunk.n_frags,
                                              int(chunk.is_paired), int(is_zero), _ptr(out_align), _ptr(junc), _ptr(bm), _ptr(verify))
        self.mappers[0]._check(rc, "srcuda_vote_chunk_multi")
        return out_align.reshape(chunk.n_frags, chunk.mates, mb), verify

    def timings(self):
        return [m.timing() for m in self.mappers]


def fastq_split(buf: np.ndarray, at_eof=True, trim_5=0, trim_3=0, max_reads=None, base_offset=0, want_names=False, threads=1,
                partial=False):
    """srcuda_fastq_split (threads > 1: srcuda_fastq_split_mt) over a uint8 array of plain FASTQ text.  Returns (off, len, batch, extra): offsets into
    ``buf`` (+ base_offset) and lengths of the records the buffer holds completely, the FastqBatch verdict, and --
    with want_names -- a dict of name_off/name_len/qual_off/qual_len.  Raises SubreadCudaError on a format issue
    (the records in front of it are lost to the caller, as the reference stops there too)."""
    L = load_library()
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    cap = int(max_reads) if max_reads is not None else buf.shape[0] // 6 + 1     # a record has at least "@\nA\n+\nI" bytes
    off = np.empty(cap, dtype=np.uint64)
    ln = np.empty(cap, dtype=np.uint16)
    extra = {}
    if want_names:
        extra = dict(name_off=np.empty(cap, np.uint64), name_len=np.empty(cap, np.uint16),
                     qual_off=np.empty(cap, np.uint64), qual_len=np.empty(cap, np.uint16))
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    b = FastqBatch()
    rc = L.srcuda_fastq_split_mt(vp(buf), buf.shape[0], int(bool(at_eof)), int(trim_5), int(trim_3), cap, int(base_offset), 1, vp(off), vp(ln),
                                 vp(extra["name_off"]) if want_names else None, vp(extra["name_len"]) if want_names else None,
                                 vp(extra["qual_off"]) if want_names else None, vp(extra["qual_len"]) if want_names else None,
                                 int(threads), ctypes.byref(b))
    n = int(b.n_reads)
    if partial:      # the records in front of a format issue too, with the return code
        return off[:n], ln[:n], b, {k: v[:n] for k, v in extra.items()}, rc
    if rc:
        raise SubreadCudaError(f"srcuda_fastq_split: format issue at byte {b.error_offset} (rc={rc}, {n} records before it)")
    return off[:n], ln[:n], b, {k: v[:n] for k, v in extra.items()}


def reads_pack(buf1, off1, len1, buf2=None, off2=None, len2=None, threads=1):
    """srcuda_reads_pack: dense (text, off[n+1], len) of the indexed reads, mates interleaved per fragment."""
    L = load_library()
    vp = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    n_frags = int(off1.shape[0])
    if buf2 is not None and off2.shape[0] != n_frags:
        raise SubreadCudaError("two input files have different amounts of reads")      # core.c:1177-1182
    mates = 2 if buf2 is not None else 1
    total = int(len1.astype(np.int64).sum()) + (int(len2.astype(np.int64).sum()) if buf2 is not None else 0)
    text = np.empty(total, dtype=np.uint8)
    off = np.zeros(n_frags * mates + 1, dtype=np.uint64)
    ln = np.empty(n_frags * mates, dtype=np.uint16)
    written = ctypes.c_uint64()
    rc = L.srcuda_reads_pack_mt(vp(buf1), vp(off1), vp(len1), vp(buf2), vp(off2), vp(len2), n_frags, vp(text), total, vp(off), vp(ln),
                                int(threads), ctypes.byref(written))
    if rc or written.value != total:
        raise SubreadCudaError(f"srcuda_reads_pack failed (rc={rc})")
    off[-1] = total
    return text, off, ln


class ChunkReader:
    """srcuda_reader_*: batches of fragments from one or two FASTQ files (plain or gzip) in the chunk layout."""

    def __init__(self, path1, path2=None, window_bytes=0, threads=1, trim_5=0, trim_3=0):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        self.mates = 2 if path2 else 1
        rc = self.lib.srcuda_reader_open(os.fsencode(path1), os.fsencode(path2) if path2 else None, int(window_bytes), int(threads),
                                         int(trim_5), int(trim_3), ctypes.byref(self.h))
        if rc:
            msg = self.lib.srcuda_reader_error(self.h).decode() if self.h else "bad argument"
            self.close()
            raise SubreadCudaError(f"srcuda_reader_open: {msg} (rc={rc})")

    def next(self, max_frags, max_read_len=1210):
        """(text, off[n+1], len, n_frags) of the next batch; n_frags == 0 at the end of the input."""
        n_reads = int(max_frags) * self.mates
        text = np.empty(max(1, n_reads * int(max_read_len)), dtype=np.uint8)
        off = np.zeros(n_reads + 1, dtype=np.uint64)
        ln = np.zeros(max(1, n_reads), dtype=np.uint16)
        n, used = ctypes.c_int64(), ctypes.c_uint64()
        rc = self.lib.srcuda_reader_next(self.h, int(max_frags), _ptr(text), text.shape[0], _ptr(off), _ptr(ln), ctypes.byref(n), ctypes.byref(used))
        if rc:
            raise SubreadCudaError(f"srcuda_reader_next: {self.lib.srcuda_reader_error(self.h).decode()} (rc={rc})")
        k = n.value * self.mates
        off[k] = used.value
        return text[:used.value], off[:k + 1], ln[:k], int(n.value)

    def skip(self, n_frags):
        """Passes over the next n_frags fragments; returns how many there were."""
        n = ctypes.c_int64()
        rc = self.lib.srcuda_reader_skip(self.h, int(n_frags), ctypes.byref(n))
        if rc:
            raise SubreadCudaError(f"srcuda_reader_skip: {self.lib.srcuda_reader_error(self.h).decode()} (rc={rc})")
        return int(n.value)

    def tell(self):
        """File offsets (inflated bytes) behind the last fragment handed out, one per file."""
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        self.lib.srcuda_reader_tell(self.h, ctypes.byref(a), ctypes.byref(b))
        return (a.value, b.value) if self.mates == 2 else (a.value,)

    def close(self):
        if self.h:
            self.lib.srcuda_reader_close(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


OUT_SAM, OUT_BAM = 0, 1


class OutputWriter:
    """ctypes mirror of the srcuda_out_* writer (include/subread_output.h): SAM text or BAM, host side only."""

    def __init__(self, path, fmt=OUT_BAM, threads=1, no_compression=False):
        self.L = load_library()
        self.h = ctypes.c_void_p()
        rc = self.L.srcuda_out_open(os.fsencode(path) if path is not None else None, fmt, threads, int(no_compression), ctypes.byref(self.h))
        if rc:
            raise SubreadCudaError(f"srcuda_out_open: {self.L.srcuda_out_error(None).decode()} (code {rc})")

    def _check(self, rc, what):
        if rc:
            raise SubreadCudaError(f"{what}: {self.L.srcuda_out_error(self.h).decode()} (code {rc})")

    def add_header(self, line, add_chro=False):
        self._check(self.L.srcuda_out_add_header(self.h, line.encode() if isinstance(line, str) else line, int(add_chro)), "add_header")

    def add_chromosome(self, name, length, add_header=True):
        self._check(self.L.srcuda_out_add_chromosome(self.h, name.encode() if isinstance(name, str) else name, length, int(add_header)), "add_chromosome")

    def finish_header(self):
        self._check(self.L.srcuda_out_finish_header(self.h), "finish_header")

    def add_read(self, rec: SamRecord, committable=True):
        self._check(self.L.srcuda_out_add_read(self.h, ctypes.byref(rec), int(committable)), "add_read")

    def add_reads(self, recs, records_per_fragment=1):
        arr = (SamRecord * len(recs))(*recs)
        self._check(self.L.srcuda_out_add_reads(self.h, arr, len(recs), records_per_fragment), "add_reads")

    def add_sam_lines(self, text: bytes, records_per_fragment=1):
        self._check(self.L.srcuda_out_add_sam_lines(self.h, text, len(text), records_per_fragment), "add_sam_lines")

    def add_fragment(self, cfg: SamConfig, m1: SamMate, m2, all_locations=1, current_location=0, counters=None):
        self._check(self.L.srcuda_out_add_fragment(self.h, ctypes.byref(cfg), ctypes.byref(m1), ctypes.byref(m2) if m2 is not None else None,
                                                   all_locations, current_location, ctypes.byref(counters) if counters is not None else None),
                    "add_fragment")

    def flush(self):
        self._check(self.L.srcuda_out_flush(self.h), "flush")

    def close(self):
        if self.h:
            n = ctypes.c_uint64()
            h, self.h = self.h, None
            rc = self.L.srcuda_out_close(h, ctypes.byref(n))
            if rc:
                raise SubreadCudaError(f"srcuda_out_close: {self.L.srcuda_out_error(None).decode()} (code {rc})")
            return n.value
        return 0
