"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It restates the reference's voting step on the CPU (oracle/sr_oracle.c).
"""
import ctypes
import os
import subprocess

import numpy as np

from subread_b200.layout import MAPPING_RESULT, VoteParams

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrIndex(ctypes.Structure):
    _fields_ = [
        ("index_gap", ctypes.c_int32),
        ("padding", ctypes.c_int32),
        ("buckets_number", ctypes.c_uint32),
        ("n_items", ctypes.c_uint64),
        ("bucket_start", ctypes.POINTER(ctypes.c_uint32)),
        ("keys", ctypes.POINTER(ctypes.c_int16)),
        ("vals", ctypes.POINTER(ctypes.c_uint32)),
        ("start_point", ctypes.c_uint32),
        ("length", ctypes.c_uint32),
        ("start_base_offset", ctypes.c_uint32),
        ("values_bytes", ctypes.c_uint32),
        ("values", ctypes.POINTER(ctypes.c_uint8)),
        ("total_offsets", ctypes.c_int32),
        ("read_offsets", ctypes.POINTER(ctypes.c_uint32)),
    ]


class OrStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in
                ("lookups", "probes", "hits", "read_bytes", "verify_bytes", "result_bytes")]

    def algorithmic_bytes(self) -> int:
        """L + sum(8 + 2*probes + 4*hits) + K*(ceil(L/4)+1) + 3*68."""
        return int(self.read_bytes + 8 * self.lookups + 2 * self.probes + 4 * self.hits
                   + self.verify_bytes + self.result_bytes)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("sr_oracle.c", "sr_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.or_index_load.argtypes = [ctypes.POINTER(OrIndex), ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
        L.or_index_load.restype = ctypes.c_int
        L.or_index_free.argtypes = [ctypes.POINTER(OrIndex)]
        L.or_genekey2int.argtypes = [ctypes.c_char_p]
        L.or_genekey2int.restype = ctypes.c_uint32
        L.or_reverse_read.argtypes = [ctypes.c_char_p, ctypes.c_int]
        L.or_match_chro.argtypes = [ctypes.c_char_p, ctypes.POINTER(OrIndex), ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
        L.or_match_chro.restype = ctypes.c_int
        L.or_subread_schedule.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        L.or_subread_schedule.restype = ctypes.c_int
        common = [ctypes.POINTER(OrIndex), ctypes.POINTER(VoteParams), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                  ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                  ctypes.POINTER(OrStats)]
        L.or_vote_chunk.argtypes = common
        L.or_vote_chunk.restype = ctypes.c_int
        L.or_vote_chunk_mt.argtypes = common + [ctypes.c_int]
        L.or_vote_chunk_mt.restype = ctypes.c_int
        _LIB = L
    return _LIB


class OracleIndex:
    """One index block (``<prefix>.NN.b.tab/.array`` + ``<prefix>.reads``) held on the host."""

    def __init__(self, prefix: str, block: int = 0):
        self.c = OrIndex()
        rc = lib().or_index_load(ctypes.byref(self.c), f"{prefix}.{block:02d}.b.tab".encode(),
                                 f"{prefix}.{block:02d}.b.array".encode(), f"{prefix}.reads".encode())
        if rc:
            raise RuntimeError(f"or_index_load({prefix}) failed with {rc}")

    @classmethod
    def from_arrays(cls, a):
        """The same block from arrays already in host memory (the dict subread_b200.index_build returns): no file
        round trip, for indexes of human size.  The arrays stay owned by Python."""
        self = cls.__new__(cls)
        keep = dict(bucket_start=np.ascontiguousarray(a["bucket_start"], dtype=np.uint32),
                    keys=np.ascontiguousarray(a["keys"], dtype=np.int16), vals=np.ascontiguousarray(a["vals"], dtype=np.uint32),
                    values=np.concatenate((np.asarray(a["values"], dtype=np.uint8), np.zeros(16, np.uint8))),  # slack like or_index_load
                    read_offsets=np.ascontiguousarray(a["read_offsets"], dtype=np.uint32))
        c = OrIndex()
        c.index_gap, c.padding = int(a["index_gap"]), int(a["padding"])
        c.buckets_number, c.n_items = keep["bucket_start"].shape[0] - 1, keep["keys"].shape[0]
        c.bucket_start = keep["bucket_start"].ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
        c.keys = keep["keys"].ctypes.data_as(ctypes.POINTER(ctypes.c_int16))
        c.vals = keep["vals"].ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
        c.start_point, c.length = int(a["start_point"]), int(a["length"])
        c.start_base_offset = c.start_point - c.start_point % 4       # gvindex_load, gene-value-index.c:208
        c.values_bytes = keep["values"].shape[0] - 16
        c.values = keep["values"].ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        c.total_offsets = keep["read_offsets"].shape[0]
        c.read_offsets = keep["read_offsets"].ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
        self.c, self._borrowed = c, keep
        return self

    def close(self):
        if self.c is not None:
            if getattr(self, "_borrowed", None) is None:
                lib().or_index_free(ctypes.byref(self.c))
            self._borrowed = None
            self.c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # numpy views of the loaded arrays (used to feed srcuda_index_upload in tests)
    def arrays(self):
        c = self.c
        nb, ni = c.buckets_number, c.n_items
        return dict(
            _owner=self,  # the views below point into memory this object frees in close()
            index_gap=c.index_gap, padding=c.padding, buckets_number=nb, n_items=ni,
            bucket_start=np.ctypeslib.as_array(c.bucket_start, shape=(nb + 1,)),
            keys=np.ctypeslib.as_array(c.keys, shape=(max(ni, 1),))[:ni],
            vals=np.ctypeslib.as_array(c.vals, shape=(max(ni, 1),))[:ni],
            start_point=c.start_point, length=c.length,
            values=np.ctypeslib.as_array(c.values, shape=(c.values_bytes,)),
            read_offsets=np.ctypeslib.as_array(c.read_offsets, shape=(c.total_offsets,)),
        )


def vote_chunk(index: OracleIndex, params: VoteParams, chunk, out_align=None, want_verify=False, threads=1, junc_state=None):
    """Run the oracle over a ReadChunk.  Returns (align[n_frags, mates, multi_best], verify|None, stats).
    With params.do_breakpoint_detection (subjunc) pass junc_state = {} (filled with 'junc' and 'big_margin'
    arrays, read-modify-write across index blocks like out_align)."""
    from subread_b200.layout import SUBJUNC_RESULT
    mb = params.multi_best_reads
    n = chunk.n_frags * chunk.mates * mb
    junc_ptr = bm_ptr = None
    if params.do_breakpoint_detection:
        if junc_state is None:
            junc_state = {}
        if "junc" not in junc_state:
            junc_state["junc"] = np.zeros(n, dtype=SUBJUNC_RESULT)
            junc_state["big_margin"] = np.zeros(chunk.n_frags * chunk.mates * 27, dtype=np.uint16)
        junc_ptr = junc_state["junc"].ctypes.data
        bm_ptr = junc_state["big_margin"].ctypes.data
    if out_align is None:
        out_align = np.zeros(n, dtype=MAPPING_RESULT)
    out_align = np.ascontiguousarray(out_align).reshape(-1)
    assert out_align.shape[0] == n
    verify = np.zeros(n, dtype=np.uint16) if want_verify else None
    stats = OrStats()
    text = np.ascontiguousarray(chunk.text)
    off = np.ascontiguousarray(chunk.off, dtype=np.uint64)
    ln = np.ascontiguousarray(chunk.len, dtype=np.uint16)
    args = [ctypes.byref(index.c), ctypes.byref(params), text.ctypes.data, off.ctypes.data, ln.ctypes.data,
            chunk.n_frags, int(chunk.is_paired), out_align.ctypes.data, junc_ptr, bm_ptr,
            verify.ctypes.data if want_verify else None, ctypes.byref(stats)]
    rc = lib().or_vote_chunk_mt(*args, threads) if threads > 1 else lib().or_vote_chunk(*args)
    if rc:
        raise RuntimeError(f"or_vote_chunk failed with {rc}")
    return out_align.reshape(chunk.n_frags, chunk.mates, mb), verify, stats


def read_bigtable_dump(path):
    """Parse a SRDUMP_BIGTABLE file written by oracle/ref_hooks.c; returns a list of chunks."""
    from subread_b200.layout import SUBJUNC_RESULT
    out = []
    with open(path, "rb") as fh:
        while True:
            hdr = fh.read(24)
            if len(hdr) < 24:
                break
            magic, n_frags, mates, mb, has_junc, bm = np.frombuffer(hdr, dtype="<u4")
            assert magic == 0x54425253
            n = int(n_frags) * int(mates) * int(mb)
            align = np.frombuffer(fh.read(n * 68), dtype=MAPPING_RESULT).reshape(n_frags, mates, mb)
            rec = dict(align=align, junc=None, big_margin=None)
            if has_junc:
                rec["junc"] = np.frombuffer(fh.read(n * 16), dtype=SUBJUNC_RESULT).reshape(n_frags, mates, mb)
                rec["big_margin"] = np.frombuffer(fh.read(int(n_frags) * int(mates) * 54), dtype="<u2").reshape(n_frags, mates, 27)
            out.append(rec)
    return out
