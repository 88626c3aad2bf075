#!/usr/bin/env python
"""tools/fastq_e2e.py -- the input side in front of the voting step (row f3), on the GPU box's host cores:
srcuda_fastq_split / _mt rates on a FASTQ buffer of the bench shape (150 bp SE, 100 Mbp genome), and the end-to-end rate
from FASTQ bytes to the bigtable three ways: (a) indexed in place, the FASTQ bytes themselves uploaded; (b) indexed, packed
(srcuda_reads_pack) and then uploaded; (c) the dense arrays handed over ready-made (what bench.py's e2e measures).  The
three bigtables must be identical.  One JSON line.

    python tools/fastq_e2e.py --reads 4000000 --steps 3
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def fastq_record_width(read_len):
    return 12 + read_len + 3 + read_len + 1


def fill_fastq(rec, text):
    """rec: uint8 [R, fastq_record_width(RL)] to fill; text: uint8 [R, RL] read characters."""
    R, RL = text.shape
    rec[:, 0], rec[:, 1] = ord("@"), ord("r")
    idx = np.arange(R, dtype=np.int64)
    for k in range(9):
        rec[:, 2 + k] = 48 + (idx // 10 ** (8 - k)) % 10
    rec[:, 11] = ord("\n")
    rec[:, 12:12 + RL] = text
    rec[:, 12 + RL:15 + RL] = np.frombuffer(b"\n+\n", np.uint8)
    rec[:, 15 + RL:15 + 2 * RL] = ord("H")
    rec[:, 15 + 2 * RL] = ord("\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-mbp", type=float, default=100.0)
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--window", type=int, default=1 << 28, help="bytes of FASTQ per pipeline stage")
    args = ap.parse_args()

    import torch
    from subread_b200 import VoteParams
    from subread_b200.binding import FastqBatch, Mapper, PinnedArray, load_library
    from subread_b200.index_build import build_index_arrays_device

    os.environ.setdefault("SRCUDA_BATCH_FRAGS", "262144")     # several batches per call: copies and kernels overlap inside a call
    L = load_library()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    CL = 500000
    n_chr = max(1, int(round(args.genome_mbp * 1e6 / CL)))
    gen = torch.Generator(device=dev)
    gen.manual_seed(20261019)
    codes = torch.randint(0, 4, (n_chr * CL,), generator=gen, device=dev, dtype=torch.uint8)
    a = build_index_arrays_device([f"chr{c + 1}" for c in range(n_chr)], codes, [CL] * n_chr, gapped=True)
    m = Mapper(0)
    m.index_upload(a["index_gap"], a["padding"], a["bucket_start"], a["keys"], a["vals"], a["start_point"], a["length"], a["values"],
                   a["read_offsets"])
    R, RL = args.reads, args.read_len
    letters = torch.tensor(list(b"AGCT"), dtype=torch.uint8, device=dev)
    lin_first = torch.from_numpy(a["contig_starts"]).to(dev)
    c = torch.randint(0, n_chr, (R,), generator=gen, device=dev)
    s = torch.randint(0, CL - RL, (R,), generator=gen, device=dev)
    ar = torch.arange(RL, device=dev)
    text = letters[codes[(c * CL + s)[:, None] + ar[None, :]].long()].cpu().numpy()          # forward-strand reads
    truth = (lin_first[c] + s).cpu().numpy()
    del codes
    torch.cuda.empty_cache()

    # FASTQ records of fixed width: "@r%09d\n" + read + "\n+\n" + qualities + "\n"
    W = fastq_record_width(RL)
    fq = PinnedArray(R * W, np.uint8)
    fill_fastq(fq.array.reshape(R, W), text)
    n_bytes = R * W

    off = PinnedArray(R, np.uint64)
    ln = PinnedArray(R, np.uint16)
    p_text = PinnedArray(R * RL, np.uint8)
    p_off = PinnedArray(R, np.uint64)
    p_len = PinnedArray(R, np.uint16)
    mb = 3
    out_a = PinnedArray(R * mb * 68, np.uint8)
    out_b = PinnedArray(R * mb * 68, np.uint8)
    params = VoteParams.defaults(False)
    b = FastqBatch()
    vp = ctypes.c_void_p

    def split(T):
        rc = L.srcuda_fastq_split_mt(vp(fq.ptr), n_bytes, 1, 0, 0, R, 0, 1, vp(off.ptr), vp(ln.ptr), None, None, None, None, T, ctypes.byref(b))
        assert rc == 0 and b.n_reads == R and b.consumed == n_bytes, (rc, b.n_reads, b.consumed)

    def pack():
        w = ctypes.c_uint64()
        rc = L.srcuda_reads_pack_mt(vp(fq.ptr), vp(off.ptr), vp(ln.ptr), None, None, None, R, vp(p_text.ptr), R * RL, vp(p_off.ptr), vp(p_len.ptr),
                                    args.threads, ctypes.byref(w))
        assert rc == 0 and w.value == R * RL

    def timed(fn, n):
        best = 1e30
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return best

    res = {"tool": "fastq_e2e", "reads": R, "read_len": RL, "fastq_bytes": n_bytes, "host_threads": args.threads, "steps": args.steps}
    for T in sorted({1, 2, 4, 8, args.threads}):
        if T <= args.threads:
            res[f"split_T{T}_reads_per_s"] = R / timed(lambda: split(T), args.steps + args.warmup)
    split(args.threads)
    res["pack_reads_per_s"] = R / timed(pack, args.steps + args.warmup)

    def in_place():
        split(args.threads)
        m.vote_chunk_raw(params, fq.ptr, off.ptr, ln.ptr, R, False, True, out_a.ptr)

    def packed():
        split(args.threads)
        pack()
        m.vote_chunk_raw(params, p_text.ptr, p_off.ptr, p_len.ptr, R, False, True, out_b.ptr)

    def ready_made():
        m.vote_chunk_raw(params, p_text.ptr, p_off.ptr, p_len.ptr, R, False, True, out_b.ptr)

    # The same two ways as a streaming caller runs them: the file arrives in windows of args.window bytes, a host thread
    # indexes (and packs) window k+1 while window k is on the GPU.  ctypes calls release the interpreter lock.
    import queue
    import threading

    def pipelined(do_pack, out):
        q = queue.Queue(maxsize=2)

        def producer():
            pos, r0, t0 = 0, 0, 0
            bb = FastqBatch()
            while pos < n_bytes:
                end = min(n_bytes, pos + args.window)
                rc = L.srcuda_fastq_split_mt(vp(fq.ptr + pos), end - pos, int(end == n_bytes), 0, 0, R - r0, pos, 1, vp(off.ptr + 8 * r0),
                                             vp(ln.ptr + 2 * r0), None, None, None, None, args.threads, ctypes.byref(bb))
                assert rc == 0 and bb.n_reads > 0, (rc, bb.n_reads)
                n = int(bb.n_reads)
                if do_pack:
                    w = ctypes.c_uint64()
                    rc = L.srcuda_reads_pack_mt(vp(fq.ptr), vp(off.ptr + 8 * r0), vp(ln.ptr + 2 * r0), None, None, None, n, vp(p_text.ptr + t0),
                                                R * RL - t0, vp(p_off.ptr + 8 * r0), vp(p_len.ptr + 2 * r0), args.threads, ctypes.byref(w))
                    assert rc == 0
                    q.put((r0, n, t0))
                    t0 += int(w.value)
                else:
                    q.put((r0, n, 0))
                pos += int(bb.consumed)
                r0 += n
            q.put(None)

        th = threading.Thread(target=producer)
        th.start()
        done = 0
        while True:
            item = q.get()
            if item is None:
                break
            r0, n, t0 = item
            if do_pack:
                m.vote_chunk_raw(params, p_text.ptr + t0, p_off.ptr + 8 * r0, p_len.ptr + 2 * r0, n, False, True, out.ptr + r0 * mb * 68)
            else:
                m.vote_chunk_raw(params, fq.ptr, off.ptr + 8 * r0, ln.ptr + 2 * r0, n, False, True, out.ptr + r0 * mb * 68)
            done += n
        th.join()
        assert done == R, done

    out_c = PinnedArray(R * mb * 68, np.uint8)
    for name, fn in (("fastq_in_place", in_place), ("fastq_packed", packed), ("dense_arrays_ready", ready_made),
                     ("fastq_in_place_pipelined", lambda: pipelined(False, out_c)), ("fastq_packed_pipelined", lambda: pipelined(True, out_c))):
        for _ in range(args.warmup):
            fn()
        dt = timed(fn, args.steps)
        t = m.timing()
        res[name + "_reads_per_s"] = R / dt
        res[name + "_h2d_bytes"] = int(t.h2d_bytes)
    from subread_b200 import MAPPING_RESULT
    ga = np.frombuffer(out_a.array, dtype=MAPPING_RESULT).reshape(R, 1, mb)
    res["in_place_equals_packed"] = bool(out_a.array.tobytes() == out_b.array.tobytes())
    res["pipelined_equals_packed"] = bool(out_c.array.tobytes() == out_b.array.tobytes())
    res["window_bytes"] = args.window
    res["batch_frags"] = int(os.environ["SRCUDA_BATCH_FRAGS"])
    res["slot0_is_true_position"] = float((ga[:, 0, 0]["selected_position"].astype(np.int64) == truth).mean())
    print(json.dumps(res), flush=True)
    m.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
