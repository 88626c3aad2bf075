#!/usr/bin/env python
"""bench.py -- reads/s of the seed-and-vote voting step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--impl reference]

Workload at N=1 (BASELINE.json configs[1]): seeded synthetic 100 Mbp genome (200 contigs x
500 kbp, gen_long_chromosomes shape), one-block gapped index in the reference's format built by
subread_b200.index_build, R = 10 M error-free 150 bp SE reads sampled uniformly from both
strands.  One "step" = the voting step of do_voting (reference src/core.c:3049) over the whole
chunk of R reads.  For N > 1 every rank maps its own R reads (weak scaling, reads shard, no
collective on the data path); rank 0 owns the index and the other ranks pull it over NVLink
through CUDA IPC (srcuda_index_import_ipc).

Reported on one JSON line (rank 0):
  value     reads/s, whole job, inputs already resident in HBM (srcuda_vote_chunk_device), timed with
            CUDA events on the library's stream, max over ranks
  e2e       the same metric through srcuda_vote_chunk with pinned HOST buffers: per step the reads are
            copied host->device and the bigtable device->host inside the timed region
  roofline  fused kernel k_map_warp: algorithmic bytes (counted exactly by the oracle on
            a sample of the same reads) / its CUDA-event duration, vs the measured HBM copy bandwidth;
            also vs a measured random-32-B-sector gather bandwidth over a table of the index's size
  cpu_baseline  the unmodified reference `subread-align -T <cores>` (oracle/_ref) on a bounded sample of
            the same reads and the same index files, on this box's host cores (N=1 only)

--impl reference times the reference's own CPU implementation instead (no GPU code on that path).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GENOME_SEED = 20261018
N_CHR, CHR_LEN, READ_LEN = 200, 500000, 150
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin", "subread-align")


# ------------------------------------------------------------------------------ workload

def make_genome(n_chr=N_CHR, chr_len=CHR_LEN, seed=GENOME_SEED):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    return [f"chr{c + 1}" for c in range(n_chr)], [acgt[rng.integers(0, 4, size=chr_len, dtype=np.uint8)] for _ in range(n_chr)]


def sample_reads_numpy(seqs, n, read_len, seed):
    """Host-side read sampler (reference arm / cpu_baseline sample): error-free, both strands."""
    rng = np.random.default_rng(seed)
    g = np.concatenate(seqs)
    chr_len = seqs[0].shape[0]
    comp = np.full(256, ord("N"), dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    c = rng.integers(0, len(seqs), size=n)
    s = rng.integers(0, chr_len - read_len, size=n)
    rc = rng.random(n) < 0.5
    idx = (c * chr_len + s)[:, None] + np.arange(read_len)[None, :]
    m = g[idx]
    m[rc] = comp[m[rc][:, ::-1]]
    return m


def sample_reads_torch(seqs, n, read_len, seed, device):
    """Same distribution, generated on the GPU (plumbing): returns a uint8 tensor [n, read_len]."""
    import torch
    g = torch.from_numpy(np.concatenate(seqs)).to(device)
    chr_len = seqs[0].shape[0]
    comp = torch.full((256,), ord("N"), dtype=torch.uint8, device=device)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.empty((n, read_len), dtype=torch.uint8, device=device)
    ar = torch.arange(read_len, device=device)
    for lo in range(0, n, 1 << 20):
        hi = min(n, lo + (1 << 20))
        c = torch.randint(0, len(seqs), (hi - lo,), generator=gen, device=device)
        s = torch.randint(0, chr_len - read_len, (hi - lo,), generator=gen, device=device)
        rc = torch.rand(hi - lo, generator=gen, device=device) < 0.5
        m = g[(c * chr_len + s)[:, None] + ar[None, :]]
        m = torch.where(rc[:, None], comp[m.flip(1).long()], m)
        out[lo:hi] = m
    return out


def write_fastq(path, reads):
    n, L = reads.shape
    rec = np.empty((n, 2 * L + 18), dtype=np.uint8)
    names = np.char.zfill(np.arange(n).astype(str), 11)
    hdr = np.frombuffer(("".join("@R" + s for s in names)).encode(), dtype=np.uint8).reshape(n, 13)
    rec[:, :13] = hdr
    rec[:, 13] = 10
    rec[:, 14:14 + L] = reads
    rec[:, 14 + L] = 10
    rec[:, 15 + L] = ord("+")
    rec[:, 16 + L] = 10
    rec[:, 17 + L:17 + 2 * L] = ord("H")
    rec[:, 17 + 2 * L] = 10
    # 13 + 1 + L + 1 + 1 + 1 + L + 1 = 2L + 18
    with open(path, "wb") as fh:
        fh.write(rec[:, :2 * L + 18].tobytes())


# ------------------------------------------------------------------------------ clocks

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ reference arm

def run_reference_align(prefix, fastq, threads, out_dir):
    """One run of the unmodified reference on the sample; returns wall seconds."""
    out = os.path.join(out_dir, "ref_out.bam")
    t0 = time.perf_counter()
    subprocess.run([REF_BIN, "-t", "1", "-T", str(threads), "-i", prefix, "-r", fastq, "-o", out], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0


def reference_setup(sample_reads, seed, tmp, names=None, seqs=None, index_arrays=None):
    """Index files + sample FASTQ on disk for the reference binary."""
    from subread_b200.index_build import build_index_arrays, write_index_files
    if seqs is None:
        names, seqs = make_genome()
    if index_arrays is None:
        index_arrays = build_index_arrays(names, seqs, gapped=True, use_gpu=False)
    prefix = os.path.join(tmp, "bench_idx")
    write_index_files(prefix, index_arrays)
    reads = sample_reads_numpy(seqs, sample_reads, READ_LEN, seed)
    fq = os.path.join(tmp, "sample.fastq")
    write_fastq(fq, reads)
    return prefix, fq, reads


def host_threads():
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return max(1, min(n, 40))  # -T is clamped to MAX_THREADS = 40 (core-interface-aligner.c:399-404)


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    config = {"workload": f"synthetic 100 Mbp genome ({N_CHR}x{CHR_LEN} bp), gapped one-block index, {READ_LEN} bp SE reads",
              "reads_per_gpu": args.reads, "read_len": READ_LEN, "inputs": "larger than L2"}
    base = {"metric": "reads/sec mapped", "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": config, "impl": "reference"}
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bin/subread-align is not built (reference sources absent at build time)"}))
        return 0
    tmp = tempfile.mkdtemp(prefix="srbench_ref_")
    try:
        threads = host_threads()
        sample = args.ref_sample
        prefix, fq, _ = reference_setup(sample, 777, tmp)
        for _ in range(args.warmup):
            run_reference_align(prefix, fq, threads, tmp)
        times = [run_reference_align(prefix, fq, threads, tmp) for _ in range(args.steps)]
        t = float(np.mean(times))
        v = sample / t
        base.update({"value": v, "ms_per_step": t * 1e3, "gpu_launches": 0,
                     "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "reference",
                                      "sample": f"{sample} of the workload's reads per step; subread-align -t 1 -T {threads}, BAM output, "
                                                "whole process wall time (index load + voting + realignment + output)"},
                     "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


# ------------------------------------------------------------------------------ our arm

def bind_to_gpu_numa_node(gpu_index):
    """Best effort: run this rank (and so first-touch its pinned chunk buffers) on the CPUs next to its GPU, so that the
    host-buffer leg does not cross the socket interconnect.  Does nothing when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        near = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        mine = near & set(os.sched_getaffinity(0))
        if mine:
            os.sched_setaffinity(0, mine)
        # ... and prefer the GPU's memory node for the pinned buffers even when none of its CPUs is ours (set_mempolicy
        # MPOL_PREFERRED; a no-op on single-node boxes and where the container's memory set forbids it)
        import ctypes
        pci = pynvml.nvmlDeviceGetPciInfo(h).busId
        pci = (pci.decode() if isinstance(pci, bytes) else pci).lower()
        if pci.startswith("0000") and len(pci.split(":")[0]) == 8:
            pci = pci[4:]
        node = int(open(f"/sys/bus/pci/devices/{pci}/numa_node").read().strip())
        if node >= 0:
            mask = ctypes.c_ulong(1 << node)
            ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))   # x86-64 set_mempolicy
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-sample", type=int, default=1_000_000, help="reads per step of the reference arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from subread_b200 import VoteParams
    from subread_b200.binding import Mapper, PinnedArray
    from subread_b200.index_build import build_index_arrays

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- index: rank 0 builds, the others pull it over NVLink (CUDA IPC) --------------------------
    names, seqs = make_genome()
    m = Mapper(local)
    index_arrays = None
    t_rep = 0.0
    if rank == 0:
        index_arrays = build_index_arrays(names, seqs, gapped=True, use_gpu=True)
        a = index_arrays
        m.index_upload(a["index_gap"], a["padding"], a["bucket_start"], a["keys"], a["vals"], a["start_point"], a["length"],
                       a["values"], a["read_offsets"])
    if world > 1:
        blob = torch.zeros(Mapper.IPC_BLOB_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            blob.copy_(torch.frombuffer(bytearray(m.index_export_ipc()), dtype=torch.uint8))
        dist.broadcast(blob, 0)
        barrier()
        t0 = time.perf_counter()
        if rank != 0:
            m.index_import_ipc(bytes(blob.cpu().numpy().tobytes()))
        t_rep = max_over_ranks(time.perf_counter() - t0)
        barrier()
    desc = m.describe()

    # ---- reads: R per rank, generated on the device, mirrored into pinned host memory -------------
    R = args.reads
    params = VoteParams.defaults(False)
    mb = params.multi_best_reads
    d_text = sample_reads_torch(seqs, R, READ_LEN, 1000 + rank, dev).reshape(-1)
    d_off = (torch.arange(R, device=dev, dtype=torch.int64) * READ_LEN)
    d_len = torch.full((R,), READ_LEN, dtype=torch.int16, device=dev)
    d_out = torch.zeros(R * mb * 68, dtype=torch.uint8, device=dev)
    h_text = PinnedArray(R * READ_LEN, np.uint8)
    h_off = PinnedArray(R, np.uint64)
    h_len = PinnedArray(R, np.uint16)
    h_out = PinnedArray(R * mb * 68, np.uint8)
    h_text.array[:] = d_text.cpu().numpy()
    h_off.array[:] = d_off.cpu().numpy().astype(np.uint64)
    h_len.array[:] = READ_LEN
    torch.cuda.synchronize()

    def step_device():
        # first index block of a chunk: the bigtable starts zeroed (core-bigtable.c:96), so it is not read
        m.vote_chunk_device(params, d_text.data_ptr(), d_text.numel(), d_off.data_ptr(), d_len.data_ptr(), R, False, READ_LEN, d_out.data_ptr(),
                            out_is_zero=True)
        return m.timing()

    def step_e2e():
        m.vote_chunk_raw(params, h_text.ptr, h_off.ptr, h_len.ptr, R, False, True, h_out.ptr)
        return m.timing()

    # ---- timed: HBM-resident ------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    w0 = time.perf_counter()
    ev_ms, kern_ms, over_ms, n_over, launches, simple_ms, n_punted = 0.0, 0.0, 0.0, 0, 0, 0.0, 0
    for _ in range(args.steps):
        t = step_device()
        ev_ms += t.ms_total
        kern_ms += t.ms_kernel_vote
        over_ms += t.ms_kernel_overflow
        n_over += t.n_overflow
        launches += t.n_launches
        simple_ms += t.ms_kernel_simple
        n_punted += t.n_simple_punted
    barrier()
    wall_dev = max_over_ranks(time.perf_counter() - w0)
    ev_ms = max_over_ranks(ev_ms)
    mapped = float((np.frombuffer(d_out[: 68 * mb * min(R, 200000)].cpu().numpy().tobytes(), dtype=np.uint8)
                    .reshape(-1, mb, 68)[:, 0, 8:10].view(np.int16) > 0).mean())

    # ---- timed: end to end through the host-buffer entry point -----------------------------------
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    w0 = time.perf_counter()
    e2e_ev, h2d, d2h = 0.0, 0, 0
    for _ in range(args.steps):
        t = step_e2e()
        e2e_ev += t.ms_total
        h2d, d2h = t.h2d_bytes, t.d2h_bytes
        launches_e2e = t.n_launches
    barrier()
    wall_e2e = max_over_ranks(time.perf_counter() - w0)
    clk = clocks.stop() if rank == 0 else None

    # e2e result must equal the device-resident result (same reads, same index)
    same = bool(np.array_equal(h_out.array[: 68 * mb * 100000], d_out[: 68 * mb * 100000].cpu().numpy()))

    # ---- gather roofline probe ------------------------------------------------------------------------
    gather_gbs = None
    if rank == 0:
        try:
            gather_gbs = m.gather_bandwidth(max(int(desc.device_bytes), 1 << 28), 3)
        except Exception:
            gather_gbs = None

    # ---- cpu baseline + algorithmic bytes (rank 0, N=1 only; the one place oracle/ is executed) -------
    cpu_baseline, alg_bytes_per_read = None, None
    if rank == 0:
        try:
            from oracle import oracle
            oidx_tmp = tempfile.mkdtemp(prefix="srbench_")
            try:
                sample_n = args.ref_sample
                prefix, fq, sample = reference_setup(sample_n, 777, oidx_tmp, names, seqs, index_arrays)
                # exact algorithmic bytes per read, counted by the oracle on 100k sample reads
                from subread_b200.reads import ReadChunk
                k = min(100000, sample_n)
                ch = ReadChunk(np.ascontiguousarray(sample[:k]).reshape(-1), np.arange(k + 1, dtype=np.uint64) * READ_LEN,
                               np.full(k, READ_LEN, np.uint16), k, False, [])
                oi = oracle.OracleIndex(prefix)
                _, _, st = oracle.vote_chunk(oi, params, ch, want_verify=False, threads=min(8, host_threads()))
                alg_bytes_per_read = st.algorithmic_bytes() / k
                oi.close()
                if world == 1 and not args.no_cpu_baseline:
                    threads = host_threads()
                    if os.path.exists(REF_BIN):
                        run_reference_align(prefix, fq, threads, oidx_tmp)  # warm the page cache
                        sec = run_reference_align(prefix, fq, threads, oidx_tmp)
                        cpu_baseline = {"value": sample_n / sec, "unit": "reads/s", "cores": threads, "kind": "reference",
                                        "sample": f"{sample_n} reads of the same workload; unmodified subread-align -t 1 -T {threads}, BAM "
                                                  "output, whole process wall time (2nd of two runs)"}
                    else:
                        ch2 = ReadChunk(np.ascontiguousarray(sample).reshape(-1), np.arange(sample_n + 1, dtype=np.uint64) * READ_LEN,
                                        np.full(sample_n, READ_LEN, np.uint16), sample_n, False, [])
                        oi = oracle.OracleIndex(prefix)
                        t0 = time.perf_counter()
                        oracle.vote_chunk(oi, params, ch2, threads=threads)
                        sec = time.perf_counter() - t0
                        cpu_baseline = {"value": sample_n / sec, "unit": "reads/s", "cores": threads, "kind": "port",
                                        "sample": f"{sample_n} reads; oracle port of the voting step only, {threads} threads"}
            finally:
                shutil.rmtree(oidx_tmp, ignore_errors=True)
        except Exception as e:  # the baseline is reporting, never the product path
            cpu_baseline = {"error": repr(e)}

    if rank == 0:
        total_reads = R * world
        ms_step = ev_ms / args.steps
        value = total_reads / (ms_step * 1e-3)
        kern_ms_step = kern_ms / args.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        roofline = None
        if alg_bytes_per_read:
            achieved = alg_bytes_per_read * R / (kern_ms_step * 1e-3) / 1e9
            traffic = None
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic = float(tr["dram_bytes_per_read"]) * R  # ncu --set full capture of this kernel, scaled per read
            except Exception:
                pass
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650",
                        "kernel": "k_map_simple<3,0> (tier 0, one thread per fragment) + k_map_warp<1,3> (queued fragments)",
                        "kernel_ms": kern_ms_step, "tier0_ms": simple_ms / args.steps, "tier0_punted_fraction": n_punted / args.steps / R, "algorithmic_bytes_per_read": alg_bytes_per_read,
                        # random 32-byte records gathered per second from a table of the index's size, through the lookup's own
                        # 64-byte-fetch load; a DRAM miss costs one line whatever is taken from it, so this is a rate of lines
                        # (profiles/r01_t0_k_map_simple.md).  The Bloom words of this workload are served by L2, so only the
                        # ~16 bucket records per read count against it.
                        "gather_peak": gather_gbs, "frac_of_gather": (achieved / gather_gbs) if gather_gbs else None,
                        "gather_records_per_s": (gather_gbs * 1e9 / 32.0) if gather_gbs else None,
                        "dram_bytes_per_read": (traffic / R) if traffic else None}
        out = {
            "metric": "reads/sec mapped", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"synthetic 100 Mbp genome ({N_CHR}x{CHR_LEN} bp), gapped one-block index "
                                   f"({desc.n_items} items, {desc.buckets_number} buckets), {READ_LEN} bp SE reads",
                       "reads_per_gpu": R, "read_len": READ_LEN, "inputs": "larger than L2 (reads %.2f GB + index %.2f GB per GPU)"
                       % (R * READ_LEN / 1e9, desc.device_bytes / 1e9), "parallelism": f"reads sharded x{world}, index replicated"},
            "wall_ms_per_step": wall_dev * 1e3 / args.steps,
            "e2e": {"value": total_reads * args.steps / wall_e2e, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step_wall": wall_e2e * 1e3 / args.steps, "ms_per_step_events": e2e_ev / args.steps, "matches_device_result": same},
            "gpu_launches": int(launches + launches_e2e * args.steps),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clk,
            "mapped_fraction": mapped, "overflow_fragments_per_step": n_over / args.steps, "overflow_kernel_ms": over_ms / args.steps,
            "index_replication_s": t_rep if world > 1 else None,
        }
        print(json.dumps(out))
    m.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
