#!/usr/bin/env python
"""bench.py -- reads/s of the seed-and-vote voting step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--reads R] [--impl reference]

Workloads (BASELINE.json configs; all synthetic and seeded, tools/synth.py):
  cfg2_se        configs[1]: 100 Mbp genome (200 x 500 kbp), one-block gapped index, 10 M x 150 bp SE reads   [default at N = 1]
  cfg2_pe        the same index, 5 M pairs of 2x150 bp
  human_pe       configs[2]: 3.1 Gbp genome (6200 x 500 kbp), one-block gapped index (1.03 G items), 2x150 bp PE, one reference
                 chunk (3 495 253 pairs, core.c:4092) per GPU per step                                        [default at N > 1]
  human_subjunc  configs[3]: the same genome with planted GT..AG introns, subjunc mode, 2x100 bp pairs, 30 % spliced
  sweep          configs[4]: SE reads of 50..250 bp on the cfg2 index; `value` is the 150 bp point, all points in "sweep"
  mixed          cfg2 genome with planted repeats; 150 bp SE reads with 1 % substitutions, 0.2 % N per base, one 1-3 bp indel in
                 14 % of the reads (0.1 % per base) and 5 % of the reads drawn from repeats -- what leaves the fast tier

One "step" = the voting step of do_voting (reference src/core.c:3049) over one chunk.  For N > 1 every rank maps its own
chunk (weak scaling, reads shard, no collective on the data path); rank 0 owns the index and the other ranks pull it over
NVLink through CUDA IPC (srcuda_index_import_ipc).  Every rank checks its own result (slot 0 = true position of the clean
reads) and rank 0 re-maps a prefix of every rank's reads with its own copy of the index and compares checksums.

One JSON line (rank 0):
  value     reads/s (mates), whole job, inputs already resident in HBM (srcuda_vote_chunk_device), CUDA events on the
            library's stream, max over ranks
  e2e       the same through srcuda_vote_chunk with pinned HOST buffers: reads host->device and bigtable device->host
            inside the timed region
  roofline  the fast tier's kernel: algorithmic bytes (counted exactly by the oracle on a sample of the same
            reads) / its CUDA-event duration vs the measured HBM copy peak; "gather": the random index fetches the kernel issued
            (counted by the kernel itself, srcuda_timing.n_gather_requests) per second vs the measured random-line rate over
            a table of the index's size; "traffic": DRAM bytes from the committed ncu capture, only if that capture was taken
            from this very kernel source (hash check), else null
  cpu_baseline / --impl reference   the UNMODIFIED reference on this box's host cores, on a bounded sample of the same
            workload: `value` = sample reads / the reference's own clock around run_maybe_threads(STEP_VOTING)
            (timecost_voting, core.c:3592-3595, read out by oracle/ref_hooks.c) -- the same step the GPU arm times; the whole
            process (index load + voting + realignment + BAM output) is reported next to it
  whole_program (N = 1, cfg2_se)   wall time of `subread-align-cuda` (the reference's CLI with STEP_VOTING bound to this
            library, integration/do_voting_cuda.c) against the stock `subread-align`, same reads, same -T
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
import zlib
from ctypes import memset as ctypes_memset

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

GENOME_SEED = 20261018
CHR_LEN = 500000
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_BIN = os.path.join(REF_DIR, "bin", "subread-align")
REF_SUBJUNC = os.path.join(REF_DIR, "bin", "subjunc")
REF_DUMP = os.path.join(REF_DIR, "ref_dump")
REF_DUMP_SUBJUNC = os.path.join(REF_DIR, "ref_dump_subjunc")
CUDA_BIN = os.path.join(REF_DIR, "subread-align-cuda")
PE_CHUNK = 3495253      # reads_per_chunk of a paired-end run (core-indel.c:4483, core.c:4092-4093)

WORKLOADS = {
    "cfg2_se": dict(n_chr=200, paired=False, subjunc=False, read_len=150, frags=10_000_000,
                    what="synthetic 100 Mbp genome (200x500000 bp), gapped one-block index, 150 bp SE reads"),
    "cfg2_pe": dict(n_chr=200, paired=True, subjunc=False, read_len=150, frags=5_000_000,
                    what="synthetic 100 Mbp genome (200x500000 bp), gapped one-block index, 2x150 bp PE reads"),
    "human_pe": dict(n_chr=6200, paired=True, subjunc=False, read_len=150, frags=PE_CHUNK,
                     what="synthetic 3.1 Gbp genome (6200x500000 bp), gapped one-block index, 2x150 bp PE reads, one reference chunk per GPU"),
    "human_subjunc": dict(n_chr=6200, paired=True, subjunc=True, read_len=100, frags=PE_CHUNK, introns=True, spliced=0.3,
                          what="synthetic 3.1 Gbp genome (6200x500000 bp) with planted GT..AG introns, gapped one-block index, subjunc mode, "
                               "2x100 bp PE reads, 30 % of the pairs spliced"),
    "sweep": dict(n_chr=200, paired=False, subjunc=False, read_len=150, frags=4_000_000, sweep=[50, 75, 100, 125, 150, 200, 250],
                  what="synthetic 100 Mbp genome (200x500000 bp), gapped one-block index, SE reads of 50-250 bp (value: 150 bp)"),
    "mixed": dict(n_chr=200, paired=False, subjunc=False, read_len=150, frags=10_000_000, repeats=(500, 5, 2000),
                  noise=dict(sub_rate=0.01, n_rate=0.002, indel_read_rate=0.14, repeat_fraction=0.05),
                  what="synthetic 100 Mbp genome (200x500000 bp) with 500 repeat families (5 copies x 2 kbp), gapped one-block index, 150 bp SE "
                       "reads: 1 % substitutions, 0.2 % N per base, one 1-3 bp indel in 14 % of the reads, 5 % of the reads from repeats"),
}


def workload_config(name, frags, world):
    """The `config` object of the JSON line -- the same function serves both arms, so the strings are identical."""
    w = WORKLOADS[name]
    mates = 2 if w["paired"] else 1
    return {"workload": f"{name}: {w['what']}", "fragments_per_gpu": int(frags), "reads_per_gpu": int(frags) * mates, "read_len": w["read_len"],
            "mode": ("subjunc " if w["subjunc"] else "subread-align ") + ("PE" if w["paired"] else "SE"),
            "inputs": "larger than L2 (126 MB): the reads of one step alone are %.2f GB per GPU" % (frags * mates * w["read_len"] / 1e9),
            "parallelism": f"reads sharded x{world}, index replicated"}


def make_genome(name, device):
    import synth
    w = WORKLOADS[name]
    g = synth.Genome(w["n_chr"], CHR_LEN, GENOME_SEED, device)
    if w.get("repeats"):
        g.plant_repeats(*w["repeats"])
    if w.get("introns"):
        g.plant_introns()
    return g


def make_reads(g, name, frags, read_len, seed):
    w = WORKLOADS[name]
    kw = dict(w.get("noise", {}))
    if w.get("spliced"):
        kw["spliced_fraction"] = w["spliced"]
    return g.reads(frags, read_len, seed, paired=w["paired"], **kw)


# ------------------------------------------------------------------------------ clocks

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
                for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ the reference on the host cores

def host_threads():
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return max(1, min(n, 40))  # -T is clamped to MAX_THREADS = 40 (core-interface-aligner.c:399-404)


def reference_files(tmp, g, index_arrays, name, sample_frags, seed):
    """Index files + sample FASTQ on disk for the reference binaries.  Returns (prefix, fastq paths, reads as a host array)."""
    import synth
    from subread_b200.index_build import write_index_files
    w = WORKLOADS[name]
    prefix = os.path.join(tmp, "bench_idx")
    write_index_files(prefix, index_arrays)
    text, _, _ = make_reads(g, name, sample_frags, w["read_len"], seed)
    mates = 2 if w["paired"] else 1
    reads = text.view(sample_frags, mates, w["read_len"]).cpu().numpy()
    fqs = []
    for m in range(mates):
        fq = os.path.join(tmp, f"sample_R{m + 1}.fastq")
        synth.write_fastq(fq, np.ascontiguousarray(reads[:, m]))
        fqs.append(fq)
    return prefix, fqs, reads


def run_reference(prefix, fqs, threads, out_dir, subjunc, exe=None, env_extra=None):
    """One run of the unmodified reference on the sample (default BAM output).  Uses the instrumented build (ref_dump: the same
    sources compiled -fPIC with oracle/ref_hooks.c reading the reference's own clocks) when it is there.  Returns
    {"wall_s", "voting_s" | None, "load_index_s" | None}."""
    dump = REF_DUMP_SUBJUNC if subjunc else REF_DUMP
    stock = REF_SUBJUNC if subjunc else REF_BIN
    exe = exe or (dump if os.path.exists(dump) else stock)
    out = os.path.join(out_dir, "ref_out.bam")
    tfile = os.path.join(out_dir, "ref_timing.txt")
    if os.path.exists(tfile):
        os.remove(tfile)
    cmd = [exe, "-T", str(threads), "-i", prefix, "-r", fqs[0], "-o", out]
    if not subjunc:
        cmd += ["-t", "1"]
    if len(fqs) > 1:
        cmd += ["-R", fqs[1]]
    t0 = time.perf_counter()
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=dict(os.environ, SRDUMP_TIMING=tfile, **(env_extra or {})))
    res = {"wall_s": time.perf_counter() - t0, "voting_s": None, "load_index_s": None, "exe": os.path.basename(exe)}
    if os.path.exists(tfile):
        for line in open(tfile):
            f = dict(x.split("=") for x in line.split()[1:])
            res["voting_s"], res["load_index_s"] = float(f["voting_s"]), float(f["load_index_s"])    # accumulated clocks: the last line counts
    return res


def reference_measure(g, index_arrays, name, sample_frags, runs, warm, tmp):
    """-> dict for cpu_baseline / the reference arm: reads/s of the reference's voting step on `sample_frags` fragments."""
    w = WORKLOADS[name]
    mates = 2 if w["paired"] else 1
    threads = host_threads()
    prefix, fqs, _ = reference_files(tmp, g, index_arrays, name, sample_frags, 777)
    for _ in range(warm):
        run_reference(prefix, fqs, threads, tmp, w["subjunc"])
    rs = [run_reference(prefix, fqs, threads, tmp, w["subjunc"]) for _ in range(max(1, runs))]
    wall = float(np.mean([r["wall_s"] for r in rs]))
    vote = [r["voting_s"] for r in rs if r["voting_s"]]
    reads = sample_frags * mates
    out = {"unit": "reads/s", "cores": threads, "kind": "reference", "whole_process_reads_per_s": reads / wall, "whole_process_s": wall,
           "index_load_s": rs[-1]["load_index_s"]}
    tool = "subjunc" if w["subjunc"] else "subread-align -t 1"
    if vote:
        v = float(np.mean(vote))
        out.update({"value": reads / v, "voting_step_s": v,
                    "sample": f"{sample_frags} fragments ({reads} reads) of the same workload per run; unmodified reference ({tool} -T {threads}, BAM "
                              f"output, built -fPIC with oracle/ref_hooks.c reading its clocks); value = reads / the reference's own timecost_voting "
                              f"(clock around run_maybe_threads(STEP_VOTING), core.c:3592-3595: FASTQ parsing, voting and the event-table merge), "
                              f"whole process next to it"})
    else:
        out.update({"value": reads / wall,
                    "sample": f"{sample_frags} fragments ({reads} reads) of the same workload per run; unmodified reference ({tool} -T {threads}, BAM "
                              f"output), whole process wall time (index load + voting + realignment + output)"})
    return out


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on all host threads; no code of this repository on
    the timed path (the index files and reads are made with torch beforehand, on the GPU when there is one)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    w = WORKLOADS[name]
    frags = args.reads or w["frags"]
    base = {"metric": "reads/sec mapped", "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(name, frags, args.gpus), "impl": "reference"}
    if not (os.path.exists(REF_BIN) or os.path.exists(REF_DUMP)):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built (reference sources absent at build time)"}))
        return 0
    import torch
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    tmp = tempfile.mkdtemp(prefix="srbench_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") and w["n_chr"] <= 200 else None)
    try:
        g = make_genome(name, dev)
        a = g.build_index(gapped=True)
        # bounded sample: the whole --steps/--warmup run stays within a few minutes
        sample = args.ref_sample or (1_000_000 if w["n_chr"] <= 200 else 200_000)
        sample = max(10_000, sample // (2 if w["paired"] else 1))
        # every run of the reference is a whole process (it loads the index: a minute for the human-sized one), so the number of runs is
        # bounded as well: the timed figure is the reference's own voting-step clock, which does not include the load
        human = w["n_chr"] > 200
        m = reference_measure(g, a, name, sample, min(args.steps, 1 if human else 3), 0 if human else min(args.warmup, 1), tmp)
        base.update({"value": m["value"], "ms_per_step": (m.get("voting_step_s") or m["whole_process_s"]) * 1e3, "gpu_launches": 0, "cpu_baseline": m,
                     "e2e": {"value": m["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


def whole_program(g, index_arrays, name, sample_frags, tmp):
    """Wall time of the reference's CLI with STEP_VOTING bound to libsubread_cuda.so against the stock binary: same index files, same
    reads, same -T, default BAM output, second of two runs each (page cache warm)."""
    if not (os.path.exists(CUDA_BIN) and os.path.exists(REF_BIN)) or WORKLOADS[name]["subjunc"]:
        return None
    threads = host_threads()
    prefix, fqs, _ = reference_files(tmp, g, index_arrays, name, sample_frags, 778)
    out = {"reads": sample_frags * len(fqs), "threads": threads}
    for tag, exe in (("stock", REF_BIN), ("cuda", CUDA_BIN)):
        try:
            run_reference(prefix, fqs, threads, tmp, False, exe)
            out[tag + "_s"] = run_reference(prefix, fqs, threads, tmp, False, exe)["wall_s"]
        except Exception as e:
            out[tag + "_error"] = repr(e)
    if "stock_s" in out and "cuda_s" in out:
        out["speedup"] = out["stock_s"] / out["cuda_s"]
        out["note"] = ("subread-align-cuda = the reference's own main, chunk loop, realignment and BAM writer with do_voting replaced "
                       "(integration/do_voting_cuda.c); everything but the voting step still runs on the host")
    # ... and with the passes after voting reading their chunk from memory (integration/read_cache.c, row f3 behind the voting step)
    try:
        cache = {"SRCUDA_READ_CACHE": "1"}
        run_reference(prefix, fqs, threads, tmp, False, CUDA_BIN, cache)
        out["cuda_read_cache_s"] = run_reference(prefix, fqs, threads, tmp, False, CUDA_BIN, cache)["wall_s"]
        if "stock_s" in out:
            out["speedup_read_cache"] = out["stock_s"] / out["cuda_read_cache_s"]
    except Exception as e:
        out["cuda_read_cache_error"] = repr(e)
    return out


def output_writer_measure(tmp, n_records=200_000, read_len=150):
    """Row f2 on this box's host cores: the library's BAM writer (srcuda_out_*, one producer + a team of deflate threads) against
    the reference's single stream (SamBam_writer_add_read_line) on the same synthetic SAM records, timed by
    oracle/_ref/ref_writer_bench (the reference half is the baseline, like cpu_baseline); the two files are compared."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_writer_bench")
    if not os.path.exists(exe):
        return None
    rng = np.random.default_rng(5)
    seqs = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n_records, read_len))]
    quals = (33 + rng.integers(2, 41, (n_records, read_len))).astype(np.uint8)
    pos = rng.integers(1, 400_000, n_records)
    sam = os.path.join(tmp, "writer_bench.sam")
    with open(sam, "wb") as fh:
        fh.write(b"@HD\tVN:1.0\tSO:unsorted\n" + b"".join(b"@SQ\tSN:chr%d\tLN:500000\n" % c for c in range(200)))
        for i in range(n_records):
            fh.write(b"read%d\t%d\tchr%d\t%d\t40\t%dM\t*\t0\t0\t%s\t%s\tHI:i:1\tNH:i:1\tNM:i:%d\n" % (
                i, 16 * (i & 1), i % 200, pos[i], read_len, seqs[i].tobytes(), quals[i].tobytes(), i % 3))
    threads = [t for t in (1, 4, 8, 16) if t <= max(1, host_threads())]
    p = subprocess.run([exe, sam, "1", tmp] + [str(t) for t in threads], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    out = {"records": n_records, "read_len": read_len, "unit": "records/s", "library": {}}
    for line in p.stdout.decode().splitlines():
        if line.startswith("reference writer"):
            out["reference_single_stream"] = float(line.split("=")[-1].split()[0]) * 1e6
        elif line.startswith("library writer"):
            out["library"]["%d_deflate_threads" % int(line.split(",")[1].split()[0])] = float(line.split("=")[-1].split()[0]) * 1e6
    ref_bam = os.path.join(tmp, "ref.bam")
    if os.path.exists(ref_bam):
        raw = open(ref_bam, "rb").read()
        out["files_identical"] = all(open(os.path.join(tmp, "native_%d.bam" % t), "rb").read() == raw for t in threads)
    if p.returncode != 0:
        out["error"] = "ref_writer_bench exit %d" % p.returncode
    return out


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


def kernel_source_hash():
    """sha256 over the kernel sources: ties profiles/traffic.json (an ncu capture) to the code it was taken from."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "subread_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cuh", ".cu", ".h")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="auto", choices=["auto"] + list(WORKLOADS))
    ap.add_argument("--reads", type=int, default=0, help="fragments per GPU per step (default: the workload's)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-sample", type=int, default=0, help="reads per run of the reference arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "auto":
        args.workload = "cfg2_se" if max(world, args.gpus) == 1 else "human_pe"
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from subread_b200 import MAPPING_RESULT, VoteParams
    from subread_b200.binding import Mapper, PinnedArray
    from subread_b200.reads import ReadChunk

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

    name = args.workload
    w = WORKLOADS[name]
    paired, subjunc = w["paired"], w["subjunc"]
    mates = 2 if paired else 1
    R = args.reads or w["frags"]

    # ---- index: rank 0 builds, the others pull it over NVLink (CUDA IPC) --------------------------
    g = make_genome(name, dev)         # every rank holds the genome (its reads are drawn from it); seeded, so identical
    m = Mapper(local)
    index_arrays = None
    t_rep = 0.0
    if rank == 0:
        index_arrays = g.build_index(gapped=True)
        a = index_arrays
        m.index_upload(a["index_gap"], a["padding"], a["bucket_start"], a["keys"], a["vals"], a["start_point"], a["length"],
                       a["values"], a["read_offsets"])
        lin_first = g.lin_first
    else:
        lin_first = torch.empty(w["n_chr"], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(lin_first, 0)
        g.lin_first = lin_first
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
    torch.cuda.empty_cache()
    params = VoteParams.defaults(paired, subjunc)
    mb = params.multi_best_reads

    def run_shape(RL, want_e2e, steps, warmup):
        """One read length: device-resident and host-buffer timing.  Returns a dict of this rank's figures (timings already
        max-reduced over the ranks)."""
        d_text, truth, clean = make_reads(g, name, R, RL, 1000 + 17 * rank + RL)
        nr = R * mates
        d_off = torch.arange(nr, device=dev, dtype=torch.int64) * RL
        d_len = torch.full((nr,), RL, dtype=torch.int16, device=dev)
        d_out = torch.zeros(nr * mb * 68, dtype=torch.uint8, device=dev)
        d_junc = d_bm = None
        if subjunc:
            d_junc = torch.zeros(nr * mb * 16, dtype=torch.uint8, device=dev)
            d_bm = torch.zeros(nr * 27 * 2, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()          # the inputs were written on torch's stream; the library launches on its own

        def step_device():
            # first index block of a chunk: the bigtable starts zeroed (core-bigtable.c:96), so it is not read
            if subjunc:
                d_junc.zero_(); d_bm.zero_()
                torch.cuda.synchronize()
            m.vote_chunk_device(params, d_text.data_ptr(), d_text.numel(), d_off.data_ptr(), d_len.data_ptr(), R, paired, RL, d_out.data_ptr(),
                                out_is_zero=True, d_junc=d_junc.data_ptr() if subjunc else None, d_big_margin=d_bm.data_ptr() if subjunc else None)
            return m.timing()

        for _ in range(warmup):
            step_device()
        barrier()
        acc = dict(ev=0.0, kern=0.0, simple=0.0, over=0.0, n_over=0, launches=0, punted=0, gathers=0)
        w0 = time.perf_counter()
        for _ in range(steps):
            t = step_device()
            acc["ev"] += t.ms_total; acc["kern"] += t.ms_kernel_vote; acc["simple"] += t.ms_kernel_simple; acc["over"] += t.ms_kernel_overflow
            acc["n_over"] += t.n_overflow; acc["launches"] += t.n_launches; acc["punted"] += t.n_simple_punted; acc["gathers"] += t.n_gather_requests
        barrier()
        wall = max_over_ranks(time.perf_counter() - w0)
        ms_step = max_over_ranks(acc["ev"]) / steps
        # ---- this rank's own check: a clean read of a random genome reports its own position in slot 0 ----
        out = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=MAPPING_RESULT).reshape(R, mates, mb)
        tr, cl = truth.cpu().numpy(), clean.cpu().numpy()
        hit = (out[:, :, 0]["selected_position"].astype(np.int64) == tr).all(1)
        res = dict(RL=RL, ms_step=ms_step, wall_ms=wall * 1e3 / steps, kern_ms=acc["kern"] / steps, simple_ms=acc["simple"] / steps, over_ms=acc["over"] / steps,
                   n_over=acc["n_over"] / steps, launches=acc["launches"], punted=acc["punted"] / steps, gathers=acc["gathers"] / steps,
                   mapped=float((out[:, :, 0]["selected_votes"] > 0).mean()), clean_fraction=float(cl.mean()),
                   slot0_true=float(hit[cl].mean()) if cl.any() else None)
        K = min(R, 100_000)
        res["crc"] = zlib.crc32(out[:K].tobytes())
        res["prefix_text"] = d_text[: K * mates * RL]
        res["K"] = K
        # ---- K_verify (match_chro / gvindex_get on the packed genome), measured on a bounded sample: the same chunk
        # prefix with per-slot matched-base counts requested -- the request takes every fragment through the warp kernel ----
        res["verify"] = None
        if rank == 0 and not subjunc and not args.no_e2e:
            try:
                V = min(R, 1_000_000)
                d_ver = torch.zeros(V * mates * mb, dtype=torch.int16, device=dev)
                v_out = torch.zeros(V * mates * mb * 68, dtype=torch.uint8, device=dev)
                torch.cuda.synchronize()
                ms = []
                for i in range(3):
                    m.vote_chunk_device(params, d_text.data_ptr(), d_text.numel(), d_off.data_ptr(), d_len.data_ptr(), V, paired, RL, v_out.data_ptr(),
                                        d_verify=d_ver.data_ptr(), out_is_zero=True)
                    ms.append(m.timing().ms_total)
                ver = d_ver.cpu().numpy().reshape(V, mates, mb)
                full = (ver[:, :, 0] == RL).all(1)                  # every base of the best candidate of a clean read matches the genome
                same = bool((v_out.cpu().numpy() == d_out[: V * mates * mb * 68].cpu().numpy()).all())
                res["verify"] = {"reads_per_s": V * mates / (min(ms[1:]) * 1e-3), "sample": "%d fragments of the same chunk, best of 2 timed launches" % V,
                                 "kernel": "k_map_warp<..., VERIFY> (packed XOR + popc over the 2-bit genome, 32-bit loads)",
                                 "clean_reads_fully_matched": float(full[cl[:V]].mean()) if cl[:V].any() else None, "bigtable_same_as_unverified": same}
                del d_ver, v_out
            except Exception as e:      # a measurement, not a gate
                res["verify"] = {"error": repr(e)}
        # ---- end to end through the host-buffer entry point ----
        if want_e2e:
            h_text = PinnedArray(nr * RL, np.uint8)
            h_off = PinnedArray(nr, np.uint64)
            h_len = PinnedArray(nr, np.uint16)
            h_out = PinnedArray(nr * mb * 68, np.uint8)
            h_junc = PinnedArray(nr * mb * 16, np.uint8) if subjunc else None
            h_bm = PinnedArray(nr * 27, np.uint16) if subjunc else None
            h_text.array[:] = d_text.cpu().numpy()
            h_off.array[:] = d_off.cpu().numpy().astype(np.uint64)
            h_len.array[:] = RL

            def step_e2e():
                if subjunc:
                    h_junc.array[:] = 0
                    h_bm.array[:] = 0
                rc = m.lib.srcuda_vote_chunk(m.h, params, h_text.ptr, h_off.ptr, h_len.ptr, R, int(paired), 1, h_out.ptr,
                                             h_junc.ptr if subjunc else None, h_bm.ptr if subjunc else None, None)
                m._check(rc, "srcuda_vote_chunk")
                return m.timing()

            for _ in range(max(2, warmup)):
                step_e2e()
            # the caller's table of a chunk's first index block is all zero (the reference clears it per chunk, core-bigtable.c:96); the
            # timed steps then write the same records over each other
            ctypes_memset(h_out.ptr, 0, nr * mb * 68)
            barrier()
            w0 = time.perf_counter()
            e2e_ev = 0.0
            for _ in range(steps):
                t = step_e2e()
                e2e_ev += t.ms_total
            barrier()
            wall_e2e = max_over_ranks(time.perf_counter() - w0)
            res.update(e2e_wall_ms=wall_e2e * 1e3 / steps, e2e_ev_ms=e2e_ev / steps, h2d=int(t.h2d_bytes), d2h=int(t.d2h_bytes), e2e_launches=int(t.n_launches),
                       e2e_same=bool(np.array_equal(h_out.array[: 68 * mb * mates * K], d_out[: 68 * mb * mates * K].cpu().numpy())))
            for x in (h_text, h_off, h_len, h_out, h_junc, h_bm):
                if x is not None:
                    x.free()
        del d_off, d_len, d_out, d_junc, d_bm, truth, clean
        return res, d_text

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    sweep_rows = None
    if w.get("sweep"):
        sweep_rows = []
        main_res = main_text = None
        for RL in w["sweep"]:
            r, txt = run_shape(RL, RL == w["read_len"] and not args.no_e2e, args.steps, args.warmup)
            sweep_rows.append({"read_len": RL, "reads_per_s": R * mates * world / (r["ms_step"] * 1e-3), "kernel_ms": r["kern_ms"],
                               "tier0_punted_fraction": r["punted"] / R, "slot0_is_true_position": r["slot0_true"]})
            if RL == w["read_len"]:
                main_res, main_text = r, txt
            else:
                del txt
            torch.cuda.empty_cache()
        res, d_text = main_res, main_text
    else:
        res, d_text = run_shape(w["read_len"], not args.no_e2e, args.steps, args.warmup)
    clk = clocks.stop() if rank == 0 else None
    RL = w["read_len"]

    # ---- every rank's result is checked: its own truth check above, and rank 0 re-maps a prefix of every rank's reads with its
    # own copy of the index (the one that was built, not imported) and compares checksums ----
    K = res["K"]
    rank_rows = None
    if world > 1:
        texts = [torch.empty_like(res["prefix_text"]) for _ in range(world)]
        dist.all_gather(texts, res["prefix_text"].contiguous())
        mine = torch.tensor([res["crc"], int(1e6 * (res["slot0_true"] if res["slot0_true"] is not None else -1)), int(1e6 * res["mapped"])],
                            dtype=torch.int64, device=dev)
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        if rank == 0:
            rank_rows = []
            nk = K * mates
            k_off = torch.arange(nk, device=dev, dtype=torch.int64) * RL
            k_len = torch.full((nk,), RL, dtype=torch.int16, device=dev)
            k_out = torch.zeros(nk * mb * 68, dtype=torch.uint8, device=dev)
            k_junc = torch.zeros(nk * mb * 16, dtype=torch.uint8, device=dev) if subjunc else None
            k_bm = torch.zeros(nk * 27 * 2, dtype=torch.uint8, device=dev) if subjunc else None
            for r in range(world):
                if subjunc:
                    k_junc.zero_(); k_bm.zero_()
                torch.cuda.synchronize()
                m.vote_chunk_device(params, texts[r].data_ptr(), texts[r].numel(), k_off.data_ptr(), k_len.data_ptr(), K, paired, RL, k_out.data_ptr(),
                                    out_is_zero=True, d_junc=k_junc.data_ptr() if subjunc else None, d_big_margin=k_bm.data_ptr() if subjunc else None)
                crc0 = zlib.crc32(k_out.cpu().numpy().tobytes())
                v = allv[r].cpu().numpy()
                st = None if v[1] < 0 else v[1] / 1e6
                rank_rows.append({"rank": r, "index": "built" if r == 0 else "imported over NVLink (CUDA IPC)", "mapped_fraction": v[2] / 1e6,
                                  "slot0_is_true_position": st, "crc_equals_rank0_recompute": bool(int(v[0]) == crc0),
                                  "ok": bool(int(v[0]) == crc0 and (st is None or st >= (0.999 if not w.get("repeats") else 0.9)))})
        barrier()

    # ---- gather roofline probe: random 64-byte lines per second over a table of the index's size ----
    gather_gbs = None
    if rank == 0:
        try:
            gather_gbs = m.gather_bandwidth(max(int(desc.device_bytes), 1 << 28), 3)
        except Exception:
            gather_gbs = None

    # ---- cpu baseline + algorithmic bytes (rank 0; the one place oracle/ is executed) -------
    cpu_baseline, alg_bytes, whole, lookups_per_frag, output_writer = None, None, None, None, None
    if rank == 0:
        try:
            from oracle import oracle
            k = min(20000 if w["n_chr"] > 200 else 100000, R)
            oi = oracle.OracleIndex.from_arrays(index_arrays)
            ch = ReadChunk(d_text[: k * mates * RL].cpu().numpy(), np.arange(k * mates + 1, dtype=np.uint64) * RL, np.full(k * mates, RL, np.uint16), k, paired, [])
            _, _, st = oracle.vote_chunk(oi, params, ch, threads=min(16, host_threads()), junc_state={} if subjunc else None)
            alg_bytes = st.algorithmic_bytes() / k          # per fragment
            lookups_per_frag = st.lookups / k
            oi.close()
        except Exception as e:
            alg_bytes = None
            cpu_baseline = {"error": "oracle: " + repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            tmp = tempfile.mkdtemp(prefix="srbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") and w["n_chr"] <= 200 else None)
            try:
                sample = args.ref_sample or (1_000_000 if w["n_chr"] <= 200 else 200_000)
                sample = max(10_000, sample // mates)
                if os.path.exists(REF_BIN) or os.path.exists(REF_DUMP):
                    cpu_baseline = reference_measure(g, index_arrays, name, sample, 1, 1, tmp)
                if name == "cfg2_se":
                    whole = whole_program(g, index_arrays, name, sample, tmp)
            except Exception as e:  # the baseline is reporting, never the product path
                cpu_baseline = {"error": repr(e)}
            try:
                if name == "cfg2_se":
                    output_writer = output_writer_measure(tmp)
            except Exception as e:
                output_writer = {"error": repr(e)}
            finally:
                shutil.rmtree(tmp, ignore_errors=True)

    if rank == 0:
        total_frags = R * world
        value = total_frags * mates / (res["ms_step"] * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        roofline = None
        if alg_bytes:
            achieved = alg_bytes * R / (res["kern_ms"] * 1e-3) / 1e9
            traffic, traffic_note = None, "no capture"
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                ent = tr.get("workloads", {}).get(name)
                if ent and tr.get("kernel_source_hash") == kernel_source_hash():
                    traffic, traffic_note = float(ent["dram_bytes_per_fragment"]) * R, ent.get("source", "ncu --set full")
                else:
                    traffic_note = "profiles/traffic.json was captured from another version of the kernel sources (or not for this workload): not quoted"
            except Exception:
                pass
            lines_per_s = res["gathers"] / (res["kern_ms"] * 1e-3) if res["gathers"] else None
            line_rate = gather_gbs * 1e9 / 32.0 if gather_gbs else None
            # the same for the kernel's whole access mix: per fragment `small` probes of the L2-resident Bloom words (sparse tables;
            # one per lookup) and `big` random fetches that miss L2, issued by a kernel that does nothing else
            mix = None
            if res["gathers"] and lookups_per_frag:
                try:
                    sparse, phase = not (desc.accelerators & 4), bool(desc.accelerators & 16)
                    # small probes: one Bloom word per lookup, or -- with the phase table of a gapped index -- one occupancy bit per subread
                    n_small = (int(round(lookups_per_frag / 3)) if phase else int(round(lookups_per_frag))) if sparse else 0
                    small_bytes = (1 << 25) if phase else 2 * desc.buckets_number
                    n_big = max(1, int(round(res["gathers"] / R)))
                    units = m.lookup_roofline(small_bytes, n_small, max(int(desc.device_bytes), 1 << 28), n_big, 3)
                    mix = {"bloom_probes_per_fragment": n_small, "dram_fetches_per_fragment": n_big, "fragments_per_s_allowed": units,
                           "frac": (R / (res["kern_ms"] * 1e-3)) / units,
                           "probe": "k_mix_probe: that many random probes of an L2-resident table (evict-last, %.0f MB) and random 64-byte fetches (%.1f GB table) "
                                    "per unit, nothing else" % (small_bytes / 1e6, max(int(desc.device_bytes), 1 << 28) / 1e9)}
                except Exception as e:
                    mix = {"error": repr(e)}
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                        "traffic_source": traffic_note,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650",
                        "kernel": ("k_map_simple_pe" if paired else "k_map_simple") + f"<3,{(4 if desc.accelerators & 16 else 2) if desc.accelerators & 4 else (3 if desc.accelerators & 16 else 0)}> (one thread per fragment; mode 3 / 4 = phase units of a sparse / dense table) + k_map_warp (queued fragments)",
                        "kernel_ms": res["kern_ms"], "tier0_ms": res["simple_ms"], "algorithmic_bytes_per_fragment": alg_bytes,
                        "tier0_punted_fraction": res["punted"] / R, "warp_kernel_ms": res["kern_ms"] - res["simple_ms"],
                        "scalar_fragments_per_step": res["n_over"], "scalar_kernel_ms": res["over_ms"],
                        # the line-rate ("HBM-gather") roofline: a random fetch that misses L2 costs one DRAM line whatever is taken from
                        # it, so the bound is lines per second, measured with the lookup's own load over a table of the index's size
                        "gather": {"requests_per_fragment": res["gathers"] / R if res["gathers"] else None, "requests_per_s": lines_per_s,
                                   "peak_lines_per_s": line_rate, "frac": (lines_per_s / line_rate) if lines_per_s and line_rate else None,
                                   "counted": "by the kernel (srcuda_timing.n_gather_requests): bucket records, direct-filter words, direct_first "
                                              "entries, positions; the L2-resident Bloom words are not counted",
                                   "probe": "k_gather_probe: random 32-byte records through ld_record (64-byte DRAM fetch) over %.1f GB" % (max(int(desc.device_bytes), 1 << 28) / 1e9),
                                   "mix": mix}}
        out = {
            "metric": "reads/sec mapped", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": dict(workload_config(name, R, world), index="%d items, %d buckets, %.2f GB per GPU with its accelerators" %
                           (desc.n_items, desc.buckets_number, desc.device_bytes / 1e9)),
            "wall_ms_per_step": res["wall_ms"],
            "e2e": ({"value": total_frags * mates / (res["e2e_wall_ms"] * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"],
                     "ms_per_step_wall": res["e2e_wall_ms"], "ms_per_step_events": res["e2e_ev_ms"], "matches_device_result": res["e2e_same"]}
                    if "e2e_wall_ms" in res else None),
            "gpu_launches": int(res["launches"] + res.get("e2e_launches", 0) * args.steps),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "whole_program": whole, "output_writer": output_writer, "verify": res.get("verify"), "clocks": clk,
            "mapped_fraction": res["mapped"], "clean_fraction": res["clean_fraction"], "slot0_is_true_position": res["slot0_true"],
            "ranks": rank_rows, "sweep": sweep_rows,
            "index_replication_s": t_rep if world > 1 else None,
        }
        print(json.dumps(out))
    m.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
