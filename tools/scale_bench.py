#!/usr/bin/env python
"""tools/scale_bench.py -- the voting step on the other BASELINE.json shapes (configs[2..4]): human-sized index,
paired-end reads, subjunc mode, read-length sweep.  One JSON line per run; not the bench.py headline (configs[1]).

    python tools/scale_bench.py --genome-mbp 3100 --reads 2000000 --read-len 150 --paired
    python tools/scale_bench.py --genome-mbp 100 --read-len 100 --paired --subjunc
    python tools/scale_bench.py --sweep 50,75,100,125,150,200,250

Everything is built on the device: a seeded uniform genome of <genome-mbp> Mbp in 500 kbp contigs
(gen_long_chromosomes shape), its one-block gapped (or --full) index (subread_b200.index_build.build_index_arrays_device,
the reference's format), error-free reads sampled from both strands (PE: FR pairs, insert 200-400).  Timing is the
HBM-resident srcuda_vote_chunk_device call with CUDA events, W warm-up and K timed steps.  --check N maps the first N
fragments on the CPU oracle as well (tests' checker, from the same in-memory index) and requires byte equality under
the appendix-A rule; --truth requires that error-free reads report their true position in slot 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-mbp", type=float, default=100.0)
    ap.add_argument("--contig-bp", type=int, default=500000)
    ap.add_argument("--reads", type=int, default=2_000_000, help="fragments per step")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--sweep", default="", help="comma-separated read lengths; one line each")
    ap.add_argument("--paired", action="store_true")
    ap.add_argument("--subjunc", action="store_true")
    ap.add_argument("--full", action="store_true", help="full (gap 1) index")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--check", type=int, default=20000, help="fragments compared with the oracle (0 = none)")
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--error-rate", type=float, default=0.0, help="per-base substitution rate of the reads (0 = error-free)")
    ap.add_argument("--second-pass", action="store_true",
                    help="also time the chunk merged into the bigtable its first pass left (what a later block of a split index costs)")
    ap.add_argument("--modes", default="", help="comma-separated list of se, pe, sj-se, sj-pe run on the same index (overrides --paired/--subjunc)")
    args = ap.parse_args()

    import torch
    from subread_b200 import MAPPING_RESULT, VoteParams
    from subread_b200.binding import Mapper
    from subread_b200.index_build import PADDING, build_index_arrays_device
    from subread_b200.layout import used_slots_equal
    from subread_b200.reads import ReadChunk

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    n_chr = max(1, int(round(args.genome_mbp * 1e6 / args.contig_bp)))
    L = args.contig_bp
    gen = torch.Generator(device=dev)
    gen.manual_seed(args.seed)
    t0 = time.perf_counter()
    codes = torch.randint(0, 4, (n_chr * L,), generator=gen, device=dev, dtype=torch.uint8)   # A0 G1 C2 T3
    names = [f"chr{c + 1}" for c in range(n_chr)]
    a = build_index_arrays_device(names, codes, [L] * n_chr, gapped=not args.full)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    m = Mapper(0)
    t0 = time.perf_counter()
    m.index_upload(a["index_gap"], a["padding"], a["bucket_start"], a["keys"], a["vals"], a["start_point"], a["length"],
                   a["values"], a["read_offsets"])
    t_upload = time.perf_counter() - t0
    desc = m.describe()
    torch.cuda.empty_cache()

    letters = torch.tensor(list(b"AGCT"), dtype=torch.uint8, device=dev)       # int2base, subread.h:248
    lin_first = torch.from_numpy(a["contig_starts"]).to(dev)
    modes = [x for x in args.modes.split(",") if x] or [("sj-" if args.subjunc else "") + ("pe" if args.paired else "se")]
    oidx = None
    for mode in modes:
      args.paired, args.subjunc = mode.endswith("pe"), mode.startswith("sj")
      mates = 2 if args.paired else 1
      params = VoteParams.defaults(args.paired, args.subjunc)
      mb = params.multi_best_reads
      oidx = run_mode(args, torch, dev, m, a, desc, codes, letters, lin_first, n_chr, L, mates, params, mb, t_build, t_upload, oidx)
    m.close()
    return 0


def run_mode(args, torch, dev, m, a, desc, codes, letters, lin_first, n_chr, L, mates, params, mb, t_build, t_upload, oidx):
    from subread_b200 import MAPPING_RESULT
    from subread_b200.layout import used_slots_equal
    from subread_b200.reads import ReadChunk

    def make_reads(R, RL, seed):
        """text [R*mates, RL] uint8 on the device + the true linear position of each mate's leftmost base."""
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        ar = torch.arange(RL, device=dev)
        text = torch.empty((R, mates, RL), dtype=torch.uint8, device=dev)
        truth = torch.empty((R, mates), dtype=torch.int64, device=dev)
        for lo in range(0, R, 1 << 20):
            hi = min(R, lo + (1 << 20))
            n = hi - lo
            c = torch.randint(0, n_chr, (n,), generator=g, device=dev)
            if args.paired:
                ins = torch.randint(max(200, RL), max(400, RL + 50), (n,), generator=g, device=dev)
                s = (torch.rand(n, generator=g, device=dev) * (L - ins).float()).long().clamp_(0, L - 450 - RL)
                flip = torch.rand(n, generator=g, device=dev) < 0.5
                left = codes[(c * L + s)[:, None] + ar[None, :]]
                right_s = s + ins - RL
                right = codes[(c * L + right_s)[:, None] + ar[None, :]]
                # FR pair: the mate at the fragment's left end reads forward, the other one is reverse-complemented;
                # `flip` decides which file (R1/R2) gets which end
                left_t = letters[left.long()]
                right_t = letters[(3 - right).flip(1).long()]
                r1 = torch.where(flip[:, None], right_t, left_t)
                r2 = torch.where(flip[:, None], left_t, right_t)
                text[lo:hi, 0], text[lo:hi, 1] = r1, r2
                pl, pr = lin_first[c] + s, lin_first[c] + right_s
                truth[lo:hi, 0] = torch.where(flip, pr, pl)
                truth[lo:hi, 1] = torch.where(flip, pl, pr)
            else:
                s = torch.randint(0, L - RL, (n,), generator=g, device=dev)
                rc = torch.rand(n, generator=g, device=dev) < 0.5
                fw = codes[(c * L + s)[:, None] + ar[None, :]]
                t = torch.where(rc[:, None], letters[(3 - fw).flip(1).long()], letters[fw.long()])
                text[lo:hi, 0] = t
                truth[lo:hi, 0] = lin_first[c] + s
        if args.error_rate > 0:      # substitutions: a wrong base is the right one's code + 1..3 (mod 4) in "AGCT"
            flat = text.reshape(-1)
            for lo in range(0, flat.numel(), 1 << 28):
                part = flat[lo:lo + (1 << 28)]
                hit = torch.rand(part.numel(), generator=g, device=dev) < args.error_rate
                shift = torch.randint(1, 4, (part.numel(),), generator=g, device=dev, dtype=torch.uint8)
                code = (part == ord("G")).to(torch.uint8) + 2 * (part == ord("C")).to(torch.uint8) + 3 * (part == ord("T")).to(torch.uint8)
                part[:] = torch.where(hit, letters[((code + shift) & 3).long()], part)
        return text.reshape(-1), truth

    lens = [int(x) for x in args.sweep.split(",")] if args.sweep else [args.read_len]
    for RL in lens:
        R = args.reads
        d_text, truth = make_reads(R, RL, 4242 + RL)
        d_off = torch.arange(R * mates, device=dev, dtype=torch.int64) * RL
        d_len = torch.full((R * mates,), RL, dtype=torch.int16, device=dev)
        d_out = torch.zeros(R * mates * mb * 68, dtype=torch.uint8, device=dev)
        d_junc = d_bm = None
        if args.subjunc:
            d_junc = torch.zeros(R * mates * mb * 16, dtype=torch.uint8, device=dev)
            d_bm = torch.zeros(R * mates * 27 * 2, dtype=torch.uint8, device=dev)

        def step():
            if args.subjunc or args.paired:
                d_out.zero_()
                if d_junc is not None:
                    d_junc.zero_(); d_bm.zero_()
                torch.cuda.synchronize()     # the library launches on its own streams
            m.vote_chunk_device(params, d_text.data_ptr(), d_text.numel(), d_off.data_ptr(), d_len.data_ptr(), R, args.paired, RL, d_out.data_ptr(),
                                out_is_zero=True, d_junc=d_junc.data_ptr() if d_junc is not None else None,
                                d_big_margin=d_bm.data_ptr() if d_bm is not None else None)
            return m.timing()

        torch.cuda.synchronize()         # the inputs were written on torch's stream; the library launches on its own
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        ms = kern = simple = over = 0.0
        punted = n_over = launches = 0
        for _ in range(args.steps):
            t = step()
            ms += t.ms_total; kern += t.ms_kernel_vote; simple += t.ms_kernel_simple; over += t.ms_kernel_overflow
            punted += t.n_simple_punted; n_over += t.n_overflow; launches += t.n_launches
        torch.cuda.synchronize()
        ms /= args.steps
        out = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=MAPPING_RESULT).reshape(R, mates, mb)
        mapped = float((out[:, :, 0]["selected_votes"] > 0).mean())
        # size-independent property: an error-free read from a random genome reports its own position
        tr = truth.cpu().numpy()
        # (PE: both mates are voted in fragment orientation -- R2 is reverse-complemented at fetch, core.c:1194 -- so each
        # mate reports the linear position of its leftmost genome base)
        correct = float((out[:, :, 0]["selected_position"].astype(np.int64) == tr).mean())
        line = {"tool": "scale_bench", "workload": f"{args.genome_mbp:g} Mbp synthetic genome ({n_chr}x{L} bp), "
                f"{'full' if args.full else 'gapped'} one-block index ({desc.n_items} items, {desc.buckets_number} buckets, "
                f"{desc.n_items / desc.buckets_number:.1f} per bucket), {RL} bp {'PE' if args.paired else 'SE'} reads"
                f"{', subjunc mode' if args.subjunc else ''}{', %.3g substitutions per base' % args.error_rate if args.error_rate else ''}",
                "fragments_per_step": R, "reads_per_s": R * mates / (ms * 1e-3), "fragments_per_s": R / (ms * 1e-3), "ms_per_step": ms,
                "kernel_ms": kern / args.steps, "tier0_ms": simple / args.steps, "tier0_punted_fraction": punted / args.steps / R,
                "overflow_fraction": n_over / args.steps / R, "overflow_ms": over / args.steps, "gpu_launches": launches,
                "mapped_fraction": mapped, "slot0_is_true_position": correct, "index_device_gb": desc.device_bytes / 1e9,
                "index_build_s": t_build, "index_upload_s": t_upload, "steps": args.steps, "warmup": args.warmup}
        if args.check > 0:
            from oracle import oracle
            if oidx is None:
                oidx = oracle.OracleIndex.from_arrays(a)
            k = min(args.check, R)
            h_text = d_text[: k * mates * RL].cpu().numpy()
            ch = ReadChunk(h_text, np.arange(k * mates + 1, dtype=np.uint64) * RL, np.full(k * mates, RL, np.uint16), k, args.paired, [])
            js = {} if args.subjunc else None
            t0 = time.perf_counter()
            exp, _, st = oracle.vote_chunk(oidx, params, ch, threads=min(16, os.cpu_count() or 1), junc_state=js)
            t_or = time.perf_counter() - t0
            ok, bad = used_slots_equal(out[:k], exp)
            line.update({"oracle_checked_fragments": k, "oracle_equal": bool(ok), "oracle_first_bad": None if ok else int(bad),
                         "oracle_reads_per_s": k * mates / t_or, "algorithmic_bytes_per_fragment": st.algorithmic_bytes() / k})
            if args.subjunc:
                from subread_b200.layout import SUBJUNC_RESULT
                gj = np.frombuffer(d_junc[: k * mates * mb * 16].cpu().numpy().tobytes(), dtype=SUBJUNC_RESULT)
                gb = np.frombuffer(d_bm[: k * mates * 54].cpu().numpy().tobytes(), dtype=np.uint16)
                used = exp.reshape(-1)["selected_votes"] > 0
                line["oracle_junc_equal"] = bool(np.array_equal(gj[used], js["junc"][used]) and np.array_equal(gb, js["big_margin"]))
        print(json.dumps(line), flush=True)
        if args.second_pass and not args.subjunc:
            # A later index block merges into a non-zero bigtable (core.c:3567-3613).  The same table stands in for the next
            # block: every step starts from the first pass's result (restored outside the library's timed region).
            first = d_out.clone()

            def step2():
                d_out.copy_(first)
                torch.cuda.synchronize()
                m.vote_chunk_device(params, d_text.data_ptr(), d_text.numel(), d_off.data_ptr(), d_len.data_ptr(), R, args.paired, RL,
                                    d_out.data_ptr(), out_is_zero=False)
                return m.timing()

            for _ in range(args.warmup):
                step2()
            ms2 = simple2 = 0.0
            punted2 = 0
            for _ in range(args.steps):
                t = step2()
                ms2 += t.ms_total; simple2 += t.ms_kernel_simple; punted2 += t.n_simple_punted
            torch.cuda.synchronize()
            ms2 /= args.steps
            out2 = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=MAPPING_RESULT).reshape(R, mates, mb)
            line2 = {"tool": "scale_bench", "workload": line["workload"] + ", second pass over the first pass's bigtable (read-modify-write)",
                     "rmw_tier": os.environ.get("SRCUDA_TIER_RMW", "1") != "0", "fragments_per_step": R, "reads_per_s": R * mates / (ms2 * 1e-3),
                     "ms_per_step": ms2, "tier0_ms": simple2 / args.steps, "tier0_punted_fraction": punted2 / args.steps / R,
                     "slot0_unchanged": bool(np.array_equal(out2[:, :, 0]["selected_position"], out[:, :, 0]["selected_position"])),
                     "steps": args.steps, "warmup": args.warmup}
            if args.check > 0:
                exp2, _, _ = oracle.vote_chunk(oidx, params, ch, threads=min(16, os.cpu_count() or 1), out_align=exp.copy())
                ok2, bad2 = used_slots_equal(out2[:k], exp2)
                line2.update({"oracle_checked_fragments": k, "oracle_equal": bool(ok2), "oracle_first_bad": None if ok2 else int(bad2)})
            print(json.dumps(line2), flush=True)
            del first
        del d_text, d_off, d_len, d_out, truth
        torch.cuda.empty_cache()
    return oidx


if __name__ == "__main__":
    sys.exit(main())
