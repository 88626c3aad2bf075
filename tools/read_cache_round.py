"""Whole-program wall time, default (unordered BAM) output: stock subread-align, subread-align-cuda, subread-align-cuda with the
passes after voting reading their chunk from memory (integration/read_cache.c, SRCUDA_READ_CACHE=1).
usage: python tools/read_cache_round.py [n_reads] [threads] [cuda|out] [contigs] [runs]   -> one JSON line"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen_data  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
    threads = sys.argv[2] if len(sys.argv) > 2 else "16"
    bound = os.path.join(REF, "subread-align-" + (sys.argv[3] if len(sys.argv) > 3 else "cuda"))
    contigs = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    runs = int(sys.argv[5]) if len(sys.argv) > 5 else 2
    d = tempfile.mkdtemp(prefix="rcache_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fa = os.path.join(d, "g.fa")
    gen_data.write_genome(fa, contigs, 500000, seed=11, repeats=0)
    subprocess.run([os.path.join(REF, "bin", "subread-buildindex"), "-B", "-o", os.path.join(d, "idx"), fa], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    gen_data.write_reads(fa, os.path.join(d, "r"), n_reads, 150, 9, False, 0.005, 0.02, True, 0.0)
    out = {"reads": n_reads, "threads": int(threads), "mode": "BAM (default output)", "genome_mbp": contigs * 0.5}
    files = {}
    for tag, exe, cache in (("stock", os.path.join(REF, "bin", "subread-align"), "0"), ("bound", bound, "0"), ("bound_read_cache", bound, "1")):
        best = None
        for _ in range(runs):
            t = time.time()
            p = subprocess.run([exe, "-t", "1", "-T", threads, "-i", os.path.join(d, "idx"), "-r", os.path.join(d, "r_R1.fastq"),
                                "-o", os.path.join(d, tag + ".bam")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                               env=dict(os.environ, SRCUDA_READ_CACHE=cache))
            dt = time.time() - t
            if p.returncode != 0:
                out[tag + "_error"] = p.stdout.decode()[-300:]
                break
            best = dt if best is None else min(best, dt)
        if best is not None:
            out[tag + "_s"] = round(best, 3)
            files[tag] = os.path.join(d, tag + ".bam")
    if len(files) == 3:
        from test_output_program import bam_records
        recs = {k: sorted(bam_records(open(v, "rb").read())[1]) for k, v in files.items()}
        out["same_records"] = recs["stock"] == recs["bound"] == recs["bound_read_cache"]
        out["records"] = len(recs["stock"])
    print(json.dumps(out))
    subprocess.run(["rm", "-rf", d])


if __name__ == "__main__":
    main()
