"""Whole-program wall time with ordered BAM output (--keepReadOrder): stock subread-align, subread-align-cuda with the reference's
writer, subread-align-cuda with the library's writer (SRCUDA_NATIVE_OUTPUT=1).  The reference compresses ordered output inline
under its output lock (core.c:1857-1866); the library's writer deflates on a team of threads and keeps the order.
usage: python tools/ordered_bam_round.py [n_reads] [threads] [cuda|out]   -> one JSON line"""
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
    d = tempfile.mkdtemp(prefix="ordered_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fa = os.path.join(d, "g.fa")
    gen_data.write_genome(fa, 40, 500000, seed=11, repeats=0)
    subprocess.run([os.path.join(REF, "bin", "subread-buildindex"), "-B", "-o", os.path.join(d, "idx"), fa], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    gen_data.write_reads(fa, os.path.join(d, "r"), n_reads, 150, 9, False, 0.005, 0.02, True, 0.0)
    out = {"reads": n_reads, "threads": int(threads), "mode": "BAM, --keepReadOrder"}
    files = {}
    for tag, exe, native in (("stock", os.path.join(REF, "bin", "subread-align"), "0"), ("bound_reference_writer", bound, "0"),
                             ("bound_library_writer", bound, "1")):
        best = None
        for _ in range(2):
            t = time.time()
            p = subprocess.run([exe, "-t", "1", "-T", threads, "-i", os.path.join(d, "idx"), "-r", os.path.join(d, "r_R1.fastq"), "--keepReadOrder",
                                "-o", os.path.join(d, tag + ".bam")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                               env=dict(os.environ, SRCUDA_NATIVE_OUTPUT=native))
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
        recs = {k: bam_records(open(v, "rb").read())[1] for k, v in files.items()}
        out["records_identical"] = recs["stock"] == recs["bound_reference_writer"] == recs["bound_library_writer"]
        out["records"] = len(recs["stock"])
    print(json.dumps(out))
    subprocess.run(["rm", "-rf", d])


if __name__ == "__main__":
    main()
