#!/usr/bin/env python
"""Regenerates tests/golden/* from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (all small, committed):
    tiny.fa.gz                           seeded 2 x 30 kbp genome with planted repeats (tools/gen_data.py)
    tiny.00.b.tab.gz / .array.gz / tiny.reads   subread-buildindex -M 1 (one gapped block, 262 153 buckets)
    tinyF.*                              subread-buildindex -F -M 1 (full index, gap 1)
    se_R1.fastq.gz, pe_R1/2.fastq.gz     reads with sequencing errors, indels, Ns and mixed lengths
    expect_<case>.bin.gz                 the reference's bigtable after STEP_VOTING (oracle/ref_hooks.c),
                                         CORE_IS_GAPPED_READ masked out (host-side find_new_indels bit);
                                         subjunc cases (tinyS: planted introns) append the subjunc_result_t table
                                         and the big-margin records
    cases.json                           case table: index, reads, CLI flags and the matching parameters
"""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_data  # noqa: E402
from oracle import oracle  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")

CASES = [
    dict(name="se_gapped", index="tiny", r1="se_R1.fastq.gz", r2=None, flags=[], params={}),
    dict(name="pe_gapped", index="tiny", r1="pe_R1.fastq.gz", r2="pe_R2.fastq.gz", flags=[], params={}),
    dict(name="se_full", index="tinyF", r1="se_R1.fastq.gz", r2=None, flags=[], params={}),
    dict(name="pe_full", index="tinyF", r1="pe_R1.fastq.gz", r2="pe_R2.fastq.gz", flags=[], params={}),
    dict(name="se_I16", index="tiny", r1="se_R1.fastq.gz", r2=None, flags=["-I", "16"], params=dict(max_indel_length=16)),
    dict(name="sj_se", index="tinyS", r1="sj_se_R1.fastq.gz", r2=None, flags=[], params={}, subjunc=True),
    dict(name="sj_pe", index="tinyS", r1="sj_pe_R1.fastq.gz", r2="sj_pe_R2.fastq.gz", flags=[], params={}, subjunc=True),
    dict(name="pe_n14_ff", index="tiny", r1="pe_R1.fastq.gz", r2="pe_R2.fastq.gz", flags=["-n", "14", "-m", "2", "-S", "ff", "-D", "400"],
         params=dict(total_subreads=14, minimum_subread_for_first_read=2, is_second_read_reversed=0, maximum_pair_distance=400)),
]


def gz(src, dst):
    with open(src, "rb") as a, gzip.GzipFile(dst, "wb", compresslevel=9, mtime=0) as b:
        shutil.copyfileobj(a, b)


def main():
    tmp = tempfile.mkdtemp(prefix="srgolden_")
    fa = os.path.join(tmp, "tiny.fa")
    gen_data.write_genome(fa, 2, 30000, seed=101, repeats=60)
    gz(fa, os.path.join(HERE, "tiny.fa.gz"))
    # subjunc: the same genome with planted GT..AG / CT..AC introns, reads sampled from the spliced transcripts
    faS = os.path.join(tmp, "tinyS.fa")
    gen_data.write_spliced(fa, faS, os.path.join(tmp, "sj_se"), 1200, 100, 301, False, mixed_len=True)
    gen_data.write_spliced(fa, faS, os.path.join(tmp, "sj_pe"), 800, 100, 301, True)
    gz(faS, os.path.join(HERE, "tinyS.fa.gz"))
    for name, extra in (("tiny", []), ("tinyF", ["-F"]), ("tinyS", [])):
        subprocess.check_call([REF + "/bin/subread-buildindex", "-o", os.path.join(tmp, name), "-M", "1"] + extra + [faS if name == "tinyS" else fa],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert not os.path.exists(os.path.join(tmp, name + ".01.b.tab")), "expected a one-block index"
        gz(os.path.join(tmp, name + ".00.b.tab"), os.path.join(HERE, name + ".00.b.tab.gz"))
        gz(os.path.join(tmp, name + ".00.b.array"), os.path.join(HERE, name + ".00.b.array.gz"))
        shutil.copy(os.path.join(tmp, name + ".reads"), os.path.join(HERE, name + ".reads"))
    gen_data.write_reads(fa, os.path.join(tmp, "se"), 1500, 160, 202, False, 0.015, 0.25, True, 0.004)
    gen_data.write_reads(fa, os.path.join(tmp, "pe"), 1000, 120, 203, True, 0.015, 0.2, True, 0.004)
    for f in ("se_R1", "pe_R1", "pe_R2", "sj_se_R1", "sj_pe_R1", "sj_pe_R2"):
        gz(os.path.join(tmp, f + ".fastq"), os.path.join(HERE, f + ".fastq.gz"))
    for c in CASES:
        bt = os.path.join(tmp, "bt.bin")
        if os.path.exists(bt):
            os.remove(bt)
        if c.get("subjunc"):
            cmd = [REF + "/ref_dump_subjunc", "-T", "1", "-i", os.path.join(tmp, c["index"]), "-r", os.path.join(HERE, c["r1"]),
                   "--SAMoutput", "-o", os.path.join(tmp, "o.sam")] + c["flags"]
        else:
            cmd = [REF + "/ref_dump", "-t", "1", "-T", "1", "-i", os.path.join(tmp, c["index"]), "-r", os.path.join(HERE, c["r1"]),
                   "--SAMoutput", "-o", os.path.join(tmp, "o.sam")] + c["flags"]
        if c["r2"]:
            cmd += ["-R", os.path.join(HERE, c["r2"])]
        subprocess.check_call(cmd, env=dict(os.environ, SRDUMP_BIGTABLE=bt), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        chunks = oracle.read_bigtable_dump(bt)
        assert len(chunks) == 1
        a = chunks[0]["align"].copy()
        a["result_flags"] &= ~np.int16(64)  # CORE_IS_GAPPED_READ: set by host-side find_new_indels after voting
        with gzip.GzipFile(os.path.join(HERE, f"expect_{c['name']}.bin.gz"), "wb", compresslevel=9, mtime=0) as fh:
            fh.write(a.tobytes())
            if c.get("subjunc"):  # + subjunc_result_t table + the 9 live u16 of every big-margin record
                fh.write(chunks[0]["junc"].tobytes())
                fh.write(np.ascontiguousarray(chunks[0]["big_margin"][:, :, :9]).tobytes())
        print(c["name"], a.shape, "mapped", float((a[:, :, 0]["selected_votes"] > 0).mean()))
    with open(os.path.join(HERE, "cases.json"), "w") as fh:
        json.dump(CASES, fh, indent=1)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
