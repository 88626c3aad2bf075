#!/usr/bin/env python
"""Seeded synthetic genomes and reads (SURVEY.md section 8d shapes).

  gen_data.py genome OUT.fa --chr 200 --len 500000 --seed 1        # gen_long_chromosomes shape, 70 cols
  gen_data.py reads  GENOME.fa OUT_PREFIX --n 100000 --len 150 --seed 2 [--paired] [--err 0.01]
                     [--indel 0.05] [--mixed-len] [--with-n 0.01]

Reads are written as FASTQ (quality 'H', like genRandomReads without --qualityRefFile).  The
reference's own genRandomReads is used where BASELINE.json names it; this generator exists
because that tool cannot make sequencing errors, indels, Ns or mixed lengths, all of which the
parity tests need.
"""
import argparse
import gzip
import sys

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def write_genome(path, n_chr, chr_len, seed, cols=70, repeats=0):
    """``repeats`` > 0 plants that many copies (1 % diverged) of a few repeat families, short tandem
    repeats and poly-A runs per chromosome, so that equal-key runs, the 24-slot vote-row limit and
    the >100-occurrence index filter all get exercised (a uniform genome never does)."""
    rng = np.random.default_rng(seed)
    families = [ACGT[rng.integers(0, 4, size=int(n))] for n in (300, 700, 1500, 120, 60)]
    with open(path, "wb") as fh:
        for c in range(n_chr):
            fh.write(b">chr%d\n" % (c + 1))
            seq = ACGT[rng.integers(0, 4, size=chr_len)]
            for _ in range(repeats):
                fam = families[int(rng.integers(0, len(families)))].copy()
                hit = rng.random(fam.shape[0]) < 0.01
                fam[hit] = ACGT[rng.integers(0, 4, size=int(hit.sum()))]
                if fam.shape[0] + 1 < chr_len:
                    at = int(rng.integers(0, chr_len - fam.shape[0]))
                    seq[at:at + fam.shape[0]] = fam
            for _ in range(repeats // 8):
                unit = ACGT[rng.integers(0, 4, size=int(rng.integers(1, 7)))]
                n = int(rng.integers(40, 400))
                at = int(rng.integers(0, max(1, chr_len - n)))
                seq[at:at + n] = np.resize(unit, n)[: len(seq[at:at + n])]
            full = (chr_len // cols) * cols
            body = seq[:full].reshape(-1, cols)
            lines = np.concatenate([body, np.full((body.shape[0], 1), 10, np.uint8)], axis=1)
            fh.write(lines.tobytes())
            if full < chr_len:
                fh.write(seq[full:].tobytes() + b"\n")


def read_fasta(path):
    names, seqs, cur = [], [], []
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as fh:
        for line in fh:
            if line.startswith(b">"):
                if cur:
                    seqs.append(b"".join(cur))
                    cur = []
                names.append(line[1:].split()[0].decode())
            else:
                cur.append(line.strip())
    if cur:
        seqs.append(b"".join(cur))
    return names, [np.frombuffer(s.upper(), dtype=np.uint8) for s in seqs]


COMP = np.full(256, ord("N"), dtype=np.uint8)
for _a, _b in zip(b"ACGTacgt", b"TGCAtgca"):
    COMP[_a] = _b


def revcomp(a):
    return COMP[a[::-1]]


def mutate(seq, rng, err, indel_p, with_n):
    seq = seq.copy()
    if err > 0:
        hit = rng.random(seq.shape[0]) < err
        seq[hit] = ACGT[rng.integers(0, 4, size=int(hit.sum()))]
    if indel_p > 0 and rng.random() < indel_p and seq.shape[0] > 60:
        p = int(rng.integers(20, seq.shape[0] - 20))
        k = int(rng.integers(1, 5))
        if rng.random() < 0.5:  # deletion from the read (= bases missing)
            seq = np.concatenate([seq[:p], seq[p + k:], ACGT[rng.integers(0, 4, size=k)]])
        else:  # insertion
            seq = np.concatenate([seq[:p], ACGT[rng.integers(0, 4, size=k)], seq[p:-k]])
    if with_n > 0:
        hit = rng.random(seq.shape[0]) < with_n
        seq[hit] = ord("N")
    return seq


def write_reads(genome_fa, prefix, n, rlen, seed, paired, err, indel_p, mixed_len, with_n, gz=False):
    names, seqs = read_fasta(genome_fa)
    rng = np.random.default_rng(seed)
    lens = np.array([s.shape[0] for s in seqs])
    opener = (lambda p: gzip.open(p + ".gz", "wb", compresslevel=1)) if gz else (lambda p: open(p, "wb"))
    f1 = opener(prefix + "_R1.fastq")
    f2 = opener(prefix + "_R2.fastq") if paired else None
    for i in range(n):
        L1 = int(rng.integers(max(20, rlen // 2), rlen + 1)) if mixed_len else rlen
        L2 = int(rng.integers(max(20, rlen // 2), rlen + 1)) if mixed_len else rlen
        while True:
            c = int(rng.integers(0, len(seqs)))
            ins = int(rng.integers(max(L1, L2), max(L1, L2) + 350)) if paired else L1
            if lens[c] > ins + 2:
                break
        start = int(rng.integers(0, lens[c] - ins))
        frag = seqs[c][start:start + ins]
        flip = rng.random() < 0.5
        if paired:
            a, b = frag[:L1], revcomp(frag[-L2:])
            if flip:
                a, b = revcomp(frag[-L1:]), frag[:L2]
            a = mutate(a, rng, err, indel_p, with_n)
            b = mutate(b, rng, err, indel_p, with_n)
            f1.write(b"@R%09d:%s:%d/1\n%s\n+\n%s\n" % (i, names[c].encode(), start + 1, a.tobytes(), b"H" * a.shape[0]))
            f2.write(b"@R%09d:%s:%d/2\n%s\n+\n%s\n" % (i, names[c].encode(), start + 1, b.tobytes(), b"H" * b.shape[0]))
        else:
            a = revcomp(frag) if flip else frag
            a = mutate(a, rng, err, indel_p, with_n)
            f1.write(b"@R%09d:%s:%d\n%s\n+\n%s\n" % (i, names[c].encode(), start + 1, a.tobytes(), b"H" * a.shape[0]))
    f1.close()
    if f2:
        f2.close()


def write_spliced(genome_fa, out_genome_fa, prefix, n, rlen, seed, paired, err=0.005, mixed_len=False):
    """RNA-seq-like input for subjunc: plants introns (GT..AG on the forward strand, CT..AC for genes on
    the reverse strand) into the genome, writes the edited genome, and samples reads from the spliced
    transcripts (introns removed), so that many reads span an exon-exon junction."""
    names, seqs = read_fasta(genome_fa)
    rng = np.random.default_rng(seed)
    seqs = [s.copy() for s in seqs]
    transcripts = []
    for s in seqs:
        pos, keep = 2000, []
        last = 0
        while pos + 4000 < s.shape[0]:
            ilen = int(rng.integers(80, 3000))
            if rng.random() < 0.5:
                s[pos:pos + 2] = np.frombuffer(b"GT", np.uint8); s[pos + ilen - 2:pos + ilen] = np.frombuffer(b"AG", np.uint8)
            else:
                s[pos:pos + 2] = np.frombuffer(b"CT", np.uint8); s[pos + ilen - 2:pos + ilen] = np.frombuffer(b"AC", np.uint8)
            keep.append((last, pos))
            last = pos + ilen
            pos = last + int(rng.integers(120, 1500))
        keep.append((last, s.shape[0]))
        transcripts.append(np.concatenate([s[a:b] for a, b in keep]))
    with open(out_genome_fa, "wb") as fh:
        for nm, s in zip(names, seqs):
            fh.write(b">" + nm.encode() + b"\n")
            for i in range(0, s.shape[0], 70):
                fh.write(s[i:i + 70].tobytes() + b"\n")
    tmp_fa = out_genome_fa + ".transcripts.fa"
    with open(tmp_fa, "wb") as fh:
        for nm, s in zip(names, transcripts):
            fh.write(b">" + nm.encode() + b"\n")
            for i in range(0, s.shape[0], 70):
                fh.write(s[i:i + 70].tobytes() + b"\n")
    write_reads(tmp_fa, prefix, n, rlen, seed + 1, paired, err, 0.0, mixed_len, 0.0)


def main(argv=None):
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    g = sub.add_parser("genome")
    g.add_argument("out")
    g.add_argument("--chr", type=int, default=1)
    g.add_argument("--len", type=int, default=500000)
    g.add_argument("--seed", type=int, default=1)
    g.add_argument("--repeats", type=int, default=0)
    r = sub.add_parser("reads")
    r.add_argument("genome")
    r.add_argument("prefix")
    r.add_argument("--n", type=int, default=10000)
    r.add_argument("--len", type=int, default=100)
    r.add_argument("--seed", type=int, default=2)
    r.add_argument("--paired", action="store_true")
    r.add_argument("--err", type=float, default=0.0)
    r.add_argument("--indel", type=float, default=0.0)
    r.add_argument("--mixed-len", action="store_true")
    r.add_argument("--with-n", type=float, default=0.0)
    r.add_argument("--gz", action="store_true")
    a = ap.parse_args(argv)
    if a.cmd == "genome":
        write_genome(a.out, a.chr, a.len, a.seed, repeats=a.repeats)
    else:
        write_reads(a.genome, a.prefix, a.n, a.len, a.seed, a.paired, a.err, a.indel, a.mixed_len, a.with_n, a.gz)


if __name__ == "__main__":
    sys.exit(main())
