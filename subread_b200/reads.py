"""Read batching: the host-side analogue of the reference's ``fetch_next_read_pair`` input half
(reference src/core.c:1121-1175) and the FASTQ branch of ``geinput_next_read_trim``
(reference src/input-files.c:982-1093), producing the flat chunk layout the C-ABI takes:

    text : uint8[...]    read characters, read r = fragment*mates + mate
    off  : uint64[n+1]   byte offset of each read inside ``text``
    len  : uint16[n]     read lengths

Reads are kept exactly as they are in the file (no upper-casing: the reference does not
upper-case FASTQ either, input-files.c:1035); strand handling happens on the device.
"""
import gzip
from dataclasses import dataclass

import numpy as np


@dataclass
class ReadChunk:
    text: np.ndarray  # uint8
    off: np.ndarray  # uint64, n_reads + 1
    len: np.ndarray  # uint16, n_reads
    n_frags: int
    is_paired: bool
    names: list

    @property
    def mates(self) -> int:
        return 2 if self.is_paired else 1


def _open(path):
    return gzip.open(path, "rb") if str(path).endswith(".gz") else open(path, "rb")


def _fastq_records(path, limit=None):
    names, seqs = [], []
    with _open(path) as fh:
        while True:
            h = fh.readline()
            if not h:
                break
            if not h.strip():
                continue
            s = fh.readline().rstrip(b"\r\n")
            fh.readline()
            fh.readline()
            names.append(h[1:].split()[0].decode())
            seqs.append(s)
            if limit is not None and len(seqs) >= limit:
                break
    return names, seqs


def pack_reads(seqs1, seqs2=None, names=None) -> ReadChunk:
    """Flatten python byte strings into the chunk layout (mates interleaved per fragment)."""
    is_paired = seqs2 is not None
    if is_paired:
        assert len(seqs1) == len(seqs2)
        seqs = [s for pair in zip(seqs1, seqs2) for s in pair]
    else:
        seqs = list(seqs1)
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    text = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    return ReadChunk(text, off, lens.astype(np.uint16), len(seqs1), is_paired, names or [])


def load_fastq(path1, path2=None, limit=None) -> ReadChunk:
    names, s1 = _fastq_records(path1, limit)
    s2 = _fastq_records(path2, limit)[1] if path2 else None
    return pack_reads(s1, s2, names)


def load_fastq_indexed(path1, path2=None, trim_5=0, trim_3=0) -> ReadChunk:
    """The same chunk as ``load_fastq`` through the library's batch indexer (srcuda_fastq_split + srcuda_reads_pack,
    subread_b200/csrc/sr_fastq.cpp): one pass over the file's bytes, no per-read Python."""
    from .binding import fastq_split, reads_pack

    def one(path):
        with _open(path) as fh:
            buf = np.frombuffer(fh.read(), dtype=np.uint8)
        off, ln, _, ex = fastq_split(buf, True, trim_5, trim_3, want_names=True)
        return buf, off, ln, ex

    b1, o1, l1, ex = one(path1)
    names = [b1[int(a):int(a) + int(n)].tobytes().decode() for a, n in zip(ex["name_off"], ex["name_len"])]
    if path2:
        b2, o2, l2, _ = one(path2)
        text, off, ln = reads_pack(b1, o1, l1, b2, o2, l2)
    else:
        text, off, ln = reads_pack(b1, o1, l1)
    return ReadChunk(text, off, ln, int(o1.shape[0]), path2 is not None, names)


def synthetic_reads(genome: np.ndarray, n_frags: int, read_len: int, seed: int, is_paired=False,
                    insert=(200, 400), error_rate=0.0, indel_rate=0.0) -> ReadChunk:
    """Uniformly sampled reads from a uint8 ACGT genome array (one contig), both strands.

    Vectorised so that BASELINE-sized chunks (10^7 reads) are generated in seconds.  For paired
    reads R1 is the forward-strand mate and R2 the reverse complement of the fragment's far end,
    i.e. the FR orientation the reference expects by default.
    """
    rng = np.random.default_rng(seed)
    comp = np.zeros(256, dtype=np.uint8)
    comp[:] = ord("N")
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    glen = genome.shape[0]
    ar = np.arange(read_len, dtype=np.int64)

    def take(starts, rc_mask):
        m = genome[starts[:, None] + ar[None, :]]
        if rc_mask.any():
            m[rc_mask] = comp[m[rc_mask][:, ::-1]]
        return m

    def mutate(m):
        if error_rate > 0:
            hit = rng.random(m.shape) < error_rate
            sub = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(hit.sum()))]
            m[hit] = sub
        return m

    if not is_paired:
        starts = rng.integers(0, glen - read_len, size=n_frags)
        rc = rng.random(n_frags) < 0.5
        m = mutate(take(starts, rc))
        if indel_rate > 0:  # delete one base (shift left, refill from genome) in a subset of reads
            sel = np.flatnonzero(rng.random(n_frags) < indel_rate)
            for r in sel:
                p = int(rng.integers(20, read_len - 20))
                if rng.random() < 0.5:
                    m[r, p:-1] = m[r, p + 1:]
                    m[r, -1] = ord("ACGT"[int(rng.integers(0, 4))])
                else:
                    m[r, p + 1:] = m[r, p:-1].copy()
                    m[r, p] = ord("ACGT"[int(rng.integers(0, 4))])
        text = m.reshape(-1)
        lens = np.full(n_frags, read_len, dtype=np.uint16)
    else:
        ins = rng.integers(max(insert[0], read_len), insert[1] + 1, size=n_frags)
        starts = rng.integers(0, glen - insert[1] - 1, size=n_frags)
        flip = rng.random(n_frags) < 0.5
        left = take(starts, np.zeros(n_frags, bool))
        right = take(starts + ins - read_len, np.ones(n_frags, bool))
        r1 = np.where(flip[:, None], right, left)
        r2 = np.where(flip[:, None], left, right)
        m = mutate(np.stack([r1, r2], axis=1))
        text = m.reshape(-1)
        lens = np.full(n_frags * 2, read_len, dtype=np.uint16)
    n_reads = lens.shape[0]
    off = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len))
    return ReadChunk(np.ascontiguousarray(text), off, lens, n_frags, is_paired, [])
