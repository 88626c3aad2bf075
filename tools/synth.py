"""tools/synth.py -- seeded synthetic genomes and read sets of the BASELINE.json shapes, generated with torch on whatever
device is at hand (plumbing; bench.py, tools/scale_bench.py and the full-size GPU tests share it).

Genome: `n_chr` contigs of `chr_len` uniform i.i.d. bases (gen_long_chromosomes shape), held as 2-bit codes A0 G1 C2 T3
(base2int, reference src/subread.h:238) in one uint8 tensor.  Optional edits before indexing:
  * plant_repeats   -- families of exact copies (reads drawn from them map to several places)
  * plant_introns   -- GT..AG introns at known places, so that spliced reads have a donor/acceptor pair to find (subjunc)
Reads: error-free by default, both strands; paired-end as FR pairs (insert 200-400); `noise` adds per-base substitutions, Ns
and one short indel per affected read (a wgsim-like mix); `spliced_fraction` makes fragments whose first mate spans a planted
intron.  Every generator returns the text of the reads (ASCII, reads back to back, mates interleaved per fragment), the true
linear position (index coordinates) of each mate's leftmost base and a per-fragment `clean` flag (no noise, not from a repeat,
not spliced): a clean read of a random genome must come back with its own position in slot 0.
"""
import numpy as np

LETTERS = b"AGCT"  # int2base, subread.h:248


class Genome:
    def __init__(self, n_chr, chr_len, seed, device):
        import torch
        self.torch = torch
        self.dev = torch.device(device)
        self.n_chr, self.chr_len = int(n_chr), int(chr_len)
        g = torch.Generator(device=self.dev)
        g.manual_seed(int(seed))
        self.gen = g
        self.codes = torch.randint(0, 4, (self.n_chr * self.chr_len,), generator=g, device=self.dev, dtype=torch.uint8)
        self.names = [f"chr{c + 1}" for c in range(self.n_chr)]
        self.repeat_starts = None      # concatenated-genome offsets of the planted repeat copies
        self.repeat_len = 0
        self.intron_starts = None      # per contig: offsets (inside the contig) of the planted introns
        self.intron_len = 0
        self.lin_first = None          # linear (index) position of each contig's first base, set by build_index

    # ---- edits ---------------------------------------------------------------------------------
    def plant_repeats(self, families, copies, length):
        torch = self.torch
        total = families * copies
        # one copy per slot of a regular grid, so that copies never overlap and never cross a contig end
        per_chr = max(1, self.chr_len // (2 * length) - 1)
        slots = torch.randperm(self.n_chr * per_chr, generator=self.gen, device=self.dev)[:total]
        starts = (slots // per_chr) * self.chr_len + (slots % per_chr) * 2 * length + length // 2
        starts = starts.view(families, copies)
        ar = torch.arange(length, device=self.dev)
        src = self.codes[starts[:, 0:1] + ar[None, :]]                      # [families, length]
        for c in range(1, copies):
            self.codes[(starts[:, c:c + 1] + ar[None, :]).reshape(-1)] = src.reshape(-1)
        self.repeat_starts, self.repeat_len = starts.reshape(-1), int(length)

    def plant_introns(self, first=20000, every=50000, length=3000):
        torch = self.torch
        k = torch.arange(first, self.chr_len - 2 * length - 1000, every, device=self.dev)
        self.intron_starts, self.intron_len = k, int(length)
        base = (torch.arange(self.n_chr, device=self.dev) * self.chr_len)[:, None] + k[None, :]
        base = base.reshape(-1)
        self.codes[base] = 1                       # G
        self.codes[base + 1] = 3                   # T
        self.codes[base + length - 2] = 0          # A
        self.codes[base + length - 1] = 1          # G

    # ---- index ---------------------------------------------------------------------------------
    def build_index(self, gapped=True):
        from subread_b200.index_build import build_index_arrays_device
        a = build_index_arrays_device(self.names, self.codes, [self.chr_len] * self.n_chr, gapped=gapped)
        self.lin_first = self.torch.from_numpy(a["contig_starts"]).to(self.dev)
        return a

    # ---- reads ---------------------------------------------------------------------------------
    def reads(self, n_frags, read_len, seed, paired=False, sub_rate=0.0, n_rate=0.0, indel_read_rate=0.0, repeat_fraction=0.0,
              spliced_fraction=0.0, piece=1 << 20):
        """-> (text uint8 [n_frags * mates * read_len], truth int64 [n_frags, mates], clean bool [n_frags]) on the device."""
        torch = self.torch
        dev, L, RL = self.dev, self.chr_len, int(read_len)
        mates = 2 if paired else 1
        g = torch.Generator(device=dev)
        g.manual_seed(int(seed))
        letters = torch.tensor(list(LETTERS), dtype=torch.uint8, device=dev)
        ar = torch.arange(RL, device=dev)
        text = torch.empty((n_frags, mates, RL), dtype=torch.uint8, device=dev)
        truth = torch.empty((n_frags, mates), dtype=torch.int64, device=dev)
        clean = torch.ones(n_frags, dtype=torch.bool, device=dev)
        span = max(450, RL + 50) if paired else RL
        for lo in range(0, n_frags, piece):
            hi = min(n_frags, lo + piece)
            n = hi - lo
            c = torch.randint(0, self.n_chr, (n,), generator=g, device=dev)
            s = torch.randint(0, L - span - 1, (n,), generator=g, device=dev)
            if repeat_fraction > 0 and self.repeat_starts is not None:
                from_rep = torch.rand(n, generator=g, device=dev) < repeat_fraction
                pick = self.repeat_starts[torch.randint(0, self.repeat_starts.numel(), (n,), generator=g, device=dev)]
                inside = torch.randint(0, max(1, self.repeat_len - RL), (n,), generator=g, device=dev)
                gpos = torch.where(from_rep, pick + inside, c * L + s)
                c, s = gpos // L, gpos % L
                clean[lo:hi] &= ~from_rep
            gap_at = torch.full((n,), 1 << 40, dtype=torch.int64, device=dev)       # concatenated-genome offset where an intron is skipped
            if spliced_fraction > 0 and self.intron_starts is not None:
                spl = torch.rand(n, generator=g, device=dev) < spliced_fraction
                k = self.intron_starts[torch.randint(0, self.intron_starts.numel(), (n,), generator=g, device=dev)]
                before = torch.randint(20, RL - 20, (n,), generator=g, device=dev)  # bases of the first mate in front of the intron
                s = torch.where(spl, k - before, s)
                gap_at = torch.where(spl, c * L + k, gap_at)
                clean[lo:hi] &= ~spl
            start = c * L + s                                                      # leftmost base, concatenated-genome offset

            def take(first):
                idx = first[:, None] + ar[None, :]
                idx = idx + (idx >= gap_at[:, None]).to(torch.int64) * self.intron_len      # the transcript skips the intron
                return self.codes[idx]

            if paired:
                ins = torch.randint(max(200, RL), max(400, RL + 50), (n,), generator=g, device=dev)
                flip = torch.rand(n, generator=g, device=dev) < 0.5
                right_first = start + ins - RL
                left, right = take(start), take(right_first)
                # FR pair: the mate at the fragment's left end reads forward, the other one is reverse-complemented; `flip`
                # decides which file (R1 / R2) gets which end
                left_t, right_t = letters[left.long()], letters[(3 - right).flip(1).long()]
                text[lo:hi, 0] = torch.where(flip[:, None], right_t, left_t)
                text[lo:hi, 1] = torch.where(flip[:, None], left_t, right_t)
                right_pos = right_first + (right_first >= gap_at).to(torch.int64) * self.intron_len
                pl = self.lin_first[c] + s
                pr = self.lin_first[c] + (right_pos - c * L)
                truth[lo:hi, 0] = torch.where(flip, pr, pl)
                truth[lo:hi, 1] = torch.where(flip, pl, pr)
            else:
                rc = torch.rand(n, generator=g, device=dev) < 0.5
                fw = take(start)
                text[lo:hi, 0] = torch.where(rc[:, None], letters[(3 - fw).flip(1).long()], letters[fw.long()])
                truth[lo:hi, 0] = self.lin_first[c] + s
        flat = text.view(n_frags, mates * RL)
        if indel_read_rate > 0:
            # one indel of 1-3 bases in a fraction of the reads: a deletion closes the gap with what followed in the read (the
            # tail is refilled with random bases), an insertion pushes the tail out
            for lo in range(0, n_frags, piece):
                hi = min(n_frags, lo + piece)
                n = hi - lo
                for m in range(mates):
                    t = text[lo:hi, m]
                    has = torch.rand(n, generator=g, device=dev) < indel_read_rate
                    at = torch.randint(20, max(21, RL - 20), (n,), generator=g, device=dev)
                    d = torch.randint(1, 4, (n,), generator=g, device=dev)
                    is_del = torch.rand(n, generator=g, device=dev) < 0.5
                    shift = torch.where(is_del, d, -d)
                    src = ar[None, :] + torch.where(ar[None, :] >= at[:, None], shift[:, None], torch.zeros_like(shift)[:, None])
                    rnd = letters[torch.randint(0, 4, (n, RL), generator=g, device=dev)]
                    fresh = (src >= RL) | ((~is_del)[:, None] & (ar[None, :] >= at[:, None]) & (ar[None, :] < (at + d)[:, None]))
                    moved = torch.gather(t, 1, src.clamp(0, RL - 1))
                    out = torch.where(fresh, rnd, moved)
                    text[lo:hi, m] = torch.where(has[:, None], out, t)
                    clean[lo:hi] &= ~has
        if sub_rate > 0 or n_rate > 0:
            per = flat.shape[1]
            for lo in range(0, n_frags, piece):
                hi = min(n_frags, lo + piece)
                part = flat[lo:hi]
                if sub_rate > 0:
                    hit = torch.rand(part.shape, generator=g, device=dev) < sub_rate
                    code = (part == ord("G")).to(torch.uint8) + 2 * (part == ord("C")).to(torch.uint8) + 3 * (part == ord("T")).to(torch.uint8)
                    shift = torch.randint(1, 4, part.shape, generator=g, device=dev, dtype=torch.uint8)
                    part[:] = torch.where(hit, letters[((code + shift) & 3).long()], part)
                    clean[lo:hi] &= ~hit.view(hi - lo, per).any(1)
                if n_rate > 0:
                    hit = torch.rand(part.shape, generator=g, device=dev) < n_rate
                    part[:] = torch.where(hit, torch.full_like(part, ord("N")), part)
                    clean[lo:hi] &= ~hit.view(hi - lo, per).any(1)
        return text.reshape(-1), truth, clean


def write_fastq(path, reads):
    """reads: uint8 numpy [n, L] -> plain FASTQ, names R<11 digits>, constant quality H."""
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
    with open(path, "wb") as fh:
        fh.write(rec.tobytes())
