"""subread_b200.index_build against the reference builder's own files (golden: subread-buildindex -M 1
[-F] on tests/golden/tiny.fa.gz).  The on-disk index format is a frozen interface, so the bar is
byte identity of .tab, .array and .reads.  CPU only."""
import gzip
import os

import numpy as np
import pytest

from conftest import GOLDEN
from subread_b200.index_build import (build_index, calculate_buckets_by_size, one_block_bucket_count)


def test_bucket_counts_match_the_reference_formula():
    assert one_block_bucket_count(3) == 16207831      # -B gapped, regardless of genome
    assert one_block_bucket_count(1) == 93018839      # -B -F
    assert calculate_buckets_by_size(131072, 3) == calculate_buckets_by_size(131072, 1)  # both floor at 0x3ffff+4


@pytest.mark.parametrize("name,gapped", [("tiny", True), ("tinyF", False)])
def test_builder_is_byte_identical_to_reference(name, gapped, tmp_path):
    expected_hash_items = int(1 * 1024.0 / 8.0) * 1024  # -M 1, index-builder.c:1269
    nb = calculate_buckets_by_size(expected_hash_items, 3 if gapped else 1)
    info = build_index(os.path.join(GOLDEN, "tiny.fa.gz"), str(tmp_path / name), gapped=gapped, buckets=nb, use_gpu=False)
    assert info["contigs"] == 2
    for ext in (".00.b.tab", ".00.b.array"):
        mine = open(tmp_path / (name + ext), "rb").read()
        ref = gzip.open(os.path.join(GOLDEN, name + ext + ".gz"), "rb").read()
        assert len(mine) == len(ref), ext
        assert mine == ref, ext
    assert open(tmp_path / (name + ".reads"), "rb").read() == open(os.path.join(GOLDEN, name + ".reads"), "rb").read()


def test_sanitising_rules(tmp_path):
    """Lower case is folded, every non-ACGT character becomes 'A', contigs of <= 16 bases are dropped
    (check_and_convert_FastA, index-builder.c:789-970)."""
    from subread_b200.index_build import read_fasta_sanitised
    fa = tmp_path / "x.fa"
    fa.write_text(">c1 some description\nacgtNNRY.-ACGTACGTACGT\nACGT\n>short\nACGTACGT\n>c3|x\nGGGGGGGGGGGGGGGGGGGG\n")
    names, seqs = read_fasta_sanitised(str(fa))
    assert names == ["c1", "c3|x"]
    assert seqs[0].tobytes() == b"ACGTAAAAAAACGTACGTACGTACGT"
    assert seqs[1].tobytes() == b"G" * 20


def test_device_builder_equals_array_builder():
    """build_index_arrays_device (torch; the builder for human-sized genomes) gives the arrays of build_index_arrays -- whose
    files are byte-identical to subread-buildindex -B (above) -- on a genome with repeats (ties of both parities, keys
    sampled more often than the threshold), gapped and full, on the CPU."""
    import torch
    from subread_b200.index_build import _CODE, build_index_arrays, build_index_arrays_device
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    unit = acgt[rng.integers(0, 4, size=300)]
    seqs = []
    for L in (5000, 777, 12001, 64):
        s = acgt[rng.integers(0, 4, size=L)]
        if L > 1000:
            for k in range(0, L - 400, 900):
                s[k:k + 300] = unit
        seqs.append(s)
    seqs.append(np.tile(acgt[rng.integers(0, 4, size=3)], 400))
    names = [f"c{i}" for i in range(len(seqs))]
    for gapped in (True, False):
        a = build_index_arrays(names, seqs, gapped=gapped, threshold=20, buckets=262147, use_gpu=False)
        codes = torch.from_numpy(np.concatenate([_CODE[s] for s in seqs]))
        b = build_index_arrays_device(names, codes, [len(s) for s in seqs], gapped=gapped, threshold=20, buckets=262147, chunk=1000)
        for k in ("bucket_start", "keys", "vals", "values", "read_offsets"):
            assert np.array_equal(a[k], b[k]), (gapped, k)
        assert (a["n_items"], a["length"], a["removed_keys"]) == (b["n_items"], b["length"], b["removed_keys"])
        assert a["removed_keys"] > 0


def _write_random_genome(path, lens, seed):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "w") as fh:
        for i, L in enumerate(lens):
            s = acgt[rng.integers(0, 4, L)].tobytes().decode()
            fh.write(f">contig{i} made up\n")
            for k in range(0, L, 70):
                fh.write(s[k:k + 70] + "\n")


@pytest.mark.parametrize("tag,lens,full,memory_mb", [
    ("cuts_at_contig_starts", [420000] * 4 + [100000, 17, 500000], False, 1),       # a block per contig: the cut waits for a contig's first 32 bases
    ("cuts_at_contig_starts_full", [150000] * 3 + [40], True, 1),
    ("cuts_inside_a_long_contig", [5500003, 30000], False, 8),                        # past 2 Mbp into a contig: cut there, step 1 999 990 bases back
])
def test_multiblock_builder_is_byte_identical_to_reference(tag, lens, full, memory_mb, tmp_path):
    """subread-buildindex WITHOUT -B cuts the index into blocks (index-builder.c:255-330, :353); build_index_multiblock follows the
    same cuts -- the overlap between consecutive blocks included -- and every .NN.b.tab / .NN.b.array / .reads is byte-identical
    to the reference builder's, run here on the same FASTA."""
    import glob
    import subprocess
    from conftest import REF_DIR
    from subread_b200.index_build import build_index_multiblock
    exe = os.path.join(REF_DIR, "bin", "subread-buildindex")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref (reference binaries) not built")
    fa = str(tmp_path / "g.fa")
    _write_random_genome(fa, lens, 3)
    subprocess.run([exe, "-o", str(tmp_path / "ref"), "-M", str(memory_mb)] + (["-F"] if full else []) + [fa], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    info = build_index_multiblock(fa, str(tmp_path / "mine"), gapped=not full, memory_mb=memory_mb)
    assert info["blocks"] >= 3
    ref_files = sorted(os.path.basename(f)[len("ref"):] for f in glob.glob(str(tmp_path / "ref.*")) if f.endswith((".tab", ".array", ".reads")))
    my_files = sorted(os.path.basename(f)[len("mine"):] for f in glob.glob(str(tmp_path / "mine.*")))
    assert ref_files == my_files and len(ref_files) == 2 * info["blocks"] + 1
    for f in ref_files:
        assert open(str(tmp_path / ("ref" + f)), "rb").read() == open(str(tmp_path / ("mine" + f)), "rb").read(), f
