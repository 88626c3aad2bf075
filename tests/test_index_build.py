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
