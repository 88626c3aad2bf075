"""The human-sized index of BASELINE configs[2]/[3] (3.1 Gbp synthetic genome, 6 200 contigs, one-block gapped index: 1.03 G
items, 64 per bucket, positions beyond 2^31) against the oracle on a STRATIFIED sample -- not the easy stratum only:
  clean       error-free reads (stay in the fast tier, every one at its own position)
  subst       1 % substitutions per base
  indel       one 1-3 bp insertion or deletion in every read
  with_n      an N in every read
  repeats     reads drawn from planted repeat families (several equally good positions: top-K ties, visit order)
  pe_* / sj   the same through the paired-end tier, and spliced pairs through subjunc's junction search
Every stratum is mapped through the C ABI (dense-table accelerators: direct filter) and compared with the oracle byte for
byte on every used slot; the oracle reads the same in-memory index arrays.  About two minutes on one B200."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, describe_mismatch
from subread_b200 import MAPPING_RESULT, SUBJUNC_RESULT, VoteParams, used_slots_equal

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, "tools"))

N_PER_STRATUM = 4000


@pytest.fixture(scope="module")
def human():
    import torch
    import synth
    from subread_b200.binding import Mapper
    dev = torch.device("cuda", 0)
    g = synth.Genome(6200, 500000, 20261018, dev)
    g.plant_repeats(300, 6, 2000)
    g.plant_introns()
    a = g.build_index(gapped=True)
    assert a["n_items"] > 1_000_000_000 and int(a["vals"].max()) > 2 ** 31
    m = Mapper(0)
    m.index_upload(a["index_gap"], a["padding"], a["bucket_start"], a["keys"], a["vals"], a["start_point"], a["length"],
                   a["values"], a["read_offsets"])
    assert m.describe().accelerators & 4, "a table of 64 items per bucket takes the direct filter"
    torch.cuda.empty_cache()
    yield dict(g=g, arrays=a, mapper=m, torch=torch)
    m.close()


STRATA = {
    "clean": dict(),
    "subst": dict(sub_rate=0.01),
    "indel": dict(indel_read_rate=1.0),
    "with_n": dict(n_rate=0.01),
    "repeats": dict(repeat_fraction=1.0),
    "mixed": dict(sub_rate=0.01, n_rate=0.002, indel_read_rate=0.14, repeat_fraction=0.05),
}


@pytest.mark.parametrize("mode", ["se150", "pe150", "subjunc_pe100"])
def test_stratified_sample_equals_the_oracle(mode, human, oracle_mod):
    from subread_b200.reads import ReadChunk
    g, m = human["g"], human["mapper"]
    paired, subjunc = mode != "se150", mode.startswith("subjunc")
    RL = 100 if subjunc else 150
    mates = 2 if paired else 1
    params = VoteParams.defaults(paired, subjunc)
    oidx = oracle_mod.OracleIndex.from_arrays(human["arrays"])
    strata = dict(STRATA)
    if subjunc:
        strata = {"clean": dict(), "spliced": dict(spliced_fraction=1.0), "spliced_noisy": dict(spliced_fraction=0.5, sub_rate=0.01, indel_read_rate=0.1)}
    report = {}
    for k, (name, kw) in enumerate(strata.items()):
        text, truth, clean = g.reads(N_PER_STRATUM, RL, 100 + 7 * k, paired=paired, **kw)
        chunk = ReadChunk(text.cpu().numpy(), np.arange(N_PER_STRATUM * mates + 1, dtype=np.uint64) * RL, np.full(N_PER_STRATUM * mates, RL, np.uint16),
                          N_PER_STRATUM, paired, [])
        js_o, js_g = ({}, {}) if subjunc else (None, None)
        exp, _, _ = oracle_mod.vote_chunk(oidx, params, chunk, threads=min(16, os.cpu_count() or 1), junc_state=js_o)
        got, _ = m.vote_chunk(chunk, params, junc_state=js_g)
        t = m.timing()
        ok, bad = used_slots_equal(got, exp)
        assert ok, f"{mode}/{name}: " + describe_mismatch(got, exp, bad)
        if subjunc:
            used = exp.reshape(-1)["selected_votes"] > 0
            assert np.array_equal(js_g["junc"][used], js_o["junc"][used]), f"{mode}/{name}: subjunc_result_t"
            assert np.array_equal(js_g["big_margin"], js_o["big_margin"]), f"{mode}/{name}: big-margin records"
            if "spliced" in name:
                assert (js_o["junc"]["minor_votes"] > 0).sum() > N_PER_STRATUM // 4, "the stratum must exercise the junction search"
        report[name] = (int(t.n_simple_punted), int(t.n_overflow), float((got[:, :, 0]["selected_votes"] > 0).mean()))
        if name == "clean":
            tr = truth.cpu().numpy()
            hit = (got[:, :, 0]["selected_position"].astype(np.int64) == tr).all(1)
            assert hit[clean.cpu().numpy()].mean() > 0.99
            assert t.n_simple_punted < N_PER_STRATUM // 10, f"{mode}: the fast tier queued {t.n_simple_punted} clean fragments"
    print(mode, {k: dict(queued=v[0], scalar=v[1], mapped=round(v[2], 3)) for k, v in report.items()})
    oidx.close()
