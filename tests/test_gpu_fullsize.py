"""BASELINE-sized parity through size-independent properties (configs[1]: 100 Mbp genome, 10 M x 150 bp SE reads; and
the paired-end shape of configs[2] on the same genome).  The oracle cannot map 10^7 reads in seconds, so the full-size
run is pinned by what must hold whatever the size:
  * an error-free read of a random genome reports its own position in slot 0 with every subread voting,
  * mapping a chunk in pieces gives the bytes of mapping it whole (fragments are independent, core.c:3093),
  * the tiers agree: tier 0, the warp kernel alone and the host-buffer entry point give the same bytes,
  * a prefix of the chunk equals the oracle byte for byte."""
import os
import sys
import zlib

import numpy as np
import pytest

from conftest import ROOT
from subread_b200 import MAPPING_RESULT, VoteParams, used_slots_equal

pytestmark = pytest.mark.gpu

N_CHR, CHR_LEN, READ_LEN = 200, 500000, 150


@pytest.fixture(scope="module")
def world():
    import torch
    from subread_b200.binding import Mapper
    from subread_b200.index_build import build_index_arrays_device
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(20261018)
    codes = torch.randint(0, 4, (N_CHR * CHR_LEN,), generator=gen, device=dev, dtype=torch.uint8)
    a = build_index_arrays_device([f"chr{c + 1}" for c in range(N_CHR)], codes, [CHR_LEN] * N_CHR, gapped=True)
    m = Mapper(0)
    m.index_upload(a["index_gap"], a["padding"], a["bucket_start"], a["keys"], a["vals"], a["start_point"], a["length"],
                   a["values"], a["read_offsets"])
    os.environ["SRCUDA_SIMPLE_TIER"] = "0"
    try:
        plain = Mapper(0)
    finally:
        del os.environ["SRCUDA_SIMPLE_TIER"]
    plain.index_upload(a["index_gap"], a["padding"], a["bucket_start"], a["keys"], a["vals"], a["start_point"], a["length"],
                       a["values"], a["read_offsets"])
    yield dict(torch=torch, dev=dev, codes=codes, arrays=a, mapper=m, plain=plain,
               lin_first=torch.from_numpy(a["contig_starts"]).to(dev))
    m.close()
    plain.close()


def make_reads(w, n_frags, paired, seed):
    """Error-free reads on the device: text [n_frags * mates * READ_LEN] and the true linear position of each mate."""
    torch, dev, codes = w["torch"], w["dev"], w["codes"]
    letters = torch.tensor(list(b"AGCT"), dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    mates = 2 if paired else 1
    ar = torch.arange(READ_LEN, device=dev)
    text = torch.empty((n_frags, mates, READ_LEN), dtype=torch.uint8, device=dev)
    truth = torch.empty((n_frags, mates), dtype=torch.int64, device=dev)
    for lo in range(0, n_frags, 1 << 20):
        hi = min(n_frags, lo + (1 << 20))
        n = hi - lo
        c = torch.randint(0, N_CHR, (n,), generator=g, device=dev)
        flip = torch.rand(n, generator=g, device=dev) < 0.5
        if paired:
            ins = torch.randint(200, 400, (n,), generator=g, device=dev)
            s = torch.randint(0, CHR_LEN - 400, (n,), generator=g, device=dev)
            left = codes[(c * CHR_LEN + s)[:, None] + ar[None, :]]
            rs = s + ins - READ_LEN
            right = codes[(c * CHR_LEN + rs)[:, None] + ar[None, :]]
            lt, rt = letters[left.long()], letters[(3 - right).flip(1).long()]
            text[lo:hi, 0] = torch.where(flip[:, None], rt, lt)
            text[lo:hi, 1] = torch.where(flip[:, None], lt, rt)
            pl, pr = w["lin_first"][c] + s, w["lin_first"][c] + rs
            truth[lo:hi, 0] = torch.where(flip, pr, pl)
            truth[lo:hi, 1] = torch.where(flip, pl, pr)
        else:
            s = torch.randint(0, CHR_LEN - READ_LEN, (n,), generator=g, device=dev)
            fw = codes[(c * CHR_LEN + s)[:, None] + ar[None, :]]
            text[lo:hi, 0] = torch.where(flip[:, None], letters[(3 - fw).flip(1).long()], letters[fw.long()])
            truth[lo:hi, 0] = w["lin_first"][c] + s
    torch.cuda.synchronize()
    return text.reshape(-1), truth


def run_device(w, mapper, params, d_text, first, n_frags, paired):
    """Maps fragments [first, first + n_frags) of the chunk; returns the bigtable slice as a device tensor."""
    torch, dev = w["torch"], w["dev"]
    mates, mb = (2 if paired else 1), params.multi_best_reads
    d_off = (torch.arange(n_frags * mates, device=dev, dtype=torch.int64) + first * mates) * READ_LEN
    d_len = torch.full((n_frags * mates,), READ_LEN, dtype=torch.int16, device=dev)
    d_out = torch.full((n_frags * mates * mb * 68,), 0xAB, dtype=torch.uint8, device=dev)   # out_is_zero: never read
    torch.cuda.synchronize()
    mapper.vote_chunk_device(params, d_text.data_ptr(), d_text.numel(), d_off.data_ptr(), d_len.data_ptr(), n_frags, paired, READ_LEN,
                             d_out.data_ptr(), out_is_zero=True)
    return d_out


@pytest.mark.parametrize("paired,n_frags", [(False, 10_000_000), (True, 5_000_000)], ids=["se_10M", "pe_5M_pairs"])
def test_full_size_properties(paired, n_frags, world, oracle_mod):
    from subread_b200.reads import ReadChunk
    w = world
    torch = w["torch"]
    params = VoteParams.defaults(paired)
    mates, mb = (2 if paired else 1), params.multi_best_reads
    d_text, truth = make_reads(w, n_frags, paired, 4242 + int(paired))
    whole = run_device(w, w["mapper"], params, d_text, 0, n_frags, paired)
    t = w["mapper"].timing()
    # error-free reads of a random genome stay in the fast tier -- but for the few with more than SR_SIMPLE_DEFER_LIMIT chance hits
    # in a batch, which are handed to the warp kernel rather than holding up their warp (0.8 % at 150 bp on 100 Mbp)
    assert t.n_simple_punted < n_frags // 50 and t.n_overflow == 0, f"the fast tier queued {t.n_simple_punted} of {n_frags}"
    out = np.frombuffer(whole.cpu().numpy().tobytes(), dtype=MAPPING_RESULT).reshape(n_frags, mates, mb)
    tr = truth.cpu().numpy()
    # every read maps to its own position with all ten subreads (one of the three phases of each hits the gapped index)
    assert (out[:, :, 0]["selected_position"].astype(np.int64) == tr).all()
    assert (out[:, :, 0]["selected_votes"] == 10).all()
    assert (out[:, :, 0]["used_subreads_in_vote"] == 10).all()
    neg = (out[:, :, 0]["result_flags"] & 8) != 0
    assert 0.45 < neg[:, 0].mean() < 0.55
    # nothing else is written: the other slots of a unique read are all zero
    assert not out[:, :, 1:].view(np.uint8).any()
    # pieces == whole (ragged split points)
    cuts = [0, n_frags // 3 + 17, n_frags // 3 + 18, 2 * n_frags // 3 - 5, n_frags]
    crc_whole = zlib.crc32(out.tobytes())
    crc = 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        part = run_device(w, w["mapper"], params, d_text, a, b - a, paired)
        crc = zlib.crc32(part.cpu().numpy().tobytes(), crc)
    assert crc == crc_whole
    # tiers agree on a prefix: warp kernel alone, host-buffer entry point, oracle
    k = 200_000
    alone = run_device(w, w["plain"], params, d_text, 0, k, paired)
    assert w["plain"].timing().n_simple_punted == 0 and w["plain"].timing().ms_kernel_simple < 0.05
    assert alone.cpu().numpy().tobytes() == out[:k].tobytes()
    h_text = d_text[: k * mates * READ_LEN].cpu().numpy()
    chunk = ReadChunk(h_text, np.arange(k * mates + 1, dtype=np.uint64) * READ_LEN, np.full(k * mates, READ_LEN, np.uint16), k, paired, [])
    host, _ = w["mapper"].vote_chunk(chunk, params)
    assert host.tobytes() == out[:k].tobytes()
    k = 20_000
    oidx = oracle_mod.OracleIndex.from_arrays(w["arrays"])
    chunk = ReadChunk(h_text[: k * mates * READ_LEN], np.arange(k * mates + 1, dtype=np.uint64) * READ_LEN,
                      np.full(k * mates, READ_LEN, np.uint16), k, paired, [])
    ora, _, _ = oracle_mod.vote_chunk(oidx, params, chunk, threads=min(16, os.cpu_count() or 1))
    ok, bad = used_slots_equal(out[:k], ora)
    assert ok, f"fragment slot {bad} differs from the oracle"
    assert out[:k].tobytes() == ora.tobytes()
    oidx.close()
