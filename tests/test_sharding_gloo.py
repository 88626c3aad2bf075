"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard a chunk's fragments the way
bench.py / a one-thread-per-GPU host does (contiguous slices, disjoint bigtable ranges, no
data-path collective) and the gathered result equals the single-process one.  The per-shard
"device" here is the oracle -- this test is about the partitioning and gather logic only."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT, load_case
from subread_b200.sharding import shard_range


def test_shard_ranges_cover_disjointly():
    for n in (0, 1, 7, 1000, 6990506):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, index_dir, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import json
    import torch
    from oracle import oracle
    from subread_b200.reads import ReadChunk
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cases = json.load(open(os.path.join(GOLDEN, "cases.json")))
    case = next(c for c in cases if c["name"] == "pe_gapped")
    prefix, chunk, params, exp = load_case(case, index_dir)
    f0, f1 = shard_range(chunk.n_frags, world, rank)
    r0, r1 = f0 * chunk.mates, f1 * chunk.mates
    sub = ReadChunk(chunk.text, chunk.off[r0:r1 + 1], chunk.len[r0:r1], f1 - f0, chunk.is_paired, [])
    idx = oracle.OracleIndex(prefix)
    part, _, _ = oracle.vote_chunk(idx, params, sub)
    # gather to rank 0 for the comparison only (the product writes disjoint host slices instead)
    mine = torch.from_numpy(np.frombuffer(part.tobytes(), dtype=np.uint8).copy())
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.numel()], dtype=torch.int64))
    bufs = [torch.zeros(int(s.item()), dtype=torch.uint8) for s in sizes]
    if rank == 0:
        bufs[0] = mine
        for r in range(1, world):
            dist.recv(bufs[r], src=r)
        whole = b"".join(b.numpy().tobytes() for b in bufs)
        open(out_path, "w").write("ok" if whole == exp.tobytes() else "mismatch")
    else:
        dist.send(mine, dst=0)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_the_single_process_table(golden_index_dir, oracle_mod, tmp_path):
    out = str(tmp_path / "result.txt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, golden_index_dir, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
