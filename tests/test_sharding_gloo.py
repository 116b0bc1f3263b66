"""N > 1 host logic on CPU: world_size-2 gloo. Blocks shard round-robin with no data-path
collective; timing is max-over-ranks; rank 0 gathers the detection summaries in block order.
The per-block compute here is the CPU oracle (test infrastructure) standing in for the GPU."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _worker(rank, world, port, tmp):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root / "simple-radio-telescope-backend_b200"))
    sys.path.insert(0, str(root / "tests"))
    import torch.distributed as dist
    from srtb_b200 import sharding
    import oracle_lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = oracle_lib.load()
    n_blocks, n = 7, 1 << 12
    mine = sharding.blocks_for_rank(n_blocks, world, rank)
    local = []
    for k in mine:
        rng = np.random.default_rng(k)
        bb = np.clip(np.round(rng.standard_normal(n) * 20), -127, 127).astype(np.int8)
        x = o.unpack(bb.view(np.uint8), n, -8)
        local.append({"block": k, "stream": 0, "rank": rank, "sum": float(np.abs(x).sum())})
    dist.barrier()
    t = sharding.max_over_ranks(10.0 + rank, dist)
    merged = sharding.gather_results(local, dist)
    if rank == 0:
        np.save(os.path.join(tmp, "result.npy"),
                np.array([t] + [m["block"] for m in merged] + [m["rank"] for m in merged], dtype=np.float64))
    dist.barrier()
    dist.destroy_process_group()


def test_block_sharding_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = np.load(tmp_path / "result.npy")
    assert r[0] == 11.0                                   # max over ranks of (10, 11)
    assert r[1:8].tolist() == list(range(7))               # merged back in block order
    assert r[8:15].tolist() == [k % 2 for k in range(7)]   # block k ran on rank k mod 2


def test_sharding_is_a_partition():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "simple-radio-telescope-backend_b200"))
    from srtb_b200 import sharding
    for world in (1, 2, 4, 8):
        seen = sorted(b for r in range(world) for b in sharding.blocks_for_rank(37, world, r))
        assert seen == list(range(37))
        assert all(sharding.owner_of_block(b, world) == r for r in range(world)
                   for b in sharding.blocks_for_rank(37, world, r))
    with pytest.raises(ValueError):
        sharding.blocks_for_rank(4, 2, 2)


def _ranksync_worker(rank, world, port, tmp):
    """bench.py's rank plumbing (RankSync) on CPU: the host-only mode replaces the lazy NCCL default group by gloo,
    everything else — barrier, gather of the per-rank step times, SUM / MAX reductions of a result vector — is the
    code the multi-GPU bench runs"""
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      SRTB_BENCH_SYNC_BACKEND="gloo")
    import torch.distributed as dist_mod
    import bench
    rs = bench.RankSync(dist_mod, torch, rank)
    assert rs.get_world_size() == world
    rs.barrier()
    times = rs.all_gather_scalar(1.0 + 0.25 * rank)
    vals = torch.tensor([float(rank + 1), 10.0 * rank], dtype=torch.float64)
    mx = vals.clone()
    rs.all_reduce(vals, op=rs.ReduceOp.SUM)
    rs.all_reduce(mx, op=rs.ReduceOp.MAX)
    if rank == 0:
        np.save(os.path.join(tmp, "ranksync.npy"), np.array(times + vals.tolist() + mx.tolist()))
    rs.barrier()
    rs.destroy_process_group()


def test_bench_ranksync_world2(tmp_path):
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_ranksync_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = np.load(tmp_path / "ranksync.npy")
    assert r[:2].tolist() == [1.0, 1.25]          # per-rank step times, in rank order (the bench reports their max)
    assert r[2:4].tolist() == [3.0, 10.0]         # SUM
    assert r[4:6].tolist() == [2.0, 10.0]         # MAX

