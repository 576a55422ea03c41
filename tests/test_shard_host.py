"""Host-side logic of the N>1 path on CPU: shard ranges, and a world_size-2 gloo rendezvous that distributes the
(fake) communicator id the way Engine.comm_init does and checks every rank derives a consistent partition."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition():
    from intrinsic3d_b200.engine import shard_range
    for n in (1, 511, 512, 100000, 2000596):
        for world in (1, 2, 3, 4, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            for a, b in zip(rs[:-1], rs[1:]):
                assert a[1] == b[0] and a[0] <= a[1]
            if n >= 512 * world:
                sizes = [e - b for b, e in rs]
                assert max(sizes) - min(sizes) <= 512 * world


def test_cut_ranges_balance_by_cost():
    """Work-balanced cuts (balanced_shard_ranges / rebalance_by_time): contiguous, aligned, monotone, and balanced in summed cost."""
    import torch
    from intrinsic3d_b200.engine import cut_ranges
    rng = np.random.default_rng(5)
    n = 200000
    cost = rng.uniform(0.0, 7.0, n)
    cost[: n // 4] *= 3.0                       # a heavy slab: equal-count cuts would be 2x off
    cost[n // 2: n // 2 + 20000] = 0.0          # an inactive region
    t = torch.from_numpy(cost)
    for world in (1, 2, 3, 8):
        rs = cut_ranges(t, world, 64)
        assert rs[0][0] == 0 and rs[-1][1] == n and len(rs) == world
        sums = []
        for (b, e), nxt in zip(rs, rs[1:] + [(n, n)]):
            assert b <= e and e == nxt[0]
            assert e == n or e % 64 == 0
            sums.append(cost[b:e].sum())
        assert max(sums) - min(sums) <= 0.02 * cost.sum() / world + 64 * 21.0
    # degenerate inputs: all-zero cost and fewer voxels than ranks still give a partition
    for cst, world in ((torch.zeros(1000, dtype=torch.float64), 4), (torch.ones(3, dtype=torch.float64), 8)):
        rs = cut_ranges(cst, world, 64)
        assert rs[0][0] == 0 and rs[-1][1] == cst.shape[0] and all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(rs[:-1], rs[1:]))


_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["I3D_ROOT"])
import torch, torch.distributed as dist
from intrinsic3d_b200.engine import shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ident = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
dist.broadcast(ident, 0)
assert ident.tolist() == list(range(128))
n = 123457
b, e = shard_range(n, rank, world)
t = torch.tensor([e - b], dtype=torch.int64)
dist.all_reduce(t)
assert t.item() == n
rows = torch.zeros(n, dtype=torch.int32); rows[b:e] = 1
dist.all_reduce(rows)
assert int(rows.min()) == 1 and int(rows.max()) == 1      # every voxel's rows owned by exactly one rank
dist.destroy_process_group()
print("OK", rank)
'''


def test_gloo_world2_partition(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, I3D_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.count("OK") == 2
