"""world_size-2 gloo test of the multi-rank plumbing (CPU): table replication
from rank 0 + contiguous batch sharding, with the oracle standing in for the
per-rank device work.  Launched as two real processes."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(%(root)r, "hexl-fpga_b200")); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import oracle_binding as ob
from sharding import shard, replicate
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, q, batch = 1024, ob.primes(1, 40, 1024)[0], 7
# tables exist on rank 0 only, then get replicated
tabs = torch.zeros((2, n), dtype=torch.int64)
if rank == 0:
    t = ob.Tables(n, q)
    tabs[0] = torch.from_numpy(t.roots.view(np.int64)); tabs[1] = torch.from_numpy(t.precon.view(np.int64))
replicate([tabs])
roots, precon = tabs[0].numpy().view(np.uint64).copy(), tabs[1].numpy().view(np.uint64).copy()
data = np.stack([ob.splitmix(n, 100 + i, q) for i in range(batch)])      # same global batch on every rank
start, count = shard(batch, world, rank)
mine = data[start:start + count].copy()
for i in range(count):
    ob.oracle().ho_fwd_ntt(ob.P(mine[i]), n, q, ob.P(roots), ob.P(precon))
gathered = [None] * world
dist.all_gather_object(gathered, (start, count, mine))
if rank == 0:
    full = ob.Tables(n, q)
    covered = np.zeros(batch, dtype=int)
    for s, c, part in gathered:
        covered[s:s + c] += 1
        for i in range(c):
            assert np.array_equal(part[i], ob.fwd_ntt(data[s + i], full)), (s, i)
    assert (covered == 1).all(), covered
    print("MULTI_RANK_OK")
dist.destroy_process_group()
'''


def test_shard_covers_batch_exactly_once():
    sys.path.insert(0, os.path.join(ROOT, "hexl-fpga_b200"))
    from sharding import shard

    for batch in (0, 1, 7, 8, 4096, 32768):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, c = shard(batch, world, r)
                seen += list(range(s, s + c))
            assert seen == list(range(batch))
    with pytest.raises(ValueError):
        shard(4, 2, 2)


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29571", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "MULTI_RANK_OK" in out.stdout
