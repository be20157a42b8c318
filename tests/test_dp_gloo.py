"""CPU tests of the N>1 host logic with the gloo backend, world_size 2 (no GPU needed)."""
import os
import subprocess
import sys
import textwrap

import pytest

from ug_stereomatcher_b200 import dp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_pairs_covers_batch():
    for n, world in [(1050, 8), (1050, 4), (7, 3), (2, 4), (0, 2)]:
        shards = [dp.shard_pairs(n, world, r) for r in range(world)]
        flat = [i for s in shards for i in s]
        assert flat == list(range(n))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    assert len(dp.shard_pairs(1050, 8, 0)) == 132 and len(dp.shard_pairs(1050, 8, 7)) == 131
    with pytest.raises(ValueError):
        dp.shard_pairs(10, 2, 2)


def test_throughput_formula():
    assert dp.throughput(1, 8, 50.0) == pytest.approx(160.0)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from ug_stereomatcher_b200 import dp
    rank, world = dp.init("gloo")
    assert world == 2 and dist.get_world_size() == 2
    mine = dp.shard_pairs(11, world, rank)
    gathered = [None, None]
    dist.all_gather_object(gathered, mine)
    assert sorted(gathered[0] + gathered[1]) == list(range(11)), gathered
    # slowest rank defines the step time; job throughput counts every rank's pairs
    ms = dp.max_over_ranks(40.0 + 10.0 * rank)
    assert ms == 50.0, ms
    total = dp.sum_over_ranks(len(mine))
    assert total == 11.0
    dist.barrier()
    if rank == 0:
        print("OK", dp.throughput(1, world, ms))
    dist.destroy_process_group()
""")


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, timeout=280)
    out = p.stdout.decode(errors="ignore")
    assert p.returncode == 0, out[-3000:]
    assert "OK 40.0" in out
