"""CPU-only, world_size 2 over gloo: the N>1 host logic of bench.py (independent targets per rank, sum of
units and max of time across ranks)."""
import os
import subprocess
import sys

from conftest import ROOT

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from ries_b200.parallel import shard_targets, aggregate
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = shard_targets(["1.5", "2.5", "3.5", "4.5", "5.5"], rank, world)
units, tmax = aggregate(dist, units=float(len(mine)), seconds=1.0 + rank, device="cpu")
if rank == 0:
    print("RESULT", mine, units, tmax)
dist.destroy_process_group()
"""


def test_two_rank_sharding_and_aggregation(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][0]
    assert "['1.5', '3.5', '5.5'] 5.0 2.0" in line
