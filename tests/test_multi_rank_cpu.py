"""Host-side logic of the N>1 paths on CPU, world_size 2, gloo: the cell partition, the exchange of the connection blobs
(the only thing torch.distributed does for a decomposed run), and bench.py's reference arm under torchrun (rank 0 alone runs
and prints, the other ranks exit 0 without work)."""
import json
import os
import subprocess
import sys

import pytest

import openrbc_b200 as orbc
from openrbc_b200 import engine
from oracle import ref as refmod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cell_partition_matches_reference_formula():
    """util_numa.h:41-42: beg = tid * range / ntd, end = (tid + 1) * range / ntd, last worker takes the remainder."""
    for nc in (1, 7, 208, 9424, 188549):
        for world in range(1, 9):
            ranges = [engine.cell_range(nc, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nc
            for r in range(world):
                assert ranges[r][0] == r * nc // world
                if r:
                    assert ranges[r][0] == ranges[r - 1][1]
    with pytest.raises(engine.OrbcError):
        engine.cell_range(10, 3, 2)


WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
import openrbc_b200 as orbc
from openrbc_b200 import engine
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = orbc.load_library().orbc_mg_blob_bytes()
mine = bytes([rank]) * n                      # stands for Simulation.mg_export() (needs a GPU)
blobs = [None] * world
dist.all_gather_object(blobs, mine)
assert [b[0] for b in blobs] == list(range(world)) and all(len(b) == n for b in blobs)
cb, ce = engine.cell_range(9424, rank, world)
ends = [None] * world
dist.all_gather_object(ends, (cb, ce))
assert ends[0][0] == 0 and ends[-1][1] == 9424 and all(ends[i][1] == ends[i + 1][0] for i in range(world - 1))
dist.barrier()
if rank == 0:
    print("ok", world, n)
dist.destroy_process_group()
"""


def torchrun(args, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port)] + args
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, cwd=ROOT)


def test_blob_exchange_and_partition_over_gloo(tmp_path):
    if not os.path.exists(orbc.library_path()):
        orbc.build_library()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = torchrun([str(script), ROOT], 29531)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().splitlines()[-1].startswith("ok 2")


@pytest.mark.skipif(not refmod.available("fast"), reason="oracle/_ref/libref_fast.so not built")
def test_reference_arm_under_torchrun():
    out = torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--workload", "sphere", "--steps", "3", "--warmup", "1", "--ref-budget", "20"], 29533)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
