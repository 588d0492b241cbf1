"""N>1 host logic on CPU: world_size-2 gloo processes shard a batch of identically structured problems exactly like
bench.py does on GPUs (contiguous split, no data-path collective), factor their items with the CPU oracle and agree on
the per-item checksums through all_gather."""
import os
import subprocess
import sys
import textwrap

import pytest

from baspacho_b200.sharding import all_shards, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_partition_properties():
    for batch in (0, 1, 7, 8, 63, 64, 65):
        for world in (1, 2, 3, 4, 8):
            shards = all_shards(batch, world)
            covered = [i for lo, hi in shards for i in range(lo, hi)]
            assert covered == list(range(batch))
            assert max(hi - lo for lo, hi in shards) == (-(-batch // world) if batch else 0)
    assert shard_range(64, 3, 8) == (24, 32)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from baspacho_b200 import _capi
    from baspacho_b200.sharding import shard_range, gather_checksums, reduce_max
    from tests import helpers as H
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"], rank=int(os.environ["RANK"]),
                            world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    batch = 5
    sizes, ptrs, inds = H.ba_problem(120, 10, seed=3, window=4)
    s = H.oracle_cpu.OracleSolver.create(sizes, ptrs, inds, [0, 120], backend=_capi.BACKEND_REF)
    lo, hi = shard_range(batch, rank, world)
    sums = []
    for q in range(lo, hi):
        d = H.make_data(s, 100 + q, np.float64)
        s.factor(d)
        sums.append(float(np.abs(d).sum()))
    allsums = gather_checksums(sums)
    tmax = reduce_max(float(rank + 1))
    if rank == 0:
        ref = []
        for q in range(batch):
            d = H.make_data(s, 100 + q, np.float64)
            s.factor(d)
            ref.append(float(np.abs(d).sum()))
        assert len(allsums) == batch and np.allclose(allsums, ref, rtol=0, atol=0), (allsums, ref)
        assert tmax == float(world)
        print("OK")
    dist.destroy_process_group()
""") % ROOT


def test_two_rank_gloo_sharded_batch(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", PORT=port, MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0]
