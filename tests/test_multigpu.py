"""N > 1 paths.  CPU (gloo, world_size 2): the host-side plumbing bench.py uses around the engine -- unique-id broadcast,
slice bounds, max-over-ranks timing.  GPU (needs >= 2 devices): results at 2 ranks are bit-identical to 1 rank."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


SLICE_ALIGN = 8192      # kSliceAlign (mcl_shard.cuh): chunks, groups and estimate partials never straddle ranks


def slice_bounds(n, rank, world):
    """Rank r owns the global particle slice [b(r), b(r+1)), b(r) = n*r/world rounded down to a multiple of 8192,
    b(world) = n (mcl_comm_init)."""
    def b(r):
        return n if r >= world else (n * r // world) // SLICE_ALIGN * SLICE_ALIGN
    return b(rank), b(rank + 1)


def test_slices_partition_the_cloud():
    for n in (2, 7, 40_001, 16_000_000, 64_000_000):
        for world in (1, 2, 4, 8):
            edges = [slice_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert all(a % SLICE_ALIGN == 0 for a, _ in edges)
            assert max(sizes) - min(sizes) < 2 * SLICE_ALIGN


GLOO_WORKER = r'''
import os, sys, torch, torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# unique-id style broadcast: 128 bytes from rank 0
uid = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
dist.broadcast(uid, 0)
assert bytes(uid.tolist()) == bytes(range(128))
# max-over-ranks of a per-rank elapsed time, the way bench.py reduces its CUDA-event timings
t = torch.tensor([10.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == 10.0 + world - 1
# slice exchange semantics (what the engine's ncclAllGather does in place): every rank ends with the full array
n = 1001
lo, hi = n * rank // world, n * (rank + 1) // world
full = torch.full((n,), -1, dtype=torch.int32)
full[lo:hi] = torch.arange(lo, hi, dtype=torch.int32)
for r in range(world):
    a, b = n * r // world, n * (r + 1) // world
    part = full[a:b].clone()
    dist.broadcast(part, r)
    full[a:b] = part
assert torch.equal(full, torch.arange(n, dtype=torch.int32))
dist.destroy_process_group()
print("ok", rank)
'''


def test_gloo_world2_host_plumbing(tmp_path):
    script = tmp_path / "gloo_worker.py"
    script.write_text(GLOO_WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def _run_worker(world, out, particles):
    cmd = [sys.executable]
    if world > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                "--master-port", str(29550 + world)]
    cmd += [os.path.join(ROOT, "tests", "mgpu_worker.py"), "--out", out, "--particles", str(particles)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("particles", [40_001, 262_144, 1_000_003])
def test_results_do_not_depend_on_gpu_count(tmp_path, particles):
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    outs = {}
    for world in [w for w in (1, 2, 4, 8) if w <= ngpu]:
        path = str(tmp_path / f"w{world}.npz")
        _run_worker(world, path, particles)
        outs[world] = np.load(path)
    base = outs[1]
    for world, o in outs.items():
        for k in ("pose", "parent_pose"):
            for f, t in (("utime", np.int64), ("x", np.uint32), ("y", np.uint32), ("theta", np.uint32)):
                assert np.array_equal(np.ascontiguousarray(o["cloud"][k][f]).view(t),
                                      np.ascontiguousarray(base["cloud"][k][f]).view(t)), (world, k, f)
        assert np.array_equal(o["cloud"]["weight"].view(np.uint64), base["cloud"]["weight"].view(np.uint64))
        assert np.array_equal(o["estimates"], base["estimates"])
        if world > 1:
            # one pose exchange (a copy-engine push) + five flag barriers (two per sequential sum, one for the estimate)
            assert int(o["collectives"]) == 6
            assert int(o["local"]) == slice_bounds(particles, 0, world)[1]
