"""Worker for the multi-GPU parity test (launched under torch.distributed.run, one rank per GPU).
Runs a short filter trajectory at WORLD_SIZE ranks and writes the final cloud of rank 0 to --out; the caller compares
the files produced at different world sizes bit for bit (the engine's multi-GPU results must not depend on the GPU
count) and against the oracle."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--particles", type=int, default=40_001)   # odd: ragged slices
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from botlab_b200 import engine, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grid = synth.make_map(300, seed=11)
    rng = np.random.default_rng(5)
    truth = synth.find_free_pose(grid, rng)
    n = args.particles
    e = engine.Engine(n, device=local)
    if world > 1:
        uid = torch.tensor(list(engine.comm_unique_id()) if rank == 0 else [0] * 128, dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        e.comm_init(bytes(uid.cpu().tolist()), rank, world)
    e.set_map(grid.cells, grid.origin_x, grid.origin_y, grid.meters_per_cell, grid.cells_per_meter)
    t = 1_000_000
    e.init_at_pose(*truth, utime=t, seed=123)
    am = engine.ActionModel()
    am.update(*truth, t)
    pose = truth
    ests = []
    for k in range(args.steps):
        pose = synth.odometry_step(rng, pose, step=(0.04, 0.02, 0.02))
        t += 100_000
        r, th, tt = synth.make_scan(grid, pose, seed=k, t0=t - 100_000 + 277)
        assert am.update(*pose, t)
        est = e.update(am, t, r, th, tt, (0.3 + 0.1 * k) / n)
        ests.append((est.x, est.y, est.theta))
    st = e.stats()
    cloud = e.export_particles()          # collective
    if rank == 0:
        np.savez(args.out, cloud=cloud, estimates=np.array(ests), collectives=st["collectives"],
                 local=st["local_particles"])
    e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
