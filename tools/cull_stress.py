#!/usr/bin/env python
"""Offline stress run of the beam-culling rule (tests/certification_model.cull_flags, the numpy restatement of
table_cull_kernel) against the oracle's per-ray scores:  python tools/cull_stress.py [cases]
Every culled beam must score 0 for every particle.  CPU only."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from botlab_b200 import synth
from oracle import port
import certification_model as cm

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
tot_beams = tot_culled = wrong = evals = 0
for case in range(cases):
    rng = np.random.default_rng(99000 + case)
    mpc = float(rng.choice([0.05, 0.025, 0.1]))
    side = int(rng.integers(300, 900))
    grid = synth.make_map(side, seed=500 + case, meters_per_cell=mpc)
    theta = float(rng.uniform(-np.pi, np.pi))
    clearance = int(rng.choice([6, 16, 30]))
    x, y, _ = synth.find_free_pose(grid, rng, clearance=clearance)
    max_range = float(rng.choice([4.0, 8.0, 12.0]))
    r, th, t = synth.make_scan(grid, (x, y, theta), seed=case, max_range=max_range, num_beams=int(rng.choice([180, 360])))
    cloud = synth.make_particles(int(rng.choice([60, 150])), (x, y, theta), seed=case, sigma_xy=float(rng.choice([0.03, 0.1, 0.3])),
                                 sigma_theta=float(rng.choice([0.01, 0.05, 0.2])), parent_utime=int(t[0]) - 100_000,
                                 pose_utime=int(t[-1]), motion=(float(rng.uniform(-0.1, 0.1)), float(rng.uniform(-0.1, 0.1)),
                                                                float(rng.uniform(-0.2, 0.2))))
    t0, t1 = int(cloud["parent_pose"]["utime"][0]), int(cloud["pose"]["utime"][0])
    ratios = (t - t0).astype(np.float64) / float(t1 - t0)
    flags = cm.cull_flags(grid, cloud, r, th, ratios, 0.15)
    pg = port.Grid(grid.cells, grid.origin_x, grid.origin_y, grid.cells_per_meter)
    per_ray = np.stack([port.ray_scores(pg, cloud[i], r, th, t) for i in range(len(cloud))])
    bad = int((per_ray[:, flags] != 0).sum())
    wrong += bad
    tot_beams += len(flags); tot_culled += int(flags.sum()); evals += per_ray.size
    print(f"case {case:3d}: {side}x{side} @ {mpc} m, max range {max_range}, {len(cloud)} particles: {int(flags.sum())} of {len(flags)} beams culled, "
          f"{bad} non-zero scores among them")
print(f"TOTAL: {cases} geometries, {evals} evaluations, {tot_culled} of {tot_beams} beams culled, {wrong} wrong")
sys.exit(1 if wrong else 0)
