"""Prints, for the randomized-geometry test cases, whether the float pass ran, its eps and the deferred fraction."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
from botlab_b200 import engine, synth

for case in range(10):
    rng = np.random.default_rng(1000 + case)
    mpc = [0.05, 0.025, 0.1, 0.05, 0.05][case % 5]
    w, h = int(rng.integers(150, 900)), int(rng.integers(150, 900))
    base = synth.make_map(max(w, h), seed=50 + case, meters_per_cell=mpc)
    cells = base.cells[:h, :w].copy()
    cells[-2:, :] = 100; cells[:, -2:] = 100
    ox, oy = [(-w * mpc / 2, -h * mpc / 2), (731.25, -412.5), (0.0, 0.0), (-2000.0, 1500.0), (5.5, 5.5)][case % 5]
    cpm = None if case != 7 else 1.0 / 0.05 * 1.01
    grid = synth.GridSpec(cells, ox, oy, mpc, cpm)
    truth = synth.find_free_pose(grid, rng)
    nb = int(rng.choice([180, 290, 360, 500, 720]))
    mr = float(rng.choice([4.0, 8.0, 12.0]))
    r, th, t = synth.make_scan(grid, truth, num_beams=nb, seed=case, max_range=mr)
    n = 20_000
    if case % 3 == 2:
        cloud = synth.make_uniform_particles(n, grid, seed=case, utime=int(t[-1]))
        cloud["parent_pose"]["utime"] = int(t[0])
    else:
        cloud = synth.make_particles(n, truth, seed=case, sigma_xy=float(rng.choice([0.05, 0.3, 1.5])),
                                     sigma_theta=float(rng.choice([0.02, 0.5])), parent_utime=int(t[0]), pose_utime=int(t[-1]))
    e = engine.Engine(n)
    e.set_map(grid.cells, grid.origin_x, grid.origin_y, grid.meters_per_cell, grid.cells_per_meter)
    e.import_particles(cloud)
    e.score(r, th, t)
    st = e.stats()
    print(case, f"{w}x{h} mpc {mpc} origin ({ox},{oy}) beams {nb} maxr {mr}: path {st['sensor_path']} tile {st['map_tile_used']} "
                f"G {st['lanes_per_particle']} eps {st['fast_eps']:.2e} deferred {st['deferred_evals'] / max(st['evals'], 1):.3f}")
    e.close()
