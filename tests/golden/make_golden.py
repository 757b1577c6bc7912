"""Generates tests/golden/*.npz by EXECUTING the compiled, unmodified reference (oracle/_ref/libbotlab_ref.so, built by
`make -C oracle` where /root/reference exists).  Inputs are stored next to the outputs so the tests never depend on
numpy's RNG stream.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from botlab_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REAL_MAP = "/root/reference/data/obstacle_slam_10mx10m_5cm.map"


def kat_scan():
    i = np.arange(360)
    ranges = (np.float32(1.0) + np.float32(0.002) * i.astype(np.float32)).astype(np.float32)
    return ranges, (i * 2 * np.pi / 360).astype(np.float32), (1000000 + 277 * i).astype(np.int64)


def part(p, q):
    a = np.zeros(1, ref.PARTICLE_DTYPE)
    a["pose"]["x"], a["pose"]["y"], a["pose"]["theta"], a["pose"]["utime"] = p
    a["parent_pose"]["x"], a["parent_pose"]["y"], a["parent_pose"]["theta"], a["parent_pose"]["utime"] = q
    return a


def mapping_golden():
    """Mapping::updateMap (mapping.cpp:17-127) over five consecutive scans on the real map: inputs and the grid after
    every call (the first call, with initialized_ false, must change nothing)."""
    g0 = ref.RefGrid.from_file(REAL_MAP)
    info = g0.info()
    # the recorded map is saturated (free cells at -128, walls at 127), where saturating updates change nothing: start
    # from a third of its log-odds so the sequence exercises both the adds and the saturation
    start = (g0.cells().astype(np.int32) // 3).astype(np.int8)
    g = ref.RefGrid.from_cells(start, info["origin_x"], info["origin_y"], info["meters_per_cell"])
    grid = synth.GridSpec(g0.cells(), info["origin_x"], info["origin_y"], info["meters_per_cell"], info["cells_per_meter"])
    rng = np.random.default_rng(77)
    pose = synth.find_free_pose(grid, rng, clearance=8)
    prev, t0 = pose, 2_000_000
    out = {"hit": 40, "miss": 25, "max_laser": 5.0, "steps": 5, "start_cells": start}
    for k in range(5):
        r, th, t = synth.make_scan(grid, pose, num_beams=290, seed=300 + k, t0=t0)
        prv = synth.make_pose(*prev, utime=int(t[0]))
        cur = synth.make_pose(*pose, utime=int(t[-1]))
        ref.map_update(g, prv, cur, k > 0, ref.Scan(r, th, t), 5.0, 40, 25)
        out[f"{k}_ranges"], out[f"{k}_thetas"], out[f"{k}_times"] = r, th, t
        out[f"{k}_previous"], out[f"{k}_pose"] = np.array(prv), np.array(cur)
        out[f"{k}_cells"] = g.cells()
        prev, pose, t0 = pose, synth.odometry_step(rng, pose, step=(0.05, 0.02, 0.04)), t0 + 100_000
    np.savez_compressed(os.path.join(OUT, "mapping.npz"), **out)


def distance_golden():
    """ObstacleDistanceGrid::setDistances (planning/obstacle_distance_grid.cpp:73-92) of the executed reference on four
    small grids (tests/test_oracle.py::_distance_cases builds the same inputs)."""
    rng = np.random.default_rng(77)
    cases = {
        "synth": synth.make_map(120, seed=5).cells,
        "random": np.where(rng.random((90, 140)) < 0.03, 60, np.where(rng.random((90, 140)) < 0.5, -9, 0)).astype(np.int8),
        "all_free": np.full((40, 60), -5, np.int8),
    }
    d = np.full((30, 30), -5, np.int8); d[11, 17] = 0
    cases["one_unknown"] = d
    out = {}
    for name, cells in cases.items():
        g = ref.RefGrid.from_cells(cells, 0.0, 0.0, 0.05)
        out[name + "_cells"] = cells
        out[name + "_dist"] = ref.distance_grid(g)
    np.savez_compressed(os.path.join(OUT, "distance.npz"), **out)


def main():
    if "--distance-only" in sys.argv:
        distance_golden()
        print("distance.npz written")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "mapping":
        mapping_golden()
        print("mapping golden written to", OUT)
        return
    mapping_golden()
    # ---- the real map fixture (data/obstacle_slam_10mx10m_5cm.map), as parsed by the reference itself
    g = ref.RefGrid.from_file(REAL_MAP)
    info = g.info()
    cells = g.cells()
    np.savez_compressed(os.path.join(OUT, "real_map.npz"), cells=cells, origin_x=info["origin_x"],
                        origin_y=info["origin_y"], meters_per_cell=info["meters_per_cell"],
                        cells_per_meter=info["cells_per_meter"])

    # ---- SURVEY Appendix B known answers, re-executed
    ranges, thetas, times = kat_scan()
    scan = ref.Scan(ranges, thetas, times)
    kat_particles = np.concatenate([
        part((0, 0, 0, 1100000), (0, 0, 0, 1100000)),
        part((0.5, -0.25, 0.3, 1100000), (0.48, -0.26, 0.28, 1000000)),
        part((-2.0, 1.5, -2.5, 1100000), (-2.02, 1.49, -2.45, 1000000)),
        part((1.0, 1.0, 3.1, 1100000), (0.98, 1.0, -3.1, 1000000)),
    ])
    kat_scores = np.concatenate([ref.likelihood(g, kat_particles[i:i + 1], scan) for i in range(4)])
    kat_rays_b = ref.moving_scan(scan, kat_particles["parent_pose"][1], kat_particles["pose"][1])
    kat_rays_d = ref.moving_scan(scan, kat_particles["parent_pose"][3], kat_particles["pose"][3])
    am = ref.RefActionModel()
    am.update(ref.make_pose(0, 0, 0))
    moved, act = am.update(ref.make_pose(0.02, 0.01, 0.01))
    two = np.concatenate([kat_particles[1:2], kat_particles[1:2]])
    act_out, act_draws = am.apply(two)
    am2 = ref.RefActionModel()
    am2.update(ref.make_pose(0, 0, 0))
    _, act_back = am2.update(ref.make_pose(-0.02, 0, 0))
    pf = ref.RefParticleFilter(8)
    ps = np.zeros(8, ref.PARTICLE_DTYPE)
    ps["weight"] = [.05, .20, .05, .30, .10, .10, .15, .05]
    for k in range(8):
        ps["pose"]["x"][k] = k
        ps["pose"]["y"][k] = np.float32(0.1) * np.float32(k)
        ps["pose"]["theta"][k] = 3 if k % 2 else -3
    pf.set_particles(ps)
    np.savez_compressed(os.path.join(OUT, "kat.npz"), ranges=ranges, thetas=thetas, times=times,
                        particles=kat_particles, scores=kat_scores, rays_b=kat_rays_b, rays_d=kat_rays_d,
                        action_forward=act, action_backward=act_back, action_in=two, action_out=act_out,
                        action_draws=act_draws, resample_particles=ps, resample_r=ref.resample_draw(1, 8),
                        resample_idx=pf.resample(1), estimate=np.array(pf.estimate(ps)))

    # ---- sensor model: real map and a synthetic map, interpolating and degenerate clouds
    rng = np.random.default_rng(20261017)
    sensor = {}
    for name, grid_spec in (("real", synth.GridSpec(cells, info["origin_x"], info["origin_y"], info["meters_per_cell"],
                                                     info["cells_per_meter"])),
                            ("synth", synth.make_map(200, seed=synth.MAP_SEED + 2))):
        rg = ref.RefGrid.from_cells(grid_spec.cells, grid_spec.origin_x, grid_spec.origin_y, grid_spec.meters_per_cell)
        truth = synth.find_free_pose(grid_spec, rng)
        r_, th_, t_ = synth.make_scan(grid_spec, truth, seed=7)
        for variant, (t0, t1) in (("interp", (int(t_[0]), int(t_[-1]))), ("degen", (int(t_[-1]), int(t_[-1])))):
            p = synth.make_particles(768, truth, seed=11, parent_utime=t0, pose_utime=t1)
            # a few hostile particles: off the map, huge, NaN heading, theta at +-pi
            p["pose"]["x"][0] = 1.0e4
            p["pose"]["x"][1] = np.float32(3.0e9); p["parent_pose"]["x"][1] = np.float32(3.0e9)
            p["pose"]["theta"][2] = np.float32(np.pi); p["parent_pose"]["theta"][2] = np.float32(-np.pi)
            p["pose"]["x"][3] = grid_spec.origin_x - 0.01; p["pose"]["y"][3] = grid_spec.origin_y - 0.01
            p["pose"]["x"][4] = np.nan
            s = ref.likelihood(rg, p, ref.Scan(r_, th_, t_))
            sensor[f"{name}_{variant}_particles"] = p
            sensor[f"{name}_{variant}_scores"] = s
        sensor[f"{name}_ranges"], sensor[f"{name}_thetas"], sensor[f"{name}_times"] = r_, th_, t_
        sensor[f"{name}_truth"] = np.array(truth)
        if name == "synth":
            sensor["synth_cells"] = grid_spec.cells
            sensor["synth_geom"] = np.array([grid_spec.origin_x, grid_spec.origin_y, grid_spec.meters_per_cell,
                                             grid_spec.cells_per_meter])
    np.savez_compressed(os.path.join(OUT, "sensor.npz"), **sensor)

    # ---- action model with the reference's own draws
    am = ref.RefActionModel(seed=12345)
    am.set_utime(2_000_000)
    am.update(ref.make_pose(0.3, -0.2, 0.1))
    moved, params = am.update(ref.make_pose(0.32, -0.19, 0.11))
    pin = synth.make_particles(512, (0.3, -0.2, 0.1), seed=5)
    pin["pose"]["theta"][:8] = np.float32([3.14, -3.14, 3.1415925, -3.1415925, 3.0, -3.0, 0.0, 1.5])
    pout, draws = am.apply(pin)
    np.savez_compressed(os.path.join(OUT, "action.npz"), params=params, moved=moved, particles_in=pin,
                        particles_out=pout, draws=draws, utime=2_000_000)

    # ---- resampling: filter-shaped weights, several sizes, the unseeded rand() draw
    res = {}
    for n in (200, 4096, 100_000):
        w = synth.filter_shaped_weights(n, seed=n)
        pf = ref.RefParticleFilter(n)
        ps = np.zeros(n, ref.PARTICLE_DTYPE)
        ps["weight"] = w
        pf.set_particles(ps)
        res[f"w_{n}"] = w
        res[f"r_{n}"] = ref.resample_draw(1, n)
        res[f"idx_{n}"] = pf.resample(1)
    np.savez_compressed(os.path.join(OUT, "resample.npz"), **res)

    # ---- normalise + estimate on the synthetic sensor case
    rg = ref.RefGrid.from_cells(sensor["synth_cells"], *[float(v) for v in sensor["synth_geom"][:3]])
    pf = ref.RefParticleFilter(768)
    post = pf.normalize(rg, sensor["synth_interp_particles"][5:], ref.Scan(sensor["synth_ranges"],
                                                                           sensor["synth_thetas"],
                                                                           sensor["synth_times"]))
    est = pf.estimate(post)
    np.savez_compressed(os.path.join(OUT, "normalize.npz"), proposal=sensor["synth_interp_particles"][5:],
                        posterior=post, estimate=np.array(est))

    # ---- full updateFilter trajectory on the real map, N = 300, both utime behaviours
    traj = {}
    grid_spec = synth.GridSpec(cells, info["origin_x"], info["origin_y"], info["meters_per_cell"], info["cells_per_meter"])
    for variant in ("interp", "legacy"):
        rng = np.random.default_rng(99)
        n = 300
        pose = (0.0, 0.0, 0.0)
        pf = ref.RefParticleFilter(n)
        cloud = synth.make_particles(n, pose, seed=3, sigma_xy=0.03, sigma_theta=0.02, parent_utime=900_000,
                                     pose_utime=900_000)
        cloud["parent_pose"] = cloud["pose"]
        pf.set_particles(cloud)
        pf.seed_action(777)
        traj[f"{variant}_init"] = cloud
        t = 1_000_000
        pf.update(g, synth.make_pose(*pose, utime=t), ref.Scan(*synth.make_scan(grid_spec, pose, seed=0, t0=t - 100_000)),
                  seed=1, action_utime=900_000)      # first call only latches the odometry (no motion)
        for step in range(4):
            pose = synth.odometry_step(rng, pose, step=(0.05, 0.02, 0.03))
            t += 100_000
            r_, th_, t_ = synth.make_scan(grid_spec, pose, seed=step + 1, t0=t - 100_000)
            odom = synth.make_pose(*pose, utime=t)
            autime = t if variant == "interp" else 900_000
            est, moved, draws, _ = pf.update(g, odom, ref.Scan(r_, th_, t_), seed=step + 1, action_utime=autime)
            traj[f"{variant}_{step}_odom"] = np.array(odom)
            traj[f"{variant}_{step}_ranges"], traj[f"{variant}_{step}_thetas"], traj[f"{variant}_{step}_times"] = r_, th_, t_
            traj[f"{variant}_{step}_r"] = ref.resample_draw(step + 1, n)
            traj[f"{variant}_{step}_draws"] = draws
            traj[f"{variant}_{step}_moved"] = moved
            traj[f"{variant}_{step}_estimate"] = np.array(est)
            traj[f"{variant}_{step}_particles"] = pf.particles()
    np.savez_compressed(os.path.join(OUT, "trajectory.npz"), **traj)
    distance_golden()
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
