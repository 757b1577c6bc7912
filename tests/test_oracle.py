"""CPU tests: the plain-C restatement (oracle/mcl_oracle.c) against golden vectors produced by the compiled,
unmodified reference (tests/golden/make_golden.py), and against the live reference library where it is present."""
import numpy as np
import pytest

from conftest import load_golden, synth_grid_from_golden
from oracle import port, ref
from botlab_b200 import synth


def port_grid(spec):
    return port.Grid(spec.cells, spec.origin_x, spec.origin_y, spec.cells_per_meter)


def same_particles(a, b):
    """Field-wise bit equality (the 4 padding bytes after each pose are indeterminate in the reference's structs)."""
    for k in ("pose", "parent_pose"):
        for f, t in (("utime", np.int64), ("x", np.uint32), ("y", np.uint32), ("theta", np.uint32)):
            if not np.array_equal(np.ascontiguousarray(a[k][f]).view(t), np.ascontiguousarray(b[k][f]).view(t)):
                return False
    return np.array_equal(a["weight"].view(np.uint64), b["weight"].view(np.uint64))


def test_layouts():
    assert port.POSE_DTYPE.itemsize == 24 and port.PARTICLE_DTYPE.itemsize == 56


def test_kat_sensor_scores(real_map):
    k = load_golden("kat")
    s, gathers, evals = port.likelihood(port_grid(real_map), k["particles"], k["ranges"], k["thetas"], k["times"])
    assert np.array_equal(s, k["scores"])
    assert np.array_equal(k["scores"], [3138.0, 1322.5, 1787.5, 2969.0])
    assert evals == 4 * 360 and 4 * 360 <= gathers <= 3 * 4 * 360


def test_kat_moving_scan():
    k = load_golden("kat")
    p = k["particles"]
    rb = port.moving_scan(k["ranges"], k["thetas"], k["times"], p["parent_pose"][1], p["pose"][1])
    rd = port.moving_scan(k["ranges"], k["thetas"], k["times"], p["parent_pose"][3], p["pose"][3])
    assert np.array_equal(rb.view(np.uint32), k["rays_b"].view(np.uint32))
    assert np.array_equal(rd.view(np.uint32), k["rays_d"].view(np.uint32))
    assert len(rd) == 360 and np.isclose(rd[359][3], 3.11791658)                    # B7


def test_kat_action():
    k = load_golden("kat")
    am = port.ActionModel()
    am.update(synth.make_pose(0, 0, 0))
    moved, params = am.update(synth.make_pose(0.02, 0.01, 0.01))
    assert moved and np.array_equal(params, k["action_forward"])                    # B8
    out = am.apply(k["action_in"], k["action_draws"])
    for f in ("x", "y", "theta"):
        assert np.array_equal(out["pose"][f], k["action_out"]["pose"][f])           # B9
    am2 = port.ActionModel()
    am2.update(synth.make_pose(0, 0, 0))
    _, back = am2.update(synth.make_pose(-0.02, 0, 0))
    assert np.array_equal(back, k["action_backward"])                               # B10


def test_kat_mt19937_and_normal():
    rng = port.Rng(5489)
    assert rng.next_u32() == 3499211612
    rng = port.Rng(5489)
    assert rng.normal(0.0, 1.0) == 0.13452965847232812
    # the draws the reference's applyAction made in B9 (default-seeded generator, after updateAction of B8)
    k = load_golden("kat")
    am = port.ActionModel()
    am.update(synth.make_pose(0, 0, 0))
    am.update(synth.make_pose(0.02, 0.01, 0.01))
    assert np.array_equal(am.draws(port.Rng(5489), 2), k["action_draws"])


def test_kat_resample_and_estimate():
    k = load_golden("kat")
    ps = k["resample_particles"]
    assert k["resample_r"] == 0.10502346464433869                                   # B11
    idx, over = port.resample(ps["weight"], float(k["resample_r"]))
    assert over == 0 and np.array_equal(idx, k["resample_idx"]) and list(idx) == [1, 1, 3, 3, 4, 5, 6, 7]
    est = port.estimate(ps)
    g = k["estimate"]
    assert est["x"] == g["x"] and est["y"] == g["y"] and est["theta"] == g["theta"]  # B12


@pytest.mark.parametrize("name", ["real", "synth"])
@pytest.mark.parametrize("variant", ["interp", "degen"])
def test_sensor_golden(name, variant, real_map, sensor_golden):
    sg = sensor_golden
    grid = real_map if name == "real" else synth_grid_from_golden(sg)
    s, gathers, evals = port.likelihood(port_grid(grid), sg[f"{name}_{variant}_particles"], sg[f"{name}_ranges"],
                                        sg[f"{name}_thetas"], sg[f"{name}_times"])
    assert np.array_equal(s, sg[f"{name}_{variant}_scores"])
    assert (s * 2 == np.round(s * 2)).all()           # scores are exact multiples of 0.5
    assert s[5:].max() > 0


def test_action_golden():
    a = load_golden("action")
    am = port.ActionModel()
    am.update(synth.make_pose(0.3, -0.2, 0.1))
    moved, params = am.update(synth.make_pose(0.32, -0.19, 0.11))
    assert moved == bool(a["moved"]) and np.array_equal(params, a["params"])
    out = am.apply(a["particles_in"], a["draws"], utime=int(a["utime"]))
    assert same_particles(out, a["particles_out"])
    # and the draws themselves from the libstdc++ restatement (generator seeded 12345)
    assert np.array_equal(am.draws(port.Rng(12345), 512), a["draws"])


@pytest.mark.parametrize("n", [200, 4096, 100_000])
def test_resample_golden(n):
    g = load_golden("resample")
    idx, over = port.resample(g[f"w_{n}"], float(g[f"r_{n}"]))
    assert over == 0 and np.array_equal(idx, g[f"idx_{n}"])


def test_normalize_estimate_golden(sensor_golden):
    g = load_golden("normalize")
    grid = synth_grid_from_golden(sensor_golden)
    sg = sensor_golden
    s, _, _ = port.likelihood(port_grid(grid), g["proposal"], sg["synth_ranges"], sg["synth_thetas"], sg["synth_times"])
    w, wsum = port.normalize(s)
    assert np.array_equal(w, g["posterior"]["weight"])
    est = port.estimate(g["posterior"])
    for f in ("x", "y", "theta"):
        assert est[f] == g["estimate"][f]


@pytest.mark.parametrize("variant", ["interp", "legacy"])
def test_trajectory_golden(variant, real_map):
    t = load_golden("trajectory")
    pf = port.ParticleFilter(t[f"{variant}_init"])
    grid = port_grid(real_map)
    # first reference call only latched the odometry at pose (0,0,0)
    pf.action.update(synth.make_pose(0.0, 0.0, 0.0, utime=1_000_000))
    for step in range(4):
        odom = t[f"{variant}_{step}_odom"]
        autime = int(odom["utime"]) if variant == "interp" else 900_000
        est, moved = pf.update(grid, odom, t[f"{variant}_{step}_ranges"], t[f"{variant}_{step}_thetas"],
                               t[f"{variant}_{step}_times"], float(t[f"{variant}_{step}_r"]),
                               t[f"{variant}_{step}_draws"], action_utime=autime)
        assert moved == bool(t[f"{variant}_{step}_moved"])
        assert same_particles(pf.particles, t[f"{variant}_{step}_particles"])
        g = t[f"{variant}_{step}_estimate"]
        assert est["x"] == g["x"] and est["y"] == g["y"] and est["theta"] == g["theta"] and est["utime"] == g["utime"]


# ---------------------------------------------------------------- live cross-checks where oracle/_ref was built
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_port_vs_reference_sensor_random(seed):
    rng = np.random.default_rng(seed)
    grid = synth.make_map(240, seed=seed)
    truth = synth.find_free_pose(grid, rng)
    r, th, t = synth.make_scan(grid, truth, num_beams=360 + 17 * seed, seed=seed)
    p = synth.make_particles(2000, truth, seed=seed, parent_utime=int(t[0]), pose_utime=int(t[-1]))
    u = synth.make_uniform_particles(2000, grid, seed=seed, utime=int(t[-1]))
    rg = ref.RefGrid.from_cells(grid.cells, grid.origin_x, grid.origin_y, grid.meters_per_cell)
    for cloud in (p, u):
        a = ref.likelihood(rg, cloud, ref.Scan(r, th, t))
        b, _, _ = port.likelihood(port_grid(grid), cloud, r, th, t)
        assert np.array_equal(a, b)


@needs_ref
def test_port_vs_reference_resample_large():
    n = 1_000_000
    w = synth.filter_shaped_weights(n, seed=42)
    pf = ref.RefParticleFilter(n)
    ps = np.zeros(n, ref.PARTICLE_DTYPE)
    ps["weight"] = w
    pf.set_particles(ps)
    idx, over = port.resample(w, ref.resample_draw(1, n))
    assert over == 0 and np.array_equal(idx, pf.resample(1))


# ---- Mapping::updateMap: the C restatement against the executed reference --------------------------
def test_map_update_golden(real_map):
    g = load_golden("mapping")
    cells = g["start_cells"].copy()
    for k in range(int(g["steps"])):
        cells = port.map_update(cells, real_map.origin_x, real_map.origin_y, real_map.cells_per_meter,
                                g[f"{k}_previous"], g[f"{k}_pose"], k > 0, g[f"{k}_ranges"], g[f"{k}_thetas"],
                                g[f"{k}_times"], float(g["max_laser"]), int(g["hit"]), int(g["miss"]))
        assert np.array_equal(cells, g[f"{k}_cells"]), k
        if k == 0:
            assert np.array_equal(cells, g["start_cells"])         # first call: initialized_ is false, no cell changes
    assert (cells != g["start_cells"]).sum() > 500
    assert (cells == 127).sum() > (g["start_cells"] == 127).sum() and (cells == -128).sum() > 100     # saturation reached


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("hit,miss,max_laser", [(3, 1, 5.0), (127, 100, 8.0), (1, 50, 3.0)])
def test_map_update_matches_live_reference(hit, miss, max_laser):
    grid = synth.make_map(300, seed=31)
    rng = np.random.default_rng(hit)
    pose = synth.find_free_pose(grid, rng)
    cells = grid.cells.copy()
    g = ref.RefGrid.from_cells(cells, grid.origin_x, grid.origin_y, grid.meters_per_cell)
    prev, t0 = pose, 1_000_000
    for k in range(5):
        r, th, t = synth.make_scan(grid, pose, seed=40 + k, t0=t0)
        prv = synth.make_pose(*prev, utime=int(t[0]) if k % 2 == 0 else int(t[-1]))
        cur = synth.make_pose(*pose, utime=int(t[-1]))
        cells = port.map_update(cells, grid.origin_x, grid.origin_y, grid.cells_per_meter, prv, cur, k > 0, r, th, t,
                                max_laser, hit, miss)
        ref.map_update(g, prv, cur, k > 0, ref.Scan(r, th, t), max_laser, hit, miss)
        assert np.array_equal(cells, g.cells()), k
        prev, pose, t0 = pose, synth.odometry_step(rng, pose), t0 + 100_000


@needs_ref
@pytest.mark.parametrize("case", [0, 1, 2, 3, 4, 5, 6, 8, 9])
def test_port_vs_reference_sensor_geometries(case):
    """The geometries of tests/test_gpu_parity.py::test_two_pass_randomized_geometry (resolutions, non-square maps, far
    origins, long ranges, beam counts): the C restatement the GPU is compared with equals the executed reference there
    too.  (Case 7 uses a stale cellsPerMeter, which the reference's constructor cannot produce.)"""
    rng = np.random.default_rng(1000 + case)
    mpc = [0.05, 0.025, 0.1, 0.05, 0.05][case % 5]
    w, h = int(rng.integers(150, 900)), int(rng.integers(150, 900))
    base = synth.make_map(max(w, h), seed=50 + case, meters_per_cell=mpc)
    cells = base.cells[:h, :w].copy()
    cells[-2:, :] = 100; cells[:, -2:] = 100
    ox, oy = [(-w * mpc / 2, -h * mpc / 2), (731.25, -412.5), (0.0, 0.0), (-2000.0, 1500.0), (5.5, 5.5)][case % 5]
    grid = synth.GridSpec(cells, ox, oy, mpc)
    truth = synth.find_free_pose(grid, rng)
    nb = int(rng.choice([180, 290, 360, 500, 720]))
    r, th, t = synth.make_scan(grid, truth, num_beams=nb, seed=case, max_range=float(rng.choice([4.0, 8.0, 12.0])))
    cloud = synth.make_particles(1500, truth, seed=case, sigma_xy=0.3, sigma_theta=0.5, parent_utime=int(t[0]),
                                 pose_utime=int(t[-1]))
    rg = ref.RefGrid.from_cells(grid.cells, grid.origin_x, grid.origin_y, grid.meters_per_cell)
    info = rg.info()
    assert info["cells_per_meter"] == np.float32(grid.cells_per_meter)
    a = ref.likelihood(rg, cloud, ref.Scan(r, th, t))
    b, _, _ = port.likelihood(port_grid(grid), cloud, r, th, t)
    assert np.array_equal(a, b)


# ---- obstacle distance grid (next row: likelihood-field sensor mode) ------------------------------------------------------
def _distance_cases():
    rng = np.random.default_rng(77)
    a = synth.make_map(120, seed=5).cells
    b = np.where(rng.random((90, 140)) < 0.03, 60, np.where(rng.random((90, 140)) < 0.5, -9, 0)).astype(np.int8)
    c = np.full((40, 60), -5, np.int8)                    # all free: the brushfire never starts
    d = np.full((30, 30), -5, np.int8); d[11, 17] = 0     # one unknown cell is a source too (log-odds >= 0)
    return {"synth": a, "random": b, "all_free": c, "one_unknown": d}


@pytest.mark.parametrize("name", ["synth", "random", "all_free", "one_unknown"])
def test_distance_grid_equals_the_compiled_reference(name):
    """oracle orc_distance_grid == ObstacleDistanceGrid::setDistances of the executed reference, bit for bit (the 0.1f
    accumulation included), on the committed fixture and -- where oracle/_ref is built -- live."""
    cells = _distance_cases()[name]
    got = port.distance_grid(cells)
    gold = load_golden("distance")
    assert np.array_equal(gold[name + "_cells"], cells)
    assert np.array_equal(got.view(np.uint32), gold[name + "_dist"].view(np.uint32))
    if name == "all_free":
        assert (got == -1.0).all()
    if name == "one_unknown":
        d = np.float32(0.0)
        for _ in range(11 + 17):                             # (0, 0) is 28 four-connected steps from the source
            d = np.float32(d + np.float32(0.1))
        assert got[11, 17] == 0.0 and got[11, 18] == np.float32(0.1) and got[0, 0] == d
    if ref.available():
        g = ref.RefGrid.from_cells(cells, 0.0, 0.0, 0.05)
        assert np.array_equal(ref.distance_grid(g).view(np.uint32), got.view(np.uint32))
