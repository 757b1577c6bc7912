"""GPU parity tests: the CUDA engine, called through the C ABI (botlab_b200/libmcl_cuda.so), against
 (a) golden vectors produced by the compiled, unmodified reference (tests/golden/), and
 (b) the oracle (oracle/mcl_oracle.c; and oracle/_ref where it travelled) on seeded inputs,
 plus size-independent properties at BASELINE.json's full sizes.

Bars (BASELINE.json north_star): resample indices bit-exact; action poses within 1e-5 m / 1e-5 rad with injected draws;
per-particle scores within 1e-5 relative (asserted EXACTLY here: they are multiples of 0.5); weights within 1e-6 abs."""
import numpy as np
import pytest

from conftest import load_golden, synth_grid_from_golden
from botlab_b200 import engine, synth
from oracle import port

pytestmark = pytest.mark.gpu

POSE_TOL = 1e-5
WEIGHT_TOL = 1e-6


def port_grid(spec):
    return port.Grid(spec.cells, spec.origin_x, spec.origin_y, spec.cells_per_meter)


def make_engine(n, grid, **params):
    e = engine.Engine(n, **params)
    e.set_map(grid.cells, grid.origin_x, grid.origin_y, grid.meters_per_cell, grid.cells_per_meter)
    return e


def angle_err(a, b):
    d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
    return np.minimum(d, 2 * np.pi - d)


# ------------------------------------------------------------------------------------------------ sensor model
def test_sensor_known_answers(real_map):
    k = load_golden("kat")
    for i, expect in enumerate([3138.0, 1322.5, 1787.5, 2969.0]):
        p = np.concatenate([k["particles"][i:i + 1]] * 2)
        e = make_engine(2, real_map)
        e.import_particles(p)
        s = e.score(k["ranges"], k["thetas"], k["times"])
        assert list(s) == [expect, expect]
        e.close()


@pytest.mark.parametrize("name", ["real", "synth"])
@pytest.mark.parametrize("variant", ["interp", "degen"])
@pytest.mark.parametrize("lanes,tile", [(0, 0), (32, 1), (1, 2), (8, 2), (4, 1)])
def test_sensor_golden(name, variant, lanes, tile, real_map, sensor_golden):
    sg = sensor_golden
    grid = real_map if name == "real" else synth_grid_from_golden(sg)
    p = sg[f"{name}_{variant}_particles"].copy()
    if tile == 2:
        # a forced tile needs a bounded cloud: keep the hostile headings, drop the far-away / NaN positions
        p = p[np.r_[2:4, 5:len(p)]]
    e = make_engine(len(p), grid, lanes_per_particle=lanes, map_tile=tile)
    e.import_particles(p)
    e.set_gather_counting(True)
    s = e.score(sg[f"{name}_ranges"], sg[f"{name}_thetas"], sg[f"{name}_times"])
    want, gathers, evals = port.likelihood(port_grid(grid), p, sg[f"{name}_ranges"], sg[f"{name}_thetas"],
                                           sg[f"{name}_times"])
    if tile != 2:
        assert np.array_equal(want, sg[f"{name}_{variant}_scores"])        # oracle == reference golden
    assert np.array_equal(s, want)                                        # engine == oracle, bit for bit
    st = e.stats()
    assert st["evals"] == evals and st["gathers"] == gathers              # algorithmic work agrees with the oracle
    assert st["map_tile_used"] in ((2, 4) if tile == 2 else (st["map_tile_used"],))    # 4: the score-table pass's own tile
    e.close()


@pytest.mark.parametrize("seed", [1, 2])
def test_sensor_random_vs_oracle(seed):
    rng = np.random.default_rng(seed)
    grid = synth.make_map(400, seed=seed)
    truth = synth.find_free_pose(grid, rng)
    r, th, t = synth.make_scan(grid, truth, num_beams=360 + 13 * seed, seed=seed)
    clouds = [synth.make_particles(30_000, truth, seed=seed, parent_utime=int(t[0]), pose_utime=int(t[-1])),
              synth.make_uniform_particles(30_000, grid, seed=seed, utime=int(t[-1]))]
    for cloud in clouds:
        e = make_engine(len(cloud), grid)
        e.import_particles(cloud)
        s = e.score(r, th, t)
        want, _, _ = port.likelihood(port_grid(grid), cloud, r, th, t)
        assert np.array_equal(s, want)
        e.close()


def test_sensor_ragged_inputs(real_map):
    """Empty scan, all-invalid scan, one beam, 720 beams."""
    p = synth.make_particles(64, (0.0, 0.0, 0.3), seed=1)
    e = make_engine(64, real_map)
    e.import_particles(p)
    z = np.zeros(0, np.float32)
    assert np.array_equal(e.score(z, z, np.zeros(0, np.int64)), np.zeros(64))
    assert np.array_equal(e.score(np.full(10, 0.1, np.float32), np.zeros(10, np.float32), np.zeros(10, np.int64)),
                          np.zeros(64))
    for nb in (1, 720):
        r, th, t = synth.make_scan(real_map, (0.0, 0.0, 0.3), num_beams=nb, seed=nb)
        want, _, _ = port.likelihood(port_grid(real_map), p, r, th, t)
        assert np.array_equal(e.score(r, th, t), want)
    e.close()


def test_map_rect_update_equals_full_upload(real_map):
    rng = np.random.default_rng(5)
    p = synth.make_particles(2000, (0.0, 0.0, 0.0), seed=2)
    r, th, t = synth.make_scan(real_map, (0.0, 0.0, 0.0), seed=3)
    e = make_engine(2000, real_map)
    e.import_particles(p)
    cells = real_map.cells.copy()
    patch = rng.integers(-128, 128, (37, 53)).astype(np.int8)
    cells[60:97, 71:124] = patch
    e.update_map_rect(71, 60, patch)
    changed = synth.GridSpec(cells, real_map.origin_x, real_map.origin_y, real_map.meters_per_cell,
                             real_map.cells_per_meter)
    want, _, _ = port.likelihood(port_grid(changed), p, r, th, t)
    assert np.array_equal(e.score(r, th, t), want)
    e.close()


# ------------------------------------------------------------------------------------------------ action model
def test_action_golden_injected_draws(real_map):
    a = load_golden("action")
    am = engine.ActionModel()
    am.update(0.3, -0.2, 0.1)
    assert am.update(0.32, -0.19, 0.11)
    e = engine.Engine(512)
    e.import_particles(a["particles_in"])
    e.apply_action(am, utime=int(a["utime"]), noise=a["draws"])
    out = e.export_particles()
    want = a["particles_out"]
    assert np.abs(out["pose"]["x"].astype(np.float64) - want["pose"]["x"]).max() <= POSE_TOL
    assert np.abs(out["pose"]["y"].astype(np.float64) - want["pose"]["y"]).max() <= POSE_TOL
    assert angle_err(out["pose"]["theta"], want["pose"]["theta"]).max() <= POSE_TOL
    # parent_pose = the old pose, utime = the planted ActionModel::utime_
    for f in ("x", "y", "theta"):
        assert np.array_equal(out["parent_pose"][f], want["parent_pose"][f])
    assert (out["pose"]["utime"] == want["pose"]["utime"]).all()
    assert (out["parent_pose"]["utime"] == want["parent_pose"]["utime"]).all()
    # in practice the poses are bit-identical except where libm's double sin/cos differs in the last place
    same = sum(np.array_equal(out["pose"][f], want["pose"][f]) for f in ("x", "y", "theta"))
    assert same >= 1
    e.close()


def test_action_philox_statistics():
    n = 400_000
    am = engine.ActionModel()
    am.update(0.0, 0.0, 0.0)
    am.update(0.10, 0.0, 0.0)
    e = engine.Engine(n)
    p = np.zeros(n, engine.PARTICLE_DTYPE)
    p["weight"] = 1.0 / n
    e.import_particles(p)
    e.apply_action(am, utime=5)
    out = e.export_particles()
    # trans ~ N(0.1, 0.005), rot1, rot2 ~ N(0, 0.05): E[x] ~ 0.1*E[cos r1], theta std ~ sqrt(2)*0.05
    assert abs(out["pose"]["x"].mean() - 0.1 * np.exp(-0.05 ** 2 / 2)) < 2e-4
    assert abs(out["pose"]["theta"].std() - np.sqrt(2) * 0.05) < 1e-3
    assert abs(out["pose"]["y"].mean()) < 2e-4
    assert (out["parent_pose"]["x"] == 0).all() and (out["pose"]["utime"] == 5).all()
    e2 = engine.Engine(n)
    e2.import_particles(p)
    e2.apply_action(am, utime=5)
    assert np.array_equal(e2.export_particles()["pose"]["x"], out["pose"]["x"])     # counter-based => reproducible
    e.close(); e2.close()


def test_init_at_pose():
    n = 200_000
    e = engine.Engine(n)
    e.init_at_pose(1.5, -2.0, 3.14, utime=77, seed=9)
    p = e.export_particles()
    assert abs(p["pose"]["x"][:-1].mean() - 1.5) < 1e-4 and abs(p["pose"]["x"][:-1].std() - 0.01) < 1e-4
    assert abs(p["pose"]["y"][:-1].std() - 0.01) < 1e-4
    assert np.abs(p["pose"]["theta"]).max() <= np.float32(np.pi)
    assert (p["pose"]["x"][-1], p["pose"]["y"][-1], p["pose"]["theta"][-1]) == (np.float32(1.5), np.float32(-2.0),
                                                                               np.float32(3.14))   # :33
    assert (p["weight"] == 1.0 / n).all() and (p["pose"]["utime"] == 77).all()
    e.close()


# ------------------------------------------------------------------------------------------------ resampling
@pytest.mark.parametrize("n", [200, 4096, 100_000])
def test_resample_golden_bit_exact(n):
    g = load_golden("resample")
    e = engine.Engine(n)
    e.init_at_pose(0, 0, 0)
    idx = e.resample(float(g[f"r_{n}"]), weights=g[f"w_{n}"])
    assert np.array_equal(idx, g[f"idx_{n}"])
    assert e.stats()["resample_overruns"] == 0
    e.close()


def test_resample_known_answer():
    k = load_golden("kat")
    e = engine.Engine(8)
    e.import_particles(k["resample_particles"])
    idx = e.resample(float(k["resample_r"]))
    assert list(idx) == [1, 1, 3, 3, 4, 5, 6, 7]
    out = e.export_particles()
    assert np.array_equal(out["pose"]["x"], k["resample_particles"]["pose"]["x"][idx])
    assert np.array_equal(out["weight"], k["resample_particles"]["weight"][idx])   # prior[m] = posterior_[i] (:100)
    e.close()


@pytest.mark.parametrize("n,seed", [(1_000_000, 1), (1_000_003, 2), (16_000_000, 3)])
def test_resample_large_vs_oracle_bit_exact(n, seed):
    """At these sizes a plain parallel double scan (and an exact one) disagrees with the sequential reference
    (SURVEY 7.2 hard part 1); the engine must not."""
    w = synth.filter_shaped_weights(n, seed=seed)
    r = 0.8401877171547095 / n                                                  # first unseeded rand()/RAND_MAX
    want, over = port.resample(w, r)
    e = engine.Engine(n)
    e.init_at_pose(0, 0, 0)
    idx = e.resample(r, weights=w)
    st = e.stats()
    assert st["resample_overruns"] == over
    assert np.array_equal(idx, want)
    e.close()


@pytest.mark.parametrize("kind", ["uniform", "one_hot", "tiny_tail", "zeros_inside", "unnormalised", "denormal"])
def test_resample_edge_weights(kind):
    n = 50_000
    rng = np.random.default_rng(3)
    if kind == "uniform":
        w = np.full(n, 1.0 / n)
    elif kind == "one_hot":
        w = np.zeros(n); w[n // 3] = 1.0
    elif kind == "tiny_tail":
        w = rng.random(n); w[n // 2:] *= 1e-18; w /= w.sum()
    elif kind == "zeros_inside":
        w = rng.random(n); w[rng.random(n) < 0.5] = 0.0; w /= w.sum()
    elif kind == "unnormalised":
        w = rng.random(n) * 37.0
    else:
        w = rng.random(n); w[:100] = 5e-324; w /= w.sum()
    r = 0.3 / n
    want, over = port.resample(w, r)
    e = engine.Engine(n)
    e.init_at_pose(0, 0, 0)
    idx = e.resample(r, weights=w)
    assert np.array_equal(idx, want)
    assert e.stats()["resample_overruns"] == over
    e.close()


# ------------------------------------------------------------------------------------------------ normalise / estimate
def test_normalize_and_estimate_golden(sensor_golden):
    g = load_golden("normalize")
    sg = sensor_golden
    grid = synth_grid_from_golden(sg)
    e = make_engine(len(g["proposal"]), grid)
    e.import_particles(g["proposal"])
    e.score(sg["synth_ranges"], sg["synth_thetas"], sg["synth_times"], want_scores=False)
    w = e.normalize()
    assert np.abs(w - g["posterior"]["weight"]).max() <= WEIGHT_TOL
    assert np.array_equal(w, g["posterior"]["weight"])          # the sequential-sum emulation makes them identical
    est = e.estimate()
    assert abs(est.x - float(g["estimate"]["x"])) <= POSE_TOL
    assert abs(est.y - float(g["estimate"]["y"])) <= POSE_TOL
    assert angle_err(est.theta, float(g["estimate"]["theta"])) <= POSE_TOL
    st = e.stats()
    scores, _, _ = port.likelihood(port_grid(grid), g["proposal"], sg["synth_ranges"], sg["synth_thetas"],
                                   sg["synth_times"])
    assert st["weight_sum"] == port.normalize(scores)[1]       # wSum with the reference's sequential rounding
    assert 1.0 <= st["effective_sample_size"] <= len(w)
    e.close()


def test_estimate_known_answer():
    k = load_golden("kat")
    e = engine.Engine(8)
    e.import_particles(k["resample_particles"])
    est = e.estimate()
    assert abs(est.x - 3.3499999) <= POSE_TOL and abs(est.y - 0.335000008) <= POSE_TOL
    assert angle_err(est.theta, 3.09885478) <= POSE_TOL
    e.close()


# ------------------------------------------------------------------------------------------------ fused update
@pytest.mark.parametrize("variant", ["interp", "legacy"])
def test_update_trajectory_golden(variant, real_map):
    """Four ParticleFilter::updateFilter calls on the real map with the reference's rand() draw and mt19937 action draws
    injected: particles, weights and estimates follow the reference."""
    t = load_golden("trajectory")
    n = len(t[f"{variant}_init"])
    e = make_engine(n, real_map, legacy_equal_utime=1 if variant == "legacy" else 0)
    e.import_particles(t[f"{variant}_init"])
    am = engine.ActionModel()
    assert am.update(0.0, 0.0, 0.0, 1_000_000) is False
    for step in range(4):
        odom = t[f"{variant}_{step}_odom"]
        moved = am.update(float(odom["x"]), float(odom["y"]), float(odom["theta"]), int(odom["utime"]))
        assert moved == bool(t[f"{variant}_{step}_moved"])
        est = e.update(am, int(odom["utime"]), t[f"{variant}_{step}_ranges"], t[f"{variant}_{step}_thetas"],
                       t[f"{variant}_{step}_times"], float(t[f"{variant}_{step}_r"]), noise=t[f"{variant}_{step}_draws"])
        want = t[f"{variant}_{step}_particles"]
        got = e.export_particles()
        # identical resample indices <=> identical parent poses (they are copies of the chosen particles)
        for f in ("x", "y", "theta"):
            assert np.array_equal(got["parent_pose"][f], want["parent_pose"][f]), (step, f)
        assert np.abs(got["pose"]["x"].astype(np.float64) - want["pose"]["x"]).max() <= POSE_TOL
        assert np.abs(got["pose"]["y"].astype(np.float64) - want["pose"]["y"]).max() <= POSE_TOL
        assert angle_err(got["pose"]["theta"], want["pose"]["theta"]).max() <= POSE_TOL
        assert np.abs(got["weight"] - want["weight"]).max() <= WEIGHT_TOL
        g = t[f"{variant}_{step}_estimate"]
        assert abs(est.x - float(g["x"])) <= POSE_TOL and abs(est.y - float(g["y"])) <= POSE_TOL
        assert angle_err(est.theta, float(g["theta"])) <= POSE_TOL and est.utime == int(g["utime"])
        if variant == "interp":
            assert (got["pose"]["utime"] == int(odom["utime"])).all()
    st = e.stats()
    assert st["updates"] == 4 and st["kernel_launches"] > 0 and st["ms_total"] > 0
    e.close()


def test_update_without_motion_returns_previous_estimate(real_map):
    e = make_engine(1000, real_map)
    e.init_at_pose(0.5, 0.25, 0.1, utime=10, seed=4)
    am = engine.ActionModel()
    am.update(0.5, 0.25, 0.1, 10)
    assert am.update(0.5, 0.25, 0.1, 20) is False
    r, th, t = synth.make_scan(real_map, (0.5, 0.25, 0.1))
    before = e.export_particles()
    est = e.update(am, 20, r, th, t, 0.5 / 1000)
    assert (est.x, est.y, est.theta, est.utime) == (np.float32(0.5), np.float32(0.25), np.float32(0.1), 20)  # :50
    assert e.export_particles().tobytes() == before.tobytes()
    e.close()


def test_update_action_only(real_map):
    a = load_golden("action")
    am = engine.ActionModel()
    am.update(0.3, -0.2, 0.1)
    am.update(0.32, -0.19, 0.11)
    e = engine.Engine(512)
    e.import_particles(a["particles_in"])
    e.update_action_only(am, int(a["utime"]), noise=a["draws"])
    out = e.export_particles()
    assert np.abs(out["pose"]["x"].astype(np.float64) - a["particles_out"]["pose"]["x"]).max() <= POSE_TOL
    assert np.array_equal(out["weight"], a["particles_in"]["weight"])          # no resample, no reweighting (:54-65)
    e.close()


def test_export_stride_and_cap():
    n = 10_000
    p = synth.make_particles(n, (0, 0, 0), seed=8)
    e = engine.Engine(n)
    e.import_particles(p)
    sub = e.export_particles(stride=7)
    assert len(sub) == -(-n // 7) and np.array_equal(sub["pose"]["x"], p["pose"]["x"][::7])
    assert len(e.export_particles(max_n=100, stride=3)) == 100
    e.close()


def test_export_weighted_subsample(real_map):
    """a weighted sub-sample for SLAM_PARTICLES consumers.  mcl_export_weighted draws `count` particles by
    systematic sampling over the weights' sequential running sum; the picks equal a numpy restatement of the rule."""
    n, count, u01 = 50_000, 777, 0.3125
    cloud = synth.make_particles(n, (0.5, -0.25, 0.3), seed=21)
    w = synth.filter_shaped_weights(n, seed=21)
    cloud["weight"] = w
    e = make_engine(n, real_map)
    e.import_particles(cloud)
    got = e.export_weighted(count, u01)
    cum = np.cumsum(w)                                   # sequential double sum, like the reference's loop
    step = 1.0 / count
    u = (u01 / count) + np.arange(count, dtype=np.float64) * step
    idx = np.minimum(np.searchsorted(cum, u, side="left"), n - 1)
    assert len(got) == count and np.array_equal(got["weight"], np.full(count, 1.0 / count))
    for k in ("pose", "parent_pose"):
        for f in ("x", "y", "theta"):
            assert np.array_equal(got[k][f], cloud[k][f][idx]), (k, f)
    assert (np.diff(idx) >= 0).all() and len(np.unique(idx)) > count // 2
    e.close()


def test_import_rejects_mixed_utimes():
    p = synth.make_particles(100, (0, 0, 0), seed=8)
    p["pose"]["utime"][50] += 1
    e = engine.Engine(100)
    with pytest.raises(engine.MclError, match="utime"):
        e.import_particles(p)
    e.close()


# ------------------------------------------------------------------------------------------------ full-size properties
@pytest.mark.parametrize("config", ["config3"])
def test_full_size_properties(config):
    """BASELINE configs at full size: size-independent properties + a sub-sample against the oracle."""
    n, side = synth.CONFIGS[config]
    rng = np.random.default_rng(17)
    grid = synth.make_map(side, seed=synth.MAP_SEED + 3)
    truth = synth.find_free_pose(grid, rng)
    r, th, t = synth.make_scan(grid, truth, seed=17)
    nb_valid = int((r > np.float32(0.15)).sum())
    e = make_engine(n, grid)
    e.init_at_pose(*truth, utime=int(t[0]) - 100_000, seed=17)
    am = engine.ActionModel()
    am.update(*truth, int(t[0]) - 100_000)
    moved = am.update(truth[0] + 0.02, truth[1] + 0.01, truth[2] + 0.01, int(t[-1]))
    assert moved
    rdraw = 0.8401877171547095 / n
    est = e.update(am, int(t[-1]), r, th, t, rdraw)
    st = e.stats()
    assert st["evals"] == n * nb_valid and st["resample_overruns"] == 0
    sub = e.export_particles(stride=997)
    scores, _, _ = port.likelihood(port_grid(grid), sub, r, th, t)
    w = np.maximum(scores, 0.001) / st["weight_sum"]
    assert np.abs(w - sub["weight"]).max() <= WEIGHT_TOL and np.array_equal(w, sub["weight"])
    assert abs(est.x - truth[0]) < 0.2 and abs(est.y - truth[1]) < 0.2
    # second update: resampling from the weights just computed; indices sorted, counts follow the weights
    am.update(truth[0] + 0.04, truth[1] + 0.02, truth[2] + 0.02, int(t[-1]) + 100_000)
    wfull = e.normalize()                      # idempotent: same scores -> same weights
    assert abs(wfull.sum() - 1.0) < 1e-9
    idx = e.resample(rdraw)
    assert (np.diff(idx) >= 0).all()
    counts = np.bincount(idx, minlength=n)
    assert np.abs(counts - wfull * n).max() <= 1.0 + 1e-6      # systematic resampling: |count - N w| < 1
    want, _ = port.resample(wfull, rdraw)
    assert np.array_equal(idx, want)
    e.close()


# ------------------------------------------------------------------------------------------------ config 1 replay
@pytest.mark.parametrize("legacy", [0, 1])
def test_config1_replay_follows_the_reference_filter(legacy, real_map):
    """BASELINE configs[0] in miniature: `slam --localization-only` shape -- the real 10 m x 10 m map, the reference's
    default 200 particles (slam_main.cpp:21), 40 scans along a synthesised trajectory (the .log is not in the checkout).
    The oracle filter draws its action noise from the libstdc++ mt19937 restatement exactly like the reference
    (default seed 5489) and its resampling offset from the reference's unseeded rand() sequence; the engine gets the
    same draws injected and must follow it particle for particle."""
    n = 200
    rng = np.random.default_rng(2026)
    pose = (0.0, 0.0, 0.0)
    t = 1_000_000
    cloud = port.init_at_pose(port.Rng(99), synth.make_pose(*pose, utime=t), n)
    pf = port.ParticleFilter(cloud)
    mt = port.Rng(5489)                               # ActionModel's default-constructed std::mt19937
    e = make_engine(n, real_map, legacy_equal_utime=legacy)
    e.import_particles(cloud)
    am = engine.ActionModel()
    grid = port_grid(real_map)
    first = synth.make_pose(*pose, utime=t)
    pf.action.update(first)
    am.update(*pose, t)
    # glibc's rand() after srand(1): the first values of the sequence the reference consumes (particle_filter.cpp:92)
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    worst_pose, worst_w = 0.0, 0.0
    for step in range(40):
        pose = synth.odometry_step(rng, pose, step=(0.03 * np.cos(pose[2]), 0.03 * np.sin(pose[2]), 0.02))
        t += 100_000
        r, th, _ = synth.make_scan(real_map, pose, seed=step, num_beams=290)      # the simulator's beam count
        tt = t - 100_000 + ((np.arange(290) + 1) * 100_000) // 290
        odom = synth.make_pose(*pose, utime=t)
        rdraw = (libc.rand() / 2147483647.0) * (1.0 / n)
        probe = port.ActionModel()
        probe.c = type(pf.action.c).from_buffer_copy(pf.action.c)
        assert probe.update(odom)[0]
        draws = probe.draws(mt, n)
        autime = t if not legacy else 1_000_000
        want_est, moved = pf.update(grid, odom, r, th, tt, rdraw, draws, action_utime=autime)
        assert moved and am.update(float(odom["x"]), float(odom["y"]), float(odom["theta"]), t)
        est = e.update(am, t, r, th, tt, rdraw, noise=draws)
        got = e.export_particles()
        for f in ("x", "y", "theta"):
            assert np.array_equal(got["parent_pose"][f], pf.particles["parent_pose"][f]), (step, f)
            worst_pose = max(worst_pose, np.abs(got["pose"][f].astype(np.float64) - pf.particles["pose"][f]).max())
        worst_w = max(worst_w, np.abs(got["weight"] - pf.particles["weight"]).max())
        assert abs(est.x - float(want_est["x"])) <= POSE_TOL and abs(est.y - float(want_est["y"])) <= POSE_TOL
        assert angle_err(est.theta, float(want_est["theta"])) <= POSE_TOL
        # the filter tracks the truth on this map
        assert abs(est.x - pose[0]) < 0.25 and abs(est.y - pose[1]) < 0.25
    assert worst_pose <= POSE_TOL and worst_w <= WEIGHT_TOL
    e.close()


# ------------------------------------------------------------------------ two-pass sensor path (certified float pass)
# The default sensor path scores every beam first with a float model whose result it can certify, and re-evaluates the
# rest with the literal restatement.  These tests pin: (a) identical scores to the exact-only
# path at scale and on hostile inputs, (b) the measured inputs of the certification's error budget.
def _scores_both_paths(grid, cloud, r, th, t, paths=(0, 1), **params):
    """Scores + stats per sensor_path: 0 = auto (the score-table pass where its window fits, else the two-pass path),
    2 = the two-pass path (certified float pass + exact re-evaluation), 1 = the literal restatement only."""
    out = []
    for path in paths:
        e = make_engine(len(cloud), grid, sensor_path=path, **params)
        e.import_particles(cloud)
        e.set_gather_counting(True)
        s = e.score(r, th, t)
        out.append((s, e.stats()))
        e.close()
    return out


@pytest.mark.parametrize("side,n,kind,variant", [
    (1000, 600_000, "tracking", "interp"), (1000, 300_000, "tracking", "degen"), (1000, 400_000, "uniform", "interp"),
    (200, 100_000, "tracking", "interp"), (200, 50_000, "uniform", "degen"), (3000, 300_000, "tracking", "interp"),
])
def test_two_pass_equals_exact_at_scale(side, n, kind, variant):
    rng = np.random.default_rng(side + n)
    grid = synth.make_map(side, seed=side)
    truth = synth.find_free_pose(grid, rng)
    r, th, t = synth.make_scan(grid, truth, seed=n)
    pu = int(t[0]) if variant == "interp" else int(t[-1])
    if kind == "tracking":
        cloud = synth.make_particles(n, truth, seed=7, parent_utime=pu, pose_utime=int(t[-1]))
    else:
        cloud = synth.make_uniform_particles(n, grid, seed=7, utime=int(t[-1]))
        if variant == "interp":
            cloud["parent_pose"]["utime"] = pu
            cloud["parent_pose"]["x"] += np.float32(0.013)
            cloud["parent_pose"]["theta"] = np.clip(cloud["parent_pose"]["theta"] - np.float32(0.02), -3.14, 3.14)
    (s0, st0), (s2, st2), (s1, st1) = _scores_both_paths(grid, cloud, r, th, t, paths=(0, 2, 1))
    assert st2["sensor_path"] == 2 and st1["sensor_path"] == 1
    # the score-table pass runs wherever the cloud's window (16-bit classes) fits shared memory: any cloud on the
    # 200 x 200 map (window clipped to the grid), tracking clouds with rays up to about 6 m; the rest falls back to the
    # two-pass path
    assert st0["sensor_path"] in (2, 3) and (side != 200 or st0["sensor_path"] == 3), st0
    assert np.array_equal(s2, s1) and np.array_equal(s0, s1)
    for st in (st0, st2):
        assert st["gathers"] == st1["gathers"] and st["evals"] == st1["evals"]
        assert 0 < st["deferred_evals"] < 0.5 * st["evals"]


def test_two_pass_hostile_particles(real_map):
    """NaN / infinite / huge / off-map positions, unwrapped headings, jumps between parent and pose: the float pass must
    hand every one of them to the exact pass."""
    r, th, t = synth.make_scan(real_map, (0.0, 0.0, 0.3), seed=4)
    cloud = synth.make_particles(4096, (0.0, 0.0, 0.3), seed=4, parent_utime=int(t[0]), pose_utime=int(t[-1]))
    x, px, h = cloud["pose"]["x"], cloud["parent_pose"]["x"], cloud["pose"]["theta"]
    x[0] = np.nan; x[1] = np.inf; x[2] = -np.inf; x[3] = 1e9; x[4] = -1e9; x[5] = 5.01; x[6] = -5.2; x[7] = 4.99
    px[8] = np.nan; px[9] = 40.0; px[10] = -4.999
    h[11] = 4.0; h[12] = -7.0; h[13] = np.nan; h[14] = 3.1415927; h[15] = -3.1415927
    cloud["pose"]["y"][16] = -4.9999; cloud["pose"]["y"][17] = 4.9999; cloud["parent_pose"]["y"][18] = 1e30
    (s0, st0), (s2, st2), (s1, st1) = _scores_both_paths(real_map, cloud, r, th, t, paths=(0, 2, 1))
    assert st2["sensor_path"] == 2
    assert np.array_equal(s2, s1) and np.array_equal(s0, s1)
    want, gathers, _ = port.likelihood(port_grid(real_map), cloud, r, th, t)
    assert np.array_equal(s2, want) and st2["gathers"] == gathers and st0["gathers"] == gathers
    # the same cloud without the non-finite positions (they blow up the window): the score-table pass runs, the
    # remaining hostile particles go to its exact drain
    finite = np.isfinite(cloud["pose"]["x"]) & np.isfinite(cloud["parent_pose"]["x"]) & (np.abs(cloud["pose"]["x"]) < 100) & \
        (np.abs(cloud["parent_pose"]["x"]) < 100) & (np.abs(cloud["parent_pose"]["y"]) < 100)
    sub = cloud[finite][:2048].copy()
    (t0, tst0), (t1, tst1) = _scores_both_paths(real_map, sub, r, th, t, paths=(0, 1))
    assert tst0["sensor_path"] == 3 and np.array_equal(t0, t1) and tst0["gathers"] == tst1["gathers"]


def test_two_pass_hostile_scans(real_map):
    """Scans outside the float pass's domain (beam angles beyond 2*pi, interpolation ratios far outside [0,1], infinite
    ranges) fall back to the exact path or defer; results stay identical to the oracle."""
    cloud = synth.make_particles(2000, (0.5, -0.25, 0.3), seed=5, parent_utime=1_000_000, pose_utime=1_100_000)
    r, th, t = synth.make_scan(real_map, (0.5, -0.25, 0.3), seed=5)
    cases = {
        "big_theta": (r, (th + np.float32(7.0)).astype(np.float32), t),
        "late_times": (r, th, t + 1_000_000),                 # ratios around 10: extrapolation
        "inf_range": (np.where(np.arange(len(r)) % 50 == 0, np.float32(np.inf), r).astype(np.float32), th, t),
        "short_ranges": (np.full_like(r, 0.16), th, t),
    }
    for name, (rr, tt, ti) in cases.items():
        for path in (0, 2):
            e = make_engine(len(cloud), real_map, sensor_path=path)
            e.import_particles(cloud)
            s = e.score(rr, tt, ti)
            st = e.stats()
            want, _, _ = port.likelihood(port_grid(real_map), cloud, rr, tt, ti)
            assert np.array_equal(s, want), (name, path)
            if name in ("big_theta", "late_times"):
                assert st["sensor_path"] == 1, name               # the float passes declared themselves not applicable
            e.close()


def test_scan_with_more_than_2048_beams(real_map):
    """Both certified passes index beams with 11 bits in their deferral queues; a longer scan must take the exact path
    (or a wider encoding), never a truncated beam index: scores equal the oracle's."""
    nb = 2500
    r0, th0, t0 = synth.make_scan(real_map, (0.5, -0.25, 0.3), num_beams=nb, seed=9)
    cloud = synth.make_particles(3000, (0.5, -0.25, 0.3), seed=9, parent_utime=int(t0[0]), pose_utime=int(t0[-1]))
    want, _, _ = port.likelihood(port_grid(real_map), cloud, r0, th0, t0)
    for path in (0, 2):
        e = make_engine(len(cloud), real_map, sensor_path=path)
        e.import_particles(cloud)
        assert np.array_equal(e.score(r0, th0, t0), want), path
        e.close()


def test_fast_trig_error_bound():
    """kFastTrigErr (mcl_device.cuh) must bound the SFU sine/cosine error over EVERY float the float pass can feed it."""
    e = engine.Engine(16)
    es, ec = e.fast_trig_error(-9.5, 9.5)
    e.close()
    assert 0 < max(es, ec) <= 2.0e-6


@pytest.mark.parametrize("side,kind", [(200, "tracking"), (2000, "tracking"), (4000, "uniform"), (1000, "uniform")])
def test_certification_margin(side, kind):
    """The float model's endpoint and extended point stay within the eps the certification assumes of the reference's
    exactly-rounded values, with margin to spare, on every evaluation of a 50 k-particle cloud."""
    rng = np.random.default_rng(side)
    grid = synth.make_map(side, seed=side + 1)
    truth = synth.find_free_pose(grid, rng)
    r, th, t = synth.make_scan(grid, truth, seed=side)
    if kind == "tracking":
        cloud = synth.make_particles(50_000, truth, seed=3, parent_utime=int(t[0]), pose_utime=int(t[-1]))
    else:
        cloud = synth.make_uniform_particles(50_000, grid, seed=3, utime=int(t[-1]))
        cloud["parent_pose"]["utime"] = int(t[0])
        cloud["parent_pose"]["y"] -= np.float32(0.021)
    e = make_engine(len(cloud), grid)
    e.import_particles(cloud)
    e.score(r, th, t)
    dev_e, dev_x2, eps = e.fast_margin()
    e.close()
    assert eps > 0
    assert dev_e <= 0.6 * eps, (dev_e, eps)
    assert dev_x2 <= 1.2 * eps, (dev_x2, eps)             # the extended point doubles the ray term; certified at 3 eps


# ------------------------------------------------------------------ per-batch map windows (global localisation clouds)
def test_init_uniform_is_stratified_and_ordered():
    """mcl_init_uniform: every particle inside the map, headings in [-pi, pi), equal counts per block, and consecutive
    particles close together (the order the per-batch windows rely on)."""
    grid = synth.make_map(2000, seed=3)
    n = 1 << 20
    e = make_engine(n, grid)
    e.init_uniform(utime=5, seed=9)
    c = e.export_particles()
    e.close()
    x, y, th = c["pose"]["x"], c["pose"]["y"], c["pose"]["theta"]
    ext = grid.width * grid.meters_per_cell
    assert x.min() >= grid.origin_x and x.max() <= grid.origin_x + ext
    assert y.min() >= grid.origin_y and y.max() <= grid.origin_y + ext
    assert th.min() >= -np.pi and th.max() <= np.pi
    assert np.array_equal(c["pose"]["x"], c["parent_pose"]["x"]) and np.all(c["weight"] == 1.0 / n)
    # uniform coverage: a 10 x 10 histogram within 3 % of flat
    hist, _, _ = np.histogram2d(x, y, bins=10)
    assert np.abs(hist / (n / 100) - 1).max() < 0.03
    # locality: any 1024 consecutive particles span a few metres, not the 100 m map
    span = np.maximum(x.reshape(-1, 1024).max(1) - x.reshape(-1, 1024).min(1),
                      y.reshape(-1, 1024).max(1) - y.reshape(-1, 1024).min(1))
    assert span.max() < 0.1 * ext


@pytest.mark.parametrize("side,n,dense,max_range,stride",
                         [(2000, 1_500_000, False, 8.0, 97), (1000, 700_001, False, 8.0, 97), (1000, 4_000_000, True, 5.0, 97),
                          (600, 2_000_003, True, 5.0, 97), (1000, 16_000_000, True, 8.0, 397)])
def test_batch_windows_equal_exact_and_oracle(side, n, dense, max_range, stride):
    """A uniformly initialised cloud over a map far larger than one shared-memory tile is scored behind per-batch
    windows (stats: map_tile_used == 5 for the score-table pass, whose CTAs rebuild their class tile and score table
    per batch, 3 for the older kernel families; the score-table pass needs a DENSE cloud -- config 5 has four particles
    per cell -- because its class tile and score table must hold the window of a whole batch of 1024 or 4096
    consecutive particles, and leaves sparse clouds to the older kernels).  The dense cases cover 16-bit class tiles
    (5 m rays) and 8-bit ones (8 m rays).  Scores must equal the exact-only path's, the L2-gather path's, and the
    oracle's on a sub-sample, before and after two full updates."""
    grid = synth.make_map(side, seed=side + 7)
    rng = np.random.default_rng(n)
    truth = synth.find_free_pose(grid, rng)
    r, th, t = synth.make_scan(grid, truth, seed=3, max_range=max_range)
    engines = {name: make_engine(n, grid, **kw) for name, kw in
               {"table": {}, "two_pass": {"sensor_path": 2}, "exact": {"sensor_path": 1}, "l2": {"map_tile": 1}}.items()}
    scores = {}
    for name, e in engines.items():
        e.init_uniform(utime=int(t[0]) - 100_000, seed=11)
        scores[name] = e.score(r, th, t)
    st = {k: e.stats() for k, e in engines.items()}
    assert (st["table"]["map_tile_used"], st["table"]["sensor_path"]) == ((5, 3) if dense else (3, 2)), st["table"]
    assert st["table"]["deferred_evals"] < 0.1 * st["table"]["evals"]
    assert np.array_equal(scores["table"], scores["two_pass"])
    assert st["two_pass"]["map_tile_used"] == 3 and st["two_pass"]["sensor_path"] == 2
    assert st["exact"]["map_tile_used"] == 3 and st["exact"]["sensor_path"] == 1
    assert st["l2"]["map_tile_used"] == 1
    assert np.array_equal(scores["two_pass"], scores["exact"]) and np.array_equal(scores["two_pass"], scores["l2"])
    sub = engines["two_pass"].export_particles(stride=stride)
    want, _, _ = port.likelihood(port_grid(grid), sub, r, th, t)
    assert np.array_equal(scores["two_pass"][::stride], want)
    if dense:
        assert st["table"]["table_variant"] == (3 if max_range > 6 else 2)
    # one full update (resample -> action -> score ...) keeps the order, hence the mode, and the paths agree
    am = engine.ActionModel()
    am.update(0.0, 0.0, 0.0, int(t[0]) - 100_000)
    assert am.update(0.02, 0.01, 0.01, int(t[-1]))
    clouds = {}
    for name in ("table", "two_pass", "exact"):
        e = engines[name]
        e.update(am, int(t[-1]), r, th, t, 0.37 / n)
        e.update(am, int(t[-1]) + 100_000, r, th, t + 100_000, 0.61 / n)
        clouds[name] = e.export_particles()
        assert e.stats()["map_tile_used"] in ((5,) if name == "table" and dense else (1, 3))
    for name in ("table", "two_pass"):
        for k in ("pose", "parent_pose"):
            for f in ("x", "y", "theta"):
                assert np.array_equal(clouds[name][k][f], clouds["exact"][k][f])
        assert np.array_equal(clouds[name]["weight"], clouds["exact"]["weight"])
    if dense:
        # the cloud collapses (here: a tight cloud is imported): the score-table pass goes back to ONE window, at the
        # latest on the pass after the one whose plan saw that the cloud fits one; scores stay the oracle's
        tight = synth.make_particles(n, truth, seed=5, parent_utime=int(t[0]) - 100_000, pose_utime=int(t[-1]))
        want, _, _ = port.likelihood(port_grid(grid), tight[::997], r, th, t)
        e = engines["table"]
        e.import_particles(tight)
        for _ in range(3):
            assert np.array_equal(e.score(r, th, t)[::997], want)
        assert e.stats()["map_tile_used"] == 4 and e.stats()["sensor_path"] == 3
        # ... and spreads out again
        e.init_uniform(utime=int(t[0]) - 100_000, seed=11)
        assert np.array_equal(e.score(r, th, t), scores["exact"]) and e.stats()["map_tile_used"] == 5
    for e in engines.values():
        e.close()


@pytest.mark.parametrize("n,paths,stride", [(8_000_000, (0, 2, 1), 10_007), (64_000_000, (0, 1), 100_003)])
def test_config5_shape_full_size(n, paths, stride):
    """BASELINE configs[4] on the 4000 x 4000 grid, uniformly initialised particles scored behind per-batch windows:
    8 M particles (one GPU's share of the count, spread over the whole map: half a particle per cell, too sparse for
    the score-table pass, so path 0 resolves to the certified float pass + exact pass) and all 64 M (four per cell: path 0
    is the score-table pass with one window per batch).  A sub-sample of the scores equals the ORACLE's, before and
    after a full update; the exact pass alone (path 1) agrees everywhere."""
    side = 4000
    grid = synth.make_map(side, seed=synth.MAP_SEED + 5)
    rng = np.random.default_rng(55)
    truth = synth.find_free_pose(grid, rng)
    r, th, t = synth.make_scan(grid, truth, seed=55)
    pg = port_grid(grid)
    am = engine.ActionModel()
    am.update(0.0, 0.0, 0.0, int(t[0]) - 100_000)
    assert am.update(0.02, 0.01, 0.01, int(t[-1]))
    dense = n >= 32_000_000
    out = {}
    for path in paths:
        e = make_engine(n, grid, sensor_path=path)
        e.init_uniform(utime=int(t[0]) - 100_000, seed=11)
        s0 = e.score(r, th, t)
        st = e.stats()
        want_tile = 5 if (path == 0 and dense) else 3
        assert st["map_tile_used"] == want_tile and st["sensor_path"] == {0: 3 if dense else 2, 2: 2, 1: 1}[path], st
        sub0 = e.export_particles(stride=stride)
        est = e.update(am, int(t[-1]), r, th, t, 0.37 / n)
        assert e.stats()["map_tile_used"] == want_tile
        s1 = e.score(r, th, t)
        sub1 = e.export_particles(stride=stride)
        out[path] = (s0, sub0, s1, sub1, (est.x, est.y, est.theta), e.stats()["weight_sum"])
        e.close()
    a = out[0]
    want0, _, _ = port.likelihood(pg, a[1], r, th, t)
    want1, _, _ = port.likelihood(pg, a[3], r, th, t)
    assert np.array_equal(a[0][::stride], want0) and np.array_equal(a[2][::stride], want1)
    for path in paths[1:]:
        b = out[path]
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and a[4] == b[4] and a[5] == b[5]
        for k in ("pose", "parent_pose"):
            for f in ("x", "y", "theta"):
                assert np.array_equal(a[3][k][f], b[3][k][f])
        assert np.array_equal(a[3]["weight"], b[3]["weight"])


# ------------------------------------------------------------------------------- map update on the device mirror
@pytest.mark.parametrize("hit,miss,max_laser", [(3, 1, 5.0), (4, 1, 5.0), (127, 100, 8.0), (0, 0, 5.0)])
def test_map_update_matches_oracle(hit, miss, max_laser):
    """Mapping::updateMap on the device mirror (mcl_map_update) leaves exactly the cells the reference's sequential
    loops leave: several scans in a row so cells saturate, moving poses so the rays are interpolated."""
    grid = synth.make_map(400, seed=9)
    rng = np.random.default_rng(hit + miss)
    pose = synth.find_free_pose(grid, rng)
    cells = grid.cells.copy()
    e = make_engine(16, grid)
    prev, t0 = pose, 1_000_000
    for k in range(6):
        r, th, t = synth.make_scan(grid, pose, seed=20 + k, t0=t0)
        cur = (pose[0], pose[1], pose[2], int(t[-1]))
        prv = (prev[0], prev[1], prev[2], int(t[0]) if k % 2 == 0 else int(t[-1]))      # interpolated / equal-utime rays
        init = k > 0                                          # the reference's first call changes no cell
        want = port.map_update(cells, grid.origin_x, grid.origin_y, grid.cells_per_meter,
                               synth.make_pose(*prv[:3], utime=prv[3]), synth.make_pose(*cur[:3], utime=cur[3]), init,
                               r, th, t, max_laser, hit, miss)
        x0, y0, w, h = e.map_update(prv, cur, init, r, th, t, max_laser, hit, miss)
        got = e.read_map_rect(0, 0, grid.width, grid.height)
        assert np.array_equal(got, want), k
        changed = np.nonzero(want != cells)
        if len(changed[0]):
            assert x0 <= changed[1].min() and changed[1].max() < x0 + w and y0 <= changed[0].min() and changed[0].max() < y0 + h
        else:
            assert not init or hit == miss == 0 or w >= 0
        cells = want
        prev, pose, t0 = pose, synth.odometry_step(rng, pose), t0 + 100_000
    e.close()


def test_map_update_edge_cases(real_map):
    """Robot next to the map border (rays leave the grid), empty and all-invalid scans, bad arguments."""
    e = make_engine(16, real_map)
    whole = lambda: e.read_map_rect(0, 0, real_map.width, real_map.height)
    base = whole()
    assert np.array_equal(base, real_map.cells)
    z = np.zeros(0, np.float32)
    assert e.map_update((0, 0, 0, 1), (0, 0, 0, 2), True, z, z, np.zeros(0, np.int64)) == (0, 0, 0, 0)
    e.map_update((0, 0, 0, 1), (0, 0, 0, 2), True, np.full(8, 0.1, np.float32), np.zeros(8, np.float32), np.zeros(8, np.int64))
    assert np.array_equal(whole(), base)
    pose = (4.9, -4.93, 0.4)                                   # 2 cells from two borders of the 10 m map
    r = np.full(360, 3.0, np.float32)
    th = (np.arange(360) * 2 * np.pi / 360).astype(np.float32)
    t = 1_000_000 + np.arange(360, dtype=np.int64) * 277
    want = port.map_update(base, real_map.origin_x, real_map.origin_y, real_map.cells_per_meter,
                           synth.make_pose(4.88, -4.92, 0.38, utime=int(t[0])), synth.make_pose(*pose, utime=int(t[-1])),
                           True, r, th, t)
    e.map_update((4.88, -4.92, 0.38, int(t[0])), (*pose, int(t[-1])), True, r, th, t)
    assert np.array_equal(whole(), want)
    with pytest.raises(engine.MclError):
        e.map_update((0, 0, 0, 1), (np.nan, 0, 0, 2), True, r, th, t)
    with pytest.raises(engine.MclError):
        e.map_update((0, 0, 0, 1), (0, 0, 0, 2), True, r, th, t, hit_odds=-1)
    e.close()


@pytest.mark.parametrize("placement", ["interior", "bench"])
def test_config4_full_size_two_pass_equals_exact(placement):
    """BASELINE configs[3] at full size (16 M particles x 360 beams, 2000 x 2000 grid): the default sensor path (the
    score-table pass: with 16-bit classes on bench.py's own workload, whose window the map's corner clips, with 8-bit
    classes for a robot in the map's interior, whose 8 m rays need a 355 x 355 window), the two-pass path and the
    literal restatement give the same 16 M scores, hence the same weight sum, estimate and resampled cloud -- and a
    sub-sample of those scores (every 10 007th particle) equals the ORACLE's, so the three cannot be wrong together."""
    n, side = synth.CONFIGS["config4"]
    if placement == "bench":
        import bench
        _, grid, truth, scans = bench.build_workload("config4")
        _, r, th, t = scans[1]
    else:
        rng = np.random.default_rng(4)
        grid = synth.make_map(side, seed=synth.MAP_SEED + 4)
        truth = synth.find_free_pose(grid, rng)
        r, th, t = synth.make_scan(grid, truth, seed=4)
    am = engine.ActionModel()
    am.update(*truth, int(t[0]) - 100_000)
    assert am.update(truth[0] + 0.02, truth[1] + 0.01, truth[2] + 0.01, int(t[-1]))
    out = {}
    for path in (0, 2, 1):
        e = make_engine(n, grid, sensor_path=path)
        e.init_at_pose(*truth, utime=int(t[0]) - 100_000, seed=21)
        est = e.update(am, int(t[-1]), r, th, t, 0.8401877171547095 / n)
        st = e.stats()
        assert st["sensor_path"] == {0: 3, 2: 2, 1: 1}[path]
        assert path != 0 or (st["map_tile_used"], st["table_variant"]) == (4, 0 if placement == "bench" else 1), st
        scores = e.score(r, th, t)                   # same cloud, same scan: the stage alone, all 16 M scores
        sub = e.export_particles(stride=10_007)
        am2 = engine.ActionModel()
        am2.c = type(am.c).from_buffer_copy(am.c)
        est2 = e.update(am2, int(t[-1]) + 1, r, th, t, 0.3 / n)      # resamples from the weights of the first update
        out[path] = (scores, st["weight_sum"], (est.x, est.y, est.theta), (est2.x, est2.y, est2.theta),
                     e.export_particles(stride=1009), st["deferred_evals"], st["evals"], sub)
        e.close()
    b = out[1]
    want, _, _ = port.likelihood(port_grid(grid), b[7], r, th, t)       # oracle on ~1600 of the 16 M particles
    for a in (out[0], out[2], out[1]):
        assert np.array_equal(a[0][::10_007], want)
        assert np.array_equal(a[0], b[0])
        assert a[1] == b[1] and a[2] == b[2] and a[3] == b[3]
        for k in ("pose", "parent_pose"):
            for f in ("x", "y", "theta"):
                assert np.array_equal(a[4][k][f], b[4][k][f])
        assert np.array_equal(a[4]["weight"], b[4]["weight"])
    for a in (out[0], out[2]):
        assert 0 < a[5] < 0.2 * a[6] and b[5] == 0


# ------------------------------------------------------------------ extension: log-sum-exp weights (weight_mode = 1)
@pytest.mark.parametrize("beta", [0.05, 0.5])
def test_lse_weight_mode(beta, real_map):
    """Non-default normalisation: w = exp(beta (s - max s)) / sum.  Scores are untouched (same sensor model), weights
    follow the formula to double rounding, the best particle's unnormalised weight is exactly 1, and the rest of the
    update (resampling on these weights, estimate) runs unchanged.  The default mode stays the reference's linear rule."""
    truth = (0.5, -0.25, 0.3)
    r, th, t = synth.make_scan(real_map, truth, seed=6)
    cloud = synth.make_particles(20_000, truth, seed=6, parent_utime=int(t[0]), pose_utime=int(t[-1]))
    e = make_engine(len(cloud), real_map, weight_mode=1, lse_beta=beta)
    e.import_particles(cloud)
    s = e.score(r, th, t)
    want_s, _, _ = port.likelihood(port_grid(real_map), cloud, r, th, t)
    assert np.array_equal(s, want_s)
    w = e.normalize()
    v = np.exp(beta * (s - s.max()))
    assert np.allclose(w, v / v.sum(), rtol=1e-12, atol=0) and abs(w.sum() - 1.0) < 1e-12
    st = e.stats()
    assert abs(st["weight_sum"] - v.sum()) <= 1e-9 * v.sum()
    assert abs(st["effective_sample_size"] - 1.0 / np.sum(w * w)) <= 1e-6 * st["effective_sample_size"]
    idx = e.resample(0.37 / len(cloud))
    want_idx, _ = port.resample(w, 0.37 / len(cloud))
    assert np.array_equal(idx, want_idx)                       # the exact sequential-sum resampling, on these weights
    e.close()
    lin = make_engine(len(cloud), real_map)
    lin.import_particles(cloud)
    lin.score(r, th, t)
    assert np.array_equal(lin.normalize(), np.maximum(s, 0.001) / lin.stats()["weight_sum"])
    lin.close()


@pytest.mark.parametrize("case", range(10))
def test_two_pass_randomized_geometry(case):
    """The certification's error budget depends on the map size, the resolution, the world coordinates and the ray
    length; sweep them: non-square maps, 2.5 / 5 / 10 cm cells (plus a stale cellsPerMeter, occupancy_grid.cpp:151-159),
    origins hundreds of metres from zero, ranges up to 12 m, 180...720 beams, beam angles in [-pi, pi) or [0, 2 pi),
    tracking and scattered clouds.  Two-pass scores == exact-only scores == oracle scores, every time."""
    rng = np.random.default_rng(1000 + case)
    mpc = [0.05, 0.025, 0.1, 0.05, 0.05][case % 5]
    w, h = int(rng.integers(150, 900)), int(rng.integers(150, 900))
    base = synth.make_map(max(w, h), seed=50 + case, meters_per_cell=mpc)
    cells = base.cells[:h, :w].copy()
    cells[-2:, :] = 100; cells[:, -2:] = 100
    ox, oy = [(-w * mpc / 2, -h * mpc / 2), (731.25, -412.5), (0.0, 0.0), (-2000.0, 1500.0), (5.5, 5.5)][case % 5]
    cpm = None if case != 7 else 1.0 / 0.05 * 1.01                # stale cells-per-metre
    grid = synth.GridSpec(cells, ox, oy, mpc, cpm)
    truth = synth.find_free_pose(grid, rng)
    nb = int(rng.choice([180, 290, 360, 500, 720]))
    r, th, t = synth.make_scan(grid, truth, num_beams=nb, seed=case, max_range=float(rng.choice([4.0, 8.0, 12.0])))
    if case % 2:
        th = np.where(th > np.pi, th - 2 * np.pi, th).astype(np.float32)
    n = 20_000
    if case % 3 == 2:
        cloud = synth.make_uniform_particles(n, grid, seed=case, utime=int(t[-1]))
        cloud["parent_pose"]["utime"] = int(t[0])
        cloud["parent_pose"]["x"] += np.float32(0.01)
    else:
        cloud = synth.make_particles(n, truth, seed=case, sigma_xy=float(rng.choice([0.05, 0.3, 1.5])),
                                     sigma_theta=float(rng.choice([0.02, 0.5])), parent_utime=int(t[0]), pose_utime=int(t[-1]))
    (s0, st0), (s2, st2), (s1, st1) = _scores_both_paths(grid, cloud, r, th, t, paths=(0, 2, 1))
    want, gathers, evals = port.likelihood(port_grid(grid), cloud, r, th, t)
    assert np.array_equal(s1, want) and np.array_equal(s2, want) and np.array_equal(s0, want)
    assert st1["gathers"] == gathers and st2["gathers"] == gathers and st2["evals"] == evals and st0["gathers"] == gathers
    assert st1["sensor_path"] == 1 and st2["sensor_path"] in (1, 2) and st0["sensor_path"] in (1, 2, 3)


def test_map_update_reference_golden(real_map):
    """mcl_map_update against the grids the EXECUTED reference left after each of five Mapping::updateMap calls
    (tests/golden/mapping.npz, tests/golden/make_golden.py: mapping_golden)."""
    g = load_golden("mapping")
    start = synth.GridSpec(g["start_cells"], real_map.origin_x, real_map.origin_y, real_map.meters_per_cell,
                           real_map.cells_per_meter)
    e = make_engine(16, start)
    for k in range(int(g["steps"])):
        prv, cur = g[f"{k}_previous"], g[f"{k}_pose"]
        e.map_update((float(prv["x"]), float(prv["y"]), float(prv["theta"]), int(prv["utime"])),
                     (float(cur["x"]), float(cur["y"]), float(cur["theta"]), int(cur["utime"])), k > 0,
                     g[f"{k}_ranges"], g[f"{k}_thetas"], g[f"{k}_times"], float(g["max_laser"]), int(g["hit"]),
                     int(g["miss"]))
        assert np.array_equal(e.read_map_rect(0, 0, start.width, start.height), g[f"{k}_cells"]), k
    e.close()


@pytest.mark.parametrize("theta,sigma_theta,expect_cull", [(0.4, 0.05, True), (3.13, 0.05, True), (-1.2, 0.9, False)])
def test_beam_culling_is_exact(theta, sigma_theta, expect_cull, monkeypatch):
    """The score-table pass leaves out the beams whose endpoints can only lie in class-0 cells (nothing positive within
    two cells) for every particle of the slice: they score 0 in the reference.  Scores with and without the culling
    equal the oracle's, map-read counts included; a robot in open space culls beams, a heading spread beyond 2.5 rad
    switches the culling off, headings around +-pi do not confuse the heading range."""
    grid = synth.make_map(700, seed=31)
    rng = np.random.default_rng(31)
    best, pose = -1, None
    for _ in range(40):                                      # the candidate with the most max-range (no-hit) rays
        x, y, _t = synth.find_free_pose(grid, rng, clearance=30)
        r, th, t = synth.make_scan(grid, (x, y, theta), seed=7)
        if (r > 7.9).sum() > best:
            best, pose = int((r > 7.9).sum()), (x, y, theta)
    r, th, t = synth.make_scan(grid, pose, seed=7)
    cloud = synth.make_particles(40_000, pose, seed=3, sigma_xy=0.10, sigma_theta=sigma_theta,
                                 parent_utime=int(t[0]) - 100_000, pose_utime=int(t[-1]))
    want, gathers, evals = port.likelihood(port_grid(grid), cloud, r, th, t)
    out = {}
    for cull in (True, False):
        if cull:
            monkeypatch.delenv("MCL_NO_CULL", raising=False)
        else:
            monkeypatch.setenv("MCL_NO_CULL", "1")
        e = make_engine(len(cloud), grid)
        e.import_particles(cloud)
        e.set_gather_counting(True)
        s = e.score(r, th, t)
        st = e.stats()
        assert st["sensor_path"] == 3 and np.array_equal(s, want)
        assert st["evals"] == evals and st["gathers"] == gathers
        out[cull] = st["culled_beams"]
        e.close()
    assert out[False] == 0
    assert (out[True] > 0) == expect_cull, (out, best)


@pytest.mark.parametrize("path", [0, 2])
def test_scoring_follows_device_map_updates(path):
    """The fast passes read a derived copy of the map (empty-neighbourhood look-ahead), the score-table pass builds its
    class tile and table from it and from the mirror; they must follow every mutation of the mirror: mcl_map_update,
    mcl_update_map_rect, mcl_set_map.  Scores after each equal the oracle's on the same map."""
    grid = synth.make_map(300, seed=12)
    rng = np.random.default_rng(12)
    pose = synth.find_free_pose(grid, rng)
    r, th, t = synth.make_scan(grid, pose, seed=12)
    cloud = synth.make_particles(20_000, pose, seed=12, sigma_xy=0.4, sigma_theta=0.3, parent_utime=int(t[0]),
                                 pose_utime=int(t[-1]))
    e = make_engine(len(cloud), grid, sensor_path=path)
    e.import_particles(cloud)
    cells = grid.cells.copy()
    seen = set()

    def check(tag):
        g = synth.GridSpec(cells, grid.origin_x, grid.origin_y, grid.meters_per_cell, grid.cells_per_meter)
        want, _, _ = port.likelihood(port_grid(g), cloud, r, th, t)
        assert np.array_equal(e.score(r, th, t), want), tag
        # (path 0: the score-table pass, until a map whose table overflows shared memory -- the random map of the last
        # step -- sends it back to the two-pass path; that launch itself scores everything exactly)
        seen.add(e.stats()["sensor_path"])
        assert e.stats()["sensor_path"] in ((2, 3) if path == 0 else (2,))

    check("initial")
    prv = (pose[0] - 0.03, pose[1], pose[2], int(t[0]))
    cur = (pose[0], pose[1], pose[2], int(t[-1]))
    for k in range(3):                                   # device-side map updates with strong odds: walls appear / vanish
        cells = port.map_update(cells, grid.origin_x, grid.origin_y, grid.cells_per_meter,
                                synth.make_pose(*prv[:3], utime=prv[3]), synth.make_pose(*cur[:3], utime=cur[3]), True,
                                r, th, t, 8.0, 90, 70)
        e.map_update(prv, cur, True, r, th, t, 8.0, 90, 70)
        check(f"map_update {k}")
    patch = rng.integers(-128, 128, (41, 57)).astype(np.int8)
    cells[100:141, 120:177] = patch
    e.update_map_rect(120, 100, patch)
    check("rect")
    cells = np.where(rng.random(cells.shape) < 0.01, 50, -5).astype(np.int8)
    e.set_map(cells, grid.origin_x, grid.origin_y, grid.meters_per_cell, grid.cells_per_meter)
    check("set_map")
    assert (3 in seen) == (path == 0)
    e.close()


# ------------------------------------------------- next row: distance grid + likelihood-field sensor mode
@pytest.mark.parametrize("name", ["real", "synth700", "all_free"])
def test_distance_grid_equals_the_oracle(name, real_map):
    """mcl_distance_grid (two separable sweeps on the device mirror) == ObstacleDistanceGrid::setDistances as the oracle
    restates it (itself bit-equal to the executed reference: tests/test_oracle.py), bit for bit."""
    grid = {"real": real_map, "synth700": synth.make_map(700, seed=31),
            "all_free": synth.GridSpec(np.full((50, 80), -3, np.int8), 0.0, 0.0, 0.05)}[name]
    e = make_engine(16, grid)
    got = e.distance_grid(grid.width, grid.height)
    want = port.distance_grid(grid.cells)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    e.close()


@pytest.mark.parametrize("side,n,kind", [(200, 30_000, "tracking"), (600, 50_000, "tracking"), (300, 20_000, "uniform")])
def test_likelihood_field_mode_equals_the_oracle(side, n, kind, real_map):
    """sensor_mode = 1 (an extension, never the default): a ray scores the field value u = max(0, 127 - 8 d^2) of its
    endpoint cell, d = steps to the nearest occupied cell.  Scores equal the oracle's restatement exactly (the float
    pass only takes cells it can certify; the rest go through the exact endpoint), they follow map changes, and the
    log-sum-exp normaliser turns them into weights."""
    grid = real_map if side == 200 else synth.make_map(side, seed=side)
    rng = np.random.default_rng(side)
    truth = synth.find_free_pose(grid, rng)
    r, th, t = synth.make_scan(grid, truth, seed=side, max_range=5.0)
    if kind == "tracking":
        cloud = synth.make_particles(n, truth, seed=3, sigma_xy=0.15, sigma_theta=0.08, parent_utime=int(t[0]), pose_utime=int(t[-1]))
    else:
        cloud = synth.make_uniform_particles(n, grid, seed=3, utime=int(t[-1]))
        cloud["parent_pose"]["utime"] = int(t[0])
        cloud["parent_pose"]["x"] += np.float32(0.02)
    e = make_engine(n, grid, sensor_mode=1, weight_mode=1, lse_beta=0.02)
    e.import_particles(cloud)
    s = e.score(r, th, t)
    want = port.likelihood_field(port_grid(grid), cloud, r, th, t)
    assert np.array_equal(s, want)
    st = e.stats()
    assert st["sensor_path"] == 3 and st["deferred_evals"] < 0.2 * st["evals"]
    assert want.max() > (300 if kind == "tracking" else 0)                  # the tracking cloud really sits on walls
    w = e.normalize()
    v = np.exp(0.02 * (s - s.max()))
    assert np.abs(w - v / v.sum()).max() <= 1e-12
    # the field follows the mirror: a wall across the map changes the scores the way the oracle says
    cells = grid.cells.copy()
    cells[grid.height // 2, :] = 90
    e.update_map_rect(0, grid.height // 2, cells[grid.height // 2:grid.height // 2 + 1, :])
    g2 = synth.GridSpec(cells, grid.origin_x, grid.origin_y, grid.meters_per_cell, grid.cells_per_meter)
    want2 = port.likelihood_field(port_grid(g2), cloud, r, th, t)
    assert np.array_equal(e.score(r, th, t), want2) and (side != 200 or not np.array_equal(want2, want))
    e.close()
    # the default mode of the same engine build is untouched
    e0 = make_engine(n, grid)
    e0.import_particles(cloud)
    ref_scores, _, _ = port.likelihood(port_grid(grid), cloud, r, th, t)
    assert np.array_equal(e0.score(r, th, t), ref_scores)
    e0.close()
