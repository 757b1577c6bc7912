"""CPU check of the certified float pass's DESIGN, independent of the CUDA code: a numpy model of
the pass (tests/certification_model.py: same arithmetic, same tests, same error budget) against the oracle's per-ray
scores.  Whatever the model calls certain must be the oracle's score -- with the sine/cosine perturbed up to the SFU
error the budget assumes -- and it must call most evaluations certain, or the two-pass path would be pointless."""
import numpy as np
import pytest

from botlab_b200 import synth
from oracle import port
import certification_model as cm

MIN_RANGE = 0.15


def make_case(case):
    rng = np.random.default_rng(7000 + case)
    mpc = [0.05, 0.025, 0.05, 0.05][case % 4]
    w, h = int(rng.integers(160, 700)), int(rng.integers(160, 700))
    base = synth.make_map(max(w, h), seed=90 + case, meters_per_cell=mpc)
    cells = base.cells[:h, :w].copy()
    cells[-2:, :] = 100; cells[:, -2:] = 100
    ox, oy = [(-w * mpc / 2, -h * mpc / 2), (731.25, -412.5), (3.0, 3.0), (-2000.0, 1500.0)][case % 4]
    grid = synth.GridSpec(cells, ox, oy, mpc)
    truth = synth.find_free_pose(grid, rng)
    r, th, t = synth.make_scan(grid, truth, num_beams=int(rng.choice([180, 360, 500])), seed=case,
                               max_range=float(rng.choice([4.0, 8.0, 12.0])))
    if case % 2:
        th = np.where(th > np.pi, th - 2 * np.pi, th).astype(np.float32)
    n = 150
    if case % 3 == 2:
        cloud = synth.make_uniform_particles(n, grid, seed=case, utime=int(t[-1]))
        cloud["parent_pose"]["utime"] = int(t[0])
        cloud["parent_pose"]["x"] += np.float32(0.01)
    else:
        cloud = synth.make_particles(n, truth, seed=case, sigma_xy=float(rng.choice([0.05, 0.5])),
                                     sigma_theta=float(rng.choice([0.02, 0.5])), parent_utime=int(t[0]), pose_utime=int(t[-1]))
    return grid, cloud, r, th, t


@pytest.mark.parametrize("case", range(8))
def test_certain_evaluations_equal_the_oracle(case):
    grid, cloud, r, th, t = make_case(case)
    ratios = (t - int(cloud["parent_pose"]["utime"][0])).astype(np.float64) / float(
        int(cloud["pose"]["utime"][0]) - int(cloud["parent_pose"]["utime"][0]))
    window = (0, 0, grid.width, grid.height) if case % 2 else cm.cloud_window(grid, cloud, r, MIN_RANGE)
    plan = cm.Plan(grid, r, th, ratios, MIN_RANGE, *window)
    assert plan.enabled and plan.eps < 0.02
    fast_cells = cm.derive_fast_map(grid.cells)
    pg = port.Grid(grid.cells, grid.origin_x, grid.origin_y, grid.cells_per_meter)
    rng = np.random.default_rng(case)
    evals = certain_total = wrong = 0
    for i in range(len(cloud)):
        v2, certain = cm.fast_pass(grid, plan, cloud[i], r, th, ratios, MIN_RANGE, fast_cells, rng)
        want2 = np.rint(2.0 * port.ray_scores(pg, cloud[i], r, th, t)).astype(np.int64)     # half units, exact
        assert len(want2) == len(v2)
        wrong += int((certain & (v2 != want2)).sum())
        evals += len(v2)
        certain_total += int(certain.sum())
    assert wrong == 0
    assert certain_total >= 0.5 * evals, (certain_total, evals)


def test_derived_map_flags():
    cells = np.zeros((12, 12), np.int8)
    cells[5, 6] = 40
    cells[0, 0] = -7
    d = cm.derive_fast_map(cells)
    assert d[5, 6] == 40
    assert (d[3:8, 4:9] >= 0).all() and d[3, 4] == 0 and d[7, 8] == 0        # within two cells of the wall: no flag
    assert d[2, 6] == -1 and d[5, 3] == -1 and d[0, 0] == -1 and d[11, 11] == -1


def test_budget_grows_with_the_geometry():
    """eps must grow with the ray length, the map size and the distance of the origin from zero (the terms of the budget)."""
    def eps(side, origin, max_range):
        g = synth.GridSpec(np.zeros((side, side), np.int8), origin, origin, 0.05)
        r = np.full(360, max_range, np.float32)
        th = (np.arange(360) * 2 * np.pi / 360).astype(np.float32)
        return cm.Plan(g, r, th, np.linspace(0, 1, 360), MIN_RANGE, 0, 0, side, side).eps
    base = eps(400, -10.0, 4.0)
    assert eps(400, -10.0, 8.0) > base and eps(3000, -75.0, 4.0) > base and eps(400, 3000.0, 4.0) > base
    assert 2e-4 < base < 2e-3


def test_the_check_is_sensitive():
    """Negative control: with the uncertainty bands removed (every evaluation called certain) the same comparison does
    find wrong scores, so a passing run above means the bands do their job rather than that the check is blind."""
    grid, cloud, r, th, t = make_case(3)
    ratios = (t - int(cloud["parent_pose"]["utime"][0])).astype(np.float64) / float(
        int(cloud["pose"]["utime"][0]) - int(cloud["parent_pose"]["utime"][0]))
    plan = cm.Plan(grid, r, th, ratios, MIN_RANGE, 0, 0, grid.width, grid.height)
    plan.fmask, plan.magic = (1 << plan.fb) - 1, plan.magic_base
    plan.t_dir = plan.t_dir_neg = np.float32(0.0)
    fast_cells = cm.derive_fast_map(grid.cells)
    pg = port.Grid(grid.cells, grid.origin_x, grid.origin_y, grid.cells_per_meter)
    rng = np.random.default_rng(3)
    wrong = 0
    for i in range(len(cloud)):
        v2, certain = cm.fast_pass(grid, plan, cloud[i], r, th, ratios, MIN_RANGE, fast_cells, rng)
        want2 = np.rint(2.0 * port.ray_scores(pg, cloud[i], r, th, t)).astype(np.int64)
        wrong += int((certain & (v2 != want2)).sum())
    assert wrong > 0


# ---- the table pass (mcl_table.cuh) ---------------------------------------------------------------------------------
def make_interior_case(case):
    """A tracking cloud well inside a large map: the table pass's EDGE = 0 variant (no window / grid tests per ray)."""
    rng = np.random.default_rng(8000 + case)
    grid = synth.make_map(900, seed=190 + case)
    while True:
        truth = synth.find_free_pose(grid, rng)
        cx, cy = (truth[0] - grid.origin_x) / 0.05, (truth[1] - grid.origin_y) / 0.05
        if 260 < cx < 640 and 260 < cy < 640:
            break
    r, th, t = synth.make_scan(grid, truth, num_beams=360, seed=case, max_range=4.0)
    cloud = synth.make_particles(150, truth, seed=case, sigma_xy=0.10, sigma_theta=0.05, parent_utime=int(t[0]),
                                 pose_utime=int(t[-1]))
    return grid, cloud, r, th, t


def make_corner_case(case):
    """A tracking cloud next to the map's low edges: rays end at negative global coordinates, where the reference's
    truncation toward zero is not a floor (the table pass's EDGE 1 / 2 variants and its never-certified cells)."""
    rng = np.random.default_rng(9000 + case)
    grid = synth.make_map(420, seed=290 + case)
    lo, hi = [(20, 70), (90, 150), (20, 150)][case % 3]
    while True:
        truth = synth.find_free_pose(grid, rng)
        cx, cy = (truth[0] - grid.origin_x) / 0.05, (truth[1] - grid.origin_y) / 0.05
        if lo < min(cx, cy) < hi:
            break
    r, th, t = synth.make_scan(grid, truth, num_beams=360, seed=case, max_range=4.0)
    r = np.where(np.arange(len(r)) % 7 == 0, np.float32(3.9), r).astype(np.float32)       # some rays leave the grid
    cloud = synth.make_particles(150, truth, seed=case, sigma_xy=0.15, sigma_theta=0.1, parent_utime=int(t[0]),
                                 pose_utime=int(t[-1]))
    return grid, cloud, r, th, t


def _table_case(case, force_edge=None, interp=True, perturb=None):
    grid, cloud, r, th, t = (make_case(case) if case < 100 else make_interior_case(case) if case < 200
                             else make_corner_case(case))
    t_b, t_a = int(cloud["parent_pose"]["utime"][0]), int(cloud["pose"]["utime"][0])
    if interp:
        ratios = (t - t_b).astype(np.float64) / float(t_a - t_b)
    else:
        ratios = np.ones(len(t))
        cloud["parent_pose"]["utime"] = t_a
    plan = cm.TabPlan(grid, cloud, r, th, ratios, MIN_RANGE)
    if not plan.ok:
        return None
    if perturb:
        perturb(plan)
    K, T, entries = cm.build_score_table(grid, plan)
    pg = port.Grid(grid.cells, grid.origin_x, grid.origin_y, grid.cells_per_meter)
    rng = np.random.default_rng(case)
    evals = certain_total = wrong = 0
    edges = set()
    for i in range(len(cloud)):
        v2, certain, edge = cm.table_pass(grid, plan, K, T, cloud[i], r, th, ratios, MIN_RANGE, rng, interp=interp,
                                          force_edge=force_edge)
        want2 = np.rint(2.0 * port.ray_scores(pg, cloud[i], r, th, t)).astype(np.int64)
        assert len(want2) == len(v2)
        wrong += int((certain & (v2 != want2)).sum())
        evals += len(v2)
        certain_total += int(certain.sum())
        edges.add(edge)
    return wrong, certain_total, evals, edges, plan


@pytest.mark.parametrize("case", range(8))
@pytest.mark.parametrize("force_edge", [None, 2])
def test_table_pass_certain_evaluations_equal_the_oracle(case, force_edge):
    res = _table_case(case, force_edge)
    if res is None:
        pytest.skip("window does not fit the table pass's budget")
    wrong, certain_total, evals, edges, plan = res
    assert wrong == 0
    assert plan.eps < 0.02
    if force_edge is None and case % 3 != 2:
        assert certain_total >= 0.5 * evals, (certain_total, evals, edges)


@pytest.mark.parametrize("case", [100, 101, 102])
@pytest.mark.parametrize("interp", [True, False])
def test_table_pass_interior_cloud(case, interp):
    res = _table_case(case, interp=interp)
    assert res is not None
    wrong, certain_total, evals, edges, plan = res
    assert wrong == 0 and edges == {0}
    assert certain_total >= 0.9 * evals, (certain_total, evals)


@pytest.mark.parametrize("case", [200, 201, 202, 203, 204, 205])
def test_table_pass_cloud_at_the_low_edges(case):
    res = _table_case(case)
    assert res is not None
    wrong, certain_total, evals, edges, plan = res
    assert wrong == 0
    assert edges & {1, 2}, edges
    assert certain_total >= 0.5 * evals, (certain_total, evals)


def test_table_pass_equal_utime_variant():
    res = _table_case(0, interp=False)
    assert res is not None and res[0] == 0 and res[1] >= 0.5 * res[2]


def test_table_check_is_sensitive():
    """Negative control: without the cell band and the direction band the same comparison finds wrong scores."""
    def blind(plan):
        plan.frac_thr = 0
        plan.t3 = np.float32(0.0)
    res = _table_case(3, perturb=blind)
    assert res is not None and res[0] > 0


# ------------------------------------------------------------------------------------------------ beam culling (CPU)
def _cull_case(case):
    """A tracking cloud in the open part of a map (so that some rays end in open space), several geometries."""
    rng = np.random.default_rng(8100 + case)
    mpc = [0.05, 0.05, 0.025, 0.05][case % 4]
    side = [500, 700, 600, 400][case % 4]
    grid = synth.make_map(side, seed=40 + case, meters_per_cell=mpc)
    theta = [0.3, 3.12, -1.6, -3.1][case % 4]
    best, pose = -1, None
    for _ in range(25):
        x, y, _t = synth.find_free_pose(grid, rng, clearance=24)
        r, th, t = synth.make_scan(grid, (x, y, theta), seed=case, max_range=[8.0, 8.0, 5.0, 12.0][case % 4])
        far = int((r > 0.95 * r.max()).sum())
        if far > best:
            best, pose = far, (x, y, theta)
    r, th, t = synth.make_scan(grid, pose, seed=case, max_range=[8.0, 8.0, 5.0, 12.0][case % 4])
    cloud = synth.make_particles(120, pose, seed=case, sigma_xy=0.08, sigma_theta=0.04, parent_utime=int(t[0]) - 100_000,
                                 pose_utime=int(t[-1]))
    return grid, cloud, r, th, t


@pytest.mark.parametrize("case", range(4))
def test_culled_beams_score_zero_for_every_particle(case):
    """The culling rule of the score-table pass, restated in numpy (certification_model.cull_flags): every beam it culls
    scores exactly 0 in the ORACLE for every particle of the cloud, and in open space it does cull beams."""
    grid, cloud, r, th, t = _cull_case(case)
    t0, t1 = int(cloud["parent_pose"]["utime"][0]), int(cloud["pose"]["utime"][0])
    ratios = (t - t0).astype(np.float64) / float(t1 - t0)
    flags = cm.cull_flags(grid, cloud, r, th, ratios, MIN_RANGE)
    pg = port.Grid(grid.cells, grid.origin_x, grid.origin_y, grid.cells_per_meter)
    per_ray = np.stack([port.ray_scores(pg, cloud[i], r, th, t) for i in range(len(cloud))])       # [particles, beams]
    assert per_ray.shape[1] == len(flags)
    assert (per_ray[:, flags] == 0).all()
    assert flags.sum() > 0, "no beam culled: the case does not exercise the rule"


def test_the_culling_check_is_sensitive():
    """Negative control: a rule that ignores the heading spread and the margins culls beams that do score somewhere."""
    wrong = 0
    for case in range(4):
        grid, cloud, r, th, t = _cull_case(case)
        t0, t1 = int(cloud["parent_pose"]["utime"][0]), int(cloud["pose"]["utime"][0])
        ratios = (t - t0).astype(np.float64) / float(t1 - t0)
        flags = cm.cull_flags(grid, cloud, r, th, ratios, MIN_RANGE, margin=-6.0, hw_scale=0.02)
        pg = port.Grid(grid.cells, grid.origin_x, grid.origin_y, grid.cells_per_meter)
        per_ray = np.stack([port.ray_scores(pg, cloud[i], r, th, t) for i in range(len(cloud))])
        wrong += int((per_ray[:, flags] != 0).any(axis=0).sum())
    assert wrong > 0


# ------------------------------------------------------------------------------------------------ 8-bit class tile (CPU)
@pytest.mark.parametrize("case", [0, 1, 2, 5])
def test_wide_tile_addresses_the_same_entries(case):
    """The 8-bit class tile + segment bases (windows too large for 16-bit classes) resolve every window cell to the same
    score-table entry as the 16-bit tile does."""
    grid, cloud, r, th, t = make_case(case)
    ratios = (t - int(cloud["parent_pose"]["utime"][0])).astype(np.float64) / float(
        int(cloud["pose"]["utime"][0]) - int(cloud["parent_pose"]["utime"][0]))
    plan = cm.TabPlan(grid, cloud, r, th, ratios, MIN_RANGE, smem_total=10 ** 7)
    if not plan.ok:
        pytest.skip("no table plan for this geometry")
    K, T, n_entries = cm.build_score_table(grid, plan)
    K8, base, entry_of, bias_x = cm.build_score_table_wide(K)
    assert len(entry_of) == n_entries and cm.tab_nseg(plan.w) == base.shape[1]
    cy, cx = np.meshgrid(np.arange(plan.h), np.arange(plan.w), indexing="ij")
    e = cm.wide_entry(K8, base, bias_x, cy, cx)
    k16 = np.where(e < cm.TAB_FIXED, e, entry_of[np.maximum(e - cm.TAB_FIXED, 0)])
    assert np.array_equal(k16, K)
    assert (T[k16] == T[K]).all()
