"""Seeded synthetic inputs for the MCL path: occupancy grids, lidar scans, odometry steps and
particle clouds.  numpy only; shared by the parity tests (fed to oracle and engine alike) and by bench.py.

Shapes follow the reference's producers: the grid geometry of OccupancyGrid's ctor (src/slam/occupancy_grid.cpp:19-36),
scans like src/sim/lidar.py:74-138 (ray march to the first occupied cell, clockwise beam angles as
src/mbot/rplidar_driver.cpp:225 reports them), odometry noise like src/sim/sim.py:178-193.
"""
import numpy as np

from .engine import PARTICLE_DTYPE, POSE_DTYPE

MAP_SEED = 0xB07A8

# BASELINE.json configs: name -> (particles, map cells per side)
CONFIGS = {
    "config1": (200, 200),
    "config2": (100_000, 200),
    "config3": (1_000_000, 1000),
    "config4": (16_000_000, 2000),
    "config5": (64_000_000, 4000),
}


class GridSpec:
    """cells (H, W) int8 + the reference's geometry fields (occupancy_grid.hpp:84-93)."""

    def __init__(self, cells, origin_x, origin_y, meters_per_cell, cells_per_meter=None):
        self.cells = np.ascontiguousarray(cells, np.int8)
        self.height, self.width = self.cells.shape
        self.origin_x = float(np.float32(origin_x))
        self.origin_y = float(np.float32(origin_y))
        self.meters_per_cell = float(np.float32(meters_per_cell))
        if cells_per_meter is None:
            cells_per_meter = np.float32(1.0) / np.float32(meters_per_cell)     # occupancy_grid.cpp:30
        self.cells_per_meter = float(np.float32(cells_per_meter))


def make_map(side_cells, seed=MAP_SEED, occupied_frac=0.015, meters_per_cell=0.05):
    """Square grid centred on the origin: outer wall ring + random wall segments/boxes until about occupied_frac of the
    cells are > 0 (reference maps: 0.7-1.8 %).  Occupied cells are uniform in [1,127] so value-dependent scoring is
    exercised; a band of free cells next to walls is negative; the rest stays 0."""
    rng = np.random.default_rng(seed)
    n = int(side_cells)
    occ = np.zeros((n, n), bool)
    occ[0:2, :] = occ[-2:, :] = True
    occ[:, 0:2] = occ[:, -2:] = True
    target = occupied_frac * n * n
    while occ.sum() < target:
        if rng.random() < 0.7:      # wall segment, 1-2 cells thick
            length = int(rng.integers(max(4, n // 40), max(8, n // 6)))
            thick = int(rng.integers(1, 3))
            x, y = int(rng.integers(2, n - 2)), int(rng.integers(2, n - 2))
            if rng.random() < 0.5:
                occ[y:y + thick, x:x + length] = True
            else:
                occ[y:y + length, x:x + thick] = True
        else:                        # hollow box
            w, h = int(rng.integers(4, max(6, n // 20))), int(rng.integers(4, max(6, n // 20)))
            x, y = int(rng.integers(2, n - w - 2)), int(rng.integers(2, n - h - 2))
            occ[y, x:x + w] = occ[y + h - 1, x:x + w] = True
            occ[y:y + h, x] = occ[y:y + h, x + w - 1] = True
    cells = np.zeros((n, n), np.int8)
    # free band: cells within 3 of an occupied cell (cheap dilation) get negative log-odds
    near = occ.copy()
    for _ in range(3):
        g = near.copy()
        g[1:, :] |= near[:-1, :]; g[:-1, :] |= near[1:, :]; g[:, 1:] |= near[:, :-1]; g[:, :-1] |= near[:, 1:]
        near = g
    free = near & ~occ
    cells[free] = rng.integers(-128, 0, size=int(free.sum()), dtype=np.int64).astype(np.int8)
    cells[occ] = rng.integers(1, 128, size=int(occ.sum()), dtype=np.int64).astype(np.int8)
    half = np.float32(n * meters_per_cell) / np.float32(2.0)                  # occupancy_grid.cpp:23
    return GridSpec(cells, -half, -half, meters_per_cell)


def load_map_file(path):
    """Parser of the reference's ASCII .map format (OccupancyGrid::loadFromFile, occupancy_grid.cpp:138-175).
    cells_per_meter is recomputed from the file's resolution (the reference keeps its ctor value)."""
    with open(path) as f:
        head = f.readline().split()
        ox, oy, w, h, mpc = float(head[0]), float(head[1]), int(head[2]), int(head[3]), float(head[4])
        vals = np.array(f.read().split(), dtype=np.int64)
    cells = vals[: w * h].astype(np.int8).reshape(h, w)
    return GridSpec(cells, ox, oy, mpc)


def find_free_pose(grid, rng, clearance=6):
    """A pose whose neighbourhood (clearance cells) holds no occupied cell."""
    occ = grid.cells > 0
    for _ in range(10000):
        cx = int(rng.integers(clearance + 2, grid.width - clearance - 2))
        cy = int(rng.integers(clearance + 2, grid.height - clearance - 2))
        if not occ[cy - clearance:cy + clearance + 1, cx - clearance:cx + clearance + 1].any():
            x = grid.origin_x + (cx + 0.5) * grid.meters_per_cell
            y = grid.origin_y + (cy + 0.5) * grid.meters_per_cell
            return float(x), float(y), float(rng.uniform(-np.pi, np.pi))
    raise RuntimeError("no free pose found")


def make_scan(grid, pose, num_beams=360, seed=0, t0=1_000_000, sweep_us=100_000, max_range=8.0, range_sigma=0.005,
              invalid_frac=0.02):
    """One lidar sweep from pose=(x, y, theta): thetas ascending (clockwise sensor: beam global angle = theta - thetas[i],
    moving_laser_scan.cpp:33), ranges by marching at half-cell steps to the first cell > 0 or max_range, plus noise;
    invalid_frac of the beams report 0 (exercises the 0.15 m gate, moving_laser_scan.cpp:24)."""
    rng = np.random.default_rng(seed)
    x, y, th = pose
    i = np.arange(num_beams)
    thetas = (2.0 * np.pi * i / num_beams).astype(np.float32)
    ang = th - thetas.astype(np.float64)
    step = grid.meters_per_cell * 0.5
    d = np.arange(1, int(max_range / step) + 1) * step                          # (S,)
    px = x + np.outer(np.cos(ang), d)
    py = y + np.outer(np.sin(ang), d)
    cx = np.floor((px - grid.origin_x) * grid.cells_per_meter).astype(np.int64)
    cy = np.floor((py - grid.origin_y) * grid.cells_per_meter).astype(np.int64)
    inside = (cx >= 0) & (cx < grid.width) & (cy >= 0) & (cy < grid.height)
    hit = np.zeros_like(inside)
    hit[inside] = grid.cells[cy[inside], cx[inside]] > 0
    first = np.where(hit.any(axis=1), hit.argmax(axis=1), len(d) - 1)
    ranges = d[first] + rng.normal(0.0, range_sigma, num_beams)
    ranges = np.clip(ranges, 0.0, None)
    ranges[rng.random(num_beams) < invalid_frac] = 0.0
    times = t0 + (i * sweep_us) // num_beams
    return ranges.astype(np.float32), thetas, times.astype(np.int64)


def odometry_step(rng, pose, step=(0.02, 0.01, 0.01), noise=(1e-3, 1e-3, 3e-3)):
    x, y, th = pose
    return (x + step[0] + rng.normal(0, noise[0]), y + step[1] + rng.normal(0, noise[1]),
            float(np.arctan2(np.sin(th + step[2] + rng.normal(0, noise[2])), np.cos(th + step[2]))))


def wrap_to_pi_f32(a):
    a = np.asarray(a, np.float64)
    return ((a + np.pi) % (2 * np.pi) - np.pi).astype(np.float32)


def make_particles(n, truth, seed=0, sigma_xy=0.10, sigma_theta=0.05, parent_utime=1_000_000, pose_utime=1_099_722,
                   motion=(0.02, 0.01, 0.01)):
    """Tracking cloud: parent = truth + N(0, sigma), pose = parent + one noisy motion step, so parent_pose != pose.
    Equal utimes give the reference's de-facto degenerate path; different utimes the per-ray interpolation."""
    rng = np.random.default_rng(seed)
    p = np.zeros(n, PARTICLE_DTYPE)
    px = truth[0] + rng.normal(0, sigma_xy, n)
    py = truth[1] + rng.normal(0, sigma_xy, n)
    pth = truth[2] + rng.normal(0, sigma_theta, n)
    p["parent_pose"]["x"] = px.astype(np.float32)
    p["parent_pose"]["y"] = py.astype(np.float32)
    p["parent_pose"]["theta"] = wrap_to_pi_f32(pth)
    p["parent_pose"]["utime"] = parent_utime
    p["pose"]["x"] = (px + motion[0] + rng.normal(0, 0.005, n)).astype(np.float32)
    p["pose"]["y"] = (py + motion[1] + rng.normal(0, 0.005, n)).astype(np.float32)
    p["pose"]["theta"] = wrap_to_pi_f32(pth + motion[2] + rng.normal(0, 0.02, n))
    p["pose"]["utime"] = pose_utime
    p["weight"] = 1.0 / n
    return p


def make_uniform_particles(n, grid, seed=0, utime=1_000_000):
    """Global-localisation cloud: x, y uniform over the map, theta uniform in [-pi, pi)."""
    rng = np.random.default_rng(seed)
    p = np.zeros(n, PARTICLE_DTYPE)
    x = grid.origin_x + rng.random(n) * grid.width * grid.meters_per_cell
    y = grid.origin_y + rng.random(n) * grid.height * grid.meters_per_cell
    th = rng.uniform(-np.pi, np.pi, n)
    for k in ("pose", "parent_pose"):
        p[k]["x"], p[k]["y"], p[k]["theta"], p[k]["utime"] = x.astype(np.float32), y.astype(np.float32), \
            th.astype(np.float32), utime
    p["weight"] = 1.0 / n
    return p


def filter_shaped_weights(n, seed=0, clamped_frac=0.3):
    """Normalised weights of the filter's own form: max(score, 0.001)/sum with scores multiples of 0.5."""
    rng = np.random.default_rng(seed)
    s = rng.integers(1, 2 * 127 * 40, n).astype(np.float64) * 0.5
    s[rng.random(n) < clamped_frac] = 0.001
    return s / s.sum()


def make_pose(x, y, theta, utime=0):
    p = np.zeros((), POSE_DTYPE)
    p["x"], p["y"], p["theta"], p["utime"] = x, y, theta, utime
    return p
