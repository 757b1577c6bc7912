"""ORACLE / TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libmcl_oracle.so (the plain-C restatement,
oracle/mcl_oracle.c).  Imported only by tests/, bench.py's CPU-baseline legs and __graft_entry__.smoke()."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcl_oracle.so")

POSE_DTYPE = np.dtype([("utime", "<i8"), ("x", "<f4"), ("y", "<f4"), ("theta", "<f4")], align=True)
PARTICLE_DTYPE = np.dtype([("pose", POSE_DTYPE), ("parent_pose", POSE_DTYPE), ("weight", "<f8")], align=True)


class _Grid(C.Structure):
    _fields_ = [("cells", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("origin_x", C.c_float),
                ("origin_y", C.c_float), ("cells_per_meter", C.c_float)]


class _Action(C.Structure):
    _fields_ = [("prev", C.c_byte * 24), ("initialized", C.c_int), ("moved", C.c_int), ("rot1", C.c_double),
                ("trans", C.c_double), ("rot2", C.c_double), ("rot1_std", C.c_double), ("trans_std", C.c_double),
                ("rot2_std", C.c_double)]


class _Rng(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int), ("saved_available", C.c_int), ("saved", C.c_double)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libmcl_oracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, ip, dp = C.c_void_p, C.c_int, C.c_double
        L.orc_wrap_to_pi.restype = C.c_float
        L.orc_wrap_to_pi.argtypes = [C.c_float]
        L.orc_moving_scan.argtypes = [vp, vp, vp, ip, vp, vp, vp]
        L.orc_score_ray.restype = dp
        L.orc_score_ray.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_float, vp]
        L.orc_likelihood.argtypes = [vp, vp, ip, vp, vp, vp, ip, vp, vp, vp]
        L.orc_action_init.argtypes = [vp]
        L.orc_action_update.argtypes = [vp, vp]
        L.orc_action_apply.argtypes = [vp, C.c_int64, vp, vp, ip, vp]
        L.orc_normalize.argtypes = [vp, ip, vp, vp]
        L.orc_resample.argtypes = [vp, ip, dp, vp]
        L.orc_estimate.argtypes = [vp, ip, vp]
        L.orc_rng_seed.argtypes = [vp, C.c_uint32]
        L.orc_rng_next.restype = C.c_uint32
        L.orc_rng_next.argtypes = [vp]
        L.orc_rng_normal.restype = dp
        L.orc_rng_normal.argtypes = [vp, dp, dp, ip]
        L.orc_action_draws.argtypes = [vp, vp, ip, vp]
        L.orc_init_at_pose.argtypes = [vp, vp, vp, ip]
        L.orc_update.argtypes = [vp, vp, vp, vp, ip, vp, C.c_int64, vp, vp, vp, ip, dp, vp, vp]
        L.orc_ray_scores.argtypes = [vp, vp, vp, vp, vp, ip, vp]
        L.orc_map_update.restype = C.c_long
        L.orc_map_update.argtypes = [vp, ip, ip, C.c_float, C.c_float, C.c_float, vp, vp, ip, vp, vp, vp, ip, C.c_float,
                                     ip, ip]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Grid:
    """int8 log-odds grid with the reference's geometry (occupancy_grid.hpp:84-123)."""

    def __init__(self, cells, origin_x, origin_y, cells_per_meter):
        self.cells = np.ascontiguousarray(cells, np.int8)
        self.height, self.width = self.cells.shape
        self.origin_x, self.origin_y, self.cells_per_meter = float(origin_x), float(origin_y), float(cells_per_meter)
        self.c = _Grid(self.cells.ctypes.data, self.width, self.height, origin_x, origin_y, cells_per_meter)

    @property
    def ptr(self):
        return C.addressof(self.c)


def moving_scan(ranges, thetas, times, begin, end):
    ranges = np.ascontiguousarray(ranges, np.float32)
    thetas = np.ascontiguousarray(thetas, np.float32)
    times = np.ascontiguousarray(times, np.int64)
    b = np.ascontiguousarray(begin, POSE_DTYPE)
    e = np.ascontiguousarray(end, POSE_DTYPE)
    rays = np.zeros((len(ranges), 4), np.float32)
    k = lib().orc_moving_scan(_p(ranges), _p(thetas), _p(times), len(ranges), _p(b), _p(e), _p(rays))
    return rays[:k]


def likelihood(grid, particles, ranges, thetas, times):
    """Returns (scores f64[N], total map reads, total valid particle-beam evaluations)."""
    particles = np.ascontiguousarray(particles, PARTICLE_DTYPE)
    ranges = np.ascontiguousarray(ranges, np.float32)
    thetas = np.ascontiguousarray(thetas, np.float32)
    times = np.ascontiguousarray(times, np.int64)
    out = np.zeros(particles.shape[0], np.float64)
    g, e = C.c_int64(0), C.c_int64(0)
    lib().orc_likelihood(grid.ptr, _p(particles), particles.shape[0], _p(ranges), _p(thetas), _p(times), len(ranges),
                         _p(out), C.addressof(g), C.addressof(e))
    return out, g.value, e.value


def map_update(cells, origin_x, origin_y, cells_per_meter, previous, pose, initialized, ranges, thetas, times,
               max_laser_distance=5.0, hit_odds=3, miss_odds=1):
    """Mapping::updateMap (mapping.cpp:17-127): returns the updated copy of `cells` (int8 [H, W])."""
    out = np.ascontiguousarray(cells, np.int8).copy()
    a = np.ascontiguousarray(previous, POSE_DTYPE).reshape(1)
    b = np.ascontiguousarray(pose, POSE_DTYPE).reshape(1)
    ranges = np.ascontiguousarray(ranges, np.float32)
    thetas = np.ascontiguousarray(thetas, np.float32)
    times = np.ascontiguousarray(times, np.int64)
    lib().orc_map_update(_p(out), out.shape[1], out.shape[0], origin_x, origin_y, cells_per_meter, _p(a), _p(b),
                         1 if initialized else 0, _p(ranges), _p(thetas), _p(times), len(ranges), max_laser_distance,
                         hit_odds, miss_odds)
    return out


def distance_grid(cells, thr=0):
    """ObstacleDistanceGrid::setDistances (planning/obstacle_distance_grid.cpp:44-188): float32 [H, W]."""
    cells = np.ascontiguousarray(cells, np.int8)
    out = np.zeros(cells.shape, np.float32)
    L = lib()
    L.orc_distance_grid.restype = C.c_long
    L.orc_distance_grid.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int, C.c_void_p]
    L.orc_distance_grid(_p(cells), cells.shape[1], cells.shape[0], thr, _p(out))
    return out


def likelihood_field(grid, particles, ranges, thetas, times):
    """Scores of the engine's likelihood-field sensor mode (extension; see mcl_oracle.c)."""
    particles = np.ascontiguousarray(particles, PARTICLE_DTYPE)
    ranges = np.ascontiguousarray(ranges, np.float32)
    thetas = np.ascontiguousarray(thetas, np.float32)
    times = np.ascontiguousarray(times, np.int64)
    out = np.zeros(particles.shape[0], np.float64)
    L = lib()
    L.orc_likelihood_field.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.orc_likelihood_field(grid.ptr, _p(particles), particles.shape[0], _p(ranges), _p(thetas), _p(times), len(ranges), _p(out))
    return out


def ray_scores(grid, particle, ranges, thetas, times):
    """Per-ray scores of one particle over its valid beams, in scan order (their sum is likelihood()'s entry)."""
    part = np.ascontiguousarray(particle, PARTICLE_DTYPE).reshape(1)
    ranges = np.ascontiguousarray(ranges, np.float32)
    thetas = np.ascontiguousarray(thetas, np.float32)
    times = np.ascontiguousarray(times, np.int64)
    out = np.zeros(len(ranges), np.float64)
    k = lib().orc_ray_scores(grid.ptr, _p(part), _p(ranges), _p(thetas), _p(times), len(ranges), _p(out))
    return out[:k]


class ActionModel:
    def __init__(self):
        self.c = _Action()
        lib().orc_action_init(C.addressof(self.c))

    def update(self, odom):
        o = np.ascontiguousarray(odom, POSE_DTYPE)
        moved = lib().orc_action_update(C.addressof(self.c), _p(o))
        return bool(moved), np.array([self.c.rot1, self.c.trans, self.c.rot2, self.c.rot1_std, self.c.trans_std,
                                      self.c.rot2_std])

    def apply(self, particles, draws, utime=0):
        particles = np.ascontiguousarray(particles, PARTICLE_DTYPE)
        draws = np.ascontiguousarray(draws, np.float32)
        out = np.zeros(particles.shape[0], PARTICLE_DTYPE)
        lib().orc_action_apply(C.addressof(self.c), utime, _p(particles), _p(out), particles.shape[0], _p(draws))
        return out

    def draws(self, rng, n):
        d = np.zeros((n, 3), np.float32)
        lib().orc_action_draws(C.addressof(rng.c), C.addressof(self.c), n, _p(d))
        return d


class Rng:
    def __init__(self, seed=5489):
        self.c = _Rng()
        lib().orc_rng_seed(C.addressof(self.c), seed)

    def next_u32(self):
        return lib().orc_rng_next(C.addressof(self.c))

    def normal(self, mean=0.0, std=1.0, fresh=True):
        return lib().orc_rng_normal(C.addressof(self.c), mean, std, 1 if fresh else 0)


def init_at_pose(rng, pose, n):
    p = np.ascontiguousarray(pose, POSE_DTYPE)
    out = np.zeros(n, PARTICLE_DTYPE)
    lib().orc_init_at_pose(C.addressof(rng.c), _p(p), _p(out), n)
    return out


def normalize(scores):
    scores = np.ascontiguousarray(scores, np.float64)
    w = np.zeros_like(scores)
    s = C.c_double()
    lib().orc_normalize(_p(scores), len(scores), _p(w), C.addressof(s))
    return w, s.value


def resample(weights, r):
    """Returns (indices int32[N], overruns)."""
    weights = np.ascontiguousarray(weights, np.float64)
    idx = np.zeros(len(weights), np.int32)
    over = lib().orc_resample(_p(weights), len(weights), r, _p(idx))
    return idx, over


def estimate(particles):
    particles = np.ascontiguousarray(particles, PARTICLE_DTYPE)
    out = np.zeros((), POSE_DTYPE)
    lib().orc_estimate(_p(particles), particles.shape[0], _p(out))
    return out


class ParticleFilter:
    """updateFilter with the uniform draw and the action draws injected (particle_filter.cpp:37-52)."""

    def __init__(self, particles):
        self.particles = np.ascontiguousarray(particles, PARTICLE_DTYPE).copy()
        self.scratch = np.zeros_like(self.particles)
        self.action = ActionModel()
        self.pose = np.zeros((), POSE_DTYPE)

    def update(self, grid, odom, ranges, thetas, times, r, draws, action_utime=0):
        o = np.ascontiguousarray(odom, POSE_DTYPE)
        ranges = np.ascontiguousarray(ranges, np.float32)
        thetas = np.ascontiguousarray(thetas, np.float32)
        times = np.ascontiguousarray(times, np.int64)
        draws = np.ascontiguousarray(draws, np.float32)
        moved = lib().orc_update(C.addressof(self.action.c), grid.ptr, _p(self.particles), _p(self.scratch),
                                 self.particles.shape[0], _p(o), action_utime, _p(ranges), _p(thetas), _p(times),
                                 len(ranges), r, _p(draws), _p(self.pose))
        return self.pose.copy(), bool(moved)
