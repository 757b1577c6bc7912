"""ORACLE / TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libbotlab_ref.so.

That library is the UNMODIFIED reference hot path (src/slam/{particle_filter,action_model,sensor_model,
moving_laser_scan,occupancy_grid}.cpp) compiled by oracle/Makefile plus oracle/ref_harness.cpp.  It is only imported by
tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke(); never by botlab_b200/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libbotlab_ref.so")

POSE_DTYPE = np.dtype([("utime", "<i8"), ("x", "<f4"), ("y", "<f4"), ("theta", "<f4")], align=True)
PARTICLE_DTYPE = np.dtype([("pose", POSE_DTYPE), ("parent_pose", POSE_DTYPE), ("weight", "<f8")], align=True)
assert POSE_DTYPE.itemsize == 24 and PARTICLE_DTYPE.itemsize == 56

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libbotlab_ref.so missing: run `make -C oracle` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        vp, ip, fp, dp = C.c_void_p, C.c_int, C.c_float, C.c_double
        L.ref_grid_new.restype = vp
        L.ref_grid_new.argtypes = [vp, ip, ip, fp, fp, fp]
        L.ref_grid_load.restype = vp
        L.ref_grid_load.argtypes = [C.c_char_p, fp, fp, fp]
        L.ref_grid_info.argtypes = [vp] + [vp] * 6
        L.ref_grid_cells.argtypes = [vp, vp]
        L.ref_grid_logodds.argtypes = [vp, ip, ip]
        L.ref_grid_free.argtypes = [vp]
        L.ref_likelihood.argtypes = [vp, vp, ip, vp, vp, vp, ip, vp]
        L.ref_moving_scan.argtypes = [vp, vp, vp, ip, vp, vp, vp]
        L.ref_action_new.restype = vp
        L.ref_action_free.argtypes = [vp]
        L.ref_action_seed.argtypes = [vp, C.c_uint]
        L.ref_action_set_utime.argtypes = [vp, C.c_int64]
        L.ref_action_update.argtypes = [vp, vp, vp]
        L.ref_action_apply.argtypes = [vp, vp, vp, ip, vp]
        L.ref_pf_new.restype = vp
        L.ref_pf_new.argtypes = [ip]
        L.ref_pf_free.argtypes = [vp]
        L.ref_pf_set_particles.argtypes = [vp, vp, ip]
        L.ref_pf_get_particles.argtypes = [vp, vp, ip]
        L.ref_pf_init_at_pose.argtypes = [vp, vp]
        L.ref_pf_set_action_utime.argtypes = [vp, C.c_int64]
        L.ref_pf_seed_action.argtypes = [vp, C.c_uint]
        L.ref_resample_draw.restype = dp
        L.ref_resample_draw.argtypes = [C.c_uint, ip]
        L.ref_pf_resample.argtypes = [vp, C.c_uint, vp]
        L.ref_pf_normalize.argtypes = [vp, vp, vp, ip, vp, vp, vp, ip, vp]
        L.ref_pf_estimate.argtypes = [vp, vp, ip, vp]
        L.ref_pf_update.argtypes = [vp, vp, vp, vp, vp, vp, ip, C.c_uint, C.c_int64, vp, vp, vp]
        L.ref_pf_update_action_only.argtypes = [vp, vp, C.c_int64, vp, vp]
        L.ref_pf_pose_estimate.argtypes = [vp, vp]
        L.ref_pf_action_params.argtypes = [vp, vp]
        L.ref_map_update.argtypes = [vp, vp, vp, ip, vp, vp, vp, ip, fp, ip, ip]
        L.ref_distance_grid.argtypes = [vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_pose(x, y, theta, utime=0):
    p = np.zeros((), POSE_DTYPE)
    p["x"], p["y"], p["theta"], p["utime"] = x, y, theta, utime
    return p


class Scan:
    """lidar_t surface (lcmtypes/lidar_t.lcm:1-14): ranges f32, thetas f32, times i64."""

    def __init__(self, ranges, thetas, times):
        self.ranges = np.ascontiguousarray(ranges, np.float32)
        self.thetas = np.ascontiguousarray(thetas, np.float32)
        self.times = np.ascontiguousarray(times, np.int64)
        assert self.ranges.shape == self.thetas.shape == self.times.shape
        self.n = int(self.ranges.shape[0])


class RefGrid:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference grid construction failed")
        self.h = handle

    @classmethod
    def from_cells(cls, cells, origin_x, origin_y, meters_per_cell):
        cells = np.ascontiguousarray(cells, np.int8)
        hgt, wid = cells.shape
        return cls(lib().ref_grid_new(_p(cells), wid, hgt, origin_x, origin_y, meters_per_cell))

    @classmethod
    def from_file(cls, path, width_m=10.0, height_m=10.0, meters_per_cell=0.05):
        return cls(lib().ref_grid_load(path.encode(), width_m, height_m, meters_per_cell))

    def info(self):
        w, h = C.c_int(), C.c_int()
        ox, oy, mpc, cpm = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        lib().ref_grid_info(self.h, *[C.addressof(v) for v in (w, h, ox, oy, mpc, cpm)])
        return dict(width=w.value, height=h.value, origin_x=ox.value, origin_y=oy.value,
                    meters_per_cell=mpc.value, cells_per_meter=cpm.value)

    def cells(self):
        i = self.info()
        out = np.zeros((i["height"], i["width"]), np.int8)
        lib().ref_grid_cells(self.h, _p(out))
        return out

    def __del__(self):
        try:
            lib().ref_grid_free(self.h)
        except Exception:
            pass


def likelihood(grid, particles, scan):
    particles = np.ascontiguousarray(particles, PARTICLE_DTYPE)
    out = np.zeros(particles.shape[0], np.float64)
    lib().ref_likelihood(grid.h, _p(particles), particles.shape[0], _p(scan.ranges), _p(scan.thetas), _p(scan.times),
                         scan.n, _p(out))
    return out


def map_update(grid, previous, pose, initialized, scan, max_laser_distance=5.0, hit_odds=3, miss_odds=1):
    """Mapping::updateMap (mapping.cpp:17-40) on `grid` (a RefGrid, modified in place)."""
    a = np.ascontiguousarray(previous, POSE_DTYPE).reshape(1)
    b = np.ascontiguousarray(pose, POSE_DTYPE).reshape(1)
    lib().ref_map_update(grid.h, _p(a), _p(b), 1 if initialized else 0, _p(scan.ranges), _p(scan.thetas),
                         _p(scan.times), len(scan.ranges), max_laser_distance, hit_odds, miss_odds)


def distance_grid(grid):
    """ObstacleDistanceGrid::setDistances of the compiled reference on `grid` (a RefGrid): float32 [H, W]."""
    w, h = grid.info()["width"], grid.info()["height"]
    out = np.zeros((h, w), np.float32)
    lib().ref_distance_grid(grid.h, _p(out))
    return out


def moving_scan(scan, begin, end):
    rays = np.zeros((scan.n, 4), np.float32)
    b = np.ascontiguousarray(begin, POSE_DTYPE)
    e = np.ascontiguousarray(end, POSE_DTYPE)
    k = lib().ref_moving_scan(_p(scan.ranges), _p(scan.thetas), _p(scan.times), scan.n, _p(b), _p(e), _p(rays))
    return rays[:k]


class RefActionModel:
    def __init__(self, seed=None):
        self.h = lib().ref_action_new()
        if seed is not None:
            lib().ref_action_seed(self.h, seed)

    def set_utime(self, t):
        lib().ref_action_set_utime(self.h, t)

    def update(self, odom):
        o = np.ascontiguousarray(odom, POSE_DTYPE)
        out = np.zeros(6, np.float64)
        moved = lib().ref_action_update(self.h, _p(o), _p(out))
        return bool(moved), out

    def apply(self, particles):
        particles = np.ascontiguousarray(particles, PARTICLE_DTYPE)
        n = particles.shape[0]
        out = np.zeros(n, PARTICLE_DTYPE)
        draws = np.zeros((n, 3), np.float32)
        lib().ref_action_apply(self.h, _p(particles), _p(out), n, _p(draws))
        return out, draws

    def __del__(self):
        try:
            lib().ref_action_free(self.h)
        except Exception:
            pass


class RefParticleFilter:
    def __init__(self, n):
        self.n = n
        self.h = lib().ref_pf_new(n)

    def set_particles(self, particles):
        particles = np.ascontiguousarray(particles, PARTICLE_DTYPE)
        assert particles.shape[0] == self.n
        lib().ref_pf_set_particles(self.h, _p(particles), self.n)

    def particles(self):
        out = np.zeros(self.n, PARTICLE_DTYPE)
        lib().ref_pf_get_particles(self.h, _p(out), self.n)
        return out

    def init_at_pose(self, pose):
        p = np.ascontiguousarray(pose, POSE_DTYPE)
        lib().ref_pf_init_at_pose(self.h, _p(p))

    def seed_action(self, seed):
        lib().ref_pf_seed_action(self.h, seed)

    def resample(self, seed=1):
        idx = np.zeros(self.n, np.int32)
        lib().ref_pf_resample(self.h, seed, _p(idx))
        return idx

    def normalize(self, grid, proposal, scan):
        proposal = np.ascontiguousarray(proposal, PARTICLE_DTYPE)
        out = np.zeros(proposal.shape[0], PARTICLE_DTYPE)
        lib().ref_pf_normalize(self.h, grid.h, _p(proposal), proposal.shape[0], _p(scan.ranges), _p(scan.thetas),
                               _p(scan.times), scan.n, _p(out))
        return out

    def estimate(self, particles):
        particles = np.ascontiguousarray(particles, PARTICLE_DTYPE)
        out = np.zeros((), POSE_DTYPE)
        lib().ref_pf_estimate(self.h, _p(particles), particles.shape[0], _p(out))
        return out

    def update(self, grid, odom, scan, seed=1, action_utime=0, want_draws=True):
        """Returns (pose, moved, draws[N,3] or None, seconds)."""
        o = np.ascontiguousarray(odom, POSE_DTYPE)
        pose = np.zeros((), POSE_DTYPE)
        draws = np.zeros((self.n, 3), np.float32) if want_draws else None
        sec = C.c_double()
        moved = lib().ref_pf_update(self.h, grid.h, _p(o), _p(scan.ranges), _p(scan.thetas), _p(scan.times), scan.n,
                                    seed, action_utime, _p(pose), _p(draws), C.addressof(sec))
        return pose, bool(moved), draws, sec.value

    def update_action_only(self, odom, action_utime=0):
        o = np.ascontiguousarray(odom, POSE_DTYPE)
        pose = np.zeros((), POSE_DTYPE)
        draws = np.zeros((self.n, 3), np.float32)
        moved = lib().ref_pf_update_action_only(self.h, _p(o), action_utime, _p(pose), _p(draws))
        return pose, bool(moved), draws

    def pose_estimate(self):
        pose = np.zeros((), POSE_DTYPE)
        lib().ref_pf_pose_estimate(self.h, _p(pose))
        return pose

    def __del__(self):
        try:
            lib().ref_pf_free(self.h)
        except Exception:
            pass


def resample_draw(seed, n):
    return lib().ref_resample_draw(seed, n)
