"""ctypes binding of botlab_b200/libmcl_cuda.so (C ABI: include/mcl_cuda.h).

This is plumbing for tests and bench.py: the product is the CUDA library and the C++ host classes in
botlab_b200/src/slam.  There is deliberately no CPU path here -- a missing library or GPU raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCL_LIB selects another build of the same library (kernel tuning sweeps); the default is the in-tree product build
LIB_PATH = os.environ.get("MCL_LIB") or os.path.join(_HERE, "libmcl_cuda.so")

POSE_DTYPE = np.dtype([("utime", "<i8"), ("x", "<f4"), ("y", "<f4"), ("theta", "<f4")], align=True)
PARTICLE_DTYPE = np.dtype([("pose", POSE_DTYPE), ("parent_pose", POSE_DTYPE), ("weight", "<f8")], align=True)
assert POSE_DTYPE.itemsize == 24 and PARTICLE_DTYPE.itemsize == 56

# every symbol include/mcl_cuda.h declares (tests check the library exports all of them)
SYMBOLS = [
    "mcl_default_params", "mcl_create", "mcl_destroy", "mcl_last_error", "mcl_stream", "mcl_sync",
    "mcl_comm_unique_id", "mcl_comm_init", "mcl_set_map", "mcl_update_map_rect", "mcl_read_map_rect", "mcl_map_update", "mcl_distance_grid", "mcl_init_at_pose",
    "mcl_init_uniform", "mcl_import_particles", "mcl_export_particles", "mcl_export_weighted", "mcl_action_reset", "mcl_action_update",
    "mcl_resample", "mcl_apply_action", "mcl_score", "mcl_normalize", "mcl_estimate", "mcl_update",
    "mcl_update_action_only", "mcl_upload_scan", "mcl_update_enqueue", "mcl_read_estimate", "mcl_get_stats",
    "mcl_set_gather_counting", "mcl_measure_gather_peak", "mcl_debug_sincosf", "mcl_debug_fast_trig_error", "mcl_debug_fast_margin", "mcl_debug_digest",
]


class Params(C.Structure):
    _fields_ = [("min_range", C.c_float), ("weight_floor", C.c_double), ("init_std", C.c_double),
                ("legacy_equal_utime", C.c_int), ("lanes_per_particle", C.c_int), ("map_tile", C.c_int),
                ("sensor_path", C.c_int), ("weight_mode", C.c_int), ("sensor_mode", C.c_int), ("lse_beta", C.c_double),
                ("reserved", C.c_int * 4)]


class Pose(C.Structure):
    _fields_ = [("utime", C.c_int64), ("x", C.c_float), ("y", C.c_float), ("theta", C.c_float)]


class Action(C.Structure):
    _fields_ = [("previous_odometry", Pose), ("initialized", C.c_int), ("moved", C.c_int), ("rot1", C.c_double),
                ("trans", C.c_double), ("rot2", C.c_double), ("rot1_std", C.c_double), ("trans_std", C.c_double),
                ("rot2_std", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("num_particles", C.c_int64), ("local_particles", C.c_int64), ("updates", C.c_int64),
                ("valid_beams", C.c_int64), ("evals", C.c_int64), ("gathers", C.c_int64),
                ("resample_overruns", C.c_int64), ("seq_fallback_chunks", C.c_int64), ("weight_sum", C.c_double),
                ("effective_sample_size", C.c_double), ("ms_resample", C.c_float), ("ms_action", C.c_float),
                ("ms_score", C.c_float), ("ms_normalize", C.c_float), ("ms_estimate", C.c_float),
                ("ms_total", C.c_float), ("lanes_per_particle", C.c_int), ("map_tile_used", C.c_int),
                ("kernel_launches", C.c_int), ("collectives", C.c_int), ("peer_push", C.c_int), ("sensor_path", C.c_int), ("table_variant", C.c_int),
                ("culled_beams", C.c_int),
                ("deferred_evals", C.c_int64), ("fast_eps", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class MclError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libmcl_cuda.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MclError(f"{LIB_PATH} is missing: build it with `make -C botlab_b200/csrc` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        vp, ip, i64, dp, fp = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_float
        L.mcl_default_params.argtypes = [vp]
        L.mcl_default_params.restype = None
        L.mcl_create.argtypes = [vp, i64, ip, vp]
        L.mcl_destroy.argtypes = [vp]
        L.mcl_destroy.restype = None
        L.mcl_last_error.argtypes = [vp]
        L.mcl_last_error.restype = C.c_char_p
        L.mcl_stream.argtypes = [vp]
        L.mcl_stream.restype = vp
        L.mcl_sync.argtypes = [vp]
        L.mcl_comm_unique_id.argtypes = [vp]
        L.mcl_comm_init.argtypes = [vp, vp, ip, ip]
        L.mcl_set_map.argtypes = [vp, vp, ip, ip, fp, fp, fp, fp]
        L.mcl_update_map_rect.argtypes = [vp, ip, ip, ip, ip, vp, ip]
        L.mcl_read_map_rect.argtypes = [vp, ip, ip, ip, ip, vp, ip]
        L.mcl_map_update.argtypes = [vp, vp, vp, ip, vp, vp, vp, ip, fp, ip, ip, vp]
        L.mcl_distance_grid.argtypes = [vp, vp]
        L.mcl_init_at_pose.argtypes = [vp, fp, fp, fp, i64, C.c_uint64]
        L.mcl_init_uniform.argtypes = [vp, i64, C.c_uint64]
        L.mcl_import_particles.argtypes = [vp, vp, i64]
        L.mcl_export_particles.argtypes = [vp, vp, i64, i64, vp]
        L.mcl_export_weighted.argtypes = [vp, vp, i64, dp, vp]
        L.mcl_action_reset.argtypes = [vp]
        L.mcl_action_reset.restype = None
        L.mcl_action_update.argtypes = [vp, vp]
        L.mcl_resample.argtypes = [vp, dp, vp, vp]
        L.mcl_apply_action.argtypes = [vp, vp, i64, vp]
        L.mcl_score.argtypes = [vp, vp, vp, vp, ip, vp]
        L.mcl_normalize.argtypes = [vp, vp]
        L.mcl_estimate.argtypes = [vp, vp]
        L.mcl_update.argtypes = [vp, vp, i64, vp, vp, vp, ip, dp, vp, vp]
        L.mcl_update_action_only.argtypes = [vp, vp, i64, vp]
        L.mcl_upload_scan.argtypes = [vp, vp, vp, vp, ip, i64]
        L.mcl_update_enqueue.argtypes = [vp, vp, i64, dp]
        L.mcl_read_estimate.argtypes = [vp, vp]
        L.mcl_get_stats.argtypes = [vp, vp]
        L.mcl_set_gather_counting.argtypes = [vp, ip]
        L.mcl_measure_gather_peak.argtypes = [vp, i64, i64, vp]
        L.mcl_debug_sincosf.argtypes = [vp, vp, i64, vp, vp]
        L.mcl_debug_fast_trig_error.argtypes = [vp, fp, fp, vp, vp]
        L.mcl_debug_fast_margin.argtypes = [vp, vp, vp, vp]
        L.mcl_debug_digest.argtypes = [vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_params(**overrides):
    p = Params()
    lib().mcl_default_params(C.addressof(p))
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


class ActionModel:
    """Host scalar half of the action model: ActionModel::updateAction (action_model.cpp:22-75)."""

    def __init__(self):
        self.c = Action()
        lib().mcl_action_reset(C.addressof(self.c))

    def update(self, x, y, theta, utime=0):
        o = Pose(utime, x, y, theta)
        return bool(lib().mcl_action_update(C.addressof(self.c), C.addressof(o)))

    @property
    def params(self):
        c = self.c
        return np.array([c.rot1, c.trans, c.rot2, c.rot1_std, c.trans_std, c.rot2_std])

    @property
    def moved(self):
        return bool(self.c.moved)


class Engine:
    """One GPU's MCL engine (mirrors ParticleFilter's life cycle, particle_filter.hpp:38-77)."""

    def __init__(self, num_particles, device=0, **params):
        self._L = lib()
        self.n = int(num_particles)
        p = default_params(**params)
        h = C.c_void_p()
        rc = self._L.mcl_create(C.addressof(p), self.n, device, C.addressof(h))
        if rc != 0:
            raise MclError(f"mcl_create failed ({rc}): {self._L.mcl_last_error(None).decode()}")
        self.h = h

    def _ck(self, rc):
        if rc != 0:
            raise MclError(f"libmcl_cuda error {rc}: {self._L.mcl_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self._L.mcl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- map
    def set_map(self, cells, origin_x, origin_y, meters_per_cell, cells_per_meter=None):
        cells = np.ascontiguousarray(cells, np.int8)
        hgt, wid = cells.shape
        if cells_per_meter is None:
            cells_per_meter = float(np.float32(1.0) / np.float32(meters_per_cell))   # occupancy_grid.cpp:30
        self._ck(self._L.mcl_set_map(self.h, _p(cells), wid, hgt, origin_x, origin_y, meters_per_cell, cells_per_meter))

    def update_map_rect(self, x0, y0, patch):
        patch = np.ascontiguousarray(patch, np.int8)
        self._ck(self._L.mcl_update_map_rect(self.h, x0, y0, patch.shape[1], patch.shape[0], _p(patch), patch.shape[1]))

    def read_map_rect(self, x0, y0, w, h):
        out = np.zeros((h, w), np.int8)
        self._ck(self._L.mcl_read_map_rect(self.h, x0, y0, w, h, _p(out), w))
        return out

    def map_update(self, previous, pose, initialized, ranges, thetas, times, max_laser_distance=5.0, hit_odds=3,
                   miss_odds=1):
        """Mapping::updateMap on the device mirror; previous / pose = (x, y, theta, utime).  Returns the rectangle
        (x0, y0, w, h) of cells that may have changed."""
        a, b = Pose(previous[3], *previous[:3]), Pose(pose[3], *pose[:3])
        ranges = np.ascontiguousarray(ranges, np.float32)
        thetas = np.ascontiguousarray(thetas, np.float32)
        times = np.ascontiguousarray(times, np.int64)
        rect = (C.c_int * 4)()
        self._ck(self._L.mcl_map_update(self.h, C.addressof(a), C.addressof(b), 1 if initialized else 0, _p(ranges),
                                        _p(thetas), _p(times), len(ranges), max_laser_distance, hit_odds, miss_odds,
                                        C.addressof(rect)))
        return tuple(rect)

    # ---- particles
    def init_at_pose(self, x, y, theta, utime=0, seed=1):
        self._ck(self._L.mcl_init_at_pose(self.h, x, y, theta, utime, seed))

    def init_uniform(self, utime=0, seed=1):
        self._ck(self._L.mcl_init_uniform(self.h, utime, seed))

    def import_particles(self, particles):
        particles = np.ascontiguousarray(particles, PARTICLE_DTYPE)
        self._ck(self._L.mcl_import_particles(self.h, _p(particles), particles.shape[0]))

    def export_particles(self, max_n=None, stride=1):
        max_n = self.n if max_n is None else max_n
        out = np.zeros(min(max_n, -(-self.n // stride)), PARTICLE_DTYPE)
        cnt = C.c_int64()
        self._ck(self._L.mcl_export_particles(self.h, _p(out), out.shape[0], stride, C.addressof(cnt)))
        return out[:cnt.value]

    # ---- stages
    def resample(self, r, weights=None, want_indices=True):
        w = None if weights is None else np.ascontiguousarray(weights, np.float64)
        idx = np.zeros(self.n, np.int32) if want_indices else None
        self._ck(self._L.mcl_resample(self.h, r, _p(w), _p(idx)))
        return idx

    def apply_action(self, action, utime=0, noise=None):
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
        self._ck(self._L.mcl_apply_action(self.h, C.addressof(action.c), utime, _p(nz)))

    def score(self, ranges, thetas, times, want_scores=True):
        ranges = np.ascontiguousarray(ranges, np.float32)
        thetas = np.ascontiguousarray(thetas, np.float32)
        times = np.ascontiguousarray(times, np.int64)
        out = np.zeros(self.n, np.float64) if want_scores else None
        self._ck(self._L.mcl_score(self.h, _p(ranges), _p(thetas), _p(times), len(ranges), _p(out)))
        return out

    def normalize(self, want_weights=True):
        out = np.zeros(self.n, np.float64) if want_weights else None
        self._ck(self._L.mcl_normalize(self.h, _p(out)))
        return out

    def estimate(self):
        o = Pose()
        self._ck(self._L.mcl_estimate(self.h, C.addressof(o)))
        return o

    # ---- fused
    def update(self, action, odometry_utime, ranges, thetas, times, r, noise=None):
        ranges = np.ascontiguousarray(ranges, np.float32)
        thetas = np.ascontiguousarray(thetas, np.float32)
        times = np.ascontiguousarray(times, np.int64)
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
        o = Pose()
        self._ck(self._L.mcl_update(self.h, C.addressof(action.c), odometry_utime, _p(ranges), _p(thetas), _p(times),
                                    len(ranges), r, _p(nz), C.addressof(o)))
        return o

    def update_action_only(self, action, odometry_utime, noise=None):
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
        self._ck(self._L.mcl_update_action_only(self.h, C.addressof(action.c), odometry_utime, _p(nz)))

    def upload_scan(self, ranges, thetas, times, odometry_utime):
        ranges = np.ascontiguousarray(ranges, np.float32)
        thetas = np.ascontiguousarray(thetas, np.float32)
        times = np.ascontiguousarray(times, np.int64)
        self._ck(self._L.mcl_upload_scan(self.h, _p(ranges), _p(thetas), _p(times), len(ranges), odometry_utime))

    def update_enqueue(self, action, odometry_utime, r=-1.0):
        self._ck(self._L.mcl_update_enqueue(self.h, C.addressof(action.c), odometry_utime, r))

    def read_estimate(self):
        o = Pose()
        self._ck(self._L.mcl_read_estimate(self.h, C.addressof(o)))
        return o

    def sync(self):
        self._ck(self._L.mcl_sync(self.h))

    @property
    def stream(self):
        return self._L.mcl_stream(self.h)

    # ---- introspection
    def stats(self):
        s = Stats()
        self._ck(self._L.mcl_get_stats(self.h, C.addressof(s)))
        return s.as_dict()

    def set_gather_counting(self, on):
        self._ck(self._L.mcl_set_gather_counting(self.h, 1 if on else 0))

    def measure_gather_peak(self, footprint_bytes, reads=1 << 30):
        out = C.c_double()
        self._ck(self._L.mcl_measure_gather_peak(self.h, footprint_bytes, reads, C.addressof(out)))
        return out.value

    def debug_sincosf(self, x):
        x = np.ascontiguousarray(x, np.float32)
        s = np.zeros_like(x)
        c = np.zeros_like(x)
        self._ck(self._L.mcl_debug_sincosf(self.h, _p(x), x.shape[0], _p(s), _p(c)))
        return s, c

    def fast_trig_error(self, lo, hi):
        """max |SFU sin/cos - double sin/cos| over every float in [lo, hi] -> (sin_err, cos_err)."""
        es, ec = C.c_double(), C.c_double()
        self._ck(self._L.mcl_debug_fast_trig_error(self.h, lo, hi, C.addressof(es), C.addressof(ec)))
        return es.value, ec.value

    def fast_margin(self):
        """(max endpoint deviation, max extended-point deviation, eps) of the float pass on the current inputs, cells."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._ck(self._L.mcl_debug_fast_margin(self.h, C.addressof(a), C.addressof(b), C.addressof(c)))
        return a.value, b.value, c.value

    def distance_grid(self, width, height):
        """The reference's ObstacleDistanceGrid of the device mirror, (H, W) float32."""
        out = np.zeros((height, width), np.float32)
        self._ck(self._L.mcl_distance_grid(self.h, _p(out)))
        return out

    def export_weighted(self, count, u01=0.5):
        """`count` particles drawn by systematic sampling over the weights, each with weight 1/count."""
        out = np.zeros(count, PARTICLE_DTYPE)
        got = C.c_int64()
        self._ck(self._L.mcl_export_weighted(self.h, _p(out), count, u01, C.addressof(got)))
        return out[:got.value]

    def digest(self):
        """Four 64-bit position-sensitive sums over this rank's slice (indices, scores, weights, poses)."""
        out = (C.c_uint64 * 4)()
        self._ck(self._L.mcl_debug_digest(self.h, out))
        return [int(v) for v in out]

    def comm_init(self, unique_id, rank, world):
        buf = (C.c_byte * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self._L.mcl_comm_init(self.h, C.addressof(buf), rank, world))


def comm_unique_id():
    buf = (C.c_byte * 128)()
    rc = lib().mcl_comm_unique_id(C.addressof(buf))
    if rc != 0:
        raise MclError(f"mcl_comm_unique_id failed: {lib().mcl_last_error(None).decode()}")
    return bytes(buf)
