"""Headless replay of OccupancyGridSLAM's loop (BASELINE configs[0]).

Engine side: HeadlessSLAM (botlab_b200/src/slam/headless_slam.cpp) = the reference's queueing / PoseTrace sampling /
init-at-first-scan / sanity gate / localise-then-map flow around the GPU ParticleFilter, driven through b200_replay_run.
Oracle side: ref_replay_run (oracle/ref_harness.cpp) = the same flow around the reference's OWN ParticleFilter, Mapping,
PoseTrace and OccupancyGrid objects.  Both replay one synthesised log (the reference's .log is not in the checkout) on
the real 10 m x 10 m map with the reference's default 200 particles, the same initial cloud, the same rand() seed and
the reference's recorded mt19937 action draws."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from botlab_b200 import synth
from oracle import port, ref

HOST_LIB = os.path.join(ROOT, "botlab_b200", "libslam_b200.so")

ARGTYPES = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]


def make_log(grid, steps=30, beams=290, seed=4):
    """10 Hz scans (one sweep per 100 ms) and 50 Hz odometry along a gentle arc from (0,0,0)."""
    rng = np.random.default_rng(seed)
    t0 = 5_000_000
    truth = lambda t: (0.3 * (t - t0) * 1e-6 * np.cos(0.15 * (t - t0) * 1e-6), 0.3 * (t - t0) * 1e-6 * np.sin(0.15 * (t - t0) * 1e-6),
                       0.3 * (t - t0) * 1e-6)
    odom_t = np.arange(t0 - 40_000, t0 + (steps + 1) * 100_000 + 40_000, 20_000, dtype=np.int64)
    odom = np.array([truth(t) for t in odom_t], np.float64)
    odom[:, :2] += rng.normal(0, 5e-4, odom[:, :2].shape)          # sim.py-like odometry noise
    odom[:, 2] += rng.normal(0, 1e-3, len(odom))
    offsets, ranges, thetas, times = [0], [], [], []
    for k in range(steps):
        t_end = t0 + (k + 1) * 100_000
        r, th, _ = synth.make_scan(grid, truth(t_end - 50_000), num_beams=beams, seed=100 + k)
        tt = t_end - 100_000 + ((np.arange(beams) + 1) * 100_000) // beams
        ranges.append(r); thetas.append(th); times.append(tt)
        offsets.append(offsets[-1] + beams)
    return dict(offsets=np.array(offsets, np.int32), ranges=np.concatenate(ranges).astype(np.float32),
                thetas=np.concatenate(thetas).astype(np.float32), times=np.concatenate(times).astype(np.int64),
                odom_t=odom_t, odom=odom.astype(np.float32), truth=truth, t0=t0, steps=steps)


def run_replay(fn, grid, log, n, init_cloud, noise, mode=1, seed=1):
    cells = np.ascontiguousarray(grid.cells, np.int8)
    poses = np.zeros((log["steps"], 5), np.float32)
    final_map = np.zeros_like(cells)
    iters = C.c_int()
    err = C.create_string_buffer(512)
    init3 = np.zeros(3, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    fn.argtypes = ARGTYPES
    rc = fn(p(cells), grid.width, grid.height, grid.origin_x, grid.origin_y, grid.meters_per_cell, 1, n, mode, 4, 1,
            5.0, log["steps"], p(log["offsets"]), p(log["ranges"]), p(log["thetas"]), p(log["times"]), len(log["odom_t"]),
            p(log["odom_t"]), p(log["odom"]), p(init3), seed, p(init_cloud), p(noise), p(poses), p(final_map),
            C.addressof(iters), err, 512)
    assert rc == 0, err.value
    return poses[:iters.value], final_map, iters.value


# ---- LCM event log: an independent (Python) writer of the format lcm-logger produces -----------------------------------
import struct

M64 = (1 << 64) - 1


def _s64(v):
    v &= M64
    return v - (1 << 64) if v >> 63 else v


def _hash_update(v, c):                     # lcmgen.c: v = ((v<<8) ^ (v>>55)) + c on an int64_t
    return _s64((_s64(v << 8) ^ (v >> 55)) + c)


def _hash_string(v, s):
    v = _hash_update(v, len(s))
    for ch in s.encode():
        v = _hash_update(v, ch)
    return v


def lcm_fingerprint(members):
    """members: [(name, primitive type, [variable-array size fields])] -> the int64 lcm-gen puts in front of a message."""
    v = 0x12345678
    for name, typ, dims in members:
        v = _hash_string(v, name)
        v = _hash_string(v, typ)
        v = _hash_update(v, len(dims))
        for d in dims:
            v = _hash_update(v, 1)          # LCM_VAR
            v = _hash_string(v, d)
    u = v & M64
    return _s64((u << 1) + (u >> 63))


LIDAR_FP = lcm_fingerprint([("utime", "int64_t", []), ("num_ranges", "int32_t", []), ("ranges", "float", ["num_ranges"]),
                            ("thetas", "float", ["num_ranges"]), ("times", "int64_t", ["num_ranges"]),
                            ("intensities", "float", ["num_ranges"])])
ODOM_FP = lcm_fingerprint([("utime", "int64_t", []), ("x", "float", []), ("y", "float", []), ("theta", "float", [])])


def write_lcm_log(path, log, garbage_after=None):
    """Events in arrival order (odometry at its utime, a scan once its last ray is measured), as b200_replay_run delivers
    them; a TRUE_POSE event the SLAM loop's replay must skip; optionally a stretch of garbage between two events."""
    events = []
    for i, t in enumerate(log["odom_t"]):
        x, y, th = (float(v) for v in log["odom"][i])
        events.append((int(t), 0, "ODOMETRY", struct.pack(">qqfff", ODOM_FP, int(t), x, y, th)))
    for k in range(log["steps"]):
        a, b = int(log["offsets"][k]), int(log["offsets"][k + 1])
        n = b - a
        body = struct.pack(">qqi", LIDAR_FP, int(log["times"][b - 1]), n)
        body += log["ranges"][a:b].astype(">f4").tobytes() + log["thetas"][a:b].astype(">f4").tobytes()
        body += log["times"][a:b].astype(">i8").tobytes() + np.zeros(n, ">f4").tobytes()
        events.append((int(log["times"][b - 1]), 1, "LIDAR", body))
    events.sort(key=lambda e: (e[0], e[1]))
    with open(path, "wb") as f:
        for num, (t, _, chan, body) in enumerate(events):
            if num == 3:
                f.write(struct.pack(">IqqII", 0xEDA1DA01, 1000 + num, t, 9, 28) + b"TRUE_POSE" + b"\0" * 28)
            f.write(struct.pack(">IqqII", 0xEDA1DA01, num, t, len(chan), len(body)) + chan.encode() + body)
            if garbage_after is not None and num == garbage_after:
                f.write(bytes(range(7, 200)))
    return len(events) + 1


def test_lcm_log_reader_decodes_an_independently_written_log(host_lib, real_map, tmp_path):
    log = make_log(real_map, steps=6)
    path = str(tmp_path / "replay.log")
    total = write_lcm_log(path, log, garbage_after=10)
    counts = (C.c_int * 6)()
    first = (C.c_float * 4)()
    fps = (C.c_int64 * 2)()
    host_lib.b200_scan_log.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
    assert host_lib.b200_scan_log(path.encode(), counts, first, fps) == 0
    events, scans, odoms, bad, fp_mismatch, resyncs = list(counts)
    assert (events, scans, odoms, bad, fp_mismatch, resyncs) == (total, 6, len(log["odom_t"]), 0, 0, 1)
    assert (fps[0], fps[1]) == (LIDAR_FP, ODOM_FP)
    b0 = int(log["offsets"][1])
    assert first[0] == b0 and first[1] == log["ranges"][0] and first[2] == log["thetas"][1]
    assert first[3] == float(int(log["times"][2]) % 1_000_000)
    assert host_lib.b200_scan_log(str(tmp_path / "missing.log").encode(), counts, first, fps) == -101


@pytest.fixture(scope="module")
def host_lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "botlab_b200", "src", "slam")])
    return C.CDLL(HOST_LIB)


def test_replay_api_is_exported(host_lib):
    assert hasattr(host_lib, "b200_replay_run") and hasattr(host_lib, "b200_replay_log") and hasattr(host_lib, "b200_scan_log")
    syms = subprocess.check_output(["nm", "-DC", "--defined-only", HOST_LIB]).decode()
    for want in ["HeadlessSLAM::runSLAMIteration()", "HeadlessSLAM::handleLaser(lidar_t const&)",
                 "HeadlessSLAM::handleOdometry(pose_xyt_t const&)", "PoseTrace::poseAt(long) const",
                 "Mapping::updateMap(lidar_t const&, pose_xyt_t const&, OccupancyGrid&)"]:
        assert want in syms, want


@pytest.mark.gpu
@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_localization_only_replay_matches_the_reference_loop(host_lib, real_map):
    n = 200                                                        # slam_main.cpp:21
    log = make_log(real_map)
    t_front = int(log["times"][0])
    init_cloud = port.init_at_pose(port.Rng(7), synth.make_pose(0, 0, 0, utime=t_front), n)
    noise = np.zeros((log["steps"], n, 3), np.float32)
    ref_poses, ref_map, ref_iters = run_replay(ref.lib().ref_replay_run, real_map, log, n, init_cloud, noise)
    eng_poses, eng_map, eng_iters = run_replay(host_lib.b200_replay_run, real_map, log, n, init_cloud, noise)
    assert ref_iters == eng_iters == log["steps"]
    assert (ref_poses[:, 3] == 1).all() and (eng_poses[:, 3] == 1).all()
    assert np.array_equal(ref_poses[:, 4], eng_poses[:, 4])        # same pose utimes (PoseTrace sampling)
    d_xy = np.abs(ref_poses[:, :2] - eng_poses[:, :2]).max()
    d_th = np.abs(ref_poses[:, 2] - eng_poses[:, 2]).max()
    # The estimates feed Mapping, whose cell writes feed the next scores, so a last-place difference in the estimate
    # (float running sum in the reference, double in the engine) can move a ray endpoint to the neighbouring cell and the
    # two runs are not required to stay bit-identical; in practice they agree to ~1e-6 m.
    assert d_xy <= 2e-3 and d_th <= 2e-3, (d_xy, d_th)
    assert (ref_map != eng_map).mean() <= 1e-3
    assert np.abs(ref_map.astype(int) - eng_map.astype(int)).max() <= 8
    # and both track the synthesised truth
    for k in (log["steps"] // 2, log["steps"] - 1):
        tx, ty, tth = log["truth"](log["t0"] + (k + 1) * 100_000)
        assert abs(eng_poses[k, 0] - tx) < 0.15 and abs(eng_poses[k, 1] - ty) < 0.15
    # the map really changed during "localization-only" (slam.cpp:276)
    assert (eng_map != real_map.cells).sum() > 100
    print(f"replay: max |dxy| {d_xy:.2e} m, max |dtheta| {d_th:.2e} rad, differing cells {(ref_map != eng_map).sum()}")
    # the same replay with Mapping::updateMap running on the device mirror (mcl_map_update): identical poses and map
    os.environ["B200_DEVICE_MAPPING"] = "1"
    try:
        dev_poses, dev_map, dev_iters = run_replay(host_lib.b200_replay_run, real_map, log, n, init_cloud, noise)
    finally:
        del os.environ["B200_DEVICE_MAPPING"]
    assert dev_iters == eng_iters
    assert np.array_equal(dev_poses, eng_poses) and np.array_equal(dev_map, eng_map)


@pytest.mark.gpu
def test_replay_from_an_lcm_log_equals_replay_from_arrays(host_lib, real_map, tmp_path):
    """`slam --localization-only map < lcm-logplayer file.log`, headless: the event-log reader feeds HeadlessSLAM the same
    message sequence b200_replay_run builds from arrays, so the pose trace and the final map are bit-identical."""
    n = 2000
    log = make_log(real_map, steps=12)
    path = str(tmp_path / "replay.log")
    write_lcm_log(path, log)
    cells = np.ascontiguousarray(real_map.cells, np.int8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    init3 = np.zeros(3, np.float32)
    # arrays: no planted cloud, no injected draws -> the filter's own Philox stream with ParticleFilter's default seed
    poses_a = np.zeros((log["steps"], 5), np.float32); map_a = np.zeros_like(cells)
    iters, err = C.c_int(), C.create_string_buffer(512)
    host_lib.b200_replay_run.argtypes = ARGTYPES
    rc = host_lib.b200_replay_run(p(cells), real_map.width, real_map.height, real_map.origin_x, real_map.origin_y,
                                  real_map.meters_per_cell, 1, n, 1, 4, 1, 5.0, log["steps"], p(log["offsets"]),
                                  p(log["ranges"]), p(log["thetas"]), p(log["times"]), len(log["odom_t"]), p(log["odom_t"]),
                                  p(log["odom"]), p(init3), 1, None, None, p(poses_a), p(map_a), C.addressof(iters), err, 512)
    assert rc == 0 and iters.value == log["steps"], err.value
    poses_l = np.zeros((log["steps"], 5), np.float32); map_l = np.zeros_like(cells)
    iters_l, counts = C.c_int(), (C.c_int * 6)()
    host_lib.b200_replay_log.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_uint, C.c_uint64,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
    rc = host_lib.b200_replay_log(path.encode(), p(cells), real_map.width, real_map.height, real_map.origin_x,
                                  real_map.origin_y, real_map.meters_per_cell, 1, n, 1, 4, 1, 5.0, p(init3), 1, 0x6d636c,
                                  log["steps"], p(poses_l), p(map_l), C.addressof(iters_l), counts, err, 512)
    assert rc == 0 and iters_l.value == log["steps"], err.value
    assert counts[1] == log["steps"] and counts[2] == len(log["odom_t"]) and counts[3] == 0 and counts[4] == 0
    assert np.array_equal(poses_l, poses_a) and np.array_equal(map_l, map_a)
    k = log["steps"] - 1
    tx, ty, _ = log["truth"](log["t0"] + (k + 1) * 100_000)
    assert abs(poses_l[k, 0] - tx) < 0.15 and abs(poses_l[k, 1] - ty) < 0.15
