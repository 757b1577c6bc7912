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


@pytest.fixture(scope="module")
def host_lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "botlab_b200", "src", "slam")])
    return C.CDLL(HOST_LIB)


def test_replay_api_is_exported(host_lib):
    assert hasattr(host_lib, "b200_replay_run")
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
