"""The C++ host classes (botlab_b200/src/slam: ParticleFilter, ActionModel, SensorModel, MovingLaserScan,
OccupancyGrid) keep the reference's public interface and reach the GPU only through the C ABI.  A C++ driver
(tests/csrc/host_api_test.cpp) uses them the way OccupancyGridSLAM does; these tests build and run it."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

HOST_DIR = os.path.join(ROOT, "botlab_b200", "src", "slam")
DRIVER = os.path.join(ROOT, "botlab_b200", "host_api_test")
HOST_LIB = os.path.join(ROOT, "botlab_b200", "libslam_b200.so")


def write_map_file(path, grid):
    """The reference's ASCII .map format (OccupancyGrid::saveToFile)."""
    with open(path, "w") as f:
        f.write(f"{grid.origin_x:g} {grid.origin_y:g} {grid.width} {grid.height} {grid.meters_per_cell:g}\n")
        for row in grid.cells:
            f.write(" ".join(str(int(v)) for v in row) + " \n")


@pytest.fixture(scope="module")
def built():
    subprocess.check_call(["make", "-s", "-C", HOST_DIR])
    return DRIVER


def test_host_library_exports_the_reference_interface(built):
    syms = subprocess.check_output(["nm", "-DC", "--defined-only", HOST_LIB]).decode()
    for want in ["ParticleFilter::ParticleFilter(int)",
                 "ParticleFilter::initializeFilterAtPose(pose_xyt_t const&)",
                 "ParticleFilter::updateFilter(pose_xyt_t const&, lidar_t const&, OccupancyGrid const&)",
                 "ParticleFilter::updateFilterActionOnly(pose_xyt_t const&)",
                 "ParticleFilter::poseEstimate() const",
                 "ParticleFilter::particles() const",
                 "ActionModel::updateAction(pose_xyt_t const&)",
                 "ActionModel::applyAction(particle_t const&)",
                 "SensorModel::likelihood(particle_t const&, lidar_t const&, OccupancyGrid const&)",
                 "MovingLaserScan::MovingLaserScan(lidar_t const&, pose_xyt_t const&, pose_xyt_t const&, int)",
                 "OccupancyGrid::logOdds(int, int) const",
                 "OccupancyGrid::loadFromFile(",
                 "OccupancyGrid::saveToFile("]:
        assert want in syms, want
    # and it reaches the device only through the C ABI
    undefined = subprocess.check_output(["nm", "-D", "--undefined-only", HOST_LIB]).decode()
    assert "mcl_update" in undefined and "mcl_create" in undefined and "cuda" not in undefined.lower()


def test_driver_fails_loudly_without_gpu(built, tmp_path, real_map):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    mp = str(tmp_path / "real.map")
    write_map_file(mp, real_map)
    res = subprocess.run([built, mp, "100", "1"], capture_output=True, text=True)
    assert res.returncode == 10 and "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_host_classes_drive_the_engine(built, tmp_path, real_map):
    mp = str(tmp_path / "real.map")
    write_map_file(mp, real_map)
    res = subprocess.run([built, mp, "20000", "6"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    out = json.loads(res.stdout)
    assert (out["width"], out["height"], out["cpm"]) == (200, 200, 20.0)
    assert out["kat_scores"] == [3138.0, 1322.5, 1787.5, 2969.0]
    k = load_golden("kat")
    assert np.array_equal(np.float32(out["ray0"]), k["rays_b"][0]) and out["rays"] == 360        # B5
    assert np.array_equal(np.float32(out["ray200"]), k["rays_b"][200])           # B6
    assert out["action"][0] == 1 and np.array_equal(out["action"][1:], k["action_forward"][:3])   # B8
    ax, ay, ath, apx, autime = out["applied"]
    assert abs(ax - 0.5) < 0.1 and abs(ay + 0.25) < 0.1 and apx == np.float32(0.5) and autime == 5
    track = np.array(out["track"])
    assert len(track) == 6
    assert np.abs(track[:, 3] - track[:, 0]).max() < 0.10 and np.abs(track[:, 4] - track[:, 1]).max() < 0.10
    assert np.abs(track[:, 5] - track[:, 2]).max() < 0.10
    assert (track[:, 6] == 1000000 + 100000 * np.arange(1, 7)).all() and (track[:, 7] == 1).all()
    assert out["exported"] == 1000 and 0.0 < out["exported_weight"] < 1.0        # every 20th particle of 20000
    assert out["updates"] == 6 and out["evals"] == 20000 * 360 and out["launches"] > 0
    inc, full, before, after, want = out["mirrors"]
    assert inc == full and inc != out["kat_scores"][0]        # the second mirror saw the block the first one consumed
    assert before == 0.0 and after == want and want > 0.0     # an assigned-to grid is mirrored again
    assert out["action_only"][0] == out["action_only"][1]
