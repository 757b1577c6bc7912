#!/usr/bin/env python
"""bench.py -- particle-beam evaluations/s of one MCL update (resample -> action -> sensor -> normalise -> estimate).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config config4] [--impl engine|reference]

A step = one full ParticleFilter::updateFilter over the whole particle cloud with a synthetic scan.
Workload (default): BASELINE.json configs[3] -- 16M particles x 360 beams on a 2000x2000 5 cm grid; the same total work
is sharded across N GPUs (strong scaling) with the map replicated.

`value`      whole-job evals/s with everything resident in HBM (updates enqueued back to back, CUDA events on the
             engine's stream, max over ranks).
`e2e`        the same metric through the reference-facing C-ABI call mcl_update() with HOST scan buffers: per step the
             scan is prepared and copied H2D, the pose estimate copied D2H, and the call blocks.
`roofline`   sensor stage (the dominant launches: score_fast_kernel + score_deferred_kernel, back to back on one
             stream): algorithmic bytes per update (32 B x map reads + 36 B x particles + map bytes) / the
             stage's mean CUDA-event duration, against MEASURED_PEAKS.json's HBM copy rate; plus the L2-gather
             microbenchmark measured on this device in this run as a second denominator.
`cpu_baseline` the reference's ParticleFilter::updateFilter (oracle/_ref when built, else the C port), 1 thread, on a
             bounded particle sub-sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from botlab_b200 import synth  # noqa: E402

METRIC = "particle_beam_evals_per_sec"
UNIT = "evals/s"
CPU_SAMPLE_PARTICLES = 100_000


def build_workload(config, seed=0):
    n, side = synth.CONFIGS[config]
    grid = synth.make_map(side, seed=synth.MAP_SEED + int(config[-1]))
    rng = np.random.default_rng(1234 + seed)
    truth = synth.find_free_pose(grid, rng)
    scans = []
    pose = truth
    t = 1_000_000
    for k in range(4):   # a few distinct scans along a short trajectory; steps cycle through them
        r, th, tt = synth.make_scan(grid, pose, seed=100 + k, t0=t)
        scans.append((pose, r, th, tt))
        pose = synth.odometry_step(rng, pose)
        t += 100_000
    return n, grid, truth, scans


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            try:
                sm.append(float(row[0])); mx.append(float(row[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, row[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def recorded_traffic(config, world, n):
    """DRAM bytes of one sensor-kernel launch from the committed ncu capture of this very workload, else None."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f).get(f"{config}_{world}gpu")
        if rec and rec["particles"] == n:
            return rec["dram_bytes_read"] + rec["dram_bytes_write"]
    except (OSError, ValueError, KeyError):
        pass
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_update(grid, truth, scans, n_sample, updates, threads=1):
    """Times the reference's ParticleFilter::updateFilter on the host.  Returns (evals/s, kind, seconds, evals)."""
    from oracle import ref, port
    cloud = synth.make_particles(n_sample, truth, seed=5, parent_utime=900_000, pose_utime=900_000)
    cloud["parent_pose"] = cloud["pose"]
    pose0, r0, th0, t0 = scans[0]
    total_s, total_evals = 0.0, 0
    if ref.available():
        kind = "reference"
        g = ref.RefGrid.from_cells(grid.cells, grid.origin_x, grid.origin_y, grid.meters_per_cell)
        pf = ref.RefParticleFilter(n_sample)
        pf.set_particles(cloud)
        pf.update(g, synth.make_pose(*pose0, utime=900_000), ref.Scan(r0, th0, t0), seed=1, want_draws=False)
        for k in range(updates):
            pose, r, th, t = scans[(k + 1) % len(scans)]
            odom = synth.make_pose(*pose, utime=int(t[-1]))
            _, moved, _, sec = pf.update(g, odom, ref.Scan(r, th, t), seed=k + 1, action_utime=int(t[-1]),
                                         want_draws=False)
            assert moved
            total_s += sec
            total_evals += n_sample * int((r > np.float32(0.15)).sum())
    else:
        kind = "port"
        g = port.Grid(grid.cells, grid.origin_x, grid.origin_y, grid.cells_per_meter)
        pf = port.ParticleFilter(cloud)
        pf.action.update(synth.make_pose(*pose0, utime=900_000))
        rng = port.Rng(5489)
        for k in range(updates):
            pose, r, th, t = scans[(k + 1) % len(scans)]
            odom = synth.make_pose(*pose, utime=int(t[-1]))
            t_a = time.perf_counter()
            # the port takes its draws as an input; generating them with the libstdc++ restatement is part of the
            # reference's applyAction cost, so it is inside the timed region
            probe = port.ActionModel()
            probe.c = type(pf.action.c).from_buffer_copy(pf.action.c)
            probe.update(odom)
            draws = probe.draws(rng, n_sample)
            pf.update(g, odom, r, th, t, 0.5 / n_sample, draws, action_utime=int(t[-1]))
            total_s += time.perf_counter() - t_a
            total_evals += n_sample * int((r > np.float32(0.15)).sum())
    return total_evals / total_s, kind, total_s, total_evals


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, grid, truth, scans = build_workload(args.config)
    n_sample = min(n, CPU_SAMPLE_PARTICLES)
    valid = int((scans[0][1] > np.float32(0.15)).sum())
    for _ in range(min(args.warmup, 1)):
        cpu_reference_update(grid, truth, scans, min(n_sample, 20_000), 1)
    t0 = time.perf_counter()
    value, kind, sec, evals = cpu_reference_update(grid, truth, scans, n_sample, args.steps)
    sample = (f"{args.steps} updateFilter calls on a {n_sample}-particle sub-sample of the {n}-particle cloud, same map "
              f"and scans; throughput is flat in N (SURVEY.md 6)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps * (n / n_sample), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64 (reference arithmetic)", "data": "synthetic",
        "config": {"workload": f"{args.config}: {n} particles x 360 beams ({valid} valid), "
                               f"{grid.width}x{grid.height} int8 grid, tracking cloud",
                   "note": "ms_per_step extrapolated linearly from the sub-sample to the full cloud; the reference's "
                           "ParticleFilter is single-threaded (no threads anywhere in src/slam), so 1 core is all it can use"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- engine arm
def run_engine(args):
    import torch
    import torch.distributed as dist
    from botlab_b200 import engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n, grid, truth, scans = build_workload(args.config)
    if args.particles:
        n = args.particles
    e = engine.Engine(n, device=local_rank, lanes_per_particle=args.lanes, map_tile=args.tile,
                      sensor_path=args.sensor_path)
    if world > 1:
        if rank == 0:
            uid = torch.tensor(list(engine.comm_unique_id()), dtype=torch.uint8, device="cuda")
        else:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        e.comm_init(bytes(uid.cpu().tolist()), rank, world)
    e.set_map(grid.cells, grid.origin_x, grid.origin_y, grid.meters_per_cell, grid.cells_per_meter)
    pose0, r0, th0, t0 = scans[0]
    uniform = args.config == "config5" or args.uniform     # global localisation: uniformly initialised cloud
    if uniform:
        e.init_uniform(utime=int(t0[0]), seed=42)
    else:
        e.init_at_pose(*pose0, utime=int(t0[0]), seed=42)
    am = engine.ActionModel()
    am.update(*pose0, int(t0[0]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        tns = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    stream = torch.cuda.ExternalStream(e.stream, device=torch.device("cuda", local_rank))
    step_no = [0]

    def next_inputs():
        k = step_no[0]
        step_no[0] += 1
        pose, r, th, _ = scans[(k + 1) % len(scans)]
        # one 10 Hz sweep ending at this update's odometry time, as OccupancyGridSLAM pairs them
        # (slam.cpp:227: odometry is sampled at scan.times.back())
        ut = int(t0[0]) + 100_000 * (k + 1)
        t = ut - 100_000 + (np.arange(len(r), dtype=np.int64) * 100_000) // len(r) + 100_000 // len(r)
        # odometry alternates so the action model always reports motion
        am.update(pose[0] + 1e-3 * (k % 7), pose[1], pose[2], ut)
        assert am.moved
        return r, th, t, ut

    valid = int((r0 > np.float32(0.15)).sum())
    h2d = valid * 16 + 64            # prepared beams (16 B each) + scalars
    d2h = 16 + 40                    # pose estimate + counters

    # ---- e2e arm: blocking C-ABI calls with host scan buffers ---------------------------------------------------
    for _ in range(args.warmup):
        r, th, t, ut = next_inputs()
        e.update(am, ut, r, th, t, 0.5 / n)
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    score_ms, stage_ms, evals_e2e, launches = [], [], 0, 0
    for _ in range(args.steps):
        r, th, t, ut = next_inputs()
        if uniform:
            e.init_uniform(utime=ut - 100_000, seed=1000 + step_no[0])   # every step scores a fresh uniform cloud
        e.update(am, ut, r, th, t, 0.5 / n)
        st = e.stats()
        score_ms.append(st["ms_score"])
        stage_ms.append([st["ms_resample"], st["ms_action"], st["ms_score"], st["ms_normalize"], st["ms_estimate"]])
        evals_e2e += n * int((r > np.float32(0.15)).sum())
        launches += st["kernel_launches"]
    ev1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1))

    # ---- resident arm: scan uploaded once, updates enqueued back to back ------------------------------------------
    r, th, t, ut = next_inputs()
    e.upload_scan(r, th, t, ut)
    for _ in range(args.warmup):
        e.update_enqueue(am, ut)
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    for k in range(args.steps):
        if uniform:
            e.init_uniform(utime=ut - 100_000, seed=2000 + k)
        e.update_enqueue(am, ut)
    ev3.record(stream)
    barrier()
    res_ms = max_over_ranks(ev2.elapsed_time(ev3))
    clock_info = clocks.stop()
    evals_res = n * int((r > np.float32(0.15)).sum()) * args.steps
    est = e.read_estimate()

    # ---- algorithmic traffic of the sensor kernel: one untimed counted pass --------------------------------------
    e.set_gather_counting(True)
    local_n = e.stats()["local_particles"]
    r, th, t, ut = next_inputs()
    if uniform:
        e.init_uniform(utime=ut - 100_000, seed=3000)
    e.update(am, ut, r, th, t, 0.5 / n)
    st = e.stats()
    e.set_gather_counting(False)
    gathers = st["gathers"]
    alg_bytes = 32 * gathers + 36 * local_n + grid.width * grid.height
    mean_score_s = float(np.mean(score_ms)) * 1e-3
    peak, peak_kind = measured_peaks()
    achieved = alg_bytes / mean_score_s / 1e9
    gather_peak = e.measure_gather_peak(grid.width * grid.height, 1 << 31)

    line = None
    if rank == 0:
        stage = np.mean(np.array(stage_ms), axis=0)
        line = {
            "metric": METRIC, "value": evals_res / (res_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32+f64 (reference arithmetic), int32 scores, int8 map",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: {n} particles x 360 beams ({valid} valid), "
                                   f"{grid.width}x{grid.height} int8 grid, "
                                   + ("uniform cloud re-initialised every step (global localisation)" if uniform
                                      else "tracking cloud"),
                       "updates_per_sec": args.steps / (res_ms * 1e-3),
                       "l2_policy": "inputs larger than L2: 28 B/particle of pose+parent+score state streams from HBM "
                                    f"every step ({28 * n / 1e6:.0f} MB); the int8 map is L2/shared-memory resident by design",
                       "lanes_per_particle": st["lanes_per_particle"], "map_tile_used": st["map_tile_used"],
                       "sensor_path": st["sensor_path"], "deferred_fraction": st["deferred_evals"] / max(st["evals"], 1),
                       "certification_eps_cells": st["fast_eps"],
                       "particles_per_gpu": local_n},
            "e2e": {"value": evals_e2e / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "updates_per_sec": args.steps / (e2e_ms * 1e-3)},
            "gpu_launches": launches,
            "clocks": clock_info,
            "roofline": {"bound": "hbm",
                         "kernel": {3: "score_table_kernel (sensor stage)",
                                    2: "score_fast_kernel + score_deferred_kernel (sensor stage)"}.get(
                                        st["sensor_path"], "score_kernel (sensor stage)"),
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": recorded_traffic(args.config, world, n),
                         "peak_kind": peak_kind,
                         "algorithmic_bytes_per_launch": alg_bytes, "map_reads_per_launch": gathers,
                         "kernel_ms": mean_score_s * 1e3,
                         "l2_gather": {"achieved_sectors_per_s": gathers / mean_score_s,
                                       "peak_sectors_per_s": gather_peak,
                                       "frac": gathers / mean_score_s / gather_peak,
                                       "peak_kind": "measured in this run: random 1-byte ld.global.cg over the map footprint"}},
            "stage_ms": {"resample": float(stage[0]), "action": float(stage[1]), "score": float(stage[2]),
                         "normalize": float(stage[3]), "estimate": float(stage[4])},
            "estimate": [est.x, est.y, est.theta],
        }
    if rank == 0 and world == 1 and not args.no_cpu:
        n_sample = min(n, CPU_SAMPLE_PARTICLES)
        v, kind, sec, ev = cpu_reference_update(grid, truth, scans, n_sample, 2)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"2 updateFilter calls on a {n_sample}-particle sub-sample of the same "
                                          f"workload ({sec:.1f} s of CPU)", "host_cores_available": os.cpu_count()}
    if rank == 0:
        print(json.dumps(line))
    e.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="config4", choices=sorted(synth.CONFIGS))
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--sensor-path", type=int, default=0, help="0 = certified float pass + exact re-evaluation, 1 = exact only")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--uniform", action="store_true", help="uniform cloud (global localisation) on any config")
    ap.add_argument("--particles", type=int, default=0, help="override the config's particle count")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
