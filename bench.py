#!/usr/bin/env python
"""bench.py -- particle-beam evaluations/s of one MCL update (resample -> action -> sensor -> normalise -> estimate).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config config4] [--impl engine|reference]

A step = one full ParticleFilter::updateFilter over the whole particle cloud with a synthetic scan.
Workload (default): BASELINE.json configs[3] -- 16M particles x 360 beams on a 2000x2000 5 cm grid, SURVEY.md 8d's
tracking cloud (truth + N(0, 0.10 m), N(0, 0.05 rad)); the same total work is sharded across N GPUs (strong scaling)
with the map replicated.

`value`      whole-job evals/s with everything resident in HBM (updates enqueued back to back, CUDA events on the
             engine's stream, max over ranks).
`e2e`        the same metric through the reference-facing C-ABI call mcl_update() with HOST scan buffers: per step the
             scan is prepared and copied H2D, the pose estimate copied D2H, and the call blocks.
`roofline`   sensor stage (the dominant kernel).  Its map reads are shared-memory lookups, so it is bound by instruction
             issue: achieved warp instructions/s (instructions per 32 evaluations from the committed ncu capture of
             that kernel x this run's evaluations / this run's CUDA-event time) against 4 schedulers x SMs x the SM
             clock sampled in this run; plus the shared-memory wavefront fraction and the HBM share.
`digest`     sums mod 2^64 over all ranks of the last update's indices / scores / weights / poses: equal across GPU
             counts iff the clouds are bit-identical.
`cpu_baseline` the reference's ParticleFilter::updateFilter (oracle/_ref when built, else the C port), 1 thread, on a
             bounded particle sub-sample of the same workload (1 M particles at config 4).
`configs`    short entries for the other BASELINE configs (1, 2, 3 on one GPU; the 64 M-particle config 5 on eight).
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


INTERIOR_POSE = [False]     # --pose interior: the robot in the middle of the map instead of wherever the seed put it


def build_workload(config, seed=0, interior=None):
    n, side = synth.CONFIGS[config]
    grid = synth.make_map(side, seed=synth.MAP_SEED + int(config[-1]))
    rng = np.random.default_rng(1234 + seed)
    truth = synth.find_free_pose(grid, rng)
    if INTERIOR_POSE[0] if interior is None else interior:
        # The seeded pose of configs 3-5 lies 1.5-5 m from the map's corner, so a window around the cloud is clipped
        # by the grid; this variant keeps the whole 8 m reach inside the map (the largest window the scan can ask for).
        cx, cy = grid.origin_x + 0.5 * grid.width * grid.meters_per_cell, grid.origin_y + 0.5 * grid.height * grid.meters_per_cell
        while max(abs(truth[0] - cx), abs(truth[1] - cy)) > 0.15 * grid.width * grid.meters_per_cell:
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


def kernel_figures():
    """Per-kernel figures taken from the committed ncu captures (profiles/r02_kernel_figures.json): warp instructions
    and shared-memory wavefronts per 32 evaluations of the sensor kernels, with the capture each came from."""
    path = os.path.join(ROOT, "profiles", "r02_kernel_figures.json")
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tracking_cloud(n, truth, t_first, seed=5):
    """SURVEY.md 8d cloud of the tracking configs: truth + N(0, 0.10 m) / N(0, 0.05 rad), then one reference-sigma action
    step so that parent_pose != pose; pose.utime = the first scan's start."""
    return synth.make_particles(n, truth, seed=seed, sigma_xy=0.10, sigma_theta=0.05, parent_utime=int(t_first) - 100_000,
                                pose_utime=int(t_first))


# ---------------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_update(grid, truth, scans, n_sample, updates, threads=1):
    """Times the reference's ParticleFilter::updateFilter on the host.  Returns (evals/s, kind, seconds, evals)."""
    from oracle import ref, port
    pose0, r0, th0, t0 = scans[0]
    cloud = tracking_cloud(n_sample, truth, 900_000 + 100_000)
    cloud["pose"]["utime"] = 900_000
    cloud["parent_pose"] = cloud["pose"]
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


def cpu_sample_size(n, updates, budget_s):
    """Particles per CPU update so that `updates` reference updates take about budget_s (the reference runs about
    3e7 evals/s on one core, flat in N), never more than the cloud or 1 M, never fewer than 100 K
    (or the cloud)."""
    per_update = budget_s * 3.0e7 / (max(updates, 1) * 357.0)
    return int(min(n, max(min(n, 100_000), min(1_000_000, per_update))))


def workload_name(config, n, valid, grid, uniform):
    return (f"{config}: {n} particles x 360 beams ({valid} valid), {grid.width}x{grid.height} int8 grid, "
            + ("uniform cloud re-initialised every step (global localisation)" if uniform
               else "tracking cloud (truth + N(0, 0.10 m), N(0, 0.05 rad))")
            + (", robot in the map's interior" if INTERIOR_POSE[0] else ""))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, grid, truth, scans = build_workload(args.config)
    if args.particles:
        n = args.particles
    uniform = args.config == "config5" or args.uniform
    n_sample = cpu_sample_size(n, args.steps + min(args.warmup, 1), 150.0)
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
        "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64 (reference arithmetic), int32 scores, int8 map",
        "data": "synthetic",
        "config": {"workload": workload_name(args.config, n, valid, grid, uniform)},
        "details": {"note": "ms_per_step extrapolated linearly from the sub-sample to the full cloud; the reference's "
                            "ParticleFilter is single-threaded (no threads anywhere in src/slam), so 1 core is all it can use"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- engine arm
class Harness:
    """torch.distributed plumbing shared by the measurements of one bench.py process."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        tns = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(tns, op=self.dist.ReduceOp.MAX)
        return float(tns.item())

    def sum_u64(self, vals):
        """Sum mod 2^64 over ranks (int64 adds wrap)."""
        as_i64 = [v - (1 << 64) if v >= (1 << 63) else v for v in vals]
        if self.world == 1:
            return [v & ((1 << 64) - 1) for v in as_i64]
        tns = self.torch.tensor(as_i64, dtype=self.torch.int64, device="cuda")
        self.dist.all_reduce(tns, op=self.dist.ReduceOp.SUM)
        return [int(v) & ((1 << 64) - 1) for v in tns.tolist()]

    def make_engine(self, n, args):
        from botlab_b200 import engine
        e = engine.Engine(n, device=self.local_rank, lanes_per_particle=args.lanes, map_tile=args.tile,
                          sensor_path=args.sensor_path)
        if self.world > 1:
            t = self.torch
            uid = t.tensor(list(engine.comm_unique_id()) if self.rank == 0 else [0] * 128, dtype=t.uint8, device="cuda")
            self.dist.broadcast(uid, 0)
            e.comm_init(bytes(uid.cpu().tolist()), self.rank, self.world)
        return e


def measure(hx, args, config, steps, warmup, particles=0, uniform=False, want_roofline=True):
    """One workload on the process's ranks: e2e arm (blocking C-ABI updates with host scan buffers), resident arm
    (updates enqueued back to back), a counted pass for the algorithmic figures.  Returns the JSON fields (rank 0)."""
    from botlab_b200 import engine
    torch = hx.torch
    rank, world = hx.rank, hx.world
    n, grid, truth, scans = build_workload(config)
    if particles:
        n = particles
    e = hx.make_engine(n, args)
    e.set_map(grid.cells, grid.origin_x, grid.origin_y, grid.meters_per_cell, grid.cells_per_meter)
    pose0, r0, th0, t0 = scans[0]
    uniform = uniform or config == "config5"
    if uniform:
        e.init_uniform(utime=int(t0[0]), seed=42)
    else:
        e.import_particles(tracking_cloud(n, truth, t0[0]))      # every rank imports the same seeded cloud
    am = engine.ActionModel()
    am.update(*pose0, int(t0[0]))
    stream = torch.cuda.ExternalStream(e.stream, device=torch.device("cuda", hx.local_rank))
    step_no = [0]

    prepared = {}

    def scan_inputs(k):
        """The k-th update's scan buffers (host arrays, made ahead of the timed loops: they are the sensor's output, not
        the filter's work)."""
        if k not in prepared:
            pose, r, th, _ = scans[(k + 1) % len(scans)]
            # one 10 Hz sweep ending at this update's odometry time, as OccupancyGridSLAM pairs them
            # (slam.cpp:227: odometry is sampled at scan.times.back())
            ut = int(t0[0]) + 100_000 * (k + 1)
            t = ut - 100_000 + (np.arange(len(r), dtype=np.int64) * 100_000) // len(r) + 100_000 // len(r)
            prepared[k] = (pose, np.ascontiguousarray(r, np.float32), np.ascontiguousarray(th, np.float32),
                           np.ascontiguousarray(t, np.int64), ut, int((r > np.float32(0.15)).sum()))
        return prepared[k]

    def next_inputs():
        k = step_no[0]
        step_no[0] += 1
        pose, r, th, t, ut, _ = scan_inputs(k)
        # odometry alternates so the action model always reports motion (ActionModel::updateAction runs per update)
        am.update(pose[0] + 1e-3 * (k % 7), pose[1], pose[2], ut)
        assert am.moved
        return r, th, t, ut

    valid = int((r0 > np.float32(0.15)).sum())
    h2d = valid * 16 + 64            # prepared beams (16 B each) + scalars
    d2h = 80                         # one readback struct: counters + pose estimate

    # ---- e2e arm: blocking C-ABI calls with host scan buffers ---------------------------------------------------
    for _ in range(warmup):
        r, th, t, ut = next_inputs()
        e.update(am, ut, r, th, t, 0.5 / n)
    for k in range(step_no[0], step_no[0] + steps):
        scan_inputs(k)
    clocks = ClockSampler(hx.local_rank)
    hx.barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    score_ms, stage_ms, evals_e2e, launches = [], [], 0, 0
    for _ in range(steps):
        r, th, t, ut = next_inputs()
        if uniform:
            e.init_uniform(utime=ut - 100_000, seed=1000 + step_no[0])   # every step scores a fresh uniform cloud
        e.update(am, ut, r, th, t, 0.5 / n)
        st = e.stats()
        score_ms.append(st["ms_score"])
        stage_ms.append([st["ms_resample"], st["ms_action"], st["ms_score"], st["ms_normalize"], st["ms_estimate"]])
        evals_e2e += n * scan_inputs(step_no[0] - 1)[5]
        launches += st["kernel_launches"]
    ev1.record(stream)
    hx.barrier()
    e2e_ms = hx.max_over_ranks(ev0.elapsed_time(ev1))

    # ---- resident arm: scan uploaded once, updates enqueued back to back ------------------------------------------
    r, th, t, ut = next_inputs()
    e.upload_scan(r, th, t, ut)
    for _ in range(warmup):
        e.update_enqueue(am, ut)
    hx.barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    for k in range(steps):
        if uniform:
            e.init_uniform(utime=ut - 100_000, seed=2000 + k)
        e.update_enqueue(am, ut)
    ev3.record(stream)
    hx.barrier()
    res_ms = hx.max_over_ranks(ev2.elapsed_time(ev3))
    clock_info = clocks.stop()
    evals_res = n * int((r > np.float32(0.15)).sum()) * steps
    est = e.read_estimate()
    digest = hx.sum_u64(e.digest())

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
    mean_score_s = float(np.mean(score_ms)) * 1e-3
    out = None
    if rank == 0:
        stage = np.mean(np.array(stage_ms), axis=0)
        out = {
            "value": evals_res / (res_ms * 1e-3), "ms_per_step": res_ms / steps, "n_gpus": world, "steps": steps,
            "warmup": warmup,
            "config": {"workload": workload_name(config, n, valid, grid, uniform)},
            "details": {"updates_per_sec": steps / (res_ms * 1e-3),
                       "l2_policy": "inputs larger than L2: 28 B/particle of pose+parent+score state streams from HBM "
                                    f"every step ({28 * n / 1e6:.0f} MB); the int8 map is shared-memory resident by design"
                                    if 28 * n > 126e6 else
                                    "particle state smaller than L2 (it is rewritten by every update; no flush between steps)",
                       "lanes_per_particle": st["lanes_per_particle"], "map_tile_used": st["map_tile_used"],
                       "sensor_path": st["sensor_path"], "table_variant": st["table_variant"], "culled_beams": st["culled_beams"],
                       "deferred_fraction": st["deferred_evals"] / max(st["evals"], 1),
                       "certification_eps_cells": st["fast_eps"], "particles_per_gpu": local_n},
            "e2e": {"value": evals_e2e / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / steps,
                    "updates_per_sec": steps / (e2e_ms * 1e-3)},
            "gpu_launches": launches, "clocks": clock_info,
            "stage_ms": {"resample": float(stage[0]), "action": float(stage[1]), "score": float(stage[2]),
                         "normalize": float(stage[3]), "estimate": float(stage[4])},
            "estimate": [est.x, est.y, est.theta],
            "digest": {"resample_indices": f"{digest[0]:016x}", "scores": f"{digest[1]:016x}",
                       "weights": f"{digest[2]:016x}", "poses": f"{digest[3]:016x}",
                       "what": "sums mod 2^64 over all ranks of mixed (global particle index, bit pattern) pairs after "
                               "the last resident update: equal across GPU counts iff the clouds are bit-identical"},
        }
    if want_roofline:
        # measured on every rank (it launches kernels), reported by rank 0
        gather_peak = e.measure_gather_peak(grid.width * grid.height, 1 << 31)
        if rank == 0:
            out["roofline"] = roofline(st, gathers, local_n, grid, mean_score_s, clock_info, gather_peak)
    e.close()
    return out, (n, grid, truth, scans)


def roofline(st, gathers, local_n, grid, score_s, clock_info, gather_peak):
    """What bounds the sensor stage (the dominant kernel).  The map reads are shared-memory lookups, so neither HBM nor
    L2 is the roofline: the kernel is bound by instruction issue.  achieved = warp instructions per second = (warp
    instructions per 32 evaluations, from the committed ncu capture of this kernel) x evaluations / 32 / the stage's
    CUDA-event time measured in this run; peak = 4 schedulers x SMs x the SM clock sampled in this run.  Beside it: the
    shared-memory wavefront rate against one wavefront per clock per SM, and the HBM share (particle state + map once)."""
    fig = kernel_figures()
    kern = {3: "score_table_kernel", 2: "score_fast_kernel+score_deferred_kernel"}.get(st["sensor_path"], "score_kernel")
    f = fig.get(kern, {})
    if st["sensor_path"] == 3:
        # four variants of the kernel (one window / one per batch of particles, 16- / 8-bit classes), each with its capture
        variant = {0: "one window, 16-bit classes", 1: "one window, 8-bit classes", 2: "window per batch, 16-bit classes",
                   3: "window per batch, 8-bit classes"}[st["table_variant"]]
        f = fig.get(f"score_table_kernel/v{st['table_variant']}", f)
        kern = f"score_table_kernel<{variant}>"
    sm_mhz = clock_info.get("sm_mhz") or 1965.0
    sms = 148
    evals = st["evals"]
    if st["sensor_path"] == 3 and st.get("culled_beams", 0) > 0 and local_n > 0:
        # beams culled for the whole slice are not executed: the instruction figure applies to the evaluations that are
        evals = evals * (1.0 - st["culled_beams"] / (evals / local_n))
    peak_issue = 4 * sms * sm_mhz * 1e6
    peak, peak_kind = measured_peaks()
    hbm_bytes = 36 * local_n + grid.width * grid.height
    out = {"bound": "issue", "kernel": kern + " (sensor stage)", "kernel_ms": score_s * 1e3, "unit": "Gwarp-inst/s",
           "peak": peak_issue / 1e9,
           "peak_kind": f"4 warp schedulers x {sms} SMs x {sm_mhz:.0f} MHz (SM clock sampled by nvidia-smi during this run)",
           "traffic": None,
           "traffic_note": "DRAM bytes need a profiler: see traffic_ncu (from the committed capture), not measured in this run",
           "traffic_ncu": f.get("dram_bytes_per_launch"),
           "hbm": {"algorithmic_bytes_per_launch": hbm_bytes, "achieved": hbm_bytes / score_s / 1e9, "peak": peak,
                   "unit": "GB/s", "frac": hbm_bytes / score_s / 1e9 / peak, "peak_kind": peak_kind,
                   "what": "36 B x particles + map bytes once (SURVEY.md 8d without the map-read term, which is "
                           "served from shared memory)"},
           "map_reads_per_launch": gathers,
           "l2_gather_reference": {"map_reads_per_s": gathers / score_s, "l2_random_sector_rate": gather_peak,
                                   "ratio": gathers / score_s / gather_peak,
                                   "what": "the rate the SURVEY's L2-gather design would be bound by (random 1-byte "
                                           "ld.global.cg over the map footprint, measured in this run); the kernel "
                                           "does not touch L2 for map reads, so this is context, not a roofline fraction"}}
    if f.get("warp_inst_per_32_evals"):
        ach = f["warp_inst_per_32_evals"] * evals / 32.0 / score_s
        out.update({"achieved": ach / 1e9, "frac": ach / peak_issue,
                    "warp_inst_per_32_evals": f["warp_inst_per_32_evals"], "figures_from": f.get("source")})
        if f.get("smem_wavefronts_per_32_evals"):
            wf = f["smem_wavefronts_per_32_evals"] * evals / 32.0 / score_s
            out["smem"] = {"achieved_wavefronts_per_s": wf, "peak_wavefronts_per_s": sms * sm_mhz * 1e6,
                           "frac": wf / (sms * sm_mhz * 1e6), "what": "shared-memory wavefronts against 1 per clock per SM"}
    else:
        out.update({"achieved": None, "frac": None,
                    "note": "no committed instruction figure for this kernel family (profiles/r02_kernel_figures.json)"})
    return out


def run_engine(args):
    hx = Harness()
    rank, world = hx.rank, hx.world
    head, (n, grid, truth, scans) = measure(hx, args, args.config, args.steps, args.warmup, particles=args.particles,
                                            uniform=args.uniform)
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32+f64 (reference arithmetic), int32 scores, int8 map",
                "data": "synthetic"}
        for k in ("config", "details", "e2e", "gpu_launches", "clocks", "roofline", "stage_ms", "estimate", "digest"):
            line[k] = head[k]
    if rank == 0 and world == 1 and not args.no_cpu:
        n_sample = cpu_sample_size(n, 2, 25.0)
        v, kind, sec, ev = cpu_reference_update(grid, truth, scans, n_sample, 2)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"2 updateFilter calls on a {n_sample}-particle sub-sample of the same "
                                          f"workload ({sec:.1f} s of CPU)", "host_cores_available": os.cpu_count()}
    # ---- the other BASELINE configs, briefly (device-timed; their own CPU baselines at BASELINE.md 3's sizes) ----------
    if args.config == "config4" and not args.no_extra and not args.particles:
        extra = []
        if world == 1:
            for cfg, k, cpu_updates in (("config1", 50, 5), ("config2", 20, 5), ("config3", 10, 1)):
                res, (cn, cgrid, ctruth, cscans) = measure(hx, args, cfg, k, 3, want_roofline=False)
                entry = {"config": res["config"], "details": res["details"], "value": res["value"], "unit": UNIT, "ms_per_step": res["ms_per_step"],
                         "steps": k, "warmup": 3, "e2e": res["e2e"], "stage_ms": res["stage_ms"],
                         "gpu_launches": res["gpu_launches"]}
                if not args.no_cpu:
                    v, kind, sec, ev = cpu_reference_update(cgrid, ctruth, cscans, cn, cpu_updates)
                    entry["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                                             "sample": f"{cpu_updates} updateFilter calls at full size ({cn} particles, "
                                                       f"{sec:.1f} s of CPU)"}
                extra.append(entry)
            # config 4 again with the robot in the map's interior: the seeded pose lies 2.5 m from the map's corner, which
            # clips the sensor kernel's window; here the whole 8 m reach is inside the map (8-bit class tile), and the
            # rays that end in open space are culled (they score 0 for every particle)
            INTERIOR_POSE[0] = True
            try:
                res, _ = measure(hx, args, "config4", 5, 3, want_roofline=False)
            finally:
                INTERIOR_POSE[0] = args.pose == "interior"
            extra.append({"config": res["config"], "details": res["details"], "value": res["value"], "unit": UNIT,
                          "ms_per_step": res["ms_per_step"], "steps": 5, "warmup": 3, "e2e": res["e2e"],
                          "stage_ms": res["stage_ms"], "gpu_launches": res["gpu_launches"]})
        elif world == 8:
            res, _ = measure(hx, args, "config5", 3, 3, want_roofline=False)
            if rank == 0:
                extra.append({"config": res["config"], "details": res["details"], "value": res["value"], "unit": UNIT,
                              "ms_per_step": res["ms_per_step"], "steps": 3, "warmup": 3, "e2e": res["e2e"],
                              "stage_ms": res["stage_ms"], "estimate": res["estimate"], "digest": res["digest"]})
        if rank == 0:
            line["configs"] = extra
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        hx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="config4", choices=sorted(synth.CONFIGS))
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--sensor-path", type=int, default=0,
                    help="0 = auto (score-table pass where it fits, else two-pass), 1 = exact only, 2 = two-pass only")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short entries for the other BASELINE configs")
    ap.add_argument("--uniform", action="store_true", help="uniform cloud (global localisation) on any config")
    ap.add_argument("--particles", type=int, default=0, help="override the config's particle count")
    ap.add_argument("--pose", default="seeded", choices=["seeded", "interior"],
                    help="interior: the robot in the middle of the map (the seeded pose of configs 3-5 is near a corner)")
    args = ap.parse_args()
    INTERIOR_POSE[0] = args.pose == "interior"
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
