#!/usr/bin/env python
"""Benchmark of the batched UAV-LoRa environment step (BASELINE.json metric: env-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[4], the scalability configuration the metric's
"1/2/4/8 B200" is quoted on -- 1000x1000 grid, N=50 LoRa sensors, 1,048,576 environments PER GPU
(weak scaling: environments are independent, each rank owns a contiguous id range, no
communication inside the step; episode statistics are combined by one NCCL all-reduce after the
timed loop), throughput build (float32 state, Philox draws), uniform random actions resident in
HBM, in-kernel auto-reset.  One "step" = one launch of the fused step kernel over all environments
of the rank.

`value`  : whole-job env-steps/s, state and actions resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : the same metric through the public host-buffer call (UAVVecEnv.step_host ->
           uavsim_step_host): actions copied host->device and obs/reward/done copied device->host
           inside the timed region, every step.
`roofline`: algorithmic bytes per env-step x envs per launch / mean launch duration,
           against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline`: the reference's algorithm on the box's host cores (python port of the reference's
           per-sensor numpy-scalar step, all cores, bounded sample) -- a reported baseline.

`--impl reference` times that CPU port alone (the reference itself is pure Python under
/root/reference, which does not exist on the GPU box; oracle/uavsim_oracle.py is its pinned
restatement) with all host cores on the same config/metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = dict(name="configs[4]: scalability sweep 1000x1000, N=50, 1M envs per GPU, random actions, Philox",
                grid=(1000, 1000), num_sensors=50, envs_per_gpu=1 << 20)
METRIC = "env-steps/sec"


def bytes_per_env_step(n_sensors: int, precision: str = "f32", obs_floats: int | None = None) -> int:
    """ALGORITHMIC bytes one env-step must move (DESIGN.md section 'byte model'): persistent bytes
    once read + once written, static per-episode data once read, outputs once written, inputs once
    read; nothing for padding lanes, caches or replay tables."""
    rb = 4 if precision == "f32" else 8
    per_sensor = 4 * rb * 2          # {buffer, transmitted, lost, avg_rssi} read + write
    per_sensor += 8                  # position float2, read
    per_sensor += rb if precision == "f32" else 2 * rb   # generated: f32 reads the reset value; f64 accumulates r+w
    if precision != "f32":
        per_sensor += 2              # spreading factor u8 read + write (f32: packed into avg_rssi)
    per_env = 128 * 2                # env record read + write
    per_env += 4                     # action int32
    per_env += 4 * (obs_floats if obs_floats is not None else 3 + 3 * n_sensors)   # observation row
    per_env += rb + 1                # reward, done
    return per_sensor * n_sensors + per_env


# ----------------------------------------------------------------------------------------------
# CPU arm: python port of the reference step on all host cores
# ----------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, n_envs, grid, n_sensors, warm, steps = args
    sys.path.insert(0, str(ROOT / "oracle"))
    import random

    import numpy as np
    import uavsim_oracle as O
    random.seed(seed)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    envs = []
    for k in range(n_envs):
        pos = rng.uniform(0, 1, size=(n_sensors, 2)) * np.asarray(grid)
        env = O.OracleEnv(grid, [tuple(p) for p in pos], (0.0, 0.0), O.default_params(normalise=False))
        env.reset(rng.uniform(0.2, 0.6, size=n_sensors))
        envs.append(env)
    acts = rng.integers(0, 5, size=(warm + steps, n_envs))
    t0 = None
    for t in range(warm + steps):
        if t == warm:
            t0 = time.perf_counter()
        for k, env in enumerate(envs):
            _, _, _, trunc = env.step(int(acts[t, k]))
            if trunc:
                env.reset(rng.uniform(0.2, 0.6, size=n_sensors))
    return time.perf_counter() - t0


def cpu_python_port(grid, n_sensors, envs_per_core, warm, steps, cores=None):
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t = pool.map(_cpu_worker, [(1000 + i, envs_per_core, grid, n_sensors, warm, steps) for i in range(cores)])
    wall = max(t)
    return dict(value=cores * envs_per_core * steps / wall, unit=METRIC, cores=cores, kind="port",
                sample="python port of the reference step (oracle/uavsim_oracle.py, numpy-scalar per-sensor loops "
                       "like the reference), %d procs x %d envs x %d steps, grid %dx%d, N=%d, random actions"
                       % (cores, envs_per_core, steps, grid[0], grid[1], n_sensors),
                ms_per_step=wall / steps * 1e3)


def cpu_c_port(grid, n_sensors, n_envs, steps, cores=None):
    """The plain-C restatement (oracle/uavsim_oracle.c) on all cores: what a tuned CPU could do."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import c_oracle
    import numpy as np
    import uavsim_oracle as O
    cores = cores or len(os.sched_getaffinity(0))
    rng = np.random.default_rng(0)
    P = n_sensors
    co = c_oracle.COracleBatch(O.default_params(), E=n_envs, P=P, threads=cores)
    g = np.tile(np.asarray(grid, dtype=np.int32), (n_envs, 1))
    co.set_scenario(g, np.full(n_envs, n_sensors, dtype=np.int32),
                    (rng.uniform(0, 1, size=(n_envs, P, 2)) * np.asarray(grid)).astype(np.float32),
                    np.zeros((n_envs, 2), dtype=np.float32))
    draws = rng.normal(0, 4.0, size=(n_envs, 6, P))
    draws[:, 2] = rng.random(size=(n_envs, P))
    co.reset(rng.uniform(0.2, 0.6, size=(n_envs, P)), draws)
    acts = rng.integers(0, 5, size=(steps + 2, n_envs)).astype(np.int32)
    for t in range(2):
        co.step(acts[t], draws, want_detail=False)
    t0 = time.perf_counter()
    for t in range(steps):
        co.step(acts[2 + t], draws, want_detail=False)   # same table every step: draw generation is not timed
    wall = time.perf_counter() - t0
    return dict(value=n_envs * steps / wall, unit=METRIC, cores=cores, kind="port",
                sample="C restatement (oracle/uavsim_oracle.c, gcc -O2), %d threads, %d envs x %d steps, draws pre-generated"
                       % (cores, n_envs, steps))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    grid, n = WORKLOAD["grid"], WORKLOAD["num_sensors"]
    cores = len(os.sched_getaffinity(0))
    envs_per_core = 8
    res = cpu_python_port(grid, n, envs_per_core, args.warmup, args.steps, cores)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": METRIC, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "grid": list(grid), "num_sensors": n,
                       "envs_per_step": cores * envs_per_core,
                       "note": "bounded sample of the workload; the reference is pure Python and scales by processes"},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import import_package
    pkg = import_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    E = args.envs or WORKLOAD["envs_per_gpu"]
    n = args.sensors or WORKLOAD["num_sensors"]
    grid = WORKLOAD["grid"]
    env = pkg.UAVVecEnv(E, grid_size=grid, num_sensors=n, precision="f32", rng="philox", seed=20261018,
                        env_id_offset=rank * E, auto_reset=True, device=str(dev), action_mask=False)
    env.reset_tensors()
    n_act = 8
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    # Steady state instead of 1M synchronised fresh episodes: start every UAV at a random state of charge, so that
    # truncations, terminal penalties and in-kernel restarts happen inside the timed region at the rate of a long run
    # (a full battery lasts ~1790 random steps -- longer than the default timed region).
    env.set_battery(5.48 + (274.0 - 5.48) * torch.rand(E, device=dev, generator=g, dtype=torch.float64))
    actions = [torch.randint(0, 5, (E,), dtype=torch.int32, device=dev, generator=g) for _ in range(n_act)]
    stream = torch.cuda.current_stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident throughput ----
    for t in range(args.warmup):
        env.step_tensors(actions[t % n_act])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = env.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for t in range(args.steps):
        env.step_tensors(actions[t % n_act])
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = env.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    env.poll_errors()
    # the one collective of the path: episode statistics (off the step's critical path)
    sums = env.stat_sums.clone()
    if world > 1:
        dist.all_reduce(sums)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())

    # ---- end to end through the host-buffer call ----
    host = env.make_host_buffers()
    e2e_steps = min(args.steps, args.e2e_steps)
    ha = [a.cpu() for a in actions]
    for t in range(min(3, args.warmup)):
        host["actions"].copy_(ha[t % n_act])
        env.step_host(host)
    barrier()
    t0 = time.perf_counter()
    ev0.record(stream)
    for t in range(e2e_steps):
        host["actions"].copy_(ha[t % n_act])      # the step's inputs arrive in pinned host memory
        env.step_host(host)
    ev1.record(stream)
    barrier()
    e2e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)
    tms = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_ms = float(tms.item())
    obs_dim = env.dims.obs_dim
    h2d = E * 4
    d2h = E * (obs_dim * 4 + 4 + 1)

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        bpe = bytes_per_env_step(n, "f32")
        launch_ms = ms / max(launches, 1)
        achieved = bpe * E / (launch_ms * 1e-3) / 1e9
        traffic = None
        tpath = ROOT / "profiles" / "dram_traffic.json"
        if tpath.exists():
            try:
                tj = json.loads(tpath.read_text())
                if tj.get("envs") == E and tj.get("num_sensors") == n:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:  # noqa: BLE001
                pass
        cpu = None
        cpu_c = None
        if not args.no_cpu:
            cpu = cpu_python_port(grid, n, envs_per_core=4, warm=5, steps=args.cpu_steps)
            cpu_c = cpu_c_port(grid, n, n_envs=8192, steps=20)
        secondary = None
        if not args.sensors and not args.no_secondary:
            # the metric is quoted for N = 20-50: the same workload at N = 20 (pair kernel: two environments per warp pass)
            env.close()
            n2 = 20
            env2 = pkg.UAVVecEnv(E, grid_size=grid, num_sensors=n2, precision="f32", rng="philox", seed=20261018,
                                 env_id_offset=rank * E, auto_reset=True, device=str(dev), action_mask=False)
            env2.reset_tensors()
            for t in range(5):
                env2.step_tensors(actions[t % n_act])
            torch.cuda.synchronize(dev)
            k2 = min(args.steps, 300)
            ev0.record(stream)
            for t in range(k2):
                env2.step_tensors(actions[t % n_act])
            ev1.record(stream)
            torch.cuda.synchronize(dev)
            ms2 = ev0.elapsed_time(ev1) / k2
            bpe2 = bytes_per_env_step(n2, "f32")
            secondary = {"num_sensors": n2, "envs_per_gpu": E, "steps": k2, "ms_per_step": ms2,
                         "value_per_gpu": E / (ms2 * 1e-3), "bytes_per_env_step": bpe2,
                         "roofline_frac": bpe2 * E / (ms2 * 1e-3) / 1e9 / peak,
                         "kernel": "uavsim::step_kernel<float, REPLAY=false, SPL=1, FPS=3, STACK=false, LEAN=plain, PAIR=true>"}
            env2.close()
        total_steps = E * world * args.steps
        line = {
            "metric": METRIC, "value": total_steps / (ms_max * 1e-3), "unit": METRIC, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "grid": list(grid), "num_sensors": n, "envs_per_gpu": E,
                       "global_envs": E * world, "precision": "f32 state / f64 battery+returns", "rng": "philox4x32-10",
                       "actions": "uniform random, %d pre-generated device tensors cycled" % n_act,
                       "auto_reset": True, "initial_state": "random state of charge per env (episodes end inside the timed region)", "parallelism": "env-sharded x%d, no data-path collective" % world,
                       "l2": "state per launch %.2f GB >> 126 MB L2 (inputs larger than L2)" % (bpe * E / 1e9)},
            "clocks": clocks,
            "e2e": {"value": E * world * e2e_steps / (e2e_ms * 1e-3), "unit": METRIC, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "UAVVecEnv.step_host -> uavsim_step_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_env_step": bpe,
                         "kernel": "uavsim::step_kernel<float, REPLAY=false, SPL=%d, FPS=3, STACK=false, LEAN=plain>" % (2 if n > 32 else 1),
                         "launch_ms": launch_ms},
            "cpu_baseline": None if cpu is None else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "cpu_baseline_native": cpu_c,
            "secondary_n20_rank0": secondary,
            "episode_stats": {"episodes_finished": float(sums[-1].item()),
                              "mean_return": float(sums[0].item() / max(sums[-1].item(), 1.0))},
        }
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's 1,048,576)")
    ap.add_argument("--sensors", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--cpu-steps", type=int, default=300)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the N=20 secondary measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
