#!/usr/bin/env python
"""Developer tool: step rate of small batches through UAVVecEnv.step_tensors (launch-bound regime) and through the
numpy SB3 path UAVVecEnv.step (BASELINE configs[1] shape: 4096 envs, 500x500, N=20)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from __graft_entry__ import import_package  # noqa: E402

pkg = import_package()
for E in (256, 4096, 65536):
    env = pkg.UAVVecEnv(E, grid_size=(500, 500), num_sensors=20, seed=1)
    env.reset_tensors()
    g = torch.Generator(device=env.device)
    g.manual_seed(1)
    acts = [torch.randint(0, 5, (E,), dtype=torch.int32, device=env.device, generator=g) for _ in range(8)]
    for t in range(50):
        env.step_tensors(acts[t % 8])
    torch.cuda.synchronize()
    K = 3000
    t0 = time.perf_counter()
    for t in range(K):
        env.step_tensors(acts[t % 8])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # the same step captured once in a CUDA graph and replayed (one frame per env: no per-step host state)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        env.step_tensors(acts[0])
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        env.step_tensors(acts[0])
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    for t in range(K):
        graph.replay()
    torch.cuda.synchronize()
    dtg = time.perf_counter() - t2
    a_np = [a.cpu().numpy() for a in acts]
    env.reset()
    K2 = 300
    t1 = time.perf_counter()
    for t in range(K2):
        env.step(a_np[t % 8])
    dt2 = time.perf_counter() - t1
    print("E=%6d  step_tensors: %.1f us/step, %.1f M env-steps/s | CUDA-graph replay: %.1f us/step, %.1f M env-steps/s | "
          "numpy VecEnv.step: %.1f us/step, %.2f M env-steps/s"
          % (E, dt / K * 1e6, E * K / dt / 1e6, dtg / K * 1e6, E * K / dtg / 1e6, dt2 / K2 * 1e6, E * K2 / dt2 / 1e6))
    env.close()
