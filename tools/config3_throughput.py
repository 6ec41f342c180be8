#!/usr/bin/env python
"""Developer tool: throughput of BASELINE.json configs[2] (65,536 envs, DomainRandEnv semantics: per-episode grid /
layout / far start, N = 10..50 padded to 50 observation slots, 4 stacked frames for the attention extractor) on one
GPU.  Not a bench line (bench.py measures configs[4]); prints env-steps/s for profiles/README.md."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from __graft_entry__ import import_package  # noqa: E402

pkg = import_package()
E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for stack in (1, 4):
    rng = np.random.default_rng(0)
    ns = rng.integers(10, 51, size=E).astype(np.int32)
    env = pkg.UAVVecEnv(E, num_sensors=ns, max_sensors=50, domain_randomisation=True, curriculum_stage=4, seed=9,
                        n_stack=stack, action_mask=False)
    env.reset_tensors()
    g = torch.Generator(device=env.device)
    g.manual_seed(1)
    acts = [torch.randint(0, 5, (E,), dtype=torch.int32, device=env.device, generator=g) for _ in range(8)]
    for t in range(20):
        env.step_tensors(acts[t % 8])
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 2000
    ev0.record()
    for t in range(K):
        env.step_tensors(acts[t % 8])
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / K
    print("configs[2] E=%d n_stack=%d: %.4f ms/step, %.1f M env-steps/s, episodes finished %d"
          % (E, stack, ms, E / ms / 1e3, int(env.stat_sums[-1].item())))
    env.close()
