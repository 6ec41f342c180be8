#!/usr/bin/env python
"""Developer tool: long-horizon soak of the benchmarked kernels (1M envs, tens of thousands of steps, many in-kernel
restarts per env) followed by invariant checks on the whole state."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from __graft_entry__ import import_package  # noqa: E402

pkg = import_package()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for n, grid in ((50, (1000, 1000)), (20, (500, 500))):
    E = 1 << 20
    env = pkg.UAVVecEnv(E, grid_size=grid, num_sensors=n, seed=99, action_mask=False)
    env.reset_tensors()
    g = torch.Generator(device=env.device)
    g.manual_seed(2)
    acts = [torch.randint(0, 5, (E,), dtype=torch.int32, device=env.device, generator=g) for _ in range(16)]
    t0 = time.perf_counter()
    for t in range(steps):
        env.step_tensors(acts[t % 16])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    env.poll_errors()
    st = env.state()
    rec = env.records()
    buf, lost, tx = st["buffer"][:, :n], st["lost"][:, :n], st["transmitted"][:, :n]
    sf = st["spreading_factor"][:, :n]
    ok = dict(obs_finite=bool(torch.isfinite(env.obs).all()), reward_finite=bool(torch.isfinite(env.reward).all()),
              buf_range=bool(((buf >= 0) & (buf <= 1000.0)).all()), lost_nonneg=bool((lost >= 0).all()), tx_nonneg=bool((tx >= 0).all()),
              sf_range=bool(((sf >= 7) & (sf <= 12)).all()), obs_01=bool((env.obs.min() >= 0) and (env.obs.max() <= 1.0)),
              battery=bool((rec["battery"] > 0).all() and (rec["battery"] <= 274.0).all()),
              steps=bool((rec["step"] >= 0).all() and (rec["step"] <= 2100).all()),
              pos=bool((rec["x"] >= 0).all() and (rec["x"] < grid[0]).all() and (rec["y"] >= 0).all() and (rec["y"] < grid[1]).all()))
    sums = env.stat_sums.cpu().numpy()
    print("N=%d: %d steps x %d envs in %.1f s (%.0f M env-steps/s incl. launch loop), episodes finished %.0f (%.1f per env), "
          "mean length %.0f, mean NDR %.1f %%, checks %s"
          % (n, steps, E, dt, E * steps / dt / 1e6, sums[-1], sums[-1] / E, sums[1] / max(sums[-1], 1), sums[6] / max(sums[-1], 1),
             "ALL OK" if all(ok.values()) else ok))
    assert all(ok.values()), ok
    env.close()
