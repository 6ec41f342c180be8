"""Helpers shared by the parity tests: load a golden fixture (produced from the
unmodified reference by oracle/gen_golden.py) and re-run it with the oracle."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
if str(ROOT / "oracle") not in sys.path:
    sys.path.insert(0, str(ROOT / "oracle"))

import uavsim_oracle as O  # noqa: E402

# reference constructor keyword (uav_env.py:266-287) -> oracle parameter
_KW = {"sensor_duty_cycle": "duty_cycle", "max_buffer_size": "max_buffer_size",
       "include_sensor_positions": "include_sensor_positions",
       "collection_duration": "collection_duration", "max_battery": "max_battery",
       "max_steps": "max_steps", "rssi_threshold": "rssi_threshold"}


class Fixture:
    def __init__(self, name, flavour="norm"):
        self.path = GOLDEN / ("%s.%s.npz" % (name, flavour))
        self.z = np.load(self.path, allow_pickle=False)
        self.meta = json.loads(str(self.z["meta"]))
        self.seeds = self.meta["seeds"]
        self.grid = tuple(self.meta["grid"])
        self.n = self.meta["n"]
        self.start = tuple(self.meta["start"])

    def ep(self, seed):
        pre = "s%d/" % seed
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def params(self, **over):
        kw = {_KW[k]: v for k, v in self.meta["env_kwargs"].items()}
        kw["normalise"] = self.meta["flavour"] == "norm"
        kw.update(over)
        return O.default_params(**kw)


def replay_stream(fx: Fixture, seed, **over):
    """Re-run one fixture episode with the python oracle in STREAM mode."""
    ep = fx.ep(seed)
    params = fx.params(**over)
    pos = O.seeded_layout(fx.grid, fx.n, seed)
    env = O.OracleEnv(fx.grid, pos, fx.start, params, O.StreamDraws(params["shadowing_std_db"]))
    first = O.traced_reset(env, O.seeded_fill(fx.n, seed))
    steps = []
    for t, a in enumerate(ep["actions"]):
        env.draws.burn_normals(ep["burn"][t])
        steps.append(O.traced_step(env, int(a)))
    S = O.stack_steps(steps)
    S["obs_sum"] = S["obs"].astype(np.float64).sum(axis=1)
    return ep, first, S, env


DENSE_KEYS = ("reward", "truncated", "pos", "battery", "sf", "visited", "capture_triggers",
              "boundary_hits", "edge_steps", "last_step_bytes", "c_collision_count",
              "c_urgency_reduced", "c_sensor_urgency", "c_was_new_sensor", "c_was_empty",
              "c_all_sensors_collected", "c_bytes_collected", "c_data_loss",
              "i_high_urgency_sensors", "i_sensors_collected", "obs_sum")
SPARSE_KEYS = ("buf", "gen", "tx", "lost", "ema", "obs", "i_ratios", "i_max_urgency",
               "i_avg_urgency", "total_reward", "total_data_collected", "i_coverage_percentage",
               "i_battery_percent")
DISCRETE_KEYS = ("truncated", "pos", "sf", "visited", "capture_triggers", "boundary_hits",
                 "edge_steps", "c_collision_count", "c_was_new_sensor", "c_was_empty",
                 "c_all_sensors_collected", "i_high_urgency_sensors", "i_sensors_collected")
