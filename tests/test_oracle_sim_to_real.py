"""Pin the restated sim-to-real wrappers (oracle/uavsim_oracle.py::SimToRealOracle) against sequences recorded from
the reference's own GPSNoisyEnv / WindDriftEnv / LatencyEnv (sim_to_real_analysis.py:170-245, exec'd slice): chained
episodes, every draw re-taken from the seeded global streams in the reference's call order, observations (padded to
50 slots), rewards, positions and sensor state compared bit for bit."""
import json
import random

import numpy as np
import pytest

from golden_util import GOLDEN, O


def load(name):
    z = np.load(GOLDEN / ("%s.norm.npz" % name))
    return z, json.loads(str(z["meta"]))


@pytest.mark.parametrize("name", ["gps5_300_n12", "wind1_200_n12", "wind2_120_n20", "lat2_200_n10"])
def test_restated_wrappers_reproduce_reference_sequences(name):
    z, meta = load(name)
    seed = {"gps5_300_n12": 31, "wind1_200_n12": 32, "wind2_120_n20": 33, "lat2_200_n10": 34}[name]
    random.seed(seed)
    np.random.seed(seed)
    np.random.uniform([0, 0], meta["grid"], (meta["n"], 2))           # the constructor's layout draw
    kw = meta["env_kwargs"]
    params = O.default_params(max_steps=kw["max_steps"], max_battery=kw["max_battery"])
    base = O.OracleEnv(tuple(meta["grid"]), z["positions"], tuple(meta["start"]), params,
                       O.StreamDraws(params["shadowing_std_db"]))
    env = O.SimToRealOracle(base, **meta["extra"])
    pcg = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))      # gymnasium reset(seed=...)
    for ep in range(meta["episodes"]):
        fill = pcg.uniform(0.2, 0.6, size=meta["n"])
        assert np.array_equal(fill, z["e%d/reset_fill" % ep])
        obs = env.reset(fill)
        assert np.array_equal(obs, z["e%d/reset_obs" % ep]), (name, ep, "reset obs")
        ex = z["e%d/reset_extra" % ep]
        assert np.array_equal(np.array(env.extra), ex[~np.isnan(ex)])
        acts = z["e%d/actions" % ep]
        for t, a in enumerate(acts):
            obs, r, _, trunc = env.step(int(a))
            tag = (name, ep, t)
            assert np.array_equal(obs, z["e%d/obs" % ep][t]), tag
            assert r == z["e%d/reward" % ep][t], tag
            assert trunc == bool(z["e%d/truncated" % ep][t]), tag
            assert np.array_equal(base.pos, z["e%d/pos" % ep][t]), tag
            assert np.array_equal(base.sf, z["e%d/sf" % ep][t]), tag
            assert np.array_equal(base.buf, z["e%d/buf" % ep][t]), tag
            ex = z["e%d/extra" % ep][t]
            assert np.array_equal(np.array(env.extra), ex[~np.isnan(ex)]), tag
