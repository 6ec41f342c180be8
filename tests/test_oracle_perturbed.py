"""Pin the restated PerturbedDQNEnv (oracle/uavsim_oracle.py::PerturbedOracle) against sequences recorded from the
reference's own class (sim_to_real_sweep.py:107-139 over compare_agents.py::AnalysisUAVEnv, exec'd slices; chained
episodes, every inner env step stored): every draw is re-taken from the seeded global streams in the reference's
call order; padded observations, rewards, positions and sensor state must be reproduced bit for bit."""
import json
import random

import numpy as np
import pytest

from golden_util import GOLDEN, O

NAMES = ["pert_gps10_200_n20", "pert_shadow8_200_n20", "pert_wind1_200_n20", "pert_all_200_n20"]


@pytest.mark.parametrize("name", NAMES)
def test_restated_perturbed_env_reproduces_reference_sequences(name):
    z = np.load(GOLDEN / ("%s.norm.npz" % name))
    meta = json.loads(str(z["meta"]))
    seed, n = meta["seed"], meta["n"]
    random.seed(seed)
    np.random.seed(seed)
    np.random.uniform([0, 0], meta["grid"], (n, 2))                   # the constructor's layout draw
    params = O.default_params(max_steps=meta["env_kwargs"]["max_steps"], shadowing_std_db=meta["shadow_sigma"])
    base = O.OracleEnv(tuple(meta["grid"]), z["positions"], tuple(meta["start"]), params,
                       O.StreamDraws(params["shadowing_std_db"]))
    env = O.PerturbedOracle(base, gps_sigma=meta["gps_sigma"], wind_drift=meta["wind_drift"])
    pcg = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))      # gymnasium reset(seed=...)
    extras = 0
    for ep in range(meta["episodes"]):
        fill = pcg.uniform(0.2, 0.6, size=n)
        assert np.array_equal(fill, z["e%d/reset_fill" % ep])
        obs = env.reset(fill)
        assert np.array_equal(obs, z["e%d/reset_obs" % ep]), (name, ep, "reset obs")
        ex = z["e%d/reset_extra" % ep]
        assert np.array_equal(np.array(env.extra), ex[~np.isnan(ex)])
        acts, is_extra = z["e%d/action" % ep], z["e%d/is_extra" % ep]
        t = 0
        while t < len(acts):
            obs, r, _, trunc = env.step(int(acts[t]))
            for (o_i, r_i, tr_i, ex_i, pre_i) in env.inner:
                tag = (name, ep, t)
                assert np.array_equal(o_i, z["e%d/obs" % ep][t]), tag
                assert r_i == z["e%d/reward" % ep][t] and tr_i == bool(z["e%d/truncated" % ep][t]), tag
                g_ex, g_pre = z["e%d/extra" % ep][t], z["e%d/pre" % ep][t]
                assert np.array_equal(np.array(ex_i), g_ex[~np.isnan(g_ex)]), tag
                assert np.array_equal(np.array(pre_i), g_pre[~np.isnan(g_pre)]), tag
                extras += int(is_extra[t])
                t += 1
            t0 = t - 1
            assert not is_extra[t0]
            assert np.array_equal(base.pos, z["e%d/pos" % ep][t0]) and np.array_equal(base.sf, z["e%d/sf" % ep][t0])
            assert np.array_equal(base.buf, z["e%d/buf" % ep][t0]) and base.total_reward == z["e%d/total_reward" % ep][t0]
            assert base.boundary_hits == z["e%d/boundary_hits" % ep][t0] and base.edge_steps == z["e%d/edge_steps" % ep][t0]
    if meta["wind_drift"] > 0:
        assert extras > 10          # the wind gate fired and its extra steps were reproduced
