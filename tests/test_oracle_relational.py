"""Pin the C oracle's relational mode against sequences produced by the reference's
`RelationalUAVEnv` (src/experiments/relational_policy/env_wrapper.py:110-313): Dict observation flattened as
uav(3) + 50 x [rel_x, rel_y, buffer, urgency, visited], potential shaping, cluster dwell bonus, 1e-5 scale."""
import json

import numpy as np
import pytest

from golden_util import GOLDEN, O  # noqa: F401  (puts oracle/ on sys.path)
import c_oracle

NAMES = ["rel_300_n20", "rel_150_n35"]


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_relational_matches_reference(name):
    z = np.load(GOLDEN / ("%s.norm.npz" % name))
    meta = json.loads(str(z["meta"]))
    n = meta["n"]
    params = O.default_params(include_sensor_positions=True, **meta["env_kwargs"])
    co = c_oracle.COracleBatch(params, E=1, P=n, obs_slots=50, relational=True)
    co.set_scenario([meta["grid"]], [n], z["positions"].astype(np.float32)[None], [meta["start"]])
    bonus_steps = 0
    for ep in range(meta["episodes"]):
        g = lambda k: z["e%d/%s" % (ep, k)]           # noqa: E731
        obs0 = co.reset(g("reset_fill")[None], np.nan_to_num(g("reset_draws"))[None]).copy()
        assert _same(obs0[0], g("reset_obs")), (name, ep)
        e = co.envs[0]
        diag = np.sqrt(meta["grid"][0] ** 2 + meta["grid"][1] ** 2)
        assert -e["prev_dist_nearest"] / diag == g("reset_potential")
        for t, a in enumerate(g("actions")):
            obs, rew, done = co.step([int(a)], np.nan_to_num(g("draws")[t])[None], dwell=np.nan_to_num(g("dwell_draws")[t])[None])
            e = co.envs[0]
            tag = (name, ep, t)
            assert _same(obs[0], g("obs")[t]), tag
            for k in ("sf", "buf", "tx", "lost", "ema", "gen"):
                assert _same(e[k][:n], g(k)[t]), (tag, k)
            assert e["dwell_steps"] == g("dwell_steps")[t], tag
            assert -e["prev_dist_nearest"] / diag == g("potential")[t], tag
            assert bool(done[0]) == bool(g("truncated")[t]), tag
            ref = g("reward")[t]
            assert abs(rew[0] - ref) <= 1e-12 * max(1.0, abs(ref)) + 1e-15, (tag, rew[0], ref)
            bonus_steps += int(g("dwell_steps")[t] > 0)
        # episode metrics of env_wrapper.py:232-268 from the final state
        gen, tx = e["gen"][:n], e["tx"][:n]
        assert abs(tx.sum() / gen.sum() - g("m_ndr")[-1]) < 1e-12
        cr = tx / np.maximum(gen, 1e-9)
        assert abs(cr.sum() ** 2 / (n * (cr ** 2).sum()) - g("m_jains")[-1]) < 1e-12
    assert bonus_steps > 10
