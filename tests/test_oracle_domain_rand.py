"""Pin the C oracle's DomainRandEnv mode against sequences produced by the reference's own
`DomainRandEnv` (src/agents/dqn/dqn.py:177-451; the file does not import as shipped, so
oracle/ref_harness.py exec's the unmodified text of the class).  Scenario (grid, layout, far start,
inherited SF) and draws are injected from the fixture; everything the wrapper computes -- fresh empty
sensors, padding to 153, proximity + Jain shaping, first-full-coverage step -- must come out identical."""
import json

import numpy as np
import pytest

from golden_util import GOLDEN, O  # noqa: F401  (puts oracle/ on sys.path)
import c_oracle

NAMES = ["dr_stage4_n20", "dr_stage2_n10", "dr_stage4_n40"]


def load(name):
    z = np.load(GOLDEN / ("%s.norm.npz" % name))
    return z, json.loads(str(z["meta"]))


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_domain_rand_matches_reference(name):
    z, meta = load(name)
    n = meta["n"]
    params = O.default_params(max_steps=meta["max_steps"])
    co = c_oracle.COracleBatch(params, E=1, P=n, obs_slots=50, domain_rand=True)
    for ep in range(meta["episodes"]):
        g = lambda k: z["e%d/%s" % (ep, k)]           # noqa: E731
        co.set_scenario([g("reset_grid")], [n], g("reset_positions").astype(np.float32)[None], [g("reset_start")])
        obs0 = co.reset(None, np.nan_to_num(g("reset_draws"))[None], init_sf=[int(g("reset_init_sf"))]).copy()
        assert _same(obs0[0], g("reset_obs")), (name, ep, "reset obs")
        e = co.envs[0]
        assert _same(e["sf"][:n], g("reset_sf")) and _same(e["ema"][:n], g("reset_ema"))
        assert (e["buf"][:n] == 0).all() and e["prev_dist_nearest"] == 0.0
        acts = g("actions")
        for t, a in enumerate(acts):
            obs, rew, done = co.step([int(a)], np.nan_to_num(g("draws")[t])[None])
            e = co.envs[0]
            tag = (name, ep, t)
            assert _same(obs[0], g("obs")[t]), tag
            for k in ("sf", "buf", "tx", "lost", "ema", "gen"):
                assert _same(e[k][:n], g(k)[t]), (tag, k)
            assert (e["x"], e["y"]) == tuple(g("pos")[t]) and e["battery"] == g("battery")[t], tag
            assert _same(e["visited"][:n].astype(bool), g("visited")[t]), tag
            assert bool(done[0]) == bool(g("truncated")[t]), tag
            assert e["capture_triggers"] == g("capture_triggers")[t] and e["boundary_hits"] == g("boundary_hits")[t], tag
            assert e["first_full_coverage_step"] == g("first_cov")[t], tag
            # shaped reward: the proximity term is a float32 norm (numpy routes it through BLAS), the Jain term
            # a builtin-sum; both agree to rounding
            assert abs(rew[0] - g("reward")[t]) <= 1e-9 * max(1.0, abs(g("reward")[t])), (tag, rew[0], g("reward")[t])
            assert abs(e["prev_dist_nearest"] - g("prev_dist")[t]) <= 1e-5, tag
            assert abs(e["total_reward"] - g("total_reward")[t]) <= 1e-9 * max(1.0, abs(g("total_reward")[t])), tag
        # DomainRandEnv.last_episode_stats (dqn.py:316-331) from the oracle's final state
        st = dict(zip(meta["stat_keys"], g("stats")))
        gen, tx, lost = e["gen"][:n], e["tx"][:n], e["lost"][:n]
        assert abs(gen.sum() - st["total_generated"]) < 1e-6 and abs(tx.sum() - st["total_collected"]) < 1e-6
        assert abs(lost.sum() - st["total_lost"]) < 1e-6 and abs(e["battery"] - st["battery_remaining"]) < 1e-12
        assert abs(e["visited"][:n].sum() / n * 100 - st["ndr"]) < 1e-9
        rates = tx / gen * 100
        jain = rates.sum() ** 2 / (n * (rates ** 2).sum()) if (rates ** 2).sum() > 0 else 1.0
        assert abs(jain - st["jains_index"]) < 1e-9 and abs(rates.std() - st["fairness_std"]) < 1e-9
