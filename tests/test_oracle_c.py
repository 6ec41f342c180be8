"""Pin the C oracle: golden fixtures (via the draws the python oracle re-derives
from the fixture seeds, which tests/test_oracle_golden.py pins to the reference)
must be reproduced bit-exactly -- floats included -- by oracle/uavsim_oracle.c."""
import numpy as np
import pytest

from golden_util import Fixture, replay_stream  # noqa: E402 (puts oracle/ on sys.path)
import c_oracle  # noqa: E402

CASES = [
    ("c1_greedy_500_n20", [0, 3]),
    ("c2_random_500_n20", [1, 2, 9]),
    ("c2_random_500_n20_full", [1]),
    ("c5_mixed_1000_n50", [0, 3]),
    ("dense_100_n40", [0, 1, 2, 3, 4, 5]),
    ("duty100_200_n30", [0, 1, 2]),
    ("smallbuf_300_n10", [0, 1]),
    ("fps5_300_n12", [0, 1]),
    ("cd2_200_n16", [0, 1]),
    ("lowbatt_200_n20", [0, 1, 2]),
    ("maxsteps_150_n8", [0, 1]),
]


def run_c_oracle_on_episode(fx, seed):
    ep, first, S, env = replay_stream(fx, seed)
    n = fx.n
    co = c_oracle.COracleBatch(fx.params(), E=1, P=n)
    co.set_scenario([fx.grid], [n], env.spos[None], [fx.start])
    obs0 = co.reset(first["fill"][None], first["draws"][None]).copy()
    out = dict(obs=[], reward=[], done=[], sf=[], buf=[], tx=[], lost=[], ema=[], gen=[], pos=[],
               battery=[], visited=[], detail=[], mask=[], capture=[], bh=[], edge=[], total_reward=[])
    for t, a in enumerate(ep["actions"]):
        o, r, d = co.step([int(a)], S["draws"][t][None])
        e = co.envs[0]
        out["obs"].append(o[0].copy()); out["reward"].append(r[0]); out["done"].append(d[0])
        for k in ("sf", "buf", "tx", "lost", "ema", "gen", "visited"):
            out[k].append(e[k][:n].copy())
        out["pos"].append((e["x"], e["y"])); out["battery"].append(e["battery"])
        out["detail"].append(co.detail[0].copy()); out["mask"].append(co.mask[0])
        out["capture"].append(e["capture_triggers"]); out["bh"].append(e["boundary_hits"])
        out["edge"].append(e["edge_steps"]); out["total_reward"].append(e["total_reward"])
    return ep, first, S, obs0, {k: np.array(v) for k, v in out.items()}


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("name,seeds", CASES)
def test_c_oracle_bit_exact(name, seeds):
    fx = Fixture(name)
    for seed in seeds:
        ep, first, S, obs0, C = run_c_oracle_on_episode(fx, seed)
        assert _same(obs0[0], first["obs"])
        assert _same(C["obs"], S["obs"]), (name, seed, "obs")
        for k in ("sf", "buf", "tx", "lost", "ema", "gen", "pos", "battery"):
            assert _same(C[k], S[k]), (name, seed, k)
        assert _same(C["visited"].astype(bool), S["visited"])
        assert _same(C["done"].astype(bool), S["truncated"])
        assert _same(C["capture"], S["capture_triggers"])
        assert _same(C["bh"], S["boundary_hits"])
        assert _same(C["edge"], S["edge_steps"])
        coll = np.nan_to_num(S["c_collision_count"], nan=0.0)
        assert _same(C["detail"][:, 0], coll)
        collect = ep["actions"] == 4
        assert _same(C["detail"][collect, 1], S["c_bytes_collected"][collect])
        assert _same(C["detail"][collect, 2], S["c_urgency_reduced"][collect])
        assert _same(C["detail"][collect, 3], S["c_sensor_urgency"][collect])
        # the step's data loss is summed directly instead of differencing two cumulative sums
        np.testing.assert_allclose(C["reward"], S["reward"], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(C["total_reward"], S["total_reward"], rtol=1e-12, atol=1e-7)


def test_c_oracle_action_mask_matches_python_oracle():
    fx = Fixture("c2_random_500_n20")
    ep, first, S, env = replay_stream(fx, 4)
    co = c_oracle.COracleBatch(fx.params(), E=1, P=fx.n)
    co.set_scenario([fx.grid], [fx.n], env.spos[None], [fx.start])
    co.reset(first["fill"][None], first["draws"][None])
    import uavsim_oracle as O
    py = O.OracleEnv(fx.grid, [tuple(p) for p in ep["positions"]], fx.start, fx.params(), O.SlotDraws())
    py.draws.load(first["draws"]); py.reset(first["fill"])
    for t, a in enumerate(ep["actions"][:200]):
        co.step([int(a)], S["draws"][t][None])
        py.draws.load(S["draws"][t]); py.step(int(a))
        bits = [(int(co.mask[0]) >> k) & 1 for k in range(5)]
        assert bits == [int(b) for b in py.action_mask()]


def test_c_oracle_rejects_invalid_action():
    fx = Fixture("maxsteps_150_n8")
    ep, first, S, env = replay_stream(fx, 0)
    co = c_oracle.COracleBatch(fx.params(), E=1, P=fx.n)
    co.set_scenario([fx.grid], [fx.n], env.spos[None], [fx.start])
    co.reset(first["fill"][None], first["draws"][None])
    with pytest.raises(ValueError):
        co.step([5], S["draws"][0][None])
    with pytest.raises(ValueError):
        co.step([-1], S["draws"][0][None])
