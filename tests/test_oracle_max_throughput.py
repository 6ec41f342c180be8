"""Pin the restated MaxThroughputGreedyV2 driver (oracle/uavsim_oracle.py::max_throughput_greedy_v2) against
rollouts recorded from the reference's own agent (greedy_agents.py:105-200) driving the reference env: from the
seed alone the restated agent must choose the same action at every step, consume the same number of shadowing
draws, and the oracle env must return the recorded rewards."""
import numpy as np
import pytest

from golden_util import Fixture, O


@pytest.mark.parametrize("name,seed", [("c1_maxtp_500_n20", 0), ("c1_maxtp_500_n20", 3), ("c1_maxtp_500_n20", 7),
                                       ("maxtp_200_n30", 0), ("maxtp_200_n30", 2)])
def test_restated_max_throughput_reproduces_reference_rollout(name, seed):
    fx = Fixture(name)
    ep = fx.ep(seed)
    params = fx.params()
    env = O.OracleEnv(fx.grid, O.seeded_layout(fx.grid, fx.n, seed), fx.start, params,
                      O.StreamDraws(params["shadowing_std_db"]))
    env.reset(O.seeded_fill(fx.n, seed))
    T = len(ep["actions"])
    for t in range(T):
        a, probes = O.max_throughput_greedy_v2(env, lambda i: env.draws.probe())
        assert a == int(ep["actions"][t]), (seed, t, a, int(ep["actions"][t]))
        assert int(np.sum(~np.isnan(probes))) == int(ep["burn"][t]), (seed, t)
        _, r, _, trunc = env.step(a)
        assert abs(r - ep["reward"][t]) <= 1e-9 * max(1.0, abs(ep["reward"][t])), (seed, t)
        assert trunc == bool(ep["truncated"][t])
    assert trunc
