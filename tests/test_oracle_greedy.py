"""Pin the restated NearestSensorGreedy driver (oracle/uavsim_oracle.py::nearest_sensor_greedy) against the
reference's recorded config-1 rollouts: the goldens were produced with the reference's own
`NearestSensorGreedy` (greedy_agents.py:73-100) driving the reference env; here the restated agent drives the
oracle env from the seed alone and must pick the same action at every step while consuming the same number of
shadowing draws from the global stream."""
import numpy as np
import pytest

from golden_util import Fixture, O


@pytest.mark.parametrize("seed", [0, 1, 5, 12, 19])
def test_restated_greedy_reproduces_reference_rollout(seed):
    fx = Fixture("c1_greedy_500_n20")
    ep = fx.ep(seed)
    params = fx.params()
    env = O.OracleEnv(fx.grid, O.seeded_layout(fx.grid, fx.n, seed), fx.start, params,
                      O.StreamDraws(params["shadowing_std_db"]))
    env.reset(O.seeded_fill(fx.n, seed))
    T = len(ep["actions"])
    for t in range(T):
        a, probes = O.nearest_sensor_greedy(env, lambda i: env.draws.probe())
        assert a == int(ep["actions"][t]), (seed, t, a, int(ep["actions"][t]))
        assert int(np.sum(~np.isnan(probes))) == int(ep["burn"][t]), (seed, t)
        _, r, _, trunc = env.step(a)
        assert abs(r - ep["reward"][t]) <= 1e-9 * max(1.0, abs(ep["reward"][t])), (seed, t)
        assert trunc == bool(ep["truncated"][t])
    assert trunc
