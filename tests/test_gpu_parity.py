"""GPU parity tests proper: the CUDA path (through the C ABI) against the C oracle on the same
seeded inputs, and against the reference's own golden trajectories.  Discrete state bit-exact,
floats within 1e-4 relative (BASELINE.json) -- the parity build is asserted much tighter."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLOAT_TOL = 1e-4          # the bar BASELINE.json states
TIGHT_TOL = 1e-9          # what the float64 parity build is expected to achieve


@pytest.fixture(scope="module")
def pkg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from __graft_entry__ import import_package
    return import_package()


def _check(res, tol=TIGHT_TOL):
    assert res["discrete_mismatches"] == 0, res
    assert res["max_rel_err"] <= tol, res


def test_config2_4096_envs_random_actions(pkg):
    """BASELINE.json configs[1]: 4096 envs, 500x500, N=20, uniform random actions."""
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=4096, N=20, grid=(500, 500), steps=96, seed=2, p_collect=0.2, compare_every=8)
    _check(res)
    assert res["bytes"] > 0


def test_n50_grid1000_two_slots_per_lane(pkg):
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=1024, N=50, grid=(1000, 1000), steps=64, seed=5, p_collect=0.5,
                          start="random", compare_every=4)
    _check(res)
    assert res["bytes"] > 0


def test_dense_grid_collisions_and_capture(pkg):
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=512, N=40, grid=(100, 100), steps=160, seed=7, p_collect=0.7, start=(50.0, 50.0),
                          compare_every=4)
    _check(res)
    assert res["collisions"] > 100 and res["captures"] > 10, res


def test_duty_cycle_100_every_sensor_transmits(pkg):
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=256, N=30, grid=(200, 200), steps=120, seed=11, p_collect=0.6, start=(100.0, 100.0),
                          params_over=dict(duty_cycle=100.0), ref_kwargs=dict(sensor_duty_cycle=100.0))
    _check(res)
    assert res["collisions"] > 1000


def test_ragged_sensor_counts_and_grids_padded_to_50(pkg):
    """BASELINE.json configs[2] geometry: grid 300-1000, N=10-50 padded to 50 observation slots."""
    from parity_util import run_parity_case
    rng = np.random.default_rng(3)
    E = 1024
    n = rng.integers(10, 51, size=E)
    g = rng.integers(300, 1001, size=E)
    res = run_parity_case(pkg, E=E, N=50, P=50, per_env_n=n, per_env_grid=np.stack([g, g], 1), steps=48, seed=13,
                          p_collect=0.4, start="random", obs_slots=50, compare_every=4)
    _check(res)


def test_auto_reset_terminal_observation_and_refill(pkg):
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=512, N=20, grid=(200, 200), steps=120, seed=17, p_collect=0.5, auto_reset=True,
                          start=(100.0, 100.0), params_over=dict(max_battery=4.0), ref_kwargs=dict(max_battery=4.0))
    _check(res)
    assert res["dones"] > 1000, res


def test_max_steps_truncation_and_terminal_penalties(pkg):
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=256, N=12, grid=(150, 150), steps=90, seed=19, p_collect=0.3, auto_reset=True,
                          params_over=dict(max_steps=40), ref_kwargs=dict(max_steps=40))
    _check(res)
    assert res["dones"] >= 512


def test_small_buffer_data_loss_path(pkg):
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=256, N=10, grid=(300, 300), steps=200, seed=23, p_collect=0.2,
                          params_over=dict(max_buffer_size=400.0), ref_kwargs=dict(max_buffer_size=400.0),
                          compare_every=5)
    _check(res)
    assert res["per_key"]["data_loss"] == 0.0 or res["per_key"]["data_loss"] < TIGHT_TOL


def test_five_feature_observation(pkg):
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=128, N=12, grid=(300, 300), steps=60, seed=29, start="random",
                          params_over=dict(include_sensor_positions=True),
                          ref_kwargs=dict(include_sensor_positions=True))
    _check(res)


def test_collect_duration_two_seconds(pkg):
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=128, N=16, grid=(200, 200), steps=100, seed=31, p_collect=0.5, start=(90.0, 110.0),
                          params_over=dict(collection_duration=2.0), ref_kwargs=dict(collection_duration=2.0))
    _check(res)


def test_domain_rand_shaping_and_in_kernel_reset(pkg):
    """DomainRandEnv semantics (dqn.py:301-444): fresh empty sensors, inherited SF, injected layout and
    far start, proximity + Jain shaping, padded 153-float rows."""
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=512, N=20, P=20, grid=(300, 300), steps=100, seed=37, p_collect=0.4,
                          auto_reset=True, start="random", domain_rand=True, obs_slots=50,
                          params_over=dict(max_battery=6.0), ref_kwargs=dict(max_battery=6.0))
    _check(res)
    assert res["dones"] > 500


def test_single_sensor_and_single_env_edge_cases(pkg):
    from parity_util import run_parity_case
    _check(run_parity_case(pkg, E=1, N=1, grid=(50, 50), steps=80, seed=41, p_collect=0.5, start=(25.0, 25.0)))
    _check(run_parity_case(pkg, E=3, N=64, grid=(120, 120), steps=80, seed=43, p_collect=0.6, start=(60.0, 60.0)))
    _check(run_parity_case(pkg, E=9, N=33, grid=(10, 10), steps=60, seed=47, p_collect=0.3))


def test_float32_throughput_build_on_replayed_draws(pkg):
    """The float32 build on the same draws: floats within the 1e-4 bar wherever no discrete decision
    flipped; flips are rare and counted (float32 state cannot promise bit-exact thresholds)."""
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=2048, N=20, grid=(500, 500), steps=24, seed=53, p_collect=0.3, precision="f32")
    frac = res["discrete_mismatches"] / (res["steps"] * 20)
    assert frac < 1e-3, res
    # per-step float quantities of the float32 build against the float64 restatement (tolerance 1e-4 relative for the
    # state and the observation; the reward cancels large terms, 1e-3)
    pk = res["per_key"]
    assert pk["obs"] <= 1e-4 and pk["ema"] <= 1e-4 and pk["buf"] <= 1e-4 and pk["tx"] <= 1e-4 and pk["bytes"] <= 1e-4, pk
    assert pk["reward"] <= 1e-3 and pk["battery"] == 0.0 and pk["lost"] <= 1e-4, pk


def test_config2_full_episodes_to_truncation(pkg):
    """configs[1] over WHOLE episodes: 1024 envs, uniform random actions until every env has truncated
    (battery-limited, ~1790 steps) and restarted in-kernel; state compared every 64 steps."""
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=1024, N=20, grid=(500, 500), steps=1900, seed=61, p_collect=0.2, auto_reset=True,
                          compare_every=64)
    _check(res)
    assert res["dones"] >= 1024, res


@pytest.mark.parametrize("N", [1, 2, 32, 33, 64])
def test_sensor_count_edges(pkg, N):
    """Sensor counts at the edges of the lane mapping: a single sensor (no fairness variance, reward_function.py:49),
    exactly one / just over one slot per lane, and the 64-sensor maximum."""
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=96, N=N, grid=(120, 120), steps=40, seed=100 + N, p_collect=0.6, start=(60.0, 60.0),
                          auto_reset=True, params_over=dict(max_battery=3.0), ref_kwargs=dict(max_battery=3.0))
    _check(res)
    assert res["dones"] > 0, res
