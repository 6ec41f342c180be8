"""BASELINE.json configs[0] entirely on the device: the NearestSensorGreedy driver kernel chooses the actions
and the step kernel executes them, for the reference's seeded 20-episode rollout (500x500, N=20).  Both the
agent's probe draws and the environment's draws are replayed (re-derived from the fixture seeds by the pinned
python oracle); actions and rewards must equal what the reference recorded."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from __graft_entry__ import import_package
    return import_package()


def _oracle_rollout(fx, seed, agent="nearest_sensor_greedy"):
    """Drive the pinned python oracle with its restated greedy agent; collect per-step probe + env draws."""
    from golden_util import O
    pick = getattr(O, agent)
    params = fx.params()
    env = O.OracleEnv(fx.grid, O.seeded_layout(fx.grid, fx.n, seed), fx.start, params,
                      O.StreamDraws(params["shadowing_std_db"]))
    fill = O.seeded_fill(fx.n, seed)
    env.reset(fill)
    out = dict(spos=env.spos.copy(), fill=np.array(fill), reset_draws=env.draws.table.copy(), probes=[], draws=[],
               actions=[], rewards=[])
    trunc = False
    while not trunc:
        a, probes = pick(env, lambda i: env.draws.probe())
        _, r, _, trunc = env.step(a)
        out["probes"].append(probes)
        out["draws"].append(env.draws.table.copy())
        out["actions"].append(a)
        out["rewards"].append(r)
    return out


def test_config1_greedy_rollouts_on_device(pkg):
    import torch
    from golden_util import Fixture
    fx = Fixture("c1_greedy_500_n20")
    seeds = fx.seeds                                 # all 20 episodes of BASELINE configs[0]
    R = [_oracle_rollout(fx, s) for s in seeds]
    gold = [fx.ep(s) for s in seeds]
    for r, g in zip(R, gold):                       # the oracle rollout IS the reference's rollout
        assert r["actions"] == [int(a) for a in g["actions"]]
    E, n = len(seeds), fx.n
    env = pkg.UAVVecEnv(E, grid_size=fx.grid, num_sensors=n, sensor_positions=np.stack([r["spos"] for r in R]),
                        uav_start_position=fx.start, precision="f64", rng="replay", auto_reset=False)
    env.set_replay(reset_draws=np.stack([np.nan_to_num(r["reset_draws"]) for r in R]),
                   reset_fill=np.stack([r["fill"] for r in R]))
    env.reset_tensors()
    lens = [len(r["actions"]) for r in R]
    total_r = np.zeros(E)
    for t in range(max(lens)):
        live = np.array([t < L for L in lens])
        probes = np.stack([r["probes"][t] if t < L else np.full(n, np.nan) for r, L in zip(R, lens)])
        acts = env.nearest_greedy_actions(probes)
        a = acts.cpu().numpy()
        want = np.array([r["actions"][t] if t < L else a[k] for k, (r, L) in enumerate(zip(R, lens))])
        assert (a[live] == want[live]).all(), (t, a, want)
        draws = np.stack([np.nan_to_num(r["draws"][t]) if t < L else np.zeros((6, n)) for r, L in zip(R, lens)])
        env.set_replay(step_draws=draws)
        env.step_tensors(acts)
        rew = env.reward.cpu().numpy()
        done = env.done.cpu().numpy().astype(bool)
        for k in np.flatnonzero(live):
            ref = float(gold[k]["reward"][t])
            assert abs(rew[k] - ref) <= 1e-9 * max(1.0, abs(ref)), (seeds[k], t)
            assert done[k] == bool(gold[k]["truncated"][t])
            total_r[k] += rew[k]
    # episode fingerprints the survey recorded from the reference: seed 0 / 1 / 2
    fp = {0: (1403, -91344.83), 1: (1381, 185618.774), 2: (1382, 349296.14)}
    for k, s in enumerate(seeds):
        if s in fp:
            assert lens[k] == fp[s][0] and abs(total_r[k] - fp[s][1]) < 0.01, (s, lens[k], total_r[k])
    env.close()


@pytest.mark.parametrize("name,nseeds", [("c1_maxtp_500_n20", 4), ("maxtp_200_n30", 3)])
def test_max_throughput_greedy_rollouts_on_device(pkg, name, nseeds):
    """MaxThroughputGreedyV2 (greedy_agents.py:105-200) on the device: actions chosen by the driver kernel and
    executed by the step kernel must equal the rollouts recorded from the reference's own agent."""
    from golden_util import Fixture
    fx = Fixture(name)
    seeds = fx.seeds[:nseeds]
    R = [_oracle_rollout(fx, s, "max_throughput_greedy_v2") for s in seeds]
    gold = [fx.ep(s) for s in seeds]
    for r, g in zip(R, gold):
        assert r["actions"] == [int(a) for a in g["actions"]]
    E, n = len(seeds), fx.n
    kw = {k: v for k, v in fx.meta["env_kwargs"].items()}
    env = pkg.UAVVecEnv(E, grid_size=fx.grid, num_sensors=n, sensor_positions=np.stack([r["spos"] for r in R]),
                        uav_start_position=fx.start, precision="f64", rng="replay", auto_reset=False, **kw)
    env.set_replay(reset_draws=np.stack([np.nan_to_num(r["reset_draws"]) for r in R]),
                   reset_fill=np.stack([r["fill"] for r in R]))
    env.reset_tensors()
    lens = [len(r["actions"]) for r in R]
    for t in range(max(lens)):
        live = np.array([t < L for L in lens])
        probes = np.stack([r["probes"][t] if t < L else np.full(n, np.nan) for r, L in zip(R, lens)])
        acts = env.max_throughput_greedy_actions(probes)
        a = acts.cpu().numpy()
        want = np.array([r["actions"][t] if t < L else a[k] for k, (r, L) in enumerate(zip(R, lens))])
        assert (a[live] == want[live]).all(), (t, a, want)
        draws = np.stack([np.nan_to_num(r["draws"][t]) if t < L else np.zeros((6, n)) for r, L in zip(R, lens)])
        env.set_replay(step_draws=draws)
        env.step_tensors(acts)
        rew = env.reward.cpu().numpy()
        done = env.done.cpu().numpy().astype(bool)
        for k in np.flatnonzero(live):
            ref = float(gold[k]["reward"][t])
            assert abs(rew[k] - ref) <= 1e-9 * max(1.0, abs(ref)), (seeds[k], t)
            assert done[k] == bool(gold[k]["truncated"][t])
    env.close()


def test_philox_max_throughput_driver_beats_random_walk(pkg):
    """Throughput build, Philox: the SF-aware driver + step entirely on the device for 4096 envs."""
    env = pkg.UAVVecEnv(4096, grid_size=(200, 200), num_sensors=20, seed=9, max_battery=60.0)
    env.reset_tensors()
    for t in range(400):
        env.step_tensors(env.max_throughput_greedy_actions())
    env.poll_errors()
    rec = env.records()
    assert rec["total_data_collected"].mean() > 1500.0
    assert np.array([bin(int(v)).count("1") for v in rec["visited_lo"]]).mean() > 1.0
    env.close()


def test_philox_greedy_driver_collects_data(pkg):
    """Throughput build: greedy + step entirely on device for 4096 envs; the driver must reach sensors."""
    import torch
    env = pkg.UAVVecEnv(4096, grid_size=(200, 200), num_sensors=20, seed=8, max_battery=60.0)
    env.reset_tensors()
    for t in range(400):
        env.step_tensors(env.nearest_greedy_actions())
    env.poll_errors()
    rec = env.records()
    assert rec["total_data_collected"].mean() > 1500.0
    visited = np.array([bin(int(v)).count("1") for v in rec["visited_lo"]])
    assert visited.mean() > 1.0
    assert float((env.state()["buffer"] >= 0).all())
    env.close()
