"""On-device frame stacking: the mirrored frame ring must equal SB3's
VecFrameStack (dqn.py:1278) applied to the oracle's observations -- shift, newest frame last, stack zeroed
on done, stacked terminal_observation."""
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


@pytest.mark.parametrize("k,N,fps5", [(4, 20, False), (10, 50, False), (3, 12, True)])
def test_frame_ring_equals_vecframestack_over_oracle(pkg, k, N, fps5):
    import torch
    from parity_util import Scenario, make_tables      # puts oracle/ on sys.path
    import c_oracle
    import uavsim_oracle as O
    E, T = 128, 60
    sc = Scenario(E, N, N, (200, 200), seed=5, start=(100.0, 100.0))
    rng = sc.rng
    params = O.default_params(max_battery=3.0, include_sensor_positions=fps5)
    co = c_oracle.COracleBatch(params, E=E, P=N)
    co.set_scenario(sc.grid, sc.n, sc.xy, sc.start)
    env = pkg.UAVVecEnv(E, grid_size=sc.grid, num_sensors=sc.n, sensor_positions=sc.xy, uav_start_position=sc.start,
                        precision="f64", rng="replay", auto_reset=True, n_stack=k, max_battery=3.0,
                        include_sensor_positions=fps5)
    od = co.obs_dim
    assert env.obs.shape == (E, k * od) and env.observation_space.shape == (k * od,)
    fill, d0 = rng.uniform(0.2, 0.6, size=(E, N)), make_tables(rng, E, N)
    co.reset(fill, d0)
    env.set_replay(reset_draws=d0, reset_fill=fill)
    env.reset_tensors()
    stack = np.zeros((E, k * od), dtype=np.float32)
    stack[:, -od:] = co.obs
    assert np.array_equal(env.obs.cpu().numpy(), stack)
    host = env.make_host_buffers()
    dones_seen = 0
    for t in range(T):
        a = np.where(rng.random(E) < 0.4, 4, rng.integers(0, 4, E)).astype(np.int32)
        dt, dr, ft = make_tables(rng, E, N), make_tables(rng, E, N), rng.uniform(0.2, 0.6, size=(E, N))
        env.set_replay(step_draws=dt, reset_draws=dr, reset_fill=ft)
        obs, rew, done = co.step(a, dt)
        term = obs.copy()
        done = done.copy().astype(bool)
        if done.any():
            co.reset(ft, dr, env_ids=np.flatnonzero(done))
        stack = np.roll(stack, -od, axis=1)                   # VecFrameStack.step_wait
        terminal = {}
        for i in np.flatnonzero(done):
            terminal[i] = np.concatenate([stack[i, :-od], term[i]])
            stack[i] = 0
        stack[:, -od:] = co.obs
        if t % 2 == 0:
            env.step_tensors(torch.as_tensor(a, device=env.device))
            got = env.obs.cpu().numpy()
        else:                                                  # the host-buffer path copies the same window
            host["actions"].copy_(torch.as_tensor(a))
            env.step_host(host)
            got = host["obs"].numpy()
        assert np.array_equal(got, stack), (t, np.argwhere(got != stack)[:4])
        tob = env.terminal_obs.cpu().numpy()
        for i, v in terminal.items():
            assert np.array_equal(tob[i], v), (t, i)
        dones_seen += int(done.sum())
    assert dones_seen > E
    env.close()


def test_numpy_vecenv_path_with_stacking(pkg):
    env = pkg.UAVVecEnv(32, grid_size=(100, 100), num_sensors=10, n_stack=4, max_battery=2.0, seed=1)
    obs = env.reset()
    assert obs.shape == (32, 4 * 33) and (obs[:, :99] == 0).all() and (obs[:, 99 + 2] == 1.0).all()
    seen = False
    for t in range(30):
        obs, rew, dones, infos = env.step(np.full(32, 4))
        for e in np.flatnonzero(dones):
            seen = True
            assert infos[e]["terminal_observation"].shape == (4 * 33,)
            assert (obs[e, :99] == 0).all() and obs[e, 99 + 2] == 1.0       # zeroed stack + fresh battery
    assert seen
    env.close()
