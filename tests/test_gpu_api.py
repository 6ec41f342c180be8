"""Host-interface behaviour of UAVVecEnv on a GPU: SB3 VecEnv contract, error behaviour of the
reference (ValueError on invalid actions), Philox determinism and shard invariance."""
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


def test_invalid_action_raises_value_error_like_reference(pkg):
    import torch
    env = pkg.UAVVecEnv(8, num_sensors=5, grid_size=(50, 50))
    env.reset()
    with pytest.raises(ValueError):
        env.step(np.array([0, 1, 2, 3, 4, 5, 0, 0]))          # uav_env.py:468
    with pytest.raises(ValueError):
        env.step(np.array([0, -1, 2, 3, 4, 1, 0, 0]))
    before = env.records()["step"].copy()
    env.step_tensors(torch.tensor([0, 1, 2, 3, 4, 7, 0, 0], dtype=torch.int32, device=env.device))
    with pytest.raises(ValueError):
        env.poll_errors()
    after = env.records()["step"]
    assert after[5] == before[5] and (after[[0, 1, 2, 3, 4, 6, 7]] == before[[0, 1, 2, 3, 4, 6, 7]] + 1).all()
    env.poll_errors()      # cleared
    env.close()


def test_vecenv_contract_shapes_dtypes_and_autoreset(pkg):
    env = pkg.UAVVecEnv(16, grid_size=(100, 100), num_sensors=10, max_battery=2.0, seed=3)
    obs = env.reset()
    assert obs.shape == (16, 33) and obs.dtype == np.float32
    assert env.observation_space.shape == (33,) and env.action_space.n == 5
    n_done = 0
    for t in range(40):
        obs, rew, dones, infos = env.step(np.random.default_rng(t).integers(0, 5, 16))
        assert obs.shape == (16, 33) and rew.shape == (16,) and dones.dtype == bool and len(infos) == 16
        for e in np.flatnonzero(dones):
            n_done += 1
            assert infos[e]["TimeLimit.truncated"] is True
            assert infos[e]["terminal_observation"].shape == (33,)
            assert set(infos[e]["episode"]) == {"r", "l"}
            # the returned obs is already the next episode's first observation: full battery
            assert obs[e, 2] == 1.0
            st = env.get_attr("last_episode_stats", [int(e)])[0]
            assert st["num_sensors"] == 10 and st["grid_size"] == (100, 100)
    assert n_done >= 16
    info = env.get_info([0])[0]
    assert len(info["sensor_collection_ratios"]) == 10 and "coverage_percentage" in info
    env.close()


def test_philox_is_deterministic_and_shard_invariant(pkg):
    import torch
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 5, size=(30, 64)).astype(np.int32)

    def run(E, offset):
        env = pkg.UAVVecEnv(E, grid_size=(300, 300), num_sensors=24, seed=99, env_id_offset=offset, max_battery=3.0)
        out = [env.reset_tensors().clone()]
        for t in range(30):
            o, r, d = env.step_tensors(torch.as_tensor(acts[t, offset:offset + E], device=env.device))
            out.append(torch.cat([o, r[:, None], d[:, None].float()], 1).clone())
        env.close()
        return out
    whole = run(64, 0)
    again = run(64, 0)
    lo, hi = run(32, 0), run(32, 32)
    for t in range(31):
        assert torch.equal(whole[t], again[t])
        assert torch.equal(whole[t][:32], lo[t]) and torch.equal(whole[t][32:], hi[t])


def test_domain_rand_philox_reset_invariants(pkg):
    """DomainRandEnv.reset in Philox mode: grid from the stage list, sensors inside the grid, far start
    >= 50 units from every sensor (or the furthest found), empty buffers, padded obs."""
    import torch
    env = pkg.UAVVecEnv(2048, num_sensors=20, domain_randomisation=True, curriculum_stage=4, seed=5)
    obs = env.reset_tensors()
    rec = env.records()
    assert obs.shape == (2048, 153)
    grids = {(w, w) for w in (100, 200, 300, 400, 500)}
    got = set(zip(rec["grid_w"].tolist(), rec["grid_h"].tolist()))
    assert got <= grids and len(got) == 5
    pos = env.state()["sensor_pos"].cpu().numpy()[:, :20]
    assert (pos >= 0).all() and (pos[..., 0] <= rec["grid_w"][:, None]).all()
    start = np.stack([rec["x"], rec["y"]], 1)
    d = np.sqrt(((pos - start[:, None, :]) ** 2).sum(-1)).min(1)
    big = rec["grid_w"] >= 300
    assert (d[big] >= 50.0 - 1e-3).mean() > 0.99
    assert (start[:, 0] >= 0.05 * rec["grid_w"] - 1e-3).all() and (start[:, 0] <= 0.95 * rec["grid_w"] + 1e-3).all()
    assert float(env.state()["buffer"][:, :20].abs().max()) == 0.0
    assert float(obs[:, 3 + 60:].abs().max()) == 0.0
    env.set_curriculum_stage(0)
    act = torch.full((2048,), 4, dtype=torch.int32, device=env.device)
    for _ in range(3):
        env.step_tensors(act)
    env.close()


def test_step_is_capturable_in_a_cuda_graph(pkg):
    """With one frame per environment a step launch carries no per-step host state (tiles are handed out by a device
    counter that the last block resets), so policy + step can be captured once and replayed: the launch-bound
    small-batch regime without Python in the loop.  Replays must equal eager stepping bit for bit."""
    import torch
    E = 2048
    kw = dict(grid_size=(300, 300), num_sensors=20, seed=5, max_battery=20.0, uav_start_position=(150.0, 150.0))
    eager = pkg.UAVVecEnv(E, **kw)
    graphed = pkg.UAVVecEnv(E, **kw)
    eager.reset_tensors()
    graphed.reset_tensors()
    dev = eager.device

    def policy(obs, out):          # a deterministic toy policy on the device: keeps everything in the graph
        out.copy_((obs[:, :3].sum(dim=1) * 7919.0).to(torch.int64).remainder(5).to(torch.int32))

    a_e = torch.zeros(E, dtype=torch.int32, device=dev)
    a_g = torch.zeros(E, dtype=torch.int32, device=dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):              # warm-up on the capture stream, as torch asks
        for _ in range(3):
            policy(graphed.obs, a_g)
            graphed.step_tensors(a_g)
    torch.cuda.current_stream(dev).wait_stream(side)
    for _ in range(3):
        policy(eager.obs, a_e)
        eager.step_tensors(a_e)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        policy(graphed.obs, a_g)
        graphed.step_tensors(a_g)
    policy(eager.obs, a_e)                      # the capture itself does not execute: eager stays one step behind ...
    for _ in range(200):
        graph.replay()
    for _ in range(199):
        eager.step_tensors(a_e)
        policy(eager.obs, a_e)
    eager.step_tensors(a_e)
    torch.cuda.synchronize()
    assert torch.equal(eager.obs, graphed.obs) and torch.equal(eager.reward, graphed.reward)
    assert torch.equal(eager.stat_sums[-1], graphed.stat_sums[-1]) and int(eager.stat_sums[-1].item()) > 0
    eager.close()
    graphed.close()


@pytest.mark.parametrize("n_sensors,n_stack", [(20, 1), (50, 1), (20, 4)])
def test_pipelined_host_step_equals_one_launch(pkg, n_sensors, n_stack):
    """uavsim_step_host steps a big batch as 4 env ranges (copy-out of range k overlapping the step of range k+1).
    Environments are independent and their draws are keyed by the global env id, so the result must be bit-identical
    to stepping the whole batch in one launch -- observations, rewards, dones, restarts, episode statistics."""
    import torch
    E = 70016 + 40                      # not a multiple of the 64-env range granularity
    kw = dict(grid_size=(300, 300), num_sensors=n_sensors, seed=17, max_battery=1.2, n_stack=n_stack)
    a, b = pkg.UAVVecEnv(E, **kw), pkg.UAVVecEnv(E, **kw)
    a.reset_tensors()
    b.reset_tensors()
    host = a.make_host_buffers()
    g = torch.Generator()
    g.manual_seed(5)
    dones = 0
    for t in range(12):
        act = torch.randint(0, 5, (E,), dtype=torch.int32, generator=g)
        host["actions"].copy_(act)
        obs, rew, done = a.step_host(host)                      # 4 ranges, second stream
        b.step_tensors(act.to(b.device))                        # one launch
        assert torch.equal(obs, b.obs.cpu()), t
        assert torch.equal(rew, b.reward.cpu()) and torch.equal(done, b.done.cpu()), t
        dones += int(done.sum())
    assert dones > 1000
    assert torch.equal(a.terminal_obs, b.terminal_obs) and torch.equal(a.episode_stats, b.episode_stats)
    assert torch.allclose(a.stat_sums, b.stat_sums, rtol=1e-12)
    assert a.launch_count - b.launch_count == 12 * 3            # 4 launches per step instead of 1
    a.close()
    b.close()
