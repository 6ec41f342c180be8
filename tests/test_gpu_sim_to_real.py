"""Sim-to-real evaluation wrappers on the device (GPS noise on the observed position, systematic wind drift,
observation latency; sim_to_real_analysis.py:170-245): the CUDA parity build replays the sequences recorded from the reference's own
GPSNoisyEnv / WindDriftEnv / LatencyEnv (tests/golden/gps*, wind*, lat*: chained episodes, one environment per recorded episode) and
must reproduce observations bit for bit, positions and spreading factors exactly, rewards to 1e-9; the throughput
build's Philox draws are checked statistically."""
import json

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


@pytest.mark.parametrize("name", ["gps5_300_n12", "wind1_200_n12", "wind2_120_n20", "lat2_200_n10"])
def test_cuda_reproduces_reference_sim_to_real_sequences(pkg, name):
    import torch
    from golden_util import GOLDEN
    z = np.load(GOLDEN / ("%s.norm.npz" % name))
    meta = json.loads(str(z["meta"]))
    E, n = meta["episodes"], meta["n"]
    kw = dict(meta["env_kwargs"])
    start = kw.pop("uav_start_position")
    env = pkg.UAVVecEnv(E, grid_size=tuple(meta["grid"]), num_sensors=n, obs_slots=50,
                        sensor_positions=np.tile(z["positions"][None], (E, 1, 1)), uav_start_position=tuple(start),
                        precision="f64", rng="replay", auto_reset=False, **meta["extra"], **kw)
    ex = np.nan_to_num(np.stack([z["e%d/reset_extra" % e] for e in range(E)]))
    if meta["cls"] == "WindDriftEnv":           # its only reset draw is the angle: slot 2 of the per-env table
        ex = np.stack([np.zeros(E), np.zeros(E), ex[:, 0], np.zeros(E)], axis=1)
    env.set_replay(reset_draws=np.nan_to_num(np.stack([z["e%d/reset_draws" % e] for e in range(E)])),
                   reset_fill=np.stack([z["e%d/reset_fill" % e] for e in range(E)]), reset_env_draws=ex)
    obs0 = env.reset_tensors().cpu().numpy()
    for e in range(E):
        assert np.array_equal(obs0[e], z["e%d/reset_obs" % e]), (name, e, "reset obs")
    lens = [len(z["e%d/actions" % e]) for e in range(E)]
    worst = 0.0
    for t in range(max(lens)):
        live = [e for e in range(E) if t < lens[e]]
        act = np.array([int(z["e%d/actions" % e][t]) if t < lens[e] else 4 for e in range(E)], dtype=np.int32)
        dr = np.stack([np.nan_to_num(z["e%d/draws" % e][t]) if t < lens[e] else np.zeros((6, n)) for e in range(E)])
        xs = np.stack([np.nan_to_num(z["e%d/extra" % e][t]) if t < lens[e] else np.zeros(4) for e in range(E)])
        env.set_replay(step_draws=dr, step_env_draws=xs)
        env.step_tensors(torch.as_tensor(act, device=env.device))
        obs = env.obs.cpu().numpy()
        rew = env.reward.cpu().numpy()
        done = env.done.cpu().numpy()
        rec = env.records()
        sf = env.state()["spreading_factor"].cpu().numpy()
        for e in live:
            tag = (name, e, t)
            assert np.array_equal(obs[e], z["e%d/obs" % e][t]), tag
            assert bool(done[e]) == bool(z["e%d/truncated" % e][t]), tag
            assert rec["x"][e] == z["e%d/pos" % e][t][0] and rec["y"][e] == z["e%d/pos" % e][t][1], tag
            assert np.array_equal(sf[e, :n], z["e%d/sf" % e][t]), tag
            ref = float(z["e%d/reward" % e][t])
            err = abs(rew[e] - ref) / max(abs(ref), 1.0)
            worst = max(worst, err)
            assert err <= 1e-9, (tag, rew[e], ref)
    assert worst <= 1e-9
    env.close()


def test_philox_gps_noise_and_wind_drift_statistics(pkg):
    import torch
    E, n, W = 16384, 10, 400
    env = pkg.UAVVecEnv(E, grid_size=(W, W), num_sensors=n, seed=3, gps_sigma=5.0, uav_start_position=(200.0, 200.0),
                        auto_reset=False)
    env.reset_tensors()
    up = torch.zeros(E, dtype=torch.int32, device=env.device)
    env.step_tensors(up)
    o = env.obs.cpu().numpy().astype(np.float64)
    nx, ny = o[:, 0] - 200.0 / W, o[:, 1] - 201.0 / W          # noise is added to the NORMALISED entries, as executed
    for v in (nx, ny):
        assert abs(v.mean()) < 0.15 and abs(v.std() - 5.0) < 0.1, (v.mean(), v.std())
    assert abs(np.corrcoef(nx, ny)[0, 1]) < 0.03
    env.close()

    env = pkg.UAVVecEnv(E, grid_size=(W, W), num_sensors=n, seed=4, wind_drift=1.5, uav_start_position=(200.0, 200.0),
                        auto_reset=False)
    env.reset_tensors()
    rec = env.records()
    ang = rec["wind_angle"]
    assert ang.min() >= 0.0 and ang.max() < 2 * np.pi
    dx, dy = 1.5 * np.cos(ang), 1.5 * np.sin(ang)
    assert abs(np.mean(np.cos(ang))) < 0.03 and abs(np.mean(np.sin(ang))) < 0.03        # uniform on the circle
    collect = torch.full((E,), 4, dtype=torch.int32, device=env.device)
    for _ in range(10):
        env.step_tensors(collect)
    rec = env.records()
    assert np.allclose(rec["x"], 200.0 + 10 * dx, atol=1e-3) and np.allclose(rec["y"], 200.0 + 10 * dy, atol=1e-3)
    o = env.obs.cpu().numpy()
    assert np.array_equal(o[:, 0], rec["x"]) and np.array_equal(o[:, 1], rec["y"])        # raw position, as executed
    for _ in range(300):
        env.step_tensors(collect)
    rec = env.records()
    assert rec["x"].min() >= 0 and rec["x"].max() <= W - 1 and rec["y"].min() >= 0 and rec["y"].max() <= W - 1
    env.close()


def test_latency_composes_with_frame_stacking(pkg):
    """LatencyEnv under VecFrameStack (sim_to_real_analysis.py:246-250 wraps it that way): the stacked window must be
    the k frames that END `latency` steps ago, primed with copies of the reset observation and zero frames before."""
    import torch
    E, n, k, L = 257, 8, 4, 3
    kw = dict(grid_size=(200, 200), num_sensors=n, seed=12, uav_start_position=(100.0, 100.0), auto_reset=False)
    plain = pkg.UAVVecEnv(E, **kw)
    late = pkg.UAVVecEnv(E, n_stack=k, latency_steps=L, **kw)
    od = plain.dims.obs_dim
    frames = [plain.reset_tensors().clone()]
    got = late.reset_tensors().clone()
    g = torch.Generator(device=plain.device)
    g.manual_seed(1)
    for t in range(12):
        # frame i of the delayed sequence: reset obs for i <= L, then the obs of step i - L; zero before the episode
        seq = [frames[max(0, i - L)] if i >= 0 else torch.zeros_like(frames[0]) for i in range(t - k + 1, t + 1)]
        assert torch.equal(got, torch.cat(seq, dim=1)), t
        a = torch.randint(0, 5, (E,), dtype=torch.int32, device=plain.device, generator=g)
        plain.step_tensors(a)
        late.step_tensors(a)
        frames.append(plain.obs.clone())
        got = late.obs.clone()
        assert got.shape == (E, k * od)
    plain.close()
    late.close()


def test_host_buffer_step_returns_the_delayed_stacked_window(pkg):
    """uavsim_step_host copies the stacked window out of the mirrored ring (2D copy): with frame stacking and
    latency it must equal the zero-copy device view, and the numpy VecEnv.step path must return the same arrays."""
    import torch
    E = 300
    env = pkg.UAVVecEnv(E, grid_size=(200, 200), num_sensors=12, seed=21, n_stack=4, latency_steps=2,
                        uav_start_position=(100.0, 100.0), max_battery=5.0)
    env.reset_tensors()
    host = env.make_host_buffers()
    g = torch.Generator()
    g.manual_seed(3)
    for t in range(40):
        host["actions"].copy_(torch.randint(0, 5, (E,), dtype=torch.int32, generator=g))
        obs, rew, done = env.step_host(host)
        assert torch.equal(obs, env.obs.cpu()), t
        assert torch.equal(rew, env.reward.cpu()) and torch.equal(done, env.done.cpu()), t
    a = np.random.default_rng(0).integers(0, 5, size=E)
    o, r, d, infos = env.step(a)
    assert o.shape == (E, 4 * env.dims.obs_dim) and np.array_equal(o, env.obs.cpu().numpy())
    assert np.array_equal(r, env.reward.cpu().numpy()) and len(infos) == E
    env.close()
