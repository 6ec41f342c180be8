"""PerturbedDQNEnv of the sim-to-real sweep on the device (sim_to_real_sweep.py:107-139: per-sensor shadowing sigma, GPS
noise applied to the position INSIDE the observation pass, probabilistic extra whole step for wind): the CUDA parity
build replays the sequences recorded from the reference's own class (tests/golden/pert_*: chained episodes, every
inner env step stored, one environment per recorded episode) -- padded observations bit for bit, positions / spreading
factors / counters exactly, rewards to 1e-9; the dropped wind repeats go through uavsim_step_ex(UAVSIM_STEP_EXTRA).
Philox mode: the wind gate's rate, per-env shadowing / duty columns."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NAMES = ["pert_gps10_200_n20", "pert_shadow8_200_n20", "pert_wind1_200_n20", "pert_all_200_n20"]


@pytest.fixture(scope="module")
def pkg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from __graft_entry__ import import_package
    return import_package()


@pytest.mark.parametrize("name", NAMES)
def test_cuda_reproduces_reference_perturbed_dqn_env(pkg, name):
    import torch
    from golden_util import GOLDEN
    z = np.load(GOLDEN / ("%s.norm.npz" % name))
    meta = json.loads(str(z["meta"]))
    E, n = meta["episodes"], meta["n"]
    env = pkg.UAVVecEnv(E, grid_size=tuple(meta["grid"]), num_sensors=n, obs_slots=50,
                        sensor_positions=np.tile(z["positions"][None], (E, 1, 1)), uav_start_position=tuple(meta["start"]),
                        precision="f64", rng="replay", auto_reset=False, max_steps=meta["env_kwargs"]["max_steps"],
                        gps_sigma=meta["gps_sigma"], gps_in_observation=True, shadowing_std_db=meta["shadow_sigma"])
    ex = np.nan_to_num(np.stack([z["e%d/reset_extra" % e] for e in range(E)]))
    ex = np.concatenate([ex, np.zeros((E, 2))], axis=1)
    env.set_replay(reset_draws=np.nan_to_num(np.stack([z["e%d/reset_draws" % e] for e in range(E)])),
                   reset_fill=np.stack([z["e%d/reset_fill" % e] for e in range(E)]), reset_env_draws=ex)
    obs0 = env.reset_tensors().cpu().numpy()
    for e in range(E):
        assert np.array_equal(obs0[e], z["e%d/reset_obs" % e]), (name, e, "reset obs")
    # every env walks its own list of INNER steps; a launch steps the envs whose next inner step is of the launch's
    # kind (dropped repeat / commanded step), the others sit it out behind the mask
    cur = [0] * E
    lens = [len(z["e%d/action" % e]) for e in range(E)]
    repeats = 0
    while any(cur[e] < lens[e] for e in range(E)):
        for kind_extra in (True, False):
            mask = np.array([cur[e] < lens[e] and bool(z["e%d/is_extra" % e][cur[e]]) == kind_extra for e in range(E)])
            if not mask.any():
                continue
            act = np.array([int(z["e%d/action" % e][cur[e]]) if mask[e] else 4 for e in range(E)], dtype=np.int32)
            dr = np.stack([np.nan_to_num(z["e%d/draws" % e][cur[e]]) if mask[e] else np.zeros((6, n)) for e in range(E)])
            xs = np.stack([np.concatenate([np.nan_to_num(z["e%d/extra" % e][cur[e]]), [0, 0]]) if mask[e] else np.zeros(4)
                           for e in range(E)])
            env.set_replay(step_draws=dr, step_env_draws=xs)
            a_dev = torch.as_tensor(act, device=env.device)
            if kind_extra:
                env.step_extra_tensors(a_dev, torch.as_tensor(mask))
                repeats += int(mask.sum())
            else:
                with torch.cuda.device(env.device):       # the commanded step, for the envs that have one left
                    pkg._native.check(env.lib.uavsim_step_ex(env._h, a_dev.data_ptr(),
                                                             torch.as_tensor(mask.astype(np.uint8), device=env.device).data_ptr(),
                                                             0, env._stream()))
            obs = env.obs.cpu().numpy()
            rew = env.reward.cpu().numpy()
            done = env.done.cpu().numpy()
            rec = env.records()
            sf = env.state()["spreading_factor"].cpu().numpy()
            for e in np.flatnonzero(mask):
                t = cur[e]
                tag = (name, int(e), t, kind_extra)
                assert np.array_equal(obs[e], z["e%d/obs" % e][t]), tag
                assert bool(done[e]) == bool(z["e%d/truncated" % e][t]), tag
                assert rec["x"][e] == z["e%d/pos" % e][t][0] and rec["y"][e] == z["e%d/pos" % e][t][1], tag
                assert rec["step"][e] == z["e%d/step" % e][t], tag
                assert np.array_equal(sf[e, :n], z["e%d/sf" % e][t]), tag
                assert rec["boundary_hits"][e] == z["e%d/boundary_hits" % e][t], tag
                assert rec["edge_steps"][e] == z["e%d/edge_steps" % e][t], tag
                ref = float(z["e%d/reward" % e][t])
                assert abs(rew[e] - ref) <= 1e-9 * max(abs(ref), 1.0), (tag, rew[e], ref)
                tr = float(z["e%d/total_reward" % e][t])        # env.total_reward includes the dropped repeats' rewards
                assert abs(rec["total_reward"][e] - tr) <= 1e-9 * max(abs(tr), 1.0), tag
                cur[e] += 1
    if meta["wind_drift"] > 0:
        assert repeats > 10
    # what run_dqn_episode reads from the last info (sim_to_real_sweep.py:222-225)
    infos = env.get_info()
    for e in range(E):
        last = dict(zip(meta["info_keys"], z["e%d/last_info" % e]))
        assert infos[e]["coverage_percentage"] == last["coverage_percentage"]
        assert infos[e]["boundary_hits"] == last["boundary_hits"] and infos[e]["edge_steps"] == last["edge_steps"]
        assert abs(infos[e]["total_reward"] - last["total_reward"]) <= 1e-9 * max(abs(last["total_reward"]), 1.0)
        assert np.array_equal(np.asarray(infos[e]["sensor_collection_ratios"]), z["e%d/last_ratios" % e])
    env.close()


def test_philox_wind_gate_rate_and_gps_in_observation(pkg):
    """Philox mode: the repeat fires for min(0.95, wind / 2) of the move actions and never for COLLECT; the in-observation
    GPS noise is N(0, sigma) cells on the reported position (clipped to the grid), the true position stays integral."""
    import torch
    E, W = 32768, 400
    env = pkg.UAVVecEnv(E, grid_size=(W, W), num_sensors=10, seed=5, wind_extra_step=1.0, auto_reset=False,
                        uav_start_position=(200.0, 200.0), max_steps=1000)
    env.reset_tensors()
    a = torch.full((E,), 3, dtype=torch.int32, device=env.device)
    a[: E // 2] = 4
    for _ in range(4):
        env.step_tensors(a)
    rec = env.records()
    assert (rec["step"][: E // 2] == 4).all()                         # COLLECT is never repeated
    extra = rec["step"][E // 2:] - 4
    assert extra.min() >= 0 and abs(extra.mean() / 4 - 0.5) < 0.01    # p = min(0.95, 1.0 / 2)
    assert np.array_equal(rec["x"][E // 2:], 200.0 + rec["step"][E // 2:])
    env.close()

    env = pkg.UAVVecEnv(E, grid_size=(W, W), num_sensors=10, seed=6, gps_sigma=10.0, gps_in_observation=True,
                        auto_reset=False, uav_start_position=(200.0, 200.0))
    o0 = env.reset_tensors().cpu().numpy().astype(np.float64)
    env.step_tensors(torch.zeros(E, dtype=torch.int32, device=env.device))      # UP: y -> 201
    o1 = env.obs.cpu().numpy().astype(np.float64)
    rec = env.records()
    assert (rec["x"] == 200.0).all() and (rec["y"] == 201.0).all()
    for v, mu in ((o0[:, 0] * W, 200.0), (o0[:, 1] * W, 200.0), (o1[:, 0] * W, 200.0), (o1[:, 1] * W, 201.0)):
        assert abs(v.mean() - mu) < 0.25 and abs(v.std() - 10.0) < 0.2, (v.mean(), v.std())
    assert abs(np.corrcoef(o1[:, 0], o1[:, 1])[0, 1]) < 0.03
    env.close()


def test_per_env_shadowing_and_duty_columns(pkg):
    """One batch holding a whole sweep: per-env shadowing sigma scales that env's draws, per-env duty cycle scales its
    attempt rate; envs left at the handle's value reproduce a plain handle bit for bit."""
    import torch
    E, n = 4096, 20
    kw = dict(grid_size=(100, 100), num_sensors=n, seed=11, auto_reset=False, uav_start_position=(50.0, 50.0),
              export_draws=True, step_detail=True)
    plain = pkg.UAVVecEnv(E, **kw)
    sweep = pkg.UAVVecEnv(E, **kw)
    sig = np.where(np.arange(E) % 2 == 0, 4.0, 8.0).astype(np.float32)
    duty = np.where(np.arange(E) % 4 < 2, 10.0, 100.0).astype(np.float32)
    sweep.set_env_params(shadowing_std_db=sig, duty_cycle=duty)
    plain.reset_tensors()
    sweep.reset_tensors()
    a = torch.full((E,), 4, dtype=torch.int32, device=plain.device)
    att = np.zeros(E)
    for _ in range(6):
        plain.step_tensors(a)
        sweep.step_tensors(a)
        d = sweep.step_detail.cpu().numpy()
        att += np.array([bin(int(x)).count("1") for x in d[:, 8]])
    ep, es = plain.draw_export.cpu().numpy(), sweep.draw_export.cpu().numpy()
    same = (sig == 4.0)
    assert np.array_equal(ep[same][:, [0, 1, 4, 5]], es[same][:, [0, 1, 4, 5]])            # untouched envs: identical draws
    assert np.allclose(es[~same][:, [4, 5]], 2.0 * ep[~same][:, [4, 5]], rtol=1e-5, atol=1e-5)   # sigma 8 = 2 x sigma 4
    base = (sig == 4.0) & (duty == 10.0)
    assert torch.equal(plain.obs[torch.as_tensor(base, device=plain.device)], sweep.obs[torch.as_tensor(base, device=plain.device)])
    lo, hi = att[duty == 10.0].mean(), att[duty == 100.0].mean()
    assert hi > 5 * lo > 0, (lo, hi)                                                        # duty 100 % attempts ~10x as often
    plain.close()
    sweep.close()
