"""RelationalUAVEnv on the device: the CUDA step in relational mode replays the
reference's own sequences (tests/golden/rel_*.npz from src/experiments/relational_policy/env_wrapper.py):
Dict observation rows, potential shaping, cluster dwell bonus, reward scale, in-kernel restarts."""
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


@pytest.mark.parametrize("name", ["rel_300_n20", "rel_150_n35"])
def test_cuda_reproduces_reference_relational_sequences(pkg, name):
    import torch
    from golden_util import GOLDEN
    z = np.load(GOLDEN / ("%s.norm.npz" % name))
    meta = json.loads(str(z["meta"]))
    n = meta["n"]
    env = pkg.UAVVecEnv(1, grid_size=meta["grid"], num_sensors=n, sensor_positions=z["positions"].astype(np.float32),
                        uav_start_position=meta["start"], relational=True, precision="f64", rng="replay",
                        auto_reset=True, **meta["env_kwargs"])
    assert env.obs.shape == (1, 253)
    g = lambda ep, k: z["e%d/%s" % (ep, k)]      # noqa: E731

    def reset_tables(ep):
        return dict(reset_draws=np.nan_to_num(g(ep, "reset_draws"))[None], reset_fill=g(ep, "reset_fill")[None])
    env.set_replay(**reset_tables(0))
    obs = env.reset_tensors().cpu().numpy()
    assert np.array_equal(obs[0], g(0, "reset_obs"))
    d = env.relational_obs()
    assert d["uav"].shape == (1, 3) and d["sensors"].shape == (1, 50, 5) and d["mask"].shape == (1, 50)
    assert np.array_equal(d["mask"].cpu().numpy()[0], g(0, "reset_mask"))
    diag = np.sqrt(meta["grid"][0] ** 2 + meta["grid"][1] ** 2)
    for ep in range(meta["episodes"]):
        nxt = reset_tables(min(ep + 1, meta["episodes"] - 1))
        acts = g(ep, "actions")
        for t, a in enumerate(acts):
            env.set_replay(step_draws=np.nan_to_num(g(ep, "draws")[t])[None],
                           dwell_draws=np.nan_to_num(g(ep, "dwell_draws")[t])[None], **nxt)
            env.step_tensors(torch.tensor([int(a)], dtype=torch.int32, device=env.device))
            tag = (name, ep, t)
            ref = float(g(ep, "reward")[t])
            rew = float(env.reward[0])
            assert abs(rew - ref) <= 1e-12 * max(1.0, abs(ref)) + 1e-15, (tag, rew, ref)
            done = bool(env.done[0])
            assert done == bool(g(ep, "truncated")[t]), tag
            if not done:
                assert np.array_equal(env.obs.cpu().numpy()[0], g(ep, "obs")[t]), tag
                rec = env.records()[0]
                assert rec["dwell_steps"] == g(ep, "dwell_steps")[t], tag
                assert -float(rec["prev_dist_nearest"]) / diag == g(ep, "potential")[t], tag
            else:
                assert np.array_equal(env.terminal_obs.cpu().numpy()[0], g(ep, "obs")[t]), tag
                if ep + 1 < meta["episodes"]:
                    assert np.array_equal(env.obs.cpu().numpy()[0], g(ep + 1, "reset_obs")), tag
    env.close()


def test_relational_philox_batch_invariants(pkg):
    import torch
    E = 2048
    env = pkg.UAVVecEnv(E, grid_size=(300, 300), num_sensors=25, relational=True, seed=4, max_battery=8.0)
    env.reset_tensors()
    tot = torch.zeros(E, device=env.device)
    for t in range(60):
        a = torch.randint(0, 5, (E,), dtype=torch.int32, device=env.device)
        _, r, _ = env.step_tensors(a)
        tot += r
    env.poll_errors()
    d = env.relational_obs()
    s = d["sensors"]
    assert float(s[..., :2].abs().max()) <= 1.0 and float(s[:, 25:].abs().max()) == 0.0
    assert set(np.unique(s[..., 4].cpu().numpy())) <= {0.0, 1.0}
    assert float(d["mask"].sum(1).min()) == 25.0
    assert float(tot.abs().max()) < 5.0                      # 1e-5 reward scale (env_wrapper.py:62)
    rec = env.records()
    assert (rec["dwell_steps"] >= 0).all() and (rec["dwell_steps"] <= 10).all()
    env.close()
