"""The CUDA path replaying the reference's OWN golden episodes (tests/golden, generated from the
unmodified reference by oracle/gen_golden.py).  Draws are re-derived from the fixture seeds by the
python oracle in stream mode (pinned to the goldens by tests/test_oracle_golden.py) and replayed
slot-addressed on the GPU; outputs are compared with the values the reference produced."""
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


CASES = [
    ("c1_greedy_500_n20", [0, 1, 2, 3, 4, 5], {}),
    ("c2_random_500_n20", [0, 1, 2, 3, 4, 5, 6, 7], {}),
    ("c2_random_500_n20_full", [0, 1], {}),
    ("c5_mixed_1000_n50", [0, 1, 2, 3], {}),
    ("dense_100_n40", [0, 1, 2, 3], {}),
    ("duty100_200_n30", [0, 1, 2], {}),
    ("lowbatt_200_n20", [0, 1, 2], {}),
    ("maxsteps_150_n8", [0, 1], {}),
    ("fps5_300_n12", [0, 1], {}),
    ("cd2_200_n16", [0, 1], {}),
    ("smallbuf_300_n10", [0, 1], {}),
]


@pytest.mark.parametrize("name,seeds,_", CASES)
def test_cuda_reproduces_reference_golden(pkg, name, seeds, _):
    import torch
    from golden_util import Fixture, replay_stream
    fx = Fixture(name)
    eps = [replay_stream(fx, s) for s in seeds]          # (golden ep, first, oracle trace, oracle env)
    E, n = len(seeds), fx.n
    T = max(len(ep[0]["actions"]) for ep in eps)
    xy = np.stack([ep[3].spos for ep in eps])
    env = pkg.UAVVecEnv(E, grid_size=fx.grid, num_sensors=n, sensor_positions=xy, uav_start_position=fx.start,
                        precision="f64", rng="replay", auto_reset=False, step_detail=True, **fx.meta["env_kwargs"])
    env.set_replay(reset_draws=np.stack([ep[1]["draws"] for ep in eps]),
                   reset_fill=np.stack([ep[1]["fill"] for ep in eps]))
    obs0 = env.reset_tensors().cpu().numpy()
    for k, ep in enumerate(eps):
        assert np.array_equal(obs0[k], ep[0]["reset_obs"]), (name, seeds[k], "reset obs")
    lens = [len(ep[0]["actions"]) for ep in eps]
    worst = 0.0
    for t in range(T):
        act = np.array([int(ep[0]["actions"][t]) if t < lens[k] else 4 for k, ep in enumerate(eps)], dtype=np.int32)
        dr = np.stack([np.nan_to_num(ep[2]["draws"][t], nan=0.0) if t < lens[k] else np.zeros((6, n))
                       for k, ep in enumerate(eps)])
        env.set_replay(step_draws=dr)
        env.step_tensors(torch.as_tensor(act, device=env.device))
        rew = env.reward.cpu().numpy()
        done = env.done.cpu().numpy()
        obs = env.obs.cpu().numpy()
        rec = env.records()
        sf = env.state()["spreading_factor"].cpu().numpy()
        det = env.step_detail.cpu().numpy()
        for k, (g, first, S, oenv) in enumerate(eps):
            if t >= lens[k]:
                continue
            tag = (name, seeds[k], t)
            # discrete: bit-exact against the REFERENCE's recorded values
            assert bool(done[k]) == bool(g["truncated"][t]), tag
            assert rec["x"][k] == g["pos"][t][0] and rec["y"][k] == g["pos"][t][1], tag
            assert np.array_equal(sf[k, :n], g["sf"][t]), tag
            vis = int(rec["visited_lo"][k]) | (int(rec["visited_hi"][k]) << 32)
            assert vis == sum(1 << i for i in range(n) if g["visited"][t][i]), tag
            assert rec["capture_triggers"][k] == g["capture_triggers"][t], tag
            assert rec["boundary_hits"][k] == g["boundary_hits"][t], tag
            assert rec["edge_steps"][k] == g["edge_steps"][t], tag
            if act[k] == 4:
                assert det[k, 0] == g["c_collision_count"][t], tag
                assert det[k, 4] == g["c_was_new_sensor"][t] and det[k, 5] == g["c_was_empty"][t], tag
            # floats: 1e-4 relative is the bar; report the worst
            for a, b in ((rew[k], g["reward"][t]), (rec["battery"][k], g["battery"][t]),
                         (obs[k].astype(np.float64).sum(), g["obs_sum"][t])):
                err = abs(a - b) / max(abs(a), abs(b), 1e-300) if a != b else 0.0
                worst = max(worst, err)
                assert err <= 1e-4, (tag, a, b)
            assert np.array_equal(obs[k], S["obs"][t]), tag      # and bit-exact vs the pinned oracle
    assert worst <= 1e-9, worst
    env.close()
