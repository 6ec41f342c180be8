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
        infos = env.get_info()                                   # _get_info() through the device kernel, every step
        for k, (g, first, S, oenv) in enumerate(eps):
            if t >= lens[k]:
                continue
            tag = (name, seeds[k], t)
            # UAVEnvironment._get_info (uav_env.py:676-700) against the values the reference's own info dicts held
            inf = infos[k]
            assert inf["high_urgency_sensors"] == g["i_high_urgency_sensors"][t], tag
            assert inf["sensors_collected"] == g["i_sensors_collected"][t], tag
            assert inf["current_step"] == t + 1 and inf["boundary_hits"] == g["boundary_hits"][t], tag
            assert inf["edge_steps"] == g["edge_steps"][t] and inf["capture_effect_triggers"] == g["capture_triggers"][t], tag
            assert abs(inf["last_step_bytes_collected"] - g["last_step_bytes"][t]) <= 1e-6 * max(1.0, g["last_step_bytes"][t]), tag
            sp = np.flatnonzero(g["sparse_idx"] == t)
            if sp.size:
                j = int(sp[0])
                assert np.array_equal(np.asarray(inf["sensor_collection_ratios"]), g["sp_i_ratios"][j]), tag
                assert inf["max_urgency"] == g["sp_i_max_urgency"][j] and inf["avg_urgency"] == g["sp_i_avg_urgency"][j], tag
                assert inf["coverage_percentage"] == g["sp_i_coverage_percentage"][j], tag
                for a, b in ((inf["total_reward"], g["sp_total_reward"][j]), (inf["battery_percent"], g["sp_i_battery_percent"][j]),
                             (inf["total_data_collected"], g["sp_total_data_collected"][j])):
                    assert abs(a - b) <= 1e-9 * max(abs(a), abs(b), 1e-300), (tag, a, b)
                assert inf["is_alive"] == (g["battery"][t] > 0.02 * env.cfg.max_battery), tag
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


def test_cuda_reproduces_reference_domain_rand_sequences(pkg):
    """The reference's DomainRandEnv sequences (tests/golden/dr_*.npz: chained episodes with per-episode grid,
    layout, far start and inherited SF injected) replayed through the CUDA step with IN-KERNEL restarts:
    three workers with N = 20 / 10 / 40 padded to 50 observation slots in one batch."""
    import json
    import torch
    from golden_util import GOLDEN
    names = ["dr_stage4_n20", "dr_stage2_n10", "dr_stage4_n40"]
    Z = [np.load(GOLDEN / ("%s.norm.npz" % nm)) for nm in names]
    M = [json.loads(str(z["meta"])) for z in Z]
    E, P = len(names), 50
    # max_steps differs per fixture, so each sequence gets its own single-env handle
    worst = 0.0
    for i, (z, m) in enumerate(zip(Z, M)):
        n = m["n"]
        g = lambda ep, k: z["e%d/%s" % (ep, k)]      # noqa: E731

        def scen(ep):
            xy = np.zeros((1, P, 2), dtype=np.float32)
            xy[0, :n] = g(ep, "reset_positions")
            dr = np.zeros((1, 6, P))
            dr[0, :, :n] = np.nan_to_num(g(ep, "reset_draws"))
            return dict(reset_sensor_xy=xy, reset_grid_wh=g(ep, "reset_grid")[None].astype(np.int32),
                        reset_start_xy=g(ep, "reset_start")[None].astype(np.float32),
                        reset_init_sf=np.array([int(g(ep, "reset_init_sf"))], dtype=np.int32), reset_draws=dr)
        env = pkg.UAVVecEnv(1, num_sensors=n, max_sensors=P, obs_slots=50, domain_randomisation=True,
                            precision="f64", rng="replay", auto_reset=True, max_steps=m["max_steps"])
        env.set_replay(**scen(0))
        obs = env.reset_tensors().cpu().numpy()
        assert np.array_equal(obs[0], g(0, "reset_obs")), (names[i], "reset obs")
        for ep in range(m["episodes"]):
            acts = g(ep, "actions")
            nxt = scen(ep + 1) if ep + 1 < m["episodes"] else scen(ep)
            for t, a in enumerate(acts):
                sd = np.zeros((1, 6, P))
                sd[0, :, :n] = np.nan_to_num(g(ep, "draws")[t])
                env.set_replay(step_draws=sd, **nxt)
                env.step_tensors(torch.tensor([int(a)], dtype=torch.int32, device=env.device))
                rew, done = float(env.reward[0]), bool(env.done[0])
                tag = (names[i], ep, t)
                assert done == bool(g(ep, "truncated")[t]), tag
                ref = float(g(ep, "reward")[t])
                err = abs(rew - ref) / max(abs(ref), 1.0)
                worst = max(worst, err)
                assert err <= 1e-9, (tag, rew, ref)
                if not done:
                    assert np.array_equal(env.obs.cpu().numpy()[0], g(ep, "obs")[t]), tag
                    rec = env.records()[0]
                    assert (rec["x"], rec["y"]) == tuple(g(ep, "pos")[t]) and rec["battery"] == g(ep, "battery")[t], tag
                    assert rec["first_full_coverage_step"] == g(ep, "first_cov")[t], tag
                    sf = env.state()["spreading_factor"].cpu().numpy()[0, :n]
                    assert np.array_equal(sf, g(ep, "sf")[t]), tag
                else:
                    assert np.array_equal(env.terminal_obs.cpu().numpy()[0], g(ep, "obs")[t]), tag
                    st = dict(zip(m["stat_keys"], g(ep, "stats")))
                    es = dict(zip(("return", "length", "total_generated", "total_collected", "total_lost",
                                   "battery_remaining", "ndr", "fairness_std", "jains_index"),
                                  env.episode_stats.cpu().numpy()[0][:9]))
                    for k in ("total_generated", "total_collected", "total_lost", "battery_remaining", "ndr",
                              "fairness_std", "jains_index"):
                        assert abs(es[k] - st[k]) <= 1e-9 * max(1.0, abs(st[k])), (tag, k, es[k], st[k])
                    assert int(es["length"]) == len(acts)
                    if ep + 1 < m["episodes"]:      # the in-kernel restart installed the next episode
                        assert np.array_equal(env.obs.cpu().numpy()[0], g(ep + 1, "reset_obs")), tag
        env.close()
    assert worst <= 1e-9
