"""The drop-in boundary on a GPU: what reference callers read from `infos[i]` at episode end,
`reset(seed=)`, checkpoints with a frame ring, and the SB3 VecEnv subclass in action."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def pkg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from __graft_entry__ import import_package
    return import_package()


@pytest.mark.parametrize("name,seeds", [("lowbatt_200_n20", [0, 1, 2]), ("maxsteps_150_n8", [0, 1])])
def test_infos_at_episode_end_carry_the_reference_get_info_payload(pkg, name, seeds):
    """numpy VecEnv path with auto-reset, replaying the reference's golden episodes: when an episode ends, infos[i]
    holds `_get_info()` of the TERMINAL state -- the keys `BestByMetricCallback` (dqn.py:1148-1155) and the
    sim-to-real sweep (sim_to_real_sweep.py:222-225) read -- equal to what the reference's own info dict held."""
    from golden_util import Fixture, replay_stream
    fx = Fixture(name)
    eps = [replay_stream(fx, s) for s in seeds]
    E, n = len(seeds), fx.n
    xy = np.stack([ep[3].spos for ep in eps])
    env = pkg.UAVVecEnv(E, grid_size=fx.grid, num_sensors=n, sensor_positions=xy, uav_start_position=fx.start,
                        precision="f64", rng="replay", auto_reset=True, **fx.meta["env_kwargs"])
    env.set_replay(reset_draws=np.stack([ep[1]["draws"] for ep in eps]), reset_fill=np.stack([ep[1]["fill"] for ep in eps]))
    env.reset()
    lens = [len(ep[0]["actions"]) for ep in eps]
    seen = 0
    for t in range(max(lens)):
        act = np.array([int(ep[0]["actions"][t]) if t < lens[k] else 0 for k, ep in enumerate(eps)])
        dr = np.stack([np.nan_to_num(ep[2]["draws"][t], nan=0.0) if t < lens[k] else np.zeros((6, n))
                       for k, ep in enumerate(eps)])
        env.set_replay(step_draws=dr)
        obs, rew, dones, infos = env.step(act)
        for k, (g, first, S, oenv) in enumerate(eps):
            if t != lens[k] - 1:
                continue
            assert dones[k] and g["truncated"][t], (name, seeds[k])
            inf = infos[k]
            assert inf["TimeLimit.truncated"] is True and inf["terminal_observation"].shape == obs[k].shape
            assert np.array_equal(inf["terminal_observation"], S["obs"][t])
            assert np.array_equal(inf["uav_position"], g["pos"][t])
            assert inf["battery"] == g["battery"][t] and inf["current_step"] == lens[k]
            assert inf["sensors_collected"] == g["i_sensors_collected"][t]
            assert inf["high_urgency_sensors"] == g["i_high_urgency_sensors"][t]
            assert inf["boundary_hits"] == g["boundary_hits"][t] and inf["edge_steps"] == g["edge_steps"][t]
            assert inf["capture_effect_triggers"] == g["capture_triggers"][t]
            assert np.array_equal(np.asarray(inf["sensor_collection_ratios"]), g["sp_i_ratios"][-1])
            assert inf["max_urgency"] == g["sp_i_max_urgency"][-1] and inf["avg_urgency"] == g["sp_i_avg_urgency"][-1]
            assert inf["coverage_percentage"] == g["sp_i_coverage_percentage"][-1]
            assert abs(inf["total_reward"] - g["sp_total_reward"][-1]) <= 1e-9 * abs(g["sp_total_reward"][-1])
            assert abs(inf["total_data_collected"] - g["sp_total_data_collected"][-1]) <= 1e-9 * max(1.0, g["sp_total_data_collected"][-1])
            assert abs(inf["episode"]["r"] - g["reward"].sum()) <= 1e-9 * abs(g["reward"].sum())
            assert inf["episode"]["l"] == lens[k]
            assert inf["is_alive"] == (g["battery"][t] > 0.02 * env.cfg.max_battery)
            # the callers' own arithmetic (dqn.py:1150-1155) then sees the reference's numbers
            crs = np.array(inf.get("sensor_collection_ratios", [1.0]), dtype=np.float64)
            assert crs.shape == (n,)
            # and the env has restarted in the same call: the obs handed back is the new episode's first one
            assert obs[k, 2] == 1.0
            seen += 1
        live = [k for k in range(E) if t < lens[k] - 1]
        for k in live:
            assert infos[k] == {}, "live envs carry no payload unless info_every_step is set"
    assert seen == E
    env.close()


def test_info_every_step_mirrors_the_reference_step_return(pkg):
    env = pkg.UAVVecEnv(8, grid_size=(100, 100), num_sensors=10, seed=4, info_every_step=True, max_battery=1.0)
    env.reset()
    rng = np.random.default_rng(0)
    for t in range(12):
        obs, rew, dones, infos = env.step(rng.integers(0, 5, 8))
        for e in range(8):
            assert set(infos[e]) >= {"uav_position", "battery", "sensor_collection_ratios", "coverage_percentage",
                                     "total_reward", "boundary_hits", "edge_steps"}
            if not dones[e]:
                assert infos[e]["current_step"] >= 1 and abs(infos[e]["battery_percent"] / 100 - obs[e, 2]) < 1e-6
    assert dones.any() or True
    env.close()


def test_reset_seed_rekeys_the_streams(pkg):
    """`env.reset(seed=s)` (uav_env.py:400-410; sweep_eval.py:415,463,584 reseed every evaluation episode)."""
    kw = dict(grid_size=(300, 300), num_sensors=20, sensor_positions=np.random.default_rng(1).uniform(0, 300, (20, 2)))
    a = pkg.UAVVecEnv(32, seed=0, **kw)
    b = pkg.UAVVecEnv(32, seed=5, **kw)
    o0 = a.reset()
    ob = b.reset()
    assert not np.array_equal(o0, ob)
    o5 = a.reset(seed=5)
    assert np.array_equal(o5, ob)                       # the same key -> the same fills and shadowing draws
    act = np.random.default_rng(2).integers(0, 5, 32)
    ra, rb = a.step(act), b.step(act)
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
    assert a.seed(7)[:2] == [7, 8]                      # VecEnv.seed: applied by the next reset()
    o7 = a.reset()
    assert not np.array_equal(o7, o5)
    assert np.array_equal(a.reset(seed=0), o0)
    a.close()
    b.close()


@pytest.mark.parametrize("n_stack,latency", [(4, 0), (3, 2)])
def test_snapshot_restore_with_a_frame_ring(pkg, n_stack, latency):
    """A checkpoint taken mid-episode with stacked (and delayed) frames resumes to the identical observations: the
    frame ring and its position are part of the snapshot."""
    import torch
    E = 64
    env = pkg.UAVVecEnv(E, grid_size=(200, 200), num_sensors=16, n_stack=n_stack, latency_steps=latency, seed=9,
                        max_battery=2.0)
    env.reset_tensors()
    g = torch.Generator(device=env.device)
    g.manual_seed(3)
    acts = [torch.randint(0, 5, (E,), dtype=torch.int32, device=env.device, generator=g) for _ in range(30)]
    for t in range(7):
        env.step_tensors(acts[t])
    snap = env.snapshot()
    ref = []
    for t in range(7, 30):
        o, r, d = env.step_tensors(acts[t])
        ref.append((o.clone(), r.clone(), d.clone(), env.terminal_obs.clone()))
    assert sum(int(x[2].sum()) for x in ref) > 0         # restarts happened after the checkpoint
    env.restore(snap)
    for t in range(7, 30):
        o, r, d = env.step_tensors(acts[t])
        ro, rr, rd, rt = ref[t - 7]
        assert torch.equal(o, ro) and torch.equal(r, rr) and torch.equal(d, rd) and torch.equal(env.terminal_obs, rt), t
    env.close()


SB3_PROBE = r"""
import sys
sys.path.insert(0, %r)
sys.path.insert(0, %r)
import numpy as np
from __graft_entry__ import import_package
pkg = import_package()
from stable_baselines3.common.vec_env import VecEnv, VecEnvWrapper
env = pkg.UAVVecEnv(8, grid_size=(100, 100), num_sensors=10, max_battery=1.0, seed=1, domain_randomisation=True)
assert isinstance(env, VecEnv) and env.num_envs == 8 and env.render_mode is None
w = VecEnvWrapper(env)              # what DQN._wrap_env / VecFrameStack do: accepted as it is, not re-wrapped
obs = w.reset()
assert obs.shape == (8, 153)
w.env_method("set_curriculum_stage", 2)
done_seen = 0
for t in range(30):
    obs, rew, dones, infos = w.step(np.random.default_rng(t).integers(0, 5, 8))
    done_seen += int(dones.sum())
assert done_seen > 0
st = [s for s in w.get_attr("last_episode_stats") if s]
assert st and st[0]["time_to_coverage"] is None and set(st[0]) >= {"ndr", "jains_index", "bytes_per_wh"}
assert w.env_is_wrapped(VecEnvWrapper) == [False] * 8
w.close()
print("ok")
"""


def test_sb3_vecenv_subclass_steps_under_a_wrapper(pkg):
    r = subprocess.run([sys.executable, "-c", SB3_PROBE % (str(ROOT / "tests" / "_sb3_stub"), str(ROOT))],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
