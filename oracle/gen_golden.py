"""Generate tests/golden/*.npz by running the UNMODIFIED reference here.

TEST INFRASTRUCTURE ONLY; runs only in the build container (needs
/root/reference).  Usage:  python oracle/gen_golden.py [scenario ...]

Every fixture is produced by `/root/reference/src/environment/uav_env.py::
UAVEnvironment` driven exactly like `sweep_eval.py::_greedy_worker` (:460-486):
`random.seed(s); np.random.seed(s); env = UAVEnvironment(...); env.reset(seed=s)`.
Because the reference's generators are all seedable and the numpy legacy stream
is frozen across numpy versions, the fixture stores seeds + actions + outputs
(not the ~1 kB/step of draws); `oracle/uavsim_oracle.py` in STREAM mode re-draws
from the same generators in the same order, and a small sample of recorded
draws is stored so that ordering itself is pinned.

Two flavours (see oracle/ref_harness.py): `norm` = the two host-ISA dependent
float32 operations (log10, scalar **2) proxied to their correctly rounded values
(host independent, compared bit-exactly), `raw` = numpy's own SIMD routines on this
host (compared with the margin-aware checker).
"""
from __future__ import annotations

import json
import platform
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_harness as H  # noqa: E402

OUT = HERE.parent / "tests" / "golden"

DENSE_KEYS = ("reward", "truncated", "pos", "battery", "sf", "visited", "capture_triggers",
              "boundary_hits", "edge_steps", "last_step_bytes", "c_collision_count",
              "c_urgency_reduced", "c_sensor_urgency", "c_was_new_sensor", "c_was_empty",
              "c_all_sensors_collected", "c_bytes_collected", "c_data_loss",
              "i_high_urgency_sensors", "i_sensors_collected")
SPARSE_KEYS = ("buf", "gen", "tx", "lost", "ema", "obs", "i_ratios", "i_max_urgency",
               "i_avg_urgency", "total_reward", "total_data_collected", "i_coverage_percentage",
               "i_battery_percent")
SPARSE_EVERY = 32
DRAW_SAMPLE_STEPS = 12


def _policy_random(seed, p_collect=None):
    rng = np.random.default_rng(seed)

    def pick(env, obs, t):
        if p_collect is None:
            return int(rng.integers(0, 5))
        return 4 if rng.random() < p_collect else int(rng.integers(0, 4))
    return pick


def _policy_greedy(env):
    agent = H.load_greedy_agents().NearestSensorGreedy(env)

    def pick(env_, obs, t):
        return int(agent.select_action(obs))
    return pick


def _policy_max_throughput(env):
    agent = H.load_greedy_agents().MaxThroughputGreedyV2(env)

    def pick(env_, obs, t):
        return int(agent.select_action(obs))
    return pick


def run_episode(sc, seed):
    kw = dict(sc.get("env_kwargs", {}))
    env = H.make_env(sc["grid"], sc["n"], seed=seed, start=sc.get("start"), **kw)
    first = H.traced_reset(env, seed=seed)
    if sc["policy"] == "greedy":
        pick = _policy_greedy(env)
    elif sc["policy"] == "max_throughput":
        pick = _policy_max_throughput(env)
    else:
        pick = _policy_random(seed, sc.get("p_collect"))
    steps, actions, burn = [], [], []
    obs = first["obs"]
    T = sc["steps"]
    t = 0
    while True:
        before = H.REC.outside_normals
        a = pick(env, obs, t)
        burn.append(H.REC.outside_normals - before)
        st = H.traced_step(env, a)
        steps.append(st)
        actions.append(a)
        obs = st["obs"]
        t += 1
        if st["truncated"] or (T is not None and t >= T):
            break
    S = H.stack_steps(steps)
    S["obs_sum"] = S["obs"].astype(np.float64).sum(axis=1)
    sparse_idx = np.unique(np.concatenate([np.arange(0, t, SPARSE_EVERY), [t - 1]]))
    out = {"actions": np.array(actions, dtype=np.int8), "burn": np.array(burn, dtype=np.int16),
           "sparse_idx": sparse_idx.astype(np.int32),
           "positions": np.array(env.sensor_positions, dtype=np.float64),
           "reset_obs": first["obs"], "reset_fill": first["fill"], "reset_sf": first["sf"],
           "reset_ema": first["ema"], "reset_buf": first["buf"],
           "draws_head": np.concatenate([first["draws"][None], S["draws"][:DRAW_SAMPLE_STEPS]])}
    if sc.get("slim"):      # driver fixtures: the env is pinned elsewhere; keep what pins the AGENT
        for k in ("reward", "truncated", "pos", "c_bytes_collected"):
            out[k] = S[k]
        for k in ("positions", "reset_obs", "reset_sf", "reset_ema", "reset_buf", "draws_head", "sparse_idx"):
            out.pop(k)
        return out
    for k in DENSE_KEYS + ("obs_sum",):
        out[k] = S[k]
    for k in SPARSE_KEYS:
        out["sp_" + k] = S[k][sparse_idx]
    return out


SCENARIOS = {
    # config 1: the reference's own CPU-runnable case
    "c1_greedy_500_n20": dict(grid=(500, 500), n=20, policy="greedy", steps=None, seeds=list(range(20))),
    # MaxThroughputGreedyV2 (greedy_agents.py:105-200) driving the same configuration: pins the restated driver
    "c1_maxtp_500_n20": dict(grid=(500, 500), n=20, policy="max_throughput", steps=None, seeds=list(range(8)),
                             slim=True),
    "maxtp_200_n30": dict(grid=(200, 200), n=30, policy="max_throughput", steps=None, seeds=[0, 1, 2],
                          start=(100.0, 100.0), slim=True, env_kwargs=dict(max_battery=60.0)),
    # config 2 episodes (per-env seed = env id, uniform random actions); 4 of them to truncation
    "c2_random_500_n20": dict(grid=(500, 500), n=20, policy="random", steps=320, seeds=list(range(24))),
    "c2_random_500_n20_full": dict(grid=(500, 500), n=20, policy="random", steps=None, seeds=[0, 1, 2, 3]),
    # config 5 geometry (1000x1000, N=50), collect-heavy so capture/collisions occur
    "c5_mixed_1000_n50": dict(grid=(1000, 1000), n=50, policy="random", p_collect=0.5, steps=300,
                              seeds=[0, 1, 2, 3]),
    # dense small grid: many simultaneous transmitters (collisions, capture effect, drains)
    "dense_100_n40": dict(grid=(100, 100), n=40, policy="random", p_collect=0.7, steps=400,
                          start=(50.0, 50.0), seeds=list(range(6))),
    # duty cycle 100 %: every in-range sensor transmits (stress for SF grouping / capture)
    "duty100_200_n30": dict(grid=(200, 200), n=30, policy="random", p_collect=0.6, steps=300,
                            start=(100.0, 100.0), seeds=[0, 1, 2],
                            env_kwargs=dict(sensor_duty_cycle=100.0)),
    # buffers saturate (data loss path) quickly: small buffer, long horizon handled by full episodes
    "smallbuf_300_n10": dict(grid=(300, 300), n=10, policy="random", steps=400, seeds=[0, 1],
                             env_kwargs=dict(max_buffer_size=400.0)),
    # 5-feature observation variant (include_sensor_positions, uav_env.py:668-672)
    "fps5_300_n12": dict(grid=(300, 300), n=12, policy="random", steps=120, seeds=[0, 1],
                         env_kwargs=dict(include_sensor_positions=True)),
    # collect duration != move duration (dt differs per action)
    "cd2_200_n16": dict(grid=(200, 200), n=16, policy="random", p_collect=0.5, steps=200,
                        start=(90.0, 110.0), seeds=[0, 1], env_kwargs=dict(collection_duration=2.0)),
    # short battery: truncation by battery + terminal penalties within a short fixture
    "lowbatt_200_n20": dict(grid=(200, 200), n=20, policy="random", p_collect=0.5, steps=None,
                            start=(100.0, 100.0), seeds=[0, 1, 2], env_kwargs=dict(max_battery=30.0)),
    # max_steps truncation
    "maxsteps_150_n8": dict(grid=(150, 150), n=8, policy="random", steps=None, seeds=[0, 1],
                            env_kwargs=dict(max_steps=150)),
}


def generate(name, flavour):
    sc = SCENARIOS[name]
    H.load_reference()
    H.REC.normalise = (flavour == "norm")
    blob = {}
    for seed in sc["seeds"]:
        ep = run_episode(sc, seed)
        for k, v in ep.items():
            blob["s%d/%s" % (seed, k)] = v
    meta = dict(name=name, flavour=flavour, grid=list(sc["grid"]), n=sc["n"], seeds=sc["seeds"],
                policy=sc["policy"], start=list(sc.get("start") or (0.0, 0.0)),
                env_kwargs=sc.get("env_kwargs", {}), sparse_every=SPARSE_EVERY,
                numpy=np.__version__, python=platform.python_version(), machine=platform.machine(),
                generator="oracle/gen_golden.py", reference_entry="src/environment/uav_env.py::UAVEnvironment")
    blob["meta"] = np.array(json.dumps(meta))
    OUT.mkdir(parents=True, exist_ok=True)
    path = OUT / ("%s.%s.npz" % (name, flavour))
    np.savez_compressed(path, **blob)
    print("%-36s %-4s %7.1f kB  steps/seed=%s" % (
        name, flavour, path.stat().st_size / 1e3,
        [int(blob["s%d/actions" % s].shape[0]) for s in sc["seeds"]][:6]))


# --------------------------------------------------------------------------
# DomainRandEnv sequences (dqn.py:177-451): several chained episodes per worker, auto-reset like DummyVecEnv.
# Its layout / grid / start draws come from unseeded or global generators, so these fixtures DO store the
# slot-addressed draw tables and the per-episode scenario.
# --------------------------------------------------------------------------
DR_SCENARIOS = {
    "dr_stage4_n20": dict(n=20, stage=4, max_steps=70, steps=300, seed=11, p_collect=0.4),
    "dr_stage2_n10": dict(n=10, stage=2, max_steps=50, steps=220, seed=12, p_collect=0.5),
    "dr_stage4_n40": dict(n=40, stage=4, max_steps=60, steps=200, seed=13, p_collect=0.45),
}
DR_STEP_KEYS = ("reward", "truncated", "pos", "battery", "sf", "visited", "capture_triggers", "boundary_hits",
                "edge_steps", "obs", "buf", "tx", "lost", "ema", "gen", "draws", "prev_dist", "first_cov",
                "total_reward", "c_collision_count", "c_bytes_collected")
DR_STAT_KEYS = ("total_generated", "total_collected", "total_lost", "battery_remaining", "ndr", "fairness_std",
                "jains_index", "data_efficiency", "bytes_per_wh")


def generate_domain_rand(name):
    import random
    sc = DR_SCENARIOS[name]
    H.load_reference()
    H.REC.normalise = True
    cls, npx = H.load_domain_rand_env()
    random.seed(sc["seed"])
    np.random.seed(sc["seed"])
    npx.layout_seed, npx.layout_calls = sc["seed"], 0
    base = dict(max_steps=sc["max_steps"], path_loss_exponent=3.8, rssi_threshold=-85.0, sensor_duty_cycle=10.0,
                max_battery=274.0, render_mode=None)
    env = cls(fixed_num_sensors=sc["n"], curriculum_stage=sc["stage"], base_config=base)
    rng = np.random.default_rng(sc["seed"])
    blob, ep, steps, actions = {}, 0, [], []

    def flush(first):
        S = H.stack_steps(steps)
        for k in DR_STEP_KEYS:
            blob["e%d/%s" % (ep, k)] = S[k]
        blob["e%d/actions" % ep] = np.array(actions, dtype=np.int8)
        for k in ("grid", "positions", "start", "init_sf", "obs", "draws", "sf", "ema", "buf"):
            blob["e%d/reset_%s" % (ep, k)] = first[k]

    first = H.dr_traced_reset(env)
    for t in range(sc["steps"]):
        a = 4 if rng.random() < sc["p_collect"] else int(rng.integers(0, 4))
        st = H.dr_traced_step(env, a)
        steps.append(st)
        actions.append(a)
        if st["truncated"]:
            flush(first)
            first = H.dr_traced_reset(env)           # DummyVecEnv: reset right after the terminal step
            les = env.last_episode_stats
            blob["e%d/stats" % ep] = np.array([les[k] for k in DR_STAT_KEYS], dtype=np.float64)
            blob["e%d/stats_ttc" % ep] = np.int32(-1 if les["time_to_coverage"] is None else les["time_to_coverage"])
            ep, steps, actions = ep + 1, [], []
    meta = dict(name=name, flavour="norm", n=sc["n"], stage=sc["stage"], episodes=ep, max_steps=sc["max_steps"],
                stat_keys=list(DR_STAT_KEYS), numpy=np.__version__, python=platform.python_version(),
                generator="oracle/gen_golden.py", reference_entry="src/agents/dqn/dqn.py::DomainRandEnv (exec'd slice)")
    blob["meta"] = np.array(json.dumps(meta))
    path = OUT / ("%s.norm.npz" % name)
    np.savez_compressed(path, **blob)
    print("%-36s norm %7.1f kB  episodes=%d" % (name, path.stat().st_size / 1e3, ep))


# --------------------------------------------------------------------------
# RelationalUAVEnv sequences (env_wrapper.py:110-313): Dict observation + potential shaping + dwell bonus.
# --------------------------------------------------------------------------
REL_SCENARIOS = {
    "rel_300_n20": dict(grid=(300, 300), n=20, seed=21, steps=260, p_collect=0.55, start=(150.0, 150.0),
                        kwargs=dict(max_steps=130, max_battery=274.0)),
    "rel_150_n35": dict(grid=(150, 150), n=35, seed=22, steps=200, p_collect=0.7, start=(70.0, 80.0),
                        kwargs=dict(max_steps=100, max_battery=274.0)),
}
REL_STEP_KEYS = ("reward", "truncated", "pos", "battery", "sf", "visited", "obs", "buf", "tx", "lost", "ema", "gen",
                 "draws", "dwell_draws", "potential", "dwell_steps", "total_reward", "m_ndr", "m_jains", "m_efficiency")


def generate_relational(name):
    import random
    sc = REL_SCENARIOS[name]
    H.load_reference()
    H.REC.normalise = True
    cls = H.load_relational_env()
    random.seed(sc["seed"])
    np.random.seed(sc["seed"])
    env = cls(grid_size=sc["grid"], num_sensors=sc["n"], n_max=50, uav_start_position=sc["start"],
              path_loss_exponent=3.8, rssi_threshold=-85.0, sensor_duty_cycle=10.0, **sc["kwargs"])
    rng = np.random.default_rng(sc["seed"])
    blob, ep, steps, actions = {}, 0, [], []
    first = H.rel_traced_reset(env, seed=sc["seed"])
    blob["positions"] = np.array(env.sensor_positions, dtype=np.float64)
    for t in range(sc["steps"]):
        a = 4 if rng.random() < sc["p_collect"] else int(rng.integers(0, 4))
        st = H.rel_traced_step(env, a)
        steps.append(st)
        actions.append(a)
        if st["truncated"]:
            S = H.stack_steps(steps)
            for k in REL_STEP_KEYS:
                blob["e%d/%s" % (ep, k)] = S[k]
            blob["e%d/actions" % ep] = np.array(actions, dtype=np.int8)
            for k in ("obs", "mask", "draws", "fill", "sf", "ema", "buf", "potential"):
                blob["e%d/reset_%s" % (ep, k)] = first[k]
            first = H.rel_traced_reset(env)
            ep, steps, actions = ep + 1, [], []
    meta = dict(name=name, flavour="norm", grid=list(sc["grid"]), n=sc["n"], start=list(sc["start"]), episodes=ep,
                env_kwargs=sc["kwargs"], numpy=np.__version__, python=platform.python_version(),
                generator="oracle/gen_golden.py",
                reference_entry="src/experiments/relational_policy/env_wrapper.py::RelationalUAVEnv")
    blob["meta"] = np.array(json.dumps(meta))
    path = OUT / ("%s.norm.npz" % name)
    np.savez_compressed(path, **blob)
    print("%-36s norm %7.1f kB  episodes=%d" % (name, path.stat().st_size / 1e3, ep))


# --------------------------------------------------------------------------
# Sim-to-real evaluation wrappers (sim_to_real_analysis.py:170-245): GPS noise on the observed position, systematic
# wind drift, observation latency.  Chained episodes with DummyVecEnv-style resets; the wrappers' own draws are
# stored next to the slot tables ("extra": [gps_x, gps_y] per step / reset, [angle] after a WindDriftEnv reset).
# --------------------------------------------------------------------------
S2R_SCENARIOS = {
    "gps5_300_n12": dict(cls="GPSNoisyEnv", extra=dict(gps_sigma=5.0), grid=(300, 300), n=12, seed=31, steps=220,
                         p_collect=0.35, kwargs=dict(max_steps=90, max_battery=274.0, uav_start_position=(140.0, 160.0))),
    "wind1_200_n12": dict(cls="WindDriftEnv", extra=dict(wind_drift=1.0), grid=(200, 200), n=12, seed=32, steps=260,
                          p_collect=0.3, kwargs=dict(max_steps=100, max_battery=274.0, uav_start_position=(90.0, 110.0))),
    "wind2_120_n20": dict(cls="WindDriftEnv", extra=dict(wind_drift=2.0), grid=(120, 120), n=20, seed=33, steps=240,
                          p_collect=0.4, kwargs=dict(max_steps=80, max_battery=274.0, uav_start_position=(5.0, 100.0))),
    "lat2_200_n10": dict(cls="LatencyEnv", extra=dict(latency_steps=2), grid=(200, 200), n=10, seed=34, steps=150,
                         p_collect=0.4, kwargs=dict(max_steps=60, max_battery=274.0, uav_start_position=(100.0, 100.0))),
}
S2R_STEP_KEYS = ("reward", "truncated", "pos", "battery", "sf", "visited", "obs", "buf", "tx", "lost", "ema", "draws",
                 "extra", "total_reward")


def generate_sim_to_real(name):
    import random
    sc = S2R_SCENARIOS[name]
    H.load_reference()
    H.REC.normalise = True
    ns, npx = H.load_sim_to_real_envs()
    random.seed(sc["seed"])
    np.random.seed(sc["seed"])
    layout = np.random.uniform([0, 0], list(sc["grid"]), (sc["n"], 2))      # what the constructor is about to draw
    np.random.seed(sc["seed"])
    kw = dict(grid_size=sc["grid"], num_sensors=sc["n"], path_loss_exponent=3.8, rssi_threshold=-85.0,
              sensor_duty_cycle=10.0)
    kw.update(sc["kwargs"])
    env = ns[sc["cls"]]([tuple(float(v) for v in p) for p in layout.astype(np.float32)], **sc["extra"], **kw)
    assert np.array_equal(np.array(env.sensor_positions, dtype=np.float32), layout.astype(np.float32))
    rng = np.random.default_rng(sc["seed"])
    blob, ep, steps, actions = {}, 0, [], []
    first = H.s2r_traced_reset(env, npx, seed=sc["seed"])
    for t in range(sc["steps"]):
        a = 4 if rng.random() < sc["p_collect"] else int(rng.integers(0, 4))
        st = H.s2r_traced_step(env, npx, a)
        steps.append(st)
        actions.append(a)
        if st["truncated"] or t == sc["steps"] - 1:
            S = H.stack_steps(steps)
            for k in S2R_STEP_KEYS:
                blob["e%d/%s" % (ep, k)] = S[k]
            blob["e%d/actions" % ep] = np.array(actions, dtype=np.int8)
            for k in ("obs", "draws", "fill", "sf", "ema", "buf", "extra", "pos"):
                blob["e%d/reset_%s" % (ep, k)] = first[k]
            ep, steps, actions = ep + 1, [], []
            if st["truncated"]:
                first = H.s2r_traced_reset(env, npx)       # DummyVecEnv: reset right after the terminal step
    meta = dict(name=name, flavour="norm", cls=sc["cls"], extra=sc["extra"], grid=list(sc["grid"]), n=sc["n"],
                start=list(sc["kwargs"]["uav_start_position"]), episodes=ep, env_kwargs=sc["kwargs"],
                numpy=np.__version__, python=platform.python_version(), generator="oracle/gen_golden.py",
                reference_entry="src/agents/dqn/dqn_evaluation_results/sim_to_real_analysis.py::%s (exec'd slice)" % sc["cls"])
    blob["positions"] = layout.astype(np.float32)
    blob["meta"] = np.array(json.dumps(meta))
    path = OUT / ("%s.norm.npz" % name)
    np.savez_compressed(path, **blob)
    print("%-36s norm %7.1f kB  episodes=%d" % (name, path.stat().st_size / 1e3, ep))


# --------------------------------------------------------------------------
# PerturbedDQNEnv of the sim-to-real sweep (sim_to_real_sweep.py:107-139 on AnalysisUAVEnv, compare_agents.py:62-147):
# per-sensor shadowing sigma, GPS noise applied to the position inside the observation pass, probabilistic extra step
# for wind.  One fixture = chained episodes; every INNER env step is stored (flag `is_extra` marks the wind repeats),
# with its slot draws, the GPS pair drawn inside its observation ("extra") and the wind gate's uniform ("pre").
# --------------------------------------------------------------------------
PERT_SCENARIOS = {
    "pert_gps10_200_n20": dict(gps_sigma=10.0, shadow_sigma=4.0, wind_drift=0.0, seed=41, steps=200, p_collect=0.35, max_steps=90),
    "pert_shadow8_200_n20": dict(gps_sigma=0.0, shadow_sigma=8.0, wind_drift=0.0, seed=42, steps=200, p_collect=0.5, max_steps=90),
    "pert_wind1_200_n20": dict(gps_sigma=0.0, shadow_sigma=4.0, wind_drift=1.0, seed=43, steps=220, p_collect=0.3, max_steps=120),
    "pert_all_200_n20": dict(gps_sigma=5.0, shadow_sigma=6.0, wind_drift=2.0, seed=44, steps=220, p_collect=0.35, max_steps=150),
}
PERT_KEYS = ("reward", "truncated", "pos", "battery", "sf", "visited", "obs", "buf", "tx", "lost", "ema", "draws", "extra", "pre",
             "total_reward", "boundary_hits", "edge_steps", "step")
PERT_INFO = ("coverage_percentage", "total_reward", "boundary_hits", "edge_steps", "total_data_collected", "battery")


def generate_perturbed(name):
    import random
    sc = PERT_SCENARIOS[name]
    H.load_reference()
    H.REC.normalise = True
    make, npx = H.load_perturbed_dqn_env()
    grid, n = (200, 200), 20                     # sim_to_real_sweep.py:68-69
    Env = make(sc["gps_sigma"], sc["shadow_sigma"], sc["wind_drift"])
    random.seed(sc["seed"])                      # run_dqn_episode, sim_to_real_sweep.py:198-205
    np.random.seed(sc["seed"])
    layout = np.random.uniform([0, 0], list(grid), (n, 2))       # what the constructor is about to draw
    np.random.seed(sc["seed"])
    env = Env(max_sensors_limit=50, grid_size=grid, num_sensors=n, max_steps=sc["max_steps"], path_loss_exponent=3.8,
              rssi_threshold=-85.0, sensor_duty_cycle=10.0, uav_start_position=(0, 0))
    assert np.array_equal(np.array(env.sensor_positions, dtype=np.float32), layout.astype(np.float32))
    rng = np.random.default_rng(sc["seed"])
    blob, ep, inner, outer_actions = {}, 0, [], []
    env.reset(seed=sc["seed"])
    first = env.inner_reset
    for t in range(sc["steps"]):
        a = 4 if rng.random() < sc["p_collect"] else int(rng.integers(0, 4))
        env.inner = []
        obs, r, term, trunc, info = env.step(a)
        for i, st in enumerate(env.inner):
            st = dict(st)
            st["is_extra"] = np.bool_(i < len(env.inner) - 1)
            st["action"] = np.int8(a)
            st.pop("info", None)
            inner.append(st)
        outer_actions.append(a)
        if trunc or t == sc["steps"] - 1:
            S = H.stack_steps(inner)
            for k in PERT_KEYS + ("is_extra", "action"):
                blob["e%d/%s" % (ep, k)] = S[k]
            for k in ("obs", "draws", "fill", "sf", "ema", "buf", "extra", "pos"):
                blob["e%d/reset_%s" % (ep, k)] = first[k]
            blob["e%d/last_info" % ep] = np.array([float(info[k]) for k in PERT_INFO])
            blob["e%d/last_ratios" % ep] = np.array(info["sensor_collection_ratios"], dtype=np.float64)
            blob["e%d/wrapper_steps" % ep] = np.int32(len(outer_actions))
            ep, inner, outer_actions = ep + 1, [], []
            if trunc:
                env.reset()                       # DummyVecEnv: reset right after the terminal step
                first = env.inner_reset
    meta = dict(name=name, flavour="norm", cls="PerturbedDQNEnv", gps_sigma=sc["gps_sigma"], shadow_sigma=sc["shadow_sigma"],
                wind_drift=sc["wind_drift"], grid=list(grid), n=n, start=[0.0, 0.0], episodes=ep, seed=sc["seed"],
                env_kwargs=dict(max_steps=sc["max_steps"]), info_keys=list(PERT_INFO),
                numpy=np.__version__, python=platform.python_version(), generator="oracle/gen_golden.py",
                reference_entry="src/agents/dqn/dqn_evaluation_results/sim_to_real_sweep.py::_perturb_positions_dqn "
                                "(exec'd slice) over compare_agents.py::AnalysisUAVEnv (exec'd slice)")
    blob["positions"] = layout.astype(np.float32)
    blob["meta"] = np.array(json.dumps(meta))
    path = OUT / ("%s.norm.npz" % name)
    np.savez_compressed(path, **blob)
    print("%-36s norm %7.1f kB  episodes=%d" % (name, path.stat().st_size / 1e3, ep))


RAW_TOO = ("c1_greedy_500_n20",)

if __name__ == "__main__":
    names = sys.argv[1:] or (list(SCENARIOS) + list(DR_SCENARIOS) + list(REL_SCENARIOS) + list(S2R_SCENARIOS) +
                             list(PERT_SCENARIOS))
    for nm in names:
        if nm in PERT_SCENARIOS:
            generate_perturbed(nm)
            continue
        if nm in S2R_SCENARIOS:
            generate_sim_to_real(nm)
            continue
        if nm in DR_SCENARIOS:
            generate_domain_rand(nm)
            continue
        if nm in REL_SCENARIOS:
            generate_relational(nm)
            continue
        generate(nm, "norm")
        if nm in RAW_TOO:
            generate(nm, "raw")
