"""Harness that runs the UNMODIFIED reference environment in this container.

TEST INFRASTRUCTURE ONLY (golden-vector generation).  Nothing in the product
imports this file and it cannot run on the GPU box: /root/reference exists only
in the build container.  `oracle/gen_golden.py` uses it to write
`tests/golden/*.npz`.

What it does
------------
* puts stub `gymnasium` / `matplotlib` packages (oracle/_stubs) on sys.path and
  imports `/root/reference/src/environment/uav_env.py` in place (read-only
  mount, so bytecode writing is disabled);
* installs *recording proxies* -- no reference source is edited:
    - `iot_sensors.np`      -> proxy whose `random.normal` records every
                               shadowing draw (`iot_sensors.py:192`) and whose
                               `log10` can optionally be normalised to the
                               correctly-rounded float32 result (see below);
    - `uav_env.random`      -> proxy whose `random()` records the duty-cycle
                               uniforms (`uav_env.py:549`);
    - `gymnasium.Env._uniform_tap` -> records the buffer-fill uniforms
                               (`uav_env.py:410`);
    - `IoTSensor.update_spreading_factor / get_success_probability /
      is_in_range` and `UAVEnvironment._get_observation` are wrapped to tag each
      draw with its SLOT (A,B,U,C,D,E of SURVEY.md section 3.2) and sensor id;
    - `RewardFunction.calculate_collection_reward` is wrapped to capture the
      collect step's intermediate quantities (collision count, bytes, ...).

float32 normalisation
---------------------
Two float32 elementary operations inside `calculate_rssi` (iot_sensors.py:163-167,
179-183) are host-ISA dependent in numpy: `np.log10(float32)` (AVX-512 SIMD loop: 62 %
of inputs differ by one ulp from the correctly rounded value) and the scalar power
`float32 ** 2` (routed through the same SIMD `power` loop: 0.07 % of inputs differ by
one ulp from `x*x`).  With `REC.normalise = True` the proxies make both correctly
rounded and host independent -- `log10` returns `float32(log10(float64(x)))`, and the
float32 scalars inside `calculate_rssi` are a float32 subclass whose `** 2` is `x*x`
-- without touching the reference source.  The oracle and the CUDA kernels use those
definitions.  Goldens are generated in both flavours (`norm`, `raw`).
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path

import numpy as np

REFERENCE_ROOT = Path(os.environ.get("UAVSIM_REFERENCE_ROOT", "/root/reference"))
_STUBS = Path(__file__).resolve().parent / "_stubs"

SLOT_A, SLOT_B, SLOT_U, SLOT_C, SLOT_D, SLOT_E, SLOT_F = range(7)
NUM_SLOTS = 7   # F: RelationalUAVEnv's dwell-bonus range probes (env_wrapper.py:303); fixtures of the base env keep A..E


def reference_available() -> bool:
    return (REFERENCE_ROOT / "src" / "environment" / "uav_env.py").is_file()


class _Recorder:
    """Collects the draws of the env step currently being executed."""

    def __init__(self):
        self.in_env = False      # inside env.reset()/env.step()
        self.in_obs = False      # inside _get_observation()
        self.ctx = None          # (sensor_id, slot) for the next normal draw
        self.num_sensors = 0
        self.draws = None        # [6, N] float64, NaN = not drawn
        self.fill = []           # buffer-fill uniforms of the current reset
        self.outside_normals = 0 # draws consumed outside env.step/reset (agents)
        self.collect_kwargs = None
        self.normalise = False

    def begin(self, n):
        self.num_sensors = n
        self.draws = np.full((NUM_SLOTS, n), np.nan, dtype=np.float64)
        self.collect_kwargs = None
        self.in_env = True

    def end(self):
        self.in_env = False
        self.ctx = None

    def on_normal(self, v):
        if not self.in_env:
            self.outside_normals += 1
            return
        sid, slot = self.ctx
        assert getattr(self, "allow_overwrite", False) or np.isnan(self.draws[slot, sid]), ("slot drawn twice", slot, sid)
        self.draws[slot, sid] = v
        self.ctx = None

    def on_uniform01(self, v):
        assert self.in_env
        sid = self._last_sensor
        assert np.isnan(self.draws[SLOT_U, sid])
        self.draws[SLOT_U, sid] = v


REC = _Recorder()
_loaded = {}


class _RandomNormalProxy:
    def __init__(self, real_random):
        self._real = real_random

    def normal(self, loc=0.0, scale=1.0, size=None):
        v = self._real.normal(loc, scale, size)
        assert size is None and loc == 0
        REC.on_normal(float(v))
        return v

    def __getattr__(self, name):
        return getattr(self._real, name)


class N32(np.float32):
    """float32 scalar whose arithmetic stays N32 and whose `** 2` is the correctly
    rounded product (see module docstring).  Only the operators calculate_rssi uses."""

    def __sub__(self, o):
        return N32(np.float32.__sub__(self, o))

    def __add__(self, o):
        return N32(np.float32.__add__(self, o))

    def __mul__(self, o):
        return N32(np.float32.__mul__(self, o))

    def __pow__(self, o):
        assert o == 2
        return N32(np.float32.__mul__(self, self))


class _NumpyProxy:
    """Stands in for the name `np` inside the reference's iot_sensors module."""

    def __init__(self, real):
        self._real = real
        self.random = _RandomNormalProxy(real.random)

    def log10(self, x):
        if REC.normalise and isinstance(x, self._real.float32):
            return self._real.float32(self._real.log10(self._real.float64(x)))
        return self._real.log10(x)

    def array(self, obj, dtype=None, **kw):
        arr = self._real.array(obj, dtype=dtype, **kw)
        if REC.normalise and sys._getframe(1).f_code.co_name == "calculate_rssi":
            return [N32(v) for v in arr]        # only [0] / [1] are read (iot_sensors.py:163-164)
        return arr

    def sqrt(self, x):
        r = self._real.sqrt(x)
        return N32(r) if isinstance(x, N32) else r

    def __getattr__(self, name):
        return getattr(self._real, name)


class _PyRandomProxy:
    def __init__(self, real):
        self._real = real

    def random(self):
        v = self._real.random()
        REC.on_uniform01(v)
        return v

    def __getattr__(self, name):
        return getattr(self._real, name)


def load_reference():
    """Import the reference modules (once) and install the recording proxies."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    src = REFERENCE_ROOT / "src"
    for p in (str(_STUBS), str(src), str(src / "environment")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gymnasium  # the stub
    import iot_sensors  # reference, unmodified
    import uav as uav_mod
    import uav_env
    import random as py_random

    iot_sensors.np = _NumpyProxy(np)
    uav_env.random = _PyRandomProxy(py_random)
    gymnasium.Env._uniform_tap = staticmethod(lambda v: REC.fill.append(float(v)))

    S = iot_sensors.IoTSensor
    E = uav_env.UAVEnvironment
    o_usf, o_gsp, o_iir = S.update_spreading_factor, S.get_success_probability, S.is_in_range
    o_obs = E._get_observation

    def usf(self, uav_position, current_step=0):
        REC.ctx = (self.sensor_id, SLOT_D if REC.in_obs else SLOT_A)
        return o_usf(self, uav_position, current_step)

    def gsp(self, uav_position, use_advanced_model=True):
        REC.ctx = (self.sensor_id, SLOT_B if use_advanced_model else SLOT_C)
        REC._last_sensor = self.sensor_id
        return o_gsp(self, uav_position, use_advanced_model)

    def iir(self, uav_position):
        REC.ctx = (self.sensor_id, SLOT_E if REC.in_obs else SLOT_F)
        return o_iir(self, uav_position)

    def obs(self):
        REC.in_obs = True
        try:
            return o_obs(self)
        finally:
            REC.in_obs = False

    S.update_spreading_factor, S.get_success_probability, S.is_in_range = usf, gsp, iir
    E._get_observation = obs

    from rewards import reward_function as rf_mod
    o_ccr = rf_mod.RewardFunction.calculate_collection_reward

    def ccr(self, **kw):
        REC.collect_kwargs = dict(kw)
        return o_ccr(self, **kw)

    rf_mod.RewardFunction.calculate_collection_reward = ccr

    _loaded.update(dict(uav_env=uav_env, iot_sensors=iot_sensors, uav=uav_mod,
                        UAVEnvironment=E, IoTSensor=S, py_random=py_random,
                        gymnasium=gymnasium))
    return _loaded


def load_greedy_agents():
    """Import the reference's greedy drivers (used unmodified for config 1)."""
    mods = load_reference()
    if "greedy" not in mods:
        p = str(REFERENCE_ROOT / "src" / "agents" / "dqn" / "dqn_evaluation_results")
        if p not in sys.path:
            sys.path.insert(0, p)
        # greedy_agents.py does `from environment.uav_env import UAVEnvironment`;
        # alias the already-imported module so the proxies stay in force.
        pkg = types.ModuleType("environment")
        pkg.__path__ = [str(REFERENCE_ROOT / "src" / "environment")]
        sys.modules.setdefault("environment", pkg)
        sys.modules.setdefault("environment.uav_env", mods["uav_env"])
        sys.modules.setdefault("environment.iot_sensors", mods["iot_sensors"])
        import greedy_agents
        mods["greedy"] = greedy_agents
    return mods["greedy"]


# --------------------------------------------------------------------------
# state snapshots
# --------------------------------------------------------------------------
def snapshot(env) -> dict:
    """Everything SURVEY.md section 8a lists as persistent state, as arrays."""
    n = env.num_sensors
    s = env.sensors
    out = dict(
        pos=np.asarray(env.uav.position, dtype=np.float32).copy(),
        battery=np.float64(env.uav.battery),
        step=np.int32(env.current_step),
        buf=np.array([x.data_buffer for x in s], dtype=np.float64),
        gen=np.array([x.total_data_generated for x in s], dtype=np.float64),
        tx=np.array([x.total_data_transmitted for x in s], dtype=np.float64),
        lost=np.array([x.total_data_lost for x in s], dtype=np.float64),
        ema=np.array([np.nan if x.avg_rssi is None else x.avg_rssi for x in s], dtype=np.float64),
        sf=np.array([x.spreading_factor for x in s], dtype=np.int8),
        visited=np.array([i in env.sensors_visited for i in range(n)], dtype=np.bool_),
        total_reward=np.float64(env.total_reward),
        total_data_collected=np.float64(env.total_data_collected),
        capture_triggers=np.int32(env.capture_effect_triggers),
        boundary_hits=np.int32(env.boundary_hits),
        edge_steps=np.int32(env.edge_steps),
        last_step_bytes=np.float64(env.last_step_bytes_collected),
    )
    return out


INFO_SCALARS = ("max_urgency", "avg_urgency", "high_urgency_sensors", "coverage_percentage",
                "battery_percent", "sensors_collected")

COLLECT_KEYS = ("bytes_collected", "was_new_sensor", "was_empty", "all_sensors_collected",
                "battery_used", "collision_count", "data_loss", "urgency_reduced",
                "sensor_urgency")


def make_env(grid, num_sensors, seed=None, sensor_positions=None, start=None, **kw):
    """Construct the reference env the way `sweep_eval.py:460-463` does."""
    mods = load_reference()
    if seed is not None:
        mods["py_random"].seed(seed)
        np.random.seed(seed)
    kwargs = dict(max_steps=2100, path_loss_exponent=3.8, rssi_threshold=-85.0,
                  sensor_duty_cycle=10.0)
    kwargs.update(kw)
    if sensor_positions is not None:
        kwargs["sensor_positions"] = [tuple(map(float, p)) for p in sensor_positions]
    if start is not None:
        kwargs["uav_start_position"] = tuple(map(float, start))
    return mods["UAVEnvironment"](grid_size=tuple(grid), num_sensors=num_sensors, **kwargs)


def traced_reset(env, seed=None):
    REC.fill = []
    REC.begin(env.num_sensors)
    try:
        obs, info = env.reset(seed=seed)
    finally:
        REC.end()
    st = snapshot(env)
    st["obs"] = np.asarray(obs, dtype=np.float32)
    st["draws"] = REC.draws[:6].copy()
    st["fill"] = np.array(REC.fill, dtype=np.float64)
    st["info"] = info
    return st


def traced_step(env, action):
    REC.begin(env.num_sensors)
    try:
        obs, reward, terminated, truncated, info = env.step(int(action))
    finally:
        REC.end()
    st = snapshot(env)
    st["obs"] = np.asarray(obs, dtype=np.float32)
    st["reward"] = np.float64(reward)
    st["truncated"] = np.bool_(truncated)
    st["terminated"] = np.bool_(terminated)
    st["draws"] = REC.draws[:6].copy()
    st["dwell_draws"] = REC.draws[6].copy()
    ck = REC.collect_kwargs
    for k in COLLECT_KEYS:
        st["c_" + k] = np.float64(ck[k]) if ck is not None else np.float64(np.nan)
    for k in INFO_SCALARS:
        st["i_" + k] = np.float64(info[k])
    st["i_ratios"] = np.array(info["sensor_collection_ratios"], dtype=np.float64)
    return st


def stack_steps(steps: list) -> dict:
    keys = [k for k in steps[0] if k != "info"]
    return {k: np.stack([s[k] for s in steps]) for k in keys}


# --------------------------------------------------------------------------
# DomainRandEnv (src/agents/dqn/dqn.py:177-451)
# --------------------------------------------------------------------------
class _DqnNumpyProxy:
    """`np` inside the exec'd slice of dqn.py: the UNSEEDED `np.random.default_rng()` that
    DomainRandEnv.reset uses for the sensor layout (dqn.py:342) is replaced by a generator seeded from a
    counter, so that fixtures are reproducible; everything else is numpy itself (the global legacy stream
    for grid choice / far start stays the one `np.random.seed` seeded)."""

    class _Random:
        def __init__(self, real, owner):
            self._real, self._owner = real, owner

        def default_rng(self, seed=None):
            self._owner.layout_calls += 1
            return self._real.default_rng(1_000_003 * self._owner.layout_seed + self._owner.layout_calls)

        def __getattr__(self, name):
            return getattr(self._real, name)

    def __init__(self, real):
        self._real = real
        self.layout_seed = 0
        self.layout_calls = 0
        self.random = self._Random(real.random, self)

    def __getattr__(self, name):
        return getattr(self._real, name)


def load_domain_rand_env():
    """Return (DomainRandEnv class, numpy proxy).  dqn.py does not import as shipped (unterminated module
    docstring -> SyntaxError, and import-time mkdir), so the text of lines 9..(GREEDY BENCHMARK marker) is
    exec'd unchanged in a namespace with stub `stable_baselines3` modules; nothing is copied into the repo."""
    mods = load_reference()
    if "DomainRandEnv" in mods:
        return mods["DomainRandEnv"], mods["dqn_np"]
    load_greedy_agents()       # installs the `environment.*` aliases onto the proxied modules
    path = REFERENCE_ROOT / "src" / "agents" / "dqn" / "dqn.py"
    lines = path.read_text().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("import sys"))
    stop = next(i for i, l in enumerate(lines) if "GREEDY BENCHMARK" in l)
    src = "\n".join(lines[start:stop])
    for name in ("stable_baselines3", "stable_baselines3.common", "stable_baselines3.common.vec_env",
                 "stable_baselines3.common.monitor", "stable_baselines3.common.callbacks",
                 "stable_baselines3.common.torch_layers"):
        m = sys.modules.setdefault(name, types.ModuleType(name))
        for attr in ("DQN", "DummyVecEnv", "VecFrameStack", "Monitor", "CheckpointCallback", "BaseCallback",
                     "CallbackList", "EvalCallback", "BaseFeaturesExtractor"):
            if not hasattr(m, attr):
                setattr(m, attr, type(attr, (), {}))
    ns = {"__name__": "dqn_slice", "__file__": str(path)}
    exec(compile(src, str(path), "exec"), ns)
    proxy = _DqnNumpyProxy(np)
    ns["np"] = proxy
    cls = ns["DomainRandEnv"]
    real_sensor = ns["IoTSensor"]

    def recording_sensor(*a, **k):      # DomainRandEnv.reset builds fresh sensors with s0's SF (dqn.py:345-360)
        REC.last_init_sf = int(k.get("spreading_factor", -1))
        return real_sensor(*a, **k)
    ns["IoTSensor"] = recording_sensor
    mods["DomainRandEnv"], mods["dqn_np"], mods["dqn_ns"] = cls, proxy, ns
    return cls, proxy


def dr_traced_reset(env):
    """DomainRandEnv.reset makes TWO observation passes (parent reset on the old sensors, then on the new
    ones, dqn.py:338,367); only the second one's draws belong to the new episode."""
    REC.fill = []
    REC.begin(env.num_sensors)
    REC.allow_overwrite = True
    try:
        obs, info = env.reset()
    finally:
        REC.allow_overwrite = False
        REC.end()
    st = snapshot(env)
    st["obs"] = np.asarray(obs, dtype=np.float32)
    st["draws"] = REC.draws[:6].copy()
    st["grid"] = np.array(env.grid_size, dtype=np.int32)
    st["positions"] = np.array(env.sensor_positions, dtype=np.float64)
    st["start"] = np.asarray(env.uav.start_position, dtype=np.float32).copy()
    st["prev_dist"] = np.float64(env._prev_dist_nearest)
    st["init_sf"] = np.int32(REC.last_init_sf)
    return st


def dr_traced_step(env, action):
    st = traced_step(env, action)
    st["prev_dist"] = np.float64(env._prev_dist_nearest)
    ffc = env._first_full_coverage_step
    st["first_cov"] = np.int32(-1 if ffc is None else ffc)
    return st


# --------------------------------------------------------------------------
# RelationalUAVEnv (src/experiments/relational_policy/env_wrapper.py:110-313)
# --------------------------------------------------------------------------
def load_relational_env():
    mods = load_reference()
    if "RelationalUAVEnv" not in mods:
        load_greedy_agents()       # `environment.*` aliases
        p = str(REFERENCE_ROOT / "src" / "experiments" / "relational_policy")
        if p not in sys.path:
            sys.path.insert(0, p)
        import env_wrapper
        mods["relational"] = env_wrapper
        mods["RelationalUAVEnv"] = env_wrapper.RelationalUAVEnv
    return mods["RelationalUAVEnv"]


def rel_flat(obs: dict) -> np.ndarray:
    """uav(3) + sensors(n_max,5) flattened: the row layout the batched simulator writes in relational mode."""
    return np.concatenate([obs["uav"], obs["sensors"].reshape(-1)]).astype(np.float32)


def rel_traced_reset(env, seed=None):
    REC.fill = []
    REC.begin(env.num_sensors)
    try:
        obs, info = env.reset(seed=seed)
    finally:
        REC.end()
    st = snapshot(env)
    st["obs"] = rel_flat(obs)
    st["mask"] = np.asarray(obs["mask"], dtype=np.float32)
    st["draws"] = REC.draws[:6].copy()
    st["fill"] = np.array(REC.fill, dtype=np.float64)
    st["potential"] = np.float64(env._prev_potential)
    return st


def rel_traced_step(env, action):
    REC.begin(env.num_sensors)
    try:
        obs, reward, terminated, truncated, info = env.step(int(action))
    finally:
        REC.end()
    st = snapshot(env)
    st["obs"] = rel_flat(obs)
    st["reward"] = np.float64(reward)
    st["truncated"] = np.bool_(truncated)
    st["draws"] = REC.draws[:6].copy()
    st["dwell_draws"] = REC.draws[6].copy()
    st["potential"] = np.float64(env._prev_potential)
    st["dwell_steps"] = np.int32(env._dwell_steps)
    for k in ("ndr", "jains", "efficiency"):
        st["m_" + k] = np.float64(info.get(k, np.nan))
    return st


# --------------------------------------------------------------------------
# Sim-to-real evaluation wrappers (src/agents/dqn/dqn_evaluation_results/sim_to_real_analysis.py:128-245):
# FixedLayoutEnv / PaddedFixedEnv / GPSNoisyEnv / WindDriftEnv / LatencyEnv
# --------------------------------------------------------------------------
class _ExtraNumpyProxy:
    """`np` inside the exec'd slice: the wrappers' OWN draws (GPS noise, wind angle) come from the global legacy
    stream like the reference's, and are also appended to `extra` so fixtures can carry them."""

    class _Random:
        def __init__(self, real, owner):
            self._real, self._owner = real, owner

        def normal(self, *a, **k):
            v = self._real.normal(*a, **k)
            self._owner.extra.append(float(v))
            return v

        def uniform(self, *a, **k):
            v = self._real.uniform(*a, **k)
            self._owner.extra.append(float(v))
            return v

        def __getattr__(self, name):
            return getattr(self._real, name)

    def __init__(self, real):
        self._real = real
        self.extra = []
        self.random = self._Random(real.random, self)

    def __getattr__(self, name):
        return getattr(self._real, name)


def load_sim_to_real_envs():
    """Return (namespace with the wrapper classes, numpy proxy).  The module itself needs SB3, pandas, a trained
    model and writes result directories at import; only the text of its ENVIRONMENT WRAPPERS section is exec'd,
    unchanged, against the proxied reference env."""
    mods = load_reference()
    if "s2r_ns" in mods:
        return mods["s2r_ns"], mods["s2r_np"]
    load_greedy_agents()
    path = REFERENCE_ROOT / "src" / "agents" / "dqn" / "dqn_evaluation_results" / "sim_to_real_analysis.py"
    lines = path.read_text().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("class FixedLayoutEnv"))
    stop = next(i for i, l in enumerate(lines) if "EPISODE RUNNERS" in l)
    src = "\n" * start + "\n".join(lines[start:stop])          # keep the reference's line numbers in tracebacks
    proxy = _ExtraNumpyProxy(np)
    import gymnasium
    ns = {"__name__": "sim_to_real_slice", "__file__": str(path), "np": proxy, "gymnasium": gymnasium,
          "UAVEnvironment": mods["UAVEnvironment"], "MAX_SEN_LIMIT": 50}
    exec(compile(src, str(path), "exec"), ns)
    mods["s2r_ns"], mods["s2r_np"] = ns, proxy
    return ns, proxy


def s2r_traced_reset(env, npx, seed=None):
    npx.extra.clear()
    st = traced_reset(env, seed=seed)
    st["extra"] = np.array(npx.extra + [np.nan] * (4 - len(npx.extra)), dtype=np.float64)
    return st


def s2r_traced_step(env, npx, action):
    npx.extra.clear()
    st = traced_step(env, action)
    st["extra"] = np.array(npx.extra + [np.nan] * (4 - len(npx.extra)), dtype=np.float64)
    return st


# --------------------------------------------------------------------------
# PerturbedDQNEnv of the sim-to-real SWEEP (src/agents/dqn/dqn_evaluation_results/sim_to_real_sweep.py:107-139) on top
# of AnalysisUAVEnv (compare_agents.py:62-147): per-sensor shadowing sigma, GPS noise applied to the position INSIDE
# the observation pass (so the pass's ADR update and range check see the noisy position), and a probabilistic extra
# whole step for wind.
# --------------------------------------------------------------------------
class _SweepNumpyProxy(_ExtraNumpyProxy):
    """`np` inside the exec'd `_perturb_positions_dqn`: records the wrapper's own draws -- np.random.random() (the wind
    gate) and np.random.normal(0, sigma, size=2) (the GPS noise pair) -- in call order."""

    class _Random(_ExtraNumpyProxy._Random):
        def normal(self, *a, **k):
            v = self._real.normal(*a, **k)
            self._owner.extra.extend(float(x) for x in np.atleast_1d(v))
            return v

        def random(self, *a, **k):
            v = self._real.random(*a, **k)
            self._owner.extra.append(float(v))
            return v

    def __init__(self, real):
        super().__init__(real)
        self.random = self._Random(real.random, self)


def load_perturbed_dqn_env():
    """Returns (make(gps_sigma, shadow_sigma, wind_drift, ple) -> PerturbedDQNEnv class, numpy proxy).  The class text
    of AnalysisUAVEnv and of `_perturb_positions_dqn` is exec'd unchanged; the base handed to the factory is a TRACING
    subclass of AnalysisUAVEnv whose step / reset record each INNER env step (draw table, state, padded observation)
    in `env.inner` -- PerturbedDQNEnv.step may run two of them."""
    mods = load_reference()
    if "pert_make" in mods:
        return mods["pert_make"], mods["pert_np"]
    import gymnasium
    base = REFERENCE_ROOT / "src" / "agents" / "dqn" / "dqn_evaluation_results"
    lines = (base / "compare_agents.py").read_text().splitlines()
    a = next(i for i, l in enumerate(lines) if l.startswith("class AnalysisUAVEnv"))
    z = next(i for i, l in enumerate(lines) if "TRAJECTORY TRACKING" in l)
    ns_a = {"__name__": "compare_agents_slice", "np": np, "gymnasium": gymnasium, "UAVEnvironment": mods["UAVEnvironment"]}
    exec(compile("\n" * a + "\n".join(lines[a:z]), str(base / "compare_agents.py"), "exec"), ns_a)
    Analysis = ns_a["AnalysisUAVEnv"]
    proxy = _SweepNumpyProxy(np)

    class TracingAnalysisEnv(Analysis):
        def __init__(self, **kw):
            self.inner = []
            super().__init__(**kw)

        def _trace(self, out, reward=None, truncated=None, pre=()):
            st = snapshot(self)
            st["obs"] = np.asarray(out, dtype=np.float32)
            st["draws"] = REC.draws[:6].copy()
            ex = list(proxy.extra)                       # the GPS pair drawn inside this inner step's observation
            st["extra"] = np.array(ex + [np.nan] * (2 - len(ex)), dtype=np.float64)
            st["pre"] = np.array(list(pre) + [np.nan] * (1 - len(pre)), dtype=np.float64)   # the wind gate's uniform
            proxy.extra.clear()
            if reward is not None:
                st["reward"] = np.float64(reward)
                st["truncated"] = np.bool_(truncated)
            return st

        def reset(self, **kw):
            proxy.extra.clear()
            REC.fill = []
            REC.begin(self.num_sensors)
            try:
                obs, info = super().reset(**kw)
            finally:
                REC.end()
            st = self._trace(obs)
            st["fill"] = np.array(REC.fill, dtype=np.float64)
            self.inner_reset = st
            self.inner = []
            return obs, info

        def step(self, action):
            pre = list(proxy.extra)
            proxy.extra.clear()
            REC.begin(self.num_sensors)
            try:
                out = super().step(action)
            finally:
                REC.end()
            st = self._trace(out[0], out[1], out[3], pre)
            st["info"] = out[4]
            self.inner.append(st)
            return out

    path = base / "sim_to_real_sweep.py"
    lines = path.read_text().splitlines()
    a = next(i for i, l in enumerate(lines) if l.startswith("def _perturb_positions_dqn"))
    z = next(i for i, l in enumerate(lines) if l.startswith("def _perturb_positions_rel"))
    ns = {"__name__": "sim_to_real_sweep_slice", "np": proxy}
    exec(compile("\n" * a + "\n".join(lines[a:z]), str(path), "exec"), ns)

    def make(gps_sigma, shadow_sigma, wind_drift, ple=3.8):
        return ns["_perturb_positions_dqn"](TracingAnalysisEnv, gps_sigma, shadow_sigma, wind_drift, ple)
    mods["pert_make"], mods["pert_np"] = make, proxy
    return make, proxy
