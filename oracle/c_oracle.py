"""ctypes front-end of the C oracle (oracle/uavsim_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of uavsim_oracle.c.  Builds
oracle/_build/liboracle.so on demand with oracle/Makefile (gcc only).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "liboracle.so"
MAXS = 64
SLOTS = 6
DETAIL_COUNT = 21   # 16 shared with the CUDA detail row + 5 oracle-only (margins, reward scale)

_PARAM_DOUBLES = (
    "data_generation_rate", "max_buffer_size", "rssi_threshold", "duty_cycle", "max_battery",
    "collection_duration", "tx_power_dbm", "noise_floor_dbm", "altitude", "adr_lambda",
    "shadowing_std_db", "capture_threshold_db", "power_move", "power_hover", "reward_per_byte",
    "reward_new_sensor", "reward_completion", "reward_urgency_reduction", "reward_movement",
    "penalty_revisit", "penalty_boundary", "penalty_collision", "penalty_battery", "penalty_hover",
    "penalty_step", "penalty_data_loss", "penalty_starvation", "penalty_unvisited",
    "penalty_starved", "starvation_cr_threshold", "prox_eta", "jain_bonus", "rel_gamma", "rel_dwell_bonus",
    "rel_reward_scale")
_PARAM_INTS = ("max_steps", "init_sf", "features_per_sensor", "obs_slots", "domain_rand", "relational",
               "rel_dwell_steps", "rel_min_cluster")


class OrcParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in _PARAM_DOUBLES] + [(k, C.c_int32) for k in _PARAM_INTS]


ENV_DTYPE = np.dtype([
    ("W", "i4"), ("H", "i4"), ("N", "i4"), ("step", "i4"),
    ("x", "f4"), ("y", "f4"), ("start_x", "f4"), ("start_y", "f4"),
    ("battery", "f8"), ("total_reward", "f8"), ("total_data_collected", "f8"),
    ("episode_return", "f8"), ("last_step_bytes", "f8"), ("prev_dist_nearest", "f8"),
    ("capture_triggers", "i4"), ("boundary_hits", "i4"), ("edge_steps", "i4"), ("collisions", "i4"),
    ("first_full_coverage_step", "i4"), ("delivered", "i4"), ("dwell_steps", "i4"), ("pad0", "i4"),
    ("spx", "f4", (MAXS,)), ("spy", "f4", (MAXS,)),
    ("buf", "f8", (MAXS,)), ("gen", "f8", (MAXS,)), ("tx", "f8", (MAXS,)), ("lost", "f8", (MAXS,)),
    ("ema", "f8", (MAXS,)), ("cur", "f8", (MAXS,)),
    ("sf", "i1", (MAXS,)), ("visited", "u1", (MAXS,)),
], align=True)


def build(force=False):
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < (HERE / "uavsim_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB_PATH))
        _lib.orc_sizeof_env.restype = C.c_size_t
        _lib.orc_sizeof_params.restype = C.c_size_t
        assert _lib.orc_sizeof_env() == ENV_DTYPE.itemsize, (_lib.orc_sizeof_env(), ENV_DTYPE.itemsize)
        assert _lib.orc_sizeof_params() == C.sizeof(OrcParams)
        _lib.orc_step_batch.restype = C.c_int
    return _lib


def make_params(params: dict, obs_slots: int, domain_rand=False, prox_eta=2.0, jain_bonus=0.5,
                relational=False) -> OrcParams:
    """`params` uses the key names of uavsim_oracle.default_params()."""
    p = OrcParams()
    for k in _PARAM_DOUBLES:
        if k in params:
            setattr(p, k, float(params[k]))
    p.prox_eta, p.jain_bonus = prox_eta, jain_bonus
    p.max_steps = int(params["max_steps"])
    p.init_sf = int(params["init_sf"])
    p.features_per_sensor = 5 if params.get("include_sensor_positions") else 3
    p.obs_slots = int(obs_slots)
    p.domain_rand = int(bool(domain_rand))
    # RelationalUAVEnv constants (env_wrapper.py:47-62); the wrapper forces include_sensor_positions=True
    p.relational = int(bool(relational))
    p.rel_gamma, p.rel_dwell_bonus, p.rel_reward_scale = 0.99, 25.0, 1e-5
    p.rel_dwell_steps, p.rel_min_cluster = 10, 2
    if relational:
        p.features_per_sensor = 5
    return p


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


class COracleBatch:
    """E independent envs stepped by the C oracle.  Draw tables are [E, 6, P] float64."""

    def __init__(self, params: dict, E: int, P: int, obs_slots=None, domain_rand=False, threads=1, relational=False):
        self.lib = lib()
        self.E, self.P = int(E), int(P)
        self.obs_slots = int(obs_slots if obs_slots is not None else P)
        self.p = make_params(params, self.obs_slots, domain_rand, relational=relational)
        self.envs = np.zeros(self.E, dtype=ENV_DTYPE)
        self.obs_dim = 3 + self.p.features_per_sensor * self.obs_slots
        self.obs = np.zeros((self.E, self.obs_dim), dtype=np.float32)
        self.reward = np.zeros(self.E, dtype=np.float64)
        self.done = np.zeros(self.E, dtype=np.uint8)
        self.mask = np.zeros(self.E, dtype=np.uint8)
        self.detail = np.zeros((self.E, DETAIL_COUNT), dtype=np.float64)
        self.threads = max(1, int(threads))
        self._pool = ThreadPoolExecutor(self.threads) if self.threads > 1 else None

    def set_scenario(self, grid_wh, num_sensors, sensor_xy, start_xy):
        """grid_wh [E,2] int, num_sensors [E] int, sensor_xy [E,P,2] float, start_xy [E,2] float."""
        e = self.envs
        e["W"], e["H"] = np.asarray(grid_wh)[:, 0], np.asarray(grid_wh)[:, 1]
        e["N"] = num_sensors
        xy = np.asarray(sensor_xy, dtype=np.float32)
        e["spx"][:, : self.P] = xy[:, :, 0]
        e["spy"][:, : self.P] = xy[:, :, 1]
        st = np.asarray(start_xy, dtype=np.float32)
        e["start_x"], e["start_y"] = st[:, 0], st[:, 1]

    def _ranges(self):
        step = -(-self.E // self.threads)
        return [(a, min(a + step, self.E)) for a in range(0, self.E, step)]

    def reset(self, fill, draws, init_sf=None, env_ids=None):
        """fill [E,P] (or None in domain-rand mode), draws [E,6,P].  env_ids: subset to reset."""
        draws = np.ascontiguousarray(draws, dtype=np.float64)
        fill = None if fill is None else np.ascontiguousarray(fill, dtype=np.float64)
        isf = None if init_sf is None else np.ascontiguousarray(init_sf, dtype=np.int32)

        def run(a, b):
            self.lib.orc_reset_batch(C.byref(self.p), C.c_void_p(self.envs.ctypes.data), a, b,
                                     _ptr(fill, C.c_double), _ptr(draws, C.c_double), self.P,
                                     _ptr(isf, C.c_int32), _ptr(self.obs, C.c_float), _ptr(self.mask, C.c_uint8),
                                     _ptr(self.detail, C.c_double))
        if env_ids is not None:
            for e in np.asarray(env_ids).tolist():
                run(e, e + 1)
        elif self._pool:
            list(self._pool.map(lambda r: run(*r), self._ranges()))
        else:
            run(0, self.E)
        return self.obs

    def step(self, actions, draws, want_detail=True, dwell=None):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        draws = np.ascontiguousarray(draws, dtype=np.float64)
        dwell = None if dwell is None else np.ascontiguousarray(dwell, dtype=np.float64)

        def run(a, b):
            return self.lib.orc_step_batch(C.byref(self.p), C.c_void_p(self.envs.ctypes.data), a, b,
                                           _ptr(actions, C.c_int32), _ptr(draws, C.c_double), _ptr(dwell, C.c_double), self.P,
                                           _ptr(self.obs, C.c_float), _ptr(self.reward, C.c_double),
                                           _ptr(self.done, C.c_uint8), _ptr(self.mask, C.c_uint8),
                                           _ptr(self.detail, C.c_double) if want_detail else None)
        if self._pool:
            bad = sum(self._pool.map(lambda r: run(*r), self._ranges()))
        else:
            bad = run(0, self.E)
        if bad:
            raise ValueError("Invalid action")      # uav_env.py:468
        return self.obs, self.reward, self.done
