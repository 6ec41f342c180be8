"""ctypes binding of libuavsim.so -- the stub a maintainer of the reference would add
(INTEGRATION.md).  Mirrors include/uavsim.h field for field; no torch types cross this
boundary, only raw pointers (`tensor.data_ptr()`), sizes and POD structs.

There is NO CPU fallback: if the library cannot be loaded the import of the package fails.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os

# UAVSIM_LIB: developer override used for A/B runs of kernel variants (tools/build_variants.py); the product
# library is libuavsim.so next to this file.
LIB_PATH = Path(os.environ["UAVSIM_LIB"]).resolve() if os.environ.get("UAVSIM_LIB") else PKG / "libuavsim.so"

MAX_SENSORS = 64
MAX_GRIDS = 8
NUM_SLOTS = 6
NUM_ACTIONS = 5
STAT_COUNT = 20
STAT_SUMS = STAT_COUNT + 1
DETAIL_COUNT = 16
INFO_COUNT = 18
# uavsim.h UAVSIM_INFO_*: the reference's _get_info() keys (uav_env.py:676-700) in row order; "uav_position" is
# (x, y) and "sensor_collection_ratios" travels as its own [P] row
INFO_NAMES = ("x", "y", "battery", "battery_percent", "sensors_collected", "current_step", "total_reward",
              "total_data_collected", "coverage_percentage", "is_alive", "max_urgency", "avg_urgency",
              "high_urgency_sensors", "capture_effect_triggers", "boundary_hits", "edge_steps",
              "last_step_bytes_collected", "num_sensors")

OK, ERR_INVALID_ARG, ERR_CUDA, ERR_STATE, ERR_INVALID_ACTION, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
F64, F32 = 0, 1
RNG_REPLAY, RNG_PHILOX = 0, 1
RESET_FIXED, RESET_DOMAIN_RAND = 0, 1
WRAP_NONE, WRAP_RELATIONAL = 0, 1
SLOT_A, SLOT_B, SLOT_U, SLOT_C, SLOT_D, SLOT_E = range(6)
FIELD_ENV_RECORD, FIELD_SENSOR_STATE, FIELD_SENSOR_POS, FIELD_SENSOR_GEN, FIELD_SENSOR_SF = range(5)

STAT_NAMES = ("return", "length", "total_generated", "total_collected", "total_lost",
              "battery_remaining", "ndr", "fairness_std", "jains_index", "grid_w", "grid_h",
              "num_sensors", "data_efficiency", "bytes_per_wh", "time_to_coverage",
              "capture_triggers", "boundary_hits", "edge_steps", "collisions", "base_return")
DETAIL_NAMES = ("collisions", "bytes", "urgency_reduced", "mean_urgency", "was_new", "was_empty",
                "all_collected", "data_loss", "attempt_lo", "attempt_hi", "winner_lo", "winner_hi",
                "base_reward", "truncated")

_d, _i32 = C.c_double, C.c_int32


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("num_envs", _i32), ("max_sensors", _i32), ("obs_slots", _i32),
        ("features_per_sensor", _i32), ("frame_stack", _i32), ("precision", _i32), ("rng_mode", _i32),
        ("reset_mode", _i32), ("auto_reset", _i32), ("seed", C.c_uint64), ("env_id_offset", C.c_int64),
        ("grid_w", _i32), ("grid_h", _i32), ("num_sensors", _i32), ("max_steps", _i32),
        ("data_generation_rate", _d), ("max_buffer_size", _d), ("rssi_threshold", _d), ("duty_cycle", _d),
        ("max_battery", _d), ("collection_duration", _d), ("start_x", _d), ("start_y", _d),
        ("tx_power_dbm", _d), ("noise_floor_dbm", _d), ("altitude_m", _d), ("adr_lambda", _d),
        ("shadowing_std_db", _d), ("capture_threshold_db", _d), ("power_move_w", _d), ("power_hover_w", _d),
        ("init_sf", _i32), ("wrapper", _i32),
        ("reward_per_byte", _d), ("reward_new_sensor", _d), ("reward_completion", _d),
        ("reward_urgency_reduction", _d), ("reward_movement", _d),
        ("penalty_revisit", _d), ("penalty_boundary", _d), ("penalty_collision", _d), ("penalty_battery", _d),
        ("penalty_hover", _d), ("penalty_step", _d), ("penalty_data_loss", _d), ("penalty_starvation", _d),
        ("penalty_unvisited", _d), ("penalty_starved", _d), ("starvation_cr_threshold", _d),
        ("num_grids", _i32), ("grids", (_i32 * 2) * MAX_GRIDS),
        ("min_start_dist", _d), ("max_start_tries", _i32), ("gps_in_observation", _i32), ("prox_eta", _d), ("jain_bonus", _d),
        ("rel_gamma", _d), ("rel_dwell_bonus", _d), ("rel_reward_scale", _d), ("rel_dwell_steps", _i32),
        ("rel_min_cluster", _i32),
        ("gps_sigma", _d), ("wind_drift", _d), ("latency_steps", _i32), ("reserved2", _i32),
    ]


class Dims(C.Structure):
    _fields_ = [("obs_dim", _i32), ("obs_row_floats", _i32), ("frame_rows", _i32), ("real_bytes", _i32),
                ("arena_bytes", C.c_int64), ("replay_step_doubles", C.c_int64)]


class Outputs(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p), ("terminal_obs", C.c_void_p),
                ("episode_stats", C.c_void_p), ("action_mask", C.c_void_p), ("step_detail", C.c_void_p),
                ("draw_export", C.c_void_p), ("terminal_info", C.c_void_p), ("terminal_ratios", C.c_void_p)]


class Scenario(C.Structure):
    _fields_ = [("sensor_xy", C.c_void_p), ("num_sensors", C.c_void_p), ("grid_wh", C.c_void_p),
                ("start_xy", C.c_void_p)]


class Replay(C.Structure):
    _fields_ = [("step_draws", C.c_void_p), ("reset_draws", C.c_void_p), ("reset_fill", C.c_void_p),
                ("reset_init_sf", C.c_void_p), ("reset_scenario", Scenario), ("dwell_draws", C.c_void_p),
                ("step_env_draws", C.c_void_p), ("reset_env_draws", C.c_void_p)]


class FieldView(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("elem_bytes", _i32), ("is_float", _i32), ("ndim", _i32),
                ("shape", C.c_int64 * 4), ("stride_bytes", C.c_int64 * 4)]


class EnvRecord(C.Structure):
    """uavsim_env_record (128 bytes)."""
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("start_x", C.c_float), ("start_y", C.c_float),
                ("battery", _d), ("total_reward", _d), ("total_data_collected", _d), ("episode_return", _d),
                ("last_step_bytes", C.c_float), ("prev_dist_nearest", C.c_float), ("elapsed_s", C.c_float),
                ("step", _i32), ("grid_w", _i32), ("grid_h", _i32), ("num_sensors", _i32),
                ("visited_lo", C.c_uint32), ("visited_hi", C.c_uint32), ("capture_triggers", _i32),
                ("boundary_hits", _i32), ("edge_steps", _i32), ("collisions", _i32),
                ("first_full_coverage_step", _i32), ("episode", C.c_uint32), ("delivered", _i32),
                ("init_sf", _i32), ("dwell_steps", _i32), ("wind_angle", C.c_double)]


assert C.sizeof(EnvRecord) == 128

# numpy view of the record array (same offsets)
ENV_RECORD_DTYPE = [
    ("x", "<f4"), ("y", "<f4"), ("start_x", "<f4"), ("start_y", "<f4"), ("battery", "<f8"),
    ("total_reward", "<f8"), ("total_data_collected", "<f8"), ("episode_return", "<f8"),
    ("last_step_bytes", "<f4"), ("prev_dist_nearest", "<f4"), ("elapsed_s", "<f4"), ("step", "<i4"),
    ("grid_w", "<i4"), ("grid_h", "<i4"), ("num_sensors", "<i4"), ("visited_lo", "<u4"), ("visited_hi", "<u4"),
    ("capture_triggers", "<i4"), ("boundary_hits", "<i4"), ("edge_steps", "<i4"), ("collisions", "<i4"),
    ("first_full_coverage_step", "<i4"), ("episode", "<u4"), ("delivered", "<i4"), ("init_sf", "<i4"),
    ("dwell_steps", "<i4"), ("wind_angle", "<f8"),
]

# Every symbol include/uavsim.h declares (tests/test_abi.py checks the header against this list).
SYMBOLS = (
    "uavsim_abi_version", "uavsim_default_config", "uavsim_query_dims", "uavsim_last_error",
    "uavsim_create", "uavsim_destroy", "uavsim_bind_outputs", "uavsim_get_field", "uavsim_set_scenario",
    "uavsim_generate_layouts", "uavsim_set_replay", "uavsim_set_stage_grids", "uavsim_reset", "uavsim_step",
    "uavsim_step_host", "uavsim_poll_error", "uavsim_stat_sums", "uavsim_launch_count", "uavsim_frame_slot",
    "uavsim_policy_nearest_greedy", "uavsim_policy_max_throughput_greedy",
    "uavsim_get_info", "uavsim_set_seed", "uavsim_set_frame_slot", "uavsim_set_host_chunks",
    "uavsim_step_ex", "uavsim_set_env_params",
)
STEP_EXTRA = 1


class UavsimError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("uavsim error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load libuavsim.so (building it with nvcc if the in-tree binary is missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    try:
        path = LIB_PATH if os.environ.get("UAVSIM_LIB") else _build.build()
    except Exception as exc:  # no nvcc on this box: use the shipped binary if there is one
        if not LIB_PATH.exists():
            raise ImportError("libuavsim.so is missing and could not be built: %s" % exc) from exc
        path = LIB_PATH
    lib = C.CDLL(str(path))
    vp = C.c_void_p
    lib.uavsim_abi_version.restype = C.c_int
    lib.uavsim_last_error.restype = C.c_char_p
    lib.uavsim_default_config.argtypes = [C.POINTER(Config)]
    lib.uavsim_query_dims.argtypes = [C.POINTER(Config), C.POINTER(Dims)]
    lib.uavsim_create.argtypes = [C.POINTER(Config), vp, C.c_size_t, vp, C.POINTER(vp)]
    lib.uavsim_destroy.argtypes = [vp]
    lib.uavsim_bind_outputs.argtypes = [vp, C.POINTER(Outputs), vp]
    lib.uavsim_get_field.argtypes = [vp, C.c_int, C.POINTER(FieldView)]
    lib.uavsim_set_scenario.argtypes = [vp, C.POINTER(Scenario), vp]
    lib.uavsim_generate_layouts.argtypes = [vp, vp]
    lib.uavsim_set_replay.argtypes = [vp, C.POINTER(Replay)]
    lib.uavsim_set_stage_grids.argtypes = [vp, _i32, C.POINTER(_i32)]
    lib.uavsim_reset.argtypes = [vp, vp, vp]
    lib.uavsim_step.argtypes = [vp, vp, vp]
    lib.uavsim_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.uavsim_poll_error.argtypes = [vp, vp, C.POINTER(C.c_int64)]
    lib.uavsim_stat_sums.argtypes = [vp, C.POINTER(vp)]
    lib.uavsim_launch_count.argtypes = [vp]
    lib.uavsim_launch_count.restype = C.c_int64
    lib.uavsim_policy_nearest_greedy.argtypes = [vp, vp, vp, vp]
    lib.uavsim_policy_max_throughput_greedy.argtypes = [vp, vp, vp, vp]
    lib.uavsim_frame_slot.argtypes = [vp]
    lib.uavsim_frame_slot.restype = C.c_int
    lib.uavsim_get_info.argtypes = [vp, vp, vp, vp]
    lib.uavsim_set_seed.argtypes = [vp, C.c_uint64]
    lib.uavsim_set_frame_slot.argtypes = [vp, _i32]
    lib.uavsim_set_host_chunks.argtypes = [vp, _i32]
    lib.uavsim_step_ex.argtypes = [vp, vp, vp, C.c_uint32, vp]
    lib.uavsim_set_env_params.argtypes = [vp, vp, vp]
    if lib.uavsim_abi_version() != 2:
        raise ImportError("libuavsim.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc != OK:
        msg = load().uavsim_last_error().decode()
        if rc == ERR_INVALID_ACTION:
            raise ValueError(msg)                  # uav_env.py:468 raises ValueError
        raise UavsimError(rc, msg)


def default_config() -> Config:
    c = Config()
    check(load().uavsim_default_config(C.byref(c)))
    return c


def query_dims(cfg: Config) -> Dims:
    d = Dims()
    check(load().uavsim_query_dims(C.byref(cfg), C.byref(d)))
    return d
