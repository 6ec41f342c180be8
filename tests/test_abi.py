"""CPU checks of the drop-in boundary: libuavsim.so builds/loads without a GPU, exports every symbol
include/uavsim.h declares, and its host-only entry points (defaults, validation, layout arithmetic)
behave.  No compute calls are made here."""
import ctypes as C
import re
from pathlib import Path

import pytest

from __graft_entry__ import ROOT, import_package


@pytest.fixture(scope="module")
def N():
    import_package()
    import sys
    return sys.modules["uavsim_b200._native"]


def test_library_exports_every_declared_symbol(N):
    header = (ROOT / "include" / "uavsim.h").read_text()
    declared = set(re.findall(r"^\s*(?:const\s+char\s*\*|int64_t|int)\s+(uavsim_\w+)\s*\(", header, re.M))
    assert declared == set(N.SYMBOLS), declared ^ set(N.SYMBOLS)
    lib = N.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.uavsim_abi_version() == 2


def test_struct_sizes_match_header(N):
    lib = N.load()
    cfg = N.default_config()
    assert cfg.struct_size == C.sizeof(N.Config)         # the library wrote ITS sizeof into the struct
    assert C.sizeof(N.EnvRecord) == 128


def test_defaults_are_the_reference_training_path_constants(N):
    c = N.default_config()
    assert (c.max_steps, c.max_battery, c.rssi_threshold, c.duty_cycle) == (2100, 274.0, -85.0, 10.0)   # dqn.py:1068-1075
    assert abs(c.data_generation_rate - 2.2) < 1e-15 and c.max_buffer_size == 1000.0                    # uav_env.py:271-272
    assert (c.penalty_data_loss, c.reward_urgency_reduction, c.penalty_battery, c.reward_movement) == (-1.0, 20.0, -0.5, 10.0)
    assert (c.reward_per_byte, c.reward_new_sensor, c.penalty_unvisited, c.penalty_starved) == (100.0, 5000.0, -5000.0, -1000.0)
    assert (c.power_move_w, c.power_hover_w, c.init_sf, c.capture_threshold_db) == (500.0, 700.0, 12, 6.0)
    assert (c.min_start_dist, c.max_start_tries, c.prox_eta, c.jain_bonus) == (50.0, 200, 2.0, 0.5)


def test_query_dims_layout_arithmetic(N):
    c = N.default_config()
    c.num_envs, c.max_sensors, c.num_sensors, c.obs_slots = 4096, 50, 50, 50
    d = N.query_dims(c)
    assert d.obs_dim == 153 and d.obs_row_floats == 153 and d.real_bytes == 4
    assert d.replay_step_doubles == 300
    per_env = 128 + 50 * (16 + 8 + 4 + 1) + 4
    assert per_env * 4096 <= d.arena_bytes <= per_env * 4096 + 16 * 256
    c.precision = N.F64
    assert N.query_dims(c).real_bytes == 8
    c.features_per_sensor = 5
    assert N.query_dims(c).obs_dim == 253                  # transformer/relational padded width


@pytest.mark.parametrize("field,value", [
    ("num_envs", 0), ("max_sensors", 65), ("num_sensors", 0), ("features_per_sensor", 4), ("precision", 7),
    ("rng_mode", 9), ("grid_w", 0), ("max_steps", 0), ("init_sf", 6), ("data_generation_rate", 0.0),
    ("struct_size", 12),
])
def test_validation_rejects_bad_configs(N, field, value):
    c = N.default_config()
    setattr(c, field, value)
    with pytest.raises(N.UavsimError) as ei:
        N.query_dims(c)
    assert ei.value.code in (N.ERR_INVALID_ARG, N.ERR_UNSUPPORTED)


def test_null_arguments_are_errors_not_crashes(N):
    lib = N.load()
    assert lib.uavsim_query_dims(None, None) == N.ERR_INVALID_ARG
    assert b"NULL" in lib.uavsim_last_error()
    assert lib.uavsim_step(None, None, None) == N.ERR_INVALID_ARG
    assert lib.uavsim_reset(None, None, None) == N.ERR_INVALID_ARG
    assert lib.uavsim_destroy(None) == N.OK
    assert lib.uavsim_launch_count(None) == 0


def test_product_package_never_imports_the_oracle():
    from __graft_entry__ import PKG_DIR
    for f in list(PKG_DIR.glob("*.py")) + list((PKG_DIR / "csrc").glob("*")):
        txt = f.read_text()
        assert "oracle" not in txt.lower(), f


def test_header_is_plain_c_and_links_from_c(N, tmp_path):
    """The boundary is a C ABI: include/uavsim.h must compile as C99 and a C program must link against the shared
    library and call its host-only entry points (no GPU needed)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi_check.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "uavsim.h"
int main(void) {
    uavsim_config cfg;
    uavsim_dims d;
    if (uavsim_abi_version() != UAVSIM_ABI_VERSION) return 2;
    if (uavsim_default_config(&cfg) != UAVSIM_OK || cfg.struct_size != sizeof(uavsim_config)) return 3;
    cfg.num_envs = 1024; cfg.max_sensors = cfg.num_sensors = cfg.obs_slots = 50;
    if (uavsim_query_dims(&cfg, &d) != UAVSIM_OK || d.obs_dim != 153) return 4;
    cfg.gps_sigma = -1.0;
    if (uavsim_query_dims(&cfg, &d) == UAVSIM_OK || strlen(uavsim_last_error()) == 0) return 5;
    if (sizeof(uavsim_env_record) != 128) return 6;
    printf("ok %d %lld\n", d.obs_dim, (long long)d.arena_bytes);
    return 0;
}
''')
    exe = tmp_path / "abi_check"
    lib_dir = N.LIB_PATH.parent
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
                        "-L", str(lib_dir), "-luavsim", "-Wl,-rpath," + str(lib_dir)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok 153"), (r.returncode, r.stdout, r.stderr)
