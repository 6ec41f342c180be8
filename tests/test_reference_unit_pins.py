"""The reference's own unit tests that still pass and touch the path,
restated against the oracles so the same facts pin them:
  /root/reference/tests/test_iot_sensors.py:80-127  (buffer fill + overflow cap)
  /root/reference/tests/test_iot_sensors.py:500-519 (LoRa data-rate table)
  /root/reference/tests/test_uav.py:76-134          (moves)      :140-185 (bounds)
  /root/reference/tests/test_uav.py:191-261         (energy, with this env's 500/700 W constants)
  /root/reference/tests/test_uav.py:267-297         (hover)      :355-409 (reset)
The stale tests of that suite (default SF 7, power_move 600, ...) contradict the executed code and are
not used."""
import numpy as np
import pytest

from golden_util import O  # puts oracle/ on sys.path
import c_oracle


def _env(grid=(10, 10), positions=((0.0, 0.0),), start=(5.0, 5.0), **over):
    p = O.default_params(**over)
    table = np.zeros((6, len(positions)))
    table[2] = 0.99                       # uniforms: nobody transmits
    d = O.SlotDraws()
    d.load(table)
    e = O.OracleEnv(grid, positions, start, p, d)
    e.reset([0.0] * len(positions))
    return e


def _c_env(grid=(10, 10), positions=((0.0, 0.0),), start=(5.0, 5.0), **over):
    p = O.default_params(**over)
    n = len(positions)
    co = c_oracle.COracleBatch(p, E=1, P=n)
    co.set_scenario([grid], [n], np.asarray(positions, dtype=np.float32)[None], [start])
    table = np.zeros((1, 6, n))
    table[:, 2] = 0.99
    co.reset(np.zeros((1, n)), table)
    return co, table


@pytest.mark.parametrize("action,expect", [(0, (5.0, 6.0)), (1, (5.0, 4.0)), (2, (4.0, 5.0)), (3, (6.0, 5.0))])
def test_moves_one_cell(action, expect):
    e = _env()
    e.step(action)
    assert tuple(e.pos) == expect
    co, t = _c_env()
    co.step([action], t)
    assert (co.envs["x"][0], co.envs["y"][0]) == expect


def test_sequential_movements():
    e = _env()
    for a in (0, 3, 3, 1):
        e.step(a)
    assert tuple(e.pos) == (7.0, 5.0)


@pytest.mark.parametrize("start,action", [((0.0, 5.0), 2), ((9.0, 5.0), 3), ((5.0, 0.0), 1), ((5.0, 9.0), 0)])
def test_boundary_blocks_move_and_costs_half_power(start, action):
    e = _env(start=start)
    b0 = e.battery
    e.step(action)
    assert tuple(e.pos) == start and e.boundary_hits == 1
    assert e.battery == b0 - (500.0 * 0.5 * 1) / 3600
    co, t = _c_env(start=start)
    co.step([action], t)
    assert (co.envs["x"][0], co.envs["y"][0]) == start and co.envs["boundary_hits"][0] == 1
    assert co.envs["battery"][0] == e.battery


def test_energy_move_and_hover():
    e = _env()
    b0 = e.battery
    e.step(0)
    assert e.battery == b0 - 500.0 / 3600
    b1 = e.battery
    e.step(4)
    assert e.battery == b1 - 700.0 / 3600
    e2 = _env(collection_duration=5.0)
    b = e2.battery
    e2.step(4)
    assert abs((b - e2.battery) - 0.9722222) < 1e-6          # 700 W * 5 s


def test_is_alive_floor_truncates_at_two_percent():
    e = _env(max_battery=1.0)
    steps = 0
    trunc = False
    while not trunc:
        _, _, _, trunc = e.step(4)
        steps += 1
    assert e.battery <= 0.02 and e.battery + 700.0 / 3600 > 0.02
    assert steps == int(np.ceil((1.0 - 0.02) / (700.0 / 3600)))


def test_data_generation_and_overflow_cap():
    e = _env(data_generation_rate=0.5)
    for _ in range(3):
        e.step(0 if e.pos[1] < 8 else 1)
    assert e.buf[0] == 1.5 and e.gen[0] == 1.5
    e = _env(data_generation_rate=100.0, max_buffer_size=50.0)
    e.step(0)
    assert e.buf[0] == 50.0 and e.gen[0] == 100.0 and e.lost[0] == 50.0
    e = _env(data_generation_rate=10.0, max_buffer_size=20.0)
    for a in (0, 1, 0):
        e.step(a)
    assert e.buf[0] == 20.0 and e.gen[0] == 30.0
    e = _env(data_generation_rate=2.0, collection_duration=5.0, positions=((900.0, 900.0),), grid=(1000, 1000))
    e.step(4)                                                  # collect step ages by collection_duration
    assert e.gen[0] == 10.0


def test_data_rate_table():
    assert O.DATA_RATE == {7: 683.75, 8: 390.625, 9: 220.0, 10: 122.5, 11: 55.0, 12: 31.25}


def test_reset_restores_position_battery_and_refills():
    e = _env(start=(3.0, 4.0))
    for a in (0, 3, 4):
        e.step(a)
    e.draws.load(e.draws.table)
    e.reset([0.25])
    assert tuple(e.pos) == (3.0, 4.0) and e.battery == 274.0
    assert e.buf[0] == 250.0 and e.gen[0] == 250.0 and e.sf[0] in (7, 9, 11, 12)
    assert e.step_count == 0 and not e.visited.any()
