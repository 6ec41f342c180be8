"""CPU oracle (numpy restatement) of the reference's UAV-LoRa environment step.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import this module,
and only as the checker / the reported CPU baseline.  The product package never
imports it: the CUDA path has no CPU fallback.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks this file against
golden vectors produced by running the unmodified reference in the build
container (`oracle/gen_golden.py`, `oracle/ref_harness.py`): every discrete
quantity bit-exact, every float bit-exact as well (same numpy scalar ops in the
same order) for the log10-normalised goldens.

It is a restatement, not a copy: state lives in flat numpy arrays, the step is
one function, the deterministic part of the RSSI is evaluated once per sensor
per step, and draws come from a pluggable source so that the same code serves
  * STREAM mode  -- draws pulled from the reference's own generators in the
    reference's call order (legacy global `np.random` MT19937 for shadowing,
    Python `random` for the duty-cycle gate, PCG64 for buffer fill), so that a
    seed alone reproduces a reference episode, and
  * SLOT mode    -- draws looked up in a slot-addressed table [6, N]
    (A,B,U,C,D,E of SURVEY.md section 3.2), which is what the C oracle and the
    CUDA kernels replay.

Reference lines followed (under /root/reference/src):
  environment/uav.py:88-125 (UAV ctor), :127-185 (move), :187-206 (hover),
      :224 (is_alive), :241-263 (reset, battery_drain_hover)
  environment/iot_sensors.py:13-37 (tables), :114-125 (step), :127-145
      (collect_data), :147-197 (calculate_rssi), :202-212
      (get_success_probability), :214-219 (is_in_range), :223-259
      (update_spreading_factor), :305-321 (reset)
  rewards/reward_function.py:46-57, :59-67, :69-79, :81-128
  environment/uav_env.py:366-394 (layout, urgency), :400-427 (reset),
      :429-488 (step), :494-516 (move action), :518-632 (collect action),
      :638-674 (observation), :676-700 (info)
  agents/dqn/dqn.py:177-451 (DomainRandEnv reset/step shaping/padding)
"""
from __future__ import annotations

import math
import random as _pyrandom

import numpy as np

SLOT_A, SLOT_B, SLOT_U, SLOT_C, SLOT_D, SLOT_E = range(6)
NUM_SLOTS = 6

# iot_sensors.py:13-37
DATA_RATE = {7: 5470 / 8, 8: 3125 / 8, 9: 1760 / 8, 10: 980 / 8, 11: 440 / 8, 12: 250 / 8}
REQUIRED_SNR = {7: -6.0, 8: -9.0, 9: -12.0, 10: -15.0, 11: -17.5, 12: -20.0}
SF_THRESHOLDS = ((-60, 7), (-70, 9), (-78, 11), (-85, 12))
LINK_QUALITY = {7: 1.0, 8: 0.8, 9: 0.6, 10: 0.4, 11: 0.2, 12: 0.1}  # uav_env.py:647

_F32 = np.float32


def default_params(**over):
    """Constants of the training / evaluation path (SURVEY.md section 8 header)."""
    p = dict(
        data_generation_rate=22.0 / 10,     # uav_env.py:271
        max_buffer_size=1000.0,             # uav_env.py:272
        rssi_threshold=-85.0,
        duty_cycle=10.0,
        max_battery=274.0,
        collection_duration=1.0,
        max_steps=2100,
        tx_power_dbm=14.0,                  # iot_sensors.py:46
        noise_floor_dbm=-105.0,             # iot_sensors.py:50
        altitude=100,                       # iot_sensors.py:51 (int)
        adr_lambda=0.1,                     # iot_sensors.py:54
        shadowing_std_db=4.0,               # iot_sensors.py:56
        capture_threshold_db=6.0,           # uav_env.py:569
        power_move=500.0, power_hover=700.0,  # uav.py:93-94
        init_sf=12,                         # iot_sensors.py:309 (reset forces 12)
        # RewardFunction as constructed at uav_env.py:339-344
        reward_per_byte=100.0, reward_new_sensor=5000.0, reward_completion=100.0,
        reward_urgency_reduction=20.0, reward_movement=10.0,
        penalty_revisit=-2.0, penalty_boundary=-50.0, penalty_collision=-10.0,
        penalty_battery=-0.5, penalty_hover=-5.0, penalty_step=-0.5,
        penalty_data_loss=-1.0, penalty_starvation=-1000.0,
        penalty_unvisited=-5000.0, penalty_starved=-1000.0,
        starvation_cr_threshold=0.20,
        normalise=True,
        include_sensor_positions=False,
    )
    p.update(over)
    return p


# --------------------------------------------------------------------------
# draw sources
# --------------------------------------------------------------------------
class StreamDraws:
    """Pulls draws from the reference's own generators in its call order and
    records them slot-addressed.  `np_random_state=None` means the global legacy
    numpy stream / global Python `random`, exactly as the reference uses them."""

    def __init__(self, sigma=4.0, np_random_state=None, py_random=None):
        self.sigma = sigma
        self._np = np_random_state if np_random_state is not None else np.random
        self._py = py_random if py_random is not None else _pyrandom
        self.table = None

    def begin(self, n):
        self.table = np.full((NUM_SLOTS, n), np.nan, dtype=np.float64)

    def normal(self, slot, i):
        v = float(self._np.normal(0, self.sigma))
        self.table[slot, i] = v
        return v

    def uniform01(self, i):
        v = self._py.random()
        self.table[SLOT_U, i] = v
        return v

    def burn_normals(self, k):
        """Consume k shadowing draws made by a caller outside the env (e.g. the
        greedy driver's `is_in_range` calls, greedy_agents.py:85); returns them."""
        return [float(self._np.normal(0, self.sigma)) for _ in range(int(k))]

    def probe(self):
        """One shadowing draw for a caller outside the env (not recorded in the slot table)."""
        return float(self._np.normal(0, self.sigma))


class SlotDraws:
    """Replays a slot-addressed table [6, N]; reading a NaN slot is an error
    (it means the oracle asked for a draw the reference never made)."""

    def __init__(self):
        self.table = None
        self.used = None

    def load(self, table):
        self.table = np.asarray(table, dtype=np.float64)
        self.used = np.zeros(self.table.shape, dtype=np.bool_)

    def begin(self, n):
        assert self.table is not None and self.table.shape == (NUM_SLOTS, n)

    def normal(self, slot, i):
        v = self.table[slot, i]
        if math.isnan(v):
            raise AssertionError("oracle requested an unrecorded draw (slot %d, sensor %d)" % (slot, i))
        self.used[slot, i] = True
        return float(v)

    def uniform01(self, i):
        return self.normal(SLOT_U, i)


# --------------------------------------------------------------------------
# numerics
# --------------------------------------------------------------------------
_D_BREAK = (4 * np.pi * 0.5 * 100) / 0.345          # iot_sensors.py:174 (ht=0.5, hr=100)
_FSPL_CONST = 20 * np.log10(868)                    # float64
_TWO_RAY_HT = 20 * np.log10(0.5)
_TWO_RAY_HR = 20 * np.log10(100)


def _log10_f32(d, normalise):
    """float32 log10 of a float32.  normalise=True is the host-independent,
    correctly-rounded definition used by every implementation in this repo;
    False is numpy's own (SIMD-dispatch dependent) float32 routine."""
    if normalise:
        return _F32(np.log10(np.float64(d)))
    return np.log10(d)


def deterministic_rssi(ux, uy, sx, sy, params):
    """iot_sensors.py:161-187 without the shadowing term.  ux,uy,sx,sy float32."""
    alt = params["altitude"]
    norm = params["normalise"]
    dx = (ux - sx) * 10
    dy = (uy - sy) * 10
    if norm:        # host-independent definitions: x**2 := x*x, log10 correctly rounded
        ground = np.sqrt(dx * dx + dy * dy)
        d = np.sqrt(ground * ground + alt ** 2)
    else:           # numpy's own SIMD `power` / `log10` float32 loops (host dependent)
        ground = np.sqrt(dx ** 2 + dy ** 2)
        d = np.sqrt(ground ** 2 + alt ** 2)
    lg = _log10_f32(d, norm)
    if alt == 100:
        d_break, c_hr = _D_BREAK, _TWO_RAY_HR
    else:
        d_break, c_hr = (4 * np.pi * 0.5 * alt) / 0.345, 20 * np.log10(alt)
    if d < d_break:
        path_loss = (20 * lg) + _FSPL_CONST - 28
    else:
        path_loss = (40 * lg) - _TWO_RAY_HT - c_hr
    return params["tx_power_dbm"] - path_loss


def np_var(x):
    return float(np.var(np.asarray(x, dtype=np.float64)))


# --------------------------------------------------------------------------
# the environment
# --------------------------------------------------------------------------
class OracleEnv:
    """One environment.  Public attributes mirror the snapshot keys produced by
    oracle/ref_harness.py::snapshot so traces compare key by key."""

    def __init__(self, grid, positions, start=(0.0, 0.0), params=None, draws=None):
        self.p = params if params is not None else default_params()
        self.W, self.H = int(grid[0]), int(grid[1])
        self.spos = np.array(positions, dtype=np.float32).reshape(-1, 2)   # iot_sensors.py:66
        self.n = self.spos.shape[0]
        self.start = np.array(start, dtype=np.float32)                     # uav.py:113
        self.draws = draws if draws is not None else StreamDraws(self.p["shadowing_std_db"])
        n = self.n
        self.pos = self.start.copy()
        self.battery = float(self.p["max_battery"])
        self.buf = np.zeros(n)
        self.gen = np.zeros(n)
        self.tx = np.zeros(n)
        self.lost = np.zeros(n)
        self.ema = np.full(n, np.nan)           # NaN <=> avg_rssi is None
        self.sf = np.full(n, 12, dtype=np.int8)
        self.cur = np.full(n, np.nan)           # current_rssi (transient)
        self.visited = np.zeros(n, dtype=np.bool_)
        self.step_count = 0
        self.total_reward = 0.0
        self.total_data_collected = 0.0
        self.capture_triggers = 0
        self.boundary_hits = 0
        self.edge_steps = 0
        self.last_step_bytes = 0.0
        self.previous_data_loss = 0.0
        self.collect_detail = None
        self.margin = math.inf                  # smallest threshold margin seen this step (dB / prob)

    # ---- helpers -------------------------------------------------------
    def _det(self):
        """Deterministic RSSI of every sensor at the current UAV position."""
        ux, uy = _F32(self.pos[0]), _F32(self.pos[1])
        return [deterministic_rssi(ux, uy, self.spos[i, 0], self.spos[i, 1], self.p)
                for i in range(self.n)]

    def _note_margin(self, a, b):
        m = abs(a - b)
        if m < self.margin:
            self.margin = m

    def _adr(self, i, rssi):
        """iot_sensors.py:223-259: EMA then first matching threshold; no match keeps SF."""
        self.cur[i] = rssi
        lam = self.p["adr_lambda"]
        if math.isnan(self.ema[i]):
            self.ema[i] = rssi
        else:
            self.ema[i] = (lam * rssi) + ((1 - lam) * self.ema[i])
        avg = self.ema[i]
        for thr, sf in SF_THRESHOLDS:
            self._note_margin(avg, thr)
            if avg > thr:
                self.sf[i] = sf
                break

    def _urgency(self, i):
        """uav_env.py:376-384."""
        util = self.buf[i] / self.p["max_buffer_size"]
        loss_rate = self.lost[i] / self.gen[i] if self.gen[i] > 0 else 0.0
        return float(np.clip(util * (1.0 + loss_rate * 10.0), 0.0, 1.0))

    def _aoi(self):
        """uav_env.py:386-394 (float32 array)."""
        out = np.zeros(self.n, dtype=np.float32)
        r = self.p["data_generation_rate"]
        for i in range(self.n):
            out[i] = self.buf[i] / r if r > 0 else 0.0
        return out

    # ---- reset -----------------------------------------------------------
    def reset(self, fill, init_sf=None):
        """uav_env.py:400-427 + iot_sensors.py:305-321.  `fill[i]` is the value of
        `np_random.uniform(0.20, 0.60)` for sensor i."""
        p = self.p
        self.pos = self.start.copy()
        self.battery = float(p["max_battery"])
        for i in range(self.n):
            self.buf[i] = p["max_buffer_size"] * np.clip(fill[i], 0.0, 1.0)
            self.gen[i] = float(self.buf[i])
        self.sf[:] = p["init_sf"] if init_sf is None else init_sf
        self.ema[:] = np.nan
        self.cur[:] = np.nan
        self.tx[:] = 0.0
        self.lost[:] = 0.0
        self.visited[:] = False
        self.step_count = 0
        self.total_reward = 0.0
        self.total_data_collected = 0.0
        self.capture_triggers = 0
        self.boundary_hits = 0
        self.edge_steps = 0
        self.last_step_bytes = 0.0
        self.previous_data_loss = 0.0
        self.draws.begin(self.n)
        return self.observation()

    # ---- observation -----------------------------------------------------
    def observation(self):
        """uav_env.py:638-674.  Mutates EMA/SF (draw D) and draws E."""
        p = self.p
        W, H = float(self.W), float(self.H)
        ux, uy = float(self.pos[0]), float(self.pos[1])
        out = [ux / W, uy / H, self.battery / p["max_battery"]]
        det = self._det()
        thr = p["rssi_threshold"]
        for i in range(self.n):
            urg = self._urgency(i)
            self._adr(i, float(det[i] + self.draws.normal(SLOT_D, i)))
            r_e = float(det[i] + self.draws.normal(SLOT_E, i))
            self._note_margin(r_e, thr)
            lq = LINK_QUALITY.get(int(self.sf[i]), 0.1) if r_e >= thr else 0.0
            out.extend([self.buf[i] / p["max_buffer_size"], urg, lq])
            if p["include_sensor_positions"]:
                out.append((float(self.spos[i, 0]) - ux) / W)
                out.append((float(self.spos[i, 1]) - uy) / H)
        return np.array(out, dtype=np.float32)

    # ---- step --------------------------------------------------------------
    def step(self, action):
        """uav_env.py:429-488."""
        p = self.p
        action = int(action)
        self.margin = math.inf
        self.draws.begin(self.n)
        self.step_count += 1
        W, H = float(self.W), float(self.H)
        ux, uy = float(self.pos[0]), float(self.pos[1])
        eps = 1e-6
        if ux <= eps or uy <= eps or ux >= W - 1 - eps or uy >= H - 1 - eps:
            self.edge_steps += 1
        dt = p["collection_duration"] if action == 4 else 1.0
        rate, cap = p["data_generation_rate"], p["max_buffer_size"]
        for i in range(self.n):                         # iot_sensors.py:114-125
            new = rate * dt
            self.gen[i] += new
            pot = self.buf[i] + new
            if pot > cap:
                self.lost[i] += pot - cap
                self.buf[i] = cap
            else:
                self.buf[i] = pot
        # uav_env.py:457-459.  The reference's builtin sum() runs over np.float64
        # objects, for which CPython adds left to right; over exact Python floats
        # (DomainRandEnv's fresh sensors) CPython >= 3.12 uses Neumaier
        # compensation instead.  Both are restated so goldens match to the bit.
        if p.get("loss_sum_compensated", False):
            cur_loss = sum(float(v) for v in self.lost)
        else:
            cur_loss = 0.0
            for v in self.lost:
                cur_loss = cur_loss + float(v)
        step_loss = cur_loss - self.previous_data_loss
        self.previous_data_loss = cur_loss

        self.collect_detail = None
        if action in (0, 1, 2, 3):
            self.last_step_bytes = 0.0
            reward = self._move(action, step_loss)
        elif action == 4:
            reward = self._collect(step_loss)
        else:
            raise ValueError("Invalid action: %r" % (action,))      # uav_env.py:468

        truncated = False
        if not (self.battery > 0.02 * p["max_battery"]):            # uav.py:224
            truncated = True
        if self.step_count >= p["max_steps"]:
            truncated = True
        if truncated:                                               # uav_env.py:481-485
            unvisited = self.n - int(self.visited.sum())
            if unvisited > 0:
                reward += p["penalty_unvisited"] * unvisited
            pen = 0.0
            for i in range(self.n):                                 # reward_function.py:59-67
                if self.gen[i] > 0:
                    cr = self.tx[i] / self.gen[i]
                    self._note_margin(cr, p["starvation_cr_threshold"])
                    if cr < p["starvation_cr_threshold"]:
                        pen += p["penalty_starved"]
            reward += pen
        self.total_reward += reward
        obs = self.observation()
        return obs, float(reward), False, truncated

    def _move(self, action, step_loss):
        """uav_env.py:494-516, uav.py:127-185, reward_function.py:69-79."""
        p = self.p
        new = self.pos.copy()
        if action == 0:
            new[1] += 1.0
        elif action == 1:
            new[1] -= 1.0
        elif action == 2:
            new[0] -= 1.0
        else:
            new[0] += 1.0
        before = self.battery
        ok = bool(0 <= new[0] < self.W and 0 <= new[1] < self.H)
        if ok:
            self.pos = new
            self.battery -= (p["power_move"] * 1) / 3600
        else:
            self.battery -= (p["power_move"] * 0.5 * 1) / 3600
            self.boundary_hits += 1
        used = before - self.battery
        reward = p["penalty_step"]
        reward += p["reward_movement"] if ok else p["penalty_boundary"]
        reward += p["penalty_battery"] * used
        reward += p["penalty_data_loss"] * step_loss
        return reward

    def _collect(self, step_loss):
        """uav_env.py:518-632."""
        p = self.p
        cd = p["collection_duration"]
        thr = p["rssi_threshold"]
        before = self._aoi()
        self.battery -= (p["power_hover"] * cd) / 3600                  # uav.py:205
        used = (p["power_hover"] / (60 * 60)) * cd                      # uav.py:263, uav_env.py:530
        det = self._det()
        attempts = {}                                                   # SF -> [sensor ids], insertion ordered
        for i in range(self.n):
            if self.buf[i] <= 0:
                continue
            self._adr(i, float(det[i] + self.draws.normal(SLOT_A, i)))
            r_b = float(det[i] + self.draws.normal(SLOT_B, i))
            self._note_margin(r_b, thr)
            if r_b < thr:                                               # iot_sensors.py:206-207
                p_link = 0.0
            else:
                snr = r_b - p["noise_floor_dbm"]
                p_link = 1.0 / (1.0 + np.exp(-(snr - REQUIRED_SNR.get(int(self.sf[i]), 7.5))))
            p_all = p_link * (p["duty_cycle"] / 100.0)
            u = self.draws.uniform01(i)
            self._note_margin(p_all, u)
            if p_all > u:
                attempts.setdefault(int(self.sf[i]), []).append(i)

        winners = {}
        collisions = 0
        for sf, ids in attempts.items():                                # uav_env.py:557-572
            if len(ids) == 1:
                winners[sf] = ids[0]
            else:
                collisions += len(ids) - 1
                order = sorted(ids, key=lambda j: self.cur[j], reverse=True)   # stable
                a, b = order[0], order[1]
                self._note_margin(self.cur[a], self.cur[b] + p["capture_threshold_db"])
                if self.cur[a] > (self.cur[b] + p["capture_threshold_db"]):
                    winners[sf] = a
                    self.capture_triggers += 1

        total = 0.0
        new_ids = []
        for sf, i in winners.items():                                   # uav_env.py:580-594
            r_c = float(det[i] + self.draws.normal(SLOT_C, i))
            self._note_margin(r_c, thr)
            in_range = r_c >= thr                                       # simple model: prob 1.0 / 0.0
            got = 0.0
            if in_range and self.buf[i] > 0:                            # iot_sensors.py:134-145
                got = min(self.buf[i], DATA_RATE[int(self.sf[i])] * cd)
                self.buf[i] -= got
                self.tx[i] += got
            if in_range and got > 0:
                total += got
                self.total_data_collected += got
                if not self.visited[i]:
                    new_ids.append(i)
                    self.visited[i] = True
        was_empty = any(self.buf[i] <= 0 for i in range(self.n))
        after = self._aoi()
        urgency_reduced = float(np.sum(np.maximum(0, before - after)))
        all_collected = all(self.buf[i] <= 0 for i in range(self.n))
        self.last_step_bytes = total
        if winners:
            mean_urg = float(np.mean([self._urgency(i) for i in winners.values()]))
        else:
            mean_urg = 0.0

        # reward_function.py:81-128, same order of additions
        reward = p["penalty_step"] + p["penalty_hover"]
        if total > 0:
            reward += p["reward_per_byte"] * total * mean_urg
            if new_ids:
                reward += p["reward_new_sensor"]
        if urgency_reduced > 0:
            reward += p["reward_urgency_reduction"] * urgency_reduced
        if was_empty and total == 0:
            reward += p["penalty_revisit"]
        reward += p["penalty_battery"] * used
        if collisions > 0:
            reward += p["penalty_collision"] * collisions
        if step_loss > 0:
            reward += p["penalty_data_loss"] * step_loss
        bufs = [float(v) for v in self.buf]                             # reward_function.py:46-57
        if len(bufs) > 1:
            mx = max(bufs)
            if mx != 0:
                reward += p["penalty_starvation"] * np_var([b / mx for b in bufs])
        if all_collected:
            reward += p["reward_completion"]
        self.collect_detail = dict(
            bytes_collected=total, was_new_sensor=float(len(new_ids) > 0), was_empty=float(was_empty),
            all_sensors_collected=float(all_collected), battery_used=used,
            collision_count=float(collisions), data_loss=step_loss,
            urgency_reduced=urgency_reduced, sensor_urgency=mean_urg,
            attempts={k: list(v) for k, v in attempts.items()}, winners=dict(winners))
        return reward

    # ---- introspection -----------------------------------------------------
    def action_mask(self):
        """Derived (the reference has no mask): the 4 move-validity bits of
        uav.py:175 plus COLLECT always valid."""
        x, y = _F32(self.pos[0]), _F32(self.pos[1])
        one = _F32(1.0)

        def ok(nx, ny):
            return bool(0 <= nx < self.W and 0 <= ny < self.H)

        return np.array([ok(x, y + one), ok(x, y - one), ok(x - one, y), ok(x + one, y), True],
                        dtype=np.bool_)

    def info(self):
        """uav_env.py:676-700 (the numeric entries)."""
        aoi = self._aoi()
        nv = int(self.visited.sum())
        return dict(
            max_urgency=float(np.max(aoi)), avg_urgency=float(np.mean(aoi)),
            high_urgency_sensors=int(np.sum(aoi > 0.8)),
            coverage_percentage=(nv / self.n) * 100,
            battery_percent=(self.battery / self.p["max_battery"]) * 100,
            sensors_collected=nv,
            sensor_collection_ratios=[self.tx[i] / max(self.gen[i], 1e-6) for i in range(self.n)],
        )

    def snapshot(self):
        return dict(
            pos=self.pos.astype(np.float32).copy(), battery=np.float64(self.battery),
            step=np.int32(self.step_count), buf=self.buf.copy(), gen=self.gen.copy(),
            tx=self.tx.copy(), lost=self.lost.copy(), ema=self.ema.copy(), sf=self.sf.copy(),
            visited=self.visited.copy(), total_reward=np.float64(self.total_reward),
            total_data_collected=np.float64(self.total_data_collected),
            capture_triggers=np.int32(self.capture_triggers),
            boundary_hits=np.int32(self.boundary_hits), edge_steps=np.int32(self.edge_steps),
            last_step_bytes=np.float64(self.last_step_bytes))


def nearest_sensor_greedy(env: "OracleEnv", probe):
    """Restatement of NearestSensorGreedy.select_action + GreedyAgent._move_toward
    (src/agents/dqn/dqn_evaluation_results/greedy_agents.py:73-100, :41-68).
    `probe(i)` returns the shadowing draw of `sensor_i.is_in_range` (one fresh draw per probe, made in index
    order and only until the first hit).  Returns (action, probes) where probes[i] is the draw used for
    sensor i (NaN = not probed)."""
    thr = env.p["rssi_threshold"]
    probes = np.full(env.n, np.nan)
    det = env._det()
    for i in range(env.n):                                  # :84-86
        if env.buf[i] > 0:
            probes[i] = probe(i)
            if float(det[i] + probes[i]) >= thr:
                return 4, probes
    cand = [i for i in range(env.n) if env.buf[i] > 0]      # :94-100
    if not cand:
        return 4, probes
    tgt = min(cand, key=lambda i: np.linalg.norm(env.spos[i] - env.pos))
    ux, uy = env.pos[0], env.pos[1]                         # float32
    dx, dy = env.spos[tgt, 0] - ux, env.spos[tgt, 1] - uy   # :43-45
    if abs(dx) <= 0.5 and abs(dy) <= 0.5:
        return 4, probes
    if abs(dx) > abs(dy):
        nx, ny = ux + (1 if dx > 0 else -1), uy
    else:
        nx, ny = ux, uy + (1 if dy > 0 else -1)
    if nx < 0 or nx >= env.W or ny < 0 or ny >= env.H:      # :57-59
        return 4, probes
    mx, my = nx - ux, ny - uy
    if mx > 0:
        return 3, probes
    if mx < 0:
        return 2, probes
    if my > 0:
        return 0, probes
    if my < 0:
        return 1, probes
    return 4, probes


def _move_toward(env: "OracleEnv", tgt):
    """GreedyAgent._move_toward (greedy_agents.py:41-68) toward sensor index `tgt`."""
    ux, uy = env.pos[0], env.pos[1]                         # float32
    dx, dy = env.spos[tgt, 0] - ux, env.spos[tgt, 1] - uy   # :43-45
    if abs(dx) <= 0.5 and abs(dy) <= 0.5:
        return 4
    if abs(dx) > abs(dy):
        nx, ny = ux + (1 if dx > 0 else -1), uy
    else:
        nx, ny = ux, uy + (1 if dy > 0 else -1)
    if nx < 0 or nx >= env.W or ny < 0 or ny >= env.H:      # :57-59
        return 4
    mx, my = nx - ux, ny - uy
    if mx > 0:
        return 3
    if mx < 0:
        return 2
    if my > 0:
        return 0
    if my < 0:
        return 1
    return 4


def max_throughput_greedy_v2(env: "OracleEnv", probe):
    """Restatement of MaxThroughputGreedyV2.select_action (greedy_agents.py:105-200).
    `probe(i)` returns the shadowing draw of `sensor_i.is_in_range`: the list comprehension at :133-138 probes
    EVERY sensor with data (no early exit), in index order.  Returns (action, probes).

    Arithmetic types follow the reference (numpy 2 promotion rules):
      * distance is a float32 norm; `distance / grid_size[0] * 5.0 * dist_w` stays float32 (:174);
      * `sensor.data_buffer` is np.float64 until the sensor first overflows -- iot_sensors.py:120 then assigns the
        Python float `max_buffer_size`, and every later value derives from it -- so for a sensor that has ever
        lost data `sf_score + buffer_score + duty_score` is a Python float and subtracting the np.float32 penalty
        rounds the whole score to float32 (weak-scalar promotion); otherwise the score is float64 (:176)."""
    thr = env.p["rssi_threshold"]
    probes = np.full(env.n, np.nan)
    det = env._det()
    battery_pct = env.battery / 274.0                       # :126 (hard-coded capacity)
    max_steps = env.p["max_steps"]
    steps_left = max_steps - env.step_count                 # :127
    steps_ratio = min(1.0, steps_left / max_steps)          # :190
    if battery_pct > 0.5 and steps_ratio > 0.5:             # :192-197
        sf_thr = 9
    elif battery_pct > 0.2 and steps_ratio > 0.2:
        sf_thr = 10
    else:
        sf_thr = 12
    immediate = False
    for i in range(env.n):                                  # :133-138
        if env.buf[i] > 0:
            probes[i] = probe(i)
            if float(det[i] + probes[i]) >= thr and int(env.sf[i]) <= sf_thr:
                immediate = True
    if immediate:
        return 4, probes
    if battery_pct < 0.1 or steps_left < 50:                # :157-162
        sf_w, dist_w = 1.0, 1.0
    elif battery_pct < 0.3 or steps_left < 150:
        sf_w, dist_w = 2.0, 1.0
    else:
        sf_w, dist_w = 5.0, 1.0
    best_score, best = -np.inf, None
    duty_prob = env.p["duty_cycle"] / 100.0                 # iot_sensors.py:105-106
    for i in range(env.n):                                  # :167-180
        if env.buf[i] <= 0:
            continue
        distance = np.linalg.norm(env.spos[i] - env.pos)    # np.float32
        sf_score = max(0, 13 - int(env.sf[i])) * 5.0 * sf_w
        py_float_buffer = env.lost[i] > 0
        buf = float(env.buf[i]) if py_float_buffer else np.float64(env.buf[i])
        buffer_score = (buf / env.p["max_buffer_size"]) * 10.0
        duty_score = duty_prob * 2.0
        distance_penalty = (distance / env.W) * 5.0 * dist_w
        score = sf_score + buffer_score + duty_score - distance_penalty
        if score > best_score:
            best_score, best = score, i
    if best is None:
        return 4, probes
    return _move_toward(env, best), probes


class SimToRealOracle:
    """Restatement of the sim-to-real evaluation wrappers
    (src/agents/dqn/dqn_evaluation_results/sim_to_real_analysis.py) over an OracleEnv:
      PaddedFixedEnv :140-167  zero padding to `obs_slots` sensor slots
      GPSNoisyEnv    :170-189  obs[0], obs[1] += N(0, gps_sigma) -- added to the NORMALISED entries, as executed
      WindDriftEnv   :192-224  per-episode drift vector (angle ~ U(0, 2 pi) drawn after reset); after every step
                               the true position moves by it, clipped to [0, W-1] x [0, H-1], and obs[0], obs[1]
                               report the RAW drifted position (not normalised), as executed
      LatencyEnv     :227-245  FIFO of `latency_steps` observations primed with copies of the reset observation
    The reference keeps these in separate classes; one wrapper instance applies whichever is switched on (wind,
    then GPS noise, then latency).  Their own draws come from the global legacy numpy stream; `extra` receives
    them in call order ([gps_x, gps_y] per step / reset, [angle] after a wind reset).  `source(kind, a, b)` may be
    given to replay recorded values instead."""

    def __init__(self, env: "OracleEnv", gps_sigma=0.0, wind_drift=0.0, latency_steps=0, obs_slots=50, source=None):
        self.env, self.gps_sigma, self.wind_drift, self.latency = env, gps_sigma, wind_drift, int(latency_steps)
        self.fps = 5 if env.p["include_sensor_positions"] else 3
        self.pad = (obs_slots - env.n) * self.fps
        self.source = source or (lambda kind, a, b: getattr(np.random, kind)(a, b))
        self.dx = self.dy = 0.0
        self.extra = []
        self.fifo = []

    def _draw(self, kind, a, b):
        v = self.source(kind, a, b)
        self.extra.append(float(v))
        return v

    def _pad(self, raw):
        return np.concatenate([raw, np.zeros(self.pad, dtype=np.float32)]).astype(np.float32)

    def _gps(self, obs):
        if self.gps_sigma > 0:                              # :176-181
            obs = obs.copy()
            obs[0] += self._draw("normal", 0, self.gps_sigma)
            obs[1] += self._draw("normal", 0, self.gps_sigma)
        return obs

    def reset(self, fill):
        self.extra = []
        obs = self._gps(self._pad(self.env.reset(fill)))
        if self.wind_drift > 0:                             # :205-210
            angle = self._draw("uniform", 0, 2 * np.pi)
            self.dx = self.wind_drift * np.cos(angle)
            self.dy = self.wind_drift * np.sin(angle)
        self.fifo = [obs.copy() for _ in range(self.latency)]   # :235-238
        return obs

    def step(self, action):
        self.extra = []
        env = self.env
        obs, reward, terminated, truncated = env.step(action)
        obs = self._pad(obs)
        if self.wind_drift > 0:                             # :212-224
            env.pos[0] = float(np.clip(env.pos[0] + self.dx, 0, env.W - 1))
            env.pos[1] = float(np.clip(env.pos[1] + self.dy, 0, env.H - 1))
            obs = obs.copy()
            obs[0] = env.pos[0]
            obs[1] = env.pos[1]
        obs = self._gps(obs)
        if self.latency:                                    # :240-245
            self.fifo.append(obs.copy())
            obs = self.fifo.pop(0)
        return obs, reward, terminated, truncated


class PerturbedOracle:
    """PerturbedDQNEnv of the sim-to-real sweep (sim_to_real_sweep.py:107-139) over AnalysisUAVEnv's zero padding
    (compare_agents.py:62-147), restated:
      shadow_sigma : every sensor's shadowing_std_db (:112-115) -- the base env's draw source is built with it
      gps_sigma    : _get_observation (:121-136) replaces uav.position by clip(position + N(0, sigma, size=2), 0,
                     [W-1, H-1]) (a float64 array) for the duration of the pass: obs[0], obs[1] report the noisy
                     position and the pass's ADR update / range check use it (cast to float32 by calculate_rssi,
                     iot_sensors.py:162); the true position is restored afterwards.  Applies to reset() too.
      wind_drift   : step (:117-119): for a move action, with probability min(0.95, wind / 2) (np.random.random(),
                     the global legacy stream) one EXTRA whole env step with the same action runs first; its reward
                     and observation are dropped, its effects on the state (and env.total_reward) stay.
    The wrapper's own draws come from the global legacy numpy stream; `extra` / `pre` record them like the fixtures
    (`source(kind, a, b)` replays recorded values instead)."""

    def __init__(self, env: "OracleEnv", gps_sigma=0.0, wind_drift=0.0, obs_slots=50, source=None):
        self.env, self.gps_sigma, self.wind_drift = env, float(gps_sigma), float(wind_drift)
        fps = 5 if env.p["include_sensor_positions"] else 3
        self.pad = (obs_slots - env.n) * fps
        self.source = source or (lambda kind, a, b: getattr(np.random, kind)(a, b) if kind != "random" else np.random.random())
        self.extra, self.pre, self.inner = [], [], []
        self._plain_obs = env.observation
        env.observation = self._observation          # OracleEnv.step / reset call self.observation()

    def _observation(self):
        env = self.env
        if self.gps_sigma <= 0:
            return self._plain_obs()
        true_pos = env.pos
        gx = float(self.source("normal", 0, self.gps_sigma))
        gy = float(self.source("normal", 0, self.gps_sigma))
        self.extra += [gx, gy]
        noisy = np.array([float(true_pos[0]) + gx, float(true_pos[1]) + gy], dtype=np.float64)
        env.pos = np.clip(noisy, 0, [env.W - 1, env.H - 1])
        try:
            return self._plain_obs()
        finally:
            env.pos = true_pos

    def _pad(self, raw):
        return np.concatenate([raw, np.zeros(self.pad, dtype=np.float32)]).astype(np.float32)

    def reset(self, fill):
        self.extra, self.pre = [], []
        return self._pad(self.env.reset(fill))

    def step(self, action):
        """Returns (padded obs, reward, terminated, truncated); self.inner lists the inner steps it ran as
        (obs, reward, truncated, extra, pre) tuples."""
        self.inner = []
        pre = []
        if self.wind_drift > 0 and action in (0, 1, 2, 3):
            u = float(self.source("random", 0, 1))
            pre = [u]
            if u < min(0.95, self.wind_drift / 2.0):
                self.extra = []
                o, r, te, tr = self.env.step(action)
                self.inner.append((self._pad(o), r, tr, list(self.extra), pre))
                pre = []
        self.extra = []
        o, r, te, tr = self.env.step(action)
        self.inner.append((self._pad(o), r, tr, list(self.extra), pre))
        return self._pad(o), r, te, tr


COLLECT_KEYS = ("bytes_collected", "was_new_sensor", "was_empty", "all_sensors_collected",
                "battery_used", "collision_count", "data_loss", "urgency_reduced",
                "sensor_urgency")
INFO_SCALARS = ("max_urgency", "avg_urgency", "high_urgency_sensors", "coverage_percentage",
                "battery_percent", "sensors_collected")


def traced_step(env: OracleEnv, action) -> dict:
    """Step and return the same record layout as ref_harness.traced_step."""
    obs, reward, terminated, truncated = env.step(action)
    st = env.snapshot()
    st["obs"] = obs
    st["reward"] = np.float64(reward)
    st["truncated"] = np.bool_(truncated)
    st["terminated"] = np.bool_(terminated)
    st["draws"] = env.draws.table.copy()
    cd = env.collect_detail
    for k in COLLECT_KEYS:
        st["c_" + k] = np.float64(cd[k]) if cd is not None else np.float64(np.nan)
    info = env.info()
    for k in INFO_SCALARS:
        st["i_" + k] = np.float64(info[k])
    st["i_ratios"] = np.array(info["sensor_collection_ratios"], dtype=np.float64)
    st["margin"] = np.float64(env.margin)
    return st


def traced_reset(env: OracleEnv, fill, init_sf=None) -> dict:
    obs = env.reset(fill, init_sf)
    st = env.snapshot()
    st["obs"] = obs
    st["draws"] = env.draws.table.copy()
    st["fill"] = np.asarray(fill, dtype=np.float64)
    return st


def stack_steps(steps: list) -> dict:
    return {k: np.stack([s[k] for s in steps]) for k in steps[0]}


# --------------------------------------------------------------------------
# stream-mode episode reproduction (a seed alone reproduces a reference episode)
# --------------------------------------------------------------------------
def seeded_layout(grid, n, seed):
    """`random.seed(s); np.random.seed(s); UAVEnvironment(...)` (sweep_eval.py:460-462):
    the layout is the first thing drawn from the global legacy stream
    (uav_env.py:366-374)."""
    _pyrandom.seed(seed)
    np.random.seed(seed)
    xy = np.random.uniform(low=[0.0, 0.0], high=[float(grid[0]), float(grid[1])], size=(n, 2))
    return [(float(x), float(y)) for x, y in xy]


def seeded_fill(n, seed):
    """`env.reset(seed=s)` -> PCG64(SeedSequence(s)).uniform(0.20, 0.60) per sensor
    (uav_env.py:407-410; gymnasium seeding as understood for 1.2.2)."""
    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
    return [float(g.uniform(0.20, 0.60)) for _ in range(n)]


def replay_seeded_episode(grid, n, seed, actions, burn=None, params=None):
    """Reproduce `sweep_eval.py::_greedy_worker`-style episodes from the seed:
    actions are given; `burn[t]` = shadowing draws the driver consumed before
    step t (NearestSensorGreedy's is_in_range probes)."""
    params = params if params is not None else default_params()
    pos = seeded_layout(grid, n, seed)
    env = OracleEnv(grid, pos, (0.0, 0.0), params, StreamDraws(params["shadowing_std_db"]))
    first = traced_reset(env, seeded_fill(n, seed))
    steps = []
    for t, a in enumerate(actions):
        if burn is not None:
            env.draws.burn_normals(burn[t])
        steps.append(traced_step(env, a))
    return first, stack_steps(steps), env
