"""Shared harness of the GPU parity tests (and of __graft_entry__.smoke()).

Runs the SAME seeded inputs (layouts, actions, slot-addressed draw tables) through
  * the C oracle (oracle/uavsim_oracle.c, pinned to the reference's golden vectors), and
  * the CUDA path, called through the C ABI exactly as a user would (UAVVecEnv ->
    libuavsim.so),
and compares every state column after every step.  Discrete quantities must match
bit-exactly; floats are compared relatively (BASELINE.json: 1e-4; the parity build is
expected to sit at ~1e-16).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT / "oracle"), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import c_oracle  # noqa: E402
import uavsim_oracle as O  # noqa: E402

SLOT_U = 2


def make_tables(rng, E, P, sigma=4.0):
    d = rng.normal(0.0, sigma, size=(E, 6, P))
    d[:, SLOT_U, :] = rng.random(size=(E, P))
    return d


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    one_nan = np.isnan(a) ^ np.isnan(b)
    if one_nan.any():
        return np.inf
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
    e = np.where(both_nan, 0.0, np.abs(a - b) / den)
    e = np.where(a == b, 0.0, e)
    return float(np.max(e)) if e.size else 0.0


class Scenario:
    def __init__(self, E, N, P, grid, seed, start=None, per_env_n=None, per_env_grid=None):
        rng = np.random.default_rng(seed)
        self.E, self.P = E, P
        self.n = np.full(E, N, dtype=np.int32) if per_env_n is None else np.asarray(per_env_n, dtype=np.int32)
        self.grid = (np.tile(np.asarray(grid, dtype=np.int32), (E, 1)) if per_env_grid is None
                     else np.asarray(per_env_grid, dtype=np.int32))
        xy = rng.uniform(0.0, 1.0, size=(E, P, 2)) * self.grid[:, None, :]
        self.xy = xy.astype(np.float32)
        if start is None:
            self.start = np.zeros((E, 2), dtype=np.float32)
        elif isinstance(start, str) and start == "random":
            self.start = (rng.uniform(0.05, 0.95, size=(E, 2)) * self.grid).astype(np.float32)
        else:
            self.start = np.tile(np.asarray(start, dtype=np.float32), (E, 1))
        self.rng = rng


def gpu_state(env):
    """Host copies of every state column of the CUDA env."""
    st = env.state()
    rec = env.records()
    vis = (rec["visited_lo"].astype(np.uint64) | (rec["visited_hi"].astype(np.uint64) << np.uint64(32)))
    return dict(rec=rec, buf=st["buffer"].cpu().numpy(), tx=st["transmitted"].cpu().numpy(),
                lost=st["lost"].cpu().numpy(), ema=st["avg_rssi"].cpu().numpy(),
                gen=env.generated().cpu().numpy(), sf=st["spreading_factor"].cpu().numpy(), visited=vis)


def oracle_visited_mask(envs):
    w = (np.uint64(1) << np.arange(64, dtype=np.uint64))
    return (envs["visited"].astype(np.uint64) * w).sum(axis=1).astype(np.uint64)


def compare_step(co, env, tag, acc, float_tol=None):
    """Compare all state after a step/reset; accumulate mismatch counts into acc."""
    g = gpu_state(env)
    oe = co.envs
    n = oe["N"]
    P = co.P
    live = np.arange(P)[None, :] < n[:, None]

    def disc(name, a, b):
        bad = int(np.sum(np.asarray(a) != np.asarray(b)))
        if bad:
            acc["discrete_mismatches"] += bad
            acc["first_bad"].append((tag, name, bad))

    def flt(name, a, b, mask=None):
        if mask is not None:
            a = np.where(mask, a, 0.0)
            b = np.where(mask, b, 0.0)
        e = rel_err(a, b)
        acc["max_rel_err"] = max(acc["max_rel_err"], e)
        acc["per_key"][name] = max(acc["per_key"].get(name, 0.0), e)
        if float_tol is not None and e > float_tol:
            acc["first_bad"].append((tag, name, e))

    r = g["rec"]
    disc("x", r["x"], oe["x"]); disc("y", r["y"], oe["y"])
    disc("step", r["step"], oe["step"])
    disc("sf", np.where(live, g["sf"], 0), np.where(live, oe["sf"][:, :P], 0))
    disc("visited", g["visited"], oracle_visited_mask(oe))
    for k in ("capture_triggers", "boundary_hits", "edge_steps", "collisions", "delivered"):
        disc(k, r[k], oe[k])
    disc("first_full_coverage_step", r["first_full_coverage_step"], oe["first_full_coverage_step"])
    flt("battery", r["battery"], oe["battery"])
    flt("total_reward", r["total_reward"], oe["total_reward"])
    flt("episode_return", r["episode_return"], oe["episode_return"])
    flt("total_data_collected", r["total_data_collected"], oe["total_data_collected"])
    if co.p.domain_rand:   # DomainRandEnv._prev_dist_nearest only exists in the wrapper
        flt("prev_dist", r["prev_dist_nearest"], oe["prev_dist_nearest"].astype(np.float32))
    for k in ("buf", "tx", "lost", "ema", "gen"):
        flt(k, g[k], oe[k][:, :P], live)
    # buffer-empty / saturation decisions are discrete
    disc("buf_empty", np.where(live, g["buf"] <= 0, False), np.where(live, oe["buf"][:, :P] <= 0, False))
    disc("buf_full", np.where(live, g["buf"] >= co.p.max_buffer_size, False),
         np.where(live, oe["buf"][:, :P] >= co.p.max_buffer_size, False))
    obs = env.obs.cpu().numpy()
    flt("obs", obs, co.obs)
    # link-quality category (discrete: SF x in-range) lives in obs[..., 3+fps*i+2]
    fps = co.p.features_per_sensor
    lq_g = obs[:, 5::fps][:, :P]
    lq_o = co.obs[:, 5::fps][:, :P]
    disc("link_quality", np.where(live[:, : lq_g.shape[1]], lq_g, 0), np.where(live[:, : lq_o.shape[1]], lq_o, 0))
    if env.action_mask is not None:
        disc("action_mask", env.action_mask.cpu().numpy(), co.mask)
    return g


def run_parity_case(pkg, E=64, N=20, P=None, grid=(500, 500), steps=32, seed=0, p_collect=0.4,
                    precision="f64", auto_reset=False, start=None, per_env_n=None, per_env_grid=None,
                    params_over=None, float_tol=None, compare_every=1, device="cuda:0", ref_kwargs=None,
                    domain_rand=False, obs_slots=None):
    """Returns a dict of mismatch statistics.  `ref_kwargs` are reference-constructor keywords
    (uav_env.py:266-287) forwarded to UAVVecEnv; `params_over` the matching oracle parameters."""
    P = int(P if P is not None else (max(per_env_n) if per_env_n is not None else N))
    sc = Scenario(E, N, P, grid, seed, start=start, per_env_n=per_env_n, per_env_grid=per_env_grid)
    rng = sc.rng
    params = O.default_params(**(params_over or {}))
    slots = obs_slots if obs_slots is not None else P
    co = c_oracle.COracleBatch(params, E=E, P=P, obs_slots=slots, domain_rand=domain_rand)
    co.set_scenario(sc.grid, sc.n, sc.xy, sc.start)

    kw = dict(ref_kwargs or {})
    env = pkg.UAVVecEnv(E, grid_size=sc.grid, num_sensors=sc.n, max_sensors=P, obs_slots=slots,
                        sensor_positions=sc.xy, uav_start_position=sc.start, precision=precision, rng="replay",
                        auto_reset=auto_reset, device=device, step_detail=True,
                        domain_randomisation=domain_rand, **kw)
    acc = dict(discrete_mismatches=0, max_rel_err=0.0, per_key={}, first_bad=[], steps=0, dones=0,
               collisions=0, captures=0, bytes=0.0)
    fill = rng.uniform(0.20, 0.60, size=(E, P))
    d0 = make_tables(rng, E, P, params["shadowing_std_db"])
    init_sf = None
    if domain_rand:
        init_sf = rng.choice(np.array([7, 9, 11, 12], dtype=np.int32), size=E)
    co.reset(None if domain_rand else fill, d0, init_sf=init_sf)
    env.set_replay(reset_draws=d0, reset_fill=fill, reset_init_sf=init_sf,
                   reset_sensor_xy=sc.xy if domain_rand else None,
                   reset_grid_wh=sc.grid if domain_rand else None,
                   reset_start_xy=sc.start if domain_rand else None)
    env.reset_tensors()
    compare_step(co, env, "reset", acc, float_tol)

    for t in range(steps):
        collect = rng.random(E) < p_collect
        actions = np.where(collect, 4, rng.integers(0, 4, size=E)).astype(np.int32)
        dt = make_tables(rng, E, P, params["shadowing_std_db"])
        if auto_reset:
            fill_t = rng.uniform(0.20, 0.60, size=(E, P))
            dr_t = make_tables(rng, E, P, params["shadowing_std_db"])
            kwargs = {}
            if domain_rand:   # a fresh injected scenario for whoever finishes this step
                sc_t = Scenario(E, N, P, grid, seed * 7919 + t + 1, start="random", per_env_n=sc.n,
                                per_env_grid=per_env_grid)
                isf_t = rng.choice(np.array([7, 9, 11, 12], dtype=np.int32), size=E)
                kwargs = dict(reset_init_sf=isf_t, reset_sensor_xy=sc_t.xy, reset_grid_wh=sc_t.grid,
                              reset_start_xy=sc_t.start)
            env.set_replay(step_draws=dt, reset_draws=dr_t, reset_fill=fill_t, **kwargs)
        else:
            env.set_replay(step_draws=dt)
        o_obs, o_rew, o_done = co.step(actions, dt)
        o_rew = o_rew.copy()
        o_done = o_done.copy()
        o_detail = co.detail.copy()
        o_term = o_obs.copy()
        if auto_reset and o_done.any():
            ids = np.flatnonzero(o_done)
            if domain_rand:
                e_ = co.envs
                e_["W"][ids], e_["H"][ids] = sc_t.grid[ids, 0], sc_t.grid[ids, 1]
                e_["spx"][ids, :P], e_["spy"][ids, :P] = sc_t.xy[ids, :, 0], sc_t.xy[ids, :, 1]
                e_["start_x"][ids], e_["start_y"][ids] = sc_t.start[ids, 0], sc_t.start[ids, 1]
            co.reset(None if domain_rand else fill_t, dr_t, init_sf=isf_t if domain_rand else None, env_ids=ids)
        import torch
        env.step_tensors(torch.as_tensor(actions, device=env.device))
        g_rew = env.reward.cpu().numpy()
        g_done = env.done.cpu().numpy()
        acc["steps"] += E
        acc["dones"] += int(o_done.sum())
        bad = int(np.sum(g_done != o_done))
        if bad:
            acc["discrete_mismatches"] += bad
            acc["first_bad"].append((t, "done", bad))
        e_r = rel_err(g_rew, o_rew)
        acc["max_rel_err"] = max(acc["max_rel_err"], e_r)
        acc["per_key"]["reward"] = max(acc["per_key"].get("reward", 0.0), e_r)
        det = env.step_detail.cpu().numpy()
        for idx, name in ((0, "collision_count"), (4, "was_new"), (5, "was_empty"), (6, "all_collected"),
                          (8, "attempt_lo"), (9, "attempt_hi"), (10, "winner_lo"), (11, "winner_hi"),
                          (13, "truncated")):
            b = int(np.sum(det[:, idx] != o_detail[:, idx]))
            if b:
                acc["discrete_mismatches"] += b
                acc["first_bad"].append((t, name, b))
        for idx, name in ((1, "bytes"), (2, "urgency_reduced"), (3, "mean_urgency"), (7, "data_loss"),
                          (12, "base_reward")):
            e_d = rel_err(det[:, idx], o_detail[:, idx])
            acc["max_rel_err"] = max(acc["max_rel_err"], e_d)
            acc["per_key"][name] = max(acc["per_key"].get(name, 0.0), e_d)
        acc["collisions"] += int(o_detail[:, 0].sum())
        acc["bytes"] += float(o_detail[:, 1].sum())
        if auto_reset and o_done.any():
            ids = np.flatnonzero(o_done)
            tob = env.terminal_obs.cpu().numpy()[ids]
            e_t = rel_err(tob, o_term[ids])
            acc["max_rel_err"] = max(acc["max_rel_err"], e_t)
            acc["per_key"]["terminal_obs"] = max(acc["per_key"].get("terminal_obs", 0.0), e_t)
        if (t % compare_every) == 0 or t == steps - 1:
            compare_step(co, env, t, acc, float_tol)
    acc["captures"] = int(co.envs["capture_triggers"].sum())
    acc["first_bad"] = acc["first_bad"][:8]
    env.poll_errors()
    env.close()
    return acc
