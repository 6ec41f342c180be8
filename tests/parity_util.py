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


def rel_err(a, b, floor=None):
    """max |a - b| / max(|a|, |b|[, floor]).  `floor` (scalar or array) keeps the error RELATIVE TO THE QUANTITY'S
    SCALE for values that are small only by cancellation (a buffer drained to a remainder of 1e-6 bytes, a reward
    whose terms cancel): without it the ratio measures the float32 state's absolute resolution, not its accuracy."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    one_nan = np.isnan(a) ^ np.isnan(b)
    if one_nan.any():
        return np.inf
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
    if floor is not None:
        den = np.maximum(den, floor)
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


# Arithmetic noise of the float32 throughput build at a discrete compare, in the compare's own units.  A float32
# trajectory may take the other branch of a compare only where the float64 oracle's margin at that env-step
# (oracle/uavsim_oracle.c, D_MARGIN_*) is below these; the environment is then counted as a justified flip and
# leaves the comparison (its trajectory is a different, equally valid one from there on).
#   dB    : MUFU lg2.approx is 2 ulp of ~25 -> 1.1e-5 dB (x2 in the far path-loss branch); draws float64 -> float32
#           2.4e-7; the stored EMA (dB above the range threshold, |value| <= 30) is rounded to 8 ulp, <= 7.6e-6 dB,
#           per step and decays by 0.9 per update.  Flips observed on the B200: margins <= 1.1e-5 dB.
#   prob  : __expf / __fdividef sigmoid, ~2 ulp of a value <= 1, times the duty fraction: 2.4e-8 -- the rssi noise
#           above moves p_link by at most its slope 0.25 / dB: 1e-5.
#   bytes : the float32 buffer (ulp 6e-5 at 1000 B) has accumulated up to a few hundred half-ulp roundings.
#   dist  : the two-ray model switches branch at d_break = 1821.2 m and the branches differ by 0.46 dB there
#           (iot_sensors.py:174-187): float32 resolves a distance of that size to 1.2e-4 m, the squared-distance
#           compare of the throughput build to ~5e-4 m.  A branch flip changes no discrete STATE by itself (it shows
#           as a 0.46 dB step in one RSSI sample), so an environment whose margin falls below the bound leaves the
#           comparison whether or not a difference is visible yet.
F32_MARGIN = dict(db=5e-5, prob=2e-5, bytes=2e-3, dist=2e-3)
_MARGIN_IDX = dict(db=16, prob=17, bytes=18, dist=20)


def _env_any(bad):
    bad = np.asarray(bad)
    return bad.reshape(bad.shape[0], -1).any(axis=1) if bad.ndim > 1 else bad


def new_acc(E):
    return dict(discrete_mismatches=0, max_rel_err=0.0, per_key={}, first_bad=[], steps=0, dones=0, collisions=0,
                captures=0, bytes=0.0, diverged=np.zeros(E, dtype=bool), flips=0, unmargined_flips=0,
                flip_margins=[], reward_abs_over_scale=0.0)


def compare_step(co, env, tag, acc, float_tol=None, extra_disc=(), extra_flt=(), margins=None, state=True):
    """Compare all state after a step/reset; accumulate mismatch counts into acc.

    extra_disc / extra_flt: (name, cuda_value, oracle_value) triples with a leading env axis, compared together
    with the state (reward, done, the per-step diagnostics).
    margins: None = strict (every discrete mismatch counts).  Otherwise the oracle's detail rows [E, >= 19] of the
    env-step just compared: an environment whose discrete state differs for the first time is a *justified flip* iff
    one of the oracle's compare margins at this step is below F32_MARGIN; it then leaves every later comparison.
    Floats are compared over the environments that have not diverged."""
    g = gpu_state(env) if state else None
    oe = co.envs
    n = oe["N"]
    P = co.P
    E = len(n)
    live = np.arange(P)[None, :] < n[:, None]
    if "diverged" not in acc:
        acc["diverged"] = np.zeros(E, dtype=bool)
        acc.setdefault("flips", 0)
        acc.setdefault("unmargined_flips", 0)
        acc.setdefault("flip_margins", [])
    out = acc["diverged"]
    env_bad = np.zeros(E, dtype=bool)

    def disc(name, a, b):
        bad = _env_any(np.asarray(a) != np.asarray(b)) & ~out
        cnt = int(np.sum((np.asarray(a) != np.asarray(b)).reshape(E, -1)[~out])) if bad.any() else 0
        if cnt:
            env_bad[:] |= bad
            if margins is None:
                acc["discrete_mismatches"] += cnt
                acc["first_bad"].append((tag, name, cnt))
            else:
                acc.setdefault("flip_names", []).append((tag, name, cnt))

    if state:
        r = g["rec"]
        disc("x", r["x"], oe["x"]); disc("y", r["y"], oe["y"])
        disc("step", r["step"], oe["step"])
        disc("sf", np.where(live, g["sf"], 0), np.where(live, oe["sf"][:, :P], 0))
        disc("visited", g["visited"], oracle_visited_mask(oe))
        for k in ("capture_triggers", "boundary_hits", "edge_steps", "collisions", "delivered"):
            disc(k, r[k], oe[k])
        disc("first_full_coverage_step", r["first_full_coverage_step"], oe["first_full_coverage_step"])
        # buffer-empty / saturation decisions are discrete
        disc("buf_empty", np.where(live, g["buf"] <= 0, False), np.where(live, oe["buf"][:, :P] <= 0, False))
        disc("buf_full", np.where(live, g["buf"] >= co.p.max_buffer_size, False),
             np.where(live, oe["buf"][:, :P] >= co.p.max_buffer_size, False))
        obs = env.obs.cpu().numpy()
        # link-quality category (discrete: SF x in-range) lives in obs[..., 3+fps*i+2]
        fps = co.p.features_per_sensor
        lq_g = obs[:, 5::fps][:, :P]
        lq_o = co.obs[:, 5::fps][:, :P]
        disc("link_quality", np.where(live[:, : lq_g.shape[1]], lq_g, 0), np.where(live[:, : lq_o.shape[1]], lq_o, 0))
        if env.action_mask is not None:
            disc("action_mask", env.action_mask.cpu().numpy(), co.mask)
    for name, a, b in extra_disc:
        disc(name, a, b)

    if margins is not None:   # a path-loss branch decided within float32 resolution of the break: see F32_MARGIN
        near = (margins[:, _MARGIN_IDX["dist"]] < F32_MARGIN["dist"]) & ~out & ~env_bad
        if near.any():
            acc["branch_margin_drops"] = acc.get("branch_margin_drops", 0) + int(near.sum())
            out |= near
    if margins is not None and env_bad.any():
        for e in np.flatnonzero(env_bad):
            m = {k: float(margins[e, i]) for k, i in _MARGIN_IDX.items()}
            ok = any(m[k] < F32_MARGIN[k] for k in m)
            acc["flips"] += 1
            acc["flip_margins"].append((tag, int(e), m))
            if not ok:
                acc["unmargined_flips"] += 1
                acc["first_bad"].append((tag, "unmargined flip", int(e), m))
        out |= env_bad
    keep = ~out

    # float32 builds (margin mode): errors relative to max(|value|, 1 % of the quantity's full scale)
    fl = (lambda v: v) if margins is not None else (lambda v: None)
    maxbuf = float(co.p.max_buffer_size)

    def flt(name, a, b, mask=None, floor=None):
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        if mask is not None:
            a = np.where(mask, a, 0.0)
            b = np.where(mask, b, 0.0)
        if floor is not None and np.ndim(floor) > 0:
            floor = np.asarray(floor)[keep]
        e = rel_err(a[keep], b[keep], floor)
        acc["max_rel_err"] = max(acc["max_rel_err"], e)
        acc["per_key"][name] = max(acc["per_key"].get(name, 0.0), e)
        if float_tol is not None and e > float_tol:
            acc["first_bad"].append((tag, name, e))
        if margins is not None and e > 1e-4 and len(acc.setdefault("float_outliers", [])) < 6:   # where, for the report
            ak, bk = a[keep], b[keep]
            den = np.maximum(np.maximum(np.abs(ak), np.abs(bk)), 1e-300)
            w = np.unravel_index(int(np.nanargmax(np.where(ak == bk, 0.0, np.abs(ak - bk) / den))), ak.shape)
            acc["float_outliers"].append((tag, name, int(np.flatnonzero(keep)[w[0]]), tuple(int(x) for x in w[1:]),
                                          float(ak[w]), float(bk[w])))

    if state:
        cum = acc.get("cum_scale")      # running sum of |reward terms| per env: the scale of the accumulated returns
        flt("battery", r["battery"], oe["battery"])
        flt("total_reward", r["total_reward"], oe["total_reward"], floor=fl(cum))
        flt("episode_return", r["episode_return"], oe["episode_return"], floor=fl(cum))
        flt("total_data_collected", r["total_data_collected"], oe["total_data_collected"], floor=fl(0.01 * maxbuf))
        if co.p.domain_rand:   # DomainRandEnv._prev_dist_nearest only exists in the wrapper
            flt("prev_dist", r["prev_dist_nearest"], oe["prev_dist_nearest"].astype(np.float32))
        for k in ("buf", "tx", "lost", "gen"):
            flt(k, g[k], oe[k][:, :P], live, floor=fl(0.01 * maxbuf))
        flt("ema", g["ema"], oe["ema"][:, :P], live)
        flt("obs", obs, co.obs, floor=fl(0.01))       # every feature lives in [0, 1] (positions in [-1, 1])
    for item in extra_flt:
        flt(item[0], item[1], item[2], floor=fl(item[3]) if len(item) > 3 else None)
    return g


_DETAIL_DISC = ((0, "collision_count"), (4, "was_new"), (5, "was_empty"), (6, "all_collected"), (8, "attempt_lo"),
                (9, "attempt_hi"), (10, "winner_lo"), (11, "winner_hi"), (13, "truncated"), (14, "starved"))
_DETAIL_FLT = ((1, "bytes"), (2, "urgency_reduced"), (3, "mean_urgency"), (7, "data_loss"))


def run_parity_case(pkg, E=64, N=20, P=None, grid=(500, 500), steps=32, seed=0, p_collect=0.4,
                    precision="f64", auto_reset=False, start=None, per_env_n=None, per_env_grid=None,
                    params_over=None, float_tol=None, compare_every=1, device="cuda:0", ref_kwargs=None,
                    domain_rand=False, obs_slots=None, margin_aware=False, draws="replay", philox_seed=2026,
                    lean_twin=False):
    """Returns a dict of mismatch statistics.  `ref_kwargs` are reference-constructor keywords
    (uav_env.py:266-287) forwarded to UAVVecEnv; `params_over` the matching oracle parameters.

    draws="replay": numpy draw tables go to both sides (the CUDA REPLAY kernels).
    draws="philox": the CUDA side runs its own Philox4x32-10 draws and exports the values it used; the oracle
        replays exactly those (no auto-reset: an in-kernel restart does not export its draws).
    lean_twin (draws="philox"): a second handle WITHOUT diagnostics / draw export -- the LEAN instantiation the
        launcher picks for production, i.e. the kernel bench.py times -- steps next to the diagnostic one and must
        agree with it bit for bit (obs, reward, done) at every step, which ties the oracle comparison to it.
    margin_aware (float32 builds): discrete differences are accepted as justified flips where the oracle's compare
        margin is below the float32 arithmetic noise (F32_MARGIN); see compare_step."""
    import torch
    P = int(P if P is not None else (max(per_env_n) if per_env_n is not None else N))
    sc = Scenario(E, N, P, grid, seed, start=start, per_env_n=per_env_n, per_env_grid=per_env_grid)
    rng = sc.rng
    params = O.default_params(**(params_over or {}))
    slots = obs_slots if obs_slots is not None else P
    co = c_oracle.COracleBatch(params, E=E, P=P, obs_slots=slots, domain_rand=domain_rand)
    philox = draws == "philox"
    assert not (philox and auto_reset), "an in-kernel restart does not export its draws"
    if margin_aware:
        compare_every = 1

    kw = dict(ref_kwargs or {})
    env = pkg.UAVVecEnv(E, grid_size=sc.grid, num_sensors=sc.n, max_sensors=P, obs_slots=slots,
                        sensor_positions=None if (philox and domain_rand) else sc.xy,
                        uav_start_position=sc.start, precision=precision,
                        rng="philox" if philox else "replay", seed=philox_seed, export_draws=philox,
                        auto_reset=auto_reset, device=device, step_detail=True,
                        domain_randomisation=domain_rand, **kw)
    acc = new_acc(E)
    sigma = params["shadowing_std_db"]
    init_sf = None
    twin = None
    if lean_twin:
        assert philox
        twin = pkg.UAVVecEnv(E, grid_size=sc.grid, num_sensors=sc.n, max_sensors=P, obs_slots=slots,
                             sensor_positions=None if domain_rand else sc.xy, uav_start_position=sc.start,
                             precision=precision, rng="philox", seed=philox_seed, auto_reset=auto_reset, device=device,
                             domain_randomisation=domain_rand, **kw)
        twin.reset_tensors()
        acc["lean_twin_steps"] = 0

    def exported():
        return np.nan_to_num(env.draw_export.cpu().numpy().astype(np.float64))

    if philox:
        env.reset_tensors()
        d0 = exported()
        fill = d0[:, SLOT_U, :].copy()           # a reset exports the fill it used in slot U
        if domain_rand:                          # the scenario the kernel drew: grid, layout, far start, inherited SF
            rec = env.records()
            sc.grid = np.stack([rec["grid_w"], rec["grid_h"]], 1).astype(np.int32)
            sc.xy = env.state()["sensor_pos"].cpu().numpy()
            sc.start = np.stack([rec["start_x"], rec["start_y"]], 1).astype(np.float32)
            init_sf = rec["init_sf"].astype(np.int32)
        co.set_scenario(sc.grid, sc.n, sc.xy, sc.start)
        co.reset(None if domain_rand else fill, d0, init_sf=init_sf)
    else:
        co.set_scenario(sc.grid, sc.n, sc.xy, sc.start)
        fill = rng.uniform(0.20, 0.60, size=(E, P))
        d0 = make_tables(rng, E, P, sigma)
        if domain_rand:
            init_sf = rng.choice(np.array([7, 9, 11, 12], dtype=np.int32), size=E)
        co.reset(None if domain_rand else fill, d0, init_sf=init_sf)
        env.set_replay(reset_draws=d0, reset_fill=fill, reset_init_sf=init_sf,
                       reset_sensor_xy=sc.xy if domain_rand else None,
                       reset_grid_wh=sc.grid if domain_rand else None,
                       reset_start_xy=sc.start if domain_rand else None)
        env.reset_tensors()
    compare_step(co, env, "reset", acc, float_tol, margins=co.detail.copy() if margin_aware else None)

    for t in range(steps):
        collect = rng.random(E) < p_collect
        actions = np.where(collect, 4, rng.integers(0, 4, size=E)).astype(np.int32)
        if philox:
            env.step_tensors(torch.as_tensor(actions, device=env.device))
            dt = exported()
            if twin is not None:
                twin.step_tensors(torch.as_tensor(actions, device=env.device))
                assert torch.equal(twin.obs, env.obs) and torch.equal(twin.reward, env.reward) and \
                    torch.equal(twin.done, env.done), ("LEAN instantiation differs from the diagnostic one", t)
                acc["lean_twin_steps"] += 1
        else:
            dt = make_tables(rng, E, P, sigma)
            if auto_reset:
                fill_t = rng.uniform(0.20, 0.60, size=(E, P))
                dr_t = make_tables(rng, E, P, sigma)
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
            # a restarted env's reset observation makes compares too: fold its margins into the step's
            for i in _MARGIN_IDX.values():
                o_detail[ids, i] = np.minimum(o_detail[ids, i], co.detail[ids, i])
        if not philox:
            env.step_tensors(torch.as_tensor(actions, device=env.device))
        g_rew = env.reward.cpu().numpy()
        g_done = env.done.cpu().numpy()
        det = env.step_detail.cpu().numpy()
        acc["steps"] += E
        acc["dones"] += int(o_done.sum())
        extra_disc = [("done", g_done, o_done)] + [(name, det[:, i], o_detail[:, i]) for i, name in _DETAIL_DISC]
        # rewards: against the scale of their terms (sum of |term|, oracle detail 19) in margin mode -- see
        # reward_metrics below; plain relative error in strict mode
        scale = np.maximum(o_detail[:, 19], 1e-300)
        acc["cum_scale"] = acc.get("cum_scale", np.zeros(E)) + scale
        if auto_reset and o_done.any():
            acc["cum_scale"][ids] = 0.0
        mb, aoi = float(params["max_buffer_size"]), float(params["max_buffer_size"]) / float(params["data_generation_rate"])
        extra_flt = [("reward", g_rew, o_rew, scale), ("base_reward", det[:, 12], o_detail[:, 12], scale),
                     ("bytes", det[:, 1], o_detail[:, 1], 0.01 * mb), ("urgency_reduced", det[:, 2], o_detail[:, 2], 0.01 * aoi),
                     ("mean_urgency", det[:, 3], o_detail[:, 3], 0.01), ("data_loss", det[:, 7], o_detail[:, 7], 0.01 * mb)]
        if auto_reset and o_done.any():
            tob = np.zeros_like(o_term)
            tob[ids] = env.terminal_obs.cpu().numpy()[ids]
            extra_flt.append(("terminal_obs", tob, np.where(o_done[:, None].astype(bool), o_term, 0.0), 0.01))
        acc["collisions"] += int(o_detail[:, 0].sum())
        acc["bytes"] += float(o_detail[:, 1].sum())
        full = (t % compare_every) == 0 or t == steps - 1
        compare_step(co, env, t, acc, float_tol, extra_disc, extra_flt, margins=o_detail if margin_aware else None,
                     state=full)
        # reward metrics over the environments still compared:
        #   reward_abs_over_scale : |r - r_ref| / sum |term|            (forward error of the sum)
        #   reward_rel_uncancelled: |r - r_ref| / |r_ref| over the rewards that keep at least half of their term
        #                           mass (|r_ref| >= scale / 2): the plain relative error where it is meaningful
        keep = ~acc["diverged"]
        if keep.any():
            err = np.abs(g_rew - o_rew)
            acc["reward_abs_over_scale"] = max(acc["reward_abs_over_scale"], float(np.max((err / scale)[keep])))
            unc = keep & (np.abs(o_rew) >= 0.5 * scale)
            if unc.any():
                acc["reward_rel_uncancelled"] = max(acc.get("reward_rel_uncancelled", 0.0),
                                                    float(np.max(err[unc] / np.abs(o_rew[unc]))))
            acc["rewards_uncancelled"] = acc.get("rewards_uncancelled", 0) + int(unc.sum())
            acc["rewards_compared"] = acc.get("rewards_compared", 0) + int(keep.sum())
            rel = np.where(keep, err / np.maximum(np.abs(o_rew), 1e-300), 0.0)
            w = int(np.argmax(rel))
            acc.setdefault("reward_worst", []).append((float(rel[w]), float(o_rew[w]), float(g_rew[w] - o_rew[w]),
                                                       float(scale[w])))
            acc["reward_worst"] = sorted(acc["reward_worst"], reverse=True)[:3]
    acc["captures"] = int(co.envs["capture_triggers"].sum())
    acc["first_bad"] = acc["first_bad"][:8]
    acc["flip_margins"] = acc["flip_margins"][:16]
    acc["diverged_envs"] = int(acc.pop("diverged").sum())
    acc.pop("cum_scale", None)
    env.poll_errors()
    env.close()
    if twin is not None:
        twin.close()
    return acc
