"""Philox mode on the GPU: the draws the kernels generate are (a) exactly the counter layout
DESIGN.md documents (checked against an independent numpy Philox4x32-10 + Box-Muller), (b) statistically
N(0, sigma) / U(0,1), and (c) when exported and replayed through the C oracle reproduce the kernel's
trajectory -- which ties the throughput configuration (Philox draws, in-kernel control flow) to the same
oracle as the replay-mode parity tests."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c = [np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ k0) & mask, lo1, (hi0 ^ c[3] ^ k1) & mask, lo0]
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return c


def box_muller(a, b, sigma):
    u1 = (a.astype(np.float64) + 0.5) * 2.0 ** -32
    u2 = (b.astype(np.float64) + 0.5) * 2.0 ** -32
    r = sigma * np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)


@pytest.fixture(scope="module")
def pkg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from __graft_entry__ import import_package
    return import_package()


@pytest.mark.parametrize("N", [20, 50])
def test_exported_draws_match_independent_philox(pkg, N):
    import torch
    E, seed, off = 96, 0xC0FFEE1234, 1000
    env = pkg.UAVVecEnv(E, grid_size=(400, 400), num_sensors=N, precision="f64", rng="philox", seed=seed,
                        env_id_offset=off, auto_reset=False, export_draws=True)
    env.reset_tensors()
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    lanes = np.arange(32, dtype=np.uint64)
    envs = (np.arange(E, dtype=np.uint64) + np.uint64(off))[:, None]
    for t, a in enumerate([4, 0, 4, 3, 4], start=1):
        env.step_tensors(torch.full((E,), a, dtype=torch.int32, device=env.device))
        exp = env.draw_export.cpu().numpy()             # [E, 6, N]
        episode = np.zeros((E, 32), dtype=np.uint64)

        def call(c):
            c3 = (np.uint64(t) << np.uint64(8)) | (lanes[None, :] << np.uint64(3)) | np.uint64(c)
            return philox4x32_10(np.broadcast_to(envs, (E, 32)), np.zeros((E, 32), dtype=np.uint64), episode,
                                 np.broadcast_to(c3, (E, 32)), k0, k1)
        q0 = call(0)
        d0, e0 = box_muller(q0[0], q0[1], 4.0)
        if N <= 32:
            # D, E of a step: one call per environment PAIR (stream 6), clocked by the even environment's episode /
            # step; words (x, y) -> even environment, (z, w) -> odd environment
            pair = np.broadcast_to(envs >> np.uint64(1), (E, 32))
            c3 = (np.uint64(t) << np.uint64(8)) | (lanes[None, :] << np.uint64(3))
            qp = philox4x32_10(pair, np.full((E, 32), 6 << 24, dtype=np.uint64), episode, np.broadcast_to(c3, (E, 32)), k0, k1)
            odd = (envs & np.uint64(1)).astype(bool)
            d0, e0 = box_muller(np.where(odd, qp[2], qp[0]), np.where(odd, qp[3], qp[1]), 4.0)
            np.testing.assert_allclose(exp[:, 4, :], d0[:, :N], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(exp[:, 5, :], e0[:, :N], rtol=1e-12, atol=1e-12)
            if a == 4:
                q1 = call(1)
                a0, b0 = box_muller(q1[0], q1[1], 4.0)
                np.testing.assert_allclose(exp[:, 0, :], a0[:, :N], rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(exp[:, 1, :], b0[:, :N], rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(exp[:, 2, :], ((q1[2].astype(np.float64) + 0.5) * 2.0 ** -32)[:, :N], rtol=1e-15)
        else:
            d1, e1 = box_muller(q0[2], q0[3], 4.0)
            np.testing.assert_allclose(exp[:, 4, :], np.concatenate([d0, d1], 1)[:, :N], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(exp[:, 5, :], np.concatenate([e0, e1], 1)[:, :N], rtol=1e-12, atol=1e-12)
            if a == 4:
                q1, q2 = call(1), call(2)
                a0, b0 = box_muller(q1[0], q1[1], 4.0)
                a1, b1 = box_muller(q1[2], q1[3], 4.0)
                np.testing.assert_allclose(exp[:, 0, :], np.concatenate([a0, a1], 1)[:, :N], rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(exp[:, 1, :], np.concatenate([b0, b1], 1)[:, :N], rtol=1e-12, atol=1e-12)
                u = np.concatenate([q2[0], q2[1]], 1).astype(np.float64)
                np.testing.assert_allclose(exp[:, 2, :], ((u + 0.5) * 2.0 ** -32)[:, :N], rtol=1e-15)
        if a != 4:
            assert np.isnan(exp[:, :4, :]).all()          # move steps draw only D and E
    env.close()


def test_float32_draw_statistics(pkg):
    import torch
    E, N = 8192, 50
    env = pkg.UAVVecEnv(E, grid_size=(1000, 1000), num_sensors=N, precision="f32", rng="philox", seed=7,
                        auto_reset=False, export_draws=True)
    env.reset_tensors()
    env.step_tensors(torch.full((E,), 4, dtype=torch.int32, device=env.device))
    exp = env.draw_export.cpu().numpy().astype(np.float64)
    for slot in (0, 1, 4, 5):
        x = exp[:, slot, :].ravel()
        assert abs(x.mean()) < 0.03 and abs(x.std() - 4.0) < 0.03, (slot, x.mean(), x.std())
        assert abs(np.mean(x ** 4) / x.var() ** 2 - 3.0) < 0.05          # kurtosis of a normal
    u = exp[:, 2, :].ravel()
    assert 0.0 < u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.003 and abs(u.var() - 1 / 12) < 0.002
    # slots are mutually uncorrelated
    c = np.corrcoef(np.stack([exp[:, s, :].ravel() for s in (0, 1, 2, 4, 5)]))
    assert np.abs(c - np.eye(5)).max() < 0.01
    env.close()


@pytest.mark.parametrize("precision,tol", [("f64", 1e-9), ("f32", None)])
def test_philox_trajectory_replayed_through_the_oracle(pkg, precision, tol):
    """Kernel in Philox mode -> export the draws it used -> C oracle replays them."""
    import torch
    from parity_util import compare_step      # puts oracle/ on sys.path
    import c_oracle
    import uavsim_oracle as O
    E, N, T = 512, 24, 40
    rng = np.random.default_rng(11)
    env = pkg.UAVVecEnv(E, grid_size=(300, 300), num_sensors=N, precision=precision, rng="philox", seed=2024,
                        auto_reset=False, export_draws=True, step_detail=True, uav_start_position=(150.0, 150.0))
    co = c_oracle.COracleBatch(O.default_params(), E=E, P=N)
    env.reset_tensors()
    exp = env.draw_export.cpu().numpy().astype(np.float64)
    xy = env.state()["sensor_pos"].cpu().numpy()
    co.set_scenario(np.tile([300, 300], (E, 1)), np.full(E, N), xy, np.tile([150.0, 150.0], (E, 1)))
    fill = exp[:, 2, :].copy()                     # reset exports the fill it used in slot U
    co.reset(fill, np.nan_to_num(exp))
    acc = dict(discrete_mismatches=0, max_rel_err=0.0, per_key={}, first_bad=[])
    compare_step(co, env, "reset", acc)
    for t in range(T):
        a = np.where(rng.random(E) < 0.5, 4, rng.integers(0, 4, E)).astype(np.int32)
        env.step_tensors(torch.as_tensor(a, device=env.device))
        exp = env.draw_export.cpu().numpy().astype(np.float64)
        co.step(a, np.nan_to_num(exp))
        if precision == "f64":
            compare_step(co, env, t, acc)
        else:                                       # float32 state drifts from the float64 oracle: resync it
            g = env.state()
            rec = env.records()
            sf_bad = int((g["spreading_factor"].cpu().numpy()[:, :N] != co.envs["sf"][:, :N]).sum())
            acc["discrete_mismatches"] += sf_bad
            gb, ob = g["buffer"].cpu().numpy(), co.envs["buf"][:, :N]
            off = np.abs(gb - ob) > 1e-4 * np.maximum(np.abs(ob), 1.0) + 1e-2
            assert off.mean() < 0.005, (t, off.mean())          # float32 state: rare threshold flips only
            np.testing.assert_allclose(rec["battery"], co.envs["battery"], rtol=1e-9)
    if precision == "f64":
        assert acc["discrete_mismatches"] == 0, acc
        assert acc["max_rel_err"] <= tol, acc
    else:
        assert acc["discrete_mismatches"] <= 0.002 * E * N * T, acc
    env.close()
