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


def test_philox_trajectory_replayed_through_the_oracle(pkg):
    """Kernel in Philox mode -> export the draws it used -> C oracle replays them: the float64 Philox trajectory is
    reproduced bit for bit.  (The float32 builds go through the same loop with the margin-aware checker:
    tests/test_gpu_parity.py::test_float32_throughput_build_against_the_oracle, draws="philox".)"""
    from parity_util import run_parity_case
    res = run_parity_case(pkg, E=512, N=24, grid=(300, 300), steps=40, seed=11, p_collect=0.5, precision="f64",
                          start=(150.0, 150.0), draws="philox", philox_seed=2024)
    assert res["discrete_mismatches"] == 0, res
    assert res["max_rel_err"] <= 1e-9, res
    assert res["bytes"] > 0
