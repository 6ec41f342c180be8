"""Size-independent properties checked at BASELINE.json's full sizes (where the oracle would take hours):
data conservation per sensor, range invariants, snapshot/restore round trip, shard invariance as a checksum
of checksums, statistics all-reduce payload consistency."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from __graft_entry__ import import_package
    return import_package()


def _run(env, steps, seed, p_collect=0.3):
    import torch
    g = torch.Generator(device=env.device)
    g.manual_seed(seed)
    E = env.num_envs
    for _ in range(steps):
        u = torch.rand(E, device=env.device, generator=g)
        a = torch.randint(0, 4, (E,), device=env.device, generator=g, dtype=torch.int32)
        a = torch.where(u < p_collect, torch.full_like(a, 4), a)
        env.step_tensors(a)


def test_config5_one_million_envs_conservation_and_ranges(pkg):
    """configs[4]: 1000x1000, N=50, 1,048,576 envs.  Every byte a sensor generated is in its buffer, was
    transmitted, or was lost: generated == buffer + transmitted + lost (uav_env / iot_sensors bookkeeping)."""
    import torch
    E, N = 1 << 20, 50
    env = pkg.UAVVecEnv(E, grid_size=(1000, 1000), num_sensors=N, seed=5, action_mask=False, max_battery=20.0)
    obs = env.reset_tensors()
    _run(env, 150, seed=1)
    env.poll_errors()
    st = env.state()
    gen = env.generated()
    lhs = st["buffer"] + st["transmitted"] + st["lost"]
    err = ((lhs - gen).abs() / gen.clamp_min(1.0)).max().item()
    assert err < 2e-5, err                                    # float32 state
    assert float(st["buffer"].min()) >= 0.0 and float(st["buffer"].max()) <= 1000.0
    sf = st["spreading_factor"]
    assert int(sf.min()) >= 7 and int(sf.max()) <= 12 and not bool((sf == 8).any() | (sf == 10).any())
    assert bool(torch.isfinite(obs).all()) and float(obs.min()) >= 0.0 and float(obs.max()) <= 1.0
    rec = env.records()
    assert (rec["battery"] > 0).all() and (rec["battery"] <= 20.0).all()
    assert (rec["x"] >= 0).all() and (rec["x"] < 1000).all() and (rec["y"] >= 0).all() and (rec["y"] < 1000).all()
    assert (rec["step"] >= 0).all() and (rec["step"] <= 150).all()      # 0: restarted by this very step
    # with a 20 Wh battery every env has finished at least one episode: the all-reduce payload says so
    sums = env.stat_sums.cpu().numpy()
    assert sums[-1] >= 0.99 * E and sums[-1] == (rec["episode"].astype(np.int64) - 1).sum()   # episodes finished, counted two ways
    mean_len = sums[1] / sums[-1]
    assert 90 < mean_len < 135, mean_len                      # 19.6 Wh / (0.139..0.194 Wh per step)
    env.close()


def test_config3_domain_rand_65536_envs_invariants(pkg):
    """configs[2]: 65,536 envs, grids from the curriculum list, N = 10..50 padded to 50 observation slots."""
    import torch
    E = 65536
    n = np.random.default_rng(0).integers(10, 51, size=E).astype(np.int32)
    env = pkg.UAVVecEnv(E, num_sensors=n, max_sensors=50, domain_randomisation=True, curriculum_stage=4, seed=9,
                        max_battery=10.0)
    obs = env.reset_tensors()
    assert obs.shape == (E, 153)
    _run(env, 120, seed=2)
    env.poll_errors()
    rec = env.records()
    assert (rec["num_sensors"] == n).all()
    assert set(np.unique(rec["grid_w"])) <= {100, 200, 300, 400, 500}
    obs = env.obs
    ghost = torch.arange(50, device=env.device)[None, :] >= torch.as_tensor(n, device=env.device)[:, None]
    slots = obs[:, 3:].view(E, 50, 3)
    assert float(slots[ghost].abs().max()) == 0.0            # ghost slots stay exactly zero (dqn.py:297-298)
    st = env.state()
    live = ~ghost
    lhs = (st["buffer"] + st["transmitted"] + st["lost"])[live]
    gen = env.generated()[live]
    assert ((lhs - gen).abs() / gen.clamp_min(1.0)).max().item() < 2e-5
    stats = env.episode_stats.cpu().numpy()
    fin = stats[:, 1] > 0
    assert fin.mean() > 0.99
    assert ((stats[fin, 6] >= 0) & (stats[fin, 6] <= 100)).all()                 # NDR %
    assert ((stats[fin, 8] > 0) & (stats[fin, 8] <= 1.0 + 1e-9)).all()           # Jain's index
    env.close()


def test_snapshot_restore_round_trip(pkg):
    import torch
    env = pkg.UAVVecEnv(4096, grid_size=(300, 300), num_sensors=30, seed=3, max_battery=8.0)
    env.reset_tensors()
    _run(env, 20, seed=3)
    snap = env.snapshot()
    _run(env, 25, seed=4)
    a = (env.obs.clone(), env.reward.clone(), env.done.clone(), env.snapshot()["arena"])
    env.restore(snap)
    _run(env, 25, seed=4)
    b = (env.obs, env.reward, env.done, env.snapshot()["arena"])
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    env.close()


def test_shard_invariance_checksum_of_checksums(pkg):
    """The whole job's per-step reward checksum equals the sum of its shards' checksums, bit for bit."""
    import torch
    E, T = 65536, 40
    acts = torch.randint(0, 5, (T, E), dtype=torch.int32, device="cuda:0", generator=torch.Generator("cuda:0").manual_seed(7))

    def checksums(count, offset):
        env = pkg.UAVVecEnv(count, grid_size=(500, 500), num_sensors=20, seed=77, env_id_offset=offset, max_battery=6.0)
        env.reset_tensors()
        out = []
        for t in range(T):
            o, r, d = env.step_tensors(acts[t, offset:offset + count])
            out.append((r.double().sum().item(), int(d.sum().item()), o.double().sum().item()))
        env.close()
        return np.array(out)
    whole = checksums(E, 0)
    parts = [checksums(E // 4, k * (E // 4)) for k in range(4)]
    assert (whole[:, 1] == sum(p[:, 1] for p in parts)).all()
    np.testing.assert_allclose(whole[:, 0], sum(p[:, 0] for p in parts), rtol=1e-12)
    np.testing.assert_allclose(whole[:, 2], sum(p[:, 2] for p in parts), rtol=1e-12)


@pytest.mark.parametrize("N,stack", [(20, 1), (50, 1), (50, 4), (33, 1), (1, 1), (2, 1), (32, 1), (16, 4), (64, 1), (10, 1), (30, 4), (31, 1)])
def test_lean_instantiation_equals_the_diagnostic_one(pkg, N, stack):
    """The benchmarked kernel is the LEAN instantiation (optional features compiled out); the instantiation with
    parity diagnostics attached is the one the oracle-replay tests exercise.  Same seed, same actions: state,
    observations, rewards and episode statistics must be bit-identical between the two, through in-kernel
    restarts (short battery)."""
    import torch
    E, T = 4096 + 7, 260
    kw = dict(grid_size=(300, 300), num_sensors=N, seed=31, max_battery=14.0, n_stack=stack, precision="f32",
              rng="philox", uav_start_position=(140.0, 150.0))
    lean = pkg.UAVVecEnv(E, **kw)
    diag = pkg.UAVVecEnv(E, step_detail=True, **kw)          # step_detail bound => not LEAN
    lean.reset_tensors()
    diag.reset_tensors()
    g = torch.Generator(device=lean.device)
    g.manual_seed(5)
    dones = 0
    for t in range(T):
        a = torch.randint(0, 5, (E,), dtype=torch.int32, device=lean.device, generator=g)
        a = torch.where(torch.rand(E, device=lean.device, generator=g) < 0.35, torch.full_like(a, 4), a)
        lean.step_tensors(a)
        diag.step_tensors(a)
        assert torch.equal(lean.reward, diag.reward), t
        assert torch.equal(lean.done, diag.done), t
        dones += int(lean.done.sum().item())
        if t % 20 == 0 or t == T - 1:
            assert torch.equal(lean.obs, diag.obs), t
            sl, sd = lean.state(), diag.state()
            for k in sl:
                assert torch.equal(sl[k], sd[k]) or (sl[k].is_floating_point() and torch.equal(torch.nan_to_num(sl[k]), torch.nan_to_num(sd[k]))), (t, k)
    assert dones > E // 2                                     # restarts happened inside the launches
    # float64 atomics from different warps: the order of additions is not reproducible, the sums agree to rounding
    assert torch.allclose(lean.stat_sums, diag.stat_sums, rtol=1e-12, atol=0.0)
    assert lean.stat_sums[-1] == diag.stat_sums[-1] == dones
    rl, rd = lean.records(), diag.records()
    for k in rl.dtype.names:
        assert (rl[k] == rd[k]).all() or np.array_equal(rl[k], rd[k], equal_nan=True), k
    lean.close()
    diag.close()


def test_pair_kernel_with_per_env_sensor_counts(pkg):
    """Two environments of a pair may have different sensor counts (DQN workers are pinned to 10/20/30, dqn.py:1066):
    the pair kernel against the one-slot diagnostic kernel, bit for bit."""
    import torch
    E, T = 2051, 120
    ns = np.resize(np.array([10, 20, 30, 7, 32, 1], dtype=np.int32), E)
    kw = dict(grid_size=(250, 250), num_sensors=ns, max_sensors=32, seed=77, max_battery=9.0, precision="f32", rng="philox",
              uav_start_position=(120.0, 130.0))
    lean = pkg.UAVVecEnv(E, **kw)
    diag = pkg.UAVVecEnv(E, step_detail=True, **kw)
    lean.reset_tensors()
    diag.reset_tensors()
    g = torch.Generator(device=lean.device)
    g.manual_seed(9)
    for t in range(T):
        a = torch.randint(0, 5, (E,), dtype=torch.int32, device=lean.device, generator=g)
        a = torch.where(torch.rand(E, device=lean.device, generator=g) < 0.4, torch.full_like(a, 4), a)
        lean.step_tensors(a)
        diag.step_tensors(a)
        assert torch.equal(lean.reward, diag.reward), t
        assert torch.equal(lean.obs, diag.obs), t
    assert int(lean.stat_sums[-1].item()) > E // 2
    sl, sd = lean.state(), diag.state()
    for k in sl:
        assert torch.equal(torch.nan_to_num(sl[k].double()), torch.nan_to_num(sd[k].double())), k
    lean.close()
    diag.close()


@pytest.mark.parametrize("max_sensors,stack", [(50, 4), (32, 1), (32, 4), (20, 1)])
def test_lean_domain_rand_instantiation_equals_the_generic_one(pkg, max_sensors, stack):
    """The DomainRandEnv training path (dqn.py:177-451) has its own LEAN instantiation; with parity diagnostics bound
    the generic kernel runs.  Same seed, same actions, in-kernel restarts with per-episode grids / layouts / far
    starts: observations, rewards, state and episode statistics must agree bit for bit."""
    import torch
    E, T = 3001, 150
    rng = np.random.default_rng(5)
    ns = rng.integers(10, max_sensors + 1, size=E).astype(np.int32)
    kw = dict(num_sensors=ns, max_sensors=max_sensors, obs_slots=50, domain_randomisation=True, curriculum_stage=4, seed=41,
              max_battery=12.0, n_stack=stack, precision="f32", rng="philox")
    lean = pkg.UAVVecEnv(E, **kw)
    diag = pkg.UAVVecEnv(E, step_detail=True, **kw)
    lean.reset_tensors()
    diag.reset_tensors()
    g = torch.Generator(device=lean.device)
    g.manual_seed(6)
    for t in range(T):
        a = torch.randint(0, 5, (E,), dtype=torch.int32, device=lean.device, generator=g)
        a = torch.where(torch.rand(E, device=lean.device, generator=g) < 0.35, torch.full_like(a, 4), a)
        lean.step_tensors(a)
        diag.step_tensors(a)
        assert torch.equal(lean.reward, diag.reward), t
        assert torch.equal(lean.done, diag.done), t
        if t % 10 == 0 or t == T - 1:
            assert torch.equal(lean.obs, diag.obs), t
    assert int(lean.stat_sums[-1].item()) > E // 2
    assert torch.allclose(lean.stat_sums, diag.stat_sums, rtol=1e-12, atol=0.0)
    sl, sd = lean.state(), diag.state()
    for k in sl:
        assert torch.equal(torch.nan_to_num(sl[k].double()), torch.nan_to_num(sd[k].double())), k
    rl, rd = lean.records(), diag.records()
    for k in rl.dtype.names:
        assert np.array_equal(rl[k], rd[k], equal_nan=True), k
    lean.close()
    diag.close()


@pytest.mark.parametrize("E", [1, 2, 3, 5, 17, 31])
def test_pair_kernel_tiny_batches(pkg, E):
    """Pair kernel with fewer environments than a tile, odd counts (an unpaired last environment) included."""
    import torch
    kw = dict(grid_size=(150, 150), num_sensors=20, seed=13, max_battery=6.0, precision="f32", rng="philox",
              uav_start_position=(70.0, 80.0))
    lean = pkg.UAVVecEnv(E, **kw)
    diag = pkg.UAVVecEnv(E, step_detail=True, **kw)
    lean.reset_tensors()
    diag.reset_tensors()
    g = torch.Generator(device=lean.device)
    g.manual_seed(E)
    for t in range(80):
        a = torch.randint(0, 5, (E,), dtype=torch.int32, device=lean.device, generator=g)
        lean.step_tensors(a)
        diag.step_tensors(a)
        assert torch.equal(lean.reward, diag.reward) and torch.equal(lean.obs, diag.obs) and torch.equal(lean.done, diag.done), t
    assert int(lean.stat_sums[-1].item()) >= E
    lean.close()
    diag.close()


@pytest.mark.parametrize("N", [20, 35])
def test_lean_relational_instantiation_equals_the_generic_one(pkg, N):
    """RelationalUAVEnv rows (env_wrapper.py:110-313) have a LEAN instantiation too; against the generic kernel."""
    import torch
    E, T = 2049, 140
    kw = dict(grid_size=(200, 200), num_sensors=N, seed=51, max_battery=10.0, relational=True, precision="f32", rng="philox",
              uav_start_position=(90.0, 110.0))
    lean = pkg.UAVVecEnv(E, **kw)
    diag = pkg.UAVVecEnv(E, step_detail=True, **kw)
    lean.reset_tensors()
    diag.reset_tensors()
    g = torch.Generator(device=lean.device)
    g.manual_seed(8)
    for t in range(T):
        a = torch.randint(0, 5, (E,), dtype=torch.int32, device=lean.device, generator=g)
        a = torch.where(torch.rand(E, device=lean.device, generator=g) < 0.5, torch.full_like(a, 4), a)
        lean.step_tensors(a)
        diag.step_tensors(a)
        assert torch.equal(lean.reward, diag.reward), t
        assert torch.equal(lean.obs, diag.obs), t
    assert int(lean.stat_sums[-1].item()) > E // 2
    rl, rd = lean.records(), diag.records()
    for k in rl.dtype.names:
        assert np.array_equal(rl[k], rd[k], equal_nan=True), k
    lean.close()
    diag.close()
