"""Host-side mirror of the reference interface for the batched step.

`UAVVecEnv` stands where the reference has

    DummyVecEnv([lambda: Monitor(DomainRandEnv(...))] * 4)      # dqn.py:1223-1234, 1276
    UAVEnvironment(grid_size=..., num_sensors=..., **ENV_BASE)  # sweep_eval.py:462

i.e. it keeps the constructor keywords of `UAVEnvironment.__init__` (uav_env.py:266-287)
and of `DomainRandEnv.__init__` (dqn.py:202-232), the Gymnasium reset/step/observation/
reward semantics, and the stable-baselines3 2.7 `VecEnv` surface the reference's call sites
use: `reset() -> obs`, `step(actions) -> (obs, rewards, dones, infos)`
with numpy arrays, `step_async/step_wait`, auto-reset with `terminal_observation` /
`TimeLimit.truncated` / Monitor's `episode` entry, `get_attr("last_episode_stats")`,
`env_method("set_curriculum_stage", s)`, `seed`, `close`.

All simulation state lives in one torch-allocated device arena that libuavsim.so borrows;
`reset_tensors()/step_tensors()` return the bound CUDA tensors themselves (zero-copy), the
numpy methods add exactly one device->host copy per output.  There is no CPU implementation
behind this class: without a CUDA device construction fails.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from .spaces import Box, Discrete

try:  # the reference trains through stable-baselines3 (dqn.py:1276-1288): be a real VecEnv when it is importable, so
    # that DQN._wrap_env / VecFrameStack take the batch as it is instead of re-wrapping it in a DummyVecEnv
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase  # type: ignore
    HAVE_SB3 = True
except Exception:  # noqa: BLE001 -- SB3 is absent from the build image: same surface, duck-typed
    _VecEnvBase = object
    HAVE_SB3 = False

# dqn.py:63-69 -- grids per curriculum stage (sensor counts are fixed per worker)
CURRICULUM_STAGES = [
    [(100, 100)],
    [(100, 100), (200, 200)],
    [(100, 100), (200, 200), (300, 300)],
    [(100, 100), (200, 200), (300, 300), (400, 400)],
    [(100, 100), (200, 200), (300, 300), (400, 400), (500, 500)],
]
MAX_SENSORS_LIMIT = 50  # dqn.py:121

_INFO_KEYS = ("uav_position", "battery", "battery_percent", "sensors_collected", "current_step",
              "total_reward", "total_data_collected", "coverage_percentage", "is_alive", "max_urgency",
              "avg_urgency", "high_urgency_sensors", "capture_effect_triggers", "boundary_hits",
              "edge_steps", "last_step_bytes_collected", "sensor_collection_ratios")


def numa_local_cpus(dev_index: int):
    """(CPUs of the NUMA node the GPU hangs off that this process may run on, node id) from sysfs; (None, node) when
    the topology cannot be read or the cpuset excludes that node."""
    import os
    from pathlib import Path
    try:
        pr = torch.cuda.get_device_properties(dev_index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(Path("/sys/bus/pci/devices/%s/numa_node" % bus).read_text())
        if node < 0:
            return None, node
        cpus = set()
        for part in Path("/sys/devices/system/node/node%d/cpulist" % node).read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        return (cpus or None), node
    except Exception:  # noqa: BLE001
        return None, None


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class UAVVecEnv(_VecEnvBase):
    """E batched UAV-LoRa environments on one GPU.  A `stable_baselines3.common.vec_env.VecEnv` subclass when SB3 is
    importable (every abstract method of the 2.7 interface is implemented: reset, step_async, step_wait, close,
    get_attr, set_attr, env_method, env_is_wrapped -- plus seed / step), the same surface duck-typed otherwise.

    Reference constructor keywords keep their names and meaning (uav_env.py:266-287):
    grid_size, sensor_positions, num_sensors, data_generation_rate, max_buffer_size,
    rssi_threshold, sensor_duty_cycle, uav_start_position, max_battery, collection_duration,
    max_steps, penalty_data_loss, reward_urgency_reduction, penalty_battery, reward_movement,
    include_sensor_positions.  `lora_spreading_factor` and `path_loss_exponent` are accepted and
    ignored exactly as the reference ignores them (iot_sensors.py:309 forces SF12 on reset,
    calculate_rssi never reads the exponent).
    """

    def __init__(self, num_envs: int, grid_size=(500, 500), num_sensors=20, *,
                 sensor_positions=None, data_generation_rate=22.0 / 10, max_buffer_size=1000.0,
                 lora_spreading_factor=7, path_loss_exponent=2.0, rssi_threshold=-85.0,
                 sensor_duty_cycle=10.0, uav_start_position=None, max_battery=274.0,
                 collection_duration=1.0, max_steps=2100, render_mode=None,
                 penalty_data_loss=-1.0, reward_urgency_reduction=20.0, penalty_battery=-0.5,
                 reward_movement=10.0, include_sensor_positions=False,
                 # batched-simulator arguments
                 max_sensors: Optional[int] = None, obs_slots: Optional[int] = None, n_stack: int = 1,
                 domain_randomisation=False, curriculum_stage=0, relational=False,
                 gps_sigma=0.0, wind_drift=0.0, latency_steps=0,
                 shadowing_std_db=4.0, gps_in_observation=False, wind_extra_step=0.0,
                 precision="f32", rng="philox", seed=0, env_id_offset=0, auto_reset=True,
                 device="cuda:0", step_detail=False, export_draws=False, action_mask=True,
                 terminal_info: Optional[bool] = None, info_every_step=False):
        if render_mode is not None:
            raise NotImplementedError("rendering is out of scope (uav_env.py:706-888)")
        if not torch.cuda.is_available():
            raise RuntimeError("UAVVecEnv needs a CUDA device: the batched step has no CPU fallback")
        self.lib = N.load()
        self.device = torch.device(device)
        self.num_envs = int(num_envs)
        ns = np.broadcast_to(np.asarray(num_sensors, dtype=np.int32), (self.num_envs,)).copy()
        self._num_sensors = ns
        P = int(max_sensors if max_sensors is not None else ns.max())
        cfg = N.default_config()
        cfg.num_envs = self.num_envs
        cfg.max_sensors = P
        cfg.obs_slots = int(obs_slots if obs_slots is not None else
                            (MAX_SENSORS_LIMIT if domain_randomisation and P <= MAX_SENSORS_LIMIT else P))
        # RelationalUAVEnv (env_wrapper.py:110-161) forces include_sensor_positions=True and pads to n_max = 50
        self.relational = bool(relational)
        if relational:
            include_sensor_positions = True
            cfg.wrapper = N.WRAP_RELATIONAL
            if obs_slots is None and P <= MAX_SENSORS_LIMIT:
                cfg.obs_slots = MAX_SENSORS_LIMIT
        cfg.features_per_sensor = 5 if include_sensor_positions else 3
        cfg.frame_stack = int(n_stack)          # VecFrameStack(n_stack) on device (dqn.py:1098,1278)
        cfg.precision = {"f64": N.F64, "f32": N.F32}[precision]
        cfg.rng_mode = {"replay": N.RNG_REPLAY, "philox": N.RNG_PHILOX}[rng]
        cfg.reset_mode = N.RESET_DOMAIN_RAND if domain_randomisation else N.RESET_FIXED
        cfg.auto_reset = int(bool(auto_reset))
        cfg.seed = int(seed) & (2 ** 64 - 1)
        cfg.env_id_offset = int(env_id_offset)
        grids = np.asarray(grid_size, dtype=np.int32)
        per_env_grid = grids.ndim == 2
        g0 = grids[0] if per_env_grid else grids
        if domain_randomisation:
            g0 = np.asarray(CURRICULUM_STAGES[0][0], dtype=np.int32)   # dqn.py:216-218 init grid
        cfg.grid_w, cfg.grid_h = int(g0[0]), int(g0[1])
        cfg.num_sensors = int(ns[0])
        cfg.max_steps = int(max_steps)
        cfg.data_generation_rate = float(data_generation_rate)
        cfg.max_buffer_size = float(max_buffer_size)
        cfg.rssi_threshold = float(rssi_threshold)
        cfg.duty_cycle = float(sensor_duty_cycle)
        cfg.max_battery = float(max_battery)
        cfg.collection_duration = float(collection_duration)
        start = (0.0, 0.0) if uav_start_position is None else uav_start_position
        starts = np.asarray(start, dtype=np.float32)
        per_env_start = starts.ndim == 2
        s0 = starts[0] if per_env_start else starts
        cfg.start_x, cfg.start_y = float(s0[0]), float(s0[1])
        cfg.penalty_data_loss = float(penalty_data_loss)
        cfg.reward_urgency_reduction = float(reward_urgency_reduction)
        cfg.penalty_battery = float(penalty_battery)
        cfg.reward_movement = float(reward_movement)
        cfg.gps_sigma = float(gps_sigma)      # sim_to_real_analysis.py:170-189
        cfg.wind_drift = float(wind_drift)    # sim_to_real_analysis.py:192-224
        cfg.latency_steps = int(latency_steps)  # sim_to_real_analysis.py:227-245
        # PerturbedDQNEnv of the sim-to-real sweep (sim_to_real_sweep.py:107-139): shadowing sigma of every sensor, GPS noise
        # inside the observation pass, and (host side, `wind_extra_step`) the probabilistic repeat of a move
        cfg.shadowing_std_db = float(shadowing_std_db)
        cfg.gps_in_observation = int(bool(gps_in_observation))
        self.wind_extra_step = float(wind_extra_step)
        self._wind_rng = None
        self.latency_steps = int(latency_steps)
        self._stage = -1
        self.cfg = cfg
        self.domain_randomisation = bool(domain_randomisation)
        if domain_randomisation:
            self._apply_stage(cfg, curriculum_stage)
        self.dims = N.query_dims(cfg)
        self.precision = precision
        self.real_dtype = torch.float64 if precision == "f64" else torch.float32

        E, dev = self.num_envs, self.device
        with torch.cuda.device(dev):
            self._arena = torch.empty(self.dims.arena_bytes, dtype=torch.uint8, device=dev)
            assert self._arena.data_ptr() % 256 == 0
            h = C.c_void_p()
            N.check(self.lib.uavsim_create(C.byref(cfg), _ptr(self._arena), self._arena.numel(),
                                           self._stream(), C.byref(h)))
            self._h = h
            self.n_stack = int(n_stack)
            self.stacked_dim = self.n_stack * self.dims.obs_dim
            # frame block per env: one row, or the 2k-row mirrored ring (include/uavsim.h, uavsim_outputs.obs)
            self._obs_buf = torch.empty((E, self.dims.obs_row_floats), dtype=torch.float32, device=dev)
            self.reward = torch.empty((E,), dtype=self.real_dtype, device=dev)
            self.done = torch.zeros((E,), dtype=torch.uint8, device=dev)
            self.terminal_obs = torch.zeros((E, self.stacked_dim), dtype=torch.float32, device=dev)
            self.episode_stats = torch.zeros((E, N.STAT_COUNT), dtype=torch.float64, device=dev)
            self.action_mask = torch.zeros((E,), dtype=torch.uint8, device=dev) if action_mask else None
            self.step_detail = (torch.zeros((E, N.DETAIL_COUNT), dtype=torch.float64, device=dev)
                                if step_detail else None)
            self.draw_export = (torch.full((E, N.NUM_SLOTS, P), float("nan"), dtype=self.real_dtype, device=dev)
                                if export_draws else None)
            # _get_info() of the terminal state of every finished episode (what callers read from infos[i] at episode
            # end: dqn.py:1148-1155, sim_to_real_sweep.py:222-225).  8 * (18 + P) bytes per env, so on by default only
            # for batches a host loop can consume; the tensor API of million-env batches reads `episode_stats`.
            if terminal_info is None:
                terminal_info = E <= 65536
            self.terminal_info = torch.zeros((E, N.INFO_COUNT), dtype=torch.float64, device=dev) if terminal_info else None
            self.terminal_ratios = torch.zeros((E, P), dtype=torch.float64, device=dev) if terminal_info else None
            self.info_every_step = bool(info_every_step)
            out = N.Outputs(_ptr(self._obs_buf), _ptr(self.reward), _ptr(self.done), _ptr(self.terminal_obs),
                            _ptr(self.episode_stats), _ptr(self.action_mask), _ptr(self.step_detail),
                            _ptr(self.draw_export), _ptr(self.terminal_info), _ptr(self.terminal_ratios))
            N.check(self.lib.uavsim_bind_outputs(self._h, C.byref(out), self._stream()))
            n_fields = 5 if precision == "f64" else 4      # f32: SF is packed into avg_rssi
            self._fields = {f: self._field_tensor(f) for f in range(n_fields)}
            sums = C.c_void_p()
            N.check(self.lib.uavsim_stat_sums(self._h, C.byref(sums)))
            off = sums.value - self._arena.data_ptr()
            self.stat_sums = self._arena[off: off + 8 * N.STAT_SUMS].view(torch.float64)
            # scenario: injected layout / per-env grids / starts, else Philox uniform layouts
            sc_xy = None
            if sensor_positions is not None:
                xy = np.asarray(sensor_positions, dtype=np.float32)
                if xy.ndim == 2:
                    xy = np.broadcast_to(xy[None], (E,) + xy.shape)
                full = np.zeros((E, P, 2), dtype=np.float32)
                full[:, : xy.shape[1]] = xy
                sc_xy = full
            self.set_scenario(sensor_xy=sc_xy,
                              num_sensors=ns if (ns != ns[0]).any() else None,
                              grid_wh=grids if per_env_grid else None,
                              start_xy=starts if per_env_start else None)
            if sensor_positions is None:
                N.check(self.lib.uavsim_generate_layouts(self._h, self._stream()))
        self._replay_keep = {}
        self._actions_dev = torch.zeros((E,), dtype=torch.int32, device=dev)
        self._pending_actions = None
        self._host = None
        self.last_episode_stats = [None] * E
        self.observation_space = Box(low=-np.inf if domain_randomisation else -1.0,
                                     high=np.inf if domain_randomisation else 1.0,
                                     shape=(self.stacked_dim,), dtype=np.float32)
        self.action_space = Discrete(N.NUM_ACTIONS)
        self.render_mode = None
        self.metadata = {"render_modes": []}
        self._closed = False
        self._np_host = None      # pinned staging buffers of the numpy VecEnv path (allocated on first use)
        self._np_infos = None
        self._np_dirty = []
        self._info_buf = None     # device rows of the on-demand _get_info kernel
        self.host_numa = None
        self._next_seed = None
        if HAVE_SB3:              # VecEnv.__init__ (sets num_envs / spaces, reset_infos, _seeds, _options, render_mode)
            _VecEnvBase.__init__(self, self.num_envs, self.observation_space, self.action_space)

    # ------------------------------------------------------------------ plumbing
    @property
    def obs(self) -> torch.Tensor:
        """The current (stacked) observation [E, n_stack*obs_dim] as a zero-copy view of the bound buffer.
        With n_stack > 1 it is the window of the k newest frames, oldest first (VecFrameStack order), of the
        mirrored frame ring; the view moves every step, so fetch it again after each step."""
        if self.n_stack + self.latency_steps == 1:
            return self._obs_buf
        s = int(self.lib.uavsim_frame_slot(self._h))
        od = self.dims.obs_dim
        return self._obs_buf.as_strided((self.num_envs, self.stacked_dim), (self.dims.obs_row_floats, 1),
                                        storage_offset=(max(s, 0) + 1) * od)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _field_tensor(self, field):
        v = N.FieldView()
        N.check(self.lib.uavsim_get_field(self._h, field, C.byref(v)))
        shape = [v.shape[i] for i in range(v.ndim)]
        nbytes = int(np.prod(shape)) * v.elem_bytes
        off = v.ptr - self._arena.data_ptr()
        raw = self._arena[off: off + nbytes]
        if v.is_float == 2:
            return raw.view(self.num_envs, v.elem_bytes)
        dt = {(1, 8): torch.float64, (1, 4): torch.float32, (0, 1): torch.uint8}[(v.is_float, v.elem_bytes)]
        return raw.view(dt).view(*shape)

    def _apply_stage(self, cfg, stage):
        stage = int(np.clip(stage, 0, len(CURRICULUM_STAGES) - 1))      # dqn.py:265
        grids = CURRICULUM_STAGES[stage]
        cfg.num_grids = len(grids)
        for i, (w, h) in enumerate(grids):
            cfg.grids[i][0], cfg.grids[i][1] = w, h
        self._stage = stage
        return grids

    # ------------------------------------------------------------------ scenario / replay
    def set_scenario(self, sensor_xy=None, num_sensors=None, grid_wh=None, start_xy=None):
        """Inject per-env layouts (the `sensor_positions=` / `uav_start_position=` ctor arguments)."""
        dev = self.device

        def up(a, dt):
            return None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev).contiguous()
        keep = [up(sensor_xy, torch.float32), up(num_sensors, torch.int32), up(grid_wh, torch.int32),
                up(start_xy, torch.float32)]
        if all(k is None for k in keep):
            return
        if keep[0] is not None:
            assert tuple(keep[0].shape) == (self.num_envs, self.cfg.max_sensors, 2)
        sc = N.Scenario(*[_ptr(k) for k in keep])
        with torch.cuda.device(dev):
            N.check(self.lib.uavsim_set_scenario(self._h, C.byref(sc), self._stream()))
            torch.cuda.current_stream(dev).synchronize()
        if num_sensors is not None:
            self._num_sensors = np.asarray(num_sensors, dtype=np.int32).copy()

    def set_replay(self, step_draws=None, reset_draws=None, reset_fill=None, reset_init_sf=None,
                   reset_sensor_xy=None, reset_num_sensors=None, reset_grid_wh=None, reset_start_xy=None,
                   dwell_draws=None, step_env_draws=None, reset_env_draws=None):
        """Bind slot-addressed draw tables (device float64 tensors or numpy arrays); see
        include/uavsim.h `uavsim_replay`.  Arguments left None keep their previous binding."""
        dev = self.device

        def up(name, a, dt):
            if a is None:
                return self._replay_keep.get(name)
            t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
            t = t.to(device=dev, dtype=dt).contiguous()
            self._replay_keep[name] = t
            return t
        E, P = self.num_envs, self.cfg.max_sensors
        sd = up("step_draws", step_draws, torch.float64)
        rd = up("reset_draws", reset_draws, torch.float64)
        rf = up("reset_fill", reset_fill, torch.float64)
        ri = up("reset_init_sf", reset_init_sf, torch.int32)
        rxy = up("rs_xy", reset_sensor_xy, torch.float32)
        rn = up("rs_n", reset_num_sensors, torch.int32)
        rg = up("rs_g", reset_grid_wh, torch.int32)
        rs = up("rs_s", reset_start_xy, torch.float32)
        dw = up("dwell", dwell_draws, torch.float64)
        assert dw is None or tuple(dw.shape) == (E, P)
        se = up("step_env", step_env_draws, torch.float64)
        re_ = up("reset_env", reset_env_draws, torch.float64)
        for t in (se, re_):
            assert t is None or tuple(t.shape) == (E, 4), "per-env draw tables are [E, 4]"
        for t in (sd, rd):
            assert t is None or tuple(t.shape) == (E, N.NUM_SLOTS, P), "draw tables are [E, 6, P]"
        assert rf is None or tuple(rf.shape) == (E, P)
        rp = N.Replay(_ptr(sd), _ptr(rd), _ptr(rf), _ptr(ri), N.Scenario(_ptr(rxy), _ptr(rn), _ptr(rg), _ptr(rs)),
                      _ptr(dw), _ptr(se), _ptr(re_))
        N.check(self.lib.uavsim_set_replay(self._h, C.byref(rp)))

    # ------------------------------------------------------------------ zero-copy tensor API
    def reset_tensors(self, env_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Gymnasium reset for all (or the masked) envs; returns the bound obs tensor [E, obs_dim]."""
        with torch.cuda.device(self.device):
            m = None
            if env_mask is not None:
                m = env_mask.to(device=self.device, dtype=torch.uint8).contiguous()
            N.check(self.lib.uavsim_reset(self._h, _ptr(m), self._stream()))
        return self.obs

    def step_tensors(self, actions: torch.Tensor):
        """One step for every env.  `actions`: int32 CUDA tensor [E] in [0, 4].  Returns the bound
        (obs, reward, done) CUDA tensors -- overwritten in place by the next call.  Asynchronous:
        invalid actions are reported by `poll_errors()`."""
        if actions.dtype != torch.int32 or actions.device != self.device or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.int32).contiguous()
        if self.wind_extra_step > 0 and self.cfg.rng_mode == N.RNG_PHILOX:
            self._wind_repeat(actions)
        with torch.cuda.device(self.device):
            N.check(self.lib.uavsim_step(self._h, _ptr(actions), self._stream()))
        return self.obs, self.reward, self.done

    def _wind_repeat(self, actions: torch.Tensor) -> None:
        """PerturbedDQNEnv.step (sim_to_real_sweep.py:117-119): with probability min(0.95, wind / 2) a move action is
        run once more first and that extra step's result is dropped."""
        if self._wind_rng is None:
            self._wind_rng = torch.Generator(device=self.device)
            self._wind_rng.manual_seed(int(self.cfg.seed) ^ 0x57494E44)
        p_extra = min(0.95, self.wind_extra_step / 2.0)
        mask = ((torch.rand(self.num_envs, device=self.device, generator=self._wind_rng) < p_extra) & (actions < 4)).to(torch.uint8)
        self.step_extra_tensors(actions, mask)

    def step_extra_tensors(self, actions: torch.Tensor, env_mask: torch.Tensor) -> None:
        """The dropped repeat itself (uavsim_step_ex with UAVSIM_STEP_EXTRA) for the envs whose mask byte is set; the
        commanded step must follow.  Replay-mode callers bind the repeat's own draw tables before this call."""
        m = env_mask.to(device=self.device, dtype=torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            N.check(self.lib.uavsim_step_ex(self._h, _ptr(actions), _ptr(m), N.STEP_EXTRA, self._stream()))

    def set_env_params(self, shadowing_std_db=None, duty_cycle=None) -> None:
        """Per-environment shadowing sigma (dB) / duty cycle (percent) columns (uavsim_set_env_params): one batch can
        hold a whole robustness sweep (sim_to_real_sweep.py:72-77 varies them per run).  None = the handle's value."""
        def col(a):
            return None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float32)).to(self.device).contiguous()
        self._env_sigma, self._env_duty = col(shadowing_std_db), col(duty_cycle)
        for t in (self._env_sigma, self._env_duty):
            assert t is None or tuple(t.shape) == (self.num_envs,)
        N.check(self.lib.uavsim_set_env_params(self._h, _ptr(self._env_sigma), _ptr(self._env_duty)))

    def make_host_buffers(self):
        """Pinned host buffers for `step_host` (actions in; obs, reward, done out), allocated NUMA-local to the GPU:
        the pages of a pinned allocation land on the node of the allocating thread, so the thread is moved onto the
        CPUs of the GPU's node for the allocation (when the cpuset allows it) -- with eight ranks on a two-socket host
        the copies of half the GPUs otherwise cross the socket interconnect.  `self.host_numa` records what was done."""
        import os
        E = self.num_envs
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        cpus, node = numa_local_cpus(idx)
        before = os.sched_getaffinity(0)
        moved = False
        if cpus and cpus != before:
            try:
                os.sched_setaffinity(0, cpus)
                moved = True
            except OSError:
                moved = False
        try:
            bufs = dict(actions=torch.zeros((E,), dtype=torch.int32).pin_memory(),
                        obs=torch.empty((E, self.stacked_dim), dtype=torch.float32).pin_memory(),
                        reward=torch.empty((E,), dtype=self.real_dtype).pin_memory(),
                        done=torch.empty((E,), dtype=torch.uint8).pin_memory())
            bufs["obs"].zero_()          # first touch on the local node
        finally:
            if moved:
                os.sched_setaffinity(0, before)
        self.host_numa = {"gpu_numa_node": node, "cpus_allowed": len(before),
                          "allocated_on_gpu_node": bool(moved or (cpus is not None and cpus == before))}
        return bufs

    def step_host(self, host: dict, copy_obs=True):
        """The end-to-end call with HOST buffers (uavsim_step_host): actions host->device, one step,
        obs/reward/done device->host, stream synchronised on return.  Actions are validated on the
        host first, as the reference raises ValueError before doing anything useful (uav_env.py:468)."""
        a = host["actions"]
        if bool(((a < 0) | (a > 4)).any()):
            raise ValueError("Invalid action")
        with torch.cuda.device(self.device):
            N.check(self.lib.uavsim_step_host(self._h, _ptr(a), _ptr(host["obs"]) if copy_obs else None,
                                              _ptr(host["reward"]), _ptr(host["done"]), self._stream()))
        return host["obs"], host["reward"], host["done"]

    def nearest_greedy_actions(self, probe_draws=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """NearestSensorGreedy.select_action (greedy_agents.py:73-100) for every env, on the device.  Returns an
        int32 CUDA tensor [E] ready for `step_tensors`.  `probe_draws` ([E, P] float64, NaN = not drawn) is the
        replay table of the agent's `is_in_range` probes; Philox mode draws its own."""
        return self._policy_actions(self.lib.uavsim_policy_nearest_greedy, probe_draws, out)

    def max_throughput_greedy_actions(self, probe_draws=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """MaxThroughputGreedyV2.select_action (greedy_agents.py:105-200) for every env, on the device; same
        conventions as `nearest_greedy_actions`."""
        return self._policy_actions(self.lib.uavsim_policy_max_throughput_greedy, probe_draws, out)

    def _policy_actions(self, entry, probe_draws, out):
        if out is None:
            out = self._actions_dev
        t = None
        if probe_draws is not None:
            t = probe_draws if isinstance(probe_draws, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(probe_draws))
            t = t.to(device=self.device, dtype=torch.float64).contiguous()
            assert tuple(t.shape) == (self.num_envs, self.cfg.max_sensors)
        with torch.cuda.device(self.device):
            N.check(entry(self._h, _ptr(out), _ptr(t), self._stream()))
        return out

    def relational_obs(self) -> dict:
        """The Dict observation of RelationalUAVEnv (env_wrapper.py:149-160, 200-228) as zero-copy views of the
        bound row: "uav" [E,3], "sensors" [E,n_max,5] = [rel_x, rel_y, buffer, urgency, visited], and the static
        "mask" [E,n_max] (1 = real sensor slot)."""
        if not self.relational:
            raise RuntimeError("constructed without relational=True")
        slots = int(self.cfg.obs_slots)
        row = self.obs[:, -self.dims.obs_dim:]          # newest frame when stacked
        if getattr(self, "_rel_mask", None) is None:
            n = torch.as_tensor(self._num_sensors, device=self.device)
            self._rel_mask = (torch.arange(slots, device=self.device)[None, :] < n[:, None]).float()
        return {"uav": row[:, :3], "sensors": row[:, 3:].unflatten(1, (slots, 5)), "mask": self._rel_mask}

    def poll_errors(self):
        """Raises ValueError if any step since the last poll saw an action outside [0, 4]
        (uav_env.py:468).  Synchronises the stream."""
        first = C.c_int64(-1)
        with torch.cuda.device(self.device):
            N.check(self.lib.uavsim_poll_error(self._h, self._stream(), C.byref(first)))

    @property
    def launch_count(self) -> int:
        return int(self.lib.uavsim_launch_count(self._h))

    def state(self) -> dict:
        """Views of the structure-of-arrays state in HBM (zero-copy, except that in the float32
        layout the spreading factor is decoded from the three low mantissa bits of avg_rssi)."""
        st = self._fields[N.FIELD_SENSOR_STATE]
        if self.precision == "f64":
            avg, sf = st[..., 3], self._fields[N.FIELD_SENSOR_SF]
        else:
            # throughput layout: avg_rssi is stored in dB above rssi_threshold (float64 here so that undoing the
            # shift does not round), with the spreading-factor code in its three low mantissa bits
            bits = st[..., 3].contiguous().view(torch.int32)
            sf = ((bits & 7) + 7).to(torch.uint8)
            avg = (bits & ~7).view(torch.float32).double() + float(self.cfg.rssi_threshold)
        return dict(records=self._fields[N.FIELD_ENV_RECORD], buffer=st[..., 0], transmitted=st[..., 1],
                    lost=st[..., 2], avg_rssi=avg, sensor_pos=self._fields[N.FIELD_SENSOR_POS],
                    generated_base=self._fields[N.FIELD_SENSOR_GEN], spreading_factor=sf)

    def records(self) -> np.ndarray:
        """Host copy of the per-env records as a numpy structured array (uavsim_env_record)."""
        raw = self._fields[N.FIELD_ENV_RECORD].cpu().numpy()
        return raw.view(np.dtype(N.ENV_RECORD_DTYPE)).reshape(self.num_envs)

    def set_battery(self, battery_wh: torch.Tensor) -> None:
        """Overwrite `uav.battery` (Wh, float64 [E]) of every env in place -- the record column is a zero-copy view.
        Evaluation from partial charge; bench.py uses it to desynchronise the episodes so that a timed region holds
        truncations and in-kernel restarts at their steady-state rate instead of none."""
        rec = self._fields[N.FIELD_ENV_RECORD].view(torch.float64)        # [E, 16]: battery is double #2 (offset 16)
        rec[:, 2] = battery_wh.to(device=self.device, dtype=torch.float64)

    def generated(self) -> torch.Tensor:
        """total_data_generated per sensor [E, P] (iot_sensors.py:116)."""
        g = self._fields[N.FIELD_SENSOR_GEN]
        if self.precision == "f64":
            return g
        el = torch.as_tensor(self.records()["elapsed_s"], device=self.device)
        return g + float(self.cfg.data_generation_rate) * el[:, None]

    def snapshot(self) -> dict:
        """Checkpoint of ALL simulation state: the arena (every SoA column, the records, the running statistic
        sums), the bound observation block with its frame-ring position (with n_stack > 1 or latency_steps > 0 the
        ring holds the k-1 (+latency) previous frames the next observations are built from) and the per-episode
        outputs."""
        snap = dict(arena=self._arena.clone(), obs=self._obs_buf.clone(),
                    frame_slot=int(self.lib.uavsim_frame_slot(self._h)), reward=self.reward.clone(),
                    done=self.done.clone(), terminal_obs=self.terminal_obs.clone(),
                    episode_stats=self.episode_stats.clone(), last_episode_stats=list(self.last_episode_stats))
        if self.action_mask is not None:
            snap["action_mask"] = self.action_mask.clone()
        if self.terminal_info is not None:
            snap["terminal_info"], snap["terminal_ratios"] = self.terminal_info.clone(), self.terminal_ratios.clone()
        return snap

    def restore(self, snap: dict):
        self._arena.copy_(snap["arena"])
        self._obs_buf.copy_(snap["obs"])
        for k in ("reward", "done", "terminal_obs", "episode_stats", "action_mask", "terminal_info", "terminal_ratios"):
            if k in snap and getattr(self, k) is not None:
                getattr(self, k).copy_(snap[k])
        self.last_episode_stats = list(snap["last_episode_stats"])
        N.check(self.lib.uavsim_set_frame_slot(self._h, int(snap["frame_slot"])))

    # ------------------------------------------------------------------ SB3 VecEnv / Gymnasium-vector API
    def reset(self, seed=None, options=None) -> np.ndarray:
        """`env.reset(seed=s)` (uav_env.py:400-410; the evaluation scripts reseed every episode,
        sweep_eval.py:415,463,584): re-keys the Philox streams, then resets every env.  SB3's `seed(s)` stores the
        seed for the next reset, as VecEnv.seed does."""
        if seed is not None:
            self._next_seed = int(seed)
        if self._next_seed is not None:
            self.reseed(self._next_seed)
            self._next_seed = None
        obs = self.reset_tensors()
        return obs.cpu().numpy()

    def reseed(self, seed: int) -> None:
        """New Philox key for every later launch (uavsim_set_seed).  Draws stay a pure function of
        (seed, global env id, episode, step, sensor, slot)."""
        self.cfg.seed = int(seed) & (2 ** 64 - 1)
        N.check(self.lib.uavsim_set_seed(self._h, C.c_uint64(self.cfg.seed)))

    def step_async(self, actions) -> None:
        a = np.asarray(actions)
        if a.shape != (self.num_envs,):
            raise ValueError("actions must have shape (%d,)" % self.num_envs)
        if ((a < 0) | (a > 4)).any():
            bad = int(np.flatnonzero((a < 0) | (a > 4))[0])
            raise ValueError("Invalid action: %r" % (a[bad],))          # uav_env.py:468
        self._pending_actions = torch.as_tensor(a.astype(np.int32, copy=False))

    def step_wait(self):
        if self._pending_actions is None:
            raise RuntimeError("step_wait() without step_async()")
        # one pinned staging set for the whole numpy path: actions H2D, step, obs / reward / done D2H, one sync
        # (uavsim_step_host); the arrays handed out are fresh copies, as DummyVecEnv's are
        if self._np_host is None:
            self._np_host = self.make_host_buffers()
            self._np_infos = [{} for _ in range(self.num_envs)]
            self._np_dirty = []
        host = self._np_host
        host["actions"].copy_(self._pending_actions)
        self._pending_actions = None
        if self.wind_extra_step > 0 and self.cfg.rng_mode == N.RNG_PHILOX:
            self._actions_dev.copy_(host["actions"], non_blocking=True)
            self._wind_repeat(self._actions_dev)
        with torch.cuda.device(self.device):
            N.check(self.lib.uavsim_step_host(self._h, _ptr(host["actions"]), _ptr(host["obs"]), _ptr(host["reward"]),
                                              _ptr(host["done"]), self._stream()))
        obs = host["obs"].numpy().copy()
        rew = host["reward"].numpy().copy()
        dones = host["done"].numpy().astype(bool)
        infos = self._np_infos                      # one dict per env, reused; last step's entries are cleared
        for e in self._np_dirty:
            infos[e] = {}
        self._np_dirty = []
        if self.info_every_step:                    # the reference returns _get_info() with every step (uav_env.py:488)
            self._np_dirty = list(range(self.num_envs))
            for e, d in enumerate(self.get_info()):
                infos[e] = d
        idx = np.flatnonzero(dones)
        if idx.size:
            didx = torch.as_tensor(idx, device=self.device)
            stats = self.episode_stats[didx].cpu().numpy()
            tobs = self.terminal_obs[didx].cpu().numpy() if self.cfg.auto_reset else None
            tinfo = tratio = None
            if self.cfg.auto_reset and self.terminal_info is not None:
                tinfo = self.terminal_info[didx].cpu().numpy()
                tratio = self.terminal_ratios[didx].cpu().numpy()
            elif not self.cfg.auto_reset and not self.info_every_step:
                term = self.get_info(idx)           # nothing has been reset: the current state IS the terminal one
            if not self.info_every_step:
                self._np_dirty = [int(e) for e in idx]
            for k, e in enumerate(idx):
                info = infos[e]
                # infos[e] at episode end carries _get_info() of the TERMINAL state, as DummyVecEnv hands it out
                # (the callers' keys: total_data_collected / battery / sensor_collection_ratios, dqn.py:1148-1155;
                # coverage_percentage / total_reward / boundary_hits / edge_steps, sim_to_real_sweep.py:222-225)
                if tinfo is not None:
                    info.update(self._info_dict(tinfo[k], tratio[k]))
                elif not self.cfg.auto_reset and not self.info_every_step:
                    info.update(term[k])
                info["TimeLimit.truncated"] = True        # terminated is never True (uav_env.py:471)
                if tobs is not None:
                    st = dict(zip(N.STAT_NAMES, stats[k].tolist()))
                    info["terminal_observation"] = tobs[k]
                    info["episode"] = {"r": st["return"], "l": int(st["length"])}   # Monitor
                    self.last_episode_stats[e] = self._episode_dict(st)
        return obs, rew, dones, list(infos)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    @staticmethod
    def _episode_dict(st):
        """DomainRandEnv.last_episode_stats (dqn.py:316-331).  `time_to_coverage` is None AS EXECUTED: the reference's
        reset() clears `_first_full_coverage_step` (dqn.py:302) before it snapshots the statistics (:330), so its
        dict always holds None; the step the coverage actually completed at is kept under `first_full_coverage_step`."""
        ttc = st["time_to_coverage"]
        return {"total_generated": st["total_generated"], "total_collected": st["total_collected"],
                "total_lost": st["total_lost"], "battery_remaining": st["battery_remaining"],
                "ndr": st["ndr"], "fairness_std": st["fairness_std"], "jains_index": st["jains_index"],
                "grid_size": (int(st["grid_w"]), int(st["grid_h"])), "num_sensors": int(st["num_sensors"]),
                "data_efficiency": st["data_efficiency"], "bytes_per_wh": st["bytes_per_wh"],
                "time_to_coverage": None, "first_full_coverage_step": None if ttc < 0 else int(ttc)}

    @staticmethod
    def _info_dict(row, ratios) -> dict:
        """One UAVSIM_INFO_* row + ratios row -> the reference's _get_info() dict (uav_env.py:676-700)."""
        v = dict(zip(N.INFO_NAMES, row.tolist()))
        n = int(v.pop("num_sensors"))
        d = {"uav_position": np.array([v.pop("x"), v.pop("y")], dtype=np.float32)}
        for k in ("sensors_collected", "current_step", "high_urgency_sensors", "capture_effect_triggers",
                  "boundary_hits", "edge_steps"):
            v[k] = int(v[k])
        v["is_alive"] = bool(v["is_alive"])
        d.update(v)
        d["sensor_collection_ratios"] = [float(x) for x in ratios[:n]]
        return {k: d[k] for k in _INFO_KEYS}

    def get_info(self, indices=None) -> list:
        """The reference's `_get_info()` dict (uav_env.py:676-700) of the CURRENT state of the requested envs: one
        small kernel (uavsim_get_info) fills a [E, 18] row block + the [E, P] collection ratios, one copy each."""
        idx = self._indices(indices)
        if self._info_buf is None:
            self._info_buf = (torch.zeros((self.num_envs, N.INFO_COUNT), dtype=torch.float64, device=self.device),
                              torch.zeros((self.num_envs, self.cfg.max_sensors), dtype=torch.float64, device=self.device))
        info, ratios = self._info_buf
        with torch.cuda.device(self.device):
            N.check(self.lib.uavsim_get_info(self._h, _ptr(info), _ptr(ratios), self._stream()))
        sel = torch.as_tensor(np.asarray(idx, dtype=np.int64), device=self.device)
        rows, rat = info[sel].cpu().numpy(), ratios[sel].cpu().numpy()
        return [self._info_dict(rows[k], rat[k]) for k in range(len(idx))]

    def _indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        if isinstance(indices, int):
            return [indices]
        return list(indices)

    def get_attr(self, attr_name: str, indices=None) -> list:
        idx = self._indices(indices)
        if attr_name == "last_episode_stats":
            return [self.last_episode_stats[i] for i in idx]
        if attr_name == "num_sensors":
            return [int(self._num_sensors[i]) for i in idx]
        if attr_name == "grid_size":
            rec = self.records()
            return [(int(rec[i]["grid_w"]), int(rec[i]["grid_h"])) for i in idx]
        if attr_name in ("_curriculum_stage", "curriculum_stage"):
            return [self._stage for _ in idx]
        if attr_name == "render_mode":
            return [None for _ in idx]
        if hasattr(self, attr_name):
            return [getattr(self, attr_name) for _ in idx]
        raise AttributeError(attr_name)

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        if attr_name in ("_curriculum_stage", "curriculum_stage"):
            self.set_curriculum_stage(value)
        else:
            raise AttributeError("cannot set %r on the batched environment" % attr_name)

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> list:
        idx = self._indices(indices)
        if method_name == "set_curriculum_stage":           # dqn.py:921-926
            self.set_curriculum_stage(*args, **kwargs)
            return [None for _ in idx]
        if method_name in ("get_info", "_get_info"):
            return self.get_info(idx)
        raise AttributeError(method_name)

    def env_is_wrapped(self, wrapper_class, indices=None) -> list:
        return [False for _ in self._indices(indices)]

    def set_curriculum_stage(self, stage: int):
        """DomainRandEnv.set_curriculum_stage (dqn.py:258-277): later resets sample grids from it."""
        grids = self._apply_stage(self.cfg, stage)
        flat = (C.c_int32 * (2 * len(grids)))(*[v for g in grids for v in g])
        N.check(self.lib.uavsim_set_stage_grids(self._h, len(grids), flat))

    def set_stage_grids(self, grids) -> None:
        """Install an explicit grid list for DomainRandEnv resets (uavsim_set_stage_grids): the reference samples
        from CURRICULUM_STAGES (100..500, dqn.py:63-69); BASELINE configs[2] words the sweep as 300..1000."""
        grids = [(int(w), int(h)) for w, h in grids]
        flat = (C.c_int32 * (2 * len(grids)))(*[v for g in grids for v in g])
        N.check(self.lib.uavsim_set_stage_grids(self._h, len(grids), flat))
        self._stage = -1

    def seed(self, seed: Optional[int] = None) -> Sequence[Optional[int]]:
        """VecEnv.seed: remembered and applied by the next `reset()` (one Philox key for the batch; env i's streams
        are keyed on its global id, so the per-env seeds SB3 expects back are seed + global id)."""
        if seed is None:
            return [None for _ in range(self.num_envs)]
        self._next_seed = int(seed)
        return [int(seed) + int(self.cfg.env_id_offset) + i for i in range(self.num_envs)]

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self.lib.uavsim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
