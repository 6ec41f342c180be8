// Episode end and reset: episode statistics, the new-episode device function (UAVEnvironment.reset / DomainRandEnv.reset)
// and the reset kernel.  Part of the fused step/reset source (see uavsim_device.cuh for the overview).
#pragma once
#include "uavsim_device.cuh"

namespace uavsim {

// ------------------------------------------------------------------------------------------------
// episode end / reset
// ------------------------------------------------------------------------------------------------
// Episode statistics written when an environment finishes: DomainRandEnv.last_episode_stats
// (dqn.py:305-331) + what Monitor reports.  Rare path: plain float64 shuffle reductions.  R is the
// committed record in shared memory.
template <typename Real, int SPL>
__device__ __forceinline__ void write_episode_stats(const Params<Real>& p, const Buffers& b, int e, int lane,
                                                    const uavsim_env_record& R, const Sensor<Real> (&s)[SPL]) {
    double g = 0, t = 0, l = 0, s1 = 0, s2 = 0, cnt = 0;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        if (s[k].active) {
            const double gen = (double)generated_of(p, s[k], R.elapsed_s);
            g += gen;
            t += (double)s[k].tx;
            l += (double)s[k].lost;
            if (gen > 0) {
                const double rate = (double)s[k].tx / gen * 100;
                s1 += rate;
                s2 += rate * rate;
                cnt += 1;
            }
        }
    }
    g = warp_sum(g);
    t = warp_sum(t);
    l = warp_sum(l);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    cnt = warp_sum(cnt);
    if (lane == 0) {
        const int N = R.num_sensors;
        const int nv = __popc(R.visited_lo) + __popc(R.visited_hi);
        const double used = p.max_battery - R.battery;
        const double mean = cnt > 0 ? s1 / cnt : 0.0;
        const double var = cnt > 0 ? fmax(s2 / cnt - mean * mean, 0.0) : 0.0;
        double v[UAVSIM_STAT_COUNT];
        v[UAVSIM_STAT_RETURN] = R.episode_return;
        v[UAVSIM_STAT_LENGTH] = R.step;
        v[UAVSIM_STAT_TOTAL_GENERATED] = g;
        v[UAVSIM_STAT_TOTAL_COLLECTED] = t;
        v[UAVSIM_STAT_TOTAL_LOST] = l;
        v[UAVSIM_STAT_BATTERY_REMAINING] = R.battery;
        v[UAVSIM_STAT_NDR] = (double)nv / N * 100;
        v[UAVSIM_STAT_FAIRNESS_STD] = sqrt(var);
        v[UAVSIM_STAT_JAINS] = (cnt > 0 && s2 > 0) ? (s1 * s1) / (cnt * s2) : 1.0;  // dqn.py:446-451
        v[UAVSIM_STAT_GRID_W] = R.grid_w;
        v[UAVSIM_STAT_GRID_H] = R.grid_h;
        v[UAVSIM_STAT_NUM_SENSORS] = N;
        v[UAVSIM_STAT_DATA_EFFICIENCY] = g > 0 ? t / g * 100 : 0.0;
        v[UAVSIM_STAT_BYTES_PER_WH] = used > 0 ? t / used : 0.0;
        v[UAVSIM_STAT_TIME_TO_COVERAGE] = R.first_full_coverage_step;
        v[UAVSIM_STAT_CAPTURE_TRIGGERS] = R.capture_triggers;
        v[UAVSIM_STAT_BOUNDARY_HITS] = R.boundary_hits;
        v[UAVSIM_STAT_EDGE_STEPS] = R.edge_steps;
        v[UAVSIM_STAT_COLLISIONS] = R.collisions;
        v[UAVSIM_STAT_BASE_RETURN] = R.total_reward;
        if (b.episode_stats)
            for (int i = 0; i < UAVSIM_STAT_COUNT; ++i) b.episode_stats[(size_t)e * UAVSIM_STAT_COUNT + i] = v[i];
        if (b.stat_sums) {
            for (int i = 0; i < UAVSIM_STAT_COUNT; ++i) atomicAdd(b.stat_sums + i, v[i]);
            atomicAdd(b.stat_sums + UAVSIM_STAT_COUNT, 1.0);
        }
    }
}

// UAVEnvironment._get_info (uav_env.py:676-700) of the state in (R, s): one float64[UAVSIM_INFO_COUNT] row plus the
// per-sensor "sensor_collection_ratios" row.  Called for the terminal state of a finished episode (the values SB3
// callers read from infos[i], dqn.py:1148-1155, sim_to_real_sweep.py:222-225) and by the on-demand info kernel.
// The urgency statistics follow the reference's float32 AoI array (uav_env.py:386-394): np.max, float32 np.mean
// (pairwise sum, float32 division) and the count of entries above float32(0.8).
template <typename Real, int SPL>
__device__ __forceinline__ void write_info(const Params<Real>& p, int lane, const uavsim_env_record& R,
                                           const Sensor<Real> (&s)[SPL], double* scratch, double* info, double* ratios) {
    constexpr bool kParity = std::is_same<Real, double>::value;
    const int N = R.num_sensors;
    float aoi[SPL], mx = 0.f;
    int high = 0;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        if constexpr (kParity)
            aoi[k] = (float)(s[k].buf / p.rate);
        else
            aoi[k] = s[k].buf * p.inv_rate;
        if (!s[k].active) aoi[k] = 0.f;
        mx = fmaxf(mx, aoi[k]);  // AoI >= 0
        high += __popc(__ballot_sync(kFull, s[k].active && aoi[k] > 0.8f));
        if (ratios && s[k].active) {
            const double gen = (double)generated_of(p, s[k], R.elapsed_s);
            ratios[lane + 32 * k] = (double)s[k].tx / fmax(gen, 1e-6);
        }
    }
    mx = warp_max(mx);
    const float sum = sensor_sum<Real, SPL>(aoi, N, lane, scratch, kSumNumpy);
    if (lane == 0 && info) {
        const int nv = __popc(R.visited_lo) + __popc(R.visited_hi);
        info[UAVSIM_INFO_X] = R.x;
        info[UAVSIM_INFO_Y] = R.y;
        info[UAVSIM_INFO_BATTERY] = R.battery;
        info[UAVSIM_INFO_BATTERY_PERCENT] = (R.battery / p.max_battery) * 100;  // uav.py:238
        info[UAVSIM_INFO_SENSORS_COLLECTED] = nv;
        info[UAVSIM_INFO_CURRENT_STEP] = R.step;
        info[UAVSIM_INFO_TOTAL_REWARD] = R.total_reward;
        info[UAVSIM_INFO_TOTAL_DATA_COLLECTED] = R.total_data_collected;
        info[UAVSIM_INFO_COVERAGE_PERCENTAGE] = ((double)nv / N) * 100;
        info[UAVSIM_INFO_IS_ALIVE] = R.battery > p.batt_floor;
        info[UAVSIM_INFO_MAX_URGENCY] = mx;
        info[UAVSIM_INFO_AVG_URGENCY] = sum / (float)N;
        info[UAVSIM_INFO_HIGH_URGENCY_SENSORS] = high;
        info[UAVSIM_INFO_CAPTURE_EFFECT_TRIGGERS] = R.capture_triggers;
        info[UAVSIM_INFO_BOUNDARY_HITS] = R.boundary_hits;
        info[UAVSIM_INFO_EDGE_STEPS] = R.edge_steps;
        info[UAVSIM_INFO_LAST_STEP_BYTES] = R.last_step_bytes;
        info[UAVSIM_INFO_NUM_SENSORS] = N;
    }
}

// Per-environment draws of the sim-to-real wrappers (sim_to_real_analysis.py:176-181, :207): GPS noise pair for
// the observation of `step` (0 = reset observation) and, at reset, the wind angle.  REPLAY: table [E][4];
// PHILOX: stream kStreamMisc, counter word 3 = step << 8 | 4 (words 0, 1 there are the grid / inherited-SF draws).
// The same pair as float64 (PerturbedDQNEnv adds it to the float64 copy of the position, sim_to_real_sweep.py:123-124).
template <typename Real, bool REPLAY>
__device__ __forceinline__ void env_draws_pos(const Params<Real>& p, const double* table, int e, uint32_t episode, uint32_t step,
                                              double& gx, double& gy) {
    if constexpr (REPLAY) {
        gx = table ? table[4 * (size_t)e + 0] : 0.0;
        gy = table ? table[4 * (size_t)e + 1] : 0.0;
    } else {
        const uint64_t gid = (uint64_t)(p.env_id_offset + e);
        const Philox4 q = philox4x32_10((uint32_t)gid, ((uint32_t)(gid >> 32) & 0xffffffu) | (kStreamMisc << 24), episode,
                                        (step << 8) | 4u, p.key0, p.key1);
        Real z0, z1;
        box_muller<Real>(q.x, q.y, p.gps_sigma, z0, z1);
        gx = (double)z0;
        gy = (double)z1;
    }
}
// PerturbedDQNEnv._get_observation (sim_to_real_sweep.py:121-136): the position the observation pass sees.
//   noisy = clip(float64(position) + N(0, sigma, 2), 0, [W-1, H-1]); obs[0], obs[1] = float32(noisy / [W, H]); the pass's
//   RSSI takes float32(noisy) (iot_sensors.py:162).
struct NoisyPos {
    float x, y;    // float32(noisy): what calculate_rssi sees
    float ox, oy;  // obs[0], obs[1]
};
template <typename Real>
__device__ __forceinline__ NoisyPos noisy_position(float x, float y, int W, int H, double gx, double gy) {
    const double nx = fmin(fmax((double)x + gx, 0.0), (double)(W - 1)), ny = fmin(fmax((double)y + gy, 0.0), (double)(H - 1));
    NoisyPos n;
    n.x = (float)nx;
    n.y = (float)ny;
    if constexpr (std::is_same<Real, double>::value) {
        n.ox = (float)(nx / (double)W);
        n.oy = (float)(ny / (double)H);
    } else {
        n.ox = __fdividef(n.x, (float)W);
        n.oy = __fdividef(n.y, (float)H);
    }
    return n;
}

template <typename Real, bool REPLAY>
__device__ __forceinline__ void env_draws(const Params<Real>& p, const double* table, int e, uint32_t episode, uint32_t step,
                                          float& gx, float& gy, double& angle) {
    if constexpr (REPLAY) {
        gx = table ? (float)table[4 * (size_t)e + 0] : 0.f;  // obs[0] += python float: added as float32
        gy = table ? (float)table[4 * (size_t)e + 1] : 0.f;
        angle = table ? table[4 * (size_t)e + 2] : 0.0;
    } else {
        const uint64_t gid = (uint64_t)(p.env_id_offset + e);
        const Philox4 q = philox4x32_10((uint32_t)gid, ((uint32_t)(gid >> 32) & 0xffffffu) | (kStreamMisc << 24), episode,
                                        (step << 8) | 4u, p.key0, p.key1);
        Real z0, z1;
        box_muller<Real>(q.x, q.y, p.gps_sigma, z0, z1);
        gx = (float)z0;
        gy = (float)z1;
        angle = 6.283185307179586 * u01<double>(q.z);
    }
}

// Start a new episode for this warp's environment: record in shared memory (ws.rec), sensors in
// registers; writes the first observation.  FIXED: UAVEnvironment.reset (uav_env.py:400-427,
// iot_sensors.py:305-321).  DOMAIN_RAND: DomainRandEnv.reset (dqn.py:301-372).  ws.rec.episode names
// the NEW episode on entry and is incremented on exit.
template <typename Real, bool REPLAY, int SPL, int FPS>
__device__ __forceinline__ void reset_env(const Params<Real>& p, const Buffers& b, int e, int lane, WarpSmem& ws,
                                       Sensor<Real> (&s)[SPL], Real* exp) {
    uavsim_env_record& R = ws.rec;
    __syncwarp();
    const uint64_t gid = (uint64_t)(p.env_id_offset + e);
    const uint32_t env_lo = (uint32_t)gid, env_hi = (uint32_t)(gid >> 32) & 0xffffffu;
    const uint32_t episode = R.episode;
    int init_sf = p.init_sf;
    int W = R.grid_w, H = R.grid_h, N = R.num_sensors;
    float start_x = R.start_x, start_y = R.start_y;

    if (p.domain_rand) {
        if constexpr (REPLAY) {
            // injected scenario (the reference draws it from unseeded / global generators)
            if (b.reset_init_sf) init_sf = b.reset_init_sf[e];
            if (b.rs_grid_wh) {
                W = b.rs_grid_wh[2 * e];
                H = b.rs_grid_wh[2 * e + 1];
            }
            if (b.rs_num_sensors) N = b.rs_num_sensors[e];
            if (b.rs_start_xy) {
                start_x = b.rs_start_xy[2 * e];
                start_y = b.rs_start_xy[2 * e + 1];
            }
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                const int j = lane + 32 * k;
                if (j < N && b.rs_sensor_xy) {
                    s[k].sx = b.rs_sensor_xy[((size_t)e * p.P + j) * 2];
                    s[k].sy = b.rs_sensor_xy[((size_t)e * p.P + j) * 2 + 1];
                }
            }
        } else {
            // (a) SF inherited from the OLD sensor 0 after the parent reset's observation pass at the
            //     OLD start position (dqn.py:338-353 with uav_env.py:407-427): one ADR update from None.
            {
                const float s0x = shfl(s[0].sx, 0), s0y = shfl(s[0].sy, 0);
                const Real det0 = det_rssi(p, start_x, start_y, s0x, s0y);
                const Philox4 q = philox4x32_10(env_lo, env_hi | (kStreamMisc << 24), episode, 1, p.key0, p.key1);
                Real z0, z1;
                box_muller<Real>(q.x, q.y, b.env_sigma ? (Real)b.env_sigma[e] : p.sigma, z0, z1);
                Real ema = none_value<Real>();
                int sf = p.init_sf;
                adr_update(p, det0 + z0, ema, sf);
                init_sf = sf;
                // (b) grid from the current curriculum stage (dqn.py:280-283)
                const Philox4 g = philox4x32_10(env_lo, env_hi | (kStreamMisc << 24), episode, 0, p.key0, p.key1);
                const int gi = (int)(g.x % (uint32_t)p.num_grids);
                W = p.grids[gi][0];
                H = p.grids[gi][1];
            }
            // (c) uniform layout (dqn.py:342-343, _layout_uniform :135-137)
            const double Wd = (double)W, Hd = (double)H;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                const int j = lane + 32 * k;
                if (j < N) {
                    const Philox4 q = philox4x32_10(env_lo, env_hi | (kStreamLayout << 24), episode, (uint32_t)j, p.key0,
                                                    p.key1);
                    s[k].sx = (float)(u01<double>(q.x) * Wd);
                    s[k].sy = (float)(u01<double>(q.y) * Hd);
                }
            }
            // (d) far start: first candidate >= min_start_dist from every sensor, else the furthest of
            //     max_start_tries (dqn.py:375-403).  32 candidates per round, one per lane.
            float2* pos = reinterpret_cast<float2*>(ws.red);  // 64 float2 = 512 B
            __syncwarp();
#pragma unroll
            for (int k = 0; k < SPL; ++k)
                if (lane + 32 * k < N) pos[lane + 32 * k] = make_float2(s[k].sx, s[k].sy);
            __syncwarp();
            float best_d = -1.0f, best_x = 0.f, best_y = 0.f;
            bool found = false;
            for (int base = 0; base < p.max_start_tries && !found; base += 32) {
                const int t = base + lane;
                const Philox4 q = philox4x32_10(env_lo, env_hi | (kStreamStart << 24), episode, (uint32_t)t, p.key0, p.key1);
                const float cx = (float)(0.05 * Wd + (0.95 * Wd - 0.05 * Wd) * u01<double>(q.x));
                const float cy = (float)(0.05 * Hd + (0.95 * Hd - 0.05 * Hd) * u01<double>(q.y));
                float dmin = CUDART_INF_F;
                for (int i = 0; i < N; ++i) {
                    const float dx = cx - pos[i].x, dy = cy - pos[i].y;
                    dmin = fminf(dmin, sqrtf(dx * dx + dy * dy));
                }
                const bool valid = t < p.max_start_tries;
                const unsigned ok = __ballot_sync(kFull, valid && (double)dmin >= p.min_start_dist);
                if (ok) {
                    const int src = __ffs(ok) - 1;
                    best_x = shfl(cx, src);
                    best_y = shfl(cy, src);
                    found = true;
                } else {
                    // running "furthest so far", first occurrence wins ties (strict > at dqn.py:397)
                    const float d = valid ? dmin : -2.0f;
                    const float m = warp_max(d);
                    const unsigned who = __ballot_sync(kFull, d == m);
                    if (m > best_d) {
                        const int src = __ffs(who) - 1;
                        best_d = m;
                        best_x = shfl(cx, src);
                        best_y = shfl(cy, src);
                    }
                }
            }
            __syncwarp();
            start_x = best_x;
            start_y = best_y;
        }
    }

    // IoTSensor.reset (iot_sensors.py:305-321) / fresh IoTSensor objects (dqn.py:345-360)
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const int j = lane + 32 * k;
        s[k].active = j < N;
        s[k].attempt = false;
        s[k].cur = Real(0);
        if (s[k].active) {
            if (p.domain_rand) {
                s[k].buf = Real(0);
            } else {
                Real fill;
                if constexpr (REPLAY) {
                    fill = (Real)b.reset_fill[(size_t)e * p.P + j];
                } else {
                    const Philox4 q = philox4x32_10(env_lo, env_hi | (kStreamFill << 24), episode, (uint32_t)j, p.key0,
                                                    p.key1);
                    fill = Real(0.20) + (Real(0.60) - Real(0.20)) * u01<Real>(q.x);  // np_random.uniform(0.20, 0.60)
                }
                fill = fill < Real(0) ? Real(0) : (fill > Real(1) ? Real(1) : fill);
                s[k].buf = p.maxbuf * fill;
            }
            s[k].gen = s[k].buf;
            s[k].tx = Real(0);
            s[k].lost = Real(0);
            s[k].ema = none_value<Real>();
            s[k].sf = init_sf;
            if (exp) exp[UAVSIM_SLOT_U * p.P + j] = fdiv(s[k].buf, p.maxbuf);  // the fill actually used (diagnostic)
        } else {
            s[k].buf = s[k].tx = s[k].lost = s[k].gen = Real(0);
            s[k].ema = none_value<Real>();
            s[k].sx = s[k].sy = 0.f;
            s[k].sf = 12;
        }
    }

    // UAV.reset (uav.py:241-258) and the episode counters (uav_env.py:412-424)
    Hot h;
    h.x = start_x;
    h.y = start_y;
    h.elapsed = 0.f;
    h.battery = p.max_battery;
    h.W = W;
    h.H = H;
    h.N = N;
    h.step = 0;
    h.vis_lo = h.vis_hi = 0u;
    h.episode = episode;
    compute_det(p, h.x, h.y, s);
    Draws<Real, REPLAY, SPL> dr;
    dr.env_lo = env_lo;
    dr.env_hi = env_hi;
    dr.episode = episode;
    dr.step = 0;
    dr.paired = false;  // a reset observation belongs to one environment: (env, episode, step 0)
    dr.pair_episode = dr.pair_step = 0;
    dr.sigma = b.env_sigma ? (Real)b.env_sigma[e] : p.sigma;
    dr.tab = REPLAY ? b.reset_draws + (size_t)e * UAVSIM_NUM_SLOTS * p.P : nullptr;
    dr.exp = exp;
    float* out = obs_frame(p, b, e);
    if (p.frame_k > 1) {  // a new episode starts with k-1 all-zero frames (VecFrameStack zeroes the stack on done)
        float* blk = b.obs + (size_t)e * p.obs_stride;
        for (int i = lane; i < p.obs_stride; i += 32) blk[i] = 0.0f;
        __syncwarp();
    }
    NoisyPos np_{};
    const bool gps_obs = p.gps_in_obs && p.gps_sigma > Real(0);
    if (gps_obs) {  // PerturbedDQNEnv: reset() observes through the same noisy position (sim_to_real_sweep.py:121-136)
        double gx, gy;
        env_draws_pos<Real, REPLAY>(p, b.reset_env_draws, e, episode, 0u, gx, gy);
        np_ = noisy_position<Real>(start_x, start_y, W, H, gx, gy);
        h.x = np_.x;
        h.y = np_.y;
        compute_det(p, h.x, h.y, s);
    }
    observe<Real, REPLAY, SPL, FPS, true, false>(p, h, s, dr, lane, out);
    if (gps_obs) {
        __syncwarp();
        if (lane == 0) {
            out[0] = np_.ox;
            out[1] = np_.oy;
        }
        h.x = start_x;  // the true position is restored (:134)
        h.y = start_y;
        __syncwarp();
    }
    double wind_angle = 0.0;
    if ((p.gps_sigma > Real(0) && !p.gps_in_obs) || p.wind_drift > 0.0) {  // GPSNoisyEnv.reset :183-185, WindDriftEnv.reset :205-210
        float gx, gy;
        double angle;
        env_draws<Real, REPLAY>(p, b.reset_env_draws, e, episode, 0u, gx, gy, angle);
        __syncwarp();
        if (lane == 0 && p.gps_sigma > Real(0)) {
            out[0] = out[0] + gx;
            out[1] = out[1] + gy;
        }
        wind_angle = angle;
        __syncwarp();
    }
    for (int i = 3 + FPS * N + lane; i < p.obs_dim; i += 32) out[i] = 0.0f;  // ghost slots
    if (p.frame_k > 1) {
        __syncwarp();
        for (int i = lane; i < p.obs_dim; i += 32) out[i + p.frame_mirror] = out[i];
        // LatencyEnv primes its FIFO with copies of the reset observation (sim_to_real_analysis.py:235-238): the
        // `latency` frames behind the newest one hold the same row
        float* blk = b.obs + (size_t)e * p.obs_stride;
        for (int j = 1; j <= p.latency; ++j) {
            const int r = (b.frame_slot - j + p.frame_k) % p.frame_k;
            for (int i = lane; i < p.obs_dim; i += 32) {
                const float v = out[i];
                blk[r * p.obs_dim + i] = v;
                blk[r * p.obs_dim + i + p.frame_mirror] = v;
            }
        }
    }
    const float prev = p.domain_rand ? nearest_with_data(h.x, h.y, s)
                                     : (p.relational ? nearest_unvisited(h.x, h.y, 0u, 0u, lane, s) : 0.f);
    __syncwarp();
    if (lane == 0) {
        R.x = h.x;
        R.y = h.y;
        R.start_x = start_x;
        R.start_y = start_y;
        R.battery = p.max_battery;
        R.total_reward = 0.0;
        R.total_data_collected = 0.0;
        R.episode_return = 0.0;
        R.last_step_bytes = 0.f;
        R.prev_dist_nearest = prev;
        R.elapsed_s = 0.f;
        R.step = 0;
        R.grid_w = W;
        R.grid_h = H;
        R.num_sensors = N;
        R.visited_lo = R.visited_hi = 0u;
        R.capture_triggers = R.boundary_hits = R.edge_steps = R.collisions = R.delivered = 0;
        R.first_full_coverage_step = -1;
        R.episode = episode + 1;
        R.init_sf = init_sf;
        R.dwell_steps = 0;
        R.wind_angle = wind_angle;
        if (b.action_mask) b.action_mask[e] = action_mask_of(h.x, h.y, W, H);
    }
    __syncwarp();
}

template <typename Real>
constexpr int min_blocks_per_sm() {
    return std::is_same<Real, double>::value ? 1 : 3;
}

// Block shape of the step kernel.  The one-slot-per-lane kernel fits 80 registers: 8 warps x 3 blocks per
// SM.  The two-slot kernel needs ~100 registers to stay out of local memory (with most of the SM given to
// shared memory the L1 is small, so a spill is an L2 round trip): 10 warps x 2 blocks per SM.
// (the UAVSIM_W*/B* macros exist for block-shape sweeps: tools/build_variants.py)
#ifndef UAVSIM_W2
#define UAVSIM_W2 10
#endif
#ifndef UAVSIM_B2
#define UAVSIM_B2 2
#endif
#ifndef UAVSIM_W1
#define UAVSIM_W1 8
#endif
#ifndef UAVSIM_B1
#define UAVSIM_B1 3
#endif
template <typename Real, int SPL, bool PAIR = false>
__host__ __device__ constexpr int step_warps() {
    return std::is_same<Real, double>::value ? 8 : ((SPL == 2 || PAIR) ? UAVSIM_W2 : UAVSIM_W1);
}
template <typename Real, int SPL, bool PAIR = false>
__host__ __device__ constexpr int step_min_blocks() {
    return std::is_same<Real, double>::value ? 1 : ((SPL == 2 || PAIR) ? UAVSIM_B2 : UAVSIM_B1);
}

template <typename Real, bool REPLAY, int SPL, int FPS>
__global__ void __launch_bounds__(kThreads, min_blocks_per_sm<Real>())
    reset_kernel(const __grid_constant__ Params<Real> p, const __grid_constant__ Buffers b, const uint8_t* mask) {
    __shared__ WarpSmem smem[kWarpsPerBlock];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSmem& ws = smem[warp];
    for (int e = blockIdx.x * kWarpsPerBlock + warp; e < b.E; e += gridDim.x * kWarpsPerBlock) {
        if (mask && !mask[e]) continue;
        load_record(ws, b.rec + e, lane);
        if (b.zero_episode) {  // env.reset(seed=s): the episode that follows is a function of s alone
            __syncwarp();
            if (lane == 0) ws.rec.episode = 0;
            __syncwarp();
        }
        Sensor<Real> s[SPL];
        load_sensors<Real, SPL>(b, p, e, ws.rec.num_sensors, lane, s);
        Real* exp = b.draw_export ? reinterpret_cast<Real*>(b.draw_export) + (size_t)e * UAVSIM_NUM_SLOTS * p.P : nullptr;
        if (exp) {
            const Real nan = Real(CUDART_NAN);
            for (int i = lane; i < UAVSIM_NUM_SLOTS * p.P; i += 32) exp[i] = nan;
            __syncwarp();
        }
        reset_env<Real, REPLAY, SPL, FPS>(p, b, e, lane, ws, s, exp);
        store_sensors<Real, SPL>(b, p, e, lane, s, true);
        store_record(ws, b.rec + e, lane);
        if (lane == 0 && b.done) b.done[e] = 0;
    }
}

// _get_info() of the current state of every environment (uavsim_get_info).
template <typename Real, int SPL>
__global__ void __launch_bounds__(kThreads) info_kernel(const __grid_constant__ Params<Real> p, const __grid_constant__ Buffers b,
                                                        double* info, double* ratios) {
    __shared__ WarpSmem smem[kWarpsPerBlock];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpSmem& ws = smem[warp];
    for (int e = blockIdx.x * kWarpsPerBlock + warp; e < b.E; e += gridDim.x * kWarpsPerBlock) {
        load_record(ws, b.rec + e, lane);
        Sensor<Real> s[SPL];
        load_sensors<Real, SPL>(b, p, e, ws.rec.num_sensors, lane, s);
        write_info<Real, SPL>(p, lane, ws.rec, s, ws.red, info + (size_t)e * UAVSIM_INFO_COUNT,
                              ratios ? ratios + (size_t)e * p.P : nullptr);
        __syncwarp();
    }
}

}  // namespace uavsim
