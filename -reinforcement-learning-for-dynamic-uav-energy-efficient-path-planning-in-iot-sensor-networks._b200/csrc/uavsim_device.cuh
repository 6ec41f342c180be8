// Fused batched UAV-LoRa environment step for sm_100a: sensors across the lanes of a warp (slot j = lane + 32*k,
// k < SPL), structure-of-arrays state in HBM.  This header: parameters, buffers and the per-sensor device functions;
// uavsim_reset.cuh: episode end / reset; uavsim_step.cuh: the step kernel.
//
// What one launch of step_kernel does for every environment (reference file:line under
// /root/reference/src):
//   1. age every sensor buffer, cap it, account overflow        environment/iot_sensors.py:114-125
//   2. MOVE: bounds-checked +-1 cell, battery drain              environment/uav.py:127-185
//      COLLECT: hover drain (uav.py:187-206), per-sensor EMA-ADR spreading-factor update
//      (iot_sensors.py:223-259) on a two-ray RSSI (iot_sensors.py:147-197), link-success
//      sigmoid (iot_sensors.py:202-212), duty-cycle gate, Capture-Effect resolution per SF
//      group and buffer drain (uav_env.py:518-632, iot_sensors.py:127-145)
//   3. multi-objective reward                                    rewards/reward_function.py:46-128
//   4. truncation + terminal penalties                           environment/uav_env.py:471-485
//   5. observation row (second ADR update + range check)         environment/uav_env.py:638-674
//   6. DomainRandEnv shaping, padding, episode statistics        agents/dqn/dqn.py:301-444
//   7. SB3-style in-kernel auto-reset with terminal observation  (DummyVecEnv semantics)
//
// The kernels are templates over
//   Real   : double = PARITY build (float64 accumulators like the reference, reference operation
//            order, numpy's pairwise summation order, this TU compiled with -fmad=false);
//            float  = THROUGHPUT build (float32 state, MUFU intrinsics, shuffle reductions).
//   REPLAY : true  = every random draw is read from a slot-addressed table (parity tests),
//            false = counter-based Philox4x32-10.
//   SPL    : sensor slots per lane (1: P <= 32, 2: P <= 64).
//   FPS / STACK / LEAN / PAIR : features per sensor, frame ring, feature set compiled in, two environments per pass.
// Execution shape (uavsim_step.cuh): persistent warps draw tiles of up to 16 environments by ticket; scalar work runs
// one environment per lane (phases A, C), per-sensor work one environment (or pair) per warp pass (phase B) with the
// sensor columns staged by cp.async two environments ahead; lanes without a sensor run the same straight-line
// arithmetic and are masked at reductions, ballots and stores.  The parity build is the SAME source, so the code
// checked bit for bit against the CPU restatement (tests/) is the code that is benchmarked.  No tensor-core work:
// the path is elementwise + warp reductions, bounded by HBM traffic and instruction issue.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <cstddef>
#include <type_traits>

#include "philox.cuh"
#include "uavsim.h"

namespace uavsim {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kMaxObsDim = 3 + 5 * UAVSIM_MAX_SENSORS;  // 323 floats
constexpr unsigned kFull = 0xffffffffu;

template <typename Real>
struct alignas(4 * sizeof(Real)) SensorState {
    Real buf, tx, lost, ema;  // data_buffer, total_data_transmitted, total_data_lost, avg_rssi (NaN = None)
};

// Everything the kernels need from uavsim_config, converted once on the host.
template <typename Real>
struct Params {
    // iot_sensors.py / uav_env.py constants
    Real rate, maxbuf, thr, duty_frac, cd, tx_power, noise_floor, lambda, one_minus_lambda, sigma, capture_db;
    Real c_fspl, c_ht, c_hr;  // 20*log10(868), 20*log10(0.5), 20*log10(alt)
    // ADR ladder thresholds -60 / -70 / -78 / -85 dBm (iot_sensors.py:32-37) in the build's RSSI scale.  The float32
    // build works in dB ABOVE the range threshold (rssi - rssi_threshold: tx_power, noise_floor, thr and these four
    // are shifted on the host, the stored avg_rssi too): the values every discrete compare looks at are then within
    // ~30 of zero instead of near -85, so one float32 ulp is <= 1.9e-6 dB at -60 dBm, 5e-7 dB at -78 dBm and next
    // to nothing at the range threshold itself, instead of 7.6e-6 dB everywhere.  rssi_shift = rssi_threshold there,
    // 0 in the parity build.
    Real adr_t60, adr_t70, adr_t78, adr_t85, rssi_shift;
    float det_near, det_far;  // float32 build: constants of the two path-loss branches, folded in float64 on the host
    double d_break;           // (4*pi*0.5*alt)/0.345
    float alt2;               // altitude^2 as the float32 the reference adds (iot_sensors.py:167)
    float d_break_sq;         // throughput build compares squared distances
    // RewardFunction
    Real r_per_byte, r_new, r_completion, r_urg_red, r_move;
    Real p_revisit, p_boundary, p_collision, p_battery, p_hover, p_step, p_data_loss, p_starvation, p_unvisited,
        p_starved, cr_thr;
    Real prox_eta, jain_bonus;
    // UAV energy (always float64: a handful of per-env scalars)
    double max_battery, batt_floor, e_move, e_fail, e_hover, used_hover;
    float inv_max_battery, inv_maxbuf, inv_rate;
    double min_start_dist;
    int32_t max_steps, init_sf, fps, obs_dim, obs_stride, P, domain_rand, auto_reset, num_grids, max_start_tries;
    int32_t relational, rel_dwell_steps, rel_min_cluster;  // RelationalUAVEnv wrapper (env_wrapper.py:47-62)
    Real rel_gamma, rel_dwell_bonus, rel_scale;
    Real gps_sigma;     // GPSNoisyEnv (sim_to_real_analysis.py:170-189), 0 = off
    int32_t gps_in_obs; // PerturbedDQNEnv (sim_to_real_sweep.py:121-136): the noise moves the POSITION the observation pass sees
    double wind_drift;  // WindDriftEnv (:192-224), 0 = off
    int32_t frame_k, frame_mirror;  // frame ring: K = stack + latency frames; offset (floats) of the mirrored copy, 0 when K == 1
    int32_t frame_window, latency;  // VecFrameStack window k (the OLDEST k frames of the ring); LatencyEnv delay
    int32_t grids[UAVSIM_MAX_GRIDS][2];
    uint32_t key0, key1;
    PhiloxKeys keys;  // round-key schedule of (key0, key1)
    int64_t env_id_offset;
};

struct Buffers {
    // state (arena)
    uavsim_env_record* rec;
    void* sstate;  // SensorState<Real>[E][P]
    float2* spos;  // [E][P]
    void* sgen;    // Real[E][P]
    uint8_t* ssf;  // [E][P]
    // outputs
    float* obs;
    void* reward;
    uint8_t* done;
    float* terminal_obs;
    double* episode_stats;
    uint8_t* action_mask;
    double* step_detail;
    void* draw_export;
    double* terminal_info;    // [E][UAVSIM_INFO_COUNT], nullable
    double* terminal_ratios;  // [E][P], nullable
    // replay
    const double* step_draws;
    const double* reset_draws;
    const double* reset_fill;
    const int32_t* reset_init_sf;
    const float* rs_sensor_xy;
    const int32_t* rs_num_sensors;
    const int32_t* rs_grid_wh;
    const float* rs_start_xy;
    const double* dwell_draws;  // [E][P] relational wrapper's range probes (replay)
    const double* step_env_draws;   // [E][4] sim-to-real wrappers, step: gps_x, gps_y (replay)
    const double* reset_env_draws;  // [E][4] sim-to-real wrappers, reset: gps_x, gps_y, wind angle (replay)
    // per-environment overrides of config.shadowing_std_db / duty_cycle (a whole robustness sweep in one batch), nullable
    const float* env_sigma;         // [E] dB
    const float* env_duty;          // [E] percent
    // uavsim_step_ex: step only the envs whose mask byte is set / the PerturbedDQNEnv extra step (reward dropped)
    const uint8_t* step_mask;
    int32_t extra_step;
    // bookkeeping
    double* stat_sums;
    unsigned long long* error_word;  // min over (step_serial << 32 | env) of offending envs, ~0 = none
    int32_t E;
    uint32_t step_serial;
    int32_t zero_episode;  // reset launch after uavsim_set_seed: episode counters restart at 0 (a seeded reset replays)
    int32_t frame_slot;  // row of the (2k-row mirrored ring) frame block this launch writes
    int32_t tile;        // environments per tile of this launch (2, 4, 8 or 16 <= kTile; chosen by the launcher)
    unsigned* tile_ticket;  // [0] tiles handed out, [1] blocks finished; zeroed by the last block; outside the arena
};

// Frame stacking (VecFrameStack, dqn.py:1278; extractor layout dqn.py:612-620) without a shift pass: each
// environment owns 2k frame rows; the frame of a step is written to rows s and s+k (s = step mod k), so the
// k most recent frames, oldest first, are always the CONTIGUOUS rows s+1 .. s+k -- the host exposes that
// window as a strided zero-copy view.  With k == 1 there is one row and no mirror.
template <typename Real>
__device__ __forceinline__ float* obs_frame(const Params<Real>& p, const Buffers& b, int e) {
    return b.obs + ((uint32_t)e * (uint32_t)p.obs_stride + (uint32_t)(b.frame_slot * p.obs_dim));  // < 2^32 floats (checked at create)
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T shfl(T v, int src) {
    return __shfl_sync(kFull, v, src);
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        T u = __shfl_xor_sync(kFull, v, o);
        v = u > v ? u : v;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        T u = __shfl_xor_sync(kFull, v, o);
        v = u < v ? u : v;
    }
    return v;
}

// numpy's pairwise summation for n <= 128 (numpy/_core/src/umath/loops_utils.h.src): what np.sum /
// np.mean / np.var do on the reference's small arrays.  Uniform across the warp (reads shared memory).
template <typename T>
__device__ __forceinline__ T np_pairwise(const T* a, int n) {
    if (n < 8) {
        T r = T(0);
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    T r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
        r0 += a[i + 0];
        r1 += a[i + 1];
        r2 += a[i + 2];
        r3 += a[i + 3];
        r4 += a[i + 4];
        r5 += a[i + 5];
        r6 += a[i + 6];
        r7 += a[i + 7];
    }
    T res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += a[i];
    return res;
}

// a / b: IEEE division in the parity build, MUFU.RCP-based __fdividef in the throughput build
template <typename Real>
__device__ __forceinline__ Real fdiv(Real a, Real b) {
    if constexpr (std::is_same<Real, double>::value)
        return a / b;
    else
        return __fdividef(a, b);
}

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// x / c for a per-handle constant c: IEEE division in the parity build, multiply by the precomputed
// reciprocal in the throughput build
template <typename Real>
__device__ __forceinline__ Real div_const(Real x, Real c, float inv_c) {
    if constexpr (std::is_same<Real, double>::value)
        return x / c;
    else
        return x * inv_c;
}

enum SumOrder { kSumNumpy, kSumSequential };

// Sum of one value per sensor slot over the N real sensors of this warp's environment.
//   parity build   : staged through shared memory and added in the reference's order
//                    (numpy pairwise, or left-to-right for Python loops) -> bit-identical floats
//   throughput build: lane-local add + xor-shuffle tree
template <typename Real, int SPL, typename T>
__device__ __forceinline__ T sensor_sum(const T (&v)[SPL], int N, int lane, double* scratch, SumOrder order) {
    if constexpr (std::is_same<Real, double>::value) {
        T* s = reinterpret_cast<T*>(scratch);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const int j = lane + 32 * k;
            if (j < N) s[j] = v[k];
        }
        __syncwarp();
        T r;
        if (order == kSumNumpy) {
            r = np_pairwise(s, N);
        } else {
            r = T(0);
            for (int i = 0; i < N; ++i) r += s[i];
        }
        __syncwarp();
        return r;
    } else {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < SPL; ++k) acc += (lane + 32 * k < N) ? v[k] : T(0);
        return warp_sum(acc);
    }
}

// LoRa tables indexed by SF - 7 (throughput build: one indexed constant load instead of a per-lane switch ladder)
__device__ __constant__ float kDataRateF32[6] = {5470 / 8.0f, 3125 / 8.0f, 1760 / 8.0f, 980 / 8.0f, 440 / 8.0f, 250 / 8.0f};
__device__ __constant__ float kRequiredSnrF32[6] = {-6.0f, -9.0f, -12.0f, -15.0f, -17.5f, -20.0f};

template <typename Real>
__device__ __forceinline__ Real data_rate_of(int sf) {  // iot_sensors.py:13-20
    if constexpr (!std::is_same<Real, double>::value) return kDataRateF32[min(max(sf - 7, 0), 5)];
    switch (sf) {
        case 7: return Real(5470 / 8.0);
        case 8: return Real(3125 / 8.0);
        case 9: return Real(1760 / 8.0);
        case 10: return Real(980 / 8.0);
        case 11: return Real(440 / 8.0);
        default: return Real(250 / 8.0);
    }
}
template <typename Real>
__device__ __forceinline__ Real required_snr_of(int sf) {  // iot_sensors.py:22-29 (:199-200 default)
    if constexpr (!std::is_same<Real, double>::value) return kRequiredSnrF32[min(max(sf - 7, 0), 5)];
    switch (sf) {
        case 7: return Real(-6.0);
        case 8: return Real(-9.0);
        case 9: return Real(-12.0);
        case 10: return Real(-15.0);
        case 11: return Real(-17.5);
        case 12: return Real(-20.0);
        default: return Real(7.5);
    }
}
template <typename Real>
__device__ __forceinline__ float link_quality_of(int sf) {  // uav_env.py:647
    if constexpr (!std::is_same<Real, double>::value) return sf >= 12 ? 0.1f : (float)(12 - sf) * 0.2f;
    switch (sf) {
        case 7: return 1.0f;
        case 8: return 0.8f;
        case 9: return 0.6f;
        case 10: return 0.4f;
        case 11: return 0.2f;
        default: return 0.1f;
    }
}

// Deterministic part of IoTSensor.calculate_rssi (iot_sensors.py:161-187).  Evaluated ONCE per
// sensor per step: the UAV does not move between the up-to-five draws a step makes on a sensor.
template <typename Real>
__device__ __forceinline__ Real det_rssi(const Params<Real>& p, float ux, float uy, float sx, float sy) {
    float dx = ux - sx, dy = uy - sy;
    if constexpr (std::is_same<Real, double>::value) {
        dx *= 10.0f;  // iot_sensors.py:161-162, float32
        dy *= 10.0f;
        // reference op order in float32 (this TU is built with -fmad=false, so no contraction)
        const float g = sqrtf(dx * dx + dy * dy);
        const float d = sqrtf(g * g + p.alt2);
        const float lg = (float)log10((double)d);  // correctly rounded float32 log10 (normalised definition)
        double pl;
        if ((double)d < p.d_break) {
            const float t = 20.0f * lg;
            pl = ((double)t + p.c_fspl) - 28;
        } else {
            const float t = 40.0f * lg;
            pl = ((double)t - p.c_ht) - p.c_hr;
        }
        return p.tx_power - pl;
    } else {
        // log10(d) = 0.5*log10(d^2): one MUFU.LG2, no square roots
        const float d2 = fmaf(fmaf(dx, dx, dy * dy), 100.0f, p.alt2);  // (10 dx)^2 + (10 dy)^2 + altitude^2
        const float l2 = __log2f(d2);
        const float k = 0.30102999566398120f * 0.5f;  // log10(2)/2
        return d2 < p.d_break_sq ? fmaf(-20.0f * k, l2, p.det_near) : fmaf(-40.0f * k, l2, p.det_far);
    }
}

// IoTSensor.update_spreading_factor (iot_sensors.py:223-259).  "No threshold matched" keeps the SF.
template <typename Real>
__device__ __forceinline__ void adr_update(const Params<Real>& p, Real rssi, Real& ema, int& sf) {
    // Straight-line selects, no per-lane branches: lanes of a warp hold different sensors and an if / else-if
    // chain compiles to a divergent branch ladder (BSSY / BRA / BSYNC per update).
    Real blended;
    if constexpr (std::is_same<Real, double>::value)
        blended = (p.lambda * rssi) + (p.one_minus_lambda * ema);  // reference order, no contraction (-fmad=false)
    else
        blended = fmaf(p.lambda, rssi, p.one_minus_lambda * ema);  // explicit: a*b + c*d has two contractions, and
                                                                    // instantiations must agree bit for bit
    ema = (ema != ema) ? rssi : blended;  // NaN = None: the first sample initialises the average
    // the reference takes the FIRST match of the descending ladder -60 / -70 / -78 / -85 (iot_sensors.py:252-255),
    // which is the LAST assignment of the ascending one
#ifdef UAVSIM_ADR_IMM  // A/B switch: thresholds as immediates (valid for rssi_threshold = -85 only)
    if constexpr (!std::is_same<Real, double>::value) {
        sf = ema > 0.0f ? 12 : sf;
        sf = ema > 7.0f ? 11 : sf;
        sf = ema > 15.0f ? 9 : sf;
        sf = ema > 25.0f ? 7 : sf;
        return;
    }
#endif
    sf = ema > p.adr_t85 ? 12 : sf;
    sf = ema > p.adr_t78 ? 11 : sf;
    sf = ema > p.adr_t70 ? 9 : sf;
    sf = ema > p.adr_t60 ? 7 : sf;
}

// UAVEnvironment._calculate_urgency (uav_env.py:376-384)
template <typename Real>
__device__ __forceinline__ Real urgency_of(const Params<Real>& p, Real buf, Real lost, Real gen) {
    const Real util = div_const(buf, p.maxbuf, p.inv_maxbuf);
    Real lr;
    if constexpr (std::is_same<Real, double>::value)
        lr = gen > 0.0 ? lost / gen : 0.0;
    else
        lr = lost * fast_rcp(fmaxf(gen, 1e-30f));  // gen == 0 implies lost == 0: 0 * 1e30 = 0 without a select
    const Real u = util * (Real(1) + lr * Real(10));
    if constexpr (std::is_same<Real, double>::value)
        return u < Real(0) ? Real(0) : (u > Real(1) ? Real(1) : u);
    else
        return __saturatef(u);  // one .SAT modifier on the producing multiply
}

// ------------------------------------------------------------------------------------------------
// draws
// ------------------------------------------------------------------------------------------------
template <typename Real>
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, Real sigma, Real& z0, Real& z1) {
    if constexpr (std::is_same<Real, double>::value) {
        const double u1 = ((double)a + 0.5) * 2.3283064365386963e-10;  // (0,1)
        const double u2 = ((double)b + 0.5) * 2.3283064365386963e-10;
        const double r = sigma * sqrt(-2.0 * log(u1));
        double s, c;
        sincospi(2.0 * u2, &s, &c);
        z0 = r * c;
        z1 = r * s;
    } else {
        const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (0,1]
        const float u2 = (float)b * 2.3283064365386963e-10f;
        const float r = sigma * fast_sqrt(-1.3862943611198906f * __log2f(u1));  // -2 ln u = -2 ln2 lg2 u
        float s, c;
        __sincosf(6.2831853071795865f * u2, &s, &c);
        // __fmul_rn is never contracted: `det + z` stays an add in every instantiation (an FMA fused at one call
        // site and not at another would make the LEAN and the diagnostic kernels differ in the last bit)
        z0 = __fmul_rn(r, c);
        z1 = __fmul_rn(r, s);
    }
}

template <typename Real>
__device__ __forceinline__ Real u01(uint32_t a) {
    if constexpr (std::is_same<Real, double>::value)
        return ((double)a + 0.5) * 2.3283064365386963e-10;
    else
        return fminf(fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f), 0.99999994f);
}

// Draws of one environment-step.  Two views, generated where they are consumed so that they are not
// live across the whole step:
//   obs(lane, d, e)        D (ADR) and E (range check) of the observation pass   uav_env.py:654-660
//   collect(lane, a, b, u) A (ADR), B (link probability), U (duty-cycle uniform) uav_env.py:535-551
//   winner(lane, k)        C, the range check of an SF-group winner              iot_sensors.py:131-133
// REPLAY: slot-addressed table [6][P] of this environment.
// PHILOX: counter = (env id, stream | env hi, episode, step << 8 | lane << 3 | call), key = seed.  One
// Philox4x32-10 call yields four words = two Box-Muller pairs:
//   call 0: D,E of slot 0 (x,y) and of slot 1 (z,w)          -> a MOVE step costs ONE call per lane
//   call 1: A,B of slot 0 (x,y) and of slot 1 (z,w); with one slot per lane U = z
//   call 2: U of slot 0 (x) and slot 1 (y)                   (two slots per lane only)
//   call 4+k: C of slot k (x,y), generated lazily by the owning lane
template <typename Real, bool REPLAY, int SPL>
struct Draws {
    const double* tab;  // replay table of this environment (REPLAY only)
    Real* exp;          // optional export of the draws actually used, [6][P]
    uint32_t env_lo, env_hi, episode, step;
    // P <= 32 (one slot per lane): the D, E draws of a STEP come from ONE Philox call per environment PAIR
    // (global ids 2m, 2m+1): counter = (m, kStreamPair | hi bits, episode and step of the EVEN environment,
    // lane << 3); words (x, y) serve the even environment, (z, w) the odd one.  A call then feeds two sensors
    // whatever the sensor count, and the pair kernel (two environments per warp pass) issues it once.
    uint32_t pair_episode, pair_step;
    bool paired;
    Real sigma;  // shadowing std of this environment (config.shadowing_std_db, or its per-env override)

    __device__ __forceinline__ Philox4 call(const Params<Real>& p, int lane, uint32_t c) const {
        return philox4x32_10(env_lo, env_hi | (kStreamStep << 24), episode, (step << 8) | ((uint32_t)lane << 3) | c,
                             p.keys);
    }
    __device__ __forceinline__ Philox4 pair_call(const Params<Real>& p, int lane) const {
        return philox4x32_10((env_lo >> 1) | (env_hi << 31), (env_hi >> 1) | (kStreamPair << 24), pair_episode,
                             (pair_step << 8) | ((uint32_t)lane << 3), p.keys);
    }
    __device__ __forceinline__ void put(const Params<Real>& p, int slot, int j, Real v) const {
        if (exp && j < p.P) exp[slot * p.P + j] = v;
    }
    __device__ __forceinline__ void obs(const Params<Real>& p, int lane, Real (&d)[SPL], Real (&e)[SPL]) const {
        if constexpr (REPLAY) {
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                const int j = lane + 32 * k;
                const bool in = j < p.P;
                d[k] = in ? (Real)tab[UAVSIM_SLOT_D * p.P + j] : Real(0);
                e[k] = in ? (Real)tab[UAVSIM_SLOT_E * p.P + j] : Real(0);
            }
        } else {
            if constexpr (SPL == 2) {
                const Philox4 q0 = call(p, lane, 0);
                box_muller<Real>(q0.x, q0.y, sigma, d[0], e[0]);
                box_muller<Real>(q0.z, q0.w, sigma, d[1], e[1]);
            } else {
                const Philox4 q0 = paired ? pair_call(p, lane) : call(p, lane, 0);
                const bool odd = paired && (env_lo & 1u);
                box_muller<Real>(odd ? q0.z : q0.x, odd ? q0.w : q0.y, sigma, d[0], e[0]);
            }
            if (exp) {
#pragma unroll
                for (int k = 0; k < SPL; ++k) {
                    put(p, UAVSIM_SLOT_D, lane + 32 * k, d[k]);
                    put(p, UAVSIM_SLOT_E, lane + 32 * k, e[k]);
                }
            }
        }
    }
    __device__ __forceinline__ void collect(const Params<Real>& p, int lane, Real (&a)[SPL], Real (&b)[SPL],
                                            Real (&u)[SPL]) const {
        if constexpr (REPLAY) {
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                const int j = lane + 32 * k;
                const bool in = j < p.P;
                a[k] = in ? (Real)tab[UAVSIM_SLOT_A * p.P + j] : Real(0);
                b[k] = in ? (Real)tab[UAVSIM_SLOT_B * p.P + j] : Real(0);
                u[k] = in ? (Real)tab[UAVSIM_SLOT_U * p.P + j] : Real(1);
            }
        } else {
            const Philox4 q1 = call(p, lane, 1);
            box_muller<Real>(q1.x, q1.y, sigma, a[0], b[0]);
            if constexpr (SPL == 1) {
                u[0] = u01<Real>(q1.z);
            } else {
                box_muller<Real>(q1.z, q1.w, sigma, a[1], b[1]);
                const Philox4 q2 = call(p, lane, 2);
                u[0] = u01<Real>(q2.x);
                u[1] = u01<Real>(q2.y);
            }
            if (exp) {
#pragma unroll
                for (int k = 0; k < SPL; ++k) {
                    put(p, UAVSIM_SLOT_A, lane + 32 * k, a[k]);
                    put(p, UAVSIM_SLOT_B, lane + 32 * k, b[k]);
                    put(p, UAVSIM_SLOT_U, lane + 32 * k, u[k]);
                }
            }
        }
    }
    // RelationalUAVEnv._cluster_dwell_bonus: one is_in_range probe per sensor on a COLLECT step (env_wrapper.py:303)
    __device__ __forceinline__ void dwell(const Params<Real>& p, const double* table, int lane, Real (&f)[SPL]) const {
        if constexpr (REPLAY) {
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                const int j = lane + 32 * k;
                f[k] = (table && j < p.P) ? (Real)table[j] : Real(0);
            }
        } else {
            const Philox4 q = call(p, lane, 3);
            Real unused;
            box_muller<Real>(q.x, q.y, sigma, f[0], unused);
            if constexpr (SPL == 2) box_muller<Real>(q.z, q.w, sigma, f[1], unused);
        }
    }
    __device__ __forceinline__ Real winner(const Params<Real>& p, int lane, int k) const {
        const int j = lane + 32 * k;
        if constexpr (REPLAY) {
            return (Real)tab[UAVSIM_SLOT_C * p.P + j];
        } else {
            const Philox4 q = call(p, lane, 4u + (uint32_t)k);
            Real c, unused;
            box_muller<Real>(q.x, q.y, sigma, c, unused);
            put(p, UAVSIM_SLOT_C, j, c);
            return c;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// per-warp shared memory and per-lane sensor registers
// ------------------------------------------------------------------------------------------------
struct alignas(16) WarpSmem {
    uavsim_env_record rec;           // this warp's environment record: one coalesced 128-byte load/store
    double red[UAVSIM_MAX_SENSORS];  // parity-build ordered reductions; far-start candidate search
};

// Scalars every lane needs in registers during a step (the rest of the record stays in shared memory).
struct Hot {
    float x, y, elapsed, prev_dist;
    double battery;
    int W, H, N, step;
    uint32_t vis_lo, vis_hi, episode;
};

__device__ __forceinline__ void load_record(WarpSmem& ws, const uavsim_env_record* g, int lane) {
    __syncwarp();
    reinterpret_cast<uint32_t*>(&ws.rec)[lane] = reinterpret_cast<const uint32_t*>(g)[lane];
    __syncwarp();
}
__device__ __forceinline__ void store_record(const WarpSmem& ws, uavsim_env_record* g, int lane) {
    __syncwarp();
    reinterpret_cast<uint32_t*>(g)[lane] = reinterpret_cast<const uint32_t*>(&ws.rec)[lane];
}
__device__ __forceinline__ Hot read_hot(const uavsim_env_record& R) {
    Hot h;
    h.x = R.x;
    h.y = R.y;
    h.elapsed = R.elapsed_s;
    h.prev_dist = R.prev_dist_nearest;
    h.battery = R.battery;
    h.W = R.grid_w;
    h.H = R.grid_h;
    h.N = R.num_sensors;
    h.step = R.step;
    h.vis_lo = R.visited_lo;
    h.vis_hi = R.visited_hi;
    h.episode = R.episode;
    return h;
}

template <typename Real>
struct Sensor {
    Real buf, tx, lost, ema, gen;  // gen: parity = total_data_generated; throughput = value at reset
    Real det, cur;                 // deterministic RSSI here; current_rssi of draw A
    float sx, sy;
    int sf;
    bool active, attempt;
};

template <typename Real>
__device__ __forceinline__ Real generated_of(const Params<Real>& p, const Sensor<Real>& s, float elapsed) {
    if constexpr (std::is_same<Real, double>::value)
        return s.gen;
    else
        return fmaf(p.rate, elapsed, s.gen);  // total_data_generated = at_reset + rate * sum(dt)
}

// Throughput layout: the spreading factor rides in the three low mantissa bits of avg_rssi (code sf-7),
// so a sensor's mutable state is ONE 16-byte vector.  avg_rssi is a decaying average of values with
// 4 dB of noise; an 8-ulp (6e-5 dB) rounded quantisation per step is far below anything the model
// resolves, and NaN (= None) keeps its payload bits.  The parity layout keeps the separate u8 column.
// None is the canonical quiet NaN 0x7fc00000 (none_value): rounding to 8 ulp maps it onto itself, so packing needs no
// special case.
__device__ __forceinline__ float pack_sf(float ema, int sf) {
    const uint32_t b = (__float_as_uint(ema) + 4u) & ~7u;  // round to a multiple of 8 ulp
    return __uint_as_float(b | (uint32_t)(sf - 7));
}
__device__ __forceinline__ void unpack_sf(float packed, float& ema, int& sf) {
    const uint32_t b = __float_as_uint(packed);
    sf = (int)(b & 7u) + 7;
    ema = __uint_as_float(b & ~7u);
}

// avg_rssi = None
template <typename Real>
__device__ __forceinline__ Real none_value() {
    if constexpr (std::is_same<Real, double>::value)
        return CUDART_NAN;
    else
        return __uint_as_float(0x7fc00000u);
}

template <typename Real, int SPL>
__device__ __forceinline__ void load_sensors(const Buffers& b, const Params<Real>& p, int e, int N, int lane,
                                             Sensor<Real> (&s)[SPL]) {
    constexpr bool kParity = std::is_same<Real, double>::value;
    const uint32_t row = (uint32_t)e * (uint32_t)p.P;  // E*P < 2^32 (checked at create)
    const SensorState<Real>* st = reinterpret_cast<const SensorState<Real>*>(b.sstate);
    const Real* gen = reinterpret_cast<const Real*>(b.sgen);
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const int j = lane + 32 * k;
        s[k].active = j < N;
        s[k].attempt = false;
        s[k].det = Real(0);
        s[k].cur = Real(0);
        // inactive lanes carry benign values and run the same arithmetic; they are masked at every
        // reduction, ballot and store
        SensorState<Real> v;
        v.buf = v.tx = v.lost = Real(0);
        v.ema = none_value<Real>();
        float2 xy = make_float2(0.f, 0.f);
        Real g = Real(0);
        int sf = 12;
        if (s[k].active) {
            const uint32_t i = row + (uint32_t)j;
            v = st[i];
            xy = b.spos[i];
            g = gen[i];
            if constexpr (kParity) sf = b.ssf[i];
        }
        if constexpr (!kParity) {
            float ema;
            unpack_sf(v.ema, ema, sf);
            if (!s[k].active) sf = 12;
            v.ema = ema;
        }
        s[k].buf = v.buf;
        s[k].tx = v.tx;
        s[k].lost = v.lost;
        s[k].ema = v.ema;
        s[k].gen = g;
        s[k].sx = xy.x;
        s[k].sy = xy.y;
        s[k].sf = sf;
    }
}

template <typename Real, int SPL>
__device__ __forceinline__ void store_sensors(const Buffers& b, const Params<Real>& p, int e, int lane,
                                              const Sensor<Real> (&s)[SPL], bool write_static) {
    constexpr bool kParity = std::is_same<Real, double>::value;
    const uint32_t row = (uint32_t)e * (uint32_t)p.P;
    SensorState<Real>* st = reinterpret_cast<SensorState<Real>*>(b.sstate);
    Real* gen = reinterpret_cast<Real*>(b.sgen);
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const int j = lane + 32 * k;
        const uint32_t i = row + (uint32_t)j;
        SensorState<Real> v;  // packed by every lane; only the stores are predicated
        v.buf = s[k].buf;
        v.tx = s[k].tx;
        v.lost = s[k].lost;
        if constexpr (kParity)
            v.ema = s[k].ema;
        else
            v.ema = pack_sf(s[k].ema, s[k].sf);
        if (s[k].active) {
            if constexpr (kParity) b.ssf[i] = (uint8_t)s[k].sf;
            st[i] = v;
            // parity build accumulates total_data_generated; throughput build only rewrites it on reset
            if (kParity || write_static) gen[i] = s[k].gen;
            if (write_static) b.spos[i] = make_float2(s[k].sx, s[k].sy);
        }
    }
}

__device__ __forceinline__ uint8_t action_mask_of(float x, float y, int W, int H) {  // derived: uav.py:175
    const float fw = (float)W, fh = (float)H;
    uint8_t m = 1u << 4;
    const bool xin = (0.f <= x) && (x < fw), yin = (0.f <= y) && (y < fh);
    if (xin && 0.f <= y + 1.0f && y + 1.0f < fh) m |= 1u << 0;
    if (xin && 0.f <= y - 1.0f && y - 1.0f < fh) m |= 1u << 1;
    if (yin && 0.f <= x - 1.0f && x - 1.0f < fw) m |= 1u << 2;
    if (yin && 0.f <= x + 1.0f && x + 1.0f < fw) m |= 1u << 3;
    return m;
}

// UAVEnvironment._get_observation (uav_env.py:638-674) + zero padding to obs_slots (dqn.py:286-298).
// Mutates EMA / SF (draw D) and range-checks (draw E).  Each lane stores its sensor's features straight
// into the row (the partial sectors of neighbouring lanes merge in L2; DRAM sees the row once); ghost
// slots are zero-filled when the buffer is bound / an episode starts and never touched by a step (N is
// fixed during an episode).
template <typename Real, int SPL, int FPS, bool HEAD, bool MIRROR>
__device__ __forceinline__ void observe_with(const Params<Real>& p, const Hot& h, Sensor<Real> (&s)[SPL],
                                             const Real (&dd)[SPL], const Real (&de)[SPL], int lane,
                                             float* __restrict__ out);

template <typename Real, bool REPLAY, int SPL, int FPS, bool HEAD, bool MIRROR>
__device__ __forceinline__ void observe(const Params<Real>& p, const Hot& h, Sensor<Real> (&s)[SPL],
                                        const Draws<Real, REPLAY, SPL>& dr, int lane, float* __restrict__ out) {
    Real dd[SPL], de[SPL];
    dr.obs(p, lane, dd, de);
    observe_with<Real, SPL, FPS, HEAD, MIRROR>(p, h, s, dd, de, lane, out);
}

// The observation pass with the D / E draws already in hand (the pair kernel makes one Philox call for two
// environments and runs this once per environment).
template <typename Real, int SPL, int FPS, bool HEAD, bool MIRROR>
__device__ __forceinline__ void observe_with(const Params<Real>& p, const Hot& h, Sensor<Real> (&s)[SPL],
                                             const Real (&dd)[SPL], const Real (&de)[SPL], int lane,
                                             float* __restrict__ out) {
    constexpr bool kParity = std::is_same<Real, double>::value;
    const int mirror = MIRROR ? p.frame_mirror : 0;
    float inv_w = 0.f, inv_h = 0.f;
    if constexpr (!kParity && FPS == 5) {
        inv_w = fast_rcp((float)h.W);
        inv_h = fast_rcp((float)h.H);
    }
    if (HEAD && lane == 0) {  // the step writes these three in its per-lane phase C
        if constexpr (kParity) {
            out[0] = (float)((double)h.x / (double)h.W);
            out[1] = (float)((double)h.y / (double)h.H);
            out[2] = (float)(h.battery / p.max_battery);
        } else {
            out[0] = __fdividef(h.x, (float)h.W);
            out[1] = __fdividef(h.y, (float)h.H);
            out[2] = (float)h.battery * p.inv_max_battery;
        }
    }
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const int j = lane + 32 * k;
        const Real urg = urgency_of(p, s[k].buf, s[k].lost, generated_of(p, s[k], h.elapsed));
        adr_update(p, s[k].det + dd[k], s[k].ema, s[k].sf);
        const bool in_range = (s[k].det + de[k]) >= p.thr;
        {   // features computed by every lane (no per-lane branch); only the stores are predicated
            float* o = out + 3 + FPS * j;
            float f[FPS];
            f[0] = (float)div_const(s[k].buf, p.maxbuf, p.inv_maxbuf);
            f[1] = (float)urg;
            f[2] = in_range ? link_quality_of<Real>(s[k].sf) : 0.0f;
            if constexpr (FPS == 5) {  // include_sensor_positions, uav_env.py:668-672
                float rx, ry;
                if constexpr (kParity) {
                    double dx = ((double)s[k].sx - (double)h.x) / (double)h.W, dy = ((double)s[k].sy - (double)h.y) / (double)h.H;
                    if (p.relational) {  // np.clip(., -1, 1), env_wrapper.py:217-218
                        dx = dx < -1.0 ? -1.0 : (dx > 1.0 ? 1.0 : dx);
                        dy = dy < -1.0 ? -1.0 : (dy > 1.0 ? 1.0 : dy);
                    }
                    rx = (float)dx;
                    ry = (float)dy;
                } else {
                    rx = (s[k].sx - h.x) * inv_w;
                    ry = (s[k].sy - h.y) * inv_h;
                    if (p.relational) {
                        rx = fminf(fmaxf(rx, -1.f), 1.f);
                        ry = fminf(fmaxf(ry, -1.f), 1.f);
                    }
                }
                if (p.relational) {  // Dict "sensors" row: [rel_x, rel_y, buffer, urgency, visited] (env_wrapper.py:216-222)
                    const bool seen = j < 32 ? (h.vis_lo >> j) & 1u : (h.vis_hi >> (j - 32)) & 1u;
                    f[2] = f[0];
                    f[3] = f[1];
                    f[0] = rx;
                    f[1] = ry;
                    f[4] = seen ? 1.0f : 0.0f;
                } else {
                    f[3] = rx;
                    f[4] = ry;
                }
            }
            if (s[k].active) {
#pragma unroll
                for (int q = 0; q < FPS; ++q) {
                    o[q] = f[q];
                    if constexpr (MIRROR) o[q + mirror] = f[q];
                }
            }
        }
    }
}

// DomainRandEnv._dist_to_nearest_with_data (dqn.py:405-412): float32 norm, min over sensors with data.
template <typename Real, int SPL>
__device__ __forceinline__ float nearest_with_data(float x, float y, const Sensor<Real> (&s)[SPL]) {
    float best = CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        if (s[k].active && s[k].buf > Real(0)) {
            const float dx = s[k].sx - x, dy = s[k].sy - y;
            best = fminf(best, sqrtf(dx * dx + dy * dy));
        }
    }
    best = warp_min(best);
    return best == CUDART_INF_F ? 0.0f : best;
}

// RelationalUAVEnv._compute_potential (env_wrapper.py:270-289): float32 distance to the nearest UNVISITED
// sensor, 0 when none is left; the potential is -dist / grid diagonal.
template <typename Real, int SPL>
__device__ __forceinline__ float nearest_unvisited(float x, float y, uint32_t vis_lo, uint32_t vis_hi, int lane,
                                                   const Sensor<Real> (&s)[SPL]) {
    float best = CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const bool seen = k == 0 ? (vis_lo >> lane) & 1u : (vis_hi >> lane) & 1u;
        if (s[k].active && !seen) {
            const float dx = s[k].sx - x, dy = s[k].sy - y;
            best = fminf(best, sqrtf(dx * dx + dy * dy));
        }
    }
    best = warp_min(best);
    return best == CUDART_INF_F ? 0.0f : best;
}

template <typename Real, int SPL>
__device__ __forceinline__ void compute_det(const Params<Real>& p, float x, float y, Sensor<Real> (&s)[SPL]) {
#pragma unroll
    for (int k = 0; k < SPL; ++k) s[k].det = det_rssi(p, x, y, s[k].sx, s[k].sy);
}

}  // namespace uavsim
