// Fused batched UAV-LoRa environment step for sm_100a: ONE WARP PER ENVIRONMENT, sensors across
// lanes (slot j = lane + 32*k, k < SPL), structure-of-arrays state in HBM.
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
// The kernel is a template over
//   Real   : double = PARITY build (float64 accumulators like the reference, reference operation
//            order, numpy's pairwise summation order, this TU compiled with -fmad=false);
//            float  = THROUGHPUT build (float32 state, MUFU intrinsics, shuffle reductions).
//   REPLAY : true  = every random draw is read from a slot-addressed table (parity tests),
//            false = counter-based Philox4x32-10.
//   SPL    : sensor slots per lane (1: P <= 32, 2: P <= 64).
// Execution shape: persistent warps (grid = SMs x resident blocks, each warp walks environments with a
// grid stride), the 128-byte environment record moves through shared memory with one coalesced
// load/store per step, inactive lanes run the same arithmetic and are masked at reductions and stores
// (no divergent branches in the per-sensor code), every draw of a step is produced up front with one
// Philox call serving both of a lane's sensor slots.
// so the code that is parity-checked bit-for-bit against the oracle is the same source as the code
// that is benchmarked.  There is no tensor-core work here: the path is elementwise + warp
// reductions and is bounded by HBM traffic (DESIGN.md has the byte model).
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
    // bookkeeping
    double* stat_sums;
    unsigned long long* error_word;  // min over (step_serial << 32 | env) of offending envs, ~0 = none
    int32_t E;
    uint32_t step_serial;
    int32_t frame_slot;  // row of the (2k-row mirrored ring) frame block this launch writes
    int32_t tile;        // environments per tile of this launch (2, 4, 8 or 16 <= kTile; chosen by the launcher)
    unsigned* tile_ticket;  // two tile counters (this launch / zeroed for the next), outside the arena
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
        return d2 < p.d_break_sq ? fmaf(-20.0f * k, l2, p.tx_power - p.c_fspl + 28.0f)
                                 : fmaf(-40.0f * k, l2, p.tx_power + p.c_ht + p.c_hr);
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
    sf = ema > Real(-85) ? 12 : sf;
    sf = ema > Real(-78) ? 11 : sf;
    sf = ema > Real(-70) ? 9 : sf;
    sf = ema > Real(-60) ? 7 : sf;
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
                box_muller<Real>(q0.x, q0.y, p.sigma, d[0], e[0]);
                box_muller<Real>(q0.z, q0.w, p.sigma, d[1], e[1]);
            } else {
                const Philox4 q0 = paired ? pair_call(p, lane) : call(p, lane, 0);
                const bool odd = paired && (env_lo & 1u);
                box_muller<Real>(odd ? q0.z : q0.x, odd ? q0.w : q0.y, p.sigma, d[0], e[0]);
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
            box_muller<Real>(q1.x, q1.y, p.sigma, a[0], b[0]);
            if constexpr (SPL == 1) {
                u[0] = u01<Real>(q1.z);
            } else {
                box_muller<Real>(q1.z, q1.w, p.sigma, a[1], b[1]);
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
            box_muller<Real>(q.x, q.y, p.sigma, f[0], unused);
            if constexpr (SPL == 2) box_muller<Real>(q.z, q.w, p.sigma, f[1], unused);
        }
    }
    __device__ __forceinline__ Real winner(const Params<Real>& p, int lane, int k) const {
        const int j = lane + 32 * k;
        if constexpr (REPLAY) {
            return (Real)tab[UAVSIM_SLOT_C * p.P + j];
        } else {
            const Philox4 q = call(p, lane, 4u + (uint32_t)k);
            Real c, unused;
            box_muller<Real>(q.x, q.y, p.sigma, c, unused);
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

// Per-environment draws of the sim-to-real wrappers (sim_to_real_analysis.py:176-181, :207): GPS noise pair for
// the observation of `step` (0 = reset observation) and, at reset, the wind angle.  REPLAY: table [E][4];
// PHILOX: stream kStreamMisc, counter word 3 = step << 8 | 4 (words 0, 1 there are the grid / inherited-SF draws).
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
                box_muller<Real>(q.x, q.y, p.sigma, z0, z1);
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
    dr.tab = REPLAY ? b.reset_draws + (size_t)e * UAVSIM_NUM_SLOTS * p.P : nullptr;
    dr.exp = exp;
    float* out = obs_frame(p, b, e);
    if (p.frame_k > 1) {  // a new episode starts with k-1 all-zero frames (VecFrameStack zeroes the stack on done)
        float* blk = b.obs + (size_t)e * p.obs_stride;
        for (int i = lane; i < p.obs_stride; i += 32) blk[i] = 0.0f;
        __syncwarp();
    }
    observe<Real, REPLAY, SPL, FPS, true, false>(p, h, s, dr, lane, out);
    double wind_angle = 0.0;
    if (p.gps_sigma > Real(0) || p.wind_drift > 0.0) {  // GPSNoisyEnv.reset :183-185, WindDriftEnv.reset :205-210
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

// ------------------------------------------------------------------------------------------------
// step
// ------------------------------------------------------------------------------------------------
// (value, index) arg-max over the lanes of one SF group; ties -> lowest sensor index, which is what
// Python's stable sorted(..., reverse=True) yields (uav_env.py:564-566).
template <typename Real>
__device__ __forceinline__ void group_argmax(Real v, int idx, Real& vmax, int& imax) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const Real ov = __shfl_xor_sync(kFull, v, o);
        const int oi = __shfl_xor_sync(kFull, idx, o);
        if (oi >= 0 && (idx < 0 || ov > v || (ov == v && oi < idx))) {
            v = ov;
            idx = oi;
        }
    }
    vmax = v;
    imax = idx;
}

template <typename Real>
struct TileSmem;

enum : int { kGeneric = 0, kLeanPlain = 1, kLeanDomainRand = 2 };

// LEAN instantiations of the step kernel compile the optional features out (DomainRandEnv shaping, the
// relational wrapper, draw export, parity diagnostics): the plain UAVEnvironment step -- the benchmarked
// configuration -- then carries neither their uniform branches nor their code in its instruction stream.
template <int LEAN, typename Real>
__device__ __forceinline__ bool feat_domain_rand(const Params<Real>& p) {
    if constexpr (LEAN == kLeanPlain) return false; else if constexpr (LEAN == kLeanDomainRand) return true; else return p.domain_rand != 0;
}
template <int LEAN, typename Real>
__device__ __forceinline__ bool feat_relational(const Params<Real>& p) {
    if constexpr (LEAN) return false; else return p.relational != 0;
}
template <int LEAN, typename Real>
__device__ __forceinline__ bool feat_gps(const Params<Real>& p) {
    if constexpr (LEAN) return false; else return p.gps_sigma > Real(0);
}
template <int LEAN, typename Real>
__device__ __forceinline__ bool feat_wind(const Params<Real>& p) {
    if constexpr (LEAN) return false; else return p.wind_drift > 0.0;
}
template <int LEAN>
__device__ __forceinline__ double* feat_step_detail(const Buffers& b) {
    if constexpr (LEAN) return nullptr; else return b.step_detail;
}
template <int LEAN>
__device__ __forceinline__ void* feat_draw_export(const Buffers& b) {
    if constexpr (LEAN) return nullptr; else return b.draw_export;
}

// Draws of the step in flight for slot k of the tile, rebuilt at each point of use (cheap) so that nothing
// of it stays live across the step.
template <typename Real, bool REPLAY, int SPL, int LEAN>
__device__ __forceinline__ Draws<Real, REPLAY, SPL> step_draws(const Params<Real>& p, const Buffers& b,
                                                                const TileSmem<Real>& ts, int k, int e);

// UAVEnvironment._execute_collect_action (uav_env.py:518-632) for one environment (slot k of the tile).
// Results go to the tile block (and the parity diagnostics) from lane 0.
template <typename Real, bool REPLAY, int SPL, int LEAN>
__device__ __forceinline__ void collect_action(const Params<Real>& p, const Buffers& b, TileSmem<Real>& ts, WarpSmem& ws,
                                               int k, int e, Sensor<Real> (&s)[SPL], Real step_loss, int lane) {
    const Draws<Real, REPLAY, SPL> dr = step_draws<Real, REPLAY, SPL, LEAN>(p, b, ts, k, e);
    Hot h;
    h.N = ts.N[k];
    h.elapsed = ts.elapsed[k];
    h.vis_lo = ts.vis_lo[k];
    h.vis_hi = ts.vis_hi[k];
    constexpr bool kParity = std::is_same<Real, double>::value;
    const int N = h.N;
    // _get_sensor_urgencies (:386-394) before the collection.  Only winners change their buffer, so the
    // throughput build accumulates the reduction winner by winner; the parity build keeps the reference's
    // float32 np.sum over all sensors.
    float before[kParity ? SPL : 1];
    float urg_red_acc = 0.f;
    if constexpr (kParity) {
#pragma unroll
        for (int k = 0; k < SPL; ++k) before[k] = (float)div_const(s[k].buf, p.rate, p.inv_rate);
    }
    const Real used = (Real)p.used_hover;                                      // uav.py:263, uav_env.py:530
    // Phase 2: ADR update (draw A), link probability (draw B), duty-cycle gate (draw U)
    Real da[SPL], db[SPL], du[SPL];
    dr.collect(p, lane, da, db, du);
    unsigned att[SPL];
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        bool a = false;
        if constexpr (kParity) {
            if (s[k].active && s[k].buf > Real(0)) {
                s[k].cur = s[k].det + da[k];
                adr_update(p, s[k].cur, s[k].ema, s[k].sf);
                const Real rb = s[k].det + db[k];
                Real plink;
                if (rb < p.thr) {
                    plink = Real(0);
                } else {
                    const Real x = (rb - p.noise_floor) - required_snr_of<Real>(s[k].sf);
                    plink = 1.0 / (1.0 + exp(-x));
                }
                a = (plink * p.duty_frac) > du[k];
            }
        } else {  // the same, straight-line: every lane computes, selects keep the state of lanes without data
            const bool has = s[k].active && s[k].buf > 0.f;
            const float cur = s[k].det + da[k];
            float ema2 = s[k].ema;
            int sf2 = s[k].sf;
            adr_update(p, cur, ema2, sf2);
            s[k].cur = cur;  // only read for attempting lanes
            s[k].ema = has ? ema2 : s[k].ema;
            s[k].sf = has ? sf2 : s[k].sf;
            const float rb = s[k].det + db[k];
            const float x = (rb - p.noise_floor) - required_snr_of<float>(sf2);
            const float plink = rb < p.thr ? 0.f : __fdividef(1.0f, 1.0f + __expf(-x));
            a = has && (plink * p.duty_frac) > du[k];
        }
        s[k].attempt = a;
        att[k] = __ballot_sync(kFull, a);
    }
    unsigned long long att_mask = (unsigned long long)att[0];
    if (SPL > 1) att_mask |= (unsigned long long)att[SPL - 1] << 32;

    // Phase 3+4: SF groups in first-appearance order; k==1 wins, k>1 collides unless the strongest beats
    // the runner-up by the capture threshold; each winner is range-checked (draw C) and drained.
    Real total_bytes = Real(0), urg_acc = Real(0);
    int nwin = 0, collisions = 0, captures = 0, delivered = 0;
    bool any_new = false;
    unsigned long long win_mask = 0ull, remaining = att_mask;
    while (remaining) {
        const int i0 = __ffsll((long long)remaining) - 1;
        bool sel[SPL];
        unsigned long long gm = 0ull;
        int sf0 = 0;
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const int v = shfl(s[k].sf, i0 & 31);
            if ((i0 >> 5) == k) sf0 = v;
        }
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            sel[k] = s[k].attempt && s[k].sf == sf0;
            gm |= (unsigned long long)__ballot_sync(kFull, sel[k]) << (32 * k);
        }
        remaining &= ~gm;
        const int cnt = __popcll(gm);
        int winner = -1;
        if (cnt == 1) {
            winner = i0;
        } else {
            collisions += cnt - 1;
            Real v1 = Real(0);
            int i1 = -1;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                if (sel[k] && (i1 < 0 || s[k].cur > v1)) {
                    v1 = s[k].cur;
                    i1 = lane + 32 * k;
                }
            }
            Real top_v, sec_v;
            int top_i, sec_i;
            group_argmax(v1, i1, top_v, top_i);
            Real v2 = Real(0);
            int i2 = -1;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                const int j = lane + 32 * k;
                if (sel[k] && j != top_i && (i2 < 0 || s[k].cur > v2)) {
                    v2 = s[k].cur;
                    i2 = j;
                }
            }
            group_argmax(v2, i2, sec_v, sec_i);
            if (top_v > (sec_v + p.capture_db)) {  // uav_env.py:569
                winner = top_i;
                captures += 1;
            }
        }
        if (winner >= 0) {
            win_mask |= 1ull << winner;
            // the owner lane performs IoTSensor.collect_data (iot_sensors.py:127-145)
            Real got = Real(0), urg_w = Real(0);
            int in_range = 0;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                if (lane + 32 * k == winner) {
                    const Real rc = s[k].det + dr.winner(p, lane, k);
                    in_range = rc >= p.thr;
                    if (in_range && s[k].buf > Real(0)) {
                        const Real cap = data_rate_of<Real>(s[k].sf) * p.cd;
                        got = s[k].buf < cap ? s[k].buf : cap;
                        s[k].buf -= got;
                        s[k].tx += got;
                    }
                    urg_w = urgency_of(p, s[k].buf, s[k].lost, generated_of(p, s[k], h.elapsed));
                }
            }
            const int owner = winner & 31;
            got = shfl(got, owner);
            if constexpr (!kParity) urg_red_acc += (float)got * p.inv_rate;  // (before - after) of the winner
            urg_w = shfl(urg_w, owner);
            in_range = shfl(in_range, owner);
            urg_acc += urg_w;  // np.mean over <= 6 winners adds left to right (uav_env.py:610-615)
            nwin += 1;
            if (in_range && got > Real(0)) {
                total_bytes += got;
                delivered += 1;
                const bool seen = winner < 32 ? (h.vis_lo >> winner) & 1u : (h.vis_hi >> (winner - 32)) & 1u;
                if (!seen) {
                    any_new = true;
                    if (winner < 32)
                        h.vis_lo |= 1u << winner;
                    else
                        h.vis_hi |= 1u << (winner - 32);
                }
            }
        }
    }
    // Phase 5: fairness metrics (uav_env.py:596-604)
    unsigned empty_b = 0u, nonempty_b = 0u;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        empty_b |= __ballot_sync(kFull, s[k].active && s[k].buf <= Real(0));
        nonempty_b |= __ballot_sync(kFull, s[k].active && !(s[k].buf <= Real(0)));
    }
    const bool was_empty = empty_b != 0u, all_collected = nonempty_b == 0u;
    Real urg_red;
    if constexpr (kParity) {
        float diff[SPL];
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const float d = before[k] - (float)div_const(s[k].buf, p.rate, p.inv_rate);
            diff[k] = d > 0.f ? d : 0.f;
        }
        urg_red = (Real)sensor_sum<Real, SPL>(diff, N, lane, ws.red, kSumNumpy);  // float32 np.sum
    } else {
        urg_red = urg_red_acc;
    }
    const Real mean_urg = nwin > 0 ? fdiv(urg_acc, (Real)nwin) : Real(0);
    // Phase 6: reward_function.py:81-128, same order of additions
    Real reward = p.p_step + p.p_hover;
    if (total_bytes > Real(0)) {
        reward += p.r_per_byte * total_bytes * mean_urg;
        if (any_new) reward += p.r_new;
    }
    if (urg_red > Real(0)) reward += p.r_urg_red * urg_red;
    if (was_empty && total_bytes == Real(0)) reward += p.p_revisit;
    reward += p.p_battery * used;
    if (collisions > 0) reward += p.p_collision * (Real)collisions;
    if (step_loss > Real(0)) reward += p.p_data_loss * step_loss;
    if (N > 1) {  // calculate_starvation_penalty, reward_function.py:46-57
        Real mx = Real(0);
#pragma unroll
        for (int k = 0; k < SPL; ++k)
            if (s[k].active) mx = s[k].buf > mx ? s[k].buf : mx;
        mx = warp_max(mx);
        if (mx != Real(0)) {
            if constexpr (kParity) {  // np.var: mean, then mean of squared deviations, numpy summation order
                Real nrm[SPL], dev[SPL];
#pragma unroll
                for (int k = 0; k < SPL; ++k) nrm[k] = s[k].active ? s[k].buf / mx : Real(0);
                const Real mean = sensor_sum<Real, SPL>(nrm, N, lane, ws.red, kSumNumpy) / (Real)N;
#pragma unroll
                for (int k = 0; k < SPL; ++k) {
                    const Real d = nrm[k] - mean;
                    dev[k] = s[k].active ? d * d : Real(0);
                }
                reward += p.p_starvation * (sensor_sum<Real, SPL>(dev, N, lane, ws.red, kSumNumpy) / (Real)N);
            } else {  // one fused shuffle tree: var = E[x^2] - E[x]^2 on x in [0, 1]
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int k = 0; k < SPL; ++k) {
                    const float v = s[k].active ? s[k].buf : 0.f;
                    s1 += v;
                    s2 = fmaf(v, v, s2);
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    s1 += __shfl_xor_sync(kFull, s1, o);
                    s2 += __shfl_xor_sync(kFull, s2, o);
                }
                const float inv = fast_rcp(mx), invn = fast_rcp((float)N);
                const float mean = s1 * inv * invn;
                reward += p.p_starvation * fmaxf(fmaf(s2 * inv * inv, invn, -mean * mean), 0.f);
            }
        }
    }
    if (all_collected) reward += p.r_completion;
    if (lane == 0) {
        ts.creward[k] = reward;
        ts.bytes[k] = total_bytes;
        ts.counts[k] = collisions | (captures << 8) | (delivered << 12);
        ts.nvis_lo[k] = h.vis_lo;
        ts.nvis_hi[k] = h.vis_hi;
        if (feat_step_detail<LEAN>(b)) {  // parity diagnostics: the collect step's intermediates (phase C adds the rest)
            double* d = feat_step_detail<LEAN>(b) + (size_t)e * UAVSIM_DETAIL_COUNT;
            d[UAVSIM_DETAIL_COLLISIONS] = collisions;
            d[UAVSIM_DETAIL_BYTES] = (double)total_bytes;
            d[UAVSIM_DETAIL_URGENCY_REDUCED] = (double)urg_red;
            d[UAVSIM_DETAIL_MEAN_URGENCY] = (double)mean_urg;
            d[UAVSIM_DETAIL_WAS_NEW] = any_new;
            d[UAVSIM_DETAIL_WAS_EMPTY] = was_empty;
            d[UAVSIM_DETAIL_ALL_COLLECTED] = all_collected;
            d[UAVSIM_DETAIL_ATTEMPT_MASK_LO] = (double)(uint32_t)att_mask;
            d[UAVSIM_DETAIL_ATTEMPT_MASK_HI] = (double)(uint32_t)(att_mask >> 32);
            d[UAVSIM_DETAIL_WINNER_MASK_LO] = (double)(uint32_t)win_mask;
            d[UAVSIM_DETAIL_WINNER_MASK_HI] = (double)(uint32_t)(win_mask >> 32);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// software pipeline of the throughput build: while a warp computes environment i, the columns of its
// next environment (i + total_warps) stream from HBM into a shared-memory stage with cp.async
// (LDGSTS: no registers held, no stall until the data is needed), double buffered per warp.
// ------------------------------------------------------------------------------------------------
// A stage holds one environment's sensor columns.  (Tried and removed, commit 0510d7b: TMA bulk copies onto an
// mbarrier -- three 320-800 B `cp.async.bulk` per environment keep the SM's TMA unit ~2/3 busy and were 3-12 %
// slower than per-lane cp.async -- and a shared-memory row image for coalesced observation stores, 7 % slower.)
template <typename Real, int SPL>
struct alignas(alignof(SensorState<Real>) > 16 ? alignof(SensorState<Real>) : 16) StageBuf {
    SensorState<Real> st[32 * SPL];
    float2 pos[32 * SPL];
    Real gen[32 * SPL];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

// Issue the asynchronous copies of environment e (N real sensors) into a stage buffer.
template <typename Real, int SPL>
__device__ __forceinline__ void stage_issue(const Buffers& b, const Params<Real>& p, int e, int N, int lane,
                                            StageBuf<Real, SPL>& sb) {
    static_assert(sizeof(SensorState<Real>) == 16, "the staged pipeline is built for the float32 layout");
    const uint32_t row = (uint32_t)e * (uint32_t)p.P;
    const SensorState<Real>* st = reinterpret_cast<const SensorState<Real>*>(b.sstate);
    const Real* gen = reinterpret_cast<const Real*>(b.sgen);
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const int j = lane + 32 * k;
        if (j < N) {
            const uint32_t i = row + (uint32_t)j;
            cp_async16(&sb.st[j], st + i);
            cp_async8(&sb.pos[j], b.spos + i);
            cp_async4(&sb.gen[j], gen + i);
        }
    }
}

template <typename Real, int SPL>
__device__ __forceinline__ void load_sensors_staged(const StageBuf<Real, SPL>& sb, int N, int lane, Sensor<Real> (&s)[SPL]) {
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const int j = lane + 32 * k;
        s[k].active = j < N;
        s[k].attempt = false;
        s[k].det = Real(0);
        s[k].cur = Real(0);
        // Unconditional loads (the stage has 32 * SPL slots): a lane without a sensor reads whatever an earlier
        // environment left there and runs the same arithmetic on it -- every reduction, ballot and store is
        // masked by `active`, so the values never escape.
        const SensorState<Real> v = sb.st[j];
        const float2 xy = sb.pos[j];
        const Real g = sb.gen[j];
        float ema;
        int sf;
        unpack_sf(v.ema, ema, sf);
        if (!s[k].active) sf = 12;
        s[k].buf = v.buf;
        s[k].tx = v.tx;
        s[k].lost = v.lost;
        s[k].ema = ema;
        s[k].gen = g;
        s[k].sx = xy.x;
        s[k].sy = xy.y;
        s[k].sf = sf;
    }
}

// Episode end inside a step (SB3 VecEnv semantics): terminal observation, episode statistics, new episode.
// Deliberately NOT inlined and self-contained -- it re-reads the sensor columns the step has just stored
// -- so that the rare path costs the hot loop neither registers nor local memory.
template <typename Real, bool REPLAY, int SPL, int FPS>
__device__ __noinline__ void finish_and_restart(const Params<Real>& p, const Buffers& b, int e, int lane, WarpSmem& ws) {
    __syncwarp();
    const int N = ws.rec.num_sensors;
    if (b.terminal_obs) {  // the k stacked frames ending with the terminal one (written by this warp a moment ago)
        const int total = p.frame_window * p.obs_dim;
        float* t = b.terminal_obs + (size_t)e * total;
        const float* win = b.obs + (size_t)e * p.obs_stride + (p.frame_k > 1 ? (b.frame_slot + 1) * p.obs_dim : 0);
        for (int i = lane; i < total; i += 32) t[i] = win[i];
    }
    Sensor<Real> s[SPL];
    load_sensors<Real, SPL>(b, p, e, N, lane, s);
    write_episode_stats<Real, SPL>(p, b, e, lane, ws.rec, s);
    reset_env<Real, REPLAY, SPL, FPS>(p, b, e, lane, ws, s, (Real*)nullptr);
    store_sensors<Real, SPL>(b, p, e, lane, s, true);
}

// ------------------------------------------------------------------------------------------------
// the step kernel: tiles of kTile consecutive environments per warp
// ------------------------------------------------------------------------------------------------
// Per-environment SCALAR work (action decode, UAV move + battery, edge bookkeeping, truncation, reward
// assembly, terminal penalties, shaping, record update, reward/done/mask outputs) is identical for every
// lane when a warp owns one environment -- 32 lanes would issue it to produce one result.  So a warp
// takes a TILE of kTile environments and runs that work one-environment-per-LANE (phases A and C),
// while the per-sensor work stays one-environment-per-WARP with sensors across lanes (phase B, a loop
// over the tile).  The phases talk through a small structure-of-arrays block in shared memory; the tile's
// 128-byte records sit in shared memory with a 4-byte XOR swizzle so that both the coalesced global
// copies (lane = word) and the per-lane field accesses (lane = environment) are bank-conflict free.
#ifndef UAVSIM_TILE
#define UAVSIM_TILE 16
#endif
constexpr int kTile = UAVSIM_TILE;

template <typename Real>
struct alignas(16) TileSmem {
    uint32_t rec[2][kTile][32];  // [buffer][l][w ^ l]: word w of record l; the other buffer receives the next tile
    int32_t action[2][kTile];
    // phase A -> phase B (read as warp-wide broadcasts) and -> phase C
    float x[kTile], y[kTile], elapsed[kTile];
    int32_t N[kTile], step[kTile], flags[kTile];  // (grid W, H are read from the record tile where needed)
    uint32_t episode[kTile], vis_lo[kTile], vis_hi[kTile];
    double battery[kTile];
    Real used[kTile];
    // phase B -> phase C
    Real step_loss[kTile], creward[kTile], bytes[kTile], jain[kTile];
    float cur_dist[kTile];
    int32_t counts[kTile];  // collisions | captures << 8 | delivered << 12 | starved << 16
    uint32_t nvis_lo[kTile], nvis_hi[kTile];
};

// Dynamic shared memory of one block of the step kernel: per warp a working record + reduction scratch,
// the tile block and (float32 layout) two stage buffers.
template <typename Real, int SPL, bool PAIR = false>
struct alignas(alignof(SensorState<Real>) > 16 ? alignof(SensorState<Real>) : 16) StepWarpSmem {
    WarpSmem ws;
    TileSmem<Real> ts;
    StageBuf<Real, SPL> stage[std::is_same<Real, double>::value ? 0 : (PAIR ? 4 : 2)];  // pair kernel: 2 stages x 2 envs
};
template <typename Real, int SPL, bool PAIR = false>
constexpr size_t step_smem_bytes() {
    return sizeof(StepWarpSmem<Real, SPL, PAIR>) * step_warps<Real, SPL, PAIR>();
}

enum : int {
    kFlagValid = 1, kFlagCollect = 2, kFlagTruncated = 4, kFlagMoveOk = 8, kFlagEdge = 16, kFlagJain = 32
};

// record word offsets (uavsim_env_record, include/uavsim.h)
enum : int {
    kRX = 0, kRY = 1, kRStartX = 2, kRStartY = 3, kRBattery = 4, kRTotalReward = 6, kRTotalData = 8, kREpReturn = 10,
    kRLastBytes = 12, kRPrevDist = 13, kRElapsed = 14, kRStep = 15, kRGridW = 16, kRGridH = 17, kRN = 18, kRVisLo = 19,
    kRVisHi = 20, kRCapture = 21, kRBoundary = 22, kREdge = 23, kRCollisions = 24, kRFirstCov = 25, kREpisode = 26,
    kRDelivered = 27, kRInitSf = 28, kRDwell = 29, kRWindAngle = 30
};
static_assert(offsetof(uavsim_env_record, battery) == 4 * kRBattery && offsetof(uavsim_env_record, step) == 4 * kRStep &&
                  offsetof(uavsim_env_record, num_sensors) == 4 * kRN && offsetof(uavsim_env_record, episode) == 4 * kREpisode &&
                  offsetof(uavsim_env_record, delivered) == 4 * kRDelivered && offsetof(uavsim_env_record, dwell_steps) == 4 * kRDwell && offsetof(uavsim_env_record, wind_angle) == 4 * kRWindAngle && offsetof(uavsim_env_record, elapsed_s) == 4 * kRElapsed &&
                  offsetof(uavsim_env_record, total_reward) == 4 * kRTotalReward && offsetof(uavsim_env_record, episode_return) == 4 * kREpReturn &&
                  offsetof(uavsim_env_record, prev_dist_nearest) == 4 * kRPrevDist && offsetof(uavsim_env_record, first_full_coverage_step) == 4 * kRFirstCov,
              "record word map out of date");

struct RecView {  // lane-private access to record l of the swizzled tile
    uint32_t* row;
    int l;
    __device__ __forceinline__ uint32_t& w(int i) const { return row[i ^ l]; }
    __device__ __forceinline__ float f(int i) const { return __uint_as_float(row[i ^ l]); }
    __device__ __forceinline__ int i32(int i) const { return (int)row[i ^ l]; }
    __device__ __forceinline__ double d(int i) const { return __hiloint2double((int)row[(i + 1) ^ l], (int)row[i ^ l]); }
    __device__ __forceinline__ void set_f(int i, float v) const { row[i ^ l] = __float_as_uint(v); }
    __device__ __forceinline__ void set_i(int i, int v) const { row[i ^ l] = (uint32_t)v; }
    __device__ __forceinline__ void set_d(int i, double v) const {
        row[i ^ l] = (uint32_t)__double2loint(v);
        row[(i + 1) ^ l] = (uint32_t)__double2hiint(v);
    }
};

template <typename Real, bool REPLAY, int SPL, int LEAN>
__device__ __forceinline__ Draws<Real, REPLAY, SPL> step_draws(const Params<Real>& p, const Buffers& b,
                                                                const TileSmem<Real>& ts, int k, int e) {
    const uint64_t gid = (uint64_t)(p.env_id_offset + e);
    Draws<Real, REPLAY, SPL> dr;
    dr.env_lo = (uint32_t)gid;
    dr.env_hi = (uint32_t)(gid >> 32) & 0xffffffu;
    dr.episode = ts.episode[k] - 1;  // the record counts resets done; the running episode is one less
    dr.step = (uint32_t)ts.step[k];
    // pair time = the even environment of the pair (tiles start at multiples of 16 and env_id_offset is even when
    // P <= 32, so it is slot k & ~1 of this tile; phase A publishes episode / step of every slot, valid or not)
    dr.paired = SPL == 1;
    dr.pair_episode = ts.episode[k & ~1] - 1;
    dr.pair_step = (uint32_t)ts.step[k & ~1];
    dr.tab = REPLAY ? b.step_draws + (size_t)e * UAVSIM_NUM_SLOTS * p.P : nullptr;
    dr.exp = (!REPLAY && feat_draw_export<LEAN>(b))
                 ? reinterpret_cast<Real*>(feat_draw_export<LEAN>(b)) + (size_t)e * UAVSIM_NUM_SLOTS * p.P
                 : nullptr;
    return dr;
}

// Asynchronous copy of the records and actions of the tile starting at `base` into buffer `tb` of the tile block
// (float32 layout: issued during phase B of the previous tile so that a tile never waits for its records).
template <typename Real>
__device__ __forceinline__ void tile_issue(const Buffers& b, const int32_t* actions, int base, int tb, int lane,
                                           TileSmem<Real>& ts) {
    const int n = min(b.tile, b.E - base);
#pragma unroll
    for (int r = 0; r < kTile; ++r)
        if (r < n) cp_async4(&ts.rec[tb][r][lane ^ r], reinterpret_cast<const uint32_t*>(b.rec + base + r) + lane);
    if (lane < n) cp_async4(&ts.action[tb][lane], actions + base + lane);
}

// IoTSensor.step (iot_sensors.py:114-125) for one sensor: the buffer grows by `nw`, is capped at max_buffer_size, and the
// excess is lost.  Returns the excess; `over` = the cap was hit.
template <typename Real>
__device__ __forceinline__ Real age_buffer(const Params<Real>& p, Sensor<Real>& s, Real nw, bool& over) {
    const Real pot = s.buf + nw;
    Real loss;
    if constexpr (std::is_same<Real, double>::value) {
        over = pot > p.maxbuf;
        loss = over ? pot - p.maxbuf : Real(0);
        s.buf = over ? p.maxbuf : pot;
    } else {  // min / max instead of compare + selects (pot - max is exact where it matters: pot ~ max)
        loss = fmaxf(pot - p.maxbuf, 0.f);
        s.buf = fminf(pot, p.maxbuf);
        over = loss > 0.f;
    }
    s.lost += loss;  // + 0 when not saturated: no effect on the stored value
    return loss;
}

// Phase B for ONE environment (slot k of the tile): everything that touches the sensor columns.  Scalars of
// the environment are read from the tile block where they are used and results are written back as soon as
// they exist, so that little besides the sensor registers is live across the phase.
template <typename Real, bool REPLAY, int SPL, int FPS, bool STACK, int LEAN>
__device__ __forceinline__ void sensor_phase(const Params<Real>& p, const Buffers& b, TileSmem<Real>& ts, WarpSmem& ws,
                                             int tb, int k, int e, int lane, Sensor<Real> (&s)[SPL]) {
    constexpr bool kParity = std::is_same<Real, double>::value;
    const int flags = ts.flags[k];
    const bool collect = flags & kFlagCollect;
    const int N = ts.N[k];

    if (!REPLAY && feat_draw_export<LEAN>(b)) {
        Real* exp = reinterpret_cast<Real*>(feat_draw_export<LEAN>(b)) + (size_t)e * UAVSIM_NUM_SLOTS * p.P;
        const Real nan = Real(CUDART_NAN);
        for (int i = lane; i < UAVSIM_NUM_SLOTS * p.P; i += 32) exp[i] = nan;
        __syncwarp();
    }

    // ---- iot_sensors.py:114-125: age the buffers ----
    Real step_loss = Real(0);
    {
        Real loss[SPL];
        const Real nw = p.rate * (collect ? p.cd : Real(1));
        bool any_over = false;
#pragma unroll
        for (int k2 = 0; k2 < SPL; ++k2) {
            if constexpr (kParity) s[k2].gen += nw;
            bool over;
            loss[k2] = age_buffer(p, s[k2], nw, over);
            any_over |= (s[k2].active && over);
        }
        if (__any_sync(kFull, any_over)) step_loss = sensor_sum<Real, SPL>(loss, N, lane, ws.red, kSumSequential);
        if (lane == 0) ts.step_loss[k] = step_loss;
    }

    compute_det(p, ts.x[k], ts.y[k], s);  // the UAV has already moved (phase A)
    if (collect) collect_action<Real, REPLAY, SPL, LEAN>(p, b, ts, ws, k, e, s, step_loss, lane);

    if (flags & kFlagTruncated) {  // reward_function.py:59-67 (counted only when the episode ends this step)
        int starved = 0;
        const float elapsed = ts.elapsed[k];
#pragma unroll
        for (int k2 = 0; k2 < SPL; ++k2) {
            const Real g = generated_of(p, s[k2], elapsed);
            starved += __popc(__ballot_sync(kFull, s[k2].active && g > Real(0) && fdiv(s[k2].tx, g) < p.cr_thr));
        }
        if (lane == 0) ts.counts[k] |= starved << 16;
    }

    // ---- observation, sensor part (uav_env.py:638-674); obs[0..2] are written by phase C ----
    {
        Hot h;
        h.x = ts.x[k];
        h.y = ts.y[k];
        h.elapsed = ts.elapsed[k];
        h.W = (int)ts.rec[tb][k][kRGridW ^ k];
        h.H = (int)ts.rec[tb][k][kRGridH ^ k];
        h.N = N;
        h.battery = 0.0;
        h.vis_lo = collect ? ts.nvis_lo[k] : ts.vis_lo[k];  // relational rows carry the visited flag
        h.vis_hi = collect ? ts.nvis_hi[k] : ts.vis_hi[k];
        observe<Real, REPLAY, SPL, FPS, false, STACK>(p, h, s, step_draws<Real, REPLAY, SPL, LEAN>(p, b, ts, k, e), lane,
                                                       obs_frame(p, b, e));
    }

    // ---- RelationalUAVEnv.step supplements (env_wrapper.py:169-198): potential distance, dwell probes ----
    if (FPS == 5 && feat_relational<LEAN>(p)) {
        const uint32_t vlo = collect ? ts.nvis_lo[k] : ts.vis_lo[k], vhi = collect ? ts.nvis_hi[k] : ts.vis_hi[k];
        const float dist = nearest_unvisited(ts.x[k], ts.y[k], vlo, vhi, lane, s);
        int n_in = 0;
        if (collect) {
            Real f[SPL];
            step_draws<Real, REPLAY, SPL, LEAN>(p, b, ts, k, e).dwell(p, b.dwell_draws ? b.dwell_draws + (size_t)e * p.P : nullptr, lane, f);
#pragma unroll
            for (int k2 = 0; k2 < SPL; ++k2)
                n_in += __popc(__ballot_sync(kFull, s[k2].active && (s[k2].det + f[k2]) >= p.thr));
        }
        if (lane == 0) {
            ts.cur_dist[k] = dist;
            ts.counts[k] |= n_in << 24;
        }
    }

    // ---- DomainRandEnv.step shaping inputs (dqn.py:415-444) ----
    if (feat_domain_rand<LEAN>(p)) {
        const float cur_dist = nearest_with_data(ts.x[k], ts.y[k], s);
        const float elapsed = ts.elapsed[k];
        Real r1[SPL], r2[SPL];
        int cnt = 0;
#pragma unroll
        for (int k2 = 0; k2 < SPL; ++k2) {
            const Real g = generated_of(p, s[k2], elapsed);
            const bool has = s[k2].active && g > Real(0);
            const Real rate = has ? fdiv(s[k2].tx, g) * Real(100) : Real(0);
            r1[k2] = rate;
            r2[k2] = rate * rate;
            cnt += __popc(__ballot_sync(kFull, has));
        }
        Real jain = Real(0);
        if (cnt > 0) {
            const Real s1 = sensor_sum<Real, SPL>(r1, N, lane, ws.red, kSumSequential);
            const Real s2 = sensor_sum<Real, SPL>(r2, N, lane, ws.red, kSumSequential);
            jain = s2 > Real(0) ? fdiv(s1 * s1, (Real)cnt * s2) : Real(1);
        }
        if (lane == 0) {
            ts.cur_dist[k] = cur_dist;
            ts.jain[k] = jain;
            if (cnt > 0) ts.flags[k] = flags | kFlagJain;
        }
    }
    store_sensors<Real, SPL>(b, p, e, lane, s, false);
}

// ------------------------------------------------------------------------------------------------
// P <= 32, plain environment (LEAN): TWO environments per warp pass.  With one sensor slot per lane a pass over
// one environment leaves 12 of 32 lanes idle at N = 20 and uses half of each Philox call; here slot q of a lane
// is sensor `lane` of environment k0 + q, so the per-sensor arithmetic of the two environments interleaves like
// the two slots of the P > 32 kernel, the observation draws of both come from ONE Philox call (the pair-keyed
// counter of Draws), and the per-pass loop control is paid once per pair.  Everything per environment -- collect
// action, reductions, stores -- runs the one-slot building blocks on a one-element view of s[q].
// ------------------------------------------------------------------------------------------------
template <bool STACK>
__device__ __forceinline__ void pair_phase(const Params<float>& p, const Buffers& b, TileSmem<float>& ts, WarpSmem& ws,
                                           int k0, int base, int lane, const StageBuf<float, 1>* st2, bool have_b) {
    using View = Sensor<float>(&)[1];
    Sensor<float> s[2];
    int flags[2], N[2];
    bool valid[2], collect[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int k = k0 + q;
        flags[q] = (q == 0 || have_b) ? ts.flags[k] : 0;
        valid[q] = flags[q] & kFlagValid;
        collect[q] = flags[q] & kFlagCollect;
        N[q] = valid[q] ? ts.N[k] : 0;  // an absent / rejected environment has no active lane: nothing is reduced or stored
        load_sensors_staged<float, 1>(st2[q], N[q], lane, reinterpret_cast<View>(s[q]));
    }
    __syncwarp();
    if (!(valid[0] || valid[1])) return;

    // ---- iot_sensors.py:114-125: age the buffers ----
    float step_loss[2];
    {
        float loss[2];
        bool any_over[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float nw = p.rate * (collect[q] ? p.cd : 1.0f);
            bool over;
            loss[q] = age_buffer(p, s[q], nw, over);
            any_over[q] = s[q].active && over;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            step_loss[q] = 0.f;
            if (__any_sync(kFull, any_over[q])) step_loss[q] = warp_sum(s[q].active ? loss[q] : 0.f);
            if (lane == 0 && valid[q]) ts.step_loss[k0 + q] = step_loss[q];
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) s[q].det = det_rssi(p, ts.x[k0 + q], ts.y[k0 + q], s[q].sx, s[q].sy);
#pragma unroll
    for (int q = 0; q < 2; ++q)
        if (valid[q] && collect[q])
            collect_action<float, false, 1, kLeanPlain>(p, b, ts, ws, k0 + q, base + k0 + q, reinterpret_cast<View>(s[q]),
                                                  step_loss[q], lane);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (flags[q] & kFlagTruncated) {  // reward_function.py:59-67
            const float g = generated_of(p, s[q], ts.elapsed[k0 + q]);
            const int starved = __popc(__ballot_sync(kFull, s[q].active && g > 0.f && fdiv(s[q].tx, g) < p.cr_thr));
            if (lane == 0) ts.counts[k0 + q] |= starved << 16;
        }
    }
    // ---- observation, sensor part: one Philox call for the pair (uav_env.py:638-674) ----
    {
        const Draws<float, false, 1> dr = step_draws<float, false, 1, kLeanPlain>(p, b, ts, k0, base + k0);
        const Philox4 qp = dr.pair_call(p, lane);
        float dd[2][1], de[2][1];
        box_muller<float>(qp.x, qp.y, p.sigma, dd[0][0], de[0][0]);
        box_muller<float>(qp.z, qp.w, p.sigma, dd[1][0], de[1][0]);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            Hot h;
            h.x = h.y = 0.f;
            h.elapsed = ts.elapsed[k0 + q];
            h.W = h.H = 1;
            h.N = N[q];
            h.battery = 0.0;
            h.vis_lo = h.vis_hi = 0u;
            observe_with<float, 1, 3, false, STACK>(p, h, reinterpret_cast<View>(s[q]), dd[q], de[q], lane,
                                                    obs_frame(p, b, base + k0 + q));
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) store_sensors<float, 1>(b, p, base + k0 + q, lane, reinterpret_cast<View>(s[q]), false);
}

// One launch steps every environment of the handle once.  Persistent warps: the grid covers the SMs at
// full residency and each warp walks tiles t = warp_id, warp_id + total_warps, ...
template <typename Real, bool REPLAY, int SPL, int FPS, bool STACK, int LEAN = 0, bool PAIR = false>
__global__ void __launch_bounds__(32 * step_warps<Real, SPL, PAIR>(), step_min_blocks<Real, SPL, PAIR>())
    step_kernel(const __grid_constant__ Params<Real> p, const __grid_constant__ Buffers b,
                const int32_t* __restrict__ actions) {
    constexpr bool kParity = std::is_same<Real, double>::value;
    constexpr bool kStaged = !kParity;  // the float32 layout runs the cp.async software pipeline
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    static_assert(!PAIR || (LEAN == kLeanPlain && SPL == 1 && FPS == 3 && !REPLAY && !std::is_same<Real, double>::value),
                  "the pair kernel is the plain throughput step with one sensor slot per lane");
    StepWarpSmem<Real, SPL, PAIR>& mine = reinterpret_cast<StepWarpSmem<Real, SPL, PAIR>*>(dyn_smem)[warp];
    WarpSmem& ws = mine.ws;
    TileSmem<Real>& ts = mine.ts;
    StageBuf<Real, SPL>* stage = mine.stage;
    // loop-invariant scalars are re-derived from the constant bank where they are used instead of being
    // carried in registers: with 184 KB of the SM given to shared memory the L1 is small and a spilled
    // loop-carried value costs an L2 round trip per environment
#define UAVSIM_E (b.E)
    // Tiles are handed out by TICKET (one atomicAdd per tile per warp, requested a whole tile ahead of its use):
    // warps that drew cheap tiles (few collect actions, few restarts) take more of them, so the launch does not end
    // on the slowest statically assigned warp, and small batches spread over the whole machine with a smaller tile.
    // Two counters alternate between launches: this one counts, the other is zeroed for the next launch.
    unsigned* const ticket_ctr = b.tile_ticket + (b.step_serial & 1u);
    if (blockIdx.x == 0 && threadIdx.x == 0) b.tile_ticket[(b.step_serial & 1u) ^ 1u] = 0u;
#define UAVSIM_TAKE_TICKET() ((lane == 0) ? (int)atomicAdd(ticket_ctr, 1u) : 0)
    int base, next_base;
    {
        const int t0 = UAVSIM_TAKE_TICKET();
        const int t1 = UAVSIM_TAKE_TICKET();
        // redux.sync (REDUX) delivers its result in a UNIFORM register: the compiler then knows that base, the
        // environment index, the row offsets and the loop counters derived from them are warp-uniform, keeps them
        // off the vector register file and turns the loop-control branches into uniform branches
        base = (int)__reduce_max_sync(kFull, (unsigned)t0) * b.tile;
        next_base = (int)__reduce_max_sync(kFull, (unsigned)t1) * b.tile;
    }
#define UAVSIM_NEXT_BASE next_base
    int tb = 0;  // tile buffer in use
    if constexpr (kStaged) {  // prime both pipelines: records of the first tile, sensor columns of its first env
        if (base < UAVSIM_E) tile_issue<Real>(b, actions, base, 0, lane, ts);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        if constexpr (!PAIR) {  // the first two environments of the warp's sequence, one stage and one group each
            if (base < UAVSIM_E) stage_issue<Real, SPL>(b, p, base, (int)ts.rec[0][0][kRN], lane, stage[0]);
            cp_async_commit();
            if (base + 1 < UAVSIM_E) stage_issue<Real, SPL>(b, p, base + 1, (int)ts.rec[0][1][kRN ^ 1], lane, stage[1]);
            cp_async_commit();
        }
        if constexpr (PAIR) {  // the first pair of the warp's sequence (sensor counts from the records just landed)
            if (base < UAVSIM_E) stage_issue<Real, SPL>(b, p, base, (int)ts.rec[0][0][kRN], lane, stage[0]);
            if (base + 1 < UAVSIM_E) stage_issue<Real, SPL>(b, p, base + 1, (int)ts.rec[0][1][kRN ^ 1], lane, stage[1]);
            cp_async_commit();
        }
    }

    while (base < UAVSIM_E) {
        const int pending_ticket = UAVSIM_TAKE_TICKET();  // the tile after the next one; consumed at the end of this tile
        const int ntile = min(b.tile, UAVSIM_E - base);
        uint32_t(*rec)[32] = ts.rec[tb];
        if constexpr (!kStaged) {
            // ---- tile records: global -> swizzled shared memory (lane = word: coalesced 128-byte rows) ----
            __syncwarp();
#pragma unroll
            for (int r = 0; r < kTile; ++r)
                if (r < ntile) rec[r][lane ^ r] = reinterpret_cast<const uint32_t*>(b.rec + base + r)[lane];
            if (lane < ntile) ts.action[tb][lane] = actions[base + lane];
            __syncwarp();
        }

        // ---- phase A: one environment per lane ----
        const RecView R{rec[lane & (kTile - 1)], lane & (kTile - 1)};
        bool my_valid = false;
        int my_n = 0;
        if (lane < ntile) {
            const int e = base + lane;
            int my_flags = 0;
            const int my_action = ts.action[tb][lane];
            ts.step[lane] = R.i32(kRStep) + 1;     // published for every slot: the even slot is its pair's clock
            ts.episode[lane] = R.w(kREpisode);
            if (my_action < 0 || my_action > 4) {  // uav_env.py:468 raises ValueError; here: flag it, leave the env untouched
                atomicMin(b.error_word, ((unsigned long long)b.step_serial << 32) | (unsigned)e);
            } else {
                my_flags = kFlagValid;
                float x = R.f(kRX), y = R.f(kRY);
                const int W = R.i32(kRGridW), H = R.i32(kRGridH);
                double battery = R.d(kRBattery);
                const int step = R.i32(kRStep) + 1;  // uav_env.py:439
                // uav_env.py:443-447: edge-cell bookkeeping on the PRE-move position
                if constexpr (kParity) {
                    const double Wd = W, Hd = H, ux = x, uy = y, eps = 1e-6;
                    if (ux <= eps || uy <= eps || ux >= Wd - 1 - eps || uy >= Hd - 1 - eps) my_flags |= kFlagEdge;
                } else {
                    const float eps = 1e-6f;
                    if (x <= eps || y <= eps || x >= (float)(W - 1) - eps || y >= (float)(H - 1) - eps) my_flags |= kFlagEdge;
                }
                const bool collect = my_action == 4;
                Real used;
                if (!collect) {  // uav.py:127-185
                    float nx = x, ny = y;
                    if (my_action == 0)
                        ny += 1.0f;
                    else if (my_action == 1)
                        ny -= 1.0f;
                    else if (my_action == 2)
                        nx -= 1.0f;
                    else
                        nx += 1.0f;
                    const double before = battery;
                    const bool ok = (0.f <= nx) && (nx < (float)W) && (0.f <= ny) && (ny < (float)H);
                    if (ok) {
                        x = nx;
                        y = ny;
                        battery -= p.e_move;
                        my_flags |= kFlagMoveOk;
                    } else {
                        battery -= p.e_fail;
                    }
                    used = (Real)(before - battery);
                } else {  // uav.py:187-206, :260-263
                    my_flags |= kFlagCollect;
                    battery -= p.e_hover;
                    used = (Real)p.used_hover;
                }
                if (!(battery > p.batt_floor) || step >= p.max_steps) my_flags |= kFlagTruncated;  // uav.py:224, uav_env.py:474-478
                ts.x[lane] = x;
                ts.y[lane] = y;
                ts.elapsed[lane] = R.f(kRElapsed) + (float)(collect ? p.cd : Real(1));
                ts.N[lane] = R.i32(kRN);
                ts.vis_lo[lane] = R.w(kRVisLo);
                ts.vis_hi[lane] = R.w(kRVisHi);
                ts.battery[lane] = battery;
                ts.used[lane] = used;
            }
            ts.flags[lane] = my_flags;
            ts.counts[lane] = 0;
            my_valid = my_flags & kFlagValid;
            my_n = R.i32(kRN);
        }
        // one bit per environment of the tile / its sensor count in the owning lane: phase B takes them from
        // registers (a shift, a shuffle) instead of re-reading shared memory per environment
        const unsigned valid_mask = __ballot_sync(kFull, my_valid);
        __syncwarp();

        // ---- phase B (pair kernel): two environments per warp pass ----
        if constexpr (PAIR) {
            // Stages: pair buffer pb holds the current pair (2 environments), pb ^ 1 receives the next pair.  cp.async
            // groups in commit order: G(j) per pair iteration j, the next tile's records + actions T right after G(0);
            // T is read from the last pair iteration on.
            for (int k = 0; k < ntile; k += 2) {
                const int pb = (k >> 1) & 1;
                int e1 = base + k + 2, n1a = 0, n1b = 0;
                bool ha = false, hb = false;
                if (k + 2 < ntile) {
                    n1a = (int)rec[k + 2][kRN ^ (k + 2)];
                    ha = true;
                    if (k + 3 < ntile) {
                        n1b = (int)rec[k + 3][kRN ^ (k + 3)];
                        hb = true;
                    }
                } else {
                    e1 = UAVSIM_NEXT_BASE;
                    cp_async_wait<0>();  // the next tile's records (T) must have landed before they are read; with
                    __syncwarp();        // the 16-tile they did long ago, with a short tile this is the only guarantee
                    if (e1 < UAVSIM_E) {
                        n1a = (int)ts.rec[tb ^ 1][0][kRN];
                        ha = true;
                        if (e1 + 1 < UAVSIM_E) {
                            n1b = (int)ts.rec[tb ^ 1][1][kRN ^ 1];
                            hb = true;
                        }
                    }
                }
                if (ha) stage_issue<Real, SPL>(b, p, e1, n1a, lane, stage[(pb ^ 1) * 2]);
                if (hb) stage_issue<Real, SPL>(b, p, e1 + 1, n1b, lane, stage[(pb ^ 1) * 2 + 1]);
                cp_async_commit();
                if (k == 0) {
                    if (UAVSIM_NEXT_BASE < UAVSIM_E)
                        tile_issue<Real>(b, actions, UAVSIM_NEXT_BASE, tb ^ 1, lane, ts);
                    cp_async_commit();
                }
                if (k <= 2)
                    cp_async_wait<2>();  // G(j) and T may still be in flight
                else
                    cp_async_wait<1>();  // only G(j): T has landed before the last iteration peeks at it
                __syncwarp();
                if constexpr (PAIR) pair_phase<STACK>(p, b, ts, ws, k, base, lane, &stage[pb * 2], k + 1 < ntile);
            }
        }
        // ---- phase B: one environment per warp, sensors across lanes ----
        for (int k = 0; !PAIR && k < ntile; ++k) {
            const int e = base + k;
            Sensor<Real> s[SPL];
            if constexpr (kStaged) {
                // Two stages, prefetch distance TWO environments: wait for this environment's stage (requested two
                // iterations ago), move it to registers, and reuse the SAME stage at once for the environment two
                // places down the warp's sequence, so a request has two full iterations to land.  Stages alternate
                // with k: tiles hold an even number of environments (a partial tile can only be the last one).
                // cp.async groups in commit order: G(k) = environment k + 2, plus the NEXT tile's records + actions
                // T right after G(0); this environment's group is G(k - 2).
                if (k == 1 || k == 2)
                    cp_async_wait<2>();  // behind G(k-2): [G(0),] T, G(k-1)
                else
                    cp_async_wait<1>();  // behind G(k-2): G(k-1) only -- from k == 3 on T has landed too
                __syncwarp();
                const bool valid = (valid_mask >> k) & 1u;
                if (valid)
                    load_sensors_staged<Real, SPL>(stage[k & 1], ts.N[k], lane, s);
                __syncwarp();
                {
                    int e2 = e + 2, n2 = 0;
                    bool have = false;
                    if (k + 2 < ntile) {
                        n2 = shfl(my_n, k + 2);
                        have = true;
                    } else {
                        const int r = k + 2 - ntile;  // 0 or 1: the next tile's first two environments
                        e2 = UAVSIM_NEXT_BASE + r;
                        if (ntile < 8) {  // a short tile has not passed the waits that retire T
                            cp_async_wait<0>();
                            __syncwarp();
                        }
                        if (e2 < UAVSIM_E) {
                            n2 = (int)ts.rec[tb ^ 1][r][kRN ^ r];
                            have = true;
                        }
                    }
                    if (have) stage_issue<Real, SPL>(b, p, e2, n2, lane, stage[k & 1]);
                    cp_async_commit();
                }
                if (k == 0) {  // T: the other record buffer is free since the previous tile's write-back
                    if (UAVSIM_NEXT_BASE < UAVSIM_E)
                        tile_issue<Real>(b, actions, UAVSIM_NEXT_BASE, tb ^ 1, lane, ts);
                    cp_async_commit();
                }
                if (!valid) continue;
            } else {
                if (!(ts.flags[k] & kFlagValid)) continue;
                load_sensors<Real, SPL>(b, p, e, ts.N[k], lane, s);
            }
            sensor_phase<Real, REPLAY, SPL, FPS, STACK, LEAN>(p, b, ts, ws, tb, k, e, lane, s);
        }
        __syncwarp();

        // ---- phase C: one environment per lane ----
        unsigned restart_mask = 0u;
        {
            bool restart = false;
            if (lane < ntile && (ts.flags[lane] & kFlagValid)) {
                const int e = base + lane;
                const int flags = ts.flags[lane];
                const bool collect = flags & kFlagCollect, truncated = flags & kFlagTruncated;
                const int N = ts.N[lane];
                const Real step_loss = ts.step_loss[lane];
                const int counts = ts.counts[lane];
                Real reward, total_bytes = Real(0);
                uint32_t vis_lo = ts.vis_lo[lane], vis_hi = ts.vis_hi[lane];
                if (!collect) {  // reward_function.py:69-79 + uav_env.py:508-514
                    reward = p.p_step;
                    reward += (flags & kFlagMoveOk) ? p.r_move : p.p_boundary;
                    reward += p.p_battery * ts.used[lane];
                    reward += p.p_data_loss * step_loss;
                } else {
                    reward = ts.creward[lane];
                    total_bytes = ts.bytes[lane];
                    vis_lo = ts.nvis_lo[lane];
                    vis_hi = ts.nvis_hi[lane];
                }
                if (truncated) {  // uav_env.py:481-485, reward_function.py:59-67
                    const int unvisited = N - (__popc(vis_lo) + __popc(vis_hi));
                    if (unvisited > 0) reward += p.p_unvisited * (Real)unvisited;
                    const int starved = (counts >> 16) & 0xff;
                    Real pen = Real(0);
                    for (int i = 0; i < starved; ++i) pen += p.p_starved;  // the reference adds -1000 once per sensor
                    reward += pen;
                }
                // DomainRandEnv.step shaping (dqn.py:415-444)
                Real shaped = reward;
                if (feat_domain_rand<LEAN>(p)) {
                    const float prev = R.f(kRPrevDist), curd = ts.cur_dist[lane];
                    if (prev > 0.f) {
                        if constexpr (kParity)
                            shaped += p.prox_eta * ((double)prev - (double)curd);
                        else
                            shaped += p.prox_eta * (prev - curd);
                    }
                    R.set_f(kRPrevDist, curd);
                    if (R.i32(kRFirstCov) < 0 && (__popc(vis_lo) + __popc(vis_hi)) == N) R.set_i(kRFirstCov, ts.step[lane]);
                    if (flags & kFlagJain) shaped += fdiv(p.jain_bonus * (ts.jain[lane] - Real(0.5)), (Real)N);
                }
                if (FPS == 5 && feat_relational<LEAN>(p)) {  // env_wrapper.py:169-198, :291-313
                    const double W = R.i32(kRGridW), H = R.i32(kRGridH);
                    const double diag = sqrt(W * W + H * H);
                    const float prev = R.f(kRPrevDist), curd = ts.cur_dist[lane];
                    R.set_f(kRPrevDist, curd);
                    int dwell = R.i32(kRDwell);
                    const int n_in = (ts.counts[lane] >> 24) & 0x7f;
                    Real bonus = Real(0);
                    if (!collect) {
                        dwell = 0;
                    } else if (n_in >= p.rel_min_cluster && dwell < p.rel_dwell_steps) {
                        if constexpr (kParity)
                            bonus = p.rel_dwell_bonus * (1.0 - (double)dwell / p.rel_dwell_steps);
                        else
                            bonus = p.rel_dwell_bonus * (1.0f - (float)dwell / (float)p.rel_dwell_steps);
                        dwell += 1;
                    } else if (n_in < p.rel_min_cluster) {
                        dwell = 0;
                    }
                    R.set_i(kRDwell, dwell);
                    if constexpr (kParity) {
                        const double shaping = p.rel_gamma * (-(double)curd / diag) - (-(double)prev / diag);
                        shaped = (reward + shaping + bonus) * p.rel_scale;
                    } else {
                        const float inv = (float)(1.0 / diag);
                        const float shaping = p.rel_gamma * (-curd * inv) + prev * inv;
                        shaped = (reward + shaping + bonus) * p.rel_scale;
                    }
                }
                // commit the record (uav_env.py:487 and the counters of :300-307)
                float x = ts.x[lane], y = ts.y[lane];
                if (feat_wind<LEAN>(p)) {  // WindDriftEnv.step, sim_to_real_analysis.py:214-219
                    const double a = R.d(kRWindAngle);
                    if constexpr (kParity) {  // np.float32 + np.float64 drift -> float64 sum, clip, one rounding to float32
                        x = (float)fmin(fmax((double)x + p.wind_drift * cos(a), 0.0), (double)(R.i32(kRGridW) - 1));
                        y = (float)fmin(fmax((double)y + p.wind_drift * sin(a), 0.0), (double)(R.i32(kRGridH) - 1));
                    } else {
                        float sn, cs;
                        __sincosf((float)a, &sn, &cs);
                        x = fminf(fmaxf(fmaf((float)p.wind_drift, cs, x), 0.f), (float)(R.i32(kRGridW) - 1));
                        y = fminf(fmaxf(fmaf((float)p.wind_drift, sn, y), 0.f), (float)(R.i32(kRGridH) - 1));
                    }
                }
                const double battery = ts.battery[lane];
                R.set_f(kRX, x);
                R.set_f(kRY, y);
                R.set_d(kRBattery, battery);
                R.set_f(kRElapsed, ts.elapsed[lane]);
                R.set_i(kRStep, ts.step[lane]);
                R.set_d(kRTotalReward, R.d(kRTotalReward) + (double)reward);
                R.set_d(kREpReturn, R.d(kREpReturn) + (double)shaped);
                R.set_f(kRLastBytes, (float)total_bytes);
                if (flags & kFlagEdge) R.set_i(kREdge, R.i32(kREdge) + 1);
                if (!collect) {
                    if (!(flags & kFlagMoveOk)) R.set_i(kRBoundary, R.i32(kRBoundary) + 1);
                } else {
                    R.w(kRVisLo) = vis_lo;
                    R.w(kRVisHi) = vis_hi;
                    R.set_d(kRTotalData, R.d(kRTotalData) + (double)total_bytes);
                    R.set_i(kRCapture, R.i32(kRCapture) + ((counts >> 8) & 0xf));
                    R.set_i(kRCollisions, R.i32(kRCollisions) + (counts & 0xff));
                    R.set_i(kRDelivered, R.i32(kRDelivered) + ((counts >> 12) & 0xf));
                }
                // outputs
                restart = truncated && p.auto_reset;
                reinterpret_cast<Real*>(b.reward)[e] = shaped;
                b.done[e] = truncated ? 1 : 0;
                float* o = obs_frame(p, b, e);  // uav_env.py:644
                float f0, f1, f2;
                if constexpr (kParity) {
                    f0 = (float)((double)x / (double)R.i32(kRGridW));
                    f1 = (float)((double)y / (double)R.i32(kRGridH));
                    f2 = (float)(battery / p.max_battery);
                } else {
                    f0 = __fdividef(x, (float)R.i32(kRGridW));
                    f1 = __fdividef(y, (float)R.i32(kRGridH));
                    f2 = (float)battery * p.inv_max_battery;
                }
                if (feat_wind<LEAN>(p)) {  // :220-223: the observation reports the raw drifted position
                    f0 = x;
                    f1 = y;
                }
                if (feat_gps<LEAN>(p)) {   // GPSNoisyEnv._add_gps_noise, :176-181
                    float gx, gy;
                    double unused;
                    env_draws<Real, REPLAY>(p, b.step_env_draws, e, ts.episode[lane] - 1, (uint32_t)ts.step[lane], gx, gy, unused);
                    f0 = f0 + gx;
                    f1 = f1 + gy;
                }
                o[0] = f0;
                o[1] = f1;
                o[2] = f2;
                if constexpr (STACK) {
                    o[p.frame_mirror] = f0;
                    o[p.frame_mirror + 1] = f1;
                    o[p.frame_mirror + 2] = f2;
                }
                if (b.action_mask && !restart) b.action_mask[e] = action_mask_of(x, y, R.i32(kRGridW), R.i32(kRGridH));
                if (feat_step_detail<LEAN>(b)) {
                    double* d = feat_step_detail<LEAN>(b) + (size_t)e * UAVSIM_DETAIL_COUNT;
                    if (!collect)  // a collect step's intermediates were written by phase B
                        for (int i = 0; i < UAVSIM_DETAIL_BASE_REWARD; ++i) d[i] = 0.0;
                    d[UAVSIM_DETAIL_DATA_LOSS] = (double)step_loss;
                    d[UAVSIM_DETAIL_BASE_REWARD] = (double)reward;
                    d[UAVSIM_DETAIL_TRUNCATED] = truncated;
                    d[14] = 0.0;
                    d[15] = 0.0;
                }
            }
            restart_mask = __ballot_sync(kFull, restart);
        }
        __syncwarp();

        // ---- SB3 VecEnv semantics: finished episodes restart inside the same call (rare) ----
        while (restart_mask) {
            const int k = __ffs(restart_mask) - 1;
            restart_mask &= restart_mask - 1;
            __syncwarp();
            reinterpret_cast<uint32_t*>(&ws.rec)[lane] = rec[k][lane ^ k];  // unswizzle into the working record
            __syncwarp();
            finish_and_restart<Real, REPLAY, SPL, FPS>(p, b, base + k, lane, ws);
            __syncwarp();
            rec[k][lane ^ k] = reinterpret_cast<const uint32_t*>(&ws.rec)[lane];
        }
        __syncwarp();

        // ---- tile records: shared -> global ----
#pragma unroll
        for (int r = 0; r < kTile; ++r)
            if ((valid_mask >> r) & 1u)
                reinterpret_cast<uint32_t*>(b.rec + base + r)[lane] = rec[r][lane ^ r];
        if constexpr (kStaged) {
            if (ntile < 4) {  // a short (= last) tile never reached the waits that retire the record group
                cp_async_wait<0>();
            }
            __syncwarp();
            tb ^= 1;
        }
        base = next_base;
        next_base = (int)__reduce_max_sync(kFull, (unsigned)pending_ticket) * b.tile;
    }
#undef UAVSIM_E
#undef UAVSIM_NEXT_BASE
#undef UAVSIM_TAKE_TICKET
}

}  // namespace uavsim
