// On-device heuristic drivers: NearestSensorGreedy and MaxThroughputGreedyV2 of the reference
// (src/agents/dqn/dqn_evaluation_results/greedy_agents.py:73-100 with GreedyAgent._move_toward :41-68),
// batched: one warp per environment, sensors across lanes, reading the same state columns the step kernel
// owns.  It is the driver of the reference's config-1 rollouts (sweep_eval.py:468-486) and of the curriculum's
// greedy-benchmark gate (dqn.py:456-543), so evaluation no longer needs a per-environment Python loop.
//
//   1. COLLECT if any sensor with data is in range: `sensor.data_buffer > 0 and sensor.is_in_range(pos)`
//      -- is_in_range draws a fresh shadowing value (iot_sensors.py:214-219); the reference stops probing at
//      the first hit, which decides the same action as probing everyone.
//   2. otherwise walk one cell toward the nearest sensor with data (float32 norm, first index on ties):
//      COLLECT when within half a cell on both axes or when the move would leave the grid.
//
// Probe draws: REPLAY reads a table [E][P] (NaN = the reference never made that draw), PHILOX uses its own
// stream id so the environment's draws are untouched.
#pragma once
#include "uavsim_kernels.cuh"

namespace uavsim {

constexpr uint32_t kStreamPolicy = 5;

// GreedyAgent._move_toward (greedy_agents.py:41-68): one cell along the dominant axis toward (tx, ty); COLLECT
// when already within half a cell on both axes or when the cell would leave the grid.
__device__ __forceinline__ int move_toward(float ux, float uy, float tx, float ty, int W, int H) {
    const float dx = tx - ux, dy = ty - uy;
    if (fabsf(dx) <= 0.5f && fabsf(dy) <= 0.5f) return 4;
    float nx = ux, ny = uy;
    if (fabsf(dx) > fabsf(dy))
        nx = ux + (dx > 0.f ? 1.0f : -1.0f);
    else
        ny = uy + (dy > 0.f ? 1.0f : -1.0f);
    if (nx < 0.f || nx >= (float)W || ny < 0.f || ny >= (float)H) return 4;
    const float mx = nx - ux, my = ny - uy;
    if (mx > 0.f) return 3;  // RIGHT
    if (mx < 0.f) return 2;  // LEFT
    if (my > 0.f) return 0;  // UP
    if (my < 0.f) return 1;  // DOWN
    return 4;
}

// Shadowing draws of the agent's is_in_range probes: REPLAY reads the table [E][P], PHILOX its own stream.
template <typename Real, bool REPLAY, int SPL>
__device__ __forceinline__ void policy_probes(const Params<Real>& p, const uavsim_env_record* R, int e, int lane,
                                              const double* __restrict__ probe_draws, Real (&g)[SPL], Real sg) {
    if constexpr (REPLAY) {
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const int j = lane + 32 * k;
            g[k] = j < p.P ? (Real)probe_draws[(size_t)e * p.P + j] : Real(CUDART_NAN);
        }
    } else {
        const uint64_t gid = (uint64_t)(p.env_id_offset + e);
        const Philox4 q = philox4x32_10((uint32_t)gid, ((uint32_t)(gid >> 32) & 0xffffffu) | (kStreamPolicy << 24),
                                        R->episode, ((uint32_t)R->step << 8) | ((uint32_t)lane << 3), p.keys);
        Real unused;
        box_muller<Real>(q.x, q.y, sg, g[0], unused);
        if constexpr (SPL == 2) box_muller<Real>(q.z, q.w, sg, g[1], unused);
    }
}

// Position of sensor `idx` (held by lane idx & 31, slot idx >> 5) broadcast to the warp.
template <typename Real, int SPL>
__device__ __forceinline__ void sensor_xy(const Sensor<Real> (&s)[SPL], int idx, float& tx, float& ty) {
    tx = ty = 0.f;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const float sx = shfl(s[k].sx, idx & 31), sy = shfl(s[k].sy, idx & 31);
        if ((idx >> 5) == k) {
            tx = sx;
            ty = sy;
        }
    }
}

template <typename Real, bool REPLAY, int SPL>
__global__ void __launch_bounds__(kThreads) nearest_greedy_kernel(const __grid_constant__ Params<Real> p,
                                                                  const __grid_constant__ Buffers b,
                                                                  const double* __restrict__ probe_draws,
                                                                  int32_t* __restrict__ actions) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = blockIdx.x * kWarpsPerBlock + warp; e < b.E; e += gridDim.x * kWarpsPerBlock) {
        const uavsim_env_record* R = b.rec + e;
        const float ux = R->x, uy = R->y;
        const int W = R->grid_w, H = R->grid_h, N = R->num_sensors;
        Sensor<Real> s[SPL];
        load_sensors<Real, SPL>(b, p, e, N, lane, s);
        compute_det(p, ux, uy, s);

        // 1. any sensor with data in range?
        Real g[SPL];
        policy_probes<Real, REPLAY, SPL>(p, R, e, lane, probe_draws, g, b.env_sigma ? (Real)b.env_sigma[e] : p.sigma);
        bool hit = false;
#pragma unroll
        for (int k = 0; k < SPL; ++k) hit |= s[k].active && s[k].buf > Real(0) && (s[k].det + g[k]) >= p.thr;
        int action = 4;
        if (!__any_sync(kFull, hit)) {
            // 2. nearest sensor with data: float32 norm, lowest index on ties (Python's min keeps the first)
            float best = CUDART_INF_F;
            int bi = -1;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                if (s[k].active && s[k].buf > Real(0)) {
                    const float dx = s[k].sx - ux, dy = s[k].sy - uy;
                    const float d = sqrtf(dx * dx + dy * dy);
                    if (d < best) {
                        best = d;
                        bi = lane + 32 * k;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const float od = __shfl_xor_sync(kFull, best, o);
                const int oi = __shfl_xor_sync(kFull, bi, o);
                if (oi >= 0 && (bi < 0 || od < best || (od == best && oi < bi))) {
                    best = od;
                    bi = oi;
                }
            }
            if (bi >= 0) {
                float tx, ty;
                sensor_xy<Real, SPL>(s, bi, tx, ty);
                action = move_toward(ux, uy, tx, ty, W, H);
            }
        }
        if (lane == 0) actions[e] = action;
    }
}


// MaxThroughputGreedyV2.select_action (greedy_agents.py:105-200), batched the same way.
//   1. COLLECT if any sensor with data is in range (every sensor with data is probed, :133-138) AND its spreading
//      factor is within the adaptive threshold (9 / 10 / 12 by remaining battery and steps, :188-197);
//   2. otherwise walk toward the best-scoring sensor with data (:155-182):
//        score = max(0, 13 - SF) * 5 * sf_w  +  buffer / max_buffer * 10  +  duty_cycle_probability * 2
//                - distance / grid_w * 5 * dist_w            (first index wins ties: strict > at :178)
//      with sf_w = 1 / 2 / 5 by remaining battery and steps (:157-162).
// Arithmetic types follow the reference under numpy-2 promotion: the distance term is float32; the sum is
// float64 unless the sensor has ever overflowed, in which case `data_buffer` descends from the Python float
// `max_buffer_size` (iot_sensors.py:120) and subtracting the float32 penalty rounds the score to float32.
template <typename Real, bool REPLAY, int SPL>
__global__ void __launch_bounds__(kThreads) max_throughput_greedy_kernel(const __grid_constant__ Params<Real> p,
                                                                         const __grid_constant__ Buffers b,
                                                                         const double* __restrict__ probe_draws,
                                                                         int32_t* __restrict__ actions) {
    constexpr bool kParity = std::is_same<Real, double>::value;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = blockIdx.x * kWarpsPerBlock + warp; e < b.E; e += gridDim.x * kWarpsPerBlock) {
        const uavsim_env_record* R = b.rec + e;
        const float ux = R->x, uy = R->y;
        const int W = R->grid_w, H = R->grid_h, N = R->num_sensors;
        Sensor<Real> s[SPL];
        load_sensors<Real, SPL>(b, p, e, N, lane, s);
        compute_det(p, ux, uy, s);
        Real g[SPL];
        policy_probes<Real, REPLAY, SPL>(p, R, e, lane, probe_draws, g, b.env_sigma ? (Real)b.env_sigma[e] : p.sigma);

        const double battery_pct = R->battery / 274.0;  // :126 (hard-coded capacity)
        const int steps_left = p.max_steps - R->step;
        const double steps_ratio = fmin(1.0, (double)steps_left / (double)p.max_steps);
        const int sf_thr = (battery_pct > 0.5 && steps_ratio > 0.5) ? 9 : ((battery_pct > 0.2 && steps_ratio > 0.2) ? 10 : 12);
        bool hit = false;
#pragma unroll
        for (int k = 0; k < SPL; ++k)
            hit |= s[k].active && s[k].buf > Real(0) && (s[k].det + g[k]) >= p.thr && s[k].sf <= sf_thr;
        int action = 4;
        if (!__any_sync(kFull, hit)) {
            const double sf_w = (battery_pct < 0.1 || steps_left < 50) ? 1.0 : ((battery_pct < 0.3 || steps_left < 150) ? 2.0 : 5.0);
            Real best = Real(0);
            int bi = -1;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                if (s[k].active && s[k].buf > Real(0)) {
                    const float dx = s[k].sx - ux, dy = s[k].sy - uy;
                    const float dist = sqrtf(dx * dx + dy * dy);
                    const int pri = 13 - s[k].sf > 0 ? 13 - s[k].sf : 0;
                    Real score;
                    if constexpr (kParity) {
                        const double sum = (((double)pri * 5.0) * sf_w + (s[k].buf / p.maxbuf) * 10.0) + p.duty_frac * 2.0;
                        const float pen = ((dist / (float)W) * 5.0f) * 1.0f;
                        score = s[k].lost > 0.0 ? (double)((float)sum - pen) : sum - (double)pen;
                    } else {
                        score = ((float)pri * 5.0f) * (float)sf_w + (s[k].buf * p.inv_maxbuf) * 10.0f + p.duty_frac * 2.0f -
                                (dist / (float)W) * 5.0f;
                    }
                    if (bi < 0 || score > best) {
                        best = score;
                        bi = lane + 32 * k;
                    }
                }
            }
            Real top;
            int ti;
            group_argmax(best, bi, top, ti);
            if (ti >= 0) {
                float tx, ty;
                sensor_xy<Real, SPL>(s, ti, tx, ty);
                action = move_toward(ux, uy, tx, ty, W, H);
            }
        }
        if (lane == 0) actions[e] = action;
    }
}

template <typename Real>
inline cudaError_t launch_nearest_greedy_t(const Params<Real>& p, const Buffers& b, const double* probe_draws,
                                           int32_t* actions, bool replay, cudaStream_t st) {
    const int grid = (b.E + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int blocks = grid < 148 * 8 ? grid : 148 * 8;
    const bool two = p.P > 32;
    if (replay) {
        if (two)
            nearest_greedy_kernel<Real, true, 2><<<blocks, kThreads, 0, st>>>(p, b, probe_draws, actions);
        else
            nearest_greedy_kernel<Real, true, 1><<<blocks, kThreads, 0, st>>>(p, b, probe_draws, actions);
    } else {
        if (two)
            nearest_greedy_kernel<Real, false, 2><<<blocks, kThreads, 0, st>>>(p, b, probe_draws, actions);
        else
            nearest_greedy_kernel<Real, false, 1><<<blocks, kThreads, 0, st>>>(p, b, probe_draws, actions);
    }
    return cudaGetLastError();
}

template <typename Real>
inline cudaError_t launch_max_throughput_greedy_t(const Params<Real>& p, const Buffers& b, const double* probe_draws,
                                                  int32_t* actions, bool replay, cudaStream_t st) {
    const int grid = (b.E + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int blocks = grid < 148 * 8 ? grid : 148 * 8;
    const bool two = p.P > 32;
    if (replay) {
        if (two)
            max_throughput_greedy_kernel<Real, true, 2><<<blocks, kThreads, 0, st>>>(p, b, probe_draws, actions);
        else
            max_throughput_greedy_kernel<Real, true, 1><<<blocks, kThreads, 0, st>>>(p, b, probe_draws, actions);
    } else {
        if (two)
            max_throughput_greedy_kernel<Real, false, 2><<<blocks, kThreads, 0, st>>>(p, b, probe_draws, actions);
        else
            max_throughput_greedy_kernel<Real, false, 1><<<blocks, kThreads, 0, st>>>(p, b, probe_draws, actions);
    }
    return cudaGetLastError();
}

}  // namespace uavsim
