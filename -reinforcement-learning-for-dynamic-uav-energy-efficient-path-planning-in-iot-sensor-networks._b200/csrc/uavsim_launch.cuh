// Launch entry points shared between the two kernel translation units and the C-ABI layer.
#pragma once
#include "uavsim_kernels.cuh"

namespace uavsim {

cudaError_t launch_step_f64(const Params<double>& p, const Buffers& b, const int32_t* actions, bool replay,
                            cudaStream_t st);
cudaError_t launch_reset_f64(const Params<double>& p, const Buffers& b, const uint8_t* mask, bool replay,
                             cudaStream_t st);
cudaError_t launch_step_f32(const Params<float>& p, const Buffers& b, const int32_t* actions, bool replay,
                            cudaStream_t st);
cudaError_t launch_reset_f32(const Params<float>& p, const Buffers& b, const uint8_t* mask, bool replay,
                             cudaStream_t st);

// Persistent grid: one wave of blocks at full residency (SM count x occupancy), never more blocks than
// there are groups of kWarpsPerBlock environments.
template <typename Kernel>
inline int persistent_grid(Kernel k, int E) {
    static int blocks_per_wave = 0;  // one static per kernel instantiation
    if (blocks_per_wave == 0) {
        int dev = 0, sms = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kThreads, 0);
        blocks_per_wave = (sms > 0 ? sms : 148) * (occ > 0 ? occ : 1);
    }
    const int need = (E + kWarpsPerBlock - 1) / kWarpsPerBlock;
    return need < blocks_per_wave ? need : blocks_per_wave;
}

template <typename Real>
inline cudaError_t launch_step_t(const Params<Real>& p, const Buffers& b, const int32_t* actions, bool replay,
                                 cudaStream_t st) {
    const bool two = p.P > 32;
#define UAVSIM_LAUNCH_STEP(R, S)                                                              \
    do {                                                                                      \
        if (p.fps == 5) {                                                                     \
            auto k = step_kernel<Real, R, S, 5>;                                              \
            k<<<persistent_grid(k, b.E), kThreads, 0, st>>>(p, b, actions);                   \
        } else {                                                                              \
            auto k = step_kernel<Real, R, S, 3>;                                              \
            k<<<persistent_grid(k, b.E), kThreads, 0, st>>>(p, b, actions);                   \
        }                                                                                     \
    } while (0)
    if (replay) {
        if (two)
            UAVSIM_LAUNCH_STEP(true, 2);
        else
            UAVSIM_LAUNCH_STEP(true, 1);
    } else {
        if (two)
            UAVSIM_LAUNCH_STEP(false, 2);
        else
            UAVSIM_LAUNCH_STEP(false, 1);
    }
#undef UAVSIM_LAUNCH_STEP
    return cudaGetLastError();
}

template <typename Real>
inline cudaError_t launch_reset_t(const Params<Real>& p, const Buffers& b, const uint8_t* mask, bool replay,
                                  cudaStream_t st) {
    const bool two = p.P > 32;
#define UAVSIM_LAUNCH_RESET(R, S)                                                             \
    do {                                                                                      \
        if (p.fps == 5) {                                                                     \
            auto k = reset_kernel<Real, R, S, 5>;                                             \
            k<<<persistent_grid(k, b.E), kThreads, 0, st>>>(p, b, mask);                      \
        } else {                                                                              \
            auto k = reset_kernel<Real, R, S, 3>;                                             \
            k<<<persistent_grid(k, b.E), kThreads, 0, st>>>(p, b, mask);                      \
        }                                                                                     \
    } while (0)
    if (replay) {
        if (two)
            UAVSIM_LAUNCH_RESET(true, 2);
        else
            UAVSIM_LAUNCH_RESET(true, 1);
    } else {
        if (two)
            UAVSIM_LAUNCH_RESET(false, 2);
        else
            UAVSIM_LAUNCH_RESET(false, 1);
    }
#undef UAVSIM_LAUNCH_RESET
    return cudaGetLastError();
}

}  // namespace uavsim
