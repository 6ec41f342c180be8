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

template <typename Real>
inline cudaError_t launch_step_t(const Params<Real>& p, const Buffers& b, const int32_t* actions, bool replay,
                                 cudaStream_t st) {
    const int grid = (b.E + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const bool two = p.P > 32;
    if (replay) {
        if (two)
            step_kernel<Real, true, 2><<<grid, kThreads, 0, st>>>(p, b, actions);
        else
            step_kernel<Real, true, 1><<<grid, kThreads, 0, st>>>(p, b, actions);
    } else {
        if (two)
            step_kernel<Real, false, 2><<<grid, kThreads, 0, st>>>(p, b, actions);
        else
            step_kernel<Real, false, 1><<<grid, kThreads, 0, st>>>(p, b, actions);
    }
    return cudaGetLastError();
}

template <typename Real>
inline cudaError_t launch_reset_t(const Params<Real>& p, const Buffers& b, const uint8_t* mask, bool replay,
                                  cudaStream_t st) {
    const int grid = (b.E + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const bool two = p.P > 32;
    if (replay) {
        if (two)
            reset_kernel<Real, true, 2><<<grid, kThreads, 0, st>>>(p, b, mask);
        else
            reset_kernel<Real, true, 1><<<grid, kThreads, 0, st>>>(p, b, mask);
    } else {
        if (two)
            reset_kernel<Real, false, 2><<<grid, kThreads, 0, st>>>(p, b, mask);
        else
            reset_kernel<Real, false, 1><<<grid, kThreads, 0, st>>>(p, b, mask);
    }
    return cudaGetLastError();
}

}  // namespace uavsim
