// PARITY build of the step/reset kernels (Real = double).  Compile with -fmad=false: the reference
// evaluates every a*b+c as two rounded numpy operations, so nothing here may be contracted.
#include "uavsim_launch.cuh"
#include "uavsim_policy.cuh"

namespace uavsim {
cudaError_t launch_step_f64(const Params<double>& p, const Buffers& b, const int32_t* actions, bool replay,
                            cudaStream_t st) {
    return launch_step_t<double>(p, b, actions, replay, st);
}
cudaError_t launch_reset_f64(const Params<double>& p, const Buffers& b, const uint8_t* mask, bool replay,
                             cudaStream_t st) {
    return launch_reset_t<double>(p, b, mask, replay, st);
}
cudaError_t launch_nearest_greedy_f64(const Params<double>& p, const Buffers& b, const double* probe_draws,
                                      int32_t* actions, bool replay, cudaStream_t st) {
    return launch_nearest_greedy_t<double>(p, b, probe_draws, actions, replay, st);
}
cudaError_t launch_max_throughput_greedy_f64(const Params<double>& p, const Buffers& b, const double* probe_draws,
                                              int32_t* actions, bool replay, cudaStream_t st) {
    return launch_max_throughput_greedy_t<double>(p, b, probe_draws, actions, replay, st);
}
cudaError_t launch_info_f64(const Params<double>& p, const Buffers& b, double* info, double* ratios, cudaStream_t st) {
    return launch_info_t<double>(p, b, info, ratios, st);
}
}  // namespace uavsim
