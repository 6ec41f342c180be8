// THROUGHPUT build of the step/reset kernels (Real = float): float32 state, MUFU intrinsics,
// FMA contraction on, shuffle-tree reductions.
#include "uavsim_launch.cuh"
#include "uavsim_policy.cuh"

namespace uavsim {
cudaError_t launch_step_f32(const Params<float>& p, const Buffers& b, const int32_t* actions, bool replay,
                            cudaStream_t st) {
    return launch_step_t<float>(p, b, actions, replay, st);
}
cudaError_t launch_reset_f32(const Params<float>& p, const Buffers& b, const uint8_t* mask, bool replay,
                             cudaStream_t st) {
    return launch_reset_t<float>(p, b, mask, replay, st);
}
cudaError_t launch_nearest_greedy_f32(const Params<float>& p, const Buffers& b, const double* probe_draws,
                                      int32_t* actions, bool replay, cudaStream_t st) {
    return launch_nearest_greedy_t<float>(p, b, probe_draws, actions, replay, st);
}
cudaError_t launch_max_throughput_greedy_f32(const Params<float>& p, const Buffers& b, const double* probe_draws,
                                              int32_t* actions, bool replay, cudaStream_t st) {
    return launch_max_throughput_greedy_t<float>(p, b, probe_draws, actions, replay, st);
}
cudaError_t launch_info_f32(const Params<float>& p, const Buffers& b, double* info, double* ratios, cudaStream_t st) {
    return launch_info_t<float>(p, b, info, ratios, st);
}
}  // namespace uavsim
