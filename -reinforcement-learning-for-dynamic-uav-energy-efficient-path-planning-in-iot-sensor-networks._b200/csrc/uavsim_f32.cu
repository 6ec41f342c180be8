// THROUGHPUT build of the step/reset kernels (Real = float): float32 state, MUFU intrinsics,
// FMA contraction on, shuffle-tree reductions.
#include "uavsim_launch.cuh"
#include "uavsim_policy.cuh"
#include "uavsim_packed.cuh"

namespace uavsim {
template <int L, int MODE, bool DIAG>
static cudaError_t launch_packed_one(const Params<float>& p, const Buffers& b, const int32_t* actions, cudaStream_t st) {
    auto k = packed_step_kernel<L, MODE, DIAG>;
    const int S = (p.P + 2 * L - 1) / (2 * L);
    const int wb = pack_warp_bytes(S);
    static std::mutex mu;
    static std::map<std::pair<int, int>, std::pair<int, int>> cache;  // (device, S) -> (warps, SMs)
    int dev = 0;
    cudaGetDevice(&dev);
    int warps = 0, sms = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find({dev, S});
        if (it != cache.end()) {
            warps = it->second.first;
            sms = it->second.second;
        }
    }
    if (!warps) {
        int max_smem = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        warps = max_smem / wb;
        if (const char* w = getenv("UAVSIM_PACK_WARPS")) warps = atoi(w) < warps ? atoi(w) : warps;
        if (warps > kPackMaxWarps) warps = kPackMaxWarps;
        if (warps < 1) return cudaErrorInvalidConfiguration;
        std::lock_guard<std::mutex> lock(mu);
        cache[{dev, S}] = {warps, sms};
    }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, warps * wb);
    const int tiles = (b.E + kPT - 1) / kPT;
    int grid = (tiles + warps - 1) / warps;
    if (grid > sms) grid = sms;
    k<<<grid, 32 * warps, (size_t)warps * wb, st>>>(p, b, actions, S, wb);
    return cudaGetLastError();
}
template <int MODE, bool DIAG>
static cudaError_t launch_packed_mode(const Params<float>& p, const Buffers& b, const int32_t* actions, cudaStream_t st) {
    static const int force_l = getenv("UAVSIM_PACK_L") ? atoi(getenv("UAVSIM_PACK_L")) : 0;
    int L = p.P <= 32 ? 2 : 4;
    if (force_l == 2 || force_l == 4 || force_l == 8) L = force_l;
    if (2 * L * 8 < p.P) L = p.P <= 32 ? 2 : (p.P <= 64 && 2 * 4 * 8 >= p.P ? 4 : 8);
    switch (L) {
        case 2: return launch_packed_one<2, MODE, DIAG>(p, b, actions, st);
        case 8: return launch_packed_one<8, MODE, DIAG>(p, b, actions, st);
        default: return launch_packed_one<4, MODE, DIAG>(p, b, actions, st);
    }
}
cudaError_t launch_packed(const Params<float>& p, const Buffers& b, const int32_t* actions, cudaStream_t st) {
    const bool diag = b.step_detail || b.draw_export;
    if (p.domain_rand) return diag ? launch_packed_mode<kLeanDomainRand, true>(p, b, actions, st) : launch_packed_mode<kLeanDomainRand, false>(p, b, actions, st);
    return diag ? launch_packed_mode<kLeanPlain, true>(p, b, actions, st) : launch_packed_mode<kLeanPlain, false>(p, b, actions, st);
}


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
