// THROUGHPUT build of the step/reset kernels (Real = float): float32 state, MUFU intrinsics,
// FMA contraction on, shuffle-tree reductions.
#include "uavsim_launch.cuh"
#include "uavsim_policy.cuh"
#include "uavsim_packed.cuh"

namespace uavsim {
template <int L, int S, int MODE>
static cudaError_t launch_packed_one(const Params<float>& p, const Buffers& b, const int32_t* actions, cudaStream_t st) {
    auto k = packed_step_kernel<L, S, MODE>;
    constexpr int wb = PackDims<L, S>::kWarpBytes;
    static std::mutex mu;
    static std::map<int, std::pair<int, int>> cache;  // device -> (warps, SMs)
    int dev = 0;
    cudaGetDevice(&dev);
    int warps = 0, sms = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(dev);
        if (it != cache.end()) {
            warps = it->second.first;
            sms = it->second.second;
        }
    }
    if (!warps) {
        int max_smem = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        // stay inside the 196 KB shared-memory configuration: the next one (228 KB) leaves 28 KB of L1, and the kernel's
        // direct loads / stores (records, state write-back, the collect path's re-reads) lose their cache: 9 warps in
        // 220 KB ran 47 % slower than 8 warps in 195 KB
        const int budget = max_smem < 196 * 1024 - 1024 ? max_smem : 196 * 1024 - 1024;
        warps = budget / wb;
        if (const char* w = getenv("UAVSIM_PACK_WARPS")) warps = atoi(w) < max_smem / wb ? atoi(w) : max_smem / wb;
        if (warps > kPackMaxWarps) warps = kPackMaxWarps;
        if (warps < 1) return cudaErrorInvalidConfiguration;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, warps * wb);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> lock(mu);
        cache[dev] = {warps, sms};
    }
    const int tiles = (b.E + kPT - 1) / kPT;
    int grid = (tiles + warps - 1) / warps;
    if (grid > sms) grid = sms;
    k<<<grid, 32 * warps, (size_t)warps * wb, st>>>(p, b, actions);
    return cudaGetLastError();
}
// (L, S) by row width: L = 2 lanes per environment up to P = 32 (S = ceil(P / 4) pair-slots per lane), L = 4 up to 64.
template <int MODE>
static cudaError_t launch_packed_mode(const Params<float>& p, const Buffers& b, const int32_t* actions, cudaStream_t st) {
    const int P = p.P;
#define UAVSIM_PACK_CASE(LL, SS) return launch_packed_one<LL, SS, MODE>(p, b, actions, st)
    if (P <= 12) UAVSIM_PACK_CASE(2, 3);
    if (P <= 16) UAVSIM_PACK_CASE(2, 4);
    if (P <= 20) UAVSIM_PACK_CASE(2, 5);
    if (P <= 24) UAVSIM_PACK_CASE(2, 6);
    if (P <= 32) UAVSIM_PACK_CASE(2, 8);
    if (P <= 40) UAVSIM_PACK_CASE(4, 5);
    if (P <= 48) UAVSIM_PACK_CASE(4, 6);
    if (P <= 56) UAVSIM_PACK_CASE(4, 7);
    UAVSIM_PACK_CASE(4, 8);
#undef UAVSIM_PACK_CASE
}
cudaError_t launch_packed(const Params<float>& p, const Buffers& b, const int32_t* actions, cudaStream_t st) {
    return p.domain_rand ? launch_packed_mode<kLeanDomainRand>(p, b, actions, st) : launch_packed_mode<kLeanPlain>(p, b, actions, st);
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
