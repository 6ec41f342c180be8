// Launch entry points shared between the two kernel translation units and the C-ABI layer.
#pragma once
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

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
cudaError_t launch_info_f64(const Params<double>& p, const Buffers& b, double* info, double* ratios, cudaStream_t st);
cudaError_t launch_info_f32(const Params<float>& p, const Buffers& b, double* info, double* ratios, cudaStream_t st);
cudaError_t launch_nearest_greedy_f64(const Params<double>& p, const Buffers& b, const double* probe_draws,
                                      int32_t* actions, bool replay, cudaStream_t st);
cudaError_t launch_nearest_greedy_f32(const Params<float>& p, const Buffers& b, const double* probe_draws,
                                      int32_t* actions, bool replay, cudaStream_t st);

cudaError_t launch_max_throughput_greedy_f64(const Params<double>& p, const Buffers& b, const double* probe_draws,
                                              int32_t* actions, bool replay, cudaStream_t st);
cudaError_t launch_max_throughput_greedy_f32(const Params<float>& p, const Buffers& b, const double* probe_draws,
                                              int32_t* actions, bool replay, cudaStream_t st);

// Blocks of `k` resident at once on the current device (SM count x occupancy).
template <typename Kernel>
inline int resident_blocks(Kernel k, size_t dyn_smem, int warps) {
    // keyed by (kernel, device): different instantiations share a function-pointer TYPE, so a plain
    // function-local static would be shared between them
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find({(const void*)k, dev});
        if (it != cache.end()) return it->second;
    }
    int sms = 0, occ = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dyn_smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 32 * warps, dyn_smem);
    const int blocks = (sms > 0 ? sms : 148) * (occ > 0 ? occ : 1);
    std::lock_guard<std::mutex> lock(mu);
    cache[{(const void*)k, dev}] = blocks;
    return blocks;
}

// Persistent grid: one wave of blocks at full residency, never more blocks than there are groups of
// warps x envs_per_warp environments.
template <typename Kernel>
inline int persistent_grid(Kernel k, int E, size_t dyn_smem = 0, int envs_per_warp = 1, int warps = kWarpsPerBlock) {
    const int blocks_per_wave = resident_blocks(k, dyn_smem, warps);
    const int per_block = warps * envs_per_warp;
    const int need = (E + per_block - 1) / per_block;
    return need < blocks_per_wave ? need : blocks_per_wave;
}

// Environments per tile of a step launch: 16 when every resident warp gets at least ~2.5 tiles, else the largest smaller
// tile that does (phases A / C run one environment per lane, so a smaller tile costs a few per cent per environment but
// lets a small batch use the whole machine).  The staged (float32) kernels prefetch two environments / one pair ahead
// across tile boundaries and need a multiple of 4.  UAVSIM_TILES_PER_WARP overrides the 2.5 for sweeps; measured flat:
// a 131,072-env shard (strong scaling over 8 GPUs) takes 0.109 / 0.105 / 0.113 ms at 2.5 / 5 / 8 tiles per warp, so the
// 30 % such a shard loses against the 1 M-env rate is launch + pipeline ramp, not the tile tail
// (profiles/r02d_tiles_per_warp_ab.txt).
inline double tiles_per_warp_target() {
    static const double v = getenv("UAVSIM_TILES_PER_WARP") ? atof(getenv("UAVSIM_TILES_PER_WARP")) : 2.5;
    return v;
}
inline int choose_tile(int E, int resident_warps, bool min4) {
    for (int t = kTile; t >= (min4 ? 4 : 2); t >>= 1)
        if ((double)E >= tiles_per_warp_target() * (double)t * (double)resident_warps) return t;
    return min4 ? 4 : 2;
}

template <typename Kernel>
inline void launch_step_kernel(Kernel k, size_t smem, int warps, bool min4, const Buffers& b, dim3* grid, Buffers* bb) {
    const int blocks = resident_blocks(k, smem, warps);
    *bb = b;
    bb->tile = choose_tile(b.E, blocks * warps, min4);
    const int per_block = warps * bb->tile;
    const int need = (b.E + per_block - 1) / per_block;
    *grid = dim3((unsigned)(need < blocks ? need : blocks));
}

// Throughput build, Philox, no diagnostics: the plain UAVEnvironment step, the DomainRandEnv training step and the
// RelationalUAVEnv step run LEAN instantiations (the features they do not use are compiled out, finished episodes restart
// outside the tile loop); everything else takes the generic kernel.
template <typename Real>
inline bool lean_eligible(const Params<Real>& p, const Buffers& b, bool replay) {
#ifdef UAVSIM_NO_LEAN
    return false;
#else
    return !std::is_same<Real, double>::value && !replay && !b.draw_export && !b.step_detail && !(p.gps_sigma > Real(0)) &&
           !b.env_sigma && !b.env_duty && !b.step_mask && !b.extra_step &&
           !(p.wind_drift > 0.0) && ((p.fps == 3 && !p.relational) || (p.fps == 5 && p.relational && !p.domain_rand));
#endif
}
// Kernels one step launches.
template <typename Real>
inline int step_launch_count(const Params<Real>&, const Buffers&, bool) {
    return 1;
}

template <typename Real>
inline cudaError_t launch_step_t(const Params<Real>& p, const Buffers& b, const int32_t* actions, bool replay,
                                 cudaStream_t st) {
    const bool two = p.P > 32;
    const bool lean_base = lean_eligible(p, b, replay);
    if constexpr (!std::is_same<Real, double>::value) {
        if (lean_base) {
            const bool stack = p.frame_k > 1;
            dim3 grid;
            Buffers bb;
#define UAVSIM_LAUNCH_LEAN(S, F, MODE, PAIRED)                                                                   \
    do {                                                                                                      \
        constexpr size_t smem = step_smem_bytes<Real, S, PAIRED>();                                           \
        constexpr int warps = step_warps<Real, S, PAIRED>();                                                  \
        auto k = stack ? step_kernel<Real, false, S, F, true, MODE, PAIRED> : step_kernel<Real, false, S, F, false, MODE, PAIRED>; \
        launch_step_kernel(k, smem, warps, true, b, &grid, &bb);                                            \
        k<<<grid, 32 * warps, smem, st>>>(p, bb, actions);                                                    \
    } while (0)
            if (p.relational) {  // RelationalUAVEnv rows (5 features per sensor)
                if (two)
                    UAVSIM_LAUNCH_LEAN(2, 5, kLeanRelational, false);
                else
                    UAVSIM_LAUNCH_LEAN(1, 5, kLeanRelational, false);
            } else if (p.domain_rand) {
                if (two)
                    UAVSIM_LAUNCH_LEAN(2, 3, kLeanDomainRand, false);
                else  // P <= 32: pair kernel
                    UAVSIM_LAUNCH_LEAN(1, 3, kLeanDomainRand, true);
            } else if (two) {
                UAVSIM_LAUNCH_LEAN(2, 3, kLeanPlain, false);
            } else {  // P <= 32: two environments per warp pass (uavsim_step.cuh, pair_phase)
                UAVSIM_LAUNCH_LEAN(1, 3, kLeanPlain, true);
            }
#undef UAVSIM_LAUNCH_LEAN
            return cudaGetLastError();
        }
    }
#define UAVSIM_LAUNCH_STEP(R, S)                                                              \
    do {                                                                                      \
        constexpr size_t smem = step_smem_bytes<Real, S>();                                   \
        const bool stack = p.frame_k > 1;                                                     \
        auto k = p.fps == 5 ? (stack ? step_kernel<Real, R, S, 5, true> : step_kernel<Real, R, S, 5, false>)  \
                            : (stack ? step_kernel<Real, R, S, 3, true> : step_kernel<Real, R, S, 3, false>); \
        constexpr int warps = step_warps<Real, S>();                                          \
        dim3 grid;                                                                            \
        Buffers bb;                                                                           \
        launch_step_kernel(k, smem, warps, !std::is_same<Real, double>::value, b, &grid, &bb); \
        k<<<grid, 32 * warps, smem, st>>>(p, bb, actions);                                    \
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

template <typename Real>
inline cudaError_t launch_info_t(const Params<Real>& p, const Buffers& b, double* info, double* ratios, cudaStream_t st) {
    if (p.P > 32) {
        auto k = info_kernel<Real, 2>;
        k<<<persistent_grid(k, b.E), kThreads, 0, st>>>(p, b, info, ratios);
    } else {
        auto k = info_kernel<Real, 1>;
        k<<<persistent_grid(k, b.E), kThreads, 0, st>>>(p, b, info, ratios);
    }
    return cudaGetLastError();
}

}  // namespace uavsim
