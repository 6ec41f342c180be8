// Construction-time layout generator (included by the C-ABI translation unit only).
#pragma once
#include "uavsim_kernels.cuh"

namespace uavsim {

// Uniform layouts at construction: np.random.uniform([0,0],[W,H],(N,2)) in distribution
// (uav_env.py:366-374), episode counter word 0xffffffff so it never collides with an episode.
__global__ void __launch_bounds__(kThreads) layout_kernel(const __grid_constant__ Buffers b, int P, uint32_t key0,
                                                          uint32_t key1, int64_t env_id_offset) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * kWarpsPerBlock + warp;
    if (e >= b.E) return;
    const uavsim_env_record r = b.rec[e];
    const uint64_t gid = (uint64_t)(env_id_offset + e);
    for (int j = lane; j < r.num_sensors; j += 32) {
        const Philox4 q = philox4x32_10((uint32_t)gid, ((uint32_t)(gid >> 32) & 0xffffffu) | (kStreamLayout << 24),
                                        0xffffffffu, (uint32_t)j, key0, key1);
        b.spos[(size_t)e * P + j] =
            make_float2((float)(u01<double>(q.x) * (double)r.grid_w), (float)(u01<double>(q.y) * (double)r.grid_h));
    }
}

}  // namespace uavsim
