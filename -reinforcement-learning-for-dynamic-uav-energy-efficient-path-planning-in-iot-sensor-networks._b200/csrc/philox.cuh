// Counter-based Philox4x32-10 (Salmon et al., SC'11) for the throughput path.
//
// The reference draws from three stateful host generators ("RNG
// substrate"): the global legacy numpy MT19937 for every shadowing draw
// (iot_sensors.py:192), Python's `random` for the duty-cycle gate (uav_env.py:549) and a
// per-env PCG64 for the buffer fill (uav_env.py:410).  None of those can be advanced
// independently by a million environments, so the batched simulator addresses each draw by
// a counter instead of by its position in a stream:
//     key     = (seed_lo, seed_hi)
//     counter = (global env id, stream id | hi env bits, episode, step << 8 | lane << 3 | call)
// (sensor j sits in lane j % 32, slot j / 32 of a call's four words; P <= 32: the observation draws of a step are keyed
// by the even / odd environment PAIR instead, stream kStreamPair -- see Draws in uavsim_device.cuh)
// A draw therefore depends only on (seed, env id, episode, step, sensor, slot) -- not on how
// environments are sharded over GPUs, not on launch geometry, not on other environments.
#pragma once
#include <stdint.h>

namespace uavsim {

struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        // one 32x32->64 multiply per lane of the bijection (IMAD.WIDE.U32), not separate hi and lo products
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// The same bijection with the ten round keys (k + r*W) supplied by the caller: the key is a per-handle
// constant, so the schedule is computed once on the host and read from the constant bank (20 fewer integer
// adds per call in the step kernel).
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};
__host__ __device__ inline PhiloxKeys philox_key_schedule(uint32_t k0, uint32_t k1) {
    PhiloxKeys ks;
    for (int r = 0; r < 10; ++r) {
        ks.k0[r] = k0 + (uint32_t)r * 0x9E3779B9u;
        ks.k1[r] = k1 + (uint32_t)r * 0xBB67AE85u;
    }
    return ks;
}
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& ks) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        c0 = hi1 ^ c1 ^ ks.k0[r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ ks.k1[r];
        c3 = lo0;
    }
    return Philox4{c0, c1, c2, c3};
}

// stream ids (counter word 1, high byte)
enum : uint32_t {
    kStreamStep = 0,    // per-sensor draws of a step / of the reset observation (step 0)
    kStreamFill = 1,    // initial buffer fill, uav_env.py:410
    kStreamLayout = 2,  // sensor layout, uav_env.py:366-374 / dqn.py:342-343
    kStreamStart = 3,   // far-start candidates, dqn.py:375-403
    kStreamMisc = 4,    // grid choice (dqn.py:283) and the inherited-SF draw (dqn.py:345-353)
    kStreamPair = 6     // observation draws of a step shared by an even/odd environment pair (P <= 32; 5 = driver probes)
};

}  // namespace uavsim
