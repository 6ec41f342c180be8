// The PACKED step kernel of the throughput build: several environments per warp pass.
//
// The first-generation kernel (uavsim_step.cuh) gives one environment the whole warp: sensor j sits in lane j % 32.
// At N = 20 only 20 of 32 lanes carry a sensor (62 %), at N = 50 50 of 64 lane-slots (78 %), and everything that is
// per ENVIRONMENT rather than per sensor -- loop control, staging issue, the five-step shuffle trees of every
// reduction, the ballots of the capture resolution -- is paid by 32 lanes to serve one environment.  Measured: 321
// warp-instructions per env-step at N = 20 and 480 at N = 50, issue-bound at 58 % / 79 % of the HBM roofline.
//
// Here a warp pass serves G = 32 / L environments, L lanes each (L = 2 for P <= 32, 4 for P <= 64):
//   * lane (q, r) = (lane / L, lane % L) owns the sensor PAIRS i = ps * L + r, ps = 0 .. S-1, of environment q
//     (sensors 2i and 2i + 1: 32 B of state, 16 B of positions, 8 B of `generated` -- the L lanes of an environment
//     read contiguous 32 L-byte runs of its rows); 2 L S >= P, so N = 20 fills L = 2, S = 5 completely and N = 50
//     fills 50 of the 56 places of L = 4, S = 7;
//   * one Philox4x32-10 call serves the two sensors of a pair: call 0 -> D, E of both (a MOVE step costs one call per
//     pair), call 1 -> A, B of both, call 2 -> U of both + the pair's two C draws (used by the rare SF-group winner);
//   * per-environment reductions run over L lanes (log2 L xor-shuffle steps) after each lane has summed its own 2 S
//     sensors;
//   * the Capture-Effect resolution (uav_env.py:551-594) leaves the lanes altogether: attempts are rare (duty cycle
//     10 % x link probability), so attempting sensors drop (current RSSI, SF) into a per-environment scratch row and
//     ONE lane per environment walks the set bits of the attempt mask in sensor order -- SF groups in first-appearance
//     order, strongest / runner-up by strict comparison (= Python's stable sort), winners range-checked and drained
//     straight in the shared-memory copy of the state -- while the observation features it changes (buffer, urgency of
//     the winner) are patched in the staged row;
//   * a tile is 32 environments, ordered COLLECT-first before its groups are formed, so that the collect-only work
//     (draws A, B, U, the first ADR update, the link sigmoid: almost as much as the rest of the step) runs in
//     ceil(collecting / G) of the 32 / G passes instead of in nearly all of them;
//   * phases A and C (one environment per lane, see uavsim_step.cuh) now use all 32 lanes.
// The sensor columns of the next group stream into the other of two shared-memory stages with cp.async while the
// current group is computed; the updated state is written back in place and leaves with 16-byte stores, the
// observation rows are staged and leave as runs of L consecutive floats per environment.
//
// Scope: float32 state, Philox, 3 features per sensor, one frame per environment, plain UAVEnvironment or
// DomainRandEnv (MODE), even P.  DIAG compiles the parity diagnostics and the draw export in: the DIAG instantiation is
// what tests/ compare with the float64 oracle (tests/parity_util.py, draws="philox"), and the production
// instantiation must agree with it bit for bit (same arithmetic, fewer stores).  Everything else -- replay tables,
// float64 parity build, relational / sim-to-real wrappers, 5-feature rows, frame rings -- stays on uavsim_step.cuh.
#pragma once
#include "uavsim_step.cuh"

namespace uavsim {

constexpr int kPT = 32;              // environments per tile
constexpr int kPackMaxWarps = 12;    // warps per block (one block per SM)
constexpr uint32_t kStreamPacked = 7;  // Philox stream of the per-pair step draws (counter word 3: step << 8 | pair << 2 | call)

struct alignas(16) PackTile {
    uint32_t rec[2][kPT][32];  // [buffer][l][w ^ l]: swizzled records of this tile / the next one
    int32_t action[2][kPT];
    uint8_t order[2][kPT];     // tile slots, COLLECT actions first
    int32_t ncollect[2];
    float x[kPT], y[kPT], elapsed[kPT], used[kPT];
    int32_t N[kPT], step[kPT], flags[kPT];
    uint32_t episode[kPT], vis_lo[kPT], vis_hi[kPT];
    double battery[kPT];
    float step_loss[kPT], creward[kPT], bytes[kPT], jain[kPT], cur_dist[kPT];
    int32_t counts[kPT];       // collisions | captures << 8 | delivered << 12 | starved << 16
    uint32_t nvis_lo[kPT], nvis_hi[kPT];
    uint32_t att[kPT][2];      // attempt mask of the environments of the group in flight (indexed by q)
};

// Per-warp dynamic shared memory: tile block, restart scratch, two input stages, observation staging, attempt scratch.
__host__ __device__ inline int pack_stage_bytes(int S) { return S * 32 * 56; }  // 2 x 16 B state, 16 B positions, 8 B generated per lane-slot
__host__ __device__ inline int pack_warp_bytes(int S) {
    int b = (int)sizeof(PackTile) + (int)sizeof(WarpSmem);
    b += 2 * pack_stage_bytes(S);
    b += S * 32 * 24;  // staged observation features: 6 floats per lane-slot
    b += S * 32 * 8;   // current RSSI (draw A) of attempting sensors: [G][2 L S] floats
    b += S * 32 * 2;   // their spreading factor after the collect-phase ADR update: [G][2 L S] bytes
    return (b + 15) & ~15;
}

struct PackStage {
    float4* st0;  // [S][32] state of the even sensor of the pair
    float4* st1;  // [S][32] state of the odd sensor
    float4* pos;  // [S][32] (x0, y0, x1, y1)
    float2* gen;  // [S][32] generated-at-reset of both
};
__device__ __forceinline__ PackStage pack_stage(unsigned char* base, int S) {
    PackStage s;
    s.st0 = reinterpret_cast<float4*>(base);
    s.st1 = s.st0 + S * 32;
    s.pos = s.st1 + S * 32;
    s.gen = reinterpret_cast<float2*>(s.pos + S * 32);
    return s;
}

template <int L>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
    for (int o = L / 2; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
template <int L>
__device__ __forceinline__ int seg_sum(int v) {
#pragma unroll
    for (int o = L / 2; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
template <int L>
__device__ __forceinline__ float seg_max(float v) {
#pragma unroll
    for (int o = L / 2; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
template <int L>
__device__ __forceinline__ float seg_min(float v) {
#pragma unroll
    for (int o = L / 2; o; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// What the resolving lane of a collecting environment hands to the reward assembly.
struct CollectOutcome {
    float total_bytes, urg_acc, urg_red;
    int nwin, collisions, captures, delivered;
    bool any_new;
    uint32_t vis_lo, vis_hi, win_lo, win_hi;
};

// uav_env.py:551-594 for ONE environment, run by one lane on the shared-memory copy of its state: SF groups in
// first-appearance order, capture test, range check (draw C) and drain of each winner (iot_sensors.py:127-145).
template <int L, bool DIAG>
__device__ __forceinline__ void pack_resolve(const Params<float>& p, const PackStage& sg, int q, uint32_t att_lo, uint32_t att_hi,
                                             const float* cur_row, const uint8_t* sfa_row, float* obs_row, float ux, float uy,
                                             float elapsed, uint32_t env_lo, uint32_t env_hi, uint32_t episode, uint32_t step,
                                             float* exp, CollectOutcome& out) {
    unsigned long long remaining = (unsigned long long)att_lo | ((unsigned long long)att_hi << 32);
    while (remaining) {
        const int i0 = __ffsll((long long)remaining) - 1;
        const int sf0 = sfa_row[i0];
        int cnt = 0, top = -1, sec = -1;
        float tv = 0.f, sv = 0.f;
        unsigned long long m = remaining, gm = 0ull;
        while (m) {  // ascending sensor index; strict comparisons keep ties in index order (Python's stable sort)
            const int j = __ffsll((long long)m) - 1;
            m &= m - 1;
            if (sfa_row[j] != sf0) continue;
            gm |= 1ull << j;
            cnt += 1;
            const float c = cur_row[j];
            if (top < 0 || c > tv) {
                sec = top;
                sv = tv;
                top = j;
                tv = c;
            } else if (sec < 0 || c > sv) {
                sec = j;
                sv = c;
            }
        }
        remaining &= ~gm;
        int winner = -1;
        if (cnt == 1) {
            winner = top;
        } else {
            out.collisions += cnt - 1;
            if (tv > (sv + p.capture_db)) {  // uav_env.py:569
                winner = top;
                out.captures += 1;
            }
        }
        if (winner < 0) continue;
        if (winner < 32)
            out.win_lo |= 1u << winner;
        else
            out.win_hi |= 1u << (winner - 32);
        // the winner's state: pair i = winner / 2 lives in lane-slot (ps, q * L + r)
        const int i = winner >> 1, ps = i / L, ln = q * L + (i % L), at = ps * 32 + ln;
        float4 v = (winner & 1) ? sg.st1[at] : sg.st0[at];
        const float4 xy = sg.pos[at];
        const float2 gg = sg.gen[at];
        const float sx = (winner & 1) ? xy.z : xy.x, sy = (winner & 1) ? xy.w : xy.y;
        const float g0 = (winner & 1) ? gg.y : gg.x;
        const Philox4 qc = philox4x32_10(env_lo, env_hi | (kStreamPacked << 24), episode, (step << 8) | ((uint32_t)i << 2) | 2u, p.keys);
        float c0, c1;
        box_muller<float>(qc.z, qc.w, p.sigma, c0, c1);
        const float c = (winner & 1) ? c1 : c0;
        if (DIAG && exp) exp[UAVSIM_SLOT_C * p.P + winner] = c;
        const float rc = det_rssi(p, ux, uy, sx, sy) + c;
        const bool in_range = rc >= p.thr;
        float got = 0.f;
        if (in_range && v.x > 0.f) {
            const float cap = data_rate_of<float>(sf0) * p.cd;
            got = fminf(v.x, cap);
            v.x -= got;
            v.y += got;
            if (winner & 1)
                sg.st1[at] = v;
            else
                sg.st0[at] = v;
        }
        const float urg_w = urgency_of(p, v.x, v.z, fmaf(p.rate, elapsed, g0));
        obs_row[3 * winner + 0] = v.x * p.inv_maxbuf;  // the observation pass ran on the pre-drain buffer
        obs_row[3 * winner + 1] = urg_w;
        out.urg_red += got * p.inv_rate;
        out.urg_acc += urg_w;  // np.mean over <= 6 winners adds left to right (uav_env.py:610-615)
        out.nwin += 1;
        if (in_range && got > 0.f) {
            out.total_bytes += got;
            out.delivered += 1;
            const bool seen = winner < 32 ? (out.vis_lo >> winner) & 1u : (out.vis_hi >> (winner - 32)) & 1u;
            if (!seen) {
                out.any_new = true;
                if (winner < 32)
                    out.vis_lo |= 1u << winner;
                else
                    out.vis_hi |= 1u << (winner - 32);
            }
        }
    }
}

template <int L, int MODE, bool DIAG>
__global__ void __launch_bounds__(32 * kPackMaxWarps, 1)
    packed_step_kernel(const __grid_constant__ Params<float> p, const __grid_constant__ Buffers b,
                       const int32_t* __restrict__ actions, const int S, const int warp_bytes) {
    constexpr int G = 32 / L;
    constexpr bool kDR = MODE == kLeanDomainRand;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane / L, r = lane % L;
    unsigned char* wbase = dyn_smem + (size_t)warp * (size_t)warp_bytes;
    PackTile& ts = *reinterpret_cast<PackTile*>(wbase);
    WarpSmem& ws = *reinterpret_cast<WarpSmem*>(wbase + sizeof(PackTile));
    unsigned char* dyn = wbase + sizeof(PackTile) + sizeof(WarpSmem);
    const int stage_bytes = pack_stage_bytes(S);
    float* const obsst = reinterpret_cast<float*>(dyn + 2 * stage_bytes);  // [G][OW]
    const int OW = 6 * L * S;                                              // staged floats per environment
    float* const cur_s = obsst + 32 * 6 * S;                               // [G][CW]
    const int CW = 2 * L * S;                                              // sensor places per environment
    uint8_t* const sfa_s = reinterpret_cast<uint8_t*>(cur_s + 32 * 2 * S);
    const float4* const g_state = reinterpret_cast<const float4*>(b.sstate);
    float4* const g_state_w = reinterpret_cast<float4*>(b.sstate);
    const float* const g_gen = reinterpret_cast<const float*>(b.sgen);
    const int E = b.E;
    unsigned* const ticket_ctr = b.tile_ticket;

    // cp.async copies of group g of the tile in record buffer tb into stage buffer sb
    auto issue_group = [&](int tb, int tbase, int nt, int g, int sb) {
        const int slot = g * G + q;
        if (slot < nt) {
            const int k = ts.order[tb][slot];
            const int N = (int)ts.rec[tb][k][kRN ^ k];
            const uint32_t row = (uint32_t)(tbase + k) * (uint32_t)p.P;
            const PackStage sg = pack_stage(dyn + sb * stage_bytes, S);
            for (int ps = 0; ps < S; ++ps) {
                const int j0 = 2 * (ps * L + r);
                if (j0 < N) {
                    const uint32_t i = row + (uint32_t)j0;
                    const int at = ps * 32 + lane;
                    cp_async16(sg.st0 + at, g_state + i);
                    cp_async16(sg.st1 + at, g_state + i + 1);
                    cp_async16(sg.pos + at, b.spos + i);
                    cp_async8(sg.gen + at, g_gen + i);
                }
            }
        }
    };
    // records + actions of the tile at tbase into record buffer tb
    auto issue_tile = [&](int tbase, int tb) {
        const int n = min(kPT, E - tbase);
#pragma unroll
        for (int rr = 0; rr < kPT; ++rr)
            if (rr < n) cp_async4(&ts.rec[tb][rr][lane ^ rr], reinterpret_cast<const uint32_t*>(b.rec + tbase + rr) + lane);
        if (lane < n) cp_async4(&ts.action[tb][lane], actions + tbase + lane);
    };
    // tile order: COLLECT actions first (the actions have landed in ts.action[tb])
    auto make_order = [&](int tb, int nt) {
        const bool in = lane < nt;
        const bool col = in && ts.action[tb][lane] == 4;
        const unsigned cm = __ballot_sync(kFull, col), im = __ballot_sync(kFull, in);
        const unsigned lt = (1u << lane) - 1u;
        const int ncol = __popc(cm);
        if (in) ts.order[tb][col ? __popc(cm & lt) : ncol + __popc(im & ~cm & lt)] = (uint8_t)lane;
        if (lane == 0) ts.ncollect[tb] = ncol;
        __syncwarp();
    };

    int base, next_base;
    {
        const int t0 = (lane == 0) ? (int)atomicAdd(ticket_ctr, 2u) : 0;
        const int t1 = (lane == 0) ? t0 + 1 : 0;
        base = (int)__reduce_max_sync(kFull, (unsigned)t0) * kPT;
        next_base = (int)__reduce_max_sync(kFull, (unsigned)t1) * kPT;
    }
    int tb = 0, sb = 0;
    if (base < E) {  // prime: records of the first tile, then its first group
        issue_tile(base, 0);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        make_order(0, min(kPT, E - base));
        issue_group(0, base, min(kPT, E - base), 0, 0);
    }
    cp_async_commit();

    while (base < E) {
        const int pending_ticket = (lane == 0) ? (int)atomicAdd(ticket_ctr, 1u) : 0;
        const int nt = min(kPT, E - base);
        const int NG = (nt + G - 1) / G;
        uint32_t(*rec)[32] = ts.rec[tb];

        // ---- phase A: one environment per lane (uav.py:127-224, uav_env.py:439-478) ----
        const RecView R{rec[lane], lane};
        if (lane < nt) {
            const int e = base + lane;
            int my_flags = 0;
            const int my_action = ts.action[tb][lane];
            ts.step[lane] = R.i32(kRStep) + 1;
            ts.episode[lane] = R.w(kREpisode);
            ts.N[lane] = R.i32(kRN);
            if (my_action < 0 || my_action > 4) {  // uav_env.py:468 raises ValueError: flag it, leave the env untouched
                atomicMin(b.error_word, ((unsigned long long)b.step_serial << 32) | (unsigned)e);
            } else {
                my_flags = kFlagValid;
                float x = R.f(kRX), y = R.f(kRY);
                const int W = R.i32(kRGridW), H = R.i32(kRGridH);
                double battery = R.d(kRBattery);
                const int step = R.i32(kRStep) + 1;  // uav_env.py:439
                const float eps = 1e-6f;             // uav_env.py:443-447: edge bookkeeping on the PRE-move position
                if (x <= eps || y <= eps || x >= (float)(W - 1) - eps || y >= (float)(H - 1) - eps) my_flags |= kFlagEdge;
                const bool collect = my_action == 4;
                float used;
                if (!collect) {  // uav.py:127-185
                    float nx = x, ny = y;
                    if (my_action == 0)
                        ny += 1.0f;
                    else if (my_action == 1)
                        ny -= 1.0f;
                    else if (my_action == 2)
                        nx -= 1.0f;
                    else
                        nx += 1.0f;
                    const double before = battery;
                    const bool ok = (0.f <= nx) && (nx < (float)W) && (0.f <= ny) && (ny < (float)H);
                    if (ok) {
                        x = nx;
                        y = ny;
                        battery -= p.e_move;
                        my_flags |= kFlagMoveOk;
                    } else {
                        battery -= p.e_fail;
                    }
                    used = (float)(before - battery);
                } else {  // uav.py:187-206, :260-263
                    my_flags |= kFlagCollect;
                    battery -= p.e_hover;
                    used = (float)p.used_hover;
                }
                if (!(battery > p.batt_floor) || step >= p.max_steps) my_flags |= kFlagTruncated;  // uav.py:224, uav_env.py:474-478
                ts.x[lane] = x;
                ts.y[lane] = y;
                ts.elapsed[lane] = R.f(kRElapsed) + (collect ? p.cd : 1.0f);
                ts.vis_lo[lane] = R.w(kRVisLo);
                ts.vis_hi[lane] = R.w(kRVisHi);
                ts.battery[lane] = battery;
                ts.used[lane] = used;
            }
            ts.flags[lane] = my_flags;
            ts.counts[lane] = 0;
        }
        __syncwarp();

        // ---- phase B: G environments per pass ----
        bool next_ordered = false;
        for (int g = 0; g < NG; ++g) {
            cp_async_wait<0>();  // this group's columns (requested one group ago) and, from g == 1 on, the next tile's records
            __syncwarp();
            if (g == 0) {  // records + actions of the next tile: the other record buffer is free since the last write-back
                if (next_base < E) issue_tile(next_base, tb ^ 1);
                cp_async_commit();
            }
            if (g + 1 < NG) {
                issue_group(tb, base, nt, g + 1, sb ^ 1);
            } else if (next_base < E) {
                if (NG == 1) {  // a one-group tile has not passed a wait that retires the record copies yet
                    cp_async_wait<0>();
                    __syncwarp();
                }
                if (!next_ordered) make_order(tb ^ 1, min(kPT, E - next_base));
                next_ordered = true;
                issue_group(tb ^ 1, next_base, min(kPT, E - next_base), 0, sb ^ 1);
            }
            cp_async_commit();

            const PackStage sg = pack_stage(dyn + sb * stage_bytes, S);
            const int slot = g * G + q;
            const bool have = slot < nt;
            const int k = have ? ts.order[tb][slot] : 0;
            const int flags = have ? ts.flags[k] : 0;
            const bool valid = flags & kFlagValid, collect = flags & kFlagCollect;
            const int N = valid ? ts.N[k] : 0;
            const int e = base + k;
            const float ux = ts.x[k], uy = ts.y[k], elapsed = ts.elapsed[k];
            const bool grp_collect = g * G < ts.ncollect[tb];  // the tile is ordered collect-first
            const uint64_t gid = (uint64_t)(p.env_id_offset + e);
            const uint32_t env_lo = (uint32_t)gid, env_hi = ((uint32_t)(gid >> 32) & 0xffffffu) | (kStreamPacked << 24);
            const uint32_t episode = ts.episode[k] - 1, stepw = (uint32_t)ts.step[k] << 8;
            const uint32_t row = (uint32_t)e * (uint32_t)p.P;
            float* const obs_row = obsst + q * OW;
            float* const cur_row = cur_s + q * CW;
            uint8_t* const sfa_row = sfa_s + q * CW;
            float* exp = nullptr;
            if constexpr (DIAG) {
                if (b.draw_export && valid) {
                    exp = reinterpret_cast<float*>(b.draw_export) + (size_t)e * UAVSIM_NUM_SLOTS * p.P;
                    const float nan = CUDART_NAN_F;
                    for (int i = r; i < UAVSIM_NUM_SLOTS * p.P; i += L) exp[i] = nan;
                }
                __syncwarp();
            }
            if (grp_collect) {
                if (r == 0) ts.att[q][0] = ts.att[q][1] = 0u;
                __syncwarp();
            }

            // ---- pass 1: everything per sensor, state updated in place in the stage ----
            float loss_acc = 0.f;
            const float nw = p.rate * (collect ? p.cd : 1.0f);
            for (int ps = 0; ps < S; ++ps) {
                const int i = ps * L + r, j0 = 2 * i, at = ps * 32 + lane;
                const bool act0 = j0 < N, act1 = j0 + 1 < N;
                float4 v0 = sg.st0[at], v1 = sg.st1[at];
                const float4 xy = sg.pos[at];
                const float2 gg = sg.gen[at];
                float ema0, ema1;
                int sf0, sf1;
                unpack_sf(v0.w, ema0, sf0);
                unpack_sf(v1.w, ema1, sf1);
                // iot_sensors.py:114-125: the buffer grows, is capped, the excess is lost
                {
                    const float pot0 = v0.x + nw, pot1 = v1.x + nw;
                    const float l0 = fmaxf(pot0 - p.maxbuf, 0.f), l1 = fmaxf(pot1 - p.maxbuf, 0.f);
                    v0.x = fminf(pot0, p.maxbuf);
                    v1.x = fminf(pot1, p.maxbuf);
                    v0.z += l0;
                    v1.z += l1;
                    loss_acc += (act0 ? l0 : 0.f) + (act1 ? l1 : 0.f);
                }
                const float det0 = det_rssi(p, ux, uy, xy.x, xy.y), det1 = det_rssi(p, ux, uy, xy.z, xy.w);
                if (grp_collect) {  // uav_env.py:535-551: ADR update (draw A), link probability (B), duty-cycle gate (U)
                    const Philox4 q1 = philox4x32_10(env_lo, env_hi, episode, stepw | ((uint32_t)i << 2) | 1u, p.keys);
                    const Philox4 q2 = philox4x32_10(env_lo, env_hi, episode, stepw | ((uint32_t)i << 2) | 2u, p.keys);
                    float da0, db0, da1, db1;
                    box_muller<float>(q1.x, q1.y, p.sigma, da0, db0);
                    box_muller<float>(q1.z, q1.w, p.sigma, da1, db1);
                    const float du0 = u01<float>(q2.x), du1 = u01<float>(q2.y);
                    if (DIAG && exp && collect) {
                        if (j0 < p.P) {
                            exp[UAVSIM_SLOT_A * p.P + j0] = da0;
                            exp[UAVSIM_SLOT_B * p.P + j0] = db0;
                            exp[UAVSIM_SLOT_U * p.P + j0] = du0;
                        }
                        if (j0 + 1 < p.P) {
                            exp[UAVSIM_SLOT_A * p.P + j0 + 1] = da1;
                            exp[UAVSIM_SLOT_B * p.P + j0 + 1] = db1;
                            exp[UAVSIM_SLOT_U * p.P + j0 + 1] = du1;
                        }
                    }
                    bool att0, att1;
                    float cur0, cur1;
                    {
                        const bool has = act0 && collect && v0.x > 0.f;
                        cur0 = det0 + da0;
                        float ema2 = ema0;
                        int sf2 = sf0;
                        adr_update(p, cur0, ema2, sf2);
                        ema0 = has ? ema2 : ema0;
                        sf0 = has ? sf2 : sf0;
                        const float rb = det0 + db0;
                        const float xx = (rb - p.noise_floor) - required_snr_of<float>(sf2);
                        const float plink = rb < p.thr ? 0.f : __fdividef(1.0f, 1.0f + __expf(-xx));
                        att0 = has && (plink * p.duty_frac) > du0;
                    }
                    {
                        const bool has = act1 && collect && v1.x > 0.f;
                        cur1 = det1 + da1;
                        float ema2 = ema1;
                        int sf2 = sf1;
                        adr_update(p, cur1, ema2, sf2);
                        ema1 = has ? ema2 : ema1;
                        sf1 = has ? sf2 : sf1;
                        const float rb = det1 + db1;
                        const float xx = (rb - p.noise_floor) - required_snr_of<float>(sf2);
                        const float plink = rb < p.thr ? 0.f : __fdividef(1.0f, 1.0f + __expf(-xx));
                        att1 = has && (plink * p.duty_frac) > du1;
                    }
                    if (att0 | att1) {  // rare: duty cycle x link probability
                        if (att0) {
                            cur_row[j0] = cur0;
                            sfa_row[j0] = (uint8_t)sf0;
                        }
                        if (att1) {
                            cur_row[j0 + 1] = cur1;
                            sfa_row[j0 + 1] = (uint8_t)sf1;
                        }
                        const unsigned long long bits = ((unsigned long long)att0 << j0) | ((unsigned long long)att1 << (j0 + 1));
                        if ((uint32_t)bits) atomicOr(&ts.att[q][0], (uint32_t)bits);
                        if ((uint32_t)(bits >> 32)) atomicOr(&ts.att[q][1], (uint32_t)(bits >> 32));
                    }
                }
                // uav_env.py:638-674: urgency, second ADR update (draw D), range check (draw E), link quality
                {
                    const Philox4 q0 = philox4x32_10(env_lo, env_hi, episode, stepw | ((uint32_t)i << 2), p.keys);
                    float dd0, de0, dd1, de1;
                    box_muller<float>(q0.x, q0.y, p.sigma, dd0, de0);
                    box_muller<float>(q0.z, q0.w, p.sigma, dd1, de1);
                    if (DIAG && exp) {
                        if (j0 < p.P) {
                            exp[UAVSIM_SLOT_D * p.P + j0] = dd0;
                            exp[UAVSIM_SLOT_E * p.P + j0] = de0;
                        }
                        if (j0 + 1 < p.P) {
                            exp[UAVSIM_SLOT_D * p.P + j0 + 1] = dd1;
                            exp[UAVSIM_SLOT_E * p.P + j0 + 1] = de1;
                        }
                    }
                    const float urg0 = urgency_of(p, v0.x, v0.z, fmaf(p.rate, elapsed, gg.x));
                    const float urg1 = urgency_of(p, v1.x, v1.z, fmaf(p.rate, elapsed, gg.y));
                    adr_update(p, det0 + dd0, ema0, sf0);
                    adr_update(p, det1 + dd1, ema1, sf1);
                    const bool in0 = (det0 + de0) >= p.thr, in1 = (det1 + de1) >= p.thr;
                    float* o = obs_row + 6 * i;
                    o[0] = v0.x * p.inv_maxbuf;
                    o[1] = urg0;
                    o[2] = in0 ? link_quality_of<float>(sf0) : 0.0f;
                    o[3] = v1.x * p.inv_maxbuf;
                    o[4] = urg1;
                    o[5] = in1 ? link_quality_of<float>(sf1) : 0.0f;
                }
                v0.w = pack_sf(ema0, sf0);
                v1.w = pack_sf(ema1, sf1);
                sg.st0[at] = v0;
                sg.st1[at] = v1;
            }
            // iot_sensors.py:114-125 summed over the environment (uav_env.py:457-459)
            {
                float step_loss = 0.f;
                if (__any_sync(kFull, loss_acc > 0.f)) step_loss = seg_sum<L>(loss_acc);
                if (r == 0 && valid) ts.step_loss[k] = step_loss;
            }
            __syncwarp();

            // ---- Capture-Effect resolution: one lane per collecting environment with attempts (rare) ----
            CollectOutcome oc;
            oc.total_bytes = oc.urg_acc = oc.urg_red = 0.f;
            oc.nwin = oc.collisions = oc.captures = oc.delivered = 0;
            oc.any_new = false;
            oc.vis_lo = ts.vis_lo[k];
            oc.vis_hi = ts.vis_hi[k];
            oc.win_lo = oc.win_hi = 0u;
            uint32_t att_lo = 0u, att_hi = 0u;
            if (grp_collect) {
                if (r == 0 && collect) {
                    att_lo = ts.att[q][0];
                    att_hi = ts.att[q][1];
                    if (att_lo | att_hi)
                        pack_resolve<L, DIAG>(p, sg, q, att_lo, att_hi, cur_row, sfa_row, obs_row, ux, uy, elapsed, env_lo, env_hi,
                                              episode, (uint32_t)ts.step[k], exp, oc);
                }
                __syncwarp();
            }

            // ---- per-environment reductions over the final state (collect reward, shaping inputs, starvation) ----
            const bool grp_trunc = __any_sync(kFull, (flags & kFlagTruncated) != 0);
            if (grp_collect || kDR || grp_trunc) {
                float mx = 0.f, s1 = 0.f, dist = CUDART_INF_F, j1s = 0.f, j2s = 0.f;
                int n_empty = 0, n_has = 0, starved = 0;
                for (int ps = 0; ps < S; ++ps) {
                    const int j0 = 2 * (ps * L + r), at = ps * 32 + lane;
                    const bool act0 = j0 < N, act1 = j0 + 1 < N;
                    const float4 v0 = sg.st0[at], v1 = sg.st1[at];
                    const float b0 = act0 ? v0.x : 0.f, b1 = act1 ? v1.x : 0.f;
                    mx = fmaxf(mx, fmaxf(b0, b1));
                    s1 += b0 + b1;
                    n_empty += (act0 && v0.x <= 0.f) + (act1 && v1.x <= 0.f);
                    if (kDR || grp_trunc) {
                        const float2 gg = sg.gen[at];
                        const float g0 = fmaf(p.rate, elapsed, gg.x), g1 = fmaf(p.rate, elapsed, gg.y);
                        const bool h0 = act0 && g0 > 0.f, h1 = act1 && g1 > 0.f;
                        const float c0 = h0 ? __fdividef(v0.y, g0) : 0.f, c1 = h1 ? __fdividef(v1.y, g1) : 0.f;
                        starved += (h0 && c0 < p.cr_thr) + (h1 && c1 < p.cr_thr);  // reward_function.py:59-67
                        if constexpr (kDR) {  // dqn.py:405-412 (nearest sensor with data), :433-444 (Jain's index of the rates)
                            const float r0 = c0 * 100.f, r1 = c1 * 100.f;
                            j1s += r0 + r1;
                            j2s = fmaf(r0, r0, fmaf(r1, r1, j2s));
                            n_has += (int)h0 + (int)h1;
                            const float4 xy = sg.pos[at];
                            const float dx0 = xy.x - ux, dy0 = xy.y - uy, dx1 = xy.z - ux, dy1 = xy.w - uy;
                            if (act0 && v0.x > 0.f) dist = fminf(dist, sqrtf(dx0 * dx0 + dy0 * dy0));
                            if (act1 && v1.x > 0.f) dist = fminf(dist, sqrtf(dx1 * dx1 + dy1 * dy1));
                        }
                    }
                }
                mx = seg_max<L>(mx);
                s1 = seg_sum<L>(s1);
                n_empty = seg_sum<L>(n_empty);
                if (kDR || grp_trunc) starved = seg_sum<L>(starved);
                if constexpr (kDR) {
                    dist = seg_min<L>(dist);
                    j1s = seg_sum<L>(j1s);
                    j2s = seg_sum<L>(j2s);
                    n_has = seg_sum<L>(n_has);
                }
                float var = 0.f;
                if (grp_collect) {  // calculate_starvation_penalty, reward_function.py:46-57: two passes
                    float inv = fast_rcp(fmaxf(mx, 1e-30f)), invn = fast_rcp((float)max(N, 1));
                    inv = fmaf(inv, fmaf(-fmaxf(mx, 1e-30f), inv, 1.0f), inv);  // one Newton step each: they scale every term
                    invn = fmaf(invn, fmaf(-(float)max(N, 1), invn, 1.0f), invn);
                    const float mean = s1 * inv * invn;
                    float s2 = 0.f;
                    for (int ps = 0; ps < S; ++ps) {
                        const int j0 = 2 * (ps * L + r), at = ps * 32 + lane;
                        const float d0 = sg.st0[at].x * inv - mean, d1 = sg.st1[at].x * inv - mean;
                        s2 = j0 < N ? fmaf(d0, d0, s2) : s2;
                        s2 = j0 + 1 < N ? fmaf(d1, d1, s2) : s2;
                    }
                    var = seg_sum<L>(s2) * invn;
                }
                if (r == 0 && valid) {
                    int counts = 0;
                    if (collect) {  // reward_function.py:81-128, assembled in float64 (see uavsim_step.cuh::collect_action)
                        const bool was_empty = n_empty > 0, all_collected = n_empty == N;
                        const float mean_urg = oc.nwin > 0 ? __fdividef(oc.urg_acc, (float)oc.nwin) : 0.f;
                        const float step_loss = ts.step_loss[k];
                        double reward = (double)p.p_step + (double)p.p_hover;
                        if (oc.total_bytes > 0.f) {
                            reward += (double)p.r_per_byte * (double)oc.total_bytes * (double)mean_urg;
                            if (oc.any_new) reward += (double)p.r_new;
                        }
                        if (oc.urg_red > 0.f) reward += (double)p.r_urg_red * (double)oc.urg_red;
                        if (was_empty && oc.total_bytes == 0.f) reward += (double)p.p_revisit;
                        reward += (double)p.p_battery * p.used_hover;
                        if (oc.collisions > 0) reward += (double)p.p_collision * (double)oc.collisions;
                        if (step_loss > 0.f) reward += (double)p.p_data_loss * (double)step_loss;
                        if (N > 1 && mx != 0.f) reward += (double)p.p_starvation * (double)var;
                        if (all_collected) reward += (double)p.r_completion;
                        ts.creward[k] = (float)reward;
                        ts.bytes[k] = oc.total_bytes;
                        counts = oc.collisions | (oc.captures << 8) | (oc.delivered << 12);
                        ts.nvis_lo[k] = oc.vis_lo;
                        ts.nvis_hi[k] = oc.vis_hi;
                        if constexpr (DIAG) {
                            if (b.step_detail) {
                                double* d = b.step_detail + (size_t)e * UAVSIM_DETAIL_COUNT;
                                d[UAVSIM_DETAIL_COLLISIONS] = oc.collisions;
                                d[UAVSIM_DETAIL_BYTES] = (double)oc.total_bytes;
                                d[UAVSIM_DETAIL_URGENCY_REDUCED] = (double)oc.urg_red;
                                d[UAVSIM_DETAIL_MEAN_URGENCY] = (double)mean_urg;
                                d[UAVSIM_DETAIL_WAS_NEW] = oc.any_new;
                                d[UAVSIM_DETAIL_WAS_EMPTY] = was_empty;
                                d[UAVSIM_DETAIL_ALL_COLLECTED] = all_collected;
                                d[UAVSIM_DETAIL_ATTEMPT_MASK_LO] = (double)att_lo;
                                d[UAVSIM_DETAIL_ATTEMPT_MASK_HI] = (double)att_hi;
                                d[UAVSIM_DETAIL_WINNER_MASK_LO] = (double)oc.win_lo;
                                d[UAVSIM_DETAIL_WINNER_MASK_HI] = (double)oc.win_hi;
                            }
                        }
                    }
                    if (flags & kFlagTruncated) counts |= starved << 16;
                    ts.counts[k] = counts;
                    if constexpr (kDR) {
                        ts.cur_dist[k] = dist == CUDART_INF_F ? 0.0f : dist;
                        if (n_has > 0) {
                            ts.jain[k] = j2s > 0.f ? __fdividef(j1s * j1s, (float)n_has * j2s) : 1.0f;
                            ts.flags[k] = flags | kFlagJain;
                        }
                    }
                }
            }
            __syncwarp();

            // ---- copy-out: state with 16-byte stores, observation rows as runs of L floats per environment ----
            {
                for (int ps = 0; ps < S; ++ps) {
                    const int j0 = 2 * (ps * L + r), at = ps * 32 + lane;
                    if (j0 < N) g_state_w[row + (uint32_t)j0] = sg.st0[at];
                    if (j0 + 1 < N) g_state_w[row + (uint32_t)j0 + 1u] = sg.st1[at];
                }
                float* out = b.obs + ((uint32_t)e * (uint32_t)p.obs_stride + 3u);
                const int n3 = 3 * N;
                for (int i = r; i < n3; i += L) out[i] = obs_row[i];
            }
            __syncwarp();
            sb ^= 1;
        }
        __syncwarp();

        // ---- phase C: one environment per lane ----
        unsigned restart_mask = 0u;
        {
            bool restart = false;
            if (lane < nt && (ts.flags[lane] & kFlagValid)) {
                const int e = base + lane;
                const int flags = ts.flags[lane];
                const bool collect = flags & kFlagCollect, truncated = flags & kFlagTruncated;
                const int N = ts.N[lane];
                const float step_loss = ts.step_loss[lane];
                const int counts = ts.counts[lane];
                float reward, total_bytes = 0.f;
                uint32_t vis_lo = ts.vis_lo[lane], vis_hi = ts.vis_hi[lane];
                if (!collect) {  // reward_function.py:69-79 + uav_env.py:508-514
                    reward = p.p_step;
                    reward += (flags & kFlagMoveOk) ? p.r_move : p.p_boundary;
                    reward += p.p_battery * ts.used[lane];
                    reward += p.p_data_loss * step_loss;
                } else {
                    reward = ts.creward[lane];
                    total_bytes = ts.bytes[lane];
                    vis_lo = ts.nvis_lo[lane];
                    vis_hi = ts.nvis_hi[lane];
                }
                if (truncated) {  // uav_env.py:481-485, reward_function.py:59-67
                    const int unvisited = N - (__popc(vis_lo) + __popc(vis_hi));
                    if (unvisited > 0) reward += p.p_unvisited * (float)unvisited;
                    const int starved = (counts >> 16) & 0xff;
                    float pen = 0.f;
                    for (int i = 0; i < starved; ++i) pen += p.p_starved;  // the reference adds -1000 once per sensor
                    reward += pen;
                }
                float shaped = reward;
                if constexpr (kDR) {  // DomainRandEnv.step shaping (dqn.py:415-444)
                    const float prev = R.f(kRPrevDist), curd = ts.cur_dist[lane];
                    if (prev > 0.f) shaped += p.prox_eta * (prev - curd);
                    R.set_f(kRPrevDist, curd);
                    if (R.i32(kRFirstCov) < 0 && (__popc(vis_lo) + __popc(vis_hi)) == N) R.set_i(kRFirstCov, ts.step[lane]);
                    if (flags & kFlagJain) shaped += __fdividef(p.jain_bonus * (ts.jain[lane] - 0.5f), (float)N);
                }
                const float x = ts.x[lane], y = ts.y[lane];
                const double battery = ts.battery[lane];
                R.set_f(kRX, x);
                R.set_f(kRY, y);
                R.set_d(kRBattery, battery);
                R.set_f(kRElapsed, ts.elapsed[lane]);
                R.set_i(kRStep, ts.step[lane]);
                R.set_d(kRTotalReward, R.d(kRTotalReward) + (double)reward);
                R.set_d(kREpReturn, R.d(kREpReturn) + (double)shaped);
                R.set_f(kRLastBytes, total_bytes);
                if (flags & kFlagEdge) R.set_i(kREdge, R.i32(kREdge) + 1);
                if (!collect) {
                    if (!(flags & kFlagMoveOk)) R.set_i(kRBoundary, R.i32(kRBoundary) + 1);
                } else {
                    R.w(kRVisLo) = vis_lo;
                    R.w(kRVisHi) = vis_hi;
                    R.set_d(kRTotalData, R.d(kRTotalData) + (double)total_bytes);
                    R.set_i(kRCapture, R.i32(kRCapture) + ((counts >> 8) & 0xf));
                    R.set_i(kRCollisions, R.i32(kRCollisions) + (counts & 0xff));
                    R.set_i(kRDelivered, R.i32(kRDelivered) + ((counts >> 12) & 0xf));
                }
                restart = truncated && p.auto_reset;
                reinterpret_cast<float*>(b.reward)[e] = shaped;
                b.done[e] = truncated ? 1 : 0;
                float* o = b.obs + (uint32_t)e * (uint32_t)p.obs_stride;  // uav_env.py:644
                o[0] = __fdividef(x, (float)R.i32(kRGridW));
                o[1] = __fdividef(y, (float)R.i32(kRGridH));
                o[2] = (float)battery * p.inv_max_battery;
                if (b.action_mask && !restart) b.action_mask[e] = action_mask_of(x, y, R.i32(kRGridW), R.i32(kRGridH));
                if constexpr (DIAG) {
                    if (b.step_detail) {
                        double* d = b.step_detail + (size_t)e * UAVSIM_DETAIL_COUNT;
                        if (!collect)
                            for (int i = 0; i < UAVSIM_DETAIL_BASE_REWARD; ++i) d[i] = 0.0;
                        d[UAVSIM_DETAIL_DATA_LOSS] = (double)step_loss;
                        d[UAVSIM_DETAIL_BASE_REWARD] = (double)reward;
                        d[UAVSIM_DETAIL_TRUNCATED] = truncated;
                        d[UAVSIM_DETAIL_STARVED] = truncated ? (double)((counts >> 16) & 0xff) : 0.0;
                        d[15] = 0.0;
                    }
                }
            }
            restart_mask = __ballot_sync(kFull, restart);
        }
        __syncwarp();

        // ---- SB3 VecEnv semantics: finished episodes restart inside the same call (rare; uavsim_step.cuh) ----
        while (restart_mask) {
            const int k = __ffs(restart_mask) - 1;
            restart_mask &= restart_mask - 1;
            __syncwarp();
            reinterpret_cast<uint32_t*>(&ws.rec)[lane] = rec[k][lane ^ k];
            __syncwarp();
            if (p.P > 32)
                finish_and_restart<float, false, 2, 3>(p, b, base + k, lane, ws);
            else
                finish_and_restart<float, false, 1, 3>(p, b, base + k, lane, ws);
            __syncwarp();
            rec[k][lane ^ k] = reinterpret_cast<const uint32_t*>(&ws.rec)[lane];
        }
        __syncwarp();

        // ---- tile records: shared -> global ----
        {
            const unsigned valid_mask = __ballot_sync(kFull, lane < nt && (ts.flags[lane] & kFlagValid));
#pragma unroll
            for (int rr = 0; rr < kPT; ++rr)
                if ((valid_mask >> rr) & 1u) reinterpret_cast<uint32_t*>(b.rec + base + rr)[lane] = rec[rr][lane ^ rr];
        }
        __syncwarp();
        tb ^= 1;
        base = next_base;
        next_base = (int)__reduce_max_sync(kFull, (unsigned)pending_ticket) * kPT;
    }
    cp_async_wait<0>();
    __syncthreads();
    if (threadIdx.x == 0) {  // the block that finishes last zeroes the ticket counters for the next launch
        __threadfence();
        if (atomicAdd(b.tile_ticket + 1, 1u) == gridDim.x - 1) {
            b.tile_ticket[0] = 0u;
            b.tile_ticket[1] = 0u;
            __threadfence();
        }
    }
}

}  // namespace uavsim
