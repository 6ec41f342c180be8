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
//     fills 50 of the 56 places of L = 4, S = 7.  S is a template parameter: the pair-slot loop is unrolled, which is
//     where the instruction-level parallelism comes from (2 S independent sensor chains per lane);
//   * one Philox4x32-10 call serves the two sensors of a pair: call 0 -> D, E of both (a MOVE step costs one call per
//     pair), call 1 -> A, B of both, call 2 -> U of both + the pair's two C draws (used by the rare SF-group winner);
//   * the sensor columns STREAM through registers: a pair-slot is read from a shared-memory ring entry, the entry is
//     re-armed at once with the cp.async copies of the same pair-slot of the NEXT group (prefetch distance: one whole
//     group of compute), the updated state leaves with two 16-byte stores, the three observation features per sensor
//     go to a staged row;
//   * per-environment reductions run over L lanes (log2 L xor-shuffle steps) after each lane has summed its own 2 S
//     sensors;
//   * the Capture-Effect resolution (uav_env.py:551-594) leaves the lanes altogether: attempts are rare (duty cycle
//     10 % x link probability), so attempting sensors drop (current RSSI, SF) into a per-environment scratch row and
//     ONE lane per environment walks the set bits of the attempt mask in sensor order -- SF groups in first-appearance
//     order, strongest / runner-up by strict comparison (= Python's stable sort), winners range-checked and drained in
//     the state just written (an L2 hit) -- patches the two observation features a drain changes, and the collect
//     reward's reductions (fairness variance, empty / all-collected tests) then re-read the final state;
//   * a tile is 32 environments, ordered COLLECT-first before its groups are formed, so that the collect-only work
//     (draws A, B, U, the first ADR update, the link sigmoid: almost as much as the rest of the step) runs in
//     ceil(collecting / G) of the 32 / G passes instead of in nearly all of them;
//   * the staged observation rows carry the 16-byte phase of their global rows (a 153-float row starts 12 bytes into
//     a 612-byte pitch), so they leave as aligned 16-byte stores, L lanes per row;
//   * phases A and C (one environment per lane, see uavsim_step.cuh) use all 32 lanes and read / write the 128-byte
//     records straight from global memory, one record per lane (four sectors each way, no staging).
// Shared memory per warp: 3.5 KB of tile scalars + the ring (S x 1792 B) + the staged rows + attempt scratch = 18 KB at
// N = 20 and 24 KB at N = 50: 12 / 9 warps per SM.
//
// Scope: float32 state, Philox, 3 features per sensor, one frame per environment, plain UAVEnvironment or
// DomainRandEnv (MODE), even P.  The parity diagnostics and the draw export are behind warp-uniform runtime tests
// (b.step_detail / b.draw_export): the SAME kernel serves tests/parity_util.py (draws="philox": the float64 oracle
// replays the exported draws) and production.  Everything else -- replay tables, float64 parity build, relational /
// sim-to-real wrappers, 5-feature rows, frame rings -- stays on uavsim_step.cuh.
#pragma once
#include "uavsim_step.cuh"

namespace uavsim {

#ifndef UAVSIM_PACK_UNROLL
#define UAVSIM_PACK_UNROLL 1
#endif
constexpr int kPackUnroll = UAVSIM_PACK_UNROLL;
constexpr int kPT = 32;                // environments per tile
constexpr int kPackMaxWarps = 12;      // warps per block (one block per SM)
constexpr uint32_t kStreamPacked = 7;  // Philox stream of the per-pair step draws (counter word 3: step << 8 | pair << 2 | call)

struct alignas(16) PackTile {
    float x[kPT], y[kPT], elapsed[kPT], used[kPT];
    int32_t N[kPT], step[kPT], flags[kPT];
    uint32_t episode[kPT], vis_lo[kPT], vis_hi[kPT];
    double battery[kPT];
    float step_loss[kPT], creward[kPT], bytes[kPT], jain[kPT], cur_dist[kPT];
    int32_t counts[kPT];       // collisions | captures << 8 | delivered << 12 | starved << 16
    uint32_t nvis_lo[kPT], nvis_hi[kPT];
    uint32_t att[kPT][2];      // attempt mask of the environments of the group in flight (indexed by q)
    uint8_t order[2][kPT];     // tile slots, COLLECT actions first: this tile / the next one
    int32_t ncollect[2];
};

// One ring entry = one pair-slot of a group: per lane 2 x 16 B of state, 16 B of positions, 8 B of `generated`.
struct PackEntry {
    float4 st0[32], st1[32], pos[32];
    float2 gen[32];
};
static_assert(sizeof(PackEntry) == 32 * 56, "ring entry layout");

template <int L, int S>
struct PackDims {
    static constexpr int G = 32 / L;
    static constexpr int CW = 2 * L * S;       // sensor places per environment
    static constexpr int RS = 3 * CW + 4;      // staged floats per environment row (+ 16-byte phase)
    static constexpr int kRing = (int)sizeof(PackEntry) * S;
    static constexpr int kObs = G * RS * 4;
    static constexpr int kCur = G * CW * 4;
    static constexpr int kSfa = (G * CW + 15) & ~15;
    static constexpr int kWarpBytes = ((int)sizeof(PackTile) + (int)sizeof(WarpSmem) + kRing + kObs + kCur + kSfa + 15) & ~15;
};

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

// The 128-byte record of one environment in eight 16-byte registers (phases A and C: one environment per lane).
struct RecRegs {
    uint4 w[8];
    __device__ __forceinline__ void load(const uavsim_env_record* g) {
        const uint4* s = reinterpret_cast<const uint4*>(g);
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = s[i];
    }
    __device__ __forceinline__ void store(uavsim_env_record* g) const {
        uint4* d = reinterpret_cast<uint4*>(g);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = w[i];
    }
    __device__ __forceinline__ uint32_t& u(int i) { return reinterpret_cast<uint32_t*>(w)[i]; }  // i is a constant everywhere
    __device__ __forceinline__ float f(int i) { return __uint_as_float(u(i)); }
    __device__ __forceinline__ int i32(int i) { return (int)u(i); }
    __device__ __forceinline__ double d(int i) { return __hiloint2double((int)u(i + 1), (int)u(i)); }
    __device__ __forceinline__ void set_f(int i, float v) { u(i) = __float_as_uint(v); }
    __device__ __forceinline__ void set_i(int i, int v) { u(i) = (uint32_t)v; }
    __device__ __forceinline__ void set_d(int i, double v) {
        u(i) = (uint32_t)__double2loint(v);
        u(i + 1) = (uint32_t)__double2hiint(v);
    }
};

// What the resolving lane of a collecting environment hands to the reward assembly.
struct CollectOutcome {
    float total_bytes, urg_acc, urg_red;
    int nwin, collisions, captures, delivered;
    bool any_new;
    uint32_t vis_lo, vis_hi, win_lo, win_hi;
};

// uav_env.py:551-594 for ONE environment, run by one lane: SF groups in first-appearance order, capture test, range
// check (draw C) and drain of each winner (iot_sensors.py:127-145) in the state the pass has just stored.
__device__ __forceinline__ void pack_resolve(const Params<float>& p, const Buffers& b, uint32_t row, uint32_t att_lo, uint32_t att_hi,
                                             const float* cur_row, const uint8_t* sfa_row, float* obs_row, float ux, float uy,
                                             float elapsed, uint32_t env_lo, uint32_t env_hi, uint32_t episode, uint32_t stepw,
                                             float* exp, CollectOutcome& out) {
    float4* const g_state = reinterpret_cast<float4*>(b.sstate);
    const float* const g_gen = reinterpret_cast<const float*>(b.sgen);
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
        const uint32_t at = row + (uint32_t)winner;
        float4 v = g_state[at];
        const float2 xy = b.spos[at];
        const float g0 = g_gen[at];
        const Philox4 qc = philox4x32_10(env_lo, env_hi, episode, stepw | ((uint32_t)(winner >> 1) << 2) | 2u, p.keys);
        float c0, c1;
        box_muller<float>(qc.z, qc.w, p.sigma, c0, c1);
        const float c = (winner & 1) ? c1 : c0;
        if (exp) exp[UAVSIM_SLOT_C * p.P + winner] = c;
        const float rc = det_rssi(p, ux, uy, xy.x, xy.y) + c;
        const bool in_range = rc >= p.thr;
        float got = 0.f;
        if (in_range && v.x > 0.f) {
            const float cap = data_rate_of<float>(sf0) * p.cd;
            got = fminf(v.x, cap);
            v.x -= got;
            v.y += got;
            g_state[at] = v;
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

// Per-environment sums a lane accumulates over its own sensors (DomainRandEnv shaping inputs, terminal starvation).
struct EnvAcc {
    float dist, j1s, j2s;
    int n_has, starved;
};
template <bool kDR>
__device__ __forceinline__ void acc_sensor(const Params<float>& p, EnvAcc& a, bool act, float buf, float tx, float gen, float sx,
                                           float sy, float ux, float uy, bool want_starved) {
    const bool h = act && gen > 0.f;
    const float c = h ? __fdividef(tx, gen) : 0.f;
    if (want_starved) a.starved += (int)(h && c < p.cr_thr);  // reward_function.py:59-67
    if constexpr (kDR) {  // dqn.py:405-412 (nearest sensor with data), :433-444 (Jain's index of the collection rates)
        const float rr = c * 100.f;
        a.j1s += rr;
        a.j2s = fmaf(rr, rr, a.j2s);
        a.n_has += (int)h;
        const float dx = sx - ux, dy = sy - uy;
        if (act && buf > 0.f) a.dist = fminf(a.dist, sqrtf(dx * dx + dy * dy));
    }
}

template <int L, int S, int MODE>
__global__ void __launch_bounds__(32 * kPackMaxWarps, 1)
    packed_step_kernel(const __grid_constant__ Params<float> p, const __grid_constant__ Buffers b,
                       const int32_t* __restrict__ actions) {
    using D = PackDims<L, S>;
    constexpr int G = D::G, CW = D::CW, RS = D::RS;
    constexpr bool kDR = MODE == kLeanDomainRand;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane / L, r = lane % L;
    unsigned char* wbase = dyn_smem + (size_t)warp * (size_t)D::kWarpBytes;
    PackTile& ts = *reinterpret_cast<PackTile*>(wbase);
    WarpSmem& ws = *reinterpret_cast<WarpSmem*>(wbase + sizeof(PackTile));
    PackEntry* const ring = reinterpret_cast<PackEntry*>(wbase + sizeof(PackTile) + sizeof(WarpSmem));
    float* const obsst = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ring) + D::kRing);  // [G][RS]
    float* const cur_s = obsst + G * RS;                                                              // [G][CW]
    uint8_t* const sfa_s = reinterpret_cast<uint8_t*>(cur_s + G * CW);                                // [G][CW]
    const float4* const g_state = reinterpret_cast<const float4*>(b.sstate);
    float4* const g_state_w = reinterpret_cast<float4*>(b.sstate);
    const float* const g_gen = reinterpret_cast<const float*>(b.sgen);
    const int E = b.E;
    unsigned* const ticket_ctr = b.tile_ticket;
    const bool diag_detail = b.step_detail != nullptr, diag_export = b.draw_export != nullptr;  // warp-uniform

    // cp.async copies of pair-slot ps of the environment at row `nrow` (N sensors) into this lane's place of ring entry ps
    auto issue_slot = [&](int ps, uint32_t nrow, int nN) {
        const int j0 = 2 * (ps * L + r);
        if (j0 < nN) {
            const uint32_t i = nrow + (uint32_t)j0;
            PackEntry& en = ring[ps];
            cp_async16(&en.st0[lane], g_state + i);
            cp_async16(&en.st1[lane], g_state + i + 1);
            cp_async16(&en.pos[lane], b.spos + i);
            cp_async8(&en.gen[lane], g_gen + i);
        }
    };
    // tile order: COLLECT actions first.  Returns this lane's environment slot -> (is in tile, action == COLLECT).
    auto make_order = [&](int ob, int nt, int my_action) {
        const bool in = lane < nt;
        const bool col = in && my_action == 4;
        const unsigned cm = __ballot_sync(kFull, col), im = __ballot_sync(kFull, in);
        const unsigned lt = (1u << lane) - 1u;
        const int ncol = __popc(cm);
        if (in) ts.order[ob][col ? __popc(cm & lt) : ncol + __popc(im & ~cm & lt)] = (uint8_t)lane;
        if (lane == 0) ts.ncollect[ob] = ncol;
        __syncwarp();
    };

    int base, next_base;
    {
        const int t0 = (lane == 0) ? (int)atomicAdd(ticket_ctr, 2u) : 0;
        const int t1 = (lane == 0) ? t0 + 1 : 0;
        base = (int)__reduce_max_sync(kFull, (unsigned)t0) * kPT;
        next_base = (int)__reduce_max_sync(kFull, (unsigned)t1) * kPT;
    }
    int ob = 0;  // order buffer of the current tile
    // what phase A needs of the NEXT tile ahead of time: its actions (tile order) and sensor counts (prefetch)
    int my_action = 0, my_n = 0;
    if (base + lane < E) {
        my_action = actions[base + lane];
        my_n = b.rec[base + lane].num_sensors;
    }
    if (base < E) {  // prime the ring with group 0 of the first tile
        const int nt = min(kPT, E - base);
        make_order(0, nt, my_action);
        const int k0 = q < nt ? ts.order[0][q] : 0;
        const int n0 = __shfl_sync(kFull, my_n, k0);
        const uint32_t row0 = (uint32_t)(base + k0) * (uint32_t)p.P;
#pragma unroll
        for (int ps = 0; ps < S; ++ps) {
            issue_slot(ps, row0, q < nt ? n0 : 0);
            cp_async_commit();
        }
    }

    while (base < E) {
        const int pending_ticket = (lane == 0) ? (int)atomicAdd(ticket_ctr, 1u) : 0;
        const int nt = min(kPT, E - base);
        const int NG = (nt + G - 1) / G;
        // the next tile's actions + sensor counts (consumed by the last group of this tile)
        int nxt_action = 0, nxt_n = 0;
        if (next_base + lane < E) {
            nxt_action = actions[next_base + lane];
            nxt_n = b.rec[next_base + lane].num_sensors;
        }

        // ---- phase A: one environment per lane (uav.py:127-224, uav_env.py:439-478) ----
        if (lane < nt) {
            const int e = base + lane;
            RecRegs R;
            R.load(b.rec + e);
            int my_flags = 0;
            ts.step[lane] = R.i32(kRStep) + 1;
            ts.episode[lane] = R.u(kREpisode);
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
                ts.vis_lo[lane] = R.u(kRVisLo);
                ts.vis_hi[lane] = R.u(kRVisHi);
                ts.battery[lane] = battery;
                ts.used[lane] = used;
            }
            ts.flags[lane] = my_flags;
            ts.counts[lane] = 0;
        }
        __syncwarp();

        // ---- phase B: G environments per pass ----
        for (int g = 0; g < NG; ++g) {
            // the environment this lane serves in the NEXT group (prefetch target)
            uint32_t nrow = 0;
            int nN = 0;
            if (g + 1 < NG) {
                const int slot = (g + 1) * G + q;
                if (slot < nt) {
                    const int k1 = ts.order[ob][slot];
                    nrow = (uint32_t)(base + k1) * (uint32_t)p.P;
                    nN = ts.N[k1];
                }
            } else if (next_base < E) {
                const int nt2 = min(kPT, E - next_base);
                make_order(ob ^ 1, nt2, nxt_action);
                const int k1 = q < nt2 ? ts.order[ob ^ 1][q] : 0;
                const int n1 = __shfl_sync(kFull, nxt_n, k1);
                nrow = (uint32_t)(next_base + k1) * (uint32_t)p.P;
                nN = q < nt2 ? n1 : 0;
            }

            const int slot = g * G + q;
            const bool have = slot < nt;
            const int k = have ? ts.order[ob][slot] : 0;
            const int flags = have ? ts.flags[k] : 0;
            const bool valid = flags & kFlagValid, collect = flags & kFlagCollect;
            const int N = valid ? ts.N[k] : 0;
            const int e = base + k;
            const float ux = ts.x[k], uy = ts.y[k], elapsed = ts.elapsed[k];
            const bool grp_collect = g * G < ts.ncollect[ob];  // the tile is ordered collect-first
            const bool grp_trunc = __any_sync(kFull, (flags & kFlagTruncated) != 0);
            const uint64_t gid = (uint64_t)(p.env_id_offset + e);
            const uint32_t env_lo = (uint32_t)gid, env_hi = ((uint32_t)(gid >> 32) & 0xffffffu) | (kStreamPacked << 24);
            const uint32_t episode = ts.episode[k] - 1, stepw = (uint32_t)ts.step[k] << 8;
            const uint32_t row = (uint32_t)e * (uint32_t)p.P;
            // staged observation row, in the 16-byte phase of its global row
            const uint32_t out_f = (uint32_t)e * (uint32_t)p.obs_stride + 3u;  // first sensor feature, in floats
            float* const obs_row = obsst + q * RS + (out_f & 3u);
            float* const cur_row = cur_s + q * CW;
            uint8_t* const sfa_row = sfa_s + q * CW;
            float* exp = nullptr;
            if (diag_export) {
                if (valid) {
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

            // ---- the pass: every sensor streams through registers once ----
            float loss_acc = 0.f;
            EnvAcc acc;
            acc.dist = CUDART_INF_F;
            acc.j1s = acc.j2s = 0.f;
            acc.n_has = acc.starved = 0;
            const float nw = p.rate * (collect ? p.cd : 1.0f);
            // NOT unrolled: the body is ~550 instructions; unrolled S times the kernel outgrows the instruction caches
            // (13 k instructions at S = 7 ran 35 % slower than the rolled loop)
#pragma unroll kPackUnroll
            for (int ps = 0; ps < S; ++ps) {
                const int i = ps * L + r, j0 = 2 * i;
                const bool act0 = j0 < N, act1 = j0 + 1 < N;
                cp_async_wait<S - 1>();  // entry ps of this group: requested S commits ago (one whole group of compute)
                float4 v0 = ring[ps].st0[lane], v1 = ring[ps].st1[lane];
                const float4 xy = ring[ps].pos[lane];
                const float2 gg = ring[ps].gen[lane];
                issue_slot(ps, nrow, nN);  // each lane re-arms its own place: no other lane reads it
                cp_async_commit();
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
                    if (exp && collect) {
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
                    if (exp) {
                        if (j0 < p.P) {
                            exp[UAVSIM_SLOT_D * p.P + j0] = dd0;
                            exp[UAVSIM_SLOT_E * p.P + j0] = de0;
                        }
                        if (j0 + 1 < p.P) {
                            exp[UAVSIM_SLOT_D * p.P + j0 + 1] = dd1;
                            exp[UAVSIM_SLOT_E * p.P + j0 + 1] = de1;
                        }
                    }
                    const float gen0 = fmaf(p.rate, elapsed, gg.x), gen1 = fmaf(p.rate, elapsed, gg.y);
                    const float urg0 = urgency_of(p, v0.x, v0.z, gen0), urg1 = urgency_of(p, v1.x, v1.z, gen1);
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
                    if (kDR || grp_trunc) {  // exact for the environments that do not collect (no drain follows)
                        acc_sensor<kDR>(p, acc, act0, v0.x, v0.y, gen0, xy.x, xy.y, ux, uy, grp_trunc);
                        acc_sensor<kDR>(p, acc, act1, v1.x, v1.y, gen1, xy.z, xy.w, ux, uy, grp_trunc);
                    }
                }
                v0.w = pack_sf(ema0, sf0);
                v1.w = pack_sf(ema1, sf1);
                if (act0) g_state_w[row + (uint32_t)j0] = v0;
                if (act1) g_state_w[row + (uint32_t)j0 + 1u] = v1;
            }
            // iot_sensors.py:114-125 summed over the environment (uav_env.py:457-459)
            {
                float step_loss = 0.f;
                if (__any_sync(kFull, loss_acc > 0.f)) step_loss = seg_sum<L>(loss_acc);
                if (r == 0 && valid) ts.step_loss[k] = step_loss;
            }
            __syncwarp();  // the state stores and the staged features of the group are visible to every lane

            // ---- Capture-Effect resolution: one lane per collecting environment with attempts (rare) ----
            CollectOutcome oc;
            oc.total_bytes = oc.urg_acc = oc.urg_red = 0.f;
            oc.nwin = oc.collisions = oc.captures = oc.delivered = 0;
            oc.any_new = false;
            oc.vis_lo = ts.vis_lo[k];
            oc.vis_hi = ts.vis_hi[k];
            oc.win_lo = oc.win_hi = 0u;
            uint32_t att_lo = 0u, att_hi = 0u;
            float mx = 0.f, var = 0.f;
            int n_empty = 0;
            if (grp_collect) {
                if (r == 0 && collect) {
                    att_lo = ts.att[q][0];
                    att_hi = ts.att[q][1];
                    if (att_lo | att_hi)
                        pack_resolve(p, b, row, att_lo, att_hi, cur_row, sfa_row, obs_row, ux, uy, elapsed, env_lo, env_hi, episode,
                                     stepw, exp, oc);
                }
                __syncwarp();
                // the collect reward's reductions over the FINAL state (an L2 hit: the pass has just written it)
                float bufs[2 * S], s1 = 0.f;
                if (kDR || grp_trunc) {
                    const bool c = collect;  // collecting environments recount over the drained state
                    acc.dist = c ? CUDART_INF_F : acc.dist;
                    acc.j1s = c ? 0.f : acc.j1s;
                    acc.j2s = c ? 0.f : acc.j2s;
                    acc.n_has = c ? 0 : acc.n_has;
                    acc.starved = c ? 0 : acc.starved;
                }
#pragma unroll
                for (int ps = 0; ps < S; ++ps) {
                    const int j0 = 2 * (ps * L + r);
                    const bool act0 = collect && j0 < N, act1 = collect && j0 + 1 < N;
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    if (act0) v0 = g_state[row + (uint32_t)j0];
                    if (act1) v1 = g_state[row + (uint32_t)j0 + 1u];
                    bufs[2 * ps] = v0.x;
                    bufs[2 * ps + 1] = v1.x;
                    mx = fmaxf(mx, fmaxf(v0.x, v1.x));
                    s1 += v0.x + v1.x;
                    n_empty += (int)(act0 && v0.x <= 0.f) + (int)(act1 && v1.x <= 0.f);
                    if (kDR || grp_trunc) {
                        float2 g2 = make_float2(0.f, 0.f);
                        float4 xy = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (act0) {
                            g2 = *reinterpret_cast<const float2*>(g_gen + row + (uint32_t)j0);
                            if (kDR) xy = *reinterpret_cast<const float4*>(b.spos + row + (uint32_t)j0);
                        }
                        acc_sensor<kDR>(p, acc, act0, v0.x, v0.y, fmaf(p.rate, elapsed, g2.x), xy.x, xy.y, ux, uy, grp_trunc);
                        acc_sensor<kDR>(p, acc, act1, v1.x, v1.y, fmaf(p.rate, elapsed, g2.y), xy.z, xy.w, ux, uy, grp_trunc);
                    }
                }
                mx = seg_max<L>(mx);
                s1 = seg_sum<L>(s1);
                n_empty = seg_sum<L>(n_empty);
                // calculate_starvation_penalty, reward_function.py:46-57: two passes (one Newton step on each reciprocal:
                // they scale every term)
                float inv = fast_rcp(fmaxf(mx, 1e-30f)), invn = fast_rcp((float)max(N, 1));
                inv = fmaf(inv, fmaf(-fmaxf(mx, 1e-30f), inv, 1.0f), inv);
                invn = fmaf(invn, fmaf(-(float)max(N, 1), invn, 1.0f), invn);
                const float mean = s1 * inv * invn;
                float s2 = 0.f;
#pragma unroll
                for (int ps = 0; ps < S; ++ps) {
                    const int j0 = 2 * (ps * L + r);
                    const float d0 = bufs[2 * ps] * inv - mean, d1 = bufs[2 * ps + 1] * inv - mean;
                    s2 = j0 < N ? fmaf(d0, d0, s2) : s2;
                    s2 = j0 + 1 < N ? fmaf(d1, d1, s2) : s2;
                }
                var = seg_sum<L>(s2) * invn;
            }
            if (kDR || grp_trunc) {
                acc.starved = seg_sum<L>(acc.starved);
                if constexpr (kDR) {
                    acc.dist = seg_min<L>(acc.dist);
                    acc.j1s = seg_sum<L>(acc.j1s);
                    acc.j2s = seg_sum<L>(acc.j2s);
                    acc.n_has = seg_sum<L>(acc.n_has);
                }
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
                    if (diag_detail) {
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
                if (flags & kFlagTruncated) counts |= acc.starved << 16;
                ts.counts[k] = counts;
                if constexpr (kDR) {
                    ts.cur_dist[k] = acc.dist == CUDART_INF_F ? 0.0f : acc.dist;
                    if (acc.n_has > 0) {
                        ts.jain[k] = acc.j2s > 0.f ? __fdividef(acc.j1s * acc.j1s, (float)acc.n_has * acc.j2s) : 1.0f;
                        ts.flags[k] = flags | kFlagJain;
                    }
                }
            }
            __syncwarp();

            // ---- staged observation rows -> global: head / aligned 16-byte body / tail, L lanes per row ----
            if (valid) {
                float* const out = b.obs + out_f;
                const int n3 = 3 * N;
                const int head = min((int)((4u - (out_f & 3u)) & 3u), n3);
                const int nb = (n3 - head) >> 2;
                for (int i = r; i < head; i += L) out[i] = obs_row[i];  // at most 3 floats
                const float4* s4 = reinterpret_cast<const float4*>(obs_row + head);
                float4* d4 = reinterpret_cast<float4*>(out + head);
                for (int t = r; t < nb; t += L) d4[t] = s4[t];
                const int done = head + 4 * nb;
                for (int i = done + r; i < n3; i += L) out[i] = obs_row[i];  // at most 3 floats
            }
            __syncwarp();
        }

        // ---- phase C: one environment per lane ----
        unsigned restart_mask = 0u;
        {
            bool restart = false;
            if (lane < nt && (ts.flags[lane] & kFlagValid)) {
                const int e = base + lane;
                RecRegs R;
                R.load(b.rec + e);  // an L2 hit: phase A read it a tile ago
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
                    R.u(kRVisLo) = vis_lo;
                    R.u(kRVisHi) = vis_hi;
                    R.set_d(kRTotalData, R.d(kRTotalData) + (double)total_bytes);
                    R.set_i(kRCapture, R.i32(kRCapture) + ((counts >> 8) & 0xf));
                    R.set_i(kRCollisions, R.i32(kRCollisions) + (counts & 0xff));
                    R.set_i(kRDelivered, R.i32(kRDelivered) + ((counts >> 12) & 0xf));
                }
                R.store(b.rec + e);
                restart = truncated && p.auto_reset;
                reinterpret_cast<float*>(b.reward)[e] = shaped;
                b.done[e] = truncated ? 1 : 0;
                float* o = b.obs + (uint32_t)e * (uint32_t)p.obs_stride;  // uav_env.py:644
                o[0] = __fdividef(x, (float)R.i32(kRGridW));
                o[1] = __fdividef(y, (float)R.i32(kRGridH));
                o[2] = (float)battery * p.inv_max_battery;
                if (b.action_mask && !restart) b.action_mask[e] = action_mask_of(x, y, R.i32(kRGridW), R.i32(kRGridH));
                if (diag_detail) {
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
            restart_mask = __ballot_sync(kFull, restart);
        }
        __syncwarp();

        // ---- SB3 VecEnv semantics: finished episodes restart inside the same call (rare; uavsim_step.cuh) ----
        while (restart_mask) {
            const int k = __ffs(restart_mask) - 1;
            restart_mask &= restart_mask - 1;
            load_record(ws, b.rec + base + k, lane);
            if (p.P > 32)
                finish_and_restart<float, false, 2, 3>(p, b, base + k, lane, ws);
            else
                finish_and_restart<float, false, 1, 3>(p, b, base + k, lane, ws);
            store_record(ws, b.rec + base + k, lane);
            __syncwarp();
        }
        __syncwarp();

        ob ^= 1;
        base = next_base;
        next_base = (int)__reduce_max_sync(kFull, (unsigned)pending_ticket) * kPT;
        my_action = nxt_action;
        my_n = nxt_n;
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
