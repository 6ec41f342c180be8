// The fused step: collect action, cp.async staging, tile phases, pair kernel and step_kernel.  Part of the fused
// step/reset source (see uavsim_device.cuh for the overview).
#pragma once
#include "uavsim_reset.cuh"

namespace uavsim {

// ------------------------------------------------------------------------------------------------
// step
// ------------------------------------------------------------------------------------------------
// (value, index) arg-max over the lanes of one SF group; ties -> lowest sensor index, which is what
// Python's stable sorted(..., reverse=True) yields (uav_env.py:564-566).
template <typename Real>
__device__ __forceinline__ void group_argmax(Real v, int idx, Real& vmax, int& imax) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const Real ov = __shfl_xor_sync(kFull, v, o);
        const int oi = __shfl_xor_sync(kFull, idx, o);
        if (oi >= 0 && (idx < 0 || ov > v || (ov == v && oi < idx))) {
            v = ov;
            idx = oi;
        }
    }
    vmax = v;
    imax = idx;
}

template <typename Real>
struct TileSmem;

enum : int { kGeneric = 0, kLeanPlain = 1, kLeanDomainRand = 2, kLeanRelational = 3 };

// LEAN instantiations of the step kernel compile the optional features out (DomainRandEnv shaping, the
// relational wrapper, draw export, parity diagnostics): the plain UAVEnvironment step -- the benchmarked
// configuration -- then carries neither their uniform branches nor their code in its instruction stream.
template <int LEAN, typename Real>
__device__ __forceinline__ bool feat_domain_rand(const Params<Real>& p) {
    if constexpr (LEAN == kLeanDomainRand) return true; else if constexpr (LEAN != kGeneric) return false; else return p.domain_rand != 0;
}
template <int LEAN, typename Real>
__device__ __forceinline__ bool feat_relational(const Params<Real>& p) {
    if constexpr (LEAN == kLeanRelational) return true; else if constexpr (LEAN != kGeneric) return false; else return p.relational != 0;
}
template <int LEAN, typename Real>
__device__ __forceinline__ bool feat_gps(const Params<Real>& p) {  // GPSNoisyEnv: noise added to the normalised obs[0], obs[1]
    if constexpr (LEAN) return false; else return p.gps_sigma > Real(0) && !p.gps_in_obs;
}
template <int LEAN, typename Real>
__device__ __forceinline__ bool feat_gps_obs(const Params<Real>& p) {  // PerturbedDQNEnv: noisy position inside the observation pass
    if constexpr (LEAN) return false; else return p.gps_sigma > Real(0) && p.gps_in_obs;
}
template <int LEAN>
__device__ __forceinline__ const float* feat_env_duty(const Buffers& b) {
    if constexpr (LEAN) return nullptr; else return b.env_duty;
}
template <int LEAN, typename Real>
__device__ __forceinline__ bool feat_wind(const Params<Real>& p) {
    if constexpr (LEAN) return false; else return p.wind_drift > 0.0;
}
template <int LEAN>
__device__ __forceinline__ double* feat_step_detail(const Buffers& b) {
    if constexpr (LEAN) return nullptr; else return b.step_detail;
}
template <int LEAN>
__device__ __forceinline__ void* feat_draw_export(const Buffers& b) {
    if constexpr (LEAN) return nullptr; else return b.draw_export;
}

// Draws of the step in flight for slot k of the tile, rebuilt at each point of use (cheap) so that nothing
// of it stays live across the step.
template <typename Real, bool REPLAY, int SPL, int LEAN>
__device__ __forceinline__ Draws<Real, REPLAY, SPL> step_draws(const Params<Real>& p, const Buffers& b,
                                                                const TileSmem<Real>& ts, int k, int e);

// UAVEnvironment._execute_collect_action (uav_env.py:518-632) for one environment (slot k of the tile).
// Results go to the tile block (and the parity diagnostics) from lane 0.
template <typename Real, bool REPLAY, int SPL, int LEAN>
__device__ __forceinline__ void collect_action(const Params<Real>& p, const Buffers& b, TileSmem<Real>& ts, WarpSmem& ws,
                                               int k, int e, Sensor<Real> (&s)[SPL], Real step_loss, int lane) {
    const Draws<Real, REPLAY, SPL> dr = step_draws<Real, REPLAY, SPL, LEAN>(p, b, ts, k, e);
    Hot h;
    h.N = ts.N[k];
    h.elapsed = ts.elapsed[k];
    h.vis_lo = ts.vis_lo[k];
    h.vis_hi = ts.vis_hi[k];
    constexpr bool kParity = std::is_same<Real, double>::value;
    const int N = h.N;
    // _get_sensor_urgencies (:386-394) before the collection.  Only winners change their buffer, so the
    // throughput build accumulates the reduction winner by winner; the parity build keeps the reference's
    // float32 np.sum over all sensors.
    float before[kParity ? SPL : 1];
    float urg_red_acc = 0.f;
    if constexpr (kParity) {
#pragma unroll
        for (int k = 0; k < SPL; ++k) before[k] = (float)div_const(s[k].buf, p.rate, p.inv_rate);
    }
    const Real used = (Real)p.used_hover;                                      // uav.py:263, uav_env.py:530
    const Real duty = feat_env_duty<LEAN>(b) ? (Real)(feat_env_duty<LEAN>(b)[e] / 100.0f) : p.duty_frac;
    // Phase 2: ADR update (draw A), link probability (draw B), duty-cycle gate (draw U)
    Real da[SPL], db[SPL], du[SPL];
    dr.collect(p, lane, da, db, du);
    unsigned att[SPL];
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        bool a = false;
        if constexpr (kParity) {
            if (s[k].active && s[k].buf > Real(0)) {
                s[k].cur = s[k].det + da[k];
                adr_update(p, s[k].cur, s[k].ema, s[k].sf);
                const Real rb = s[k].det + db[k];
                Real plink;
                if (rb < p.thr) {
                    plink = Real(0);
                } else {
                    const Real x = (rb - p.noise_floor) - required_snr_of<Real>(s[k].sf);
                    plink = 1.0 / (1.0 + exp(-x));
                }
                a = (plink * duty) > du[k];
            }
        } else {  // the same, straight-line: every lane computes, selects keep the state of lanes without data
            const bool has = s[k].active && s[k].buf > 0.f;
            const float cur = s[k].det + da[k];
            float ema2 = s[k].ema;
            int sf2 = s[k].sf;
            adr_update(p, cur, ema2, sf2);
            s[k].cur = cur;  // only read for attempting lanes
            s[k].ema = has ? ema2 : s[k].ema;
            s[k].sf = has ? sf2 : s[k].sf;
            const float rb = s[k].det + db[k];
            const float x = (rb - p.noise_floor) - required_snr_of<float>(sf2);
            const float plink = rb < p.thr ? 0.f : __fdividef(1.0f, 1.0f + __expf(-x));
            a = has && (plink * duty) > du[k];
        }
        s[k].attempt = a;
        att[k] = __ballot_sync(kFull, a);
    }
    unsigned long long att_mask = (unsigned long long)att[0];
    if (SPL > 1) att_mask |= (unsigned long long)att[SPL - 1] << 32;

    // Phase 3+4: SF groups in first-appearance order; k==1 wins, k>1 collides unless the strongest beats
    // the runner-up by the capture threshold; each winner is range-checked (draw C) and drained.
    Real total_bytes = Real(0), urg_acc = Real(0);
    int nwin = 0, collisions = 0, captures = 0, delivered = 0;
    bool any_new = false;
    unsigned long long win_mask = 0ull, remaining = att_mask;
    while (remaining) {
        const int i0 = __ffsll((long long)remaining) - 1;
        bool sel[SPL];
        unsigned long long gm = 0ull;
        int sf0 = 0;
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const int v = shfl(s[k].sf, i0 & 31);
            if ((i0 >> 5) == k) sf0 = v;
        }
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            sel[k] = s[k].attempt && s[k].sf == sf0;
            gm |= (unsigned long long)__ballot_sync(kFull, sel[k]) << (32 * k);
        }
        remaining &= ~gm;
        const int cnt = __popcll(gm);
        int winner = -1;
        if (cnt == 1) {
            winner = i0;
        } else {
            collisions += cnt - 1;
            Real v1 = Real(0);
            int i1 = -1;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                if (sel[k] && (i1 < 0 || s[k].cur > v1)) {
                    v1 = s[k].cur;
                    i1 = lane + 32 * k;
                }
            }
            Real top_v, sec_v;
            int top_i, sec_i;
            group_argmax(v1, i1, top_v, top_i);
            Real v2 = Real(0);
            int i2 = -1;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                const int j = lane + 32 * k;
                if (sel[k] && j != top_i && (i2 < 0 || s[k].cur > v2)) {
                    v2 = s[k].cur;
                    i2 = j;
                }
            }
            group_argmax(v2, i2, sec_v, sec_i);
            if (top_v > (sec_v + p.capture_db)) {  // uav_env.py:569
                winner = top_i;
                captures += 1;
            }
        }
        if (winner >= 0) {
            win_mask |= 1ull << winner;
            // the owner lane performs IoTSensor.collect_data (iot_sensors.py:127-145)
            Real got = Real(0), urg_w = Real(0);
            int in_range = 0;
#pragma unroll
            for (int k = 0; k < SPL; ++k) {
                if (lane + 32 * k == winner) {
                    const Real rc = s[k].det + dr.winner(p, lane, k);
                    in_range = rc >= p.thr;
                    if (in_range && s[k].buf > Real(0)) {
                        const Real cap = data_rate_of<Real>(s[k].sf) * p.cd;
                        got = s[k].buf < cap ? s[k].buf : cap;
                        s[k].buf -= got;
                        s[k].tx += got;
                    }
                    urg_w = urgency_of(p, s[k].buf, s[k].lost, generated_of(p, s[k], h.elapsed));
                }
            }
            const int owner = winner & 31;
            got = shfl(got, owner);
            if constexpr (!kParity) urg_red_acc += (float)got * p.inv_rate;  // (before - after) of the winner
            urg_w = shfl(urg_w, owner);
            in_range = shfl(in_range, owner);
            urg_acc += urg_w;  // np.mean over <= 6 winners adds left to right (uav_env.py:610-615)
            nwin += 1;
            if (in_range && got > Real(0)) {
                total_bytes += got;
                delivered += 1;
                const bool seen = winner < 32 ? (h.vis_lo >> winner) & 1u : (h.vis_hi >> (winner - 32)) & 1u;
                if (!seen) {
                    any_new = true;
                    if (winner < 32)
                        h.vis_lo |= 1u << winner;
                    else
                        h.vis_hi |= 1u << (winner - 32);
                }
            }
        }
    }
    // Phase 5: fairness metrics (uav_env.py:596-604)
    unsigned empty_b = 0u, nonempty_b = 0u;
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        empty_b |= __ballot_sync(kFull, s[k].active && s[k].buf <= Real(0));
        nonempty_b |= __ballot_sync(kFull, s[k].active && !(s[k].buf <= Real(0)));
    }
    const bool was_empty = empty_b != 0u, all_collected = nonempty_b == 0u;
    Real urg_red;
    if constexpr (kParity) {
        float diff[SPL];
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
            const float d = before[k] - (float)div_const(s[k].buf, p.rate, p.inv_rate);
            diff[k] = d > 0.f ? d : 0.f;
        }
        urg_red = (Real)sensor_sum<Real, SPL>(diff, N, lane, ws.red, kSumNumpy);  // float32 np.sum
    } else {
        urg_red = urg_red_acc;
    }
    const Real mean_urg = nwin > 0 ? fdiv(urg_acc, (Real)nwin) : Real(0);
    // Phase 6: reward_function.py:81-128, same order of additions.  The throughput build assembles it in FLOAT64
    // from its float32 inputs: the terms reach 7e4 (100 * bytes * urgency) and can cancel down to O(1), where a
    // float32 running sum would leave ~1e-4 of absolute rounding; the cost is a dozen DADD / DFMA per COLLECT step,
    // identical in every lane.
#ifdef UAVSIM_F32_REWARD  // A/B switch (tools/build_variants.py)
    using Acc = Real;
#else
    using Acc = double;
#endif
    Acc reward = (Acc)p.p_step + (Acc)p.p_hover;
    if (total_bytes > Real(0)) {
        reward += (Acc)p.r_per_byte * (Acc)total_bytes * (Acc)mean_urg;
        if (any_new) reward += (Acc)p.r_new;
    }
    if (urg_red > Real(0)) reward += (Acc)p.r_urg_red * (Acc)urg_red;
    if (was_empty && total_bytes == Real(0)) reward += (Acc)p.p_revisit;
    reward += (Acc)p.p_battery * (Acc)used;
    if (collisions > 0) reward += (Acc)p.p_collision * (Acc)collisions;
    if (step_loss > Real(0)) reward += (Acc)p.p_data_loss * (Acc)step_loss;
    if (N > 1) {  // calculate_starvation_penalty, reward_function.py:46-57
        Real mx = Real(0);
#pragma unroll
        for (int k = 0; k < SPL; ++k)
            if (s[k].active) mx = s[k].buf > mx ? s[k].buf : mx;
        mx = warp_max(mx);
        if (mx != Real(0)) {
            if constexpr (kParity) {  // np.var: mean, then mean of squared deviations, numpy summation order
                Real nrm[SPL], dev[SPL];
#pragma unroll
                for (int k = 0; k < SPL; ++k) nrm[k] = s[k].active ? s[k].buf / mx : Real(0);
                const Real mean = sensor_sum<Real, SPL>(nrm, N, lane, ws.red, kSumNumpy) / (Real)N;
#pragma unroll
                for (int k = 0; k < SPL; ++k) {
                    const Real d = nrm[k] - mean;
                    dev[k] = s[k].active ? d * d : Real(0);
                }
                reward += p.p_starvation * (sensor_sum<Real, SPL>(dev, N, lane, ws.red, kSumNumpy) / (Real)N);
            } else {
#ifdef UAVSIM_ONEPASS_VAR  // A/B switch: round-1 form, var = E[x^2] - E[x]^2 on one fused shuffle tree
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int k = 0; k < SPL; ++k) {
                    const float v = s[k].active ? s[k].buf : 0.f;
                    s1 += v;
                    s2 = fmaf(v, v, s2);
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    s1 += __shfl_xor_sync(kFull, s1, o);
                    s2 += __shfl_xor_sync(kFull, s2, o);
                }
                const float inv = fast_rcp(mx), invn = fast_rcp((float)N);
                const float mean = s1 * inv * invn;
                reward += (Acc)p.p_starvation * (Acc)fmaxf(fmaf(s2 * inv * inv, invn, -mean * mean), 0.f);
#else
                // the same two passes on shuffle trees (the one-pass E[x^2] - E[x]^2 cancels: x is in [0, 1])
                // MUFU.RCP + one Newton step: the reciprocals scale every x (and the variance twice)
                float inv = fast_rcp(mx), invn = fast_rcp((float)N);
                inv = fmaf(inv, fmaf(-mx, inv, 1.0f), inv);
                invn = fmaf(invn, fmaf(-(float)N, invn, 1.0f), invn);
                float x[SPL], s1 = 0.f;
#pragma unroll
                for (int k = 0; k < SPL; ++k) {
                    x[k] = s[k].active ? s[k].buf * inv : 0.f;
                    s1 += x[k];
                }
                const float mean = warp_sum(s1) * invn;
                float s2 = 0.f;
#pragma unroll
                for (int k = 0; k < SPL; ++k) {
                    const float d = x[k] - mean;
                    s2 = s[k].active ? fmaf(d, d, s2) : s2;
                }
                reward += (Acc)p.p_starvation * (Acc)(warp_sum(s2) * invn);
#endif
            }
        }
    }
    if (all_collected) reward += (Acc)p.r_completion;
    if (lane == 0) {
        ts.creward[k] = (Real)reward;
        ts.bytes[k] = total_bytes;
        ts.counts[k] = collisions | (captures << 8) | (delivered << 12);
        ts.nvis_lo[k] = h.vis_lo;
        ts.nvis_hi[k] = h.vis_hi;
        if (feat_step_detail<LEAN>(b)) {  // parity diagnostics: the collect step's intermediates (phase C adds the rest)
            double* d = feat_step_detail<LEAN>(b) + (size_t)e * UAVSIM_DETAIL_COUNT;
            d[UAVSIM_DETAIL_COLLISIONS] = collisions;
            d[UAVSIM_DETAIL_BYTES] = (double)total_bytes;
            d[UAVSIM_DETAIL_URGENCY_REDUCED] = (double)urg_red;
            d[UAVSIM_DETAIL_MEAN_URGENCY] = (double)mean_urg;
            d[UAVSIM_DETAIL_WAS_NEW] = any_new;
            d[UAVSIM_DETAIL_WAS_EMPTY] = was_empty;
            d[UAVSIM_DETAIL_ALL_COLLECTED] = all_collected;
            d[UAVSIM_DETAIL_ATTEMPT_MASK_LO] = (double)(uint32_t)att_mask;
            d[UAVSIM_DETAIL_ATTEMPT_MASK_HI] = (double)(uint32_t)(att_mask >> 32);
            d[UAVSIM_DETAIL_WINNER_MASK_LO] = (double)(uint32_t)win_mask;
            d[UAVSIM_DETAIL_WINNER_MASK_HI] = (double)(uint32_t)(win_mask >> 32);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// software pipeline of the throughput build: while a warp computes environment i, the columns of its
// next environment (i + total_warps) stream from HBM into a shared-memory stage with cp.async
// (LDGSTS: no registers held, no stall until the data is needed), double buffered per warp.
// ------------------------------------------------------------------------------------------------
// A stage holds one environment's sensor columns.  (Tried and removed, commit 0510d7b: TMA bulk copies onto an
// mbarrier -- three 320-800 B `cp.async.bulk` per environment keep the SM's TMA unit ~2/3 busy and were 3-12 %
// slower than per-lane cp.async -- and a shared-memory row image for coalesced observation stores, 7 % slower.)
template <typename Real, int SPL>
struct alignas(alignof(SensorState<Real>) > 16 ? alignof(SensorState<Real>) : 16) StageBuf {
    SensorState<Real> st[32 * SPL];
    float2 pos[32 * SPL];
    Real gen[32 * SPL];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

// Issue the asynchronous copies of environment e (N real sensors) into a stage buffer.
template <typename Real, int SPL>
__device__ __forceinline__ void stage_issue(const Buffers& b, const Params<Real>& p, int e, int N, int lane,
                                            StageBuf<Real, SPL>& sb) {
    static_assert(sizeof(SensorState<Real>) == 16, "the staged pipeline is built for the float32 layout");
    const uint32_t row = (uint32_t)e * (uint32_t)p.P;
    const SensorState<Real>* st = reinterpret_cast<const SensorState<Real>*>(b.sstate);
    const Real* gen = reinterpret_cast<const Real*>(b.sgen);
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const int j = lane + 32 * k;
        if (j < N) {
            const uint32_t i = row + (uint32_t)j;
            cp_async16(&sb.st[j], st + i);
            cp_async8(&sb.pos[j], b.spos + i);
            cp_async4(&sb.gen[j], gen + i);
        }
    }
}

template <typename Real, int SPL>
__device__ __forceinline__ void load_sensors_staged(const StageBuf<Real, SPL>& sb, int N, int lane, Sensor<Real> (&s)[SPL]) {
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
        const int j = lane + 32 * k;
        s[k].active = j < N;
        s[k].attempt = false;
        s[k].det = Real(0);
        s[k].cur = Real(0);
        // Unconditional loads (the stage has 32 * SPL slots): a lane without a sensor reads whatever an earlier
        // environment left there and runs the same arithmetic on it -- every reduction, ballot and store is
        // masked by `active`, so the values never escape.
        const SensorState<Real> v = sb.st[j];
        const float2 xy = sb.pos[j];
        const Real g = sb.gen[j];
        float ema;
        int sf;
        unpack_sf(v.ema, ema, sf);
        if (!s[k].active) sf = 12;
        s[k].buf = v.buf;
        s[k].tx = v.tx;
        s[k].lost = v.lost;
        s[k].ema = ema;
        s[k].gen = g;
        s[k].sx = xy.x;
        s[k].sy = xy.y;
        s[k].sf = sf;
    }
}

// Episode end inside a step (SB3 VecEnv semantics): terminal observation, episode statistics, new episode.
// Deliberately NOT inlined and self-contained -- it re-reads the sensor columns the step has just stored
// -- so that the rare path costs the hot loop neither registers nor local memory.
template <typename Real, bool REPLAY, int SPL, int FPS>
__device__ __noinline__ void finish_and_restart(const Params<Real>& p, const Buffers& b, int e, int lane, WarpSmem& ws) {
    __syncwarp();
    const int N = ws.rec.num_sensors;
    if (b.terminal_obs) {  // the k stacked frames ending with the terminal one (written by this warp a moment ago)
        const int total = p.frame_window * p.obs_dim;
        float* t = b.terminal_obs + (size_t)e * total;
        const float* win = b.obs + (size_t)e * p.obs_stride + (p.frame_k > 1 ? (b.frame_slot + 1) * p.obs_dim : 0);
        for (int i = lane; i < total; i += 32) t[i] = win[i];
    }
    Sensor<Real> s[SPL];
    load_sensors<Real, SPL>(b, p, e, N, lane, s);
    write_episode_stats<Real, SPL>(p, b, e, lane, ws.rec, s);
#ifndef UAVSIM_NO_TINFO  // A/B switch
    if (b.terminal_info || b.terminal_ratios)
        write_info<Real, SPL>(p, lane, ws.rec, s, ws.red, b.terminal_info ? b.terminal_info + (size_t)e * UAVSIM_INFO_COUNT : nullptr,
                              b.terminal_ratios ? b.terminal_ratios + (size_t)e * p.P : nullptr);
#endif
    reset_env<Real, REPLAY, SPL, FPS>(p, b, e, lane, ws, s, (Real*)nullptr);
    store_sensors<Real, SPL>(b, p, e, lane, s, true);
}

// ------------------------------------------------------------------------------------------------
// the step kernel: tiles of kTile consecutive environments per warp
// ------------------------------------------------------------------------------------------------
// Per-environment SCALAR work (action decode, UAV move + battery, edge bookkeeping, truncation, reward
// assembly, terminal penalties, shaping, record update, reward/done/mask outputs) is identical for every
// lane when a warp owns one environment -- 32 lanes would issue it to produce one result.  So a warp
// takes a TILE of kTile environments and runs that work one-environment-per-LANE (phases A and C),
// while the per-sensor work stays one-environment-per-WARP with sensors across lanes (phase B, a loop
// over the tile).  The phases talk through a small structure-of-arrays block in shared memory; the tile's
// 128-byte records sit in shared memory with a 4-byte XOR swizzle so that both the coalesced global
// copies (lane = word) and the per-lane field accesses (lane = environment) are bank-conflict free.
#ifndef UAVSIM_TILE
#define UAVSIM_TILE 16
#endif
constexpr int kTile = UAVSIM_TILE;

template <typename Real>
struct alignas(16) TileSmem {
    uint32_t rec[2][kTile][32];  // [buffer][l][w ^ l]: word w of record l; the other buffer receives the next tile
    int32_t action[2][kTile];
    // phase A -> phase B (read as warp-wide broadcasts) and -> phase C
    float x[kTile], y[kTile], elapsed[kTile];
    int32_t N[kTile], step[kTile], flags[kTile];  // (grid W, H are read from the record tile where needed)
    uint32_t episode[kTile], vis_lo[kTile], vis_hi[kTile];
    double battery[kTile];
    Real used[kTile];
    // phase B -> phase C
    Real step_loss[kTile], creward[kTile], bytes[kTile], jain[kTile];
    float cur_dist[kTile];
    int32_t counts[kTile];  // collisions | captures << 8 | delivered << 12 | starved << 16
    uint32_t nvis_lo[kTile], nvis_hi[kTile];
};

// Dynamic shared memory of one block of the step kernel: per warp a working record + reduction scratch,
// the tile block and (float32 layout) two stage buffers.
template <typename Real, int SPL, bool PAIR = false>
struct alignas(alignof(SensorState<Real>) > 16 ? alignof(SensorState<Real>) : 16) StepWarpSmem {
    WarpSmem ws;
    TileSmem<Real> ts;
    StageBuf<Real, SPL> stage[std::is_same<Real, double>::value ? 0 : (PAIR ? 4 : 2)];  // pair kernel: 2 stages x 2 envs
};
template <typename Real, int SPL, bool PAIR = false>
constexpr size_t step_smem_bytes() {
    return sizeof(StepWarpSmem<Real, SPL, PAIR>) * step_warps<Real, SPL, PAIR>();
}

enum : int {
    kFlagValid = 1, kFlagCollect = 2, kFlagTruncated = 4, kFlagMoveOk = 8, kFlagEdge = 16, kFlagJain = 32
};

// record word offsets (uavsim_env_record, include/uavsim.h)
enum : int {
    kRX = 0, kRY = 1, kRStartX = 2, kRStartY = 3, kRBattery = 4, kRTotalReward = 6, kRTotalData = 8, kREpReturn = 10,
    kRLastBytes = 12, kRPrevDist = 13, kRElapsed = 14, kRStep = 15, kRGridW = 16, kRGridH = 17, kRN = 18, kRVisLo = 19,
    kRVisHi = 20, kRCapture = 21, kRBoundary = 22, kREdge = 23, kRCollisions = 24, kRFirstCov = 25, kREpisode = 26,
    kRDelivered = 27, kRInitSf = 28, kRDwell = 29, kRWindAngle = 30
};
static_assert(offsetof(uavsim_env_record, battery) == 4 * kRBattery && offsetof(uavsim_env_record, step) == 4 * kRStep &&
                  offsetof(uavsim_env_record, num_sensors) == 4 * kRN && offsetof(uavsim_env_record, episode) == 4 * kREpisode &&
                  offsetof(uavsim_env_record, delivered) == 4 * kRDelivered && offsetof(uavsim_env_record, dwell_steps) == 4 * kRDwell && offsetof(uavsim_env_record, wind_angle) == 4 * kRWindAngle && offsetof(uavsim_env_record, elapsed_s) == 4 * kRElapsed &&
                  offsetof(uavsim_env_record, total_reward) == 4 * kRTotalReward && offsetof(uavsim_env_record, episode_return) == 4 * kREpReturn &&
                  offsetof(uavsim_env_record, prev_dist_nearest) == 4 * kRPrevDist && offsetof(uavsim_env_record, first_full_coverage_step) == 4 * kRFirstCov,
              "record word map out of date");

struct RecView {  // lane-private access to record l of the swizzled tile
    uint32_t* row;
    int l;
    __device__ __forceinline__ uint32_t& w(int i) const { return row[i ^ l]; }
    __device__ __forceinline__ float f(int i) const { return __uint_as_float(row[i ^ l]); }
    __device__ __forceinline__ int i32(int i) const { return (int)row[i ^ l]; }
    __device__ __forceinline__ double d(int i) const { return __hiloint2double((int)row[(i + 1) ^ l], (int)row[i ^ l]); }
    __device__ __forceinline__ void set_f(int i, float v) const { row[i ^ l] = __float_as_uint(v); }
    __device__ __forceinline__ void set_i(int i, int v) const { row[i ^ l] = (uint32_t)v; }
    __device__ __forceinline__ void set_d(int i, double v) const {
        row[i ^ l] = (uint32_t)__double2loint(v);
        row[(i + 1) ^ l] = (uint32_t)__double2hiint(v);
    }
};

template <typename Real, bool REPLAY, int SPL, int LEAN>
__device__ __forceinline__ Draws<Real, REPLAY, SPL> step_draws(const Params<Real>& p, const Buffers& b,
                                                                const TileSmem<Real>& ts, int k, int e) {
    const uint64_t gid = (uint64_t)(p.env_id_offset + e);
    Draws<Real, REPLAY, SPL> dr;
    dr.env_lo = (uint32_t)gid;
    dr.env_hi = (uint32_t)(gid >> 32) & 0xffffffu;
    dr.episode = ts.episode[k] - 1;  // the record counts resets done; the running episode is one less
    dr.step = (uint32_t)ts.step[k];
    // pair time = the even environment of the pair (tiles start at multiples of 16 and env_id_offset is even when
    // P <= 32, so it is slot k & ~1 of this tile; phase A publishes episode / step of every slot, valid or not)
    dr.paired = SPL == 1;
    dr.pair_episode = ts.episode[k & ~1] - 1;
    dr.pair_step = (uint32_t)ts.step[k & ~1];
    if constexpr (LEAN)
        dr.sigma = p.sigma;
    else
        dr.sigma = b.env_sigma ? (Real)b.env_sigma[e] : p.sigma;
    dr.tab = REPLAY ? b.step_draws + (size_t)e * UAVSIM_NUM_SLOTS * p.P : nullptr;
    dr.exp = (!REPLAY && feat_draw_export<LEAN>(b))
                 ? reinterpret_cast<Real*>(feat_draw_export<LEAN>(b)) + (size_t)e * UAVSIM_NUM_SLOTS * p.P
                 : nullptr;
    return dr;
}

// Asynchronous copy of the records and actions of the tile starting at `base` into buffer `tb` of the tile block
// (float32 layout: issued during phase B of the previous tile so that a tile never waits for its records).
template <typename Real>
__device__ __forceinline__ void tile_issue(const Buffers& b, const int32_t* actions, int base, int tb, int lane,
                                           TileSmem<Real>& ts) {
    const int n = min(b.tile, b.E - base);
#pragma unroll
    for (int r = 0; r < kTile; ++r)
        if (r < n) cp_async4(&ts.rec[tb][r][lane ^ r], reinterpret_cast<const uint32_t*>(b.rec + base + r) + lane);
    if (lane < n) cp_async4(&ts.action[tb][lane], actions + base + lane);
}

// IoTSensor.step (iot_sensors.py:114-125) for one sensor: the buffer grows by `nw`, is capped at max_buffer_size, and the
// excess is lost.  Returns the excess; `over` = the cap was hit.
template <typename Real>
__device__ __forceinline__ Real age_buffer(const Params<Real>& p, Sensor<Real>& s, Real nw, bool& over) {
    const Real pot = s.buf + nw;
    Real loss;
    if constexpr (std::is_same<Real, double>::value) {
        over = pot > p.maxbuf;
        loss = over ? pot - p.maxbuf : Real(0);
        s.buf = over ? p.maxbuf : pot;
    } else {  // min / max instead of compare + selects (pot - max is exact where it matters: pot ~ max)
        loss = fmaxf(pot - p.maxbuf, 0.f);
        s.buf = fminf(pot, p.maxbuf);
        over = loss > 0.f;
    }
    s.lost += loss;  // + 0 when not saturated: no effect on the stored value
    return loss;
}

// DomainRandEnv.step shaping inputs of one environment (dqn.py:415-444): distance to the nearest sensor with data and
// Jain's index of the per-sensor collection rates, left in the tile block for phase C.
template <typename Real, int SPL>
__device__ __forceinline__ void domain_rand_inputs(const Params<Real>& p, TileSmem<Real>& ts, WarpSmem& ws, int k, int N,
                                                   int flags, int lane, Sensor<Real> (&s)[SPL]) {
    const float cur_dist = nearest_with_data(ts.x[k], ts.y[k], s);
    const float elapsed = ts.elapsed[k];
    Real r1[SPL], r2[SPL];
    int cnt = 0;
#pragma unroll
    for (int k2 = 0; k2 < SPL; ++k2) {
        const Real g = generated_of(p, s[k2], elapsed);
        const bool has = s[k2].active && g > Real(0);
        const Real rate = has ? fdiv(s[k2].tx, g) * Real(100) : Real(0);
        r1[k2] = rate;
        r2[k2] = rate * rate;
        cnt += __popc(__ballot_sync(kFull, has));
    }
    Real jain = Real(0);
    if (cnt > 0) {
        const Real s1 = sensor_sum<Real, SPL>(r1, N, lane, ws.red, kSumSequential);
        const Real s2 = sensor_sum<Real, SPL>(r2, N, lane, ws.red, kSumSequential);
        jain = s2 > Real(0) ? fdiv(s1 * s1, (Real)cnt * s2) : Real(1);
    }
    if (lane == 0) {
        ts.cur_dist[k] = cur_dist;
        ts.jain[k] = jain;
        if (cnt > 0) ts.flags[k] = flags | kFlagJain;
    }
}

// Phase B for ONE environment (slot k of the tile): everything that touches the sensor columns.  Scalars of
// the environment are read from the tile block where they are used and results are written back as soon as
// they exist, so that little besides the sensor registers is live across the phase.
template <typename Real, bool REPLAY, int SPL, int FPS, bool STACK, int LEAN>
__device__ __forceinline__ void sensor_phase(const Params<Real>& p, const Buffers& b, TileSmem<Real>& ts, WarpSmem& ws,
                                             int tb, int k, int e, int lane, Sensor<Real> (&s)[SPL]) {
    constexpr bool kParity = std::is_same<Real, double>::value;
    const int flags = ts.flags[k];
    const bool collect = flags & kFlagCollect;
    const int N = ts.N[k];

    if (!REPLAY && feat_draw_export<LEAN>(b)) {
        Real* exp = reinterpret_cast<Real*>(feat_draw_export<LEAN>(b)) + (size_t)e * UAVSIM_NUM_SLOTS * p.P;
        const Real nan = Real(CUDART_NAN);
        for (int i = lane; i < UAVSIM_NUM_SLOTS * p.P; i += 32) exp[i] = nan;
        __syncwarp();
    }

    // ---- iot_sensors.py:114-125: age the buffers ----
    Real step_loss = Real(0);
    {
        Real loss[SPL];
        const Real nw = p.rate * (collect ? p.cd : Real(1));
        bool any_over = false;
#pragma unroll
        for (int k2 = 0; k2 < SPL; ++k2) {
            if constexpr (kParity) s[k2].gen += nw;
            bool over;
            loss[k2] = age_buffer(p, s[k2], nw, over);
            any_over |= (s[k2].active && over);
        }
        if (__any_sync(kFull, any_over)) step_loss = sensor_sum<Real, SPL>(loss, N, lane, ws.red, kSumSequential);
        if (lane == 0) ts.step_loss[k] = step_loss;
    }

    compute_det(p, ts.x[k], ts.y[k], s);  // the UAV has already moved (phase A)
    if (collect) collect_action<Real, REPLAY, SPL, LEAN>(p, b, ts, ws, k, e, s, step_loss, lane);

    if (flags & kFlagTruncated) {  // reward_function.py:59-67 (counted only when the episode ends this step)
        int starved = 0;
        const float elapsed = ts.elapsed[k];
#pragma unroll
        for (int k2 = 0; k2 < SPL; ++k2) {
            const Real g = generated_of(p, s[k2], elapsed);
            starved += __popc(__ballot_sync(kFull, s[k2].active && g > Real(0) && fdiv(s[k2].tx, g) < p.cr_thr));
        }
        if (lane == 0) ts.counts[k] |= starved << 16;
    }

    // ---- observation, sensor part (uav_env.py:638-674); obs[0..2] are written by phase C ----
    {
        Hot h;
        h.x = ts.x[k];
        h.y = ts.y[k];
        h.elapsed = ts.elapsed[k];
        h.W = (int)ts.rec[tb][k][kRGridW ^ k];
        h.H = (int)ts.rec[tb][k][kRGridH ^ k];
        h.N = N;
        h.battery = 0.0;
        h.vis_lo = collect ? ts.nvis_lo[k] : ts.vis_lo[k];  // relational rows carry the visited flag
        h.vis_hi = collect ? ts.nvis_hi[k] : ts.vis_hi[k];
        if (feat_gps_obs<LEAN>(p)) {  // PerturbedDQNEnv._get_observation: the pass runs at the noisy position
            double gx, gy;
            env_draws_pos<Real, REPLAY>(p, b.step_env_draws, e, ts.episode[k] - 1, (uint32_t)ts.step[k], gx, gy);
            const NoisyPos np_ = noisy_position<Real>(h.x, h.y, h.W, h.H, gx, gy);
            h.x = np_.x;
            h.y = np_.y;
            compute_det(p, h.x, h.y, s);
            if (lane == 0) {  // phase C writes obs[0], obs[1]: handed over in two fields only DomainRandEnv uses
                ts.cur_dist[k] = np_.ox;
                ts.jain[k] = (Real)np_.oy;
            }
        }
        observe<Real, REPLAY, SPL, FPS, false, STACK>(p, h, s, step_draws<Real, REPLAY, SPL, LEAN>(p, b, ts, k, e), lane,
                                                       obs_frame(p, b, e));
    }

    // ---- RelationalUAVEnv.step supplements (env_wrapper.py:169-198): potential distance, dwell probes ----
    if (FPS == 5 && feat_relational<LEAN>(p)) {
        const uint32_t vlo = collect ? ts.nvis_lo[k] : ts.vis_lo[k], vhi = collect ? ts.nvis_hi[k] : ts.vis_hi[k];
        const float dist = nearest_unvisited(ts.x[k], ts.y[k], vlo, vhi, lane, s);
        int n_in = 0;
        if (collect) {
            Real f[SPL];
            step_draws<Real, REPLAY, SPL, LEAN>(p, b, ts, k, e).dwell(p, b.dwell_draws ? b.dwell_draws + (size_t)e * p.P : nullptr, lane, f);
#pragma unroll
            for (int k2 = 0; k2 < SPL; ++k2)
                n_in += __popc(__ballot_sync(kFull, s[k2].active && (s[k2].det + f[k2]) >= p.thr));
        }
        if (lane == 0) {
            ts.cur_dist[k] = dist;
            ts.counts[k] |= n_in << 24;
        }
    }

    if (feat_domain_rand<LEAN>(p)) domain_rand_inputs<Real, SPL>(p, ts, ws, k, N, flags, lane, s);
    store_sensors<Real, SPL>(b, p, e, lane, s, false);
}

// ------------------------------------------------------------------------------------------------
// P <= 32, plain environment (LEAN): TWO environments per warp pass.  With one sensor slot per lane a pass over
// one environment leaves 12 of 32 lanes idle at N = 20 and uses half of each Philox call; here slot q of a lane
// is sensor `lane` of environment k0 + q, so the per-sensor arithmetic of the two environments interleaves like
// the two slots of the P > 32 kernel, the observation draws of both come from ONE Philox call (the pair-keyed
// counter of Draws), and the per-pass loop control is paid once per pair.  Everything per environment -- collect
// action, reductions, stores -- runs the one-slot building blocks on a one-element view of s[q].
// ------------------------------------------------------------------------------------------------
template <bool STACK, int LEAN>
__device__ __forceinline__ void pair_phase(const Params<float>& p, const Buffers& b, TileSmem<float>& ts, WarpSmem& ws,
                                           int k0, int base, int lane, const StageBuf<float, 1>* st2, bool have_b) {
    using View = Sensor<float>(&)[1];
    Sensor<float> s[2];
    int flags[2], N[2];
    bool valid[2], collect[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int k = k0 + q;
        flags[q] = (q == 0 || have_b) ? ts.flags[k] : 0;
        valid[q] = flags[q] & kFlagValid;
        collect[q] = flags[q] & kFlagCollect;
        N[q] = valid[q] ? ts.N[k] : 0;  // an absent / rejected environment has no active lane: nothing is reduced or stored
        load_sensors_staged<float, 1>(st2[q], N[q], lane, reinterpret_cast<View>(s[q]));
    }
    __syncwarp();
    if (!(valid[0] || valid[1])) return;

    // ---- iot_sensors.py:114-125: age the buffers ----
    float step_loss[2];
    {
        float loss[2];
        bool any_over[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float nw = p.rate * (collect[q] ? p.cd : 1.0f);
            bool over;
            loss[q] = age_buffer(p, s[q], nw, over);
            any_over[q] = s[q].active && over;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            step_loss[q] = 0.f;
            if (__any_sync(kFull, any_over[q])) step_loss[q] = warp_sum(s[q].active ? loss[q] : 0.f);
            if (lane == 0 && valid[q]) ts.step_loss[k0 + q] = step_loss[q];
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) s[q].det = det_rssi(p, ts.x[k0 + q], ts.y[k0 + q], s[q].sx, s[q].sy);
#ifndef UAVSIM_PAIR_FOLD  // the collect path inlined once per slot; -DUAVSIM_PAIR_FOLD: one copy (3 % slower, measured)
#pragma unroll
    for (int q = 0; q < 2; ++q)
        if (valid[q] && collect[q])
            collect_action<float, false, 1, LEAN>(p, b, ts, ws, k0 + q, base + k0 + q, reinterpret_cast<View>(s[q]),
                                                  step_loss[q], lane);
#else
    {   // A/B variant: ONE copy of the collect path for both slots, the collecting slot selected into a temporary (at most
        // two rounds).  840 SASS instructions shorter, but 0.426 vs 0.413 ms per launch at N = 20: rejected.
        const bool c0 = valid[0] && collect[0], c1 = valid[1] && collect[1];
        if (c0 || c1) {
            int q = c0 ? 0 : 1;
            for (;;) {
                Sensor<float> t = q ? s[1] : s[0];
                collect_action<float, false, 1, LEAN>(p, b, ts, ws, k0 + q, base + k0 + q, reinterpret_cast<View>(t),
                                                      q ? step_loss[1] : step_loss[0], lane);
                if (q)
                    s[1] = t;
                else
                    s[0] = t;
                if (q == 0 && c1)
                    q = 1;
                else
                    break;
            }
        }
    }
#endif
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (flags[q] & kFlagTruncated) {  // reward_function.py:59-67
            const float g = generated_of(p, s[q], ts.elapsed[k0 + q]);
            const int starved = __popc(__ballot_sync(kFull, s[q].active && g > 0.f && fdiv(s[q].tx, g) < p.cr_thr));
            if (lane == 0) ts.counts[k0 + q] |= starved << 16;
        }
    }
    // ---- observation, sensor part: one Philox call for the pair (uav_env.py:638-674) ----
    {
        const Draws<float, false, 1> dr = step_draws<float, false, 1, LEAN>(p, b, ts, k0, base + k0);
        const Philox4 qp = dr.pair_call(p, lane);
        float dd[2][1], de[2][1];
        box_muller<float>(qp.x, qp.y, p.sigma, dd[0][0], de[0][0]);
        box_muller<float>(qp.z, qp.w, p.sigma, dd[1][0], de[1][0]);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            Hot h;
            h.x = h.y = 0.f;
            h.elapsed = ts.elapsed[k0 + q];
            h.W = h.H = 1;
            h.N = N[q];
            h.battery = 0.0;
            h.vis_lo = h.vis_hi = 0u;
            observe_with<float, 1, 3, false, STACK>(p, h, reinterpret_cast<View>(s[q]), dd[q], de[q], lane,
                                                    obs_frame(p, b, base + k0 + q));
        }
    }
    if constexpr (LEAN == kLeanDomainRand) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (valid[q]) domain_rand_inputs<float, 1>(p, ts, ws, k0 + q, N[q], flags[q], lane, reinterpret_cast<View>(s[q]));
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) store_sensors<float, 1>(b, p, base + k0 + q, lane, reinterpret_cast<View>(s[q]), false);
}

// One launch steps every environment of the handle once.  Persistent warps: the grid covers the SMs at
// full residency and each warp walks tiles t = warp_id, warp_id + total_warps, ...
template <typename Real, bool REPLAY, int SPL, int FPS, bool STACK, int LEAN = 0, bool PAIR = false>
__global__ void __launch_bounds__(32 * step_warps<Real, SPL, PAIR>(), step_min_blocks<Real, SPL, PAIR>())
    step_kernel(const __grid_constant__ Params<Real> p, const __grid_constant__ Buffers b,
                const int32_t* __restrict__ actions) {
    constexpr bool kParity = std::is_same<Real, double>::value;
    constexpr bool kStaged = !kParity;  // the float32 layout runs the cp.async software pipeline
    // LEAN instantiations restart finished episodes OUTSIDE the tile loop: a tile that ends episodes leaves the loop
    // with the mask of its finished environments, they are restarted (out-of-line call), and the loop is re-entered.
    // The restart path -- episode statistics, terminal info, DomainRandEnv scenario draws with the far-start search,
    // first observation -- almost never runs, but a call to it INSIDE the hot loop costs the loop registers at the call
    // site: adding one more callee to it (the terminal-info row) took the N = 50 step from 0.65 to 0.74 ms per launch
    // (profiles/README.md).  (Round 2 first moved it to a second launch: 30-38 us per step on the critical path.)
    constexpr bool kDeferredRestart = LEAN != kGeneric;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    static_assert(!PAIR || ((LEAN == kLeanPlain || LEAN == kLeanDomainRand) && SPL == 1 && FPS == 3 && !REPLAY && !std::is_same<Real, double>::value),
                  "the pair kernel is the plain throughput step with one sensor slot per lane");
    StepWarpSmem<Real, SPL, PAIR>& mine = reinterpret_cast<StepWarpSmem<Real, SPL, PAIR>*>(dyn_smem)[warp];
    WarpSmem& ws = mine.ws;
    TileSmem<Real>& ts = mine.ts;
    StageBuf<Real, SPL>* stage = mine.stage;
    // loop-invariant scalars are re-derived from the constant bank where they are used instead of being
    // carried in registers: with 184 KB of the SM given to shared memory the L1 is small and a spilled
    // loop-carried value costs an L2 round trip per environment
#define UAVSIM_E (b.E)
    // Tiles are handed out by TICKET (one atomicAdd per tile per warp, requested a whole tile ahead of its use):
    // warps that drew cheap tiles (few collect actions, few restarts) take more of them, so the launch does not end
    // on the slowest statically assigned warp, and small batches spread over the whole machine with a smaller tile.
    // The block that finishes last zeroes the counters for the next launch (end of the kernel), so a launch carries no
    // per-step host state besides the frame slot: with one frame per environment it can be captured in a CUDA graph
    // and replayed (tests/test_gpu_api.py).
    unsigned* const ticket_ctr = b.tile_ticket;
#define UAVSIM_TAKE_TICKET() ((lane == 0) ? (int)atomicAdd(ticket_ctr, 1u) : 0)
    int base, next_base;
    {
        const int t0 = (lane == 0) ? (int)atomicAdd(ticket_ctr, 2u) : 0;  // the first two tiles in one round trip
        const int t1 = (lane == 0) ? t0 + 1 : 0;
        // redux.sync (REDUX) delivers its result in a UNIFORM register: the compiler then knows that base, the
        // environment index, the row offsets and the loop counters derived from them are warp-uniform, keeps them
        // off the vector register file and turns the loop-control branches into uniform branches
        base = (int)__reduce_max_sync(kFull, (unsigned)t0) * b.tile;
        next_base = (int)__reduce_max_sync(kFull, (unsigned)t1) * b.tile;
    }
#define UAVSIM_NEXT_BASE next_base
    int tb = 0;  // tile buffer in use
    if constexpr (kStaged) {  // prime both pipelines: records of the first tile, sensor columns of its first env
        if (base < UAVSIM_E) tile_issue<Real>(b, actions, base, 0, lane, ts);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        if constexpr (!PAIR) {  // the first two environments of the warp's sequence, one stage and one group each
            if (base < UAVSIM_E) stage_issue<Real, SPL>(b, p, base, (int)ts.rec[0][0][kRN], lane, stage[0]);
            cp_async_commit();
            if (base + 1 < UAVSIM_E) stage_issue<Real, SPL>(b, p, base + 1, (int)ts.rec[0][1][kRN ^ 1], lane, stage[1]);
            cp_async_commit();
        }
        if constexpr (PAIR) {  // the first pair of the warp's sequence (sensor counts from the records just landed)
            if (base < UAVSIM_E) stage_issue<Real, SPL>(b, p, base, (int)ts.rec[0][0][kRN], lane, stage[0]);
            if (base + 1 < UAVSIM_E) stage_issue<Real, SPL>(b, p, base + 1, (int)ts.rec[0][1][kRN ^ 1], lane, stage[1]);
            cp_async_commit();
        }
    }

    unsigned pend_mask = 0u;  // finished environments of the tile that has just left the loop (kDeferredRestart)
    int pend_base = 0;
  for (;;) {
    while (base < UAVSIM_E && pend_mask == 0u) {
        const int pending_ticket = UAVSIM_TAKE_TICKET();  // the tile after the next one; consumed at the end of this tile
        const int ntile = min(b.tile, UAVSIM_E - base);
        uint32_t(*rec)[32] = ts.rec[tb];
        if constexpr (!kStaged) {
            // ---- tile records: global -> swizzled shared memory (lane = word: coalesced 128-byte rows) ----
            __syncwarp();
#pragma unroll
            for (int r = 0; r < kTile; ++r)
                if (r < ntile) rec[r][lane ^ r] = reinterpret_cast<const uint32_t*>(b.rec + base + r)[lane];
            if (lane < ntile) ts.action[tb][lane] = actions[base + lane];
            __syncwarp();
        }

        // ---- phase A: one environment per lane ----
        const RecView R{rec[lane & (kTile - 1)], lane & (kTile - 1)};
        bool my_valid = false;
        int my_n = 0;
        if (lane < ntile) {
            const int e = base + lane;
            int my_flags = 0;
            const int my_action = ts.action[tb][lane];
            ts.step[lane] = R.i32(kRStep) + 1;     // published for every slot: the even slot is its pair's clock
            ts.episode[lane] = R.w(kREpisode);
            bool masked_out = false;
            if constexpr (!LEAN) masked_out = b.step_mask && !b.step_mask[e];  // uavsim_step_ex: this env sits the launch out
            if (masked_out) {
            } else if (my_action < 0 || my_action > 4) {  // uav_env.py:468 raises ValueError; here: flag it, leave the env untouched
                atomicMin(b.error_word, ((unsigned long long)b.step_serial << 32) | (unsigned)e);
            } else {
                my_flags = kFlagValid;
                float x = R.f(kRX), y = R.f(kRY);
                const int W = R.i32(kRGridW), H = R.i32(kRGridH);
                double battery = R.d(kRBattery);
                const int step = R.i32(kRStep) + 1;  // uav_env.py:439
                // uav_env.py:443-447: edge-cell bookkeeping on the PRE-move position
                if constexpr (kParity) {
                    const double Wd = W, Hd = H, ux = x, uy = y, eps = 1e-6;
                    if (ux <= eps || uy <= eps || ux >= Wd - 1 - eps || uy >= Hd - 1 - eps) my_flags |= kFlagEdge;
                } else {
                    const float eps = 1e-6f;
                    if (x <= eps || y <= eps || x >= (float)(W - 1) - eps || y >= (float)(H - 1) - eps) my_flags |= kFlagEdge;
                }
                const bool collect = my_action == 4;
                Real used;
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
                    used = (Real)(before - battery);
                } else {  // uav.py:187-206, :260-263
                    my_flags |= kFlagCollect;
                    battery -= p.e_hover;
                    used = (Real)p.used_hover;
                }
                if (!(battery > p.batt_floor) || step >= p.max_steps) my_flags |= kFlagTruncated;  // uav.py:224, uav_env.py:474-478
                ts.x[lane] = x;
                ts.y[lane] = y;
                ts.elapsed[lane] = R.f(kRElapsed) + (float)(collect ? p.cd : Real(1));
                ts.N[lane] = R.i32(kRN);
                ts.vis_lo[lane] = R.w(kRVisLo);
                ts.vis_hi[lane] = R.w(kRVisHi);
                ts.battery[lane] = battery;
                ts.used[lane] = used;
            }
            ts.flags[lane] = my_flags;
            ts.counts[lane] = 0;
            my_valid = my_flags & kFlagValid;
            my_n = R.i32(kRN);
        }
        // one bit per environment of the tile / its sensor count in the owning lane: phase B takes them from
        // registers (a shift, a shuffle) instead of re-reading shared memory per environment
        const unsigned valid_mask = __ballot_sync(kFull, my_valid);
        __syncwarp();

        // ---- phase B (pair kernel): two environments per warp pass ----
        if constexpr (PAIR) {
            // Stages: pair buffer pb holds the current pair (2 environments), pb ^ 1 receives the next pair.  cp.async
            // groups in commit order: G(j) per pair iteration j, the next tile's records + actions T right after G(0);
            // T is read from the last pair iteration on.
            for (int k = 0; k < ntile; k += 2) {
                const int pb = (k >> 1) & 1;
                int e1 = base + k + 2, n1a = 0, n1b = 0;
                bool ha = false, hb = false;
                if (k + 2 < ntile) {
                    n1a = (int)rec[k + 2][kRN ^ (k + 2)];
                    ha = true;
                    if (k + 3 < ntile) {
                        n1b = (int)rec[k + 3][kRN ^ (k + 3)];
                        hb = true;
                    }
                } else {
                    e1 = UAVSIM_NEXT_BASE;
                    cp_async_wait<0>();  // the next tile's records (T) must have landed before they are read; with
                    __syncwarp();        // the 16-tile they did long ago, with a short tile this is the only guarantee
                    if (e1 < UAVSIM_E) {
                        n1a = (int)ts.rec[tb ^ 1][0][kRN];
                        ha = true;
                        if (e1 + 1 < UAVSIM_E) {
                            n1b = (int)ts.rec[tb ^ 1][1][kRN ^ 1];
                            hb = true;
                        }
                    }
                }
                if (ha) stage_issue<Real, SPL>(b, p, e1, n1a, lane, stage[(pb ^ 1) * 2]);
                if (hb) stage_issue<Real, SPL>(b, p, e1 + 1, n1b, lane, stage[(pb ^ 1) * 2 + 1]);
                cp_async_commit();
                if (k == 0) {
                    if (UAVSIM_NEXT_BASE < UAVSIM_E)
                        tile_issue<Real>(b, actions, UAVSIM_NEXT_BASE, tb ^ 1, lane, ts);
                    cp_async_commit();
                }
                if (k <= 2)
                    cp_async_wait<2>();  // G(j) and T may still be in flight
                else
                    cp_async_wait<1>();  // only G(j): T has landed before the last iteration peeks at it
                __syncwarp();
                if constexpr (PAIR) pair_phase<STACK, LEAN>(p, b, ts, ws, k, base, lane, &stage[pb * 2], k + 1 < ntile);
            }
        }
        // ---- phase B: one environment per warp, sensors across lanes ----
        for (int k = 0; !PAIR && k < ntile; ++k) {
            const int e = base + k;
            Sensor<Real> s[SPL];
            if constexpr (kStaged) {
                // Two stages, prefetch distance TWO environments: wait for this environment's stage (requested two
                // iterations ago), move it to registers, and reuse the SAME stage at once for the environment two
                // places down the warp's sequence, so a request has two full iterations to land.  Stages alternate
                // with k: tiles hold an even number of environments (a partial tile can only be the last one).
                // cp.async groups in commit order: G(k) = environment k + 2, plus the NEXT tile's records + actions
                // T right after G(0); this environment's group is G(k - 2).
                if (k == 1 || k == 2)
                    cp_async_wait<2>();  // behind G(k-2): [G(0),] T, G(k-1)
                else
                    cp_async_wait<1>();  // behind G(k-2): G(k-1) only -- from k == 3 on T has landed too
                __syncwarp();
                const bool valid = (valid_mask >> k) & 1u;
                if (valid)
                    load_sensors_staged<Real, SPL>(stage[k & 1], ts.N[k], lane, s);
                __syncwarp();
                {
                    int e2 = e + 2, n2 = 0;
                    bool have = false;
                    if (k + 2 < ntile) {
                        n2 = shfl(my_n, k + 2);
                        have = true;
                    } else {
                        const int r = k + 2 - ntile;  // 0 or 1: the next tile's first two environments
                        e2 = UAVSIM_NEXT_BASE + r;
                        if (ntile < 8) {  // a short tile has not passed the waits that retire T
                            cp_async_wait<0>();
                            __syncwarp();
                        }
                        if (e2 < UAVSIM_E) {
                            n2 = (int)ts.rec[tb ^ 1][r][kRN ^ r];
                            have = true;
                        }
                    }
                    if (have) stage_issue<Real, SPL>(b, p, e2, n2, lane, stage[k & 1]);
                    cp_async_commit();
                }
                if (k == 0) {  // T: the other record buffer is free since the previous tile's write-back
                    if (UAVSIM_NEXT_BASE < UAVSIM_E)
                        tile_issue<Real>(b, actions, UAVSIM_NEXT_BASE, tb ^ 1, lane, ts);
                    cp_async_commit();
                }
                if (!valid) continue;
            } else {
                if (!(ts.flags[k] & kFlagValid)) continue;
                load_sensors<Real, SPL>(b, p, e, ts.N[k], lane, s);
            }
            sensor_phase<Real, REPLAY, SPL, FPS, STACK, LEAN>(p, b, ts, ws, tb, k, e, lane, s);
        }
        __syncwarp();

        // ---- phase C: one environment per lane ----
        unsigned restart_mask = 0u;
        {
            bool restart = false;
            if (lane < ntile && (ts.flags[lane] & kFlagValid)) {
                const int e = base + lane;
                const int flags = ts.flags[lane];
                const bool collect = flags & kFlagCollect, truncated = flags & kFlagTruncated;
                const int N = ts.N[lane];
                const Real step_loss = ts.step_loss[lane];
                const int counts = ts.counts[lane];
                Real reward, total_bytes = Real(0);
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
                    if (unvisited > 0) reward += p.p_unvisited * (Real)unvisited;
                    const int starved = (counts >> 16) & 0xff;
                    Real pen = Real(0);
                    for (int i = 0; i < starved; ++i) pen += p.p_starved;  // the reference adds -1000 once per sensor
                    reward += pen;
                }
                // DomainRandEnv.step shaping (dqn.py:415-444)
                Real shaped = reward;
                if (feat_domain_rand<LEAN>(p)) {
                    const float prev = R.f(kRPrevDist), curd = ts.cur_dist[lane];
                    if (prev > 0.f) {
                        if constexpr (kParity)
                            shaped += p.prox_eta * ((double)prev - (double)curd);
                        else
                            shaped += p.prox_eta * (prev - curd);
                    }
                    R.set_f(kRPrevDist, curd);
                    if (R.i32(kRFirstCov) < 0 && (__popc(vis_lo) + __popc(vis_hi)) == N) R.set_i(kRFirstCov, ts.step[lane]);
                    if (flags & kFlagJain) shaped += fdiv(p.jain_bonus * (ts.jain[lane] - Real(0.5)), (Real)N);
                }
                if (FPS == 5 && feat_relational<LEAN>(p)) {  // env_wrapper.py:169-198, :291-313
                    const double W = R.i32(kRGridW), H = R.i32(kRGridH);
                    const double diag = sqrt(W * W + H * H);
                    const float prev = R.f(kRPrevDist), curd = ts.cur_dist[lane];
                    R.set_f(kRPrevDist, curd);
                    int dwell = R.i32(kRDwell);
                    const int n_in = (ts.counts[lane] >> 24) & 0x7f;
                    Real bonus = Real(0);
                    if (!collect) {
                        dwell = 0;
                    } else if (n_in >= p.rel_min_cluster && dwell < p.rel_dwell_steps) {
                        if constexpr (kParity)
                            bonus = p.rel_dwell_bonus * (1.0 - (double)dwell / p.rel_dwell_steps);
                        else
                            bonus = p.rel_dwell_bonus * (1.0f - (float)dwell / (float)p.rel_dwell_steps);
                        dwell += 1;
                    } else if (n_in < p.rel_min_cluster) {
                        dwell = 0;
                    }
                    R.set_i(kRDwell, dwell);
                    if constexpr (kParity) {
                        const double shaping = p.rel_gamma * (-(double)curd / diag) - (-(double)prev / diag);
                        shaped = (reward + shaping + bonus) * p.rel_scale;
                    } else {
                        const float inv = (float)(1.0 / diag);
                        const float shaping = p.rel_gamma * (-curd * inv) + prev * inv;
                        shaped = (reward + shaping + bonus) * p.rel_scale;
                    }
                }
                // commit the record (uav_env.py:487 and the counters of :300-307)
                float x = ts.x[lane], y = ts.y[lane];
                if (feat_wind<LEAN>(p)) {  // WindDriftEnv.step, sim_to_real_analysis.py:214-219
                    const double a = R.d(kRWindAngle);
                    if constexpr (kParity) {  // np.float32 + np.float64 drift -> float64 sum, clip, one rounding to float32
                        x = (float)fmin(fmax((double)x + p.wind_drift * cos(a), 0.0), (double)(R.i32(kRGridW) - 1));
                        y = (float)fmin(fmax((double)y + p.wind_drift * sin(a), 0.0), (double)(R.i32(kRGridH) - 1));
                    } else {
                        float sn, cs;
                        __sincosf((float)a, &sn, &cs);
                        x = fminf(fmaxf(fmaf((float)p.wind_drift, cs, x), 0.f), (float)(R.i32(kRGridW) - 1));
                        y = fminf(fmaxf(fmaf((float)p.wind_drift, sn, y), 0.f), (float)(R.i32(kRGridH) - 1));
                    }
                }
                const double battery = ts.battery[lane];
                R.set_f(kRX, x);
                R.set_f(kRY, y);
                R.set_d(kRBattery, battery);
                R.set_f(kRElapsed, ts.elapsed[lane]);
                R.set_i(kRStep, ts.step[lane]);
                R.set_d(kRTotalReward, R.d(kRTotalReward) + (double)reward);
                bool extra = false;
                if constexpr (!LEAN) extra = b.extra_step != 0;  // PerturbedDQNEnv's wind repeat: its reward is dropped (:119)
                if (!extra) R.set_d(kREpReturn, R.d(kREpReturn) + (double)shaped);
                R.set_f(kRLastBytes, (float)total_bytes);
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
                // outputs
                restart = truncated && p.auto_reset && !extra;  // the commanded step that follows ends the episode
                reinterpret_cast<Real*>(b.reward)[e] = shaped;
                b.done[e] = truncated ? 1 : 0;
                float* o = obs_frame(p, b, e);  // uav_env.py:644
                float f0, f1, f2;
                if constexpr (kParity) {
                    f0 = (float)((double)x / (double)R.i32(kRGridW));
                    f1 = (float)((double)y / (double)R.i32(kRGridH));
                    f2 = (float)(battery / p.max_battery);
                } else {
                    f0 = __fdividef(x, (float)R.i32(kRGridW));
                    f1 = __fdividef(y, (float)R.i32(kRGridH));
                    f2 = (float)battery * p.inv_max_battery;
                }
                if (feat_wind<LEAN>(p)) {  // :220-223: the observation reports the raw drifted position
                    f0 = x;
                    f1 = y;
                }
                if (feat_gps_obs<LEAN>(p)) {  // PerturbedDQNEnv: the noisy position of the observation pass
                    f0 = ts.cur_dist[lane];
                    f1 = (float)ts.jain[lane];
                }
                if (feat_gps<LEAN>(p)) {   // GPSNoisyEnv._add_gps_noise, :176-181
                    float gx, gy;
                    double unused;
                    env_draws<Real, REPLAY>(p, b.step_env_draws, e, ts.episode[lane] - 1, (uint32_t)ts.step[lane], gx, gy, unused);
                    f0 = f0 + gx;
                    f1 = f1 + gy;
                }
                o[0] = f0;
                o[1] = f1;
                o[2] = f2;
                if constexpr (STACK) {
                    o[p.frame_mirror] = f0;
                    o[p.frame_mirror + 1] = f1;
                    o[p.frame_mirror + 2] = f2;
                }
                if (b.action_mask && !restart) b.action_mask[e] = action_mask_of(x, y, R.i32(kRGridW), R.i32(kRGridH));
                if (feat_step_detail<LEAN>(b)) {
                    double* d = feat_step_detail<LEAN>(b) + (size_t)e * UAVSIM_DETAIL_COUNT;
                    if (!collect)  // a collect step's intermediates were written by phase B
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

        // ---- SB3 VecEnv semantics: finished episodes restart inside the same call (rare) ----
        if constexpr (kDeferredRestart) {
            pend_mask = restart_mask;
            pend_base = base;
            restart_mask = 0u;
        }
        while (restart_mask) {
            const int k = __ffs(restart_mask) - 1;
            restart_mask &= restart_mask - 1;
            __syncwarp();
            reinterpret_cast<uint32_t*>(&ws.rec)[lane] = rec[k][lane ^ k];  // unswizzle into the working record
            __syncwarp();
            finish_and_restart<Real, REPLAY, SPL, FPS>(p, b, base + k, lane, ws);
            __syncwarp();
            rec[k][lane ^ k] = reinterpret_cast<const uint32_t*>(&ws.rec)[lane];
        }
        __syncwarp();

        // ---- tile records: shared -> global ----
#pragma unroll
        for (int r = 0; r < kTile; ++r)
            if ((valid_mask >> r) & 1u)
                reinterpret_cast<uint32_t*>(b.rec + base + r)[lane] = rec[r][lane ^ r];
        if constexpr (kStaged) {
            if (ntile < 4) {  // a short (= last) tile never reached the waits that retire the record group
                cp_async_wait<0>();
            }
            __syncwarp();
            tb ^= 1;
        }
        base = next_base;
        next_base = (int)__reduce_max_sync(kFull, (unsigned)pending_ticket) * b.tile;
    }
    if constexpr (kDeferredRestart) {
        while (pend_mask) {  // the tile's records are back in global memory: restart straight from there
            const int e = pend_base + __ffs(pend_mask) - 1;
            pend_mask &= pend_mask - 1;
            load_record(ws, b.rec + e, lane);
            finish_and_restart<Real, REPLAY, SPL, FPS>(p, b, e, lane, ws);
            store_record(ws, b.rec + e, lane);
            __syncwarp();
        }
    }
    if (!(base < UAVSIM_E)) break;
  }
    // every warp of this block has drawn its last ticket; the last block to get here resets the counters
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(b.tile_ticket + 1, 1u) == gridDim.x - 1) {
            b.tile_ticket[0] = 0u;
            b.tile_ticket[1] = 0u;
            __threadfence();
        }
    }
#undef UAVSIM_E
#undef UAVSIM_NEXT_BASE
#undef UAVSIM_TAKE_TICKET
}

}  // namespace uavsim
