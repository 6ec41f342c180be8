/*
 * uavsim_oracle.c -- plain-C CPU restatement of the reference's environment step.
 *
 * TEST INFRASTRUCTURE ONLY.  Built into oracle/_build/liboracle.so by
 * oracle/Makefile; loaded only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs, and only as the checker or the reported
 * CPU baseline.  Nothing in the product package links or calls it.
 *
 * Parity status: PINNED -- tests/test_oracle_c.py checks it (a) against the golden
 * vectors generated from the unmodified reference (tests/golden, via the draws the
 * python oracle re-derives from the fixture seeds) and (b) against
 * oracle/uavsim_oracle.py on random slot tables; all state bit-exact.
 *
 * Draws are always SLOT addressed: table[slot][sensor], slots A,B,U,C,D,E
 * (SURVEY.md section 3.2); normal slots hold np.random.normal(0, sigma) itself.
 *
 * Reference lines followed (under /root/reference/src):
 *   environment/uav.py:127-185 (move) :187-206 (hover) :224 (is_alive) :241-263
 *   environment/iot_sensors.py:13-37 :114-125 :127-145 :147-197 :202-212 :214-219
 *       :223-259 :305-321
 *   rewards/reward_function.py:46-57 :59-67 :69-79 :81-128
 *   environment/uav_env.py:376-394 :400-427 :429-488 :494-516 :518-632 :638-674
 *   agents/dqn/dqn.py:301-372 (DomainRandEnv.reset) :405-444 (shaping) :446-451 (Jain)
 *   experiments/relational_policy/env_wrapper.py:163-228 (reset/step/Dict observation) :270-313
 *       (potential shaping, cluster dwell bonus)
 *
 * Float discipline: the reference mixes float32 geometry with float64 everything
 * else; every expression below keeps the reference's type and association, and the
 * file must be compiled with -ffp-contract=off (see Makefile).  float32 log10 is the
 * normalised, host-independent definition float32(log10(float64(d))) (see
 * oracle/ref_harness.py) and float32 x**2 is the correctly rounded x*x (numpy's scalar
 * float32 power runs a SIMD loop that is off by one ulp on 0.07 % of inputs; the
 * normalised goldens remove that host dependence the same way as for log10).
 * Reductions reproduce numpy's pairwise summation order.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXS 64
#define ORC_SLOTS 6
enum { SLOT_A = 0, SLOT_B, SLOT_U, SLOT_C, SLOT_D, SLOT_E };

typedef struct orc_params {
    double data_generation_rate, max_buffer_size, rssi_threshold, duty_cycle, max_battery,
        collection_duration;
    double tx_power_dbm, noise_floor_dbm, altitude, adr_lambda, shadowing_std_db,
        capture_threshold_db, power_move, power_hover;
    double reward_per_byte, reward_new_sensor, reward_completion, reward_urgency_reduction,
        reward_movement;
    double penalty_revisit, penalty_boundary, penalty_collision, penalty_battery, penalty_hover,
        penalty_step;
    double penalty_data_loss, penalty_starvation, penalty_unvisited, penalty_starved,
        starvation_cr_threshold;
    double prox_eta, jain_bonus;
    double rel_gamma, rel_dwell_bonus, rel_reward_scale; /* env_wrapper.py:47-62 */
    int32_t max_steps, init_sf, features_per_sensor, obs_slots, domain_rand, relational;
    int32_t rel_dwell_steps, rel_min_cluster;
} orc_params;

typedef struct orc_env {
    int32_t W, H, N, step;
    float x, y, start_x, start_y;
    double battery, total_reward, total_data_collected, episode_return, last_step_bytes;
    double prev_dist_nearest;
    int32_t capture_triggers, boundary_hits, edge_steps, collisions, first_full_coverage_step,
        delivered, dwell_steps, pad0;
    float spx[ORC_MAXS], spy[ORC_MAXS];
    double buf[ORC_MAXS], gen[ORC_MAXS], tx[ORC_MAXS], lost[ORC_MAXS], ema[ORC_MAXS], cur[ORC_MAXS];
    int8_t sf[ORC_MAXS];
    uint8_t visited[ORC_MAXS];
} orc_env;

/* per-step diagnostics (same order as UAVSIM_DETAIL_* in include/uavsim.h) */
enum { D_COLLISIONS = 0, D_BYTES, D_URG_RED, D_MEAN_URG, D_WAS_NEW, D_WAS_EMPTY, D_ALL_COLLECTED,
       D_DATA_LOSS, D_ATT_LO, D_ATT_HI, D_WIN_LO, D_WIN_HI, D_BASE_REWARD, D_TRUNCATED,
       D_STARVED = 14, /* sensors below the collection-ratio threshold when the episode ended (reward_function.py:59-67) */
       /* oracle-only entries (the CUDA detail row has 16 doubles): how close this env-step came to flipping one
        * of its discrete decisions, and the scale of the reward's terms.  A float32 build is allowed to differ in
        * a discrete decision only where the float64 margin recorded here is below its arithmetic noise
        * (tests/parity_util.py, SURVEY.md section 7 hard part 1c). */
       D_MARGIN_DB = 16,    /* min |lhs - rhs| over the dB compares: ADR ladder, range checks, capture threshold */
       D_MARGIN_PROB = 17,  /* min |p_link * duty - u| over the duty-cycle gates, and |tx / gen - threshold| over the
                               terminal starvation tests */
       D_MARGIN_BYTES = 18, /* min |buffer + new - max_buffer| over the saturation compares and |buffer - rate * dt|
                               over the drains (whether the buffer empties), in bytes */
       D_REWARD_SCALE = 19, /* sum of |term| over the terms the base reward was assembled from */
       D_MARGIN_DIST = 20,  /* min |d - d_break| (metres) over the two-ray branch selections (iot_sensors.py:174-187): the
                               two branches differ by 0.46 dB at the break, so the selection is a discrete decision */
       D_COUNT = 21 };

/* margin / scale accumulators of the env-step in flight (NULL = not recorded) */
static __thread double* g_mg = NULL;
static inline void mg_note(int slot, double lhs, double rhs) {
    if (!g_mg) return;
    double d = fabs(lhs - rhs);
    if (d < g_mg[slot]) g_mg[slot] = d;
}
static inline double term(double v) { /* a reward term: returns v, accumulates |v| */
    if (g_mg) g_mg[D_REWARD_SCALE] += fabs(v);
    return v;
}

size_t orc_sizeof_env(void) { return sizeof(orc_env); }
size_t orc_sizeof_params(void) { return sizeof(orc_params); }

/* ---- numpy-order reductions (numpy/_core/src/umath/loops_utils.h.src pairwise sum) ---- */
static double np_sum_f64(const double* a, int n) {
    if (n < 8) {
        double r = 0.;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    }
    double r[8];
    int i;
    for (i = 0; i < 8; i++) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res; /* n <= 128 always here (n <= 64) */
}
static float np_sum_f32(const float* a, int n) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    }
    float r[8];
    int i;
    for (i = 0; i < 8; i++) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

static double data_rate(int sf) { /* iot_sensors.py:13-20 */
    switch (sf) {
        case 7: return 5470 / 8.0;
        case 8: return 3125 / 8.0;
        case 9: return 1760 / 8.0;
        case 10: return 980 / 8.0;
        case 11: return 440 / 8.0;
        default: return 250 / 8.0;
    }
}
static double required_snr(int sf) { /* iot_sensors.py:22-29, default 7.5 at :199-200 */
    switch (sf) {
        case 7: return -6.0;
        case 8: return -9.0;
        case 9: return -12.0;
        case 10: return -15.0;
        case 11: return -17.5;
        case 12: return -20.0;
        default: return 7.5;
    }
}
static double link_quality(int sf) { /* uav_env.py:647 */
    switch (sf) {
        case 7: return 1.0;
        case 8: return 0.8;
        case 9: return 0.6;
        case 10: return 0.4;
        case 11: return 0.2;
        default: return 0.1;
    }
}

/* iot_sensors.py:161-187: deterministic part of calculate_rssi */
static double det_rssi(const orc_params* p, float ux, float uy, float sx, float sy) {
    float dx = (ux - sx) * 10.0f;
    float dy = (uy - sy) * 10.0f;
    float g = sqrtf(dx * dx + dy * dy);
    float alt2 = (float)((long)p->altitude * (long)p->altitude); /* python int, then cast by the f32 add */
    float d = sqrtf(g * g + alt2);
    float lg = (float)log10((double)d);
    double d_break = (4 * M_PI * 0.5 * p->altitude) / 0.345;
    mg_note(D_MARGIN_DIST, (double)d, d_break);
    double pl;
    if ((double)d < d_break) {
        float t = 20.0f * lg;
        pl = ((double)t + 20 * log10(868.0)) - 28;
    } else {
        float t = 40.0f * lg;
        pl = ((double)t - 20 * log10(0.5)) - 20 * log10(p->altitude);
    }
    return p->tx_power_dbm - pl;
}

static void adr_update(const orc_params* p, orc_env* e, int i, double rssi) { /* iot_sensors.py:223-259 */
    e->cur[i] = rssi;
    if (isnan(e->ema[i]))
        e->ema[i] = rssi;
    else
        e->ema[i] = (p->adr_lambda * rssi) + ((1 - p->adr_lambda) * e->ema[i]);
    double avg = e->ema[i];
    mg_note(D_MARGIN_DB, avg, -60);
    mg_note(D_MARGIN_DB, avg, -70);
    mg_note(D_MARGIN_DB, avg, -78);
    mg_note(D_MARGIN_DB, avg, -85);
    if (avg > -60)
        e->sf[i] = 7;
    else if (avg > -70)
        e->sf[i] = 9;
    else if (avg > -78)
        e->sf[i] = 11;
    else if (avg > -85)
        e->sf[i] = 12;
}

static double urgency(const orc_params* p, const orc_env* e, int i) { /* uav_env.py:376-384 */
    double util = e->buf[i] / p->max_buffer_size;
    double lr = e->gen[i] > 0 ? e->lost[i] / e->gen[i] : 0.0;
    double u = util * (1.0 + lr * 10.0);
    return u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
}

static void aoi(const orc_params* p, const orc_env* e, float* out) { /* uav_env.py:386-394 */
    for (int i = 0; i < e->N; i++)
        out[i] = p->data_generation_rate > 0 ? (float)(e->buf[i] / p->data_generation_rate) : 0.0f;
}

/* uav_env.py:638-674; obs row is obs_dim = 3 + fps*obs_slots floats, ghost slots zero (dqn.py:286-298) */
static void observe(const orc_params* p, orc_env* e, const double* draws, int P, float* obs) {
    double W = (double)e->W, H = (double)e->H;
    double ux = (double)e->x, uy = (double)e->y;
    int fps = p->features_per_sensor;
    int dim = 3 + fps * p->obs_slots;
    for (int k = 0; k < dim; k++) obs[k] = 0.0f;
    obs[0] = (float)(ux / W);
    obs[1] = (float)(uy / H);
    obs[2] = (float)(e->battery / p->max_battery);
    for (int i = 0; i < e->N; i++) {
        double urg = urgency(p, e, i);
        double det = det_rssi(p, e->x, e->y, e->spx[i], e->spy[i]);
        adr_update(p, e, i, det + draws[SLOT_D * P + i]);
        double r_e = det + draws[SLOT_E * P + i];
        mg_note(D_MARGIN_DB, r_e, p->rssi_threshold);
        double lq = r_e >= p->rssi_threshold ? link_quality(e->sf[i]) : 0.0;
        float* o = obs + 3 + fps * i;
        o[0] = (float)(e->buf[i] / p->max_buffer_size);
        o[1] = (float)urg;
        o[2] = (float)lq;
        if (fps == 5) {
            o[3] = (float)(((double)e->spx[i] - ux) / W);
            o[4] = (float)(((double)e->spy[i] - uy) / H);
        }
    }
}

/* dqn.py:405-412: float32 norm of the float32 difference, min over sensors with data */
static double dist_nearest_with_data(const orc_env* e) {
    int any = 0;
    float best = 0.f;
    for (int i = 0; i < e->N; i++) {
        if (!(e->buf[i] > 0)) continue;
        float dx = e->spx[i] - e->x, dy = e->spy[i] - e->y;
        float d = sqrtf(dx * dx + dy * dy);
        if (!any || d < best) best = d;
        any = 1;
    }
    return any ? (double)best : 0.0;
}

/* env_wrapper.py:270-289: distance (float32 norm) to the nearest UNVISITED sensor, 0 when all are visited;
 * the potential is -dist / grid_diagonal. */
static float nearest_unvisited(const orc_env* e) {
    int any = 0;
    float best = 0.f;
    for (int i = 0; i < e->N; i++) {
        if (e->visited[i]) continue;
        float dx = e->spx[i] - e->x, dy = e->spy[i] - e->y;
        float d = sqrtf(dx * dx + dy * dy);
        if (!any || d < best) best = d;
        any = 1;
    }
    return any ? best : 0.f;
}
static double potential_of(const orc_env* e, float dist) {
    double diag = sqrt((double)((long)e->W * e->W + (long)e->H * e->H));
    return -(double)dist / diag;
}

/* env_wrapper.py:200-228: row = uav(3) then per slot [rel_x, rel_y, buffer, urgency, visited]; the mask is
 * (slot < N) and is not part of the row.  No draws, no ADR. */
static void observe_relational(const orc_params* p, const orc_env* e, float* obs) {
    double W = (double)e->W, H = (double)e->H, ux = (double)e->x, uy = (double)e->y;
    int dim = 3 + 5 * p->obs_slots;
    for (int k = 0; k < dim; k++) obs[k] = 0.0f;
    obs[0] = (float)(ux / W);
    obs[1] = (float)(uy / H);
    obs[2] = (float)(e->battery / p->max_battery);
    for (int i = 0; i < e->N; i++) {
        double rx = ((double)e->spx[i] - ux) / W, ry = ((double)e->spy[i] - uy) / H;
        rx = rx < -1.0 ? -1.0 : (rx > 1.0 ? 1.0 : rx);
        ry = ry < -1.0 ? -1.0 : (ry > 1.0 ? 1.0 : ry);
        float* o = obs + 3 + 5 * i;
        o[0] = (float)rx;
        o[1] = (float)ry;
        o[2] = (float)(e->buf[i] / p->max_buffer_size);
        o[3] = (float)urgency(p, e, i);
        o[4] = e->visited[i] ? 1.0f : 0.0f;
    }
}

static uint8_t action_mask(const orc_env* e) { /* derived from uav.py:175 */
    float x = e->x, y = e->y;
    uint8_t m = 1u << 4;
#define OK(nx, ny) (0 <= (nx) && (nx) < (float)e->W && 0 <= (ny) && (ny) < (float)e->H)
    if (OK(x, y + 1.0f)) m |= 1u << 0;
    if (OK(x, y - 1.0f)) m |= 1u << 1;
    if (OK(x - 1.0f, y)) m |= 1u << 2;
    if (OK(x + 1.0f, y)) m |= 1u << 3;
#undef OK
    return m;
}

/* Reset one env.  FIXED mode: uav_env.py:400-427 + iot_sensors.py:305-321 (fill[i] =
 * np_random.uniform(0.2,0.6)).  DOMAIN_RAND mode (p->domain_rand): the caller has already stored
 * the new grid/layout/start in *e (they are injected); sensors are fresh objects with empty
 * buffers and SF = init_sf (dqn.py:342-360).  Only draw slots D,E are read. */
void orc_reset_env(const orc_params* p, orc_env* e, const double* fill, const double* draws, int P,
                   int init_sf, float* obs, uint8_t* mask) {
    e->x = e->start_x;
    e->y = e->start_y;
    e->battery = p->max_battery;
    for (int i = 0; i < e->N; i++) {
        if (p->domain_rand) {
            e->buf[i] = 0.0;
            e->gen[i] = 0.0;
        } else {
            double f = fill[i] < 0.0 ? 0.0 : (fill[i] > 1.0 ? 1.0 : fill[i]);
            e->buf[i] = p->max_buffer_size * f;
            e->gen[i] = e->buf[i];
        }
        e->tx[i] = 0.0;
        e->lost[i] = 0.0;
        e->ema[i] = NAN;
        e->cur[i] = NAN;
        e->sf[i] = (int8_t)(init_sf > 0 ? init_sf : p->init_sf);
        e->visited[i] = 0;
    }
    e->step = 0;
    e->total_reward = 0.0;
    e->total_data_collected = 0.0;
    e->episode_return = 0.0;
    e->last_step_bytes = 0.0;
    e->capture_triggers = e->boundary_hits = e->edge_steps = e->collisions = e->delivered = 0;
    e->first_full_coverage_step = -1;
    observe(p, e, draws, P, obs); /* the parent builds (and draws for) its flat observation in every mode */
    e->prev_dist_nearest = dist_nearest_with_data(e);
    if (p->relational) { /* env_wrapper.py:163-167 */
        observe_relational(p, e, obs);
        e->dwell_steps = 0;
        e->prev_dist_nearest = (double)nearest_unvisited(e);
    }
    if (mask) *mask = action_mask(e);
}

static double move_action(const orc_params* p, orc_env* e, int a, double step_loss) {
    float nx = e->x, ny = e->y; /* uav.py:152-163 */
    if (a == 0) ny += 1.0f;
    else if (a == 1) ny -= 1.0f;
    else if (a == 2) nx -= 1.0f;
    else nx += 1.0f;
    double before = e->battery;
    int ok = (0 <= nx && nx < (float)e->W && 0 <= ny && ny < (float)e->H);
    if (ok) {
        e->x = nx;
        e->y = ny;
        e->battery -= (p->power_move * 1) / 3600;
    } else {
        e->battery -= ((p->power_move * 0.5) * 1) / 3600;
        e->boundary_hits++;
    }
    double used = before - e->battery;
    double r = term(p->penalty_step); /* reward_function.py:69-79 */
    r += term(ok ? p->reward_movement : p->penalty_boundary);
    r += term(p->penalty_battery * used);
    r += term(p->penalty_data_loss * step_loss); /* uav_env.py:514 */
    return r;
}

static double collect_action(const orc_params* p, orc_env* e, const double* draws, int P,
                             double step_loss, double* detail) {
    int N = e->N;
    double cd = p->collection_duration, thr = p->rssi_threshold;
    float before[ORC_MAXS], after[ORC_MAXS], diff[ORC_MAXS];
    aoi(p, e, before);
    e->battery -= (p->power_hover * cd) / 3600;
    double used = (p->power_hover / (60 * 60)) * cd;
    double det[ORC_MAXS];
    int attempt[ORC_MAXS];
    uint64_t att_mask = 0, win_mask = 0;
    for (int i = 0; i < N; i++) det[i] = det_rssi(p, e->x, e->y, e->spx[i], e->spy[i]);
    for (int i = 0; i < N; i++) {
        attempt[i] = 0;
        if (e->buf[i] <= 0) continue;
        adr_update(p, e, i, det[i] + draws[SLOT_A * P + i]);
        double rb = det[i] + draws[SLOT_B * P + i];
        mg_note(D_MARGIN_DB, rb, thr);
        double pl;
        if (rb < thr)
            pl = 0.0;
        else {
            double snr = rb - p->noise_floor_dbm;
            pl = 1.0 / (1.0 + exp(-(snr - required_snr(e->sf[i]))));
        }
        double pall = pl * (p->duty_cycle / 100.0);
        mg_note(D_MARGIN_PROB, pall, draws[SLOT_U * P + i]);
        if (pall > draws[SLOT_U * P + i]) {
            attempt[i] = 1;
            att_mask |= 1ull << i;
        }
    }
    /* SF groups in first-appearance order (dict insertion order, uav_env.py:551,557) */
    int order_sf[ORC_MAXS], ngroups = 0;
    for (int i = 0; i < N; i++) {
        if (!attempt[i]) continue;
        int seen = 0;
        for (int g = 0; g < ngroups; g++) seen |= (order_sf[g] == e->sf[i]);
        if (!seen) order_sf[ngroups++] = e->sf[i];
    }
    int winners[ORC_MAXS], nwin = 0, collisions = 0;
    for (int g = 0; g < ngroups; g++) {
        int k = 0, a = -1, b = -1;
        for (int i = 0; i < N; i++) {
            if (!attempt[i] || e->sf[i] != order_sf[g]) continue;
            k++;
            /* stable descending sort: strictly greater moves ahead, ties keep index order */
            if (a < 0 || e->cur[i] > e->cur[a]) {
                b = a;
                a = i;
            } else if (b < 0 || e->cur[i] > e->cur[b]) {
                b = i;
            }
        }
        if (k == 1) {
            winners[nwin++] = a;
        } else {
            collisions += k - 1;
            mg_note(D_MARGIN_DB, e->cur[a], e->cur[b] + p->capture_threshold_db);
            if (e->cur[a] > (e->cur[b] + p->capture_threshold_db)) {
                winners[nwin++] = a;
                e->capture_triggers++;
            }
        }
    }
    double total = 0.0;
    int any_new = 0;
    for (int w = 0; w < nwin; w++) { /* uav_env.py:580-594, iot_sensors.py:127-145 */
        int i = winners[w];
        win_mask |= 1ull << i;
        double rc = det[i] + draws[SLOT_C * P + i];
        mg_note(D_MARGIN_DB, rc, thr);
        int in_range = rc >= thr;
        double got = 0.0;
        if (in_range && e->buf[i] > 0) {
            double cap = data_rate(e->sf[i]) * cd;
            mg_note(D_MARGIN_BYTES, e->buf[i], cap);
            got = e->buf[i] < cap ? e->buf[i] : cap;
            e->buf[i] -= got;
            e->tx[i] += got;
            if (got > 0) e->delivered++;
        }
        if (in_range && got > 0) {
            total += got;
            e->total_data_collected += got;
            if (!e->visited[i]) {
                any_new = 1;
                e->visited[i] = 1;
            }
        }
    }
    int was_empty = 0, all_collected = 1;
    for (int i = 0; i < N; i++) {
        was_empty |= (e->buf[i] <= 0);
        all_collected &= (e->buf[i] <= 0);
    }
    aoi(p, e, after);
    for (int i = 0; i < N; i++) {
        float d = before[i] - after[i];
        diff[i] = d > 0.f ? d : 0.f;
    }
    double urg_red = (double)np_sum_f32(diff, N);
    e->last_step_bytes = total;
    double mean_urg = 0.0;
    if (nwin > 0) {
        double u[ORC_MAXS];
        for (int w = 0; w < nwin; w++) u[w] = urgency(p, e, winners[w]);
        mean_urg = np_sum_f64(u, nwin) / nwin;
    }
    double r = term(p->penalty_step + p->penalty_hover); /* reward_function.py:81-128 */
    if (total > 0) {
        r += term(p->reward_per_byte * total * mean_urg);
        if (any_new) r += term(p->reward_new_sensor);
    }
    if (urg_red > 0) r += term(p->reward_urgency_reduction * urg_red);
    if (was_empty && total == 0) r += term(p->penalty_revisit);
    r += term(p->penalty_battery * used);
    if (collisions > 0) r += term(p->penalty_collision * collisions);
    if (step_loss > 0) r += term(p->penalty_data_loss * step_loss);
    if (N > 1) { /* reward_function.py:46-57 */
        double mx = e->buf[0];
        for (int i = 1; i < N; i++) mx = e->buf[i] > mx ? e->buf[i] : mx;
        if (mx != 0) {
            double nrm[ORC_MAXS], dev[ORC_MAXS];
            for (int i = 0; i < N; i++) nrm[i] = e->buf[i] / mx;
            double mean = np_sum_f64(nrm, N) / N;
            for (int i = 0; i < N; i++) {
                double d = nrm[i] - mean;
                dev[i] = d * d;
            }
            r += term(p->penalty_starvation * (np_sum_f64(dev, N) / N));
        }
    }
    if (all_collected) r += term(p->reward_completion);
    e->collisions += collisions;
    if (detail) {
        detail[D_COLLISIONS] = collisions;
        detail[D_BYTES] = total;
        detail[D_URG_RED] = urg_red;
        detail[D_MEAN_URG] = mean_urg;
        detail[D_WAS_NEW] = any_new;
        detail[D_WAS_EMPTY] = was_empty;
        detail[D_ALL_COLLECTED] = all_collected;
        detail[D_ATT_LO] = (double)(uint32_t)att_mask;
        detail[D_ATT_HI] = (double)(uint32_t)(att_mask >> 32);
        detail[D_WIN_LO] = (double)(uint32_t)win_mask;
        detail[D_WIN_HI] = (double)(uint32_t)(win_mask >> 32);
    }
    return r;
}

/* One env, one step.  Returns truncated.  reward_out = what the (possibly wrapped) env returns:
 * base reward, plus DomainRandEnv shaping in domain_rand mode.  Invalid action -> returns -1 and
 * leaves the env untouched (the reference raises ValueError, uav_env.py:468, after having aged the
 * sensors; callers validate first, as the host wrapper does). */
int orc_step_env(const orc_params* p, orc_env* e, int action, const double* draws, const double* dwell, int P,
                 float* obs, double* reward_out, uint8_t* mask, double* detail) {
    if (action < 0 || action > 4) return -1;
    if (detail) {
        for (int k = 0; k < D_COUNT; k++) detail[k] = 0.0;
        detail[D_MARGIN_DB] = detail[D_MARGIN_PROB] = detail[D_MARGIN_BYTES] = detail[D_MARGIN_DIST] = INFINITY;
    }
    g_mg = detail;
    double prev_dist = e->prev_dist_nearest;
    e->step += 1;
    double W = (double)e->W, H = (double)e->H, ux = (double)e->x, uy = (double)e->y, eps = 1e-6;
    if (ux <= eps || uy <= eps || ux >= W - 1 - eps || uy >= H - 1 - eps) e->edge_steps++;
    double dt = action == 4 ? p->collection_duration : 1.0;
    double step_loss = 0.0;
    for (int i = 0; i < e->N; i++) { /* iot_sensors.py:114-125 */
        double nw = p->data_generation_rate * dt;
        e->gen[i] += nw;
        double pot = e->buf[i] + nw;
        mg_note(D_MARGIN_BYTES, pot, p->max_buffer_size);
        if (pot > p->max_buffer_size) {
            double loss = pot - p->max_buffer_size;
            e->buf[i] = p->max_buffer_size;
            e->lost[i] += loss;
            step_loss += loss; /* the reference differences two cumulative sums (uav_env.py:457-459): equal to <= 1e-11 */
        } else {
            e->buf[i] = pot;
        }
    }
    double r;
    if (action < 4) {
        e->last_step_bytes = 0.0;
        r = move_action(p, e, action, step_loss);
    } else {
        r = collect_action(p, e, draws, P, step_loss, detail);
    }
    int truncated = 0;
    if (!(e->battery > 0.02 * p->max_battery)) truncated = 1; /* uav.py:224 */
    if (e->step >= p->max_steps) truncated = 1;
    if (truncated) { /* uav_env.py:481-485, reward_function.py:59-67 */
        int unvisited = 0;
        for (int i = 0; i < e->N; i++) unvisited += !e->visited[i];
        if (unvisited > 0) r += term(p->penalty_unvisited * unvisited);
        double pen = 0.0;
        int starved = 0;
        for (int i = 0; i < e->N; i++) {
            if (e->gen[i] > 0) mg_note(D_MARGIN_PROB, e->tx[i] / e->gen[i], p->starvation_cr_threshold);
            if (e->gen[i] > 0 && e->tx[i] / e->gen[i] < p->starvation_cr_threshold) {
                pen += p->penalty_starved;
                starved++;
            }
        }
        r += term(pen);
        if (detail) detail[D_STARVED] = starved;
    }
    e->total_reward += r;
    observe(p, e, draws, P, obs);
    double shaped = r;
    if (p->domain_rand) { /* dqn.py:415-444 */
        double cur = dist_nearest_with_data(e);
        if (prev_dist > 0) shaped += term(p->prox_eta * (prev_dist - cur));
        e->prev_dist_nearest = cur;
        int nv = 0;
        for (int i = 0; i < e->N; i++) nv += e->visited[i];
        if (e->first_full_coverage_step < 0 && nv == e->N) e->first_full_coverage_step = e->step;
        double s1 = 0.0, s2 = 0.0;
        int cnt = 0;
        for (int i = 0; i < e->N; i++) {
            if (e->gen[i] > 0) {
                double rate = (e->tx[i] / e->gen[i]) * 100;
                s1 += rate;
                s2 += rate * rate;
                cnt++;
            }
        }
        if (cnt > 0) {
            double j = (cnt > 0 && s2 > 0) ? (s1 * s1) / (cnt * s2) : 1.0;
            shaped += term(p->jain_bonus * (j - 0.5) / e->N);
        }
    }
    if (p->relational) { /* env_wrapper.py:169-198 */
        double phi_s = potential_of(e, (float)prev_dist);
        float nd = nearest_unvisited(e);
        double phi_n = potential_of(e, nd);
        e->prev_dist_nearest = (double)nd;
        double shaping = p->rel_gamma * phi_n - phi_s;
        double bonus = 0.0;
        if (action != 4) { /* :291-313 */
            e->dwell_steps = 0;
        } else {
            int n_in = 0;
            for (int i = 0; i < e->N; i++) {
                double det = det_rssi(p, e->x, e->y, e->spx[i], e->spy[i]);
                n_in += (det + dwell[i]) >= p->rssi_threshold;
            }
            if (n_in >= p->rel_min_cluster && e->dwell_steps < p->rel_dwell_steps) {
                double frac = 1.0 - (double)e->dwell_steps / p->rel_dwell_steps;
                bonus = p->rel_dwell_bonus * frac;
                e->dwell_steps += 1;
            } else if (n_in < p->rel_min_cluster) {
                e->dwell_steps = 0;
            }
        }
        shaped = (r + shaping + bonus) * p->rel_reward_scale;
        observe_relational(p, e, obs);
    }
    e->episode_return += shaped;
    *reward_out = shaped;
    if (mask) *mask = action_mask(e);
    if (detail) {
        detail[D_DATA_LOSS] = step_loss;
        detail[D_BASE_REWARD] = r;
        detail[D_TRUNCATED] = truncated;
    }
    g_mg = NULL;
    return truncated;
}

/* Batched drivers over the env range [e0, e1): callers parallelise by giving disjoint ranges to
 * host threads (ctypes releases the GIL).  draws: [E][6][P]; fill: [E][P]; obs: [E][obs_dim]. */
void orc_reset_batch(const orc_params* p, orc_env* envs, int e0, int e1, const double* fill,
                     const double* draws, int P, const int32_t* init_sf, float* obs, uint8_t* mask,
                     double* detail /* nullable [E][D_COUNT]: only the margin entries are written */) {
    int dim = 3 + p->features_per_sensor * p->obs_slots;
    for (int e = e0; e < e1; e++) {
        if (detail) {
            g_mg = detail + (size_t)e * D_COUNT;
            g_mg[D_MARGIN_DB] = g_mg[D_MARGIN_PROB] = g_mg[D_MARGIN_BYTES] = g_mg[D_MARGIN_DIST] = INFINITY;
        }
        orc_reset_env(p, &envs[e], fill ? fill + (size_t)e * P : NULL, draws + (size_t)e * ORC_SLOTS * P, P,
                      init_sf ? init_sf[e] : 0, obs + (size_t)e * dim, mask ? mask + e : NULL);
        g_mg = NULL;
    }
}

/* Returns the number of invalid actions. */
int orc_step_batch(const orc_params* p, orc_env* envs, int e0, int e1, const int32_t* actions,
                   const double* draws, const double* dwell, int P, float* obs, double* reward, uint8_t* done,
                   uint8_t* mask, double* detail) {
    int dim = 3 + p->features_per_sensor * p->obs_slots;
    int bad = 0;
    for (int e = e0; e < e1; e++) {
        int t = orc_step_env(p, &envs[e], actions[e], draws + (size_t)e * ORC_SLOTS * P,
                             dwell ? dwell + (size_t)e * P : NULL, P, obs + (size_t)e * dim, &reward[e],
                             mask ? mask + e : NULL,
                             detail ? detail + (size_t)e * D_COUNT : NULL);
        if (t < 0)
            bad++;
        else
            done[e] = (uint8_t)t;
    }
    return bad;
}
