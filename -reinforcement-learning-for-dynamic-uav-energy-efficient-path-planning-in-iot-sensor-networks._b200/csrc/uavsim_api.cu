// C-ABI layer of libuavsim.so (see include/uavsim.h for the contract and the reference lines each
// entry point stands in for).  Host logic only: validation, arena layout, parameter conversion,
// kernel dispatch.  All compute lives in uavsim_kernels.cuh.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "uavsim_launch.cuh"
#include "uavsim_layout.cuh"

using namespace uavsim;

static_assert(sizeof(uavsim_env_record) == 128, "env record must stay one 128-byte line");

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    return fail(UAVSIM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define CUDA_TRY(expr)                                   \
    do {                                                 \
        cudaError_t _e = (expr);                         \
        if (_e != cudaSuccess) return cuda_fail(_e, #expr); \
    } while (0)

constexpr int64_t kAlign = 256;
// frames kept per env: the stacked window plus the LatencyEnv delay line (sim_to_real_analysis.py:227-245)
int ring_frames(const uavsim_config& c) { return c.frame_stack + c.latency_steps; }

int64_t align_up(int64_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct Layout {
    int64_t rec, sstate, spos, sgen, ssf, stat_sums, error_word, actions, total;
};

int validate(const uavsim_config* c) {
    if (!c) return fail(UAVSIM_ERR_INVALID_ARG, "config is NULL");
    if (c->struct_size != sizeof(uavsim_config))
        return fail(UAVSIM_ERR_INVALID_ARG, "config.struct_size %u != %zu (ABI mismatch)", c->struct_size,
                    sizeof(uavsim_config));
    if (c->num_envs <= 0) return fail(UAVSIM_ERR_INVALID_ARG, "num_envs must be positive");
    if (c->max_sensors <= 0 || c->max_sensors > UAVSIM_MAX_SENSORS)
        return fail(UAVSIM_ERR_INVALID_ARG, "max_sensors must be in [1, %d]", UAVSIM_MAX_SENSORS);
    if (c->num_sensors <= 0 || c->num_sensors > c->max_sensors)
        return fail(UAVSIM_ERR_INVALID_ARG, "num_sensors must be in [1, max_sensors]");
    if (c->obs_slots < c->max_sensors || c->obs_slots > UAVSIM_MAX_SENSORS)
        return fail(UAVSIM_ERR_INVALID_ARG, "obs_slots must be in [max_sensors, %d]", UAVSIM_MAX_SENSORS);
    if (c->features_per_sensor != 3 && c->features_per_sensor != 5)
        return fail(UAVSIM_ERR_INVALID_ARG, "features_per_sensor must be 3 or 5 (uav_env.py:349)");
    if (c->frame_stack < 1 || c->frame_stack > 16) return fail(UAVSIM_ERR_INVALID_ARG, "frame_stack must be in [1, 16]");
    if (c->rng_mode == UAVSIM_RNG_PHILOX && c->max_sensors <= 32 && (c->env_id_offset & 1))
        return fail(UAVSIM_ERR_INVALID_ARG, "env_id_offset must be even when max_sensors <= 32 (observation draws are shared by even/odd environment pairs)");
    if (c->latency_steps < 0 || c->latency_steps > 16) return fail(UAVSIM_ERR_INVALID_ARG, "latency_steps must be in [0, 16]");
    if (c->precision != UAVSIM_F64 && c->precision != UAVSIM_F32)
        return fail(UAVSIM_ERR_INVALID_ARG, "precision must be UAVSIM_F64 or UAVSIM_F32");
    if (c->rng_mode != UAVSIM_RNG_REPLAY && c->rng_mode != UAVSIM_RNG_PHILOX)
        return fail(UAVSIM_ERR_INVALID_ARG, "rng_mode must be UAVSIM_RNG_REPLAY or UAVSIM_RNG_PHILOX");
    if (c->reset_mode != UAVSIM_RESET_FIXED && c->reset_mode != UAVSIM_RESET_DOMAIN_RAND)
        return fail(UAVSIM_ERR_INVALID_ARG, "reset_mode must be UAVSIM_RESET_FIXED or UAVSIM_RESET_DOMAIN_RAND");
    if (c->grid_w <= 0 || c->grid_h <= 0) return fail(UAVSIM_ERR_INVALID_ARG, "grid must be positive");
    if ((int64_t)c->num_envs * (3 + c->features_per_sensor * c->obs_slots) * (ring_frames(*c) > 1 ? 2 * ring_frames(*c) : 1) >= (1ll << 32))
        return fail(UAVSIM_ERR_INVALID_ARG, "num_envs * observation floats per env must be < 2^32 per handle");
    if ((int64_t)c->num_envs * c->max_sensors >= (1ll << 32))
        return fail(UAVSIM_ERR_INVALID_ARG, "num_envs * max_sensors must be < 2^32 per handle");
    if (c->max_steps <= 0 || c->max_steps >= (1 << 24))
        return fail(UAVSIM_ERR_INVALID_ARG, "max_steps must be in [1, 2^24)");
    if (!(c->data_generation_rate > 0)) return fail(UAVSIM_ERR_INVALID_ARG, "data_generation_rate must be > 0");
    if (!(c->max_buffer_size > 0)) return fail(UAVSIM_ERR_INVALID_ARG, "max_buffer_size must be > 0");
    if (!(c->max_battery > 0)) return fail(UAVSIM_ERR_INVALID_ARG, "max_battery must be > 0");
    if (!(c->collection_duration > 0)) return fail(UAVSIM_ERR_INVALID_ARG, "collection_duration must be > 0");
    if (c->init_sf < 7 || c->init_sf > 12) return fail(UAVSIM_ERR_INVALID_ARG, "init_sf must be in [7, 12]");
    if (!(c->gps_sigma >= 0.0) || !(c->wind_drift >= 0.0)) return fail(UAVSIM_ERR_INVALID_ARG, "gps_sigma / wind_drift must be >= 0");
    if (c->wrapper != UAVSIM_WRAP_NONE && c->wrapper != UAVSIM_WRAP_RELATIONAL)
        return fail(UAVSIM_ERR_INVALID_ARG, "wrapper must be UAVSIM_WRAP_NONE or UAVSIM_WRAP_RELATIONAL");
    if (c->gps_in_observation != 0 && c->gps_in_observation != 1)
        return fail(UAVSIM_ERR_INVALID_ARG, "gps_in_observation must be 0 or 1");
    if (c->gps_in_observation && (c->reset_mode != UAVSIM_RESET_FIXED || c->wrapper != UAVSIM_WRAP_NONE))
        return fail(UAVSIM_ERR_INVALID_ARG, "gps_in_observation (PerturbedDQNEnv) sits on the plain UAVEnvironment");
    if (c->wrapper == UAVSIM_WRAP_RELATIONAL) {
        if (c->features_per_sensor != 5)
            return fail(UAVSIM_ERR_INVALID_ARG, "the relational wrapper needs features_per_sensor == 5 (env_wrapper.py:147)");
        if (c->reset_mode != UAVSIM_RESET_FIXED)
            return fail(UAVSIM_ERR_INVALID_ARG, "the relational wrapper sits on UAVEnvironment (UAVSIM_RESET_FIXED)");
        if (c->rel_dwell_steps <= 0) return fail(UAVSIM_ERR_INVALID_ARG, "rel_dwell_steps must be positive");
    }
    if (c->reset_mode == UAVSIM_RESET_DOMAIN_RAND && c->rng_mode == UAVSIM_RNG_PHILOX) {
        if (c->num_grids <= 0 || c->num_grids > UAVSIM_MAX_GRIDS)
            return fail(UAVSIM_ERR_INVALID_ARG, "num_grids must be in [1, %d]", UAVSIM_MAX_GRIDS);
        for (int i = 0; i < c->num_grids; ++i)
            if (c->grids[i][0] <= 0 || c->grids[i][1] <= 0)
                return fail(UAVSIM_ERR_INVALID_ARG, "stage grid %d is not positive", i);
        if (c->max_start_tries <= 0) return fail(UAVSIM_ERR_INVALID_ARG, "max_start_tries must be positive");
    }
    return UAVSIM_OK;
}

int real_bytes(const uavsim_config* c) { return c->precision == UAVSIM_F64 ? 8 : 4; }

Layout make_layout(const uavsim_config* c) {
    const int64_t E = c->num_envs, P = c->max_sensors, rb = real_bytes(c);
    Layout L;
    int64_t off = 0;
    L.rec = off;
    off = align_up(off + E * (int64_t)sizeof(uavsim_env_record));
    L.sstate = off;
    off = align_up(off + E * P * 4 * rb);
    L.spos = off;
    off = align_up(off + E * P * 8);
    L.sgen = off;
    off = align_up(off + E * P * rb);
    L.ssf = off;
    off = align_up(off + E * P);
    L.stat_sums = off;
    off = align_up(off + UAVSIM_STAT_SUMS * 8);
    L.error_word = off;
    off = align_up(off + 8);
    L.actions = off;
    off = align_up(off + E * 4);
    L.total = off;
    return L;
}

template <typename Real>
Params<Real> make_params(const uavsim_config& c) {
    Params<Real> p;
    std::memset(&p, 0, sizeof p);
    p.rate = (Real)c.data_generation_rate;
    p.maxbuf = (Real)c.max_buffer_size;
    // float32 build: RSSI in dB above the range threshold (see Params::rssi_shift)
    const double shift = std::is_same<Real, double>::value ? 0.0 : c.rssi_threshold;
    p.rssi_shift = (Real)shift;
    p.thr = (Real)(c.rssi_threshold - shift);
    p.adr_t60 = (Real)(-60.0 - shift);  // iot_sensors.py:32-37
    p.adr_t70 = (Real)(-70.0 - shift);
    p.adr_t78 = (Real)(-78.0 - shift);
    p.adr_t85 = (Real)(-85.0 - shift);
    p.duty_frac = (Real)(c.duty_cycle / 100.0);  // IoTSensor.duty_cycle_probability, iot_sensors.py:104-106
    p.cd = (Real)c.collection_duration;
    p.tx_power = (Real)(c.tx_power_dbm - shift);
    p.noise_floor = (Real)(c.noise_floor_dbm - shift);
    p.lambda = (Real)c.adr_lambda;
    p.one_minus_lambda = (Real)(1 - c.adr_lambda);
    p.sigma = (Real)c.shadowing_std_db;
    p.capture_db = (Real)c.capture_threshold_db;
    p.c_fspl = (Real)(20 * std::log10(868.0));  // iot_sensors.py:179
    p.c_ht = (Real)(20 * std::log10(0.5));      // iot_sensors.py:183
    p.c_hr = (Real)(20 * std::log10(c.altitude_m));
    p.d_break = (4 * M_PI * 0.5 * c.altitude_m) / 0.345;  // iot_sensors.py:174
    p.det_near = (float)((c.tx_power_dbm - shift) - 20 * std::log10(868.0) + 28.0);                       // :179-180
    p.det_far = (float)((c.tx_power_dbm - shift) + 20 * std::log10(0.5) + 20 * std::log10(c.altitude_m));  // :183-185
    p.alt2 = (float)((long)c.altitude_m * (long)c.altitude_m);
    p.d_break_sq = (float)(p.d_break * p.d_break);
    p.r_per_byte = (Real)c.reward_per_byte;
    p.r_new = (Real)c.reward_new_sensor;
    p.r_completion = (Real)c.reward_completion;
    p.r_urg_red = (Real)c.reward_urgency_reduction;
    p.r_move = (Real)c.reward_movement;
    p.p_revisit = (Real)c.penalty_revisit;
    p.p_boundary = (Real)c.penalty_boundary;
    p.p_collision = (Real)c.penalty_collision;
    p.p_battery = (Real)c.penalty_battery;
    p.p_hover = (Real)c.penalty_hover;
    p.p_step = (Real)c.penalty_step;
    p.p_data_loss = (Real)c.penalty_data_loss;
    p.p_starvation = (Real)c.penalty_starvation;
    p.p_unvisited = (Real)c.penalty_unvisited;
    p.p_starved = (Real)c.penalty_starved;
    p.cr_thr = (Real)c.starvation_cr_threshold;
    p.prox_eta = (Real)c.prox_eta;
    p.jain_bonus = (Real)c.jain_bonus;
    p.relational = c.wrapper == UAVSIM_WRAP_RELATIONAL;
    p.rel_dwell_steps = c.rel_dwell_steps;
    p.rel_min_cluster = c.rel_min_cluster;
    p.rel_gamma = (Real)c.rel_gamma;
    p.rel_dwell_bonus = (Real)c.rel_dwell_bonus;
    p.rel_scale = (Real)c.rel_reward_scale;
    p.gps_sigma = (Real)c.gps_sigma;
    p.gps_in_obs = c.gps_in_observation;
    p.wind_drift = c.wind_drift;
    p.max_battery = c.max_battery;
    p.batt_floor = 0.02 * c.max_battery;                       // uav.py:224
    p.e_move = (c.power_move_w * 1) / 3600;                    // uav.py:177-178
    p.e_fail = ((c.power_move_w * 0.5) * 1) / 3600;            // uav.py:125,182-183
    p.e_hover = (c.power_hover_w * c.collection_duration) / 3600;          // uav.py:205
    p.used_hover = (c.power_hover_w / (60 * 60)) * c.collection_duration;  // uav.py:263, uav_env.py:530
    p.inv_max_battery = (float)(1.0 / c.max_battery);
    p.inv_maxbuf = (float)(1.0 / c.max_buffer_size);
    p.inv_rate = (float)(1.0 / c.data_generation_rate);
    p.min_start_dist = c.min_start_dist;
    p.max_steps = c.max_steps;
    p.init_sf = c.init_sf;
    p.fps = c.features_per_sensor;
    p.obs_dim = 3 + c.features_per_sensor * c.obs_slots;
    p.frame_k = ring_frames(c);
    p.frame_window = c.frame_stack;
    p.latency = c.latency_steps;
    p.frame_mirror = p.frame_k > 1 ? p.frame_k * p.obs_dim : 0;
    p.obs_stride = p.obs_dim * (p.frame_k > 1 ? 2 * p.frame_k : 1);
    p.P = c.max_sensors;
    p.domain_rand = c.reset_mode == UAVSIM_RESET_DOMAIN_RAND;
    p.auto_reset = c.auto_reset != 0;
    p.num_grids = c.num_grids;
    p.max_start_tries = c.max_start_tries;
    std::memcpy(p.grids, c.grids, sizeof p.grids);
    p.key0 = (uint32_t)c.seed;
    p.key1 = (uint32_t)(c.seed >> 32);
    p.keys = philox_key_schedule(p.key0, p.key1);
    p.env_id_offset = c.env_id_offset;
    return p;
}

__global__ void init_records_kernel(uavsim_env_record* rec, int E, int W, int H, int N, float sx, float sy,
                                    double battery, int init_sf, uint8_t* ssf, int P, unsigned long long* error_word) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) *error_word = ~0ull;
    if (e >= E) return;
    uavsim_env_record r;
    memset(&r, 0, sizeof r);
    r.x = r.start_x = sx;
    r.y = r.start_y = sy;
    r.battery = battery;
    r.grid_w = W;
    r.grid_h = H;
    r.num_sensors = N;
    r.first_full_coverage_step = -1;
    rec[e] = r;
    for (int j = 0; j < P; ++j) ssf[(size_t)e * P + j] = (uint8_t)init_sf;
}

__global__ void scenario_kernel(Buffers b, int P, const float* xy, const int32_t* n, const int32_t* wh, const float* st) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= b.E) return;
    uavsim_env_record r = b.rec[e];
    if (n) r.num_sensors = n[e];
    if (wh) {
        r.grid_w = wh[2 * e];
        r.grid_h = wh[2 * e + 1];
    }
    if (st) {
        r.start_x = st[2 * e];
        r.start_y = st[2 * e + 1];
        r.x = r.start_x;
        r.y = r.start_y;
    }
    b.rec[e] = r;
    if (xy)
        for (int j = 0; j < P; ++j)
            b.spos[(size_t)e * P + j] = make_float2(xy[((size_t)e * P + j) * 2], xy[((size_t)e * P + j) * 2 + 1]);
}

__global__ void check_scenario_kernel(const uavsim_env_record* rec, int E, int P, int* bad) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const uavsim_env_record r = rec[e];
    if (r.num_sensors <= 0 || r.num_sensors > P || r.grid_w <= 0 || r.grid_h <= 0) atomicAdd(bad, 1);
}

}  // namespace

struct uavsim_handle {
    uavsim_config cfg;
    uavsim_dims dims;
    Layout lay;
    char* arena = nullptr;
    bool owns_arena = false;
    unsigned* tickets = nullptr;  // tile-ticket + finished-block counters of the step kernel (deliberately NOT in the
                                  // arena: a snapshot / restore of the arena must not touch them)
    Params<double> p64;
    Params<float> p32;
    Buffers buf;
    bool outputs_bound = false;
    bool reseeded = false;  // uavsim_set_seed since the last reset: the next reset restarts the episode counters
    int frame_slot = -1;  // row the newest frame was written to; -1 before the first reset
    int64_t launches = 0;
    int32_t* actions_dev = nullptr;
    // uavsim_step_host pipeline: the copy-out of env range k runs on `copy_stream` while range k+1 is stepped
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_range[8] = {}, ev_copied = nullptr;
    int host_chunks = 4;
};

extern "C" {

int uavsim_abi_version(void) { return UAVSIM_ABI_VERSION; }

const char* uavsim_last_error(void) { return g_err.c_str(); }

int uavsim_default_config(uavsim_config* c) {
    if (!c) return fail(UAVSIM_ERR_INVALID_ARG, "config is NULL");
    std::memset(c, 0, sizeof *c);
    c->struct_size = sizeof *c;
    c->num_envs = 1;
    c->max_sensors = 20;
    c->obs_slots = 20;
    c->features_per_sensor = 3;
    c->frame_stack = 1;
    c->precision = UAVSIM_F32;
    c->rng_mode = UAVSIM_RNG_PHILOX;
    c->reset_mode = UAVSIM_RESET_FIXED;
    c->auto_reset = 1;
    c->seed = 0;
    c->env_id_offset = 0;
    c->grid_w = 500;  // evaluation grid (dqn.py:116-117)
    c->grid_h = 500;
    c->num_sensors = 20;
    c->max_steps = 2100;                   // dqn.py:1069
    c->data_generation_rate = 22.0 / 10;   // uav_env.py:271
    c->max_buffer_size = 1000.0;           // uav_env.py:272
    c->rssi_threshold = -85.0;             // dqn.py:1071
    c->duty_cycle = 10.0;                  // dqn.py:1072
    c->max_battery = 274.0;                // dqn.py:1073
    c->collection_duration = 1.0;          // uav_env.py:279
    c->start_x = 0.0;                      // uav_env.py:332-333
    c->start_y = 0.0;
    c->tx_power_dbm = 14.0;                // iot_sensors.py:46
    c->noise_floor_dbm = -105.0;           // iot_sensors.py:50
    c->altitude_m = 100.0;                 // iot_sensors.py:51
    c->adr_lambda = 0.1;                   // iot_sensors.py:54
    c->shadowing_std_db = 4.0;             // iot_sensors.py:56
    c->capture_threshold_db = 6.0;         // uav_env.py:569
    c->power_move_w = 500.0;               // uav.py:93
    c->power_hover_w = 700.0;              // uav.py:94
    c->init_sf = 12;                       // iot_sensors.py:309
    c->reward_per_byte = 100.0;            // reward_function.py:10-26 with uav_env.py:282-285
    c->reward_new_sensor = 5000.0;
    c->reward_completion = 100.0;
    c->reward_urgency_reduction = 20.0;
    c->reward_movement = 10.0;
    c->penalty_revisit = -2.0;
    c->penalty_boundary = -50.0;
    c->penalty_collision = -10.0;
    c->penalty_battery = -0.5;
    c->penalty_hover = -5.0;
    c->penalty_step = -0.5;
    c->penalty_data_loss = -1.0;
    c->penalty_starvation = -1000.0;
    c->penalty_unvisited = -5000.0;
    c->penalty_starved = -1000.0;
    c->starvation_cr_threshold = 0.20;
    c->num_grids = 1;                      // curriculum stage 0 (dqn.py:64)
    c->grids[0][0] = 100;
    c->grids[0][1] = 100;
    c->min_start_dist = 50.0;              // dqn.py:127
    c->max_start_tries = 200;              // dqn.py:128
    c->prox_eta = 2.0;                     // dqn.py:129
    c->jain_bonus = 0.5;                   // dqn.py:442
    c->wrapper = UAVSIM_WRAP_NONE;
    c->rel_gamma = 0.99;                   // env_wrapper.py:48
    c->rel_dwell_bonus = 25.0;             // env_wrapper.py:49
    c->rel_reward_scale = 1e-5;            // env_wrapper.py:62
    c->rel_dwell_steps = 10;               // env_wrapper.py:50
    c->rel_min_cluster = 2;                // env_wrapper.py:51
    c->gps_sigma = 0.0;
    c->wind_drift = 0.0;
    c->latency_steps = 0;
    c->gps_in_observation = 0;
    c->reserved2 = 0;
    return UAVSIM_OK;
}

int uavsim_query_dims(const uavsim_config* c, uavsim_dims* out) {
    if (int rc = validate(c)) return rc;
    if (!out) return fail(UAVSIM_ERR_INVALID_ARG, "out is NULL");
    std::memset(out, 0, sizeof *out);
    out->obs_dim = 3 + c->features_per_sensor * c->obs_slots;
    out->frame_rows = ring_frames(*c) > 1 ? 2 * ring_frames(*c) : 1;
    out->obs_row_floats = out->obs_dim * out->frame_rows;
    out->real_bytes = real_bytes(c);
    out->arena_bytes = make_layout(c).total;
    out->replay_step_doubles = (int64_t)UAVSIM_NUM_SLOTS * c->max_sensors;
    return UAVSIM_OK;
}

int uavsim_create(const uavsim_config* c, void* arena_dev, size_t arena_bytes, void* stream, uavsim_handle** out) {
    if (!out) return fail(UAVSIM_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (int rc = validate(c)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    uavsim_handle* h = new (std::nothrow) uavsim_handle();
    if (!h) return fail(UAVSIM_ERR_INVALID_ARG, "out of host memory");
    h->cfg = *c;
    uavsim_query_dims(c, &h->dims);
    h->lay = make_layout(c);
    if (arena_dev) {
        if ((int64_t)arena_bytes < h->lay.total || ((uintptr_t)arena_dev % kAlign) != 0) {
            delete h;
            return fail(UAVSIM_ERR_INVALID_ARG, "arena must be >= %lld bytes and %lld-byte aligned",
                        (long long)h->lay.total, (long long)kAlign);
        }
        h->arena = (char*)arena_dev;
    } else {
        cudaError_t e = cudaMalloc((void**)&h->arena, (size_t)h->lay.total);
        if (e != cudaSuccess) {
            delete h;
            return cuda_fail(e, "cudaMalloc(arena)");
        }
        h->owns_arena = true;
    }
    {
        cudaError_t e = cudaMalloc((void**)&h->tickets, 2 * sizeof(unsigned));
        if (e == cudaSuccess) e = cudaMemsetAsync(h->tickets, 0, 2 * sizeof(unsigned), st);
        if (e != cudaSuccess) {
            if (h->owns_arena) cudaFree(h->arena);
            delete h;
            return cuda_fail(e, "cudaMalloc(tile tickets)");
        }
    }
    h->p64 = make_params<double>(*c);
    h->p32 = make_params<float>(*c);
    Buffers& b = h->buf;
    std::memset(&b, 0, sizeof b);
    b.rec = (uavsim_env_record*)(h->arena + h->lay.rec);
    b.sstate = h->arena + h->lay.sstate;
    b.spos = (float2*)(h->arena + h->lay.spos);
    b.sgen = h->arena + h->lay.sgen;
    b.ssf = (uint8_t*)(h->arena + h->lay.ssf);
    b.stat_sums = (double*)(h->arena + h->lay.stat_sums);
    b.error_word = (unsigned long long*)(h->arena + h->lay.error_word);
    b.tile_ticket = h->tickets;
    b.tile = UAVSIM_TILE;
    b.E = c->num_envs;
    h->actions_dev = (int32_t*)(h->arena + h->lay.actions);
    cudaError_t e = cudaMemsetAsync(h->arena, 0, (size_t)h->lay.total, st);
    if (e == cudaSuccess) {
        const int E = c->num_envs;
        init_records_kernel<<<(E + 255) / 256, 256, 0, st>>>(b.rec, E, c->grid_w, c->grid_h, c->num_sensors,
                                                            (float)c->start_x, (float)c->start_y, c->max_battery,
                                                            c->init_sf, b.ssf, c->max_sensors, b.error_word);
        e = cudaGetLastError();
        h->launches += 1;
    }
    if (e != cudaSuccess) {
        if (h->owns_arena) cudaFree(h->arena);
        cudaFree(h->tickets);
        delete h;
        return cuda_fail(e, "arena initialisation");
    }
    *out = h;
    return UAVSIM_OK;
}

int uavsim_destroy(uavsim_handle* h) {
    if (!h) return UAVSIM_OK;
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (cudaEvent_t& ev : h->ev_range)
        if (ev) cudaEventDestroy(ev);
    if (h->ev_copied) cudaEventDestroy(h->ev_copied);
    if (h->owns_arena) cudaFree(h->arena);
    cudaFree(h->tickets);
    delete h;
    return UAVSIM_OK;
}

int uavsim_bind_outputs(uavsim_handle* h, const uavsim_outputs* o, void* stream) {
    if (!h || !o) return fail(UAVSIM_ERR_INVALID_ARG, "handle/outputs is NULL");
    if (!o->obs || !o->reward || !o->done) return fail(UAVSIM_ERR_INVALID_ARG, "obs, reward and done are required");
    Buffers& b = h->buf;
    b.obs = o->obs;
    b.reward = o->reward;
    b.done = o->done;
    b.terminal_obs = o->terminal_obs;
    b.episode_stats = o->episode_stats;
    b.action_mask = o->action_mask;
    b.step_detail = o->step_detail;
    b.draw_export = o->draw_export;
    b.terminal_info = o->terminal_info;
    b.terminal_ratios = o->terminal_ratios;
    // ghost sensor slots are written once here and never again (N is fixed per environment)
    CUDA_TRY(cudaMemsetAsync(b.obs, 0, (size_t)h->cfg.num_envs * h->dims.obs_row_floats * sizeof(float),
                             (cudaStream_t)stream));
    h->outputs_bound = true;
    return UAVSIM_OK;
}

int uavsim_get_field(uavsim_handle* h, int field, uavsim_field_view* v) {
    if (!h || !v) return fail(UAVSIM_ERR_INVALID_ARG, "handle/view is NULL");
    std::memset(v, 0, sizeof *v);
    const int64_t E = h->cfg.num_envs, P = h->cfg.max_sensors, rb = h->dims.real_bytes;
    switch (field) {
        case UAVSIM_FIELD_ENV_RECORD:
            v->ptr = h->buf.rec;
            v->elem_bytes = sizeof(uavsim_env_record);
            v->is_float = 2;
            v->ndim = 1;
            v->shape[0] = E;
            v->stride_bytes[0] = sizeof(uavsim_env_record);
            break;
        case UAVSIM_FIELD_SENSOR_STATE:
            v->ptr = h->buf.sstate;
            v->elem_bytes = (int32_t)rb;
            v->is_float = 1;
            v->ndim = 3;
            v->shape[0] = E;
            v->shape[1] = P;
            v->shape[2] = 4;
            v->stride_bytes[0] = P * 4 * rb;
            v->stride_bytes[1] = 4 * rb;
            v->stride_bytes[2] = rb;
            break;
        case UAVSIM_FIELD_SENSOR_POS:
            v->ptr = h->buf.spos;
            v->elem_bytes = 4;
            v->is_float = 1;
            v->ndim = 3;
            v->shape[0] = E;
            v->shape[1] = P;
            v->shape[2] = 2;
            v->stride_bytes[0] = P * 8;
            v->stride_bytes[1] = 8;
            v->stride_bytes[2] = 4;
            break;
        case UAVSIM_FIELD_SENSOR_GEN:
            v->ptr = h->buf.sgen;
            v->elem_bytes = (int32_t)rb;
            v->is_float = 1;
            v->ndim = 2;
            v->shape[0] = E;
            v->shape[1] = P;
            v->stride_bytes[0] = P * rb;
            v->stride_bytes[1] = rb;
            break;
        case UAVSIM_FIELD_SENSOR_SF:
            if (h->cfg.precision == UAVSIM_F32)
                return fail(UAVSIM_ERR_UNSUPPORTED,
                            "throughput layout: the spreading factor is packed into the low 3 mantissa bits of avg_rssi "
                            "(sf = (bits & 7) + 7)");
            v->ptr = h->buf.ssf;
            v->elem_bytes = 1;
            v->is_float = 0;
            v->ndim = 2;
            v->shape[0] = E;
            v->shape[1] = P;
            v->stride_bytes[0] = P;
            v->stride_bytes[1] = 1;
            break;
        default:
            return fail(UAVSIM_ERR_INVALID_ARG, "unknown field %d", field);
    }
    return UAVSIM_OK;
}

int uavsim_set_scenario(uavsim_handle* h, const uavsim_scenario* sc, void* stream) {
    if (!h || !sc) return fail(UAVSIM_ERR_INVALID_ARG, "handle/scenario is NULL");
    const int E = h->cfg.num_envs;
    cudaStream_t st = (cudaStream_t)stream;
    scenario_kernel<<<(E + 255) / 256, 256, 0, st>>>(h->buf, h->cfg.max_sensors, sc->sensor_xy, sc->num_sensors,
                                                     sc->grid_wh, sc->start_xy);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    if (sc->num_sensors || sc->grid_wh) {  // validated with one sync: scenarios are installed rarely
        int* bad = nullptr;
        CUDA_TRY(cudaMalloc((void**)&bad, sizeof(int)));
        cudaMemsetAsync(bad, 0, sizeof(int), st);
        check_scenario_kernel<<<(E + 255) / 256, 256, 0, st>>>(h->buf.rec, E, h->cfg.max_sensors, bad);
        h->launches += 1;
        int host_bad = 0;
        cudaError_t e = cudaMemcpyAsync(&host_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFree(bad);
        if (e != cudaSuccess) return cuda_fail(e, "scenario check");
        if (host_bad)
            return fail(UAVSIM_ERR_INVALID_ARG, "%d environments have num_sensors outside [1, max_sensors] or a non-positive grid",
                        host_bad);
        if (h->outputs_bound)  // N may have changed: re-zero ghost slots
            CUDA_TRY(cudaMemsetAsync(h->buf.obs, 0, (size_t)E * h->dims.obs_row_floats * sizeof(float), st));
    }
    return UAVSIM_OK;
}

int uavsim_generate_layouts(uavsim_handle* h, void* stream) {
    if (!h) return fail(UAVSIM_ERR_INVALID_ARG, "handle is NULL");
    const int E = h->cfg.num_envs;
    layout_kernel<<<(E + kWarpsPerBlock - 1) / kWarpsPerBlock, kThreads, 0, (cudaStream_t)stream>>>(
        h->buf, h->cfg.max_sensors, (uint32_t)h->cfg.seed, (uint32_t)(h->cfg.seed >> 32), h->cfg.env_id_offset);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return UAVSIM_OK;
}

int uavsim_set_replay(uavsim_handle* h, const uavsim_replay* rp) {
    if (!h || !rp) return fail(UAVSIM_ERR_INVALID_ARG, "handle/replay is NULL");
    if (h->cfg.rng_mode != UAVSIM_RNG_REPLAY) return fail(UAVSIM_ERR_STATE, "handle was not created with UAVSIM_RNG_REPLAY");
    Buffers& b = h->buf;
    b.step_draws = rp->step_draws;
    b.reset_draws = rp->reset_draws;
    b.reset_fill = rp->reset_fill;
    b.reset_init_sf = rp->reset_init_sf;
    b.rs_sensor_xy = rp->reset_scenario.sensor_xy;
    b.rs_num_sensors = rp->reset_scenario.num_sensors;
    b.rs_grid_wh = rp->reset_scenario.grid_wh;
    b.rs_start_xy = rp->reset_scenario.start_xy;
    b.dwell_draws = rp->dwell_draws;
    b.step_env_draws = rp->step_env_draws;
    b.reset_env_draws = rp->reset_env_draws;
    return UAVSIM_OK;
}

int uavsim_set_stage_grids(uavsim_handle* h, int32_t n, const int32_t* wh) {
    if (!h || !wh) return fail(UAVSIM_ERR_INVALID_ARG, "handle/grids is NULL");
    if (n <= 0 || n > UAVSIM_MAX_GRIDS) return fail(UAVSIM_ERR_INVALID_ARG, "num_grids must be in [1, %d]", UAVSIM_MAX_GRIDS);
    for (int i = 0; i < n; ++i)
        if (wh[2 * i] <= 0 || wh[2 * i + 1] <= 0) return fail(UAVSIM_ERR_INVALID_ARG, "stage grid %d is not positive", i);
    h->cfg.num_grids = n;
    for (int i = 0; i < n; ++i) {
        h->cfg.grids[i][0] = wh[2 * i];
        h->cfg.grids[i][1] = wh[2 * i + 1];
    }
    h->p64.num_grids = h->p32.num_grids = n;
    std::memcpy(h->p64.grids, h->cfg.grids, sizeof h->p64.grids);
    std::memcpy(h->p32.grids, h->cfg.grids, sizeof h->p32.grids);
    return UAVSIM_OK;
}

static int check_ready(uavsim_handle* h, bool reset) {
    if (!h) return fail(UAVSIM_ERR_INVALID_ARG, "handle is NULL");
    if (!h->outputs_bound) return fail(UAVSIM_ERR_STATE, "uavsim_bind_outputs has not been called");
    if (h->cfg.rng_mode == UAVSIM_RNG_REPLAY) {
        const Buffers& b = h->buf;
        const bool need_reset_tables = reset || h->cfg.auto_reset;
        if (!reset && !b.step_draws) return fail(UAVSIM_ERR_STATE, "replay mode: step_draws not set (uavsim_set_replay)");
        if (need_reset_tables && !b.reset_draws) return fail(UAVSIM_ERR_STATE, "replay mode: reset_draws not set");
        if (need_reset_tables && h->cfg.reset_mode == UAVSIM_RESET_FIXED && !b.reset_fill)
            return fail(UAVSIM_ERR_STATE, "replay mode: reset_fill not set");
    }
    return UAVSIM_OK;
}

int uavsim_reset(uavsim_handle* h, const uint8_t* mask, void* stream) {
    if (int rc = check_ready(h, true)) return rc;
    const bool replay = h->cfg.rng_mode == UAVSIM_RNG_REPLAY;
    const int k = ring_frames(h->cfg);
    if (!mask || h->frame_slot < 0) h->frame_slot = k - 1;  // a full reset restarts the ring: window = k-1 zero frames + obs
    h->buf.frame_slot = h->frame_slot;
    h->buf.zero_episode = h->reseeded ? 1 : 0;  // the envs this reset covers restart their episode counters
    h->reseeded = false;
    cudaError_t e = h->cfg.precision == UAVSIM_F64 ? launch_reset_f64(h->p64, h->buf, mask, replay, (cudaStream_t)stream)
                                                   : launch_reset_f32(h->p32, h->buf, mask, replay, (cudaStream_t)stream);
    h->buf.zero_episode = 0;
    if (e != cudaSuccess) return cuda_fail(e, "reset_kernel launch");
    h->launches += 1;
    return UAVSIM_OK;
}

// The buffers of the env range [e0, e0 + n): every per-env pointer advanced by e0.  Environments are independent and
// their Philox streams are keyed by the GLOBAL env id (env_id_offset + e0 here), so stepping a batch as several ranges is
// bit-identical to stepping it in one launch (tests/test_gpu_api.py).
static Buffers range_buffers(const uavsim_handle* h, int e0, int n) {
    Buffers b = h->buf;
    const size_t P = (size_t)h->cfg.max_sensors, rb = (size_t)h->dims.real_bytes, k = (size_t)e0;
    b.E = n;
    b.rec += k;
    b.sstate = (char*)b.sstate + k * P * 4 * rb;
    b.spos += k * P;
    b.sgen = (char*)b.sgen + k * P * rb;
    b.ssf += k * P;
    b.obs += k * (size_t)h->dims.obs_row_floats;
    b.reward = (char*)b.reward + k * rb;
    b.done += k;
    if (b.terminal_obs) b.terminal_obs += k * (size_t)h->cfg.frame_stack * (size_t)h->dims.obs_dim;
    if (b.episode_stats) b.episode_stats += k * UAVSIM_STAT_COUNT;
    if (b.action_mask) b.action_mask += k;
    if (b.step_detail) b.step_detail += k * UAVSIM_DETAIL_COUNT;
    if (b.draw_export) b.draw_export = (char*)b.draw_export + k * UAVSIM_NUM_SLOTS * P * rb;
    if (b.terminal_info) b.terminal_info += k * UAVSIM_INFO_COUNT;
    if (b.terminal_ratios) b.terminal_ratios += k * P;
    if (b.step_draws) b.step_draws += k * UAVSIM_NUM_SLOTS * P;
    if (b.reset_draws) b.reset_draws += k * UAVSIM_NUM_SLOTS * P;
    if (b.reset_fill) b.reset_fill += k * P;
    if (b.reset_init_sf) b.reset_init_sf += k;
    if (b.rs_sensor_xy) b.rs_sensor_xy += k * P * 2;
    if (b.rs_num_sensors) b.rs_num_sensors += k;
    if (b.rs_grid_wh) b.rs_grid_wh += k * 2;
    if (b.rs_start_xy) b.rs_start_xy += k * 2;
    if (b.dwell_draws) b.dwell_draws += k * P;
    if (b.step_env_draws) b.step_env_draws += k * 4;
    if (b.reset_env_draws) b.reset_env_draws += k * 4;
    if (b.env_sigma) b.env_sigma += k;
    if (b.env_duty) b.env_duty += k;
    return b;
}

static int step_launches(const uavsim_handle* h) {
    const bool replay = h->cfg.rng_mode == UAVSIM_RNG_REPLAY;
    return h->cfg.precision == UAVSIM_F64 ? step_launch_count(h->p64, h->buf, replay) : step_launch_count(h->p32, h->buf, replay);
}

static cudaError_t launch_step_range(uavsim_handle* h, int e0, int n, const int32_t* actions, cudaStream_t st) {
    const bool replay = h->cfg.rng_mode == UAVSIM_RNG_REPLAY;
    const Buffers b = range_buffers(h, e0, n);
    if (h->cfg.precision == UAVSIM_F64) {
        Params<double> p = h->p64;
        p.env_id_offset += e0;
        return launch_step_f64(p, b, actions + e0, replay, st);
    }
    Params<float> p = h->p32;
    p.env_id_offset += e0;
    return launch_step_f32(p, b, actions + e0, replay, st);
}

int uavsim_step(uavsim_handle* h, const int32_t* actions, void* stream) {
    if (int rc = check_ready(h, false)) return rc;
    if (!actions) return fail(UAVSIM_ERR_INVALID_ARG, "actions is NULL");
    const bool replay = h->cfg.rng_mode == UAVSIM_RNG_REPLAY;
    if (h->frame_slot < 0) return fail(UAVSIM_ERR_STATE, "uavsim_reset has not been called");
    h->buf.step_serial += 1;
    h->frame_slot = (h->frame_slot + 1) % ring_frames(h->cfg);
    h->buf.frame_slot = h->frame_slot;
    cudaError_t e = h->cfg.precision == UAVSIM_F64 ? launch_step_f64(h->p64, h->buf, actions, replay, (cudaStream_t)stream)
                                                   : launch_step_f32(h->p32, h->buf, actions, replay, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "step_kernel launch");
    h->launches += step_launches(h);
    return UAVSIM_OK;
}

int uavsim_step_ex(uavsim_handle* h, const int32_t* actions, const uint8_t* env_mask, uint32_t flags, void* stream) {
    if (int rc = check_ready(h, false)) return rc;
    if (!actions) return fail(UAVSIM_ERR_INVALID_ARG, "actions is NULL");
    if (flags & ~UAVSIM_STEP_EXTRA) return fail(UAVSIM_ERR_INVALID_ARG, "unknown step flags");
    if (h->frame_slot < 0) return fail(UAVSIM_ERR_STATE, "uavsim_reset has not been called");
    const bool replay = h->cfg.rng_mode == UAVSIM_RNG_REPLAY, extra = (flags & UAVSIM_STEP_EXTRA) != 0;
    h->buf.step_serial += 1;
    const int next_slot = (h->frame_slot + 1) % ring_frames(h->cfg);
    if (!extra) h->frame_slot = next_slot;  // a dropped repeat writes the row the commanded step is about to overwrite
    h->buf.frame_slot = next_slot;
    h->buf.step_mask = env_mask;
    h->buf.extra_step = extra ? 1 : 0;
    cudaError_t e = h->cfg.precision == UAVSIM_F64 ? launch_step_f64(h->p64, h->buf, actions, replay, (cudaStream_t)stream)
                                                   : launch_step_f32(h->p32, h->buf, actions, replay, (cudaStream_t)stream);
    const int n = step_launches(h);
    h->buf.step_mask = nullptr;
    h->buf.extra_step = 0;
    h->buf.frame_slot = h->frame_slot;
    if (e != cudaSuccess) return cuda_fail(e, "step_kernel launch");
    h->launches += n;
    return UAVSIM_OK;
}

int uavsim_set_env_params(uavsim_handle* h, const float* sigma_dev, const float* duty_dev) {
    if (!h) return fail(UAVSIM_ERR_INVALID_ARG, "handle is NULL");
    h->buf.env_sigma = sigma_dev;
    h->buf.env_duty = duty_dev;
    return UAVSIM_OK;
}

int uavsim_policy_nearest_greedy(uavsim_handle* h, int32_t* actions_dev, const double* probe_draws_dev, void* stream) {
    if (!h || !actions_dev) return fail(UAVSIM_ERR_INVALID_ARG, "handle/actions is NULL");
    const bool replay = h->cfg.rng_mode == UAVSIM_RNG_REPLAY;
    if (replay && !probe_draws_dev) return fail(UAVSIM_ERR_STATE, "replay mode: probe_draws is required");
    if (h->frame_slot < 0) return fail(UAVSIM_ERR_STATE, "uavsim_reset has not been called");
    cudaError_t e = h->cfg.precision == UAVSIM_F64
                        ? launch_nearest_greedy_f64(h->p64, h->buf, probe_draws_dev, actions_dev, replay, (cudaStream_t)stream)
                        : launch_nearest_greedy_f32(h->p32, h->buf, probe_draws_dev, actions_dev, replay, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "nearest_greedy_kernel launch");
    h->launches += 1;
    return UAVSIM_OK;
}

int uavsim_policy_max_throughput_greedy(uavsim_handle* h, int32_t* actions_dev, const double* probe_draws_dev,
                                        void* stream) {
    if (!h || !actions_dev) return fail(UAVSIM_ERR_INVALID_ARG, "handle/actions is NULL");
    const bool replay = h->cfg.rng_mode == UAVSIM_RNG_REPLAY;
    if (replay && !probe_draws_dev) return fail(UAVSIM_ERR_STATE, "replay mode: probe_draws is required");
    if (h->frame_slot < 0) return fail(UAVSIM_ERR_STATE, "uavsim_reset has not been called");
    cudaError_t e = h->cfg.precision == UAVSIM_F64
                        ? launch_max_throughput_greedy_f64(h->p64, h->buf, probe_draws_dev, actions_dev, replay, (cudaStream_t)stream)
                        : launch_max_throughput_greedy_f32(h->p32, h->buf, probe_draws_dev, actions_dev, replay, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "max_throughput_greedy_kernel launch");
    h->launches += 1;
    return UAVSIM_OK;
}

int uavsim_step_host(uavsim_handle* h, const int32_t* actions_host, float* obs_host, void* reward_host,
                     uint8_t* done_host, void* stream) {
    if (!h || !actions_host || !reward_host || !done_host)
        return fail(UAVSIM_ERR_INVALID_ARG, "handle/actions/reward/done is NULL");
    if (int rc = check_ready(h, false)) return rc;
    if (h->frame_slot < 0) return fail(UAVSIM_ERR_STATE, "uavsim_reset has not been called");
    cudaStream_t st = (cudaStream_t)stream;
    const int E = h->cfg.num_envs;
    const size_t rb = (size_t)h->dims.real_bytes, od = (size_t)h->dims.obs_dim, k = (size_t)h->cfg.frame_stack;
    // Ranges of a multiple of 64 environments (tile- and pair-aligned), at least 16 Ki each: below that the copies are
    // latency-bound and one launch is as good.
    int chunks = h->host_chunks;
    while (chunks > 1 && E / chunks < 16384) chunks >>= 1;
    const int per = ((E + chunks - 1) / chunks + 63) / 64 * 64;
    if (chunks > 1 && !h->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (cudaEvent_t& ev : h->ev_range) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaMemcpyAsync(h->actions_dev, actions_host, (size_t)E * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    h->buf.step_serial += 1;
    h->frame_slot = (h->frame_slot + 1) % ring_frames(h->cfg);
    h->buf.frame_slot = h->frame_slot;
    int c = 0;
    for (int e0 = 0; e0 < E; e0 += per, ++c) {
        const int n = E - e0 < per ? E - e0 : per;
        cudaError_t le = launch_step_range(h, e0, n, h->actions_dev, st);
        if (le != cudaSuccess) return cuda_fail(le, "step_kernel launch");
        h->launches += step_launches(h);
        cudaStream_t cs = st;
        if (chunks > 1) {  // the copy-out of this range overlaps the step of the next one
            cs = h->copy_stream;
            CUDA_TRY(cudaEventRecord(h->ev_range[c], st));
            CUDA_TRY(cudaStreamWaitEvent(cs, h->ev_range[c], 0));
        }
        if (obs_host) {  // the stacked window [n][k * obs_dim] (a strided slice of the mirrored ring when K > 1)
            const float* src = h->buf.obs + (size_t)e0 * (size_t)h->dims.obs_row_floats;
            float* dst = obs_host + (size_t)e0 * k * od;
            if (ring_frames(h->cfg) == 1)
                CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)n * od * sizeof(float), cudaMemcpyDeviceToHost, cs));
            else
                CUDA_TRY(cudaMemcpy2DAsync(dst, k * od * sizeof(float), src + (size_t)(h->frame_slot + 1) * od,
                                           (size_t)h->dims.obs_row_floats * sizeof(float), k * od * sizeof(float), (size_t)n,
                                           cudaMemcpyDeviceToHost, cs));
        }
        CUDA_TRY(cudaMemcpyAsync((char*)reward_host + (size_t)e0 * rb, (const char*)h->buf.reward + (size_t)e0 * rb, (size_t)n * rb,
                                 cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(done_host + e0, h->buf.done + e0, (size_t)n, cudaMemcpyDeviceToHost, cs));
    }
    if (chunks > 1) {  // the caller's stream sees the copies as done, then the host waits for both
        CUDA_TRY(cudaEventRecord(h->ev_copied, h->copy_stream));
        CUDA_TRY(cudaStreamWaitEvent(st, h->ev_copied, 0));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return UAVSIM_OK;
}

int uavsim_set_host_chunks(uavsim_handle* h, int32_t chunks) {
    if (!h) return fail(UAVSIM_ERR_INVALID_ARG, "handle is NULL");
    if (chunks != 1 && chunks != 2 && chunks != 4 && chunks != 8) return fail(UAVSIM_ERR_INVALID_ARG, "chunks must be 1, 2, 4 or 8");
    h->host_chunks = chunks;
    return UAVSIM_OK;
}

int uavsim_get_info(uavsim_handle* h, double* info_dev, double* ratios_dev, void* stream) {
    if (!h || !info_dev) return fail(UAVSIM_ERR_INVALID_ARG, "handle/info is NULL");
    if (h->frame_slot < 0) return fail(UAVSIM_ERR_STATE, "uavsim_reset has not been called");
    cudaError_t e = h->cfg.precision == UAVSIM_F64 ? launch_info_f64(h->p64, h->buf, info_dev, ratios_dev, (cudaStream_t)stream)
                                                   : launch_info_f32(h->p32, h->buf, info_dev, ratios_dev, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "info_kernel launch");
    h->launches += 1;
    return UAVSIM_OK;
}

int uavsim_set_seed(uavsim_handle* h, uint64_t seed) {
    if (!h) return fail(UAVSIM_ERR_INVALID_ARG, "handle is NULL");
    h->cfg.seed = seed;
    h->p64.key0 = h->p32.key0 = (uint32_t)seed;
    h->p64.key1 = h->p32.key1 = (uint32_t)(seed >> 32);
    h->p64.keys = h->p32.keys = philox_key_schedule((uint32_t)seed, (uint32_t)(seed >> 32));
    h->reseeded = true;
    return UAVSIM_OK;
}

int uavsim_set_frame_slot(uavsim_handle* h, int32_t slot) {
    if (!h) return fail(UAVSIM_ERR_INVALID_ARG, "handle is NULL");
    if (slot < -1 || slot >= ring_frames(h->cfg)) return fail(UAVSIM_ERR_INVALID_ARG, "frame slot must be in [-1, %d)", ring_frames(h->cfg));
    h->frame_slot = slot;
    h->buf.frame_slot = slot;
    return UAVSIM_OK;
}

int uavsim_poll_error(uavsim_handle* h, void* stream, int64_t* first_env) {
    if (!h) return fail(UAVSIM_ERR_INVALID_ARG, "handle is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long w = ~0ull;
    CUDA_TRY(cudaMemcpyAsync(&w, h->buf.error_word, sizeof w, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (w == ~0ull) return UAVSIM_OK;
    CUDA_TRY(cudaMemsetAsync(h->buf.error_word, 0xff, sizeof w, st));
    if (first_env) *first_env = (int64_t)(w & 0xffffffffull);
    return fail(UAVSIM_ERR_INVALID_ACTION, "Invalid action for env %llu (step call %llu): actions must be in [0, 4]",
                w & 0xffffffffull, w >> 32);
}

int uavsim_stat_sums(uavsim_handle* h, double** sums) {
    if (!h || !sums) return fail(UAVSIM_ERR_INVALID_ARG, "handle/sums is NULL");
    *sums = h->buf.stat_sums;
    return UAVSIM_OK;
}

int64_t uavsim_launch_count(const uavsim_handle* h) { return h ? h->launches : 0; }

int uavsim_frame_slot(const uavsim_handle* h) { return h ? h->frame_slot : -1; }

}  // extern "C"
