/*
 * uavsim.h -- C ABI of the B200-native batched UAV-LoRa environment step.
 *
 * This is the drop-in boundary for ONE path of the reference project
 * (atiladeokegab/...-UAV-...-IoT-Sensor-Networks): the Gymnasium environment
 *     src/environment/uav_env.py::UAVEnvironment.reset / .step   (:400-427, :429-488)
 * together with the objects it drives
 *     src/environment/uav.py::UAV                 (:63-263)
 *     src/environment/iot_sensors.py::IoTSensor   (:10-321)
 *     src/rewards/reward_function.py::RewardFunction (:4-128)
 * and the per-env training wrapper
 *     src/agents/dqn/dqn.py::DomainRandEnv        (:177-451)
 * as consumed through stable-baselines3's VecEnv (dqn.py:1276-1278).
 *
 * The reference is pure Python and has no FFI of its own; the binding a
 * maintainer would add is the ctypes loader shown in INTEGRATION.md (and shipped
 * as <package>/_native.py).  Everything here is plain C: pointers, sizes and
 * POD structs, no C++ or torch types.  All device work is stream-ordered on the
 * stream the caller passes (a cudaStream_t as void*); no entry point
 * synchronises unless its comment says so.  One host thread per handle.  The
 * device is the one current (cudaSetDevice) when uavsim_create is called.
 *
 * Every function returns UAVSIM_OK (0) or a negative uavsim_status;
 * uavsim_last_error() gives the message for the calling thread.
 */
#ifndef UAVSIM_H_
#define UAVSIM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UAVSIM_ABI_VERSION 2
#define UAVSIM_MAX_SENSORS 64 /* visited set is one 64-bit word; the reference tops out at 50 (dqn.py:121) */
#define UAVSIM_MAX_GRIDS 8    /* grids per curriculum stage (dqn.py:63-69 uses <= 5) */
#define UAVSIM_NUM_SLOTS 6    /* replay draw slots A,B,U,C,D,E */
#define UAVSIM_NUM_ACTIONS 5  /* uav_env.py:347  Discrete(5): UP, DOWN, LEFT, RIGHT, COLLECT */

typedef enum uavsim_status {
    UAVSIM_OK = 0,
    UAVSIM_ERR_INVALID_ARG = -1,
    UAVSIM_ERR_CUDA = -2,
    UAVSIM_ERR_STATE = -3,          /* call order (e.g. step before outputs are bound) */
    UAVSIM_ERR_INVALID_ACTION = -4, /* an action outside [0,4] was seen (uav_env.py:468 raises ValueError) */
    UAVSIM_ERR_UNSUPPORTED = -5
} uavsim_status;

typedef enum uavsim_precision {
    UAVSIM_F64 = 0, /* parity layout: float64 accumulators like the reference, IEEE op order, no FMA contraction */
    UAVSIM_F32 = 1  /* throughput layout: float32 state, fast intrinsics */
} uavsim_precision;

typedef enum uavsim_rng_mode {
    UAVSIM_RNG_REPLAY = 0, /* every random draw is read from a slot-addressed table bound with uavsim_set_replay */
    UAVSIM_RNG_PHILOX = 1  /* counter-based Philox4x32-10 keyed by (seed; global env id, episode, step, sensor, slot) */
} uavsim_rng_mode;

typedef enum uavsim_reset_mode {
    UAVSIM_RESET_FIXED = 0,      /* UAVEnvironment: layout + start fixed at construction, buffers refilled U(0.2,0.6) (uav_env.py:400-427) */
    UAVSIM_RESET_DOMAIN_RAND = 1 /* DomainRandEnv: new grid from the stage list, new uniform layout, empty fresh sensors,
                                    far-start rejection sampling, proximity + Jain shaping (dqn.py:301-444) */
} uavsim_reset_mode;

typedef enum uavsim_wrapper {
    UAVSIM_WRAP_NONE = 0,
    UAVSIM_WRAP_RELATIONAL = 1 /* RelationalUAVEnv (env_wrapper.py:110-313): needs features_per_sensor == 5; the observation
                                  row becomes uav(3) + obs_slots x [rel_x, rel_y, buffer, urgency, visited] (the Dict's
                                  "uav" and "sensors" entries back to back; "mask" is slot < N), and the reward
                                  (base + potential shaping + cluster dwell bonus) * rel_reward_scale */
} uavsim_wrapper;

/* Replay draw slots.  A/B/U are consumed per sensor with data on a COLLECT
 * step (ADR draw, link-probability draw, duty-cycle uniform), C by each SF-group
 * winner (range check in collect_data), D/E per sensor by every observation
 * (ADR draw, range draw).  Normal slots hold the shadowing value itself, i.e.
 * the result of np.random.normal(0, shadowing_std_db) (iot_sensors.py:192);
 * U holds random.random() (uav_env.py:549). */
enum { UAVSIM_SLOT_A = 0, UAVSIM_SLOT_B = 1, UAVSIM_SLOT_U = 2, UAVSIM_SLOT_C = 3, UAVSIM_SLOT_D = 4, UAVSIM_SLOT_E = 5 };

typedef struct uavsim_config {
    uint32_t struct_size; /* = sizeof(uavsim_config); checked by uavsim_create */
    int32_t num_envs;     /* E, environments owned by this handle (one device) */
    int32_t max_sensors;  /* P, sensor slots per env row (>= every env's N, <= UAVSIM_MAX_SENSORS) */
    int32_t obs_slots;    /* sensor slots in the observation row (zero padded; 50 for the attention extractor, dqn.py:121,567-571) */
    int32_t features_per_sensor; /* 3, or 5 with include_sensor_positions (uav_env.py:349,668-672) */
    int32_t frame_stack;  /* k frames kept per env, newest last (VecFrameStack, dqn.py:1278; k=4 training, k=10 extractor
                             layout dqn.py:570,612-620); 1 = none.  See uavsim_dims.frame_rows. */
    int32_t precision;    /* uavsim_precision */
    int32_t rng_mode;     /* uavsim_rng_mode */
    int32_t reset_mode;   /* uavsim_reset_mode */
    int32_t auto_reset;   /* 1: SB3 VecEnv semantics -- a finished env is reset inside the same step call,
                             obs holds the new episode's first observation and terminal_obs the last one */
    uint64_t seed;        /* Philox key */
    int64_t env_id_offset; /* global id of local env 0: RNG streams do not depend on how envs are sharded */

    /* --- environment constructor arguments (uav_env.py:266-287) --- */
    int32_t grid_w, grid_h;      /* default grid for every env (FIXED mode); per-env override via uavsim_set_scenario */
    int32_t num_sensors;         /* default N for every env; per-env override via uavsim_set_scenario */
    int32_t max_steps;           /* 2100 on the training/eval path (dqn.py:1069) */
    double data_generation_rate; /* 2.2 B/s */
    double max_buffer_size;      /* 1000 B */
    double rssi_threshold;       /* -85 dBm */
    double duty_cycle;           /* percent, 10 */
    double max_battery;          /* 274 Wh */
    double collection_duration;  /* 1 s */
    double start_x, start_y;     /* uav_start_position, default (0,0) (uav_env.py:332-333) */

    /* --- IoTSensor / UAV constants (iot_sensors.py:39-57, uav.py:88-125) --- */
    double tx_power_dbm;         /* 14 */
    double noise_floor_dbm;      /* -105 */
    double altitude_m;           /* 100 */
    double adr_lambda;           /* 0.1 */
    double shadowing_std_db;     /* 4 */
    double capture_threshold_db; /* 6 (uav_env.py:569) */
    double power_move_w;         /* 500 */
    double power_hover_w;        /* 700 */
    int32_t init_sf;             /* 12: IoTSensor.reset forces SF12 (iot_sensors.py:309) */
    int32_t wrapper;             /* uavsim_wrapper: observation / reward wrapper applied on top of the step */

    /* --- RewardFunction (reward_function.py:7-44 with the overrides of uav_env.py:339-344) --- */
    double reward_per_byte, reward_new_sensor, reward_completion, reward_urgency_reduction, reward_movement;
    double penalty_revisit, penalty_boundary, penalty_collision, penalty_battery, penalty_hover, penalty_step;
    double penalty_data_loss, penalty_starvation, penalty_unvisited, penalty_starved, starvation_cr_threshold;

    /* --- DomainRandEnv (dqn.py:63-69, 126-130, 415-444) --- */
    int32_t num_grids;                   /* grids of the current curriculum stage */
    int32_t grids[UAVSIM_MAX_GRIDS][2];
    double min_start_dist;               /* 50 */
    int32_t max_start_tries;             /* 200 */
    int32_t gps_in_observation;          /* with gps_sigma > 0: 0 = GPSNoisyEnv (noise added to the normalised obs[0], obs[1],
                                            sim_to_real_analysis.py:176-181); 1 = PerturbedDQNEnv of the sweep
                                            (sim_to_real_sweep.py:121-136): the noise, in cells, moves the POSITION the
                                            observation pass sees -- clipped to the grid; obs[0], obs[1] report it and the
                                            pass's ADR update / range check run at it; the true position is untouched.
                                            Needs reset_mode FIXED and no wrapper. */
    double prox_eta;                     /* 2.0 */
    double jain_bonus;                   /* 0.5: reward += jain_bonus*(J-0.5)/N */

    /* --- RelationalUAVEnv (experiments/relational_policy/env_wrapper.py:47-62), wrapper == UAVSIM_WRAP_RELATIONAL --- */
    double rel_gamma;                    /* 0.99: shaping = gamma*Phi(s') - Phi(s), Phi = -dist(nearest unvisited)/diagonal */
    double rel_dwell_bonus;              /* 25: cluster dwell bonus at dwell step 0, decaying linearly */
    double rel_reward_scale;             /* 1e-5 */
    int32_t rel_dwell_steps;             /* 10 */
    int32_t rel_min_cluster;             /* 2 in-range sensors */

    /* --- sim-to-real evaluation wrappers (dqn_evaluation_results/sim_to_real_analysis.py:170-224), 0 = off --- */
    double gps_sigma;                    /* GPSNoisyEnv: obs[0], obs[1] += N(0, gps_sigma) every step and at reset (:176-181) */
    double wind_drift;                   /* WindDriftEnv: per-episode drift vector of this length (angle ~ U(0, 2 pi) after
                                            reset); after every step the true position moves by it, clipped to the grid, and
                                            obs[0], obs[1] report the raw drifted position (:205-224) */
    int32_t latency_steps;               /* LatencyEnv: the observation handed out is the one from this many steps ago, the
                                            FIFO primed with copies of the reset observation (:227-245).  On the device the
                                            frame ring simply holds frame_stack + latency_steps frames and the stacked
                                            window is its OLDEST frame_stack frames. */
    int32_t reserved2;
} uavsim_config;

/* State columns visible through uavsim_get_field (structure-of-arrays in HBM). */
typedef enum uavsim_field {
    UAVSIM_FIELD_ENV_RECORD = 0,  /* [E] uavsim_env_record (128 B each) */
    UAVSIM_FIELD_SENSOR_STATE = 1,/* [E][P][4] real: buffer, transmitted, lost, avg_rssi (NaN = None).  F32 layout: avg_rssi is
                                     stored as dB ABOVE config.rssi_threshold (avg_rssi - rssi_threshold, so that float32
                                     resolves the ADR / range thresholds to <= 2e-6 dB) with the spreading-factor code
                                     (sf - 7) in its three low mantissa bits */
    UAVSIM_FIELD_SENSOR_POS = 2,  /* [E][P][2] float32 */
    UAVSIM_FIELD_SENSOR_GEN = 3,  /* [E][P] real: F64: total_data_generated; F32: its value at reset (generated = this + rate*elapsed) */
    UAVSIM_FIELD_SENSOR_SF = 4,   /* [E][P] uint8 spreading factor 7..12 */
    UAVSIM_FIELD_COUNT = 5
} uavsim_field;

/* Per-environment scalars, one 128-byte record per env (same in both precisions). */
typedef struct uavsim_env_record {
    float x, y;                 /* uav.position (float32, uav.py:113) */
    float start_x, start_y;     /* uav.start_position */
    double battery;             /* Wh */
    double total_reward;        /* env.total_reward (uav_env.py:487; base reward, no wrapper shaping) */
    double total_data_collected;
    double episode_return;      /* what Monitor would report: base reward + DomainRandEnv shaping */
    float last_step_bytes;
    float prev_dist_nearest;    /* DomainRandEnv._prev_dist_nearest; relational wrapper: distance behind _prev_potential */
    float elapsed_s;            /* sum of step durations this episode */
    int32_t step;               /* current_step */
    int32_t grid_w, grid_h;
    int32_t num_sensors;        /* N */
    uint32_t visited_lo, visited_hi; /* sensors_visited as a bit set */
    int32_t capture_triggers;
    int32_t boundary_hits;
    int32_t edge_steps;
    int32_t collisions;         /* cumulative collision_count */
    int32_t first_full_coverage_step; /* -1 = not yet (dqn.py:428-431) */
    uint32_t episode;           /* episodes completed by this env (RNG counter word) */
    int32_t delivered;          /* cumulative successful transmissions (iot_sensors.py:141-143) */
    int32_t init_sf;            /* SF the episode's sensors started with (12, or the inherited one of DomainRandEnv) */
    int32_t dwell_steps;        /* RelationalUAVEnv._dwell_steps */
    double wind_angle;          /* WindDriftEnv: direction of the running episode's drift vector (:207); the reference adds
                                   wind_drift * cos / sin of it to the position in float64 */
} uavsim_env_record;

typedef struct uavsim_field_view {
    void* ptr;              /* device pointer */
    int32_t elem_bytes;     /* bytes per scalar element */
    int32_t is_float;       /* 1 float, 0 integer, 2 = struct record */
    int32_t ndim;
    int64_t shape[4];
    int64_t stride_bytes[4];
} uavsim_field_view;

/* Per-episode statistics row written when an env finishes (DomainRandEnv.last_episode_stats,
 * dqn.py:316-331, plus Monitor's r/l).  Index into a float64[UAVSIM_STAT_COUNT] row. */
enum {
    UAVSIM_STAT_RETURN = 0, UAVSIM_STAT_LENGTH = 1, UAVSIM_STAT_TOTAL_GENERATED = 2,
    UAVSIM_STAT_TOTAL_COLLECTED = 3, UAVSIM_STAT_TOTAL_LOST = 4, UAVSIM_STAT_BATTERY_REMAINING = 5,
    UAVSIM_STAT_NDR = 6, UAVSIM_STAT_FAIRNESS_STD = 7, UAVSIM_STAT_JAINS = 8, UAVSIM_STAT_GRID_W = 9,
    UAVSIM_STAT_GRID_H = 10, UAVSIM_STAT_NUM_SENSORS = 11, UAVSIM_STAT_DATA_EFFICIENCY = 12,
    UAVSIM_STAT_BYTES_PER_WH = 13, UAVSIM_STAT_TIME_TO_COVERAGE = 14, UAVSIM_STAT_CAPTURE_TRIGGERS = 15,
    UAVSIM_STAT_BOUNDARY_HITS = 16, UAVSIM_STAT_EDGE_STEPS = 17, UAVSIM_STAT_COLLISIONS = 18,
    UAVSIM_STAT_BASE_RETURN = 19, UAVSIM_STAT_COUNT = 20
};
/* Device-side running sums over finished episodes (payload of the one NCCL all-reduce):
 * float64[UAVSIM_STAT_COUNT + 1]; element 0..COUNT-1 = sum of the stat, element COUNT = episodes finished. */
#define UAVSIM_STAT_SUMS (UAVSIM_STAT_COUNT + 1)

/* UAVEnvironment._get_info() (uav_env.py:676-700) as a float64[UAVSIM_INFO_COUNT] row; the per-sensor
 * "sensor_collection_ratios" list travels as a separate [P] row.  Written for the TERMINAL state of a finished
 * episode (uavsim_outputs.terminal_info / terminal_ratios: what SB3 callers read from infos[i] at episode end,
 * dqn.py:1148-1155, sim_to_real_sweep.py:222-225) and on demand by uavsim_get_info. */
enum {
    UAVSIM_INFO_X = 0, UAVSIM_INFO_Y = 1,           /* "uav_position" */
    UAVSIM_INFO_BATTERY = 2, UAVSIM_INFO_BATTERY_PERCENT = 3, UAVSIM_INFO_SENSORS_COLLECTED = 4,
    UAVSIM_INFO_CURRENT_STEP = 5, UAVSIM_INFO_TOTAL_REWARD = 6, UAVSIM_INFO_TOTAL_DATA_COLLECTED = 7,
    UAVSIM_INFO_COVERAGE_PERCENTAGE = 8, UAVSIM_INFO_IS_ALIVE = 9,
    UAVSIM_INFO_MAX_URGENCY = 10, UAVSIM_INFO_AVG_URGENCY = 11, /* float32 np.max / np.mean of the AoI buf / rate */
    UAVSIM_INFO_HIGH_URGENCY_SENSORS = 12,                      /* count of AoI > 0.8 */
    UAVSIM_INFO_CAPTURE_EFFECT_TRIGGERS = 13, UAVSIM_INFO_BOUNDARY_HITS = 14, UAVSIM_INFO_EDGE_STEPS = 15,
    UAVSIM_INFO_LAST_STEP_BYTES = 16,               /* kept as float32 in the record */
    UAVSIM_INFO_NUM_SENSORS = 17,                   /* length of the ratios row (not a reference key) */
    UAVSIM_INFO_COUNT = 18
};

/* Output buffers, allocated by the caller (e.g. torch tensors -> zero-copy), borrowed by the library.
 * reward dtype follows precision (float64 / float32).  Nullable members are optional. */
typedef struct uavsim_outputs {
    float* obs;            /* [E][frame_rows][obs_dim].  K = frame_stack + latency_steps frames are kept.  K == 1: one row.
                              K > 1: a mirrored ring of 2K rows; the step's frame goes to rows s and s+K
                              (s = uavsim_frame_slot), so the K most recent frames, oldest first, are the contiguous rows
                              s+1..s+K and the stacked observation of env e (the frame_stack oldest of them) is the
                              frame_stack*obs_dim floats at obs + e*frame_rows*obs_dim + (s+1)*obs_dim. */
    void* reward;          /* [E] real */
    uint8_t* done;         /* [E] terminated|truncated (terminated is always false in the reference, uav_env.py:471) */
    float* terminal_obs;   /* [E][frame_stack*obs_dim] last (stacked) observation of a finished episode (auto_reset), nullable */
    double* episode_stats; /* [E][UAVSIM_STAT_COUNT] written when an env finishes, nullable */
    uint8_t* action_mask;  /* [E] bit a set = action a valid (derived: uav.py:175; COLLECT always), nullable */
    double* step_detail;   /* [E][UAVSIM_DETAIL_COUNT] per-step diagnostics for parity tests, nullable */
    void* draw_export;     /* [E][6][P] real: draws actually used this step (Philox mode), nullable */
    double* terminal_info;   /* [E][UAVSIM_INFO_COUNT] _get_info() of the terminal state, written when an env finishes
                                (auto_reset; a caller that resets by hand reads uavsim_get_info first), nullable */
    double* terminal_ratios; /* [E][P] "sensor_collection_ratios" of the terminal state: transmitted / max(generated,
                                1e-6) per sensor (uav_env.py:696-699), nullable */
} uavsim_outputs;

enum {
    UAVSIM_DETAIL_COLLISIONS = 0, UAVSIM_DETAIL_BYTES = 1, UAVSIM_DETAIL_URGENCY_REDUCED = 2,
    UAVSIM_DETAIL_MEAN_URGENCY = 3, UAVSIM_DETAIL_WAS_NEW = 4, UAVSIM_DETAIL_WAS_EMPTY = 5,
    UAVSIM_DETAIL_ALL_COLLECTED = 6, UAVSIM_DETAIL_DATA_LOSS = 7, UAVSIM_DETAIL_ATTEMPT_MASK_LO = 8,
    UAVSIM_DETAIL_ATTEMPT_MASK_HI = 9, UAVSIM_DETAIL_WINNER_MASK_LO = 10, UAVSIM_DETAIL_WINNER_MASK_HI = 11,
    UAVSIM_DETAIL_BASE_REWARD = 12, UAVSIM_DETAIL_TRUNCATED = 13,
    UAVSIM_DETAIL_STARVED = 14, /* sensors that drew the terminal starvation penalty (reward_function.py:59-67) */
    UAVSIM_DETAIL_COUNT = 16
};

typedef struct uavsim_dims {
    int32_t obs_dim;        /* 3 + features_per_sensor*obs_slots */
    int32_t obs_row_floats; /* floats per env in outputs.obs (= frame rows * obs_dim) */
    int32_t frame_rows;     /* physical frame rows per env: 2K (mirrored ring) when K = frame_stack + latency_steps > 1, else 1 */
    int32_t real_bytes;     /* 8 or 4 */
    int64_t arena_bytes;    /* device bytes uavsim_create needs for all state */
    int64_t replay_step_doubles;  /* doubles per env in a replay draw table: 6*P */
} uavsim_dims;

/* Injected scenario (per env).  All device pointers, nullable = keep. */
typedef struct uavsim_scenario {
    const float* sensor_xy;     /* [E][P][2]; rows beyond an env's N ignored */
    const int32_t* num_sensors; /* [E] */
    const int32_t* grid_wh;     /* [E][2] */
    const float* start_xy;      /* [E][2] */
} uavsim_scenario;

/* Replay tables (UAVSIM_RNG_REPLAY).  All device pointers to float64.
 * step_draws: draws of the step about to run, [E][6][P].
 * reset_draws/reset_fill: consumed by uavsim_reset and by in-step auto-reset:
 *   reset_draws [E][6][P] (only slots D,E read), reset_fill [E][P] = np_random.uniform(0.2,0.6)
 *   (uav_env.py:410).  reset_init_sf [E] int32 (DomainRand: SF inherited from the old sensor 0,
 *   dqn.py:345-353), nullable -> config.init_sf.
 * reset_scenario: layout/grid/start the reset installs (DomainRand), nullable members keep. */
typedef struct uavsim_replay {
    const double* step_draws;
    const double* reset_draws;
    const double* reset_fill;
    const int32_t* reset_init_sf;
    uavsim_scenario reset_scenario;
    const double* dwell_draws; /* [E][P]: shadowing values of the relational wrapper's is_in_range probes on a COLLECT step
                                  (env_wrapper.py:303), nullable unless wrapper == UAVSIM_WRAP_RELATIONAL */
    const double* step_env_draws;  /* [E][4] per-env draws of the sim-to-real wrappers for the step about to run:
                                      [0], [1] = GPS noise on obs[0], obs[1]; nullable unless gps_sigma > 0 */
    const double* reset_env_draws; /* [E][4] the same for a reset: [0], [1] = GPS noise on the reset observation,
                                      [2] = wind angle in [0, 2 pi); nullable unless gps_sigma > 0 or wind_drift > 0 */
} uavsim_replay;

typedef struct uavsim_handle uavsim_handle;

/* ---- host-only helpers (no device needed) -------------------------------- */
int uavsim_abi_version(void);
/* Fills *cfg with the reference's training-path defaults (SURVEY.md section 8 header). */
int uavsim_default_config(uavsim_config* cfg);
/* Validates cfg and reports sizes.  Pure host arithmetic. */
int uavsim_query_dims(const uavsim_config* cfg, uavsim_dims* out);
const char* uavsim_last_error(void);

/* ---- lifecycle ------------------------------------------------------------ */
/* arena_dev: device memory of at least dims.arena_bytes, 256-byte aligned, owned by the caller
 * (torch allocates, the library borrows); NULL -> the library cudaMallocs (and frees) its own.
 * Zero-fills the arena on `stream`. */
int uavsim_create(const uavsim_config* cfg, void* arena_dev, size_t arena_bytes, void* stream,
                  uavsim_handle** out);
int uavsim_destroy(uavsim_handle* h);
int uavsim_bind_outputs(uavsim_handle* h, const uavsim_outputs* out, void* stream);
int uavsim_get_field(uavsim_handle* h, int field, uavsim_field_view* out);

/* ---- scenario / replay ---------------------------------------------------- */
/* Install injected layouts etc. (replaces `sensor_positions=` / `uav_start_position=` ctor args). */
int uavsim_set_scenario(uavsim_handle* h, const uavsim_scenario* sc, void* stream);
/* Generate per-env uniform layouts with Philox, np.random.uniform([0,0],[W,H]) in distribution
 * (uav_env.py:366-374).  FIXED mode construction when no layout is injected. */
int uavsim_generate_layouts(uavsim_handle* h, void* stream);
int uavsim_set_replay(uavsim_handle* h, const uavsim_replay* rp);
/* DomainRandEnv.set_curriculum_stage (dqn.py:258-277): the grid list later resets sample from. */
int uavsim_set_stage_grids(uavsim_handle* h, int32_t num_grids, const int32_t* grids_wh);

/* ---- the hot path --------------------------------------------------------- */
/* Gymnasium reset for the envs whose mask byte is non-zero (NULL = all): uav_env.py:400-427 /
 * dqn.py:301-372.  Writes obs (and action_mask). */
int uavsim_reset(uavsim_handle* h, const uint8_t* env_mask_dev, void* stream);
/* One environment step for every env: uav_env.py:429-488 (+ dqn.py:415-444 in DOMAIN_RAND mode).
 * actions_dev: [E] int32 in [0,4].  Writes obs, reward, done (+ optional outputs). */
int uavsim_step(uavsim_handle* h, const int32_t* actions_dev, void* stream);
/* Same step with HOST buffers: copies actions host->device, steps, copies obs/reward/done back, then synchronises
 * `stream`.  Pipelined: the batch is stepped as up to uavsim_set_host_chunks (default 4) env ranges, and the copy-out
 * of range k runs on an internal stream while range k+1 is stepped (environments are independent and their draws are
 * keyed by the global env id, so the result is bit-identical to one launch); batches under 32 Ki envs take one launch.
 * Buffers should be pinned (and NUMA-local to the GPU) for full PCIe rate.  obs_host may be NULL to skip that copy. */
int uavsim_step_host(uavsim_handle* h, const int32_t* actions_host, float* obs_host, void* reward_host,
                     uint8_t* done_host, void* stream);
/* Number of env ranges uavsim_step_host pipelines over: 1 (no overlap), 2, 4 or 8. */
int uavsim_set_host_chunks(uavsim_handle* h, int32_t chunks);

/* uavsim_step for a SUBSET of the envs, and PerturbedDQNEnv's wind repeat (sim_to_real_sweep.py:117-119: "with
 * probability min(0.95, wind / 2) a move is run once more first, and that extra step's result is dropped").
 * env_mask_dev: [E] bytes, only envs with a non-zero byte step (NULL = all).  flags: */
#define UAVSIM_STEP_EXTRA 1u /* the launch is such a dropped repeat: no episode restarts, the reward does not enter
                                episode_return (it does enter total_reward, as env.total_reward in the reference), the
                                frame ring does not advance -- the commanded step that follows overwrites reward / done
                                / observation and ends the episode if the repeat exhausted it */
int uavsim_step_ex(uavsim_handle* h, const int32_t* actions_dev, const uint8_t* env_mask_dev, uint32_t flags, void* stream);
/* Per-environment overrides of config.shadowing_std_db (dB) and config.duty_cycle (percent): a whole robustness sweep
 * (sim_to_real_sweep.py:72-77 varies them per run) in one batch.  [E] float32 device arrays, borrowed; NULL = config. */
int uavsim_set_env_params(uavsim_handle* h, const float* shadowing_std_db_dev, const float* duty_cycle_dev);

/* ---- on-device heuristic driver ------------------------------- */
/* NearestSensorGreedy.select_action for every env (greedy_agents.py:73-100, :41-68): COLLECT if a sensor with
 * data is in range (one fresh shadowing draw per probe, iot_sensors.py:214-219), else one cell toward the nearest
 * sensor with data.  actions_dev: [E] int32 out.  probe_draws_dev: replay mode only, [E][P] float64 shadowing
 * values of the probes (NaN = not drawn); Philox mode draws from its own stream and takes NULL. */
int uavsim_policy_nearest_greedy(uavsim_handle* h, int32_t* actions_dev, const double* probe_draws_dev, void* stream);
/* MaxThroughputGreedyV2.select_action for every env (greedy_agents.py:105-200): COLLECT if a sensor with data is
 * in range (every sensor with data is probed, :133-138) and its spreading factor is within the adaptive threshold
 * (:188-197); else one cell toward the sensor maximising SF priority + buffer fill + duty cycle - distance
 * (:155-182).  Same argument conventions as uavsim_policy_nearest_greedy. */
int uavsim_policy_max_throughput_greedy(uavsim_handle* h, int32_t* actions_dev, const double* probe_draws_dev,
                                        void* stream);

/* ---- introspection / control ---------------------------------------------- */
/* UAVEnvironment._get_info() of the CURRENT state of every env: info_dev [E][UAVSIM_INFO_COUNT] float64,
 * ratios_dev [E][P] float64 (nullable).  One small kernel on `stream`. */
int uavsim_get_info(uavsim_handle* h, double* info_dev, double* ratios_dev, void* stream);
/* env.reset(seed=s) (uav_env.py:400-410; callers reseed per evaluation episode, sweep_eval.py:415,463,584):
 * re-keys the Philox streams of every later launch, and the next uavsim_reset restarts the episode counters of the
 * envs it covers at 0 -- the episode that follows is a function of (seed, global env id) alone, as a reseeded
 * reference env replays.  Replay mode: accepted, no effect on the draws (they are injected). */
int uavsim_set_seed(uavsim_handle* h, uint64_t seed);
/* Restore the frame-ring position after the caller has restored the bound obs buffer from a checkpoint taken
 * together with uavsim_frame_slot() (UAVVecEnv.snapshot / restore). */
int uavsim_set_frame_slot(uavsim_handle* h, int32_t slot);

/* ---- diagnostics ---------------------------------------------------------- */
/* Synchronises `stream`, then reports and clears the device error word: returns
 * UAVSIM_ERR_INVALID_ACTION if any env saw an action outside [0,4] since the last poll
 * (*first_env = lowest such local env id of the first offending step, nullable). */
int uavsim_poll_error(uavsim_handle* h, void* stream, int64_t* first_env);
/* float64[UAVSIM_STAT_SUMS] device array of running episode-stat sums (all-reduce payload). */
int uavsim_stat_sums(uavsim_handle* h, double** sums_dev);
/* Kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t uavsim_launch_count(const uavsim_handle* h);
/* Row s of the mirrored frame ring the newest frame was written to (0 when frame_stack == 1; -1 before the
 * first reset).  The stacked observation window starts at row s+1 (see uavsim_outputs.obs). */
int uavsim_frame_slot(const uavsim_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* UAVSIM_H_ */
