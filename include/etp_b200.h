/*
 * etp_b200.h - C ABI of the B200-native candidate-trajectory scoring path.
 *
 * The reference (TUMFTM/EthicalTrajectoryPlanning) is pure Python and has no FFI: its
 * boundary for this path is a set of Python call signatures.  Each
 * entry point below names the reference interface it replaces (paths relative to the
 * reference root).  Signatures carry only plain pointers and sizes; no torch types.
 *
 * Conventions
 *   - every function returns ETP_OK (0) or a negative etp_status; no exceptions cross the ABI;
 *     etp_last_error() returns a human-readable message for the last failure on that context;
 *   - one etp_ctx per (device, stream); a context is not thread-safe, distinct contexts are
 *     independent;
 *   - INPUT pointers (parameters, spline, obstacles, grids, optional host masks) are HOST
 *     memory: they are small (KBs) and are staged through pinned memory by the library;
 *   - OUTPUT tensors in etp_out are caller-allocated DEVICE memory (e.g. torch tensors'
 *     data_ptr()); any of them may be NULL, in which case that tensor is not materialised;
 *   - the library's own device buffers (profile table, selection workspace, arenas of a sweep
 *     or a planning cycle) live in the context and GROW on demand: a call that meets a larger
 *     grid, more obstacles or more scenarios than any before it allocates (cudaMalloc, a
 *     synchronising call); repeated calls of one shape allocate nothing;
 *   - etp_plan_step() is asynchronous on the context's stream; etp_get_best() /
 *     etp_get_trajectory() synchronise and read back a few hundred bytes.
 *
 * Index conventions
 *   dense candidate index  c = (iv * nt + it) * nd + id   (generation order v -> t -> d of
 *   planner/Frenet/utils/frenet_functions.py:247,253,330).  Candidates the reference skips
 *   (`ds <= |dT|`, frenet_functions.py:333-334) keep their dense slot with level ETP_LEVEL_SKIPPED.
 */
#ifndef ETP_B200_H
#define ETP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct etp_ctx etp_ctx;

typedef enum {
    ETP_OK = 0,
    ETP_ERR_CUDA = -1,         /* a CUDA runtime call failed */
    ETP_ERR_INVALID = -2,      /* bad argument / inconsistent sizes */
    ETP_ERR_UNSUPPORTED = -3,  /* mode outside the scored path (e.g. gidas harm, reach-set responsibility) */
    ETP_ERR_NO_CANDIDATE = -4, /* every candidate skipped: the reference raises ValueError (frenet_functions.py:562-564) */
    ETP_ERR_STATE = -5         /* params / spline / obstacles not set before plan_step */
} etp_status;

/* order of etp_params.weights: keys of planner/Frenet/configs/weights*.json */
enum {
    ETP_W_BAYES = 0, ETP_W_EQUALITY, ETP_W_MAXIMIN, ETP_W_RESPONSIBILITY, ETP_W_EGO, ETP_W_RISK_COST,
    ETP_W_VISIBLE_AREA, ETP_W_LON_JERK, ETP_W_LAT_JERK, ETP_W_VELOCITY, ETP_W_DIST_TO_GLOBAL_PATH,
    ETP_W_TRAVELLED_DIST, ETP_W_DIST_TO_GOAL_POS, ETP_W_DIST_TO_LANE_CENTER, ETP_NUM_WEIGHTS
};

/* rows of etp_out.traj: fields of FrenetTrajectory (frenet_functions.py:32-102) minus the shared t */
enum {
    ETP_F_D = 0, ETP_F_D_D, ETP_F_D_DD, ETP_F_D_DDD, ETP_F_S, ETP_F_S_D, ETP_F_S_DD, ETP_F_S_DDD,
    ETP_F_X, ETP_F_Y, ETP_F_YAW, ETP_F_V, ETP_F_CURV, ETP_NUM_FIELDS
};

/* rows of etp_out.red: per (candidate, obstacle) reductions over time (risk_costs.py:98-101) */
enum { ETP_R_EGO_RISK = 0, ETP_R_OBST_RISK, ETP_R_EGO_HARM, ETP_R_OBST_HARM, ETP_NUM_RED };

/* rows of etp_out.cost (calc_trajectory_cost.py:130-146, 269-298, 321-346) */
enum {
    ETP_C_BAYES = 0, ETP_C_EQUALITY, ETP_C_MAXIMIN, ETP_C_RESPONSIBILITY, ETP_C_EGO, ETP_C_RISK_TOTAL,
    ETP_C_VELOCITY, ETP_C_DIST_TO_GLOBAL_PATH, ETP_C_LON_JERK, ETP_C_LAT_JERK, ETP_C_TRAVELLED_DIST,
    ETP_C_FACTOR, ETP_C_TOTAL, ETP_C_DIST_TO_GOAL_POS, ETP_NUM_COST
};

/* validity levels (validity_checks.py:22-28) and the slot marker for skipped candidates */
enum { ETP_LEVEL_PHYSICAL = 0, ETP_LEVEL_COLLISION = 1, ETP_LEVEL_BOUNDARY = 2, ETP_LEVEL_MAX_RISK = 3,
       ETP_LEVEL_VALID = 10, ETP_LEVEL_SKIPPED = 255 };

/* reason_invalid strings of validity_checks.py:68-122 as codes */
enum { ETP_REASON_NONE = 0, ETP_REASON_VELOCITY = 1, ETP_REASON_ACCELERATION = 2, ETP_REASON_CURV_LVC = 3,
       ETP_REASON_CURV_HVC = 4, ETP_REASON_COLLISION = 5, ETP_REASON_BOUNDARIES = 6, ETP_REASON_MAX_RISK = 7 };

/* planner mode (frenet_planner.py:199-202, frenet_functions.py:524-527, validity_checks.py:287) */
enum { ETP_MODE_RISK = 0, ETP_MODE_WALENET = 1, ETP_MODE_GROUND_TRUTH = 2 };

/*
 * Opt-outs of reference quirks.  Strict reproduction is the default; a bit switches ONE quirk to what its
 * surrounding code and comments intend.  The reference cannot produce these variants: they are checked against this
 * repository's own restatement (oracle/), not against reference fixtures.
 *   Q1_CURVATURE   : kappa = (dx ddy - ddx dy) / (dx^2 + dy^2)^(3/2) instead of the reference's denominator
 *                    dx^2 + (dy^2)^(3/2) (frenet_functions.py:308-310)
 *   Q2_ACCELERATION: the "acceleration" test checks all(|s_dd| <= a_max) (level 0, reason "acceleration") instead of
 *                    repeating the velocity test (validity_checks.py:70-75)
 *   Q3_MAXIMIN     : the maximin principle keeps the harm of obstacles whose risk is NOT below eps ("set maximin to 0 if
 *                    probability (or risk) is 0", risk_costs.py:204-206) instead of the inverted gate
 *   Q9_VIEW_CONE   : the bearing of an obstacle is compared with the ego orientation modulo 2 pi in the action-space
 *                    responsibility flag (responsibility.py:78-89; etp_set_raw_predictions)
 */
enum { ETP_RELAX_Q1_CURVATURE = 1, ETP_RELAX_Q2_ACCELERATION = 2, ETP_RELAX_Q3_MAXIMIN = 4, ETP_RELAX_Q9_VIEW_CONE = 8 };

/*
 * Collision of the ego box with the predicted obstacle boxes (validity_checks.py:91-103,216-274;
 * prediction_helpers.py:138-210) in planner modes "risk" and "WaleNet".  The reference asks pycrcc: ego boxes
 * RectOBB(l/2, w/2, yaw[i], x[i], y[i]) at time steps time_step + i, prediction boxes RectOBB(length/2, width/2,
 * orientation_list[j], pos_list[j]) at time_step + 1 + j for j < min(len(ft.t), len(pos_list)), each object first run
 * through trajectory_preprocess_obb_sum and, if that fails, used as it is.
 *   HOST     : the device tests nothing; pycrcc's answer arrives in etp_step.ext_level (the strict default).
 *   DISCRETE : the un-preprocessed objects: box i = j + 1 against box j (separating-axis test).  Exact semantics
 *              of the reference's fallback branch (validity_checks.py:136-141, prediction_helpers.py:189-194).
 *   SWEPT    : consecutive boxes (k, k + 1) of each object are replaced by the box in the frame of box k that
 *              encloses both, then tested pairwise.  pycrcc's hull construction is not part of the reference
 *              sources (commonroad-drivability-checker, un-vendored): same intent, not pinned to it.
 * A hit is OR-ed with ext_level == 1.
 */
enum { ETP_COLLISION_HOST = 0, ETP_COLLISION_DISCRETE = 1, ETP_COLLISION_SWEPT = 2 };

/* element type of the real-valued output tensors and of the risk arithmetic */
enum { ETP_PRECISION_FP32 = 0, ETP_PRECISION_FP64 = 1 };

#define ETP_MAX_ANGLE_BINS 8

/*
 * Flattened params_dict = {'weights', 'modes', 'harm'} (frenet_planner.py:149-153) plus the
 * vehicle constants of planner/utils/vehicleparams.py:63-91.
 *
 * The protected-occupant log-reg model selected by risk.json (harm_estimation.py:358-400) is
 * passed as thresholds on |angle| plus one coefficient per band and sign:
 *   band 0: |a| < thr[0]  (front, coefficient 0), band i: thr[i-1] <= |a| < thr[i], last band: else.
 * (logistic_regression_symmetrical.py:7-130, logistic_regression_asymmetrical.py:7-96.)
 */
typedef struct {
    double weights[ETP_NUM_WEIGHTS];
    int32_t planner_mode;            /* ETP_MODE_* */
    int32_t fast_prob_mahalanobis;   /* risk.json: collision_probability.py:259-292 instead of 136-256 */
    int32_t multiple_cost_functions; /* risk.json: calc_trajectory_cost.py:322-327 */
    int32_t n_angle_thresholds;      /* 0 = ignore angle (LR1S) */
    double max_acceptable_risk;      /* validity_checks.py:277-294 */
    double angle_thr[ETP_MAX_ANGLE_BINS];
    double angle_coef_pos[ETP_MAX_ANGLE_BINS + 1];
    double angle_coef_neg[ETP_MAX_ANGLE_BINS + 1];
    double lr_const, lr_speed;       /* selected protected model */
    double lr1s_const, lr1s_speed;   /* ignore-angle model: ego harm against unprotected obstacles */
    double ped_const, ped_speed;     /* harm_parameters.json "pedestrian" (harm_estimation.py:408-414) */
    double veh_l, veh_w, veh_l_r, veh_m, veh_steer_max, veh_lat_a_max, veh_v_max, veh_a_max;
    int32_t collision_check;         /* ETP_COLLISION_*: collision with the predictions (validity level 1) on the device */
    int32_t relaxed_quirks;          /* ETP_RELAX_* bits; 0 = strict reference behaviour (the default) */
} etp_params;

/*
 * Data contract 1 of SURVEY.md 8b: the `predictions` dict after
 * get_orientation_velocity_and_shape_of_prediction (prediction_helpers.py:75-135) and
 * assign_responsibility_by_action_space (responsibility.py:57-75), as SoA host arrays.
 * n_obstacles = -1 means `predictions is None` (no risk assessment at all).
 */
typedef struct {
    int32_t n_obstacles;          /* O */
    int32_t max_pred;             /* row stride Tp of the arrays below */
    const int32_t* pred_len;      /* [O] valid prediction length per obstacle (>= 1) */
    const double* pos;            /* [O][max_pred][2]  pos_list */
    const double* cov;            /* [O][max_pred][4]  cov_list, row major 2x2 */
    const double* yaw;            /* [O][max_pred]     orientation_list */
    const double* vel;            /* [O][max_pred]     v_list */
    const double* length;         /* [O] shape['length'] (margin included) */
    const double* width;          /* [O] shape['width'] */
    const double* mass;           /* [O] get_obstacle_mass (properties.py:14-46) */
    const int32_t* is_protected;  /* [O] obstacle_protection (harm_estimation.py:41-58): 1 / 0 */
    const int32_t* responsibility;/* [O] predictions[id]['responsibility'] */
} etp_obstacles;

/*
 * Goal context of one planning step: what get_cost_factor (calc_trajectory_cost.py:349-415) and
 * velocity_costs (:438-519) read from planning_problem.goal, goal_area and ego_state.  With a goal attached to a step
 * the device derives, per candidate, the velocity cost and the cost factor; velocity_cost_mode / velocity_target /
 * cost_factor / cost_factor_list of the step are then ignored.
 *   goal area  : union of simple polygons (lanelet.convert_to_polygon() / goal_state.position of
 *                helper_functions.py:226-271) and circles; has_area = 0 stands for goal_area None (survival scenario:
 *                every position counts as inside, calc_trajectory_cost.py:585-586).  Containment is the even-odd rule
 *                in fp64; pycrcc's treatment of points exactly on an edge is not reproduced.
 *   velocity / orientation : goal.state_list[0].velocity / .orientation intervals if present (open intervals,
 *                helper_functions.py:408-421; the orientation test is check_orientation, :881-932).
 *   time_step  : goal.state_list[0].time_step interval (the reference requires it once a goal state is reached).
 */
typedef struct {
    int32_t has_area;
    int32_t n_polygons;
    const int32_t* poly_start;      /* [n_polygons + 1] first vertex of each polygon */
    const double* vertices;         /* [poly_start[n_polygons]][2] */
    int32_t n_circles;
    const double* circles;          /* [n_circles][3] x, y, radius */
    int32_t has_velocity;
    double velocity_start, velocity_end;
    int32_t has_orientation;
    double orientation_start, orientation_end;
    int32_t time_start, time_end;
    int32_t ego_time_step;          /* ego_state.time_step */
    double ego_x, ego_y;            /* ego_state.position */
    int32_t has_centers;            /* velocity_costs found goal centre points (:473-489); 0: survival scenario, cost 0 outside the
                                     * goal; -1: a goal position without a centre point -- the reference averages an empty list
                                     * (NaN) there, the call fails with ETP_ERR_UNSUPPORTED */
    double avg_goal_distance;       /* mean distance from ego_state.position to them (:492-497) */
    /* calc_dist_to_goal_pos (calc_trajectory_cost.py:760-803; weights['dist_to_goal_pos'] > 0): distance of the trajectory's
     * last position to the closest goal position.  Goal given in lanelets: their centre lines as polylines (closest point
     * of the polyline, :835-853); goal given as shapes: the centres of goal.state_list[i].position; neither (survival
     * scenario): 0. */
    int32_t n_goal_lines;
    const int32_t* goal_line_start; /* [n_goal_lines + 1] first vertex of each centre line */
    const double* goal_line_vertices; /* [goal_line_start[n_goal_lines]][2] */
    int32_t n_goal_points;
    const double* goal_points;      /* [n_goal_points][2] */
} etp_goal;

/*
 * Arguments of calc_frenet_trajectories (frenet_functions.py:207-221) plus the host-only
 * cost context of sort_frenet_trajectories the GPU cannot derive (last row).
 */
typedef struct {
    double c_s, c_s_d, c_s_dd, c_d, c_d_d, c_d_dd;
    const double* v_list; int32_t nv;
    const double* t_list; int32_t nt;
    const int32_t* n_samples;       /* [nt] len(np.arange(0, tT, dt)) (frenet_functions.py:267) */
    const double* d_list; int32_t nd;
    double dt, v_thr;
    int32_t velocity_cost_mode;     /* outcome of velocity_costs (calc_trajectory_cost.py:438-519):
                                       0 |target - mean v|, 1 30 - mean v, 2 mean v, 3 zero */
    double velocity_target;
    double cost_factor;             /* get_cost_factor (calc_trajectory_cost.py:349-415) */
    const uint8_t* ext_level;       /* optional [nv*nt*nd]: 1 collision, 2 boundary, else pass (pycrcc results) */
    const double* bd_harm;          /* optional [nv*nt*nd]: boundary harm (risk_costs.py:104-123) */
    const double* cost_factor_list; /* optional [nv*nt*nd]: per-candidate get_cost_factor (goal reached / reached in time /
                                       reached and left, calc_trajectory_cost.py:349-415); overrides cost_factor */
    int32_t c_begin, c_end;         /* dense candidate range of this call; must be aligned to whole (v,t) profiles */
    int32_t shard_index, shard_count; /* multi-GPU: score only the profiles p with (p - p_begin) % shard_count == shard_index
                                       (interleaved, so slow and fast profiles spread evenly); shard_count 0 or 1 = all */
    const etp_goal* goal;           /* optional: velocity cost and cost factor per candidate on the device */
} etp_step;

/* Caller-allocated device tensors for the scored candidates.  Cs = (number of profiles of this shard) * nd;
 * local column cl = k * nd + id is dense candidate (p_begin + shard_index + k * shard_count) * nd + id.
 * Every real-valued tensor is CANDIDATE-MINOR (structure of arrays over candidates): element [..][cl] lives at
 * row * c_stride + cl, so the 32 candidates a warp scores together are one contiguous 128-byte line per row.
 * c_stride >= Cs; a multiple of 32 keeps every line aligned (c_stride = 0 means Cs). */
typedef struct {
    int32_t precision;   /* ETP_PRECISION_*; element type of traj/prob/red/cost */
    void* traj;          /* [ETP_NUM_FIELDS][Tmax][c_stride]  (entries past a candidate's T are 0) */
    void* prob;          /* [Tmax-1][O][c_stride]             collision probability per timestep and obstacle */
    void* red;           /* [ETP_NUM_RED][O][c_stride] */
    void* cost;          /* [ETP_NUM_COST][c_stride] */
    uint8_t* level;      /* [Cs] validity level or ETP_LEVEL_SKIPPED */
    uint8_t* reason;     /* [Cs] ETP_REASON_* */
    int64_t c_stride;    /* elements between consecutive rows of the tensors above */
    void* harm;          /* optional [2][Tmax-1][O][c_stride]: ego / obstacle harm per sample, 0 past min(T-1, Tp)
                            (the arrays get_harm returns, harm_estimation.py:202-338) */
    double* bd_harm;     /* optional [Cs] fp64: boundary harm of each candidate (fp.bd_harm, risk_costs.py:104-123): the
                            device's own when a road boundary is set, else the caller's etp_step.bd_harm (or 0) */
} etp_out;

/*
 * Caller-provided trajectories (the reference's per-trajectory entry points: calc_risk risk_costs.py:16,
 * get_collision_probability_fast collision_probability.py:136, get_harm harm_estimation.py:202, check_validity
 * validity_checks.py:31, calc_trajectory_costs calc_trajectory_cost.py:33, and sort_frenet_trajectories on an
 * arbitrary fp_list).  HOST memory, candidate-minor like the outputs.
 */
typedef struct {
    int32_t n_traj;                 /* C */
    int32_t t_max;                  /* samples of the longest trajectory (>= 2) */
    const int32_t* n_samples;       /* [C] samples of each trajectory (>= 2) */
    const double* fields;           /* [ETP_NUM_FIELDS][t_max][C]: FrenetTrajectory attributes, rows ETP_F_* */
    int32_t velocity_cost_mode;     /* as in etp_step */
    double velocity_target;
    double cost_factor;
    const uint8_t* ext_level;       /* optional [C] */
    const double* bd_harm;          /* optional [C] */
    const double* cost_factor_list; /* optional [C], overrides cost_factor */
    const etp_goal* goal;           /* optional, like etp_step.goal; sample i of a trajectory is taken at time i * dt */
    double dt;                      /* scenario time step, only read with a goal */
} etp_trajectories;

/* Selected candidate: sort by (highest non-empty level, cost, generation order)
 * (frenet_functions.py:562-570, frenet_planner.py:457,555-558). */
typedef struct {
    uint64_t key_hi;     /* (4 - level rank) << 61 | m >> 3, m = order-preserving 64-bit image of the cost: smaller is better */
    uint64_t key_lo;     /* (m & 7) << 61 | dense candidate index: (key_hi, key_lo) orders like (level, cost, index) exactly */
    double cost;
    int32_t index;       /* dense index, -1 if no candidate */
    int32_t level;
    int32_t n_scored;    /* non-skipped candidates in the range */
    int32_t list_index;  /* position in the reference's fp_list: rank among the non-skipped candidates of [c_begin, c_end) */
} etp_best;

int etp_create(etp_ctx** ctx, int device);
int etp_destroy(etp_ctx* ctx);
const char* etp_last_error(const etp_ctx* ctx);
const char* etp_version(void);

/* Use an existing CUDA stream (cudaStream_t as void*; NULL is CUDA's legacy default stream);
 * (void*)-1 restores the context's own non-blocking stream. */
int etp_set_stream(etp_ctx* ctx, void* cuda_stream);

/* replaces params_dict / vehicle_params arguments of sort_frenet_trajectories (frenet_functions.py:477-495).
 * The packed obstacle tile depends on veh_m (mass ratios, harm_estimation.py:327-328) and, for raw predictions, on
 * ETP_RELAX_Q9_VIEW_CONE: when a later call changes either, the obstacles have to be set again -- etp_plan_step returns
 * ETP_ERR_STATE until they are, instead of scoring against a stale tile. */
int etp_set_params(etp_ctx* ctx, const etp_params* params);

/* replaces the `csp` argument of calc_frenet_trajectories (frenet_functions.py:218,278):
 * knots[n_seg+1], coef_x/coef_y[n_seg][4] with value = a + b*ds + c*ds^2 + d*ds^3 */
int etp_set_spline(etp_ctx* ctx, const double* knots, const double* coef_x, const double* coef_y, int n_seg);

/*
 * Road boundary of the scenario as triangles (boundary.create_road_boundary_obstacle(method="aligned_triangulation"),
 * frenet_planner.py:226-236): the region OUTSIDE the road.  With a boundary set, every planning step tests the
 * candidates' ego boxes against it on the device (trajectories_collision_static_obstacles, validity_checks.py:232-245
 * and risk_costs.py:104-112): the first colliding box i gives boundary_harm = LR1S(traj.v[i]) (risk_costs.py:114-121),
 * which enters the risk costs and makes the candidate level 2 (validity_checks.py:108-115); without predictions any
 * collision is level 2.  The device result replaces etp_step.bd_harm.  `check` = ETP_COLLISION_DISCRETE tests the boxes
 * of create_collision_object as they are (the reference's fallback branch), ETP_COLLISION_SWEPT merges consecutive
 * boxes first like etp_params.collision_check (not pinned to pycrcc's hull).  n_triangles = 0 or check =
 * ETP_COLLISION_HOST removes the boundary.  Persistent across steps; a uniform grid over the triangles is built here.
 */
typedef struct {
    int32_t n_triangles;
    const double* triangles;   /* [n_triangles][3][2] */
    int32_t check;             /* ETP_COLLISION_* */
} etp_road_boundary;
int etp_set_road_boundary(etp_ctx* ctx, const etp_road_boundary* boundary);

/* replaces the `predictions` argument (frenet_functions.py:481; collision_probability.py:136; harm_estimation.py:202) */
int etp_set_obstacles(etp_ctx* ctx, const etp_obstacles* obstacles, int precision);

/*
 * The step BEFORE the scored path, on the device: raw predictions {pos_list, cov_list} as the predictor
 * returns them are extended by orientation, velocity, shape margins and the action-space responsibility flag, and packed,
 * without the host ever forming the enriched dict:
 *   get_orientation_velocity_and_shape_of_prediction (planner/Frenet/utils/prediction_helpers.py:75-135): np.gradient of
 *     the positions over t_i = 0.0 + i * dt, "barely moves" rule (all dx < 1e-4 and all dy < 1e-4, no abs: sic) ->
 *     initial orientation, one-sample predictions -> initial orientation / velocity, shape + safety margins;
 *   assign_responsibility_by_action_space (planner/utils/responsibility.py:57-89): view cone of +-pi/4, no angle wrap.
 * Replaces etp_set_obstacles for that step.  etp_get_enriched reads the derived arrays back (the log line prints them).
 */
typedef struct {
    int32_t n_obstacles;            /* O */
    int32_t max_pred;               /* row stride Tp */
    const int32_t* pred_len;        /* [O] >= 1 (the reference deletes empty predictions, prediction_helpers.py:98-100) */
    const double* pos;              /* [O][max_pred][2] */
    const double* cov;              /* [O][max_pred][4] */
    const double* length;           /* [O] obstacle_shape.length (no margin) */
    const double* width;            /* [O] obstacle_shape.width */
    const double* init_orientation; /* [O] obstacle.initial_state.orientation */
    const double* init_velocity;    /* [O] obstacle.initial_state.velocity */
    const double* mass;             /* [O] get_obstacle_mass of the shape INCLUDING margins (properties.py:14-46) */
    const int32_t* is_protected;    /* [O] */
    double dt;                      /* scenario.dt */
    double safety_margin_length, safety_margin_width;
    double ego_x, ego_y, ego_orientation;
} etp_raw_predictions;

int etp_set_raw_predictions(etp_ctx* ctx, const etp_raw_predictions* raw, int precision);

/* orientation_list / v_list [O][max_pred], shape length / width incl. margins [O], responsibility [O] of the last
 * etp_set_raw_predictions (any pointer may be NULL).  Synchronises. */
int etp_get_enriched(etp_ctx* ctx, double* yaw, double* vel, double* length, double* width, int32_t* responsibility);

/* calc_frenet_trajectories + sort_frenet_trajectories + selection for the dense range of `step`:
 * frenet_planner.py:326-340, 433-457, 555-558.  `out` may be NULL (argmin-only mode). */
int etp_plan_step(etp_ctx* ctx, const etp_step* step, const etp_out* out);

/* Scores caller-provided trajectories with the same fused kernel (generation replaced by loads); the output tensors have
 * Tmax = t_max and Cs = n_traj; etp_get_best returns the arg-min over them (index = position in the list). */
int etp_score_trajectories(etp_ctx* ctx, const etp_trajectories* trajectories, const etp_out* out);

/* The five principle costs from per-obstacle maxima (host arrays of n values; n = 0 gives zeros):
 * get_bayesian_costs / get_equality_costs / get_maximin_costs / get_ego_costs / get_responsibility_cost
 * (risk_costs.py:128-255); responsibility[] are the action-space flags (responsibility.py:57-89).
 * maximin_eps / maximin_scale are get_maximin_costs' eps and scale_factor arguments (defaults 10e-10 and 10).
 * out[5] = bayes, equality, maximin, responsibility, ego. */
int etp_principle_costs(etp_ctx* ctx, int n, const double* ego_risk, const double* obst_risk, const double* ego_harm,
                        const double* obst_harm, const int32_t* responsibility, double boundary_harm, double maximin_eps,
                        double maximin_scale, double* out);

/*
 * Scenario sweep (planner/Frenet/plannertools/evaluatefrenet.py:16-55, planner/plannertools/evaluate.py:160-169: one
 * planner per scenario in a process pool): n independent planning steps in one call.  Each scenario brings its own
 * spline, predictions and sampling grid (and goal, etp_step.goal); the parameters and the road boundary are the context's
 * (etp_set_params, etp_set_road_boundary).  The steps run on a pool of
 * internal streams so that their small launches and uploads overlap; only the arg-min of each scenario is produced
 * (bests[i], index -1 if every candidate of scenario i was skipped).  Synchronises before returning.
 */
typedef struct {
    const etp_step* step;
    const etp_obstacles* obstacles;
    const double* knots;     /* spline of this scenario, as in etp_set_spline */
    const double* coef_x;
    const double* coef_y;
    int32_t n_seg;
} etp_scenario;

int etp_plan_batch(etp_ctx* ctx, int n, const etp_scenario* scenarios, int precision, etp_best* bests);

/*
 * One planning cycle of the closed loop in ONE call (frenet_planner.py:265-576 between "predictions in hand" and
 * `self._trajectory`): predictions of this cycle -> (enrichment) -> obstacle tile -> profiles -> fused scoring -> exact
 * selection -> the winner's 13 fp64 rows.  Exactly one of `raw` (predictor output, enriched on the device like
 * etp_set_raw_predictions) and `obstacles` (the enriched dict, like etp_set_obstacles) is given; parameters, spline and road
 * boundary are the context's; the step may carry a goal context (etp_step.goal).  All inputs of the cycle travel in one
 * upload, every kernel takes its step from device memory, the launch sequence of a cycle whose shapes equal the previous
 * cycle's is replayed as a captured CUDA graph (ETP_NO_GRAPH=1: launched one by one) -- three launches: profiles +
 * enrichment + tile | fused scoring (a small step's tile spread over a thread-block cluster) + selection | exact
 * re-scoring + the winner's trajectory --, and the result is written into pinned host memory by the last kernel.  Steps
 * with per-candidate host arrays (ext_level, bd_harm, cost_factor_list), a sharded range or more than 65536 candidates
 * run as etp_set_*_predictions + etp_plan_step + etp_get_best_trajectory inside the same call.  Arg-min only (no etp_out).
 * Returns like etp_get_best_trajectory.  Afterwards the context holds the step (etp_get_trajectory etc. work), but the
 * obstacles live in the cycle's arena: etp_plan_step needs etp_set_obstacles / etp_set_raw_predictions again.
 */
typedef struct {
    const etp_raw_predictions* raw;
    const etp_obstacles* obstacles;
    const etp_step* step;
} etp_cycle;
int etp_plan_cycle(etp_ctx* ctx, const etp_cycle* cycle, int precision, etp_best* best, double* fields, int* n_samples);

/* number of timesteps Tmax the output tensors must be strided with for `step` */
int etp_tmax(const etp_step* step);

/* synchronises; best of the last etp_plan_step on this context */
int etp_get_best(etp_ctx* ctx, etp_best* best);

/* synchronises; the 13 fp64 field rows [ETP_NUM_FIELDS][Tmax] of dense candidate `index` of the last
 * step (recomputed in fp64) -> `self._trajectory` (frenet_planner.py:564-576). n_samples receives T. */
int etp_get_trajectory(etp_ctx* ctx, int index, double* fields, int* n_samples);

/* etp_get_best + etp_get_trajectory of the winner with one launch, one device-to-host copy and one synchronisation
 * (the closed loop's read-back: frenet_planner.py:555-576).  Returns ETP_ERR_NO_CANDIDATE like etp_get_best. */
int etp_get_best_trajectory(etp_ctx* ctx, etp_best* best, double* fields, int* n_samples);

/* Device address of the etp_best record the last etp_plan_step wrote (for collectives on it). */
int etp_best_device_ptr(etp_ctx* ctx, void** ptr);

/* Optional CUDA-event timing of the kernels inside etp_plan_step (off by default).
 * etp_get_timing synchronises on the recorded events and returns milliseconds of the last step. */
int etp_enable_timing(etp_ctx* ctx, int enable);
int etp_get_timing(etp_ctx* ctx, float* profile_ms, float* score_ms);
/* the same per stage: ms[0] profile stage, ms[1] fused scoring kernel, ms[2] selection (arg-min and the fp64 re-decisions) */
int etp_get_stage_timing(etp_ctx* ctx, float* ms, int n);

/* Diagnostics: out[0] = fp32-path evaluations that needed the fp64 fallback (|rho| > 0.55) since load;
 * since the context was created: out[1] = candidates the selection re-scored in fp64 (near-ties of the arg-min, max-risk /
 * maximin-gate decisions too close to call in fp32), out[2] = how many of them differed from their fp32 cost by more than the
 * bound the fp32 pass had claimed (must stay 0), out[3] = largest such error / bound seen (float bits), out[4] / out[5] =
 * bytes this context copied host-to-device / device-to-host (output tensors stay on the device and are not in here),
 * out[6] = kernels it launched (directly or through a replayed graph), out[7] = graph replays (etp_plan_cycle). */
int etp_debug_counters(etp_ctx* ctx, uint64_t* out, int n);

/* combine shard-local bests (e.g. after an all-gather of etp_best records across ranks).
 * consecutive_ranges = 0: interleaved shards of ONE range (list_index is already global);
 * consecutive_ranges = 1: records of consecutive candidate ranges, given in dense order. */
int etp_combine_best(const etp_best* bests, int n, int consecutive_ranges, etp_best* out);

/* The same on the device: `gathered` is DEVICE memory holding n records (e.g. the output of an all-gather of
 * etp_best_device_ptr records over the ranks, on this context's stream); the result replaces this context's best record, so
 * that etp_get_best / etp_get_best_trajectory return the global selection with one read-back.  Asynchronous. */
int etp_combine_best_device(etp_ctx* ctx, const etp_best* gathered, int n, int consecutive_ranges);

#ifdef __cplusplus
}
#endif
#endif /* ETP_B200_H */
