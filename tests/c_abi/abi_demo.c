/* The drop-in boundary used from plain C: no Python, no torch, no CUDA headers -- only include/etp_b200.h and
 * libetp_b200.so.  Reads one planning step (numbers in the order written by tests/test_gpu_c_abi.py), runs
 * etp_set_params / etp_set_spline / etp_set_obstacles / etp_plan_step (arg-min only) and prints the selection and the
 * first samples of the selected trajectory.  Build:  gcc abi_demo.c -I include -L <dir of the .so> -letp_b200 */
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "etp_b200.h"

static FILE* in;

static double num(void) {
    double v;
    if (fscanf(in, "%lf", &v) != 1) {
        fprintf(stderr, "input file too short\n");
        exit(2);
    }
    return v;
}

static double* doubles(int n) {
    double* p = (double*)malloc(sizeof(double) * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) p[i] = num();
    return p;
}

static int32_t* ints(int n) {
    int32_t* p = (int32_t*)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) p[i] = (int32_t)num();
    return p;
}

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != 0) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, etp_last_error(ctx)); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 2 || !(in = fopen(argv[1], "r"))) {
        fprintf(stderr, "usage: abi_demo <step file>\n");
        return 2;
    }
    etp_ctx* ctx = NULL;
    etp_params p = {0};
    for (int i = 0; i < ETP_NUM_WEIGHTS; ++i) p.weights[i] = num();
    p.planner_mode = (int32_t)num();
    p.fast_prob_mahalanobis = (int32_t)num();
    p.multiple_cost_functions = (int32_t)num();
    p.n_angle_thresholds = (int32_t)num();
    p.max_acceptable_risk = num();
    for (int i = 0; i < ETP_MAX_ANGLE_BINS; ++i) p.angle_thr[i] = num();
    for (int i = 0; i <= ETP_MAX_ANGLE_BINS; ++i) p.angle_coef_pos[i] = num();
    for (int i = 0; i <= ETP_MAX_ANGLE_BINS; ++i) p.angle_coef_neg[i] = num();
    p.lr_const = num(); p.lr_speed = num(); p.lr1s_const = num(); p.lr1s_speed = num(); p.ped_const = num(); p.ped_speed = num();
    p.veh_l = num(); p.veh_w = num(); p.veh_l_r = num(); p.veh_m = num(); p.veh_steer_max = num(); p.veh_lat_a_max = num();
    p.veh_v_max = num(); p.veh_a_max = num();
    p.collision_check = (int32_t)num();

    const int n_seg = (int)num();
    double* knots = doubles(n_seg + 1);
    double* cx = doubles(4 * n_seg);
    double* cy = doubles(4 * n_seg);

    etp_obstacles o = {0};
    o.n_obstacles = (int32_t)num();
    o.max_pred = (int32_t)num();
    const int O = o.n_obstacles, Tp = o.max_pred;
    o.pred_len = ints(O);
    o.pos = doubles(O * Tp * 2);
    o.cov = doubles(O * Tp * 4);
    o.yaw = doubles(O * Tp);
    o.vel = doubles(O * Tp);
    o.length = doubles(O);
    o.width = doubles(O);
    o.mass = doubles(O);
    o.is_protected = ints(O);
    o.responsibility = ints(O);

    etp_step s = {0};
    s.c_s = num(); s.c_s_d = num(); s.c_s_dd = num(); s.c_d = num(); s.c_d_d = num(); s.c_d_dd = num();
    s.nv = (int32_t)num(); s.v_list = doubles(s.nv);
    s.nt = (int32_t)num(); s.t_list = doubles(s.nt); s.n_samples = ints(s.nt);
    s.nd = (int32_t)num(); s.d_list = doubles(s.nd);
    s.dt = num(); s.v_thr = num();
    s.velocity_cost_mode = (int32_t)num(); s.velocity_target = num(); s.cost_factor = num();
    s.c_begin = 0; s.c_end = s.nv * s.nt * s.nd;

    if (etp_create(&ctx, 0) != 0) {
        fprintf(stderr, "etp_create failed (no CUDA device?)\n");
        return 1;
    }
    CHECK(etp_set_params(ctx, &p));
    CHECK(etp_set_spline(ctx, knots, cx, cy, n_seg));
    CHECK(etp_set_obstacles(ctx, &o, ETP_PRECISION_FP32));
    CHECK(etp_plan_step(ctx, &s, NULL)); /* arg-min only: no output tensors */
    etp_best b;
    CHECK(etp_get_best(ctx, &b));
    printf("best %d %d %d %d %.17g\n", b.index, b.list_index, b.level, b.n_scored, b.cost);
    const int tmax = etp_tmax(&s);
    double* fields = (double*)malloc(sizeof(double) * ETP_NUM_FIELDS * tmax);
    int n = 0;
    CHECK(etp_get_trajectory(ctx, b.index, fields, &n));
    printf("traj %d", n);
    for (int t = 0; t < 3 && t < n; ++t) printf(" %.17g %.17g", fields[ETP_F_X * tmax + t], fields[ETP_F_Y * tmax + t]);
    printf("\n%s\n", etp_version());
    /* the same step as one planning cycle (etp_plan_cycle: predictions + step in, winner + trajectory out), three times:
     * launched directly, captured, replayed */
    etp_cycle cyc;
    memset(&cyc, 0, sizeof(cyc));
    cyc.obstacles = &o;
    cyc.step = &s;
    etp_best bc;
    int nc = 0;
    for (int k = 0; k < 3; ++k) CHECK(etp_plan_cycle(ctx, &cyc, ETP_PRECISION_FP32, &bc, fields, &nc));
    printf("cycle %d %d %d %d %.17g %d", bc.index, bc.list_index, bc.level, bc.n_scored, bc.cost, nc);
    for (int t = 0; t < 3 && t < nc; ++t) printf(" %.17g %.17g", fields[ETP_F_X * tmax + t], fields[ETP_F_Y * tmax + t]);
    printf("\n");
    CHECK(etp_destroy(ctx));
    return 0;
}
