// C ABI of the scoring path (include/etp_b200.h): context, input staging, launches, readback.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "etp_launch.h"

using namespace etp;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool view = false;  // p points into another buffer (stage_group): not owned, never freed here
    cudaError_t reserve(size_t bytes) {
        if (!view && bytes <= cap) return cudaSuccess;
        if (p && !view) cudaFree(p);
        p = nullptr;
        cap = 0;
        view = false;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void set_view(void* q) {
        if (p && !view) cudaFree(p);
        p = q;
        cap = 0;
        view = true;
    }
    void release() {
        if (p && !view) cudaFree(p);
        p = nullptr;
        cap = 0;
        view = false;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
    unsigned char* p = nullptr;
    size_t cap = 0, used = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost((void**)&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct etp_ctx {
    int device = 0;
    int num_sms = 0;
    int smem_optin = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t staged = nullptr;
    bool staged_pending = false;
    std::string err;
    bool have_params = false, have_spline = false, have_obs = false, have_step = false;
    DevParams pr{};
    int precision = ETP_PRECISION_FP32;
    // spline
    DevBuf knots, cx, cy;
    int nseg = 0;
    // obstacles
    DevBuf o_pos, o_cov, o_yaw, o_vel, o_len, o_mass, o_plen, o_prot, o_resp, o_tile, o_const;
    DevBuf o_wid, o_rawlen, o_rawwid, o_ori0, o_vel0;  // etp_set_raw_predictions
    PinnedBuf enr_host;
    int O = 0, Tp = 0, has_pred = 0;
    double origin_x = 0.0, origin_y = 0.0;  // of the fp32 coordinates: first obstacle sample, else the spline start
    std::vector<double> spline_pts, obs_box;  // knot positions (+ reach) and the bounding box of the predictions
    // step
    DevBuf grids, ext_level, bd_harm, factor_list, goal_geom, prof, prof_meta, nskip;
    DevBuf rb_tri, rb_trif, rb_start, rb_items;  // etp_set_road_boundary
    DevBuf obs_blob;                 // the input arrays of etp_set_obstacles / etp_set_raw_predictions, one upload
    std::vector<unsigned char> group_host;
    DevBoundary boundary{};
    DevBuf counters, best_dev, traj_dev, traj_n, given, given_T, prin;
    // selection (etp_select.cu): per-candidate workspace of the fused kernel, survivors of select_kernel, their exact keys
    DevBuf ws_cost, ws_bound, ws_meta, ws_bd, ws_gf, sel_blocks, sel_list, sel_res, sel_small;
    // planning cycle (etp_plan_cycle): host copy of the spline, a version of everything the captured graph bakes in, the
    // cycle's arena / pinned images, the captured launch sequence and the shapes it was captured for
    std::vector<double> sp_knots, sp_cx, sp_cy;
    unsigned long long config_version = 0;
    DevBuf cyc_dev;
    PinnedBuf cyc_in, cyc_out;
    cudaGraph_t cyc_graph = nullptr;
    cudaGraphExec_t cyc_exec = nullptr;
    std::vector<long long> cyc_key, cyc_seen;
    std::vector<unsigned char> goal_img, goal_img_next;  // goal geometry resident in goal_geom (stage_goal)
    const void* goal_img_base = nullptr;
    bool no_graph = false, cycle_profile = false, cycle_split = false;
    double cyc_prof[3] = {0.0, 0.0, 0.0};
    int cyc_prof_n = 0;
    cudaEvent_t cyc_ev[12] = {};  // ETP_CYCLE_PROFILE=2: one event between any two operations of a cycle
    double cyc_stage[12] = {};
    int cyc_stage_n = 0;
    unsigned long long graph_launches = 0;
    // what crossed the bus and what was launched since the context was created (etp_debug_counters)
    unsigned long long h2d_bytes = 0, d2h_bytes = 0, launches = 0;
    bool refine_off = false;  // ETP_NO_REFINE=1: fp32 selection as it comes out of the fused kernel (kernel experiments)
    bool last_given = false;
    PinnedBuf stage, best_host, traj_host;
    DevStep st{};
    int last_grid = 0;
    bool timing = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // step start | profiles done | fused kernel done | selection done
    // scenario sweep: lazily created child contexts (own stream, own buffers) and the pinned result array
    std::vector<etp_ctx*> pool;
    PinnedBuf batch_host;
    // fused sweep (plan_batch_fused): device arena of a chunk, two pinned images of its input region, their copy events
    DevBuf batch_dev;
    PinnedBuf batch_in[2];
    cudaEvent_t batch_ev[2] = {nullptr, nullptr};
};

#define ETP_FAIL(ctx, code, ...)                            \
    do {                                                    \
        char _b[512];                                       \
        snprintf(_b, sizeof(_b), __VA_ARGS__);              \
        (ctx)->err = _b;                                    \
        return (code);                                      \
    } while (0)

#define ETP_CUDA(ctx, expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            ETP_FAIL(ctx, ETP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

namespace {

// host bytes -> pinned staging -> device, asynchronous on the context stream
int stage_to_device(etp_ctx* c, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return ETP_OK;
    const size_t aligned = (bytes + 255) & ~size_t(255);
    if (c->stage.used + aligned > c->stage.cap) {
        // drain the copies in flight, then reuse (or grow) the staging area
        ETP_CUDA(c, cudaStreamSynchronize(c->stream));
        c->stage.used = 0;
        if (aligned > c->stage.cap) {
            size_t want = c->stage.cap ? c->stage.cap : (size_t)1 << 20;
            while (want < aligned) want *= 2;
            ETP_CUDA(c, c->stage.reserve(want));
        }
    }
    unsigned char* h = c->stage.p + c->stage.used;
    memcpy(h, src, bytes);
    c->stage.used += aligned;
    ETP_CUDA(c, cudaMemcpyAsync(dst, h, bytes, cudaMemcpyHostToDevice, c->stream));
    c->h2d_bytes += bytes;
    return ETP_OK;
}

// The pinned staging area is a ring: copies stay in flight behind each other on the context's stream and the host
// only waits (stage_to_device) when the area wraps, i.e. every few hundred small uploads.
int begin_staging(etp_ctx*) { return ETP_OK; }

int end_staging(etp_ctx*) { return ETP_OK; }

// Several small host arrays -> ONE device blob with ONE copy (a copy costs a few microseconds of host and stream time
// whatever its size; a closed loop uploads nine arrays per cycle).  Every item becomes a view into the blob.
struct GroupItem {
    DevBuf* buf;
    const void* src;
    size_t bytes;
};

int stage_group(etp_ctx* c, DevBuf& blob, std::initializer_list<GroupItem> items) {
    size_t total = 0;
    for (const GroupItem& it : items) total += (it.bytes + 255) & ~size_t(255);
    ETP_CUDA(c, blob.reserve(total));
    c->group_host.resize(total);
    size_t off = 0;
    for (const GroupItem& it : items) {
        memcpy(c->group_host.data() + off, it.src, it.bytes);
        it.buf->set_view(static_cast<unsigned char*>(blob.p) + off);
        off += (it.bytes + 255) & ~size_t(255);
    }
    return stage_to_device(c, blob.p, c->group_host.data(), total);
}

// selection workspace of a step with st.n_local candidates; fills st.ws_* / st.refine / st.raw_*
int prepare_selection(etp_ctx* c, DevStep& st) {
    const size_t n = st.n_local > 0 ? (size_t)st.n_local : 1;
    ETP_CUDA(c, c->ws_cost.reserve(sizeof(double) * n));
    ETP_CUDA(c, c->ws_meta.reserve(n));
    st.refine = (c->precision == ETP_PRECISION_FP32 && !c->refine_off) ? 1 : 0;
    if (st.refine) {
        ETP_CUDA(c, c->ws_bound.reserve(sizeof(float) * n));
        ETP_CUDA(c, c->ws_bd.reserve(sizeof(double) * n));
        ETP_CUDA(c, c->ws_gf.reserve(n));
        ETP_CUDA(c, c->sel_list.reserve(sizeof(int) * n));
        ETP_CUDA(c, c->sel_res.reserve(rescore_rec_bytes(st.n_local)));
    }
    ETP_CUDA(c, c->sel_blocks.reserve(select_blocks_bytes(st.n_local)));
    st.ws_cost = c->ws_cost.as<double>(); st.ws_meta = c->ws_meta.as<uint8_t>();
    st.ws_bound = c->ws_bound.as<float>(); st.ws_bd = c->ws_bd.as<double>(); st.ws_gf = c->ws_gf.as<uint8_t>();
    st.raw_cov = c->o_cov.as<double>(); st.raw_vel = c->o_vel.as<double>();
    return ETP_OK;
}

SelectBufs select_bufs(etp_ctx* c) {
    SelectBufs sb;
    unsigned char* sm = c->sel_small.as<unsigned char>();
    sb.blocks = c->sel_blocks.p;
    sb.list = c->sel_list.as<int>();
    sb.res = c->sel_res.p;
    sb.hdr = reinterpret_cast<int*>(sm);
    sb.counters = reinterpret_cast<unsigned*>(sm + 16);
    sb.stats = reinterpret_cast<unsigned long long*>(sm + 32);
    sb.state = sm + 64;
    return sb;
}

// fused kernel + selection of the step in c->st
int launch_step(etp_ctx* c, const DevOut& dout) {
    const DevStep& st = c->st;
    unsigned* cnt = c->counters.as<unsigned>();
    c->launches += 1;
    ETP_CUDA(c, launch_score(c->precision, st, c->pr, dout, cnt, c->num_sms, &c->last_grid, c->stream));
    if (c->timing) ETP_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
    const SelectBufs sb = select_bufs(c);
    // shards of one step must compare exact keys with each other: a shard's survivor is always re-scored
    c->launches += (st.refine && st.n_local > select_chunk()) ? 2 : 1;
    ETP_CUDA(c, launch_select(st, sb, c->best_dev.as<BestRec>(), c->nskip.as<int>(), cnt + 1, st.given != nullptr, st.shard_count > 1, c->stream));
    if (st.refine) {
        c->launches += 1;
        ETP_CUDA(c, launch_rescore(c->precision, st, c->pr, dout, sb, c->best_dev.as<BestRec>(), c->nskip.as<int>(), c->num_sms, c->stream));
    }
    return ETP_OK;
}

}  // namespace

extern "C" {

const char* etp_version(void) { return "etp_b200 0.1 (sm_100a)"; }

int etp_create(etp_ctx** out, int device) {
    if (!out) return ETP_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return ETP_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return ETP_ERR_CUDA;
    etp_ctx* c = new etp_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return ETP_ERR_CUDA; }
    c->num_sms = prop.multiProcessorCount;
    c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return ETP_ERR_CUDA; }
    c->stream = c->own_stream;
    if (cudaEventCreateWithFlags(&c->staged, cudaEventDisableTiming) != cudaSuccess) { delete c; return ETP_ERR_CUDA; }
    bool ok = c->stage.reserve((size_t)4 << 20) == cudaSuccess && c->best_host.reserve(sizeof(BestRec)) == cudaSuccess &&
              c->counters.reserve(64) == cudaSuccess && c->best_dev.reserve(sizeof(BestRec)) == cudaSuccess &&
              c->traj_n.reserve(sizeof(int)) == cudaSuccess && c->sel_small.reserve(256) == cudaSuccess;
    if (ok) ok = cudaMemset(c->counters.p, 0, 64) == cudaSuccess && cudaMemset(c->best_dev.p, 0, sizeof(BestRec)) == cudaSuccess &&
                 cudaMemset(c->sel_small.p, 0, 256) == cudaSuccess;
    {
        const char* e = getenv("ETP_NO_REFINE");
        c->refine_off = e && e[0] == '1';
        e = getenv("ETP_NO_GRAPH");
        c->no_graph = e && e[0] == '1';
        c->cycle_profile = getenv("ETP_CYCLE_PROFILE") != nullptr;
        c->cycle_split = getenv("ETP_CYCLE_SPLIT") != nullptr;
        if (c->cycle_profile && getenv("ETP_CYCLE_PROFILE")[0] == '2')
            for (auto& e2 : c->cyc_ev) ok = ok && cudaEventCreate(&e2) == cudaSuccess;
    }
    if (!ok) { etp_destroy(c); return ETP_ERR_CUDA; }
    *out = c;
    return ETP_OK;
}

int etp_destroy(etp_ctx* c) {
    if (!c) return ETP_OK;
    cudaSetDevice(c->device);
    for (etp_ctx* k : c->pool) etp_destroy(k);
    c->pool.clear();
    c->batch_host.release();
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    c->batch_dev.release();
    c->cyc_dev.release();
    c->cyc_in.release();
    c->cyc_out.release();
    for (auto& e2 : c->cyc_ev) if (e2) cudaEventDestroy(e2);
    if (c->cyc_exec) cudaGraphExecDestroy(c->cyc_exec);
    if (c->cyc_graph) cudaGraphDestroy(c->cyc_graph);
    for (int i = 0; i < 2; ++i) {
        c->batch_in[i].release();
        if (c->batch_ev[i]) cudaEventDestroy(c->batch_ev[i]);
    }
    DevBuf* bufs[] = {&c->knots, &c->cx, &c->cy, &c->o_pos, &c->o_cov, &c->o_yaw, &c->o_vel, &c->o_len, &c->o_mass,
                      &c->o_plen, &c->o_prot, &c->o_resp, &c->o_tile, &c->o_const, &c->o_wid, &c->o_rawlen, &c->o_rawwid, &c->o_ori0, &c->o_vel0, &c->grids, &c->ext_level, &c->bd_harm, &c->factor_list, &c->goal_geom, &c->rb_tri, &c->rb_trif, &c->rb_start, &c->rb_items, &c->obs_blob,
                      &c->prof, &c->prof_meta, &c->nskip, &c->counters, &c->best_dev, &c->traj_dev,
                      &c->traj_n, &c->given, &c->given_T, &c->prin, &c->ws_cost, &c->ws_bound, &c->ws_meta, &c->ws_bd, &c->ws_gf,
                      &c->sel_blocks, &c->sel_list, &c->sel_res, &c->sel_small};
    for (DevBuf* b : bufs) b->release();
    c->stage.release();
    c->enr_host.release();
    c->best_host.release();
    c->traj_host.release();
    if (c->staged) cudaEventDestroy(c->staged);
    for (int i = 0; i < 4; ++i)
        if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return ETP_OK;
}

const char* etp_last_error(const etp_ctx* c) { return c ? c->err.c_str() : "null context"; }

int etp_set_stream(etp_ctx* c, void* s) {
    if (!c) return ETP_ERR_INVALID;
    ETP_CUDA(c, cudaSetDevice(c->device));
    ETP_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stream = (s == (void*)-1) ? c->own_stream : (cudaStream_t)s;
    return ETP_OK;
}

int etp_set_params(etp_ctx* c, const etp_params* p) {
    if (!c || !p) return ETP_ERR_INVALID;
    if (p->weights[ETP_W_RESPONSIBILITY] > 0.0)
        ETP_FAIL(c, ETP_ERR_UNSUPPORTED, "weights['responsibility'] > 0 needs reachable sets (planner/utils/reachable_set.py): out of scope");
    if (p->weights[ETP_W_VISIBLE_AREA] > 0.0 || p->weights[ETP_W_DIST_TO_LANE_CENTER] > 0.0)
        ETP_FAIL(c, ETP_ERR_UNSUPPORTED, "visible_area / dist_to_lane_center are host-side cost terms (sensor model, lanelet network)");
    if (p->n_angle_thresholds < 0 || p->n_angle_thresholds > ETP_MAX_ANGLE_BINS)
        ETP_FAIL(c, ETP_ERR_INVALID, "n_angle_thresholds out of range");
    if (p->planner_mode < ETP_MODE_RISK || p->planner_mode > ETP_MODE_GROUND_TRUTH) ETP_FAIL(c, ETP_ERR_INVALID, "planner_mode");
    DevParams& d = c->pr;
    // the packed obstacle tile holds the mass ratios m_o / (m_e + m_o) (harm_estimation.py:327-328) and, for raw predictions,
    // the responsibility flag under the Q9 rule: a change of either makes the tile stale -- the obstacles have to be set again
    if (c->have_params && c->have_obs && c->has_pred &&
        (d.veh_m != p->veh_m || ((d.relaxed ^ p->relaxed_quirks) & ETP_RELAX_Q9_VIEW_CONE)))
        c->have_obs = false;
    for (int i = 0; i < ETP_NUM_WEIGHTS; ++i) d.w[i] = p->weights[i];
    d.planner_mode = p->planner_mode;
    d.mahalanobis = p->fast_prob_mahalanobis;
    d.multi_cost = p->multiple_cost_functions;
    if (p->collision_check < ETP_COLLISION_HOST || p->collision_check > ETP_COLLISION_SWEPT) ETP_FAIL(c, ETP_ERR_INVALID, "collision_check");
    // validity_checks.py:253-268: predictions are tested in modes "WaleNet" and "risk"; "ground_truth" asks the scenario's
    // own collision checker (host)
    d.collision_check = p->planner_mode == ETP_MODE_GROUND_TRUTH ? ETP_COLLISION_HOST : p->collision_check;
    d.n_thr = p->n_angle_thresholds;
    d.max_risk = p->max_acceptable_risk;
    for (int i = 0; i < ETP_MAX_ANGLE_BINS; ++i) {
        d.thr[i] = p->angle_thr[i];
        d.fcos[i] = std::cos(p->angle_thr[i]) * std::fabs(std::cos(p->angle_thr[i]));
    }
    for (int i = 0; i <= ETP_MAX_ANGLE_BINS; ++i) { d.coef_pos[i] = p->angle_coef_pos[i]; d.coef_neg[i] = p->angle_coef_neg[i]; }
    d.lr_const = p->lr_const; d.lr_speed = p->lr_speed;
    d.lr1s_const = p->lr1s_const; d.lr1s_speed = p->lr1s_speed;
    d.ped_const = p->ped_const; d.ped_speed = p->ped_speed;
    d.veh_l = p->veh_l; d.veh_w = p->veh_w; d.veh_m = p->veh_m; d.v_max = p->veh_v_max;
    if (p->relaxed_quirks < 0 || p->relaxed_quirks > 15) ETP_FAIL(c, ETP_ERR_INVALID, "relaxed_quirks");
    d.relaxed = p->relaxed_quirks;
    d.a_max = p->veh_a_max;
    // validity_checks.py:193-199
    const double tn = std::tan(p->veh_steer_max);
    const double radius = std::sqrt((p->veh_l * p->veh_l / (tn * tn)) + (p->veh_l_r * p->veh_l_r));
    d.inv_turn_radius = 1.0 / radius;
    d.v_thr_curv = std::sqrt(p->veh_lat_a_max * radius);
    d.lat_a_max = p->veh_lat_a_max;
    for (int i = 0; i < ETP_MAX_ANGLE_BINS; ++i) d.f_fcos[i] = (float)d.fcos[i];
    for (int i = 0; i <= ETP_MAX_ANGLE_BINS; ++i) { d.f_coef_pos[i] = (float)d.coef_pos[i]; d.f_coef_neg[i] = (float)d.coef_neg[i]; }
    d.f_lr_speed = (float)d.lr_speed; d.f_lr1s_speed = (float)d.lr1s_speed; d.f_ped_speed = (float)d.ped_speed;
    if (d.n_thr == 0) d.wrapfree_yaw = 1e30;
    else if (d.coef_pos[d.n_thr] == d.coef_neg[d.n_thr]) d.wrapfree_yaw = 3.14159265358979323846 - d.thr[d.n_thr - 1] - 1e-9;
    else d.wrapfree_yaw = -1.0;
    c->have_params = true;
    c->config_version++;
    return ETP_OK;
}

int etp_set_spline(etp_ctx* c, const double* knots, const double* coef_x, const double* coef_y, int n_seg) {
    if (!c || !knots || !coef_x || !coef_y || n_seg < 1) return ETP_ERR_INVALID;
    for (int i = 0; i < n_seg; ++i)
        if (!(knots[i + 1] > knots[i])) ETP_FAIL(c, ETP_ERR_INVALID, "spline knots must be strictly increasing");
    ETP_CUDA(c, cudaSetDevice(c->device));
    int r = begin_staging(c);
    if (r) return r;
    ETP_CUDA(c, c->knots.reserve(sizeof(double) * (n_seg + 1)));
    ETP_CUDA(c, c->cx.reserve(sizeof(double) * 4 * n_seg));
    ETP_CUDA(c, c->cy.reserve(sizeof(double) * 4 * n_seg));
    if ((r = stage_to_device(c, c->knots.p, knots, sizeof(double) * (n_seg + 1)))) return r;
    if ((r = stage_to_device(c, c->cx.p, coef_x, sizeof(double) * 4 * n_seg))) return r;
    if ((r = stage_to_device(c, c->cy.p, coef_y, sizeof(double) * 4 * n_seg))) return r;
    c->nseg = n_seg;
    c->sp_knots.assign(knots, knots + n_seg + 1);
    c->sp_cx.assign(coef_x, coef_x + 4 * (size_t)n_seg);
    c->sp_cy.assign(coef_y, coef_y + 4 * (size_t)n_seg);
    c->config_version++;
    c->spline_pts.clear();
    for (int i = 0; i < n_seg; ++i) {
        // knot position and the reach of the segment (|b| ds bounds the excursion to first order, doubled for slack)
        const double ds = knots[i + 1] - knots[i];
        c->spline_pts.push_back(coef_x[4 * i]);
        c->spline_pts.push_back(coef_y[4 * i]);
        c->spline_pts.push_back(2.0 * std::fmax(std::fabs(coef_x[4 * i + 1]), std::fabs(coef_y[4 * i + 1])) * ds);
    }
    c->have_spline = true;
    return end_staging(c);
}

int etp_set_road_boundary(etp_ctx* c, const etp_road_boundary* b) {
    if (!c || !b) return ETP_ERR_INVALID;
    if (b->check < ETP_COLLISION_HOST || b->check > ETP_COLLISION_SWEPT) ETP_FAIL(c, ETP_ERR_INVALID, "road boundary check mode");
    if (b->n_triangles < 0 || (b->n_triangles > 0 && !b->triangles)) ETP_FAIL(c, ETP_ERR_INVALID, "road boundary triangles missing");
    c->boundary = DevBoundary{};
    c->config_version++;
    if (b->n_triangles == 0 || b->check == ETP_COLLISION_HOST) return ETP_OK;
    const int n = b->n_triangles;
    const double* T = b->triangles;
    double lo[2] = {T[0], T[1]}, hi[2] = {T[0], T[1]};
    for (int i = 0; i < 3 * n; ++i)
        for (int a = 0; a < 2; ++a) {
            if (!std::isfinite(T[2 * i + a])) ETP_FAIL(c, ETP_ERR_INVALID, "road boundary vertex %d is not finite", i);
            lo[a] = std::fmin(lo[a], T[2 * i + a]);
            hi[a] = std::fmax(hi[a], T[2 * i + a]);
        }
    // uniform grid: cells of at least 4 m, at most 512 x 512 of them
    double cell = std::fmax(4.0, std::fmax(hi[0] - lo[0], hi[1] - lo[1]) / 512.0);
    const int nx = (int)std::floor((hi[0] - lo[0]) / cell) + 1, ny = (int)std::floor((hi[1] - lo[1]) / cell) + 1;
    auto cell_of = [&](double v, int a) {
        const int k = (int)std::floor((v - lo[a]) / cell);
        const int m = a == 0 ? nx : ny;
        return k < 0 ? 0 : (k >= m ? m - 1 : k);
    };
    // two passes over the triangles' bounding boxes: count per cell, then fill (CSR layout)
    std::vector<int> start((size_t)nx * ny + 1, 0), items, fill;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            int run = 0;
            for (size_t k = 0; k < start.size(); ++k) {
                const int v = start[k];
                start[k] = run;
                run += v;
            }
            items.assign(run > 0 ? run : 1, 0);
            fill.assign(start.begin(), start.end());
        }
        for (int i = 0; i < n; ++i) {
            const double* q = T + 6 * (size_t)i;
            const int ix0 = cell_of(std::fmin(q[0], std::fmin(q[2], q[4])), 0), ix1 = cell_of(std::fmax(q[0], std::fmax(q[2], q[4])), 0);
            const int iy0 = cell_of(std::fmin(q[1], std::fmin(q[3], q[5])), 1), iy1 = cell_of(std::fmax(q[1], std::fmax(q[3], q[5])), 1);
            for (int iy = iy0; iy <= iy1; ++iy)
                for (int ix = ix0; ix <= ix1; ++ix) {
                    // the triangle's bounding box reaches this cell; keep it only if no edge normal separates the two
                    // (cell grown by a millimetre so that rounding can only add entries)
                    const double cx = lo[0] + (ix + 0.5) * cell, cy = lo[1] + (iy + 0.5) * cell, h = 0.5 * cell + 1e-3;
                    bool apart = false;
                    for (int k = 0; k < 3 && !apart; ++k) {
                        const int j = (k + 1) % 3, m = (k + 2) % 3;
                        const double ex = -(q[2 * j + 1] - q[2 * k + 1]), ey = q[2 * j] - q[2 * k];  // edge normal
                        const double a = ex * (q[2 * k] - cx) + ey * (q[2 * k + 1] - cy), bb = ex * (q[2 * m] - cx) + ey * (q[2 * m + 1] - cy);
                        const double r = h * (std::fabs(ex) + std::fabs(ey));
                        apart = std::fmin(a, bb) > r || std::fmax(a, bb) < -r;
                    }
                    if (apart) continue;
                    if (pass == 0) ++start[(size_t)iy * nx + ix];
                    else items[fill[(size_t)iy * nx + ix]++] = i;
                }
        }
    }
    {
        ETP_CUDA(c, cudaSetDevice(c->device));
        int r = begin_staging(c);
        if (r) return r;
        ETP_CUDA(c, c->rb_tri.reserve(sizeof(double) * 6 * (size_t)n));
        ETP_CUDA(c, c->rb_trif.reserve(sizeof(float) * 6 * (size_t)n));
        std::vector<float> tf(6 * (size_t)n);
        for (size_t i = 0; i < tf.size(); ++i) tf[i] = (float)(T[i] - lo[i & 1]);
        if ((r = stage_to_device(c, c->rb_trif.p, tf.data(), sizeof(float) * tf.size()))) return r;
        ETP_CUDA(c, c->rb_start.reserve(sizeof(int) * start.size()));
        ETP_CUDA(c, c->rb_items.reserve(sizeof(int) * items.size()));
        if ((r = stage_to_device(c, c->rb_tri.p, T, sizeof(double) * 6 * (size_t)n))) return r;
        if ((r = stage_to_device(c, c->rb_start.p, start.data(), sizeof(int) * start.size()))) return r;
        if ((r = stage_to_device(c, c->rb_items.p, items.data(), sizeof(int) * items.size()))) return r;
        if ((r = end_staging(c))) return r;
    }
    DevBoundary& d = c->boundary;
    d.check = b->check;
    d.n_tri = n; d.nx = nx; d.ny = ny;
    d.x0 = lo[0]; d.y0 = lo[1]; d.inv_cell = 1.0 / cell;
    d.tri = c->rb_tri.as<double>();
    d.trif = c->rb_trif.as<float>();
    // fp32 pass: vertices and box centres are rounded to fp32 relative to the grid origin (half an ulp each)
    d.tol_f = (float)(2e-4 + 4.0 * 1.2e-7 * (std::fmax(hi[0] - lo[0], hi[1] - lo[1]) + 16.0));
    d.cell_start = c->rb_start.as<int>();
    d.cell_items = c->rb_items.as<int>();
    return ETP_OK;
}

int etp_set_obstacles(etp_ctx* c, const etp_obstacles* o, int precision) {
    if (!c || !o) return ETP_ERR_INVALID;
    if (precision != ETP_PRECISION_FP32 && precision != ETP_PRECISION_FP64) ETP_FAIL(c, ETP_ERR_INVALID, "precision");
    if (!c->have_params) ETP_FAIL(c, ETP_ERR_STATE, "etp_set_params must precede etp_set_obstacles (mass ratios use the ego mass)");
    ETP_CUDA(c, cudaSetDevice(c->device));
    c->precision = precision;
    c->obs_box.clear();
    c->origin_x = c->origin_y = 0.0;
    if (o->n_obstacles < 0) {  // predictions is None
        c->has_pred = 0; c->O = 0; c->Tp = 0; c->have_obs = true;
        return ETP_OK;
    }
    c->has_pred = 1;
    const int O = o->n_obstacles, Tp = o->max_pred;
    if (O > 0) {
        if (Tp < 1 || !o->pred_len || !o->pos || !o->cov || !o->yaw || !o->vel || !o->length || !o->width || !o->mass || !o->is_protected ||
            !o->responsibility)
            ETP_FAIL(c, ETP_ERR_INVALID, "obstacle arrays missing");
        for (int i = 0; i < O; ++i) {
            if (o->pred_len[i] < 1 || o->pred_len[i] > Tp)
                ETP_FAIL(c, ETP_ERR_INVALID, "pred_len[%d]=%d outside [1, max_pred]; empty predictions are dropped by the caller "
                         "(prediction_helpers.py:98-100)", i, o->pred_len[i]);
            if (o->is_protected[i] != 0 && o->is_protected[i] != 1)
                ETP_FAIL(c, ETP_ERR_UNSUPPORTED, "obstacle %d has no protection class (harm_estimation.py:52-56): unsupported input", i);
        }
        int r = begin_staging(c);
        if (r) return r;
        const size_t n = (size_t)O * Tp;
        c->origin_x = o->pos[0];
        c->origin_y = o->pos[1];
        c->obs_box.assign({o->pos[0], o->pos[0], o->pos[1], o->pos[1]});
        for (int i = 0; i < O; ++i)
            for (int t = 0; t < o->pred_len[i]; ++t) {
                const double* q = o->pos + ((size_t)i * Tp + t) * 2;
                c->obs_box[0] = std::fmin(c->obs_box[0], q[0]); c->obs_box[1] = std::fmax(c->obs_box[1], q[0]);
                c->obs_box[2] = std::fmin(c->obs_box[2], q[1]); c->obs_box[3] = std::fmax(c->obs_box[3], q[1]);
            }
        ETP_CUDA(c, c->o_tile.reserve(Tile<double>::bytes(O, Tp)));
        ETP_CUDA(c, c->o_const.reserve(sizeof(double) * OC_COUNT * O));
        if ((r = stage_group(c, c->obs_blob,
                             {{&c->o_pos, o->pos, sizeof(double) * 2 * n}, {&c->o_cov, o->cov, sizeof(double) * 4 * n},
                              {&c->o_yaw, o->yaw, sizeof(double) * n}, {&c->o_vel, o->vel, sizeof(double) * n},
                              {&c->o_len, o->length, sizeof(double) * O}, {&c->o_wid, o->width, sizeof(double) * O},
                              {&c->o_mass, o->mass, sizeof(double) * O}, {&c->o_plen, o->pred_len, sizeof(int) * O},
                              {&c->o_prot, o->is_protected, sizeof(int) * O}, {&c->o_resp, o->responsibility, sizeof(int) * O}})))
            return r;
        if ((r = end_staging(c))) return r;
        c->launches += 1;
    ETP_CUDA(c, launch_pack_obstacles(precision, O, Tp, c->o_plen.as<int>(), c->o_pos.as<double>(), c->o_cov.as<double>(),
                                          c->o_yaw.as<double>(), c->o_vel.as<double>(), c->o_len.as<double>(),
                                          c->o_wid.as<double>(), c->o_mass.as<double>(), c->o_prot.as<int>(), c->o_resp.as<int>(), c->pr.veh_m,
                                          c->origin_x, c->origin_y, c->o_tile.p, c->o_const.as<double>(), c->stream));
    }
    c->O = O;
    c->Tp = Tp;
    c->have_obs = true;
    return ETP_OK;
}

int etp_set_raw_predictions(etp_ctx* c, const etp_raw_predictions* o, int precision) {
    if (!c || !o) return ETP_ERR_INVALID;
    if (precision != ETP_PRECISION_FP32 && precision != ETP_PRECISION_FP64) ETP_FAIL(c, ETP_ERR_INVALID, "precision");
    if (!c->have_params) ETP_FAIL(c, ETP_ERR_STATE, "etp_set_params must precede etp_set_raw_predictions");
    ETP_CUDA(c, cudaSetDevice(c->device));
    c->precision = precision;
    c->obs_box.clear();
    c->origin_x = c->origin_y = 0.0;
    const int O = o->n_obstacles, Tp = o->max_pred;
    if (O < 0) ETP_FAIL(c, ETP_ERR_INVALID, "n_obstacles < 0: use etp_set_obstacles for `predictions is None`");
    c->has_pred = 1;
    if (O > 0) {
        if (Tp < 1 || !o->pred_len || !o->pos || !o->cov || !o->length || !o->width || !o->init_orientation || !o->init_velocity ||
            !o->mass || !o->is_protected)
            ETP_FAIL(c, ETP_ERR_INVALID, "raw prediction arrays missing");
        for (int i = 0; i < O; ++i) {
            if (o->pred_len[i] < 1 || o->pred_len[i] > Tp)
                ETP_FAIL(c, ETP_ERR_INVALID, "pred_len[%d]=%d outside [1, max_pred]; empty predictions are dropped by the caller "
                         "(prediction_helpers.py:98-100)", i, o->pred_len[i]);
            if (o->is_protected[i] != 0 && o->is_protected[i] != 1)
                ETP_FAIL(c, ETP_ERR_UNSUPPORTED, "obstacle %d has no protection class (harm_estimation.py:52-56): unsupported input", i);
        }
        int r = begin_staging(c);
        if (r) return r;
        const size_t n = (size_t)O * Tp;
        c->origin_x = o->pos[0];
        c->origin_y = o->pos[1];
        c->obs_box.assign({o->pos[0], o->pos[0], o->pos[1], o->pos[1]});
        for (int i = 0; i < O; ++i)
            for (int t = 0; t < o->pred_len[i]; ++t) {
                const double* q = o->pos + ((size_t)i * Tp + t) * 2;
                c->obs_box[0] = std::fmin(c->obs_box[0], q[0]); c->obs_box[1] = std::fmax(c->obs_box[1], q[0]);
                c->obs_box[2] = std::fmin(c->obs_box[2], q[1]); c->obs_box[3] = std::fmax(c->obs_box[3], q[1]);
            }
        ETP_CUDA(c, c->o_yaw.reserve(sizeof(double) * n));
        ETP_CUDA(c, c->o_vel.reserve(sizeof(double) * n));
        ETP_CUDA(c, c->o_len.reserve(sizeof(double) * O));
        ETP_CUDA(c, c->o_wid.reserve(sizeof(double) * O));
        ETP_CUDA(c, c->o_resp.reserve(sizeof(int) * O));
        ETP_CUDA(c, c->o_tile.reserve(Tile<double>::bytes(O, Tp)));
        ETP_CUDA(c, c->o_const.reserve(sizeof(double) * OC_COUNT * O));
        if ((r = stage_group(c, c->obs_blob,
                             {{&c->o_pos, o->pos, sizeof(double) * 2 * n}, {&c->o_cov, o->cov, sizeof(double) * 4 * n},
                              {&c->o_rawlen, o->length, sizeof(double) * O}, {&c->o_rawwid, o->width, sizeof(double) * O},
                              {&c->o_ori0, o->init_orientation, sizeof(double) * O}, {&c->o_vel0, o->init_velocity, sizeof(double) * O},
                              {&c->o_mass, o->mass, sizeof(double) * O}, {&c->o_plen, o->pred_len, sizeof(int) * O},
                              {&c->o_prot, o->is_protected, sizeof(int) * O}})))
            return r;
        if ((r = end_staging(c))) return r;
        c->launches += 1;
    ETP_CUDA(c, launch_enrich(O, Tp, c->o_plen.as<int>(), c->o_pos.as<double>(), c->o_rawlen.as<double>(), c->o_rawwid.as<double>(),
                                  c->o_ori0.as<double>(), c->o_vel0.as<double>(), o->dt, o->safety_margin_length,
                                  o->safety_margin_width, o->ego_x, o->ego_y, o->ego_orientation, (c->pr.relaxed & ETP_RELAX_Q9_VIEW_CONE) != 0, c->o_yaw.as<double>(),
                                  c->o_vel.as<double>(), c->o_len.as<double>(), c->o_wid.as<double>(), c->o_resp.as<int>(), c->stream));
        c->launches += 1;
    ETP_CUDA(c, launch_pack_obstacles(precision, O, Tp, c->o_plen.as<int>(), c->o_pos.as<double>(), c->o_cov.as<double>(),
                                          c->o_yaw.as<double>(), c->o_vel.as<double>(), c->o_len.as<double>(),
                                          c->o_wid.as<double>(), c->o_mass.as<double>(), c->o_prot.as<int>(), c->o_resp.as<int>(), c->pr.veh_m,
                                          c->origin_x, c->origin_y, c->o_tile.p, c->o_const.as<double>(), c->stream));
    }
    c->O = O;
    c->Tp = Tp;
    c->have_obs = true;
    return ETP_OK;
}

int etp_get_enriched(etp_ctx* c, double* yaw, double* vel, double* length, double* width, int32_t* responsibility) {
    if (!c) return ETP_ERR_INVALID;
    if (!c->have_obs || !c->has_pred) ETP_FAIL(c, ETP_ERR_STATE, "no predictions set on this context");
    ETP_CUDA(c, cudaSetDevice(c->device));
    const size_t n = (size_t)c->O * c->Tp, O = (size_t)c->O;
    if (O == 0) return ETP_OK;
    const size_t bytes = sizeof(double) * (2 * n + 2 * O) + sizeof(int) * O;
    ETP_CUDA(c, c->enr_host.reserve(bytes));
    unsigned char* h = c->enr_host.p;
    ETP_CUDA(c, cudaMemcpyAsync(h, c->o_yaw.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    ETP_CUDA(c, cudaMemcpyAsync(h + sizeof(double) * n, c->o_vel.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    ETP_CUDA(c, cudaMemcpyAsync(h + sizeof(double) * 2 * n, c->o_len.p, sizeof(double) * O, cudaMemcpyDeviceToHost, c->stream));
    if (c->o_wid.p)
        ETP_CUDA(c, cudaMemcpyAsync(h + sizeof(double) * (2 * n + O), c->o_wid.p, sizeof(double) * O, cudaMemcpyDeviceToHost, c->stream));
    ETP_CUDA(c, cudaMemcpyAsync(h + sizeof(double) * (2 * n + 2 * O), c->o_resp.p, sizeof(int) * O, cudaMemcpyDeviceToHost, c->stream));
    c->d2h_bytes += bytes;
    ETP_CUDA(c, cudaStreamSynchronize(c->stream));
    if (yaw) memcpy(yaw, h, sizeof(double) * n);
    if (vel) memcpy(vel, h + sizeof(double) * n, sizeof(double) * n);
    if (length) memcpy(length, h + sizeof(double) * 2 * n, sizeof(double) * O);
    if (width) memcpy(width, h + sizeof(double) * (2 * n + O), sizeof(double) * O);
    if (responsibility) memcpy(responsibility, h + sizeof(double) * (2 * n + 2 * O), sizeof(int) * O);
    return ETP_OK;
}

int etp_tmax(const etp_step* s) {
    if (!s || !s->n_samples || s->nt < 1) return ETP_ERR_INVALID;
    int m = 0;
    for (int i = 0; i < s->nt; ++i) m = s->n_samples[i] > m ? s->n_samples[i] : m;
    return m;
}

// Digest an etp_goal for the device (calc_trajectory_cost.py:349-415, 438-519): stage the goal geometry, fold the
// orientation interval the way check_orientation does (:881-932), find the first trajectory sample whose time step
// lies in the goal's time window (reached_target_time, :639-660; calc_remaining_time_steps, :702-723) and evaluate
// what depends on the ego state only.  Must run between begin_staging and end_staging.
static int stage_goal(etp_ctx* c, const etp_goal* g, double dt, int Tmax, DevGoal* out) {
    DevGoal d{};
    if (!g) {
        *out = d;
        return ETP_OK;
    }
    if (g->n_polygons < 0 || g->n_circles < 0 || (g->n_polygons > 0 && (!g->poly_start || !g->vertices)) ||
        (g->n_circles > 0 && !g->circles))
        ETP_FAIL(c, ETP_ERR_INVALID, "goal geometry arrays missing");
    if (!(dt > 0.0)) ETP_FAIL(c, ETP_ERR_INVALID, "goal context needs the scenario time step dt > 0");
    d.on = 1;
    d.has_area = g->has_area ? 1 : 0;
    d.n_poly = d.has_area ? g->n_polygons : 0;
    d.n_circ = d.has_area ? g->n_circles : 0;
    const int n_vert = d.n_poly > 0 ? g->poly_start[d.n_poly] : 0;
    for (int p = 0; p < d.n_poly; ++p)
        if (g->poly_start[p + 1] - g->poly_start[p] < 3) ETP_FAIL(c, ETP_ERR_INVALID, "goal polygon %d has fewer than 3 vertices", p);
    const size_t b_start = sizeof(int) * (size_t)(d.n_poly + 1), b_start_pad = (b_start + 7) / 8 * 8;
    const size_t b_vert = sizeof(double) * 2 * (size_t)n_vert, b_circ = sizeof(double) * 3 * (size_t)d.n_circ;
    // goal positions of calc_dist_to_goal_pos: centre lines, or shape centres
    if (g->n_goal_lines < 0 || g->n_goal_points < 0 || (g->n_goal_lines > 0 && (!g->goal_line_start || !g->goal_line_vertices)) ||
        (g->n_goal_points > 0 && !g->goal_points))
        ETP_FAIL(c, ETP_ERR_INVALID, "goal position arrays missing");
    d.n_lines = g->n_goal_lines;
    d.n_pts = g->n_goal_points;
    const int n_lv = d.n_lines > 0 ? g->goal_line_start[d.n_lines] : 0;
    for (int l = 0; l < d.n_lines; ++l)
        if (g->goal_line_start[l + 1] - g->goal_line_start[l] < 1) ETP_FAIL(c, ETP_ERR_INVALID, "goal centre line %d is empty", l);
    const size_t b_ls = (sizeof(int) * (size_t)(d.n_lines + 1) + 7) / 8 * 8, b_lv = sizeof(double) * 2 * (size_t)n_lv,
                 b_pt = sizeof(double) * 2 * (size_t)d.n_pts;
    // the geometry of a goal rarely changes between steps: one image [polygon starts | vertices | circles | line starts | line
    // vertices | points], uploaded with one copy and only when its bytes (or the buffer) differ from what is resident
    const size_t total = b_start_pad + b_vert + b_circ + b_ls + b_lv + b_pt;
    ETP_CUDA(c, c->goal_geom.reserve(total + 8));
    unsigned char* base = c->goal_geom.as<unsigned char>();
    {
        std::vector<unsigned char>& img = c->goal_img_next;
        img.assign(total, 0);
        if (d.n_poly > 0) {
            memcpy(img.data(), g->poly_start, b_start);
            memcpy(img.data() + b_start_pad, g->vertices, b_vert);
        }
        if (d.n_circ > 0) memcpy(img.data() + b_start_pad + b_vert, g->circles, b_circ);
        unsigned char* q = img.data() + b_start_pad + b_vert + b_circ;
        if (d.n_lines > 0) {
            memcpy(q, g->goal_line_start, sizeof(int) * (size_t)(d.n_lines + 1));
            memcpy(q + b_ls, g->goal_line_vertices, b_lv);
        }
        if (d.n_pts > 0) memcpy(q + b_ls + b_lv, g->goal_points, b_pt);
        if (total > 0 && (base != c->goal_img_base || img != c->goal_img)) {
            const int r = stage_to_device(c, base, img.data(), total);
            if (r) return r;
            c->goal_img.swap(img);
            c->goal_img_base = base;
        }
    }
    d.poly_start = reinterpret_cast<const int*>(base);
    d.verts = reinterpret_cast<const double*>(base + b_start_pad);
    d.circles = reinterpret_cast<const double*>(base + b_start_pad + b_vert);
    d.line_start = reinterpret_cast<const int*>(base + b_start_pad + b_vert + b_circ);
    d.line_verts = reinterpret_cast<const double*>(base + b_start_pad + b_vert + b_circ + b_ls);
    d.pts = reinterpret_cast<const double*>(base + b_start_pad + b_vert + b_circ + b_ls + b_lv);
    d.bx0 = d.by0 = INFINITY;
    d.bx1 = d.by1 = -INFINITY;
    for (int i = 0; i < n_vert; ++i) {
        d.bx0 = std::fmin(d.bx0, g->vertices[2 * i]); d.bx1 = std::fmax(d.bx1, g->vertices[2 * i]);
        d.by0 = std::fmin(d.by0, g->vertices[2 * i + 1]); d.by1 = std::fmax(d.by1, g->vertices[2 * i + 1]);
    }
    for (int k = 0; k < d.n_circ; ++k) {
        const double* q = g->circles + 3 * k;
        d.bx0 = std::fmin(d.bx0, q[0] - q[2]); d.bx1 = std::fmax(d.bx1, q[0] + q[2]);
        d.by0 = std::fmin(d.by0, q[1] - q[2]); d.by1 = std::fmax(d.by1, q[1] + q[2]);
    }
    d.has_vel = g->has_velocity ? 1 : 0;
    d.v_lo = g->velocity_start;
    d.v_hi = g->velocity_end;
    d.vel_mid = g->velocity_start + (g->velocity_end - g->velocity_start) / 2;  // :456-463
    d.has_ori = g->has_orientation ? 1 : 0;
    {
        const double PI = 3.14159265358979323846;
        double v[2] = {g->orientation_start, g->orientation_end};
        int n_in_range = 0;
        for (int k = 0; k < 2; ++k) {
            if (v[k] > PI) {
                while (v[k] > PI) v[k] -= 2 * PI;
            } else if (v[k] < -PI) {
                while (v[k] < -PI) v[k] += 2 * PI;
            } else {
                ++n_in_range;
            }
        }
        d.o_lo = std::fmin(v[0], v[1]);
        d.o_hi = std::fmax(v[0], v[1]);
        d.ori_invert = n_in_range == 1 ? 1 : 0;
    }
    d.time_first = 0x7fffffff;
    for (int i = 0; i < Tmax; ++i) {
        const double t = 0.0 + i * dt;                                   // traj.t = np.arange(0, tT, dt)
        const int considered = (int)((double)g->ego_time_step + t / dt);  // int(ego_state_time + t / dt)
        if (g->time_start - considered <= 0 && g->time_end - considered >= 0) {
            d.time_first = i;
            break;
        }
    }
    // host-side containment of the ego position: same function, same fp64 arithmetic as the device
    d.ego_in_goal = !d.has_area || point_in_goal(d, g->poly_start, g->vertices, g->circles, g->ego_x, g->ego_y);
    if (g->has_centers < 0 && c->pr.w[ETP_W_VELOCITY] > 0.0)  // velocity_costs runs only with a positive weight (:279)
        ETP_FAIL(c, ETP_ERR_UNSUPPORTED,
                 "the goal has a position but no centre point: the reference's velocity cost is the mean of an empty list (NaN) here "
                 "(calc_trajectory_cost.py:473-497)");
    if (g->has_centers <= 0) {
        d.vel_kind = 3;  // "no velocity can be proposed" (:489-490)
    } else {
        const double remaining = (double)(g->time_end - g->ego_time_step) * dt;  // t = 0: considered = ego time step
        if (remaining > 0.0) {
            d.vel_kind = 0;
            d.vel_out = g->avg_goal_distance / remaining;
        } else {
            d.vel_kind = 1;
        }
    }
    *out = d;
    return ETP_OK;
}

int etp_plan_step(etp_ctx* c, const etp_step* s, const etp_out* out) {
    if (!c || !s) return ETP_ERR_INVALID;
    if (!c->have_params || !c->have_spline || !c->have_obs)
        ETP_FAIL(c, ETP_ERR_STATE, "etp_set_params / etp_set_spline / etp_set_obstacles must be called before etp_plan_step");
    if (s->nv < 1 || s->nt < 1 || s->nd < 1 || !s->v_list || !s->t_list || !s->d_list || !s->n_samples)
        ETP_FAIL(c, ETP_ERR_INVALID, "empty sampling grid");
    const long long dense = (long long)s->nv * s->nt * s->nd;
    if (dense >= (1ll << 31)) ETP_FAIL(c, ETP_ERR_INVALID, "more than 2^31 candidates");
    if (s->c_begin < 0 || s->c_end > dense || s->c_begin >= s->c_end) ETP_FAIL(c, ETP_ERR_INVALID, "candidate range");
    if (s->shard_count > 0 && (s->shard_index < 0 || s->shard_index >= s->shard_count))
        ETP_FAIL(c, ETP_ERR_INVALID, "shard_index outside [0, shard_count)");
    if (s->c_begin % s->nd != 0 || s->c_end % s->nd != 0)
        ETP_FAIL(c, ETP_ERR_INVALID, "candidate range must be aligned to whole (v,t) profiles (multiples of nd=%d)", s->nd);
    const int Tmax = etp_tmax(s);
    for (int i = 0; i < s->nt; ++i)
        if (s->n_samples[i] < 2) ETP_FAIL(c, ETP_ERR_INVALID, "n_samples[%d] < 2", i);
    if (Tmax > 256) ETP_FAIL(c, ETP_ERR_INVALID, "more than 256 samples per trajectory");
    if (out && out->precision != c->precision)
        ETP_FAIL(c, ETP_ERR_INVALID, "etp_out.precision differs from the precision given to etp_set_obstacles");
    ETP_CUDA(c, cudaSetDevice(c->device));
    int r = begin_staging(c);
    if (r) return r;

    // grids blob: v_list | t_list | d_list | n_samples
    const size_t off_v = 0, off_t = off_v + sizeof(double) * s->nv, off_d = off_t + sizeof(double) * s->nt,
                 off_n = off_d + sizeof(double) * s->nd, total = off_n + sizeof(int) * s->nt;
    ETP_CUDA(c, c->grids.reserve(total));
    {
        std::vector<unsigned char> blob(total);
        memcpy(blob.data() + off_v, s->v_list, sizeof(double) * s->nv);
        memcpy(blob.data() + off_t, s->t_list, sizeof(double) * s->nt);
        memcpy(blob.data() + off_d, s->d_list, sizeof(double) * s->nd);
        memcpy(blob.data() + off_n, s->n_samples, sizeof(int) * s->nt);
        if ((r = stage_to_device(c, c->grids.p, blob.data(), total))) return r;
    }
    DevStep& st = c->st;
    st = DevStep{};
    if (s->ext_level) {
        ETP_CUDA(c, c->ext_level.reserve((size_t)dense));
        if ((r = stage_to_device(c, c->ext_level.p, s->ext_level, (size_t)dense))) return r;
        st.ext_level = c->ext_level.as<uint8_t>();
    }
    if (s->bd_harm) {
        ETP_CUDA(c, c->bd_harm.reserve(sizeof(double) * (size_t)dense));
        if ((r = stage_to_device(c, c->bd_harm.p, s->bd_harm, sizeof(double) * (size_t)dense))) return r;
        st.bd_harm = c->bd_harm.as<double>();
    }
    if (s->cost_factor_list) {
        ETP_CUDA(c, c->factor_list.reserve(sizeof(double) * (size_t)dense));
        if ((r = stage_to_device(c, c->factor_list.p, s->cost_factor_list, sizeof(double) * (size_t)dense))) return r;
        st.factor_list = c->factor_list.as<double>();
    }
    if (c->pr.w[ETP_W_DIST_TO_GOAL_POS] > 0.0 && !s->goal)
        ETP_FAIL(c, ETP_ERR_INVALID, "weights['dist_to_goal_pos'] > 0 needs the goal positions of the step (etp_step.goal)");
    if ((r = stage_goal(c, s->goal, s->dt, Tmax, &st.goal))) return r;
    if ((r = end_staging(c))) return r;

    unsigned char* g = c->grids.as<unsigned char>();
    st.c_s = s->c_s; st.c_s_d = s->c_s_d; st.c_s_dd = s->c_s_dd;
    st.c_d = s->c_d; st.c_d_d = s->c_d_d; st.c_d_dd = s->c_d_dd;
    st.dt = s->dt; st.v_thr = s->v_thr;
    st.v_list = reinterpret_cast<const double*>(g + off_v);
    st.t_list = reinterpret_cast<const double*>(g + off_t);
    st.d_list = reinterpret_cast<const double*>(g + off_d);
    st.nsamp = reinterpret_cast<const int*>(g + off_n);
    st.nv = s->nv; st.nt = s->nt; st.nd = s->nd; st.Tmax = Tmax;
    st.c_begin = s->c_begin; st.c_end = s->c_end;
    st.p_begin = s->c_begin / s->nd; st.p_end = s->c_end / s->nd;
    st.shard_count = s->shard_count > 0 ? s->shard_count : 1;
    st.shard_index = s->shard_count > 0 ? s->shard_index : 0;
    {
        const int np_all = st.p_end - st.p_begin;
        const int np_loc = st.shard_index < np_all ? (np_all - st.shard_index + st.shard_count - 1) / st.shard_count : 0;
        st.n_local = np_loc * s->nd;
    }
    st.vel_mode = s->velocity_cost_mode; st.vel_target = s->velocity_target; st.factor = s->cost_factor;
    {
        // 2 ulp (fp32) of the largest coordinate, relative to the origin, a candidate or an obstacle can take
        double dmax = 0.0, m = 0.0;
        for (int i = 0; i < s->nd; ++i) dmax = std::fmax(dmax, std::fabs(s->d_list[i]));
        if (c->obs_box.empty()) {  // no obstacles: the fp32 coordinates are not used for anything position critical
            c->origin_x = c->spline_pts.empty() ? 0.0 : c->spline_pts[0];
            c->origin_y = c->spline_pts.empty() ? 0.0 : c->spline_pts[1];
        } else {
            m = std::fmax(std::fmax(std::fabs(c->obs_box[0] - c->origin_x), std::fabs(c->obs_box[1] - c->origin_x)),
                          std::fmax(std::fabs(c->obs_box[2] - c->origin_y), std::fabs(c->obs_box[3] - c->origin_y)));
        }
        for (size_t i = 0; i + 2 < c->spline_pts.size(); i += 3)
            m = std::fmax(m, std::fmax(std::fabs(c->spline_pts[i] - c->origin_x), std::fabs(c->spline_pts[i + 1] - c->origin_y)) +
                                 c->spline_pts[i + 2]);
        st.eps_pos = 2.4e-7 * (m + dmax + 16.0);
        st.origin_x = c->origin_x;
        st.origin_y = c->origin_y;
    }
    st.knots = c->knots.as<double>(); st.cx = c->cx.as<double>(); st.cy = c->cy.as<double>(); st.nseg = c->nseg;
    const int n_prof = st.p_end - st.p_begin;
    ETP_CUDA(c, c->prof.reserve(sizeof(double) * PF_COUNT * Tmax * (size_t)n_prof));
    ETP_CUDA(c, c->prof_meta.reserve(sizeof(double) * PM_COUNT * (size_t)n_prof));
    ETP_CUDA(c, c->nskip.reserve(sizeof(int) * (size_t)n_prof));
    st.prof = c->prof.as<double>();
    st.prof_meta = c->prof_meta.as<double>();
    st.nskip = c->nskip.as<int>();
    st.has_pred = c->has_pred; st.O = c->O; st.Tp = c->Tp;
    st.boundary = c->boundary;
    st.boundary.lr1s_const = c->pr.lr1s_const; st.boundary.lr1s_speed = c->pr.lr1s_speed;
    st.boundary.off_x = st.origin_x - st.boundary.x0; st.boundary.off_y = st.origin_y - st.boundary.y0;
    st.obs = c->o_tile.p; st.obs_const = c->o_const.as<double>(); st.pred_len = c->o_plen.as<int>();
    st.raw_pos = c->o_pos.as<double>(); st.raw_yaw = c->o_yaw.as<double>();

    DevOut dout{};
    if (out) {
        dout.traj = out->traj; dout.prob = out->prob; dout.red = out->red; dout.cost = out->cost;
        dout.level = out->level; dout.reason = out->reason; dout.harm = out->harm; dout.bd_harm = out->bd_harm;
        dout.c_stride = out->c_stride > 0 ? (size_t)out->c_stride : (size_t)st.n_local;
        if (dout.c_stride < (size_t)st.n_local)
            ETP_FAIL(c, ETP_ERR_INVALID, "etp_out.c_stride=%lld is smaller than the %d candidates of this call",
                     (long long)out->c_stride, st.n_local);
    }
    {
        const size_t smem = score_smem_bytes(c->precision, Tmax, c->has_pred ? c->O : 0);
        if (smem > (size_t)c->smem_optin)
            ETP_FAIL(c, ETP_ERR_UNSUPPORTED, "Tmax=%d samples x %d obstacles need %zu bytes of shared memory per CTA (limit %d)",
                     Tmax, c->O, smem, c->smem_optin);
    }
    if ((r = prepare_selection(c, st))) return r;
    if (c->timing) ETP_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
    c->launches += 1;
    ETP_CUDA(c, launch_profiles(st, c->pr, c->nskip.as<int>(), c->stream));
    if (c->timing) ETP_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    if ((r = launch_step(c, dout))) return r;
    if (c->timing) ETP_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    c->have_step = true;
    c->last_given = false;
    return ETP_OK;
}

int etp_score_trajectories(etp_ctx* c, const etp_trajectories* g, const etp_out* out) {
    if (!c || !g) return ETP_ERR_INVALID;
    if (!c->have_params || !c->have_obs)
        ETP_FAIL(c, ETP_ERR_STATE, "etp_set_params / etp_set_obstacles must be called before etp_score_trajectories");
    if (g->n_traj < 1 || g->t_max < 2 || !g->n_samples || !g->fields) ETP_FAIL(c, ETP_ERR_INVALID, "empty trajectory list");
    if (g->t_max > 256) ETP_FAIL(c, ETP_ERR_INVALID, "more than 256 samples per trajectory");
    for (int i = 0; i < g->n_traj; ++i)
        if (g->n_samples[i] < 2 || g->n_samples[i] > g->t_max) ETP_FAIL(c, ETP_ERR_INVALID, "n_samples[%d] outside [2, t_max]", i);
    if (out && out->precision != c->precision)
        ETP_FAIL(c, ETP_ERR_INVALID, "etp_out.precision differs from the precision given to etp_set_obstacles");
    ETP_CUDA(c, cudaSetDevice(c->device));
    int r = begin_staging(c);
    if (r) return r;
    const size_t n = (size_t)g->n_traj, nf = sizeof(double) * ETP_NUM_FIELDS * g->t_max * n;
    ETP_CUDA(c, c->given.reserve(nf));
    ETP_CUDA(c, c->given_T.reserve(sizeof(int) * n));
    if ((r = stage_to_device(c, c->given.p, g->fields, nf))) return r;
    if ((r = stage_to_device(c, c->given_T.p, g->n_samples, sizeof(int) * n))) return r;
    DevStep& st = c->st;
    st = DevStep{};
    if (g->ext_level) {
        ETP_CUDA(c, c->ext_level.reserve(n));
        if ((r = stage_to_device(c, c->ext_level.p, g->ext_level, n))) return r;
        st.ext_level = c->ext_level.as<uint8_t>();
    }
    if (g->bd_harm) {
        ETP_CUDA(c, c->bd_harm.reserve(sizeof(double) * n));
        if ((r = stage_to_device(c, c->bd_harm.p, g->bd_harm, sizeof(double) * n))) return r;
        st.bd_harm = c->bd_harm.as<double>();
    }
    if (g->cost_factor_list) {
        ETP_CUDA(c, c->factor_list.reserve(sizeof(double) * n));
        if ((r = stage_to_device(c, c->factor_list.p, g->cost_factor_list, sizeof(double) * n))) return r;
        st.factor_list = c->factor_list.as<double>();
    }
    if (c->pr.w[ETP_W_DIST_TO_GOAL_POS] > 0.0 && !g->goal)
        ETP_FAIL(c, ETP_ERR_INVALID, "weights['dist_to_goal_pos'] > 0 needs the goal positions (etp_trajectories.goal)");
    if ((r = stage_goal(c, g->goal, g->dt, g->t_max, &st.goal))) return r;
    if ((r = end_staging(c))) return r;
    st.given = c->given.as<double>(); st.given_T = c->given_T.as<int>(); st.given_stride = g->n_traj;
    st.nv = 1; st.nt = 1; st.nd = g->n_traj; st.Tmax = g->t_max;
    st.c_begin = 0; st.c_end = g->n_traj; st.p_begin = 0; st.p_end = 1;
    st.shard_count = 1; st.shard_index = 0; st.n_local = g->n_traj;
    st.vel_mode = g->velocity_cost_mode; st.vel_target = g->velocity_target; st.factor = g->cost_factor;
    {
        if (c->obs_box.empty()) {
            c->origin_x = g->fields[(size_t)ETP_F_X * g->t_max * n];
            c->origin_y = g->fields[(size_t)ETP_F_Y * g->t_max * n];
        }
        double m = 0.0;
        if (!c->obs_box.empty())
            m = std::fmax(std::fmax(std::fabs(c->obs_box[0] - c->origin_x), std::fabs(c->obs_box[1] - c->origin_x)),
                          std::fmax(std::fabs(c->obs_box[2] - c->origin_y), std::fabs(c->obs_box[3] - c->origin_y)));
        const size_t plane = (size_t)g->t_max * n;
        for (size_t i = 0; i < plane; ++i)
            m = std::fmax(m, std::fmax(std::fabs(g->fields[ETP_F_X * plane + i] - c->origin_x),
                                       std::fabs(g->fields[ETP_F_Y * plane + i] - c->origin_y)));
        st.eps_pos = 2.4e-7 * (m + 16.0);
        st.origin_x = c->origin_x;
        st.origin_y = c->origin_y;
    }
    st.has_pred = c->has_pred; st.O = c->O; st.Tp = c->Tp;
    st.boundary = c->boundary;
    st.boundary.lr1s_const = c->pr.lr1s_const; st.boundary.lr1s_speed = c->pr.lr1s_speed;
    st.boundary.off_x = st.origin_x - st.boundary.x0; st.boundary.off_y = st.origin_y - st.boundary.y0;
    st.obs = c->o_tile.p; st.obs_const = c->o_const.as<double>(); st.pred_len = c->o_plen.as<int>();
    st.raw_pos = c->o_pos.as<double>(); st.raw_yaw = c->o_yaw.as<double>();
    DevOut dout{};
    if (out) {
        dout.traj = out->traj; dout.prob = out->prob; dout.red = out->red; dout.cost = out->cost;
        dout.level = out->level; dout.reason = out->reason; dout.harm = out->harm; dout.bd_harm = out->bd_harm;
        dout.c_stride = out->c_stride > 0 ? (size_t)out->c_stride : (size_t)st.n_local;
        if (dout.c_stride < (size_t)st.n_local) ETP_FAIL(c, ETP_ERR_INVALID, "etp_out.c_stride too small");
    }
    {
        const size_t smem = score_smem_bytes(c->precision, st.Tmax, c->has_pred ? c->O : 0);
        if (smem > (size_t)c->smem_optin)
            ETP_FAIL(c, ETP_ERR_UNSUPPORTED, "t_max=%d samples x %d obstacles need %zu bytes of shared memory per CTA (limit %d)",
                     st.Tmax, c->O, smem, c->smem_optin);
    }
    ETP_CUDA(c, c->nskip.reserve(sizeof(int)));
    if ((r = prepare_selection(c, st))) return r;
    if (c->timing) {
        ETP_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
        ETP_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
    }
    if ((r = launch_step(c, dout))) return r;
    if (c->timing) ETP_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
    c->have_step = true;
    c->last_given = true;
    return ETP_OK;
}

int etp_principle_costs(etp_ctx* c, int n, const double* ego_risk, const double* obst_risk, const double* ego_harm,
                        const double* obst_harm, const int32_t* responsibility, double boundary_harm, double maximin_eps,
                        double maximin_scale, double* out) {
    if (!c || !out || n < 0) return ETP_ERR_INVALID;
    if (n > 0 && (!ego_risk || !obst_risk || !ego_harm || !obst_harm || !responsibility)) return ETP_ERR_INVALID;
    ETP_CUDA(c, cudaSetDevice(c->device));
    int r = begin_staging(c);
    if (r) return r;
    const size_t nb = sizeof(double) * (size_t)n;
    ETP_CUDA(c, c->prin.reserve(4 * nb + sizeof(int) * (size_t)n + 8 * sizeof(double) + 64));
    double* vals = c->prin.as<double>();
    double* res = vals + 4 * (size_t)n;
    int* resp = reinterpret_cast<int*>(res + 8);
    if (n > 0) {
        if ((r = stage_to_device(c, vals, ego_risk, nb))) return r;
        if ((r = stage_to_device(c, vals + n, obst_risk, nb))) return r;
        if ((r = stage_to_device(c, vals + 2 * (size_t)n, ego_harm, nb))) return r;
        if ((r = stage_to_device(c, vals + 3 * (size_t)n, obst_harm, nb))) return r;
        if ((r = stage_to_device(c, resp, responsibility, sizeof(int) * (size_t)n))) return r;
    }
    if ((r = end_staging(c))) return r;
    c->launches += 1;
    ETP_CUDA(c, launch_principles(n, vals, resp, boundary_harm, maximin_eps, maximin_scale, (c->pr.relaxed & ETP_RELAX_Q3_MAXIMIN) != 0, res, c->stream));
    ETP_CUDA(c, c->traj_host.reserve(5 * sizeof(double)));
    ETP_CUDA(c, cudaMemcpyAsync(c->traj_host.p, res, 5 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    c->d2h_bytes += 5 * sizeof(double);
    ETP_CUDA(c, cudaStreamSynchronize(c->stream));
    memcpy(out, c->traj_host.p, 5 * sizeof(double));
    return ETP_OK;
}

namespace {

// Can this scenario go through the fused sweep?  Per-candidate host arrays and goal contexts take the scenario-by-scenario
// path; so do steps too large for one selection block.
bool batchable(const etp_scenario& s) {
    if (!s.step || !s.obstacles || !s.knots || !s.coef_x || !s.coef_y || s.n_seg < 1) return false;
    const etp_step& p = *s.step;
    if (p.ext_level || p.bd_harm || p.cost_factor_list || p.goal) return false;
    if (p.nv < 1 || p.nt < 1 || p.nd < 1 || !p.v_list || !p.t_list || !p.d_list || !p.n_samples) return false;
    const long long dense = (long long)p.nv * p.nt * p.nd;
    if (dense > select_chunk() || p.c_begin != 0 || p.c_end != dense || p.shard_count > 1) return false;
    const etp_obstacles& o = *s.obstacles;
    if (o.n_obstacles < 0 || (o.n_obstacles > 0 && (o.max_pred < 1 || !o.pred_len || !o.pos || !o.cov || !o.yaw || !o.vel || !o.length ||
                                                    !o.width || !o.mass || !o.is_protected || !o.responsibility)))
        return false;
    for (int i = 0; i < p.nt; ++i)
        if (p.n_samples[i] < 2 || p.n_samples[i] > 256) return false;
    for (int i = 0; i < o.n_obstacles; ++i)
        if (o.pred_len[i] < 1 || o.pred_len[i] > o.max_pred || (o.is_protected[i] != 0 && o.is_protected[i] != 1)) return false;
    for (int i = 0; i < s.n_seg; ++i)
        if (!(s.knots[i + 1] > s.knots[i])) return false;
    return true;
}

struct Arena {  // offsets into a chunk's device arena, 256-byte aligned
    size_t off = 0;
    size_t take(size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) & ~size_t(255);
        return at;
    }
};

// One chunk of the sweep as five launches: all inputs in one upload, pack / profile / fused scoring / selection / exact
// re-scoring each over every scenario of the chunk, one read-back of the winners.  Asynchronous; `slot` picks the pinned
// image (two alternate so that the next chunk is assembled while this one runs).
int plan_chunk_fused(etp_ctx* c, int n, const etp_scenario* sc, int precision, BestRec* host_out, int slot) {
    const size_t rsz = precision == ETP_PRECISION_FP64 ? sizeof(double) : sizeof(float);
    (void)rsz;
    // ---- layout -----------------------------------------------------------------------------------
    Arena in, scratch;
    const size_t o_steps = in.take(sizeof(DevStep) * n), o_pack = in.take(sizeof(PackArgs) * n), o_sb = in.take(sizeof(SelectBufs) * n);
    const size_t o_first = in.take(sizeof(int) * n);
    struct Lay {
        size_t knots, cx, cy, v, t, d, ns, pos, cov, yaw, vel, len, wid, mass, plen, prot, resp, small;
        size_t tile, oconst, prof, meta, nskip, wc, wb, wm, wbd, wgf, blocks, list, res;
        int Tmax, n_prof, Cs, tiles, O, Tp;
        bool shared_spline;
    };
    std::vector<Lay> lay(n);
    long long n_tiles = 0;
    int max_items = 0, max_prof = 0, max_T = 2, max_O = 0;
    size_t max_smem = 0;
    for (int i = 0; i < n; ++i) {
        const etp_scenario& s = sc[i];
        const etp_step& p = *s.step;
        const etp_obstacles& o = *s.obstacles;
        Lay& L = lay[i];
        L.Tmax = etp_tmax(&p);
        L.n_prof = p.nv * p.nt;
        L.Cs = L.n_prof * p.nd;
        L.tiles = (L.Cs + 31) / 32;
        L.O = o.n_obstacles;
        L.Tp = L.O > 0 ? o.max_pred : 0;
        n_tiles += L.tiles;
        max_items = std::max(max_items, L.O * L.Tp);
        max_prof = std::max(max_prof, L.n_prof);
        max_T = std::max(max_T, L.Tmax);
        max_O = std::max(max_O, L.O);
        const size_t smem = score_smem_bytes(precision, L.Tmax, L.O);
        if (smem > (size_t)c->smem_optin) ETP_FAIL(c, ETP_ERR_UNSUPPORTED, "scenario %d needs %zu bytes of shared memory per CTA", i, smem);
        max_smem = std::max(max_smem, smem);
        // a spline the previous scenario of the chunk brought already (sweeps over one map) is not copied again
        L.shared_spline = i > 0 && sc[i - 1].knots == s.knots && sc[i - 1].coef_x == s.coef_x && sc[i - 1].coef_y == s.coef_y &&
                          sc[i - 1].n_seg == s.n_seg;
        if (L.shared_spline) {
            L.knots = lay[i - 1].knots; L.cx = lay[i - 1].cx; L.cy = lay[i - 1].cy;
        } else {
            L.knots = in.take(sizeof(double) * (s.n_seg + 1));
            L.cx = in.take(sizeof(double) * 4 * s.n_seg);
            L.cy = in.take(sizeof(double) * 4 * s.n_seg);
        }
        L.v = in.take(sizeof(double) * p.nv); L.t = in.take(sizeof(double) * p.nt); L.d = in.take(sizeof(double) * p.nd);
        L.ns = in.take(sizeof(int) * p.nt);
        const size_t m = (size_t)L.O * L.Tp;
        L.pos = in.take(sizeof(double) * 2 * m); L.cov = in.take(sizeof(double) * 4 * m); L.yaw = in.take(sizeof(double) * m);
        L.vel = in.take(sizeof(double) * m);
        L.len = in.take(sizeof(double) * L.O); L.wid = in.take(sizeof(double) * L.O); L.mass = in.take(sizeof(double) * L.O);
        L.plen = in.take(sizeof(int) * L.O); L.prot = in.take(sizeof(int) * L.O); L.resp = in.take(sizeof(int) * L.O);
        L.small = in.take(64);  // zeroed counters of the selection
        L.tile = scratch.take(Tile<double>::bytes(L.O > 0 ? L.O : 1, L.Tp > 0 ? L.Tp : 1));
        L.oconst = scratch.take(sizeof(double) * OC_COUNT * (L.O > 0 ? L.O : 1));
        L.prof = scratch.take(sizeof(double) * PF_COUNT * L.Tmax * (size_t)L.n_prof);
        L.meta = scratch.take(sizeof(double) * PM_COUNT * (size_t)L.n_prof);
        L.nskip = scratch.take(sizeof(int) * (size_t)L.n_prof);
        L.wc = scratch.take(sizeof(double) * L.Cs); L.wb = scratch.take(sizeof(float) * L.Cs); L.wm = scratch.take(L.Cs);
        L.wbd = scratch.take(sizeof(double) * L.Cs); L.wgf = scratch.take(L.Cs);
        L.blocks = scratch.take(select_blocks_bytes(L.Cs)); L.list = scratch.take(sizeof(int) * L.Cs);
        L.res = scratch.take(rescore_rec_bytes(L.Cs));
    }
    if (n_tiles >= (1ll << 30)) ETP_FAIL(c, ETP_ERR_INVALID, "too many tiles in one chunk");
    const size_t o_tscn = in.take(sizeof(int) * (size_t)n_tiles);
    const size_t in_bytes = in.off;
    const size_t o_bests = scratch.take(sizeof(BestRec) * n);
    ETP_CUDA(c, c->batch_dev.reserve(in_bytes + scratch.off));
    unsigned char* dev = c->batch_dev.as<unsigned char>();
    unsigned char* sdev = dev + in_bytes;
    // the pinned image must not be rewritten while its previous copy is in flight
    if (!c->batch_ev[slot]) ETP_CUDA(c, cudaEventCreateWithFlags(&c->batch_ev[slot], cudaEventDisableTiming));
    else ETP_CUDA(c, cudaEventSynchronize(c->batch_ev[slot]));
    if (c->batch_in[slot].cap < in_bytes) ETP_CUDA(c, c->batch_in[slot].reserve(in_bytes + in_bytes / 4));
    unsigned char* h = c->batch_in[slot].p;
    DevStep* steps = reinterpret_cast<DevStep*>(h + o_steps);
    PackArgs* packs = reinterpret_cast<PackArgs*>(h + o_pack);
    SelectBufs* sbs = reinterpret_cast<SelectBufs*>(h + o_sb);
    int* tile_first = reinterpret_cast<int*>(h + o_first);
    int* tile_scn = reinterpret_cast<int*>(h + o_tscn);
    const SelectBufs ctx_sb = select_bufs(c);
    const bool refine = precision == ETP_PRECISION_FP32 && !c->refine_off;
    int tile_at = 0;
    // ---- fill ---------------------------------------------------------------------------------------
    for (int i = 0; i < n; ++i) {
        const etp_scenario& s = sc[i];
        const etp_step& p = *s.step;
        const etp_obstacles& o = *s.obstacles;
        const Lay& L = lay[i];
        if (!L.shared_spline) {
            memcpy(h + L.knots, s.knots, sizeof(double) * (s.n_seg + 1));
            memcpy(h + L.cx, s.coef_x, sizeof(double) * 4 * s.n_seg);
            memcpy(h + L.cy, s.coef_y, sizeof(double) * 4 * s.n_seg);
        }
        memcpy(h + L.v, p.v_list, sizeof(double) * p.nv);
        memcpy(h + L.t, p.t_list, sizeof(double) * p.nt);
        memcpy(h + L.d, p.d_list, sizeof(double) * p.nd);
        memcpy(h + L.ns, p.n_samples, sizeof(int) * p.nt);
        const size_t m = (size_t)L.O * L.Tp;
        if (L.O > 0) {
            memcpy(h + L.pos, o.pos, sizeof(double) * 2 * m); memcpy(h + L.cov, o.cov, sizeof(double) * 4 * m);
            memcpy(h + L.yaw, o.yaw, sizeof(double) * m); memcpy(h + L.vel, o.vel, sizeof(double) * m);
            memcpy(h + L.len, o.length, sizeof(double) * L.O); memcpy(h + L.wid, o.width, sizeof(double) * L.O);
            memcpy(h + L.mass, o.mass, sizeof(double) * L.O); memcpy(h + L.plen, o.pred_len, sizeof(int) * L.O);
            memcpy(h + L.prot, o.is_protected, sizeof(int) * L.O); memcpy(h + L.resp, o.responsibility, sizeof(int) * L.O);
        }
        memset(h + L.small, 0, 64);
        // step origin and the fp32 rounding of a position in this scenario's coordinate range (as in etp_plan_step)
        double ox, oy, m_abs = 0.0, dmax = 0.0;
        if (L.O > 0) {
            ox = o.pos[0]; oy = o.pos[1];
            double bx0 = ox, bx1 = ox, by0 = oy, by1 = oy;
            for (int k = 0; k < L.O; ++k)
                for (int t = 0; t < o.pred_len[k]; ++t) {
                    const double* q = o.pos + ((size_t)k * L.Tp + t) * 2;
                    bx0 = std::fmin(bx0, q[0]); bx1 = std::fmax(bx1, q[0]); by0 = std::fmin(by0, q[1]); by1 = std::fmax(by1, q[1]);
                }
            m_abs = std::fmax(std::fmax(std::fabs(bx0 - ox), std::fabs(bx1 - ox)), std::fmax(std::fabs(by0 - oy), std::fabs(by1 - oy)));
        } else {
            ox = s.coef_x[0]; oy = s.coef_y[0];
        }
        for (int k = 0; k < s.n_seg; ++k) {
            const double ds = s.knots[k + 1] - s.knots[k];
            const double reach = 2.0 * std::fmax(std::fabs(s.coef_x[4 * k + 1]), std::fabs(s.coef_y[4 * k + 1])) * ds;
            m_abs = std::fmax(m_abs, std::fmax(std::fabs(s.coef_x[4 * k] - ox), std::fabs(s.coef_y[4 * k] - oy)) + reach);
        }
        for (int k = 0; k < p.nd; ++k) dmax = std::fmax(dmax, std::fabs(p.d_list[k]));
        DevStep st{};
        st.c_s = p.c_s; st.c_s_d = p.c_s_d; st.c_s_dd = p.c_s_dd; st.c_d = p.c_d; st.c_d_d = p.c_d_d; st.c_d_dd = p.c_d_dd;
        st.dt = p.dt; st.v_thr = p.v_thr;
        st.v_list = reinterpret_cast<const double*>(dev + L.v); st.t_list = reinterpret_cast<const double*>(dev + L.t);
        st.d_list = reinterpret_cast<const double*>(dev + L.d); st.nsamp = reinterpret_cast<const int*>(dev + L.ns);
        st.nv = p.nv; st.nt = p.nt; st.nd = p.nd; st.Tmax = L.Tmax;
        st.c_begin = 0; st.c_end = L.Cs; st.p_begin = 0; st.p_end = L.n_prof;
        st.shard_index = 0; st.shard_count = 1; st.n_local = L.Cs;
        st.vel_mode = p.velocity_cost_mode; st.vel_target = p.velocity_target; st.factor = p.cost_factor;
        st.eps_pos = 2.4e-7 * (m_abs + dmax + 16.0);
        st.origin_x = ox; st.origin_y = oy;
        st.boundary = c->boundary;
        st.boundary.lr1s_const = c->pr.lr1s_const; st.boundary.lr1s_speed = c->pr.lr1s_speed;
        st.boundary.off_x = ox - st.boundary.x0; st.boundary.off_y = oy - st.boundary.y0;
        st.knots = reinterpret_cast<const double*>(dev + L.knots); st.cx = reinterpret_cast<const double*>(dev + L.cx);
        st.cy = reinterpret_cast<const double*>(dev + L.cy); st.nseg = s.n_seg;
        st.prof = reinterpret_cast<double*>(sdev + L.prof); st.prof_meta = reinterpret_cast<double*>(sdev + L.meta);
        st.nskip = reinterpret_cast<int*>(sdev + L.nskip);
        st.has_pred = 1; st.O = L.O; st.Tp = L.Tp;
        st.obs = sdev + L.tile; st.obs_const = reinterpret_cast<const double*>(sdev + L.oconst);
        st.pred_len = reinterpret_cast<const int*>(dev + L.plen);
        st.raw_pos = reinterpret_cast<const double*>(dev + L.pos); st.raw_yaw = reinterpret_cast<const double*>(dev + L.yaw);
        st.raw_cov = reinterpret_cast<const double*>(dev + L.cov); st.raw_vel = reinterpret_cast<const double*>(dev + L.vel);
        st.ws_cost = reinterpret_cast<double*>(sdev + L.wc); st.ws_bound = reinterpret_cast<float*>(sdev + L.wb);
        st.ws_meta = sdev + L.wm; st.ws_bd = reinterpret_cast<double*>(sdev + L.wbd); st.ws_gf = sdev + L.wgf;
        st.refine = refine ? 1 : 0;
        steps[i] = st;
        packs[i] = PackArgs{L.O, L.Tp, st.pred_len, st.raw_pos, st.raw_cov, st.raw_yaw, st.raw_vel,
                            reinterpret_cast<const double*>(dev + L.len), reinterpret_cast<const double*>(dev + L.wid),
                            reinterpret_cast<const double*>(dev + L.mass), reinterpret_cast<const int*>(dev + L.prot),
                            reinterpret_cast<const int*>(dev + L.resp), c->pr.veh_m, ox, oy, sdev + L.tile,
                            reinterpret_cast<double*>(sdev + L.oconst)};
        SelectBufs sb;
        sb.blocks = sdev + L.blocks; sb.list = reinterpret_cast<int*>(sdev + L.list); sb.res = sdev + L.res;
        sb.hdr = reinterpret_cast<int*>(dev + L.small); sb.counters = reinterpret_cast<unsigned*>(dev + L.small + 16);
        sb.stats = ctx_sb.stats;
        sb.state = nullptr;  // one selection block per scenario: no second launch
        sbs[i] = sb;
        tile_first[i] = tile_at;
        for (int k = 0; k < L.tiles; ++k) tile_scn[tile_at + k] = i;
        tile_at += L.tiles;
    }
    // ---- run ----------------------------------------------------------------------------------------
    ETP_CUDA(c, cudaMemcpyAsync(dev, h, in_bytes, cudaMemcpyHostToDevice, c->stream));
    ETP_CUDA(c, cudaEventRecord(c->batch_ev[slot], c->stream));
    c->h2d_bytes += in_bytes;
    const DevStep* d_steps = reinterpret_cast<const DevStep*>(dev + o_steps);
    const SelectBufs* d_sbs = reinterpret_cast<const SelectBufs*>(dev + o_sb);
    BestRec* d_bests = reinterpret_cast<BestRec*>(sdev + o_bests);
    ETP_CUDA(c, launch_pack_obstacles_batch(precision, reinterpret_cast<const PackArgs*>(dev + o_pack), n, max_items, c->stream));
    ETP_CUDA(c, launch_profiles_batch(d_steps, c->pr, n, max_prof, max_T, c->stream));
    BatchArgs ba{d_steps, reinterpret_cast<const int*>(dev + o_tscn), reinterpret_cast<const int*>(dev + o_first), (int)n_tiles, n};
    unsigned* cnt = c->counters.as<unsigned>();
    ETP_CUDA(c, launch_score(precision, DevStep{}, c->pr, DevOut{}, cnt, c->num_sms, &c->last_grid, c->stream, &ba, max_smem));
    ETP_CUDA(c, launch_select_batch(d_steps, d_sbs, d_bests, cnt + 1, n, c->stream));
    c->launches += 4;
    if (refine) {
        ETP_CUDA(c, launch_rescore_batch(d_steps, c->pr, d_sbs, d_bests, n, max_T, max_O, c->stream));
        c->launches += 1;
    }
    ETP_CUDA(c, cudaMemcpyAsync(host_out, d_bests, sizeof(BestRec) * n, cudaMemcpyDeviceToHost, c->stream));
    c->d2h_bytes += sizeof(BestRec) * n;
    return ETP_OK;
}


// ---------------------------------------------------------------------------------------------
// etp_plan_cycle: one closed-loop cycle as "a sweep of one scenario" whose launch sequence is a captured graph
// ---------------------------------------------------------------------------------------------
// a planning cycle's selection runs in ONE CTA (the fused kernel's last), block after block: bounded so that it stays short
constexpr long long kCycleMaxCandidates = 16 * 4096;

bool cycle_fusable(const etp_ctx* c, const etp_cycle& cy) {
    if (!c->have_spline || !cy.step) return false;
    const etp_step& p = *cy.step;
    if (p.ext_level || p.bd_harm || p.cost_factor_list) return false;  // per-candidate host arrays: the step-by-step path
    if (p.nv < 1 || p.nt < 1 || p.nd < 1 || !p.v_list || !p.t_list || !p.d_list || !p.n_samples) return false;
    const long long dense = (long long)p.nv * p.nt * p.nd;
    if (c->cycle_split && dense > select_chunk()) return false;  // the measurement aid launches the one-block selection
    if (dense > kCycleMaxCandidates || p.c_begin != 0 || p.c_end != dense || p.shard_count > 1) return false;
    for (int i = 0; i < p.nt; ++i)
        if (p.n_samples[i] < 2 || p.n_samples[i] > 256) return false;
    if (cy.raw) {
        const etp_raw_predictions& o = *cy.raw;
        if (o.n_obstacles < 1 || o.max_pred < 1 || !o.pred_len || !o.pos || !o.cov || !o.length || !o.width || !o.init_orientation ||
            !o.init_velocity || !o.mass || !o.is_protected)
            return false;
        for (int i = 0; i < o.n_obstacles; ++i)
            if (o.pred_len[i] < 1 || o.pred_len[i] > o.max_pred || (o.is_protected[i] != 0 && o.is_protected[i] != 1)) return false;
    } else {
        const etp_obstacles& o = *cy.obstacles;
        if (o.n_obstacles < 0 || (o.n_obstacles > 0 && (o.max_pred < 1 || !o.pred_len || !o.pos || !o.cov || !o.yaw || !o.vel || !o.length ||
                                                        !o.width || !o.mass || !o.is_protected || !o.responsibility)))
            return false;
        for (int i = 0; i < o.n_obstacles; ++i)
            if (o.pred_len[i] < 1 || o.pred_len[i] > o.max_pred || (o.is_protected[i] != 0 && o.is_protected[i] != 1)) return false;
    }
    return true;
}

int plan_cycle_fused(etp_ctx* c, const etp_cycle& cy, int precision, etp_best* best, double* fields, int* n_samples) {
    const auto t_enter = std::chrono::steady_clock::now();
    const etp_step& p = *cy.step;
    const bool raw = cy.raw != nullptr;
    const int O = raw ? cy.raw->n_obstacles : cy.obstacles->n_obstacles;
    const int Tp = O > 0 ? (raw ? cy.raw->max_pred : cy.obstacles->max_pred) : 0;
    const int* plen = raw ? cy.raw->pred_len : cy.obstacles->pred_len;
    const double* pos = raw ? cy.raw->pos : cy.obstacles->pos;
    const double* cov = raw ? cy.raw->cov : cy.obstacles->cov;
    const int Tmax = etp_tmax(&p), n_prof = p.nv * p.nt, Cs = n_prof * p.nd, tiles = (Cs + 31) / 32;
    const size_t m = (size_t)O * Tp;
    const size_t smem = score_smem_bytes(precision, Tmax, O);
    if (smem > (size_t)c->smem_optin) ETP_FAIL(c, ETP_ERR_UNSUPPORTED, "the step needs %zu bytes of shared memory per CTA", smem);
    const bool refine = precision == ETP_PRECISION_FP32 && !c->refine_off;
    // ---- layout ----
    Arena in, sc;
    const size_t o_step = in.take(sizeof(DevStep)), o_pack = in.take(sizeof(PackArgs)), o_sb = in.take(sizeof(SelectBufs));
    const size_t o_enr = in.take(sizeof(EnrichArgs)), o_first = in.take(sizeof(int)), o_tscn = in.take(sizeof(int) * tiles);
    const size_t o_v = in.take(sizeof(double) * p.nv), o_t = in.take(sizeof(double) * p.nt), o_d = in.take(sizeof(double) * p.nd);
    const size_t o_ns = in.take(sizeof(int) * p.nt);
    const size_t o_pos = in.take(sizeof(double) * 2 * m), o_cov = in.take(sizeof(double) * 4 * m), o_plen = in.take(sizeof(int) * O);
    const size_t o_mass = in.take(sizeof(double) * O), o_prot = in.take(sizeof(int) * O);
    // raw: shape without margins and initial state; enriched: orientation, velocity, shape with margins, responsibility
    const size_t o_a = in.take(sizeof(double) * (raw ? O : m)), o_b = in.take(sizeof(double) * (raw ? O : m));
    const size_t o_c = in.take(sizeof(double) * O), o_e = in.take(sizeof(double) * O), o_resp_in = in.take(sizeof(int) * O);
    const size_t o_small = in.take(64);
    const size_t in_bytes = in.off;
    const size_t s_yaw = sc.take(sizeof(double) * m), s_vel = sc.take(sizeof(double) * m), s_len = sc.take(sizeof(double) * O);
    const size_t s_wid = sc.take(sizeof(double) * O), s_resp = sc.take(sizeof(int) * O);
    const size_t s_tile = sc.take(Tile<double>::bytes(O > 0 ? O : 1, Tp > 0 ? Tp : 1)), s_oconst = sc.take(sizeof(double) * OC_COUNT * (O > 0 ? O : 1));
    const size_t s_prof = sc.take(sizeof(double) * PF_COUNT * Tmax * (size_t)n_prof), s_meta = sc.take(sizeof(double) * PM_COUNT * (size_t)n_prof);
    const size_t s_nskip = sc.take(sizeof(int) * (size_t)n_prof);
    const size_t s_wc = sc.take(sizeof(double) * Cs), s_wb = sc.take(sizeof(float) * Cs), s_wm = sc.take(Cs), s_wbd = sc.take(sizeof(double) * Cs);
    const size_t s_wgf = sc.take(Cs), s_blocks = sc.take(select_blocks_bytes(Cs)), s_list = sc.take(sizeof(int) * Cs);
    const size_t s_res = sc.take(rescore_rec_bytes(Cs)), s_best = sc.take(sizeof(BestRec));
    const size_t s_state = sc.take(64);  // SelState: hand-over between the two halves of a selection over several blocks
    const size_t combo_bytes = sizeof(BestRec) + 8 + sizeof(double) * ETP_NUM_FIELDS * Tmax;
    const size_t s_combo = sc.take(combo_bytes);
    ETP_CUDA(c, c->cyc_dev.reserve(in_bytes + sc.off));
    if (c->cyc_in.cap < in_bytes) ETP_CUDA(c, c->cyc_in.reserve(in_bytes + in_bytes / 4));
    if (c->cyc_out.cap < combo_bytes) ETP_CUDA(c, c->cyc_out.reserve(combo_bytes + 256));
    unsigned char* dev = c->cyc_dev.as<unsigned char>();
    unsigned char* sdev = dev + in_bytes;
    unsigned char* h = c->cyc_in.p;
    // ---- fill the input image ----
    memcpy(h + o_v, p.v_list, sizeof(double) * p.nv);
    memcpy(h + o_t, p.t_list, sizeof(double) * p.nt);
    memcpy(h + o_d, p.d_list, sizeof(double) * p.nd);
    memcpy(h + o_ns, p.n_samples, sizeof(int) * p.nt);
    if (O > 0) {
        memcpy(h + o_pos, pos, sizeof(double) * 2 * m);
        memcpy(h + o_cov, cov, sizeof(double) * 4 * m);
        memcpy(h + o_plen, plen, sizeof(int) * O);
        if (raw) {
            memcpy(h + o_mass, cy.raw->mass, sizeof(double) * O); memcpy(h + o_prot, cy.raw->is_protected, sizeof(int) * O);
            memcpy(h + o_a, cy.raw->length, sizeof(double) * O); memcpy(h + o_b, cy.raw->width, sizeof(double) * O);
            memcpy(h + o_c, cy.raw->init_orientation, sizeof(double) * O); memcpy(h + o_e, cy.raw->init_velocity, sizeof(double) * O);
        } else {
            const etp_obstacles& o = *cy.obstacles;
            memcpy(h + o_mass, o.mass, sizeof(double) * O); memcpy(h + o_prot, o.is_protected, sizeof(int) * O);
            memcpy(h + o_a, o.yaw, sizeof(double) * m); memcpy(h + o_b, o.vel, sizeof(double) * m);
            memcpy(h + o_c, o.length, sizeof(double) * O); memcpy(h + o_e, o.width, sizeof(double) * O);
            memcpy(h + o_resp_in, o.responsibility, sizeof(int) * O);
        }
    }
    memset(h + o_small, 0, 64);
    double ox, oy, m_abs = 0.0, dmax = 0.0;
    if (O > 0) {
        ox = pos[0]; oy = pos[1];
        double bx0 = ox, bx1 = ox, by0 = oy, by1 = oy;
        for (int k = 0; k < O; ++k)
            for (int t = 0; t < plen[k]; ++t) {
                const double* q = pos + ((size_t)k * Tp + t) * 2;
                bx0 = std::fmin(bx0, q[0]); bx1 = std::fmax(bx1, q[0]); by0 = std::fmin(by0, q[1]); by1 = std::fmax(by1, q[1]);
            }
        m_abs = std::fmax(std::fmax(std::fabs(bx0 - ox), std::fabs(bx1 - ox)), std::fmax(std::fabs(by0 - oy), std::fabs(by1 - oy)));
    } else {
        ox = c->spline_pts.empty() ? 0.0 : c->spline_pts[0];
        oy = c->spline_pts.empty() ? 0.0 : c->spline_pts[1];
    }
    for (size_t i = 0; i + 2 < c->spline_pts.size(); i += 3)
        m_abs = std::fmax(m_abs, std::fmax(std::fabs(c->spline_pts[i] - ox), std::fabs(c->spline_pts[i + 1] - oy)) + c->spline_pts[i + 2]);
    for (int k = 0; k < p.nd; ++k) dmax = std::fmax(dmax, std::fabs(p.d_list[k]));
    DevStep st{};
    st.c_s = p.c_s; st.c_s_d = p.c_s_d; st.c_s_dd = p.c_s_dd; st.c_d = p.c_d; st.c_d_d = p.c_d_d; st.c_d_dd = p.c_d_dd;
    st.dt = p.dt; st.v_thr = p.v_thr;
    st.v_list = reinterpret_cast<const double*>(dev + o_v); st.t_list = reinterpret_cast<const double*>(dev + o_t);
    st.d_list = reinterpret_cast<const double*>(dev + o_d); st.nsamp = reinterpret_cast<const int*>(dev + o_ns);
    st.nv = p.nv; st.nt = p.nt; st.nd = p.nd; st.Tmax = Tmax;
    st.c_begin = 0; st.c_end = Cs; st.p_begin = 0; st.p_end = n_prof; st.shard_index = 0; st.shard_count = 1; st.n_local = Cs;
    st.vel_mode = p.velocity_cost_mode; st.vel_target = p.velocity_target; st.factor = p.cost_factor;
    st.eps_pos = 2.4e-7 * (m_abs + dmax + 16.0);
    st.origin_x = ox; st.origin_y = oy;
    st.boundary = c->boundary;
    st.boundary.lr1s_const = c->pr.lr1s_const; st.boundary.lr1s_speed = c->pr.lr1s_speed;
    st.boundary.off_x = ox - st.boundary.x0; st.boundary.off_y = oy - st.boundary.y0;
    st.knots = c->knots.as<double>(); st.cx = c->cx.as<double>(); st.cy = c->cy.as<double>(); st.nseg = c->nseg;
    st.prof = reinterpret_cast<double*>(sdev + s_prof); st.prof_meta = reinterpret_cast<double*>(sdev + s_meta);
    st.nskip = reinterpret_cast<int*>(sdev + s_nskip);
    st.has_pred = 1; st.O = O; st.Tp = Tp;
    st.obs = sdev + s_tile; st.obs_const = reinterpret_cast<const double*>(sdev + s_oconst);
    st.pred_len = reinterpret_cast<const int*>(dev + o_plen);
    const double* d_yaw = raw ? reinterpret_cast<const double*>(sdev + s_yaw) : reinterpret_cast<const double*>(dev + o_a);
    const double* d_vel = raw ? reinterpret_cast<const double*>(sdev + s_vel) : reinterpret_cast<const double*>(dev + o_b);
    const double* d_len = raw ? reinterpret_cast<const double*>(sdev + s_len) : reinterpret_cast<const double*>(dev + o_c);
    const double* d_wid = raw ? reinterpret_cast<const double*>(sdev + s_wid) : reinterpret_cast<const double*>(dev + o_e);
    const int* d_resp = raw ? reinterpret_cast<const int*>(sdev + s_resp) : reinterpret_cast<const int*>(dev + o_resp_in);
    st.raw_pos = reinterpret_cast<const double*>(dev + o_pos); st.raw_yaw = d_yaw;
    st.raw_cov = reinterpret_cast<const double*>(dev + o_cov); st.raw_vel = d_vel;
    st.ws_cost = reinterpret_cast<double*>(sdev + s_wc); st.ws_bound = reinterpret_cast<float*>(sdev + s_wb);
    st.ws_meta = sdev + s_wm; st.ws_bd = reinterpret_cast<double*>(sdev + s_wbd); st.ws_gf = sdev + s_wgf;
    st.refine = refine ? 1 : 0;
    // goal context of the step (cost factor, velocity cost, distance to the goal positions per candidate): its per-cycle part
    // travels in the step record, its geometry is uploaded when it differs from what is resident
    if (c->pr.w[ETP_W_DIST_TO_GOAL_POS] > 0.0 && !p.goal)
        ETP_FAIL(c, ETP_ERR_INVALID, "weights['dist_to_goal_pos'] > 0 needs the goal positions of the step (etp_step.goal)");
    {
        const int rg = stage_goal(c, p.goal, p.dt, Tmax, &st.goal);
        if (rg) return rg;
    }
    memcpy(h + o_step, &st, sizeof(st));
    const PackArgs pk{O, Tp, st.pred_len, st.raw_pos, st.raw_cov, d_yaw, d_vel, d_len, d_wid, reinterpret_cast<const double*>(dev + o_mass),
                      reinterpret_cast<const int*>(dev + o_prot), d_resp, c->pr.veh_m, ox, oy, sdev + s_tile,
                      reinterpret_cast<double*>(sdev + s_oconst)};
    memcpy(h + o_pack, &pk, sizeof(pk));
    if (raw) {
        const etp_raw_predictions& o = *cy.raw;
        const EnrichArgs ea{O, Tp, st.pred_len, st.raw_pos, reinterpret_cast<const double*>(dev + o_a), reinterpret_cast<const double*>(dev + o_b),
                            reinterpret_cast<const double*>(dev + o_c), reinterpret_cast<const double*>(dev + o_e), o.dt,
                            o.safety_margin_length, o.safety_margin_width, o.ego_x, o.ego_y, o.ego_orientation,
                            (c->pr.relaxed & ETP_RELAX_Q9_VIEW_CONE) ? 1 : 0, reinterpret_cast<double*>(sdev + s_yaw),
                            reinterpret_cast<double*>(sdev + s_vel), reinterpret_cast<double*>(sdev + s_len),
                            reinterpret_cast<double*>(sdev + s_wid), reinterpret_cast<int*>(sdev + s_resp)};
        memcpy(h + o_enr, &ea, sizeof(ea));
    }
    SelectBufs sb;
    sb.blocks = sdev + s_blocks; sb.list = reinterpret_cast<int*>(sdev + s_list); sb.res = sdev + s_res;
    sb.hdr = reinterpret_cast<int*>(dev + o_small); sb.counters = reinterpret_cast<unsigned*>(dev + o_small + 16);
    sb.stats = select_bufs(c).stats; sb.state = sdev + s_state;
    memcpy(h + o_sb, &sb, sizeof(sb));
    *reinterpret_cast<int*>(h + o_first) = 0;
    for (int k = 0; k < tiles; ++k) reinterpret_cast<int*>(h + o_tscn)[k] = 0;

    const auto t_filled = std::chrono::steady_clock::now();
    // ---- the launch sequence: replayed from the captured graph when nothing it bakes in has changed ----
    const DevStep* d_step = reinterpret_cast<const DevStep*>(dev + o_step);
    const SelectBufs* d_sb = reinterpret_cast<const SelectBufs*>(dev + o_sb);
    BestRec* d_best = reinterpret_cast<BestRec*>(sdev + s_best);
    unsigned char* d_combo = sdev + s_combo;
    const int n_launch = c->cycle_split ? 5 + (raw && O > 0 ? 1 : 0) + (refine ? 1 : 0) : 3;
    const bool timed = c->timing;  // CUDA-event timing of the stages (etp_get_stage_timing): launched one by one then
    const bool staged = c->cyc_ev[0] != nullptr;  // ETP_CYCLE_PROFILE=2
#define ETP_STAGE_EV(k) do { if (staged) ETP_CUDA(c, cudaEventRecord(c->cyc_ev[k], c->stream)); } while (0)
    // the result lands in pinned host memory straight from the last kernel (unified addressing: the device writes through the
    // host pointer), so the cycle is one upload and three launches: head | fused scoring + selection | tail
    unsigned char* h_combo = c->cyc_out.p;
    auto enqueue = [&]() -> int {
        ETP_STAGE_EV(0);
        ETP_CUDA(c, cudaMemcpyAsync(dev, h, in_bytes, cudaMemcpyHostToDevice, c->stream));
        ETP_STAGE_EV(1);
        BatchArgs ba{d_step, reinterpret_cast<const int*>(dev + o_tscn), reinterpret_cast<const int*>(dev + o_first), tiles, 1};
        if (!c->cycle_split) {  // the selection runs in the fused kernel's last CTA
            ba.sel = d_sb;
            ba.best = d_best;
            ba.done = c->counters.as<unsigned>() + 4;
        }
        if (c->cycle_split) {  // ETP_CYCLE_SPLIT=1: the stages as separate launches (measurement aid)
            if (raw) ETP_CUDA(c, launch_enrich_args(reinterpret_cast<const EnrichArgs*>(dev + o_enr), O, Tp, c->stream));
            ETP_STAGE_EV(2);
            ETP_CUDA(c, launch_pack_obstacles_batch(precision, reinterpret_cast<const PackArgs*>(dev + o_pack), 1, O * Tp, c->stream));
            if (timed) ETP_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
            ETP_STAGE_EV(3);
            ETP_CUDA(c, launch_profiles_batch(d_step, c->pr, 1, n_prof, Tmax, c->stream));
        } else {
            ETP_STAGE_EV(2);
            if (timed) ETP_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
            ETP_STAGE_EV(3);
            ETP_CUDA(c, launch_cycle_head(precision, d_step, c->pr, reinterpret_cast<const PackArgs*>(dev + o_pack),
                                          raw ? reinterpret_cast<const EnrichArgs*>(dev + o_enr) : nullptr, n_prof, O, Tp, Tmax, c->stream));
        }
        if (timed) ETP_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
        ETP_STAGE_EV(4);
        ETP_CUDA(c, launch_score(precision, DevStep{}, c->pr, DevOut{}, c->counters.as<unsigned>(), c->num_sms, &c->last_grid, c->stream, &ba, smem));
        if (timed) ETP_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
        ETP_STAGE_EV(5);
        if (c->cycle_split) ETP_CUDA(c, launch_select_batch(d_step, d_sb, d_best, c->counters.as<unsigned>() + 1, 1, c->stream));
        ETP_STAGE_EV(6);
        if (c->cycle_split) {
            if (refine) ETP_CUDA(c, launch_rescore_batch(d_step, c->pr, d_sb, d_best, 1, Tmax, O, c->stream));
            if (timed) ETP_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
            ETP_STAGE_EV(7);
            ETP_CUDA(c, launch_best_trajectory_steps(d_step, d_best, d_combo, c->stream));
            ETP_STAGE_EV(8);
            ETP_CUDA(c, cudaMemcpyAsync(c->cyc_out.p, d_combo, combo_bytes, cudaMemcpyDeviceToHost, c->stream));
        } else {
            ETP_STAGE_EV(7);
            ETP_CUDA(c, launch_cycle_tail(d_step, c->pr, d_sb, d_best, h_combo, refine, Tmax, O, c->stream));
            if (timed) ETP_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
            ETP_STAGE_EV(8);
        }
        ETP_STAGE_EV(9);
        return ETP_OK;
    };
#undef ETP_STAGE_EV
    const std::vector<long long> key = {precision, raw ? 1 : 0, O, Tp, p.nv, p.nt, p.nd, Tmax, (long long)in_bytes, (long long)(size_t)dev,
                                        (long long)(size_t)h, (long long)(size_t)c->cyc_out.p, (long long)c->config_version, refine ? 1 : 0,
                                        (long long)(size_t)c->stream, (long long)(size_t)c->knots.p, c->cycle_split ? 1 : 0};
    int r = ETP_OK;
    if (timed || staged) {
        if ((r = enqueue())) return r;
    } else if (!c->no_graph && c->cyc_exec && key == c->cyc_key) {
        ETP_CUDA(c, cudaGraphLaunch(c->cyc_exec, c->stream));
        c->graph_launches++;
    } else if (!c->no_graph && key == c->cyc_seen) {
        // second cycle of this shape: capture what the first one launched directly (every one-time runtime query is done)
        if (c->cyc_exec) { cudaGraphExecDestroy(c->cyc_exec); c->cyc_exec = nullptr; }
        if (c->cyc_graph) { cudaGraphDestroy(c->cyc_graph); c->cyc_graph = nullptr; }
        ETP_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
        r = enqueue();
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        if (r != ETP_OK || e != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            if (r == ETP_OK) ETP_FAIL(c, ETP_ERR_CUDA, "graph capture of the planning cycle failed: %s", cudaGetErrorString(e));
            return r;
        }
        c->cyc_graph = g;
        ETP_CUDA(c, cudaGraphInstantiate(&c->cyc_exec, g, 0));
        c->cyc_key = key;
        ETP_CUDA(c, cudaGraphLaunch(c->cyc_exec, c->stream));
        c->graph_launches++;
    } else {
        c->cyc_seen = key;
        if ((r = enqueue())) return r;
    }
    c->launches += n_launch;
    c->h2d_bytes += in_bytes;
    c->d2h_bytes += combo_bytes;  // written by the device into pinned host memory, or copied (ETP_CYCLE_SPLIT)
    const auto t_launched = std::chrono::steady_clock::now();
    ETP_CUDA(c, cudaStreamSynchronize(c->stream));
    if (staged) {
        static const char* names[9] = {"upload", "enrich", "pack", "profiles|head", "score", "select", "rescore", "trajectory|tail", "read-back"};
        for (int k = 0; k < 9; ++k) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c->cyc_ev[k], c->cyc_ev[k + 1]);
            c->cyc_stage[k] += ms * 1e3;
        }
        if (++c->cyc_stage_n == 50) {
            fprintf(stderr, "etp_plan_cycle stages (mean of 50, us, launched one by one):");
            for (int k = 0; k < 9; ++k) { fprintf(stderr, " %s %.1f", names[k], c->cyc_stage[k] / 50); c->cyc_stage[k] = 0.0; }
            fprintf(stderr, "\n");
            c->cyc_stage_n = 0;
        }
    }
    if (c->cycle_profile) {
        const auto t_done = std::chrono::steady_clock::now();
        c->cyc_prof[0] += std::chrono::duration<double, std::micro>(t_filled - t_enter).count();
        c->cyc_prof[1] += std::chrono::duration<double, std::micro>(t_launched - t_filled).count();
        c->cyc_prof[2] += std::chrono::duration<double, std::micro>(t_done - t_launched).count();
        if (++c->cyc_prof_n == 50) {
            fprintf(stderr, "etp_plan_cycle (mean of 50): input image %.1f us, launches %.1f us, wait for the device %.1f us, graph launches so far %llu\n",
                    c->cyc_prof[0] / 50, c->cyc_prof[1] / 50, c->cyc_prof[2] / 50, c->graph_launches);
            c->cyc_prof[0] = c->cyc_prof[1] = c->cyc_prof[2] = 0.0;
            c->cyc_prof_n = 0;
        }
    }
    // the context now holds this step like after etp_plan_step (etp_get_trajectory etc. keep working on it)
    c->st = st;
    c->precision = precision;
    c->has_pred = 1; c->O = O; c->Tp = Tp;
    c->have_step = true;
    c->last_given = false;
    const BestRec* b = reinterpret_cast<const BestRec*>(c->cyc_out.p);
    best->key_hi = b->key_hi; best->key_lo = b->key_lo; best->cost = b->cost; best->index = b->index;
    best->level = b->level; best->n_scored = b->n_scored; best->list_index = b->list_index;
    if (b->index < 0)
        ETP_FAIL(c, ETP_ERR_NO_CANDIDATE, "every candidate of the range was skipped (ds <= |dT|); the reference raises ValueError here");
    memcpy(fields, c->cyc_out.p + sizeof(BestRec) + 8, sizeof(double) * ETP_NUM_FIELDS * Tmax);
    if (n_samples) memcpy(n_samples, c->cyc_out.p + sizeof(BestRec), sizeof(int));
    return ETP_OK;
}

}  // namespace

int etp_plan_batch(etp_ctx* c, int n, const etp_scenario* sc, int precision, etp_best* bests) {
    if (!c || n < 0 || (n > 0 && (!sc || !bests))) return ETP_ERR_INVALID;
    if (!c->have_params) ETP_FAIL(c, ETP_ERR_STATE, "etp_set_params must precede etp_plan_batch");
    if (precision != ETP_PRECISION_FP32 && precision != ETP_PRECISION_FP64) ETP_FAIL(c, ETP_ERR_INVALID, "precision");
    if (n == 0) return ETP_OK;
    ETP_CUDA(c, cudaSetDevice(c->device));
    {
        bool fused = getenv("ETP_BATCH_STREAMS") == nullptr;
        for (int i = 0; i < n && fused; ++i) fused = batchable(sc[i]);
        if (fused) {
            // chunks of scenarios: every stage of a chunk is ONE launch over all its scenarios, and the next chunk's inputs
            // are assembled on the host while this one runs.  A chunk's arena is reused by the next one (stream order).
            constexpr int kChunk = 512;
            ETP_CUDA(c, c->batch_host.reserve(sizeof(BestRec) * (size_t)n));
            BestRec* host = reinterpret_cast<BestRec*>(c->batch_host.p);
            c->precision = precision;
            int slot = 0;
            const auto t0 = std::chrono::steady_clock::now();
            // optional CUDA-event timing of the whole sweep on the stream (etp_get_stage_timing: all of it under ms[1])
            if (c->timing) {
                ETP_CUDA(c, cudaEventRecord(c->ev[0], c->stream));
                ETP_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
            }
            for (int at = 0; at < n; at += kChunk, slot ^= 1) {
                const int rc = plan_chunk_fused(c, std::min(kChunk, n - at), sc + at, precision, host + at, slot);
                if (rc != ETP_OK) {
                    cudaStreamSynchronize(c->stream);
                    return rc;
                }
            }
            if (c->timing) {
                ETP_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
                ETP_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
                c->have_step = true;
            }
            const auto t1 = std::chrono::steady_clock::now();
            ETP_CUDA(c, cudaStreamSynchronize(c->stream));
            if (getenv("ETP_BATCH_PROFILE")) {
                const auto t2 = std::chrono::steady_clock::now();
                fprintf(stderr, "etp_plan_batch: %d scenarios, host assembly + launches %.3f ms, then waited %.3f ms for the device\n", n,
                        std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
            }
            for (int i = 0; i < n; ++i) {
                const BestRec& b = host[i];
                bests[i].key_hi = b.key_hi; bests[i].key_lo = b.key_lo; bests[i].cost = b.cost; bests[i].index = b.index;
                bests[i].level = b.level; bests[i].n_scored = b.n_scored; bests[i].list_index = b.list_index;
            }
            return ETP_OK;
        }
    }
    constexpr int kPool = 8;  // scenarios in flight
    while ((int)c->pool.size() < kPool) {
        etp_ctx* k = nullptr;
        const int rc = etp_create(&k, c->device);
        if (rc != ETP_OK) ETP_FAIL(c, rc, "etp_plan_batch: could not create a child context");
        c->pool.push_back(k);
    }
    ETP_CUDA(c, c->batch_host.reserve(sizeof(BestRec) * (size_t)n));
    BestRec* host = reinterpret_cast<BestRec*>(c->batch_host.p);
    // the road boundary of this context (grid and triangles stay in its buffers) serves every scenario of the batch;
    // its upload ran on this context's stream
    if (c->boundary.check != ETP_COLLISION_HOST) ETP_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) {
        etp_ctx* k = c->pool[i % kPool];
        const etp_scenario& s = sc[i];
        if (!s.step || !s.obstacles || !s.knots || !s.coef_x || !s.coef_y) ETP_FAIL(c, ETP_ERR_INVALID, "scenario %d incomplete", i);
        k->pr = c->pr;
        k->boundary = c->boundary;
        k->have_params = true;
        int rc = ETP_OK;
        // the spline travels with every scenario (8 KB): a cache keyed on the caller's pointer would mistake a new map at a
        // recycled address for the old one
        if ((rc = etp_set_spline(k, s.knots, s.coef_x, s.coef_y, s.n_seg)) || (rc = etp_set_obstacles(k, s.obstacles, precision)) ||
            (rc = etp_plan_step(k, s.step, nullptr))) {
            for (etp_ctx* q : c->pool) cudaStreamSynchronize(q->stream);  // nothing of this call may still write into batch_host
            ETP_FAIL(c, rc, "scenario %d: %s", i, k->err.c_str());
        }
        ETP_CUDA(c, cudaMemcpyAsync(host + i, k->best_dev.p, sizeof(BestRec), cudaMemcpyDeviceToHost, k->stream));
    }
    for (etp_ctx* k : c->pool) ETP_CUDA(c, cudaStreamSynchronize(k->stream));
    for (int i = 0; i < n; ++i) {
        const BestRec& b = host[i];
        bests[i].key_hi = b.key_hi; bests[i].key_lo = b.key_lo; bests[i].cost = b.cost; bests[i].index = b.index;
        bests[i].level = b.level; bests[i].n_scored = b.n_scored; bests[i].list_index = b.list_index;
    }
    return ETP_OK;
}

int etp_plan_cycle(etp_ctx* c, const etp_cycle* cy, int precision, etp_best* best, double* fields, int* n_samples) {
    if (!c || !cy || !cy->step || !best || !fields || (!cy->raw) == (!cy->obstacles)) return ETP_ERR_INVALID;
    if (precision != ETP_PRECISION_FP32 && precision != ETP_PRECISION_FP64) ETP_FAIL(c, ETP_ERR_INVALID, "precision");
    if (!c->have_params || !c->have_spline) ETP_FAIL(c, ETP_ERR_STATE, "etp_set_params / etp_set_spline must precede etp_plan_cycle");
    ETP_CUDA(c, cudaSetDevice(c->device));
    if (cycle_fusable(c, *cy)) return plan_cycle_fused(c, *cy, precision, best, fields, n_samples);
    int r = cy->raw ? etp_set_raw_predictions(c, cy->raw, precision) : etp_set_obstacles(c, cy->obstacles, precision);
    if (r) return r;
    if ((r = etp_plan_step(c, cy->step, nullptr))) return r;
    return etp_get_best_trajectory(c, best, fields, n_samples);
}

int etp_get_best(etp_ctx* c, etp_best* best) {
    if (!c || !best) return ETP_ERR_INVALID;
    if (!c->have_step) ETP_FAIL(c, ETP_ERR_STATE, "no planning step has run on this context");
    ETP_CUDA(c, cudaSetDevice(c->device));
    ETP_CUDA(c, cudaMemcpyAsync(c->best_host.p, c->best_dev.p, sizeof(BestRec), cudaMemcpyDeviceToHost, c->stream));
    c->d2h_bytes += sizeof(BestRec);
    ETP_CUDA(c, cudaStreamSynchronize(c->stream));
    const BestRec* b = reinterpret_cast<const BestRec*>(c->best_host.p);
    best->key_hi = b->key_hi; best->key_lo = b->key_lo; best->cost = b->cost; best->index = b->index;
    best->level = b->level; best->n_scored = b->n_scored; best->list_index = b->list_index;
    if (b->index < 0)
        ETP_FAIL(c, ETP_ERR_NO_CANDIDATE, "every candidate of the range was skipped (ds <= |dT|); the reference raises ValueError here");
    return ETP_OK;
}

int etp_get_trajectory(etp_ctx* c, int index, double* fields, int* n_samples) {
    if (!c || !fields) return ETP_ERR_INVALID;
    if (!c->have_step) ETP_FAIL(c, ETP_ERR_STATE, "no planning step has run on this context");
    const DevStep& st = c->st;
    if (c->last_given) ETP_FAIL(c, ETP_ERR_STATE, "the last call scored caller-provided trajectories: the caller holds them");
    if (index < st.c_begin || index >= st.c_end) ETP_FAIL(c, ETP_ERR_INVALID, "candidate %d outside the last scored range", index);
    ETP_CUDA(c, cudaSetDevice(c->device));
    const size_t bytes = sizeof(double) * ETP_NUM_FIELDS * st.Tmax;
    ETP_CUDA(c, c->traj_dev.reserve(bytes));
    ETP_CUDA(c, c->traj_host.reserve(bytes + sizeof(int)));
    c->launches += 1;
    ETP_CUDA(c, launch_trajectory(st, index, c->traj_dev.as<double>(), c->traj_n.as<int>(), c->stream));
    ETP_CUDA(c, cudaMemcpyAsync(c->traj_host.p, c->traj_dev.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    ETP_CUDA(c, cudaMemcpyAsync(c->traj_host.p + bytes, c->traj_n.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    c->d2h_bytes += bytes + sizeof(int);
    ETP_CUDA(c, cudaStreamSynchronize(c->stream));
    memcpy(fields, c->traj_host.p, bytes);
    if (n_samples) memcpy(n_samples, c->traj_host.p + bytes, sizeof(int));
    return ETP_OK;
}

int etp_get_best_trajectory(etp_ctx* c, etp_best* best, double* fields, int* n_samples) {
    if (!c || !best || !fields) return ETP_ERR_INVALID;
    if (!c->have_step) ETP_FAIL(c, ETP_ERR_STATE, "no planning step has run on this context");
    if (c->last_given) ETP_FAIL(c, ETP_ERR_STATE, "the last call scored caller-provided trajectories: the caller holds them");
    const DevStep& st = c->st;
    ETP_CUDA(c, cudaSetDevice(c->device));
    const size_t fbytes = sizeof(double) * ETP_NUM_FIELDS * st.Tmax, bytes = sizeof(BestRec) + 8 + fbytes;
    ETP_CUDA(c, c->traj_dev.reserve(bytes));
    ETP_CUDA(c, c->traj_host.reserve(bytes));
    c->launches += 1;
    ETP_CUDA(c, launch_best_trajectory(st, c->best_dev.as<BestRec>(), c->traj_dev.as<unsigned char>(), c->stream));
    ETP_CUDA(c, cudaMemcpyAsync(c->traj_host.p, c->traj_dev.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    c->d2h_bytes += bytes;
    ETP_CUDA(c, cudaStreamSynchronize(c->stream));
    const BestRec* b = reinterpret_cast<const BestRec*>(c->traj_host.p);
    best->key_hi = b->key_hi; best->key_lo = b->key_lo; best->cost = b->cost; best->index = b->index;
    best->level = b->level; best->n_scored = b->n_scored; best->list_index = b->list_index;
    if (b->index < 0)
        ETP_FAIL(c, ETP_ERR_NO_CANDIDATE, "every candidate of the range was skipped (ds <= |dT|); the reference raises ValueError here");
    memcpy(fields, c->traj_host.p + sizeof(BestRec) + 8, fbytes);
    if (n_samples) memcpy(n_samples, c->traj_host.p + sizeof(BestRec), sizeof(int));
    return ETP_OK;
}

int etp_best_device_ptr(etp_ctx* c, void** ptr) {
    if (!c || !ptr) return ETP_ERR_INVALID;
    *ptr = c->best_dev.p;
    return ETP_OK;
}

int etp_enable_timing(etp_ctx* c, int enable) {
    if (!c) return ETP_ERR_INVALID;
    ETP_CUDA(c, cudaSetDevice(c->device));
    if (enable && !c->ev[0])
        for (int i = 0; i < 4; ++i) ETP_CUDA(c, cudaEventCreate(&c->ev[i]));
    c->timing = enable != 0;
    return ETP_OK;
}

int etp_get_timing(etp_ctx* c, float* profile_ms, float* score_ms) {
    float ms[3] = {0.f, 0.f, 0.f};
    const int r = etp_get_stage_timing(c, ms, 3);
    if (r) return r;
    if (profile_ms) *profile_ms = ms[0];
    if (score_ms) *score_ms = ms[1];
    return ETP_OK;
}

int etp_get_stage_timing(etp_ctx* c, float* ms, int n) {
    if (!c || !ms || n < 1) return ETP_ERR_INVALID;
    if (!c->timing || !c->have_step) ETP_FAIL(c, ETP_ERR_STATE, "timing not enabled or no step run");
    ETP_CUDA(c, cudaEventSynchronize(c->ev[3]));
    for (int i = 0; i < 3 && i < n; ++i) ETP_CUDA(c, cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
    return ETP_OK;
}

int etp_debug_counters(etp_ctx* c, uint64_t* out, int n) {
    if (!c || !out || n < 1) return ETP_ERR_INVALID;
    ETP_CUDA(c, cudaSetDevice(c->device));
    ETP_CUDA(c, cudaStreamSynchronize(c->stream));
    unsigned long long v[1] = {0};
    ETP_CUDA(c, read_debug_counters(v, 1));
    out[0] = v[0];
    if (n > 1) {
        unsigned long long st[3] = {0, 0, 0};
        ETP_CUDA(c, cudaMemcpy(st, select_bufs(c).stats, sizeof(st), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 3 && 1 + i < n; ++i) out[1 + i] = st[i];
    }
    if (n > 4) out[4] = c->h2d_bytes;
    if (n > 5) out[5] = c->d2h_bytes;
    if (n > 6) out[6] = c->launches;
    if (n > 7) out[7] = c->graph_launches;
    return ETP_OK;
}

int etp_combine_best_device(etp_ctx* c, const etp_best* gathered, int n, int consecutive_ranges) {
    if (!c || !gathered || n < 1) return ETP_ERR_INVALID;
    if (!c->have_step) ETP_FAIL(c, ETP_ERR_STATE, "no planning step has run on this context");
    static_assert(sizeof(etp_best) == sizeof(BestRec), "etp_best and the device record share one layout");
    ETP_CUDA(c, cudaSetDevice(c->device));
    c->launches += 1;
    ETP_CUDA(c, launch_combine_best(reinterpret_cast<const BestRec*>(gathered), n, consecutive_ranges, c->best_dev.as<BestRec>(), c->stream));
    return ETP_OK;
}

int etp_combine_best(const etp_best* b, int n, int consecutive_ranges, etp_best* out) {
    if (!b || !out || n < 1) return ETP_ERR_INVALID;
    int win = -1, scored_before_win = 0, scored = 0;
    for (int i = 0; i < n; ++i) {
        if (b[i].index >= 0 &&
            (win < 0 || b[i].key_hi < b[win].key_hi || (b[i].key_hi == b[win].key_hi && b[i].key_lo < b[win].key_lo))) {
            win = i;
            scored_before_win = scored;
        }
        scored += b[i].n_scored;
    }
    if (win < 0) {
        *out = b[0];
        out->index = -1;
        out->n_scored = scored;
        return ETP_ERR_NO_CANDIDATE;
    }
    *out = b[win];
    out->n_scored = scored;
    // interleaved shards of one range already report the global rank; consecutive ranges add the preceding counts
    out->list_index = (consecutive_ranges ? scored_before_win : 0) + b[win].list_index;
    return ETP_OK;
}

}  // extern "C"
