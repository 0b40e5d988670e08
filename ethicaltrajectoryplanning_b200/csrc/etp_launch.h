// Host-callable launchers implemented in etp_kernels.cu / etp_score.cu (internal to the shared library).
#pragma once

#include <cuda_runtime.h>

#include "etp_device.cuh"

namespace etp {

constexpr int kMaxGrid = 4096;  // upper bound of persistent CTAs (148 SMs x resident CTAs per SM)

cudaError_t launch_pack_obstacles(int precision, int O, int Tp, const int* pred_len, const double* pos, const double* cov,
                                  const double* yaw, const double* vel, const double* length, const double* width,
                                  const double* mass, const int* prot, const int* resp, double veh_m, double ox, double oy,
                                  void* tile, double* oconst, cudaStream_t stream);

cudaError_t launch_profiles(const DevStep& st, const DevParams& pr, int* nskip, cudaStream_t stream);

// scenario sweep: one launch per stage for all scenarios of a chunk (arrays in device memory)
cudaError_t launch_pack_obstacles_batch(int precision, const PackArgs* args, int n_scn, int max_items, cudaStream_t stream);
cudaError_t launch_profiles_batch(const DevStep* steps, const DevParams& pr, int n_scn, int max_prof, int max_T, cudaStream_t stream);

// fused kernel: scores the candidates of `st`, materialises `out`, leaves the selection workspace (st.ws_*);
// counters[1] is its tile counter (re-armed by select_kernel)
// `ba` != null: the scenario sweep -- the tiles of ba->n_tiles come from many steps (ba->steps), `st` is ignored,
// batch_smem = the largest score_smem_bytes of the scenarios
cudaError_t launch_score(int precision, const DevStep& st, const DevParams& pr, const DevOut& out, unsigned* counters, int num_sms,
                         int* grid_out, cudaStream_t stream, const BatchArgs* ba = nullptr, size_t batch_smem = 0);

// buffers of the selection (etp_select.cu)
struct SelectBufs {
    void* blocks;               // per-block records of select_kernel (select_blocks_bytes)
    int* list;                  // [n_local] columns left for the exact re-scoring
    void* res;                  // their exact keys (rescore_rec_bytes)
    int* hdr;                   // [0] candidates left for rescore_kernel, [1] non-skipped candidates, [2] fill of `list` during select_kernel
    unsigned* counters;         // [2] completion counters of the two kernels
    unsigned long long* stats;  // [0] candidates re-scored, [1] of them outside their fp32 bound, [2] largest error / bound (float bits)
    void* state;                // hand-over between the two selection launches of a step with several blocks (32 bytes)
};
size_t select_blocks_bytes(int n_local);
size_t rescore_rec_bytes(int n_local);
size_t rescore_smem_bytes(int Tmax, int O);

// arg-min over the workspace -> *best_out, or (fp32 mode) the candidates that can still win -> rescore_kernel
cudaError_t launch_select(const DevStep& st, const SelectBufs& sb, BestRec* best_out, const int* nskip, unsigned* tile_counter,
                          bool given, bool always_rescore, cudaStream_t stream);
// scenario sweep: per-scenario buffers in device arrays; every scenario has at most select_chunk() candidates
int select_chunk();
cudaError_t launch_select_batch(const DevStep* steps, const SelectBufs* sbs, BestRec* bests, unsigned* tile_counter, int n_scn,
                                cudaStream_t stream);
cudaError_t launch_rescore_batch(const DevStep* steps, const DevParams& pr, const SelectBufs* sbs, BestRec* bests, int n_scn, int max_T,
                                 int max_O, cudaStream_t stream);
cudaError_t launch_rescore(int precision, const DevStep& st, const DevParams& pr, const DevOut& out, const SelectBufs& sb,
                           BestRec* best_out, const int* nskip, int num_sms, cudaStream_t stream);

// dynamic shared memory one CTA of the scoring kernel needs for (Tmax, O)
size_t score_smem_bytes(int precision, int Tmax, int O);

cudaError_t read_debug_counters(unsigned long long* out, int n);

cudaError_t launch_enrich(int O, int Tp, const int* pred_len, const double* pos, const double* raw_len, const double* raw_wid,
                          const double* ori0, const double* vel0, double dt, double margin_l, double margin_w, double ego_x,
                          double ego_y, double ego_ori, bool wrap_view_cone, double* yaw, double* vel, double* length, double* width, int* resp,
                          cudaStream_t stream);

cudaError_t launch_principles(int n, const double* vals /*[4][n]*/, const int* resp, double bd, double eps, double scale,
                              bool relaxed_gate, double* out5, cudaStream_t stream);

cudaError_t launch_combine_best(const BestRec* recs, int n, int consecutive, BestRec* out, cudaStream_t stream);

cudaError_t launch_best_trajectory(const DevStep& st, const BestRec* best, unsigned char* combo, cudaStream_t stream);

// forms that take their step / arguments from device memory (planning cycle as a captured graph)
cudaError_t launch_enrich_args(const EnrichArgs* args, int O, int Tp, cudaStream_t stream);
// planning cycle: [profiles | enrich + pack per obstacle] in one launch; [re-score survivors | winner's trajectory] in one launch
cudaError_t launch_cycle_head(int precision, const DevStep* steps, const DevParams& pr, const PackArgs* pack, const EnrichArgs* enrich,
                              int n_prof, int O, int Tp, int Tmax, cudaStream_t stream);
cudaError_t launch_cycle_tail(const DevStep* steps, const DevParams& pr, const SelectBufs* sbs, BestRec* bests, unsigned char* combo,
                              bool refine, int Tmax, int O, cudaStream_t stream);
cudaError_t launch_best_trajectory_steps(const DevStep* steps, const BestRec* best, unsigned char* combo, cudaStream_t stream);

cudaError_t launch_trajectory(const DevStep& st, int c_dense, double* fields, int* n_out, cudaStream_t stream);

}  // namespace etp
