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

cudaError_t launch_score(int precision, const DevStep& st, const DevParams& pr, const DevOut& out, BestRec* block_best,
                         unsigned* counters, BestRec* best_out, const int* nskip, int num_sms, int* grid_out,
                         cudaStream_t stream);

// dynamic shared memory one CTA of the scoring kernel needs for (Tmax, O)
size_t score_smem_bytes(int precision, int Tmax, int O);

cudaError_t read_debug_counters(unsigned long long* out, int n);

cudaError_t launch_enrich(int O, int Tp, const int* pred_len, const double* pos, const double* raw_len, const double* raw_wid,
                          const double* ori0, const double* vel0, double dt, double margin_l, double margin_w, double ego_x,
                          double ego_y, double ego_ori, bool wrap_view_cone, double* yaw, double* vel, double* length, double* width, int* resp,
                          cudaStream_t stream);

cudaError_t launch_principles(int n, const double* vals /*[4][n]*/, const int* resp, double bd, double eps, double scale,
                              bool relaxed_gate, double* out5, cudaStream_t stream);

cudaError_t launch_best_trajectory(const DevStep& st, const BestRec* best, unsigned char* combo, cudaStream_t stream);

cudaError_t launch_trajectory(const DevStep& st, int c_dense, double* fields, int* n_out, cudaStream_t stream);

}  // namespace etp
