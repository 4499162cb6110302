// fp32 (SIMT) MLP path: workspace plan + core entry points shared with pipeline.cu
#pragma once
#include "common.cuh"

namespace knerf {

struct Fp32Plan {
  int64_t rows;
  int ldx, ldd;                       // leading dims of the encoded inputs inside the workspace
  size_t off_x0, off_dir;             // PE(xyz) [rows, ldx], PE(dir) [rows, ldd]
  size_t off_h[kMaxLayers];           // post-ReLU hidden activations [rows, U] (ping-pong when !training)
  size_t off_f, off_g;                // features [rows, U], rgb_features [rows, U/2]
  size_t off_d0, off_d1, off_dg;      // backward scratch (training only)
  size_t total;                       // bytes
};

Fp32Plan make_fp32_plan(const Model& m, int64_t rows, bool training);

int fp32_forward_core(const Model& m, const float* params, const float* X0, int ldx, const float* DIR, int ldd,
                      int64_t rows, char* ws, const Fp32Plan& p, float* out_rgb, int ld_rgb, float* out_sigma,
                      int ld_sigma, cudaStream_t st);

int fp32_backward_core(const Model& m, const float* params, const float* X0, int ldx, const float* DIR, int ldd,
                       const float* d_pre, int64_t rows, char* ws, const Fp32Plan& p, float* grads,
                       cudaStream_t st);

}  // namespace knerf
