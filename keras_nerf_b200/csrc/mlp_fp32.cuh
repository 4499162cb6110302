// fp32 (SIMT) MLP path: workspace plan + core entry points shared with pipeline.cu
#pragma once
#include "common.cuh"

namespace knerf {

// C[M,N] = epi( A1[M,K1] @ op(B1) + A2[M,K2] @ op(B2) + bias ): one Dense layer forward (op(B) = B[K,N], the Keras
// kernel layout) or its dgrad (op(B) = B[N,K]^T); the two sources express the [h, x] concats without materialising them
enum Epilogue { EPI_NONE = 0, EPI_RELU = 1, EPI_SIGMOID = 2, EPI_MASK = 3 };
struct GemmArgs {
  const float* A1; int lda1; int K1; const float* B1; int ldb1;
  const float* A2; int lda2; int K2; const float* B2; int ldb2;
  const float* bias; float* C; int ldc; int64_t M; int N; int epi;
  const float* mask; int ldmask;   // EPI_MASK: C = acc * (mask > 0)
};

// ---- KNERF_FP32_TC: the same GEMMs on tcgen05 with 3-way bf16 split operands (mlp_fp32_tc.cu) ----
size_t tcx_blob_bytes(int N, int K);                 // split weight operand blob of a [K x N] block
int tcx_pack(const float* src, int ld, int N, int K, bool trans, void* blob, cudaStream_t st);
bool tcx_gemm_eligible(const GemmArgs& g);
int launch_gemm_tc(const GemmArgs& g, const void* blob1, const void* blob2, cudaStream_t st);
bool tcx_wgrad_eligible(int K, int N);
// db != nullptr: also db[N] += column sums of Z
int launch_wgrad_tc(const float* A, int lda, int K, const float* Z, int ldz, int N, int64_t M, float* dW, int ldw,
                    float* db, cudaStream_t st);

struct Fp32Plan {
  int64_t rows;
  int ldx, ldd;                       // leading dims of the encoded inputs inside the workspace
  size_t off_x0, off_dir;             // PE(xyz) [rows, ldx], PE(dir) [rows, ldd]
  size_t off_h[kMaxLayers];           // post-ReLU hidden activations [rows, U] (ping-pong when !training)
  size_t off_f, off_g;                // features [rows, U], rgb_features [rows, U/2]
  size_t off_d0, off_d1, off_dg;      // backward scratch (training only)
  // KNERF_FP32_TC: split weight operand blobs, rebuilt by every forward call.  fwd1/fwd2: the hidden / encoding rows
  // of layer i's kernel as a forward B operand; bwd: the hidden rows as a dgrad B operand (training only)
  bool tc;
  size_t off_fwd1[kMaxLayers], off_fwd2[kMaxLayers], off_bwd[kMaxLayers];
  size_t total;                       // bytes
};

Fp32Plan make_fp32_plan(const Model& m, int64_t rows, bool training, bool tc = false);

// (p.tc: the eligible GEMMs run on the tensor cores; `training` also prepares the dgrad operand blobs)
int fp32_forward_core(const Model& m, const float* params, const float* X0, int ldx, const float* DIR, int ldd,
                      int64_t rows, char* ws, const Fp32Plan& p, float* out_rgb, int ld_rgb, float* out_sigma,
                      int ld_sigma, cudaStream_t st, bool training = false);

int fp32_backward_core(const Model& m, const float* params, const float* X0, int ldx, const float* DIR, int ldd,
                       const float* d_pre, int64_t rows, char* ws, const Fp32Plan& p, float* grads,
                       cudaStream_t st);

}  // namespace knerf
