// a4 / a5: positional encoding and ray-point construction
// (keras_nerf/model/nerf/utils.py:176-186, :188-210).  fp32 parity path only: the tcgen05 MLP fuses
// this into its first-layer prologue and never writes encodings to HBM.
#include "common.cuh"

namespace knerf {

// value of encoding column `col` (< dim*(1+2L)) for input vector v[dim]:
// [v, sin(2^0 v), cos(2^0 v), sin(2^1 v), cos(2^1 v), ...]  -- no pi, full-accuracy sinf/cosf
// (|arg| reaches ~3e3 at L=10: fast intrinsics / double-angle recurrences break the 1e-5 budget).
__device__ __forceinline__ float pe_value(const float* v, int dim, int col) {
  if (col < dim) return v[col];
  const int k = col - dim;
  const int blk = k / dim, comp = k - blk * dim;
  const float arg = ldexpf(v[comp], blk >> 1);   // 2.0**i * x, exact
  return (blk & 1) ? cosf(arg) : sinf(arg);
}

__global__ void __launch_bounds__(256) posenc_kernel(const float* __restrict__ x, int64_t n_rows, int dim,
                                                     int L, float* __restrict__ out, int ld_out) {
  const int width = dim * (1 + 2 * L);
  const int64_t total = n_rows * ld_out;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = g / ld_out;
    const int col = (int)(g - row * ld_out);
    float val = 0.f;
    if (col < width) {
      if (col < dim) {
        val = x[row * dim + col];
      } else {
        const int k = col - dim;
        const int blk = k / dim, comp = k - blk * dim;
        const float arg = ldexpf(x[row * dim + comp], blk >> 1);
        val = (blk & 1) ? cosf(arg) : sinf(arg);
      }
    }
    out[g] = val;
  }
}

__global__ void __launch_bounds__(256)
encode_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ t,
              int64_t R, int S, int L_xyz, int L_dir, float* __restrict__ xyz, int ld_xyz,
              float* __restrict__ dirs, int ld_dir) {
  const int ld = ld_xyz + ld_dir;
  const int wx = 3 * (1 + 2 * L_xyz), wd = 3 * (1 + 2 * L_dir);
  const int64_t total = R * S * ld;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = g / ld;
    const int col = (int)(g - row * ld);
    const int64_t ray = row / S;
    if (col < ld_xyz) {
      float val = 0.f;
      if (col < wx) {
        const float tt = t[row];
        float p[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)   // utils.py:193-194: o + (d * t), product rounded first
          p[c] = __fadd_rn(o[ray * 3 + c], __fmul_rn(d[ray * 3 + c], tt));
        val = pe_value(p, 3, col);
      }
      xyz[row * ld_xyz + col] = val;
    } else {
      const int cd = col - ld_xyz;
      float val = 0.f;
      if (cd < wd) {
        const float v[3] = {d[ray * 3], d[ray * 3 + 1], d[ray * 3 + 2]};   // utils.py:203-207
        val = pe_value(v, 3, cd);
      }
      dirs[row * ld_dir + cd] = val;
    }
  }
}

}  // namespace knerf

using namespace knerf;

extern "C" int knerf_positional_encoding(const float* x, int64_t n_rows, int dim, int L, float* out,
                                         int ld_out, void* stream) {
  KN_CHECK_ARG(x && out && n_rows >= 0 && dim > 0 && L >= 0, "knerf_positional_encoding: bad arguments");
  KN_CHECK_ARG(ld_out >= dim * (1 + 2 * L), "knerf_positional_encoding: ld_out %d < %d", ld_out, dim * (1 + 2 * L));
  if (n_rows == 0) return KNERF_OK;
  const int grid = (int)std::min<int64_t>(cdiv(n_rows * ld_out, 256), (int64_t)kNumSMs * 16);
  posenc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n_rows, dim, L, out, ld_out);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

extern "C" int knerf_encode_position_and_directions(const float* o, const float* d, const float* t, int64_t R,
                                                    int S, int L_xyz, int L_dir, float* xyz, int ld_xyz,
                                                    float* dirs, int ld_dir, void* stream) {
  KN_CHECK_ARG(o && d && t && xyz && dirs && R >= 0 && S > 0, "knerf_encode_position_and_directions: bad arguments");
  KN_CHECK_ARG(ld_xyz >= 3 * (1 + 2 * L_xyz) && ld_dir >= 3 * (1 + 2 * L_dir), "knerf_encode: leading dims too small");
  if (R == 0) return KNERF_OK;
  const int64_t total = R * S * (int64_t)(ld_xyz + ld_dir);
  const int grid = (int)std::min<int64_t>(cdiv(total, 256), (int64_t)kNumSMs * 16);
  encode_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(o, d, t, R, S, L_xyz, L_dir, xyz, ld_xyz, dirs, ld_dir);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}
