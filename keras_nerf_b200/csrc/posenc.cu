// a4 / a5: positional encoding and ray-point construction
// (keras_nerf/model/nerf/utils.py:176-186, :188-210).  fp32 parity path only: the tcgen05 MLP fuses
// this into its first-layer prologue and never writes encodings to HBM.
#include "common.cuh"

namespace knerf {

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

// One thread per (sample, unit): unit 0 copies the vector itself, unit 1 + f is frequency f -- ONE sincosf per
// component gives the unit's three sines and three cosines (same argument reduction and polynomials as sinf / cosf:
// the values are those of posenc_kernel's).  The direction encoding is the same for every sample of a ray but is written
// per sample like the reference does (utils.py:203-207 broadcasts before encoding).
// (One thread per output element -- a separate sinf or cosf each and two 64-bit divisions per element -- took 0.62 ms
// for 786 k samples; this form 0.2 ms.)
__global__ void __launch_bounds__(256)
encode_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ t,
              int64_t R, int S, int L_xyz, int L_dir, float* __restrict__ xyz, int ld_xyz,
              float* __restrict__ dirs, int ld_dir) {
  const int ux = 1 + L_xyz, ud = 1 + L_dir, units = ux + ud;
  const int wx = 3 * (1 + 2 * L_xyz), wd = 3 * (1 + 2 * L_dir);
  const int64_t total = R * S * units;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = g / units;
    const int u = (int)(g - row * units);
    const int64_t ray = row / S;
    const bool is_dir = u >= ux;
    float v[3];
    if (!is_dir) {
      const float tt = t[row];
#pragma unroll
      for (int c = 0; c < 3; ++c)   // utils.py:193-194: o + (d * t), product rounded first
        v[c] = __fadd_rn(o[ray * 3 + c], __fmul_rn(d[ray * 3 + c], tt));
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = d[ray * 3 + c];
    }
    const int f = (is_dir ? u - ux : u) - 1;                       // -1: the identity block
    float* out = is_dir ? dirs + row * ld_dir : xyz + row * ld_xyz;
    if (f < 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) out[c] = v[c];
      // zero the padding columns of the row (ld > width)
      const int w = is_dir ? wd : wx, ld = is_dir ? ld_dir : ld_xyz;
      for (int c = w; c < ld; ++c) out[c] = 0.f;
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float sn, cs;
        sincosf(ldexpf(v[c], f), &sn, &cs);                        // 2.0**f * x, exact
        out[3 + 6 * f + c] = sn;
        out[6 + 6 * f + c] = cs;
      }
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
  const int64_t total = R * S * (int64_t)(2 + L_xyz + L_dir);
  const int grid = (int)std::min<int64_t>(cdiv(total, 256), (int64_t)kNumSMs * 16);
  encode_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(o, d, t, R, S, L_xyz, L_dir, xyz, ld_xyz, dirs, ld_dir);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}
