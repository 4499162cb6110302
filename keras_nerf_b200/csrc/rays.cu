// a3: camera-ray generation + stratified coarse samples  (keras_nerf/data/rays.py:69-130)
// HBM-write-bound: 24 + 4*N bytes per ray out (+ 4*N in when the uniforms are supplied).
#include "common.cuh"

namespace knerf {

struct Pose {
  float r[3][3];
  float t[3];
};

// One warp per 32 consecutive rays.  Phase 1: lane l computes origin and direction of ray base + l (every lane busy:
// the two divisions, the normalisation and its three divisions cost the warp one pass per 32 rays instead of one per
// ray).  Phase 2: the warp walks the 32 * N samples of those rays in units of VEC (4 = one 16-byte store), consecutive
// lanes on consecutive addresses of t.  Index arithmetic is 32-bit.
template <int VEC>
__global__ void __launch_bounds__(256) rays_kernel(Pose pose, int H, int W, float focal, float near_,
                                                   float far_, int N, const float* __restrict__ u,
                                                   uint64_t seed, float* __restrict__ o,
                                                   float* __restrict__ d, float* __restrict__ t) {
  const uint32_t per_ray = (uint32_t)(N / VEC);
  const uint32_t n_rays = (uint32_t)H * (uint32_t)W;
  const float Wf = (float)W, Hf = (float)H, Nf = (float)N;
  const float delta = __fdiv_rn(__fsub_rn(far_, near_), (float)(N - 1));   // tf.linspace step
  const float interval = __fdiv_rn(__fsub_rn(far_, near_), Nf);            // rays.py:120 (N, not N-1)
  const float half_iv = interval * 0.5f;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t base = warp * 32; base < n_rays; base += n_warps * 32) {
    {
      const uint32_t ray = base + lane;
      if (ray < n_rays) {
        const uint32_t y = ray / (uint32_t)W, x = ray - y * (uint32_t)W;
        // rays.py:89-94 -- no pixel-centre offset; cam = (xc, -yc, -1)
        const float xc = __fdiv_rn(__fsub_rn((float)x, Wf * 0.5f), focal);
        const float yc = __fdiv_rn(__fsub_rn((float)y, Hf * 0.5f), focal);
        float dir[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          // rays.py:103-107: broadcast multiply then reduce_sum over j = 0,1,2 (products rounded separately)
          const float p0 = __fmul_rn(xc, pose.r[i][0]);
          const float p1 = __fmul_rn(-yc, pose.r[i][1]);
          const float p2 = __fmul_rn(-1.0f, pose.r[i][2]);
          dir[i] = __fadd_rn(__fadd_rn(p0, p1), p2);
        }
        const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dir[0], dir[0]), __fmul_rn(dir[1], dir[1])),
                                               __fmul_rn(dir[2], dir[2])));
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          d[(size_t)ray * 3 + i] = __fdiv_rn(dir[i], nrm);   // rays.py:108-109
          o[(size_t)ray * 3 + i] = pose.t[i];                // rays.py:112-113
        }
      }
    }
    const uint32_t units = min(32u, n_rays - base) * per_ray;
    for (uint32_t unit = lane; unit < units; unit += 32) {
    const uint32_t r = unit / per_ray;
    const int s0 = (int)(unit - r * per_ray) * VEC;
    const int64_t ray = (int64_t)base + r;
    float uu[VEC];
    const int64_t e0 = ray * N + s0;
    if (u != nullptr) {
      if (VEC == 4) {
        const float4 v = ld_stream4(reinterpret_cast<const float4*>(u + e0));
        uu[0] = v.x; uu[1 % VEC] = v.y; uu[2 % VEC] = v.z; uu[3 % VEC] = v.w;
      } else {
        uu[0] = ld_stream(u + e0);
      }
    } else {
      if (VEC == 4) {
        uint32_t r[4];
        philox4x32(seed, (uint64_t)e0 >> 2, r);   // e0 % 4 == 0
#pragma unroll
        for (int k = 0; k < VEC; ++k) uu[k] = u01(r[k]);
      } else {
        uu[0] = philox_uniform(seed, (uint64_t)e0);
      }
    }
    float out[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int i = s0 + k;
      // tf.linspace: exact end points, start + delta*i in between (rays.py:116-117)
      float lin = __fadd_rn(near_, __fmul_rn(delta, (float)i));
      if (i == N - 1) lin = far_;
      if (i == 0) lin = near_;
      const float noise = __fsub_rn(__fmul_rn(uu[k], interval), half_iv);   // rays.py:122-123
      out[k] = fminf(fmaxf(__fadd_rn(lin, noise), near_), far_);            // rays.py:126-127
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(t + e0) = make_float4(out[0], out[1 % VEC], out[2 % VEC], out[3 % VEC]);
    } else {
      t[e0] = out[0];
    }
    }
  }
}

__global__ void __launch_bounds__(256) uniform_kernel(float* __restrict__ out, int64_t n, uint64_t seed,
                                                      uint64_t offset) {
  const int64_t n4 = (n + 3) / 4;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n4;
       g += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32(seed, offset + (uint64_t)g, r);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t e = g * 4 + k;
      if (e < n) out[e] = u01(r[k]);
    }
  }
}

}  // namespace knerf

using namespace knerf;

extern "C" int knerf_generate_rays(const float* c2w_host, int H, int W, float focal, float near_,
                                   float far_, int n_samples, const float* u, uint64_t seed, float* o,
                                   float* d, float* t, void* stream) {
  KN_CHECK_ARG(c2w_host && o && d && t, "knerf_generate_rays: null pointer");
  KN_CHECK_ARG(H > 0 && W > 0 && n_samples > 0 && focal > 0.f, "knerf_generate_rays: bad shape H=%d W=%d N=%d focal=%g",
               H, W, n_samples, (double)focal);
  Pose p;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) p.r[i][j] = c2w_host[i * 4 + j];   // rays.py:99
    p.t[i] = c2w_host[i * 4 + 3];                                  // rays.py:100
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (n_samples % 4 == 0) && ((reinterpret_cast<uintptr_t>(t) & 15) == 0) &&
                   (u == nullptr || (reinterpret_cast<uintptr_t>(u) & 15) == 0);
  KN_CHECK_ARG((int64_t)H * W < (int64_t)1 << 31 && (int64_t)n_samples * 32 < (int64_t)1 << 31,
               "knerf_generate_rays: image too large (%d x %d)", H, W);
  const int grid = (int)std::min<int64_t>(cdiv((int64_t)H * W, 256), (int64_t)kNumSMs * 16);
  if (vec)
    rays_kernel<4><<<grid, 256, 0, st>>>(p, H, W, focal, near_, far_, n_samples, u, seed, o, d, t);
  else
    rays_kernel<1><<<grid, 256, 0, st>>>(p, H, W, focal, near_, far_, n_samples, u, seed, o, d, t);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

extern "C" int knerf_uniform(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
  KN_CHECK_ARG(out && n >= 0, "knerf_uniform: bad arguments");
  if (n == 0) return KNERF_OK;
  const int grid = (int)std::min<int64_t>(cdiv(cdiv(n, 4), 256), (int64_t)kNumSMs * 16);
  uniform_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, n, seed, offset);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}
