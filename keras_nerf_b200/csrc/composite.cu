// a7: alpha-compositing volume rendering, forward and fused backward
// (keras_nerf/model/nerf/utils.py:16-58; backward = what tf.GradientTape yields for nerf.py:361-377).
//
// One warp per ray.  Sample i of the ray lives in lane (i % 32), register block (i / 32), so every
// global access of the warp is a contiguous 128 B (t, weights) or 512 B (packed rgb-sigma float4)
// segment.  The transmittance cumprod is a warp-shuffle product scan per 32-sample block with a
// carried prefix; the backward's "sum over later samples" is the mirrored suffix scan.
// HBM bytes per ray (algorithmic): forward 24*S + 20, backward 36*S + 24.
#include "common.cuh"

namespace knerf {

constexpr int kWarpsPerBlock = 8;

template <int NB, bool PACKED>
__device__ __forceinline__ void load_ray(const float* __restrict__ rgbsigma, const float* __restrict__ rgb,
                                         const float* __restrict__ sigma, const float* __restrict__ t,
                                         int64_t ray, int S, int lane, float4 (&c)[NB], float (&tt)[NB]) {
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int i = j * 32 + lane;
    if (i < S) {
      const int64_t e = ray * S + i;
      if (PACKED) {
        c[j] = ld_stream4(reinterpret_cast<const float4*>(rgbsigma) + e);
      } else {
        c[j] = make_float4(ld_stream(rgb + e * 3), ld_stream(rgb + e * 3 + 1), ld_stream(rgb + e * 3 + 2),
                           ld_stream(sigma + e));
      }
      tt[j] = ld_stream(t + e);
    } else {
      c[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      tt[j] = 0.f;
    }
  }
}

// alpha, exp(-sigma*delta), e = (1-alpha)+eps, exclusive transmittance T and weight w for every sample
template <int NB>
__device__ __forceinline__ void transmittance(const float4 (&c)[NB], const float (&tt)[NB], int S, int lane,
                                              float eps, float (&delta)[NB], float (&ex)[NB], float (&e)[NB],
                                              float (&T)[NB], float (&w)[NB]) {
  float carry = 1.0f;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int i = j * 32 + lane;
    float tn = __shfl_down_sync(kFullMask, tt[j], 1);
    if (j + 1 < NB) {
      const float first_next = __shfl_sync(kFullMask, tt[(j + 1 < NB) ? j + 1 : j], 0);
      if (lane == 31) tn = first_next;
    }
    // utils.py:35-37: delta_i = t_{i+1}-t_i, last = epsilon (1e-10, not 1e10)
    delta[j] = (i == S - 1) ? eps : __fsub_rn(tn, tt[j]);
    const bool valid = i < S;
    ex[j] = expf(-c[j].w * delta[j]);
    const float alpha = __fsub_rn(1.0f, ex[j]);                      // utils.py:41
    e[j] = valid ? __fadd_rn(__fsub_rn(1.0f, alpha), eps) : 1.0f;    // utils.py:43,47  (1-(1-exp))+eps
    const float incl = warp_scan_mul(e[j], lane);
    float excl = __shfl_up_sync(kFullMask, incl, 1);
    if (lane == 0) excl = 1.0f;
    T[j] = carry * excl;                                             // exclusive cumprod, T_0 = 1
    carry *= __shfl_sync(kFullMask, incl, 31);
    w[j] = valid ? alpha * T[j] : 0.0f;                              // utils.py:48
  }
}

template <int NB, bool PACKED>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd_kernel(const float* __restrict__ rgbsigma, const float* __restrict__ rgb,
                     const float* __restrict__ sigma, const float* __restrict__ t, int64_t R, int S,
                     int white, int clip, float eps, float* __restrict__ image, float* __restrict__ depth,
                     float* __restrict__ weights, float* __restrict__ acc_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  // Short rays (S <= 64) carry only 1.3 KB per warp: the next ray of the warp is fetched before this one is
  // composited, which doubles the bytes in flight per SM (a long ray already fills the memory pipeline by itself).
  constexpr bool kPrefetch = NB <= 2;
  float4 c[NB], c_next[NB];
  float tt[NB], tt_next[NB];
  if (kPrefetch && warp0 < R) load_ray<NB, PACKED>(rgbsigma, rgb, sigma, t, warp0, S, lane, c_next, tt_next);
  for (int64_t ray = warp0; ray < R; ray += nwarps) {
    float delta[NB], ex[NB], e[NB], T[NB], w[NB];
    if (kPrefetch) {
#pragma unroll
      for (int j = 0; j < NB; ++j) { c[j] = c_next[j]; tt[j] = tt_next[j]; }
      if (ray + nwarps < R) load_ray<NB, PACKED>(rgbsigma, rgb, sigma, t, ray + nwarps, S, lane, c_next, tt_next);
    } else {
      load_ray<NB, PACKED>(rgbsigma, rgb, sigma, t, ray, S, lane, c, tt);
    }
    transmittance<NB>(c, tt, S, lane, eps, delta, ex, e, T, w);
    float cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, acc = 0.f;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      cr += w[j] * c[j].x; cg += w[j] * c[j].y; cb += w[j] * c[j].z;   // utils.py:50
      dep += w[j] * tt[j];                                             // utils.py:51
      acc += w[j];
      const int i = j * 32 + lane;
      if (weights != nullptr && i < S) weights[ray * S + i] = w[j];
    }
    cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); dep = warp_sum(dep); acc = warp_sum(acc);
    if (lane == 0) {
      if (white) { const float bg = 1.0f - acc; cr += bg; cg += bg; cb += bg; }   // utils.py:53-54
      if (clip) {                                                                  // utils.py:56
        cr = fminf(fmaxf(cr, 0.f), 1.f); cg = fminf(fmaxf(cg, 0.f), 1.f); cb = fminf(fmaxf(cb, 0.f), 1.f);
      }
      if (image != nullptr) { image[ray * 3] = cr; image[ray * 3 + 1] = cg; image[ray * 3 + 2] = cb; }
      if (depth != nullptr) depth[ray] = dep;
      if (acc_out != nullptr) acc_out[ray] = acc;
    }
  }
}

// Backward.  With G_c = dL/dC_c * 1[0 <= C_c(pre-clip) <= 1], bg = white ? 1 : 0,
//   g_i = sum_c G_c (rgb_ic - bg)
//   dL/drgb_ic   = G_c w_i
//   dL/dalpha_i  = g_i T_i - (sum_{k>i} g_k w_k) / e_i
//   dL/dsigma_i  = dL/dalpha_i * delta_i * exp(-sigma_i delta_i)
// through_act folds rgb=sigmoid(.), sigma=relu(.) derivatives in.
template <int NB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, (NB >= 3 && NB <= 6) ? 4 : 0)
composite_bwd_kernel(const float* __restrict__ rgbsigma, const float* __restrict__ t, int64_t R, int S,
                     int white, int clip, float eps, const float* __restrict__ dimage,
                     const float* __restrict__ target, float loss_scale, int through_act,
                     float* __restrict__ d_out, float* __restrict__ sqerr) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  const float bg = white ? 1.0f : 0.0f;
  constexpr bool kPrefetch = NB <= 2;   // as in the forward kernel (no gain for longer rays: measured)
  float4 c[NB], c_next[NB];
  float tt[NB], tt_next[NB];
  if (kPrefetch && warp0 < R) load_ray<NB, true>(rgbsigma, nullptr, nullptr, t, warp0, S, lane, c_next, tt_next);
  for (int64_t ray = warp0; ray < R; ray += nwarps) {
    float delta[NB], ex[NB], e[NB], T[NB], w[NB];
    if (kPrefetch) {
#pragma unroll
      for (int j = 0; j < NB; ++j) { c[j] = c_next[j]; tt[j] = tt_next[j]; }
      if (ray + nwarps < R) load_ray<NB, true>(rgbsigma, nullptr, nullptr, t, ray + nwarps, S, lane, c_next, tt_next);
    } else {
      load_ray<NB, true>(rgbsigma, nullptr, nullptr, t, ray, S, lane, c, tt);
    }
    transmittance<NB>(c, tt, S, lane, eps, delta, ex, e, T, w);
    float cr = 0.f, cg = 0.f, cb = 0.f, acc = 0.f;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      cr += w[j] * c[j].x; cg += w[j] * c[j].y; cb += w[j] * c[j].z; acc += w[j];
    }
    cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); acc = warp_sum(acc);
    if (white) { const float b = 1.0f - acc; cr += b; cg += b; cb += b; }
    float G[3];
    const float pre[3] = {cr, cg, cb};
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float clipped = clip ? fminf(fmaxf(pre[k], 0.f), 1.f) : pre[k];
      float gk;
      if (target != nullptr) {
        const float diff = clipped - target[ray * 3 + k];
        se += diff * diff;
        gk = loss_scale * diff;
      } else {
        gk = dimage[ray * 3 + k];
      }
      // tf.clip_by_value passes the gradient on [lo, hi] inclusive, zero outside [TF-sem]
      if (clip && (pre[k] < 0.f || pre[k] > 1.f)) gk = 0.f;
      G[k] = gk;
    }
    if (sqerr != nullptr && lane == 0) sqerr[ray] = se;

    float g[NB], gw[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      g[j] = G[0] * (c[j].x - bg) + G[1] * (c[j].y - bg) + G[2] * (c[j].z - bg);
      gw[j] = g[j] * w[j];   // 0 for padding lanes (w = 0)
    }
    // exclusive suffix sums, blocks from last to first
    float carry = 0.f;
#pragma unroll
    for (int j = NB - 1; j >= 0; --j) {
      float incl = gw[j];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_down_sync(kFullMask, incl, o);
        if (lane + o < 32) incl += n;
      }
      float excl = __shfl_down_sync(kFullMask, incl, 1);
      if (lane == 31) excl = 0.f;
      const float suffix = excl + carry;
      carry += __shfl_sync(kFullMask, incl, 0);
      const int i = j * 32 + lane;
      if (i < S) {
        const float d_alpha = g[j] * T[j] - suffix / e[j];
        float d_sigma = d_alpha * delta[j] * ex[j];
        float dr = G[0] * w[j], dg = G[1] * w[j], db = G[2] * w[j];
        if (through_act) {
          dr *= c[j].x * (1.0f - c[j].x);          // sigmoid'
          dg *= c[j].y * (1.0f - c[j].y);
          db *= c[j].z * (1.0f - c[j].z);
          if (!(c[j].w > 0.0f)) d_sigma = 0.0f;    // relu' (0 at 0, as tf ReluGrad)
        }
        reinterpret_cast<float4*>(d_out)[ray * S + i] = make_float4(dr, dg, db, d_sigma);
      }
    }
  }
}

static int pick_nb(int S) {
  const int nb = (S + 31) / 32;
  const int opts[] = {1, 2, 3, 4, 6, 8, 10, 12, 16};
  for (int o : opts) if (nb <= o) return o;
  return -1;
}

}  // namespace knerf

using namespace knerf;

#define KN_DISPATCH_NB(nb, ...)                 \
  switch (nb) {                                 \
    case 1: { constexpr int NB = 1; __VA_ARGS__; } break;   \
    case 2: { constexpr int NB = 2; __VA_ARGS__; } break;   \
    case 3: { constexpr int NB = 3; __VA_ARGS__; } break;   \
    case 4: { constexpr int NB = 4; __VA_ARGS__; } break;   \
    case 6: { constexpr int NB = 6; __VA_ARGS__; } break;   \
    case 8: { constexpr int NB = 8; __VA_ARGS__; } break;   \
    case 10: { constexpr int NB = 10; __VA_ARGS__; } break; \
    case 12: { constexpr int NB = 12; __VA_ARGS__; } break; \
    default: { constexpr int NB = 16; __VA_ARGS__; } break; \
  }

extern "C" int knerf_composite_forward(const float* rgbsigma, const float* rgb, const float* sigma,
                                       const float* t, int64_t R, int S, int white_background, int clip,
                                       float epsilon, float* image, float* depth, float* weights,
                                       float* acc, void* stream) {
  KN_CHECK_ARG(t != nullptr && R >= 0 && S > 0, "knerf_composite_forward: bad arguments");
  KN_CHECK_ARG((rgbsigma != nullptr) != (rgb != nullptr && sigma != nullptr),
               "knerf_composite_forward: pass either rgbsigma or (rgb, sigma)");
  const int nb = pick_nb(S);
  KN_CHECK_ARG(nb > 0, "knerf_composite_forward: S=%d exceeds 512 samples per ray", S);
  if (R == 0) return KNERF_OK;
  const int grid = (int)std::min<int64_t>(cdiv(R, kWarpsPerBlock), (int64_t)kNumSMs * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (rgbsigma != nullptr) {
    KN_CHECK_ARG((reinterpret_cast<uintptr_t>(rgbsigma) & 15) == 0, "rgbsigma must be 16-byte aligned");
    KN_DISPATCH_NB(nb, (composite_fwd_kernel<NB, true><<<grid, kWarpsPerBlock * 32, 0, st>>>(
                           rgbsigma, nullptr, nullptr, t, R, S, white_background, clip, epsilon, image, depth,
                           weights, acc)));
  } else {
    KN_DISPATCH_NB(nb, (composite_fwd_kernel<NB, false><<<grid, kWarpsPerBlock * 32, 0, st>>>(
                           nullptr, rgb, sigma, t, R, S, white_background, clip, epsilon, image, depth,
                           weights, acc)));
  }
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

extern "C" int knerf_composite_backward(const float* rgbsigma, const float* t, int64_t R, int S,
                                        int white_background, int clip, float epsilon, const float* dimage,
                                        const float* target, float loss_scale, int through_activations,
                                        float* d_out, float* sqerr, void* stream) {
  KN_CHECK_ARG(rgbsigma && t && d_out && R >= 0 && S > 0, "knerf_composite_backward: bad arguments");
  KN_CHECK_ARG((dimage != nullptr) != (target != nullptr),
               "knerf_composite_backward: pass either dimage or target");
  KN_CHECK_ARG(((reinterpret_cast<uintptr_t>(rgbsigma) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0,
               "rgbsigma / d_out must be 16-byte aligned");
  const int nb = pick_nb(S);
  KN_CHECK_ARG(nb > 0, "knerf_composite_backward: S=%d exceeds 512 samples per ray", S);
  if (R == 0) return KNERF_OK;
  const int grid = (int)std::min<int64_t>(cdiv(R, kWarpsPerBlock), (int64_t)kNumSMs * 8);
  cudaStream_t st = (cudaStream_t)stream;
  KN_DISPATCH_NB(nb, (composite_bwd_kernel<NB><<<grid, kWarpsPerBlock * 32, 0, st>>>(
                         rgbsigma, t, R, S, white_background, clip, epsilon, dimage, target, loss_scale,
                         through_activations, d_out, sqerr)));
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}
