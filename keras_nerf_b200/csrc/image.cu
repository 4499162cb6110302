// f3: ImageLoader.__call__ after the PNG decode (keras_nerf/data/image.py:17-35):
//   convert_image_dtype(uint8 -> float32)  ->  tf.image.resize(bilinear, antialias=True)
//   ->  alpha-composite on a white / black background  ->  concat alpha  ->  clip [0,1].
//
// tf.image.resize(antialias=True) is the ScaleAndTranslate op with the triangle kernel: per axis an output
// sample x looks at the input positions within `radius * kernel_scale` of (x + 0.5) / scale, weights
// max(0, 1 - |pos| / kernel_scale) normalised to sum 1, rows resized first, columns second, every sum taken in
// span order in fp32.  One thread per output pixel does exactly that for its 4 channels: for each column tap the
// row sum (the value TF's intermediate [out_h, in_w] image holds), then the column sum.  The weights are
// computed in-kernel (registers for row spans of up to 8 taps, per tap beyond) instead of being tabulated;
// products and sums are left unfused (__fmul_rn / __fadd_rn) so the result equals the two-pass fp32
// formulation bit for bit.
// HBM: reads the uint8 source once (4 B per source pixel; the taps of neighbouring outputs overlap in L1/L2),
// writes 16 B per output pixel.
#include "common.cuh"

namespace knerf {

struct ResizeAxis {
  float inv_scale;         // float(1.0 / double(scale)), scale = float(out) / float(in)
  float support;           // radius (1) * kernel_scale, kernel_scale = max(inv_scale, 1)
  float inv_kernel_scale;  // 1.0f / kernel_scale
  int in_size, out_size;
};

struct Span {
  int start, n;
  float sample, inv_total;   // inv_total == 0: all-zero weights (TF leaves the weight row at 0)
};

__device__ __forceinline__ float tap_weight(const ResizeAxis& ax, const Span& sp, int k) {
  const float pos = __fsub_rn(__fadd_rn((float)(sp.start + k), 0.5f), sp.sample);
  const float x = fabsf(__fmul_rn(pos, ax.inv_kernel_scale));
  const float w = x < 1.0f ? __fsub_rn(1.0f, x) : 0.0f;
  return __fmul_rn(w, sp.inv_total);
}

__device__ __forceinline__ Span make_span(const ResizeAxis& ax, int x) {
  Span sp;
  sp.sample = __fmul_rn(__fadd_rn((float)x, 0.5f), ax.inv_scale);   // + inv_translate (= -0)
  sp.start = 0; sp.n = 0; sp.inv_total = 0.f;
  if (sp.sample < 0.f || sp.sample > (float)ax.in_size) return sp;
  int s0 = (int)ceilf(__fsub_rn(__fsub_rn(sp.sample, ax.support), 0.5f));
  int s1 = (int)floorf(__fsub_rn(__fadd_rn(sp.sample, ax.support), 0.5f));
  s0 = min(max(s0, 0), ax.in_size - 1);
  s1 = min(max(s1, 0), ax.in_size - 1) + 1;
  sp.start = s0; sp.n = s1 - s0;
  sp.inv_total = 1.0f;
  float total = 0.f;
  for (int k = 0; k < sp.n; ++k) total = __fadd_rn(total, tap_weight(ax, sp, k));
  sp.inv_total = (fabsf(total) >= 1000.0f * 1.17549435e-38f) ? __fdiv_rn(1.0f, total) : 0.f;
  return sp;
}

__global__ void __launch_bounds__(256) image_prepare_kernel(const uchar4* __restrict__ src, ResizeAxis ay,
                                                            ResizeAxis ax, int white, float4* __restrict__ out) {
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31);
  const int oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= ax.out_size || oy >= ay.out_size) return;
  const Span sy = make_span(ay, oy), sx = make_span(ax, ox);
  const float k255 = 1.0f / 255.0f;   // convert_image_dtype: cast * (1 / 255)
  // row weights do not depend on the column tap: spans of up to kRegTaps rows (reductions up to 3.5x) keep them
  // in registers, longer ones recompute them per tap
  constexpr int kRegTaps = 8;
  float wy[kRegTaps];
#pragma unroll
  for (int k = 0; k < kRegTaps; ++k) wy[k] = (k < sy.n) ? tap_weight(ay, sy, k) : 0.f;
  float r = 0.f, g = 0.f, b = 0.f, a = 0.f;
  for (int kx = 0; kx < sx.n; ++kx) {
    float ir = 0.f, ig = 0.f, ib = 0.f, ia = 0.f;   // the [oy, sx.start + kx] pixel of TF's row-resized image
    const uchar4* col = src + (size_t)sy.start * ax.in_size + (sx.start + kx);
    if (sy.n <= kRegTaps) {
#pragma unroll
      for (int ky = 0; ky < kRegTaps; ++ky) {
        if (ky < sy.n) {
          const uchar4 p = __ldg(col + (size_t)ky * ax.in_size);
          ir = __fadd_rn(ir, __fmul_rn(wy[ky], __fmul_rn((float)p.x, k255)));
          ig = __fadd_rn(ig, __fmul_rn(wy[ky], __fmul_rn((float)p.y, k255)));
          ib = __fadd_rn(ib, __fmul_rn(wy[ky], __fmul_rn((float)p.z, k255)));
          ia = __fadd_rn(ia, __fmul_rn(wy[ky], __fmul_rn((float)p.w, k255)));
        }
      }
    } else {
      for (int ky = 0; ky < sy.n; ++ky) {
        const uchar4 p = __ldg(col + (size_t)ky * ax.in_size);
        const float w = tap_weight(ay, sy, ky);
        ir = __fadd_rn(ir, __fmul_rn(w, __fmul_rn((float)p.x, k255)));
        ig = __fadd_rn(ig, __fmul_rn(w, __fmul_rn((float)p.y, k255)));
        ib = __fadd_rn(ib, __fmul_rn(w, __fmul_rn((float)p.z, k255)));
        ia = __fadd_rn(ia, __fmul_rn(w, __fmul_rn((float)p.w, k255)));
      }
    }
    const float w = tap_weight(ax, sx, kx);
    r = __fadd_rn(r, __fmul_rn(w, ir));
    g = __fadd_rn(g, __fmul_rn(w, ig));
    b = __fadd_rn(b, __fmul_rn(w, ib));
    a = __fadd_rn(a, __fmul_rn(w, ia));
  }
  // image.py:25-33: alpha * rgb + (1 - alpha) * background, concat alpha, clip
  const float bg = white ? 1.0f : 0.0f;
  const float na = __fmul_rn(__fsub_rn(1.0f, a), bg);
  float4 o;
  o.x = fminf(fmaxf(__fadd_rn(__fmul_rn(a, r), na), 0.f), 1.f);
  o.y = fminf(fmaxf(__fadd_rn(__fmul_rn(a, g), na), 0.f), 1.f);
  o.z = fminf(fmaxf(__fadd_rn(__fmul_rn(a, b), na), 0.f), 1.f);
  o.w = fminf(fmaxf(a, 0.f), 1.f);
  out[(size_t)oy * ax.out_size + ox] = o;
}

static ResizeAxis make_axis(int in_size, int out_size) {
  ResizeAxis ax;
  const float scale = (float)out_size / (float)in_size;   // tf.image.resize: cast(size) / cast(shape), fp32
  ax.inv_scale = (float)(1.0 / (double)scale);
  const float kernel_scale = ax.inv_scale > 1.0f ? ax.inv_scale : 1.0f;
  ax.support = 1.0f * kernel_scale;
  ax.inv_kernel_scale = 1.0f / kernel_scale;
  ax.in_size = in_size;
  ax.out_size = out_size;
  return ax;
}

}  // namespace knerf

using namespace knerf;

extern "C" int knerf_image_prepare(const uint8_t* rgba, int in_h, int in_w, int out_h, int out_w,
                                   int white_background, float* out, void* stream) {
  KN_CHECK_ARG(rgba != nullptr && out != nullptr, "knerf_image_prepare: null pointer");
  KN_CHECK_ARG(in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "knerf_image_prepare: bad shape %dx%d -> %dx%d",
               in_h, in_w, out_h, out_w);
  KN_CHECK_ARG((reinterpret_cast<uintptr_t>(rgba) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "knerf_image_prepare: rgba must be 4-byte and out 16-byte aligned");
  const dim3 grid((unsigned)cdiv(out_w, 32), (unsigned)cdiv(out_h, 8));
  image_prepare_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uchar4*>(rgba),
                                                               make_axis(in_h, out_h), make_axis(in_w, out_w),
                                                               white_background ? 1 : 0,
                                                               reinterpret_cast<float4*>(out));
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}
