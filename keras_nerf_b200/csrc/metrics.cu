// f1 / a18: the per-image metrics of NeRF.update_and_return_metrics (keras_nerf/model/nerf/nerf.py:306-330):
// tf.image.psnr(max_val=1) needs the per-image mean squared error, tf.image.ssim(max_val=1) the mean over the
// VALID positions of an 11x11 gaussian window (sigma 1.5) of luminance * contrast-structure, per channel.
//
// One block per 16x16 tile of window positions of one (image, channel): the 26x26 input patch of both images is
// staged in shared memory, the four window means TF forms (x, y, x*y, x^2 + y^2) are taken separably (11 taps
// along rows into shared memory, 11 along columns), the per-position SSIM and the squared differences of the
// pixels the tile owns are block-reduced to one partial each.  A second one-block-per-image kernel adds the
// partials in a fixed order, so the result is bit-reproducible.  HBM: both images read once (+ 10-pixel halos,
// L2 hits).
#include "common.cuh"

namespace knerf {

constexpr int kMT = 16;                  // window positions per tile edge
constexpr int kMW = 11;                  // window size (tf.image.ssim default)
constexpr int kMP = kMT + kMW - 1;       // 26: patch edge

struct GaussWindow { float g[kMW]; };

__global__ void __launch_bounds__(kMT * kMT) metrics_tile_kernel(const float* __restrict__ a,
                                                                 const float* __restrict__ b, int H, int W, int C,
                                                                 GaussWindow gw, float c1, float c2,
                                                                 float* __restrict__ part_ssim,
                                                                 float* __restrict__ part_sq) {
  __shared__ float sx[kMP][kMP + 1], sy[kMP][kMP + 1];
  __shared__ float rows[4][kMP][kMT + 1];
  __shared__ float red[2][kMT * kMT / 32];
  const int tid = threadIdx.x;
  const int tiles_x = gridDim.x, tiles_y = gridDim.y;
  const int bc = blockIdx.z, img = bc / C, ch = bc % C;
  const int y0 = blockIdx.y * kMT, x0 = blockIdx.x * kMT;
  const int Ho = H - kMW + 1, Wo = W - kMW + 1;
  const float* pa = a + (size_t)img * H * W * C + ch;
  const float* pb = b + (size_t)img * H * W * C + ch;
  // rows / columns of squared differences this tile accounts for (the last tile of a row / column takes the rest)
  const int own_y1 = (blockIdx.y == tiles_y - 1) ? H : y0 + kMT;
  const int own_x1 = (blockIdx.x == tiles_x - 1) ? W : x0 + kMT;
  float sq = 0.f;
  for (int i = tid; i < kMP * kMP; i += kMT * kMT) {
    const int py = i / kMP, px = i % kMP;
    const int y = y0 + py, x = x0 + px;
    float va = 0.f, vb = 0.f;
    if (y < H && x < W) {
      va = __ldg(pa + ((size_t)y * W + x) * C);
      vb = __ldg(pb + ((size_t)y * W + x) * C);
      if (y < own_y1 && x < own_x1) { const float d = va - vb; sq = fmaf(d, d, sq); }
    }
    sx[py][px] = va;
    sy[py][px] = vb;
  }
  __syncthreads();
  for (int i = tid; i < kMP * kMT; i += kMT * kMT) {      // 11 taps along x for every patch row
    const int py = i / kMT, ox = i % kMT;
    float mx = 0.f, my = 0.f, mxy = 0.f, mss = 0.f;
#pragma unroll
    for (int k = 0; k < kMW; ++k) {
      const float vx = sx[py][ox + k], vy = sy[py][ox + k], w = gw.g[k];
      mx = fmaf(w, vx, mx);
      my = fmaf(w, vy, my);
      mxy = fmaf(w, vx * vy, mxy);
      mss = fmaf(w, fmaf(vx, vx, vy * vy), mss);
    }
    rows[0][py][ox] = mx; rows[1][py][ox] = my; rows[2][py][ox] = mxy; rows[3][py][ox] = mss;
  }
  __syncthreads();
  const int oy = tid / kMT, ox = tid % kMT;
  float s = 0.f;
  if (y0 + oy < Ho && x0 + ox < Wo) {
    float mx = 0.f, my = 0.f, mxy = 0.f, mss = 0.f;
#pragma unroll
    for (int k = 0; k < kMW; ++k) {
      const float w = gw.g[k];
      mx = fmaf(w, rows[0][oy + k][ox], mx);
      my = fmaf(w, rows[1][oy + k][ox], my);
      mxy = fmaf(w, rows[2][oy + k][ox], mxy);
      mss = fmaf(w, rows[3][oy + k][ox], mss);
    }
    // image_ops_impl.py _ssim_helper: luminance = (2 mu_x mu_y + c1) / (mu_x^2 + mu_y^2 + c1),
    // cs = (2 E[xy] - 2 mu_x mu_y + c2) / (E[x^2 + y^2] - mu_x^2 - mu_y^2 + c2)
    const float num0 = mx * my * 2.0f, den0 = mx * mx + my * my;
    const float lum = (num0 + c1) / (den0 + c1);
    const float cs = (mxy * 2.0f - num0 + c2) / (mss - den0 + c2);
    s = lum * cs;
  }
  s = warp_sum(s);
  sq = warp_sum(sq);
  if ((tid & 31) == 0) { red[0][tid >> 5] = s; red[1][tid >> 5] = sq; }
  __syncthreads();
  if (tid == 0) {
    float ts = 0.f, tq = 0.f;
    for (int i = 0; i < kMT * kMT / 32; ++i) { ts += red[0][i]; tq += red[1][i]; }
    const size_t slot = ((size_t)bc * tiles_y + blockIdx.y) * tiles_x + blockIdx.x;
    part_ssim[slot] = ts;
    part_sq[slot] = tq;
  }
}

// one block per image: fixed-order sum of its C * tiles partials
__global__ void __launch_bounds__(256) metrics_finish_kernel(const float* __restrict__ part_ssim,
                                                             const float* __restrict__ part_sq, int n_per_image,
                                                             float inv_ssim_count, float inv_sq_count,
                                                             float* __restrict__ mse, float* __restrict__ ssim) {
  __shared__ float red[2][256];
  const int img = blockIdx.x, tid = threadIdx.x;
  float s = 0.f, q = 0.f;
  for (int i = tid; i < n_per_image; i += 256) {
    s += part_ssim[(size_t)img * n_per_image + i];
    q += part_sq[(size_t)img * n_per_image + i];
  }
  red[0][tid] = s; red[1][tid] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { red[0][tid] += red[0][tid + o]; red[1][tid] += red[1][tid + o]; }
    __syncthreads();
  }
  if (tid == 0) {
    if (mse != nullptr) mse[img] = red[1][0] * inv_sq_count;
    if (ssim != nullptr) ssim[img] = red[0][0] * inv_ssim_count;
  }
}

}  // namespace knerf

using namespace knerf;

extern "C" int64_t knerf_image_metrics_workspace_floats(int B, int H, int W, int C) {
  if (B <= 0 || H < kMW || W < kMW || C <= 0) return 0;
  return 2 * (int64_t)B * C * cdiv(H - kMW + 1, kMT) * cdiv(W - kMW + 1, kMT);
}

extern "C" int knerf_image_metrics(const float* a, const float* b, int B, int H, int W, int C, float max_val,
                                   float* mse, float* ssim, float* workspace, int64_t workspace_floats,
                                   void* stream) {
  KN_CHECK_ARG(a && b && workspace && (mse || ssim), "knerf_image_metrics: null pointer");
  KN_CHECK_ARG(B > 0 && C > 0 && H >= kMW && W >= kMW,
               "knerf_image_metrics: images must be at least %dx%d (got B=%d H=%d W=%d C=%d)", kMW, kMW, B, H, W, C);
  const int ty = (int)cdiv(H - kMW + 1, kMT), tx = (int)cdiv(W - kMW + 1, kMT);
  const int64_t need = knerf_image_metrics_workspace_floats(B, H, W, C);
  KN_CHECK_ARG(workspace_floats >= need, "knerf_image_metrics: workspace of %lld floats needed", (long long)need);
  KN_CHECK_ARG((int64_t)B * C <= 65535 && ty <= 65535, "knerf_image_metrics: too many images / rows");
  GaussWindow gw;                                   // _fspecial_gauss(11, 1.5): softmax of -(i - 5)^2 / (2 sigma^2)
  double sum = 0.0, e[kMW];
  for (int i = 0; i < kMW; ++i) { const double d = i - (kMW - 1) / 2.0; e[i] = std::exp(-0.5 * d * d / (1.5 * 1.5)); sum += e[i]; }
  for (int i = 0; i < kMW; ++i) gw.g[i] = (float)(e[i] / sum);
  const float c1 = (0.01f * max_val) * (0.01f * max_val), c2 = (0.03f * max_val) * (0.03f * max_val);
  float* part_ssim = workspace;
  float* part_sq = workspace + need / 2;
  cudaStream_t st = (cudaStream_t)stream;
  metrics_tile_kernel<<<dim3(tx, ty, B * C), kMT * kMT, 0, st>>>(a, b, H, W, C, gw, c1, c2, part_ssim, part_sq);
  KN_LAUNCH_CHECK();
  metrics_finish_kernel<<<B, 256, 0, st>>>(part_ssim, part_sq, C * ty * tx,
                                           1.0f / ((float)C * (float)(H - kMW + 1) * (float)(W - kMW + 1)),
                                           1.0f / ((float)C * (float)H * (float)W), mse, ssim);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}
