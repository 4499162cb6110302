// a6 (KNERF_BF16 mode): the NeRF MLP on Blackwell tensor cores -- tcgen05.mma with TMEM accumulators, weights
// streamed by TMA bulk copies, positional encoding fused into the first-layer prologue and the sigma / rgb
// heads into the epilogue, so per-sample activations never round-trip through HBM at inference
// (keras_nerf/model/nerf/mlp.py:29-50 + keras_nerf/model/nerf/utils.py:176-210).
//
// Flagship shape only (8x256, skip 4, L = 10/4); other shapes run the fp32 SIMT path.
//
// Forward kernel (persistent, 1 CTA / SM, 320 threads):
//   warp 0      TMA producer: streams the pre-packed bf16 weight blob, 32-wide K stages, 4-slot mbarrier ring
//   warp 1      MMA issuer:   one thread issues tcgen05.mma (M=128, N=256|128, K=16) and tcgen05.commit
//   warps 2-9   compute:      PE prologue, TMEM -> register epilogues (bias, ReLU, bf16 pack) that write the
//                             next layer's A operand straight into shared memory, heads, output
// Each CTA keeps TWO 128-sample tiles in flight (accumulators in TMEM columns [0,256) and [256,512)): while
// the tensor core runs tile 1's layer, the compute warps drain tile 0's accumulator and build its next A
// operand, and vice versa, so the MMA pipe only waits when an epilogue is slower than a layer of MMAs.
// Shared memory: 2 x 64 KB hidden activations + 2 x 16 KB encodings + 4 x 16 KB weight stages = 224 KB.
#include "mlp_tc.cuh"

#include "tc_ptx.cuh"

namespace knerf {
using namespace tc;

namespace {

constexpr int kTileM = 128;
constexpr int kU = 256;
constexpr int kKStage = 32;                       // K elements per weight stage
constexpr int kStageBytes = kU * kKStage * 2;     // 16 KB (N = 256); N = 128 stages use half
constexpr int kNumStages = 4;
constexpr int kHSBytes = kTileM * kU * 2;         // 64 KB
constexpr int kXSBytes = kTileM * 64 * 2;         // 16 KB
constexpr int kChunkA = kTileM * 16;              // bytes between 8-element K chunks of an A operand (2048)
constexpr int kFwdSteps = 10;
constexpr int kThreads = 320;
constexpr int kComputeThreads = 256;

// ---- forward step table --------------------------------------------------------------------------------
// step:        0    1..4   5      6,7   8          9
// layer:       L0   L1-4   L5     L6,7  features   rgb_features      (sigma and rgb heads run on CUDA cores)
__host__ __device__ constexpr int fwd_layer(int s) { return s < 8 ? s : (s == 8 ? 9 : 10); }
__host__ __device__ constexpr int fwd_nk_h(int s) { return s == 0 ? 0 : 8; }          // K stages fed by HS
__host__ __device__ constexpr int fwd_nk_x(int s) { return (s == 0 || s == 5) ? 2 : (s == 9 ? 1 : 0); }
__host__ __device__ constexpr int fwd_N(int s) { return s == 9 ? 128 : 256; }
__host__ __device__ constexpr int fwd_stage_bytes(int s) { return fwd_N(s) * kKStage * 2; }
__host__ __device__ constexpr int fwd_blob_off(int s) {
  int off = 0;
  for (int i = 0; i < s; ++i) off += (fwd_nk_h(i) + fwd_nk_x(i)) * fwd_stage_bytes(i);
  return off;
}
constexpr int kFwdBlobBytes = fwd_blob_off(kFwdSteps);
// fp32 side table appended to the blob (16-byte aligned copies; the Keras flat buffer is not: the 1-wide sigma
// bias shifts everything after it): bias[l] at l*256 (l = 0..11), sigma kernel at 12*256, rgb kernel at 13*256
constexpr int kAuxFloats = 12 * 256 + 256 + 512;
constexpr int kAuxOff = kFwdBlobBytes;
constexpr int kPackedBytes = kAuxOff + kAuxFloats * 4;   // + dgrad blob (added with the backward kernels)

struct TcParams {
  int64_t w_off[12], b_off[12];   // float offsets into the flat Keras-order parameter buffer
};

// per-tile record of the activations the backward needs (training only), all in chunk-major bf16
constexpr int kRecXS = 0;                          // PE(xyz)   [8 chunks][128][8]   16 KB
constexpr int kRecDS = 16384;                      // PE(dir)   [8 chunks][128][8]   16 KB (chunks 4..7 zero)
constexpr int kRecH0 = 32768;                      // h0..h7    8 x 64 KB
constexpr int kRecF = kRecH0 + 8 * kHSBytes;       // features  64 KB
constexpr int kRecG = kRecF + kHSBytes;            // rgb_features [16 chunks][128][8] 32 KB
constexpr int kRecBytes = kRecG + 32768;           // 640 KB per 128 samples

struct FwdSmem {
  uint8_t hs[2][kHSBytes];
  uint8_t xs[2][kXSBytes];
  uint8_t stage[kNumStages][kStageBytes];
  float part[kTileM][4];
  uint64_t full[kNumStages], empty[kNumStages], a_ready[2], acc_ready[2];
  uint32_t tmem_base;
};

// ---- weight packing --------------------------------------------------------------------------------------
// forward blob: per step, per K stage: [4 chunks][N][8] with element (c, n, e) = W[row(ks, c, e)][n]
__global__ void __launch_bounds__(256) pack_fwd_kernel(const float* __restrict__ params, TcParams P,
                                                       uint8_t* __restrict__ packed) {
  const int total_vec = kFwdBlobBytes / 16;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < total_vec; v += gridDim.x * blockDim.x) {
    int byte = v * 16, s = 0;
    while (s + 1 < kFwdSteps && byte >= fwd_blob_off(s + 1)) ++s;
    const int local = byte - fwd_blob_off(s);
    const int N = fwd_N(s), sb = fwd_stage_bytes(s);
    const int ks = local / sb, r = local - ks * sb;
    const int c = r / (N * 16), n = (r - c * N * 16) / 16;
    const int L = fwd_layer(s);
    const int fan_in = (L == 0) ? 63 : (L == 5) ? 319 : (L == 10) ? 283 : 256;
    const float* W = params + P.w_off[L];
    int row0;
    if (ks < fwd_nk_h(s)) row0 = ks * kKStage + c * 8;
    else row0 = (fwd_nk_h(s) > 0 ? 256 : 0) + (ks - fwd_nk_h(s)) * kKStage + c * 8;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ra = row0 + 2 * e, rb = ra + 1;
      const float a = ra < fan_in ? W[(int64_t)ra * N + n] : 0.f;
      const float b = rb < fan_in ? W[(int64_t)rb * N + n] : 0.f;
      w[e] = pack_bf16x2(a, b);
    }
    *reinterpret_cast<uint4*>(packed + byte) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  float* aux = reinterpret_cast<float*>(packed + kAuxOff);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kAuxFloats; i += gridDim.x * blockDim.x) {
    const int blk = i >> 8, j = i & 255;
    float v = 0.f;
    if (blk < 12) {
      const int fo = (blk == 8) ? 1 : (blk == 10) ? 128 : (blk == 11) ? 3 : 256;
      if (j < fo) v = params[P.b_off[blk] + j];
    } else if (blk == 12) {
      v = params[P.w_off[8] + j];                       // sigma kernel [256,1]
    } else {
      const int k = i - 13 * 256;
      if (k < 384) v = params[P.w_off[11] + k];         // rgb kernel [128,3]
    }
    aux[i] = v;
  }
}

// ---- positional encoding into an A operand -----------------------------------------------------------------
// thread (row r, half h) of the 256 compute threads; element (row, col) of a [128 x K] chunk-major operand
__device__ __forceinline__ void store_elem(uint8_t* base, int r, int col, float v) {
  *reinterpret_cast<__nv_bfloat16*>(base + (col >> 3) * kChunkA + r * 16 + (col & 7) * 2) = __float2bfloat16_rn(v);
}

// PE_10(o + d t) -> 63 columns (+ zero pad column 63): h=0 writes identity + frequencies 0..4, h=1 5..9
__device__ __forceinline__ void pe_xyz(uint8_t* xs, int r, int h, bool valid, const float* __restrict__ o,
                                       const float* __restrict__ d, const float* __restrict__ t, int64_t g, int S) {
  float p[3] = {0.f, 0.f, 0.f};
  if (valid) {
    const int64_t ray = g / S;
    const float tt = __ldg(t + g);
#pragma unroll
    for (int c = 0; c < 3; ++c) p[c] = __fadd_rn(__ldg(o + ray * 3 + c), __fmul_rn(__ldg(d + ray * 3 + c), tt));
  }
  if (h == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) store_elem(xs, r, c, p[c]);
  } else {
    store_elem(xs, r, 63, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int f = h * 5 + i;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sn, cs;
      sincosf(ldexpf(p[c], f), &sn, &cs);     // accurate range reduction: |arg| reaches ~3e3
      store_elem(xs, r, 3 + 6 * f + c, sn);
      store_elem(xs, r, 3 + 6 * f + 3 + c, cs);
    }
  }
}

// PE_4(d) -> 27 columns (+ zero pad 27..31) into chunks 0..3 of the xs buffer; h=0: identity + f 0,1; h=1: f 2,3
__device__ __forceinline__ void pe_dir(uint8_t* xs, int r, int h, bool valid, const float* __restrict__ d, int64_t g,
                                       int S) {
  float v[3] = {0.f, 0.f, 0.f};
  if (valid) {
    const int64_t ray = g / S;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = __ldg(d + ray * 3 + c);
  }
  if (h == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) store_elem(xs, r, c, v[c]);
  } else {
#pragma unroll
    for (int c = 27; c < 32; ++c) store_elem(xs, r, c, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int f = h * 2 + i;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float sn, cs;
      sincosf(ldexpf(v[c], f), &sn, &cs);
      store_elem(xs, r, 3 + 6 * f + c, sn);
      store_elem(xs, r, 3 + 6 * f + 3 + c, cs);
    }
  }
}

// copy `bytes` of a smem operand image to global (all 256 compute threads, 16 B vectors)
__device__ __forceinline__ void copy_smem_to_global(const uint8_t* src, uint8_t* dst, int bytes, int ctid) {
  for (int i = ctid * 16; i < bytes; i += kComputeThreads * 16)
    *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
}

// ---- the fused forward kernel ----------------------------------------------------------------------------
template <bool TRAIN>
__global__ void __launch_bounds__(kThreads, 1)
tc_mlp_fwd_kernel(const uint8_t* __restrict__ packed,
                  const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ t,
                  int64_t M, int S, float4* __restrict__ rgbsigma, uint8_t* __restrict__ rec) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (M + kTileM - 1) / kTileM;
  const int64_t n_pairs = (n_tiles + 1) / 2;

  if (tid == 0) {
    for (int i = 0; i < kNumStages; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.a_ready[i], kComputeThreads); mbar_init(&sm.acc_ready[i], 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (int s = 0; s < kFwdSteps; ++s) {
          const int nk = fwd_nk_h(s) + fwd_nk_x(s);
          const uint32_t sb = fwd_stage_bytes(s);
          const uint8_t* src = packed + fwd_blob_off(s);
          for (int tl = 0; tl < 2; ++tl) {
            for (int ks = 0; ks < nk; ++ks, ++it) {
              const uint32_t slot = it % kNumStages, ph = (it / kNumStages) & 1;
              mbar_wait(&sm.empty[slot], ph ^ 1);
              mbar_arrive_expect_tx(&sm.full[slot], sb);
              tma_load_1d(sm.stage[slot], src + (size_t)ks * sb, sb, &sm.full[slot]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =============================
    if (lane == 0) {
      uint32_t it = 0, a_par[2] = {0, 0};
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (int s = 0; s < kFwdSteps; ++s) {
          const int nkh = fwd_nk_h(s), nk = nkh + fwd_nk_x(s);
          const int N = fwd_N(s);
          const uint32_t idesc = umma_idesc_bf16(kTileM, N, 0, 0);
          const uint32_t chunk_b = (uint32_t)N * 16;
          for (int tl = 0; tl < 2; ++tl) {
            mbar_wait(&sm.a_ready[tl], a_par[tl]);
            a_par[tl] ^= 1;
            tc_fence_after();
            const uint32_t d_tmem = tmem + tl * 256;
            for (int ks = 0; ks < nk; ++ks, ++it) {
              const uint32_t slot = it % kNumStages, ph = (it / kNumStages) & 1;
              mbar_wait(&sm.full[slot], ph);
              tc_fence_after();
              const uint32_t a_base = (ks < nkh) ? smem_u32(sm.hs[tl]) + ks * 4 * kChunkA
                                                 : smem_u32(sm.xs[tl]) + (ks - nkh) * 4 * kChunkA;
              const uint32_t b_base = smem_u32(sm.stage[slot]);
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint64_t da = umma_smem_desc(a_base + j * 2 * kChunkA, kChunkA, 128);
                const uint64_t db = umma_smem_desc(b_base + j * 2 * chunk_b, chunk_b, 128);
                umma_bf16(d_tmem, da, db, idesc, (ks > 0 || j > 0) ? 1u : 0u);
              }
              umma_commit(&sm.empty[slot]);
            }
            umma_commit(&sm.acc_ready[tl]);
          }
        }
      }
    }
  } else {
    // =========================== compute warps ==========================
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int ctid = tid - 64;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t acc_par[2] = {0, 0};
    float sig_keep[2] = {0.f, 0.f};

    auto prologue = [&](int64_t pair, int tl) {
      const int64_t tile = pair * 2 + tl;
      const int64_t g = tile * kTileM + r;
      pe_xyz(sm.xs[tl], r, h, g < M, o, d, t, g, S);
      if (TRAIN && tile < n_tiles) {
        named_bar_sync(1, kComputeThreads);
        copy_smem_to_global(sm.xs[tl], rec + tile * kRecBytes + kRecXS, kXSBytes, ctid);
      }
      fence_async_smem();
      mbar_arrive(&sm.a_ready[tl]);
    };

    if ((int64_t)blockIdx.x < n_pairs) {
      prologue(blockIdx.x, 0);
      prologue(blockIdx.x, 1);
    }
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      for (int s = 0; s < kFwdSteps; ++s) {
        const int L = fwd_layer(s);
        const float* aux = reinterpret_cast<const float*>(packed + kAuxOff);
        const float* bias = aux + L * 256;
#pragma unroll
        for (int tl = 0; tl < 2; ++tl) {
          const int64_t tile = pair * 2 + tl;
          const int64_t g = tile * kTileM + r;
          const bool valid = g < M;
          const bool save = TRAIN && tile < n_tiles;
          uint8_t* rec_t = rec + tile * kRecBytes;
          mbar_wait(&sm.acc_ready[tl], acc_par[tl]);
          acc_par[tl] ^= 1;
          tc_fence_after();

          if (s < 9) {
            // hidden layers (bias + ReLU) and `features` (bias, linear): 128 of the 256 columns per thread
            uint8_t* rec_out = rec_t + (s < 8 ? kRecH0 + s * kHSBytes : kRecF);
            float sigdot = 0.f;
            const float* wsig = aux + 12 * 256;
#pragma unroll 1
            for (int gI = 0; gI < 4; ++gI) {
              const int col0 = h * 128 + gI * 32;
              float v[32];
              tmem_ld32(tmem + lane_base + tl * 256 + col0, v);
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                const int col = col0 + c8 * 8;
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float x[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  x[e] = v[c8 * 8 + e] + bb[e];
                  if (s < 8) x[e] = fmaxf(x[e], 0.f);                    // mlp.py:33-34 (features is linear: :42)
                }
                if (s == 7) {                                            // sigma head on the fp32 h7 (mlp.py:40)
                  const float4 w0 = __ldg(reinterpret_cast<const float4*>(wsig + col));
                  const float4 w1 = __ldg(reinterpret_cast<const float4*>(wsig + col + 4));
                  sigdot += x[0] * w0.x + x[1] * w0.y + x[2] * w0.z + x[3] * w0.w + x[4] * w1.x + x[5] * w1.y +
                            x[6] * w1.z + x[7] * w1.w;
                }
                const uint4 pk = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                                            pack_bf16x2(x[6], x[7]));
                const int off = (col >> 3) * kChunkA + r * 16;
                *reinterpret_cast<uint4*>(sm.hs[tl] + off) = pk;        // next layer's A operand, in place
                if (save) *reinterpret_cast<uint4*>(rec_out + off) = pk;
              }
            }
            if (s == 7) {
              if (h == 1) sm.part[r][0] = sigdot;
              named_bar_sync(1, kComputeThreads);
              if (h == 0) sig_keep[tl] = fmaxf(sigdot + sm.part[r][0] + __ldg(aux + 8 * 256), 0.f);
              named_bar_sync(1, kComputeThreads);
            }
            if (s == 5) {
              // xs[tl] is dead after layer 5 (skip concat consumed): it now carries PE(dir) for rgb_features
              pe_dir(sm.xs[tl], r, h, valid, d, g, S);
              if (save) {
                named_bar_sync(1, kComputeThreads);
                copy_smem_to_global(sm.xs[tl], rec_t + kRecDS, 8192, ctid);
                for (int i = ctid * 16; i < 8192; i += kComputeThreads * 16)
                  *reinterpret_cast<uint4*>(rec_t + kRecDS + 8192 + i) = make_uint4(0, 0, 0, 0);
              }
            }
            tc_fence_before();
            fence_async_smem();
            mbar_arrive(&sm.a_ready[tl]);
          } else {
            // rgb_features (bias, linear; mlp.py:43-46) then the rgb head + sigmoid on CUDA cores (mlp.py:48)
            const float* wrgb = aux + 13 * 256;
            float pr = 0.f, pg = 0.f, pb = 0.f;
#pragma unroll 1
            for (int gI = 0; gI < 2; ++gI) {
              const int col0 = h * 64 + gI * 32;
              float v[32];
              tmem_ld32(tmem + lane_base + tl * 256 + col0, v);
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                const int col = col0 + c8 * 8;
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float x[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  x[e] = v[c8 * 8 + e] + bb[e];
                  const float* wr = wrgb + (col + e) * 3;
                  pr += x[e] * __ldg(wr); pg += x[e] * __ldg(wr + 1); pb += x[e] * __ldg(wr + 2);
                }
                if (save) {
                  const uint4 pk = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]),
                                              pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                  *reinterpret_cast<uint4*>(rec_t + kRecG + (col >> 3) * kChunkA + r * 16) = pk;
                }
              }
            }
            if (h == 1) { sm.part[r][0] = pr; sm.part[r][1] = pg; sm.part[r][2] = pb; }
            named_bar_sync(1, kComputeThreads);
            if (h == 0 && valid) {
              const float* brgb = aux + 11 * 256;
              const float zr = pr + sm.part[r][0] + __ldg(brgb), zg = pg + sm.part[r][1] + __ldg(brgb + 1),
                          zb = pb + sm.part[r][2] + __ldg(brgb + 2);
              rgbsigma[g] = make_float4(1.f / (1.f + expf(-zr)), 1.f / (1.f + expf(-zg)), 1.f / (1.f + expf(-zb)),
                                        sig_keep[tl]);
            }
            named_bar_sync(1, kComputeThreads);
            tc_fence_before();
            // this tile is finished: start the next pair's tile in the same slot right away
            const int64_t next = pair + gridDim.x;
            if (next < n_pairs) prologue(next, tl);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace

bool tc_path_compiled() { return true; }

int64_t tc_packed_weight_bytes(const Model& m) { return is_flagship(m) ? kPackedBytes : -1; }

int64_t tc_workspace_bytes(const Model& m, int64_t rows, bool training) {
  if (!is_flagship(m)) return -1;
  if (!training) return 256;
  return cdiv(rows, kTileM) * (int64_t)kRecBytes + 256;
}

static TcParams make_params(const Model& m) {
  TcParams P;
  for (int i = 0; i < 12; ++i) { P.w_off[i] = m.L[i].w_off; P.b_off[i] = m.L[i].b_off; }
  return P;
}

int tc_pack_weights(const Model& m, const float* params, void* packed, cudaStream_t st) {
  if (!is_flagship(m)) return fail(KNERF_ERR_UNSUPPORTED, "KNERF_BF16 implements the 8x256 / skip 4 / L=10,4 model only");
  pack_fwd_kernel<<<kNumSMs, 256, 0, st>>>(params, make_params(m), (uint8_t*)packed);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

int tc_forward(const Model& m, const float* params, const void* packed, const float* o, const float* d, const float* t,
               int64_t R, int S, bool training, float* rgbsigma, char* ws, int64_t ws_bytes, cudaStream_t st) {
  if (!is_flagship(m)) return fail(KNERF_ERR_UNSUPPORTED, "KNERF_BF16 implements the 8x256 / skip 4 / L=10,4 model only");
  const int64_t M = R * S;
  if (training && ws_bytes < tc_workspace_bytes(m, M, true))
    return fail(KNERF_ERR_WORKSPACE, "tc_forward: workspace %lld < %lld bytes", (long long)ws_bytes,
                (long long)tc_workspace_bytes(m, M, true));
  KN_CHECK_ARG((reinterpret_cast<uintptr_t>(rgbsigma) & 15) == 0 && (reinterpret_cast<uintptr_t>(packed) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(params) & 15) == 0,
               "tc_forward: rgbsigma / packed / params must be 16-byte aligned");
  const int64_t n_pairs = cdiv(cdiv(M, kTileM), 2);
  const int grid = (int)std::min<int64_t>(n_pairs, kNumSMs);
  const size_t smem = sizeof(FwdSmem);
  if (training) {
    KN_CUDA(cudaFuncSetAttribute(tc_mlp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_mlp_fwd_kernel<true><<<grid, kThreads, smem, st>>>((const uint8_t*)packed, o, d, t, M, S,
                                                          (float4*)rgbsigma, (uint8_t*)ws);
  } else {
    KN_CUDA(cudaFuncSetAttribute(tc_mlp_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_mlp_fwd_kernel<false><<<grid, kThreads, smem, st>>>((const uint8_t*)packed, o, d, t, M, S,
                                                           (float4*)rgbsigma, nullptr);
  }
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

int tc_backward(const Model&, const float*, const void*, const float*, int64_t, int, float*, char*, int64_t,
                cudaStream_t) {
  return fail(KNERF_ERR_UNSUPPORTED, "KNERF_BF16 backward not built yet");
}

}  // namespace knerf
