// a6 (KNERF_FP32_TC mode): the fp32 NeRF MLP on Blackwell tensor cores (keras_nerf/model/nerf/mlp.py:29-50; the
// reference is fp32 end to end: mlp.py:11-27 sets no mixed-precision policy).
//
// tcgen05 has no fp32 x fp32 MMA.  Every fp32 operand x is therefore split into THREE bf16 values
//     x = x1 + x2 + x3,   x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2)      (24 mantissa bits in total)
// and a product a*b is formed as the six leading terms  a1 b1 + a1 b2 + a2 b1 + a1 b3 + a2 b2 + a3 b1
// (the dropped ones are below 2^-24 |a b|), each an exact bf16 x bf16 product accumulated in fp32 in TMEM: fp32-grade
// results at one sixth of the bf16 tensor rate -- ~9x what the SIMT FFMA path reaches (mlp_fp32.cu).
//
// Two kernels, both one GEMM per Dense layer with fp32 activations in HBM (like the SIMT mode, whose host code
// -- fp32_forward_core / fp32_backward_core -- drives them):
//   tcx_gemm_kernel   C[M,N] = epi(A1 B1 + A2 B2 + bias): forward layers and dgrad.  A tile: fp32 rows loaded and
//                     split by the compute warps into the chunk-major K-major layout of tc_ptx.cuh; B: the weights,
//                     split ONCE per call into operand blobs (tcx_pack_kernel) and streamed by bulk copies.
//   tcx_wgrad_kernel  dW[K,N] += A[M,K]^T Z[M,N], reduction over samples: both operands split on the fly into the
//                     MN-major form (same bytes, other axis), accumulators resident in TMEM over a slab of samples.
#include "mlp_fp32.cuh"
#include "tc_ptx.cuh"

namespace knerf {
using namespace tc;

namespace {

constexpr int kXThreads = 320;          // warps 0-7 compute (split + epilogue), warp 8 = MMA issue, warp 9 = bulk copies
// the six products, ordered from the smallest terms to the largest (the order only matters for rounding)
__device__ __constant__ int8_t kProdA[6] = {2, 1, 0, 1, 0, 0};
__device__ __constant__ int8_t kProdB[6] = {0, 1, 2, 0, 1, 0};

// 8 consecutive fp32 -> three 16-byte vectors of bf16 (x1 | x2 | x3), element e in half (e & 1) of word e >> 1
__device__ __forceinline__ void split3x8(const float (&x)[8], uint4& p1, uint4& p2, uint4& p3) {
  uint32_t w1[4], w2[4], w3[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = x[2 * i], b = x[2 * i + 1];
    w1[i] = pack_bf16x2(a, b);
    const float ra = a - bf16_lo(w1[i]), rb = b - bf16_hi(w1[i]);      // exact
    w2[i] = pack_bf16x2(ra, rb);
    w3[i] = pack_bf16x2(ra - bf16_lo(w2[i]), rb - bf16_hi(w2[i]));
  }
  p1 = make_uint4(w1[0], w1[1], w1[2], w1[3]);
  p2 = make_uint4(w2[0], w2[1], w2[2], w2[3]);
  p3 = make_uint4(w3[0], w3[1], w3[2], w3[3]);
}

// ---- weight operand blobs -------------------------------------------------------------------------------------
// B[n][k] (n < N outputs, k < K reduction) = src[k * ld + n] (TRANS = false: a Keras kernel [in,out] used forward) or
// src[n * ld + k] (TRANS = true: the same kernel used by dgrad).  Blob: [K16 steps][3 splits][2 chunks][N][8] bf16,
// zero beyond K.  One thread per 16-byte vector of split 0 (writes the same vector of all three splits).
__global__ void __launch_bounds__(256) tcx_pack_kernel(const float* __restrict__ src, int ld, int N, int K, int trans,
                                                       uint8_t* __restrict__ blob) {
  const int ksteps = (K + 15) / 16;
  const int total = ksteps * 2 * N;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < total; v += gridDim.x * blockDim.x) {
    const int n = v % N, c = (v / N) & 1, ks = v / (2 * N);
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = ks * 16 + c * 8 + e;
      x[e] = (k < K) ? (trans ? src[(int64_t)n * ld + k] : src[(int64_t)k * ld + n]) : 0.f;
    }
    uint4 p1, p2, p3;
    split3x8(x, p1, p2, p3);
    uint8_t* base = blob + (size_t)ks * 96 * N + (size_t)(c * N + n) * 16;
    *reinterpret_cast<uint4*>(base) = p1;
    *reinterpret_cast<uint4*>(base + 32 * N) = p2;
    *reinterpret_cast<uint4*>(base + 64 * N) = p3;
  }
}

// ---- C = epi(A1 B1 + A2 B2 + bias) ----------------------------------------------------------------------------
// One CTA per SM takes 256 rows = TWO 128-row tiles that share every weight stage: the split weights are 6 bytes per
// element, and with one tile per stage the bulk copies alone (384 KB per tile and layer) sat at the ~42 B/clk an SM
// can ingest from L2 (measured: 36 % of the MMA-bound time with 128-row CTAs, two per SM).
constexpr int kGStages = 4;
struct GemmSmem {
  // per stage: A [2 tiles][3 splits][2 chunks][128 rows][8] = 24 KB, B [3][2][N <= 256][8] = 24 KB
  uint8_t a[kGStages][2 * 3 * 4096];
  uint8_t b[kGStages][3 * 8192];
  uint64_t full[kGStages], a_ready[kGStages], empty[kGStages], acc_ready;
  uint32_t tmem_base;
};

struct XGemmArgs {
  const float* A1; int lda1; int K1; const uint8_t* B1;
  const float* A2; int lda2; int K2; const uint8_t* B2;
  const float* bias; float* C; int ldc; int64_t M; int N; int epi;
  const float* mask; int ldmask;
};

// 8 fp32 of row `grow`, reduction columns k0 .. k0 + 7 of source A (zero beyond M / K)
__device__ __forceinline__ void load_a8(const float* __restrict__ A, int lda, int K, int64_t grow, int64_t M, int k0,
                                        float (&x)[8]) {
  if (grow < M && k0 + 8 <= K && (lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0) {
    const float4* p = reinterpret_cast<const float4*>(A + grow * lda + k0);
    const float4 v0 = __ldg(p), v1 = __ldg(p + 1);
    x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = (grow < M && k0 + e < K) ? __ldg(A + grow * lda + k0 + e) : 0.f;
  }
}

__global__ void __launch_bounds__(kXThreads, 1) tcx_gemm_kernel(XGemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  GemmSmem& sm = *reinterpret_cast<GemmSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * 256;
  const int N = g.N;
  const int ks1 = (g.K1 + 15) / 16, ks2 = (g.K2 + 15) / 16, ksteps = ks1 + ks2;
  if (tid == 0) {
    for (int i = 0; i < kGStages; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.a_ready[i], 8); mbar_init(&sm.empty[i], 1); }
    mbar_init(&sm.acc_ready, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t stage_bytes = 96u * (uint32_t)N;

  if (warp == 9) {
    // ---- weight-operand producer: one bulk copy per K = 16 step ----
    if (lane == 0) {
      for (int kt = 0; kt < ksteps; ++kt) {
        const int s = kt % kGStages;
        if (kt >= kGStages) mbar_wait(&sm.empty[s], ((kt / kGStages) - 1) & 1);
        const uint8_t* src = (kt < ks1) ? g.B1 + (size_t)kt * stage_bytes : g.B2 + (size_t)(kt - ks1) * stage_bytes;
        mbar_arrive_expect_tx(&sm.full[s], stage_bytes);
        tma_load_1d(sm.b[s], src, stage_bytes, &sm.full[s]);
      }
    }
  } else if (warp == 8) {
    // ---- MMA issuer: per step and tile six bf16 x bf16 products into the tile's fp32 accumulator ----
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
      for (int kt = 0; kt < ksteps; ++kt) {
        const int s = kt % kGStages;
        const uint32_t ph = (kt / kGStages) & 1;
        mbar_wait(&sm.full[s], ph);
        mbar_wait(&sm.a_ready[s], ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sm.a[s]), b0 = smem_u32(sm.b[s]);
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            const uint64_t da = umma_smem_desc(a0 + tile * 12288 + kProdA[p] * 4096, 2048, 128);
            const uint64_t db = umma_smem_desc(b0 + kProdB[p] * 32 * N, 16 * N, 128);
            umma_bf16(tmem + tile * 256, da, db, idesc, (kt > 0 || p > 0) ? 1u : 0u);
          }
        }
        umma_commit(&sm.empty[s]);
      }
      umma_commit(&sm.acc_ready);
    }
  } else {
    // ---- compute warps: thread = one row of one of the two tiles; it splits the row's 16 reduction columns of every
    //      K = 16 step (the next step's are already in flight), then the epilogue ----
    const int tile = tid >> 7, row = tid & 127;
    const int64_t grow = m0 + tile * 128 + row;
    auto fetch = [&](int kt, float (&lo)[8], float (&hi)[8]) {
      const bool second = kt >= ks1;
      const float* A = second ? g.A2 : g.A1;
      const int lda = second ? g.lda2 : g.lda1, K = second ? g.K2 : g.K1, k0 = (second ? kt - ks1 : kt) * 16;
      load_a8(A, lda, K, grow, g.M, k0, lo);
      load_a8(A, lda, K, grow, g.M, k0 + 8, hi);
    };
    // The rows are 64 bytes per step and thread: with one step in flight the loop ran at the memory LATENCY (~4,000
    // cycles per step against 1,536 of MMA time).  Three steps are kept in flight in a register ring (the ring index
    // is a compile-time constant of the 4x unrolled body).
    float xlo[4][8], xhi[4][8];
#pragma unroll
    for (int u = 0; u < 3; ++u)
      if (u < ksteps) fetch(u, xlo[u], xhi[u]);
    for (int kt0 = 0; kt0 < ksteps; kt0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int kt = kt0 + u;
        if (kt < ksteps) {
          if (kt + 3 < ksteps) fetch(kt + 3, xlo[(u + 3) & 3], xhi[(u + 3) & 3]);
          uint4 p1, p2, p3, q1, q2, q3;
          split3x8(xlo[u], p1, p2, p3);
          split3x8(xhi[u], q1, q2, q3);
          const int s = kt % kGStages;
          if (kt >= kGStages) mbar_wait(&sm.empty[s], ((kt / kGStages) - 1) & 1);
          uint8_t* dst = sm.a[s] + tile * 12288 + row * 16;
          *reinterpret_cast<uint4*>(dst) = p1;
          *reinterpret_cast<uint4*>(dst + 2048) = q1;
          *reinterpret_cast<uint4*>(dst + 4096) = p2;
          *reinterpret_cast<uint4*>(dst + 4096 + 2048) = q2;
          *reinterpret_cast<uint4*>(dst + 8192) = p3;
          *reinterpret_cast<uint4*>(dst + 8192 + 2048) = q3;
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.a_ready[s]);
        }
      }
    }
    // ---- epilogue: TMEM lane quadrant = warp & 3 (rows), column half = warp >> 2, both tiles ----
    mbar_wait(&sm.acc_ready, 0);
    tc_fence_after();
    const int q = warp & 3, hcol = warp >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int ncol_half = N / 2;                               // N is a multiple of 64
#pragma unroll 1
    for (int tl = 0; tl < 2; ++tl) {
      const int64_t orow = m0 + tl * 128 + q * 32 + lane;
      for (int c0 = hcol * ncol_half; c0 < (hcol + 1) * ncol_half; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + lane_base + tl * 256 + c0, v);        // warp-collective: outside the row guard
        if (orow < g.M) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            if (g.bias != nullptr) {
              // (scalar loads: the flat Keras-order parameter buffer is not 16-byte aligned past the 1-wide sigma bias)
              o.x += __ldg(g.bias + c0 + i); o.y += __ldg(g.bias + c0 + i + 1);
              o.z += __ldg(g.bias + c0 + i + 2); o.w += __ldg(g.bias + c0 + i + 3);
            }
            if (g.epi == 1) {          // EPI_RELU
              o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            } else if (g.epi == 3) {   // EPI_MASK: C = acc * (mask > 0)
              const float4 mk = __ldg(reinterpret_cast<const float4*>(g.mask + orow * g.ldmask + c0 + i));
              o.x = mk.x > 0.f ? o.x : 0.f; o.y = mk.y > 0.f ? o.y : 0.f;
              o.z = mk.z > 0.f ? o.z : 0.f; o.w = mk.w > 0.f ? o.w : 0.f;
            }
            *reinterpret_cast<float4*>(g.C + orow * g.ldc + c0 + i) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

// ---- dW[K,N] += A[M,K]^T Z[M,N] -------------------------------------------------------------------------------
constexpr int kWSamples = 32;           // samples per stage (two K = 16 MMA steps)
constexpr int kWStages = 2;
struct WgSmem {
  // per stage: A [3 splits][32 feature chunks][32 samples][8] = 48 KB, Z likewise ([N/8 chunks])
  uint8_t a[kWStages][3 * 16384];
  uint8_t z[kWStages][3 * 16384];
  uint64_t ready[kWStages], empty[kWStages], acc_ready;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kXThreads, 1)
tcx_wgrad_kernel(const float* __restrict__ A, int lda, int K, const float* __restrict__ Z, int ldz, int N, int64_t M,
                 int64_t slab, float* __restrict__ dW, int ldw) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  WgSmem& sm = *reinterpret_cast<WgSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t mb = (int64_t)blockIdx.x * slab, me = min(mb + slab, M);
  const int nst = (int)((me - mb + kWSamples - 1) / kWSamples);
  const int nch = N / 8;                             // 8-wide feature chunks of Z
  const int halves = (K + 127) / 128;                // 128-row accumulator blocks (TMEM columns h * 256)
  if (tid == 0) {
    for (int i = 0; i < kWStages; ++i) { mbar_init(&sm.ready[i], 8); mbar_init(&sm.empty[i], 1); }
    mbar_init(&sm.acc_ready, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 8) {
    if (lane == 0 && nst > 0) {
      const uint32_t idesc = umma_idesc_bf16(128, N, 1, 1);
      for (int st = 0; st < nst; ++st) {
        const int s = st % kWStages;
        mbar_wait(&sm.ready[s], (st / kWStages) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sm.a[s]), z0 = smem_u32(sm.z[s]);
        for (int h = 0; h < halves; ++h) {
#pragma unroll
          for (int p = 0; p < 6; ++p) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {   // MN-major: LBO = 128 B (next 8 samples), SBO = 512 B (next 8 features)
              const uint64_t da = umma_smem_desc(a0 + kProdA[p] * 16384 + h * 8192 + ks * 256, 128, 512);
              const uint64_t dz = umma_smem_desc(z0 + kProdB[p] * 16384 + ks * 256, 128, 512);
              umma_bf16(tmem + h * 256, da, dz, idesc, (st > 0 || p > 0 || ks > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(&sm.empty[s]);
      }
      umma_commit(&sm.acc_ready);
    }
  } else if (warp < 8) {
    // ---- compute warps: lane = sample of the stage, the warp walks the feature chunks (conflict-free smem stores,
    //      32-byte global sectors fully used) ----
    for (int st = 0; st < nst; ++st) {
      const int s = st % kWStages;
      const int64_t r = mb + (int64_t)st * kWSamples + lane;
      const bool live = r < me;
#pragma unroll 1
      for (int op = 0; op < 2; ++op) {
        const float* src = op ? Z : A;
        const int ld = op ? ldz : lda, width = op ? N : K;
        uint8_t* dst = op ? sm.z[s] : sm.a[s];
        const int cmax = op ? nch : 16 * halves;       // A: zero-fill up to the next 128-row block
        // this warp's (up to four) feature chunks: all loads first, then split + store
        float x[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cidx = warp + 8 * j;
          if (cidx < cmax) load_a8(src, ld, width, live ? r : me, me, cidx * 8, x[j]);
        }
        if (op == 0 && st >= kWStages) mbar_wait(&sm.empty[s], ((st / kWStages) - 1) & 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cidx = warp + 8 * j;
          if (cidx < cmax) {
            uint4 p1, p2, p3;
            split3x8(x[j], p1, p2, p3);
            uint8_t* d = dst + cidx * 512 + lane * 16;
            *reinterpret_cast<uint4*>(d) = p1;
            *reinterpret_cast<uint4*>(d + 16384) = p2;
            *reinterpret_cast<uint4*>(d + 32768) = p3;
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.ready[s]);
    }
    if (nst > 0) {
      mbar_wait(&sm.acc_ready, 0);
      tc_fence_after();
      const int q = warp & 3, hcol = warp >> 2;
      const uint32_t lane_base = (uint32_t)(q * 32) << 16;
      for (int h = 0; h < halves; ++h) {
        const int krow = h * 128 + q * 32 + lane;
        for (int c0 = hcol * (N / 2); c0 < (hcol + 1) * (N / 2); c0 += 32) {
          float v[32];
          tmem_ld32(tmem + lane_base + h * 256 + c0, v);
          if (krow < K) {
            float* dst = dW + (int64_t)krow * ldw + c0;
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(dst + i, v[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------
size_t tcx_blob_bytes(int N, int K) { return (size_t)((K + 15) / 16) * 96 * N; }

int tcx_pack(const float* src, int ld, int N, int K, bool trans, void* blob, cudaStream_t st) {
  const int total = ((K + 15) / 16) * 2 * N;
  if (total == 0) return KNERF_OK;
  tcx_pack_kernel<<<std::min((total + 255) / 256, kNumSMs * 4), 256, 0, st>>>(src, ld, N, K, trans ? 1 : 0, (uint8_t*)blob);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

// shapes the tensor kernels take: N a multiple of 64 up to 256, 16-byte aligned rows of C / bias / mask
bool tcx_gemm_eligible(const GemmArgs& g) {
  return g.N >= 64 && g.N <= 256 && g.N % 64 == 0 && (g.ldc & 3) == 0 && (g.K1 + g.K2) >= 16 &&
         (g.epi == EPI_NONE || g.epi == EPI_RELU || (g.epi == EPI_MASK && (g.ldmask & 3) == 0)) &&
         (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 &&
         (g.epi != EPI_MASK || (reinterpret_cast<uintptr_t>(g.mask) & 15) == 0);
}

int launch_gemm_tc(const GemmArgs& g, const void* blob1, const void* blob2, cudaStream_t st) {
  if (g.M == 0) return KNERF_OK;
  XGemmArgs x{g.A1, g.lda1, g.K1, (const uint8_t*)blob1, g.A2, g.lda2, g.K2, (const uint8_t*)blob2,
              g.bias, g.C, g.ldc, g.M, g.N, g.epi, g.mask, g.ldmask};
  if (x.A1 == nullptr) x.K1 = 0;
  if (x.A2 == nullptr) x.K2 = 0;
  const size_t smem = sizeof(GemmSmem);
  KN_CUDA(cudaFuncSetAttribute(tcx_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tcx_gemm_kernel<<<(unsigned)cdiv(g.M, 256), kXThreads, smem, st>>>(x);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

bool tcx_wgrad_eligible(int K, int N) { return K >= 16 && K <= 256 && N >= 64 && N <= 256 && N % 64 == 0; }

int launch_wgrad_tc(const float* A, int lda, int K, const float* Z, int ldz, int N, int64_t M, float* dW, int ldw,
                    cudaStream_t st) {
  if (M == 0) return KNERF_OK;
  const int64_t stages = cdiv(M, kWSamples);
  const int grid = (int)std::min<int64_t>(kNumSMs, stages);
  const int64_t slab = cdiv(stages, grid) * kWSamples;
  const size_t smem = sizeof(WgSmem);
  KN_CUDA(cudaFuncSetAttribute(tcx_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tcx_wgrad_kernel<<<(unsigned)cdiv(M, slab), kXThreads, smem, st>>>(A, lda, K, Z, ldz, N, M, slab, dW, ldw);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

}  // namespace knerf
