// a6 (fp32 parity mode): NeRFMLP forward and backward on SIMT FFMA  (keras_nerf/model/nerf/mlp.py:29-50).
//
// This is the KNERF_FP32 precision mode: fp32 operands, fp32 accumulation, any layer widths.  It exists
// because the reference is fp32 end to end and the north-star parity bar (1e-5 on composited RGB) cannot
// be met with bf16/TF32 operands; the throughput path is the tcgen05 kernel in mlp_tc.cu.
// Activations round-trip through HBM here (one tiled SGEMM per Dense layer), concats are expressed as
// two-source GEMMs ([h, x] @ W == h @ W[:U] + x @ W[U:]) so no concat tensor is materialised.
#include "common.cuh"
#include "mlp_fp32.cuh"

namespace knerf {

// ---------------------------------------------------------------------------------------------------
// C[M,N] = epi( A1[M,K1] @ op(B1) + A2[M,K2] @ op(B2) + bias )
//   op(B) = B[K,N] row-major (TRANS_B = false, Keras kernel layout) or B[N,K]^T (TRANS_B = true: dgrad)
// 128x128x8 tiles, 256 threads, 8x8 register micro-tile, all loads bounds-checked (any M, N, K, ld).
// ---------------------------------------------------------------------------------------------------
constexpr int GBM = 128, GBN = 128, GBK = 8, GPAD = 4;

template <bool TRANS_B>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][GBK][GBM + GPAD];
  __shared__ __align__(16) float Bs[2][GBK][GBN + GPAD];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * GBM;
  const int n0 = blockIdx.y * GBN;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // global -> register staging: A tile 128x8 (row = tid/2, 4 consecutive k), B tile 8x128
  const int a_row = tid >> 1, a_k0 = (tid & 1) * 4;
  const int b_k = tid >> 5, b_n0 = (tid & 31) * 4;     // !TRANS_B: row k, 4 consecutive n
  const int bt_n = tid >> 1, bt_k0 = (tid & 1) * 4;    //  TRANS_B: row n, 4 consecutive k
  float ra[4], rb[4];

  const int kt1 = (g.K1 + GBK - 1) / GBK, kt2 = (g.K2 + GBK - 1) / GBK;
  const int ktiles = kt1 + kt2;

  auto load_tile = [&](int kt) {
    const bool second = kt >= kt1;
    const float* A = second ? g.A2 : g.A1;
    const float* B = second ? g.B2 : g.B1;
    const int lda = second ? g.lda2 : g.lda1, ldb = second ? g.ldb2 : g.ldb1;
    const int K = second ? g.K2 : g.K1;
    const int k0 = (second ? kt - kt1 : kt) * GBK;
    const int64_t row = m0 + a_row;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + a_k0 + q;
      ra[q] = (row < g.M && k < K) ? A[row * lda + k] : 0.f;
    }
    if (!TRANS_B) {
      const int k = k0 + b_k;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int n = n0 + b_n0 + q;
        rb[q] = (k < K && n < g.N) ? B[(int64_t)k * ldb + n] : 0.f;
      }
    } else {
      const int n = n0 + bt_n;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = k0 + bt_k0 + q;
        rb[q] = (k < K && n < g.N) ? B[(int64_t)n * ldb + k] : 0.f;
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q) As[buf][a_k0 + q][a_row] = ra[q];
    if (!TRANS_B) {
      *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n0]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) Bs[buf][bt_k0 + q][bt_n] = rb[q];
    }
  };

  if (ktiles > 0) {
    load_tile(0);
    store_tile(0);
  }
  __syncthreads();
  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < ktiles) store_tile(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (col >= g.N) continue;
      float v = acc[i][j];
      if (g.bias != nullptr) v += g.bias[col];
      if (g.epi == EPI_RELU) v = fmaxf(v, 0.f);
      else if (g.epi == EPI_SIGMOID) v = 1.0f / (1.0f + expf(-v));
      else if (g.epi == EPI_MASK) v = (g.mask[row * g.ldmask + col] > 0.f) ? v : 0.f;
      g.C[row * g.ldc + col] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Heads with 1..4 outputs (sigma: N = 1, rgb: N = 3): C[M,N] = epi(A1 B1 + A2 B2 + bias) is a bandwidth problem (it
// reads every activation row once), which the 128-wide tiles above turn into a compute one.  One warp per row: the
// lanes stride over the reduction, N <= 4 partial sums each, butterfly reduction in a fixed order.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) skinny_gemm_kernel(GemmArgs g) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * 8;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < g.M; row += warps) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int src = 0; src < 2; ++src) {
      const float* A = src ? g.A2 : g.A1;
      const float* B = src ? g.B2 : g.B1;
      const int lda = src ? g.lda2 : g.lda1, ldb = src ? g.ldb2 : g.ldb1, K = (A == nullptr) ? 0 : (src ? g.K2 : g.K1);
      for (int k = lane; k < K; k += 32) {
        const float a = A[row * lda + k];
#pragma unroll
        for (int n = 0; n < 4; ++n)
          if (n < g.N) acc[n] = fmaf(a, __ldg(B + (int64_t)k * ldb + n), acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[n] = warp_sum(acc[n]);
    if (lane < g.N) {
      float v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
      if (g.bias != nullptr) v += g.bias[lane];
      if (g.epi == EPI_RELU) v = fmaxf(v, 0.f);
      else if (g.epi == EPI_SIGMOID) v = 1.0f / (1.0f + expf(-v));
      g.C[row * g.ldc + lane] = v;
    }
  }
}

// dW[K,N] += A[M,K]^T Z[M,N] for N <= 4: thread k of a block owns row k of dW, the block walks a slab of samples
// (A rows read coalesced, the Z row broadcast), one red.global.add per element and slab
__global__ void __launch_bounds__(256)
skinny_wgrad_kernel(const float* __restrict__ A, int lda, int K, const float* __restrict__ Z, int ldz, int N, int64_t M,
                    int64_t slab, float* __restrict__ dW, int ldw) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int64_t mb = (int64_t)blockIdx.y * slab, me = min(mb + slab, M);
  if (k >= K) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t r = mb; r < me; ++r) {
    const float a = A[r * lda + k];
#pragma unroll
    for (int n = 0; n < 4; ++n)
      if (n < N) acc[n] = fmaf(a, __ldg(Z + r * ldz + n), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < 4; ++n)
    if (n < N) atomicAdd(&dW[(int64_t)k * ldw + n], acc[n]);
}

// blob1 / blob2: split weight operands of the two sources (KNERF_FP32_TC) -- when given and the shape qualifies, the
// GEMM runs on the tensor cores (mlp_fp32_tc.cu); otherwise SIMT
static int launch_gemm(const GemmArgs& g, bool trans_b, cudaStream_t st, const void* blob1 = nullptr,
                       const void* blob2 = nullptr) {
  if (g.M == 0 || g.N == 0) return KNERF_OK;
  if (!trans_b && g.N <= 4 && g.epi != EPI_MASK) {
    skinny_gemm_kernel<<<(unsigned)std::min<int64_t>(cdiv(g.M, 8), (int64_t)kNumSMs * 16), 256, 0, st>>>(g);
    KN_LAUNCH_CHECK();
    return KNERF_OK;
  }
  if ((blob1 != nullptr || g.K1 == 0 || g.A1 == nullptr) && (blob2 != nullptr || g.K2 == 0 || g.A2 == nullptr) &&
      (blob1 != nullptr || blob2 != nullptr) && tcx_gemm_eligible(g))
    return launch_gemm_tc(g, blob1, blob2, st);
  dim3 grid((unsigned)cdiv(g.M, GBM), (unsigned)cdiv(g.N, GBN));
  if (trans_b) sgemm_kernel<true><<<grid, 256, 0, st>>>(g);
  else sgemm_kernel<false><<<grid, 256, 0, st>>>(g);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

// ---------------------------------------------------------------------------------------------------
// dW[K,N] += A[M,K]^T @ Z[M,N]   (weight gradient; reduction over the M samples, split over blockIdx.z,
// partial tiles added with red.global.add.f32)
// ---------------------------------------------------------------------------------------------------
constexpr int WBK = 64, WBN = 64, WBM = 16;

__global__ void __launch_bounds__(256)
wgrad_kernel(const float* __restrict__ A, int lda, int K, const float* __restrict__ Z, int ldz, int N,
             int64_t M, int64_t slab, float* __restrict__ dW, int ldw) {
  __shared__ __align__(16) float As[WBM][WBK];
  __shared__ __align__(16) float Zs[WBM][WBN];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int k0 = blockIdx.x * WBK, n0 = blockIdx.y * WBN;
  const int64_t mb = (int64_t)blockIdx.z * slab, me = min(mb + slab, M);
  const int l_m = tid >> 4, l_c0 = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t m = mb; m < me; m += WBM) {
    const int64_t row = m + l_m;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + l_c0 + q, n = n0 + l_c0 + q;
      As[l_m][l_c0 + q] = (row < me && k < K) ? A[row * lda + k] : 0.f;
      Zs[l_m][l_c0 + q] = (row < me && n < N) ? Z[row * ldz + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < WBM; ++mm) {
      const float4 a = *reinterpret_cast<const float4*>(&As[mm][ty * 4]);
      const float4 z = *reinterpret_cast<const float4*>(&Zs[mm][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, zv[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], zv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) atomicAdd(&dW[(int64_t)k * ldw + n], acc[i][j]);
    }
  }
}

// db != nullptr: the caller also wants db[N] += column sums of Z (the layer's bias gradient); *db_done tells whether
// this launch took them along (the tensor-core kernel has every Z value in registers anyway) or a colsum launch is due
static int launch_wgrad(const float* A, int lda, int K, const float* Z, int ldz, int N, int64_t M, float* dW,
                        int ldw, cudaStream_t st, bool tc = false, float* db = nullptr, bool* db_done = nullptr) {
  if (M == 0 || K == 0 || N == 0) return KNERF_OK;
  if (N <= 4) {
    const int kb = (int)cdiv(K, 256);
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>(cdiv(M, 64), cdiv((int64_t)kNumSMs * 8, kb)));
    const int64_t slab = cdiv(M, splits);
    splits = cdiv(M, slab);
    skinny_wgrad_kernel<<<dim3((unsigned)kb, (unsigned)splits), 256, 0, st>>>(A, lda, K, Z, ldz, N, M, slab, dW, ldw);
    KN_LAUNCH_CHECK();
    return KNERF_OK;
  }
  if (tc && tcx_wgrad_eligible(K, N)) {
    const bool fuse = db != nullptr && db_done != nullptr && !*db_done;
    if (fuse) *db_done = true;
    return launch_wgrad_tc(A, lda, K, Z, ldz, N, M, dW, ldw, fuse ? db : nullptr, st);
  }
  const int tiles = (int)(cdiv(K, WBK) * cdiv(N, WBN));
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>(cdiv(M, 4 * WBM), cdiv((int64_t)kNumSMs * 8, tiles)));
  int64_t slab = align_up(cdiv(M, splits), WBM);
  splits = cdiv(M, slab);
  dim3 grid((unsigned)cdiv(K, WBK), (unsigned)cdiv(N, WBN), (unsigned)splits);
  wgrad_kernel<<<grid, 256, 0, st>>>(A, lda, K, Z, ldz, N, M, slab, dW, ldw);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

// db[N] += column sums of Z[M,N]
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ Z, int ldz, int N, int64_t M, int64_t slab, float* __restrict__ db) {
  __shared__ float red[8][33];
  const int c = threadIdx.x & 31, rr = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + c;
  const int64_t mb = (int64_t)blockIdx.y * slab, me = min(mb + slab, M);
  float s = 0.f;
  if (col < N)
    for (int64_t m = mb + rr; m < me; m += 8) s += Z[m * ldz + col];
  red[rr][c] = s;
  __syncthreads();
  if (rr == 0 && col < N) {
    float tsum = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) tsum += red[q][c];
    atomicAdd(&db[col], tsum);
  }
}

static int launch_colsum(const float* Z, int ldz, int N, int64_t M, float* db, cudaStream_t st) {
  if (M == 0 || N == 0) return KNERF_OK;
  const int cb = (int)cdiv(N, 32);
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>(cdiv(M, 256), cdiv((int64_t)kNumSMs * 4, cb)));
  const int64_t slab = cdiv(M, splits);
  splits = cdiv(M, slab);
  dim3 grid((unsigned)cb, (unsigned)splits);
  colsum_kernel<<<grid, 256, 0, st>>>(Z, ldz, N, M, slab, db);
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

// ---------------------------------------------------------------------------------------------------
// workspace plan
// ---------------------------------------------------------------------------------------------------
Fp32Plan make_fp32_plan(const Model& m, int64_t rows, bool training, bool tc) {
  Fp32Plan p{};
  p.rows = rows;
  p.tc = tc;
  p.ldx = (m.dx + 3) & ~3;
  p.ldd = (m.dd + 3) & ~3;
  size_t off = 0;
  auto take = [&](size_t floats) { size_t o = off; off += (floats * 4 + 255) & ~(size_t)255; return o; };
  p.off_x0 = take((size_t)rows * p.ldx);
  p.off_dir = take((size_t)rows * p.ldd);
  const int nbuf = training ? m.n_layers : std::min(2, m.n_layers);
  size_t hb[kMaxLayers];
  for (int i = 0; i < nbuf; ++i) hb[i] = take((size_t)rows * m.U);
  for (int i = 0; i < m.n_layers; ++i) p.off_h[i] = training ? hb[i] : hb[i & 1];
  p.off_f = take((size_t)rows * m.U);
  p.off_g = take((size_t)rows * (m.U / 2));
  if (training) {
    p.off_d0 = take((size_t)rows * m.U);
    p.off_d1 = take((size_t)rows * m.U);
    p.off_dg = take((size_t)rows * (m.U / 2));
  }
  if (tc) {
    for (int i = 0; i < m.n_dense; ++i) {
      const LayerDesc& L = m.L[i];
      p.off_fwd1[i] = take(tcx_blob_bytes(L.fan_out, L.k_h) / 4 + 1);
      p.off_fwd2[i] = take(tcx_blob_bytes(L.fan_out, L.k_x) / 4 + 1);
      p.off_bwd[i] = training ? take(tcx_blob_bytes(L.k_h, L.fan_out) / 4 + 1) : 0;
    }
  }
  p.total = off;
  return p;
}

// ---------------------------------------------------------------------------------------------------
// forward: hidden stack + heads.  X0/DIR are the encoded inputs (any leading dims).
// out_rgb/out_sigma: ld 3/1 (separate, NeRFMLP.call) or both into a packed [rows,4] buffer (ld 4).
// ---------------------------------------------------------------------------------------------------
static bool tc_width(int n) { return n >= 64 && n <= 256 && n % 64 == 0; }
// operand blobs of layer i (nullptr: SIMT)
static const void* fwd_blob1(const Model& m, const Fp32Plan& p, const char* ws, int i) {
  return (p.tc && tc_width(m.L[i].fan_out) && m.L[i].k_h > 0) ? ws + p.off_fwd1[i] : nullptr;
}
static const void* fwd_blob2(const Model& m, const Fp32Plan& p, const char* ws, int i) {
  return (p.tc && tc_width(m.L[i].fan_out) && m.L[i].k_x > 0) ? ws + p.off_fwd2[i] : nullptr;
}
static const void* bwd_blob(const Model& m, const Fp32Plan& p, const char* ws, int i) {
  return (p.tc && tc_width(m.L[i].k_h)) ? ws + p.off_bwd[i] : nullptr;
}

// KNERF_FP32_TC: split every kernel into its bf16 x 3 operand blobs (forward form; dgrad form too when training)
static int tc_pack_all(const Model& m, const float* params, char* ws, const Fp32Plan& p, bool training, cudaStream_t st) {
  for (int i = 0; i < m.n_dense; ++i) {
    const LayerDesc& L = m.L[i];
    const float* W = params + L.w_off;
    if (fwd_blob1(m, p, ws, i)) KN_TRY(tcx_pack(W, L.fan_out, L.fan_out, L.k_h, false, ws + p.off_fwd1[i], st));
    if (fwd_blob2(m, p, ws, i))
      KN_TRY(tcx_pack(W + (int64_t)L.k_h * L.fan_out, L.fan_out, L.fan_out, L.k_x, false, ws + p.off_fwd2[i], st));
    if (training && bwd_blob(m, p, ws, i)) KN_TRY(tcx_pack(W, L.fan_out, L.k_h, L.fan_out, true, ws + p.off_bwd[i], st));
  }
  return KNERF_OK;
}

int fp32_forward_core(const Model& m, const float* params, const float* X0, int ldx, const float* DIR, int ldd,
                      int64_t rows, char* ws, const Fp32Plan& p, float* out_rgb, int ld_rgb, float* out_sigma,
                      int ld_sigma, cudaStream_t st, bool training) {
  const int n = m.n_layers, U = m.U;
  if (p.tc) KN_TRY(tc_pack_all(m, params, ws, p, training, st));
  auto H = [&](int i) { return reinterpret_cast<float*>(ws + p.off_h[i]); };
  float* F = reinterpret_cast<float*>(ws + p.off_f);
  float* G = reinterpret_cast<float*>(ws + p.off_g);
  for (int i = 0; i < n; ++i) {
    const LayerDesc& L = m.L[i];
    GemmArgs g{};
    g.A1 = (i == 0) ? nullptr : H(i - 1); g.lda1 = U; g.K1 = L.k_h; g.B1 = params + L.w_off; g.ldb1 = L.fan_out;
    g.A2 = X0; g.lda2 = ldx; g.K2 = L.k_x; g.B2 = params + L.w_off + (int64_t)L.k_h * L.fan_out; g.ldb2 = L.fan_out;
    g.bias = params + L.b_off; g.C = H(i); g.ldc = U; g.M = rows; g.N = L.fan_out; g.epi = EPI_RELU;   // mlp.py:33-34
    KN_TRY(launch_gemm(g, false, st, fwd_blob1(m, p, ws, i), fwd_blob2(m, p, ws, i)));
  }
  const float* Hl = H(n - 1);
  {  // sigma = relu(h @ Ws + bs)   (mlp.py:40)
    const LayerDesc& L = m.L[n];
    GemmArgs g{};
    g.A1 = Hl; g.lda1 = U; g.K1 = L.k_h; g.B1 = params + L.w_off; g.ldb1 = 1;
    g.A2 = X0; g.lda2 = ldx; g.K2 = L.k_x; g.B2 = params + L.w_off + L.k_h; g.ldb2 = 1;
    g.bias = params + L.b_off; g.C = out_sigma; g.ldc = ld_sigma; g.M = rows; g.N = 1; g.epi = EPI_RELU;
    KN_TRY(launch_gemm(g, false, st));
  }
  {  // features = h @ Wf + bf   (linear, mlp.py:42)
    const LayerDesc& L = m.L[n + 1];
    GemmArgs g{};
    g.A1 = Hl; g.lda1 = U; g.K1 = L.k_h; g.B1 = params + L.w_off; g.ldb1 = L.fan_out;
    g.A2 = X0; g.lda2 = ldx; g.K2 = L.k_x; g.B2 = params + L.w_off + (int64_t)L.k_h * L.fan_out; g.ldb2 = L.fan_out;
    g.bias = params + L.b_off; g.C = F; g.ldc = U; g.M = rows; g.N = L.fan_out; g.epi = EPI_NONE;
    KN_TRY(launch_gemm(g, false, st, fwd_blob1(m, p, ws, n + 1), fwd_blob2(m, p, ws, n + 1)));
  }
  {  // rgb_features = [features, dir] @ Wg + bg   (linear, no activation: mlp.py:43-46)
    const LayerDesc& L = m.L[n + 2];
    GemmArgs g{};
    g.A1 = F; g.lda1 = U; g.K1 = L.k_h; g.B1 = params + L.w_off; g.ldb1 = L.fan_out;
    g.A2 = DIR; g.lda2 = ldd; g.K2 = L.k_x; g.B2 = params + L.w_off + (int64_t)L.k_h * L.fan_out; g.ldb2 = L.fan_out;
    g.bias = params + L.b_off; g.C = G; g.ldc = U / 2; g.M = rows; g.N = L.fan_out; g.epi = EPI_NONE;
    KN_TRY(launch_gemm(g, false, st, fwd_blob1(m, p, ws, n + 2), fwd_blob2(m, p, ws, n + 2)));
  }
  {  // rgb = sigmoid(g @ Wc + bc)   (mlp.py:48)
    const LayerDesc& L = m.L[n + 3];
    GemmArgs g{};
    g.A1 = G; g.lda1 = U / 2; g.K1 = L.k_h; g.B1 = params + L.w_off; g.ldb1 = 3;
    g.K2 = 0;
    g.bias = params + L.b_off; g.C = out_rgb; g.ldc = ld_rgb; g.M = rows; g.N = 3; g.epi = EPI_SIGMOID;
    KN_TRY(launch_gemm(g, false, st));
  }
  return KNERF_OK;
}

// ---------------------------------------------------------------------------------------------------
// backward: d_pre[rows,4] = (d rgb_pre[3], d sigma_pre) ; grads += dL/dtheta (flat Keras order)
// ---------------------------------------------------------------------------------------------------
int fp32_backward_core(const Model& m, const float* params, const float* X0, int ldx, const float* DIR, int ldd,
                       const float* d_pre, int64_t rows, char* ws, const Fp32Plan& p, float* grads,
                       cudaStream_t st) {
  const int n = m.n_layers, U = m.U, U2 = m.U / 2;
  auto H = [&](int i) { return reinterpret_cast<float*>(ws + p.off_h[i]); };
  const float* F = reinterpret_cast<float*>(ws + p.off_f);
  const float* G = reinterpret_cast<float*>(ws + p.off_g);
  float* D0 = reinterpret_cast<float*>(ws + p.off_d0);
  float* D1 = reinterpret_cast<float*>(ws + p.off_d1);
  float* DG = reinterpret_cast<float*>(ws + p.off_dg);
  const float* Hl = H(n - 1);
  const LayerDesc &Ls = m.L[n], &Lf = m.L[n + 1], &Lg = m.L[n + 2], &Lc = m.L[n + 3];

  // rgb layer: dWc += G^T d_rgbpre ; dbc ; dG = d_rgbpre @ Wc^T
  KN_TRY(launch_wgrad(G, U2, U2, d_pre, 4, 3, rows, grads + Lc.w_off, 3, st, p.tc));
  KN_TRY(launch_colsum(d_pre, 4, 3, rows, grads + Lc.b_off, st));
  {
    GemmArgs g{};
    g.A1 = d_pre; g.lda1 = 4; g.K1 = 3; g.B1 = params + Lc.w_off; g.ldb1 = 3; g.K2 = 0;
    g.C = DG; g.ldc = U2; g.M = rows; g.N = U2; g.epi = EPI_NONE;
    KN_TRY(launch_gemm(g, true, st));
  }
  // rgb_features (linear): dWg[:U] += F^T dG ; dWg[U:] += DIR^T dG ; dbg ; dF = dG @ Wg[:U]^T
  {
    bool done = false;
    KN_TRY(launch_wgrad(F, U, Lg.k_h, DG, U2, U2, rows, grads + Lg.w_off, U2, st, p.tc, grads + Lg.b_off, &done));
    KN_TRY(launch_wgrad(DIR, ldd, Lg.k_x, DG, U2, U2, rows, grads + Lg.w_off + (int64_t)Lg.k_h * U2, U2, st, p.tc,
                        grads + Lg.b_off, &done));
    if (!done) KN_TRY(launch_colsum(DG, U2, U2, rows, grads + Lg.b_off, st));
  }
  {
    GemmArgs g{};
    g.A1 = DG; g.lda1 = U2; g.K1 = U2; g.B1 = params + Lg.w_off; g.ldb1 = U2; g.K2 = 0;
    g.C = D0; g.ldc = U; g.M = rows; g.N = U; g.epi = EPI_NONE;
    KN_TRY(launch_gemm(g, true, st, bwd_blob(m, p, ws, n + 2), nullptr));
  }
  // features (linear) + sigma head share the input h_last (and x after a trailing skip)
  {
    bool done = false;
    KN_TRY(launch_wgrad(Hl, U, Lf.k_h, D0, U, U, rows, grads + Lf.w_off, U, st, p.tc, grads + Lf.b_off, &done));
    KN_TRY(launch_wgrad(X0, ldx, Lf.k_x, D0, U, U, rows, grads + Lf.w_off + (int64_t)Lf.k_h * U, U, st, p.tc,
                        grads + Lf.b_off, &done));
    if (!done) KN_TRY(launch_colsum(D0, U, U, rows, grads + Lf.b_off, st));
  }
  KN_TRY(launch_wgrad(Hl, U, Ls.k_h, d_pre + 3, 4, 1, rows, grads + Ls.w_off, 1, st, p.tc));
  KN_TRY(launch_wgrad(X0, ldx, Ls.k_x, d_pre + 3, 4, 1, rows, grads + Ls.w_off + Ls.k_h, 1, st, p.tc));
  KN_TRY(launch_colsum(d_pre + 3, 4, 1, rows, grads + Ls.b_off, st));
  {  // dZ_{n-1} = (dF @ Wf[:U]^T + d_sigmapre @ Ws[:U]^T) * (h_last > 0)
    GemmArgs g{};
    g.A1 = D0; g.lda1 = U; g.K1 = U; g.B1 = params + Lf.w_off; g.ldb1 = U;
    g.A2 = d_pre + 3; g.lda2 = 4; g.K2 = 1; g.B2 = params + Ls.w_off; g.ldb2 = 1;
    g.C = D1; g.ldc = U; g.M = rows; g.N = U; g.epi = EPI_MASK; g.mask = Hl; g.ldmask = U;
    KN_TRY(launch_gemm(g, true, st, bwd_blob(m, p, ws, n + 1), bwd_blob(m, p, ws, n)));
  }
  float* dz = D1;
  float* other = D0;
  for (int i = n - 1; i >= 0; --i) {
    const LayerDesc& L = m.L[i];
    bool done = false;
    if (i > 0)
      KN_TRY(launch_wgrad(H(i - 1), U, L.k_h, dz, U, U, rows, grads + L.w_off, U, st, p.tc, grads + L.b_off, &done));
    KN_TRY(launch_wgrad(X0, ldx, L.k_x, dz, U, U, rows, grads + L.w_off + (int64_t)L.k_h * U, U, st, p.tc,
                        grads + L.b_off, &done));
    if (!done) KN_TRY(launch_colsum(dz, U, U, rows, grads + L.b_off, st));
    if (i > 0) {  // dZ_{i-1} = (dZ_i @ W_i[:U]^T) * (h_{i-1} > 0)   (no gradient into the x / dir inputs)
      GemmArgs g{};
      g.A1 = dz; g.lda1 = U; g.K1 = U; g.B1 = params + L.w_off; g.ldb1 = U; g.K2 = 0;
      g.C = other; g.ldc = U; g.M = rows; g.N = L.k_h; g.epi = EPI_MASK; g.mask = H(i - 1); g.ldmask = U;
      KN_TRY(launch_gemm(g, true, st, bwd_blob(m, p, ws, i), nullptr));
      std::swap(dz, other);
    }
  }
  return KNERF_OK;
}

}  // namespace knerf

using namespace knerf;

extern "C" int knerf_mlp_forward_encoded(const knerf_config* cfg, const float* params, const float* xyz,
                                         int ld_xyz, const float* dirs, int ld_dir, int64_t rows, float* rgb,
                                         float* sigma, void* workspace, int64_t workspace_bytes, void* stream) {
  KN_CHECK_ARG(cfg && params && xyz && dirs && rgb && sigma && rows >= 0, "knerf_mlp_forward_encoded: null argument");
  Model m;
  KN_TRY(build_model(cfg, &m));
  KN_CHECK_ARG(ld_xyz >= m.dx && ld_dir >= m.dd, "knerf_mlp_forward_encoded: leading dims smaller than dx/dd");
  const Fp32Plan p = make_fp32_plan(m, rows, false);
  if ((int64_t)p.total > workspace_bytes || workspace == nullptr)
    return fail(KNERF_ERR_WORKSPACE, "knerf_mlp_forward_encoded: workspace %lld < %lld bytes",
                (long long)workspace_bytes, (long long)p.total);
  return fp32_forward_core(m, params, xyz, ld_xyz, dirs, ld_dir, rows, (char*)workspace, p, rgb, 3, sigma, 1,
                           (cudaStream_t)stream);
}
