// a11 (KNERF_BF16 mode): backward of the NeRF MLP on tcgen05 tensor cores.
//
//  1. tc_mlp_dgrad_kernel -- the forward's chain structure run backwards (same CTA-pair roles): starting from the
//     gradient w.r.t. the head pre-activations (from the fused compositing backward) it goes from d(rgb_features)
//     straight to d(h7) through W'^T (tc_layout.cuh: `features` is folded away), then layer 7 .. layer 1, each step
//     one [128 x 256] x [256 x 256] GEMM followed by a ReLU' epilogue (1-bit masks written by the forward).
//     Every pre-activation gradient tile dZ_l is written once to HBM in chunk-major bf16 for step 2.
//     No gradient flows into the xyz / direction encodings (mlp.py inputs are constants: nerf.py:361-369).
//  2. tc_wgrad_kernel -- dW_l = X_l^T dZ_l with the reduction over SAMPLES: the saved activation / gradient
//     blobs are consumed as MN-major UMMA operands (same bytes, other axis; see tc_ptx.cuh), fp32 partial
//     sums stay in TMEM across all tiles of a work item and are flushed once with red.global.add.f32; idle warps
//     take the bias gradients (column sums of the dZ units in shared memory).
//  3. tc_finish_kernel -- every gradient that contains d(rgb_features) = d(rgb_pre) Wc^T (rank 3): features,
//     rgb_features and rgb kernels / biases from Y = h7^T d(rgb_pre), Yd = PE(dir)^T d(rgb_pre), sum d(rgb_pre).
#include "mlp_tc.cuh"

#include <cstdlib>
#include "tc_roles2.cuh"

namespace knerf {
using namespace tc;
using namespace tcl;

TcParams tc_make_params(const Model& m);

namespace {

// =============================================================================================================
// dgrad chain
// =============================================================================================================
// epilogue of one dgrad step: accumulator (dH = dZ_next W^T) -> pre-activation gradient tile, bf16.
// KIND 1: dZ7 = (dG W'^T + d(sigma_pre) Ws^T) * ReLU'(h7) -- the sigma head hangs off h7 too;
// 2: dZ_l = dH * ReLU'(h_l), l = 6..1;  3: dZ0, the last step: feeds no further GEMM, so it goes to HBM directly
// (dst = record or nullptr) and must NOT touch hs[tl], which the other half-row thread of this row may already be
// rebuilding for the next tile.  KIND < 3: dst = hs[tl], the next A operand and (via warp 10) the dZ record.
// mb = this thread's ReLU' bits (4 x 32 columns; bit 8 t + q = column 4 q + t of the group).
// REC8 (fp8 records): the dZ record leaves from here as e5m2(dZ * scale), 16 columns per 16-byte vector (rec8, nullptr
// for a padding tile); dst is only the next A operand then (unused for KIND 3).
template <int KIND, bool REC8>
__device__ __noinline__ void epi_dgrad(uint32_t tacc, uint8_t* __restrict__ dst, int h, int r, uint4 mbv, float dsig,
                                       const float* __restrict__ wsig, uint8_t* __restrict__ rec8) {
  const uint32_t mb[4] = {mbv.x, mbv.y, mbv.z, mbv.w};   // by value: no local memory (no L1 on this SM)
#pragma unroll
  for (int gI = 0; gI < 4; ++gI) {
    const int col0 = h * 128 + gI * 32;
    uint32_t v[32];
    tmem_ld32_issue(tacc + col0, v);
    float4 ws[8];
    if (KIND == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) ws[i] = *(reinterpret_cast<const float4*>(wsig + col0) + i);   // shared memory
    }
    tmem_ld32_wait(v);
    uint32_t q8[8];
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      float x[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[c8 * 8 + e]);
      if (KIND == 1) {
        const float4 w0 = ws[2 * c8], w1 = ws[2 * c8 + 1];
        x[0] += dsig * w0.x; x[1] += dsig * w0.y; x[2] += dsig * w0.z; x[3] += dsig * w0.w;
        x[4] += dsig * w1.x; x[5] += dsig * w1.y; x[6] += dsig * w1.z; x[7] += dsig * w1.w;
      }
      uint4 pk = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                            pack_bf16x2(x[6], x[7]));
      // ReLU' of columns 4q .. 4q+3 (q = 2 c8, 2 c8 + 1): bit q of bytes 0..3 of mb (tc_layout.cuh kRecMask).  Shifted left
      // by 7 - q they are the four byte msbs, and ONE prmt with sign replication expands two of them to a half-word
      // mask (the shift is an IMAD on the FMA pipe; the epilogue is bound by the half-rate ALU pipe)
      const uint32_t ta = mb[gI] << (7 - 2 * c8), tb = mb[gI] << (6 - 2 * c8);
      pk.x &= prmt(ta, 0u, 0x9988u); pk.y &= prmt(ta, 0u, 0xBBAAu);
      pk.z &= prmt(tb, 0u, 0x9988u); pk.w &= prmt(tb, 0u, 0xBBAAu);
      if (REC8) {   // byte masks of columns (0,1,2,3) / (4,5,6,7): the same flag bits of two words.  (The chain runs on
                    // gradients that already carry the record scale -- see the prologue -- so there is no multiply here.)
        q8[2 * c8] = pack_e5m2x4(x[0], x[1], x[2], x[3]) & prmt(ta, 0u, 0xBA98u);
        q8[2 * c8 + 1] = pack_e5m2x4(x[4], x[5], x[6], x[7]) & prmt(tb, 0u, 0xBA98u);
        if (KIND < 3) *reinterpret_cast<uint4*>(dst + ((col0 >> 3) + c8) * kChunkA + r * 16) = pk;
      } else {
        if (KIND < 3 || dst != nullptr) *reinterpret_cast<uint4*>(dst + ((col0 >> 3) + c8) * kChunkA + r * 16) = pk;
      }
    }
    if (REC8 && rec8 != nullptr) {
      __stcs(reinterpret_cast<uint4*>(rec8 + ((col0 >> 4) + 0) * kChunkA + r * 16), make_uint4(q8[0], q8[1], q8[2], q8[3]));
      __stcs(reinterpret_cast<uint4*>(rec8 + ((col0 >> 4) + 1) * kChunkA + r * 16), make_uint4(q8[4], q8[5], q8[6], q8[7]));
    }
  }
}

// fp8 records: max|d_pre| of the call (the bit pattern of a non-negative float orders like the float)
__global__ void __launch_bounds__(256) dpre_amax_kernel(const float4* __restrict__ d_pre, int64_t M,
                                                        uint32_t* __restrict__ amax_bits) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(d_pre + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
}

// clusters of 2, cta_group::2 MMAs (tc_roles2.cuh); a work unit is four tiles (two per CTA)
template <bool REC8>
__global__ void __launch_bounds__(kThreads, 1)
tc_mlp_dgrad_kernel(const uint8_t* __restrict__ packed, const float4* __restrict__ d_pre, int64_t M,
                    const uint8_t* __restrict__ rec, uint8_t* __restrict__ dz, float* __restrict__ grads, TcParams P,
                    float* __restrict__ xbuf) {
  constexpr int kRB = REC8 ? kRec8Bytes : kRecBytes, kOffMask = REC8 ? kRec8Mask : kRecMask,
                kDB = REC8 ? kDz8Bytes : kDzBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  ChainSmem& sm = *reinterpret_cast<ChainSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (M + kTileM - 1) / kTileM;
  const int64_t n_pairs = (n_tiles + 3) / 4;   // work units: four tiles per CTA pair
  const uint32_t cta = cluster_ctarank();
  const int64_t first = blockIdx.x >> 1, stride = gridDim.x >> 1;
  const uint32_t tmem = chain2_setup(sm, tid, warp, cta);
  auto tile_of = [&](int64_t unit, int tl) -> int64_t { return unit * 4 + tl * 2 + (int64_t)cta; };

  if (warp == 0) {
    if (lane == 0) producer2_role<BwdProg>(sm, packed + kBwdPairOff, cta, n_pairs, first, stride);
  } else if (warp == 1) {
    if (lane == 0 && cta == 0) mma2_role<BwdProg>(sm, tmem, 0u, true, n_pairs, first, stride);
    else if (lane == 0) relay_role<BwdProg>(sm, n_pairs, first, stride);
  } else if (warp == 11) {
    if (lane == 0 && cta == 0) mma2_role<BwdProg>(sm, tmem, 1u, true, n_pairs, first, stride);
  } else if (warp == 10) {
    // record store: the A operands dZ7..dZ1 of the chain (items 1..7) are also the dZ records of the weight-gradient
    // kernel (tc_roles.cuh store_role); item 0 = dG only takes part in the hand-shake: nothing downstream reads it
    // from HBM (its weight gradients factor through d(rgb_pre), see build_task_table)
    if (!REC8 && lane == 0) {   // (fp8 records leave from the epilogue's registers)
      store_role(sm, 8, n_tiles, n_pairs, first, stride, tile_of,
                 [&](int item, int64_t tile) {
                   return dz + tile * kDzBytes + (item == 0 ? kDzG : kDzZ0 + (8 - item) * kHSBytes);
                 },
                 [](int item) { return (uint32_t)(item == 0 ? 0 : kHSBytes); });   // dG stays on chip (item 0)
    }
  } else {
    uint32_t st_pending = 0, st_par = 0;      // bit tl = hs[tl] is being stored / parity of st_done[tl]
    auto a_ready_arrive = [&](int tl) {   // this thread's (warp's) part of the next A operand (= dZ record) is in smem
      a_ready_arrive2(sm, tl, lane);
      if (!REC8) {
        st_ready_arrive(&sm.st_ready[tl], lane);
        st_pending |= 1u << tl;
      }
    };
    auto hs_writable = [&](int tl) {          // the bulk store of the previous contents of hs[tl] has read them
      if ((st_pending >> tl) & 1) {
        mbar_wait(&sm.st_done[tl], (st_par >> tl) & 1);
        st_par ^= 1u << tl;
        st_pending &= ~(1u << tl);
      }
    };
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    // The fp32 kernels of the two CUDA-core heads (sigma [256], rgb [128,3]) live in shared memory for the whole
    // kernel -- the dgrad chain has no encoding operand, so xs[] is free.  With 227 KB of shared memory the SM has
    // no L1 left: reading them with __ldg cost an L2 round trip each (224 per thread and tile).
    const float* aux = reinterpret_cast<const float*>(packed + kAuxOff);
    float* head = reinterpret_cast<float*>(sm.xs[0]);
    for (int i = tid - 64; i < 160; i += kComputeThreads)
      *(reinterpret_cast<float4*>(head) + i) = __ldg(reinterpret_cast<const float4*>(aux + 12 * 256) + i);
    named_bar_sync(1, kComputeThreads);
    const float* wsig = head;
    const float* wrgb = head + 256;
    uint32_t acc_par[2] = {0, 0};
    float dsig_keep0 = 0.f, dsig_keep1 = 0.f;
    // fp8 records: one power-of-two scale per call, from max|d_pre| (written by dpre_amax_kernel before this launch)
    const float scale = REC8 ? dz8_scale(__ldg(reinterpret_cast<const uint32_t*>(xbuf) + kXOffAmax), false) : 1.f;

    // tile start: dG = d(rgb_pre) Wc^T (K = 3, CUDA cores) becomes the first A operand
    auto prologue = [&](int64_t pair, int tl) {
      const int64_t tile = tile_of(pair, tl);
      const int64_t g = tile * kTileM + r;
      const bool active = tile < n_tiles;
      float4 dp = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < M) dp = d_pre[g];
      // fp8 records: the whole chain runs on gradients times the record scale (a power of two: the same mantissas, so
      // every dZ is the same number times 2^k), and the e5m2 records need no multiply in the epilogues
      const float4 dq = REC8 ? make_float4(dp.x * scale, dp.y * scale, dp.z * scale, dp.w * scale) : dp;
      if (tl == 0) dsig_keep0 = dq.w; else dsig_keep1 = dq.w;
      uint8_t* dz_t = dz + tile * kDB;
      if (active && !REC8) {
        const uint4 pk = (h == 0) ? make_uint4(pack_bf16x2(dp.x, dp.y), pack_bf16x2(dp.z, dp.w), 0u, 0u)
                                  : make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(dz_t + kDzP + h * kChunkA + r * 16) = pk;
      }
      if (active && REC8 && h == 0)   // one 16-column chunk: (d rgb_pre, d sigma_pre) * scale, then zeros
        *reinterpret_cast<uint4*>(dz_t + kDz8P + r * 16) =
            make_uint4(pack_e5m2x4(dq.x, dq.y, dq.z, dq.w), 0u, 0u, 0u);
      if (h == 0) {   // bias gradients of the two heads: db_rgb = sum d(rgb_pre), db_sigma = sum d(sigma_pre)
        const float s0 = warp_sum(dp.x), s1 = warp_sum(dp.y), s2 = warp_sum(dp.z), s3 = warp_sum(dp.w);
        if (lane == 0) {
          atomicAdd(grads + P.b_off[11], s0); atomicAdd(grads + P.b_off[11] + 1, s1);
          atomicAdd(grads + P.b_off[11] + 2, s2); atomicAdd(grads + P.b_off[8], s3);
          if (REC8) {   // sum d(rgb_pre) of THIS call for tc_finish_kernel (the bf16 path sums its operand copy instead)
            atomicAdd(xbuf + kXOffS, s0); atomicAdd(xbuf + kXOffS + 1, s1); atomicAdd(xbuf + kXOffS + 2, s2);
          }
        }
      }
      hs_writable(tl);
#pragma unroll 2
      for (int c8 = 0; c8 < 8; ++c8) {
        const int col = h * 64 + c8 * 8;
        float x[8];
        // rgb kernel rows col..col+7: 24 consecutive floats, 16-byte aligned (shared memory, broadcast reads)
        const float4* w4 = reinterpret_cast<const float4*>(wrgb + col * 3);
        const float4 a0 = w4[0], a1 = w4[1], a2 = w4[2], a3 = w4[3], a4 = w4[4], a5 = w4[5];
        x[0] = dq.x * a0.x + dq.y * a0.y + dq.z * a0.z;
        x[1] = dq.x * a0.w + dq.y * a1.x + dq.z * a1.y;
        x[2] = dq.x * a1.z + dq.y * a1.w + dq.z * a2.x;
        x[3] = dq.x * a2.y + dq.y * a2.z + dq.z * a2.w;
        x[4] = dq.x * a3.x + dq.y * a3.y + dq.z * a3.z;
        x[5] = dq.x * a3.w + dq.y * a4.x + dq.z * a4.y;
        x[6] = dq.x * a4.z + dq.y * a4.w + dq.z * a5.x;
        x[7] = dq.x * a5.y + dq.y * a5.z + dq.z * a5.w;
        const uint4 pk = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                                    pack_bf16x2(x[6], x[7]));
        const int off = (col >> 3) * kChunkA + r * 16;
        *reinterpret_cast<uint4*>(sm.hs[tl] + off) = pk;   // first A operand and the dG record (warp 10 stores it)
      }
      a_ready_arrive(tl);
    };

    if (first < n_pairs) {
#pragma unroll 1
      for (int tl = 0; tl < 2; ++tl) prologue(first, tl);
    }
    for (int64_t pair = first; pair < n_pairs; pair += stride) {
      for (int b = 0; b < BwdProg::kSteps; ++b) {
        const int zi = 7 - b;                                  // index of the dZ this step produces
#pragma unroll 1
        for (int tl = 0; tl < 2; ++tl) {
          const int64_t tile = tile_of(pair, tl);
          const bool active = tile < n_tiles;
          uint8_t* out = REC8 ? dz + tile * kDB + kDz8Z0 + zi * kH8Bytes : dz + tile * kDB + kDzZ0 + zi * kHSBytes;
          const uint8_t* mask = rec + tile * kRB + kOffMask + zi * kMaskLayerBytes;
          const float dsig = tl == 0 ? dsig_keep0 : dsig_keep1;
          // the ReLU' bits (written by the forward, 1 bit per activation) do not depend on the accumulator: fetch
          // this thread's 4 x 32 columns BEFORE waiting for the MMA so that the latency overlaps it
          uint32_t mb[4] = {0u, 0u, 0u, 0u};
          if (active) {
            const uint4 mv = __ldg(reinterpret_cast<const uint4*>(mask + h * 2048 + r * 16));
            mb[0] = mv.x; mb[1] = mv.y; mb[2] = mv.z; mb[3] = mv.w;
          }
          mbar_wait_cluster(&sm.acc_ready[tl], acc_par[tl]);
          acc_par[tl] ^= 1;
          tc_fence_after();
          if (b + 1 < BwdProg::kSteps) hs_writable(tl);
          // separately instantiated bodies (a merged loop gets if-converted: the d(sigma) term and its loads
          // would run on every step)
          const uint32_t tacc = tmem + lane_base + tl * 256;
          const uint4 mbv = make_uint4(mb[0], mb[1], mb[2], mb[3]);
          uint8_t* out8 = (REC8 && active) ? out : nullptr;
          if (b == 0) epi_dgrad<1, REC8>(tacc, sm.hs[tl], h, r, mbv, dsig, wsig, out8);
          else if (b + 1 < BwdProg::kSteps) epi_dgrad<2, REC8>(tacc, sm.hs[tl], h, r, mbv, 0.f, nullptr, out8);
          else epi_dgrad<3, REC8>(tacc, active ? out : nullptr, h, r, mbv, 0.f, nullptr, out8);
          if (b + 1 < BwdProg::kSteps) {
            a_ready_arrive(tl);
          } else {
            tc_fence_before();
            // (bias gradients = column sums of the dZ tiles are taken by tc_wgrad_kernel, which has them in smem)
            const int64_t next = pair + stride;
            if (next < n_pairs) prologue(next, tl);
          }
        }
      }
    }
  }
  chain2_teardown(tmem, warp);
}

// =============================================================================================================
// weight gradients
// =============================================================================================================
constexpr int kWUnitBytes = 32768;     // 16 chunks x 128 samples x 16 B: 128 features (A) or 128 outputs (B)
constexpr int kWSlots = 7;            // 7 x 32 KB in flight per CTA (224 KB of the 227 KB)
constexpr int kWThreads = 224;         // warp 0 producer, warps 1 and 6 MMA issuers, warps 2-5 flush
constexpr int kNumTasks = 10;
// Item share of the heads task (a `big` task is 128).  It moves only 76 KB per tile but issues 24 N = 16 MMAs, each of
// which holds its issuing thread as long as a wide one: measured 2.08 / 1.76 / 1.65 / 1.65 / 1.63 ns per sample for
// the whole kernel at 50 / 80 / 110 / 140 / 178; 178 gives 217 + 36 + 43 = 296 items = exactly two per CTA.
constexpr int kHeadsCost = 178;
// fp8 records (build_task_table8): 64 KB and 16 MMAs per tile for a `big` task (cost 128), 40 KB / 8 MMAs for the two
// encoding tasks, 38 KB / 12 narrow MMAs for the heads; 66 slabs -> 462 + 2 x 41 + 48 = 592 items = four per CTA
// (swept on the GPU, profiles/r02_fp8_records.md: the item count must be a multiple of the 148 CTAs -- a partly filled
// extra round costs 4-35 % -- and four per CTA beat two by 1.5-2 %)
#ifndef KNERF_WG8_X   // (scratch builds for the balance measurements pass the three numbers as macros)
#define KNERF_WG8_X 80
#define KNERF_WG8_H 94
#define KNERF_WG8_S 66
#endif
constexpr int kXpartCost8 = KNERF_WG8_X;
constexpr int kHeadsCost8 = KNERF_WG8_H;
constexpr int kSlabs8 = KNERF_WG8_S;
constexpr int kWMaxUnits = 6;

// one ring slot: `bytes` from the forward (src 0) or dz (src 1) record at offset 0 and, optionally, `bytes2` from the
// dz record at slot offset 8192 (two small operands share a slot); bias_layer >= 0: column sums of this dZ unit
// are the bias gradient of that layer
struct WUnit { int src, off, bytes, bias_layer, bias_col0, off2, bytes2; };
// accumulator group: A = unit a (+ a_sub bytes), B = unit b (+ b_sub bytes)
struct WGroup { int a, b, col, N, layer, row_base, row_limit, col_base, mode, free_a, free_b, a_sub, b_sub, issuer; };
// last_grp[x][k]: the last group of issuer x that reads unit k (the one after which x releases the slot), -1 = none
struct WTask { int n_units; WUnit u[kWMaxUnits]; int n_groups; WGroup g[kWMaxUnits]; int cost; int8_t last_grp[2][kWMaxUnits]; };
// mode 0: dW[layer][(row_base+row), col_base+col]
//      1: A = h7, B = the packed d_pre operand: column 3 = the sigma kernel's gradient, columns 0..2 = Y = h7^T d(rgb_pre)
//         into the fp32 scratch
//      4: A = PE(dir): columns 0..2 = Yd = PE(dir)^T d(rgb_pre) into the scratch     (tc_finish_kernel uses Y, Yd)

struct WTaskTable { WTask t[kNumTasks]; };

// Fit the task table of the chain's fixed program to the model embedded in it (api.cu tc_chain_map).  Identity layers
// (w_off < 0) keep their tasks, so that the item layout does not depend on the model, but flush nothing: row_limit 0,
// no bias gradient.  A narrower model (U < 256) flushes only its own rows / columns, and the encoding rows of the
// concat layer start at row U of its kernel.
static void fit_tasks_to_model(WTaskTable& T, const TcParams& P) {
  for (int k = 0; k < kNumTasks; ++k) {
    WTask& t = T.t[k];
    for (int gi = 0; gi < t.n_groups; ++gi) {
      WGroup& g = t.g[gi];
      const int l = g.layer;
      if (g.mode == 0 && l >= 0 && l < 8) {
        if (l == 5 && g.row_base == 256) {            // encoding rows of the concat layer
          g.row_base = P.U;
          if (!P.x5) g.row_limit = 0;
        } else if (l > 0) {                           // hidden rows (layer 0's rows are encoding columns)
          g.row_limit = std::max(0, std::min(g.row_limit, P.U - g.row_base));
        }
        if (P.w_off[l] < 0) g.row_limit = 0;
      } else if (g.mode == 1) {                       // Y = h7^T d(rgb_pre) and the sigma kernel: rows of h7
        g.row_limit = std::max(0, std::min(g.row_limit, P.U - g.row_base));
      }
    }
    for (int u = 0; u < t.n_units; ++u)
      if (t.u[u].bias_layer >= 0 && P.b_off[t.u[u].bias_layer] < 0) t.u[u].bias_layer = -1;
  }
}

static WTaskTable build_task_table(int dx, int dd) {
  WTaskTable T{};
  int n = 0;
  auto big = [&](int layer, int a_off, int b_off, int row_base) {   // A: two 128-feature halves, B: two 128-output halves
    WTask& t = T.t[n++];
    t.n_units = 4;
    t.u[0] = {0, a_off, kWUnitBytes, -1, 0}; t.u[1] = {1, b_off, kWUnitBytes, layer, 0};
    t.u[2] = {1, b_off + kWUnitBytes, kWUnitBytes, layer, 128}; t.u[3] = {0, a_off + kWUnitBytes, kWUnitBytes, -1, 0};
    t.n_groups = 4;
    t.g[0] = {0, 1, 0, 128, layer, row_base, 128, 0, 0, 0, 0};
    t.g[1] = {0, 2, 128, 128, layer, row_base, 128, 128, 0, 1, 0};
    t.g[2] = {3, 1, 256, 128, layer, row_base + 128, 128, 0, 0, 0, 1};
    t.g[3] = {3, 2, 384, 128, layer, row_base + 128, 128, 128, 0, 1, 1};
    t.cost = 128;
  };
  auto xpart = [&](int layer, int b_off, int row_base) {           // A: PE(xyz) (63 valid rows), B: two halves
    WTask& t = T.t[n++];
    t.n_units = 3;
    const int bl = (layer == 0) ? 0 : -1;   // layer 5's dZ column sums are taken by its h-part task
    // PE(xyz) is 64 features = 16 KB: only that much is fetched.  The MMA still reads a 128-feature unit; what the
    // rest of the ring slot holds (stale bytes) only reaches accumulator rows >= 64, which are never flushed.
    t.u[0] = {0, kRecXS, kXSBytes, -1, 0}; t.u[1] = {1, b_off, kWUnitBytes, bl, 0};
    t.u[2] = {1, b_off + kWUnitBytes, kWUnitBytes, bl, 128};
    t.n_groups = 2;
    t.g[0] = {0, 1, 0, 128, layer, row_base, dx, 0, 0, 0, 1};
    t.g[1] = {0, 2, 128, 128, layer, row_base, dx, 128, 0, 1, 1};
    t.cost = 76;    // costs = measured time per tile of a CTA working alone on the task (x 128 / the `big` tasks'):
                    // 3,700 / 6,300 / 8,400 cycles -- the kernel is bound by its CTAs' per-tile time, so equal item
                    // TIMES (not bytes) remove the tail
  };
  xpart(0, kDzZ0, 0);                                                        // layer 0
  for (int l = 1; l <= 7; ++l) big(l, kRecH0 + (l - 1) * kHSBytes, kDzZ0 + l * kHSBytes, 0);   // layers 1..7 (h part)
  xpart(5, kDzZ0 + 5 * kHSBytes, 256);                                       // layer 5, skip rows 256..318
  {  // Everything that hangs off h7 and d(rgb_features).  d(rgb_features) = dG = d(rgb_pre) Wc^T has rank 3, so every
     // gradient that contains it factors through three columns:  h7^T dG = (h7^T d(rgb_pre)) Wc^T,
     // PE(dir)^T dG = (PE(dir)^T d(rgb_pre)) Wc^T,  sum dG = (sum d(rgb_pre)) Wc^T.  The task therefore reads h7, PE(dir)
     // and the packed (d rgb_pre, d sigma_pre) operand only -- N = 16 MMAs -- and leaves Y = h7^T d(rgb_pre) [256 x 3]
     // (+ the sigma kernel's gradient, column 3 of the same accumulator), Yd = PE(dir)^T d(rgb_pre) [27 x 3] and
     // s3 = sum d(rgb_pre) in the scratch; tc_finish_kernel expands them into dW / db of features, rgb_features and
     // rgb.  Neither dG nor the rgb_features activations travel through HBM (64 KB per tile less than round 1).
    WTask& t = T.t[n++];
    const int h7 = kRecH0 + 7 * kHSBytes;
    t.n_units = 3;
    t.u[0] = {0, h7, kWUnitBytes, -1, 0, 0, 0};                 // h7 features 0..127
    t.u[1] = {0, kRecDS, 8192, -1, 0, kDzP, 4096};              // PE(dir) (32 columns, 27 valid accumulator rows) and,
                                                                // at +8192, the packed (d rgb_pre, d sigma_pre) operand
    t.u[2] = {0, h7 + kWUnitBytes, kWUnitBytes, -1, 0, 0, 0};   // h7 features 128..255
    t.n_groups = 3;
    t.g[0] = {0, 1, 0, 16, 8, 0, 128, 0, 1, 0, 0, 0, 8192};     // Y rows 0..127 (columns 0..2), sigma kernel (column 3)
    t.g[1] = {2, 1, 32, 16, 8, 128, 128, 0, 1, 0, 0, 0, 8192};  // rows 128..255
    t.g[2] = {1, 1, 64, 16, -1, 0, dd, 0, 4, 0, 0, 0, 8192};    // Yd: A and B share the unit
    t.g[0].issuer = 0; t.g[1].issuer = 1; t.g[2].issuer = 0;
    t.cost = kHeadsCost;
  }
  // Two MMA-issuing threads (tcgen05.mma holds its issuer for ~160 cycles per N = 128 MMA whose pipe time is 64):
  // every accumulator group belongs to ONE issuer, so the fp32 accumulation order of each gradient element -- and
  // with it every bit -- is that of a single issuer.  Default split: groups alternate.
  for (int k = 0; k < n; ++k) {
    WTask& t = T.t[k];
    if (k != n - 1)
      for (int gi = 0; gi < t.n_groups; ++gi) t.g[gi].issuer = gi & 1;
    for (int x = 0; x < 2; ++x)
      for (int u = 0; u < kWMaxUnits; ++u) {
        t.last_grp[x][u] = -1;
        for (int gi = 0; gi < t.n_groups; ++gi)
          if (t.g[gi].issuer == x && (t.g[gi].a == u || t.g[gi].b == u)) t.last_grp[x][u] = (int8_t)gi;
      }
  }
  return T;
}

// fp8 records: a unit is a WHOLE 32 KB record (256 features x 128 samples); the 128 x 128 accumulator groups address
// its halves through a_sub / b_sub.  64 KB per tile and `big` task instead of 128.
static WTaskTable build_task_table8(int dx, int dd) {
  WTaskTable T{};
  int n = 0;
  constexpr int kHalf = kH8Bytes / 2;
  auto big = [&](int layer, int a_off, int b_off, int row_base) {
    WTask& t = T.t[n++];
    t.n_units = 2;
    t.u[0] = {0, a_off, kH8Bytes, -1, 0, 0, 0}; t.u[1] = {1, b_off, kH8Bytes, layer, 0, 0, 0};
    t.n_groups = 4;
    t.g[0] = {0, 1, 0, 128, layer, row_base, 128, 0, 0, 0, 0, 0, 0};
    t.g[1] = {0, 1, 128, 128, layer, row_base, 128, 128, 0, 0, 0, 0, kHalf};
    t.g[2] = {0, 1, 256, 128, layer, row_base + 128, 128, 0, 0, 0, 0, kHalf, 0};
    t.g[3] = {0, 1, 384, 128, layer, row_base + 128, 128, 128, 0, 0, 0, kHalf, kHalf};
    t.cost = 128;
  };
  auto xpart = [&](int layer, int b_off, int row_base) {
    WTask& t = T.t[n++];
    t.n_units = 2;
    const int bl = (layer == 0) ? 0 : -1;
    t.u[0] = {0, kRec8XS, 8192, -1, 0, 0, 0}; t.u[1] = {1, b_off, kH8Bytes, bl, 0, 0, 0};
    t.n_groups = 2;
    t.g[0] = {0, 1, 0, 128, layer, row_base, dx, 0, 0, 0, 0, 0, 0};
    t.g[1] = {0, 1, 128, 128, layer, row_base, dx, 128, 0, 0, 0, 0, kHalf};
    t.cost = kXpartCost8;
  };
  xpart(0, kDz8Z0, 0);
  for (int l = 1; l <= 7; ++l) big(l, kRec8H0 + (l - 1) * kH8Bytes, kDz8Z0 + l * kH8Bytes, 0);
  xpart(5, kDz8Z0 + 5 * kH8Bytes, 256);
  {
    WTask& t = T.t[n++];
    t.n_units = 2;
    t.u[0] = {0, kRec8H0 + 7 * kH8Bytes, kH8Bytes, -1, 0, 0, 0};   // h7
    t.u[1] = {0, kRec8DS, 4096, -1, 0, kDz8P, 2048};               // PE(dir) and, at +4096, the packed d_pre operand
    t.n_groups = 3;
    t.g[0] = {0, 1, 0, 16, 8, 0, 128, 0, 1, 0, 0, 0, 4096};
    t.g[1] = {0, 1, 32, 16, 8, 128, 128, 0, 1, 0, 0, kHalf, 4096};
    t.g[2] = {1, 1, 64, 16, -1, 0, dd, 0, 4, 0, 0, 0, 4096};
    t.g[0].issuer = 0; t.g[1].issuer = 1; t.g[2].issuer = 0;
    t.cost = kHeadsCost8;
  }
  for (int k = 0; k < n; ++k) {
    WTask& t = T.t[k];
    if (k != n - 1)
      for (int gi = 0; gi < t.n_groups; ++gi) t.g[gi].issuer = gi & 1;
    for (int x = 0; x < 2; ++x)
      for (int u = 0; u < kWMaxUnits; ++u) {
        t.last_grp[x][u] = -1;
        for (int gi = 0; gi < t.n_groups; ++gi)
          if (t.g[gi].issuer == x && (t.g[gi].a == u || t.g[gi].b == u)) t.last_grp[x][u] = (int8_t)gi;
      }
  }
  return T;
}

struct WSmem {
  uint8_t slot[kWSlots][kWUnitBytes];
  uint64_t full[kWSlots], empty[kWSlots], acc_done, acc_free;
  uint32_t tmem_base;
};

struct WItem { int task; int64_t tile_lo, tile_hi; };

__device__ __forceinline__ WItem decode_item(const WTaskTable& T, int item, int slabs_per_cost_x128, int64_t n_tiles) {
  // item ids are laid out task after task; task k owns max(1, cost_k * slabs / 128) slabs
  WItem it{-1, 0, 0};
  int base = 0;
  for (int k = 0; k < kNumTasks; ++k) {
    const int ns = max(1, T.t[k].cost * slabs_per_cost_x128 / 128);
    if (item < base + ns) {
      const int s = item - base;
      it.task = k;
      it.tile_lo = n_tiles * s / ns;
      it.tile_hi = n_tiles * (s + 1) / ns;
      return it;
    }
    base += ns;
  }
  return it;
}

static int count_items(const WTaskTable& T, int slabs_per_cost_x128) {
  int n = 0;
  for (int k = 0; k < kNumTasks; ++k) n += std::max(1, T.t[k].cost * slabs_per_cost_x128 / 128);
  return n;
}

template <bool REC8>
__global__ void __launch_bounds__(kWThreads, 1)
tc_wgrad_kernel(const uint8_t* __restrict__ rec, const uint8_t* __restrict__ dz, int64_t n_tiles,
                float* __restrict__ grads, float* __restrict__ xbuf, TcParams P, const __grid_constant__ WTaskTable T,
                int n_items, int slabs) {
  constexpr int kRB = REC8 ? kRec8Bytes : kRecBytes, kDB = REC8 ? kDz8Bytes : kDzBytes;
  constexpr int kSub2 = REC8 ? 4096 : 8192;   // slot offset of a unit's second part (the packed d_pre operand)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  WSmem& sm = *reinterpret_cast<WSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    // a slot is free again when the MMAs that read it have completed AND the flush warps are done with it
    // (one release per issuer + one from the flush warps)
    for (int i = 0; i < kWSlots; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 3); }
    mbar_init(&sm.acc_done, 2);
    mbar_init(&sm.acc_free, 128);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const WItem w = decode_item(T, item, slabs, n_tiles);
        const WTask& t = T.t[w.task];
        for (int64_t tile = w.tile_lo; tile < w.tile_hi; ++tile) {
          for (int k = 0; k < t.n_units; ++k, ++it) {
            const uint32_t slot = it % kWSlots, ph = (it / kWSlots) & 1;
            const uint8_t* src = (t.u[k].src == 0 ? rec + tile * kRB : dz + tile * kDB) + t.u[k].off;
            mbar_wait(&sm.empty[slot], ph ^ 1);
            mbar_arrive_expect_tx(&sm.full[slot], t.u[k].bytes + t.u[k].bytes2);
            tma_load_1d(sm.slot[slot], src, t.u[k].bytes, &sm.full[slot]);
            if (t.u[k].bytes2)
              tma_load_1d(sm.slot[slot] + kSub2, dz + tile * kDB + t.u[k].off2, t.u[k].bytes2, &sm.full[slot]);
          }
        }
      }
    }
  } else if (warp == 1 || warp == 6) {
    if (lane == 0) {
      const int me = (warp == 1) ? 0 : 1;
      uint32_t it = 0, free_par = 0;
      bool first_item = true;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const WItem w = decode_item(T, item, slabs, n_tiles);
        const WTask& t = T.t[w.task];
        if (w.tile_lo >= w.tile_hi) continue;
        if (!first_item) {   // the previous item's accumulators must have been flushed
          mbar_wait(&sm.acc_free, free_par);
          free_par ^= 1;
          tc_fence_after();
        }
        first_item = false;
        for (int64_t tile = w.tile_lo; tile < w.tile_hi; ++tile, it += t.n_units) {
          for (int gi = 0; gi < t.n_groups; ++gi) {
            const WGroup& g = t.g[gi];
            if (g.issuer != me) continue;
            const uint32_t ia = it + g.a, ib = it + g.b;
            const uint32_t sa = ia % kWSlots, sb = ib % kWSlots;
            mbar_wait(&sm.full[sa], (ia / kWSlots) & 1);
            mbar_wait(&sm.full[sb], (ib / kWSlots) & 1);
            tc_fence_after();
            const uint32_t a_base = smem_u32(sm.slot[sa]) + g.a_sub, b_base = smem_u32(sm.slot[sb]) + g.b_sub;
            if (REC8) {   // e4m3 activations x e5m2 gradients, 32 samples per MMA (512 B further along every chunk)
              const uint32_t idesc = umma_idesc_f8(128, g.N, kE4M3, kE5M2, 1, 1);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_smem_desc(a_base + k * 512, 128, kChunkA);
                const uint64_t db = umma_smem_desc(b_base + k * 512, 128, kChunkA);
                umma_f8(tmem + g.col, da, db, idesc, (tile > w.tile_lo || k > 0) ? 1u : 0u);
              }
            } else {
              const uint32_t idesc = umma_idesc_bf16(128, g.N, 1, 1);
#pragma unroll
              for (int k = 0; k < 8; ++k) {   // K = 128 samples per tile, 16 per MMA; MN-major: LBO = 128 B, SBO = 2 KB
                const uint64_t da = umma_smem_desc(a_base + k * 256, 128, kChunkA);
                const uint64_t db = umma_smem_desc(b_base + k * 256, 128, kChunkA);
                umma_bf16(tmem + g.col, da, db, idesc, (tile > w.tile_lo || k > 0) ? 1u : 0u);
              }
            }
            if (t.last_grp[me][g.a] == gi) umma_commit(&sm.empty[sa]);
            if (g.b != g.a && t.last_grp[me][g.b] == gi) umma_commit(&sm.empty[sb]);
          }
          // units this issuer never reads: release them too (after their load has landed, so that the arrival
          // counts for this use of the slot and not for the previous one)
          for (int k = 0; k < t.n_units; ++k) {
            if (t.last_grp[me][k] >= 0) continue;
            const uint32_t iu = it + k;
            mbar_wait(&sm.full[iu % kWSlots], (iu / kWSlots) & 1);
            mbar_arrive(&sm.empty[iu % kWSlots]);
          }
        }
        umma_commit(&sm.acc_done);
      }
    }
  } else {
    // flush warps: TMEM lane quadrant = warp & 3
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int ftid = tid - 64, fc = ftid >> 3, fsub = ftid & 7;
    uint32_t done_par = 0, uit = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const WItem w = decode_item(T, item, slabs, n_tiles);
      const WTask& t = T.t[w.task];
      if (w.tile_lo >= w.tile_hi) continue;
      // While the tensor core works through the item, these warps take the bias gradients: column sums (over
      // samples) of every dZ unit passing through shared memory.  Thread (fc, fsub) owns 8 columns of chunk fc
      // and every 8th sample; partial sums stay in registers for the whole item.
      // (fp8 records: a task has ONE dZ unit of 16 chunks x 16 columns; bacc[0] / bacc[1] = columns 0..7 / 8..15 of chunk fc)
      const float inv_scale = REC8 ? dz8_scale(__ldg(reinterpret_cast<const uint32_t*>(xbuf) + kXOffAmax), true) : 1.f;
      float bacc[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int e = 0; e < 8; ++e) bacc[u][e] = 0.f;
      float pacc[3] = {0.f, 0.f, 0.f};          // this thread's row of the packed d_pre operand, summed over the tiles
      for (int64_t tile = w.tile_lo; tile < w.tile_hi; ++tile) {
        int nb = 0;
        for (int k = 0; k < t.n_units; ++k, ++uit) {
          const uint32_t slot = uit % kWSlots;
          mbar_wait(&sm.full[slot], (uit / kWSlots) & 1);
          if (t.u[k].bias_layer >= 0) {
            const uint8_t* base = sm.slot[slot] + fc * kChunkA + fsub * 16;
            if (REC8) {
#pragma unroll 4
              for (int i = 0; i < 16; ++i) {
                const uint4 v = *reinterpret_cast<const uint4*>(base + i * 128);
                const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 lo = e5m2x2_lo(wv[j]), hi = e5m2x2_hi(wv[j]);
                  float* a = bacc[j >> 1] + (j & 1) * 4;
                  a[0] += lo.x; a[1] += lo.y; a[2] += hi.x; a[3] += hi.y;
                }
              }
            } else {
              float* a = bacc[nb & 1];
#pragma unroll 4
              for (int i = 0; i < 16; ++i) {
                const uint4 v = *reinterpret_cast<const uint4*>(base + i * 128);
                a[0] += bf16_lo(v.x); a[1] += bf16_hi(v.x); a[2] += bf16_lo(v.y); a[3] += bf16_hi(v.y);
                a[4] += bf16_lo(v.z); a[5] += bf16_hi(v.z); a[6] += bf16_lo(v.w); a[7] += bf16_hi(v.w);
              }
            }
            ++nb;
          }
          if (!REC8 && t.u[k].bytes2 != 0) {   // the d_pre operand at +8192: [2 chunks][128 rows][8], chunk 0 = (r, g, b, sigma, 0..)
            const uint2 v = *reinterpret_cast<const uint2*>(sm.slot[slot] + 8192 + ftid * 16);   // one row per thread: this
            pacc[0] += bf16_lo(v.x); pacc[1] += bf16_hi(v.x); pacc[2] += bf16_lo(v.y);           // sits on the slot-release path
          }
          named_bar_sync(2, 128);               // all 128 readers are done with the slot
          if (ftid == 0) mbar_arrive(&sm.empty[slot]);
        }
      }
      {
        int nb = 0;
        for (int k = 0; k < t.n_units; ++k) {
          if (t.u[k].bias_layer < 0) continue;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (!REC8 && u != (nb & 1)) continue;
            float* a = bacc[u];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              a[e] += __shfl_xor_sync(0xffffffffu, a[e], 1);
              a[e] += __shfl_xor_sync(0xffffffffu, a[e], 2);
              a[e] += __shfl_xor_sync(0xffffffffu, a[e], 4);
            }
            if (fsub == 0) {
              const int c0 = t.u[k].bias_col0 + (REC8 ? fc * 16 + u * 8 : fc * 8);
              float* dst = grads + P.b_off[t.u[k].bias_layer] + c0;
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (c0 + e < P.U) atomicAdd(dst + e, a[e] * inv_scale);
            }
          }
          ++nb;
        }
      }
      if (!REC8 && w.task == kNumTasks - 1) {   // heads task: sum d(rgb_pre) of this call for tc_finish_kernel
#pragma unroll
        for (int e = 0; e < 3; ++e) {
          const float sres = warp_sum(pacc[e]);
          if (lane == 0) atomicAdd(xbuf + kXOffS + e, sres);
        }
      }
      mbar_wait(&sm.acc_done, done_par);
      done_par ^= 1;
      tc_fence_after();
      for (int gi = 0; gi < t.n_groups; ++gi) {
        const WGroup& g = t.g[gi];
        const int ncol = g.N < 32 ? 32 : g.N;
        for (int c0 = 0; c0 < ncol; c0 += 32) {
          float v[32];
          tmem_ld32(tmem + lane_base + g.col + c0, v);   // warp-collective: outside the row guard
          if (REC8) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= inv_scale;
          }
          if (row < g.row_limit) {
            if (g.mode == 0) {   // kernel [fan_in, U] of a hidden layer: columns >= U belong to the padding
              float* dst = grads + P.w_off[g.layer] + (int64_t)(g.row_base + row) * P.U + g.col_base + c0;
              const int nv = P.U - (g.col_base + c0);
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < nv) atomicAdd(dst + i, v[i]);
            } else if (g.mode == 1) {
              atomicAdd(grads + P.w_off[8] + g.row_base + row, v[3]);
              float* dst = xbuf + kXOffY + (g.row_base + row) * 4;
              atomicAdd(dst, v[0]); atomicAdd(dst + 1, v[1]); atomicAdd(dst + 2, v[2]);
            } else {
              float* dst = xbuf + kXOffYd + row * 4;
              atomicAdd(dst, v[0]); atomicAdd(dst + 1, v[1]); atomicAdd(dst + 2, v[2]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&sm.acc_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// =============================================================================================================
// heads: features / rgb_features / rgb gradients from the rank-3 factors
// =============================================================================================================
// With Y = h7^T d(rgb_pre) [256 x 3], Yd = PE(dir)^T d(rgb_pre) [27 x 3], s3 = sum d(rgb_pre) [3] of THIS call (scratch,
// 4 floats per row) and the weights W_f [256,256], b_f, W_g [283,128], W_c [128,3], W' = W_f W_g[:256], b' (fold):
//   X = h7^T dG = Y W_c^T,  sum dG = s3 W_c^T   (never formed: T[j][c] = sum_n W_g[j][n] W_c[n][c],
//                                                U[j][c] = sum_i W_f[i][j] Y[i][c] + b_f[j] s3[c])
//   dW_f[i][j]        += sum_c Y[i][c] T[j][c]                    db_f[j]  += sum_c s3[c] T[j][c]
//   dW_g[j][n]        += sum_c U[j][c] W_c[n][c]      (j < 256)   db_g[n]  += sum_c s3[c] W_c[n][c]
//   dW_g[256 + i][n]  += sum_c Yd[i][c] W_c[n][c]     (i < 27)
//   dW_c[k][c]        += sum_j W'[j][k] Y[j][c] + sum_i W_g[256+i][k] Yd[i][c] + b'[k] s3[c]
// tc_finish_prep_kernel forms T and U (256 x 3 each), tc_finish_kernel one output element per thread.
constexpr int kFinF = 256 * 256, kFinG = 256 * 128, kFinBf = 256, kFinGd = 27 * 128, kFinBg = 128, kFinC = 128 * 3;   // (kFinGd: rows >= dd idle)
constexpr int kFinTotal = kFinF + kFinG + kFinBf + kFinGd + kFinBg + kFinC;

// the two U x 3 factors every output of tc_finish_kernel needs (one thread each; 0.3 M MAC)
// (U = dense_units <= 256, H = U / 2: the kernels are [U, U], [U + dd, H], [H, 3] in the flat buffer; the scratch and the
//  fold keep their 256 / 128 strides)
__global__ void __launch_bounds__(256) tc_finish_prep_kernel(const float* __restrict__ params, TcParams P,
                                                             float* __restrict__ xbuf) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // < 2 * 256 * 3
  const int which = idx / 768, e = idx - which * 768, j = e / 3, c = e - j * 3;
  const int U = P.U, H = P.U / 2;
  if (which > 1) return;
  const float* Wc = params + P.w_off[11];
  float acc = 0.f;
  if (which == 0) {
    if (j < U) {
      const float* w = params + P.w_off[10] + (int64_t)j * H;
      for (int n = 0; n < H; ++n) acc = fmaf(w[n], Wc[n * 3 + c], acc);
    }
    xbuf[kXOffT + j * 4 + c] = acc;
  } else {
    if (j < U) {
      const float* Wf = params + P.w_off[9];
      const float* Y = xbuf + kXOffY;
      acc = params[P.b_off[9] + j] * xbuf[kXOffS + c];
      for (int i = 0; i < U; ++i) acc = fmaf(Wf[(int64_t)i * U + j], Y[i * 4 + c], acc);
    }
    xbuf[kXOffU + j * 4 + c] = acc;
  }
}

__global__ void __launch_bounds__(256) tc_finish_kernel(const float* __restrict__ params, TcParams P,
                                                        const float* __restrict__ xbuf, float* __restrict__ grads,
                                                        const float* __restrict__ fold) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int U = P.U, H = P.U / 2;
  const float* Wg = params + P.w_off[10];   // rgb_features  [U + dd, H]
  const float* Wc = params + P.w_off[11];   // rgb           [H, 3]
  const float* Y = xbuf + kXOffY;
  const float* Yd = xbuf + kXOffYd;
  const float* T = xbuf + kXOffT;
  const float* Uf = xbuf + kXOffU;
  const float s3[3] = {xbuf[kXOffS], xbuf[kXOffS + 1], xbuf[kXOffS + 2]};
  if (idx < kFinF) {
    const int i = idx >> 8, j = idx & 255;
    if (i < U && j < U)
      grads[P.w_off[9] + (int64_t)i * U + j] += Y[i * 4] * T[j * 4] + Y[i * 4 + 1] * T[j * 4 + 1] + Y[i * 4 + 2] * T[j * 4 + 2];
    return;
  }
  idx -= kFinF;
  if (idx < kFinG) {
    const int j = idx >> 7, n = idx & 127;
    if (j < U && n < H)
      grads[P.w_off[10] + (int64_t)j * H + n] += Uf[j * 4] * Wc[n * 3] + Uf[j * 4 + 1] * Wc[n * 3 + 1] + Uf[j * 4 + 2] * Wc[n * 3 + 2];
    return;
  }
  idx -= kFinG;
  if (idx < kFinBf) {
    if (idx < U) grads[P.b_off[9] + idx] += s3[0] * T[idx * 4] + s3[1] * T[idx * 4 + 1] + s3[2] * T[idx * 4 + 2];
    return;
  }
  idx -= kFinBf;
  if (idx < kFinGd) {
    const int i = idx >> 7, n = idx & 127;
    if (i < P.dd && n < H)
      grads[P.w_off[10] + (int64_t)(U + i) * H + n] += Yd[i * 4] * Wc[n * 3] + Yd[i * 4 + 1] * Wc[n * 3 + 1] + Yd[i * 4 + 2] * Wc[n * 3 + 2];
    return;
  }
  idx -= kFinGd;
  if (idx < kFinBg) {
    if (idx < H) grads[P.b_off[10] + idx] += s3[0] * Wc[idx * 3] + s3[1] * Wc[idx * 3 + 1] + s3[2] * Wc[idx * 3 + 2];
    return;
  }
  idx -= kFinBg;
  if (idx < kFinC) {
    const int k = idx / 3, c = idx - k * 3;
    if (k >= H) return;
    float acc = fold[256 * 128 + k] * s3[c];
    for (int j = 0; j < U; ++j) acc = fmaf(fold[j * 128 + k], Y[j * 4 + c], acc);
    for (int i = 0; i < P.dd; ++i) acc = fmaf(Wg[(int64_t)(U + i) * H + k], Yd[i * 4 + c], acc);
    grads[P.w_off[11] + idx] += acc;
  }
}

}  // namespace

// parts: which of the backward kernels to launch (bit 0 dgrad, bit 1 wgrad + finish; KNERF_BWD_*_ONLY)
int tc_backward(const Model& m, const float* params, const void* packed, const float* d_pre, int64_t R, int S,
                float* grads, char* ws, int64_t ws_bytes, int parts, bool rec8, cudaStream_t st) {
  if (!is_flagship(m)) return fail(KNERF_ERR_UNSUPPORTED, "KNERF_BF16 implements models of dense_units <= 256 (even), up to 8 layers, at most one skip concat (not into the heads), L_xyz <= 10, L_dir <= 4");
  KN_CHECK_ARG(packed != nullptr, "KNERF_BF16 needs packed weights (knerf_pack_weights)");
  const int64_t M = R * S;
  if (ws_bytes < tc_workspace_bytes(m, M, true, rec8))
    return fail(KNERF_ERR_WORKSPACE, "tc_backward: workspace %lld < %lld bytes", (long long)ws_bytes,
                (long long)tc_workspace_bytes(m, M, true, rec8));
  KN_CHECK_ARG((reinterpret_cast<uintptr_t>(d_pre) & 15) == 0, "tc_backward: d_pre must be 16-byte aligned");
  const int64_t n_tiles = cdiv(M, kTileM);
  float* xbuf = (float*)ws;                                // X = h7^T dG and sum(dG) of this call
  uint8_t* rec = (uint8_t*)ws + kXBytes;
  uint8_t* dz = rec + n_tiles * (int64_t)(rec8 ? kRec8Bytes : kRecBytes);
  const TcParams P = tc_make_params(m);
  // the fp32 scratch (rank-3 factors, sum d(rgb_pre), max|d_pre|) starts from zero in every call -- except that a
  // weight-gradient-only call (KNERF_BWD_WGRAD_ONLY) over fp8 records keeps what the dgrad-only call before it left
  // there: sum d(rgb_pre) and the scale of the gradient records
  KN_CUDA(cudaMemsetAsync(xbuf, 0, ((rec8 && parts == 2) ? kXOffS : kXFloats) * sizeof(float), st));

  if (parts & 1) {
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(kThreads);
    cfg.stream = st;
    cfg.gridDim = dim3((unsigned)(2 * std::min<int64_t>(cdiv(n_tiles, 4), kNumSMs / 2)));
    cfg.dynamicSmemBytes = sizeof(ChainSmem);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const uint8_t* pk = (const uint8_t*)packed;
    const float4* dp = (const float4*)d_pre;
    const uint8_t* rec_c = rec;
    if (rec8) {
      dpre_amax_kernel<<<kNumSMs * 4, 256, 0, st>>>(dp, M, reinterpret_cast<uint32_t*>(xbuf) + kXOffAmax);
      KN_LAUNCH_CHECK();
      KN_CUDA(cudaFuncSetAttribute(tc_mlp_dgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes));
      KN_CUDA(cudaLaunchKernelEx(&cfg, tc_mlp_dgrad_kernel<true>, pk, dp, M, rec_c, dz, grads, P, xbuf));
    } else {
      KN_CUDA(cudaFuncSetAttribute(tc_mlp_dgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes));
      KN_CUDA(cudaLaunchKernelEx(&cfg, tc_mlp_dgrad_kernel<false>, pk, dp, M, rec_c, dz, grads, P, xbuf));
    }
    KN_LAUNCH_CHECK();
  }
  if (parts & 2) {
    // ~3 KB, passed by value as a __grid_constant__
    WTaskTable h_table = rec8 ? build_task_table8(m.dx, m.dd) : build_task_table(m.dx, m.dd);
    fit_tasks_to_model(h_table, P);
    // items = 2 x 148: 7 tasks of cost 128, 2 of cost 76, 1 of cost 178 -> 217 + 36 + 43 = 296 items for 31 slabs
    const int slabs = (int)std::max<int64_t>(1, std::min<int64_t>(rec8 ? kSlabs8 : 31, n_tiles / 4));
    const int n_items = count_items(h_table, slabs);
    const int grid = std::min(n_items, kNumSMs);
    const size_t smem = sizeof(WSmem);
    if (rec8) {
      KN_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tc_wgrad_kernel<true><<<grid, kWThreads, smem, st>>>(rec, dz, n_tiles, grads, xbuf, P, h_table, n_items, slabs);
    } else {
      KN_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tc_wgrad_kernel<false><<<grid, kWThreads, smem, st>>>(rec, dz, n_tiles, grads, xbuf, P, h_table, n_items, slabs);
    }
    KN_LAUNCH_CHECK();
    tc_finish_prep_kernel<<<6, 256, 0, st>>>(params, P, xbuf);
    KN_LAUNCH_CHECK();
    tc_finish_kernel<<<(kFinTotal + 255) / 256, 256, 0, st>>>(
        params, P, xbuf, grads, reinterpret_cast<const float*>((const uint8_t*)packed + kFoldOff));
    KN_LAUNCH_CHECK();
  }
  return KNERF_OK;
}

}  // namespace knerf
