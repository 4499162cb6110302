// Warp roles shared by the fused forward and dgrad "chain" kernels: setup, the TMA weight producer and the
// single-thread tcgen05.mma issuer.  A chain kernel walks Prog::kSteps GEMM steps for two 128-sample tiles in
// ping-pong; the compute warps (kernel specific) turn each accumulator into the next step's A operand.
#pragma once

#include "tc_layout.cuh"
#include "tc_ptx.cuh"

namespace knerf {
namespace tcl {
using namespace tc;

// ---- optional in-kernel timing (build with -DKNERF_TC_TIMING; off by default, zero cost) ----------------------
// per CTA, clock64() cycles: slot 0 producer waits for a free stage | 1 MMA thread waits for the A operand |
// 2 MMA thread waits for a weight stage | 3 MMA thread total | 4 compute warp 2 waits for an accumulator |
// 5 compute total | 6.. per-step epilogue cycles of compute warp 2 (16 + s for tile slot 1)
#ifdef KNERF_TC_TIMING
static __device__ unsigned long long g_tc_prof[160][40];
// counters live in the thread's local memory (L1) and are flushed once at the end of the role
#define KN_PROF_DECL() unsigned long long kn_prof[40]; for (int kn_i = 0; kn_i < 40; ++kn_i) kn_prof[kn_i] = 0ull
#define KN_PROF_BEGIN(var) const long long var = clock64()
#define KN_PROF_END(var, slot) kn_prof[slot] += (unsigned long long)(clock64() - var)
#define KN_PROF_FLUSH() for (int kn_i = 0; kn_i < 40; ++kn_i) if (kn_prof[kn_i]) g_tc_prof[blockIdx.x][kn_i] += kn_prof[kn_i]
#else
#define KN_PROF_DECL()
#define KN_PROF_BEGIN(var)
#define KN_PROF_END(var, slot)
#define KN_PROF_FLUSH()
#endif

__device__ __forceinline__ uint32_t chain_setup(ChainSmem& sm, int tid, int warp) {
  if (tid == 0) {
    for (int i = 0; i < kNumStages; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.a_ready[i], kComputeThreads); mbar_init(&sm.acc_ready[i], 1);
      mbar_init(&sm.st_ready[i], 8); mbar_init(&sm.st_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&sm.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return sm.tmem_base;
}

__device__ __forceinline__ void chain_teardown(uint32_t tmem, int warp) {
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// warp 0, one lane: stream the step's weight stages (once per tile) through the 4-slot ring
template <class Prog>
__device__ __forceinline__ void producer_role(ChainSmem& sm, const uint8_t* __restrict__ blob, int64_t n_pairs) {
  uint32_t it = 0;
  const uint64_t pol = l2_policy_evict_last();
  KN_PROF_DECL();
  for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
    for (int s = 0; s < Prog::kSteps; ++s) {
      const int nk = Prog::nk_h(s) + Prog::nk_x(s);
      const uint32_t sb = Prog::stage_bytes(s);
      const uint8_t* src = blob + Prog::blob_off(s);
      for (int tl = 0; tl < 2; ++tl) {
        for (int ks = 0; ks < nk; ++ks, ++it) {
          const uint32_t slot = it % kNumStages, ph = (it / kNumStages) & 1;
          KN_PROF_BEGIN(t0);
          mbar_wait(&sm.empty[slot], ph ^ 1);
          KN_PROF_END(t0, 0);
          mbar_arrive_expect_tx(&sm.full[slot], sb);
          tma_load_1d_hint(sm.stage[slot], src + (size_t)ks * sb, sb, &sm.full[slot], pol);
        }
        if (Prog::kHasBias) {   // the bias "stage": [2 chunks][N][8]
          const uint32_t slot = it % kNumStages, ph = (it / kNumStages) & 1, bb = Prog::bias_bytes(s);
          mbar_wait(&sm.empty[slot], ph ^ 1);
          mbar_arrive_expect_tx(&sm.full[slot], bb);
          tma_load_1d_hint(sm.stage[slot], src + (size_t)nk * sb, bb, &sm.full[slot], pol);
          ++it;
        }
      }
    }
  }
  KN_PROF_FLUSH();
}

// warp 1, one lane: D[tile] = A[tile] * W_step^T, K = 32 per stage = two K=16 tcgen05.mma
template <class Prog>
__device__ __forceinline__ void mma_role(ChainSmem& sm, uint32_t tmem, int64_t n_pairs) {
  uint32_t it = 0, a_par[2] = {0, 0};
  KN_PROF_DECL();
  KN_PROF_BEGIN(t_all);
  for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
    for (int s = 0; s < Prog::kSteps; ++s) {
      const int nkh = Prog::nk_h(s), nk = nkh + Prog::nk_x(s);
      const int N = Prog::N(s);
      const uint32_t idesc = umma_idesc_bf16(kTileM, N, 0, 0);
      const uint32_t chunk_b = (uint32_t)N * 16;
      for (int tl = 0; tl < 2; ++tl) {
        KN_PROF_BEGIN(t_a);
        mbar_wait(&sm.a_ready[tl], a_par[tl]);
        KN_PROF_END(t_a, 1);
        a_par[tl] ^= 1;
        tc_fence_after();
        const uint32_t d_tmem = tmem + tl * 256;
        for (int ks = 0; ks < nk; ++ks, ++it) {
          const uint32_t slot = it % kNumStages, ph = (it / kNumStages) & 1;
          KN_PROF_BEGIN(t_f);
          mbar_wait(&sm.full[slot], ph);
          KN_PROF_END(t_f, 2);
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
        if (Prog::kHasBias) {   // + 1 * bias: A = the two encoding chunks holding the constant-1 column
          const uint32_t slot = it % kNumStages, ph = (it / kNumStages) & 1;
          mbar_wait(&sm.full[slot], ph);
          tc_fence_after();
          const uint64_t da = umma_smem_desc(smem_u32(sm.xs[tl]) + Prog::bias_a_chunk(s) * kChunkA, kChunkA, 128);
          const uint64_t db = umma_smem_desc(smem_u32(sm.stage[slot]), chunk_b, 128);
          umma_bf16(d_tmem, da, db, idesc, 1u);
          umma_commit(&sm.empty[slot]);
          ++it;
        }
        umma_commit(&sm.acc_ready[tl]);
      }
    }
  }
  KN_PROF_END(t_all, 3);
  KN_PROF_FLUSH();
}

// warp 10, one lane (training kernels): every operand tile the compute warps leave in hs[tl] is also a saved
// record -- written to HBM by bulk copies instead of 16 STG.128 per compute thread.  Protocol per tile slot:
// compute warps arrive on st_ready[tl] (one arrive per warp) once their part of hs[tl] is written and fenced;
// this thread stores the tile, waits until the copy engine has read shared memory and arrives on st_done[tl],
// which the compute warps wait for before they overwrite hs[tl] in their next epilogue.
// dst(item, tile) -> destination, bytes(item) -> size.
// The tile goes out in 16 KB pieces with at most two in flight: a whole-tile copy ahead of them in the SM's copy
// queue delays the weight stages (measured: the MMA thread then waits for stages a third of the time).
constexpr uint32_t kStorePiece = 16384;
template <class Smem, class TileOf, class Dst, class Bytes>
__device__ __forceinline__ void store_role(Smem& sm, int items_per_tile, int64_t n_tiles, int64_t n_units, int64_t first,
                                           int64_t stride, TileOf tile_of, Dst dst, Bytes bytes) {
  uint32_t par = 0;
  const uint64_t pol = l2_policy_evict_first();
  for (int64_t unit = first; unit < n_units; unit += stride) {
#pragma unroll 1
    for (int item = 0; item < items_per_tile; ++item) {
#pragma unroll 1
      for (int tl = 0; tl < 2; ++tl) {
        const int64_t tile = tile_of(unit, tl);
        mbar_wait(&sm.st_ready[tl], (par >> tl) & 1);
        par ^= 1u << tl;
        if (tile < n_tiles) {
          uint8_t* g = dst(item, tile);
          const uint32_t nb = bytes(item);
#pragma unroll 1
          for (uint32_t off = 0; off < nb; off += kStorePiece) {
            tma_store_1d_hint(g + off, sm.hs[tl] + off, kStorePiece, pol);
            tma_store_commit();
            tma_store_wait_read_1();
          }
          tma_store_wait_read();
        }
        mbar_arrive(&sm.st_done[tl]);
      }
    }
  }
}

// compute warps, after their writes to hs[tl] are fenced (fence.proxy.async) and the warp has converged
__device__ __forceinline__ void st_ready_arrive(uint64_t* bar, int lane) {
  if (lane == 0) mbar_arrive(bar);
}

// column sums of a [128 x ncols] bf16 operand sitting in shared memory (chunk-major) -> atomically added to
// dst[0..ncols) (bias gradients).  All 256 compute threads; caller has synchronised them after the writes.
static __device__ __noinline__ void colsum_to_global(const uint8_t* hs, int ncols, float* __restrict__ dst, int ctid) {
  const int c = ctid >> 3, sub = ctid & 7;
  if (c * 8 < ncols) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      // rows i*8 + sub: the 8 lanes of a quarter-warp read 128 contiguous bytes (conflict-free LDS.128)
      const uint4 v = *reinterpret_cast<const uint4*>(hs + c * kChunkA + (i * 8 + sub) * 16);
      acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
      acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 1);
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 2);
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 4);
    }
    if (sub == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(dst + c * 8 + e, acc[e]);
    }
  }
}

}  // namespace tcl
}  // namespace knerf
