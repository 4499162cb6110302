// cta_group::2 variant of the chain-kernel roles (tc_roles.cuh): a CTA PAIR (cluster of 2 = one TPC) works on
// four 128-sample tiles -- two per CTA, ping-pong as before -- and every GEMM step is ONE M=256 tcgen05.mma per
// K=16 issued by the leader CTA: each CTA feeds its own A tile and HALF of the weight rows (N/2), so the
// shared-memory traffic per SM (operand reads + TMA writes) and the L2->SM weight traffic are both halved, which
// is what bounds the single-CTA kernel.
//
//   warp 0 lane 0 (both CTAs)  TMA producer: this CTA's half of every weight stage, 8-slot ring of 8 KB
//   warp 1 lane 0 (peer CTA)   relay: local "stage landed" -> arrive on the leader's peer_full barrier
//   warp 1 lane 0 (leader)     MMA issuer; tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs
//   warps 2-9    (both CTAs)   compute warps; "A operand ready" = one arrive per warp on the LEADER's barrier
#pragma once

#include "tc_layout.cuh"
#include "tc_ptx.cuh"

namespace knerf {
namespace tcl {
using namespace tc;

constexpr int kNumStages2 = 8;
constexpr int kStageBytes2 = kStageBytes / 2;   // 8 KB: [4 chunks][128 rows][8]

struct Chain2Smem {
  uint8_t hs[2][kHSBytes];
  uint8_t xs[2][kXSBytes];
  uint8_t stage[kNumStages2][kStageBytes2];
  float part[kTileM][4];
  uint64_t full[kNumStages2], peer_full[kNumStages2], empty[kNumStages2], a_ready[2], acc_ready[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t chain2_setup(Chain2Smem& sm, int tid, int warp) {
  if (tid == 0) {
    for (int i = 0; i < kNumStages2; ++i) {
      mbar_init(&sm.full[i], 1); mbar_init(&sm.peer_full[i], 1); mbar_init(&sm.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.a_ready[i], 16); mbar_init(&sm.acc_ready[i], 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<512>(&sm.tmem_base);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  return sm.tmem_base;
}

__device__ __forceinline__ void chain2_teardown(uint32_t tmem, int warp) {
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta<512>(tmem);
}

// the 1-CTA blob layout [chunks][N][8] is reused: this CTA's half of a stage = `chunks` pieces of N/2 rows
template <class Prog>
__device__ __forceinline__ void producer2_role(Chain2Smem& sm, const uint8_t* __restrict__ blob, uint32_t cta,
                                               int64_t n_quads, int64_t first, int64_t stride) {
  uint32_t it = 0;
  for (int64_t quad = first; quad < n_quads; quad += stride) {
    for (int s = 0; s < Prog::kSteps; ++s) {
      const int nk = Prog::nk_h(s) + Prog::nk_x(s);
      const uint32_t sb = Prog::stage_bytes(s);            // full-N stage bytes in the blob
      const uint32_t piece = Prog::N(s) * 8;               // N/2 rows x 16 B
      const uint8_t* src = blob + Prog::blob_off(s) + cta * piece;
      for (int tl = 0; tl < 2; ++tl) {
        for (int ks = 0; ks < nk; ++ks, ++it) {
          const uint32_t slot = it % kNumStages2, ph = (it / kNumStages2) & 1;
          mbar_wait_cluster(&sm.empty[slot], ph ^ 1);
          mbar_arrive_expect_tx(&sm.full[slot], 4 * piece);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            tma_load_1d(sm.stage[slot] + c * piece, src + (size_t)ks * sb + c * 2 * piece, piece, &sm.full[slot]);
        }
        if (Prog::kHasBias) {
          const uint32_t slot = it % kNumStages2, ph = (it / kNumStages2) & 1;
          mbar_wait_cluster(&sm.empty[slot], ph ^ 1);
          mbar_arrive_expect_tx(&sm.full[slot], 2 * piece);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            tma_load_1d(sm.stage[slot] + c * piece, src + (size_t)nk * sb + c * 2 * piece, piece, &sm.full[slot]);
          ++it;
        }
      }
    }
  }
}

// peer CTA: forward "my half of slot k has landed" to the leader, in ring order
template <class Prog>
__device__ __forceinline__ void relay_role(Chain2Smem& sm, int64_t n_quads, int64_t first, int64_t stride) {
  uint32_t it = 0;
  for (int64_t quad = first; quad < n_quads; quad += stride) {
    for (int s = 0; s < Prog::kSteps; ++s) {
      const int n = (Prog::nk_h(s) + Prog::nk_x(s) + (Prog::kHasBias ? 1 : 0)) * 2;
      for (int i = 0; i < n; ++i, ++it) {
        const uint32_t slot = it % kNumStages2, ph = (it / kNumStages2) & 1;
        mbar_wait(&sm.full[slot], ph);
        mbar_arrive_cluster(&sm.peer_full[slot], 0);
      }
    }
  }
}

template <class Prog>
__device__ __forceinline__ void mma2_role(Chain2Smem& sm, uint32_t tmem, int64_t n_quads, int64_t first, int64_t stride) {
  uint32_t it = 0, a_par[2] = {0, 0};
  for (int64_t quad = first; quad < n_quads; quad += stride) {
    for (int s = 0; s < Prog::kSteps; ++s) {
      const int nkh = Prog::nk_h(s), nk = nkh + Prog::nk_x(s);
      const int N = Prog::N(s);
      const uint32_t idesc = umma_idesc_bf16(2 * kTileM, N, 0, 0);
      const uint32_t chunk_b = (uint32_t)N * 8;            // N/2 rows x 16 B
      for (int tl = 0; tl < 2; ++tl) {
        mbar_wait_cluster(&sm.a_ready[tl], a_par[tl]);
        a_par[tl] ^= 1;
        tc_fence_after();
        const uint32_t d_tmem = tmem + tl * 256;
        for (int ks = 0; ks < nk; ++ks, ++it) {
          const uint32_t slot = it % kNumStages2, ph = (it / kNumStages2) & 1;
          mbar_wait(&sm.full[slot], ph);
          mbar_wait_cluster(&sm.peer_full[slot], ph);
          tc_fence_after();
          const uint32_t a_base = (ks < nkh) ? smem_u32(sm.hs[tl]) + ks * 4 * kChunkA
                                             : smem_u32(sm.xs[tl]) + (ks - nkh) * 4 * kChunkA;
          const uint32_t b_base = smem_u32(sm.stage[slot]);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t da = umma_smem_desc(a_base + j * 2 * kChunkA, kChunkA, 128);
            const uint64_t db = umma_smem_desc(b_base + j * 2 * chunk_b, chunk_b, 128);
            umma_bf16_2cta(d_tmem, da, db, idesc, (ks > 0 || j > 0) ? 1u : 0u);
          }
          umma_commit_2cta(&sm.empty[slot], 3);
        }
        if (Prog::kHasBias) {
          const uint32_t slot = it % kNumStages2, ph = (it / kNumStages2) & 1;
          mbar_wait(&sm.full[slot], ph);
          mbar_wait_cluster(&sm.peer_full[slot], ph);
          tc_fence_after();
          const uint64_t da = umma_smem_desc(smem_u32(sm.xs[tl]) + Prog::bias_a_chunk(s) * kChunkA, kChunkA, 128);
          const uint64_t db = umma_smem_desc(smem_u32(sm.stage[slot]), chunk_b, 128);
          umma_bf16_2cta(d_tmem, da, db, idesc, 1u);
          umma_commit_2cta(&sm.empty[slot], 3);
          ++it;
        }
        umma_commit_2cta(&sm.acc_ready[tl], 3);
      }
    }
  }
}

// compute warps: "this warp's part of tile slot tl is written": publish to the async proxy, then ONE arrive per
// warp on the leader's barrier (16 = 8 warps x 2 CTAs)
__device__ __forceinline__ void a_ready_arrive2(Chain2Smem& sm, int tl, int lane) {
  tc_fence_before();
  fence_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(&sm.a_ready[tl], 0);
}

}  // namespace tcl
}  // namespace knerf
