// a8 + the sort of a9: inverse-CDF hierarchical resampling and merge with the coarse depths
// (keras_nerf/model/nerf/utils.py:60-97; keras_nerf/model/nerf/nerf.py:182-191).
//
// One warp per ray.  pdf/cdf: warp-shuffle inclusive scan over 32-sample blocks with a carried prefix,
// staged in shared memory; searchsorted(side='right') = per-lane upper-bound binary search over the
// staged cdf (== "count of cdf_j <= u" because the cdf is non-decreasing); the Nf fine samples are
// bitonic-sorted in registers (P per lane) and rank-merged with the already ascending coarse depths,
// so the Nc+Nf output row is written once, coalesced.
// HBM bytes per ray (algorithmic): 8*Nc + 4*Nf read, 4*(Nc+Nf) written.
#include <math_constants.h>

#include "common.cuh"

namespace knerf {

constexpr int kSampWarps = 4;

template <int P>
__device__ __forceinline__ void bitonic_sort_regs(float (&v)[P], int lane) {
  // element index e = lane*P + r ; ascending over e
#pragma unroll
  for (int k = 2; k <= 32 * P; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= P) {
        // j >= P and k >= 2P: both bits lie in the lane part of e, so one predicate serves all P registers
        const int lj = j / P;
        const bool take_min = (((lane * P) & k) == 0) == (((lane * P) & j) == 0);
#pragma unroll
        for (int r = 0; r < P; ++r) {
          const float o = __shfl_xor_sync(kFullMask, v[r], lj);
          v[r] = take_min ? fminf(v[r], o) : fmaxf(v[r], o);
        }
      } else {
#pragma unroll
        for (int r = 0; r < P; ++r) {
          if ((r & j) == 0) {
            const int e = lane * P + r;
            const bool up = (e & k) == 0;
            const float a = v[r], b = v[r | j];
            const float mn = fminf(a, b), mx = fmaxf(a, b);
            v[r] = up ? mn : mx;
            v[r | j] = up ? mx : mn;
          }
        }
      }
    }
  }
}

// Level-order (Eytzinger) slot of sorted position p in a complete tree of 2^LOG - 1 nodes stored at [1, 2^LOG):
// the nodes of one level are contiguous, so the 32 lanes' probes of a search level fall into consecutive
// shared-memory banks (a plain binary search probes addresses = step - 1 mod 2*step: one or two banks per level).
__device__ __forceinline__ int eyt_slot(int p, int LOG) {
  const int t = __ffs(p + 1) - 1;
  return (1 << (LOG - 1 - t)) + ((p + 1) >> (t + 1));
}

// #{entries <= v} (INCL) or #{entries < v} of the tree `e` (2^LOG - 1 nodes, missing ones = +inf), for NQ
// independent queries at once (interleaved LDS chains)
template <int LOG, int NQ, bool INCL>
__device__ __forceinline__ void eyt_count(const float* e, const float (&v)[NQ], int (&cnt)[NQ]) {
  int k[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) k[q] = 1;
#pragma unroll
  for (int l = 0; l < LOG; ++l) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const float x = e[k[q]];
      k[q] = 2 * k[q] + ((INCL ? (x <= v[q]) : (x < v[q])) ? 1 : 0);
    }
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q) cnt[q] = k[q] - (1 << LOG);
}

template <int N> struct Log2 { static constexpr int value = 1 + Log2<N / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

struct SampleArgs {
  const float* t_coarse; const float* mid_points; const float* weights; const float* u;
  uint64_t seed; const float* cdf_in; int64_t R; int Nc; int Nf; int oob_mode; int sequential;
  float* t_sorted; float* samples; int32_t* indices; float* cdf_out; int32_t* oob_count;
  int smem_per_warp;   // floats
};

template <int NCB, int P>
__global__ void __launch_bounds__(kSampWarps * 32) sample_fine_kernel(SampleArgs a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int Nc = a.Nc, Nf = a.Nf;
  const int nc1 = (Nc + 1 + 3) & ~3;
  // Searches run over level-order copies (eyt_slot) of the cdf, the coarse depths and the sorted fine samples,
  // padded with +inf to a complete tree: fixed depth, branchless, conflict-free, all of a lane's searches interleaved.
  constexpr int CAP = 64 * NCB;                        // > Nc + 1
  constexpr int LOGC = Log2<CAP>::value, LOGF = Log2<32 * P>::value;
  float* cdf = smem + (size_t)wib * a.smem_per_warp;   // [Nc+1] linear (gathers, cdf_out)
  float* cdf_e = cdf + nc1;                            // [CAP] level order
  float* tcs_e = cdf_e + CAP;                          // [CAP] level order, Nc coarse depths
  float* midp = tcs_e + CAP;                           // [Nc+1]: Nc-1 mid points + 2 out-of-range slots
  float* fs_e = midp + nc1;                            // [32*P] level order: the first 32P-1 sorted fine samples
  float* outs = fs_e + 32 * P;                         // [Nc+Nf]
  const int64_t nwarps = (int64_t)gridDim.x * kSampWarps;
  for (int i = lane; i < CAP; i += 32) { cdf_e[i] = CUDART_INF_F; tcs_e[i] = CUDART_INF_F; }
  __syncwarp();
  // 4 consecutive draws per lane (one 16-byte load / one Philox block) when the row length allows it
  const bool vec4 = (P >= 4) && (Nf % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.u) & 15) == 0);

  // The trip count depends on blockIdx only: the compiler can see that the body is not under divergent control
  // flow, so the shuffles are bare SHFLs (a per-warp bound wrapped each one in WARPSYNC / collective barriers).
  // Warps past the end redo the last ray and skip the stores.
  for (int64_t base = (int64_t)blockIdx.x * kSampWarps; base < a.R; base += nwarps) {
    const bool live = base + wib < a.R;
    const int64_t ray = live ? base + wib : a.R - 1;
    // ---- pdf / cdf (utils.py:63-69) -----------------------------------------------------------
    float w[NCB], tc[NCB];
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < NCB; ++j) {
      const int i = j * 32 + lane;
      const bool valid = i < Nc;
      w[j] = valid ? __fadd_rn(ld_stream(a.weights + ray * Nc + i), 1e-5f) : 0.f;   // weights += 1e-5
      tc[j] = (valid && a.t_coarse != nullptr) ? ld_stream(a.t_coarse + ray * Nc + i) : 0.f;
      part += w[j];
    }
    if (a.cdf_in == nullptr && a.sequential) {
      // TF-CPU / NumPy order: normaliser and cdf summed strictly left to right (KNERF_SCAN_SEQUENTIAL)
#pragma unroll
      for (int j = 0; j < NCB; ++j) {
        const int i = j * 32 + lane;
        if (i < Nc) outs[i] = w[j];
      }
      __syncwarp();
      if (lane == 0) {
        float tot = 0.f;
        for (int i = 0; i < Nc; ++i) tot = __fadd_rn(tot, outs[i]);
        float run = 0.f;
        cdf[0] = 0.f;
        for (int i = 0; i < Nc; ++i) {
          run = __fadd_rn(run, __fdiv_rn(outs[i], tot));
          cdf[i + 1] = run;
        }
      }
      __syncwarp();
    }
    const float total = warp_sum(part);
    if (a.cdf_in == nullptr && !a.sequential) {
      float carry = 0.f;
#pragma unroll
      for (int j = 0; j < NCB; ++j) {
        const int i = j * 32 + lane;
        const float pdf = __fdiv_rn(w[j], total);
        const float incl = warp_scan_add(pdf, lane);
        if (i < Nc) cdf[i + 1] = carry + incl;
        carry += __shfl_sync(kFullMask, incl, 31);
      }
      if (lane == 0) cdf[0] = 0.f;
    } else if (a.cdf_in != nullptr) {
      for (int i = lane; i <= Nc; i += 32) cdf[i] = a.cdf_in[ray * (Nc + 1) + i];
    }
    // ---- mid points (nerf.py:182-183) + the two out-of-range gather slots (App. C-1) -----------
    if (a.t_coarse != nullptr) {
#pragma unroll
      for (int j = 0; j < NCB; ++j) {
        const int i = j * 32 + lane;
        float tn = __shfl_down_sync(kFullMask, tc[j], 1);
        if (j + 1 < NCB) {
          const float first_next = __shfl_sync(kFullMask, tc[(j + 1 < NCB) ? j + 1 : j], 0);
          if (lane == 31) tn = first_next;
        }
        if (i < Nc - 1) midp[i] = 0.5f * __fadd_rn(tn, tc[j]);
        if (i < Nc) tcs_e[eyt_slot(i, LOGC)] = tc[j];
      }
    } else {
      for (int i = lane; i < Nc - 1; i += 32) midp[i] = a.mid_points[ray * (Nc - 1) + i];
    }
    __syncwarp();
    if (lane < 2) midp[Nc - 1 + lane] = (a.oob_mode == KNERF_OOB_CLAMP && Nc >= 2) ? midp[Nc - 2] : 0.f;
    for (int i = lane; i <= Nc; i += 32) {
      const float c = cdf[i];
      cdf_e[eyt_slot(i, LOGC)] = c;
      if (a.cdf_out != nullptr && live) a.cdf_out[ray * (Nc + 1) + i] = c;
    }
    __syncwarp();

    // ---- inverse-CDF samples (utils.py:72-94) ---------------------------------------------------
    float s[P], uu[P];
    int fidx[P];
#pragma unroll
    for (int r = 0; r < P; ++r) fidx[r] = vec4 ? ((r >> 2) * 128 + lane * 4 + (r & 3)) : (r * 32 + lane);
    if (vec4) {
#pragma unroll
      for (int q = 0; q < (P >= 4 ? P / 4 : 1); ++q) {
        const int f0 = q * 128 + lane * 4;
        float4 v4 = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
        if (f0 < Nf) {
          const int64_t e0 = ray * Nf + f0;                 // e0 % 4 == 0
          if (a.u != nullptr) {
            v4 = __ldcs(reinterpret_cast<const float4*>(a.u + e0));
          } else {
            uint32_t rr[4];
            philox4x32(a.seed, (uint64_t)e0 >> 2, rr);     // == philox_uniform(seed, e0 + 0..3)
            v4 = make_float4(u01(rr[0]), u01(rr[1]), u01(rr[2]), u01(rr[3]));
          }
        }
        if (4 * q + 3 < P) { uu[4 * q] = v4.x; uu[4 * q + 1] = v4.y; uu[4 * q + 2] = v4.z; uu[4 * q + 3] = v4.w; }
      }
    } else {
#pragma unroll
      for (int r = 0; r < P; ++r) {
        const int64_t e = ray * Nf + fidx[r];
        uu[r] = CUDART_NAN_F;
        if (fidx[r] < Nf) uu[r] = (a.u != nullptr) ? ld_stream(a.u + e) : philox_uniform(a.seed, (uint64_t)e);
      }
    }
    int idx[P];                                       // searchsorted(cdf, u, side='right') = #{cdf_j <= u}
    eyt_count<LOGC, P, true>(cdf_e, uu, idx);
    int oob = 0;
#pragma unroll
    for (int r = 0; r < P; ++r) {
      s[r] = CUDART_INF_F;
      if (fidx[r] < Nf) {
        const int below = max(idx[r] - 1, 0);         // utils.py:78
        const int above = min(idx[r], Nc);            // utils.py:79 (cdf.shape[-1]-1 == Nc)
        const float c0 = cdf[below], c1 = cdf[above];
        const float m0 = midp[below], m1 = midp[above];
        if (above >= Nc - 1) oob = 1;
        float den = __fsub_rn(c1, c0);
        if (den < 1e-5f) den = 1.0f;                  // utils.py:91
        const float tt = __fdiv_rn(__fsub_rn(uu[r], c0), den);
        s[r] = __fadd_rn(m0, __fmul_rn(tt, __fsub_rn(m1, m0)));   // utils.py:93-94
        const int64_t e = ray * Nf + fidx[r];
        if (a.samples != nullptr && live) a.samples[e] = s[r];
        if (a.indices != nullptr && live) a.indices[e] = idx[r];
      }
    }
    if (a.oob_mode == KNERF_OOB_COUNT && a.oob_count != nullptr) {
      const unsigned any = __ballot_sync(kFullMask, oob);
      if (lane == 0 && any && live) atomicAdd(a.oob_count, 1);   // rays (not samples) with an out-of-range gather
    }
    if (a.t_sorted == nullptr) { __syncwarp(); continue; }

    // ---- sort(concat(t_coarse, samples)) (nerf.py:190-191) -------------------------------------
    bitonic_sort_regs<P>(s, lane);
#pragma unroll
    for (int r = 0; r < P; ++r) {
      const int e = lane * P + r;
      if (e < 32 * P - 1) fs_e[eyt_slot(e, LOGF)] = s[r];
    }
    const float fs_last = __shfl_sync(kFullMask, s[P - 1], 31);   // sorted position 32P-1 (not in the tree)
    __syncwarp();
    {
      int cnt[P];                                     // coarse entries <= v (a fine sample goes after equal coarse ones)
      eyt_count<LOGC, P, true>(tcs_e, s, cnt);
#pragma unroll
      for (int r = 0; r < P; ++r) {
        const int e = lane * P + r;
        if (e < Nf) outs[e + cnt[r]] = s[r];
      }
    }
    {
      int cnt[NCB];                                   // fine entries < v
      eyt_count<LOGF, NCB, false>(fs_e, tc, cnt);
#pragma unroll
      for (int j = 0; j < NCB; ++j) {
        const int i = j * 32 + lane;
        cnt[j] += (fs_last < tc[j]) ? 1 : 0;          // all 32P-1 tree entries smaller is implied when the last one is
        if (i < Nc) outs[i + cnt[j]] = tc[j];
      }
    }
    __syncwarp();
    if (live)
      for (int i = lane; i < Nc + Nf; i += 32) a.t_sorted[ray * (Nc + Nf) + i] = outs[i];
    __syncwarp();
  }
}

}  // namespace knerf

using namespace knerf;

template <int NCB>
static int launch_sampler(const SampleArgs& a, int P, int grid, size_t smem, cudaStream_t st) {
#define KN_LAUNCH_P(PP)                                                                             \
  case PP:                                                                                          \
    KN_CUDA(cudaFuncSetAttribute(sample_fine_kernel<NCB, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 (int)smem));                                                       \
    sample_fine_kernel<NCB, PP><<<grid, kSampWarps * 32, smem, st>>>(a);                            \
    break;
  switch (P) {
    KN_LAUNCH_P(1) KN_LAUNCH_P(2) KN_LAUNCH_P(4) KN_LAUNCH_P(8) KN_LAUNCH_P(16)
    default: return fail(KNERF_ERR_INVALID, "knerf_sample_fine: Nf too large");
  }
#undef KN_LAUNCH_P
  KN_LAUNCH_CHECK();
  return KNERF_OK;
}

extern "C" int knerf_sample_fine(const float* t_coarse, const float* mid_points, const float* weights,
                                 const float* u, uint64_t seed, const float* cdf_in, int64_t R, int Nc,
                                 int Nf, int oob_mode, float* t_sorted, float* samples, int32_t* indices,
                                 float* cdf_out, int32_t* oob_count, void* stream) {
  KN_CHECK_ARG(weights != nullptr && R >= 0, "knerf_sample_fine: null weights");
  KN_CHECK_ARG((t_coarse != nullptr) != (mid_points != nullptr),
               "knerf_sample_fine: pass either t_coarse or mid_points");
  KN_CHECK_ARG(Nc >= 2 && Nc <= 256 && Nf >= 1 && Nf <= 512, "knerf_sample_fine: Nc=%d (2..256) Nf=%d (1..512)", Nc, Nf);
  KN_CHECK_ARG(t_sorted == nullptr || t_coarse != nullptr, "knerf_sample_fine: t_sorted needs t_coarse");
  const int sequential = (oob_mode & KNERF_SCAN_SEQUENTIAL) ? 1 : 0;
  oob_mode &= ~KNERF_SCAN_SEQUENTIAL;
  KN_CHECK_ARG(oob_mode >= 0 && oob_mode <= 2, "knerf_sample_fine: bad oob_mode %d", oob_mode);
  if (R == 0) return KNERF_OK;
  int P = 1;
  while (32 * P < Nf) P <<= 1;
  const int ncb = (Nc + 31) / 32;
  SampleArgs a{t_coarse, mid_points, weights, u, seed, cdf_in, R, Nc, Nf, oob_mode, sequential,
               t_sorted, samples, indices, cdf_out, oob_count, 0};
  const int nc1 = (Nc + 1 + 3) & ~3;
  const int cap = 64 * (ncb <= 2 ? ncb : (ncb <= 4 ? 4 : 8));
  a.smem_per_warp = 2 * nc1 + 2 * cap + 32 * P + ((Nc + Nf + 3) & ~3);
  const size_t smem = (size_t)a.smem_per_warp * kSampWarps * sizeof(float);
  const int grid = (int)std::min<int64_t>(cdiv(R, kSampWarps), (int64_t)kNumSMs * 12);
  cudaStream_t st = (cudaStream_t)stream;
  switch (ncb) {
    case 1: return launch_sampler<1>(a, P, grid, smem, st);
    case 2: return launch_sampler<2>(a, P, grid, smem, st);
    case 3: case 4: return launch_sampler<4>(a, P, grid, smem, st);
    default: return launch_sampler<8>(a, P, grid, smem, st);
  }
}
