// Microbenchmark: raw issue/execute rate of tcgen05.mma from shared-memory operands in the library's chunk-major
// (SWIZZLE_NONE) layout, 1-CTA (M=128) and cta_group::2 (M=256), with and without a tcgen05.commit per K=32 stage.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I keras_nerf_b200/csrc -o benchmarks/micro/umma_rate benchmarks/micro/umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace knerf::tc;

constexpr int kA = 128 * 256 * 2;      // one A tile [128 x 256] bf16
constexpr int kStage = 256 * 32 * 2;   // one full-N weight stage (K = 32)

template <bool TWO, bool MN = false>
__global__ void __launch_bounds__(128) rate_kernel(int n_stages, int N, int commit_every, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_done, bar_full, bar_stage[8];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t cta = TWO ? cluster_ctarank() : 0u;
  for (int i = tid; i < (kA + 4 * kStage) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (tid == 0) {
    mbar_init(&bar_done, 1);
    mbar_init(&bar_full, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&bar_stage[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) { if (TWO) tmem_alloc_2cta<512>(&tmem_base_s); else tmem_alloc<512>(&tmem_base_s); }
  fence_async_smem();
  tc_fence_before();
  if (TWO) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (cta == 0 && tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(TWO ? 256 : 128, N, MN ? 1 : 0, MN ? 1 : 0);
    const uint32_t chunk_b = (uint32_t)(TWO ? N / 2 : N) * 16;
    t0 = clock64();
    for (int s = 0; s < n_stages; ++s) {
      const uint32_t a_base = smem_u32(smem) + (MN ? 0 : (s % 8) * 4 * 2048);
      const uint32_t b_base = smem_u32(smem) + (MN ? 32768 : kA + (s % 4) * kStage);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        // MN: the weight-gradient form (tc_ptx.cuh): K = rows of the blobs, LBO = 128 B (next 8 k), SBO = 2 KB (next 8 m|n)
        const uint64_t da = MN ? umma_smem_desc(a_base + j * 256, 128, 2048) : umma_smem_desc(a_base + j * 2 * 2048, 2048, 128);
        const uint64_t db = MN ? umma_smem_desc(b_base + j * 256, 128, 2048) : umma_smem_desc(b_base + j * 2 * chunk_b, chunk_b, 128);
        if (TWO) umma_bf16_2cta(tmem + (s & 256), da, db, idesc, s > 0);
        else umma_bf16(tmem + (s & 256), da, db, idesc, s > 0);
      }
      // commit_every is a bit mask here: 1 = commit per stage, 2 = wait on an (already complete) barrier per stage,
      // 4 = tcgen05.fence::after_thread_sync per stage, 8 = commit every 2nd stage only
      if (commit_every & 2) mbar_wait(&bar_full, 1);
      if (commit_every & 4) tc_fence_after();
      if ((commit_every & 1) || ((commit_every & 8) && (s & 1))) {
        if (TWO) umma_commit_2cta(&bar_stage[s % 8], 3); else umma_commit(&bar_stage[s % 8]);
      }
    }
    t1 = clock64();
    if (TWO) umma_commit_2cta(&bar_done, 3); else umma_commit(&bar_done);
  }
  if (TWO) mbar_wait_cluster(&bar_done, 0); else mbar_wait(&bar_done, 0);
  t2 = clock64();
  tc_fence_after();
  if (cta == 0 && tid == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  tc_fence_before();
  if (TWO) cluster_sync_all(); else __syncthreads();
  if (warp == 0) { if (TWO) tmem_dealloc_2cta<512>(tmem); else tmem_dealloc<512>(tmem); }
}

template <bool TWO, bool MN = false>
void run(int grid, int n_stages, int N, int commit_every) {
  long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  cudaMemset(d, 0, 148 * 2 * sizeof(long long));
  const size_t smem = kA + 4 * kStage;
  cudaFuncSetAttribute(rate_kernel<TWO, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  if (TWO) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<TWO, MN>, n_stages, N, commit_every, d);
    if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long h[296];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double issue = 0, total = 0; int n = 0;
  for (int b = 0; b < grid; b += (TWO ? 2 : 1)) { issue += h[b * 2]; total += h[b * 2 + 1]; ++n; }
  printf("%s grid %3d N %3d stages %5d mode %2d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)\n",
         MN ? "1-CTA M=128 MN-major A and B" : TWO ? "2-CTA M=256" : "1-CTA M=128", grid, N, n_stages, commit_every, issue / n / (2.0 * n_stages),
         total / n / (2.0 * n_stages), N / 2);
  cudaFree(d);
}

int main() {
  for (int mode : {0, 1, 2, 4, 3, 7, 8, 14}) {
    run<false>(148, 4096, 256, mode);
    run<true>(148, 4096, 256, mode);
  }
  run<true>(148, 4096, 128, 7);
  // weight-gradient form: both operands MN-major, N = 128 (floor 64) and N = 256
  run<false>(148, 4096, 128, 0);
  run<false, true>(148, 4096, 128, 0);
  run<false, true>(148, 4096, 256, 0);
  return 0;
}
