// Microbenchmark: does epilogue traffic (tcgen05.ld by 8 warps, optional bf16 pack + st.shared) slow the tensor
// pipe?  One thread issues back-to-back M=128 N=256 K=16 MMAs into TMEM columns [0,256) while 8 warps read
// columns [256,512) in a loop.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I keras_nerf_b200/csrc -o benchmarks/micro/umma_interf benchmarks/micro/umma_interf.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace knerf::tc;

constexpr int kA = 128 * 256 * 2;
constexpr int kStage = 256 * 32 * 2;

__global__ void __launch_bounds__(320) interf_kernel(int n_stages, int bg_mode, int bg_iters, long long* out, int tail) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_done, bar_a;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (2 * kA + 4 * kStage) / 4; i += 320) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (tid == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_a, (tail == 2) ? 256 : 8); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    if (lane == 0 && n_stages > 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
      const long long t0 = clock64();
      for (int s = 0; s < n_stages; ++s) {
        const uint32_t a_base = smem_u32(smem) + (s % 8) * 4 * 2048;
        const uint32_t b_base = smem_u32(smem) + 2 * kA + (s % 4) * kStage;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint64_t da = umma_smem_desc(a_base + j * 2 * 2048, 2048, 128);
          const uint64_t db = umma_smem_desc(b_base + j * 2 * 4096, 4096, 128);
          umma_bf16(tmem, da, db, idesc, s > 0);
        }
      }
      umma_commit(&bar_done);
      mbar_wait(&bar_done, 0);
      out[blockIdx.x * 4] = clock64() - t0;
    }
  } else if (warp >= 2 && bg_mode) {
    const int q = warp & 3, h = (warp - 2) >> 2, r = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint8_t* hs = smem + kA;   // second A-sized buffer: the "next operand"
    const long long t0 = clock64();
    uint32_t sink = 0;
    for (int it = 0; it < bg_iters; ++it) {
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        const int col0 = h * 128 + g * 32;
        uint32_t v[32];
        tmem_ld32_issue(tmem + lane_base + 256 + col0, v);
        tmem_ld32_wait(v);
        if (bg_mode >= 2) {
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            const float* x = reinterpret_cast<const float*>(&v[c8 * 8]);
            const uint4 pk = make_uint4(pack_bf16x2_relu(x[0], x[1]), pack_bf16x2_relu(x[2], x[3]),
                                        pack_bf16x2_relu(x[4], x[5]), pack_bf16x2_relu(x[6], x[7]));
            *reinterpret_cast<uint4*>(hs + ((col0 >> 3) + c8) * 2048 + r * 16) = pk;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) sink ^= v[i];
        }
      }
      // tail variants: how the compute warps publish "next A operand ready"
      if (tail == 1) { fence_async_smem(); }
      else if (tail == 2) { tc_fence_before(); fence_async_smem(); mbar_arrive(&bar_a); }
      else if (tail == 3) { tc_fence_before(); fence_async_smem(); __syncwarp(); if (lane == 0) mbar_arrive(&bar_a); }
      else if (tail == 4) { tc_fence_before(); fence_async_smem(); __syncwarp(); if (lane == 0) mbar_arrive_cluster(&bar_a, 0); }
      else if (tail == 5) {
        tc_fence_before(); fence_async_smem(); __syncwarp();
        if (lane == 0) asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                                    "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(&bar_a)), "r"(0) : "memory");
      }
      else if (tail == 6) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&bar_a); }
    }
    if (sink == 0x12345678u) out[1000] = sink;
    if (tid == 64) out[blockIdx.x * 4 + 1] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

void run(int n_stages, int bg_mode, int bg_iters, int tail = 0) {
  long long* d;
  cudaMalloc(&d, 8192 * sizeof(long long));
  cudaMemset(d, 0, 8192 * sizeof(long long));
  const size_t smem = 2 * kA + 4 * kStage;
  cudaFuncSetAttribute(interf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) {
    interf_kernel<<<148, 320, smem>>>(n_stages, bg_mode, bg_iters, d, tail);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long h[148 * 4];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double mma = 0, bg = 0;
  for (int b = 0; b < 148; ++b) { mma += h[b * 4]; bg += h[b * 4 + 1]; }
  printf("tail %d mma stages %5d bg_mode %d iters %4d: %.1f cyc/MMA; epilogue pass (128x256 tile) %.0f cyc\n", tail, n_stages, bg_mode,
         bg_iters, n_stages ? mma / 148 / (2.0 * n_stages) : 0.0, bg_iters && bg_mode ? bg / 148 / bg_iters : 0.0);
  cudaFree(d);
}

int main() {
  for (int tail = 0; tail <= 6; ++tail) run(4096, 2, 600, tail);
  return 0;
}
