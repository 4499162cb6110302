// Microbenchmark: tcgen05.mma rate for the weight-gradient operand form (both operands MN-major, SWIZZLE_NONE,
// M = 128) against the K-major form, clean unrolled issue loop (8 MMAs per iteration, descriptors precomputed).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I keras_nerf_b200/csrc -o benchmarks/micro/umma_mn benchmarks/micro/umma_mn.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace knerf::tc;

template <bool MN, int N>
__global__ void __launch_bounds__(128) mn_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (32768 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (tid == 0) { mbar_init(&bar_done, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, MN ? 1 : 0, MN ? 1 : 0);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 32768;
    // MN-major: K = 128 rows of the blobs, 16 per MMA (256 B apart); K-major: K chunks 4 KB / N*32 B apart
    const uint64_t da0 = MN ? umma_smem_desc(a, 128, 2048) : umma_smem_desc(a, 2048, 128);
    const uint64_t db0 = MN ? umma_smem_desc(b, 128, 2048) : umma_smem_desc(b, N * 16, 128);
    const uint32_t sa = MN ? 16 : (2 * 2048) >> 4, sb = MN ? 16 : (2 * N * 16) >> 4;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_bf16(tmem, da0 + k * sa, db0 + k * sb, idesc, 1u);
    }
    umma_commit(&bar_done);
    mbar_wait(&bar_done, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <bool MN, int N>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const size_t smem = 32768 + 65536;
  cudaFuncSetAttribute(mn_kernel<MN, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2048;
  for (int rep = 0; rep < 2; ++rep) {
    mn_kernel<MN, N><<<148, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double t = 0;
  for (int i = 0; i < 148; ++i) t += h[i];
  printf("%-28s N %3d: %.1f cycles per MMA (pipe floor %d)\n", name, N, t / 148 / (8.0 * iters), N / 2);
  cudaFree(d);
}

int main() {
  run<false, 128>("K-major A and B");
  run<false, 256>("K-major A and B");
  run<true, 128>("MN-major A and B (wgrad)");
  run<true, 256>("MN-major A and B (wgrad)");
  run<true, 64>("MN-major A and B (wgrad)");
  run<true, 16>("MN-major A and B (wgrad)");
  return 0;
}
