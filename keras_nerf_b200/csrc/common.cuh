// Shared helpers for libknerf (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>

#include "../../include/knerf.h"

namespace knerf {

// ---- thread-local error string -------------------------------------------------------------------
char* last_error_buffer();
int fail(int code, const char* fmt, ...);

#define KN_CHECK_ARG(cond, ...)                                              \
  do {                                                                       \
    if (!(cond)) return ::knerf::fail(KNERF_ERR_INVALID, __VA_ARGS__);       \
  } while (0)

#define KN_CUDA(expr)                                                                             \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::knerf::fail(KNERF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                   \
  } while (0)

void count_launch();
#define KN_LAUNCH_CHECK()          \
  do {                             \
    ::knerf::count_launch();       \
    KN_CUDA(cudaGetLastError());   \
  } while (0)

#define KN_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != KNERF_OK) return _s; \
  } while (0)

constexpr int kNumSMs = 148;  // B200

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

// ---- Philox4x32-10 (counter based; keyed by seed, counter = element index / 4) --------------------
__host__ __device__ inline void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
  uint32_t hi0 = __umulhi(M0, c[0]), hi1 = __umulhi(M1, c[2]);
#else
  uint32_t hi0 = (uint32_t)(((uint64_t)M0 * c[0]) >> 32), hi1 = (uint32_t)(((uint64_t)M1 * c[2]) >> 32);
#endif
  uint32_t lo0 = M0 * c[0], lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__host__ __device__ inline void philox4x32(uint64_t seed, uint64_t counter, uint32_t (&out)[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), 0x6b6e6572u /*"knerf"*/, 0u};
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// uniform in [0,1): 24 random bits * 2^-24 (same lattice as the test fixtures)
__host__ __device__ inline float u01(uint32_t bits) { return (float)(bits >> 8) * (1.0f / 16777216.0f); }

__host__ __device__ inline float philox_uniform(uint64_t seed, uint64_t index) {
  uint32_t r[4];
  philox4x32(seed, index >> 2, r);
  return u01(r[index & 3]);
}

// ---- warp helpers --------------------------------------------------------------------------------
constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// inclusive scans across the 32 lanes
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v += n;
  }
  return v;
}
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}

// streaming (read-once) loads: bypass L1 allocation
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// ---- model geometry (host) ------------------------------------------------------------------------
struct LayerDesc {
  int64_t w_off, b_off;  // offsets in floats into the flat parameter buffer
  int fan_in, fan_out;
  int k_h;               // rows of the kernel fed by the hidden activation (0 for layer 0)
  int k_x;               // rows fed by the skip / xyz encoding (dx for layer 0 and after a skip)
};

constexpr int kMaxLayers = 36;

struct Model {
  knerf_config cfg;
  int dx, dd, U, n_layers;
  int n_dense;                 // n_layers + 4
  LayerDesc L[kMaxLayers];     // hidden 0..n-1, then sigma, features, rgb_features, rgb
  bool head_skip;              // concat applied after the last hidden layer
  int64_t n_params;
};

int build_model(const knerf_config* cfg, Model* m);
bool is_flagship(const Model& m);  // the shapes the fused bf16 path implements = tc_chain_map succeeds (api.cu)
bool tc_chain_map(const Model& m, int chain_map[8]);

}  // namespace knerf
