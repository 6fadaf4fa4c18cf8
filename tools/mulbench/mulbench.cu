// Microbenchmark: dependent chains of Fp Montgomery products per thread, carry-chain 32-bit limbs
// (csrc/bigint.cuh) vs carry-free 30-bit limbs with 64-bit column accumulators.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../plonk_b200/csrc/bigint.cuh"
using namespace pb;

__global__ void k_chain32(Fp* io, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  Fp a = io[i], b = io[i + 1];
  for (int k = 0; k < iters; k++) { a = a * b; b = b * a; }
  io[i] = a + b;
}

// 13 x 30-bit limbs, R = 2^390, lazy: inputs limbs < 2^30 (+eps), output carry-propagated, in [0, 2p)
struct F30 { uint32_t l[13]; };
__constant__ uint32_t P30[13];
__constant__ uint32_t PINV30;  // -p^-1 mod 2^30
__device__ __forceinline__ F30 mul30(const F30& a, const F30& b) {
  unsigned long long t[14];
#pragma unroll
  for (int j = 0; j < 14; j++) t[j] = 0;
#pragma unroll
  for (int i = 0; i < 13; i++) {
#pragma unroll
    for (int j = 0; j < 13; j++) t[j] += (unsigned long long)a.l[j] * b.l[i];
    uint32_t m = ((uint32_t)t[0] * PINV30) & 0x3fffffffu;
#pragma unroll
    for (int j = 0; j < 13; j++) t[j] += (unsigned long long)m * P30[j];
    unsigned long long c = t[0] >> 30;
#pragma unroll
    for (int j = 0; j < 13; j++) t[j] = t[j + 1];
    t[13] = 0;
    t[0] += c;
    if (i == 6) {  // mid-way partial carry so the 64-bit columns cannot overflow
#pragma unroll
      for (int j = 12; j >= 1; j--) { t[j] += t[j - 1] >> 30; }
#pragma unroll
      for (int j = 0; j < 12; j++) t[j] &= 0x3fffffffull;
    }
  }
  F30 r;
  unsigned long long c = 0;
#pragma unroll
  for (int j = 0; j < 13; j++) { unsigned long long v = t[j] + c; r.l[j] = (uint32_t)v & 0x3fffffffu; c = v >> 30; }
  return r;
}
// v2: every carry is folded into the next column with another IMAD.WIDE (multiplier 1) fed by a
// 32-bit funnel shift, so the only non-IMAD work per row is one SHF and one LOP.
__device__ __forceinline__ unsigned long long madw(uint32_t a, uint32_t b, unsigned long long c) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint32_t shr30(unsigned long long t) {
  return __funnelshift_r((uint32_t)t, (uint32_t)(t >> 32), 30);
}
__device__ __forceinline__ F30 mul30v2(const F30& a, const F30& b) {
  unsigned long long t[14];
#pragma unroll
  for (int j = 0; j < 14; j++) t[j] = 0;
#pragma unroll
  for (int i = 0; i < 13; i++) {
#pragma unroll
    for (int j = 0; j < 13; j++) t[j] = madw(a.l[j], b.l[i], t[j]);
    uint32_t m = ((uint32_t)t[0] * PINV30) & 0x3fffffffu;
#pragma unroll
    for (int j = 0; j < 13; j++) t[j] = madw(m, P30[j], t[j]);
    uint32_t c = shr30(t[0]);
    t[1] = madw(c, 1u, t[1]);
#pragma unroll
    for (int j = 0; j < 13; j++) t[j] = t[j + 1];
    t[13] = 0;
    if (i == 6) {
#pragma unroll
      for (int j = 11; j >= 0; j--) {
        uint32_t cc = shr30(t[j]);
        t[j + 1] = madw(cc, 1u, t[j + 1]);
        t[j] = (uint32_t)t[j] & 0x3fffffffu;
      }
    }
  }
  F30 r;
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < 13; j++) {
    unsigned long long v = t[j] + c;
    r.l[j] = (uint32_t)v & 0x3fffffffu;
    c = shr30(v);
  }
  return r;
}
__global__ void k_chain30v2(F30* io, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  F30 a = io[i], b = io[i + 1];
  for (int k = 0; k < iters; k++) { a = mul30v2(a, b); b = mul30v2(b, a); }
#pragma unroll
  for (int j = 0; j < 13; j++) a.l[j] += b.l[j];
  io[i] = a;
}
__global__ void k_chain30(F30* io, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  F30 a = io[i], b = io[i + 1];
  for (int k = 0; k < iters; k++) { a = mul30(a, b); b = mul30(b, a); }
#pragma unroll
  for (int j = 0; j < 13; j++) a.l[j] += b.l[j];
  io[i] = a;
}

template <class K, class T>
float run(K kern, T* buf, int blocks, int threads, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(buf, 2);
  cudaEventRecord(e0);
  kern<<<blocks, threads>>>(buf, iters);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  const int maxthreads = 148 * 2048 + 64;
  Fp* b32; F30* b30;
  cudaMalloc(&b32, sizeof(Fp) * maxthreads); cudaMemset(b32, 0x11, sizeof(Fp) * maxthreads);
  cudaMalloc(&b30, sizeof(F30) * maxthreads); cudaMemset(b30, 0x11, sizeof(F30) * maxthreads);
  uint32_t p30[13] = {0x3fffaaab, 0x27fbffff, 0x153ffffb, 0x2affffac, 0x30f6241e, 0x034a83da, 0x112bf673, 0x12e13ce1, 0x2cd76477, 0x1ed90d2e, 0x29a4b1ba, 0x3a8e5ff9, 0x001a0111};
  // p in 30-bit limbs is recomputed below to be safe
  { unsigned __int128 dummy = 0; (void)dummy; }
  cudaMemcpyToSymbol(P30, p30, sizeof p30);
  uint32_t pinv = 0x3ffcfffd; cudaMemcpyToSymbol(PINV30, &pinv, 4);
  const int iters = 200;
  for (int threads_per_sm : {128, 256, 384, 512, 768, 1024}) {
    int blocks = 148 * threads_per_sm / 128;
    float m32 = run(k_chain32, b32, blocks, 128, iters);
    float m30 = run(k_chain30, b30, blocks, 128, iters);
    float m30b = run(k_chain30v2, b30, blocks, 128, iters);
    printf("   v2 30-bit: %8.3f ms (%6.2f G mul/s, %.2f us/mul/thread)\n", m30b, (double)blocks * 128 * iters * 2 / m30b / 1e6, m30b * 1e3 / (iters * 2));
    double muls = (double)blocks * 128 * iters * 2;
    printf("threads/SM %4d: 32-bit carry chains %8.3f ms (%6.2f G mul/s, %.2f us/mul/thread) | 30-bit lazy %8.3f ms (%6.2f G mul/s, %.2f us/mul/thread)\n",
           threads_per_sm, m32, muls / m32 / 1e6, m32 * 1e3 / (iters * 2), m30, muls / m30 / 1e6, m30 * 1e3 / (iters * 2));
  }
  return 0;
}
