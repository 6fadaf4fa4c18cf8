// Microbenchmark: dependent chains of Fp Montgomery products per thread, carry-chain 32-bit limbs
// (csrc/bigint.cuh) vs carry-free 30-bit limbs with 64-bit column accumulators.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../plonk_b200/csrc/bigint.cuh"
#include "fp_dfma52.cuh"
using namespace pb;

__global__ void k_chain32(Fp* io, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  Fp a = io[i], b = io[i + 1];
  for (int k = 0; k < iters; k++) { a = Fp::mul_imad(a, b); b = Fp::mul_imad(b, a); }
  io[i] = a + b;
}

// IMAD product + DFMA Montgomery reduction (Field::mul_hybrid)
__global__ void k_chainhyb(Fp* io, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  Fp a = io[i], b = io[i + 1];
  for (int k = 0; k < iters; k++) { a = Fp::mul_hybrid(a, b); b = Fp::mul_hybrid(b, a); }
  io[i] = a + b;
}
// All-FP64 product on 52-bit limbs (fp_dfma52.cuh)
__global__ void k_chain52(pb52::F52* io, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  pb52::F52 a = io[i], b = io[i + 1];
  for (int k = 0; k < iters; k++) { a = pb52::mul52(a, b); b = pb52::mul52(b, a); }
#pragma unroll
  for (int j = 0; j < 8; j++) a.l[j] += b.l[j];
  io[i] = a;
}
// Pipe probes: dependent DFMA.RZ chains (8 per thread), dependent IMAD.WIDE.U32.X chains, and both
// in the same thread, to see whether the FP64 pipe and the integer multiply pipe overlap.
template <int MODE>  // bit 0 = DFMA, bit 1 = IMAD.WIDE.X, bit 2 = 64-bit IADD3 pairs
__global__ void k_pipes(double* io, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double d[8];
  uint32_t u[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { d[k] = io[i + k]; u[k] = (uint32_t)__double_as_longlong(d[k]) | 1u; }
  const double m = io[i + 9];
  unsigned long long w[4];
#pragma unroll
  for (int k = 0; k < 4; k++) w[k] = (unsigned long long)__double_as_longlong(d[k]) * 3u;
  for (int it = 0; it < iters; it++) {
    if (MODE & 4) {  // 4 three-input 64-bit adds = 8 IADD3 / IADD3.X
#pragma unroll
      for (int k = 0; k < 4; k++) w[k] += w[(k + 1) & 3] + w[(k + 2) & 3];
    }
    if (MODE & 1) {
#pragma unroll
      for (int k = 0; k < 8; k++) d[k] = __fma_rz(d[k], m, d[(k + 1) & 7]);
    }
    if (MODE & 2) {
      pb::mad_pair<false, true>(u[0], u[1], u[6], u[7], u[0], u[1]);
      pb::mad_pair<true, true>(u[2], u[3], u[6], u[7], u[2], u[3]);
      pb::mad_pair<true, true>(u[4], u[5], u[0], u[7], u[4], u[5]);
      pb::mad_pair<true, false>(u[6], u[7], u[2], u[5], u[6], u[7]);
      pb::mad_pair<false, true>(u[0], u[1], u[4], u[7], u[0], u[1]);
      pb::mad_pair<true, true>(u[2], u[3], u[6], u[5], u[2], u[3]);
      pb::mad_pair<true, true>(u[4], u[5], u[2], u[7], u[4], u[5]);
      pb::mad_pair<true, false>(u[6], u[7], u[0], u[3], u[6], u[7]);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += d[k] + (double)u[k] + (double)w[k & 3];
  io[i] = s;
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
  pb52::F52* b52; cudaMalloc(&b52, sizeof(pb52::F52) * maxthreads);
  {  // limbs = 2^51 + small (exact integers below 2^52)
    double* hb = (double*)malloc(sizeof(pb52::F52) * maxthreads);
    for (size_t k = 0; k < (size_t)maxthreads * 8; k++) hb[k] = (double)((1ull << 51) + (k * 2654435761ull) % (1ull << 50));
    cudaMemcpy(b52, hb, sizeof(pb52::F52) * maxthreads, cudaMemcpyHostToDevice); free(hb);
  }
  {
    double* pd; cudaMalloc(&pd, 8 * (maxthreads + 16)); cudaMemset(pd, 0x3f, 8 * (maxthreads + 16));
    for (int threads_per_sm : {256, 512, 1024}) {
      int blocks = 148 * threads_per_sm / 128;
      double ops = (double)blocks * 128 * 2000 * 8;
      float t1 = run(k_pipes<1>, pd, blocks, 128, 2000);
      float t2 = run(k_pipes<2>, pd, blocks, 128, 2000);
      float t3 = run(k_pipes<3>, pd, blocks, 128, 2000);
      float t4 = run(k_pipes<4>, pd, blocks, 128, 2000);
      float t5 = run(k_pipes<5>, pd, blocks, 128, 2000);
      float t6 = run(k_pipes<6>, pd, blocks, 128, 2000);
      printf("pipes threads/SM %4d: 8 IADD3 alone %.3f ms | with 8 DFMA %.3f ms | with 8 IMAD.WIDE.X %.3f ms\n", threads_per_sm, t4, t5, t6);
      printf("pipes threads/SM %4d: DFMA.RZ %.2f T/s | IMAD.WIDE.X %.2f T/s | both in one thread: %.2f T/s each (%.3f ms vs %.3f + %.3f)\n",
             threads_per_sm, ops / t1 / 1e9, ops / t2 / 1e9, ops / t3 / 1e9, t3, t1, t2);
    }
  }
  for (int threads_per_sm : {128, 256, 384, 512, 768, 1024}) {
    int blocks = 148 * threads_per_sm / 128;
    float m32 = run(k_chain32, b32, blocks, 128, iters);
    float m30 = run(k_chain30, b30, blocks, 128, iters);
    float m30b = run(k_chain30v2, b30, blocks, 128, iters);
    printf("   v2 30-bit: %8.3f ms (%6.2f G mul/s, %.2f us/mul/thread)\n", m30b, (double)blocks * 128 * iters * 2 / m30b / 1e6, m30b * 1e3 / (iters * 2));
    double muls = (double)blocks * 128 * iters * 2;
    float m52 = run(k_chain52, b52, blocks, 128, iters);
    printf("   all-FP64 52-bit limbs: %8.3f ms (%6.2f G mul/s, %.2f us/mul/thread)\n", m52, muls / m52 / 1e6, m52 * 1e3 / (iters * 2));
    float mh = run(k_chainhyb, b32, blocks, 128, iters);
    printf("   hybrid IMAD product + DFMA reduction: %8.3f ms (%6.2f G mul/s, %.2f us/mul/thread)\n", mh, muls / mh / 1e6, mh * 1e3 / (iters * 2));
    printf("threads/SM %4d: 32-bit carry chains %8.3f ms (%6.2f G mul/s, %.2f us/mul/thread) | 30-bit lazy %8.3f ms (%6.2f G mul/s, %.2f us/mul/thread)\n",
           threads_per_sm, m32, muls / m32 / 1e6, m32 * 1e3 / (iters * 2), m30, muls / m30 / 1e6, m30 * 1e3 / (iters * 2));
  }
  return 0;
}
