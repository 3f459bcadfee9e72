// Probe for the shift-GEMM convolution (conv_shift.cu): can a SWIZZLE_128B K-major UMMA
// descriptor start at an arbitrary 128-byte ROW offset inside a larger, 1024-byte aligned
// shared-memory image whose 16-byte chunks were swizzled by ABSOLUTE row index?
//   variant 0: descriptor base_offset field = 0      (swizzle taken from address bits)
//   variant 1: descriptor base_offset field = off & 7
// Also exercises kind::f16 with FP16 operands (the fp32-parity mode splits fp32 into
// fp16 hi + 2^11-scaled fp16 lo).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o shift_probe shift_probe.cu
// Run:   ./shift_probe      prints one line per (offset, variant) and a final verdict.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../sgtapose_b200/csrc/umma.cuh"

using namespace sgta::umma;

constexpr int ROWS = 160, N = 64, BK = 64;

__device__ __forceinline__ uint64_t desc_sw128_off(uint32_t saddr, uint32_t base_off) {
  uint64_t d = smem_desc_sw128(saddr);
  d |= (uint64_t)(base_off & 7u) << 49;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_f16_f32(int M, int Nn) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(Nn >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(192) probe(const __half* __restrict__ A,   // [ROWS][64] row-major
                                             const __half* __restrict__ Bm,  // [N][64] row-major
                                             float* __restrict__ D, int off, int variant) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                 // ROWS * 128
  unsigned char* sB = smem + ROWS * 128;    // N * 128 (ROWS*128 is a multiple of 1024)
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 4) tmem_alloc<64>(&tmem_slot);
  for (int e = tid; e < ROWS * 8; e += blockDim.x) {
    int r = e >> 3, c = e & 7;
    *reinterpret_cast<uint4*>(sA + sw128_offset(r, c)) = *reinterpret_cast<const uint4*>(A + (size_t)r * 64 + c * 8);
  }
  for (int e = tid; e < N * 8; e += blockDim.x) {
    int r = e >> 3, c = e & 7;
    *reinterpret_cast<uint4*>(sB + sw128_offset(r, c)) = *reinterpret_cast<const uint4*>(Bm + (size_t)r * 64 + c * 8);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 160) {
    const uint32_t idesc = idesc_f16_f32(128, N);
    const uint64_t ad = desc_sw128_off(smem_u32(sA) + off * 128, variant ? (uint32_t)off : 0u);
    const uint64_t bd = smem_desc_sw128(smem_u32(sB));
#pragma unroll
    for (int k = 0; k < BK / 16; ++k) mma_bf16_ss(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, k > 0);
    mma_commit(&bar_mma);
  }
  if (tid < 128) {
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc<64>(tmem);
}

int main() {
  std::vector<float> Af(ROWS * 64), Bf(N * 64);
  std::vector<__half> Ah(ROWS * 64), Bh(N * 64);
  srand(2);
  for (int i = 0; i < ROWS * 64; ++i) { Ah[i] = __float2half((rand() % 2001 - 1000) / 1000.f); Af[i] = __half2float(Ah[i]); }
  for (int i = 0; i < N * 64; ++i) { Bh[i] = __float2half((rand() % 2001 - 1000) / 1000.f); Bf[i] = __half2float(Bh[i]); }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, Ah.size() * 2); cudaMalloc(&dB, Bh.size() * 2); cudaMalloc(&dD, sizeof(float) * 128 * N);
  cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice);
  size_t smem = ROWS * 128 + N * 128 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int ok[2] = {1, 1};
  std::vector<float> Dh(128 * N);
  for (int variant = 0; variant < 2; ++variant)
    for (int off = 0; off <= 19; ++off) {
      cudaMemset(dD, 0, sizeof(float) * 128 * N);
      probe<<<1, 192, smem>>>(dA, dB, dD, off, variant);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d off %d CUDA error %s\n", variant, off, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(Dh.data(), dD, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost);
      double maxerr = 0, maxref = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)Af[(m + off) * 64 + k] * Bf[n * 64 + k];
          maxerr = fmax(maxerr, fabs(ref - Dh[m * N + n]));
          maxref = fmax(maxref, fabs(ref));
        }
      bool pass = maxerr < 1e-3 * maxref;
      if (!pass) ok[variant] = 0;
      printf("variant %d off %2d max|err|=%.3e %s\n", variant, off, maxerr, pass ? "PASS" : "FAIL");
    }
  printf("VERDICT base_offset=0:%s base_offset=off&7:%s\n", ok[0] ? "PASS" : "FAIL", ok[1] ? "PASS" : "FAIL");
  return (ok[0] || ok[1]) ? 0 : 1;
}
