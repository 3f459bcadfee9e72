// Standalone probe of the tcgen05 building blocks used by dcn_umma.cu / conv kernels:
// thread-written SWIZZLE_128B A tiles, bulk-copied pre-swizzled B tiles, UMMA descriptors,
// commit -> mbarrier, TMEM load layout.  D[128 x N] = A[128 x K] * B[N x K]^T in bf16/fp32.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
// Run:   ./umma_probe [N] [K]      prints max |err| vs a host reference and PASS/FAIL.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../sgtapose_b200/csrc/umma.cuh"

using namespace sgta::umma;

constexpr int BM = 128, BK = 64;

template <int N>
__global__ void __launch_bounds__(192) probe_kernel(const __nv_bfloat16* __restrict__ A,      // [128][K] row-major
                                                    const unsigned char* __restrict__ Bpack,  // per K block: N x 64 SW128 image
                                                    float* __restrict__ D, int K) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                   // 16 KB
  unsigned char* sB = smem + BM * 128;        // N * 128 B
  __shared__ uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar_b, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 4) tmem_alloc<(N < 32 ? 32 : N)>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int nkb = K / BK;
  for (int kb = 0; kb < nkb; ++kb) {
    if (kb > 0) {
      // previous MMAs must be done reading smem before we overwrite it
      if (tid == 0) mbar_wait(&bar_mma, (kb - 1) & 1);
      __syncthreads();
    }
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_b, N * 128);
      bulk_g2s(sB, Bpack + (size_t)kb * N * 128, N * 128, &bar_b);
    }
    // 128 rows x 8 chunks of 16 B written by 128 threads (warps 0-3)
    if (tid < 128) {
      const int r = tid;
      for (int c = 0; c < 8; ++c) {
        uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)r * K + kb * BK + c * 8);
        *reinterpret_cast<uint4*>(sA + sw128_offset(r, c)) = v;
      }
      fence_async_smem();
    }
    __syncthreads();
    if (tid == 160) {   // one lane of warp 5 issues
      mbar_wait(&bar_b, kb & 1);
      tc_fence_after();
      const uint32_t idesc = idesc_bf16_f32(BM, N);
      const uint64_t ad = smem_desc_sw128(smem_u32(sA));
      const uint64_t bd = smem_desc_sw128(smem_u32(sB));
#pragma unroll
      for (int k = 0; k < BK / 16; ++k)
        mma_bf16_ss(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
      mma_commit(&bar_mma);
    }
  }
  if (tid < 128) {
    mbar_wait(&bar_mma, (nkb - 1) & 1);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 32 && c0 + j < N; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc<(N < 32 ? 32 : N)>(tmem);
}

template <int N>
int run(int K) {
  std::vector<float> Af(BM * K), Bf((size_t)N * K);
  std::vector<__nv_bfloat16> Ah(BM * K);
  std::vector<unsigned char> Bp((size_t)(K / BK) * N * 128);
  srand(1);
  for (auto& v : Af) v = (rand() % 2001 - 1000) / 1000.f;
  for (auto& v : Bf) v = (rand() % 2001 - 1000) / 1000.f;
  for (int i = 0; i < BM * K; ++i) { Ah[i] = __float2bfloat16(Af[i]); Af[i] = __bfloat162float(Ah[i]); }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      __nv_bfloat16 b = __float2bfloat16(Bf[(size_t)n * K + k]);
      Bf[(size_t)n * K + k] = __bfloat162float(b);
      int kb = k / BK, kk = k % BK;
      size_t off = (size_t)kb * N * 128 + sw128_offset(n, kk / 8) + (kk % 8) * 2;
      *reinterpret_cast<__nv_bfloat16*>(&Bp[off]) = b;
    }
  __nv_bfloat16* dA; unsigned char* dB; float* dD;
  cudaMalloc(&dA, Ah.size() * 2); cudaMalloc(&dB, Bp.size()); cudaMalloc(&dD, sizeof(float) * BM * N);
  cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bp.data(), Bp.size(), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, sizeof(float) * BM * N);
  size_t smem = BM * 128 + N * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<N><<<1, 192, smem>>>(dA, dB, dD, K);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d K=%d CUDA error: %s\nFAIL\n", N, K, cudaGetErrorString(e)); return 1; }
  std::vector<float> Dh(BM * N);
  cudaMemcpy(Dh.data(), dD, sizeof(float) * BM * N, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < BM; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)Af[m * K + k] * Bf[(size_t)n * K + k];
      maxerr = fmax(maxerr, fabs(ref - Dh[m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("N=%d K=%d max|err|=%.3e max|ref|=%.3e %s\n", N, K, maxerr, maxref, maxerr < 1e-3 * maxref ? "PASS" : "FAIL");
  return maxerr < 1e-3 * maxref ? 0 : 1;
}

int main(int argc, char** argv) {
  int N = argc > 1 ? atoi(argv[1]) : 64;
  int K = argc > 2 ? atoi(argv[2]) : 128;
  switch (N) {
    case 32: return run<32>(K);
    case 64: return run<64>(K);
    case 128: return run<128>(K);
    case 256: return run<256>(K);
  }
  printf("unsupported N\n");
  return 2;
}
