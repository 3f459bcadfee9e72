// Raw tcgen05.mma issue/execute rate on B200: one CTA per SM, one elected thread issues `iters`
// K blocks of 4 MMAs (M=128, K=16 each, bf16, SW128 operands resident in smem), N in {64,128,256},
// optionally rotating over `nacc` accumulators; clock64 around issue .. commit completion.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mma_rate_probe mma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "../../sgtapose_b200/csrc/umma.cuh"
using namespace sgta::umma;

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int nacc, int shifted, int commit_every, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, dummy, ready;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (160 * 128 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&dummy, 1 << 20); mbar_init(&ready, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t idesc = idesc_bf16_f32(128, N);
    const uint32_t a0 = smem_u32(smem) + (shifted ? 3 * 128 : 0), b0 = smem_u32(smem + 160 * 128);
    const uint64_t ad = smem_desc_sw128(a0), bd = smem_desc_sw128(b0);
    long long t0 = clock64();
    int r = 0;
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          mma_bf16_ss(tmem + (uint32_t)(r * N), ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, 1u);
          r = r + 1 == nacc ? 0 : r + 1;
        }
        if (commit_every == 101 || commit_every == 102) mma_commit(&dummy);
      }
      if (commit_every == 100 || commit_every == 101) { mbar_wait(&ready, 1); tc_fence_after(); }
      r = (r + 0);
      __syncwarp();
    }
    if (elect_one()) mma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int N>
void run(int nacc, int commit_every, unsigned long long* d) {
  const int shifted = 0;
  const int iters = 2000;
  size_t smem = 160 * 128 + N * 128 + 2048;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) rate_kernel<N><<<148, 128, smem>>>(iters, nacc, shifted, commit_every, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  unsigned long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  printf("N=%3d nacc=%d commit_every=%d : %.1f clk per MMA (ideal %d)\n", N, nacc, commit_every, avg / (iters * 4.0), N / 2);
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 148 * 8);
  for (int ce : {0, 100, 102, 101}) {
    run<64>(1, ce, d); run<128>(1, ce, d); run<256>(1, ce, d);
  }
  return 0;
}
