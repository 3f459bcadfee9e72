// Per-SM global->shared copy bandwidth of the TMA engine on B200, L2-resident source:
//   mode 0: cp.async.bulk (1-D, UBLKCP)            chunks of `bytes`
//   mode 1: cp.async.bulk.tensor.2d (UTMALDG)      boxes of 64 x rows 16-bit elements, no swizzle
// `depth` copies in flight per CTA, one CTA per SM.  Prints GB/s per SM and aggregate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_bw_probe tma_bw_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#include "../../sgtapose_b200/csrc/umma.cuh"
using namespace sgta::umma;

__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) bw_kernel(const unsigned char* __restrict__ src, size_t src_bytes,
                                                    const __grid_constant__ CUtensorMap map, int mode, int bytes,
                                                    int depth, int iters, unsigned long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (mode == 2) {
    // one issuing thread per warp, `depth` warps (<= 4), each with 2 copies in flight
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < depth) {
      const size_t nchunks = src_bytes / bytes;
      size_t c = ((size_t)blockIdx.x * 977 + w * 131) % nchunks;
      const long long t0 = clock64();
      const int per = iters / depth;
      for (int it = 0; it < per + 2; ++it) {
        const int s = w * 2 + (it & 1);
        if (it >= 2) mbar_wait(&bar[s], ((it / 2) - 1) & 1);
        if (it < per) {
          mbar_arrive_expect_tx(&bar[s], bytes);
          bulk_g2s(smem + (size_t)s * bytes, src + c * bytes, bytes, &bar[s]);
          c = (c + 148 * 4) % nchunks;
        }
      }
      if (w == 0) cycles[blockIdx.x] = clock64() - t0;
    }
    return;
  }
  if (threadIdx.x == 0) {
    const size_t nchunks = src_bytes / bytes;
    size_t c = (size_t)blockIdx.x * 977 % nchunks;
    const long long t0 = clock64();
    for (int it = 0; it < iters + depth; ++it) {
      const int s = it % depth;
      if (it >= depth) mbar_wait(&bar[s], ((it / depth) - 1) & 1);
      if (it < iters) {
        mbar_arrive_expect_tx(&bar[s], bytes);
        if (mode == 0) bulk_g2s(smem + (size_t)s * bytes, src + c * bytes, bytes, &bar[s]);
        else tma_2d(smem + (size_t)s * bytes, &map, 0, (int)(c * (bytes / 128)), &bar[s]);
        c = (c + 148) % nchunks;
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const size_t src_bytes = 48u << 20;           // L2 resident
  unsigned char* src;
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 1, src_bytes);
  unsigned long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
  int sms = 148, clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int cfgs[][2] = {{16384, 4}, {16384, 8}, {18432, 8}, {36864, 4}, {8192, 12}, {32768, 6}, {65536, 3}};
  for (int mode = 0; mode < 3; ++mode)
    for (auto& cfg : cfgs) {
      const int bytes = cfg[0], iters = 2000;
      int depth = cfg[1];
      if (mode == 2) depth = 4;
      if ((size_t)bytes * depth * (mode == 2 ? 2 : 1) > 196 * 1024) continue;
      CUtensorMap map;
      memset(&map, 0, sizeof(map));
      if (mode == 1) {
        if (!encode) { printf("no cuTensorMapEncodeTiled\n"); continue; }
        const int rows_box = bytes / 128;
        if (rows_box > 256) continue;
        cuuint64_t dims[2] = {64, src_bytes / 128};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)rows_box};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, src, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
      }
      for (int rep = 0; rep < 2; ++rep) {
        bw_kernel<<<sms, 128, 200 * 1024>>>(src, src_bytes, map, mode, bytes, depth, iters, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      }
      unsigned long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < sms; ++i) avg += (double)h[i];
      avg /= sms;
      const double bpc = (double)bytes * iters / avg;
      printf("mode %d (%s) bytes %6d depth %2d : %6.1f B/clk/SM  -> %.2f TB/s aggregate at %.2f GHz\n", mode,
             mode == 1 ? "tensor 2d" : (mode == 2 ? "bulk 1d x4 warps" : "bulk 1d"), bytes, depth, bpc, bpc * sms * clk_khz * 1e3 / 1e12, clk_khz / 1e6);
    }
  return 0;
}
