// Memory-bound NHWC helpers of the inference engine: layout converters at the module
// boundary (the reference is NCHW fp32), 2x2 max-pool (dla.py:209-210), the
// depth-wise bilinear ConvTranspose2d up-sampler fused with the IDAUp skip add
// (dla.py:561-577), and dtype-aware token gather / scatter on NHWC maps.
#include "common.cuh"

namespace sgta {

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_from_float(T* p, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// src NCHW fp32 [B,C,HW] -> dst[(b*HW + p)*ld + coff + c]; 32x32 shared-memory transpose
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int HW,
                                    long long ld, int coff) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    int c = c0 + i, p = p0 + tx;
    tile[i][tx] = (c < C && p < HW) ? __ldg(src + ((long long)b * C + c) * HW + p) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int p = p0 + i, c = c0 + tx;
    if (p < HW && c < C) st_from_float(dst + ((long long)b * HW + p) * ld + coff + c, tile[tx][i]);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int HW,
                                    long long ld, int coff) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    int p = p0 + i, c = c0 + tx;
    tile[i][tx] = (p < HW && c < C) ? ld_as_float(src + ((long long)b * HW + p) * ld + coff + c) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int c = c0 + i, p = p0 + tx;
    if (c < C && p < HW) dst[((long long)b * C + c) * HW + p] = tile[tx][i];
  }
}

// one thread per (output pixel, 4 channels)
template <typename T>
__global__ void maxpool2_kernel(const T* __restrict__ x, long long ldx, T* __restrict__ y, long long ldy,
                                int H, int W, int C, long long total) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int C4 = C / 4;
  int c = (int)(e % C4) * 4;
  long long po = e / C4;
  const int Wo = W / 2, Ho = H / 2;
  int ox = (int)(po % Wo), oy = (int)((po / Wo) % Ho), b = (int)(po / ((long long)Wo * Ho));
  const T* s = x + (((long long)b * H + 2 * oy) * W + 2 * ox) * ldx + c;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float m = fmaxf(fmaxf(ld_as_float(s + j), ld_as_float(s + ldx + j)),
                    fmaxf(ld_as_float(s + (long long)W * ldx + j), ld_as_float(s + (long long)W * ldx + ldx + j)));
    st_from_float(y + po * ldy + c + j, m);
  }
}

// y[b,oy,ox,c] = skip[b,oy,ox,c] + sum_{ky,kx} x[b,iy,ix,c] * w[c,ky,kx],  oy = iy*f - f/2 + ky,
// k = 2f (ConvTranspose2d(o,o,2f,stride=f,padding=f//2,groups=o), dla.py:561-563).
template <typename T>
__global__ void upsample_add_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                    const T* __restrict__ skip, long long ldskip, T* __restrict__ y,
                                    long long ldy, int h, int wd, int C, int f, long long total) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int C4 = C / 4;
  int c = (int)(e % C4) * 4;
  long long po = e / C4;
  const int Wo = wd * f, Ho = h * f, k = 2 * f, pad = f / 2;
  int ox = (int)(po % Wo), oy = (int)((po / Wo) % Ho), b = (int)(po / ((long long)Wo * Ho));
  float acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = skip ? ld_as_float(skip + po * ldskip + c + j) : 0.f;
  // ky = oy + pad - iy*f in [0, k)  ->  iy in ((oy+pad-k)/f, (oy+pad)/f]
  int iy_hi = (oy + pad) / f, ix_hi = (ox + pad) / f;
  for (int iy = iy_hi; iy >= 0 && iy > iy_hi - 2; --iy) {
    int ky = oy + pad - iy * f;
    if (ky >= k || iy >= h) continue;
    for (int ix = ix_hi; ix >= 0 && ix > ix_hi - 2; --ix) {
      int kx = ox + pad - ix * f;
      if (kx >= k || ix >= wd) continue;
      const T* s = x + (((long long)b * h + iy) * wd + ix) * C + c;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        acc[j] = fmaf(ld_as_float(s + j), __ldg(w + ((long long)(c + j) * k + ky) * k + kx), acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) st_from_float(y + po * ldy + c + j, acc[j]);
}

template <typename T>
__global__ void gather_tokens_nhwc_kernel(const T* __restrict__ feats, long long ld,
                                          const long long* __restrict__ ids, float* __restrict__ rows,
                                          int C, int HW, int n, long long total) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  int c = (int)(e % C);
  long long bt = e / C;
  int b = (int)(bt / n);
  rows[e] = ld_as_float(feats + ((long long)b * HW + ids[bt]) * ld + c);
}

template <typename T>
__global__ void __launch_bounds__(256)
scatter_tokens_nhwc_kernel(T* __restrict__ feats, long long ld, const long long* __restrict__ ids,
                           const float* __restrict__ rows, int C, int HW, int n) {
  extern __shared__ int s_ids[];
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < n; t += blockDim.x) s_ids[t] = (int)ids[(long long)b * n + t];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int t = warp; t < n; t += nw) {
    int id = s_ids[t];
    bool dup = false;
    for (int u = t + 1 + lane; u < n; u += 32) dup |= (s_ids[u] == id);
    if (__any_sync(0xffffffffu, dup)) continue;      // highest token index wins
    for (int c = lane; c < C; c += 32)
      st_from_float(feats + ((long long)b * HW + id) * ld + c, rows[((long long)b * n + t) * C + c]);
  }
}

}  // namespace sgta

using namespace sgta;

#define DISPATCH_DTYPE(dt, ...)                                   \
  if ((dt) == SGTA_DTYPE_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
  else { using T = float; __VA_ARGS__; }

extern "C" int sgta_nchw_to_nhwc(const void* src, void* dst, int B, int C, int HW, int64_t ld, int coff,
                                 int dst_dtype, void* stream) {
  SGTA_REQUIRE(src && dst && B > 0 && C > 0 && HW > 0 && ld >= coff + C && B <= 65535, "sgta_nchw_to_nhwc: bad arguments");
  dim3 grid(cdiv(HW, 32), cdiv(C, 32), B), block(32, 8);
  DISPATCH_DTYPE(dst_dtype, (nchw_to_nhwc_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>(
                                (const float*)src, (T*)dst, C, HW, ld, coff)));
  return check_launch("nchw_to_nhwc_kernel");
}

extern "C" int sgta_nhwc_to_nchw(const void* src, void* dst, int B, int C, int HW, int64_t ld, int coff,
                                 int src_dtype, void* stream) {
  SGTA_REQUIRE(src && dst && B > 0 && C > 0 && HW > 0 && ld >= coff + C && B <= 65535, "sgta_nhwc_to_nchw: bad arguments");
  dim3 grid(cdiv(HW, 32), cdiv(C, 32), B), block(32, 8);
  DISPATCH_DTYPE(src_dtype, (nhwc_to_nchw_kernel<T><<<grid, block, 0, (cudaStream_t)stream>>>(
                                (const T*)src, (float*)dst, C, HW, ld, coff)));
  return check_launch("nhwc_to_nchw_kernel");
}

extern "C" int sgta_maxpool2x2_nhwc(const void* x, int64_t ldx, void* y, int64_t ldy, int B, int H, int W,
                                    int C, int dtype, void* stream) {
  SGTA_REQUIRE(x && y && B > 0 && H > 1 && W > 1 && C > 0 && C % 4 == 0 && H % 2 == 0 && W % 2 == 0,
               "sgta_maxpool2x2_nhwc: bad arguments (even H, W; C %% 4 == 0)");
  long long total = (long long)B * (H / 2) * (W / 2) * (C / 4);
  DISPATCH_DTYPE(dtype, (maxpool2_kernel<T><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)x, ldx, (T*)y, ldy, H, W, C, total)));
  return check_launch("maxpool2_kernel");
}

extern "C" int sgta_upsample_add_nhwc(const void* x, const void* w_up, const void* skip, int64_t ldskip,
                                      void* y, int64_t ldy, int B, int h, int w, int C, int f, int dtype,
                                      void* stream) {
  SGTA_REQUIRE(x && w_up && y && B > 0 && h > 0 && w > 0 && C > 0 && C % 4 == 0 && f >= 1,
               "sgta_upsample_add_nhwc: bad arguments");
  long long total = (long long)B * h * f * w * f * (C / 4);
  DISPATCH_DTYPE(dtype, (upsample_add_kernel<T><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)x, (const float*)w_up, (const T*)skip, ldskip, (T*)y, ldy, h, w, C, f, total)));
  return check_launch("upsample_add_kernel");
}

extern "C" int sgta_gather_tokens_nhwc(const void* feats, int64_t ld, const void* ids, void* rows, int B,
                                       int C, int HW, int n, int dtype, void* stream) {
  SGTA_REQUIRE(feats && ids && rows && B > 0 && C > 0 && HW > 0 && n > 0, "sgta_gather_tokens_nhwc: bad arguments");
  long long total = (long long)B * n * C;
  DISPATCH_DTYPE(dtype, (gather_tokens_nhwc_kernel<T><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)feats, ld, (const long long*)ids, (float*)rows, C, HW, n, total)));
  return check_launch("gather_tokens_nhwc_kernel");
}

extern "C" int sgta_scatter_tokens_nhwc(void* feats, int64_t ld, const void* ids, const void* rows, int B,
                                        int C, int HW, int n, int dtype, void* stream) {
  SGTA_REQUIRE(feats && ids && rows && B > 0 && C > 0 && HW > 0 && n > 0 && n <= 48 * 1024,
               "sgta_scatter_tokens_nhwc: bad arguments");
  size_t smem = sizeof(int) * n;
  DISPATCH_DTYPE(dtype, {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(scatter_tokens_nhwc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    scatter_tokens_nhwc_kernel<T><<<B, 256, smem, (cudaStream_t)stream>>>((T*)feats, ld, (const long long*)ids,
                                                                        (const float*)rows, C, HW, n);
  });
  return check_launch("scatter_tokens_nhwc_kernel");
}
