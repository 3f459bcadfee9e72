// Prior-guided token selection: top-K of the prior heatmaps, window index arithmetic,
// token gather and the deterministic write-back.
//
// Replaces get_topk_index (reference sgtapose/lib/model/networks/dla.py:898-913),
// get_topk_features_scale (:915-968) and substitute_topk_features_scale (:1006-1018).
// The reference builds its index tensors on the CPU (dla.py:903,909,959), forcing a
// device->host sync and a host->device copy per level per frame, and permutes the whole
// feature map to NHWC (9.4 MB at level 0) to pick 1183 rows; here everything stays on the
// device and only the selected rows move.
#include "common.cuh"

namespace sgta {

// One CTA per (sample, channel).  K rounds of block arg-max; ties -> lowest index
// (== CPU torch.topk for K=1); selected entries are excluded by index.
__global__ void __launch_bounds__(256)
topk_index_kernel(const float* __restrict__ hm, long long* __restrict__ idx, int HW, int K) {
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ int s_taken[64];
  pdl_trigger();
  pdl_wait();
  const float* src = hm + (long long)blockIdx.x * HW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < K; ++r) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int e = threadIdx.x; e < HW; e += blockDim.x) {
      bool taken = false;
      for (int t = 0; t < r; ++t) taken |= (s_taken[t] == e);
      if (taken) continue;
      float v = __ldg(src + e);
      // NaN never wins unless everything is NaN; -inf entries still selectable by index
      if (v > best || (v == best && e < bi) || bi == 0x7fffffff) { best = v; bi = e; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float bv = s_val[0]; int bb = s_idx[0];
      for (int wv = 1; wv < (int)(blockDim.x >> 5); ++wv) {
        float ov = s_val[wv]; int oi = s_idx[wv];
        if (oi != 0x7fffffff && (bb == 0x7fffffff || ov > bv || (ov == bv && oi < bb))) { bv = ov; bb = oi; }
      }
      s_taken[r] = bb;
      idx[(long long)blockIdx.x * K + r] = bb;
    }
    __syncthreads();
  }
}

// ids[b, ck*win2 + w] for w = a*win + c  (a: x offset index, c: y offset index; the
// reference's x-major meshgrid, dla.py:932-942).  fp32 arithmetic without FMA contraction.
__global__ void window_ids_kernel(const long long* __restrict__ idx, long long* __restrict__ ids,
                                  int total, int win, int Whm, float scale, int half, int H, int W) {
  pdl_trigger();
  pdl_wait();
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  int win2 = win * win;
  int w = e % win2;
  long long src = idx[e / win2];
  float ix = (float)(src % Whm), iy = (float)(src / Whm);
  float ox = (float)(w / win - half), oy = (float)(w % win - half);
  float cx = __fadd_rn(__fmul_rn(ix, scale), ox);
  float cy = __fadd_rn(__fmul_rn(iy, scale), oy);
  float hi = (float)(H - 1);
  cx = fminf(fmaxf(cx, 0.f), hi);
  cy = fminf(fmaxf(cy, 0.f), hi);
  float fid = __fadd_rn(__fmul_rn(cy, (float)W), cx);
  ids[e] = (long long)fid;   // truncation toward zero, as .type(torch.long)
}

// rows[b,t,c] = feats[b,c,id] (NCHW) or feats[b,id,c] (NHWC)
__global__ void gather_tokens_kernel(const float* __restrict__ feats, const long long* __restrict__ ids,
                                     float* __restrict__ rows, int C, int HW, int n, int nhwc,
                                     long long total) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  int c = (int)(e % C);
  long long bt = e / C;
  int b = (int)(bt / n);
  long long id = ids[bt];
  const float* fb = feats + (long long)b * C * HW;
  rows[e] = nhwc ? __ldg(fb + id * C + c) : __ldg(fb + (long long)c * HW + id);
}

// One CTA per sample: ids staged in shared memory; token t writes iff no later token
// carries the same id (highest token index wins -> deterministic, == sequential index_put_).
__global__ void __launch_bounds__(256)
scatter_tokens_kernel(float* __restrict__ feats, const long long* __restrict__ ids,
                      const float* __restrict__ rows, int C, int HW, int n, int nhwc) {
  extern __shared__ int s_ids[];
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < n; t += blockDim.x) s_ids[t] = (int)ids[(long long)b * n + t];
  __syncthreads();
  float* fb = feats + (long long)b * C * HW;
  const float* rb = rows + (long long)b * n * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int t = warp; t < n; t += nw) {
    int id = s_ids[t];
    bool dup = false;
    for (int u = t + 1 + lane; u < n; u += 32) dup |= (s_ids[u] == id);
    if (__any_sync(0xffffffffu, dup)) continue;
    for (int c = lane; c < C; c += 32) {
      float v = rb[(long long)t * C + c];
      if (nhwc) fb[(long long)id * C + c] = v; else fb[(long long)c * HW + id] = v;
    }
  }
}

}  // namespace sgta

using namespace sgta;

extern "C" int sgta_topk_index(const void* hm, void* idx, int B, int C, int HW, int K, void* stream) {
  SGTA_REQUIRE(hm && idx, "sgta_topk_index: null pointer");
  SGTA_REQUIRE(B > 0 && C > 0 && HW > 0 && K > 0 && K <= 64 && K <= HW, "sgta_topk_index: bad shape (K <= 64)");
  launch_k(topk_index_kernel, B * C, 256, 0, (cudaStream_t)stream, (const float*)hm, (long long*)idx, HW, K);
  return check_launch("topk_index_kernel");
}

extern "C" int sgta_window_ids(const void* idx, void* ids, int B, int CK, int Whm, float scale,
                               int kernel, int H, int W, void* stream) {
  SGTA_REQUIRE(idx && ids, "sgta_window_ids: null pointer");
  SGTA_REQUIRE(B > 0 && CK > 0 && Whm > 0 && kernel > 0 && H > 0 && W > 0, "sgta_window_ids: bad shape");
  int half = kernel / 2, win = 2 * half + 1;
  long long total = (long long)B * CK * win * win;
  SGTA_REQUIRE(total < (1ll << 31), "sgta_window_ids: too many tokens");
  launch_k(window_ids_kernel, cdiv(total, 256), 256, 0, (cudaStream_t)stream,
           (const long long*)idx, (long long*)ids, (int)total, win, Whm, scale, half, H, W);
  return check_launch("window_ids_kernel");
}

extern "C" int sgta_gather_tokens(const void* feats, const void* ids, void* rows, int B, int C,
                                  int HW, int n, int nhwc, void* stream) {
  SGTA_REQUIRE(feats && ids && rows, "sgta_gather_tokens: null pointer");
  SGTA_REQUIRE(B > 0 && C > 0 && HW > 0 && n > 0, "sgta_gather_tokens: bad shape");
  long long total = (long long)B * n * C;
  gather_tokens_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float*)feats, (const long long*)ids, (float*)rows, C, HW, n, nhwc, total);
  return check_launch("gather_tokens_kernel");
}

extern "C" int sgta_scatter_tokens(void* feats, const void* ids, const void* rows, int B, int C,
                                   int HW, int n, int nhwc, void* stream) {
  SGTA_REQUIRE(feats && ids && rows, "sgta_scatter_tokens: null pointer");
  SGTA_REQUIRE(B > 0 && C > 0 && HW > 0 && n > 0 && n <= 48 * 1024, "sgta_scatter_tokens: bad shape");
  size_t smem = sizeof(int) * n;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(scatter_tokens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  scatter_tokens_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(
      (float*)feats, (const long long*)ids, (const float*)rows, C, HW, n, nhwc);
  return check_launch("scatter_tokens_kernel");
}
