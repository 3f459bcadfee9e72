// Memory-bound operators on the "planes" layouts (planes.cuh): converters at the module
// boundary (the reference is NCHW fp32), the stem input packer, 2x2 max-pool
// (dla.py:209-210), the depth-wise ConvTranspose2d up-sampler fused with the IDAUp skip add
// (dla.py:561-577) and the prior-guided token gather / deterministic write-back
// (dla.py:915-968, :1006-1018).  One thread moves one 16-byte group (8 channels) per plane.
#include "common.cuh"
#include "planes.cuh"

namespace sgta {

__device__ __forceinline__ long long frame_row(const View& v, int b, int y, int x) {   // unpadded (y,x)
  const int Wp = v.W + 2, Hp = v.H + 2;
  return ((long long)b * Hp + (y + 1)) * Wp + (x + 1);
}

template <int NS>
__device__ __forceinline__ void load8(const View& v, long long p, int c, float (&f)[8]) {
  if (v.layout == SGTA_LAYOUT_PL) pl_load8<NS>(v, c >> 6, p, (c & 63) >> 3, f);
  else sc_load8<NS>(v, p, c, f);
}
template <int NS>
__device__ __forceinline__ void store8(const View& v, long long p, int c, const float (&f)[8]) {
  if (v.layout == SGTA_LAYOUT_PL) pl_store8<NS>(v, c >> 6, p, (c & 63) >> 3, f);
  else sc_store8<NS>(v, p, c, f);
}

// ---- NCHW fp32 <-> planes -------------------------------------------------------------------
template <int NS>
__global__ void pl_from_nchw_kernel(const float* __restrict__ src, View y, int C, int c_off, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int HW = y.H * y.W;
  const int pix = (int)(e % HW);
  const long long r = e / HW;
  const int g = (int)(r % (C / 8)), b = (int)(r / (C / 8));
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = __ldg(src + ((long long)b * C + g * 8 + j) * HW + pix);
  store8<NS>(y, frame_row(y, b, pix / y.W, pix % y.W), c_off + g * 8, f);
}
template <int NS>
__global__ void pl_to_nchw_kernel(View x, float* __restrict__ dst, int C, int c_off, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int HW = x.H * x.W;
  const int pix = (int)(e % HW);
  const long long r = e / HW;
  const int g = (int)(r % (C / 8)), b = (int)(r / (C / 8));
  float f[8];
  load8<NS>(x, frame_row(x, b, pix / x.W, pix % x.W), c_off + g * 8, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) dst[((long long)b * C + g * 8 + j) * HW + pix] = f[j];
}

// img [B,3,H,W] + hm [B,1,H,W] fp32 -> SC view with C = 4 and a 3-pixel zero border (7x7 stems,
// dla.py:241-270), images [b_off, b_off+B) of the view
template <int NS>
__global__ void pl_pack_stem_kernel(const float* __restrict__ img, const float* __restrict__ hm, View y, int b_off,
                                 int B, int border, long long total) {
  pdl_trigger();
  pdl_wait();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int HW = y.H * y.W;
  const int pix = (int)(e % HW), b = (int)(e / HW);
  const int Wp = y.W + 2 * border, Hp = y.H + 2 * border;
  const long long row = ((long long)(b + b_off) * Hp + pix / y.W + border) * Wp + pix % y.W + border;
  const float v0 = __ldg(img + ((long long)b * 3 + 0) * HW + pix), v1 = __ldg(img + ((long long)b * 3 + 1) * HW + pix);
  const float v2 = __ldg(img + ((long long)b * 3 + 2) * HW + pix), v3 = __ldg(hm + (long long)b * HW + pix);
  if (NS == 2) {
    uint32_t h0, l0, h1, l1;
    split2(v0, v1, h0, l0);
    split2(v2, v3, h1, l1);
    *reinterpret_cast<uint2*>(y.base + sc_offset(y, 0, row, 0)) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(y.base + sc_offset(y, 1, row, 0)) = make_uint2(l0, l1);
  } else {
    *reinterpret_cast<uint2*>(y.base + sc_offset(y, 0, row, 0)) = make_uint2(pack_b2(v0, v1), pack_b2(v2, v3));
  }
}

// ---- 2x2 / stride-2 max-pool ------------------------------------------------------------------
template <int NS>
__global__ void pl_maxpool2_kernel(View x, View y, int C, int xc_off, int yc_off, long long total) {
  pdl_trigger();
  pdl_wait();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int G = C / 8;
  const int g = (int)(e % G);
  long long r = e / G;
  const int ox = (int)(r % y.W); r /= y.W;
  const int oy = (int)(r % y.H), b = (int)(r / y.H);
  float m[8], f[8];
  const long long p00 = frame_row(x, b, 2 * oy, 2 * ox);
  const int Wp = x.W + 2;
  load8<NS>(x, p00, xc_off + g * 8, m);
  const long long ps[3] = {p00 + 1, p00 + Wp, p00 + Wp + 1};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    load8<NS>(x, ps[k], xc_off + g * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
  }
  store8<NS>(y, frame_row(y, b, oy, ox), yc_off + g * 8, m);
}

// ---- y = ConvTranspose2d(C,C,2f,stride f,pad f/2,groups C)(x) + skip ---------------------------
// grid (x blocks, row bands, B); the depth-wise kernel is staged once per CTA as [ky*k+kx][C] so a tap is
// two LDS.128 for a thread's 8 channels (it was 32 scalar loads), and (b, oy) come from the grid
// instead of 64-bit divisions per thread.
constexpr int UPS_ROWS = 8;            // output rows per CTA
template <int NS>
__global__ void __launch_bounds__(256)
pl_upsample_add_kernel(View x, const float* __restrict__ w, View skip, int has_skip, View y, int C, int f) {
  extern __shared__ __align__(16) float ups_w[];
  pdl_trigger();
  const int k = 2 * f, pad = f / 2, kk = k * k;
  for (int i = threadIdx.x; i < kk * C; i += 256) {
    const int c = i % C, t = i / C;
    ups_w[i] = __ldg(w + (long long)c * kk + t);
  }
  pdl_wait();
  __syncthreads();
  const int G = C >> 3;
  const int b = blockIdx.z, oy0 = blockIdx.y * UPS_ROWS;
  const int oy1 = min(oy0 + UPS_ROWS, y.H);
  for (int t = blockIdx.x * 256 + threadIdx.x; t < y.W * G; t += gridDim.x * 256) {
    const int ox = t / G, g = t - ox * G;
    const int ix_hi = (ox + pad) / f;
    for (int oy = oy0; oy < oy1; ++oy) {
      float acc[8];
      const long long po = frame_row(y, b, oy, ox);
      if (has_skip) load8<NS>(skip, po, g * 8, acc);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      }
      const int iy_hi = (oy + pad) / f;
      for (int iy = iy_hi; iy >= 0 && iy > iy_hi - 2; --iy) {
        const int ky = oy + pad - iy * f;
        if (ky >= k || iy >= x.H) continue;
        for (int ix = ix_hi; ix >= 0 && ix > ix_hi - 2; --ix) {
          const int kx = ox + pad - ix * f;
          if (kx >= k || ix >= x.W) continue;
          float v[8];
          load8<NS>(x, frame_row(x, b, iy, ix), g * 8, v);
          const float4* wt = reinterpret_cast<const float4*>(ups_w + (ky * k + kx) * C + g * 8);
          const float4 w0 = wt[0], w1 = wt[1];
          acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]);
          acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
          acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]);
          acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
        }
      }
      store8<NS>(y, po, g * 8, acc);
    }
  }
}

// f = 2 (seven of the eight up-samplers of the network: dla.py:561-577 with f = 2): the output block
// {2i-1, 2i} x {2j-1, 2j} reads the same four input pixels (i-1..i, j-1..j), so one thread loads them once (the zero
// border of the frame stands in for the out-of-range ones) and produces the four outputs -- a quarter of the input
// loads and of the address arithmetic of the general kernel.  Measured: 0.530 -> 0.519 ms over the 8 launches of a
// step only -- the kernel is bound by memory latency (ncu: DRAM 32 %, L1 42 %, issue 55 %), not by instructions.
// Same tap order per output as the general kernel: skip, then (i, j), (i, j-1), (i-1, j), (i-1, j-1).
constexpr int UP2_ROWS = 2;            // input rows per CTA
template <int NS>
__global__ void __launch_bounds__(256, 2)
pl_upsample2_add_kernel(View x, const float* __restrict__ w, View skip, int has_skip, View y, int C, int rows) {
  extern __shared__ __align__(16) float ups_w[];
  pdl_trigger();
  for (int i = threadIdx.x; i < 16 * C; i += 256) {
    const int c = i % C, t = i / C;
    ups_w[i] = __ldg(w + (long long)c * 16 + t);
  }
  pdl_wait();
  __syncthreads();
  const int G = C >> 3;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= (x.W + 1) * G) return;
  const int j = t / G, g = t - j * G;
  const int b = blockIdx.z;
  const int i1 = min((int)(blockIdx.y + 1) * rows, x.H + 1);
  const int chunk = g >> 3, c8 = g & 7;            // PL views: 64-channel chunk and 16-byte group of this thread's 8 channels
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (int i = blockIdx.y * rows; i < i1; ++i) {
    // all 16 loads of the iteration (4 inputs + 4 skips, per plane) are issued before anything waits on them: the
    // per-output load -> wait -> FMA -> store chain made the kernel latency-bound (2 waves of ~20 us CTAs)
    uint4 xi[2][2][NS], sk[2][2][NS];
    long long po[2][2];
    bool ok[2][2];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const long long pi = frame_row(x, b, i - 1 + dy, j - 1 + dx);
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2)
          xi[dy][dx][s2] = __ldg(reinterpret_cast<const uint4*>(x.base + pl_offset(x, s2, chunk, pi, c8)));
      }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const int oy = 2 * i - 1 + a, ox = 2 * j - 1 + bb;
        ok[a][bb] = oy >= 0 && oy < y.H && ox >= 0 && ox < y.W;
        po[a][bb] = frame_row(y, b, oy, ox);
#pragma unroll
        for (int s2 = 0; s2 < NS; ++s2)
          sk[a][bb][s2] = (has_skip && ok[a][bb])
                              ? __ldg(reinterpret_cast<const uint4*>(skip.base + pl_offset(skip, s2, chunk, po[a][bb], c8)))
                              : zero4;
      }
    float v[2][2][8];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) decode8<NS>(xi[dy][dx][0], xi[dy][dx][NS - 1], v[dy][dx]);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        if (!ok[a][bb]) continue;
        float acc[8];
        if (has_skip) decode8<NS>(sk[a][bb][0], sk[a][bb][NS - 1], acc);
        else {
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        }
#pragma unroll
        for (int dy = 1; dy >= 0; --dy)
#pragma unroll
          for (int dx = 1; dx >= 0; --dx) {
            const int ky = a + 2 * (1 - dy), kx = bb + 2 * (1 - dx);
            const float4* wt = reinterpret_cast<const float4*>(ups_w + (ky * 4 + kx) * C + g * 8);
            const float4 w0 = wt[0], w1 = wt[1];
            const float (&u)[8] = v[dy][dx];
            acc[0] = fmaf(u[0], w0.x, acc[0]); acc[1] = fmaf(u[1], w0.y, acc[1]);
            acc[2] = fmaf(u[2], w0.z, acc[2]); acc[3] = fmaf(u[3], w0.w, acc[3]);
            acc[4] = fmaf(u[4], w1.x, acc[4]); acc[5] = fmaf(u[5], w1.y, acc[5]);
            acc[6] = fmaf(u[6], w1.z, acc[6]); acc[7] = fmaf(u[7], w1.w, acc[7]);
          }
        pl_store8<NS>(y, chunk, po[a][bb], c8, acc);
      }
    }
  }
}

// ---- super-pixel views ----------------------------------------------------------
// A [rows][16 ch] SC map read as [rows / 4][4 px x 16 ch]: one 128-byte PL row per SUPER-PIXEL (channel index
// j*16 + c for pixel j of the group), frame (H+2) x (W/4+2) with the usual one-(super-)pixel zero border, so
// that a small-channel 3x3 convolution becomes a plain PL convolution with Toeplitz-expanded weights
// (planes.superpixel_weight) that conv_shift_kernel runs as it stands.  Raw 16-byte moves, no arithmetic:
// group g of a super-pixel row = channels [8*(g&1), +8) of pixel g>>1.
template <int NS>
__global__ void pl_sc16_super4_kernel(View sc, View sp, int to_super, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int g = (int)(e & 7);
  long long r = e >> 3;
  const int xs = (int)(r % sp.W);
  r /= sp.W;
  const int y = (int)(r % sp.H), b = (int)(r / sp.H);
  const long long p_sc = frame_row(sc, b, y, 4 * xs + (g >> 1)), p_sp = frame_row(sp, b, y, xs);
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    unsigned char* a = sc.base + sc_offset(sc, s, p_sc, (g & 1) * 8);
    unsigned char* d = sp.base + pl_offset(sp, s, 0, p_sp, g);
    if (to_super) *reinterpret_cast<uint4*>(d) = __ldg(reinterpret_cast<const uint4*>(a));
    else *reinterpret_cast<uint4*>(a) = __ldg(reinterpret_cast<const uint4*>(d));
  }
}

// ---- tokens -----------------------------------------------------------------------------------
// rows[b,t,:] = feats[b, :, ids[b,t]]  (ids = y*W + x in the unpadded map); images
// [b_off, b_off+B) of the view
template <int NS>
__global__ void pl_gather_tokens_kernel(View x, int b_off, const long long* __restrict__ ids, float* __restrict__ rows,
                                     int C, int n, long long total) {
  pdl_trigger();
  pdl_wait();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int G = C / 8;
  const int g = (int)(e % G);
  const long long bt = e / G;
  const int b = (int)(bt / n);
  const long long id = ids[bt];
  float f[8];
  load8<NS>(x, frame_row(x, b + b_off, (int)(id / x.W), (int)(id % x.W)), g * 8, f);
  float4* dst = reinterpret_cast<float4*>(rows + bt * C + g * 8);
  dst[0] = make_float4(f[0], f[1], f[2], f[3]);
  dst[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// grid (samples, token chunks); token t writes iff no later token carries the same id (highest token
// index wins: deterministic, == sequential index_put_).
template <int NS>
__global__ void __launch_bounds__(256)
pl_scatter_tokens_kernel(View x, int b_off, const long long* __restrict__ ids, const float* __restrict__ rows, int C, int n) {
  extern __shared__ int s_ids[];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < n; t += blockDim.x) s_ids[t] = (int)ids[(long long)b * n + t];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int G = C / 8;
  // grid.y CTAs share a sample: each takes a contiguous chunk of tokens (the duplicate scan is O(n^2))
  const int per = (n + gridDim.y - 1) / gridDim.y;
  const int t_end = min(n, (int)(blockIdx.y + 1) * per);
  for (int t = blockIdx.y * per + warp; t < t_end; t += nw) {
    const int id = s_ids[t];
    bool dup = false;
    for (int u = t + 1 + lane; u < n; u += 32) dup |= (s_ids[u] == id);
    if (__any_sync(0xffffffffu, dup)) continue;
    const long long p = frame_row(x, b + b_off, id / x.W, id % x.W);
    for (int g = lane; g < G; g += 32) {
      const float4* src = reinterpret_cast<const float4*>(rows + ((long long)b * n + t) * C + g * 8);
      const float4 a = __ldg(src), c = __ldg(src + 1);
      const float f[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
      store8<NS>(x, p, g * 8, f);
    }
  }
}

static bool geom_ok(const sgta_planes* v) {
  return v && v->data && (v->nplanes == 1 || v->nplanes == 2) && v->B > 0 && v->H > 0 && v->W > 0 &&
         (v->layout == SGTA_LAYOUT_PL || v->layout == SGTA_LAYOUT_SC) &&
         v->rows >= (int64_t)v->guard + (int64_t)v->B * (v->H + 2 * v->border) * (v->W + 2 * v->border);
}
static bool chan_ok(const sgta_planes* v, int c_off, int C) {
  if (C <= 0 || C % 8 || c_off % 8 || c_off < 0) return false;
  if (v->layout == SGTA_LAYOUT_PL) return (long long)v->chunk0 * 64 + c_off + C <= (long long)v->nchunks * 64;
  return c_off + C <= v->nchunks;
}

}  // namespace sgta

using namespace sgta;

#define NS_DISPATCH(ns, ...)              \
  if ((ns) == 2) { constexpr int NS = 2; __VA_ARGS__; } \
  else { constexpr int NS = 1; __VA_ARGS__; }

extern "C" int sgta_planes_from_nchw(const void* src_f32, const sgta_planes* y, int C, int c_off, void* stream) {
  SGTA_REQUIRE(src_f32 && geom_ok(y) && y->border == 1 && chan_ok(y, c_off, C), "sgta_planes_from_nchw: bad arguments");
  const long long total = (long long)y->B * (C / 8) * y->H * y->W;
  View v = make_view(y);
  NS_DISPATCH(y->nplanes, (pl_from_nchw_kernel<NS><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const float*)src_f32, v, C, c_off, total)));
  return check_launch("pl_from_nchw_kernel");
}

extern "C" int sgta_planes_to_nchw(const sgta_planes* x, void* dst_f32, int C, int c_off, void* stream) {
  SGTA_REQUIRE(dst_f32 && geom_ok(x) && x->border == 1 && chan_ok(x, c_off, C), "sgta_planes_to_nchw: bad arguments");
  const long long total = (long long)x->B * (C / 8) * x->H * x->W;
  View v = make_view(x);
  NS_DISPATCH(x->nplanes, (pl_to_nchw_kernel<NS><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(v, (float*)dst_f32, C, c_off, total)));
  return check_launch("pl_to_nchw_kernel");
}

extern "C" int sgta_planes_pack_stem(const void* img_f32, const void* hm_f32, const sgta_planes* y, int b_off, int B,
                                     void* stream) {
  SGTA_REQUIRE(img_f32 && hm_f32 && geom_ok(y) && y->layout == SGTA_LAYOUT_SC && y->nchunks == 4 && y->border >= 1 &&
               b_off >= 0 && B > 0 && b_off + B <= y->B, "sgta_planes_pack_stem: bad arguments");
  const long long total = (long long)B * y->H * y->W;
  View v = make_view(y);
  NS_DISPATCH(y->nplanes, launch_k(pl_pack_stem_kernel<NS>, cdiv(total, 256), 256, 0, (cudaStream_t)stream,
                                   (const float*)img_f32, (const float*)hm_f32, v, b_off, B, y->border, total));
  return check_launch("pl_pack_stem_kernel");
}

extern "C" int sgta_planes_maxpool2(const sgta_planes* x, int xc_off, const sgta_planes* y, int yc_off, int C, void* stream) {
  SGTA_REQUIRE(geom_ok(x) && geom_ok(y) && x->border == 1 && y->border == 1 && x->nplanes == y->nplanes &&
               x->B == y->B && x->H == 2 * y->H && x->W == 2 * y->W && chan_ok(x, xc_off, C) && chan_ok(y, yc_off, C),
               "sgta_planes_maxpool2: bad arguments");
  const long long total = (long long)y->B * y->H * y->W * (C / 8);
  View vx = make_view(x), vy = make_view(y);
  NS_DISPATCH(x->nplanes, launch_k(pl_maxpool2_kernel<NS>, cdiv(total, 256), 256, 0, (cudaStream_t)stream, vx, vy, C, xc_off, yc_off, total));
  return check_launch("pl_maxpool2_kernel");
}

extern "C" int sgta_planes_superpixels(const sgta_planes* sc, const sgta_planes* sp, int to_super, void* stream) {
  SGTA_REQUIRE(geom_ok(sc) && geom_ok(sp) && sc->layout == SGTA_LAYOUT_SC && sc->nchunks == 16 && sc->border == 1 &&
               sp->layout == SGTA_LAYOUT_PL && sp->border == 1 && chan_ok(sp, 0, 64) && sc->nplanes == sp->nplanes &&
               sc->B == sp->B && sc->H == sp->H && sc->W == 4 * sp->W,
               "sgta_planes_superpixels: sc = SC view with 16 channels, sp = PL view with 64 channels and W / 4 columns");
  const long long total = (long long)sp->B * sp->H * sp->W * 8;
  View vsc = make_view(sc), vsp = make_view(sp);
  NS_DISPATCH(sc->nplanes, (pl_sc16_super4_kernel<NS><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(vsc, vsp, to_super, total)));
  return check_launch("pl_sc16_super4_kernel");
}

extern "C" int sgta_planes_upsample_add(const sgta_planes* x, const void* w_up, const sgta_planes* skip,
                                        const sgta_planes* y, int C, int f, void* stream) {
  SGTA_REQUIRE(geom_ok(x) && geom_ok(y) && w_up && f >= 1 && x->border == 1 && y->border == 1 &&
               x->nplanes == y->nplanes && x->B == y->B && y->H == x->H * f && y->W == x->W * f &&
               chan_ok(x, 0, C) && chan_ok(y, 0, C), "sgta_planes_upsample_add: bad arguments");
  if (skip)
    SGTA_REQUIRE(geom_ok(skip) && skip->border == 1 && skip->nplanes == y->nplanes && skip->B == y->B &&
                 skip->H == y->H && skip->W == y->W && chan_ok(skip, 0, C), "sgta_planes_upsample_add: bad skip view");
  View vx = make_view(x), vy = make_view(y), vs = make_view(skip);
  const size_t smem = sizeof(float) * 4 * f * f * C;
  SGTA_REQUIRE(smem <= 96 * 1024 && y->B <= 65535, "sgta_planes_upsample_add: kernel %dx%d x %d channels too large", 2 * f, 2 * f, C);
  if (f == 2 && !(debug_flags() & 65536) && x->layout == SGTA_LAYOUT_PL && y->layout == SGTA_LAYOUT_PL &&
      (!skip || skip->layout == SGTA_LAYOUT_PL)) {
    static const int rows = getenv("SGTA_UP2_ROWS") ? atoi(getenv("SGTA_UP2_ROWS")) : UP2_ROWS;
    dim3 grid2(cdiv((long long)(x->W + 1) * (C / 8), 256), cdiv(x->H + 1, rows), y->B);
    NS_DISPATCH(x->nplanes, {
      launch_k(pl_upsample2_add_kernel<NS>, grid2, 256, sizeof(float) * 16 * C, (cudaStream_t)stream, vx, (const float*)w_up, vs, (int)(skip != nullptr), vy, C, rows);
    });
    return check_launch("pl_upsample2_add_kernel");
  }
  int gx = cdiv((long long)y->W * (C / 8), 256);
  if (gx > 8) gx = 8;
  dim3 grid(gx, cdiv(y->H, UPS_ROWS), y->B);
  NS_DISPATCH(x->nplanes, {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(pl_upsample_add_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(pl_upsample_add_kernel<NS>, grid, 256, smem, (cudaStream_t)stream, vx, (const float*)w_up, vs, (int)(skip != nullptr), vy, C, f);
  });
  return check_launch("pl_upsample_add_kernel");
}

extern "C" int sgta_planes_gather_tokens(const sgta_planes* x, int b_off, const void* ids, void* rows, int B, int C,
                                         int n, void* stream) {
  SGTA_REQUIRE(geom_ok(x) && x->border == 1 && ids && rows && B > 0 && n > 0 && b_off >= 0 && b_off + B <= x->B &&
               chan_ok(x, 0, C), "sgta_planes_gather_tokens: bad arguments");
  const long long total = (long long)B * n * (C / 8);
  View v = make_view(x);
  NS_DISPATCH(x->nplanes, launch_k(pl_gather_tokens_kernel<NS>, cdiv(total, 256), 256, 0, (cudaStream_t)stream,
                                   v, b_off, (const long long*)ids, (float*)rows, C, n, total));
  return check_launch("pl_gather_tokens_kernel");
}

extern "C" int sgta_planes_scatter_tokens(const sgta_planes* x, int b_off, const void* ids, const void* rows, int B,
                                          int C, int n, void* stream) {
  SGTA_REQUIRE(geom_ok(x) && x->border == 1 && ids && rows && B > 0 && n > 0 && n <= 48 * 1024 && b_off >= 0 &&
               b_off + B <= x->B && chan_ok(x, 0, C), "sgta_planes_scatter_tokens: bad arguments");
  const size_t smem = sizeof(int) * n;
  View v = make_view(x);
  NS_DISPATCH(x->nplanes, {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(pl_scatter_tokens_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(pl_scatter_tokens_kernel<NS>, dim3(B, cdiv(n, 96)), 256, smem, (cudaStream_t)stream, v, b_off, (const long long*)ids, (const float*)rows, C, n);
  });
  return check_launch("pl_scatter_tokens_kernel");
}
