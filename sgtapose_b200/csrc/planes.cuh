// The engine's native activation layouts (DESIGN.md "Data layout in HBM").
//
// Every feature map lives in a zero-bordered frame of (H+2) x (W+2) pixels per image,
// flattened to one row index  p = (b*(H+2) + y+1)*(W+2) + x+1;  `guard` rows precede p = 0
// and a tail guard follows the last image, so halo reads never leave the allocation and a
// 3x3 / stride-1 convolution is a GEMM over nine row-SHIFTED views of the same buffer.
//
//   layout PL (C % 64 == 0):  [plane][chunk = c/64][row][64 channels], 128 bytes per row, with
//       the eight 16-byte groups of a row XOR-swizzled by (row & 7) -- exactly the
//       SWIZZLE_128B K-major image tcgen05.mma reads, so one linear bulk copy
//       (cp.async.bulk, TMA engine) of consecutive rows IS a valid UMMA operand tile;
//   layout SC (C in {4,16,32}): [plane][row][C], plain row-major (stem / level0 / level1).
//
// planes: 1 = bf16;  2 = fp16 hi + fp16 (residual * 2^11): fp32 values to ~2^-22 relative,
// multiplied on the tensor cores as hi*hi + (hi*lo + lo*hi) * 2^-11 (two TMEM accumulators).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/sgta_b200.h"

namespace sgta {

constexpr float LO_SCALE = 2048.f;
constexpr float LO_INV = 1.f / 2048.f;

struct View {                // device-side copy of sgta_planes
  unsigned char* base;
  long long rows;
  int guard, nchunks, chunk0, nplanes, layout, border;
  int B, H, W;
};

inline View make_view(const sgta_planes* p) {
  View v{};
  if (p) {
    v.base = (unsigned char*)p->data; v.rows = p->rows; v.guard = p->guard; v.nchunks = p->nchunks;
    v.chunk0 = p->chunk0; v.nplanes = p->nplanes; v.layout = p->layout; v.border = p->border; v.B = p->B; v.H = p->H; v.W = p->W;
  }
  return v;
}

// byte offset of the 16-byte group holding channels [c8*8, c8*8+8) of chunk `chunk` (relative
// to the view) at padded row p, plane s
__device__ __forceinline__ size_t pl_offset(const View& v, int s, int chunk, long long p, int c8) {
  const long long r = v.guard + p;
  return ((((size_t)s * v.nchunks + v.chunk0 + chunk) * v.rows + r) << 7) + (size_t)(((c8 ^ (int)(r & 7)) & 7) << 4);
}
__device__ __forceinline__ size_t sc_offset(const View& v, int s, long long p, int c) {
  return (((size_t)s * v.rows + v.guard + p) * v.nchunks + c) * 2;
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_b2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
__device__ __forceinline__ float2 unpack_b2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ float clamp_h(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }

// fp32 pair -> (hi, lo) fp16 pairs
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  a = clamp_h(a); b = clamp_h(b);
  __half2 h = __floats2half2_rn(a, b);
  float2 hf = __half22float2(h);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = pack_h2((a - hf.x) * LO_SCALE, (b - hf.y) * LO_SCALE);
}

// 8 fp32 values -> one 16-byte group per plane
template <int NS>
__device__ __forceinline__ void encode8(const float (&f)[8], uint4& p0, uint4& p1) {
  if (NS == 2) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(f[2 * i], f[2 * i + 1], h[i], l[i]);
    p0 = make_uint4(h[0], h[1], h[2], h[3]);
    p1 = make_uint4(l[0], l[1], l[2], l[3]);
  } else {
    p0 = make_uint4(pack_b2(f[0], f[1]), pack_b2(f[2], f[3]), pack_b2(f[4], f[5]), pack_b2(f[6], f[7]));
  }
}
template <int NS>
__device__ __forceinline__ void decode8(const uint4& p0, const uint4& p1, float (&f)[8]) {
  const uint32_t a[4] = {p0.x, p0.y, p0.z, p0.w};
  if (NS == 2) {
    const uint32_t b[4] = {p1.x, p1.y, p1.z, p1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 h = unpack_h2(a[i]), l = unpack_h2(b[i]);
      f[2 * i] = fmaf(l.x, LO_INV, h.x);
      f[2 * i + 1] = fmaf(l.y, LO_INV, h.y);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 h = unpack_b2(a[i]);
      f[2 * i] = h.x; f[2 * i + 1] = h.y;
    }
  }
}

// load / store 8 consecutive channels (c8 = group index inside the chunk) of a PL view
template <int NS>
__device__ __forceinline__ void pl_load8(const View& v, int chunk, long long p, int c8, float (&f)[8]) {
  uint4 a = __ldg(reinterpret_cast<const uint4*>(v.base + pl_offset(v, 0, chunk, p, c8)));
  uint4 b = make_uint4(0, 0, 0, 0);
  if (NS == 2) b = __ldg(reinterpret_cast<const uint4*>(v.base + pl_offset(v, 1, chunk, p, c8)));
  decode8<NS>(a, b, f);
}
template <int NS>
__device__ __forceinline__ void pl_store8(const View& v, int chunk, long long p, int c8, const float (&f)[8]) {
  uint4 a, b;
  encode8<NS>(f, a, b);
  *reinterpret_cast<uint4*>(v.base + pl_offset(v, 0, chunk, p, c8)) = a;
  if (NS == 2) *reinterpret_cast<uint4*>(v.base + pl_offset(v, 1, chunk, p, c8)) = b;
}
// 32-byte store (STG.256, sm_100): lo = the 16 bytes at the lower address
__device__ __forceinline__ void stg256(void* ptr, const uint4& lo, const uint4& hi) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(lo.x), "r"(lo.y), "r"(lo.z),
               "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}
// 16 consecutive channels (groups c8, c8 + 1, c8 even) of a PL view: the swizzle XORs the group index with the row's low
// bits, so the pair stays inside one aligned 32-byte sector (swapped when the row is odd) -- one STG.256 per plane
// instead of two STG.128: half the store instructions and half the L1 wavefronts of the epilogues
template <int NS>
__device__ __forceinline__ void pl_store16(const View& v, int chunk, long long p, int c8, const float (&f)[16]) {
  float f0[8], f1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { f0[j] = f[j]; f1[j] = f[8 + j]; }
  uint4 a0, b0, a1, b1;
  encode8<NS>(f0, a0, b0);
  encode8<NS>(f1, a1, b1);
  const long long r = v.guard + p;
  const bool odd = r & 1;
  unsigned char* dst = v.base + ((((size_t)v.chunk0 + chunk) * v.rows + r) << 7) + (size_t)((((c8 ^ (int)(r & 7)) & 6)) << 4);
  stg256(dst, odd ? a1 : a0, odd ? a0 : a1);
  if (NS == 2) stg256(dst + (((size_t)v.nchunks * v.rows) << 7), odd ? b1 : b0, odd ? b0 : b1);
}
template <int NS>
__device__ __forceinline__ void sc_load8(const View& v, long long p, int c, float (&f)[8]) {
  uint4 a = __ldg(reinterpret_cast<const uint4*>(v.base + sc_offset(v, 0, p, c)));
  uint4 b = make_uint4(0, 0, 0, 0);
  if (NS == 2) b = __ldg(reinterpret_cast<const uint4*>(v.base + sc_offset(v, 1, p, c)));
  decode8<NS>(a, b, f);
}
template <int NS>
__device__ __forceinline__ void sc_store8(const View& v, long long p, int c, const float (&f)[8]) {
  uint4 a, b;
  encode8<NS>(f, a, b);
  *reinterpret_cast<uint4*>(v.base + sc_offset(v, 0, p, c)) = a;
  if (NS == 2) *reinterpret_cast<uint4*>(v.base + sc_offset(v, 1, p, c)) = b;
}

// 16 consecutive channels of an SC view (c % 16 == 0: 32-byte aligned in every SC layout used, C in {16, 32})
template <int NS>
__device__ __forceinline__ void sc_store16(const View& v, long long p, int c, const float (&f)[16]) {
  float f0[8], f1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { f0[j] = f[j]; f1[j] = f[8 + j]; }
  uint4 a0, b0, a1, b1;
  encode8<NS>(f0, a0, b0);
  encode8<NS>(f1, a1, b1);
  stg256(v.base + sc_offset(v, 0, p, c), a0, a1);
  if (NS == 2) stg256(v.base + sc_offset(v, 1, p, c), b0, b1);
}

}  // namespace sgta
