// Implicit-GEMM convolutions on tcgen05 / TMEM (sm_100a): the modulated deformable conv of
// DeformConv (reference sgtapose/lib/model/networks/dla.py:538-550) and every plain
// convolution of the DLA-34 base, roots, projections and heads (dla.py:41-69, :157-175,
// :234-337; base_model.py:121-135), all NHWC.
//
//   out[p, o] = act( scale[o] * sum_k A[p, k] * Wt[k, o] + shift[o] (+ residual[p, o]) )
//
// K = (tap, channel) with the channel fastest, cut into blocks of 64 (one 128-byte
// SWIZZLE_128B row of bf16).  A never exists in global memory:
//   * DCN producer:  A[p,(tap,c)] = sigmoid(mask[p,tap]) * bilinear(x[:,c], p + tap + offset[p,tap])
//     -- four contiguous NHWC channel vectors per sample, blended in fp32 (upstream DCNv2
//     writes the [9*Cin, H*W] column matrix to HBM and reads it back for cuBLAS);
//   * conv producers: A[p,(tap,c)] = x[pixel(p)+tap, c] (zero outside the image).
// A CTA owns 128 output pixels x NT output channels.  8 producer warps write the A tile into
// shared memory directly in the UMMA canonical layout; one thread streams the pre-swizzled
// weight tile of each K block with a single bulk async copy (TMA engine) onto the same
// mbarrier; one thread issues tcgen05.mma (M=128, N=NT, K=16) into a TMEM accumulator and
// releases stages with tcgen05.commit; the producer warps then become the epilogue: tcgen05.ld,
// folded bias / eval-BatchNorm scale+shift, residual, ReLU / sigmoid, NHWC (fp32 or bf16) or
// NCHW fp32 store.
//
// MODE BF16 : bf16 activations and weights, fp32 accumulate, 1 MMA per K step.
// MODE F32X3: fp32 activations; A and W split into bf16 hi+lo, 3 MMAs per K step
//             (hi*hi + lo*hi + hi*lo) -> ~2^-16 relative: the fp32 parity mode.
#include "common.cuh"
#include "umma.cuh"

namespace sgta {
using namespace umma;

constexpr int CM = 128;           // pixels per CTA (UMMA M)
constexpr int CK = 64;            // K per block
constexpr int PROD_WARPS = 8;
constexpr int CONV_THREADS = (PROD_WARPS + 2) * 32;   // + MMA warp + weight-loader warp
constexpr int OMS = 33;           // padded row stride of the offset/mask tile in smem
constexpr int A_TILE = CM * 128;  // bytes of one bf16 A tile
constexpr int SMEM_MAX = 227 * 1024;

enum { PROD_GEN = 0, PROD_C64 = 1, PROD_DCN = 2 };
enum { EPI_NHWC = 0, EPI_STEM = 1, EPI_NCHW = 2 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

struct ConvArgs {
  const void* x; long long ldx;
  const float* om;
  const unsigned char* wpack;
  const float* scale; const float* shift;
  const void* res; long long ldres;
  void* y; long long ldy;
  int Mtot, H, W, Ho, Wo, Cin, Cout, kh, kw, stride, pad;
  int nkb, ntile, stages, act, y_bf16, res_bf16, tmem_cols, epi, n_valid;
};

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(u[i] << 16);
    f[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void split_hi_lo(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat16 h0 = __float2bfloat16_rn(a), h1 = __float2bfloat16_rn(b);
  hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  lo = pack_bf16(a - __bfloat162float(h0), b - __bfloat162float(h1));
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

template <int PROD, int MODE>
__global__ void __launch_bounds__(CONV_THREADS, 1) conv_umma_kernel(const ConvArgs a) {
  constexpr int NT2 = MODE == SGTA_MMA_F32X3 ? 2 : 1;
  extern __shared__ unsigned char conv_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(conv_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int NT = a.ntile;
  const int a_bytes = A_TILE * NT2, b_bytes = NT * 128 * NT2;
  const int stage_bytes = a_bytes + b_bytes;
  const int stages = a.stages;
  float* om_s = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(om_s + (PROD == PROD_DCN ? CM * OMS : 64));
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* accum_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile0 = blockIdx.x * CM;
  const int n0 = blockIdx.y * NT;
  const int nkb = a.nkb;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], PROD_WARPS + 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == PROD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (PROD == PROD_DCN) {
    if (warp < PROD_WARPS) {          // raw offset/mask tile: [128 pixels][32 floats], contiguous in NHWC
      for (int e = tid; e < CM * 32; e += PROD_WARPS * 32) {
        int r = e >> 5, c = e & 31;
        int p = tile0 + r;
        om_s[r * OMS + c] = (p < a.Mtot) ? __ldg(a.om + (size_t)p * 32 + c) : 0.f;
      }
    }
  } else if (PROD == PROD_GEN) {
    if (tid < 64) {                   // tap -> (ky, kx) table; 0xffff = beyond the filter
      int taps = a.kh * a.kw;
      reinterpret_cast<uint32_t*>(om_s)[tid] = tid < taps ? (uint32_t)((tid / a.kw) << 8 | (tid % a.kw)) : 0xffffu;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < PROD_WARPS) {
    // ================================================================== A producers
    const int r = tid >> 1, half = tid & 1;          // tile row, which 32-wide half of the K block
    const int p = tile0 + r;
    const bool pvalid = p < a.Mtot;
    const int pp = pvalid ? p : 0;
    const int ox = pp % a.Wo, oy = (pp / a.Wo) % a.Ho, pb = pp / (a.Wo * a.Ho);
    const size_t img_base = (size_t)pb * a.H * a.W;
    const int H = a.H, W = a.W;
    const long long ldx = a.ldx;

    auto wait_stage = [&](int kb) -> unsigned char* {
      const int s = kb % stages;
      mbar_wait(&empty_bar[s], (((uint32_t)(kb / stages)) & 1u) ^ 1u);
      return smem + (size_t)s * stage_bytes;
    };
    auto publish = [&](int kb) {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[kb % stages]);
    };

    if (PROD == PROD_DCN) {
      const int cchunks = a.Cin / CK;
      int kb = 0;
      for (int tap = 0; tap < 9; ++tap) {
        float sy = (float)(oy - 1 + tap / 3) + om_s[r * OMS + 2 * tap];
        float sx = (float)(ox - 1 + tap % 3) + om_s[r * OMS + 2 * tap + 1];
        float m = 1.f / (1.f + __expf(-om_s[r * OMS + 18 + tap]));
        if (!pvalid) m = 0.f;
        sy = fminf(fmaxf(sy, -2.f), (float)H + 1.f);
        sx = fminf(fmaxf(sx, -2.f), (float)W + 1.f);
        float yf = floorf(sy), xf = floorf(sx);
        int y0 = (int)yf, x0 = (int)xf;
        float ly = sy - yf, lx = sx - xf, hy = 1.f - ly, hx = 1.f - lx;
        bool y0ok = y0 >= 0 && y0 < H, y1ok = y0 + 1 >= 0 && y0 + 1 < H;
        bool x0ok = x0 >= 0 && x0 < W, x1ok = x0 + 1 >= 0 && x0 + 1 < W;
        float w00 = (y0ok && x0ok) ? m * hy * hx : 0.f;
        float w01 = (y0ok && x1ok) ? m * hy * lx : 0.f;
        float w10 = (y1ok && x0ok) ? m * ly * hx : 0.f;
        float w11 = (y1ok && x1ok) ? m * ly * lx : 0.f;
        int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
        int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
        const size_t o00 = (img_base + (size_t)yc0 * W + xc0) * ldx + half * 32;
        const size_t o01 = (img_base + (size_t)yc0 * W + xc1) * ldx + half * 32;
        const size_t o10 = (img_base + (size_t)yc1 * W + xc0) * ldx + half * 32;
        const size_t o11 = (img_base + (size_t)yc1 * W + xc1) * ldx + half * 32;
        for (int cc = 0; cc < cchunks; ++cc, ++kb) {
          unsigned char* sA = wait_stage(kb);
          if (MODE == SGTA_MMA_BF16) {
            const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(a.x) + cc * CK;
            uint4 c00[4], c01[4], c10[4], c11[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              c00[j] = __ldg(reinterpret_cast<const uint4*>(xb + o00) + j);
              c01[j] = __ldg(reinterpret_cast<const uint4*>(xb + o01) + j);
              c10[j] = __ldg(reinterpret_cast<const uint4*>(xb + o10) + j);
              c11[j] = __ldg(reinterpret_cast<const uint4*>(xb + o11) + j);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float fa[8], fb[8], fc[8], fd[8];
              unpack8(c00[j], fa); unpack8(c01[j], fb); unpack8(c10[j], fc); unpack8(c11[j], fd);
              uint32_t o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float v0 = w00 * fa[2 * i] + w01 * fb[2 * i] + w10 * fc[2 * i] + w11 * fd[2 * i];
                float v1 = w00 * fa[2 * i + 1] + w01 * fb[2 * i + 1] + w10 * fc[2 * i + 1] + w11 * fd[2 * i + 1];
                o[i] = pack_bf16(v0, v1);
              }
              *reinterpret_cast<uint4*>(sA + sw128_offset(r, half * 4 + j)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          } else {
            const float* xb = reinterpret_cast<const float*>(a.x) + cc * CK;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float v[8];
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                float4 fa = __ldg(reinterpret_cast<const float4*>(xb + o00) + 2 * j + q);
                float4 fb = __ldg(reinterpret_cast<const float4*>(xb + o01) + 2 * j + q);
                float4 fc = __ldg(reinterpret_cast<const float4*>(xb + o10) + 2 * j + q);
                float4 fd = __ldg(reinterpret_cast<const float4*>(xb + o11) + 2 * j + q);
                v[4 * q + 0] = w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x;
                v[4 * q + 1] = w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y;
                v[4 * q + 2] = w00 * fa.z + w01 * fb.z + w10 * fc.z + w11 * fd.z;
                v[4 * q + 3] = w00 * fa.w + w01 * fb.w + w10 * fc.w + w11 * fd.w;
              }
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) split_hi_lo(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
              const uint32_t off = sw128_offset(r, half * 4 + j);
              *reinterpret_cast<uint4*>(sA + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(sA + A_TILE + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
          publish(kb);
        }
      }
    } else if (PROD == PROD_C64) {
      // plain conv, Cin % 64 == 0: one tap per K block, 32 contiguous channels per thread
      const int cchunks = a.Cin / CK;
      const int iy0 = oy * a.stride - a.pad, ix0 = ox * a.stride - a.pad;
      int kb = 0;
      for (int tap = 0; tap < a.kh * a.kw; ++tap) {
        const int iy = iy0 + tap / a.kw, ix = ix0 + tap % a.kw;
        const bool ok = pvalid && iy >= 0 && iy < H && ix >= 0 && ix < W;
        const size_t off0 = ok ? (img_base + (size_t)iy * W + ix) * ldx + half * 32 : 0;
        for (int cc = 0; cc < cchunks; ++cc, ++kb) {
          unsigned char* sA = wait_stage(kb);
          if (MODE == SGTA_MMA_BF16) {
            const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.x) + off0 + cc * CK);
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = ok ? __ldg(src + j) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(sA + sw128_offset(r, half * 4 + j)) = v[j];
          } else {
            const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.x) + off0 + cc * CK);
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = ok ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t hi[4], lo[4];
              split_hi_lo(v[2 * j].x, v[2 * j].y, hi[0], lo[0]);
              split_hi_lo(v[2 * j].z, v[2 * j].w, hi[1], lo[1]);
              split_hi_lo(v[2 * j + 1].x, v[2 * j + 1].y, hi[2], lo[2]);
              split_hi_lo(v[2 * j + 1].z, v[2 * j + 1].w, hi[3], lo[3]);
              const uint32_t off = sw128_offset(r, half * 4 + j);
              *reinterpret_cast<uint4*>(sA + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(sA + A_TILE + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
          publish(kb);
        }
      }
    } else {
      // plain conv, small power-of-two Cin (4..32): several taps per K block, quads of 4 channels
      const uint32_t* taptab = reinterpret_cast<const uint32_t*>(om_s);
      const int cshift = 31 - __clz(a.Cin);
      const int cmask = a.Cin - 1;
      const int iy0 = oy * a.stride - a.pad, ix0 = ox * a.stride - a.pad;
      for (int kb = 0; kb < nkb; ++kb) {
        unsigned char* sA = wait_stage(kb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int k = kb * CK + (half * 4 + j) * 8 + q * 4;
            const int tap = k >> cshift, c = k & cmask;
            const uint32_t t = tap < 64 ? taptab[tap] : 0xffffu;
            const int iy = iy0 + (int)(t >> 8), ix = ix0 + (int)(t & 0xff);
            const bool ok = pvalid && t != 0xffffu && iy >= 0 && iy < H && ix >= 0 && ix < W;
            const size_t off = ok ? (img_base + (size_t)iy * W + ix) * ldx + c : 0;
            if (MODE == SGTA_MMA_BF16) {
              uint2 v = ok ? __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(a.x) + off))
                           : make_uint2(0, 0);
              hi[2 * q] = v.x; hi[2 * q + 1] = v.y;
            } else {
              float4 v = ok ? __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.x) + off))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
              split_hi_lo(v.x, v.y, hi[2 * q], lo[2 * q]);
              split_hi_lo(v.z, v.w, hi[2 * q + 1], lo[2 * q + 1]);
            }
          }
          const uint32_t off = sw128_offset(r, half * 4 + j);
          *reinterpret_cast<uint4*>(sA + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (MODE == SGTA_MMA_F32X3)
            *reinterpret_cast<uint4*>(sA + A_TILE + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        publish(kb);
      }
    }

    // ================================================================== epilogue
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3, hsel = warp >> 2;
    const int row = q * 32 + lane;
    const int po = tile0 + row;
    const bool rvalid = po < a.Mtot;
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16);
    if (a.epi == EPI_STEM) {
      // N = 32: out[c] = relu(bn_a(acc[c])) + relu(bn_b(acc[16+c])), c < 16   (dla.py:325-331)
      if (hsel == 0) {
        uint32_t va[16], vb[16];
        tmem_ld16(tbase, va);
        tmem_ld16(tbase + 16, vb);
        tmem_ld_wait();
        if (rvalid) {
          float o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float ta = fmaf(__uint_as_float(va[j]), __ldg(a.scale + j), __ldg(a.shift + j));
            float tb = fmaf(__uint_as_float(vb[j]), __ldg(a.scale + 16 + j), __ldg(a.shift + 16 + j));
            o[j] = fmaxf(ta, 0.f) + fmaxf(tb, 0.f);
          }
          if (a.y_bf16) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + (size_t)po * a.ldy);
            dst[0] = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
            dst[1] = make_uint4(pack_bf16(o[8], o[9]), pack_bf16(o[10], o[11]), pack_bf16(o[12], o[13]), pack_bf16(o[14], o[15]));
          } else {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.y) + (size_t)po * a.ldy);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          }
        }
      }
    } else {
      for (int c0 = hsel * 16; c0 < NT; c0 += 32) {
        uint32_t v[16];
        tmem_ld16(tbase + (uint32_t)c0, v);
        tmem_ld_wait();
        const int n = n0 + c0;
        if (!rvalid || n >= a.n_valid) continue;
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = fmaf(__uint_as_float(v[j]), __ldg(a.scale + n + j), __ldg(a.shift + n + j));
        if (a.res) {
          if (a.res_bf16) {
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.res) + (size_t)po * a.ldres + n);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              float f[8];
              unpack8(__ldg(rp + j), f);
#pragma unroll
              for (int i = 0; i < 8; ++i) o[8 * j + i] += f[i];
            }
          } else {
            const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.res) + (size_t)po * a.ldres + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 f = __ldg(rp + j);
              o[4 * j] += f.x; o[4 * j + 1] += f.y; o[4 * j + 2] += f.z; o[4 * j + 3] += f.w;
            }
          }
        }
        if (a.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
        } else if (a.act == ACT_SIGMOID) {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = 1.f / (1.f + expf(-o[j]));
        }
        if (a.epi == EPI_NCHW) {
          const int hw = a.Ho * a.Wo;
          const int bb = po / hw, pix = po % hw;
          float* dst = reinterpret_cast<float*>(a.y) + ((size_t)bb * a.n_valid + n) * hw + pix;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n + j < a.n_valid) dst[(size_t)j * hw] = o[j];
        } else if (a.y_bf16) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + (size_t)po * a.ldy + n);
          dst[0] = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
          dst[1] = make_uint4(pack_bf16(o[8], o[9]), pack_bf16(o[10], o[11]), pack_bf16(o[12], o[13]), pack_bf16(o[14], o[15]));
        } else {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.y) + (size_t)po * a.ldy + n);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        }
      }
    }
    tc_fence_before();
  } else if (warp == PROD_WARPS) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_f32(CM, NT);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % stages;
        mbar_wait(&full_bar[s], ((uint32_t)(kb / stages)) & 1u);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b0 = a0 + a_bytes;
        const uint64_t ahi = smem_desc_sw128(a0), bhi = smem_desc_sw128(b0);
        const uint64_t alo = smem_desc_sw128(a0 + A_TILE), blo = smem_desc_sw128(b0 + NT * 128);
#pragma unroll
        for (int k = 0; k < CK / 16; ++k) {
          mma_bf16_ss(tmem, ahi + (uint64_t)(2 * k), bhi + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          if (MODE == SGTA_MMA_F32X3) {
            mma_bf16_ss(tmem, alo + (uint64_t)(2 * k), bhi + (uint64_t)(2 * k), idesc, 1u);
            mma_bf16_ss(tmem, ahi + (uint64_t)(2 * k), blo + (uint64_t)(2 * k), idesc, 1u);
          }
        }
        mma_commit(&empty_bar[s]);
      }
      mma_commit(accum_bar);
    }
  } else {
    // ================================================================== weight loader
    if (lane == 0) {
      const unsigned char* wsrc = a.wpack + (size_t)blockIdx.y * nkb * b_bytes;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % stages;
        mbar_wait(&empty_bar[s], (((uint32_t)(kb / stages)) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)b_bytes);
        bulk_g2s(smem + (size_t)s * stage_bytes + a_bytes, wsrc + (size_t)kb * b_bytes, (uint32_t)b_bytes,
                 &full_bar[s]);
      }
    }
  }
  __syncthreads();
  if (warp == PROD_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(a.tmem_cols) : "memory");
  }
}

// Wm [Cout][Kpad] fp32 (K ordered tap-major / channel-minor, zero padded to a multiple of 64)
// -> wpack[n tile][K block][hi tile | lo tile], each tile an NT x 64 SWIZZLE_128B image
__global__ void pack_weight_kernel(const float* __restrict__ wm, unsigned char* __restrict__ wpack,
                                   int Cout, int Kpad, int NT, int mode) {
  const int nt2 = mode == SGTA_MMA_F32X3 ? 2 : 1;
  const int nkb = Kpad / CK;
  const long long total = (long long)Cout * Kpad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int k = (int)(e % Kpad), n = (int)(e / Kpad);
    int kb = k / CK, kk = k % CK, t = n / NT, row = n % NT;
    float v = wm[e];
    __nv_bfloat16 hi = __float2bfloat16_rn(v);
    size_t base = ((size_t)t * nkb + kb) * (size_t)NT * 128 * nt2;
    size_t off = sw128_offset(row, kk / 8) + (kk % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(wpack + base + off) = hi;
    if (nt2 == 2)
      *reinterpret_cast<__nv_bfloat16*>(wpack + base + (size_t)NT * 128 + off) =
          __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

struct Plan { int ntile, stages, smem, tmem_cols; };
static bool make_plan(int Cout, int mode, bool dcn, Plan& pl) {
  if (Cout < 16 || Cout % 16) return false;
  int cap = mode == SGTA_MMA_F32X3 ? 128 : 256;
  int nt = Cout;
  if (nt > cap) {
    nt = cap;
    while (Cout % nt) nt -= 16;
  }
  int nt2 = mode == SGTA_MMA_F32X3 ? 2 : 1;
  int stage = (A_TILE + nt * 128) * nt2;
  int fixed = 1024 + (dcn ? CM * OMS * 4 : 256) + 256;
  int stages = (SMEM_MAX - fixed) / stage;
  if (stages > 6) stages = 6;
  if (stages < 2) return false;
  pl.ntile = nt; pl.stages = stages; pl.smem = stages * stage + fixed;
  pl.tmem_cols = 32;
  while (pl.tmem_cols < nt) pl.tmem_cols <<= 1;
  return true;
}

template <int PROD>
static int launch_conv(const ConvArgs& a, int mode, int smem, cudaStream_t st) {
  dim3 grid(cdiv(a.Mtot, CM), a.Cout / a.ntile);
  if (mode == SGTA_MMA_BF16) {
    cudaFuncSetAttribute(conv_umma_kernel<PROD, SGTA_MMA_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    conv_umma_kernel<PROD, SGTA_MMA_BF16><<<grid, CONV_THREADS, smem, st>>>(a);
  } else {
    cudaFuncSetAttribute(conv_umma_kernel<PROD, SGTA_MMA_F32X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    conv_umma_kernel<PROD, SGTA_MMA_F32X3><<<grid, CONV_THREADS, smem, st>>>(a);
  }
  return check_launch("conv_umma_kernel");
}

}  // namespace sgta

using namespace sgta;

extern "C" int sgta_conv_ntile(int Cout, int mode) {
  Plan pl;
  return make_plan(Cout, mode, false, pl) ? pl.ntile : -1;
}

extern "C" int64_t sgta_conv_wpack_bytes(int Cout, int Kpad, int mode) {
  Plan pl;
  if (!make_plan(Cout, mode, false, pl) || Kpad <= 0 || Kpad % 64) return -1;
  return (int64_t)Cout * Kpad * 2 * (mode == SGTA_MMA_F32X3 ? 2 : 1);
}

extern "C" int sgta_conv_pack_weight(const void* wm_f32, void* wpack, int Cout, int Kpad, int mode,
                                     void* stream) {
  SGTA_REQUIRE(wm_f32 && wpack, "sgta_conv_pack_weight: null pointer");
  SGTA_REQUIRE(mode == SGTA_MMA_BF16 || mode == SGTA_MMA_F32X3, "sgta_conv_pack_weight: bad mode");
  Plan pl;
  SGTA_REQUIRE(make_plan(Cout, mode, false, pl) && Kpad > 0 && Kpad % 64 == 0,
               "sgta_conv_pack_weight: need Cout %% 16 == 0 and Kpad %% 64 == 0 (got %d, %d)", Cout, Kpad);
  long long total = (long long)Cout * Kpad;
  int blocks = cdiv(total, 256) > 2368 ? 2368 : cdiv(total, 256);
  pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)wm_f32, (unsigned char*)wpack, Cout,
                                                             Kpad, pl.ntile, mode);
  return check_launch("pack_weight_kernel");
}

extern "C" int sgta_conv_forward_nhwc(const void* x, int64_t ldx, const void* wpack, const void* scale,
                                      const void* shift, const void* res, int64_t ldres, void* y,
                                      int64_t ldy, int B, int H, int W, int Cin, int Cout, int kh, int kw,
                                      int stride, int pad, int mode, int act, int out_dtype, int res_dtype,
                                      int epi, int n_valid, void* stream) {
  SGTA_REQUIRE(x && wpack && scale && shift && y, "sgta_conv_forward_nhwc: null pointer");
  SGTA_REQUIRE(B > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, "sgta_conv_forward_nhwc: bad shape");
  SGTA_REQUIRE(mode == SGTA_MMA_BF16 || mode == SGTA_MMA_F32X3, "sgta_conv_forward_nhwc: bad mode");
  SGTA_REQUIRE(Cin >= 4 && (Cin % 64 == 0 || ((Cin & (Cin - 1)) == 0 && Cin <= 32)),
               "sgta_conv_forward_nhwc: Cin must be a multiple of 64 or a power of two in [4,32] (got %d)", Cin);
  SGTA_REQUIRE(kh * kw <= 64, "sgta_conv_forward_nhwc: filter too large");
  Plan pl;
  SGTA_REQUIRE(make_plan(Cout, mode, false, pl), "sgta_conv_forward_nhwc: need Cout %% 16 == 0 (got %d)", Cout);
  SGTA_REQUIRE(epi >= 0 && epi <= 2 && (epi != EPI_STEM || Cout == 32), "sgta_conv_forward_nhwc: bad epilogue");
  int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  SGTA_REQUIRE(Ho > 0 && Wo > 0, "sgta_conv_forward_nhwc: empty output");
  long long M = (long long)B * Ho * Wo;
  SGTA_REQUIRE(M < (1ll << 31) - CM, "sgta_conv_forward_nhwc: too many pixels");
  ConvArgs a{};
  a.x = x; a.ldx = ldx; a.om = nullptr; a.wpack = (const unsigned char*)wpack;
  a.scale = (const float*)scale; a.shift = (const float*)shift;
  a.res = res; a.ldres = ldres; a.y = y; a.ldy = ldy;
  a.Mtot = (int)M; a.H = H; a.W = W; a.Ho = Ho; a.Wo = Wo; a.Cin = Cin; a.Cout = Cout;
  a.kh = kh; a.kw = kw; a.stride = stride; a.pad = pad;
  a.nkb = (kh * kw * Cin + CK - 1) / CK;
  a.ntile = pl.ntile; a.stages = pl.stages; a.act = act;
  a.y_bf16 = out_dtype == SGTA_DTYPE_BF16; a.res_bf16 = res_dtype == SGTA_DTYPE_BF16;
  a.tmem_cols = pl.tmem_cols; a.epi = epi; a.n_valid = n_valid > 0 ? n_valid : Cout;
  if (Cin % 64 == 0) return launch_conv<PROD_C64>(a, mode, pl.smem, (cudaStream_t)stream);
  return launch_conv<PROD_GEN>(a, mode, pl.smem, (cudaStream_t)stream);
}

// ---- DCN entry points (weights [Cout][Cin][3][3] fp32; K order = (tap, channel)) -----------------
static bool dcn_nhwc_supported(int Cin, int Cout) {
  return Cin > 0 && Cin % 64 == 0 && Cout >= 16 && Cout <= 256 && Cout % 16 == 0;
}

__global__ void dcn_weight_to_matrix_kernel(const float* __restrict__ w, float* __restrict__ wm, int Cin,
                                            int Cout) {
  long long total = (long long)Cout * 9 * Cin;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % Cin), tap = (int)((e / Cin) % 9), n = (int)(e / ((long long)9 * Cin));
    wm[e] = w[((long long)n * Cin + c) * 9 + tap];
  }
}

extern "C" int64_t sgta_dcn_wpack_bytes(int Cin, int Cout, int mode) {
  if (!dcn_nhwc_supported(Cin, Cout)) return -1;
  // packed image + fp32 scratch matrix used only while packing
  return (int64_t)9 * Cin * Cout * 2 * (mode == SGTA_MMA_F32X3 ? 2 : 1) + (int64_t)9 * Cin * Cout * 4;
}

extern "C" int sgta_dcn_pack_weight(const void* weight_f32, void* wpack, int Cin, int Cout, int mode,
                                    void* stream) {
  SGTA_REQUIRE(weight_f32 && wpack, "sgta_dcn_pack_weight: null pointer");
  SGTA_REQUIRE(dcn_nhwc_supported(Cin, Cout), "sgta_dcn_pack_weight: need Cin %% 64 == 0, Cout %% 16 == 0, Cout <= 256");
  SGTA_REQUIRE(mode == SGTA_MMA_BF16 || mode == SGTA_MMA_F32X3, "sgta_dcn_pack_weight: bad mode");
  size_t packed = (size_t)9 * Cin * Cout * 2 * (mode == SGTA_MMA_F32X3 ? 2 : 1);
  float* wm = reinterpret_cast<float*>((unsigned char*)wpack + packed);
  long long total = (long long)9 * Cin * Cout;
  int blocks = cdiv(total, 256) > 2368 ? 2368 : cdiv(total, 256);
  dcn_weight_to_matrix_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)weight_f32, wm, Cin, Cout);
  int rc = check_launch("dcn_weight_to_matrix_kernel");
  if (rc) return rc;
  return sgta_conv_pack_weight(wm, wpack, Cout, 9 * Cin, mode, stream);
}

extern "C" int sgta_dcn_forward_nhwc(const void* x, const void* offset_mask, const void* wpack,
                                     const void* scale, const void* shift, void* y, int B, int Cin,
                                     int Cout, int H, int W, int mode, int relu, int out_dtype,
                                     void* stream) {
  SGTA_REQUIRE(x && offset_mask && wpack && scale && shift && y, "sgta_dcn_forward_nhwc: null pointer");
  SGTA_REQUIRE(B > 0 && H > 0 && W > 0, "sgta_dcn_forward_nhwc: bad shape");
  SGTA_REQUIRE(dcn_nhwc_supported(Cin, Cout),
               "sgta_dcn_forward_nhwc: need Cin %% 64 == 0, Cout %% 16 == 0, 16 <= Cout <= 256 (got %d -> %d)", Cin, Cout);
  SGTA_REQUIRE(mode == SGTA_MMA_BF16 || mode == SGTA_MMA_F32X3, "sgta_dcn_forward_nhwc: bad mode");
  SGTA_REQUIRE(out_dtype == SGTA_DTYPE_F32 || out_dtype == SGTA_DTYPE_BF16, "sgta_dcn_forward_nhwc: bad out dtype");
  long long M = (long long)B * H * W;
  SGTA_REQUIRE(M < (1ll << 31) - CM, "sgta_dcn_forward_nhwc: too many pixels");
  Plan pl;
  SGTA_REQUIRE(make_plan(Cout, mode, true, pl), "sgta_dcn_forward_nhwc: no tiling for Cout=%d", Cout);
  ConvArgs a{};
  a.x = x; a.ldx = Cin; a.om = (const float*)offset_mask; a.wpack = (const unsigned char*)wpack;
  a.scale = (const float*)scale; a.shift = (const float*)shift; a.res = nullptr; a.ldres = 0;
  a.y = y; a.ldy = Cout;
  a.Mtot = (int)M; a.H = H; a.W = W; a.Ho = H; a.Wo = W; a.Cin = Cin; a.Cout = Cout;
  a.kh = 3; a.kw = 3; a.stride = 1; a.pad = 1; a.nkb = 9 * (Cin / CK);
  a.ntile = pl.ntile; a.stages = pl.stages; a.act = relu ? ACT_RELU : ACT_NONE;
  a.y_bf16 = out_dtype == SGTA_DTYPE_BF16; a.res_bf16 = 0; a.tmem_cols = pl.tmem_cols;
  a.epi = EPI_NHWC; a.n_valid = Cout;
  return launch_conv<PROD_DCN>(a, mode, pl.smem, (cudaStream_t)stream);
}
