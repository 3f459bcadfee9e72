// Modulated deformable convolution v2 as an implicit GEMM on tcgen05 / TMEM (sm_100a).
//
// Replaces, for the one configuration reference sgtapose/lib/model/networks/dla.py:545
// constructs (3x3, stride 1, pad 1, dil 1, one deformable group), upstream DCNv2's
// `modulated_deformable_im2col` + cuBLAS GEMM pair, which writes the [9*Cin, H*W] column
// matrix to HBM and reads it back.  Here:
//
//   out[p, o] = act( scale[o] * sum_{tap, c} A[p, (tap,c)] * Wt[(tap,c), o] + shift[o] )
//   A[p, (tap,c)] = sigmoid(mask[p,tap]) * bilinear(x[:, c], p + tap + offset[p,tap])
//
// * activations are NHWC, so one (pixel, tap) sample is four contiguous channel vectors;
// * a CTA owns 128 output pixels x all Cout channels; 8 producer warps gather the four
//   corners with 16-byte loads, blend them in fp32, and write the A tile straight into shared
//   memory in the SWIZZLE_128B K-major layout the UMMA descriptor expects (the column matrix
//   never exists in global memory);
// * the weight tile of each K block (one tap x 64 channels) is a pre-swizzled shared-memory
//   image in global memory, fetched by one bulk async copy (TMA engine, UBLKCP) that signals
//   the same mbarrier the producers arrive on;
// * one elected thread issues tcgen05.mma (M=128, N=Cout, K=16) into a TMEM accumulator;
//   tcgen05.commit releases pipeline stages and finally hands the accumulator to the epilogue;
// * the epilogue reads TMEM with tcgen05.ld, applies the folded bias / eval-BatchNorm scale
//   and shift and the ReLU of DeformConv (dla.py:547-550), and writes NHWC fp32 or bf16.
//
// Modes: BF16  -- bf16 operands, fp32 accumulate (1 MMA per K step);
//        F32X3 -- fp32 input, both operands split hi/lo into bf16 pairs, 3 MMAs per K step
//                 (hi*hi + lo*hi + hi*lo), ~2^-16 relative: the fp32 parity mode.
#include "common.cuh"
#include "umma.cuh"

namespace sgta {
using namespace umma;

constexpr int DM = 128;           // pixels per CTA (UMMA M)
constexpr int DK = 64;            // channels per K block (one 128-byte swizzle row of bf16)
constexpr int PROD_WARPS = 8;
constexpr int DCN_THREADS = (PROD_WARPS + 2) * 32;   // + MMA warp + weight-loader warp
constexpr int OMS = 33;           // padded row stride of the offset/mask tile in smem
constexpr int A_TILE = DM * 128;  // bytes

struct DcnSmemLayout {
  int stages, a_bytes, b_bytes, stage_bytes, total;
};
static DcnSmemLayout dcn_layout(int Cout, int mode) {
  DcnSmemLayout L;
  int nt = mode == SGTA_MMA_F32X3 ? 2 : 1;
  L.a_bytes = A_TILE * nt;
  L.b_bytes = Cout * 128 * nt;
  L.stage_bytes = L.a_bytes + L.b_bytes;
  int budget = 200 * 1024 - DM * OMS * 4 - 1024;
  L.stages = budget / L.stage_bytes;
  if (L.stages > 6) L.stages = 6;
  L.total = L.stages * L.stage_bytes + DM * OMS * 4 + 1024 /*align slack*/ + 256 /*barriers*/;
  return L;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(u[i] << 16);
    f[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}

template <int MODE>
__global__ void __launch_bounds__(DCN_THREADS, 1)
dcn_umma_kernel(const void* __restrict__ xin, const float* __restrict__ om,
                const unsigned char* __restrict__ wpack, const float* __restrict__ scale,
                const float* __restrict__ shift, void* __restrict__ yout, int Mtot, int H, int W,
                int Cin, int Cout, int stages, int relu, int out_bf16, int tmem_cols) {
  constexpr int NT = MODE == SGTA_MMA_F32X3 ? 2 : 1;
  extern __shared__ unsigned char dcn_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(dcn_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = A_TILE * NT, b_bytes = Cout * 128 * NT;
  const int stage_bytes = a_bytes + b_bytes;
  float* om_s = reinterpret_cast<float*>(smem + (size_t)stages * stage_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(om_s + DM * OMS);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* accum_bar = empty_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile0 = blockIdx.x * DM;
  const int cchunks = Cin / DK;
  const int nkb = 9 * cchunks;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], PROD_WARPS + 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == PROD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // stage the raw offset/mask tile: [128 pixels][32 floats] contiguous in NHWC
  if (warp < PROD_WARPS) {
    for (int e = tid; e < DM * 32; e += PROD_WARPS * 32) {
      int r = e >> 5, c = e & 31;
      int p = tile0 + r;
      om_s[r * OMS + c] = (p < Mtot) ? __ldg(om + (size_t)p * 32 + c) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < PROD_WARPS) {
    // ------------------------------------------------------------------ A producers
    const int r = tid >> 1, half = tid & 1;          // pixel row in the tile, which 32-channel half
    const int p = tile0 + r;
    const bool pvalid = p < Mtot;
    const int pp = pvalid ? p : 0;
    const int px = pp % W, py = (pp / W) % H, pb = pp / (W * H);
    const size_t img_base = (size_t)pb * H * W;
    int kb = 0;
    for (int tap = 0; tap < 9; ++tap) {
      float sy = (float)(py - 1 + tap / 3) + om_s[r * OMS + 2 * tap];
      float sx = (float)(px - 1 + tap % 3) + om_s[r * OMS + 2 * tap + 1];
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
      const size_t o00 = (img_base + (size_t)yc0 * W + xc0) * Cin + half * 32;
      const size_t o01 = (img_base + (size_t)yc0 * W + xc1) * Cin + half * 32;
      const size_t o10 = (img_base + (size_t)yc1 * W + xc0) * Cin + half * 32;
      const size_t o11 = (img_base + (size_t)yc1 * W + xc1) * Cin + half * 32;
      for (int cc = 0; cc < cchunks; ++cc, ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (uint32_t)(kb / stages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        unsigned char* sA = smem + (size_t)s * stage_bytes;
        if (MODE == SGTA_MMA_BF16) {
          const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xin) + cc * DK;
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
            float a[8], b[8], c[8], d[8];
            unpack8(c00[j], a); unpack8(c01[j], b); unpack8(c10[j], c); unpack8(c11[j], d);
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float v0 = w00 * a[2 * i] + w01 * b[2 * i] + w10 * c[2 * i] + w11 * d[2 * i];
              float v1 = w00 * a[2 * i + 1] + w01 * b[2 * i + 1] + w10 * c[2 * i + 1] + w11 * d[2 * i + 1];
              o[i] = pack_bf16(v0, v1);
            }
            *reinterpret_cast<uint4*>(sA + sw128_offset(r, half * 4 + j)) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        } else {
          const float* xb = reinterpret_cast<const float*>(xin) + cc * DK;
#pragma unroll
          for (int j = 0; j < 4; ++j) {            // 8 channels per step (two float4 per corner)
            float v[8];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              float4 a = __ldg(reinterpret_cast<const float4*>(xb + o00) + 2 * j + q);
              float4 b = __ldg(reinterpret_cast<const float4*>(xb + o01) + 2 * j + q);
              float4 c = __ldg(reinterpret_cast<const float4*>(xb + o10) + 2 * j + q);
              float4 d = __ldg(reinterpret_cast<const float4*>(xb + o11) + 2 * j + q);
              v[4 * q + 0] = w00 * a.x + w01 * b.x + w10 * c.x + w11 * d.x;
              v[4 * q + 1] = w00 * a.y + w01 * b.y + w10 * c.y + w11 * d.y;
              v[4 * q + 2] = w00 * a.z + w01 * b.z + w10 * c.z + w11 * d.z;
              v[4 * q + 3] = w00 * a.w + w01 * b.w + w10 * c.w + w11 * d.w;
            }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
              float r0 = v[2 * i] - __bfloat162float(h0), r1 = v[2 * i + 1] - __bfloat162float(h1);
              hi[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
              lo[i] = pack_bf16(r0, r1);
            }
            const uint32_t off = sw128_offset(r, half * 4 + j);
            *reinterpret_cast<uint4*>(sA + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(sA + A_TILE + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
      }
    }
    // ------------------------------------------------------------------ epilogue
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3, hsel = warp >> 2;
    const int row = q * 32 + lane;
    const int po = tile0 + row;
    for (int c0 = hsel * 32; c0 < Cout; c0 += 64) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (po < Mtot) {
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float t = fmaf(__uint_as_float(v[j]), __ldg(scale + c0 + j), __ldg(shift + c0 + j));
          o[j] = relu ? fmaxf(t, 0.f) : t;
        }
        if (out_bf16) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(yout) + (size_t)po * Cout + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                                pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
        } else {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(yout) + (size_t)po * Cout + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        }
      }
    }
    tc_fence_before();
  } else if (warp == PROD_WARPS) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_f32(DM, Cout);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (uint32_t)(kb / stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b0 = a0 + a_bytes;
        const uint64_t ahi = smem_desc_sw128(a0), bhi = smem_desc_sw128(b0);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) {
          const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
          mma_bf16_ss(tmem, ahi + (uint64_t)(2 * k), bhi + (uint64_t)(2 * k), idesc, acc);
          if (MODE == SGTA_MMA_F32X3) {
            const uint64_t alo = smem_desc_sw128(a0 + A_TILE), blo = smem_desc_sw128(b0 + Cout * 128);
            mma_bf16_ss(tmem, alo + (uint64_t)(2 * k), bhi + (uint64_t)(2 * k), idesc, 1u);
            mma_bf16_ss(tmem, ahi + (uint64_t)(2 * k), blo + (uint64_t)(2 * k), idesc, 1u);
          }
        }
        mma_commit(&empty_bar[s]);
      }
      mma_commit(accum_bar);
    }
  } else {
    // ------------------------------------------------------------------ weight loader
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (uint32_t)(kb / stages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)b_bytes);
        bulk_g2s(smem + (size_t)s * stage_bytes + a_bytes, wpack + (size_t)kb * b_bytes, (uint32_t)b_bytes,
                 &full_bar[s]);
      }
    }
  }
  __syncthreads();
  if (warp == PROD_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
  }
}

// weight [Cout][Cin][3][3] fp32 -> per K block (tap, 64-channel chunk) SW128 images
__global__ void dcn_pack_weight_kernel(const float* __restrict__ w, unsigned char* __restrict__ wpack,
                                       int Cin, int Cout, int mode) {
  const int nt = mode == SGTA_MMA_F32X3 ? 2 : 1;
  const int cchunks = Cin / DK;
  const long long total = (long long)9 * cchunks * Cout * DK;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int kk = (int)(e % DK);
    int n = (int)((e / DK) % Cout);
    int kb = (int)(e / ((long long)DK * Cout));
    int tap = kb / cchunks, cc = kb % cchunks;
    float v = w[((long long)n * Cin + cc * DK + kk) * 9 + tap];
    __nv_bfloat16 hi = __float2bfloat16_rn(v);
    size_t base = (size_t)kb * Cout * 128 * nt;
    size_t off = sw128_offset(n, kk / 8) + (kk % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(wpack + base + off) = hi;
    if (nt == 2)
      *reinterpret_cast<__nv_bfloat16*>(wpack + base + (size_t)Cout * 128 + off) =
          __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

}  // namespace sgta

using namespace sgta;

static bool dcn_nhwc_supported(int Cin, int Cout) {
  return Cin > 0 && Cin % 64 == 0 && Cout >= 32 && Cout <= 256 && Cout % 32 == 0;
}

extern "C" int64_t sgta_dcn_wpack_bytes(int Cin, int Cout, int mode) {
  if (!dcn_nhwc_supported(Cin, Cout)) return -1;
  return (int64_t)9 * (Cin / 64) * Cout * 128 * (mode == SGTA_MMA_F32X3 ? 2 : 1);
}

extern "C" int sgta_dcn_pack_weight(const void* weight_f32, void* wpack, int Cin, int Cout, int mode,
                                    void* stream) {
  SGTA_REQUIRE(weight_f32 && wpack, "sgta_dcn_pack_weight: null pointer");
  SGTA_REQUIRE(dcn_nhwc_supported(Cin, Cout), "sgta_dcn_pack_weight: need Cin %% 64 == 0, Cout %% 32 == 0, Cout <= 256");
  SGTA_REQUIRE(mode == SGTA_MMA_BF16 || mode == SGTA_MMA_F32X3, "sgta_dcn_pack_weight: bad mode");
  long long total = (long long)9 * Cin * Cout;
  dcn_pack_weight_kernel<<<cdiv(total, 256) > 1184 ? 1184 : cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const float*)weight_f32, (unsigned char*)wpack, Cin, Cout, mode);
  return check_launch("dcn_pack_weight_kernel");
}

extern "C" int sgta_dcn_forward_nhwc(const void* x, const void* offset_mask, const void* wpack,
                                     const void* scale, const void* shift, void* y, int B, int Cin,
                                     int Cout, int H, int W, int mode, int relu, int out_dtype,
                                     void* stream) {
  SGTA_REQUIRE(x && offset_mask && wpack && scale && shift && y, "sgta_dcn_forward_nhwc: null pointer");
  SGTA_REQUIRE(B > 0 && H > 0 && W > 0, "sgta_dcn_forward_nhwc: bad shape");
  SGTA_REQUIRE(dcn_nhwc_supported(Cin, Cout),
               "sgta_dcn_forward_nhwc: need Cin %% 64 == 0, Cout %% 32 == 0, 32 <= Cout <= 256 (got %d -> %d)", Cin, Cout);
  SGTA_REQUIRE(mode == SGTA_MMA_BF16 || mode == SGTA_MMA_F32X3, "sgta_dcn_forward_nhwc: bad mode");
  SGTA_REQUIRE(out_dtype == SGTA_DTYPE_F32 || out_dtype == SGTA_DTYPE_BF16, "sgta_dcn_forward_nhwc: bad out dtype");
  long long M = (long long)B * H * W;
  SGTA_REQUIRE(M < (1ll << 31) - DM, "sgta_dcn_forward_nhwc: too many pixels");
  DcnSmemLayout L = dcn_layout(Cout, mode);
  SGTA_REQUIRE(L.stages >= 2, "sgta_dcn_forward_nhwc: not enough shared memory for 2 stages");
  int tmem_cols = 32;
  while (tmem_cols < Cout) tmem_cols <<= 1;
  int grid = cdiv(M, DM);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == SGTA_MMA_BF16) {
    cudaFuncSetAttribute(dcn_umma_kernel<SGTA_MMA_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    dcn_umma_kernel<SGTA_MMA_BF16><<<grid, DCN_THREADS, L.total, st>>>(
        x, (const float*)offset_mask, (const unsigned char*)wpack, (const float*)scale, (const float*)shift, y,
        (int)M, H, W, Cin, Cout, L.stages, relu, out_dtype == SGTA_DTYPE_BF16, tmem_cols);
  } else {
    cudaFuncSetAttribute(dcn_umma_kernel<SGTA_MMA_F32X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    dcn_umma_kernel<SGTA_MMA_F32X3><<<grid, DCN_THREADS, L.total, st>>>(
        x, (const float*)offset_mask, (const unsigned char*)wpack, (const float*)scale, (const float*)shift, y,
        (int)M, H, W, Cin, Cout, L.stages, relu, out_dtype == SGTA_DTYPE_BF16, tmem_cols);
  }
  return check_launch("dcn_umma_kernel");
}
