// Structure-prior temporal attention core: fused QK^T/scale + pos_embed -> softmax -> V.
//
// Replaces the einsum / softmax / einsum triple of MHCA_ein.forward (reference
// sgtapose/lib/model/networks/dla.py:878-885), which materialises the [B,8,n,n] energy
// and attention tensors (45 MB per sample at level 0) in HBM.  Here one warp owns one
// query row of one head: its 32 lanes stride over the keys with an online softmax
// (running max / sum in the log2 domain), K and V of the (sample, head) live in shared
// memory, pos_embed rows are read once, coalesced, straight from L2/HBM, and the lane
// partials are merged with warp-shuffle reductions.  Nothing n x n is ever stored.
//
// Layout: q,k,v,out are the Linear outputs as they are, [B, n, heads*d] ("b n (h d)").
#include <stdlib.h>

#include "common.cuh"

namespace sgta {

constexpr int ATT_WARPS = 8;
constexpr int ATT_ROWS = 32;          // query rows per CTA (4 per warp)
constexpr float LOG2E = 1.4426950408889634f;

template <int D> struct KPad { static constexpr int stride = (D == 4) ? 4 : D + 4; };

template <int D>
__device__ __forceinline__ void load_row(float (&dst)[D], const float* __restrict__ src) {
#pragma unroll
  for (int i = 0; i < D; i += 4) {
    float4 t = *reinterpret_cast<const float4*>(src + i);
    dst[i] = t.x; dst[i + 1] = t.y; dst[i + 2] = t.z; dst[i + 3] = t.w;
  }
}

template <int D>
__device__ __forceinline__ void stage_kv(float* Ks, float* Vs, const float* __restrict__ kb,
                                         const float* __restrict__ vb, int nk, int HD, int h) {
  constexpr int S = KPad<D>::stride;
  constexpr int V4 = D / 4;
  for (int e = threadIdx.x; e < nk * V4; e += blockDim.x) {
    int j = e / V4, c = (e % V4) * 4;
    float4 kk = __ldg(reinterpret_cast<const float4*>(kb + (long long)j * HD + h * D + c));
    float4 vv = __ldg(reinterpret_cast<const float4*>(vb + (long long)j * HD + h * D + c));
    *reinterpret_cast<float4*>(Ks + j * S + c) = kk;
    *reinterpret_cast<float4*>(Vs + j * S + c) = vv;
  }
}

// grid (ceil(nq/ATT_ROWS), heads, B), 256 threads, smem 2*nk*stride floats
template <int D>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                const float* __restrict__ v, const float* __restrict__ pos,
                float* __restrict__ out, int heads, int nq, int nk, float inv_scale) {
  constexpr int S = KPad<D>::stride;
  extern __shared__ __align__(16) float att_smem[];
  float* Ks = att_smem;
  float* Vs = att_smem + (size_t)nk * S;
  const int h = blockIdx.y, b = blockIdx.z;
  const int HD = heads * D;
  pdl_trigger();
  pdl_wait();
  stage_kv<D>(Ks, Vs, k + (long long)b * nk * HD, v + (long long)b * nk * HD, nk, HD, h);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float sc = inv_scale * LOG2E;
  for (int r = warp; r < ATT_ROWS; r += ATT_WARPS) {
    const int i = blockIdx.x * ATT_ROWS + r;
    if (i >= nq) break;
    float qr[D];
    load_row<D>(qr, q + ((long long)b * nq + i) * HD + h * D);
#pragma unroll
    for (int c = 0; c < D; ++c) qr[c] *= sc;
    const float* prow = pos ? pos + ((long long)h * nq + i) * nk : nullptr;
    float m = -INFINITY, l = 0.f, acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.f;
    for (int j = lane; j < nk; j += 32) {
      float kr[D];
      load_row<D>(kr, Ks + j * S);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) s = fmaf(qr[c], kr[c], s);
      if (prow) s = fmaf(__ldg(prow + j), LOG2E, s);
      if (s > m) {
        float corr = exp2f(m - s);
        l *= corr;
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] *= corr;
        m = s;
      }
      float p = exp2f(s - m);
      l += p;
      float vr[D];
      load_row<D>(vr, Vs + j * S);
#pragma unroll
      for (int c = 0; c < D; ++c) acc[c] = fmaf(p, vr[c], acc[c]);
    }
    float M = warp_max(m);
    float f = (m == -INFINITY) ? 0.f : exp2f(m - M);
    l = warp_sum(l * f);
    float inv = 1.f / l;
    float* orow = out + ((long long)b * nq + i) * HD + h * D;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      float a = warp_sum(acc[c] * f);
      if (lane == c) orow[c] = a * inv;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Forward, batch-looping variant (the one the engine runs).  pos_embed[h, rows, :] is shared by every
// sample of the batch: re-reading it per sample costs B x 44.8 MB of L2 traffic per launch at level 0
// (1.4 GB at B = 32, as long as the whole old kernel).  Here a CTA owns (32 query rows, head) and
// keeps that pos block in shared memory (151 KB at n = 1183) while it loops over its samples; K / V
// of the next sample are prefetched into registers during the current sample's math.  A warp
// processes RW query rows per pass so each K / V shared-memory read feeds RW rows, the running max is
// only rescaled when some key of the pass raises it, and exp2 is the bare MUFU instruction.
template <int D> struct RowsPerPass { static constexpr int value = D == 4 ? 4 : D == 8 ? 2 : 1; };

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr float LAZY = 8.f;
constexpr int ATB_WARPS = 16;          // batched kernel: 4 warps per SM sub-partition hide the exp2 -> fma chains
constexpr int ATB_THREADS = ATB_WARPS * 32;

template <int D>
__global__ void __launch_bounds__(ATB_THREADS, 1)
attn_fwd_batched_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                        const float* __restrict__ pos, float* __restrict__ out, int B, int heads, int nq, int nk,
                        int nk_pad, int b_per_cta, float inv_scale) {
  constexpr int S = KPad<D>::stride;
  constexpr int V4 = D / 4;
  constexpr int RW = RowsPerPass<D>::value;
  constexpr int GROUPS = ATT_ROWS / RW;                               // row groups per CTA
  constexpr int KS = ATB_WARPS / GROUPS >= 1 ? ATB_WARPS / GROUPS : 1;   // warps that split the keys of one group
  constexpr int WG = ATB_WARPS / KS;                                  // row groups processed concurrently
  constexpr int PF = 5;                       // float4 pairs of K/V prefetch per thread (covers nk*V4 <= 2560)
  extern __shared__ __align__(16) float att_smem[];
  float* Ks = att_smem;
  float* Vs = Ks + (size_t)nk * S;
  float* Part = Vs + (size_t)nk * S;         // [KS][ATT_ROWS][D + 2] partial (acc, m, l) per key split
  float* Qs = Part + KS * ATT_ROWS * (D + 2);   // [ATT_ROWS][D] query rows of the current sample (prefetched like K / V:
                                                // read straight from global they cost one exposed miss per sample and warp)
  float* Ps = Qs + ATT_ROWS * D;                // [ATT_ROWS][nk_pad]
  const int h = blockIdx.y;
  const int HD = heads * D;
  const int row0 = blockIdx.x * ATT_ROWS;
  const int b_begin = blockIdx.z * b_per_cta, b_end = min(B, b_begin + b_per_cta);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nvec = nk * V4;
  const int grp = warp % WG, ks = warp / WG;

  pdl_trigger();
  if (pos) {
    for (int r = warp; r < ATT_ROWS; r += ATB_WARPS) {
      const int i = row0 + r;
      if (i >= nq) break;
      const float* prow = pos + ((long long)h * nq + i) * nk;
      for (int j = lane; j < nk; j += 32) Ps[r * nk_pad + j] = __ldg(prow + j) * LOG2E;
    }
  }
  pdl_wait();                                   // pos_embed is a parameter; q / k / v come from the preceding kernels
  float4 kreg[PF], vreg[PF], qreg = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&](int b) {
    const float* kb = k + (long long)b * nk * HD + h * D;
    const float* vb = v + (long long)b * nk * HD + h * D;
    if (tid < ATT_ROWS * V4) {
      const int i = min(row0 + tid / V4, nq - 1);       // rows past the end recompute the last row (not stored)
      qreg = __ldg(reinterpret_cast<const float4*>(q + ((long long)b * nq + i) * HD + h * D + (tid % V4) * 4));
    }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int e = tid + u * ATB_THREADS;
      if (e < nvec) {
        const int j = e / V4, c = (e % V4) * 4;
        kreg[u] = __ldg(reinterpret_cast<const float4*>(kb + (long long)j * HD + c));
        vreg[u] = __ldg(reinterpret_cast<const float4*>(vb + (long long)j * HD + c));
      }
    }
  };
  if (b_begin < b_end) prefetch(b_begin);
  const float sc = inv_scale * LOG2E;

  for (int b = b_begin; b < b_end; ++b) {
    __syncthreads();                          // previous sample: K / V consumed, partials merged
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int e = tid + u * ATB_THREADS;
      if (e < nvec) {
        const int j = e / V4, c = (e % V4) * 4;
        *reinterpret_cast<float4*>(Ks + j * S + c) = kreg[u];
        *reinterpret_cast<float4*>(Vs + j * S + c) = vreg[u];
      }
    }
    if (tid < ATT_ROWS * V4) *reinterpret_cast<float4*>(Qs + tid * 4) = qreg;
    __syncthreads();
    if (b + 1 < b_end) prefetch(b + 1);

    for (int g = grp; g < GROUPS; g += WG) {
      const int r = g * RW;
      float qr[RW][D], m[RW], l[RW], acc[RW][D];
#pragma unroll
      for (int w = 0; w < RW; ++w) {
        load_row<D>(qr[w], Qs + (r + w) * D);
        m[w] = -INFINITY; l[w] = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) { qr[w][c] *= sc; acc[w][c] = 0.f; }
      }
      const float* prow = Ps + (size_t)r * nk_pad;
      // lazy rescale: a row's reference point mt - LAZY only moves when a score exceeds it by 2^LAZY;
      // exponentials stay <= 2^LAZY (exact in the final ratio).  The rescale is behind a WARP-UNIFORM
      // vote so it stays a real, rarely taken branch (a per-lane `if` is if-converted by the compiler
      // and its ex2 + D+1 multiplies then run on every key)
      float mt[RW];
#pragma unroll
      for (int w = 0; w < RW; ++w) mt[w] = -INFINITY;
#pragma unroll 2
      for (int jb = ks * 32; jb < nk; jb += 32 * KS) {       // warp-uniform trip count (the vote below)
        const bool live = jb + lane < nk;
        const int j = live ? jb + lane : nk - 1;
        float kr[D], vr[D], sv[RW];
        load_row<D>(kr, Ks + j * S);
        load_row<D>(vr, Vs + j * S);
        bool up = false;
#pragma unroll
        for (int w = 0; w < RW; ++w) {
          sv[w] = pos ? prow[w * nk_pad + j] : 0.f;
#pragma unroll
          for (int c = 0; c < D; ++c) sv[w] = fmaf(qr[w][c], kr[c], sv[w]);
          up |= live && sv[w] > mt[w];
        }
        if (__any_sync(0xffffffffu, up)) {
          // the reference point is shared by the warp (the max over all 32 keys of this step), so it
          // converges after a few steps and the final merge needs no per-lane rescale
#pragma unroll
          for (int w = 0; w < RW; ++w) {
            if (__any_sync(0xffffffffu, live && sv[w] > mt[w])) {
              const float nm = fmaxf(warp_max(live ? sv[w] : -INFINITY), m[w]);
              const float corr = (m[w] == -INFINITY) ? 0.f : ex2_approx(m[w] - nm);
              l[w] *= corr;
#pragma unroll
              for (int c = 0; c < D; ++c) acc[w][c] *= corr;
              m[w] = nm;
              mt[w] = nm + LAZY;
            }
          }
        }
#pragma unroll
        for (int w = 0; w < RW; ++w) {
          const float pw = live ? ex2_approx(sv[w] - m[w]) : 0.f;
          l[w] += pw;
#pragma unroll
          for (int c = 0; c < D; ++c) acc[w][c] = fmaf(pw, vr[c], acc[w][c]);
        }
      }
#pragma unroll
      for (int w = 0; w < RW; ++w) {
        const float M = m[w];                                    // warp-uniform
        const float L = warp_sum(l[w]);
        float* pr = Part + ((size_t)ks * ATT_ROWS + r + w) * (D + 2);
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const float a = warp_sum(acc[w][c]);
          if (lane == c) pr[c] = a;
        }
        if (lane == 0) { pr[D] = M; pr[D + 1] = L; }
      }
    }
    __syncthreads();
    // merge the key splits: one thread per (row, channel)
    for (int e = tid; e < ATT_ROWS * D; e += ATB_THREADS) {
      const int r = e / D, c = e % D;
      if (row0 + r >= nq) continue;
      float M = -INFINITY;
#pragma unroll
      for (int s2 = 0; s2 < KS; ++s2) M = fmaxf(M, Part[((size_t)s2 * ATT_ROWS + r) * (D + 2) + D]);
      float L = 0.f, a = 0.f;
#pragma unroll
      for (int s2 = 0; s2 < KS; ++s2) {
        const float* pr = Part + ((size_t)s2 * ATT_ROWS + r) * (D + 2);
        const float f = pr[D] == -INFINITY ? 0.f : ex2_approx(pr[D] - M);
        L = fmaf(pr[D + 1], f, L);
        a = fmaf(pr[c], f, a);
      }
      out[((long long)b * nq + row0 + r) * HD + h * D + c] = a / L;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Forward, batch-looping variant with LANES = QUERY ROWS (D = 4: level 0, the one the engine runs).
// Same CTA shape as attn_fwd_batched_kernel -- (32 query rows, head), pos block resident, loop over the samples --
// but a warp's 32 lanes are the CTA's 32 query rows and the 16 warps split the KEYS: K / V rows are warp-wide
// broadcast reads (one LDS.128 each per key and warp), the pos value of (row, key) is one conflict-free LDS per lane
// (row stride = 1 mod 32), every lane runs its own online softmax, and nothing is reduced across lanes: the per-
// sample epilogue is 6 stores per lane and a 16-way merge by 128 threads, where the lanes-as-keys kernel spends 100
// shuffles + 100 adds per warp (41 % of its stall samples sat in that per-sample prologue / epilogue, ncu source page).
// KVHM: K / V are given head-major, [B, heads, nk, D] (sgta_token_linear_heads writes them that way): the slab of a
// (sample, head) is contiguous, so the per-sample prefetch is coalesced 16-byte loads instead of one 16-byte piece per
// 128-byte line ("b n (h d)": a third of this kernel's stall samples were LSU-queue throttling on those gathers).
// Measured and rejected (round 2, 0.985 ms per step for the three level-0 launches): K / V transposed in shared memory
// with the scores and P V as FFMA2 on key pairs (a third fewer instructions per key: 0.984 ms, no change -- the inner
// loop is ~2 us of the ~5 us a CTA spends per sample; the rest is the three barriers, the staging and the merge);
// merging the 16 slices with all 512 threads and xor-shuffles (1.035 ms with a conflict-free layout, 1.185 without).
template <int D, bool KVHM>
__global__ void __launch_bounds__(ATB_THREADS, 1)
attn_fwd_rows_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                     const float* __restrict__ pos, float* __restrict__ out, int B, int heads, int nq, int nk,
                     int nk_pad, int b_per_cta, float inv_scale) {
  static_assert(D == 4, "lanes-as-rows attention: head dim 4");
  constexpr int PF = 5;                       // float4 pairs of K/V prefetch per thread (covers nk <= 2560)
  constexpr int UN = 4;                       // keys per inner iteration
  extern __shared__ __align__(16) float att_smem[];
  float* Ks = att_smem;                                       // [nk][4]
  float* Vs = Ks + (size_t)nk * 4;                            // [nk][4]
  float* Part = Vs + (size_t)nk * 4;                          // [ATB_WARPS][ATT_ROWS][D + 2]: acc, m, l per key slice
  float* Qs = Part + ATB_WARPS * ATT_ROWS * (D + 2);          // [ATT_ROWS][D]
  float* Ps = Qs + ATT_ROWS * D;                              // [ATT_ROWS][nk_pad], pre-scaled by log2(e)
  const int h = blockIdx.y;
  const int HD = heads * D;
  const int row0 = blockIdx.x * ATT_ROWS;
  const int b_begin = blockIdx.z * b_per_cta, b_end = min(B, b_begin + b_per_cta);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  pdl_trigger();
  for (int r = warp; r < ATT_ROWS; r += ATB_WARPS) {
    const int i = min(row0 + r, nq - 1);
    const float* prow = pos ? pos + ((long long)h * nq + i) * nk : nullptr;
    for (int j = lane; j < nk; j += 32) Ps[r * nk_pad + j] = prow ? __ldg(prow + j) * LOG2E : 0.f;
  }
  pdl_wait();                                   // pos_embed is a parameter; q / k / v come from the preceding kernels
  float4 kreg[PF], vreg[PF], qreg = make_float4(0.f, 0.f, 0.f, 0.f);
  auto prefetch = [&](int b) {
    const long long kv0 = KVHM ? ((long long)b * heads + h) * nk * D : (long long)b * nk * HD + h * D;
    const int kvs = KVHM ? D : HD;                           // floats between consecutive keys of this (sample, head)
    const float* kb = k + kv0;
    const float* vb = v + kv0;
    if (tid < ATT_ROWS) {
      const int i = min(row0 + tid, nq - 1);              // rows past the end recompute the last row (not stored)
      qreg = __ldg(reinterpret_cast<const float4*>(q + ((long long)b * nq + i) * HD + h * D));
    }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int j = tid + u * ATB_THREADS;
      if (j < nk) {
        kreg[u] = __ldg(reinterpret_cast<const float4*>(kb + (long long)j * kvs));
        vreg[u] = __ldg(reinterpret_cast<const float4*>(vb + (long long)j * kvs));
      }
    }
  };
  if (b_begin < b_end) prefetch(b_begin);
  const float sc = inv_scale * LOG2E;
  // this warp's key slice (multiples of UN except the last one)
  const int per = ((nk + ATB_WARPS - 1) / ATB_WARPS + UN - 1) / UN * UN;
  const int j0 = min(warp * per, nk), j1 = min(j0 + per, nk);
  const float* prow = Ps + (size_t)lane * nk_pad;

  for (int b = b_begin; b < b_end; ++b) {
    __syncthreads();                          // previous sample: K / V consumed, partials merged
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int j = tid + u * ATB_THREADS;
      if (j < nk) {
        *reinterpret_cast<float4*>(Ks + j * 4) = kreg[u];
        *reinterpret_cast<float4*>(Vs + j * 4) = vreg[u];
      }
    }
    if (tid < ATT_ROWS) *reinterpret_cast<float4*>(Qs + tid * 4) = qreg;
    __syncthreads();
    if (b + 1 < b_end) prefetch(b + 1);

    const float4 q4 = *reinterpret_cast<const float4*>(Qs + lane * 4);
    const float q0 = q4.x * sc, q1 = q4.y * sc, q2 = q4.z * sc, q3 = q4.w * sc;
    float m = -INFINITY, mt = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    // lazy rescale: the reference point m only moves when a score exceeds it by 2^LAZY; the test is a warp vote
    // so that the rescale stays a real, rarely taken branch
    auto bump = [&](float smax) {
      if (__any_sync(0xffffffffu, smax > mt)) {
        if (smax > mt) {
          const float corr = (m == -INFINITY) ? 0.f : ex2_approx(m - smax);
          l *= corr; a0 *= corr; a1 *= corr; a2 *= corr; a3 *= corr;
          m = smax;
          mt = smax + LAZY;
        }
      }
    };
    int j = j0;
    for (; j + UN <= j1; j += UN) {
      float4 kk[UN], vv[UN];
      float s[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        kk[u] = *reinterpret_cast<const float4*>(Ks + (j + u) * 4);
        vv[u] = *reinterpret_cast<const float4*>(Vs + (j + u) * 4);
        s[u] = prow[j + u];
      }
#pragma unroll
      for (int u = 0; u < UN; ++u)
        s[u] = fmaf(q3, kk[u].w, fmaf(q2, kk[u].z, fmaf(q1, kk[u].y, fmaf(q0, kk[u].x, s[u]))));
      bump(fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3])));
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const float e = ex2_approx(s[u] - m);
        l += e;
        a0 = fmaf(e, vv[u].x, a0); a1 = fmaf(e, vv[u].y, a1); a2 = fmaf(e, vv[u].z, a2); a3 = fmaf(e, vv[u].w, a3);
      }
    }
    for (; j < j1; ++j) {                     // ragged tail of the last slice
      const float4 kk = *reinterpret_cast<const float4*>(Ks + j * 4);
      const float4 vv = *reinterpret_cast<const float4*>(Vs + j * 4);
      const float s1 = fmaf(q3, kk.w, fmaf(q2, kk.z, fmaf(q1, kk.y, fmaf(q0, kk.x, prow[j]))));
      bump(s1);
      const float e = ex2_approx(s1 - m);
      l += e;
      a0 = fmaf(e, vv.x, a0); a1 = fmaf(e, vv.y, a1); a2 = fmaf(e, vv.z, a2); a3 = fmaf(e, vv.w, a3);
    }
    float* pr = Part + ((size_t)warp * ATT_ROWS + lane) * (D + 2);
    pr[0] = a0; pr[1] = a1; pr[2] = a2; pr[3] = a3; pr[4] = m; pr[5] = l;
    __syncthreads();
    // merge the 16 key slices: one thread per (row, channel)
    if (tid < ATT_ROWS * D) {
      const int r = tid / D, c = tid % D;
      if (row0 + r < nq) {
        float M = -INFINITY;
#pragma unroll
        for (int s2 = 0; s2 < ATB_WARPS; ++s2) M = fmaxf(M, Part[((size_t)s2 * ATT_ROWS + r) * (D + 2) + D]);
        float L = 0.f, a = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < ATB_WARPS; ++s2) {
          const float* p2 = Part + ((size_t)s2 * ATT_ROWS + r) * (D + 2);
          const float f = p2[D] == -INFINITY ? 0.f : ex2_approx(p2[D] - M);
          L = fmaf(p2[D + 1], f, L);
          a = fmaf(p2[c], f, a);
        }
        out[((long long)b * nq + row0 + r) * HD + h * D + c] = a / L;
      }
    }
  }
}

// Backward of the same core.  One warp per query row; recomputes the softmax statistics,
// then ds_j = p_j (dp_j - delta) with delta = gout_i . out_i.
//   gq_i = inv_scale * sum_j ds_j k_j          (registers + warp reduce)
//   gk_j += inv_scale * ds_j q_i, gv_j += p_j gout_i, gpos[h,i,j] += ds_j   (atomicAdd)
template <int D>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k,
                const float* __restrict__ v, const float* __restrict__ pos,
                const float* __restrict__ gout, float* __restrict__ gq, float* __restrict__ gk,
                float* __restrict__ gv, float* __restrict__ gpos, int heads, int nq, int nk,
                float inv_scale) {
  constexpr int S = KPad<D>::stride;
  extern __shared__ __align__(16) float att_smem[];
  float* Ks = att_smem;
  float* Vs = att_smem + (size_t)nk * S;
  const int h = blockIdx.y, b = blockIdx.z;
  const int HD = heads * D;
  stage_kv<D>(Ks, Vs, k + (long long)b * nk * HD, v + (long long)b * nk * HD, nk, HD, h);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float sc = inv_scale * LOG2E;
  for (int r = warp; r < ATT_ROWS; r += ATT_WARPS) {
    const int i = blockIdx.x * ATT_ROWS + r;
    if (i >= nq) break;
    float qr[D], go[D];
    const long long rowoff = ((long long)b * nq + i) * HD + h * D;
    load_row<D>(qr, q + rowoff);
    load_row<D>(go, gout + rowoff);
    const float* prow = pos ? pos + ((long long)h * nq + i) * nk : nullptr;
    // pass 1: statistics and delta
    float m = -INFINITY, l = 0.f, dacc = 0.f;
    for (int j = lane; j < nk; j += 32) {
      float kr[D], vr[D];
      load_row<D>(kr, Ks + j * S);
      load_row<D>(vr, Vs + j * S);
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) { s = fmaf(qr[c] * sc, kr[c], s); dp = fmaf(go[c], vr[c], dp); }
      if (prow) s = fmaf(__ldg(prow + j), LOG2E, s);
      if (s > m) { float corr = exp2f(m - s); l *= corr; dacc *= corr; m = s; }
      float p = exp2f(s - m);
      l += p; dacc = fmaf(p, dp, dacc);
    }
    float M = warp_max(m);
    float f = (m == -INFINITY) ? 0.f : exp2f(m - M);
    float L = warp_sum(l * f);
    float delta = warp_sum(dacc * f) / L;
    float invL = 1.f / L;
    // pass 2
    float gqa[D];
#pragma unroll
    for (int c = 0; c < D; ++c) gqa[c] = 0.f;
    for (int j = lane; j < nk; j += 32) {
      float kr[D], vr[D];
      load_row<D>(kr, Ks + j * S);
      load_row<D>(vr, Vs + j * S);
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < D; ++c) { s = fmaf(qr[c] * sc, kr[c], s); dp = fmaf(go[c], vr[c], dp); }
      if (prow) s = fmaf(__ldg(prow + j), LOG2E, s);
      float p = exp2f(s - M) * invL;
      float ds = p * (dp - delta);
      if (gpos) atomicAdd(gpos + ((long long)h * nq + i) * nk + j, ds);
      float* gkr = gk + ((long long)b * nk + j) * HD + h * D;
      float* gvr = gv + ((long long)b * nk + j) * HD + h * D;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        gqa[c] = fmaf(ds, kr[c], gqa[c]);
        atomicAdd(gkr + c, ds * inv_scale * qr[c]);
        atomicAdd(gvr + c, p * go[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < D; ++c) {
      float a = warp_sum(gqa[c]);
      if (lane == c) gq[rowoff + c] = a * inv_scale;
    }
  }
}

// the batch-looping kernels serve LARGE pos_embed blocks (level 0: 44.8 MB re-read per sample otherwise)
static bool big_pos_case(const float* pos, int B, int heads, int nq, int nk) {
  static const size_t thr_mb = getenv("SGTA_ATTN_POS_MB") ? (size_t)atoi(getenv("SGTA_ATTN_POS_MB")) : 16;
  return pos != nullptr && (size_t)heads * nq * nk * sizeof(float) > (thr_mb << 20) && B >= 4;
}
static size_t rows_kernel_smem(int nk, int D) {
  const int nk_pad = nk + ((33 - (nk & 31)) & 31);
  return sizeof(float) * (2 * (size_t)nk * 4 + (size_t)ATB_WARPS * ATT_ROWS * (D + 2) + (size_t)ATT_ROWS * D +
                          (size_t)ATT_ROWS * nk_pad);
}

template <int D>
static int launch_fwd(const float* q, const float* k, const float* v, const float* pos, float* out,
                      int B, int heads, int nq, int nk, float inv_scale, cudaStream_t st, bool kvhm = false) {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  // batch-looping kernel for LARGE pos_embed blocks (level 0: 44.8 MB re-read per sample otherwise):
  // pos block + K/V of one sample in shared memory; the prefetch registers cover nk*D/4 <= 2560
  constexpr int RWv = RowsPerPass<D>::value;
  constexpr int KSv = ATB_WARPS / (ATT_ROWS / RWv) >= 1 ? ATB_WARPS / (ATT_ROWS / RWv) : 1;
  const int nk_pad = nk + ((33 - (nk & 31)) & 31);           // row stride = 1 mod 32: the RW rows of a pass hit distinct banks
  const size_t smem_b = sizeof(float) * (2 * (size_t)nk * KPad<D>::stride + (size_t)KSv * ATT_ROWS * (D + 2) +
                                         (size_t)ATT_ROWS * D + (size_t)ATT_ROWS * nk_pad);
  const bool big_pos = big_pos_case(pos, B, heads, nq, nk);
  if (big_pos && smem_b <= 226 * 1024 && (long long)nk * (D / 4) <= 2560) {
    const int blocks_xy = cdiv(nq, ATT_ROWS) * heads;
    if constexpr (D == 4) {
      // lanes-as-rows variant (see attn_fwd_rows_kernel); SGTA_ATTN_KEYS=1 in the environment keeps the lanes-as-keys one
      static const bool keys_variant = getenv("SGTA_ATTN_KEYS") != nullptr;
      const size_t smem_r = rows_kernel_smem(nk, D);
      if ((!keys_variant || kvhm) && smem_r <= 226 * 1024) {
        int zc = cdiv(2 * sms, blocks_xy);
        if (zc > B) zc = B;
        if (zc < 1) zc = 1;
        const int bpc = cdiv(B, zc);
        zc = cdiv(B, bpc);
        auto kern = kvhm ? attn_fwd_rows_kernel<D, true> : attn_fwd_rows_kernel<D, false>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
        launch_k(kern, dim3(cdiv(nq, ATT_ROWS), heads, zc), ATB_THREADS, smem_r, st, q, k, v, pos, out, B, heads, nq, nk, nk_pad,
                 bpc, inv_scale);
        return check_launch("attn_fwd_rows_kernel");
      }
    }
    int zchunks = cdiv(2 * sms, blocks_xy);                    // >= 2 CTAs per SM worth of parallelism when B allows
    if (zchunks > B) zchunks = B;
    if (zchunks < 1) zchunks = 1;
    const int b_per_cta = cdiv(B, zchunks);
    zchunks = cdiv(B, b_per_cta);
    cudaFuncSetAttribute(attn_fwd_batched_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
    dim3 grid(cdiv(nq, ATT_ROWS), heads, zchunks);
    launch_k(attn_fwd_batched_kernel<D>, grid, ATB_THREADS, smem_b, st, q, k, v, pos, out, B, heads, nq, nk, nk_pad,
             b_per_cta, inv_scale);
    return check_launch("attn_fwd_batched_kernel");
  }
  size_t smem = sizeof(float) * 2 * (size_t)nk * KPad<D>::stride;
  SGTA_REQUIRE(smem <= 220 * 1024, "sgta_attn_forward: nk=%d too large for shared memory", nk);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(cdiv(nq, ATT_ROWS), heads, B);
  launch_k(attn_fwd_kernel<D>, grid, ATT_WARPS * 32, smem, st, q, k, v, pos, out, heads, nq, nk, inv_scale);
  return check_launch("attn_fwd_kernel");
}

template <int D>
static int launch_bwd(const float* q, const float* k, const float* v, const float* pos,
                      const float* go, float* gq, float* gk, float* gv, float* gpos, int B,
                      int heads, int nq, int nk, float inv_scale, cudaStream_t st) {
  size_t smem = sizeof(float) * 2 * (size_t)nk * KPad<D>::stride;
  SGTA_REQUIRE(smem <= 220 * 1024, "sgta_attn_backward: nk=%d too large for shared memory", nk);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(attn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  size_t kvbytes = sizeof(float) * (size_t)B * nk * heads * D;
  cudaMemsetAsync(gk, 0, kvbytes, st);
  cudaMemsetAsync(gv, 0, kvbytes, st);
  dim3 grid(cdiv(nq, ATT_ROWS), heads, B);
  attn_bwd_kernel<D><<<grid, ATT_WARPS * 32, smem, st>>>(q, k, v, pos, go, gq, gk, gv, gpos, heads,
                                                        nq, nk, inv_scale);
  return check_launch("attn_bwd_kernel");
}

}  // namespace sgta

using namespace sgta;

extern "C" int sgta_attn_kvhm_supported(int B, int heads, int nq, int nk, int d, int has_pos) {
  return d == 4 && has_pos && B >= 4 && (size_t)heads * nq * nk * sizeof(float) > (16u << 20) && nk <= 2560 &&
         rows_kernel_smem(nk, 4) <= 226 * 1024;
}

extern "C" int sgta_attn_forward_kvhm(const void* q, const void* k_hm, const void* v_hm, const void* pos, void* out,
                                      int B, int heads, int nq, int nk, int d, float inv_scale, void* stream) {
  SGTA_REQUIRE(q && k_hm && v_hm && pos && out, "sgta_attn_forward_kvhm: null pointer");
  SGTA_REQUIRE(sgta_attn_kvhm_supported(B, heads, nq, nk, d, 1),
               "sgta_attn_forward_kvhm: shape not served by the head-major kernel (see sgta_attn_kvhm_supported)");
  return launch_fwd<4>((const float*)q, (const float*)k_hm, (const float*)v_hm, (const float*)pos, (float*)out, B, heads,
                       nq, nk, inv_scale, (cudaStream_t)stream, true);
}

extern "C" int sgta_attn_forward(const void* q, const void* k, const void* v, const void* pos,
                                 void* out, int B, int heads, int nq, int nk, int d,
                                 float inv_scale, void* stream) {
  SGTA_REQUIRE(q && k && v && out, "sgta_attn_forward: null pointer");
  SGTA_REQUIRE(B > 0 && B <= 65535 && heads > 0 && heads <= 65535 && nq > 0 && nk > 0,
               "sgta_attn_forward: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const float *Q = (const float*)q, *K = (const float*)k, *V = (const float*)v, *P = (const float*)pos;
  switch (d) {
    case 4: return launch_fwd<4>(Q, K, V, P, (float*)out, B, heads, nq, nk, inv_scale, st);
    case 8: return launch_fwd<8>(Q, K, V, P, (float*)out, B, heads, nq, nk, inv_scale, st);
    case 16: return launch_fwd<16>(Q, K, V, P, (float*)out, B, heads, nq, nk, inv_scale, st);
    case 32: return launch_fwd<32>(Q, K, V, P, (float*)out, B, heads, nq, nk, inv_scale, st);
    default:
      set_error("sgta_attn_forward: head dim %d not in {4,8,16,32}", d);
      return SGTA_EUNSUPPORTED;
  }
}

extern "C" int sgta_attn_backward(const void* q, const void* k, const void* v, const void* pos,
                                  const void* grad_out, void* grad_q, void* grad_k, void* grad_v,
                                  void* grad_pos, int B, int heads, int nq, int nk, int d,
                                  float inv_scale, void* stream) {
  SGTA_REQUIRE(q && k && v && grad_out && grad_q && grad_k && grad_v, "sgta_attn_backward: null pointer");
  SGTA_REQUIRE(B > 0 && B <= 65535 && heads > 0 && nq > 0 && nk > 0, "sgta_attn_backward: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const float *Q = (const float*)q, *K = (const float*)k, *V = (const float*)v, *P = (const float*)pos;
  const float* G = (const float*)grad_out;
  float *GQ = (float*)grad_q, *GK = (float*)grad_k, *GV = (float*)grad_v, *GP = (float*)grad_pos;
  switch (d) {
    case 4: return launch_bwd<4>(Q, K, V, P, G, GQ, GK, GV, GP, B, heads, nq, nk, inv_scale, st);
    case 8: return launch_bwd<8>(Q, K, V, P, G, GQ, GK, GV, GP, B, heads, nq, nk, inv_scale, st);
    case 16: return launch_bwd<16>(Q, K, V, P, G, GQ, GK, GV, GP, B, heads, nq, nk, inv_scale, st);
    case 32: return launch_bwd<32>(Q, K, V, P, G, GQ, GK, GV, GP, B, heads, nq, nk, inv_scale, st);
    default:
      set_error("sgta_attn_backward: head dim %d not in {4,8,16,32}", d);
      return SGTA_EUNSUPPORTED;
  }
}
