// Heatmap decode kernels (memory-bound: one read of the heatmap per launch).
//
//   decode_peaks_kernel   -- the LIVE decode of the reference: dream_generic_decode ->
//       _peaks_info -> peaks_from_belief_maps (sgtapose/lib/model/decode.py:184-313,
//       lib/model/utils.py:207-284, sgtapose/image_proc.py:1032-1143).  The reference does
//       this on the CPU with scipy + Python loops (44 ms/frame, >= 7 device syncs); here one
//       CTA owns one (sample, keypoint) map in shared memory and reproduces scipy's
//       gaussian_filter(sigma=3) and numpy.average bit for bit (float64, same operation
//       order, no FMA contraction), so integer peak indices are bit-exact.
//       decode_peaks_f32_kernel is the production path: the blur runs in float32 (25-tap FMA
//       chains from registers), every peak predicate whose float32 margin exceeds a rigorous
//       rounding bound is decided there, and only the undecided pixels are re-evaluated with
//       the float64 restatement -- same integer outputs, ~10x fewer cycles per map.
//       decode_peaks_kernel (all-float64) stays as the cross-check and the large-map path.
//   nms_topk kernels      -- the alternate decode: _nms + _topk (utils.py:59-103).
//   soft_argmax_kernel    -- SoftArgmaxPavlo (sgtapose/spatial_softmax.py:24-95).
#include <cfloat>

#include "common.cuh"

namespace sgta {

struct GaussW { double w[25]; };
constexpr int GR = 12;                    // int(4.0 * 3 + 0.5)
constexpr double PEAK_OFFSET = 0.4395;    // utils.py:212
constexpr float BLUR_THRESH = 0.01f;      // image_proc.py:1043
constexpr float AMBIG_GAP = 0.25f;        // utils.py:230-233

__device__ __forceinline__ int reflect_idx(int i, int n) {
  // scipy 'reflect' (half-sample symmetric), any distance outside
  if (n == 1) return 0;
  int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

struct Cand {
  double cx, cy;
  float score;
  int pos;   // row-major pixel position (np.nonzero order); INT_MAX = empty
};

__device__ __forceinline__ bool cand_before(const Cand& a, const Cand& b) {
  // order of sorted(peaks, key=cy, reverse=True) with Python's stable sort
  if (a.pos == 0x7fffffff) return false;
  if (b.pos == 0x7fffffff) return true;
  return a.cy > b.cy || (a.cy == b.cy && a.pos < b.pos);
}
__device__ __forceinline__ void cand_insert(const Cand& c, Cand& first, Cand& second) {
  if (cand_before(c, first)) { second = first; first = c; }
  else if (cand_before(c, second)) { second = c; }
}

__device__ __forceinline__ double np_sum25(const double* a) {
  // numpy pairwise_sum for n = 25 (< 128): 8 interleaved accumulators over the first 24
  // elements, tree-combined, then the tail
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
#pragma unroll
  for (int i = 8; i < 24; i += 8)
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  return __dadd_rn(res, a[24]);
}

// numpy.average of the 5x5 window of the un-blurred map around an accepted peak (image_proc.py:1075-1110)
template <class F>
__device__ __forceinline__ Cand make_cand(F orig, int y, int x, int h, int w) {
  double wts[25], xv[25], yv[25];
#pragma unroll
  for (int j = -2; j <= 2; ++j) {        // x offset -> first array index
#pragma unroll
    for (int i = -2; i <= 2; ++i) {      // y offset -> second array index
      int f = (j + 2) * 5 + (i + 2);
      bool in = (y + i >= 0) && (y + i < h) && (x + j >= 0) && (x + j < w);
      double wt = in ? (double)orig((y + i) * w + (x + j)) : 0.0;
      wts[f] = wt;
      xv[f] = in ? __dmul_rn((double)(x + j), wt) : 0.0;
      yv[f] = in ? __dmul_rn((double)(y + i), wt) : 0.0;
    }
  }
  double scl = np_sum25(wts);
  Cand c;
  if (scl == 0.0) {
    c.cx = __dadd_rn((double)x, PEAK_OFFSET);
    c.cy = __dadd_rn((double)y, PEAK_OFFSET);
  } else {
    c.cx = __dadd_rn(__ddiv_rn(np_sum25(xv), scl), PEAK_OFFSET);
    c.cy = __dadd_rn(__ddiv_rn(np_sum25(yv), scl), PEAK_OFFSET);
  }
  c.score = orig(y * w + x);
  c.pos = y * w + x;
  return c;
}

// block-wide top-2 of the candidates in sorted(key=cy, reverse=True) order, then the reference's
// selection rule (utils.py:222-240) and the reg / tracking gathers (decode.py:218-276)
template <class F>
__device__ __forceinline__ void reduce_and_emit(Cand first, Cand second, int mine, Cand* cands, int* s_count,
                                                F orig, const float* __restrict__ reg,
                                                const float* __restrict__ tracking, float* __restrict__ scores,
                                                long long* __restrict__ inds, long long* __restrict__ xs,
                                                long long* __restrict__ ys, float* __restrict__ cts_wreg,
                                                float* __restrict__ trk, int bc, int b, int h, int w) {
  const int tid = threadIdx.x, hw = h * w;
  if (mine) atomicAdd(s_count, mine);
  cands[2 * tid] = first;
  cands[2 * tid + 1] = second;
  __syncthreads();
  for (int stride = 128; stride > 0; stride >>= 1) {
    if (tid < stride) {
      Cand a = cands[2 * tid], c2 = cands[2 * tid + 1];
      cand_insert(cands[2 * (tid + stride)], a, c2);
      cand_insert(cands[2 * (tid + stride) + 1], a, c2);
      cands[2 * tid] = a; cands[2 * tid + 1] = c2;
    }
    __syncthreads();
  }
  if (tid == 0) {
    int count = *s_count;
    Cand a = cands[0], s2 = cands[1];
    bool found = false;
    if (count == 1) found = true;
    else if (count > 1) found = (a.score - s2.score) >= AMBIG_GAP;   // float32 subtraction
    long long xi = 0, yi = 0;
    if (found) {
      xi = (long long)a.cx;   // int(): truncation toward zero
      yi = (long long)a.cy;
      // non-negative (post-sigmoid) maps keep the centroid inside the map; clamp only guards
      // the gathers against negative-weight inputs the reference would fault on
      xi = xi < 0 ? 0 : (xi > w - 1 ? w - 1 : xi);
      yi = yi < 0 ? 0 : (yi > h - 1 ? h - 1 : yi);
    }
    const long long ind = yi * w + xi;
    // the five gathers depend on `ind` only: issue them together (one L2 round trip, not three)
    const float* rb = reg ? reg + (long long)b * 2 * hw : nullptr;
    const float* tb = (tracking && trk) ? tracking + (long long)b * 2 * hw : nullptr;
    const float sv = orig((int)ind);
    const float r0 = rb ? __ldg(rb + ind) : 0.5f, r1 = rb ? __ldg(rb + hw + ind) : 0.5f;
    const float t0 = tb ? __ldg(tb + ind) : 0.f, t1 = tb ? __ldg(tb + hw + ind) : 0.f;
    scores[bc] = found ? sv : -1.f; inds[bc] = ind; xs[bc] = xi; ys[bc] = yi;
    cts_wreg[2 * bc] = (float)xi + r0; cts_wreg[2 * bc + 1] = (float)yi + r1;
    if (tb) { trk[2 * bc] = t0; trk[2 * bc + 1] = t1; }
  }
}

// grid: B*C CTAs, 256 threads; smem: 3 maps of h*w floats + 256 Cand pairs.
// FAST (maps up to ~100x100): the two blur passes read float64 copies of their input that are PADDED
// with scipy's 'reflect' halo along the pass axis ((h+24) x w and h x (w+24) doubles), so the 25-tap
// inner loop is 2 LDS.64 + DADD + DMUL + DADD per tap pair -- no index reflection (two integer modulos),
// no float->double conversion per tap.  Same operations in the same order: results are bit-identical.
// The kernel is bound by the fp64 pipe (37 fp64 ops per pixel and pass, no FMA contraction allowed).
template <bool FAST>
__global__ void __launch_bounds__(256)
decode_peaks_kernel(const float* __restrict__ hm, const float* __restrict__ reg,
                    const float* __restrict__ tracking, float* __restrict__ scores,
                    long long* __restrict__ inds, long long* __restrict__ xs,
                    long long* __restrict__ ys, float* __restrict__ cts_wreg,
                    float* __restrict__ trk, GaussW gw, int C, int h, int w) {
  extern __shared__ __align__(16) unsigned char dsm[];
  const int hw = h * w;
  __shared__ int s_count;
  const int tid = threadIdx.x;
  const int bc = blockIdx.x, b = bc / C;
  const float* src = hm + (long long)bc * hw;
  if (tid == 0) s_count = 0;
  float *ori, *blr;
  Cand* cands;
  if (FAST) {
    double* pv = reinterpret_cast<double*>(dsm);                 // [(h + 2 GR)][w]
    double* ph = pv + (size_t)(h + 2 * GR) * w;                  // [h][w + 2 GR]
    blr = reinterpret_cast<float*>(ph + (size_t)h * (w + 2 * GR));
    ori = nullptr;                                               // un-blurred values are re-read from L2
    cands = reinterpret_cast<Cand*>(dsm);                        // aliases pv (dead after pass 1)
    const int wp = w + 2 * GR;
    for (int e = tid; e < (h + 2 * GR) * w; e += 256) {
      const int yy = e / w, x = e - yy * w;
      pv[e] = (double)__ldg(src + reflect_idx(yy - GR, h) * w + x);
    }
    __syncthreads();
    // pass 1: correlate along axis 0 (rows), float64 accumulate, float32 store
    // four outputs per thread and trip: four independent DADD chains hide the fp64 latency
    for (int e0 = tid; e0 < hw; e0 += 4 * 256) {
      const double* c0[4];
      double t[4];
      int yy[4], xx[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = min(e0 + u * 256, hw - 1);
        yy[u] = e / w; xx[u] = e - yy[u] * w;
        c0[u] = pv + (size_t)(yy[u] + GR) * w + xx[u];
        t[u] = __dmul_rn(c0[u][0], gw.w[GR]);
      }
#pragma unroll
      for (int ii = -GR; ii < 0; ++ii)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          t[u] = __dadd_rn(t[u], __dmul_rn(__dadd_rn(c0[u][ii * w], c0[u][-ii * w]), gw.w[ii + GR]));
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (e0 + u * 256 < hw) ph[(size_t)yy[u] * wp + GR + xx[u]] = (double)(float)t[u];
    }
    __syncthreads();
    for (int e = tid; e < h * 2 * GR; e += 256) {                // 'reflect' halo of pass 2
      const int y = e / (2 * GR), k = e - y * 2 * GR;
      const int xp = k < GR ? k - GR : w + (k - GR);
      ph[(size_t)y * wp + GR + xp] = ph[(size_t)y * wp + GR + reflect_idx(xp, w)];
    }
    __syncthreads();
    // pass 2: along axis 1 (columns)
    for (int e0 = tid; e0 < hw; e0 += 4 * 256) {
      const double* c0[4];
      double t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = min(e0 + u * 256, hw - 1);
        const int y = e / w, x = e - y * w;
        c0[u] = ph + (size_t)y * wp + GR + x;
        t[u] = __dmul_rn(c0[u][0], gw.w[GR]);
      }
#pragma unroll
      for (int ii = -GR; ii < 0; ++ii)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          t[u] = __dadd_rn(t[u], __dmul_rn(__dadd_rn(c0[u][ii], c0[u][-ii]), gw.w[ii + GR]));
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (e0 + u * 256 < hw) blr[e0 + u * 256] = (float)t[u];
    }
    __syncthreads();
  } else {
    ori = reinterpret_cast<float*>(dsm);
    float* tmp = ori + hw;
    blr = tmp + hw;
    cands = reinterpret_cast<Cand*>(dsm + ((sizeof(float) * 3 * hw + 15) / 16) * 16);
    for (int e = tid; e < hw; e += 256) ori[e] = __ldg(src + e);
    __syncthreads();
    // pass 1: correlate along axis 0 (rows), float64 accumulate, float32 store
    for (int e = tid; e < hw; e += 256) {
      int y = e / w, x = e % w;
      double t = __dmul_rn((double)ori[e], gw.w[GR]);
#pragma unroll 4
      for (int ii = -GR; ii < 0; ++ii) {
        double a = (double)ori[reflect_idx(y + ii, h) * w + x];
        double c = (double)ori[reflect_idx(y - ii, h) * w + x];
        t = __dadd_rn(t, __dmul_rn(__dadd_rn(a, c), gw.w[ii + GR]));
      }
      tmp[e] = (float)t;
    }
    __syncthreads();
    // pass 2: along axis 1 (columns)
    for (int e = tid; e < hw; e += 256) {
      int y = e / w, x = e % w;
      const float* row = tmp + y * w;
      double t = __dmul_rn((double)row[x], gw.w[GR]);
#pragma unroll 4
      for (int ii = -GR; ii < 0; ++ii) {
        double a = (double)row[reflect_idx(x + ii, w)];
        double c = (double)row[reflect_idx(x - ii, w)];
        t = __dadd_rn(t, __dmul_rn(__dadd_rn(a, c), gw.w[ii + GR]));
      }
      blr[e] = (float)t;
    }
    __syncthreads();
  }
  auto orig = [&](int idx) -> float { return FAST ? __ldg(src + idx) : ori[idx]; };

  // peak test on the blurred map (0 outside), centroid on the un-blurred one
  Cand first, second;
  first.pos = second.pos = 0x7fffffff;
  first.cx = first.cy = second.cx = second.cy = 0.0;
  first.score = second.score = 0.f;
  int mine = 0;
  for (int e = tid; e < hw; e += 256) {
    int y = e / w, x = e % w;
    float v = blr[e];
    float up = y > 0 ? blr[e - w] : 0.f, dn = y < h - 1 ? blr[e + w] : 0.f;
    float lf = x > 0 ? blr[e - 1] : 0.f, rt = x < w - 1 ? blr[e + 1] : 0.f;
    if (!(v >= up && v >= dn && v >= lf && v >= rt && v > BLUR_THRESH)) continue;
    ++mine;
    Cand c = make_cand(orig, y, x, h, w);
    cand_insert(c, first, second);
  }
  reduce_and_emit(first, second, mine, cands, &s_count, orig, reg, tracking, scores, inds, xs, ys, cts_wreg,
                  trk, bc, b, h, w);
}

// ---------------------------------------------------------------------------------------
// Production live decode: float32 blur + rigorous margin test + float64 re-check of the rest.
//
// The reference's peak predicate (image_proc.py:1047-1073) only COMPARES blurred values; the
// outputs (centroid, score, indices) come from the un-blurred map.  So the blur is evaluated in
// float32 and a comparison is accepted only when its margin exceeds the worst-case distance
// between this evaluation and the reference's (float64 accumulate, float32 store per pass):
//   per pass   |fma-chain - real| <= 26u * sum w|x|  (25 fused steps + tap rounding), reference store u|t|
//   two passes |v32 - v_ref| <= (26 + 27 + 1) u S = 54 u S,   S = blur of |x|,  u = 2^-24
// For a non-negative map (every post-sigmoid heatmap) S is the blurred value itself, so the bound is
// RELATIVE: E(p) = 80 u v32(p); otherwise S <= M = max|x| and E = 80 u M.  A comparison v ? n is accepted
// when |v - n| > E(v) + E(n); anything inside the band (or NaN/Inf) is "undecided" and is re-evaluated with
// the exact restatement (exact_pass1/exact_pass2: same operations and order as decode_peaks_kernel) --
// pixel by pixel by the whole CTA when there are <= UND_CAP of them, else by re-blurring the whole map in
// float64 (flat / constant maps).  Integer outputs are therefore identical by construction, not by luck.
//
// grid: B*C CTAs of 256 threads, 3 resident per SM; smem: two h x (w|1) float planes.
//   phase 0  map -> bufA (dense rows), nine 16-byte loads in flight per thread; block max|x| and sign
//   phase 1  vertical taps   -> bufB.  PACKED (even sizes, the production shape): lane = two adjacent columns,
//            item = 6 output rows, 30 LDS.64 -> 150 FFMA2 (taps broadcast from uniform registers);
//            scalar fallback: item = (column, 12 output rows), 36 LDS -> 300 FFMA
//   phase 2  horizontal taps -> bufA, same shapes transposed (PACKED: lane = two adjacent rows; pass 1 stores
//            row pairs interleaved so this is again one LDS.64 per input; odd strides keep it conflict-free)
//            each item also reports whether its largest output may exceed the 0.01 threshold ("hot")
//   phase 3  one thread per pixel of the hot items (a few dozen of 768): margin test against the four
//            neighbours, lists of accepted / undecided pixels; centroids spread over the CTA, block top-2,
//            emit.  (NaN/Inf inputs, more than 256 hot items and the float64 round scan every pixel instead.)
// Measured and rejected: a per-pixel candidate bit mask set by phase 2 (+2.6 k instructions there, scan no
// shorter: 0.36 ms vs 0.325 ms per 1024 frames); persistent CTAs with a third plane that prefetches the next
// map by cp.async (2 CTAs/SM instead of 3: 0.37 ms vs 0.30 ms); 384-thread CTAs (12 warps x 3 CTAs, 56
// registers, no spills in the blur passes: 0.33 ms).
// ---------------------------------------------------------------------------------------
struct GaussWF { float w[25]; int full_map; };     // full_map != 0: blur every pixel (test hook: sgta_debug_flags bit 12)
constexpr int SEG = 12;
constexpr float ERR_BOUND = 80.f * 5.9604644775390625e-08f;   // 80 * 2^-24, see the derivation above

__device__ unsigned long long g_exact_rechecks = 0ull;

__device__ __noinline__ int reflect_idx_far(int i, int n) { return reflect_idx(i, n); }
__device__ __forceinline__ int refl(int i, int n) {
  int r = i < 0 ? -1 - i : (i >= n ? 2 * n - 1 - i : i);
  if ((unsigned)r >= (unsigned)n) r = reflect_idx_far(i, n);    // more than one fold (ragged last segment, tiny maps)
  return r;
}

// one fold, branch-free: valid for -n <= i < 2n (maps of at least SEG + 2*GR rows / columns)
__device__ __forceinline__ int refl1(int i, int n) {
  i = i < 0 ? -1 - i : i;
  return i >= n ? 2 * n - 1 - i : i;
}

// exact restatement of one output of pass 1 (reads the un-blurred map from global memory) and of
// pass 2 (reads a row of exact pass-1 values): same operations and order as decode_peaks_kernel
__device__ __noinline__ float exact_pass1(const float* __restrict__ src, int y, int x, int h, int w,
                                          const GaussW& gw) {
  double t = __dmul_rn((double)__ldg(src + y * w + x), gw.w[GR]);
  for (int ii = -GR; ii < 0; ++ii) {
    double a = (double)__ldg(src + reflect_idx(y + ii, h) * w + x);
    double c = (double)__ldg(src + reflect_idx(y - ii, h) * w + x);
    t = __dadd_rn(t, __dmul_rn(__dadd_rn(a, c), gw.w[ii + GR]));
  }
  return (float)t;
}
__device__ __noinline__ float exact_pass2(const float* row, int x, int w, const GaussW& gw) {
  double t = __dmul_rn((double)row[x], gw.w[GR]);
  for (int ii = -GR; ii < 0; ++ii) {
    double a = (double)row[reflect_idx(x + ii, w)];
    double c = (double)row[reflect_idx(x - ii, w)];
    t = __dadd_rn(t, __dmul_rn(__dadd_rn(a, c), gw.w[ii + GR]));
  }
  return (float)t;
}

// ---- packed float32 pairs (FFMA2, sm_100): two maps columns (pass 1) or two map rows (pass 2) per lane-op.
// Each half is an IEEE fused multiply-add, so the results equal the scalar chains bit for bit.
constexpr int SEG2 = 6;                     // outputs per item and half

// One item of a packed pass: SEG2 outputs x 2 halves from SEG2 + 24 float2 inputs at `base + idx(i)`.
// Input-major order keeps one input pair live; per output the taps still accumulate in order t = 0..24.
template <bool INTERIOR>
__device__ __forceinline__ void packed_taps(const float2* __restrict__ base, int stride, int first, int n,
                                            const float2 (&w2)[2 * GR + 1], float2 (&a2)[SEG2]) {
#pragma unroll
  for (int k = 0; k < SEG2; ++k) a2[k] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < SEG2 + 2 * GR; ++i) {
    const int pos = INTERIOR ? first + i : refl1(first + i, n);
    const float2 v = base[pos * stride];
#pragma unroll
    for (int k = 0; k < SEG2; ++k)
      if (i - k >= 0 && i - k <= 2 * GR) a2[k] = __ffma2_rn(v, w2[i - k], a2[k]);
  }
}

constexpr int UND_CAP = 64;
constexpr int HOT_CAP = 256;    // more hot items than this (or NaN/Inf input): the scan visits every pixel     // more undecided pixels than this: the whole map is re-blurred in float64

template <bool PACKED>
__global__ void __launch_bounds__(256, 3)
decode_peaks_f32_kernel(const float* __restrict__ hm, const float* __restrict__ reg,
                        const float* __restrict__ tracking, float* __restrict__ scores,
                        long long* __restrict__ inds, long long* __restrict__ xs,
                        long long* __restrict__ ys, float* __restrict__ cts_wreg,
                        float* __restrict__ trk, const __grid_constant__ GaussW gw,
                        const __grid_constant__ GaussWF gf, int C, int h, int w, int plane_floats) {
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ int s_count, s_nund, s_nacc, s_min_ok, s_nhot, s_finite;
  __shared__ int s_box[4];                              // active rows [0..1] / columns [2..3] (see "active box" below)
  __shared__ int s_hot[HOT_CAP];                        // pass-2 items (12 pixels each) that may hold a pixel above the threshold
  __shared__ float s_max[8];
  __shared__ int s_list[UND_CAP];                       // undecided pixels (row-major position)
  __shared__ float s_t[3 * 32];                         // re-check: exact pass-1 values, 3 rows x 27
  __shared__ float s_v[8];                              // re-check: the five exact blurred values
  const int hw = h * w, wp = w | 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bc = blockIdx.x, b = bc / C;
  const float* src = hm + (long long)bc * hw;
  float* bufA = reinterpret_cast<float*>(dsm);
  float* bufB = bufA + plane_floats;
  Cand* cands = reinterpret_cast<Cand*>(bufB);          // aliases bufB (dead once the scans are done)
  pdl_trigger();
  if (tid == 0) {
    s_count = 0; s_nund = 0; s_nacc = 0; s_min_ok = 1; s_nhot = 0; s_finite = 1;
    s_box[0] = h; s_box[1] = -1; s_box[2] = w; s_box[3] = -1;
  }
  pdl_wait();
  __syncthreads();

  // phase 0
  float m = 0.f;
  bool pos = true;                                              // false on any negative or NaN element
  bool fin = true;                                              // false on any NaN / Inf element
  if ((hw & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    // nine 16-byte loads in flight per thread (a whole 96x96 map per CTA) before the first use
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(bufA);
    const int n4 = hw >> 2;
    for (int base = tid; base < n4; base += 9 * 256) {
      float4 v[9];
#pragma unroll
      for (int j = 0; j < 9; ++j)
        v[j] = base + j * 256 < n4 ? __ldg(s4 + base + j * 256) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        if (base + j * 256 < n4) d4[base + j * 256] = v[j];
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v[j].x), fabsf(v[j].y))), fmaxf(fabsf(v[j].z), fabsf(v[j].w)));
        pos = pos && (v[j].x >= 0.f) && (v[j].y >= 0.f) && (v[j].z >= 0.f) && (v[j].w >= 0.f);
        fin = fin && (fabsf(v[j].x) <= FLT_MAX) && (fabsf(v[j].y) <= FLT_MAX) && (fabsf(v[j].z) <= FLT_MAX) &&
              (fabsf(v[j].w) <= FLT_MAX);
      }
    }
  } else {
    for (int e = tid; e < hw; e += 256) {
      float v = __ldg(src + e);
      bufA[e] = v;
      m = fmaxf(m, fabsf(v));
      pos = pos && (v >= 0.f);
      fin = fin && (fabsf(v) <= FLT_MAX);
    }
  }
  m = warp_max(m);
  if (lane == 0) s_max[warp] = m;
  for (int x = tid; x < w; x += 256) bufB[h + x] = 0.f;         // column maxima of the active-box test (bufB is still free)
  if (!pos) s_min_ok = 0;                                       // benign race: every writer stores 0
  if (!fin) s_finite = 0;
  __syncthreads();
  m = s_max[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, s_max[i]);
  const float EM = ERR_BOUND * m;                               // sign-agnostic bound
  const bool nonneg = s_min_ok;                                 // x >= 0 everywhere: sum w|x| is the blur itself
  const float er0 = nonneg ? ERR_BOUND : 0.f, ea0 = nonneg ? 0.f : EM;   // |v32 - v_ref| <= er0 * v32 + ea0
  // A pass-2 item is "hot" unless its largest output is surely below the threshold (the test is monotone in
  // v, so the maximum decides for all of them).  Hot items are the only ones the scan has to visit.
  auto note_hot = [&](float mx, int e0) {
    if (!(mx < BLUR_THRESH - fmaf(er0, mx, ea0))) {
      const int slot = atomicAdd(&s_nhot, 1);
      if (slot < HOT_CAP) s_hot[slot] = e0;
    }
  };

  // Active box.  Only pixels whose blurred value may reach the 0.01 threshold can be peaks, and for a non-negative
  // map the blur (weights >= 0, sum 1 in either pass) is bounded by the blurred row / column maxima:
  //     blur(y, x) <= sum_i w_i * rowmax[refl(y + i)] =: R(y),   blur(y, x) <= sum_j w_j * colmax[refl(x + j)] =: C(x).
  // Rows with R(y) < TAU and columns with C(x) < TAU (TAU = 0.0099: 1 % below the threshold, the float32 evaluation of
  // R / C is good to 1e-6) hold no peak and are never "undecided", so the two passes only run on the bounding box of
  // the rest, grown by one pixel (the neighbours a comparison reads; they are below the threshold themselves, which
  // is all the predicate needs to know about them).  A real heat map is a few blobs on a flat floor: the box is a
  // few per cent of the map and the kernel turns from FP32-issue-bound into a read of the map.
  int ya0 = 0, ya1 = h - 1, xa0 = 0, xa1 = w - 1;               // box in which bufA will hold blurred values
  if (PACKED && nonneg && s_finite && !gf.full_map) {
    // (non-negative finite floats order like their bit patterns: REDUX / integer atomicMax do the reductions)
    float* rmax = bufB;                                         // bufB is free until pass 1
    unsigned* cmax_u = reinterpret_cast<unsigned*>(bufB + h);   // zeroed in phase 0
    for (int y = warp; y < h; y += 8) {
      float v = 0.f;
      for (int x = lane; x < w; x += 32) v = fmaxf(v, bufA[y * w + x]);
      const unsigned mu = __reduce_max_sync(0xffffffffu, __float_as_uint(v));
      if (lane == 0) rmax[y] = __uint_as_float(mu);
    }
    {
      const int G = 256 / w > 0 ? 256 / w : 1;                  // row groups: (group, column) per thread
      for (int e = tid; e < G * w; e += 256) {
        const int g = e / w, x = e - g * w;
        const int y0 = g * h / G, y1 = (g + 1) * h / G;
        float v = 0.f;
#pragma unroll 8
        for (int y = y0; y < y1; ++y) v = fmaxf(v, bufA[y * w + x]);
        atomicMax(cmax_u + x, __float_as_uint(v));
      }
    }
    const float* cmax = bufB + h;
    __syncthreads();
    constexpr float TAU = 0.0099f;
    for (int e = tid; e < h + w; e += 256) {
      const bool isrow = e < h;
      const int i0 = isrow ? e : e - h, n = isrow ? h : w;
      const float* mx = isrow ? rmax : cmax;
      float r = 0.f;
#pragma unroll
      for (int t = 0; t < 2 * GR + 1; ++t) r = fmaf(gf.w[t], mx[refl(i0 - GR + t, n)], r);
      if (r >= TAU) {
        atomicMin(&s_box[isrow ? 0 : 2], i0);
        atomicMax(&s_box[isrow ? 1 : 3], i0);
      }
    }
    __syncthreads();
    ya0 = max(s_box[0] - 1, 0); ya1 = min(s_box[1] + 1, h - 1);
    xa0 = max(s_box[2] - 1, 0); xa1 = min(s_box[3] + 1, w - 1);
    if (s_box[1] < 0 || s_box[3] < 0) { ya0 = 0; ya1 = -1; xa0 = 0; xa1 = -1; }    // nothing can reach the threshold
  }

  if (PACKED) {
    // h, w even and >= SEG2 + 2*GR (launcher).  Pass 1: lane = two adjacent columns (one LDS.64 of the dense
    // map per input row); its output goes to bufB interleaved by ROW PAIR, [h/2][w|1] float2 = (row 2j, row 2j+1),
    // so that pass 2 (lane = two adjacent rows) also reads one LDS.64 per input column.
    float2 w2[2 * GR + 1];
#pragma unroll
    for (int t = 0; t < 2 * GR + 1; ++t) w2[t] = make_float2(gf.w[t], gf.w[t]);
    float2* tmpI = reinterpret_cast<float2*>(bufB);
    const int wh = w >> 1, hh = h >> 1;
    // items of the active box only: row segments that hold a box row x column pairs within GR of a box column
    // (pass-2 items are whole column segments: pass 1 covers every column they read, so that every value pass 2
    // writes -- also the ones outside the box -- is a true blurred value)
    const bool any = ya1 >= ya0;
    const int sgy0 = ya0 / SEG2, nsy = any ? ya1 / SEG2 - sgy0 + 1 : 0;
    const int sgx0 = xa0 / SEG2, nsx = any ? xa1 / SEG2 - sgx0 + 1 : 0;
    const int xp0 = max(sgx0 * SEG2 - GR, 0) >> 1;
    const int nxp = any ? (min((sgx0 + nsx) * SEG2 - 1 + GR, w - 1) >> 1) - xp0 + 1 : 0;
    for (int it = tid; it < nsy * nxp; it += 256) {
      const int sg = sgy0 + it / nxp, x = 2 * (xp0 + it % nxp), y0 = sg * SEG2;
      const float2* col = reinterpret_cast<const float2*>(bufA + x);    // row stride w floats = w/2 float2
      float2 a2[SEG2];
      if (y0 >= GR && y0 + SEG2 + GR <= h) packed_taps<true>(col, wh, y0 - GR, h, w2, a2);
      else packed_taps<false>(col, wh, y0 - GR, h, w2, a2);
#pragma unroll
      for (int k = 0; k < SEG2; k += 2) {
        if (y0 + k < h) {
          float2* o = tmpI + ((y0 + k) >> 1) * wp + x;
          o[0] = make_float2(a2[k].x, a2[k + 1].x);
          o[1] = make_float2(a2[k].y, a2[k + 1].y);
        }
      }
    }
    __syncthreads();
    const int yp0 = ya0 >> 1, nyp = any ? (ya1 >> 1) - yp0 + 1 : 0;
    for (int it = tid; it < nsx * nyp; it += 256) {
      const int sg = sgx0 + it / nyp, yp = yp0 + it % nyp, x0 = sg * SEG2;
      const float2* row = tmpI + yp * wp;
      float2 a2[SEG2];
      if (x0 >= GR && x0 + SEG2 + GR <= w) packed_taps<true>(row, 1, x0 - GR, w, w2, a2);
      else packed_taps<false>(row, 1, x0 - GR, w, w2, a2);
      float* o0 = bufA + (2 * yp) * wp + x0;
      float mx = fmaxf(a2[0].x, a2[0].y);                       // column x0 always exists
#pragma unroll
      for (int k = 0; k < SEG2; ++k)
        if (x0 + k < w) { o0[k] = a2[k].x; o0[wp + k] = a2[k].y; mx = fmaxf(mx, fmaxf(a2[k].x, a2[k].y)); }
      note_hot(mx, 2 * yp * w + x0);                            // item = rows 2yp, 2yp+1 x columns x0 .. x0+5
    }
    __syncthreads();
  } else {
    // phase 1: vertical taps
    const int nsy = (h + SEG - 1) / SEG;
    for (int it = tid; it < nsy * w; it += 256) {
      const int sg = it / w, x = it - sg * w, y0 = sg * SEG;
      float in[SEG + 2 * GR];
      if (y0 >= GR && y0 + SEG + GR <= h) {                       // interior segment: no fold
        const float* col = bufA + (y0 - GR) * w + x;
#pragma unroll
        for (int r = 0; r < SEG + 2 * GR; ++r) in[r] = col[r * w];
      } else if (h >= SEG + 2 * GR) {                             // first / last segments: one fold
#pragma unroll
        for (int r = 0; r < SEG + 2 * GR; ++r) in[r] = bufA[refl1(y0 - GR + r, h) * w + x];
      } else {
#pragma unroll
        for (int r = 0; r < SEG + 2 * GR; ++r) in[r] = bufA[refl(y0 - GR + r, h) * w + x];
      }
      float a[SEG];                                               // SEG independent chains: tap-major order
#pragma unroll
      for (int k = 0; k < SEG; ++k) a[k] = in[k] * gf.w[0];
#pragma unroll
      for (int t = 1; t < 2 * GR + 1; ++t)
#pragma unroll
        for (int k = 0; k < SEG; ++k) a[k] = fmaf(in[k + t], gf.w[t], a[k]);
      float* ocol = bufB + y0 * wp + x;
      if (y0 + SEG <= h) {
#pragma unroll
        for (int k = 0; k < SEG; ++k) ocol[k * wp] = a[k];
      } else {
#pragma unroll
        for (int k = 0; k < SEG; ++k)
          if (y0 + k < h) ocol[k * wp] = a[k];
      }
    }
    __syncthreads();

    // phase 2: horizontal taps
    const int nsx = (w + SEG - 1) / SEG;
    for (int it = tid; it < nsx * h; it += 256) {
      const int sg = it / h, y = it - sg * h, x0 = sg * SEG;
      const float* row = bufB + y * wp;
      float in[SEG + 2 * GR];
      if (x0 >= GR && x0 + SEG + GR <= w) {
#pragma unroll
        for (int r = 0; r < SEG + 2 * GR; ++r) in[r] = row[x0 - GR + r];
      } else if (w >= SEG + 2 * GR) {
#pragma unroll
        for (int r = 0; r < SEG + 2 * GR; ++r) in[r] = row[refl1(x0 - GR + r, w)];
      } else {
#pragma unroll
        for (int r = 0; r < SEG + 2 * GR; ++r) in[r] = row[refl(x0 - GR + r, w)];
      }
      float* orow = bufA + y * wp + x0;
      float a[SEG];
#pragma unroll
      for (int k = 0; k < SEG; ++k) a[k] = in[k] * gf.w[0];
#pragma unroll
      for (int t = 1; t < 2 * GR + 1; ++t)
#pragma unroll
        for (int k = 0; k < SEG; ++k) a[k] = fmaf(in[k + t], gf.w[t], a[k]);
      float mx = a[0];
#pragma unroll
      for (int k = 0; k < SEG; ++k)
        if (x0 + k < w) { orow[k] = a[k]; mx = fmaxf(mx, a[k]); }
      note_hot(mx, y * w + x0);                                 // item = row y x columns x0 .. x0+11
    }
    __syncthreads();
  }

  // phase 3: scan.  Round 0 tests the float32 blur with margins: accepted peaks and undecided pixels go to
  // shared-memory lists (the loop itself stays light); if more than UND_CAP pixels fall inside the rounding
  // band (flat or constant maps) the whole map is re-blurred in float64 and round 1 applies the reference
  // predicate to exact values.
  auto orig = [&](int idx) -> float { return __ldg(src + idx); };
  int* acc_list = reinterpret_cast<int*>(bufB);                 // <= h*w accepted peaks; bufB is dead here
  bool exact_mode = false;
  for (;;) {
    const float er = exact_mode ? 0.f : er0, ea = exact_mode ? 0.f : ea0;   // |v - v_ref| <= er * v + ea
    auto test_pixel = [&](int y, int x) {
      const float* p = bufA + y * wp + x;
      const float v = *p;
      const float ev = fmaf(er, v, ea);
      if (v < BLUR_THRESH - ev) return;                         // surely not above the threshold
      const float nb[4] = {y > 0 ? p[-wp] : 0.f, y < h - 1 ? p[wp] : 0.f, x > 0 ? p[-1] : 0.f,
                           x < w - 1 ? p[1] : 0.f};
      bool sure = v > BLUR_THRESH + ev, drop = false;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float band = ev + fmaf(er, nb[k], ea);
        const float d = v - nb[k];
        drop = drop || d < -band;                               // surely below a neighbour
        sure = sure && d > band;
      }
      if (drop) return;
      if (!sure) {                                              // inside the rounding band, a tie, or NaN/Inf
        if (exact_mode) {                                       // image_proc.py:1054-1073 on exact values
          if (!(v > BLUR_THRESH && v >= nb[0] && v >= nb[1] && v >= nb[2] && v >= nb[3])) return;
        } else {
          const int slot = atomicAdd(&s_nund, 1);
          if (slot < UND_CAP) s_list[slot] = y * w + x;         // re-checked by the whole CTA below
          return;
        }
      }
      acc_list[atomicAdd(&s_nacc, 1)] = y * w + x;
    };
    const int nhot = s_nhot;
    if (!exact_mode && s_finite && nhot <= HOT_CAP) {
      // the usual case: a few dozen hot items of 12 pixels, one pixel per thread
      constexpr int IR = PACKED ? 2 : 1, IC = PACKED ? SEG2 : SEG;     // item rows x columns
      for (int i = tid; i < nhot * (IR * IC); i += 256) {
        const int it = i / (IR * IC), j = i - it * (IR * IC);
        const int e0 = s_hot[it], y0 = e0 / w, x = e0 - y0 * w + j % IC;
        if (x < w) test_pixel(y0 + j / IC, x);
      }
    } else {
      // (outside the active box bufA still holds the raw map; in the float64 round it is valid everywhere, but no
      // pixel out there can pass the threshold either way)
      const int xe = xa1 + 1;
      for (int y = ya0 + warp; y <= ya1; y += 8) {
        const float* rowp = bufA + y * wp;
        for (int x0 = xa0 + lane; x0 < xe; x0 += 128) {
          // threshold test of four pixels per lane first
          float v4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v4[j] = x0 + 32 * j < xe ? rowp[x0 + 32 * j] : 0.f;
          unsigned pass = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (x0 + 32 * j < xe && !(v4[j] < BLUR_THRESH - fmaf(er, v4[j], ea))) pass |= 1u << j;
          while (pass) {
            const int x = x0 + 32 * (__ffs(pass) - 1);
            pass &= pass - 1;
            test_pixel(y, x);
          }
        }
      }
    }
    __syncthreads();
    if (exact_mode || s_nund <= UND_CAP) break;
    for (int e = tid; e < hw; e += 256) {                       // float64 pass 1 -> bufB
      const int y = e / w, x = e - y * w;
      bufB[y * wp + x] = exact_pass1(src, y, x, h, w, gw);
    }
    __syncthreads();
    for (int e = tid; e < hw; e += 256) {                       // float64 pass 2 -> bufA
      const int y = e / w, x = e - y * w;
      bufA[y * wp + x] = exact_pass2(bufB + y * wp, x, w, gw);
    }
    if (tid == 0) s_nacc = 0;
    __syncthreads();
    exact_mode = true;
  }

  // exact re-check of the undecided pixels, one at a time by the whole CTA: 81 threads evaluate the
  // float64 pass 1 at the 3 x 27 positions the five blurred values need, 5 threads finish pass 2.
  const int nund = exact_mode ? 0 : s_nund;
  for (int i = 0; i < nund; ++i) {
    const int e = s_list[i], y = e / w, x = e - y * w;
    if (tid < 81) {
      const int r = tid / 27, k = tid - r * 27, yy = y - 1 + r;
      if (yy >= 0 && yy < h) s_t[r * 32 + k] = exact_pass1(src, yy, reflect_idx(x - 13 + k, w), h, w, gw);
    }
    __syncthreads();
    if (tid < 5) {
      // 0 centre, 1 up, 2 down, 3 left, 4 right
      const int r = tid == 1 ? 0 : (tid == 2 ? 2 : 1), k0 = tid == 3 ? 12 : (tid == 4 ? 14 : 13);
      const int yy = y - 1 + r, xx = x - 13 + k0;
      float val = 0.f;                                          // outside the map
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
        const float* t1 = s_t + r * 32 + k0;
        double t = __dmul_rn((double)t1[0], gw.w[GR]);
        for (int ii = -GR; ii < 0; ++ii)
          t = __dadd_rn(t, __dmul_rn(__dadd_rn((double)t1[ii], (double)t1[-ii]), gw.w[ii + GR]));
        val = (float)t;
      }
      s_v[tid] = val;
    }
    __syncthreads();
    if (tid == 0) {
      const float ve = s_v[0];
      const bool pk = ve > BLUR_THRESH && ve >= s_v[1] && ve >= s_v[2] && ve >= s_v[3] && ve >= s_v[4];
      if (pk) acc_list[s_nacc++] = e;                           // only this thread touches the list here
    }
  }
  __syncthreads();

  // centroids of the accepted peaks, spread over the CTA
  Cand first, second;
  first.pos = second.pos = 0x7fffffff;
  first.cx = first.cy = second.cx = second.cy = 0.0;
  first.score = second.score = 0.f;
  int mine = 0;
  const int nacc = s_nacc;
  for (int i = tid; i < nacc; i += 256) {
    const int e = acc_list[i], y = e / w;
    ++mine;
    Cand c = make_cand(orig, y, e - y * w, h, w);
    cand_insert(c, first, second);
  }
  if (tid == 0 && s_nund) atomicAdd(&g_exact_rechecks, (unsigned long long)(exact_mode ? hw : s_nund));
  __syncthreads();                                              // cands alias bufB: every reader is done
  reduce_and_emit(first, second, mine, cands, &s_count, orig, reg, tracking, scores, inds, xs, ys, cts_wreg,
                  trk, bc, b, h, w);
}

// ---------------------------------------------------------------------------------------
// _nms (+ optional per-channel top-K).  grid B*C, 256 threads, smem 2*h*w floats.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void block_argmax_excluding(const float* vals, int n, const int* taken,
                                                       int ntaken, float* s_val, int* s_idx,
                                                       float& outv, int& outi) {
  float best = -INFINITY; int bi = 0x7fffffff;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    bool tk = false;
    for (int t = 0; t < ntaken; ++t) tk |= (taken[t] == e);
    if (tk) continue;
    float v = vals[e];
    if (bi == 0x7fffffff || v > best) { best = v; bi = e; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bv = s_val[0]; int bb = s_idx[0];
    for (int wv = 1; wv < (int)(blockDim.x >> 5); ++wv) {
      float ov = s_val[wv]; int oi = s_idx[wv];
      if (oi != 0x7fffffff && (bb == 0x7fffffff || ov > bv || (ov == bv && oi < bb))) { bv = ov; bb = oi; }
    }
    s_val[0] = bv; s_idx[0] = bb;
  }
  __syncthreads();
  outv = s_val[0]; outi = s_idx[0];
  __syncthreads();
}

__global__ void __launch_bounds__(256)
nms_topk_channel_kernel(const float* __restrict__ hm, float* __restrict__ nms_out,
                        float* __restrict__ ws_val, int* __restrict__ ws_idx, int h, int w, int K) {
  extern __shared__ __align__(16) float nsm[];
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ int s_taken[64];
  const int hw = h * w;
  float* ori = nsm;
  float* kept = nsm + hw;
  const float* src = hm + (long long)blockIdx.x * hw;
  for (int e = threadIdx.x; e < hw; e += 256) ori[e] = __ldg(src + e);
  __syncthreads();
  for (int e = threadIdx.x; e < hw; e += 256) {
    int y = e / w, x = e % w;
    float v = ori[e], mx = v;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        int yy = y + dy, xx = x + dx;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) mx = fmaxf(mx, ori[yy * w + xx]);
      }
    float kv = v * ((mx == v) ? 1.f : 0.f);
    kept[e] = kv;
    if (nms_out) nms_out[(long long)blockIdx.x * hw + e] = kv;
  }
  __syncthreads();
  for (int r = 0; r < K; ++r) {
    float bv; int bi;
    block_argmax_excluding(kept, hw, s_taken, r, s_val, s_idx, bv, bi);
    if (threadIdx.x == 0) {
      s_taken[r] = bi;
      ws_val[(long long)blockIdx.x * K + r] = bv;
      ws_idx[(long long)blockIdx.x * K + r] = bi;
    }
    __syncthreads();
  }
}

// grid B, 256 threads: top-K over the C*K per-channel candidates (value desc, position asc)
__global__ void __launch_bounds__(256)
topk_merge_kernel(const float* __restrict__ ws_val, const int* __restrict__ ws_idx,
                  float* __restrict__ scores, long long* __restrict__ inds, int* __restrict__ clses,
                  int CK, int K) {
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ int s_taken[64];
  const float* vals = ws_val + (long long)blockIdx.x * CK;
  for (int r = 0; r < K; ++r) {
    float bv; int bi;
    block_argmax_excluding(vals, CK, s_taken, r, s_val, s_idx, bv, bi);
    if (threadIdx.x == 0) {
      s_taken[r] = bi;
      scores[(long long)blockIdx.x * K + r] = bv;
      inds[(long long)blockIdx.x * K + r] = ws_idx[(long long)blockIdx.x * CK + bi];
      clses[(long long)blockIdx.x * K + r] = bi / K;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// SoftArgmaxPavlo: 7x7 average pool (zeros counted), softmax(beta * (pooled - max)), E[x], E[y]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
soft_argmax_kernel(const float* __restrict__ hm, float* __restrict__ out, int h, int w, float beta,
                   float size_mult) {
  extern __shared__ __align__(16) float ssm[];
  __shared__ float s_red[3][8];
  const int hw = h * w;
  float* ori = ssm;
  float* pooled = ssm + hw;
  const float* src = hm + (long long)blockIdx.x * hw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < hw; e += 256) ori[e] = __ldg(src + e);
  __syncthreads();
  float mx = -INFINITY;
  for (int e = threadIdx.x; e < hw; e += 256) {
    int y = e / w, x = e % w;
    float s = 0.f;
    for (int dy = -3; dy <= 3; ++dy) {
      int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
      for (int dx = -3; dx <= 3; ++dx) {
        int xx = x + dx;
        if (xx >= 0 && xx < w) s += ori[yy * w + xx];
      }
    }
    s = s / 49.f;
    pooled[e] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[0][warp] = mx;
  __syncthreads();
  mx = s_red[0][0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_red[0][i]);
  __syncthreads();
  float se = 0.f, sx = 0.f, sy = 0.f;
  for (int e = threadIdx.x; e < hw; e += 256) {
    int y = e / w, x = e % w;
    float ex = expf(beta * (pooled[e] - mx));
    se += ex; sx += ex * ((float)x * size_mult); sy += ex * ((float)y * size_mult);
  }
  se = warp_sum(se); sx = warp_sum(sx); sy = warp_sum(sy);
  if (lane == 0) { s_red[0][warp] = se; s_red[1][warp] = sx; s_red[2][warp] = sy; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bx = 0.f, by = 0.f;
    for (int i = 0; i < 8; ++i) { a += s_red[0][i]; bx += s_red[1][i]; by += s_red[2][i]; }
    float inv = 1.f / (a + 1e-8f);
    out[2 * blockIdx.x] = bx * inv;
    out[2 * blockIdx.x + 1] = by * inv;
  }
}

}  // namespace sgta

using namespace sgta;

static int launch_peaks_f64(const void* hm, const void* reg, const void* tracking, void* scores, void* inds,
                            void* xs, void* ys, void* cts_wreg, void* trk, const GaussW& gw, int B, int C,
                            int h, int w, cudaStream_t st) {
  const size_t fast = sizeof(double) * ((size_t)(h + 2 * GR) * w + (size_t)h * (w + 2 * GR)) + sizeof(float) * (size_t)h * w;
  if (fast <= 226 * 1024 && sizeof(double) * (size_t)(h + 2 * GR) * w >= sizeof(Cand) * 512) {
    cudaFuncSetAttribute(decode_peaks_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast);
    decode_peaks_kernel<true><<<B * C, 256, fast, st>>>(
        (const float*)hm, (const float*)reg, (const float*)tracking, (float*)scores, (long long*)inds,
        (long long*)xs, (long long*)ys, (float*)cts_wreg, (float*)trk, gw, C, h, w);
    return check_launch("decode_peaks_kernel");
  }
  size_t maps = ((sizeof(float) * 3 * (size_t)h * w + 15) / 16) * 16;
  size_t smem = maps + sizeof(Cand) * 512;
  SGTA_REQUIRE(smem <= 220 * 1024, "sgta_decode_peaks: heatmap %dx%d too large for shared memory", h, w);
  cudaFuncSetAttribute(decode_peaks_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  decode_peaks_kernel<false><<<B * C, 256, smem, st>>>(
      (const float*)hm, (const float*)reg, (const float*)tracking, (float*)scores, (long long*)inds,
      (long long*)xs, (long long*)ys, (float*)cts_wreg, (float*)trk, gw, C, h, w);
  return check_launch("decode_peaks_kernel");
}

static int check_peaks_args(const void* hm, void* scores, void* inds, void* xs, void* ys, void* cts_wreg,
                            const void* tracking, void* trk, const double* gauss_w, int B, int C, int h, int w) {
  SGTA_REQUIRE(hm && scores && inds && xs && ys && cts_wreg && gauss_w, "sgta_decode_peaks: null pointer");
  SGTA_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0, "sgta_decode_peaks: bad shape");
  SGTA_REQUIRE(!tracking || trk, "sgta_decode_peaks: tracking given without an output buffer");
  return SGTA_OK;
}

static int g_decode_full_map = 0;
extern "C" int sgta_decode_full_map(int on) { int old = g_decode_full_map; g_decode_full_map = on ? 1 : 0; return old; }

extern "C" int sgta_decode_peaks(const void* hm, const void* reg, const void* tracking, void* scores,
                                 void* inds, void* xs, void* ys, void* cts_wreg, void* trk,
                                 const double* gauss_w, int B, int C, int h, int w, void* stream) {
  int rc = check_peaks_args(hm, scores, inds, xs, ys, cts_wreg, tracking, trk, gauss_w, B, C, h, w);
  if (rc) return rc;
  GaussW gw;
  GaussWF gf;
  for (int i = 0; i < 25; ++i) { gw.w[i] = gauss_w[i]; gf.w[i] = (float)gauss_w[i]; }
  gf.full_map = g_decode_full_map;
  // one plane: h rows of odd stride, at least the 512 block-reduction candidates that alias the second one
  size_t plane = (size_t)h * (w | 1);
  if (plane * sizeof(float) < sizeof(Cand) * 512) plane = sizeof(Cand) * 512 / sizeof(float);
  plane = (plane + 3) & ~(size_t)3;
  const size_t smem = 2 * plane * sizeof(float);
  if (smem <= 226 * 1024) {
    // packed pairs need even sizes (8-byte aligned pair loads) and single-fold reflection
    const bool packed = !(h & 1) && !(w & 1) && h >= SEG2 + 2 * GR && w >= SEG2 + 2 * GR;
    auto kern = packed ? decode_peaks_f32_kernel<true> : decode_peaks_f32_kernel<false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    launch_k(kern, B * C, 256, smem, (cudaStream_t)stream,
             (const float*)hm, (const float*)reg, (const float*)tracking, (float*)scores, (long long*)inds,
             (long long*)xs, (long long*)ys, (float*)cts_wreg, (float*)trk, gw, gf, C, h, w, (int)plane);
    return check_launch("decode_peaks_f32_kernel");
  }
  return launch_peaks_f64(hm, reg, tracking, scores, inds, xs, ys, cts_wreg, trk, gw, B, C, h, w,
                          (cudaStream_t)stream);
}

extern "C" int sgta_decode_peaks_exact64(const void* hm, const void* reg, const void* tracking, void* scores,
                                         void* inds, void* xs, void* ys, void* cts_wreg, void* trk,
                                         const double* gauss_w, int B, int C, int h, int w, void* stream) {
  int rc = check_peaks_args(hm, scores, inds, xs, ys, cts_wreg, tracking, trk, gauss_w, B, C, h, w);
  if (rc) return rc;
  GaussW gw;
  for (int i = 0; i < 25; ++i) gw.w[i] = gauss_w[i];
  return launch_peaks_f64(hm, reg, tracking, scores, inds, xs, ys, cts_wreg, trk, gw, B, C, h, w,
                          (cudaStream_t)stream);
}

extern "C" int sgta_decode_recheck_count(unsigned long long* count, int reset) {
  SGTA_REQUIRE(count, "sgta_decode_recheck_count: null pointer");
  if (cudaMemcpyFromSymbol(count, g_exact_rechecks, sizeof(*count)) != cudaSuccess) {
    set_error("sgta_decode_recheck_count: %s", cudaGetErrorString(cudaGetLastError()));
    return SGTA_ECUDA;
  }
  if (reset) {
    const unsigned long long z = 0ull;
    cudaMemcpyToSymbol(g_exact_rechecks, &z, sizeof(z));
  }
  return SGTA_OK;
}

static int launch_nms(const void* hm, void* nms_out, float* ws_val, int* ws_idx, int B, int C, int h,
                      int w, int K, cudaStream_t st) {
  size_t smem = sizeof(float) * 2 * (size_t)h * w;
  SGTA_REQUIRE(smem <= 220 * 1024, "nms: heatmap %dx%d too large for shared memory", h, w);
  cudaFuncSetAttribute(nms_topk_channel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  nms_topk_channel_kernel<<<B * C, 256, smem, st>>>((const float*)hm, (float*)nms_out, ws_val, ws_idx,
                                                   h, w, K);
  return check_launch("nms_topk_channel_kernel");
}

extern "C" int sgta_decode_nms_topk(const void* hm, void* scores, void* inds, void* clses,
                                    void* workspace, int B, int C, int h, int w, int K, void* stream) {
  SGTA_REQUIRE(hm && scores && inds && clses && workspace, "sgta_decode_nms_topk: null pointer");
  SGTA_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0 && K > 0 && K <= 64 && K <= h * w,
               "sgta_decode_nms_topk: bad shape (K <= 64)");
  float* ws_val = (float*)workspace;
  int* ws_idx = (int*)(ws_val + (size_t)B * C * K);
  int rc = launch_nms(hm, nullptr, ws_val, ws_idx, B, C, h, w, K, (cudaStream_t)stream);
  if (rc) return rc;
  topk_merge_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(ws_val, ws_idx, (float*)scores,
                                                        (long long*)inds, (int*)clses, C * K, K);
  return check_launch("topk_merge_kernel");
}

extern "C" int sgta_nms3x3(const void* hm, void* out, int B, int C, int h, int w, void* stream) {
  SGTA_REQUIRE(hm && out, "sgta_nms3x3: null pointer");
  SGTA_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0, "sgta_nms3x3: bad shape");
  return launch_nms(hm, out, nullptr, nullptr, B, C, h, w, 0, (cudaStream_t)stream);
}

extern "C" int sgta_soft_argmax(const void* hm, void* out, int B, int C, int h, int w, float beta,
                                float size_mult, void* stream) {
  SGTA_REQUIRE(hm && out, "sgta_soft_argmax: null pointer");
  SGTA_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0, "sgta_soft_argmax: bad shape");
  size_t smem = sizeof(float) * 2 * (size_t)h * w;
  SGTA_REQUIRE(smem <= 220 * 1024, "sgta_soft_argmax: heatmap too large for shared memory");
  cudaFuncSetAttribute(soft_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  soft_argmax_kernel<<<B * C, 256, smem, (cudaStream_t)stream>>>((const float*)hm, (float*)out, h, w,
                                                                beta, size_mult);
  return check_launch("soft_argmax_kernel");
}
