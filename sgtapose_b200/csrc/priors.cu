// Structure-prior heat maps rendered on the device.
//
// Reference: the host code that turns the re-projected keypoints of the previous frame into the
// four prior maps of the next forward -- sgtapose/utilities.py:1045-1057 get_prev_hm_wo_noise
// (all keypoints max-blended into one [H,W] map), :1085-1098 get_prev_hm_wo_noise_cls (one
// [H/4,W/4] map per keypoint), both through :800-824 draw_umich_gaussian (radius 4, sigma 2,
// INTEGER centre = int(c), nothing drawn unless the 9x9 patch plus one pixel fits) and :846-853
// gaussian2D; called per clip and frame at lib/sgta_detector.py:528-540 and followed by four
// host->device copies (1.7 MB per clip at 384^2).  Here the host sends 2 x 7 centres per clip and
// one launch writes every pixel of both maps (the zero fill is part of the same pass), for all
// clips of the lock-step batch.  The affine + clip of the centres (utilities.py:943-972) stays
// on the host in float64 exactly as the reference does it (sgtapose_b200/priors.py).
//
// HBM-bound: algorithmic bytes = the maps written, B * (H*W + K*h*w) * 4.
#include "common.cuh"

namespace sgta {

constexpr int PR = 4;                       // radius
struct GaussTab { float g[(2 * PR + 1) * (2 * PR + 1)]; };
constexpr int PRI_MAXK = 16;

// one thread = four consecutive pixels of one row of one map
__global__ void __launch_bounds__(256)
render_priors_kernel(const double* __restrict__ c_in, const double* __restrict__ c_out, float* __restrict__ hm,
                     float* __restrict__ cls, GaussTab tab, int B, int K, int H, int W, int h, int w,
                     long long n_hm4, long long n_total4) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_total4;
       e += (long long)gridDim.x * blockDim.x) {
    const bool is_hm = e < n_hm4;
    const long long q = is_hm ? e : e - n_hm4;
    const int mw = is_hm ? W : w, mh = is_hm ? H : h;
    const int w4 = mw >> 2;
    const int x0 = (int)(q % w4) * 4;
    const long long t = q / w4;
    const int y = (int)(t % mh);
    const long long map = t / mh;                       // hm: b ; cls: b*K + k
    const int b = is_hm ? (int)map : (int)(map / K);
    const int k0 = is_hm ? 0 : (int)(map % K), k1 = is_hm ? K : k0 + 1;
    const double* cc = (is_hm ? c_in : c_out) + (size_t)b * K * 2;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = k0; k < k1; ++k) {
      const int cx = (int)__ldg(cc + 2 * k), cy = (int)__ldg(cc + 2 * k + 1);     // int(): truncation
      if (!(cx - PR >= 0 && cx + PR + 1 < mw && cy - PR >= 0 && cy + PR + 1 < mh)) continue;
      const int dy = y - cy;
      if (dy < -PR || dy > PR) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int dx = x0 + i - cx;
        if (dx >= -PR && dx <= PR) v[i] = fmaxf(v[i], tab.g[(dy + PR) * (2 * PR + 1) + dx + PR]);
      }
    }
    float* dst = is_hm ? hm + (size_t)q * 4 : cls + (size_t)q * 4;
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

}  // namespace sgta

using namespace sgta;

extern "C" int sgta_render_priors(const void* centres_in, const void* centres_out, void* hm, void* hm_cls,
                                  const float* gauss9x9, int B, int K, int H, int W, int h, int w, void* stream) {
  SGTA_REQUIRE(gauss9x9 && (hm || hm_cls), "sgta_render_priors: null pointer");
  SGTA_REQUIRE((!hm || centres_in) && (!hm_cls || centres_out), "sgta_render_priors: a map needs its centres");
  SGTA_REQUIRE(B > 0 && K > 0 && K <= PRI_MAXK && H > 0 && W > 0 && h > 0 && w > 0, "sgta_render_priors: bad shape");
  SGTA_REQUIRE(W % 4 == 0 && w % 4 == 0, "sgta_render_priors: map widths must be multiples of 4 (got %d, %d)", W, w);
  GaussTab tab;
  for (int i = 0; i < 81; ++i) tab.g[i] = gauss9x9[i];
  const long long n_hm4 = hm ? (long long)B * H * (W / 4) : 0;
  const long long n_cls4 = hm_cls ? (long long)B * K * h * (w / 4) : 0;
  const long long total = n_hm4 + n_cls4;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;      // grid-stride, a multiple of the SM count
  render_priors_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const double*)centres_in, (const double*)centres_out, (float*)hm, (float*)hm_cls, tab, B, K, H, W, h, w, n_hm4,
      total);
  return check_launch("render_priors_kernel");
}
