// Image pre-processing on the device.
//
// Reference: SGTADetector.pre_process (sgtapose/lib/sgta_detector.py:368-399) --
//   cv2.warpAffine(image, trans_input, (inp_w, inp_h), flags=cv2.INTER_LINEAR)   (:381-383)
//   ((img / 255.) - mean) / std  in float32, HWC -> [1,3,H,W]                    (:384-386, :402-403)
// per frame on the host, followed by one host->device copy of the float32 tensor (:154).  Here the raw
// uint8 frames of the whole lock-step batch are uploaded as they are (4x fewer bytes) and one launch
// writes the network input.
//
// cv2.warpAffine for 8-bit images is integer arithmetic (OpenCV imgwarp.cpp: warpAffine ->
// WarpAffineInvoker -> remapBilinear<FixedPtCast<int, uchar, 15>>), restated here exactly:
//   dst->src matrix m = inverse of M in float64 (host, same operation order);
//   X = (rint((m1*y + m2) * 2^10) + 16 + rint(m0*x * 2^10)) >> 5, Y likewise with m4, m5, m3;
//   source pixel (X >> 5, Y >> 5) saturated to int16, 5-bit fractions fx = X & 31, fy = Y & 31;
//   weights (32-fy)(32-fx)*32, (32-fy)fx*32, fy(32-fx)*32, fy*fx*32 (the 2^15 fixed-point table; its
//   products are exact, so OpenCV's sum fix-up never fires);
//   out = (sum w_i * src_i + 2^14) >> 15, BORDER_CONSTANT 0 for samples outside the source.
// Bit-exact against cv2 4.13 and the reference's pre_process output (tests/golden/preprocess.npz).
//
// HBM-bound: algorithmic bytes = raw frame read once (3 h w) + network input written (12 H W) per frame.
#include "common.cuh"

namespace sgta {

constexpr int PP_MAXM = 64;
struct WarpMats { double m[PP_MAXM][6]; };      // dst->src matrices, one per frame of the launch (or one shared)

// grid (ceil(W/128), H, frames of this launch), 128 threads: one output pixel (3 channels) per thread
__global__ void __launch_bounds__(128)
preprocess_kernel(const unsigned char* __restrict__ img, float* __restrict__ out, unsigned char* __restrict__ out_u8,
                  const __grid_constant__ WarpMats mats, int shared_m, int b0, int h, int w, int H, int W,
                  float mean0, float mean1, float mean2, float std0, float std1, float std2) {
  const int x = blockIdx.x * 128 + threadIdx.x, y = blockIdx.y, bl = blockIdx.z, b = b0 + bl;
  if (x >= W) return;
  const double* m = mats.m[shared_m ? 0 : bl];
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), 1024.0)) + 16;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), 1024.0)) + 16;
  const int X = (X0 + __double2int_rn(__dmul_rn(__dmul_rn(m[0], (double)x), 1024.0))) >> 5;
  const int Y = (Y0 + __double2int_rn(__dmul_rn(__dmul_rn(m[3], (double)x), 1024.0))) >> 5;
  const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);
  const int fx = X & 31, fy = Y & 31;
  const int wt[4] = {(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32};
  int acc[3] = {0, 0, 0};
  const unsigned char* src = img + (size_t)b * h * w * 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int yy = sy + (k >> 1), xx = sx + (k & 1);
    if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
    const unsigned char* p = src + ((size_t)yy * w + xx) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += wt[k] * (int)__ldg(p + c);
  }
  const float mean[3] = {mean0, mean1, mean2}, sd[3] = {std0, std1, std2};
  const size_t plane = (size_t)H * W, o = (size_t)b * 3 * plane + (size_t)y * W + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int u = (acc[c] + (1 << 14)) >> 15;
    out[o + c * plane] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.f), mean[c]), sd[c]);
    if (out_u8) out_u8[((size_t)b * plane + (size_t)y * W + x) * 3 + c] = (unsigned char)u;
  }
}

// imgwarp.cpp warpAffine: forward 2x3 matrix -> dst->src matrix, float64, same operation order
static void invert_affine(const double* M, double* m) {
  for (int i = 0; i < 6; ++i) m[i] = M[i];
  double D = m[0] * m[4] - m[1] * m[3];
  D = D != 0 ? 1. / D : 0;
  const double A11 = m[4] * D, A22 = m[0] * D;
  m[0] = A11; m[1] *= -D;
  m[3] *= -D; m[4] = A22;
  const double b1 = -m[0] * m[2] - m[1] * m[5];
  const double b2 = -m[3] * m[2] - m[4] * m[5];
  m[2] = b1; m[5] = b2;
}

// Post-processing of the live decode (one detection per class): dream_generic_post_process (lib/utils/post_process.py:93-117:
// detections with score < out_thresh dropped, ct_wreg mapped to raw-image pixels with the float32 inverse output
// affine, image.py:20-26), merge_outputs (sgta_detector.py:955-961: score > out_thresh) and _get_final_kps (:608-651,
// is_ct branch: best score per class).  One thread per (clip, keypoint); float32 arithmetic like the reference's
// np.dot of float32 operands, result widened to float64 (the reference's kps array).
__global__ void post_process_kernel(const float* __restrict__ scores, const float* __restrict__ cts, double* __restrict__ kps,
                                    float t00, float t01, float t02, float t10, float t11, float t12, float thresh,
                                    double missing, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float s = scores[e], x = cts[2 * e], y = cts[2 * e + 1];
  const bool keep = s >= thresh && s > thresh;
  // dot([t0, t1, t2], [x, y, 1]): products summed left to right, no contraction
  const float rx = __fadd_rn(__fadd_rn(__fmul_rn(t00, x), __fmul_rn(t01, y)), t02);
  const float ry = __fadd_rn(__fadd_rn(__fmul_rn(t10, x), __fmul_rn(t11, y)), t12);
  kps[2 * e] = keep ? (double)rx : missing;
  kps[2 * e + 1] = keep ? (double)ry : missing;
}

}  // namespace sgta

using namespace sgta;

extern "C" int sgta_post_process(const void* scores, const void* cts_wreg, void* kps_raw, const float* trans_inv6,
                                 float out_thresh, double missing, int B, int K, void* stream) {
  SGTA_REQUIRE(scores && cts_wreg && kps_raw && trans_inv6 && B > 0 && K > 0, "sgta_post_process: bad arguments");
  const int n = B * K;
  post_process_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      (const float*)scores, (const float*)cts_wreg, (double*)kps_raw, trans_inv6[0], trans_inv6[1], trans_inv6[2],
      trans_inv6[3], trans_inv6[4], trans_inv6[5], out_thresh, missing, n);
  return check_launch("post_process_kernel");
}

extern "C" int sgta_preprocess(const void* img_u8, void* out, void* out_u8, const double* trans, int n_trans,
                               const float* mean3, const float* std3, int B, int h, int w, int H, int W,
                               void* stream) {
  SGTA_REQUIRE(img_u8 && out && trans && mean3 && std3, "sgta_preprocess: null pointer");
  SGTA_REQUIRE(B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "sgta_preprocess: bad shape");
  SGTA_REQUIRE(h <= 32767 && w <= 32767 && H <= 65535, "sgta_preprocess: image too large");
  SGTA_REQUIRE(n_trans == 1 || n_trans == B, "sgta_preprocess: n_trans must be 1 or B (got %d)", n_trans);
  for (int b0 = 0; b0 < B; b0 += PP_MAXM) {
    const int nb = B - b0 < PP_MAXM ? B - b0 : PP_MAXM;
    WarpMats mats;
    if (n_trans == 1) invert_affine(trans, mats.m[0]);
    else for (int i = 0; i < nb; ++i) invert_affine(trans + (size_t)(b0 + i) * 6, mats.m[i]);
    dim3 grid((W + 127) / 128, H, nb);
    preprocess_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
        (const unsigned char*)img_u8, (float*)out, (unsigned char*)out_u8, mats, n_trans == 1, b0, h, w, H, W,
        mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
    int rc = check_launch("preprocess_kernel");
    if (rc) return rc;
  }
  return SGTA_OK;
}
