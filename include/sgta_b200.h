/*
 * sgta_b200.h -- C ABI of libsgta_b200.so: the B200 (sm_100a) hot path of SGTAPose's
 * per-frame dense inference.
 *
 * Conventions:
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     (the Python host passes torch allocator pointers), unless marked HOST;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream) and never synchronise;
 *   - return 0 on success, a negative SGTA_E* code otherwise; nothing throws across the
 *     ABI; sgta_last_error() returns a HOST string describing the last failure of the
 *     calling thread.
 *
 * Each entry point names the reference interface it replaces (file:line in
 * Nimolty/SGTAPose; the DCNv2 extension itself is third party).
 */
#ifndef SGTA_B200_H_
#define SGTA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGTA_OK 0
#define SGTA_EINVAL (-1)      /* bad argument / unsupported shape            */
#define SGTA_ECUDA (-2)       /* CUDA runtime error at launch                */
#define SGTA_EUNSUPPORTED (-3) /* valid call, but no kernel for this config   */

#define SGTA_DTYPE_F32 0
#define SGTA_DTYPE_BF16 1

int sgta_abi_version(void);
const char* sgta_last_error(void);
/* number of kernels launched by this library since load (bench.py "gpu_launches") */
int64_t sgta_launch_count(void);

/* ---------------------------------------------------------------------------------
 * DCNv2 forward, reference layout.  Replaces `_ext.dcn_v2_forward` of lbin/DCNv2 as
 * reached from `DCN.forward`, constructed at sgtapose/lib/model/networks/dla.py:545 and
 * called at dla.py:548.
 *   x            [B, Cin, H, W]              fp32 NCHW
 *   offset_mask  [B, 3*dg*kh*kw, Ho, Wo]     RAW output of conv_offset_mask: first
 *                2*dg*kh*kw channels are (dy,dx) interleaved per tap, the rest are mask
 *                logits; the sigmoid is applied inside the kernel
 *   weight       [Cout, Cin, kh, kw], bias [Cout] (may be NULL)
 *   y            [B, Cout, Ho, Wo]           fp32 NCHW
 * dtype must be SGTA_DTYPE_F32 (exact fp32 FMA arithmetic, CUDA cores).
 * --------------------------------------------------------------------------------- */
int sgta_dcn_forward(const void* x, const void* offset_mask, const void* weight,
                     const void* bias, void* y, int B, int Cin, int Cout, int H, int W,
                     int kh, int kw, int stride, int pad, int dil, int dgroups, int dtype,
                     void* stream);

/* DCNv2 backward (training config 5).  Replaces `_ext.dcn_v2_backward`.
 * grad_offset_mask is w.r.t. the RAW conv_offset_mask output (sigmoid' folded in).
 * grad_* buffers are OVERWRITTEN, except grad_weight / grad_bias which are accumulated
 * into (caller zero-fills).  Any grad_* may be NULL to skip it. */
int sgta_dcn_backward(const void* x, const void* offset_mask, const void* weight,
                      const void* grad_y, void* grad_x, void* grad_offset_mask,
                      void* grad_weight, void* grad_bias, int B, int Cin, int Cout, int H,
                      int W, int kh, int kw, int stride, int pad, int dil, int dgroups,
                      int dtype, void* stream);

/* activation applied by the convolution epilogues of the engine path below */
#define SGTA_ACT_NONE 0
#define SGTA_ACT_RELU 1
#define SGTA_ACT_SIGMOID 2

/* ---------------------------------------------------------------------------------
 * B200 engine path: zero-bordered "planes" activation layouts + tcgen05 shift-GEMM /
 * gather-GEMM convolutions (csrc/planes.cuh, conv_planes.cu, planes_ops.cu).
 *
 * A view describes images [.. B) of H x W pixels, each inside a frame of
 * (H + 2*border) x (W + 2*border) zero-bordered pixels, flattened to rows
 *   p = (b*(H+2*border) + y+border)*(W+2*border) + x+border,
 * preceded by `guard` rows (and followed by a tail guard inside `rows`).
 *   layout PL: [plane][chunk][row][64 ch] (16-bit), 16-byte groups XOR-swizzled by (row & 7);
 *              `nchunks` = 64-channel chunks per plane of the buffer, `chunk0` = first chunk
 *              of the view; border must be 1;
 *   layout SC: [plane][row][nchunks channels] (16-bit) row-major, channels in {4, 16, 32}.
 * nplanes: 1 = bf16;  2 = fp16 hi + fp16 (v - hi) * 2^11 (fp32 values to ~2^-22 relative;
 * |v| <= 65504).  The zero border must be kept zero by the caller (kernels never write it).
 * --------------------------------------------------------------------------------- */
#define SGTA_LAYOUT_PL 0
#define SGTA_LAYOUT_SC 1
typedef struct sgta_planes {
  void* data;       /* device pointer: plane 0, chunk 0, row 0 (first guard row)        */
  int64_t rows;     /* rows per (plane, chunk), guards included                          */
  int32_t guard;    /* rows before padded pixel 0 of this view                           */
  int32_t nchunks;  /* PL: chunks per plane of the buffer;  SC: channels per row         */
  int32_t chunk0;   /* PL: first chunk of this view                                      */
  int32_t nplanes;  /* 1 or 2                                                            */
  int32_t layout;   /* SGTA_LAYOUT_*                                                     */
  int32_t border;   /* zero-border width of the frame (PL: 1; SC stem input: 3)          */
  int32_t B, H, W;  /* images of the view and their unpadded size                        */
} sgta_planes;

#define SGTA_EPI_PL 0       /* y view, layout PL, same nplanes as x (+ optional residual, act) */
#define SGTA_EPI_SC 1       /* y view, layout SC                                               */
#define SGTA_EPI_F32ROWS 2  /* y_f32[p*ld_f32 + o] for every padded row p (DCN offset/mask)    */
#define SGTA_EPI_NCHW 3     /* y_f32 [B, n_valid, Ho, Wo] fp32 (heads; act may be sigmoid)     */
#define SGTA_EPI_STEM 4     /* Cout == 32 -> SC view with 16 ch: relu(a[c]) + relu(a[16+c])    */
/* super-pixel forms of the first two layers: a GEMM row is a SUPER-PIXEL = 4 consecutive pixels of an
 * image row; a 16-channel map [B,H,W] is then a 64-channel PL view [B,H,W/4] (channel j*16 + c = pixel j, channel c) */
#define SGTA_EPI_STEM_SP 5  /* Cout == 128 = 4 px x (16+16): the dual-stem epilogue per pixel -> 64-channel PL view      */
#define SGTA_EPI_SP2SC 6    /* Cout == 64 = 4 px x 16 ch -> 16-channel SC view [B,H,4*W]: pixel j of row (y,X) -> (y,4X+j) */
#define SGTA_EPI_HEADS 7    /* internal to sgta_planes_conv_heads: fused 1x1 output convolutions, fp32 NCHW outputs            */
/* A 3x3 convolution with EPI_SP2SC and Cin == 64 is taken to be a 16 -> 16 convolution over super-pixels: its weight
 * matrix MUST be the Toeplitz expansion planes.superpixel_weight(w, 4, 4, 1) -- the left / right neighbour taps are
 * zero except for their last / first pixel, and the kernel does not issue those all-zero K steps. */

/* performance experiments only (tools/conv_bench.py): bit 0 skip A copies, bit 1 skip B copies,
 * bit 2 skip the epilogue, bit 4 skip the MMAs in the planes convolutions (results are then garbage); returns the old value */
int sgta_debug_flags(int flags);
/* rows of guard recommended before / after the frames of a W-wide map */
int sgta_planes_guard(int W);
/* output-channel tile the kernels use for (Cout, nplanes); -1 if unsupported */
int sgta_planes_ntile(int Cout, int nplanes);
/* Weight matrix Wm [Cout][Kpad] fp32 (K = (tap, channel), channel fastest, zero padded to
 * Kpad % 64 == 0; Cout % 16 == 0) -> packed SWIZZLE_128B tiles, hi/lo split if nplanes == 2. */
int64_t sgta_planes_wpack_bytes(int Cout, int Kpad, int nplanes);
int sgta_planes_pack_weight(const void* wm_f32, void* wpack, int Cout, int Kpad, int nplanes,
                            void* stream);
/* Convolution on a PL input (dla.py:41-69, :157-175, :212-216; base_model.py:121-135):
 * ksize 1 or 3 (pad = ksize/2), stride 1 (TMA shift-GEMM) or 3x3 stride 2 (gather-GEMM).
 * y = act(scale*conv + shift (+ res)); scale/shift fold bias and eval BatchNorm. */
int sgta_planes_conv(const sgta_planes* x, const void* wpack, const void* scale, const void* shift,
                     const sgta_planes* res, const sgta_planes* y, void* y_f32, int64_t ld_f32,
                     int Cin, int Cout, int ksize, int stride, int act, int epi, int n_valid,
                     void* stream);
/* The head stack of the network in ONE launch (base_model.py:121-135 build, :190-199 call): for every head h
 *   out[h] = act_h( W2_h * relu(scale * conv3x3(x)[h*hid .. (h+1)*hid) + shift) + b2_h )     fp32 [B, nout[h], H, W]
 * -- the 3x3 convolutions of all heads as one Cin -> n_heads*hid shift-GEMM (wpack / scale / shift as for
 * sgta_planes_conv) with the 1x1 output convolutions folded into its epilogue in fp32 FMA arithmetic; the hidden map is
 * never written.  fp32 mode (2 planes) only; hid must be a multiple of the kernel's N tile (128); nout[h] <= 8.
 * w2: device fp32, per head [hid][stride_h] with stride_h = 2 / 4 / 8 for nout <= 2 / 4 / 8 (row k = the nout weights of
 * hidden channel k, zero padded), heads back to back; b2: device fp32 [n_heads][8]; out / nout: HOST arrays [n_heads];
 * bit h of sigmoid_mask applies the detector's sigmoid (sgta_detector.py:854-862) to head h. */
int sgta_planes_conv_heads(const sgta_planes* x, const void* wpack, const void* scale, const void* shift,
                           const void* w2, const void* b2, void* const* out, const int* nout, int n_heads,
                           int sigmoid_mask, int Cin, int hid, void* stream);
/* Convolution on an SC input (stems dla.py:241-270, level0/1 :302-312, level2 entry).  K block
 * kb (64 wide = 128 bytes per output pixel) is made of 8/seg_groups segments; segment j is a
 * run of seg_groups*16 contiguous bytes starting at input row
 *   anchor(b,oy,ox) + seg_off[kb*2 + j],  anchor = frame(b) + (oy*stride)*(W+2*border) + ox*stride_x
 * (HOST table; the caller builds the matching weight matrix). */
int sgta_planes_conv_sc(const sgta_planes* x, const void* wpack, const void* scale, const void* shift,
                        const sgta_planes* y, int Cout, int stride, int stride_x, int Ho, int Wo, int nkb,
                        int seg_groups, const int* seg_off, int act, int epi, void* stream);
/* DeformConv main GEMM (dla.py:538-550): DCNv2 3x3/s1/p1/d1/dg1 + folded bias/BN (+ReLU).
 * offset_mask: fp32 [padded rows][32] raw conv_offset_mask output (SGTA_EPI_F32ROWS). */
int sgta_planes_dcn(const sgta_planes* x, const void* offset_mask, const void* wpack,
                    const void* scale, const void* shift, const sgta_planes* y, int Cin, int Cout,
                    int relu, void* stream);
/* layout converters / memory-bound ops (planes_ops.cu); C, c_off multiples of 8 */
int sgta_planes_from_nchw(const void* src_f32, const sgta_planes* y, int C, int c_off, void* stream);
int sgta_planes_to_nchw(const sgta_planes* x, void* dst_f32, int C, int c_off, void* stream);
/* img [B,3,H,W] + hm [B,1,H,W] fp32 -> images [b_off, b_off+B) of an SC view with 4 channels */
int sgta_planes_pack_stem(const void* img_f32, const void* hm_f32, const sgta_planes* y, int b_off,
                          int B, void* stream);
/* nn.MaxPool2d(2, 2) of Tree.downsample (dla.py:209-210) */
int sgta_planes_maxpool2(const sgta_planes* x, int xc_off, const sgta_planes* y, int yc_off, int C,
                         void* stream);
/* IDAUp: y = ConvTranspose2d(C,C,2f,stride=f,padding=f/2,groups=C)(x) + skip (dla.py:561-577) */
int sgta_planes_upsample_add(const sgta_planes* x, const void* w_up, const sgta_planes* skip,
                             const sgta_planes* y, int C, int f, void* stream);
/* rows[b,t,:] (fp32 [B,n,C]) <-> images [b_off, b_off+B) of the view at ids[b,t] = y*W + x;
 * write-back with duplicate ids: the HIGHEST t wins (dla.py:915-968, :1006-1018) */
int sgta_planes_gather_tokens(const sgta_planes* x, int b_off, const void* ids, void* rows, int B,
                              int C, int n, void* stream);
int sgta_planes_scatter_tokens(const sgta_planes* x, int b_off, const void* ids, const void* rows,
                               int B, int C, int n, void* stream);

/* ---------------------------------------------------------------------------------
 * Structure-prior temporal attention (dla.py:868-887 MHCA_ein.forward core):
 *   out[b,i,h,:] = softmax_j( q[b,i,h,:].k[b,j,h,:] * inv_scale + pos[h,i,j] ) v[b,j,h,:]
 * q [B,nq,heads*d], k,v [B,nk,heads*d], out [B,nq,heads*d] fp32, "b n (h d)" layout (the
 * Linear outputs as they are, no rearrange); pos [heads,nq,nk] fp32 or NULL.
 * d in {4, 8, 16, 32}.
 * --------------------------------------------------------------------------------- */
int sgta_attn_forward(const void* q, const void* k, const void* v, const void* pos,
                      void* out, int B, int heads, int nq, int nk, int d, float inv_scale,
                      void* stream);
/* The same core with K and V given HEAD-MAJOR, [B, heads, nk, d] (written that way by sgta_token_linear_heads): the
 * batch-looping level-0 kernel then prefetches a (sample, head) slab with coalesced loads.  Served shapes:
 * sgta_attn_kvhm_supported(...) != 0 (d == 4, pos given, large pos block, B >= 4); q / out stay "b n (h d)". */
int sgta_attn_kvhm_supported(int B, int heads, int nq, int nk, int d, int has_pos);
int sgta_attn_forward_kvhm(const void* q, const void* k_hm, const void* v_hm, const void* pos, void* out,
                           int B, int heads, int nq, int nk, int d, float inv_scale, void* stream);
/* backward of the same core (config 5): grads w.r.t. q, k, v, pos (pos grad accumulated) */
int sgta_attn_backward(const void* q, const void* k, const void* v, const void* pos,
                       const void* grad_out, void* grad_q, void* grad_k, void* grad_v,
                       void* grad_pos, int B, int heads, int nq, int nk, int d,
                       float inv_scale, void* stream);

/* ---------------------------------------------------------------------------------
 * Prior-guided token selection (dla.py:898-913 get_topk_index, :915-968
 * get_topk_features_scale, :1006-1018 substitute_topk_features_scale).
 * --------------------------------------------------------------------------------- */
/* top-K flat index per (sample, channel); ties: value descending, lowest index first.
 * hm [B,C,HW] fp32 -> idx [B, C*K] int64 */
int sgta_topk_index(const void* hm, void* idx, int B, int C, int HW, int K, void* stream);
/* window ids with the reference's fp32 index arithmetic:
 * idx [B,CK] int64 (flat index in a Whm-wide map) -> ids [B, CK*win*win] int64,
 * win = 2*(kernel/2)+1, coords = (x,y)*scale + offset, clamp [0,H-1], id = y*W + x (fp32),
 * truncated. */
int sgta_window_ids(const void* idx, void* ids, int B, int CK, int Whm, float scale,
                    int kernel, int H, int W, void* stream);
/* rows[b,t,:] = feats[b,:,ids[b,t]]   (feats NCHW fp32 if nhwc==0 else NHWC) */
int sgta_gather_tokens(const void* feats, const void* ids, void* rows, int B, int C, int HW,
                       int n, int nhwc, void* stream);
/* feats[b,:,ids[b,t]] = rows[b,t,:]; duplicate ids: the HIGHEST t wins (deterministic). */
int sgta_scatter_tokens(void* feats, const void* ids, const void* rows, int B, int C, int HW,
                        int n, int nhwc, void* stream);

/* ---------------------------------------------------------------------------------
 * Heatmap decode.
 * --------------------------------------------------------------------------------- */
/* Live decode: dream_generic_decode -> _peaks_info -> peaks_from_belief_maps
 * (sgtapose/lib/model/decode.py:184-313, lib/model/utils.py:207-284,
 *  sgtapose/image_proc.py:1032-1143), one keypoint per channel, every sample decoded.
 *   hm [B,C,h,w] fp32 (post-sigmoid), reg [B,2,h,w] or NULL, tracking [B,2,h,w] or NULL
 *   scores [B,C] f32 (-1 = missing), inds/xs/ys [B,C] int64,
 *   cts_wreg [B,C,2] f32 (x_int + reg_x, y_int + reg_y; +0.5 if reg NULL),
 *   trk [B,C,2] f32 (ignored if tracking NULL)
 * gauss_w: HOST pointer to the 25 float64 blur taps (scipy _gaussian_kernel1d(3,0,12));
 * integer outputs are bit-exact w.r.t. scipy/numpy double arithmetic.
 * The blur is evaluated in float32 and every peak comparison whose margin lies inside a
 * rigorous rounding bound (80 * 2^-24 relative) is re-evaluated with the float64 restatement,
 * so the result equals sgta_decode_peaks_exact64's by construction. */
int sgta_decode_peaks(const void* hm, const void* reg, const void* tracking, void* scores,
                      void* inds, void* xs, void* ys, void* cts_wreg, void* trk,
                      const double* gauss_w, int B, int C, int h, int w, void* stream);
/* Same contract, every pixel blurred with the float64 restatement (the cross-check of the
 * production entry in the parity tests, and the path for maps too large for its shared memory). */
int sgta_decode_peaks_exact64(const void* hm, const void* reg, const void* tracking, void* scores,
                              void* inds, void* xs, void* ys, void* cts_wreg, void* trk,
                              const double* gauss_w, int B, int C, int h, int w, void* stream);
/* Test hook: on != 0 makes sgta_decode_peaks blur every pixel of every map instead of the active box only (the box =
 * rows / columns whose blurred row / column maxima can reach the 0.01 threshold; outputs are identical either way);
 * returns the previous setting. */
int sgta_decode_full_map(int on);
/* Test hook: number of pixels sgta_decode_peaks re-evaluated in float64 since the last reset
 * (synchronises the device; count is a HOST pointer). */
int sgta_decode_recheck_count(unsigned long long* count, int reset);
/* Alternate decode: _nms + _topk (lib/model/utils.py:59-103; generic_decode decode.py:93-94).
 *   scores [B,K] f32, inds [B,K] int64, clses [B,K] int32; workspace >= B*C*K*12 bytes */
int sgta_decode_nms_topk(const void* hm, void* scores, void* inds, void* clses,
                         void* workspace, int B, int C, int h, int w, int K, void* stream);
/* 3x3 max-pool NMS alone (utils.py:59-65): out = hm * (maxpool3x3(hm) == hm) */
int sgta_nms3x3(const void* hm, void* out, int B, int C, int h, int w, void* stream);
/* SoftArgmaxPavlo.forward (sgtapose/spatial_softmax.py:24-95): out [B,C,2] = (E[x], E[y]) */
int sgta_soft_argmax(const void* hm, void* out, int B, int C, int h, int w, float beta,
                     float size_mult, void* stream);

/* ---------------------------------------------------------------------------------
 * Post-attention half of the shared TransformerEncoderLayer, one launch per layer
 * (sgtapose/lib/model/networks/dla.py:734-743 forward, :728-732 forward_ffn, :886-887 fc):
 *   q1 = LayerNorm1(att fc_w^T + fc_b + q);  q2 = LayerNorm3(q1 + w2 relu(w1 q1 + b1) + b2)  -> q_out [T,C]
 *   qp_out [T,hid] = q2 wq_next^T   (next layer's query projection; wq_next NULL = skip)
 * att [T,hid], q [T,C], fc_wt [hid,C] (= fc.weight^T), w1 [dffn,C], w2t [dffn,C] (= linear2.weight^T),
 * wq_next [hid,C]; fp32, row-major,
 * C in {16,32,64}, hid % 4 == 0.  Replaces ~9 library launches (2 GEMMs through a [T,dffn] HBM
 * intermediate, 2 LayerNorms, bias/ReLU/residual kernels). */
int sgta_token_mlp(const void* att, const void* q, const void* fc_wt, const void* fc_b, const void* ln1_w,
                   const void* ln1_b, const void* w1, const void* b1, const void* w2t, const void* b2,
                   const void* ln3_w, const void* ln3_b, const void* wq_next, void* q_out, void* qp_out,
                   int T, int C, int hid, int dffn, float eps, void* stream);

/* Token-row Linear layers of the fusion (dla.py:868-876 w_q / w_k / w_v of MHCA_ein; :1499-1502 cat_layer applied to
 * cat([out, cur_query], -1) at :1006-1018):   y[m,n] = act( sum_k [x1 | x2][m,k] * w[n,k] + bias[n] )
 * x1 [M,K1], x2 [M,K2] or NULL (K2 = 0), w [N,K1+K2] (nn.Linear weight as stored), bias [N] or NULL, y [M,N]; fp32,
 * row-major; K1, K2 multiples of 16, N of 4; relu != 0 applies ReLU.  fp32 FMA, ascending k (replaces library SGEMMs). */
int sgta_token_linear(const void* x1, int K1, const void* x2, int K2, const void* w, const void* bias,
                      void* y, int M, int N, int relu, void* stream);
/* bias-free projection with a head-major result (w_k / w_v of MHCA_ein, dla.py:872-876, for sgta_attn_forward_kvhm):
 * x [B*n_tokens, K], w [N, K] -> y [B, heads, n_tokens, N / heads] */
int sgta_token_linear_heads(const void* x, int K, const void* w, void* y, int M, int N, int n_tokens, int heads,
                            void* stream);

/* ---------------------------------------------------------------------------------
 * Structure-prior maps: replaces the host rendering + 4 H2D copies per clip
 * and frame of lib/sgta_detector.py:528-540 (sgtapose/utilities.py:1045-1057 get_prev_hm_wo_noise,
 * :1085-1098 get_prev_hm_wo_noise_cls, :800-824 draw_umich_gaussian, :846-853 gaussian2D).
 *   centres_in  [B,K,2] float64 DEVICE: keypoint centres (x,y) in network-INPUT pixels, already
 *               affine-transformed and clipped like utilities.py:943-972 (points outside the raw
 *               image = (0,0)); rendered max-blended into hm [B,1,H,W] fp32 (NULL = skip)
 *   centres_out [B,K,2] float64 DEVICE: the same in network-OUTPUT pixels; rendered one map per
 *               keypoint into hm_cls [B,K,h,w] fp32 (NULL = skip)
 *   gauss9x9    HOST pointer, 81 floats: gaussian2D((9,9), sigma=2) rounded to fp32
 * Every pixel of both maps is written (zeros included).  W and w must be multiples of 4. */
int sgta_render_priors(const void* centres_in, const void* centres_out, void* hm, void* hm_cls,
                       const float* gauss9x9, int B, int K, int H, int W, int h, int w, void* stream);

/* Super-pixel view of a 16-channel SC map (round-2 plan; not on the default path yet):
 * sc = SC view, 16 channels, [B,H,W];  sp = PL view, 64 channels (4 pixels x 16 channels), [B,H,W/4].
 * to_super != 0: sc -> sp, else sp -> sc.  Raw 16-byte moves; border super-pixels are left untouched (zero). */
int sgta_planes_superpixels(const sgta_planes* sc, const sgta_planes* sp, int to_super, void* stream);

/* ---------------------------------------------------------------------------------
 * Image pre-processing: SGTADetector.pre_process
 * (sgtapose/lib/sgta_detector.py:368-399) = cv2.warpAffine(image, trans_input, (W, H), INTER_LINEAR)
 * (:381-383) + ((img / 255.) - mean) / std in float32 and HWC -> CHW (:384-386, :402-403), for B frames.
 *   img_u8 [B,h,w,3] uint8 DEVICE (raw frames as cv2.imread returns them)
 *   out    [B,3,H,W] fp32 DEVICE (network input)
 *   out_u8 [B,H,W,3] uint8 DEVICE or NULL: the warped 8-bit image (bit-exact vs cv2, for parity tests)
 *   trans  HOST pointer, n_trans x 6 float64: the FORWARD 2x3 matrices the reference passes to
 *          cv2.warpAffine (trans_input); n_trans = 1 (shared by all frames) or B
 *   mean3 / std3  HOST pointers, 3 floats (sgta_detector.py:58-59)
 * 8-bit results are bit-exact w.r.t. OpenCV's fixed-point bilinear remap (BORDER_CONSTANT 0). */
int sgta_preprocess(const void* img_u8, void* out, void* out_u8, const double* trans, int n_trans,
                    const float* mean3, const float* std3, int B, int h, int w, int H, int W, void* stream);

/* Post-processing of the live decode on the device: dream_generic_post_process
 * (sgtapose/lib/utils/post_process.py:93-117), merge_outputs (lib/sgta_detector.py:955-961) and _get_final_kps
 * (:608-651, is_ct branch) for one detection per class.
 *   scores [B,K] f32, cts_wreg [B,K,2] f32 (network-output pixels) DEVICE
 *   kps_raw [B,K,2] f64 DEVICE: raw-image pixels, `missing` where score < / <= out_thresh
 *   trans_inv6 HOST, 6 floats: get_affine_transform(c, s, 0, (w, h), inv=1).astype(float32), row-major 2x3 */
int sgta_post_process(const void* scores, const void* cts_wreg, void* kps_raw, const float* trans_inv6,
                      float out_thresh, double missing, int B, int K, void* stream);

/* ---------------------------------------------------------------------------------
 * Host-side pose refinement (no device work): the `LM` entry of the reference's
 * binary-only rf_tools/libtestso_final.so, bound at sgtapose/rf_tools/LM.py:10 and called from
 * `register_GN_C` (:256-266).  All pointers are HOST memory.
 *   value_init [7]  qw qx qy qz tx ty tz (start: the PnP pose)
 *   x2d [n,2] detected keypoints (pixels), x3d [n,3] keypoints in the robot frame
 *   weights [2n+2]  per-coordinate weights, last two = the unit-quaternion constraint weights (1e8)
 *   camera [9]      row-major 3x3 intrinsics;   ans [7] refined qw qx qy qz tx ty tz
 * Minimises sum_k F_k^2 with F as in LM.py `fun` (:128-156), any num_points >= 1.
 * `LM` is the drop-in: the reference's own symbol name, argument list and iteration, the final iterate
 * returned as it is (finite or NaN/Inf, exactly what the reference binary hands back).
 * `sgta_lm_refine` is the same iteration plus ONE DELIBERATE DEVIATION: a result that puts a weighted
 * keypoint behind the camera (a mirror pose the chaotic Gauss-Newton can land on) is returned as NaN, so
 * that the caller's existing NaN fall-back (analysis.py:206-210) keeps the PnP pose. */
int sgta_lm_refine(const double* value_init, const double* x2d, const double* x3d, const double* weights,
                   const double* camera, double* ans, int num_points);
void LM(double* value_init, double* x2d, double* x3d, double* weights, double* camera, double* ans, int num_points);

#ifdef __cplusplus
}
#endif
#endif /* SGTA_B200_H_ */
