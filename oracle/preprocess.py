"""CPU restatement of the reference's image pre-processing -- TEST INFRASTRUCTURE.

Only tests/ and the golden generator may import this.  Restates, for uint8 HWC images:

  transform_scale      SGTADetector._transform_scale   sgtapose/lib/sgta_detector.py:334-366
  get_affine_transform lib/utils/image.py:45-78 (rot = 0; the 3-point solve is cv2.getAffineTransform,
                       exactly as the reference calls it -- a host-side 6x6 solve, not device work)
  warp_affine_u8       cv2.warpAffine(img, M, (W, H), flags=cv2.INTER_LINEAR) as called at
                       sgta_detector.py:381-383.  OpenCV is a third-party dependency of the reference
                       (requirements.txt:7 `opencv-python`, un-pinned; 4.13.0 in this image); its published
                       algorithm (modules/imgproc/src/imgwarp.cpp: warpAffine -> WarpAffineInvoker ->
                       remapBilinear<FixedPtCast<int, uchar, 15>>) is restated here in integer arithmetic:
                         inverse matrix in float64;  X = (rint((M1 y + M2) 2^10) + 16 + rint(M0 x 2^10)) >> 5
                         integer source pixel X >> 5, 5-bit fraction X & 31 (same for Y);
                         weights (32-fy)(32-fx)*32 ... (exact products: the table's sum fix-up never fires);
                         out = (sum w_i src_i + 2^14) >> 15, BORDER_CONSTANT 0 outside.
                       Pinned bit-exactly to cv2 itself (tests/test_oracle.py) and to the reference's
                       pre_process output (tests/golden/preprocess.npz, oracle/make_golden_preprocess.py).
  normalize            SGTADetector.normalize_img :402-403: ((img / 255.) - mean) / std in float32.
  pre_process          :368-399: the network input [1,3,H,W] float32 and the meta matrices.
"""
import numpy as np

AB_BITS, INTER_BITS, COEF_BITS = 10, 5, 15
TAB = 1 << INTER_BITS


def transform_scale(height, width, input_h, input_w, scale=1):
    """fix_res branch (opt.fix_res is True unless --keep_res, opts_parallel.py:341)."""
    new_height, new_width = int(height * scale), int(width * scale)
    c = np.array([new_width / 2., new_height / 2.], dtype=np.float32)
    s = max(height, width) * 1.0
    return c, s, input_w, input_h, new_width, new_height


def get_affine_transform(center, scale, output_size):
    import cv2
    src_w, dst_w, dst_h = scale, output_size[0], output_size[1]
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center
    src[1, :] = center + np.array([0, src_w * -0.5])          # get_dir with rot_rad = 0
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5], np.float32) + np.array([0, dst_w * -0.5], np.float32)
    for p in (src, dst):                                      # get_3rd_point
        d = p[0] - p[1]
        p[2] = p[1] + np.array([-d[1], d[0]], dtype=np.float32)
    return cv2.getAffineTransform(np.float32(src), np.float32(dst))


def invert_affine(M):
    """imgwarp.cpp warpAffine: the 2x3 forward matrix -> the dst->src matrix, float64, same operation order."""
    M = np.array(M, np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    return M


def warp_affine_u8(src, M, dsize):
    W, H = dsize
    h, w = src.shape[:2]
    m = invert_affine(M)
    scale = float(1 << AB_BITS)
    xs = np.arange(W, dtype=np.float64)
    ys = np.arange(H, dtype=np.float64)
    adelta = np.rint(m[0] * xs * scale).astype(np.int64)
    bdelta = np.rint(m[3] * xs * scale).astype(np.int64)
    rd = (1 << AB_BITS) // TAB // 2
    X0 = np.rint((m[1] * ys + m[2]) * scale).astype(np.int64) + rd
    Y0 = np.rint((m[4] * ys + m[5]) * scale).astype(np.int64) + rd
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)             # saturate_cast<short>
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    fx, fy = X & (TAB - 1), Y & (TAB - 1)
    wts = [[(TAB - fy) * (TAB - fx) * 32, (TAB - fy) * fx * 32], [fy * (TAB - fx) * 32, fy * fx * 32]]
    acc = np.zeros((H, W) + src.shape[2:], np.int64)
    for k1 in range(2):
        for k2 in range(2):
            yy, xx = sy + k1, sx + k2
            ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
            v = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].astype(np.int64)
            wk = np.where(ok, wts[k1][k2], 0)
            acc += v * (wk[..., None] if src.ndim == 3 else wk)
    return ((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS).astype(np.uint8)


def normalize(img_u8, mean, std):
    f = img_u8.astype(np.float32) / np.float32(255.)
    return ((f - mean.astype(np.float32)) / std.astype(np.float32)).astype(np.float32)


def pre_process(image, input_h, input_w, down_ratio=4, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """-> (images [1,3,H,W] float32, meta dict with c, s, trans_input, trans_output, sizes)."""
    height, width = image.shape[:2]
    c, s, inp_w, inp_h, _, _ = transform_scale(height, width, input_h, input_w)
    trans_input = get_affine_transform(c, s, [inp_w, inp_h])
    out_h, out_w = inp_h // down_ratio, inp_w // down_ratio
    trans_output = get_affine_transform(c, s, [out_w, out_h])
    warped = warp_affine_u8(image, trans_input, (inp_w, inp_h))          # cv2.resize to the same size is a copy
    x = normalize(warped, np.array(mean, np.float32).reshape(1, 1, 3), np.array(std, np.float32).reshape(1, 1, 3))
    images = np.ascontiguousarray(x.transpose(2, 0, 1))[None]
    meta = {"c": c, "s": s, "height": height, "width": width, "out_height": out_h, "out_width": out_w,
            "inp_height": inp_h, "inp_width": inp_w, "trans_input": trans_input, "trans_output": trans_output}
    return images, meta, warped
