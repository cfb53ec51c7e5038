"""CPU restatement of the reference's HOST EDGES - TEST INFRASTRUCTURE, like the rest
of oracle/: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.  Not yet used by a
product path; it pins the arithmetic a device-side uint8 front/back end has to reproduce.

    load_frame()        test_online_tra.py:252-264  cv2.imread -> fp32 CHW hr frame, cv2.resize to 360x480,
                                                    /127.5 - 1 -> fp32 CHW network input
    resize_u8_linear()  cv2.resize(img, (w, h)) on uint8, INTER_LINEAR: the algorithm lives in OpenCV (third party,
                        cv2 4.13.0 in this image, not vendored by the reference); restated from its published
                        fixed-point scheme (11-bit coefficients, HResizeLinear / VResizeLinear for uchar) and pinned
                        against cv2 itself in tests/test_oracle_golden.py: bit-exact for the reference's shapes
                        (down-scaling 720p / 1080p frames); on UP-scaling cv2 differs by one grey level in ~0.2 % of
                        the samples (not on the reference's path).
    to_video_frame()    test_online_tra.py:414  fused fp32 HWC frame -> uint8 (numpy astype: truncation toward zero,
                        wrap modulo 256)
"""
import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _linear_coeffs(ssize, dsize):
    """per destination index: first source index and the two 11-bit weights (OpenCV resize.cpp, INTER_LINEAR)"""
    scale = ssize / dsize
    idx = np.empty(dsize, np.int64)
    w = np.empty((dsize, 2), np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if s < 0:
            f, s = np.float32(0), 0
        if s >= ssize - 1:
            f, s = np.float32(0), ssize - 1
        w[d, 0] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(COEF_SCALE)))
        w[d, 1] = int(np.rint(f * np.float32(COEF_SCALE)))
        idx[d] = s
    return idx, w


def resize_u8_linear(src, out_w, out_h):
    """src uint8 [H,W,C] -> uint8 [out_h,out_w,C], cv2.resize(src, (out_w, out_h)) (INTER_LINEAR)"""
    assert src.dtype == np.uint8 and src.ndim == 3
    sh, sw = src.shape[:2]
    xi, xw = _linear_coeffs(sw, out_w)
    yi, yw = _linear_coeffs(sh, out_h)
    s = src.astype(np.int64)
    x1 = np.minimum(xi + 1, sw - 1)
    rows = s[:, xi] * xw[None, :, 0, None] + s[:, x1] * xw[None, :, 1, None]          # horizontal pass, int
    y1 = np.minimum(yi + 1, sh - 1)
    b0, b1 = yw[:, 0][:, None, None], yw[:, 1][:, None, None]
    out = (((b0 * (rows[yi] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2   # vertical pass, uchar cast
    return out.astype(np.uint8)


def load_frame(img_u8, net_h=360, net_w=480):
    """uint8 HWC (as cv2.imread returns it) -> (hr fp32 [1,3,H,W] 0..255, lr fp32 [1,3,net_h,net_w] in [-1,1])"""
    hr = np.transpose(img_u8.astype(np.float32), [2, 0, 1])[None]
    lr = resize_u8_linear(img_u8, net_w, net_h).astype(np.float32)
    lr = np.transpose(lr, [2, 0, 1])
    lr = (lr / 127.5) - 1.0
    return hr, lr[None].astype(np.float32)


def to_video_frame(fused_hwc):
    """fp32 [Ho,Wo,3] -> uint8, numpy's astype(np.uint8) as the reference applies before VideoWriter.write: C
    conversion through a signed integer, i.e. truncation toward zero and wrap modulo 256"""
    return (np.trunc(fused_hwc).astype(np.int64) & 0xFF).astype(np.uint8)
