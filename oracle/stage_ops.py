"""Closed-form CPU restatements of every cv2/numpy call on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function cites the
reference call site it restates (paths are into /root/reference) and is
checked against the real ``cv2`` call in tests/test_oracle_vs_cv2.py.
Integer ops are written with numpy integer arithmetic only, so they are an
independent statement of what OpenCV computes, not a wrapper around it.  The
one exception is the float32 DCT (``cv2.dct``/``cv2.idct``, Intel IPP inside
opencv-python): that algorithm is closed source, so ``block_dct_quantise``
calls cv2 itself -- batched through ``cv2.DCT_ROWS``, which
tests/test_oracle_vs_cv2.py shows is bit-identical to the per-block 2-D call
the reference makes.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# colour conversions
# ----------------------------------------------------------------------------

def bgr2gray(bgr: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(frame, COLOR_BGR2GRAY) on uint8 (frame_differencing.py:75,92;
    motion_compression_opt.py:60,71,149,181).  15-bit fixed point, round half up."""
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


def _sat_u8(x: np.ndarray) -> np.ndarray:
    return np.clip(x, 0, 255).astype(np.uint8)


def bgr2ycrcb(bgr: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(frame, COLOR_BGR2YCrCb) on uint8 (frame_differencing.py:115;
    motion_compression_opt.py:152).  14-bit fixed point."""
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    y = (1868 * b + 9617 * g + 4899 * r + 8192) >> 14
    half = 128 << 14
    cr = ((r - y) * 11682 + half + 8192) >> 14
    cb = ((b - y) * 9241 + half + 8192) >> 14
    return np.stack([_sat_u8(y), _sat_u8(cr), _sat_u8(cb)], axis=-1)


def ycrcb2bgr(ycrcb: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(frame, COLOR_YCrCb2BGR) on uint8 (frame_differencing.py:130;
    motion_compression_opt.py:171)."""
    y = ycrcb[..., 0].astype(np.int32)
    cr = ycrcb[..., 1].astype(np.int32) - 128
    cb = ycrcb[..., 2].astype(np.int32) - 128
    b = y + ((29049 * cb + 8192) >> 14)
    g = y + ((-5636 * cb - 11698 * cr + 8192) >> 14)
    r = y + ((22987 * cr + 8192) >> 14)
    return np.stack([_sat_u8(b), _sat_u8(g), _sat_u8(r)], axis=-1)


# ----------------------------------------------------------------------------
# mask front end
# ----------------------------------------------------------------------------

def _reflect101_pad(img: np.ndarray, p: int) -> np.ndarray:
    return np.pad(img, p, mode="reflect")       # numpy 'reflect' == BORDER_REFLECT_101


def gaussian_blur5(gray: np.ndarray) -> np.ndarray:
    """cv2.GaussianBlur(gray, (5, 5), 0) on uint8 (frame_differencing.py:93).
    sigma=0, ksize 5 selects the fixed binomial kernel [1,4,6,4,1]/16; OpenCV
    evaluates it in 8.8 fixed point, which reduces to (sum + 128) >> 8 over the
    separable 5x5 integer kernel with BORDER_REFLECT_101."""
    h, w = gray.shape
    if h < 3 or w < 3:
        # reflect-101 needs at least 3 samples for a 2-pixel border; cv2 handles
        # tiny images by repeated reflection -- do the same by index arithmetic.
        return _gaussian_blur5_small(gray)
    k = np.array([1, 4, 6, 4, 1], np.int32)
    p = _reflect101_pad(gray.astype(np.int32), 2)
    hor = sum(k[i] * p[:, i:i + w] for i in range(5))
    ver = sum(k[i] * hor[i:i + h, :] for i in range(5))
    return ((ver + 128) >> 8).astype(np.uint8)


def _reflect101_index(i: int, n: int) -> int:
    if n == 1:
        return 0
    period = 2 * (n - 1)
    i %= period
    return i if i < n else period - i


def _gaussian_blur5_small(gray: np.ndarray) -> np.ndarray:
    h, w = gray.shape
    k = (1, 4, 6, 4, 1)
    g = gray.astype(np.int64)
    out = np.zeros((h, w), np.int64)
    for y in range(h):
        for x in range(w):
            s = 0
            for j in range(5):
                yy = _reflect101_index(y + j - 2, h)
                for i in range(5):
                    xx = _reflect101_index(x + i - 2, w)
                    s += k[j] * k[i] * g[yy, xx]
            out[y, x] = (s + 128) >> 8
    return out.astype(np.uint8)


_SMALL_GAUSS = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
                7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125]}


def gaussian_kernel_fixed(ksize: int, sigma: float) -> np.ndarray:
    """OpenCV's 8.8 fixed-point Gaussian kernel for uint8 images (getGaussianKernelBitExact +
    getGaussianKernelFixedPoint_ED): the normalised double kernel rounded to 8 fractional bits from the ends towards the
    middle with the rounding error carried along; the taps sum to exactly 256.  (25, 30) -> 10 10 10 10 10 10 10 11 10 11 ..."""
    n = int(ksize)
    if sigma <= 0 and n in _SMALL_GAUSS:
        k = np.array(_SMALL_GAUSS[n], np.float64)
    else:
        sg = sigma if sigma > 0 else ((n - 1) * 0.5 - 1) * 0.3 + 0.8
        c = (n - 1) * 0.5
        k = np.exp((-0.5 / (sg * sg)) * (np.arange(n) - c) ** 2)
        k = k / k.sum()
    res = np.zeros(n, np.int64)
    err = 0.0
    for i in range(n // 2):
        adj = k[i] * 256.0 + err
        v = int(np.rint(adj))
        err = adj - v
        res[i] = res[n - 1 - i] = v
    res[n // 2] = int(np.rint(k[n // 2] * 256.0 + err))
    return res


def gaussian_blur_fixed(gray: np.ndarray, ksize: int, sigma: float) -> np.ndarray:
    """cv2.GaussianBlur(gray, (ksize, ksize), sigma) on uint8 (frame_differencing.py:77,93): separable 8.8 fixed point,
    BORDER_REFLECT_101, (v + 2^15) >> 16."""
    k = gaussian_kernel_fixed(ksize, sigma)
    r = len(k) // 2
    h, w = gray.shape
    yi = np.array([_reflect101_index(i, h) for i in range(-r, h + r)])
    xi = np.array([_reflect101_index(i, w) for i in range(-r, w + r)])
    p = gray.astype(np.int64)[yi][:, xi]
    hs = sum(int(c) * p[:, i:i + w] for i, c in enumerate(k))
    vs = sum(int(c) * hs[j:j + h, :] for j, c in enumerate(k))
    return np.clip((vs + (1 << 15)) >> 16, 0, 255).astype(np.uint8)


def absdiff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """cv2.absdiff on uint8 (frame_differencing.py:96)."""
    return np.abs(a.astype(np.int16) - b.astype(np.int16)).astype(np.uint8)


def threshold_floor(thresh: float) -> int:
    """cv2.threshold on uint8 floors the float threshold: 0.5 -> 0, 25 -> 25."""
    return int(np.floor(thresh))


def threshold_binary(diff: np.ndarray, thresh: float, maxval: int = 255) -> np.ndarray:
    """cv2.threshold(diff, thresh, 255, THRESH_BINARY) on uint8 (frame_differencing.py:97)."""
    return np.where(diff.astype(np.int32) > threshold_floor(thresh), maxval, 0).astype(np.uint8)


# ----------------------------------------------------------------------------
# morphology
# ----------------------------------------------------------------------------

def structuring_rect(k: int) -> np.ndarray:
    """np.ones((k, k), np.uint8) (frame_differencing.py:80)."""
    return np.ones((k, k), np.uint8)


def structuring_ellipse(k: int) -> np.ndarray:
    """cv2.getStructuringElement(MORPH_ELLIPSE, (k, k)) (motion_compression_opt.py:62),
    restating OpenCV >= 4.x morph.dispatch.cpp: row i spans |dx| <= rint(c*sqrt(r^2-dy^2)/r)
    (here r = c since the element is square)."""
    r = k // 2
    c = k // 2
    inv_r2 = 1.0 / (r * r) if r else 0.0
    out = np.zeros((k, k), np.uint8)
    for i in range(k):
        dy = i - r
        if abs(dy) <= r:
            dx = int(np.rint(c * np.sqrt((r * r - dy * dy) * inv_r2)))
            j1, j2 = max(c - dx, 0), min(c + dx + 1, k)
            out[i, j1:j2] = 1
    return out


def _morph(img: np.ndarray, kernel: np.ndarray, is_dilate: bool) -> np.ndarray:
    kh, kw = kernel.shape
    ay, ax = kh // 2, kw // 2                      # default anchor (-1,-1) -> centre = k//2
    h, w = img.shape
    pad_val = 0 if is_dilate else 255               # morphologyDefaultBorderValue: ignored pixels
    top, left = ay, ax
    bottom, right = kh - 1 - ay, kw - 1 - ax
    p = np.pad(img, ((top, bottom), (left, right)), mode="constant", constant_values=pad_val)
    out = np.full((h, w), 0 if is_dilate else 255, np.uint8)
    for j in range(kh):
        for i in range(kw):
            if kernel[j, i]:
                win = p[j:j + h, i:i + w]
                out = np.maximum(out, win) if is_dilate else np.minimum(out, win)
    return out


def dilate(img: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """cv2.dilate(mask, kernel, iterations=1) (frame_differencing.py:106)."""
    return _morph(img, kernel, True)


def erode(img: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    return _morph(img, kernel, False)


def morph_close(img: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """cv2.morphologyEx(mask, MORPH_CLOSE, kernel) = erode(dilate(.)) (motion_compression_opt.py:89)."""
    return erode(dilate(img, kernel), kernel)


def morph_open(img: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """cv2.morphologyEx(mask, MORPH_OPEN, kernel) = dilate(erode(.)) (motion_compression_opt.py:90)."""
    return dilate(erode(img, kernel), kernel)


# ----------------------------------------------------------------------------
# temporal smoothing
# ----------------------------------------------------------------------------

def _fma32(a: np.ndarray, b: np.float32, c: np.ndarray) -> np.ndarray:
    # a*b is exact in float64 for a <= 255 and b float32; one rounding to float32.
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def add_weighted(a: np.ndarray, alpha: float, b: np.ndarray, beta: float) -> np.ndarray:
    """cv2.addWeighted(a, alpha, b, beta, 0) on uint8 (frame_differencing.py:107):
    float32 arithmetic, t = b*beta rounded, then fma(a, alpha, t), round half to even, saturate."""
    al, be = np.float32(alpha), np.float32(beta)
    t = (b.astype(np.float32) * be).astype(np.float32)
    s = _fma32(a.astype(np.float32), al, t)
    return _sat_u8(np.rint(s).astype(np.int64))


def window_min_counts(alpha_fraction: float, window_size: int) -> list[int]:
    """For L = 1..K, the smallest count c of 255-valued masks with
    ``c*255 >= alpha_fraction * L * 255`` evaluated exactly as the reference
    writes it in Python floats (motion_compression_opt.py:85-86)."""
    out = []
    for L in range(1, window_size + 1):
        rhs = alpha_fraction * L * 255
        c = 0
        while not (c * 255 >= rhs):
            c += 1
            if c > L:                       # vote can never pass
                break
        out.append(c)
    return out


def window_vote(masks: list[np.ndarray], alpha_fraction: float) -> np.ndarray:
    """np.sum over the deque + compare (motion_compression_opt.py:85-86); ``masks`` is the
    current deque content (at most window_size entries of 0/255 uint8)."""
    cumulative = np.sum(np.array(masks), axis=0)
    return (cumulative >= (alpha_fraction * len(masks) * 255)).astype(np.uint8) * 255


# ----------------------------------------------------------------------------
# contour filter (frame_differencing.py:100-104)
# ----------------------------------------------------------------------------

def contour_filter(mask: np.ndarray, min_area: float) -> np.ndarray:
    """findContours(RETR_EXTERNAL) -> keep contourArea > min_area -> drawContours(FILLED),
    restated without contour tracing:

    O = background pixels 4-connected to the outside of the image;  F = not O
    (foreground plus enclosed holes);  label F with 8-connectivity;  for each
    label, twice the polygon area through pixel centres is 2*Q4 + Q3, where
    Q4/Q3 count the 2x2 windows with 4/3 pixels of that label;  keep labels
    with area > min_area.
    """
    from scipy import ndimage
    fg = mask != 0
    h, w = fg.shape
    bg = np.pad(~fg, 1, mode="constant", constant_values=True)
    lab4, _ = ndimage.label(bg, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    outside = lab4 == lab4[0, 0]
    filled = ~outside[1:-1, 1:-1]
    lab8, n = ndimage.label(filled, structure=np.ones((3, 3), int))
    if n == 0:
        return np.zeros((h, w), np.uint8)
    f = filled.astype(np.int32)
    cnt = f[:-1, :-1] + f[:-1, 1:] + f[1:, :-1] + f[1:, 1:]
    # a 2x2 window with >=3 pixels of F lies in exactly one 8-component; take its label as the max.
    lab_win = np.maximum(np.maximum(lab8[:-1, :-1], lab8[:-1, 1:]), np.maximum(lab8[1:, :-1], lab8[1:, 1:]))
    twice_area = (np.bincount(lab_win[cnt == 4], minlength=n + 1) * 2
                  + np.bincount(lab_win[cnt == 3], minlength=n + 1))
    keep = twice_area > 2 * min_area
    keep[0] = False
    return np.where(keep[lab8], 255, 0).astype(np.uint8)


def contour_filter_cv2(mask: np.ndarray, min_area: float) -> np.ndarray:
    """The literal reference lines (frame_differencing.py:100-104)."""
    import cv2
    contours, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    out = np.zeros_like(mask)
    for c in contours:
        if cv2.contourArea(c) > min_area:
            cv2.drawContours(out, [c], -1, 255, thickness=cv2.FILLED)
    return out


def resize_linear(src: np.ndarray, dsize_wh) -> np.ndarray:
    """cv2.resize(src, (w, h)) with the default INTER_LINEAR on uint8 (frame_differencing.py:74,91), restated:
    OpenCV's two-pass fixed-point bilinear with 11-bit coefficients.  Columns clamp (index, fraction) at the border,
    rows keep the fraction and clip the row index; out = (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2."""
    dw, dh = int(dsize_wh[0]), int(dsize_wh[1])
    sh, sw = src.shape[:2]
    s = src.reshape(sh, sw, -1).astype(np.int64)

    def tables(dn, sn, clamp):
        scale = 1.0 / (dn / sn)
        d = np.arange(dn, dtype=np.float64)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        i = np.floor(f).astype(np.int64)
        f = (f - i.astype(np.float32)).astype(np.float32)
        if clamp:
            lo, hi = i < 0, i >= sn - 1
            i = np.where(lo, 0, np.where(hi, sn - 1, i))
            f = np.where(lo | hi, np.float32(0), f).astype(np.float32)
        a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
        a1 = np.rint(f * np.float32(2048)).astype(np.int64)
        return i, a0, a1

    xi, a0, a1 = tables(dw, sw, True)
    yi, b0, b1 = tables(dh, sh, False)
    x1 = np.minimum(xi + 1, sw - 1)
    rows = s[:, xi, :] * a0[None, :, None] + s[:, x1, :] * a1[None, :, None]
    r0, r1 = rows[np.clip(yi, 0, sh - 1)], rows[np.clip(yi + 1, 0, sh - 1)]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out.reshape((dh, dw) if src.ndim == 2 else (dh, dw, src.shape[2]))


def mask_rectangles(mask: np.ndarray) -> np.ndarray:
    """contours -> bounding rectangles (motion_compression_opt.py:93-97), restated without contour tracing:
    every 8-connected component is replaced by the rectangle columns min_x..max_x+1, rows min_y..max_y+1
    (cv2.rectangle includes the corner (x + w, y + h)), clipped to the image.  Components nested inside a hole of
    another one are not returned by RETR_EXTERNAL, but their rectangles lie inside the outer rectangle."""
    from scipy import ndimage
    h, w = mask.shape
    lab, n = ndimage.label(mask != 0, structure=np.ones((3, 3), int))
    out = np.zeros((h, w), np.uint8)
    for sl in ndimage.find_objects(lab):
        out[sl[0].start:min(h, sl[0].stop + 1), sl[1].start:min(w, sl[1].stop + 1)] = 255
    return out


def mask_rectangles_cv2(mask: np.ndarray) -> np.ndarray:
    """The literal reference lines (motion_compression_opt.py:93-97)."""
    import cv2
    contours, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    out = np.zeros_like(mask)
    for c in contours:
        x, y, w, h = cv2.boundingRect(c)
        cv2.rectangle(out, (x, y), (x + w, y + h), 255, -1)
    return out


# ----------------------------------------------------------------------------
# block DCT degrade (frame_differencing.py:117-127; motion_compression_opt.py:156-183)
# ----------------------------------------------------------------------------

def block_all_zero(mask: np.ndarray, bs: int, full_blocks_only: bool = False) -> np.ndarray:
    """``mask[y:y+bs, x:x+bs].mean() == 0`` per block (frame_differencing.py:120).  Returns a bool
    array [ceil(H/bs), ceil(W/bs)]; with ``full_blocks_only`` clipped edge blocks are False
    (motion_compression_opt.py:159 skips them)."""
    h, w = mask.shape
    nby, nbx = -(-h // bs), -(-w // bs)
    p = np.zeros((nby * bs, nbx * bs), bool)
    p[:h, :w] = mask != 0
    static = ~p.reshape(nby, bs, nbx, bs).any(axis=(1, 3))
    if full_blocks_only:
        if h % bs:
            static[-1, :] = False
        if w % bs:
            static[:, -1] = False
    return static


def _dct_rows(x: np.ndarray, inverse: bool) -> np.ndarray:
    import cv2
    flags = cv2.DCT_ROWS | (cv2.DCT_INVERSE if inverse else 0)
    return cv2.dct(np.ascontiguousarray(x, np.float32), flags=flags)


_BATCH_OK: dict = {}


def _batched_rows_is_exact(bs: int, inverse: bool) -> bool:
    """Does rows-pass + columns-pass through cv2.DCT_ROWS reproduce the per-block 2-D call bit for
    bit on THIS host?  (True for 4x4 here; False for 8x8, where IPP has a dedicated 2-D routine.)
    Decided by a one-off self-check so the oracle never silently drifts from the literal call."""
    import cv2
    key = (bs, inverse)
    if key not in _BATCH_OK:
        rng = np.random.default_rng(12345)
        blocks = rng.integers(-128, 128, (256, bs, bs)).astype(np.float32)
        if inverse:
            blocks = (np.round(blocks / 3) * 100).astype(np.float32)
        lit = np.stack([(cv2.idct if inverse else cv2.dct)(b) for b in blocks])
        _BATCH_OK[key] = bool(np.array_equal(lit, _dct2_blocks_batched(blocks, inverse)))
    return _BATCH_OK[key]


def _dct2_blocks_batched(blocks: np.ndarray, inverse: bool) -> np.ndarray:
    n, bh, bw = blocks.shape
    r = _dct_rows(blocks.reshape(n * bh, bw), inverse).reshape(n, bh, bw)
    rt = np.ascontiguousarray(r.transpose(0, 2, 1)).reshape(n * bw, bh)
    c = _dct_rows(rt, inverse).reshape(n, bw, bh).transpose(0, 2, 1)
    return np.ascontiguousarray(c)


def dct2_blocks(blocks: np.ndarray, inverse: bool = False) -> np.ndarray:
    """2-D DCT of a stack [N, bs, bs] float32, bit-identical to calling cv2.dct / cv2.idct on each
    block as the reference does (frame_differencing.py:122,124).  Uses one batched DCT_ROWS call per
    pass where that is verified to be bit-identical on this host, else the literal per-block calls."""
    import cv2
    blocks = np.ascontiguousarray(blocks, np.float32)
    n, bh, bw = blocks.shape
    if n == 0:
        return blocks.copy()
    if bh == bw and _batched_rows_is_exact(bh, inverse):
        return _dct2_blocks_batched(blocks, inverse)
    f = cv2.idct if inverse else cv2.dct
    return np.stack([f(b) for b in blocks])


def quantise_plane_blocks(plane: np.ndarray, static: np.ndarray, bs: int, q: float) -> np.ndarray:
    """For every full bs x bs block flagged static:
    ``clip(idct(round(dct(block - 128) / q) * q) + 128, 0, 255)`` stored to uint8 (truncation),
    frame_differencing.py:121-125.  Clipped edge blocks are handled by the callers."""
    h, w = plane.shape
    nby, nbx = h // bs, w // bs
    out = plane.copy()
    st = static[:nby, :nbx]
    if not st.any():
        return out
    view = plane[:nby * bs, :nbx * bs].reshape(nby, bs, nbx, bs).transpose(0, 2, 1, 3)
    blocks = view[st].astype(np.float32) - np.float32(128)
    d = dct2_blocks(blocks)
    qd = np.round(d / q) * q                      # numpy: float32 / python scalar stays float32
    r = dct2_blocks(qd.astype(np.float32), inverse=True) + np.float32(128)
    res = np.clip(r, 0, 255).astype(np.uint8)       # assignment into a uint8 array truncates
    oview = out[:nby * bs, :nbx * bs].reshape(nby, bs, nbx, bs).transpose(0, 2, 1, 3)
    oview[st] = res                                  # writes through: oview is a view of ``out``
    return out


def tie_blocks(plane: np.ndarray, static: np.ndarray, bs: int, q: float, eps: float = 2e-3) -> np.ndarray:
    """Blocks (bool [H//bs, W//bs]) having a DCT coefficient within ``eps`` of an exact quantiser tie
    (d/q = n + 1/2).  IPP's float32 rounding decides such ties, so no other float32 implementation
    can be expected to agree on them."""
    h, w = plane.shape
    nby, nbx = h // bs, w // bs
    view = plane[:nby * bs, :nbx * bs].reshape(nby, bs, nbx, bs).transpose(0, 2, 1, 3)
    blocks = view.reshape(-1, bs, bs).astype(np.float64) - 128.0
    from scipy.fft import dctn
    d = dctn(blocks, axes=(1, 2), norm="ortho")
    frac = np.abs(d / q - np.floor(d / q) - 0.5)
    tie = (frac < eps / q).any(axis=(1, 2)).reshape(nby, nbx)
    return tie & static[:nby, :nbx]


def degrade_fd(bgr: np.ndarray, acc_mask: np.ndarray, bs: int, q: float) -> np.ndarray:
    """frame_differencing.py:115-130 on one frame: BGR->YCrCb, static blocks (mask block all zero)
    get DCT-quantised luma and neutral chroma, YCrCb->BGR.  Requires H, W multiples of ``bs`` or
    even-sized clipped edge blocks (cv2.dct rejects odd sizes, as it does in the reference)."""
    import cv2
    ycc = bgr2ycrcb(bgr)
    h, w = acc_mask.shape
    static = block_all_zero(acc_mask, bs)
    y = quantise_plane_blocks(ycc[..., 0], static, bs, q)
    cr, cb = ycc[..., 1].copy(), ycc[..., 2].copy()
    nby, nbx = static.shape
    st_px = np.repeat(np.repeat(static, bs, axis=0), bs, axis=1)[:h, :w]
    cr[st_px] = 128
    cb[st_px] = 128
    # clipped edge blocks: literal per-block path, exactly as the reference slices them
    for by in range(nby):
        for bx in range(nbx):
            y0, x0 = by * bs, bx * bs
            full = (y0 + bs <= h) and (x0 + bs <= w)
            if full or not static[by, bx]:
                continue
            blk = ycc[y0:y0 + bs, x0:x0 + bs, 0]
            d = cv2.dct(blk.astype(np.float32) - 128)
            qd = np.round(d / q) * q
            r = cv2.idct(qd) + 128
            y[y0:y0 + bs, x0:x0 + bs] = np.clip(r, 0, 255)
    return ycrcb2bgr(np.stack([y, cr, cb], axis=-1))


def degrade_mco(bgr: np.ndarray, mask: np.ndarray, q: float = 100.0) -> np.ndarray:
    """motion_compression_opt.py:152-183 on one frame: 8x8 full blocks whose mask block is all zero get
    Y, Cr and Cb DCT-quantised, then after YCrCb->BGR the same blocks are re-grayed."""
    bs = 8
    ycc = bgr2ycrcb(bgr)
    static = block_all_zero(mask, bs, full_blocks_only=True)
    planes = [quantise_plane_blocks(ycc[..., c], static, bs, q) for c in range(3)]
    out = ycrcb2bgr(np.stack(planes, axis=-1))
    h, w = mask.shape
    st_px = np.repeat(np.repeat(static, bs, axis=0), bs, axis=1)[:h, :w]
    g = bgr2gray(out)
    out[st_px] = g[st_px][:, None]
    return out


def overlay_paint(bgr: np.ndarray, acc_mask: np.ndarray) -> np.ndarray:
    """``ov = frame.copy(); ov[acc > 127] = [0, 0, 255]`` (frame_differencing.py:110-111)."""
    ov = bgr.copy()
    ov[acc_mask > 127] = (0, 0, 255)
    return ov


# ----------------------------------------------------------------------------
# closed form of cv2's float32 4-point DCT (recovered by search)
# ----------------------------------------------------------------------------
_C1 = np.float32(np.cos(np.pi / 8) / np.sqrt(2.0))
_C3 = np.float32(np.cos(3 * np.pi / 8) / np.sqrt(2.0))
_H = np.float32(0.5)


def _fma(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64."""
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def dct4_rows_closed_form(x: np.ndarray, inverse: bool = False) -> np.ndarray:
    """The 1-D 4-point transform cv2.dct(..., DCT_ROWS) applies to each row of a float32 [N, 4] array, as
    an explicit float32 operation sequence.  The CUDA kernel (csrc/k_degrade.cuh dct4_fwd / dct4_inv) is the
    same sequence."""
    x = np.asarray(x, np.float32)
    x0, x1, x2, x3 = x[:, 0], x[:, 1], x[:, 2], x[:, 3]
    if not inverse:
        s0, s1, d0, d1 = x0 + x3, x1 + x2, x0 - x3, x1 - x2
        y0 = (s0 + s1) * _H
        y2 = (s0 - s1) * _H
        y1 = _fma(d1, _C3, (_C1 * d0).astype(np.float32))
        y3 = _fma(d0, _C3, -(_C1 * d1).astype(np.float32))
        return np.stack([y0, y1, y2, y3], axis=1).astype(np.float32)
    e0, e1 = (x0 + x2) * _H, (x0 - x2) * _H
    o0 = _fma(x3, _C3, (_C1 * x1).astype(np.float32))
    o1 = _fma(x1, _C3, -(_C1 * x3).astype(np.float32))
    return np.stack([e0 + o0, e1 + o1, e1 - o1, e0 - o0], axis=1).astype(np.float32)


def cv2_dct4_matches_closed_form(n: int = 20000) -> bool:
    """Does this host's cv2 (its IPP code path depends on the CPU) agree bit for bit with the closed form?"""
    import cv2
    rng = np.random.default_rng(777)
    x = (rng.standard_normal((n, 4)) * 60).astype(np.float32)
    xi = rng.integers(-128, 128, (n, 4)).astype(np.float32)
    ok = True
    for arr in (x, xi):
        ok &= np.array_equal(cv2.dct(arr, flags=cv2.DCT_ROWS), dct4_rows_closed_form(arr))
        ok &= np.array_equal(cv2.dct(arr, flags=cv2.DCT_ROWS | cv2.DCT_INVERSE), dct4_rows_closed_form(arr, True))
    return bool(ok)


# ----------------------------------------------------------------------------
# closed forms of cv2's other float32 DCT routines (recovered by search in round 2):
# the dedicated 2-D 8x8 routine and the 1-D routines of length 2, 6 and 8 that every other (clipped) block shape
# goes through, rows first, then columns, forward and inverse alike.  csrc/k_dct8.cuh is the same sequences in CUDA.
# ----------------------------------------------------------------------------
def _fma_any(a, b, c):
    """float32 fma for arrays or scalars (float32 products are exact in float64)."""
    return (np.asarray(a, np.float64) * np.float64(b) + np.asarray(c, np.float64)).astype(np.float32)


_A8 = np.array([[(np.sqrt(1 / 8) if k == 0 else 0.5) * np.cos(np.pi * (2 * n + 1) * k / 16) for n in range(8)]
                for k in range(8)]).astype(np.float32)                       # orthonormal 8-point DCT-II matrix
_C8 = [None] + [np.float32(0.5 * np.cos(j * np.pi / 16)) for j in range(1, 8)]   # 0.5 cos(j pi / 16)
_TG = [None] + [np.float32(np.tan(j * np.pi / 16)) for j in range(1, 4)]
_R2 = np.float32(np.sqrt(0.5))
_B8 = [None] + [np.float32(np.cos(j * np.pi / 16) / np.sqrt(2.0)) for j in range(1, 8)]
_ROWSCALE8 = np.array([_C8[4], _C8[1], _C8[2], _C8[3], _C8[4], _C8[3], _C8[2], _C8[1]], np.float32)


def _dct8x8_rows_fwd(x):
    v = [x[..., i] for i in range(8)]
    s = [v[i] + v[7 - i] for i in range(4)]
    d = [v[i] - v[7 - i] for i in range(4)]
    out = []
    for l in range(8):
        if l % 2 == 0:
            acc = _A8[l, 0] * s[0]
            for i in (1, 2, 3):
                acc = _fma_any(s[i], _A8[l, i], acc)
        else:
            acc = _A8[l, 3] * d[3]
            for i in (2, 1, 0):
                acc = _fma_any(d[i], _A8[l, i], acc)
        out.append(acc)
    return np.stack(out, -1)


def _dct8x8_cols_fwd(x):
    v = [x[..., i, :] for i in range(8)]
    t = [v[i] + v[7 - i] for i in range(4)]
    m = [v[i] - v[7 - i] for i in range(4)]
    tp03, tm03, tp12, tm12 = t[0] + t[3], t[0] - t[3], t[1] + t[2], t[1] - t[2]
    y = [None] * 8
    y[0] = _C8[4] * (tp03 + tp12)
    y[4] = _C8[4] * (tp03 - tp12)
    y[2] = _C8[2] * _fma_any(tm12, _TG[2], tm03)
    y[6] = _C8[2] * _fma_any(tm03, _TG[2], -tm12)
    tp65, tm65 = (m[1] + m[2]) * _R2, (m[1] - m[2]) * _R2
    tp765, tm765, tp465, tm465 = m[0] + tp65, m[0] - tp65, m[3] + tm65, m[3] - tm65
    y[1] = _C8[1] * _fma_any(tp465, _TG[1], tp765)
    y[7] = _C8[1] * _fma_any(tp765, _TG[1], -tp465)
    y[5] = _C8[3] * _fma_any(tm765, _TG[3], tm465)
    y[3] = _C8[3] * _fma_any(tm465, -_TG[3], tm765)
    return np.stack(y, -2)


def _dct8x8_rows_inv(w):
    u = [w[..., i] for i in range(8)]
    out = [None] * 8
    for c in range(4):
        e = _A8[0, c] * u[0]
        o = _A8[1, c] * u[1]
        for l in (2, 4, 6):
            e = _fma_any(u[l], _A8[l, c], e)
            o = _fma_any(u[l + 1], _A8[l + 1, c], o)
        out[c], out[7 - c] = e + o, e - o
    return np.stack(out, -1)


def _dct8x8_cols_inv(z):
    x = [z[..., k, :] for k in range(8)]
    tp765, tp465 = _fma_any(x[7], _TG[1], x[1]), _fma_any(x[1], _TG[1], -x[7])
    tm765, tm465 = _fma_any(x[5], _TG[3], x[3]), _fma_any(x[3], -_TG[3], x[5])
    tm03, tm12 = _fma_any(x[6], _TG[2], x[2]), _fma_any(x[2], _TG[2], -x[6])
    t7, tp65, t4, tm65 = tp765 + tm765, tp765 - tm765, tp465 + tm465, tp465 - tm465
    p65, m65 = tp65 * _R2, tm65 * _R2
    t6, t5 = p65 + m65, p65 - m65
    tp03, tp12 = x[0] + x[4], x[0] - x[4]
    t0, t3, t1, t2 = tp03 + tm03, tp03 - tm03, tp12 + tm12, tp12 - tm12
    return np.stack([t0 + t7, t1 + t6, t2 + t5, t3 + t4, t3 - t4, t2 - t5, t1 - t6, t0 - t7], -2)


def dct8x8_closed_form(blocks: np.ndarray, inverse: bool = False) -> np.ndarray:
    """cv2.dct / cv2.idct of float32 [..., 8, 8] blocks as an explicit float32 operation sequence: forward = rows as
    even/odd FMA chains, then columns through a tangent-rotation butterfly scaled last; inverse = row k scaled by
    0.5 cos(k pi / 16), rows as FMA chains, then the column butterfly."""
    b = np.asarray(blocks, np.float32)
    if not inverse:
        return _dct8x8_cols_fwd(_dct8x8_rows_fwd(b))
    return _dct8x8_cols_inv(_dct8x8_rows_inv(b * _ROWSCALE8[:, None]))


def _dct1d_generic(v, n: int, inverse: bool):
    """Lengths 3, 5, 6, 7: fold about the middle; even / odd outputs as FMA chains over the unnormalised cosines
    float32(cos(pi (2 i + 1) k / (2 n))) (odd lengths start the even chains with the middle sample); the scale
    sqrt(2/n) (sqrt(1/n) for k = 0) is applied last (forward) or first (inverse)."""
    m = np.array([[np.cos(np.pi * (2 * i + 1) * k / (2 * n)) for i in range(n)] for k in range(n)]).astype(np.float32)
    k1, k0 = np.float32(np.sqrt(2 / n)), np.float32(np.sqrt(1 / n))
    h, odd = n // 2, n % 2
    if not inverse:
        s = [v[i] + v[n - 1 - i] for i in range(h)]
        d = [v[i] - v[n - 1 - i] for i in range(h)]
        out = []
        for k in range(n):
            if k % 2:
                terms = [(d[i], m[k, i]) for i in range(h)]
            else:
                terms = ([(v[h], m[k, h])] if odd else []) + [(s[i], m[k, i]) for i in range(h)]
            acc = terms[0][1] * terms[0][0]
            for u, c in terms[1:]:
                acc = _fma_any(u, c, acc)
            out.append(acc * (k0 if k == 0 else k1))
        return np.stack(out, -1)
    w = [v[k] * (k0 if k == 0 else k1) for k in range(n)]
    out = [None] * n
    for i in range(h):
        e = m[0, i] * w[0]
        for k in range(2, n, 2):
            e = _fma_any(w[k], m[k, i], e)
        o = m[1, i] * w[1]
        for k in range(3, n, 2):
            o = _fma_any(w[k], m[k, i], o)
        out[i], out[n - 1 - i] = e + o, e - o
    if odd:
        p, q = w[0], w[2]
        if n > 4:
            p = p + w[4]
        if n > 6:
            q = q + w[6]
        out[h] = p - q
    return np.stack(out, -1)


def dct1d_closed_form(x: np.ndarray, inverse: bool = False) -> np.ndarray:
    """The 1-D transform cv2.dct(..., DCT_ROWS) applies along the last axis (length 1..8)."""
    x = np.asarray(x, np.float32)
    n = x.shape[-1]
    v = [x[..., i] for i in range(n)]
    h = np.float32(0.5)
    if n == 1:
        return x.copy()
    if n == 2:
        a, b = v[0] * _R2, v[1] * _R2
        return np.stack([a + b, a - b], -1)
    if n == 4:
        return dct4_rows_closed_form(x.reshape(-1, 4), inverse).reshape(x.shape)
    if n in (3, 5, 6, 7):
        return _dct1d_generic(v, n, inverse)
    if n == 8:
        c0, c2, c6 = _C8[4], _C8[2], _C8[6]
        if not inverse:
            s = [v[i] + v[7 - i] for i in range(4)]
            d = [v[i] - v[7 - i] for i in range(4)]
            e0, e1, f0, f1 = s[0] + s[3], s[1] + s[2], s[0] - s[3], s[1] - s[2]
            u0, u3 = d[0] * _R2, d[3] * _R2
            p65, m65 = (d[1] + d[2]) * h, (d[1] - d[2]) * h
            tp765, tm765, tp465, tm465 = u0 + p65, u0 - p65, u3 + m65, u3 - m65
            y = [None] * 8
            y[0], y[4] = c0 * (e0 + e1), c0 * (e0 - e1)
            y[2] = _fma_any(f0, c2, c6 * f1)
            y[6] = _fma_any(f0, c6, -(c2 * f1))
            y[1] = _fma_any(tp465, _B8[7], _B8[1] * tp765)
            y[7] = _fma_any(tp765, _B8[7], -(_B8[1] * tp465))
            y[5] = _fma_any(tm465, _B8[3], _B8[5] * tm765)
            y[3] = _fma_any(tm765, _B8[3], -(_B8[5] * tm465))
            return np.stack(y, -1)
        a0, a4 = c0 * v[0], c0 * v[4]
        ap, am = a0 + a4, a0 - a4
        b0 = _fma_any(v[6], c6, c2 * v[2])
        b1 = _fma_any(v[2], c6, -(c2 * v[6]))
        e = [ap + b0, am + b1, am - b1, ap - b0]
        tp765 = _fma_any(v[7], _B8[7], _B8[1] * v[1])
        tp465 = _fma_any(v[1], _B8[7], -(_B8[1] * v[7]))
        tm765 = _fma_any(v[3], _B8[3], _B8[5] * v[5])
        tm465 = _fma_any(v[5], _B8[3], -(_B8[5] * v[3]))
        p65, m65 = (tp765 - tm765) * h, (tp465 - tm465) * h
        o = [_R2 * (tp765 + tm765), p65 + m65, p65 - m65, _R2 * (tp465 + tm465)]
        out = [None] * 8
        for i in range(4):
            out[i], out[7 - i] = e[i] + o[i], e[i] - o[i]
        return np.stack(out, -1)
    raise ValueError(f"no closed form for length {n}")


def dct_block_closed_form(x: np.ndarray, inverse: bool = False) -> np.ndarray:
    """cv2.dct / cv2.idct of float32 [..., h, w] blocks (h, w in 1..8) as explicit float32 sequences."""
    x = np.asarray(x, np.float32)
    h, w = x.shape[-2:]
    if (h, w) == (8, 8):
        return dct8x8_closed_form(x, inverse)
    r = dct1d_closed_form(x, inverse)
    c = dct1d_closed_form(np.swapaxes(r, -1, -2), inverse)
    return np.ascontiguousarray(np.swapaxes(c, -1, -2))


_CF_OK: dict = {}


def cv2_dct_matches_closed_form(h: int = 8, w: int = 8, n: int = 4000) -> bool:
    """Does this host's cv2 (its IPP code path depends on the CPU) agree bit for bit with the closed forms for
    h x w blocks, forward and inverse, on integer, quantised and generic inputs?"""
    import cv2
    key = (h, w)
    if key not in _CF_OK:
        rng = np.random.default_rng(4242 + 16 * h + w)
        xi = rng.integers(-128, 128, (n, h, w)).astype(np.float32)
        xg = (rng.standard_normal((n, h, w)) * 80).astype(np.float32)
        ok = True
        for arr in (xi, xg):
            f = np.stack([cv2.dct(b) for b in arr])
            ok &= np.array_equal(f, dct_block_closed_form(arr))
            for q in (100.0, 3.0):
                qd = (np.round(f / np.float32(q)) * np.float32(q)).astype(np.float32)
                ok &= np.array_equal(np.stack([cv2.idct(b) for b in qd]), dct_block_closed_form(qd, True))
        _CF_OK[key] = bool(ok)
    return _CF_OK[key]
