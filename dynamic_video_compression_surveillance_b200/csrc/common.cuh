// Shared device helpers and the bit-plane layout used by every mask kernel.
//
// Masks on this path are binary by construction (cv2.threshold output, frame_differencing.py:97),
// so between kernels they live as bit-planes: one bit per pixel, LSB-first inside little-endian
// 32-bit words, `wpr` words per row (row pitch padded to 16 bytes so rows and row bands can be moved
// with cp.async.bulk).  Bits at x >= W are always zero.  A 1080p plane is 259 KB: every mask stage
// (window vote, morphology, contour filter, EMA flags) works out of L2 / shared memory and the HBM
// traffic of the loop is the BGR frames themselves.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DEVI __device__ __forceinline__

namespace dvc {

__host__ __device__ inline int words_per_row(int W) { return (((W + 31) / 32) + 3) & ~3; }

// valid-bit mask of word j of a row of width W
DEVI uint32_t valid_mask(int j, int W) {
    int rem = W - j * 32;
    return rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}

// byte i (compile-time constant after unrolling) of a packed little-endian word array
template <int N>
DEVI uint32_t byte_at(const uint32_t (&w)[N], int i) {
    return (w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
}

// cv2.cvtColor(BGR2GRAY) on uint8: 15-bit fixed point, round half up
// (frame_differencing.py:75,92; motion_compression_opt.py:60,71,149,181)
DEVI uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r) {
    return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

// zero-extended byte k of a word: one PRMT
DEVI uint32_t bytek(uint32_t w, int k) { return __byte_perm(w, 0u, 0x4440u | (uint32_t)k); }

// 2 * (3735 B + 19235 G + 9798 R + 16384): the gray value is then byte 2 of the sum, so no shift is needed
DEVI uint32_t gray2x(uint32_t b, uint32_t g, uint32_t r) { return 7470u * b + 38470u * g + 19596u * r + 32768u; }

// 16 BGR pixels (12 words) -> 16 gray bytes (4 words).  Per 4 pixels: 12 byte extractions (PRMT / shift),
// 12 IMAD, 3 PRMT to pack the four result bytes.
DEVI void gray16(const uint32_t (&w)[12], uint32_t (&g)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        const uint32_t s0 = gray2x(bytek(a, 0), bytek(a, 1), bytek(a, 2));
        const uint32_t s1 = gray2x(a >> 24, bytek(b, 0), bytek(b, 1));
        const uint32_t s2 = gray2x(bytek(b, 2), b >> 24, bytek(c, 0));
        const uint32_t s3 = gray2x(bytek(c, 1), bytek(c, 2), c >> 24);
        const uint32_t lo = __byte_perm(s0, s1, 0x0062u), hi = __byte_perm(s2, s3, 0x0062u);   // byte 2 of each
        g[q] = __byte_perm(lo, hi, 0x5410u);
    }
}

// Same result through IDP.4A: each coefficient of 2*(3735, 19235, 9798) is split into a high and a low byte,
// s = 256 * dot(px, hi) + dot(px, lo) + 32768; pixels that straddle two words chain two dot products.
DEVI void gray16_dp4a(const uint32_t (&w)[12], uint32_t (&g)[4]) {
    // little-endian coefficient words for a pixel whose B byte sits at byte 0 / 3 / 2 / 1 of its first word
    const uint32_t L0 = 46u | (70u << 8) | (140u << 16), H0 = 29u | (150u << 8) | (76u << 16);          // B G R .
    const uint32_t L1a = 46u << 24, H1a = 29u << 24, L1b = 70u | (140u << 8), H1b = 150u | (76u << 8);   // ...B | G R
    const uint32_t L2a = (46u << 16) | (70u << 24), H2a = (29u << 16) | (150u << 24), L2b = 140u, H2b = 76u;   // ..BG | R
    const uint32_t L3 = (46u << 8) | (70u << 16) | (140u << 24), H3 = (29u << 8) | (150u << 16) | (76u << 24); // .BGR
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        const uint32_t s0 = (__dp4a(a, H0, 0u) << 8) + __dp4a(a, L0, 32768u);
        const uint32_t s1 = (__dp4a(b, H1b, __dp4a(a, H1a, 0u)) << 8) + __dp4a(b, L1b, __dp4a(a, L1a, 32768u));
        const uint32_t s2 = (__dp4a(c, H2b, __dp4a(b, H2a, 0u)) << 8) + __dp4a(c, L2b, __dp4a(b, L2a, 32768u));
        const uint32_t s3 = (__dp4a(c, H3, 0u) << 8) + __dp4a(c, L3, 32768u);
        const uint32_t lo = __byte_perm(s0, s1, 0x0062u), hi = __byte_perm(s2, s3, 0x0062u);
        g[q] = __byte_perm(lo, hi, 0x5410u);
    }
}

// Same result through IDP.2A (u16 x u8 pairs): the doubled coefficients 7470 / 38470 / 19596 fit 16 bits, so a pixel
// is two dot products (three source bytes, one of the four products multiplies by zero) and the gray value is byte 2
// of the sum.  dp2a.lo: a.lo16 * b.byte0 + a.hi16 * b.byte1; dp2a.hi: a.lo16 * b.byte2 + a.hi16 * b.byte3.
DEVI void gray16_dp2a(const uint32_t (&w)[12], uint32_t (&g)[4]) {
    const uint32_t cBG = 7470u | (38470u << 16), cR_ = 19596u, c_B = 7470u << 16, cGR = 38470u | (19596u << 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        const uint32_t s0 = __dp2a_hi(cR_, a, __dp2a_lo(cBG, a, 32768u));      // B G R .
        const uint32_t s1 = __dp2a_lo(cGR, b, __dp2a_hi(c_B, a, 32768u));      // . . . B | G R
        const uint32_t s2 = __dp2a_lo(cR_, c, __dp2a_hi(cBG, b, 32768u));      // . . B G | R
        const uint32_t s3 = __dp2a_hi(cGR, c, __dp2a_lo(c_B, c, 32768u));      // . B G R
        const uint32_t lo = __byte_perm(s0, s1, 0x0062u), hi = __byte_perm(s2, s3, 0x0062u);
        g[q] = __byte_perm(lo, hi, 0x5410u);
    }
}

// per-byte |a - b| > thr  ->  0x80 in that byte.  thr < 128 uses a SWAR compare.
DEVI uint32_t diff_gt_msb4(uint32_t a, uint32_t b, uint32_t thr) {
    if (thr == 0u) {
        // the reference's default (motion_threshold 0.5 -> diff > 0): a byte differs iff its XOR is non-zero
        const uint32_t x = a ^ b;
        return (x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;
    }
    const uint32_t d = __vabsdiffu4(a, b);
    if (thr < 128u) {
        const uint32_t k7 = (0x7fu - thr) * 0x01010101u;
        return (((d & 0x7f7f7f7fu) + k7) | d) & 0x80808080u;                            // bit 8p+7 set iff byte p > thr
    }
    uint32_t m = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) m |= (((d >> (8 * p)) & 0xffu) > thr ? 0x80u : 0u) << (8 * p);
    return m;
}
// the flag bytes of 8 pixels (two words of diff_gt_msb4) gathered by two dot products: 128 * (bit p = pixel p)
DEVI uint32_t gather8_x128(uint32_t m0, uint32_t m1) { return __dp4a(m1, 0x80402010u, __dp4a(m0, 0x08040201u, 0u)); }
// per-byte |a - b| > thr  ->  4 mask bits (bit p = byte p)
DEVI uint32_t diff_gt_bits4(uint32_t a, uint32_t b, uint32_t thr) { return __dp4a(diff_gt_msb4(a, b, thr), 0x08040201u, 0u) >> 7; }
// 16 pixels -> 16 mask bits
DEVI uint32_t diff_gt_bits16(const uint32_t (&a)[4], const uint32_t (&b)[4], uint32_t thr) {
    const uint32_t lo = gather8_x128(diff_gt_msb4(a[0], b[0], thr), diff_gt_msb4(a[1], b[1], thr));
    const uint32_t hi = gather8_x128(diff_gt_msb4(a[2], b[2], thr), diff_gt_msb4(a[3], b[3], thr));
    return (lo >> 7) | (hi << 1);
}

DEVI void load16(const uint8_t* p, uint32_t (&v)[4]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
DEVI void store16(uint8_t* p, const uint32_t (&v)[4]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]);
}
// streaming (evict-first) variants for data touched once
DEVI void load16_cs(const uint8_t* p, uint32_t* v) {
    uint4 t = __ldcs(reinterpret_cast<const uint4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
DEVI void store16_cs(uint8_t* p, const uint32_t* v) {
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(v[0], v[1], v[2], v[3]));
}

}  // namespace dvc
