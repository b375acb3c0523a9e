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

// 16 BGR pixels (12 words) -> 16 gray bytes (4 words)
DEVI void gray16(const uint32_t (&w)[12], uint32_t (&g)[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t acc = 0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            int px = q * 4 + p;
            acc |= gray_of(byte_at(w, 3 * px), byte_at(w, 3 * px + 1), byte_at(w, 3 * px + 2)) << (8 * p);
        }
        g[q] = acc;
    }
}

// per-byte |a - b| > thr  ->  4 mask bits (bit p = byte p)
DEVI uint32_t diff_gt_bits4(uint32_t a, uint32_t b, uint32_t thr) {
    uint32_t d = __vabsdiffu4(a, b);
    uint32_t bits = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) bits |= (((d >> (8 * p)) & 0xffu) > thr ? 1u : 0u) << p;
    return bits;
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
