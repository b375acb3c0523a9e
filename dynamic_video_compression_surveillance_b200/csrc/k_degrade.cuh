// K4: overlay paint + BGR->YCrCb + block-DCT degrade of static blocks + YCrCb->BGR, and the
// statistics, in one pass over the frame (frame_differencing.py:110-111,115-130;
// motion_compression_opt.py:152-183).
//
// Inputs are the BGR frame and two bit-planes of the accumulated mask: over127 (acc > 127, the
// overlay test, frame_differencing.py:111) and nonzero (acc != 0; a block is static when its
// nonzero bits are all clear, which is what `.mean() == 0` says, :120).
//
// Arithmetic contract (re-verified in tests/test_oracle_vs_cv2.py):
//   BGR->YCrCb   Y  = (1868 B + 9617 G + 4899 R + 8192) >> 14
//                Cr = sat(((R - Y) * 11682 + (128 << 14) + 8192) >> 14),  Cb likewise with B, 9241
//   YCrCb->BGR   B = sat(Y + ((29049 Cb' + 8192) >> 14)),  R = sat(Y + ((22987 Cr' + 8192) >> 14)),
//                G = sat(Y + ((-5636 Cb' - 11698 Cr' + 8192) >> 14)),  ' = minus 128
//   static block Y' = trunc(clip(idct(rint(dct(Y - 128) / q) * q) + 128, 0, 255)), chroma 128 => B=G=R=Y'
// The 4-point DCT below reproduces cv2.dct / cv2.idct (IPP build, float32) bit for bit: the operation
// order and the two FMAs were recovered by search against cv2 on 200 000 random vectors.
#pragma once
#include "common.cuh"

namespace dvc {

// 0.5*cos(pi/8)*sqrt(2) ... : orthonormal 4-point DCT-II coefficients, correctly rounded to float32
#define DVC_C1 0x1.4e7aeap-1f   /* cos(pi/8)  / sqrt(2) = 0.6532815 */
#define DVC_C3 0x1.1517a8p-2f   /* cos(3pi/8) / sqrt(2) = 0.2705981 */

DEVI void dct4_fwd(float& x0, float& x1, float& x2, float& x3) {
    const float s0 = __fadd_rn(x0, x3), s1 = __fadd_rn(x1, x2);
    const float d0 = __fsub_rn(x0, x3), d1 = __fsub_rn(x1, x2);
    x0 = __fmul_rn(__fadd_rn(s0, s1), 0.5f);
    x2 = __fmul_rn(__fsub_rn(s0, s1), 0.5f);
    x1 = __fmaf_rn(DVC_C3, d1, __fmul_rn(DVC_C1, d0));
    x3 = __fmaf_rn(DVC_C3, d0, -__fmul_rn(DVC_C1, d1));
}
DEVI void dct4_inv(float& x0, float& x1, float& x2, float& x3) {
    const float e0 = __fmul_rn(__fadd_rn(x0, x2), 0.5f), e1 = __fmul_rn(__fsub_rn(x0, x2), 0.5f);
    const float o0 = __fmaf_rn(DVC_C3, x3, __fmul_rn(DVC_C1, x1));
    const float o1 = __fmaf_rn(DVC_C3, x1, -__fmul_rn(DVC_C1, x3));
    x0 = __fadd_rn(e0, o0);
    x3 = __fsub_rn(e0, o0);
    x1 = __fadd_rn(e1, o1);
    x2 = __fsub_rn(e1, o1);
}

// np.round(d / q) * q in float32: IEEE division, round half to even, exact product
DEVI float quantise(float d, float q) { return __fmul_rn(rintf(__fdiv_rn(d, q)), q); }

DEVI uint32_t clip_trunc_u8(float v) {   // np.clip(v, 0, 255) stored into a uint8 array
    return (uint32_t)__float2int_rz(fminf(fmaxf(v, 0.0f), 255.0f));
}

DEVI int luma_of(int b, int g, int r) { return (1868 * b + 9617 * g + 4899 * r + 8192) >> 14; }
DEVI int sat8(int v) { return min(255, max(0, v)); }

// BGR -> YCrCb -> BGR of one pixel (the non-static path: pure integer round trip)
DEVI void ycc_roundtrip(int& b, int& g, int& r) {
    const int y = luma_of(b, g, r);
    const int cr = sat8(((r - y) * 11682 + (128 << 14) + 8192) >> 14) - 128;
    const int cb = sat8(((b - y) * 9241 + (128 << 14) + 8192) >> 14) - 128;
    b = sat8(y + ((29049 * cb + 8192) >> 14));
    g = sat8(y + ((-5636 * cb - 11698 * cr + 8192) >> 14));
    r = sat8(y + ((22987 * cr + 8192) >> 14));
}

// 4x4 block, full 2-D transform pair on luma (rows first, then columns, as cv2 does)
DEVI void degrade_block4(float (&v)[4][4], float q) {
#pragma unroll
    for (int r = 0; r < 4; ++r) dct4_fwd(v[r][0], v[r][1], v[r][2], v[r][3]);
#pragma unroll
    for (int c = 0; c < 4; ++c) dct4_fwd(v[0][c], v[1][c], v[2][c], v[3][c]);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) v[r][c] = quantise(v[r][c], q);
#pragma unroll
    for (int r = 0; r < 4; ++r) dct4_inv(v[r][0], v[r][1], v[r][2], v[r][3]);
#pragma unroll
    for (int c = 0; c < 4; ++c) dct4_inv(v[0][c], v[1][c], v[2][c], v[3][c]);
}

struct Counters { unsigned long long frames, pixels, motion_pixels, blocks, static_blocks; };

__global__ void k_counters_add(Counters* c, unsigned long long frames, unsigned long long pixels, unsigned long long blocks) {
    atomicAdd(&c->frames, frames);
    atomicAdd(&c->pixels, pixels);
    atomicAdd(&c->blocks, blocks);
}

// ------------------------------------------------------------------------------------------------
// Fast path: block_size 4, W % 16 == 0, H % 4 == 0.  One thread owns 16 pixels x 4 rows = four 4x4
// blocks = 4 x 48 contiguous bytes: twelve 16-byte loads in flight per thread, twelve (or twenty-four
// with the overlay) 16-byte streaming stores.  Algorithmic HBM bytes: 3 read + 3 (+3) written per pixel
// plus 2/8 of mask bits.
// grid: (ceil(W/16 * H/4 / 128), n_frames)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_degrade4(const uint8_t* __restrict__ frames, const uint32_t* __restrict__ over127,
           const uint32_t* __restrict__ nonzero, uint8_t* __restrict__ compressed, uint8_t* __restrict__ overlay,
           int H, int W, int wpr, float q, Counters* __restrict__ counters) {
    const int gpr = W >> 4, nbr = H >> 2;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = gid < gpr * nbr;
    unsigned n_motion = 0, n_static = 0;
    if (active) {
        const int br = gid / gpr, gx = gid - br * gpr;
        const size_t frame_off = (size_t)blockIdx.y * H * W * 3;
        const size_t plane_off = (size_t)blockIdx.y * H * wpr;
        const size_t base = frame_off + ((size_t)(br * 4) * W + (size_t)gx * 16) * 3;
        const size_t pitch = (size_t)W * 3;
        uint32_t w[4][12];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            load16_cs(frames + base + r * pitch, &w[r][0]);
            load16_cs(frames + base + r * pitch + 16, &w[r][4]);
            load16_cs(frames + base + r * pitch + 32, &w[r][8]);
        }
        uint32_t hi[4], nz = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const size_t wo = plane_off + (size_t)(br * 4 + r) * wpr;
            hi[r] = reinterpret_cast<const uint16_t*>(over127 + wo)[gx];
            nz |= reinterpret_cast<const uint16_t*>(nonzero + wo)[gx];
        }
        // ---- overlay: paint (B,G,R) = (0,0,255) where acc > 127 ----
        if (overlay) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                uint32_t o[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) o[i] = w[r][i];
                if (hi[r]) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const uint32_t m = hi[r] >> (4 * b);
                        if (m & 1u) { o[3 * b] = (o[3 * b] & 0xff000000u) | 0x00ff0000u; }
                        if (m & 2u) { o[3 * b] &= 0x00ffffffu; o[3 * b + 1] = (o[3 * b + 1] & 0xffff0000u) | 0x0000ff00u; }
                        if (m & 4u) { o[3 * b + 1] &= 0x0000ffffu; o[3 * b + 2] = (o[3 * b + 2] & 0xffffff00u) | 0x000000ffu; }
                        if (m & 8u) { o[3 * b + 2] = (o[3 * b + 2] & 0x000000ffu) | 0xff000000u; }
                    }
                    n_motion += __popc(hi[r]);
                }
                store16_cs(overlay + base + r * pitch, &o[0]);
                store16_cs(overlay + base + r * pitch + 16, &o[4]);
                store16_cs(overlay + base + r * pitch + 32, &o[8]);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) n_motion += __popc(hi[r]);
        }
        // ---- compressed: per 4x4 block ----
        if (compressed) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const bool is_static = ((nz >> (4 * b)) & 0xfu) == 0u;
                if (is_static) {
                    ++n_static;
                    float v[4][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const uint32_t (&ww)[12] = w[r];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const int i = 12 * b + 3 * p;
                            v[r][p] = (float)(luma_of(byte_at(ww, i), byte_at(ww, i + 1), byte_at(ww, i + 2)) - 128);
                        }
                    }
                    degrade_block4(v, q);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const uint32_t y0 = clip_trunc_u8(__fadd_rn(v[r][0], 128.0f));
                        const uint32_t y1 = clip_trunc_u8(__fadd_rn(v[r][1], 128.0f));
                        const uint32_t y2 = clip_trunc_u8(__fadd_rn(v[r][2], 128.0f));
                        const uint32_t y3 = clip_trunc_u8(__fadd_rn(v[r][3], 128.0f));
                        w[r][3 * b] = y0 * 0x00010101u | (y1 << 24);
                        w[r][3 * b + 1] = y1 * 0x00000101u | (y2 * 0x01010000u);
                        w[r][3 * b + 2] = y2 | (y3 * 0x01010100u);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        uint32_t o[3] = {0, 0, 0};
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const int i = 12 * b + 3 * p;
                            int bb = byte_at(w[r], i), gg = byte_at(w[r], i + 1), rr = byte_at(w[r], i + 2);
                            ycc_roundtrip(bb, gg, rr);
                            const int j = 3 * p;
                            o[j >> 2] |= (uint32_t)bb << ((j & 3) * 8);
                            o[(j + 1) >> 2] |= (uint32_t)gg << (((j + 1) & 3) * 8);
                            o[(j + 2) >> 2] |= (uint32_t)rr << (((j + 2) & 3) * 8);
                        }
                        w[r][3 * b] = o[0]; w[r][3 * b + 1] = o[1]; w[r][3 * b + 2] = o[2];
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                store16_cs(compressed + base + r * pitch, &w[r][0]);
                store16_cs(compressed + base + r * pitch + 16, &w[r][4]);
                store16_cs(compressed + base + r * pitch + 32, &w[r][8]);
            }
        } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) n_static += ((nz >> (4 * b)) & 0xfu) == 0u ? 1u : 0u;
        }
    }
    // ---- statistics: warp shuffle reduction, one atomic pair per warp ----
    if (counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_motion += __shfl_xor_sync(0xffffffffu, n_motion, o);
            n_static += __shfl_xor_sync(0xffffffffu, n_static, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (n_motion) atomicAdd(&counters->motion_pixels, (unsigned long long)n_motion);
            if (n_static) atomicAdd(&counters->static_blocks, (unsigned long long)n_static);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// General path: block_size 4 or 8, any W, H that are multiples of block_size, both flavours.
// One thread per block, bytewise access.  Used for block_size 8 (frame_differencing.py:203 main
// config), for the MCO flavour (motion_compression_opt.py:152-183) and for widths the fast path
// cannot take.  The 8-point DCT is the orthonormal DCT-II in even/odd matrix form in float32; cv2's
// 8x8 routine (IPP) is not reproduced bit for bit -- results agree within 1 grey level away from
// exact quantiser ties.
// ------------------------------------------------------------------------------------------------
__constant__ float c_dct8[8][4];   // c_dct8[k][n] = s(k) cos(pi (2n+1) k / 16), n < 4

DEVI void dct8_fwd(float (&x)[8]) {
    float a[4], b[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) { a[n] = x[n] + x[7 - n]; b[n] = x[n] - x[7 - n]; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float (&s)[4] = (k & 1) ? b : a;
        float acc = 0.0f;
#pragma unroll
        for (int n = 0; n < 4; ++n) acc = fmaf(c_dct8[k][n], s[n], acc);
        x[k] = acc;
    }
}
DEVI void dct8_inv(float (&x)[8]) {
    float e[4], o[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        float ae = 0.0f, ao = 0.0f;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            ae = fmaf(c_dct8[2 * m][n], x[2 * m], ae);
            ao = fmaf(c_dct8[2 * m + 1][n], x[2 * m + 1], ao);
        }
        e[n] = ae; o[n] = ao;
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) { x[n] = e[n] + o[n]; x[7 - n] = e[n] - o[n]; }
}

template <int BS>
DEVI void degrade_plane(float (&v)[BS][BS], float q) {
    if constexpr (BS == 4) {
        degrade_block4(v, q);
    } else {
        float t[8];
#pragma unroll
        for (int r = 0; r < BS; ++r) {
#pragma unroll
            for (int c = 0; c < 8; ++c) t[c] = v[r][c];
            dct8_fwd(t);
#pragma unroll
            for (int c = 0; c < 8; ++c) v[r][c] = t[c];
        }
#pragma unroll
        for (int c = 0; c < BS; ++c) {
#pragma unroll
            for (int r = 0; r < 8; ++r) t[r] = v[r][c];
            dct8_fwd(t);
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r][c] = quantise(t[r], q);
        }
#pragma unroll
        for (int r = 0; r < BS; ++r) {
#pragma unroll
            for (int c = 0; c < 8; ++c) t[c] = v[r][c];
            dct8_inv(t);
#pragma unroll
            for (int c = 0; c < 8; ++c) v[r][c] = t[c];
        }
#pragma unroll
        for (int c = 0; c < BS; ++c) {
#pragma unroll
            for (int r = 0; r < 8; ++r) t[r] = v[r][c];
            dct8_inv(t);
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r][c] = t[r];
        }
    }
}

template <int BS, int FLAVOUR>   // FLAVOUR 0 = FD, 1 = MCO
__global__ void __launch_bounds__(128)
k_degrade_generic(const uint8_t* __restrict__ frames, const uint32_t* __restrict__ over127,
                  const uint32_t* __restrict__ nonzero, uint8_t* __restrict__ compressed,
                  uint8_t* __restrict__ overlay, int H, int W, int wpr, float q, Counters* __restrict__ counters) {
    const int nbx = W / BS, nby = H / BS;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned n_motion = 0, n_static = 0;
    if (gid < nbx * nby) {
        const int by = gid / nbx, bx = gid - by * nbx;
        const int x0 = bx * BS, y0 = by * BS;
        const uint8_t* fr = frames + (size_t)blockIdx.y * H * W * 3;
        const size_t plane_off = (size_t)blockIdx.y * H * wpr;
        uint32_t nz = 0;
        for (int r = 0; r < BS; ++r) {
            const size_t wo = plane_off + (size_t)(y0 + r) * wpr;
            const uint32_t hb = (over127[wo + (x0 >> 5)] >> (x0 & 31)) & ((1u << BS) - 1u);
            nz |= (nonzero[wo + (x0 >> 5)] >> (x0 & 31)) & ((1u << BS) - 1u);
            n_motion += __popc(hb);
            if (overlay) {
                uint8_t* orow = overlay + (size_t)blockIdx.y * H * W * 3 + ((size_t)(y0 + r) * W + x0) * 3;
                const uint8_t* irow = fr + ((size_t)(y0 + r) * W + x0) * 3;
                for (int c = 0; c < BS; ++c) {
                    const bool m = (hb >> c) & 1u;
                    orow[3 * c] = m ? 0 : irow[3 * c];
                    orow[3 * c + 1] = m ? 0 : irow[3 * c + 1];
                    orow[3 * c + 2] = m ? 255 : irow[3 * c + 2];
                }
            }
        }
        const bool is_static = nz == 0u;
        if (is_static) ++n_static;
        if (compressed) {
            uint8_t* out = compressed + (size_t)blockIdx.y * H * W * 3;
            if (!is_static) {
                for (int r = 0; r < BS; ++r)
                    for (int c = 0; c < BS; ++c) {
                        const size_t o = ((size_t)(y0 + r) * W + x0 + c) * 3;
                        int b = fr[o], g = fr[o + 1], rr = fr[o + 2];
                        ycc_roundtrip(b, g, rr);
                        out[o] = (uint8_t)b; out[o + 1] = (uint8_t)g; out[o + 2] = (uint8_t)rr;
                    }
            } else if (FLAVOUR == 0) {
                float v[BS][BS];
#pragma unroll
                for (int r = 0; r < BS; ++r)
#pragma unroll
                    for (int c = 0; c < BS; ++c) {
                        const size_t o = ((size_t)(y0 + r) * W + x0 + c) * 3;
                        v[r][c] = (float)(luma_of(fr[o], fr[o + 1], fr[o + 2]) - 128);
                    }
                degrade_plane<BS>(v, q);
#pragma unroll
                for (int r = 0; r < BS; ++r)
#pragma unroll
                    for (int c = 0; c < BS; ++c) {
                        const size_t o = ((size_t)(y0 + r) * W + x0 + c) * 3;
                        const uint8_t y = (uint8_t)clip_trunc_u8(__fadd_rn(v[r][c], 128.0f));
                        out[o] = y; out[o + 1] = y; out[o + 2] = y;
                    }
            } else {
                // MCO: quantise Y, Cr, Cb; YCrCb->BGR; BGR->gray; replicate (motion_compression_opt.py:162-183)
                float v[BS][BS];
                uint8_t ch[3][BS][BS];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
#pragma unroll
                    for (int r = 0; r < BS; ++r)
#pragma unroll
                        for (int c = 0; c < BS; ++c) {
                            const size_t o = ((size_t)(y0 + r) * W + x0 + c) * 3;
                            const int b = fr[o], g = fr[o + 1], rr = fr[o + 2];
                            const int y = luma_of(b, g, rr);
                            int val = y;
                            if (k == 1) val = sat8(((rr - y) * 11682 + (128 << 14) + 8192) >> 14);
                            if (k == 2) val = sat8(((b - y) * 9241 + (128 << 14) + 8192) >> 14);
                            v[r][c] = (float)(val - 128);
                        }
                    degrade_plane<BS>(v, q);
#pragma unroll
                    for (int r = 0; r < BS; ++r)
#pragma unroll
                        for (int c = 0; c < BS; ++c) ch[k][r][c] = (uint8_t)clip_trunc_u8(__fadd_rn(v[r][c], 128.0f));
                }
#pragma unroll
                for (int r = 0; r < BS; ++r)
#pragma unroll
                    for (int c = 0; c < BS; ++c) {
                        const int y = ch[0][r][c], cr = ch[1][r][c] - 128, cb = ch[2][r][c] - 128;
                        const int b = sat8(y + ((29049 * cb + 8192) >> 14));
                        const int g = sat8(y + ((-5636 * cb - 11698 * cr + 8192) >> 14));
                        const int rr = sat8(y + ((22987 * cr + 8192) >> 14));
                        const uint8_t gy = (uint8_t)gray_of(b, g, rr);
                        const size_t o = ((size_t)(y0 + r) * W + x0 + c) * 3;
                        out[o] = gy; out[o + 1] = gy; out[o + 2] = gy;
                    }
            }
        }
    }
    if (counters) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_motion += __shfl_xor_sync(0xffffffffu, n_motion, o);
            n_static += __shfl_xor_sync(0xffffffffu, n_static, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (n_motion) atomicAdd(&counters->motion_pixels, (unsigned long long)n_motion);
            if (n_static) atomicAdd(&counters->static_blocks, (unsigned long long)n_static);
        }
    }
}

}  // namespace dvc
