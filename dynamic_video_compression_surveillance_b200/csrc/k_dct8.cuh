// Exact float32 restatement of the transforms cv2.dct / cv2.idct (OpenCV with IPP, float32) apply to the block sizes the
// reference can produce (frame_differencing.py:117-125, block_size 4 or 8 with clipped even-sized edge blocks;
// motion_compression_opt.py:156-168, 8x8).  Every operation order, FMA placement and constant below was recovered by
// search against cv2 on random vectors (oracle/stage_ops.py holds the same sequences in numpy and checks the host's
// cv2 against them):
//
//   * 8x8 blocks go through a dedicated 2-D routine.  Forward: each ROW as an even/odd 4x4 matrix product with FMA
//     chains (even outputs accumulate s0..s3, odd outputs d3..d0), then each COLUMN through a tangent-rotation
//     butterfly (the Intel AP-922 flow graph) with the output scale 0.5 cos(k pi / 16) applied last.  Inverse: scale
//     row k by 0.5 cos(k pi / 16), each ROW as even/odd FMA chains (0,2,4,6 / 1,3,5,7), then the column butterfly.
//   * every other shape is rows-then-columns (forward and inverse alike) through one 1-D routine per length:
//     N = 2, 4, 8 have dedicated butterflies, N = 3, 5, 6, 7 share a generic fold + FMA-chain routine.
//
// All arithmetic uses the explicit-rounding intrinsics, so nvcc can neither contract nor reassociate it.
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


// ---- packed pair of binary32 lanes in a 64-bit register pair ----
struct P2 { unsigned long long v; };
DEVI P2 p2(float lo, float hi) { P2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
DEVI void unp2(P2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }

DEVI P2 add(P2 a, P2 b) { P2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DEVI P2 sub(P2 a, P2 b) { P2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DEVI P2 mul(P2 a, P2 b) { P2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DEVI P2 fma_(P2 a, P2 b, P2 c) { P2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
DEVI float add(float a, float b) { return __fadd_rn(a, b); }
DEVI float sub(float a, float b) { return __fsub_rn(a, b); }
DEVI float mul(float a, float b) { return __fmul_rn(a, b); }
DEVI float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }

template <typename T> DEVI T splat(float x);
template <> DEVI float splat<float>(float x) { return x; }
template <> DEVI P2 splat<P2>(float x) { return p2(x, x); }

// 0.5 cos(j pi / 16), correctly rounded to float32 (j = 4: 1 / sqrt(8))
#define DCT8_C1 0.490392625f
#define DCT8_C2 0.461939752f
#define DCT8_C3 0.415734798f
#define DCT8_C4 0.353553385f
#define DCT8_C5 0.277785122f
#define DCT8_C6 0.191341713f
#define DCT8_C7 0.0975451618f
// tan(j pi / 16), cos(pi / 4)
#define DCT8_TG1 0.198912367f
#define DCT8_TG2 0.414213568f
#define DCT8_TG3 0.668178618f
#define DCT8_R 0.707106769f
// cos(j pi / 16) / sqrt(2): the 1-D 8-point routine's rotation constants
#define DCT8_B1 0.69351995f
#define DCT8_B3 0.587937772f
#define DCT8_B5 0.392847478f
#define DCT8_B7 0.13794969f

#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))
#define FF(a, b, c) __fmaf_rn((a), (b), (c))

// A[l][i] = s(l) cos(pi (2 i + 1) l / 16), i < 4 (the matrix is (anti)symmetric about the middle)
DEVI constexpr float dct8_a(int l, int i) {
    constexpr float t[8][4] = {
        {DCT8_C4, DCT8_C4, DCT8_C4, DCT8_C4},   {DCT8_C1, DCT8_C3, DCT8_C5, DCT8_C7},
        {DCT8_C2, DCT8_C6, -DCT8_C6, -DCT8_C2}, {DCT8_C3, -DCT8_C7, -DCT8_C1, -DCT8_C5},
        {DCT8_C4, -DCT8_C4, -DCT8_C4, DCT8_C4}, {DCT8_C5, -DCT8_C1, DCT8_C7, DCT8_C3},
        {DCT8_C6, -DCT8_C2, DCT8_C2, -DCT8_C6}, {DCT8_C7, -DCT8_C5, DCT8_C3, -DCT8_C1}};
    return t[l][i];
}

// ---- 2-D 8x8, forward ----
DEVI void dct8x8_row_fwd(float (&x)[8]) {
    float s[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { s[i] = FA(x[i], x[7 - i]); d[i] = FS(x[i], x[7 - i]); }
#pragma unroll
    for (int l = 0; l < 8; l += 2) {
        float acc = FM(dct8_a(l, 0), s[0]);
#pragma unroll
        for (int i = 1; i < 4; ++i) acc = FF(s[i], dct8_a(l, i), acc);
        x[l] = acc;
    }
#pragma unroll
    for (int l = 1; l < 8; l += 2) {
        float acc = FM(dct8_a(l, 3), d[3]);
#pragma unroll
        for (int i = 2; i >= 0; --i) acc = FF(d[i], dct8_a(l, i), acc);
        x[l] = acc;
    }
}
DEVI void dct8x8_col_fwd(float& x0, float& x1, float& x2, float& x3, float& x4, float& x5, float& x6, float& x7) {
    const float t0 = FA(x0, x7), t1 = FA(x1, x6), t2 = FA(x2, x5), t3 = FA(x3, x4);
    const float m0 = FS(x0, x7), m1 = FS(x1, x6), m2 = FS(x2, x5), m3 = FS(x3, x4);
    const float tp03 = FA(t0, t3), tm03 = FS(t0, t3), tp12 = FA(t1, t2), tm12 = FS(t1, t2);
    x0 = FM(DCT8_C4, FA(tp03, tp12));
    x4 = FM(DCT8_C4, FS(tp03, tp12));
    x2 = FM(DCT8_C2, FF(tm12, DCT8_TG2, tm03));
    x6 = FM(DCT8_C2, FF(tm03, DCT8_TG2, -tm12));
    const float tp65 = FM(FA(m1, m2), DCT8_R), tm65 = FM(FS(m1, m2), DCT8_R);
    const float tp765 = FA(m0, tp65), tm765 = FS(m0, tp65), tp465 = FA(m3, tm65), tm465 = FS(m3, tm65);
    x1 = FM(DCT8_C1, FF(tp465, DCT8_TG1, tp765));
    x7 = FM(DCT8_C1, FF(tp765, DCT8_TG1, -tp465));
    x5 = FM(DCT8_C3, FF(tm765, DCT8_TG3, tm465));
    x3 = FM(DCT8_C3, FF(tm465, -DCT8_TG3, tm765));
}
// ---- 2-D 8x8, inverse (the caller has already multiplied row k by dct8_rowscale(k)) ----
DEVI constexpr float dct8_rowscale(int k) {
    constexpr float t[8] = {DCT8_C4, DCT8_C1, DCT8_C2, DCT8_C3, DCT8_C4, DCT8_C3, DCT8_C2, DCT8_C1};
    return t[k];
}
DEVI void dct8x8_row_inv(float (&u)[8]) {
    float o[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float e = FM(dct8_a(0, c), u[0]), od = FM(dct8_a(1, c), u[1]);
#pragma unroll
        for (int l = 2; l < 8; l += 2) { e = FF(u[l], dct8_a(l, c), e); od = FF(u[l + 1], dct8_a(l + 1, c), od); }
        o[c] = FA(e, od);
        o[7 - c] = FS(e, od);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) u[c] = o[c];
}
DEVI void dct8x8_col_inv(float& x0, float& x1, float& x2, float& x3, float& x4, float& x5, float& x6, float& x7) {
    const float tp765 = FF(x7, DCT8_TG1, x1), tp465 = FF(x1, DCT8_TG1, -x7);
    const float tm765 = FF(x5, DCT8_TG3, x3), tm465 = FF(x3, -DCT8_TG3, x5);
    const float tm03 = FF(x6, DCT8_TG2, x2), tm12 = FF(x2, DCT8_TG2, -x6);
    const float t7 = FA(tp765, tm765), tp65 = FS(tp765, tm765), t4 = FA(tp465, tm465), tm65 = FS(tp465, tm465);
    const float p65 = FM(tp65, DCT8_R), m65 = FM(tm65, DCT8_R);
    const float t6 = FA(p65, m65), t5 = FS(p65, m65);
    const float tp03 = FA(x0, x4), tp12 = FS(x0, x4);
    const float t0 = FA(tp03, tm03), t3 = FS(tp03, tm03), t1 = FA(tp12, tm12), t2 = FS(tp12, tm12);
    x0 = FA(t0, t7); x7 = FS(t0, t7);
    x1 = FA(t1, t6); x6 = FS(t1, t6);
    x2 = FA(t2, t5); x5 = FS(t2, t5);
    x3 = FA(t3, t4); x4 = FS(t3, t4);
}

// ---- the same 2-D sequences over T = float or the packed pair P2 (two rows, or two columns, per instruction) ----
// A product that feeds an addition must stay a separate rounding: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2
// in spite of the modifiers (k_degrade4p.cuh), so those few products go through scalar FMULs (mul_sep).  Negated addends are
// folded into the constants (round(ab - c) = -round(c - ab)), because the packed fma has no negate modifier.
DEVI float mul_sep(float a, float b) { return __fmul_rn(a, b); }
DEVI P2 mul_sep(P2 a, float b) { float lo, hi; unp2(a, lo, hi); return p2(__fmul_rn(lo, b), __fmul_rn(hi, b)); }

template <typename T>
DEVI void dct8x8_row_fwd_t(T (&x)[8]) {
    T s[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { s[i] = add(x[i], x[7 - i]); d[i] = sub(x[i], x[7 - i]); }
#pragma unroll
    for (int l = 0; l < 8; l += 2) {
        T acc = mul(splat<T>(dct8_a(l, 0)), s[0]);
#pragma unroll
        for (int i = 1; i < 4; ++i) acc = fma_(s[i], splat<T>(dct8_a(l, i)), acc);
        x[l] = acc;
    }
#pragma unroll
    for (int l = 1; l < 8; l += 2) {
        T acc = mul(splat<T>(dct8_a(l, 3)), d[3]);
#pragma unroll
        for (int i = 2; i >= 0; --i) acc = fma_(d[i], splat<T>(dct8_a(l, i)), acc);
        x[l] = acc;
    }
}
template <typename T>
DEVI void dct8x8_col_fwd_t(T& x0, T& x1, T& x2, T& x3, T& x4, T& x5, T& x6, T& x7) {
    const T t0 = add(x0, x7), t1 = add(x1, x6), t2 = add(x2, x5), t3 = add(x3, x4);
    const T m0 = sub(x0, x7), m1 = sub(x1, x6), m2 = sub(x2, x5), m3 = sub(x3, x4);
    const T tp03 = add(t0, t3), tm03 = sub(t0, t3), tp12 = add(t1, t2), tm12 = sub(t1, t2);
    x0 = mul(splat<T>(DCT8_C4), add(tp03, tp12));
    x4 = mul(splat<T>(DCT8_C4), sub(tp03, tp12));
    x2 = mul(splat<T>(DCT8_C2), fma_(tm12, splat<T>(DCT8_TG2), tm03));
    x6 = mul(splat<T>(-DCT8_C2), fma_(tm03, splat<T>(-DCT8_TG2), tm12));          // C2 * fma(tm03, TG2, -tm12)
    const T tp65 = mul_sep(add(m1, m2), DCT8_R), tm65 = mul_sep(sub(m1, m2), DCT8_R);   // products that feed additions
    const T tp765 = add(m0, tp65), tm765 = sub(m0, tp65), tp465 = add(m3, tm65), tm465 = sub(m3, tm65);
    x1 = mul(splat<T>(DCT8_C1), fma_(tp465, splat<T>(DCT8_TG1), tp765));
    x7 = mul(splat<T>(-DCT8_C1), fma_(tp765, splat<T>(-DCT8_TG1), tp465));        // C1 * fma(tp765, TG1, -tp465)
    x5 = mul(splat<T>(DCT8_C3), fma_(tm765, splat<T>(DCT8_TG3), tm465));
    x3 = mul(splat<T>(DCT8_C3), fma_(tm465, splat<T>(-DCT8_TG3), tm765));
}
template <typename T>
DEVI void dct8x8_row_inv_t(T (&u)[8]) {
    T o[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        T e = mul(splat<T>(dct8_a(0, c)), u[0]), od = mul(splat<T>(dct8_a(1, c)), u[1]);
#pragma unroll
        for (int l = 2; l < 8; l += 2) { e = fma_(u[l], splat<T>(dct8_a(l, c)), e); od = fma_(u[l + 1], splat<T>(dct8_a(l + 1, c)), od); }
        o[c] = add(e, od);
        o[7 - c] = sub(e, od);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) u[c] = o[c];
}
template <typename T>
DEVI void dct8x8_col_inv_t(T& x0, T& x1, T& x2, T& x3, T& x4, T& x5, T& x6, T& x7) {
    const T tp765 = fma_(x7, splat<T>(DCT8_TG1), x1), q465 = fma_(x1, splat<T>(-DCT8_TG1), x7);       // q465 = -tp465
    const T tm765 = fma_(x5, splat<T>(DCT8_TG3), x3), tm465 = fma_(x3, splat<T>(-DCT8_TG3), x5);
    const T tm03 = fma_(x6, splat<T>(DCT8_TG2), x2), q12 = fma_(x2, splat<T>(-DCT8_TG2), x6);          // q12 = -tm12
    const T t7 = add(tp765, tm765), tp65 = sub(tp765, tm765);
    const T t4 = sub(tm465, q465);                                      // tp465 + tm465
    const T p65 = mul_sep(tp65, DCT8_R), m65 = mul_sep(add(q465, tm465), -DCT8_R);     // m65 = (tp465 - tm465) * R
    const T t6 = add(p65, m65), t5 = sub(p65, m65);
    const T tp03 = add(x0, x4), tp12 = sub(x0, x4);
    const T t0 = add(tp03, tm03), t3 = sub(tp03, tm03), t1 = sub(tp12, q12), t2 = add(tp12, q12);
    x0 = add(t0, t7); x7 = sub(t0, t7);
    x1 = add(t1, t6); x6 = sub(t1, t6);
    x2 = add(t2, t5); x5 = sub(t2, t5);
    x3 = add(t3, t4); x4 = sub(t3, t4);
}

// One 8x8 block through the packed pipe: rows are transformed two at a time (rp[i][c] = rows 2i, 2i+1 of column c), columns
// two at a time (cp[r][j] = columns 2j, 2j+1 of row r); between passes the 2x2 sub-blocks are re-paired (register moves).
// quant: P2 -> P2, np.round(d / q) * q on both lanes.  Same values as degrade_block8_exact, about 0.6x the instructions.
template <typename Quant>
DEVI void degrade_block8_packed(float (&v)[8][8], Quant quant) {
    P2 rp[4][8], cp[8][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) rp[i][c] = p2(v[2 * i][c], v[2 * i + 1][c]);
#pragma unroll
    for (int i = 0; i < 4; ++i) dct8x8_row_fwd_t(rp[i]);
    auto rows_to_cols = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a, c, b, d;
                unp2(rp[i][2 * j], a, c);            // (row 2i, row 2i+1) of column 2j
                unp2(rp[i][2 * j + 1], b, d);        // ... of column 2j+1
                cp[2 * i][j] = p2(a, b);
                cp[2 * i + 1][j] = p2(c, d);
            }
    };
    auto cols_to_rows = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a, b, c, d;
                unp2(cp[2 * i][j], a, b);
                unp2(cp[2 * i + 1][j], c, d);
                rp[i][2 * j] = p2(a, c);
                rp[i][2 * j + 1] = p2(b, d);
            }
    };
    rows_to_cols();
#pragma unroll
    for (int j = 0; j < 4; ++j) dct8x8_col_fwd_t(cp[0][j], cp[1][j], cp[2][j], cp[3][j], cp[4][j], cp[5][j], cp[6][j], cp[7][j]);
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) cp[r][j] = mul(quant(cp[r][j]), splat<P2>(dct8_rowscale(r)));
    cols_to_rows();
#pragma unroll
    for (int i = 0; i < 4; ++i) dct8x8_row_inv_t(rp[i]);
    rows_to_cols();
#pragma unroll
    for (int j = 0; j < 4; ++j) dct8x8_col_inv_t(cp[0][j], cp[1][j], cp[2][j], cp[3][j], cp[4][j], cp[5][j], cp[6][j], cp[7][j]);
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) unp2(cp[r][j], v[r][2 * j], v[r][2 * j + 1]);
}

// clip(idct(round(dct(v) / q) * q)) of one 8x8 block held in registers; Q(d) is the caller's quantiser (same value as
// np.round(d / q) * q in float32).
template <typename Quant>
DEVI void degrade_block8_exact(float (&v)[8][8], Quant quant) {
#pragma unroll
    for (int r = 0; r < 8; ++r) dct8x8_row_fwd(v[r]);
#pragma unroll
    for (int c = 0; c < 8; ++c) dct8x8_col_fwd(v[0][c], v[1][c], v[2][c], v[3][c], v[4][c], v[5][c], v[6][c], v[7][c]);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int c = 0; c < 8; ++c) v[r][c] = FM(quant(v[r][c]), dct8_rowscale(r));
        dct8x8_row_inv(v[r]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) dct8x8_col_inv(v[0][c], v[1][c], v[2][c], v[3][c], v[4][c], v[5][c], v[6][c], v[7][c]);
}

// ---- 1-D routines (clipped edge blocks; x has stride `st` floats) ----
DEVI void dct1d_2(float* x, int st, bool) {
    const float a = FM(x[0], DCT8_R), b = FM(x[st], DCT8_R);
    x[0] = FA(a, b);
    x[st] = FS(a, b);
}
DEVI void dct1d_4(float* x, int st, bool inverse) {
    if (inverse) dct4_inv(x[0], x[st], x[2 * st], x[3 * st]);
    else dct4_fwd(x[0], x[st], x[2 * st], x[3 * st]);
}
// generated by tools/gen_dct_tables.py -- c_dctg[slot][k][i] = float32(cos(pi (2 i + 1) k / (2 N))), slot: N = 3, 5, 6, 7
__constant__ float c_dctg[4][7][4] = {
    {{1.0f, 1.0f, 0.0f, 0.0f},
     {0.866025388f, 6.12323426e-17f, 0.0f, 0.0f},
     {0.5f, -1.0f, 0.0f, 0.0f},
     {0.0f, 0.0f, 0.0f, 0.0f},
     {0.0f, 0.0f, 0.0f, 0.0f},
     {0.0f, 0.0f, 0.0f, 0.0f},
     {0.0f, 0.0f, 0.0f, 0.0f}},
    {{1.0f, 1.0f, 1.0f, 0.0f},
     {0.95105654f, 0.587785244f, 6.12323426e-17f, 0.0f},
     {0.809017003f, -0.309017003f, -1.0f, 0.0f},
     {0.587785244f, -0.95105654f, -1.83697015e-16f, 0.0f},
     {0.309017003f, -0.809017003f, 1.0f, 0.0f},
     {0.0f, 0.0f, 0.0f, 0.0f},
     {0.0f, 0.0f, 0.0f, 0.0f}},
    {{1.0f, 1.0f, 1.0f, 0.0f},
     {0.965925813f, 0.707106769f, 0.258819044f, 0.0f},
     {0.866025388f, 6.12323426e-17f, -0.866025388f, 0.0f},
     {0.707106769f, -0.707106769f, -0.707106769f, 0.0f},
     {0.5f, -1.0f, 0.5f, 0.0f},
     {0.258819044f, -0.707106769f, 0.965925813f, 0.0f},
     {0.0f, 0.0f, 0.0f, 0.0f}},
    {{1.0f, 1.0f, 1.0f, 1.0f},
     {0.974927902f, 0.781831503f, 0.433883727f, 6.12323426e-17f},
     {0.90096885f, 0.222520933f, -0.623489797f, -1.0f},
     {0.781831503f, -0.433883727f, -0.974927902f, -1.83697015e-16f},
     {0.623489797f, -0.90096885f, -0.222520933f, 1.0f},
     {0.433883727f, -0.974927902f, 0.781831503f, 3.061617e-16f},
     {0.222520933f, -0.623489797f, 0.90096885f, -1.0f}},
};
__constant__ float c_dctg_scale[4][2] = {{0.577350259f, 0.816496611f}, {0.44721359f, 0.632455528f}, {0.408248305f, 0.577350259f}, {0.377964467f, 0.534522474f}};   // sqrt(1/N), sqrt(2/N)
// Lengths 3, 5, 6, 7: fold about the middle, even / odd outputs as FMA chains over the unnormalised cosines (odd
// lengths start the even chains with the middle sample), scale by sqrt(2/N) (sqrt(1/N) for k = 0) last.  Inverse: scale
// first, even / odd FMA chains per output pair; the middle sample of an odd length is (w0 + w4) - (w2 + w6).
DEVI void dct1d_generic(float* x, int n, int st, bool inverse) {
    const int slot = n == 3 ? 0 : n == 5 ? 1 : n == 6 ? 2 : 3, h = n >> 1;
    const bool odd = n & 1;
    const float (*M)[4] = c_dctg[slot];
    const float k0 = c_dctg_scale[slot][0], k1 = c_dctg_scale[slot][1];
    float v[7], o[7];
    for (int i = 0; i < n; ++i) v[i] = x[i * st];
    if (!inverse) {
        float s[4], d[4];
        for (int i = 0; i < h; ++i) { s[i] = FA(v[i], v[n - 1 - i]); d[i] = FS(v[i], v[n - 1 - i]); }
        for (int k = 0; k < n; ++k) {
            float acc;
            if (k & 1) {
                acc = FM(M[k][0], d[0]);
                for (int i = 1; i < h; ++i) acc = FF(d[i], M[k][i], acc);
            } else if (odd) {
                acc = FM(M[k][h], v[h]);
                for (int i = 0; i < h; ++i) acc = FF(s[i], M[k][i], acc);
            } else {
                acc = FM(M[k][0], s[0]);
                for (int i = 1; i < h; ++i) acc = FF(s[i], M[k][i], acc);
            }
            o[k] = FM(acc, k == 0 ? k0 : k1);
        }
    } else {
        float w[7];
        for (int k = 0; k < n; ++k) w[k] = FM(v[k], k == 0 ? k0 : k1);
        for (int i = 0; i < h; ++i) {
            float e = FM(M[0][i], w[0]), od = FM(M[1][i], w[1]);
            for (int k = 2; k < n; k += 2) e = FF(w[k], M[k][i], e);
            for (int k = 3; k < n; k += 2) od = FF(w[k], M[k][i], od);
            o[i] = FA(e, od);
            o[n - 1 - i] = FS(e, od);
        }
        if (odd) {
            float p = w[0], m = w[2];
            if (n > 4) p = FA(p, w[4]);
            if (n > 6) m = FA(m, w[6]);
            o[h] = FS(p, m);
        }
    }
    for (int i = 0; i < n; ++i) x[i * st] = o[i];
}
DEVI void dct1d_8(float* x, int st, bool inverse) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = x[i * st];
    if (!inverse) {
        float s[4], d[4];
        for (int i = 0; i < 4; ++i) { s[i] = FA(v[i], v[7 - i]); d[i] = FS(v[i], v[7 - i]); }
        const float e0 = FA(s[0], s[3]), e1 = FA(s[1], s[2]), f0 = FS(s[0], s[3]), f1 = FS(s[1], s[2]);
        const float u0 = FM(d[0], DCT8_R), u3 = FM(d[3], DCT8_R);
        const float p65 = FM(FA(d[1], d[2]), 0.5f), m65 = FM(FS(d[1], d[2]), 0.5f);
        const float tp765 = FA(u0, p65), tm765 = FS(u0, p65), tp465 = FA(u3, m65), tm465 = FS(u3, m65);
        x[0] = FM(DCT8_C4, FA(e0, e1));
        x[4 * st] = FM(DCT8_C4, FS(e0, e1));
        x[2 * st] = FF(f0, DCT8_C2, FM(DCT8_C6, f1));
        x[6 * st] = FF(f0, DCT8_C6, -FM(DCT8_C2, f1));
        x[1 * st] = FF(tp465, DCT8_B7, FM(DCT8_B1, tp765));
        x[7 * st] = FF(tp765, DCT8_B7, -FM(DCT8_B1, tp465));
        x[5 * st] = FF(tm465, DCT8_B3, FM(DCT8_B5, tm765));
        x[3 * st] = FF(tm765, DCT8_B3, -FM(DCT8_B5, tm465));
    } else {
        const float a0 = FM(DCT8_C4, v[0]), a4 = FM(DCT8_C4, v[4]);
        const float ap = FA(a0, a4), am = FS(a0, a4);
        const float b0 = FF(v[6], DCT8_C6, FM(DCT8_C2, v[2])), b1 = FF(v[2], DCT8_C6, -FM(DCT8_C2, v[6]));
        const float e[4] = {FA(ap, b0), FA(am, b1), FS(am, b1), FS(ap, b0)};
        const float tp765 = FF(v[7], DCT8_B7, FM(DCT8_B1, v[1])), tp465 = FF(v[1], DCT8_B7, -FM(DCT8_B1, v[7]));
        const float tm765 = FF(v[3], DCT8_B3, FM(DCT8_B5, v[5])), tm465 = FF(v[5], DCT8_B3, -FM(DCT8_B5, v[3]));
        const float p65 = FM(FS(tp765, tm765), 0.5f), m65 = FM(FS(tp465, tm465), 0.5f);
        const float o[4] = {FM(DCT8_R, FA(tp765, tm765)), FA(p65, m65), FS(p65, m65), FM(DCT8_R, FA(tp465, tm465))};
        for (int n = 0; n < 4; ++n) { x[n * st] = FA(e[n], o[n]); x[(7 - n) * st] = FS(e[n], o[n]); }
    }
}
// length-n transform of a strided vector (n = 1..8); n = 1 is the identity (a one-pixel-wide clipped block is a 1-D
// transform in cv2).
DEVI bool dct1d(float* x, int n, int st, bool inverse) {
    switch (n) {
        case 1: return true;
        case 2: dct1d_2(x, st, inverse); return true;
        case 4: dct1d_4(x, st, inverse); return true;
        case 8: dct1d_8(x, st, inverse); return true;
        case 3: case 5: case 6: case 7: dct1d_generic(x, n, st, inverse); return true;
        default: return false;
    }
}

#undef FM
#undef FA
#undef FS
#undef FF

}  // namespace dvc
