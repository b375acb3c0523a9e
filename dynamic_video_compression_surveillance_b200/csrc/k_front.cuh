// K1: the mask front end.
//   k_gray_diff_thresh  BGR -> gray -> absdiff(prev) -> threshold, time loop inside the kernel
//                       (frame_differencing.py:92,96-97 without the pre-blur; window mode)
//   k_gray_blur5        BGR -> gray -> GaussianBlur(5,5),0 (frame_differencing.py:92-93; fd mode)
//   k_diff_thresh_planes  absdiff + threshold on blurred gray planes (frame_differencing.py:96-97)
//   k_bgr2gray, k_pack_bits, k_unpack_bits  stage-level helpers
#pragma once
#include "common.cuh"

namespace dvc {

// ------------------------------------------------------------------------------------------------
// scalar fallbacks for widths whose BGR row pitch is not a multiple of 16 bytes
// ------------------------------------------------------------------------------------------------
DEVI void load_bgr16_generic(const uint8_t* row, int x0, int W, uint32_t (&w)[12]) {
#pragma unroll
    for (int i = 0; i < 12; ++i) w[i] = 0;
    int nb = min(16, W - x0) * 3;
    const uint8_t* p = row + (size_t)x0 * 3;
    for (int i = 0; i < nb; ++i) w[i >> 2] |= (uint32_t)p[i] << ((i & 3) * 8);
}
DEVI void load_u8x16_generic(const uint8_t* row, int x0, int W, uint32_t (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = 0;
    int nb = min(16, W - x0);
    for (int i = 0; i < nb; ++i) v[i >> 2] |= (uint32_t)row[x0 + i] << ((i & 3) * 8);
}
DEVI void store_u8x16_generic(uint8_t* row, int x0, int W, const uint32_t (&v)[4]) {
    int nb = min(16, W - x0);
    for (int i = 0; i < nb; ++i) row[x0 + i] = (uint8_t)(v[i >> 2] >> ((i & 3) * 8));
}

// ------------------------------------------------------------------------------------------------
// K1 (window mode).  One thread owns 16 horizontally adjacent pixels and walks a segment of frames,
// keeping the previous gray values in registers: per frame it loads 48 B of BGR (3 x LDG.128),
// and stores one 16-bit piece of the raw-mask bit-plane.  Algorithmic HBM bytes: 3 B/px read
// (+ 1/8 B/px written); the previous gray never touches memory inside a segment.
//
// ring:   raw-mask ring buffer, plane of frame f lives in slot f % ring_cap
// f0:     global index of frames[0] within the stream
// blockIdx.z = stream of a lock-step stream group (dvc_config.n_streams): frames [S][T], one prev-gray plane, ring and
// gray output run per stream, laid out stream-major.
// ------------------------------------------------------------------------------------------------
template <bool ALIGNED, int GRAY = 0>   // GRAY: 0 = PRMT + IMAD, 1 = IDP.4A, 2 = IDP.2A
__global__ void __launch_bounds__(256)
k_gray_diff_thresh(const uint8_t* __restrict__ frames, int T, int H, int W,
                   const uint8_t* __restrict__ prev_gray_in, uint8_t* __restrict__ gray_state_out,
                   uint8_t* __restrict__ gray_all_out, uint32_t* __restrict__ ring, int wpr, int ring_cap,
                   long long f0, uint32_t thr, int seg_len) {
    const int gpr = (W + 15) >> 4;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)gpr * H) return;
    const int y = (int)(gid / gpr), gx = (int)(gid % gpr), x0 = gx << 4;
    const int t0 = blockIdx.y * seg_len, t1 = min(T, t0 + seg_len);
    const size_t frame_bytes = (size_t)H * W * 3, plane_bytes = (size_t)H * W;
    const size_t row_off = (size_t)y * W * 3, px_off = (size_t)y * W + x0;
    const size_t plane_words = (size_t)H * wpr;
    const uint32_t vmask = (W - x0 >= 16) ? 0xffffu : ((1u << (W - x0)) - 1u);
    {
        const size_t s = blockIdx.z;
        frames += s * T * frame_bytes;
        prev_gray_in += s * plane_bytes;
        if (gray_state_out) gray_state_out += s * plane_bytes;
        if (gray_all_out) gray_all_out += s * T * plane_bytes;
        ring += s * ring_cap * plane_words;
    }

    uint32_t pg[4], w[12];
    if (t0 == 0) {
        if (ALIGNED) load16(prev_gray_in + px_off, pg);
        else load_u8x16_generic(prev_gray_in + (size_t)y * W, x0, W, pg);
    } else {
        const uint8_t* fr = frames + (size_t)(t0 - 1) * frame_bytes + row_off;
        if (ALIGNED) {
            load16(fr + x0 * 3, *reinterpret_cast<uint32_t(*)[4]>(&w[0]));
            load16(fr + x0 * 3 + 16, *reinterpret_cast<uint32_t(*)[4]>(&w[4]));
            load16(fr + x0 * 3 + 32, *reinterpret_cast<uint32_t(*)[4]>(&w[8]));
        } else load_bgr16_generic(fr, x0, W, w);
        if (GRAY == 2) gray16_dp2a(w, pg); else if (GRAY == 1) gray16_dp4a(w, pg); else gray16(w, pg);
    }
    for (int t = t0; t < t1; ++t) {
        const uint8_t* fr = frames + (size_t)t * frame_bytes + row_off;
        if (ALIGNED) {
            load16(fr + x0 * 3, *reinterpret_cast<uint32_t(*)[4]>(&w[0]));
            load16(fr + x0 * 3 + 16, *reinterpret_cast<uint32_t(*)[4]>(&w[4]));
            load16(fr + x0 * 3 + 32, *reinterpret_cast<uint32_t(*)[4]>(&w[8]));
        } else load_bgr16_generic(fr, x0, W, w);
        uint32_t g[4];
        if (GRAY == 2) gray16_dp2a(w, g); else if (GRAY == 1) gray16_dp4a(w, g); else gray16(w, g);
        uint32_t bits = diff_gt_bits16(g, pg, thr);
        bits &= vmask;
        const int slot = (int)((f0 + t) % ring_cap);
        reinterpret_cast<uint16_t*>(ring + (size_t)slot * plane_words + (size_t)y * wpr)[gx] = (uint16_t)bits;
        if (gray_all_out) {
            if (ALIGNED) store16(gray_all_out + (size_t)t * plane_bytes + px_off, g);
            else store_u8x16_generic(gray_all_out + (size_t)t * plane_bytes + (size_t)y * W, x0, W, g);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) pg[q] = g[q];
    }
    if (t1 == T && gray_state_out) {
        if (ALIGNED) store16(gray_state_out + px_off, pg);
        else store_u8x16_generic(gray_state_out + (size_t)y * W, x0, W, pg);
    }
}

// ------------------------------------------------------------------------------------------------
// stage helper: BGR -> gray for n images
// ------------------------------------------------------------------------------------------------
template <bool ALIGNED>
__global__ void __launch_bounds__(256)
k_bgr2gray(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray, int H, int W) {
    const int gpr = (W + 15) >> 4;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)gpr * H) return;
    const int y = (int)(gid / gpr), x0 = (int)(gid % gpr) << 4;
    const uint8_t* fr = bgr + (size_t)blockIdx.y * H * W * 3 + (size_t)y * W * 3;
    uint8_t* out = gray + (size_t)blockIdx.y * H * W + (size_t)y * W;
    uint32_t w[12], g[4];
    if (ALIGNED) {
        load16(fr + x0 * 3, *reinterpret_cast<uint32_t(*)[4]>(&w[0]));
        load16(fr + x0 * 3 + 16, *reinterpret_cast<uint32_t(*)[4]>(&w[4]));
        load16(fr + x0 * 3 + 32, *reinterpret_cast<uint32_t(*)[4]>(&w[8]));
    } else load_bgr16_generic(fr, x0, W, w);
    gray16_dp2a(w, g);
    if (ALIGNED) store16(out + x0, g);
    else store_u8x16_generic(out, x0, W, g);
}

// ------------------------------------------------------------------------------------------------
// K1 (fd mode), part 1: BGR -> gray -> GaussianBlur(5,5),0.
// cv2 evaluates the binomial kernel [1,4,6,4,1]/16 in 8.8 fixed point; the result is
// (sum over the separable 5x5 integer kernel + 128) >> 8 with BORDER_REFLECT_101
// (frame_differencing.py:93).  Tile 128 x 32 output pixels per CTA; gray of tile + 2-pixel halo is
// built in shared memory (interior through 48-byte vector loads, halo columns scalar), then the
// horizontal and vertical 5-tap passes run out of shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int BL_TW = 128, BL_TH = 32, BL_PAD = 16, BL_GP = BL_TW + 2 * BL_PAD;  // gray pitch, 16 B pad each side

DEVI int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256)
k_gray_blur5(const uint8_t* __restrict__ frames, uint8_t* __restrict__ blurred, int H, int W) {
    __shared__ __align__(16) uint8_t sg[(BL_TH + 4) * BL_GP];      // gray, column c <-> x = x0 - BL_PAD + c
    __shared__ __align__(16) uint2 sh[(BL_TH + 4) * (BL_TW / 4)];   // horizontal pass: (even, odd) 16-bit lane pairs per 4 px
    const int x0 = blockIdx.x * BL_TW, y0 = blockIdx.y * BL_TH;
    const uint8_t* fr = frames + (size_t)blockIdx.z * H * W * 3;
    uint8_t* out = blurred + (size_t)blockIdx.z * H * W;
    const int tid = threadIdx.x;

    // phase 1a: interior gray, 16 pixels per task
    for (int task = tid; task < (BL_TH + 4) * (BL_TW / 16); task += 256) {
        const int r = task / (BL_TW / 16), g = task % (BL_TW / 16);
        const int x = x0 + g * 16;
        if (x >= W) continue;
        const int yy = reflect101(y0 - 2 + r, H);
        const uint8_t* row = fr + (size_t)yy * W * 3;
        uint32_t w[12], gg[4];
        if (ALIGNED && x + 16 <= W) {
            load16(row + x * 3, *reinterpret_cast<uint32_t(*)[4]>(&w[0]));
            load16(row + x * 3 + 16, *reinterpret_cast<uint32_t(*)[4]>(&w[4]));
            load16(row + x * 3 + 32, *reinterpret_cast<uint32_t(*)[4]>(&w[8]));
        } else load_bgr16_generic(row, x, W, w);
        gray16_dp2a(w, gg);
        *reinterpret_cast<uint4*>(&sg[r * BL_GP + BL_PAD + g * 16]) = make_uint4(gg[0], gg[1], gg[2], gg[3]);
    }
    __syncthreads();
    // phase 1b: halo columns.  Left: x0-2, x0-1.  Right: the two columns after the last valid
    // column of this tile (tile may be clipped by the image edge).
    const int tw = min(BL_TW, W - x0);              // valid columns in this tile
    for (int task = tid; task < (BL_TH + 4) * 4; task += 256) {
        const int r = task >> 2, k = task & 3;
        const int dx = k < 2 ? k - 2 : tw + (k - 2);   // column offset relative to x0
        const int xx = reflect101(x0 + dx, W);
        const int yy = reflect101(y0 - 2 + r, H);
        uint8_t v;
        if (xx >= x0 && xx < x0 + tw) v = sg[r * BL_GP + BL_PAD + (xx - x0)];   // reflected into this tile
        else {
            const uint8_t* p = fr + ((size_t)yy * W + xx) * 3;
            v = (uint8_t)gray_of(p[0], p[1], p[2]);
        }
        sg[r * BL_GP + BL_PAD + dx] = v;
    }
    __syncthreads();
    // phase 2: horizontal 5 taps on 16-bit lanes (two pixels per 32-bit operation).  A task is one aligned group of
    // four pixels: the shifted byte windows come from funnel shifts of (prev, cur, next), PRMT widens even / odd bytes
    // to 16-bit lanes, and h = (a + e) + 4 (b + d) + 6 c <= 4080 per lane.  Result per group: (even lanes, odd lanes).
    for (int task = tid; task < (BL_TH + 4) * (BL_TW / 4); task += 256) {
        const int r = task / (BL_TW / 4), cg = task % (BL_TW / 4);
        const uint32_t* g = reinterpret_cast<const uint32_t*>(&sg[r * BL_GP + BL_PAD]) + cg;
        const uint32_t prev = g[-1], cur = g[0], next = g[1];
        const uint32_t a = __funnelshift_r(prev, cur, 16), b = __funnelshift_r(prev, cur, 24);
        const uint32_t d = __funnelshift_r(cur, next, 8), e = __funnelshift_r(cur, next, 16);
        uint2 h;
        {
            const uint32_t ae = __byte_perm(a, 0u, 0x4240u) + __byte_perm(e, 0u, 0x4240u);
            const uint32_t bd = __byte_perm(b, 0u, 0x4240u) + __byte_perm(d, 0u, 0x4240u);
            h.x = ae + 4u * bd + 6u * __byte_perm(cur, 0u, 0x4240u);
        }
        {
            const uint32_t ae = __byte_perm(a, 0u, 0x4341u) + __byte_perm(e, 0u, 0x4341u);
            const uint32_t bd = __byte_perm(b, 0u, 0x4341u) + __byte_perm(d, 0u, 0x4341u);
            h.y = ae + 4u * bd + 6u * __byte_perm(cur, 0u, 0x4341u);
        }
        sh[r * (BL_TW / 4) + cg] = h;
    }
    __syncthreads();
    // phase 3: vertical 5 taps + rounding, still two pixels per operation (v <= 65280 fits a 16-bit lane).  A thread
    // owns one 4-pixel column group and four consecutive output rows, sliding a 5-row window through registers; a warp
    // writes 128 contiguous bytes per row.
    {
        const int cg = tid & 31, r0 = (tid >> 5) * 4;
        const uint2* col = sh + r0 * (BL_TW / 4) + cg;
        uint2 w0 = col[0], w1 = col[BL_TW / 4], w2 = col[2 * (BL_TW / 4)], w3 = col[3 * (BL_TW / 4)];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint2 w4 = col[(4 + i) * (BL_TW / 4)];
            const uint32_t ve = (w0.x + w4.x) + 4u * (w1.x + w3.x) + 6u * w2.x;
            const uint32_t vo = (w0.y + w4.y) + 4u * (w1.y + w3.y) + 6u * w2.y;
            const uint32_t re = ((ve + 0x00800080u) >> 8) & 0x00ff00ffu, ro = ((vo + 0x00800080u) >> 8) & 0x00ff00ffu;
            const uint32_t o = re | (ro << 8);
            const int y = y0 + r0 + i, x = x0 + cg * 4;
            if (y < H && x < W) {
                uint8_t* q = out + (size_t)y * W + x;
                if ((W & 3) == 0) *reinterpret_cast<uint32_t*>(q) = o;
                else for (int k = 0; k < min(4, W - x); ++k) q[k] = (uint8_t)(o >> (8 * k));
            }
            w0 = w1; w1 = w2; w2 = w3; w3 = w4;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1 (fd mode), part 2: absdiff + threshold between consecutive blurred planes -> raw-mask bit-planes.
// planes[t] is differenced against planes[t-1]; planes[-1] is prev_gray (stream state).
// ------------------------------------------------------------------------------------------------
template <bool ALIGNED>
__global__ void __launch_bounds__(256)
k_diff_thresh_planes(const uint8_t* __restrict__ planes, const uint8_t* __restrict__ prev_gray, int H, int W,
                     uint32_t* __restrict__ bits_out, int wpr, uint32_t thr) {     // grid (.., T, S): planes [S][T], prev_gray [S]
    const int gpr = (W + 15) >> 4;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)gpr * H) return;
    const int y = (int)(gid / gpr), gx = (int)(gid % gpr), x0 = gx << 4;
    const int t = blockIdx.y;
    const size_t plane_bytes = (size_t)H * W;
    planes += (size_t)blockIdx.z * gridDim.y * plane_bytes;
    prev_gray += (size_t)blockIdx.z * plane_bytes;
    bits_out += (size_t)blockIdx.z * gridDim.y * H * wpr;
    const uint8_t* cur = planes + (size_t)t * plane_bytes + (size_t)y * W;
    const uint8_t* prv = (t == 0 ? prev_gray : planes + (size_t)(t - 1) * plane_bytes) + (size_t)y * W;
    uint32_t a[4], b[4];
    if (ALIGNED) { load16(cur + x0, a); load16(prv + x0, b); }
    else { load_u8x16_generic(cur, x0, W, a); load_u8x16_generic(prv, x0, W, b); }
    uint32_t bits = diff_gt_bits16(a, b, thr);
    if (W - x0 < 16) bits &= (1u << (W - x0)) - 1u;
    reinterpret_cast<uint16_t*>(bits_out + (size_t)t * H * wpr + (size_t)y * wpr)[gx] = (uint16_t)bits;
}

// ------------------------------------------------------------------------------------------------
// K1 (fd mode), fused and time-walking: BGR -> gray -> GaussianBlur(5,5),0 -> absdiff(previous blurred frame) ->
// threshold -> raw-mask bit-plane (frame_differencing.py:92-97), one CTA per 128 x 64 tile walking a segment of frames.
// The blurred frames never touch HBM: each thread keeps the previous blurred values of its 8 px x 4 rows in registers, so
// the kernel reads 3 B/px and writes 1/8 B/px (k_gray_blur5 + k_diff_thresh_planes moved 3 + 1 + 2 B/px); only the last
// blurred frame of the batch is stored (it is the stream's prev_gray, :133).  The thread -> task mapping and every
// reflected source offset are computed once, outside the frame loop.
// grid (tiles_x, tiles_y, S * nseg): z = stream * nseg + segment; a segment that does not start the batch first blurs
// the frame before it (1 / seg_len extra work) instead of waiting for its neighbour.
// ------------------------------------------------------------------------------------------------
// 60 output rows + 4 halo rows = 64 staged rows: 64 x 8 groups of 16 px are exactly two rounds of the CTA's 256 threads for the gray
// and the horizontal pass and 64 x 4 halo pixels exactly one (a 64-row tile needed three rounds, the third 1/8 full), and 1080 =
// 18 x 60.  The vertical pass (8 px x 4 rows per thread) uses 240 of the 256 threads.
constexpr int FF_TW = 128, FF_TH = 60, FF_PAD = 16, FF_GP = FF_TW + 2 * FF_PAD, FF_ROWS = FF_TH + 4;
constexpr int FF_GR = (FF_ROWS * 8 + 255) / 256, FF_HR = (FF_ROWS * 4 + 255) / 256;      // rounds of gray / halo tasks

// 6 CTAs per SM = 40 registers (a few spills): alone the kernel is 1 % slower than with 48, beside the contour filter and the EMA of
// the previous batch (three-stream pipeline) the smaller footprint is worth +0.7 % on the loop; 32 registers: 50 % slower
#ifndef DVC_FF_MINB
#define DVC_FF_MINB 6
#endif
template <bool ALIGNED>
__global__ void __launch_bounds__(256, DVC_FF_MINB)
k_fd_front(const uint8_t* __restrict__ frames, int T, int H, int W, const uint8_t* __restrict__ prev_gray_in,
           uint8_t* __restrict__ prev_gray_out, uint32_t* __restrict__ bits_out, int wpr, uint32_t thr, int seg_len, int nseg) {
    __shared__ __align__(16) uint8_t sg[FF_ROWS * FF_GP];            // gray, column c <-> x = x0 - FF_PAD + c
    __shared__ __align__(16) uint2 sh[FF_ROWS * (FF_TW / 4)];         // horizontal pass: (even, odd) 16-bit lane pairs per 4 px
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * FF_TW, y0 = blockIdx.y * FF_TH;
    const int stream = blockIdx.z / nseg, seg = blockIdx.z - stream * nseg;
    const int t0 = seg * seg_len, t1 = min(T, t0 + seg_len);
    const size_t frame_bytes = (size_t)H * W * 3, plane_bytes = (size_t)H * W, plane_words = (size_t)H * wpr;
    frames += (size_t)stream * T * frame_bytes;
    bits_out += (size_t)stream * T * plane_words;
    prev_gray_in += (size_t)stream * plane_bytes;
    if (prev_gray_out) prev_gray_out += (size_t)stream * plane_bytes;
    const int tw = min(FF_TW, W - x0);                                // valid columns of this tile (> 0 by construction)

    // ---- per-thread task tables (frame independent) ----
    // gray: 68 rows x 8 groups of 16 px = 544 tasks, three rounds
    int g_src[FF_GR], g_dst[FF_GR];
#pragma unroll
    for (int i = 0; i < FF_GR; ++i) {
        const int task = tid + 256 * i, r = task >> 3, g = task & 7, x = x0 + g * 16;
        g_src[i] = -1; g_dst[i] = 0;
        if (task < FF_ROWS * 8 && x < W) {
            g_src[i] = (reflect101(y0 - 2 + r, H) * W + x) * 3;       // frames are < 2^31 bytes (checked on the host)
            g_dst[i] = r * FF_GP + FF_PAD + g * 16;
        }
    }
    // halo columns: 68 rows x 4 = 272 tasks (left x0-2, x0-1; right: the two columns after the tile's last valid one)
    int h_src[FF_HR], h_dst[FF_HR];
#pragma unroll
    for (int i = 0; i < FF_HR; ++i) {
        const int task = tid + 256 * i;
        h_src[i] = -2; h_dst[i] = 0;
        if (task < FF_ROWS * 4) {
            const int r = task >> 2, k = task & 3;
            const int dx = k < 2 ? k - 2 : tw + (k - 2);
            const int xx = reflect101(x0 + dx, W);
            h_dst[i] = r * FF_GP + FF_PAD + dx;
            if (xx >= x0 && xx < x0 + tw) h_src[i] = -(r * FF_GP + FF_PAD + (xx - x0)) - 16;      // < -2: copy inside the tile
            else h_src[i] = (reflect101(y0 - 2 + r, H) * W + xx) * 3;
        }
    }
    const bool edge_tile = x0 < 2 || x0 + tw + 2 > W;                 // some halo column reflects into the tile (CTA-uniform)
    // vertical / output: thread = 8 px x 4 rows
    const int cgp = tid & 15, rg = tid >> 4;
    const int ox = x0 + cgp * 8, oy = y0 + rg * 4;
    const int lane = tid & 31;
    const bool vact = rg * 4 < FF_TH;                                 // the last 16 threads have no rows (they still take part in the shuffles)
    uint32_t pv[4][2];                                                // previous blurred values of this thread's pixels
#pragma unroll
    for (int i = 0; i < 4; ++i) { pv[i][0] = pv[i][1] = 0u; }
    if (t0 == 0 && vact) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int y = oy + i;
            if (y < H && ox < W) {
                const uint8_t* q = prev_gray_in + (size_t)y * W + ox;
                if (ALIGNED) { const uint2 v = *reinterpret_cast<const uint2*>(q); pv[i][0] = v.x; pv[i][1] = v.y; }
                else for (int b = 0; b < min(8, W - ox); ++b) pv[i][b >> 2] |= (uint32_t)q[b] << ((b & 3) * 8);
            }
        }
    }
    // words of the bit-plane: lanes 4k..4k+3 hold the four bytes of one 32-bit word; after the 4 x 4 byte transpose below
    // lane (lane & 3) = q owns the word of row q
    const int wj = (x0 >> 5) + (cgp >> 2);
    const uint32_t wvm = valid_mask(wj, W);

    for (int t = (t0 == 0 ? 0 : t0 - 1); t < t1; ++t) {
        const uint8_t* fr = frames + (size_t)t * frame_bytes;
        const bool emit = t >= t0;
        // phase 1a: interior gray
#pragma unroll
        for (int i = 0; i < FF_GR; ++i) {
            if (g_src[i] >= 0) {
                uint32_t w[12], gg[4];
                const uint8_t* row = fr + g_src[i];
                if constexpr (ALIGNED) {       // W % 16 == 0: every 16-px group inside the image is whole and 16-byte aligned
                    const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(row)), q1 = __ldg(reinterpret_cast<const uint4*>(row) + 1),
                                q2 = __ldg(reinterpret_cast<const uint4*>(row) + 2);
                    w[0] = q0.x; w[1] = q0.y; w[2] = q0.z; w[3] = q0.w; w[4] = q1.x; w[5] = q1.y; w[6] = q1.z; w[7] = q1.w;
                    w[8] = q2.x; w[9] = q2.y; w[10] = q2.z; w[11] = q2.w;
                } else {
                    const int x = x0 + ((tid + 256 * i) & 7) * 16;
                    load_bgr16_generic(row - (size_t)x * 3, x, W, w);
                }
                gray16_dp2a(w, gg);
                *reinterpret_cast<uint4*>(&sg[g_dst[i]]) = make_uint4(gg[0], gg[1], gg[2], gg[3]);
            }
        }
        // phase 1b: halo columns that come from the frame (the neighbouring tiles' pixels).  Widths that are a multiple of 16: same
        // phase as the interior, so that their loads are in flight together with the interior's.  Other widths: a clipped tile's
        // right halo columns lie inside its last 16-pixel interior group, so they are written after the interior (phase 1c).
        if (ALIGNED) {
#pragma unroll
            for (int i = 0; i < FF_HR; ++i) {
                if (h_src[i] >= 0) {
                    const uint8_t* p = fr + h_src[i];
                    sg[h_dst[i]] = (uint8_t)gray_of(p[0], p[1], p[2]);
                }
            }
        }
        __syncthreads();
        // phase 1c: halo columns reflected at the image border are copies of tile pixels (tiles at the left / right edge only)
        if (!ALIGNED || edge_tile) {
#pragma unroll
            for (int i = 0; i < FF_HR; ++i) {
                if (h_src[i] < -2) sg[h_dst[i]] = sg[-(h_src[i] + 16)];
                else if (!ALIGNED && h_src[i] >= 0) {
                    const uint8_t* p = fr + h_src[i];
                    sg[h_dst[i]] = (uint8_t)gray_of(p[0], p[1], p[2]);
                }
            }
            __syncthreads();
        }
        // phase 2: horizontal 5 taps on 16-bit lanes, 16 px (four gray words) per task, same 544-task table as phase 1a.  Every
        // word is widened once into even / odd pixel lanes (E = px0 | px2 << 16, O = px1 | px3 << 16); the shifted operands
        // of the taps are funnel shifts of neighbouring E / O words: even outputs E(-1) + E(+1) + 4 (O(-1) + O) + 6 E, odd
        // outputs O(-1) + O(+1) + 4 (E + E(+1)) + 6 O, where (-1) / (+1) is the lane-shifted word pair.  h <= 4080 per lane.
#pragma unroll
        for (int i = 0; i < FF_GR; ++i) {
            const int task = tid + 256 * i;
            if (task < FF_ROWS * 8) {
                const int r = task >> 3, g = task & 7;
                const uint32_t* gw = reinterpret_cast<const uint32_t*>(&sg[r * FF_GP + FF_PAD + g * 16]);
                const uint4 c4 = *reinterpret_cast<const uint4*>(gw);
                const uint32_t wv[6] = {gw[-1], c4.x, c4.y, c4.z, c4.w, gw[4]};
                uint32_t E[6], O[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) { E[k] = __byte_perm(wv[k], 0u, 0x4240u); O[k] = __byte_perm(wv[k], 0u, 0x4341u); }
                uint32_t hx[4], hy[4];
#pragma unroll
                for (int k = 1; k <= 4; ++k) {
                    const uint32_t Em = __funnelshift_r(E[k - 1], E[k], 16), Om = __funnelshift_r(O[k - 1], O[k], 16);
                    const uint32_t Ep = __funnelshift_r(E[k], E[k + 1], 16), Op = __funnelshift_r(O[k], O[k + 1], 16);
                    hx[k - 1] = (Em + Ep) + 4u * (Om + O[k]) + 6u * E[k];
                    hy[k - 1] = (Om + Op) + 4u * (E[k] + Ep) + 6u * O[k];
                }
                uint4* dst = reinterpret_cast<uint4*>(&sh[r * (FF_TW / 4) + g * 4]);
                dst[0] = make_uint4(hx[0], hy[0], hx[1], hy[1]);
                dst[1] = make_uint4(hx[2], hy[2], hx[3], hy[3]);
            }
        }
        __syncthreads();
        // phase 3: vertical 5 taps + rounding for 8 px x 4 rows, absdiff + threshold against the previous blurred values
        {
            const uint4* col = reinterpret_cast<const uint4*>(sh) + (rg * 4) * (FF_TW / 8) + cgp;
            uint32_t rv[4] = {0u, 0u, 0u, 0u};                        // 128 * (the 8 mask bits of row i)
            if (vact) {
            uint4 w0 = col[0], w1 = col[FF_TW / 8], w2 = col[2 * (FF_TW / 8)], w3 = col[3 * (FF_TW / 8)];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 w4 = col[(4 + i) * (FF_TW / 8)];
                const uint32_t ve0 = (w0.x + w4.x) + 4u * (w1.x + w3.x) + 6u * w2.x + 0x00800080u;
                const uint32_t vo0 = (w0.y + w4.y) + 4u * (w1.y + w3.y) + 6u * w2.y + 0x00800080u;
                const uint32_t ve1 = (w0.z + w4.z) + 4u * (w1.z + w3.z) + 6u * w2.z + 0x00800080u;
                const uint32_t vo1 = (w0.w + w4.w) + 4u * (w1.w + w3.w) + 6u * w2.w + 0x00800080u;
                // every 16-bit lane is < 2^16 (16 * 4080 + 128): the blurred pixel is the lane's high byte
                const uint32_t o0 = __byte_perm(ve0, vo0, 0x7351u), o1 = __byte_perm(ve1, vo1, 0x7351u);
                if (emit) rv[i] = gather8_x128(diff_gt_msb4(o0, pv[i][0], thr), diff_gt_msb4(o1, pv[i][1], thr));
                pv[i][0] = o0; pv[i][1] = o1;
                w0 = w1; w1 = w2; w2 = w3; w3 = w4;
            }
            }
            const uint32_t rows_bits = (rv[0] >> 7) | (rv[1] << 1) | (rv[2] << 9) | (rv[3] << 17);      // byte i = the mask bits of row i
            if (emit) {
                // 4 x 4 byte transpose over lanes 4k..4k+3: lane q ends with the 32-bit plane word of row q
                const uint32_t p1 = __shfl_xor_sync(0xffffffffu, rows_bits, 1);
                const uint32_t a1 = (lane & 1) ? __byte_perm(rows_bits, p1, 0x3715u) : __byte_perm(rows_bits, p1, 0x6240u);
                const uint32_t p2 = __shfl_xor_sync(0xffffffffu, a1, 2);
                const uint32_t word = (lane & 2) ? __byte_perm(a1, p2, 0x3276u) : __byte_perm(a1, p2, 0x5410u);
                const int y = oy + (lane & 3);
                if (vact && y < H && wj < wpr) bits_out[(size_t)t * plane_words + (size_t)y * wpr + wj] = word & wvm;
            }
        }
        // (the next frame's phase 1 writes sg, which nobody reads after the barrier above; its phase 2 writes sh only after
        //  two more barriers, by which time every thread has left phase 3)
    }
    if (t1 == T && prev_gray_out && vact) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int y = oy + i;
            if (y < H && ox < W) {
                uint8_t* q = prev_gray_out + (size_t)y * W + ox;
                if (ALIGNED) *reinterpret_cast<uint2*>(q) = make_uint2(pv[i][0], pv[i][1]);
                else for (int b = 0; b < min(8, W - ox); ++b) q[b] = (uint8_t)(pv[i][b >> 2] >> ((b & 3) * 8));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cv2.GaussianBlur on uint8, any odd ksize <= 33 and sigma (frame_differencing.py:77: the first frame's (25,25), sigma 30
// blur that seeds prev_gray).  OpenCV's uint8 path is fixed point: an 8.8 kernel made from the double-precision Gaussian
// by error-diffusion rounding so that it sums to exactly 256 (host: gaussian_taps_fixed), a horizontal pass into 8.8
// values, a vertical pass into 16.16, then (v + 2^15) >> 16, BORDER_REFLECT_101.  Runs once per stream: two plain passes.
// ------------------------------------------------------------------------------------------------
struct GaussTaps { int n; uint16_t k[33]; };

__global__ void __launch_bounds__(256)
k_gauss_h(const uint8_t* __restrict__ src, uint16_t* __restrict__ tmp, int H, int W, const __grid_constant__ GaussTaps tp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)H * W) return;
    const int y = (int)(i / W), x = (int)(i - (long long)y * W), r = tp.n >> 1;
    const uint8_t* row = src + (size_t)blockIdx.y * H * W + (size_t)y * W;
    uint32_t acc = 0;
    for (int j = 0; j < tp.n; ++j) acc += (uint32_t)tp.k[j] * row[reflect101(x - r + j, W)];
    tmp[(size_t)blockIdx.y * H * W + i] = (uint16_t)acc;                 // <= 255 * 256
}
__global__ void __launch_bounds__(256)
k_gauss_v(const uint16_t* __restrict__ tmp, uint8_t* __restrict__ dst, int H, int W, const __grid_constant__ GaussTaps tp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)H * W) return;
    const int y = (int)(i / W), x = (int)(i - (long long)y * W), r = tp.n >> 1;
    const uint16_t* pl = tmp + (size_t)blockIdx.y * H * W;
    uint32_t acc = 0;
    for (int j = 0; j < tp.n; ++j) acc += (uint32_t)tp.k[j] * pl[(size_t)reflect101(y - r + j, H) * W + x];
    dst[(size_t)blockIdx.y * H * W + i] = (uint8_t)min(255u, (acc + 32768u) >> 16);
}

// ------------------------------------------------------------------------------------------------
// u8 mask (0 / non-zero) <-> bit-plane, n images per launch (blockIdx.y).  One thread per 16 pixels: a 16-byte load,
// SWAR byte tests, multiply-gather of the four flag bits per word, one 16-bit store (and the reverse).
// ------------------------------------------------------------------------------------------------
DEVI uint32_t gather_msb4(uint32_t m) { return (((m >> 7) & 0x01010101u) * 0x00204081u >> 21) & 0xfu; }   // bit 8p+7 -> bit p
DEVI uint32_t nonzero_msb4(uint32_t x) { return (x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u; }   // MSB set iff byte != 0

template <bool ALIGNED, bool FLAGS>   // FLAGS: also emit the (> 127) plane
__global__ void __launch_bounds__(256)
k_pack_bits(const uint8_t* __restrict__ src, uint32_t* __restrict__ nonzero, uint32_t* __restrict__ over127, int H, int W,
            int wpr) {
    const int gpr = (W + 15) >> 4;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)gpr * H) return;
    const int y = (int)(gid / gpr), gx = (int)(gid % gpr), x0 = gx << 4;
    const uint8_t* row = src + (size_t)blockIdx.y * H * W + (size_t)y * W;
    uint32_t v[4];
    if (ALIGNED) load16(row + x0, v);
    else load_u8x16_generic(row, x0, W, v);
    uint32_t nz = 0, hi = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        nz |= gather_msb4(nonzero_msb4(v[q])) << (4 * q);
        if (FLAGS) hi |= gather_msb4(v[q] & 0x80808080u) << (4 * q);
    }
    const size_t wo = (size_t)blockIdx.y * H * wpr + (size_t)y * wpr;
    reinterpret_cast<uint16_t*>(nonzero + wo)[gx] = (uint16_t)nz;
    if (FLAGS) reinterpret_cast<uint16_t*>(over127 + wo)[gx] = (uint16_t)hi;
    // an odd number of 16-pixel groups leaves the upper half of the row's last used word: clear it
    if (gx == gpr - 1 && (gpr & 1)) {
        reinterpret_cast<uint16_t*>(nonzero + wo)[gx + 1] = 0;
        if (FLAGS) reinterpret_cast<uint16_t*>(over127 + wo)[gx + 1] = 0;
    }
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256)
k_unpack_bits(const uint32_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, int wpr) {
    const int gpr = (W + 15) >> 4;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)gpr * H) return;
    const int y = (int)(gid / gpr), gx = (int)(gid % gpr), x0 = gx << 4;
    const uint32_t bits = reinterpret_cast<const uint16_t*>(src + (size_t)blockIdx.y * H * wpr + (size_t)y * wpr)[gx];
    uint32_t v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t nib = (bits >> (4 * q)) & 0xfu;
        v[q] = ((nib * 0x00204081u) & 0x01010101u) * 0xffu;         // bit p -> byte p = 0xff
    }
    uint8_t* row = dst + (size_t)blockIdx.y * H * W + (size_t)y * W;
    if (ALIGNED) store16(row + x0, v);
    else store_u8x16_generic(row, x0, W, v);
}

// ------------------------------------------------------------------------------------------------
// cv2.resize(frame, (w, h)) with the default INTER_LINEAR on uint8 (frame_differencing.py:74,91): two-pass fixed
// point with 11-bit coefficients.  Host tables (make_resize_tables in dvc_b200.cu) hold, per output column, the left
// source column and the coefficient pair (a0, a1), per output row the top source row and (b0, b1), computed with the
// float / double expressions of OpenCV so the integers are the same.  Per output value:
//     h_r = S[r][x0] * a0 + S[r][x1] * a1              (r = the two source rows, clipped to the image)
//     out = (((b0 * (h_0 >> 4)) >> 16) + ((b1 * (h_1 >> 4)) >> 16) + 2) >> 2
// Checked bit for bit against cv2 for down- and upscales (oracle/stage_ops.py::resize_linear).
// One thread per output pixel (all channels); grid (ceil(dW * dH / 256), n).
// ------------------------------------------------------------------------------------------------
struct ResizeTables { const int* xofs; const short2* xa; const int* yofs; const short2* yb; };

template <int CN>
__global__ void __launch_bounds__(256)
k_resize_linear(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int sH, int sW, int dH, int dW, ResizeTables t) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= dW * dH) return;
    const int dy = gid / dW, dx = gid - dy * dW;
    const int x0 = t.xofs[dx], x1 = min(x0 + 1, sW - 1);
    const short2 a = t.xa[dx], b = t.yb[dy];
    const int sy = t.yofs[dy];
    const int y0 = min(max(sy, 0), sH - 1), y1 = min(max(sy + 1, 0), sH - 1);
    const uint8_t* s = src + (size_t)blockIdx.y * sH * sW * CN;
    const uint8_t* r0 = s + (size_t)y0 * sW * CN;
    const uint8_t* r1 = s + (size_t)y1 * sW * CN;
    uint8_t* o = dst + ((size_t)blockIdx.y * dH * dW + gid) * CN;
#pragma unroll
    for (int c = 0; c < CN; ++c) {
        const int h0 = (int)r0[x0 * CN + c] * a.x + (int)r0[x1 * CN + c] * a.y;
        const int h1 = (int)r1[x0 * CN + c] * a.x + (int)r1[x1 * CN + c] * a.y;
        const int v = (((b.x * (h0 >> 4)) >> 16) + ((b.y * (h1 >> 4)) >> 16) + 2) >> 2;
        o[c] = (uint8_t)min(255, max(0, v));
    }
}

}  // namespace dvc
