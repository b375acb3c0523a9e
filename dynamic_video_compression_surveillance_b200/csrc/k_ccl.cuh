// K-ccl: the contour filter of frame_differencing.py:100-104
//     contours = findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)
//     keep contours with contourArea > min_area, drawContours(FILLED)
// restated without contour tracing (oracle/stage_ops.py::contour_filter,
// pinned against cv2 in tests/test_oracle_vs_cv2.py):
//   phase A  O = background 4-connected to the outside of the image;  F = not O (blobs + their holes)
//   phase B  label F with 8-connectivity;  twice the polygon area of a label is 2*Q4 + Q3, the number
//            of 2x2 windows holding 4 / 3 of its pixels;  keep labels with 2*area > 2*min_area.
//
// Both phases are a union-find over *bit runs*: the graph nodes are the maximal runs of set bits inside
// each 32-bit word of the bit-plane, so an empty 1080p mask is 65 k nodes instead of 2 M pixels and a thread
// owns one word.  Node id = ((word index + 1) << 4) | slot, slot = ordinal of the run inside its word
// (a word holds at most 16 runs); id 0 is the "outside" node of phase A.  Slot-0 parents live in a dense
// array indexed by word (coalesced, L2 resident: 259 KB per 1080p frame); the rarely used slots 1..15 live in
// an overflow array.  Horizontal runs are linked by a warp-per-row kernel without atomics; vertical links are
// lock-free unions (atomicMin on the larger root) with path halving.
#pragma once
#include "common.cuh"

namespace dvc {

struct UF {
    int* base;     // dense part: base[0] = outside, base[1 + w] = slot 0 of word w
    int ov_off;    // offset (in ints, from base) of the overflow part [plane_words * 15]: slots 1..15
};

DEVI int node_id(int word, int slot) { return ((word + 1) << 4) | slot; }
// One pointer + a 32-bit offset: the dense and overflow arrays of a pair live in one allocation (see ccl_pair_alloc),
// so the address is a single IMAD.WIDE instead of two 64-bit multiply-adds and a pointer select.
DEVI int* uf_addr(const UF& u, int id) {
    const int s = id & 15, w = id >> 4;
    return u.base + (s == 0 ? w : u.ov_off + (w - 1) * 15 + (s - 1));
}
DEVI UF uf_of_frame(int* p0, int* pov, size_t plane_words, int frame) {
    UF u;
    u.base = p0 + (size_t)frame * (plane_words + 1);
    u.ov_off = (int)((pov + (size_t)frame * plane_words * 15) - u.base);
    return u;
}

DEVI int uf_find(const UF& u, int x) {
    while (true) {
        const int p = __ldcg(uf_addr(u, x));       // L2 reads: other SMs link roots with atomics at L2
        if (p == x) return x;
        const int gp = __ldcg(uf_addr(u, p));
        if (gp == p) return p;
        __stcg(uf_addr(u, x), gp);                 // path halving; parents only move to smaller ancestors of the set
        x = gp;
    }
}
// After the unions of a phase are complete (a later kernel), roots no longer change: a parent read from the SM's L1 may be
// stale, but it is still an ancestor, and a root still reads as its own parent.  The 60 threads of a row share their row
// head and its chain, so most steps hit L1 instead of making an L2 round trip.
DEVI int uf_find_settled(const UF& u, int x) {
    while (true) {
        const int p = *uf_addr(u, x);
        if (p == x) return x;
        const int gp = *uf_addr(u, p);
        if (gp == p) return p;
        __stcg(uf_addr(u, x), gp);
        x = gp;
    }
}
DEVI void uf_union(const UF& u, int a, int b) {
    while (true) {
        a = uf_find(u, a);
        b = uf_find(u, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(uf_addr(u, a), b);
        if (old == a) return;
        a = old;
    }
}

// lowest run of set bits of m: returns its mask, lo = first bit
DEVI uint32_t lowest_run(uint32_t m, int& lo) {
    lo = __ffs(m) - 1;
    const uint32_t t = m + (1u << lo);
    return m & ~t;
}
// first bit of the run of u that contains bit p (bit p must be set)
DEVI int run_start(uint32_t u, int p) {
    const uint32_t z = ~u & ((1u << p) - 1u);
    return z ? 32 - __clz(z) : 0;
}
DEVI uint32_t run_mask_from(uint32_t u, int start) { return u & ~(u + (1u << start)); }
DEVI uint32_t run_starts(uint32_t u) { return u & ~(u << 1); }
// ordinal of the run of u that contains bit p (bit p must be set)
DEVI int run_slot(uint32_t u, int p) { return __popc(run_starts(u) & (0xffffffffu >> (31 - p))) - 1; }

template <bool INVERT>
DEVI uint32_t plane_word(const uint32_t* plane, int y, int j, int H, int W, int wpr) {
    if (y < 0 || y >= H || j < 0 || j >= wpr) return 0u;
    const uint32_t w = plane[y * wpr + j];
    return INVERT ? (~w & valid_mask(j, W)) : w;
}

// ---- init + horizontal linking: one warp per image row ------------------------------------------------
// Every run inside a word becomes a node; a run that continues from the previous word (bit 31 of word j-1
// and bit 0 of word j both set) is pointed straight at the head of the whole horizontal run, found with
// warp ballots over "word is all ones and linked" -- no atomics, depth 1.  With BORDER (phase A) runs that
// touch the image border are rooted at node 0 ("outside") directly, so an empty mask needs no union at all.
template <bool INVERT, bool BORDER>
__global__ void __launch_bounds__(256)
k_ccl_rowlink(const uint32_t* __restrict__ planes, int* __restrict__ p0, int* __restrict__ pov,
              int* __restrict__ a0, int* __restrict__ aov, int H, int W, int wpr, uint8_t* __restrict__ rowflag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int y = blockIdx.x * 8 + warp;
    if (y >= H) return;
    // rowflag[frame][y] = the row holds foreground.  Phase A (INVERT) computes it; every later kernel skips rows without
    // foreground: their background is rooted at 'outside' by this kernel, nothing of them is filled, labelled or selected.
    if (!INVERT && rowflag && rowflag[(size_t)blockIdx.y * H + y] == 0) return;
    const size_t plane_words = (size_t)H * wpr;
    const uint32_t* plane = planes + (size_t)blockIdx.y * plane_words;
    const UF P = uf_of_frame(p0, pov, plane_words, blockIdx.y);
    const UF A = uf_of_frame(a0, aov, plane_words, blockIdx.y);
    if (BORDER && y == 0 && lane == 0) P.base[0] = 0;
    const bool edge_row = BORDER && (y == 0 || y == H - 1);
    const int last_word = (W - 1) >> 5, last_bit = (W - 1) & 31;
    int carry_root = -1;                      // root of the run leaving the previous chunk through bit 31 (-1: none)
    bool any_fg = false;
    for (int j0 = 0; j0 < wpr; j0 += 32) {
        const int j = j0 + lane;
        const uint32_t w = j < wpr ? plane_word<INVERT>(plane, y, j, H, W, wpr) : 0u;
        if (INVERT && rowflag) any_fg |= __any_sync(0xffffffffu, j < wpr && (~w & valid_mask(j, W)) != 0u);
        const int word = y * wpr + j;
        const int nruns = __popc(run_starts(w));
        const bool full = w == 0xffffffffu;
        const uint32_t w_prev = __shfl_up_sync(0xffffffffu, w, 1);
        const bool prev_msb = lane == 0 ? carry_root >= 0 : (w_prev >> 31) != 0;
        const bool link = (w & 1u) && prev_msb;
        const uint32_t T = __ballot_sync(0xffffffffu, full && link);
        // root of a run that starts inside this word and leaves it through bit 31 (used when !(full && link))
        int own_out = full ? ((BORDER && j == 0) ? 0 : node_id(word, 0)) : node_id(word, nruns - 1);
        if (edge_row) own_out = 0;
        const uint32_t z = ~T & ((1u << lane) - 1u);
        const int h = z ? 31 - __clz(z) : 0;
        const int src = __shfl_sync(0xffffffffu, own_out, h);
        int root_first;                        // root of the run containing bit 0 (when set)
        if (link) root_first = z ? src : carry_root;
        else root_first = (BORDER && j == 0) ? 0 : node_id(word, 0);
        if (edge_row) root_first = 0;
        // write the nodes of this word
        uint32_t m = w;
        int slot = 0;
        while (m) {
            int lo;
            const uint32_t run = lowest_run(m, lo);
            m &= ~run;
            const int id = node_id(word, slot);
            *uf_addr(P, id) = edge_row ? 0 : (lo == 0 ? root_first : id);
            if (a0) *uf_addr(A, id) = 0;
            ++slot;
        }
        const bool last_is_first = nruns == 1 && (w & 1u);     // the run through bit 31 is also the run through bit 0
        int out_root = -1;
        if (w >> 31) out_root = edge_row ? 0 : (last_is_first ? root_first : node_id(word, nruns - 1));
        __syncwarp();
        if (BORDER && !edge_row && j == last_word && ((w >> last_bit) & 1u)) {
            // the run holding the image's last column is outside: root its head at 0
            const int s = run_slot(w, last_bit);
            const int head = (s == 0 && (w & 1u)) ? root_first : node_id(word, s);
            if (head != 0) *uf_addr(P, head) = 0;
        }
        carry_root = __shfl_sync(0xffffffffu, out_root, 31);
    }
    if (INVERT && rowflag && lane == 0) rowflag[(size_t)blockIdx.y * H + y] = any_fg ? 1 : 0;
}

// thread -> (row y, word j, frame) without a division: CTAs of 64 words x 4 rows, grid (ceil(wpr / 64), ceil(H / 4), frames)
#define CCL_GRID(wpr, H, frames) dim3((unsigned)(((wpr) + 63) / 64), (unsigned)(((H) + 3) / 4), (unsigned)(frames))
#define CCL_THREAD_POS()                                                         \
    const int j = (int)blockIdx.x * 64 + (int)(threadIdx.x & 63u);               \
    const int y = (int)blockIdx.y * 4 + (int)(threadIdx.x >> 6);                 \
    if (j >= wpr || y >= H) return;                                              \
    const uint32_t idx = (uint32_t)y * (uint32_t)wpr + (uint32_t)j;              \
    const uint32_t frame = blockIdx.z;

// ---- unions with the row above (4- or 8-connected); horizontal links already exist ----------------------
// All vertical links of one word (y, j) with row y - 1.
template <bool INVERT, int CONN>
DEVI void ccl_union_word(const uint32_t* __restrict__ plane, const UF& P, int y, int j, int H, int W, int wpr) {
    const uint32_t cur = plane_word<INVERT>(plane, y, j, H, W, wpr);
    if (!cur) return;
    const uint32_t up = plane_word<INVERT>(plane, y - 1, j, H, W, wpr);
    const uint32_t upl = CONN == 8 ? plane_word<INVERT>(plane, y - 1, j - 1, H, W, wpr) : 0u;
    const uint32_t upr = CONN == 8 ? plane_word<INVERT>(plane, y - 1, j + 1, H, W, wpr) : 0u;
    if (!(up | (upl >> 31) | (upr & 1u))) return;
    const int word = y * wpr + j, word_up = word - wpr;
    // The run through bit 0 of this word and the run through bit 0 of the word above both continue from the words to
    // the left: the thread of word j-1 already links those two horizontal runs, so this pair is skipped.  Inside wide
    // regions only the left-most word of every row does a union.
    const uint32_t left = plane_word<INVERT>(plane, y, j - 1, H, W, wpr);
    const uint32_t up_left = CONN == 8 ? upl : plane_word<INVERT>(plane, y - 1, j - 1, H, W, wpr);
    const bool skip_first = (cur & 1u) && (up & 1u) && (left >> 31) && (up_left >> 31);
    uint32_t m = cur;
    int slot = 0;
    while (m) {
        int lo;
        const uint32_t run = lowest_run(m, lo);
        m &= ~run;
        const int id = node_id(word, slot++);
        const int hi = 31 - __clz(run);
        const int pa = __ldcg(uf_addr(P, id));               // quick test: same parent already (e.g. both outside)
        uint32_t nm = run;
        if (CONN == 8) nm |= (run << 1) | (run >> 1);
        uint32_t n = up & nm;
        if (lo == 0 && skip_first) n &= ~run_mask_from(up, 0);
        while (n) {
            const int p = __ffs(n) - 1;
            const int o = node_id(word_up, run_slot(up, p));
            if (__ldcg(uf_addr(P, o)) != pa) uf_union(P, id, o);
            n &= ~run_mask_from(up, run_start(up, p));
        }
        if (CONN == 8) {
            if (lo == 0 && (upl >> 31) && !skip_first) {
                const int o = node_id(word_up - 1, __popc(run_starts(upl)) - 1);
                if (__ldcg(uf_addr(P, o)) != pa) uf_union(P, id, o);
            }
            if (hi == 31 && (upr & 1u)) {
                const int o = node_id(word_up + 1, 0);
                if (__ldcg(uf_addr(P, o)) != pa) uf_union(P, id, o);
            }
        }
    }
}

template <bool INVERT, int CONN>
__global__ void __launch_bounds__(256)
k_ccl_union(const uint32_t* __restrict__ planes, int* __restrict__ p0, int* __restrict__ pov, int H, int W, int wpr,
            const uint8_t* __restrict__ rowflag) {
    const size_t plane_words = (size_t)H * wpr;
    CCL_THREAD_POS();
    (void)idx;
    if (y == 0) return;
    if (rowflag) {
        const uint8_t* f = rowflag + (size_t)frame * H;
        // background flood (INVERT): two foreground-free rows are both rooted at 'outside' already; foreground labelling: a row
        // without foreground has no nodes
        if (INVERT ? (f[y] == 0 && f[y - 1] == 0) : (f[y] == 0)) return;
    }
    const uint32_t* plane = planes + (size_t)frame * plane_words;
    const UF P = uf_of_frame(p0, pov, plane_words, frame);
    ccl_union_word<INVERT, CONN>(plane, P, y, j, H, W, wpr);
}

// ---- phase A result: F = complement of the background reachable from outside ---------------------
__global__ void __launch_bounds__(256)
k_ccl_fill(const uint32_t* __restrict__ planes, int* __restrict__ p0, int* __restrict__ pov,
           uint32_t* __restrict__ filled, int H, int W, int wpr, const uint8_t* __restrict__ rowflag) {
    const size_t plane_words = (size_t)H * wpr;
    CCL_THREAD_POS();
    if (rowflag && rowflag[(size_t)frame * H + y] == 0) {                  // no foreground in the row: all of it is outside
        filled[(size_t)frame * plane_words + idx] = 0u;
        return;
    }
    const UF P = uf_of_frame(p0, pov, plane_words, frame);
    uint32_t m = plane_word<true>(planes + (size_t)frame * plane_words, y, j, H, W, wpr);
    uint32_t outside = 0;
    int slot = 0;
    while (m) {
        int lo;
        const uint32_t run = lowest_run(m, lo);
        m &= ~run;
        if (uf_find_settled(P, node_id((int)idx, slot++)) == 0) outside |= run;
    }
    filled[(size_t)frame * plane_words + idx] = ~outside & valid_mask(j, W);
}

// ---- phase B: 2*area = 2*Q4 + Q3 per label --------------------------------------------------------
__global__ void __launch_bounds__(256)
k_ccl_area(const uint32_t* __restrict__ filled, int* __restrict__ p0, int* __restrict__ pov, int* __restrict__ a0,
           int* __restrict__ aov, int H, int W, int wpr, const uint8_t* __restrict__ rowflag) {
    const size_t plane_words = (size_t)H * wpr;
    CCL_THREAD_POS();
    if (rowflag) {
        const uint8_t* f = rowflag + (size_t)frame * H;
        if (f[y] == 0 && (y + 1 >= H || f[y + 1] == 0)) return;           // both rows of the 2x2 windows are empty
    }
    const uint32_t* F = filled + (size_t)frame * plane_words;
    const UF P = uf_of_frame(p0, pov, plane_words, frame);
    const UF A = uf_of_frame(a0, aov, plane_words, frame);
    const uint32_t a = plane_word<false>(F, y, j, H, W, wpr), b = plane_word<false>(F, y + 1, j, H, W, wpr);
    if (!(a | b)) return;
    const uint32_t an = plane_word<false>(F, y, j + 1, H, W, wpr), bn = plane_word<false>(F, y + 1, j + 1, H, W, wpr);
    const uint32_t a1 = (a >> 1) | (an << 31), b1 = (b >> 1) | (bn << 31);
    const uint32_t q4 = a & a1 & b & b1;
    const uint32_t q3a = a & ((a1 & b & ~b1) | (a1 & ~b & b1) | (~a1 & b & b1));   // top-left pixel in F
    const uint32_t q3b = ~a & a1 & b & b1;                                         // top-left pixel not in F
    if (!(q4 | q3a | q3b)) return;
    uint32_t m = a;
    int slot = 0;
    while (m) {
        int lo;
        const uint32_t run = lowest_run(m, lo);
        m &= ~run;
        const int c = 2 * __popc(q4 & run) + __popc(q3a & run);
        if (c) atomicAdd(uf_addr(A, uf_find_settled(P, node_id((int)idx, slot))), c);
        ++slot;
    }
    m = q3b ? b : 0u;
    slot = 0;
    while (m) {
        int lo;
        const uint32_t run = lowest_run(m, lo);
        m &= ~run;
        const int c = __popc(q3b & run);
        if (c) atomicAdd(uf_addr(A, uf_find_settled(P, node_id((int)idx + wpr, slot))), c);
        ++slot;
    }
}

__global__ void __launch_bounds__(256)
k_ccl_select(const uint32_t* __restrict__ filled, int* __restrict__ p0, int* __restrict__ pov, int* __restrict__ a0,
             int* __restrict__ aov, uint32_t* __restrict__ out, int H, int W, int wpr, int twice_min_area_floor,
             const uint8_t* __restrict__ rowflag) {
    const size_t plane_words = (size_t)H * wpr;
    CCL_THREAD_POS();
    (void)j;
    if (rowflag && rowflag[(size_t)frame * H + y] == 0) {
        out[(size_t)frame * plane_words + idx] = 0u;
        return;
    }
    const UF P = uf_of_frame(p0, pov, plane_words, frame);
    const UF A = uf_of_frame(a0, aov, plane_words, frame);
    uint32_t m = filled[(size_t)frame * plane_words + idx];
    uint32_t keep = 0;
    int slot = 0;
    while (m) {
        int lo;
        const uint32_t run = lowest_run(m, lo);
        m &= ~run;
        if (__ldcg(uf_addr(A, uf_find_settled(P, node_id((int)idx, slot++)))) > twice_min_area_floor) keep |= run;
    }
    out[(size_t)frame * plane_words + idx] = keep;
}

// ---- bounding rectangles of the 8-connected components (motion_compression_opt.py:93-97) ------------------
//     contours = findContours(mask, RETR_EXTERNAL); for c: x, y, w, h = boundingRect(c); rectangle((x, y), (x + w, y + h), 255, FILLED)
// Only outermost contours are returned, but a component nested in a hole of another lies inside that one's
// rectangle, so painting the rectangle of every 8-connected component gives the same image.  cv2.rectangle
// includes both corners: columns min_x .. max_x + 1 and rows min_y .. max_y + 1, clipped to the image.
// The root of a set is its smallest node id = its first run in raster order, so min_y is the root's own row;
// min_x, -max_x and -max_y are reduced with atomicMin into per-node arrays (same dense / overflow layout).
struct BBoxArrays { int* d[3]; int* ov[3]; };        // [0] min_x, [1] -max_x, [2] -max_y
DEVI UF bbox_view(const BBoxArrays& b, int k, size_t plane_words, int frame) { return uf_of_frame(b.d[k], b.ov[k], plane_words, frame); }

__global__ void __launch_bounds__(256)
k_ccl_bbox(const uint32_t* __restrict__ planes, int* __restrict__ p0, int* __restrict__ pov, BBoxArrays bb, int H, int W, int wpr) {
    const size_t plane_words = (size_t)H * wpr;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;            // a plane has far fewer than 2^31 words: 32-bit index math
    if (idx >= (uint32_t)plane_words) return;
    const int y = (int)(idx / (uint32_t)wpr), j = (int)(idx - (uint32_t)y * (uint32_t)wpr);
    const UF P = uf_of_frame(p0, pov, plane_words, blockIdx.y);
    const UF MINX = bbox_view(bb, 0, plane_words, blockIdx.y), NMAXX = bbox_view(bb, 1, plane_words, blockIdx.y),
             NMAXY = bbox_view(bb, 2, plane_words, blockIdx.y);
    uint32_t m = planes[(size_t)blockIdx.y * plane_words + idx];
    int slot = 0;
    while (m) {
        int lo;
        const uint32_t run = lowest_run(m, lo);
        m &= ~run;
        const int hi = 31 - __clz(run);
        const int root = uf_find_settled(P, node_id((int)idx, slot++));
        atomicMin(uf_addr(MINX, root), j * 32 + lo);
        atomicMin(uf_addr(NMAXX, root), -(j * 32 + hi));
        atomicMin(uf_addr(NMAXY, root), -y);
    }
}

// One warp paints one rectangle at a time: lanes that own a root publish it by ballot, all 32 lanes OR its words.
__global__ void __launch_bounds__(256)
k_ccl_paint_rects(const uint32_t* __restrict__ planes, int* __restrict__ p0, int* __restrict__ pov, BBoxArrays bb,
                  uint32_t* __restrict__ out, int H, int W, int wpr) {
    const size_t plane_words = (size_t)H * wpr;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in_range = idx < plane_words;
    const UF P = uf_of_frame(p0, pov, plane_words, blockIdx.y);
    const UF MINX = bbox_view(bb, 0, plane_words, blockIdx.y), NMAXX = bbox_view(bb, 1, plane_words, blockIdx.y),
             NMAXY = bbox_view(bb, 2, plane_words, blockIdx.y);
    uint32_t* o = out + (size_t)blockIdx.y * plane_words;
    uint32_t m = in_range ? planes[(size_t)blockIdx.y * plane_words + idx] : 0u;
    int slot = 0;
    while (__any_sync(0xffffffffu, m != 0u)) {
        int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
        if (m) {
            int lo;
            const uint32_t run = lowest_run(m, lo);
            m &= ~run;
            const int id = node_id((int)idx, slot++);
            if (__ldcg(uf_addr(P, id)) == id) {                 // a root: owns the rectangle of its component
                x0 = __ldcg(uf_addr(MINX, id));
                x1 = min(W - 1, -__ldcg(uf_addr(NMAXX, id)) + 1);
                y0 = (int)(idx / wpr);
                y1 = min(H - 1, -__ldcg(uf_addr(NMAXY, id)) + 1);
            }
        }
        uint32_t owners = __ballot_sync(0xffffffffu, x1 >= x0 && y1 >= y0);
        while (owners) {
            const int src = __ffs(owners) - 1;
            owners &= owners - 1;
            const int rx0 = __shfl_sync(0xffffffffu, x0, src), rx1 = __shfl_sync(0xffffffffu, x1, src);
            const int ry0 = __shfl_sync(0xffffffffu, y0, src), ry1 = __shfl_sync(0xffffffffu, y1, src);
            const int jw0 = rx0 >> 5, nw = (rx1 >> 5) - jw0 + 1, total = nw * (ry1 - ry0 + 1);
            for (int i = lane; i < total; i += 32) {
                const int r = i / nw, c = i - r * nw, jw = jw0 + c;
                const int b0 = max(rx0, jw * 32) - jw * 32, b1 = min(rx1, jw * 32 + 31) - jw * 32;
                const uint32_t bits = (0xffffffffu >> (31 - b1)) & (0xffffffffu << b0);
                atomicOr(o + (size_t)(ry0 + r) * wpr + jw, bits);
            }
        }
    }
}

}  // namespace dvc
