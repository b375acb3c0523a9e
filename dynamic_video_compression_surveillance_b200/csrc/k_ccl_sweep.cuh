// K-ccl, second generation (round 2): the contour filter of frame_differencing.py:100-104 in ONE launch, one CTA per frame.
//
// Same restatement as k_ccl.cuh (oracle/stage_ops.py::contour_filter):
//   phase A  O = background 4-connected to the outside of the image;  F = not O (blobs + their holes)
//   phase B  label F with 8-connectivity;  2 * area of a label = 2*Q4 + Q3;  keep labels with 2*area > 2*min_area.
// but built around what the data is: a bit-plane whose rows fit the registers of ONE warp (a lane holds NW 64-bit words,
// 32 lanes x 64 bits = 2048 pixels per NW), so every row operation is a handful of word-parallel instructions plus a shuffle
// or a ballot, and rows without foreground cost one load and one vote.
//
//   Both phases are a union-find over *row runs* (maximal horizontal runs of a row): phase A over the background runs of the rows
//   that hold foreground (node 0 = "outside": runs touching the image border or a foreground-free row are rooted there; what does
//   not end up in that set is a hole), phase B over the runs of F.  Strips of a few rows are loaded into registers in one go, their
//   runs counted, a contiguous range of node ids taken from a shared counter, and the strip is swept top-down: a run links to
//   runs of the row above whose parents are already (near) their roots, so trees stay shallow; the node arrays of a typical frame
//   (a few thousand runs) live in shared memory and a find is a couple of shared-memory reads instead of a chain of L2 round
//   trips.  Frames with more than SW_CAP runs redo the phase with node arrays in global scratch (same code, static id layout).
//   Links across strip boundaries are made after a barrier.  The 2x2-window counts of a row are taken in the phase B sweep (the
//   row below is in registers) and added to the run's node; one pass over the nodes moves them to the roots, a second one flags
//   the rows that hold a run of a too-small component, and only those rows are rewritten.
//
// tools/ccl_sweep_model.py is a lane-level model of exactly these steps with a configurable word width; it is checked against the
// oracle on tens of thousands of small masks (runs and neighbour bits on lane boundaries all the time).
#pragma once
#include "common.cuh"

namespace dvc {

#ifndef DVC_SW_WARPS
#define DVC_SW_WARPS 24
#endif
constexpr int SW_WARPS = DVC_SW_WARPS;               // warps of a frame's CTA (16 / 24 measured: 15.1 / 14.0 ms per 10 800 frames)
constexpr int SW_THREADS = SW_WARPS * 32;
constexpr int SW_CAP = 8192;                       // nodes (row runs) per frame held in shared memory
constexpr unsigned SW_FULL = 0xffffffffu;

DEVI uint64_t valid64(int d, int W) {
    const int rem = W - d * 64;
    return rem >= 64 ? ~0ull : (rem <= 0 ? 0ull : ((1ull << rem) - 1ull));
}
DEVI uint64_t lowest_run64(uint64_t m, int& lo) {
    lo = __ffsll((long long)m) - 1;
    return m & ~(m + (1ull << lo));
}
DEVI uint64_t lowmask64(int p) { return (2ull << p) - 1ull; }       // bits 0..p (p = 63: all)

template <int NW>
DEVI void sw_load_row(const uint32_t* row, int ndw, int lane, uint64_t (&r)[NW]) {
    const uint64_t* p = reinterpret_cast<const uint64_t*>(row);
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const int d = lane * NW + k;
        r[k] = d < ndw ? p[d] : 0ull;
    }
}
template <int NW>
DEVI void sw_store_row(uint32_t* row, int ndw, int lane, const uint64_t (&r)[NW]) {
    uint64_t* p = reinterpret_cast<uint64_t*>(row);
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const int d = lane * NW + k;
        if (d < ndw) p[d] = r[k];
    }
}

// ---- union-find over row runs --------------------------------------------------------------------------------------------------------
template <bool G> DEVI int sw_ld(const int* p) { return G ? __ldcg(p) : *(const volatile int*)p; }
template <bool G> DEVI void sw_st(int* p, int v) { if (G) __stcg(p, v); else *(volatile int*)p = v; }

template <bool G>
DEVI int sw_find(int* P, int x) {
    while (true) {
        const int p = sw_ld<G>(P + x);
        if (p == x) return x;
        const int gp = sw_ld<G>(P + p);
        if (gp == p) return p;
        sw_st<G>(P + x, gp);                       // path halving; parents only move to smaller ancestors of the set
        x = gp;
    }
}
template <bool G>
DEVI void sw_union(int* P, int a, int b) {
    while (true) {
        a = sw_find<G>(P, a);
        b = sw_find<G>(P, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(P + a, b);
        if (old == a) return;
        a = old;
    }
}

// bit 63 of the word to the left / bit 0 of the word to the right of every word of a row (0 outside the row)
template <int NW>
DEVI void sw_neighbour_bits(const uint64_t (&x)[NW], int lane, unsigned (&lb)[NW], unsigned (&rb)[NW]) {
    unsigned l = __shfl_up_sync(SW_FULL, (unsigned)(x[NW - 1] >> 63), 1);
    unsigned r = __shfl_down_sync(SW_FULL, (unsigned)(x[0] & 1ull), 1);
    if (lane == 0) l = 0;
    if (lane == 31) r = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        lb[k] = k ? (unsigned)(x[k - 1] >> 63) : l;
        rb[k] = k + 1 < NW ? (unsigned)(x[k + 1] & 1ull) : r;
    }
}

// run starts of every word, the number of run starts of the row before each word, the row's run count
template <int NW>
struct RowMeta {
    uint64_t st[NW];
    int pre[NW];
    unsigned lb[NW], rb[NW];
};
template <int NW>
DEVI int sw_row_meta(const uint64_t (&c)[NW], int lane, RowMeta<NW>& m) {
    sw_neighbour_bits<NW>(c, lane, m.lb, m.rb);
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        m.st[k] = c[k] & ~((c[k] << 1) | (uint64_t)m.lb[k]);
        m.pre[k] = cnt;
        cnt += __popcll(m.st[k]);
    }
    // exclusive prefix of cnt over the lanes, one ballot per bit of the counts (most rows: one or two bits); the ballots do not
    // depend on each other, a shuffle scan is five dependent steps
    const unsigned bits_used = __reduce_or_sync(SW_FULL, (unsigned)cnt);
    const unsigned below = (1u << lane) - 1u;
    int ex = 0, total = 0;
    for (int b = 0; (bits_used >> b) != 0u; ++b) {
        const unsigned mb = __ballot_sync(SW_FULL, (cnt >> b) & 1);
        ex += __popc(mb & below) << b;
        total += __popc(mb) << b;
    }
#pragma unroll
    for (int k = 0; k < NW; ++k) m.pre[k] += ex;
    return total;
}
template <int NW>
DEVI int sw_node_of(const RowMeta<NW>& m, int k, int base, int bit) {
    return base + m.pre[k] + __popcll(m.st[k] & lowmask64(bit)) - 1;
}

// all 8-connected links between the runs of row y (c) and of row y - 1 (u): tools/ccl_sweep_model.py::union_row
template <int NW, bool G>
DEVI void sw_union_row(int* P, const uint64_t (&c)[NW], const RowMeta<NW>& cm, int cbase, const uint64_t (&u)[NW],
                       const RowMeta<NW>& um, int ubase) {
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const uint64_t cw = c[k], uw = u[k];
        if (cw == 0ull) continue;
        const unsigned cl = cm.lb[k], ul = um.lb[k], ur = um.rb[k];
        const uint64_t ud = uw | (uw << 1) | (uw >> 1) | (uint64_t)ul | ((uint64_t)ur << 63);
        uint64_t I = cw & ud;
        while (I) {
            int a;
            const uint64_t run = lowest_run64(I, a);
            I &= ~run;
            const int b = a + __popcll(run) - 1;
            const int cnode = sw_node_of<NW>(cm, k, cbase, a);
            uint64_t E = uw & (run | (run << 1) | (run >> 1));
            // both runs continue from the word to the left: that word links them
            if (a == 0 && cl && (uw & 1ull) && ul) E &= ~(uw & ~(uw + 1ull));
            if (a == 0 && ul && !(uw & 1ull)) sw_union<G>(P, cnode, ubase + um.pre[k] - 1);
            if (b == 63 && ur && !(uw >> 63)) sw_union<G>(P, cnode, ubase + um.pre[k] + __popcll(um.st[k]));
            while (E) {
                int p;
                const uint64_t er = lowest_run64(E, p);
                E &= ~er;
                sw_union<G>(P, cnode, sw_node_of<NW>(um, k, ubase, p));
            }
        }
    }
}

// 2x2 windows whose top row is row y (a), bottom row y + 1 (b): 2*Q4 + Q3 added to the node of the row-y run that holds the
// window's top-left pixel (or its top-right pixel when the top-left one is not in F): tools/ccl_sweep_model.py::area_row
template <int NW>
DEVI void sw_area_row(int* A, const uint64_t (&a)[NW], const RowMeta<NW>& am, int abase, const uint64_t (&b)[NW], int lane) {
    unsigned bl[NW], br[NW];
    sw_neighbour_bits<NW>(b, lane, bl, br);
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const uint64_t aw = a[k], bw = b[k];
        if (aw == 0ull) continue;
        const uint64_t a1 = (aw >> 1) | ((uint64_t)am.rb[k] << 63), b1 = (bw >> 1) | ((uint64_t)br[k] << 63);
        const uint64_t q4 = aw & a1 & bw & b1;
        const uint64_t q3a = aw & ((a1 & bw & ~b1) | (a1 & ~bw & b1) | (~a1 & bw & b1));
        const uint64_t q3b = ~aw & a1 & bw & b1;
        const uint64_t carry = (!am.lb[k] && (aw & 1ull) && bl[k] && (bw & 1ull)) ? 1ull : 0ull;
        const uint64_t q3s = (q3b << 1) | carry;
        if (!(q4 | q3a | q3s)) continue;
        uint64_t m = aw;
        while (m) {
            int lo;
            const uint64_t run = lowest_run64(m, lo);
            m &= ~run;
            const int w = 2 * __popcll(q4 & run) + __popcll(q3a & run) + __popcll(q3s & run);
            if (w) atomicAdd(A + sw_node_of<NW>(am, k, abase, lo), w);
        }
    }
}

template <int NW, bool G>
DEVI void sw_init_nodes(int* P, int* A, const RowMeta<NW>& m, int base) {
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const int n = __popcll(m.st[k]);
        for (int i = 0; i < n; ++i) {
            const int id = base + m.pre[k] + i;
            sw_st<G>(P + id, id);
            sw_st<G>(A + id, 0);
        }
    }
}

// all 4-connected links between the background runs of row y (c) and of row y - 1 (u): one per maximal run of c & u, made by
// the lane in which that run starts (tools/ccl_sweep_model.py::fill_holes_uf)
template <int NW, bool G>
DEVI void sw_union_row4(int* P, const uint64_t (&c)[NW], const RowMeta<NW>& cm, int cbase, const uint64_t (&u)[NW],
                        const RowMeta<NW>& um, int ubase) {
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const uint64_t I = c[k] & u[k];
        uint64_t st = I & ~((I << 1) | (uint64_t)(cm.lb[k] & um.lb[k]));
        while (st) {
            const int p = __ffsll((long long)st) - 1;
            st &= st - 1ull;
            sw_union<G>(P, sw_node_of<NW>(cm, k, cbase, p), sw_node_of<NW>(um, k, ubase, p));
        }
    }
}
// every run of a row reaches the outside (the row above or below holds no foreground)
template <int NW, bool G>
DEVI void sw_all_outside(int* P, const RowMeta<NW>& m, int base) {
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        const int n = __popcll(m.st[k]);
        for (int i = 0; i < n; ++i) sw_union<G>(P, base + m.pre[k] + i, 0);
    }
}
// background runs of a row that are not in the outside set
template <int NW, bool G>
DEVI void sw_holes(int* P, const uint64_t (&b)[NW], const RowMeta<NW>& m, int base, uint64_t (&holes)[NW]) {
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        uint64_t w = b[k];
        holes[k] = 0ull;
        while (w) {
            int lo;
            const uint64_t run = lowest_run64(w, lo);
            w &= ~run;
            if (sw_find<G>(P, sw_node_of<NW>(m, k, base, lo)) != 0) holes[k] |= run;
        }
    }
}

constexpr int SW_PHASES = 4;

struct SwShared {
    int* rowbase;             // [H]  first node id of the row (phase A, then phase B); with the global arrays: its run count
    int* P;                   // [SW_CAP]
    int* A;                   // [SW_CAP]
    unsigned short* nrow;     // [SW_CAP] row of a phase B node
    volatile uint8_t* rowst;  // [H]  bit 0: the row holds foreground; bit 1: it holds a run of a too-small component
    int* counter;             // next free node id
    int* overflow;            // a row did not get its ids: redo the phase with the global arrays
    int* next;                // [SW_PHASES] next chunk of rows to hand out
    int chunk;                // consecutive rows a warp sweeps in one go
};

// chunks of rows are handed out dynamically: rows with foreground cluster, a static split leaves most warps at the barrier
DEVI int sw_next_chunk(int* ctr, int lane) {
    int c = 0;
    if (lane == 0) c = atomicAdd(ctr, 1);
    return __shfl_sync(SW_FULL, c, 0);
}
// node ids of a row: a range from the shared counter, or the row's fixed range of the global arrays
// Phase A has no use for the area and row arrays that follow P in shared memory: its parents may run over them.
constexpr int SW_CAP_A = SW_CAP * 10 / 4;
template <bool G, int CAP>
DEVI int sw_alloc(const SwShared& sh, int y, int cnt, int maxr, int lane, bool& ok) {
    if (G) {
        if (lane == 0) sh.rowbase[y] = cnt;
        return 1 + y * maxr;
    }
    int b = 0;
    if (lane == 0 && cnt) b = atomicAdd(sh.counter, cnt);
    b = __shfl_sync(SW_FULL, b, 0);
    if (b + cnt > CAP) {
        if (lane == 0) *sh.overflow = 1;
        ok = false;
    }
    if (lane == 0) sh.rowbase[y] = b;
    return b;
}
template <bool G> DEVI int sw_base_of(const SwShared& sh, int y, int maxr) { return G ? 1 + y * maxr : sh.rowbase[y]; }

// ---- phase A ----------------------------------------------------------------------------------------------------------------
// rows with foreground get nodes for their background runs, linked to the row above (same chunk)
template <int NW, bool G>
DEVI void sw_phase_a(const SwShared& sh, int* P, const uint32_t* __restrict__ M, uint32_t* out, int H, int W, int wpr,
                     int ndw, int maxr, const uint64_t (&vm)[NW]) {
    const int lane = threadIdx.x & 31;
    const int nchunks = (H + sh.chunk - 1) / sh.chunk;
    const int last_d = (W - 1) >> 6, last_bit = (W - 1) & 63;
    uint64_t zero[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) zero[k] = 0ull;
    for (int c = sw_next_chunk(sh.next + 0, lane); c < nchunks; c = sw_next_chunk(sh.next + 0, lane)) {
        const int ya = c * sh.chunk, yb = min(H, ya + sh.chunk);
        uint64_t cur[NW], nxt[NW], u[NW];
        RowMeta<NW> um;
        int ub = 0;
        bool have_u = false, ok = true;
        sw_load_row<NW>(M + (size_t)ya * wpr, ndw, lane, cur);
#pragma unroll 1
        for (int y = ya; y < yb && ok; ++y) {
            if (y + 1 < yb) sw_load_row<NW>(M + (size_t)(y + 1) * wpr, ndw, lane, nxt);
            uint64_t any = 0ull;
#pragma unroll
            for (int k = 0; k < NW; ++k) any |= cur[k];
            const bool fg = __any_sync(SW_FULL, any != 0ull);
            if (lane == 0) sh.rowst[y] = fg ? 1 : 0;
            if (!fg) {
                sw_store_row<NW>(out + (size_t)y * wpr, ndw, lane, zero);
                if (have_u) sw_all_outside<NW, G>(P, um, ub);
                have_u = false;
            } else {
                uint64_t b[NW];
#pragma unroll
                for (int k = 0; k < NW; ++k) b[k] = ~cur[k] & vm[k];
                RowMeta<NW> cm;
                const int cnt = sw_row_meta<NW>(b, lane, cm);
                const int rb = sw_alloc<G, SW_CAP_A>(sh, y, cnt, maxr, lane, ok);
                if (ok) {
                    const bool all_out = y == 0 || y == H - 1 || (y > ya && !have_u);
#pragma unroll
                    for (int k = 0; k < NW; ++k) {
                        const int n = __popcll(cm.st[k]);
                        for (int j = 0; j < n; ++j) {
                            const int id = rb + cm.pre[k] + j;
                            // the run that starts at column 0 is outside
                            sw_st<G>(P + id, (all_out || (lane == 0 && k == 0 && j == 0 && (b[0] & 1ull))) ? 0 : id);
                        }
                    }
                    __syncwarp();
                    if (!all_out) {
#pragma unroll
                        for (int k = 0; k < NW; ++k)
                            if (lane * NW + k == last_d && ((b[k] >> last_bit) & 1ull))      // the run that holds the last column
                                sw_union<G>(P, sw_node_of<NW>(cm, k, rb, last_bit), 0);
                    }
                    if (have_u) sw_union_row4<NW, G>(P, b, cm, rb, u, um, ub);
                    um = cm;
                    ub = rb;
#pragma unroll
                    for (int k = 0; k < NW; ++k) u[k] = b[k];
                    have_u = true;
                }
            }
#pragma unroll
            for (int k = 0; k < NW; ++k) cur[k] = nxt[k];
        }
    }
}
// links across chunk boundaries, and rows whose neighbour in the other chunk holds no foreground
template <int NW, bool G>
DEVI void sw_seams_a(const SwShared& sh, int* P, const uint32_t* __restrict__ M, int H, int wpr, int ndw, int maxr,
                     const uint64_t (&vm)[NW]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = (H + sh.chunk - 1) / sh.chunk;
    for (int c = 1 + warp; c < nchunks; c += SW_WARPS) {
        const int yb = c * sh.chunk;
        const bool up = sh.rowst[yb - 1] & 1, cu = sh.rowst[yb] & 1;
        if (!up && !cu) continue;
        uint64_t u[NW], w[NW];
        RowMeta<NW> um, cm;
        if (up) {
            sw_load_row<NW>(M + (size_t)(yb - 1) * wpr, ndw, lane, u);
#pragma unroll
            for (int k = 0; k < NW; ++k) u[k] = ~u[k] & vm[k];
            sw_row_meta<NW>(u, lane, um);
        }
        if (cu) {
            sw_load_row<NW>(M + (size_t)yb * wpr, ndw, lane, w);
#pragma unroll
            for (int k = 0; k < NW; ++k) w[k] = ~w[k] & vm[k];
            sw_row_meta<NW>(w, lane, cm);
        }
        if (up && cu) sw_union_row4<NW, G>(P, w, cm, sw_base_of<G>(sh, yb, maxr), u, um, sw_base_of<G>(sh, yb - 1, maxr));
        else if (up) sw_all_outside<NW, G>(P, um, sw_base_of<G>(sh, yb - 1, maxr));
        else sw_all_outside<NW, G>(P, cm, sw_base_of<G>(sh, yb, maxr));
    }
}
// F = foreground + holes, written to the output plane (a row of nothing but foreground has no background run and no node: F = M)
template <int NW, bool G>
DEVI void sw_fill(const SwShared& sh, int* P, const uint32_t* __restrict__ M, uint32_t* out, int H, int wpr, int ndw,
                  int maxr, const uint64_t (&vm)[NW]) {
    const int lane = threadIdx.x & 31;
    const int nchunks = (H + sh.chunk - 1) / sh.chunk;
    for (int c = sw_next_chunk(sh.next + 1, lane); c < nchunks; c = sw_next_chunk(sh.next + 1, lane)) {
        const int ya = c * sh.chunk, yb = min(H, ya + sh.chunk);
        uint64_t cur[NW], nxt[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) cur[k] = nxt[k] = 0ull;
        if (sh.rowst[ya] & 1) sw_load_row<NW>(M + (size_t)ya * wpr, ndw, lane, cur);
#pragma unroll 1
        for (int y = ya; y < yb; ++y) {
            if (y + 1 < yb && (sh.rowst[y + 1] & 1)) sw_load_row<NW>(M + (size_t)(y + 1) * wpr, ndw, lane, nxt);
            if (sh.rowst[y] & 1) {
                uint64_t b[NW], holes[NW];
#pragma unroll
                for (int k = 0; k < NW; ++k) b[k] = ~cur[k] & vm[k];
                RowMeta<NW> m;
                sw_row_meta<NW>(b, lane, m);
                sw_holes<NW, G>(P, b, m, sw_base_of<G>(sh, y, maxr), holes);
#pragma unroll
                for (int k = 0; k < NW; ++k) holes[k] |= cur[k];
                sw_store_row<NW>(out + (size_t)y * wpr, ndw, lane, holes);
            }
#pragma unroll
            for (int k = 0; k < NW; ++k) cur[k] = nxt[k];
        }
    }
}

// ---- phase B ----------------------------------------------------------------------------------------------------------------
template <int NW, bool G>
DEVI void sw_phase_b(const SwShared& sh, int* P, int* A, const uint32_t* out, int H, int wpr, int ndw, int maxr) {
    const int lane = threadIdx.x & 31;
    const int nchunks = (H + sh.chunk - 1) / sh.chunk;
    for (int c = sw_next_chunk(sh.next + 2, lane); c < nchunks; c = sw_next_chunk(sh.next + 2, lane)) {
        const int ya = c * sh.chunk, yb = min(H, ya + sh.chunk);
        uint64_t cur[NW], nxt[NW], u[NW];
        RowMeta<NW> um;
        int ub = 0;
        bool have_u = false, ok = true;
#pragma unroll
        for (int k = 0; k < NW; ++k) cur[k] = nxt[k] = 0ull;
        if (sh.rowst[ya] & 1) sw_load_row<NW>(out + (size_t)ya * wpr, ndw, lane, cur);
#pragma unroll 1
        for (int y = ya; y < yb && ok; ++y) {
            // the row below, also past the end of the chunk: the 2x2 windows of row y need it (rows without foreground are zero)
#pragma unroll
            for (int k = 0; k < NW; ++k) nxt[k] = 0ull;
            if (y + 1 < H && (sh.rowst[y + 1] & 1)) sw_load_row<NW>(out + (size_t)(y + 1) * wpr, ndw, lane, nxt);
            if (!(sh.rowst[y] & 1)) {
                have_u = false;
            } else {
                RowMeta<NW> cm;
                const int cnt = sw_row_meta<NW>(cur, lane, cm);
                const int rb = sw_alloc<G, SW_CAP>(sh, y, cnt, maxr, lane, ok);
                if (ok) {
#pragma unroll
                    for (int k = 0; k < NW; ++k) {
                        const int n = __popcll(cm.st[k]);
                        for (int j = 0; j < n; ++j) {
                            const int id = rb + cm.pre[k] + j;
                            sw_st<G>(P + id, id);
                            sw_st<G>(A + id, 0);
                            if (!G) sh.nrow[id] = (unsigned short)y;
                        }
                    }
                    __syncwarp();
                    if (have_u) sw_union_row<NW, G>(P, cur, cm, rb, u, um, ub);
                    sw_area_row<NW>(A, cur, cm, rb, nxt, lane);
                    um = cm;
                    ub = rb;
#pragma unroll
                    for (int k = 0; k < NW; ++k) u[k] = cur[k];
                    have_u = true;
                }
            }
#pragma unroll
            for (int k = 0; k < NW; ++k) cur[k] = nxt[k];
        }
    }
}
template <int NW, bool G>
DEVI void sw_seams_b(const SwShared& sh, int* P, const uint32_t* out, int H, int wpr, int ndw, int maxr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = (H + sh.chunk - 1) / sh.chunk;
    for (int c = 1 + warp; c < nchunks; c += SW_WARPS) {
        const int yb = c * sh.chunk;
        if (!(sh.rowst[yb - 1] & 1) || !(sh.rowst[yb] & 1)) continue;
        uint64_t u[NW], w[NW];
        RowMeta<NW> um, cm;
        sw_load_row<NW>(out + (size_t)(yb - 1) * wpr, ndw, lane, u);
        sw_load_row<NW>(out + (size_t)yb * wpr, ndw, lane, w);
        sw_row_meta<NW>(u, lane, um);
        sw_row_meta<NW>(w, lane, cm);
        sw_union_row<NW, G>(P, w, cm, sw_base_of<G>(sh, yb, maxr), u, um, sw_base_of<G>(sh, yb - 1, maxr));
    }
}
// window counts to the roots; rows that hold a run of a component that is not kept; those rows rewritten
template <int NW, bool G>
DEVI void sw_select(const SwShared& sh, int* P, int* A, uint32_t* out, int H, int wpr, int ndw, int maxr, int thr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (G) {
        for (int y = warp; y < H; y += SW_WARPS) {
            if (!(sh.rowst[y] & 1)) continue;
            const int b0 = 1 + y * maxr, b1 = b0 + sh.rowbase[y];
            for (int i = b0 + lane; i < b1; i += 32) {
                const int r = sw_find<G>(P, i);
                if (r != i) {
                    const int a = sw_ld<G>(A + i);
                    if (a) atomicAdd(A + r, a);
                }
            }
        }
    } else {
        const int n = min(*sh.counter, SW_CAP);
        for (int i = 1 + threadIdx.x; i < n; i += SW_THREADS) {
            const int r = sw_find<G>(P, i);
            if (r != i) {
                const int a = sw_ld<G>(A + i);
                if (a) atomicAdd(A + r, a);
            }
        }
    }
    __syncthreads();
    bool fix = false;
    if (G) {
        for (int y = warp; y < H; y += SW_WARPS) {
            if (!(sh.rowst[y] & 1)) continue;
            const int b0 = 1 + y * maxr, b1 = b0 + sh.rowbase[y];
            bool small = false;
            for (int i = b0 + lane; i < b1; i += 32) small = small || sw_ld<G>(A + sw_find<G>(P, i)) <= thr;
            if (__any_sync(SW_FULL, small)) {
                if (lane == 0) sh.rowst[y] = 3;
                fix = true;
            }
        }
    } else {
        const int n = min(*sh.counter, SW_CAP);
        for (int i = 1 + threadIdx.x; i < n; i += SW_THREADS) {
            if (sw_ld<G>(A + sw_find<G>(P, i)) <= thr) {
                sh.rowst[sh.nrow[i]] = 3;
                fix = true;
            }
        }
    }
    if (!__syncthreads_or(fix ? 1 : 0)) return;
    for (int y = warp; y < H; y += SW_WARPS) {
        if (sh.rowst[y] != 3) continue;
        uint64_t f[NW], keep[NW];
        RowMeta<NW> m;
        sw_load_row<NW>(out + (size_t)y * wpr, ndw, lane, f);
        sw_row_meta<NW>(f, lane, m);
        const int base = sw_base_of<G>(sh, y, maxr);
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            uint64_t w = f[k];
            keep[k] = 0ull;
            while (w) {
                int lo;
                const uint64_t run = lowest_run64(w, lo);
                w &= ~run;
                if (sw_ld<G>(A + sw_find<G>(P, sw_node_of<NW>(m, k, base, lo))) > thr) keep[k] |= run;
            }
        }
        sw_store_row<NW>(out + (size_t)y * wpr, ndw, lane, keep);
    }
}

template <int NW>
__global__ void __launch_bounds__(SW_THREADS, 2)
k_ccl_sweep(const uint32_t* __restrict__ planes, uint32_t* __restrict__ outp, int* __restrict__ gP, int* __restrict__ gA,
            size_t g_stride, int H, int W, int wpr, int thr, int chunk_rows) {
    extern __shared__ __align__(16) unsigned char sw_smem[];
    __shared__ int s_counter, s_overflow, s_next[SW_PHASES];
    SwShared sh;
    sh.rowbase = reinterpret_cast<int*>(sw_smem);
    sh.P = sh.rowbase + ((H + 3) & ~3);
    sh.A = sh.P + SW_CAP;
    sh.nrow = reinterpret_cast<unsigned short*>(sh.A + SW_CAP);
    sh.rowst = reinterpret_cast<volatile uint8_t*>(sh.nrow + SW_CAP);
    sh.counter = &s_counter;
    sh.overflow = &s_overflow;
    sh.next = s_next;
    sh.chunk = chunk_rows;

    const int lane = threadIdx.x & 31;
    const int frame = blockIdx.x;
    const size_t pw = (size_t)H * wpr;
    const uint32_t* M = planes + (size_t)frame * pw;
    uint32_t* out = outp + (size_t)frame * pw;
    int* GP = gP + (size_t)frame * g_stride;
    int* GA = gA + (size_t)frame * g_stride;
    const int ndw = wpr >> 1;
    const int maxr = (W + 1) / 2 + 1;                  // runs (of either kind) a row can hold

    uint64_t vm[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) vm[k] = valid64(lane * NW + k, W);

    if (threadIdx.x == 0) { s_counter = 1; s_overflow = 0; sh.P[0] = 0; }
    if (threadIdx.x < SW_PHASES) s_next[threadIdx.x] = 0;
    __syncthreads();

    // ---- phase A: holes ----
    sw_phase_a<NW, false>(sh, sh.P, M, out, H, W, wpr, ndw, maxr, vm);
    __syncthreads();
    if (!s_overflow) {
        sw_seams_a<NW, false>(sh, sh.P, M, H, wpr, ndw, maxr, vm);
        __syncthreads();
        sw_fill<NW, false>(sh, sh.P, M, out, H, wpr, ndw, maxr, vm);
    } else {
        if (threadIdx.x == 0) { __stcg(GP, 0); s_next[0] = 0; }
        __syncthreads();
        sw_phase_a<NW, true>(sh, GP, M, out, H, W, wpr, ndw, maxr, vm);
        __syncthreads();
        sw_seams_a<NW, true>(sh, GP, M, H, wpr, ndw, maxr, vm);
        __syncthreads();
        sw_fill<NW, true>(sh, GP, M, out, H, wpr, ndw, maxr, vm);
    }
    __syncthreads();
    if (threadIdx.x == 0) { s_counter = 1; s_overflow = 0; }
    __syncthreads();

    // ---- phase B: components of F, their areas, selection ----
    sw_phase_b<NW, false>(sh, sh.P, sh.A, out, H, wpr, ndw, maxr);
    __syncthreads();
    if (!s_overflow) {
        sw_seams_b<NW, false>(sh, sh.P, out, H, wpr, ndw, maxr);
        __syncthreads();
        sw_select<NW, false>(sh, sh.P, sh.A, out, H, wpr, ndw, maxr, thr);
    } else {
        if (threadIdx.x == 0) s_next[2] = 0;
        __syncthreads();
        sw_phase_b<NW, true>(sh, GP, GA, out, H, wpr, ndw, maxr);
        __syncthreads();
        sw_seams_b<NW, true>(sh, GP, out, H, wpr, ndw, maxr);
        __syncthreads();
        sw_select<NW, true>(sh, GP, GA, out, H, wpr, ndw, maxr, thr);
    }
}

static inline size_t ccl_sweep_smem_bytes(int H, int W) {
    (void)W;
    return (size_t)((H + 3) & ~3) * 4 + (size_t)SW_CAP * 10 + (size_t)((H + 15) & ~15);
}
// node ids of the global fallback: 1 + y * maxr + ordinal of the run in its row
static inline size_t ccl_sweep_max_nodes(int H, int W) { return 1 + (size_t)H * (size_t)((W + 1) / 2 + 1); }

}  // namespace dvc
