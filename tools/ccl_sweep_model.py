"""Lane-level model of k_ccl_sweep (csrc/k_ccl_sweep.cuh): the bit tricks and enumeration rules of the CUDA kernel
restated with Python integers so that they can be checked on the CPU against oracle/stage_ops.contour_filter.

The word width WB and the strip height are parameters: small words (4 or 8 bits) put run boundaries, carries and
neighbour bits on lane boundaries all the time, which 64-bit words on random images almost never do.

    python tools/ccl_sweep_model.py            # randomised check, a few thousand small masks
"""
from __future__ import annotations

import sys
import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])


class Model:
    def __init__(self, H, W, WB=8, strips=4):
        self.H, self.W, self.WB = H, W, WB
        self.NL = (W + WB - 1) // WB                      # lanes (one word per lane)
        self.FULL = (1 << WB) - 1
        self.vm = [self._valid(l) for l in range(self.NL)]
        self.strips = strips

    def _valid(self, l):
        rem = self.W - l * self.WB
        return self.FULL if rem >= self.WB else ((1 << rem) - 1 if rem > 0 else 0)

    def pack(self, img):
        rows = []
        for y in range(self.H):
            r = []
            for l in range(self.NL):
                v = 0
                for i in range(self.WB):
                    x = l * self.WB + i
                    if x < self.W and img[y, x]:
                        v |= 1 << i
                r.append(v)
            rows.append(r)
        return rows

    def unpack(self, rows):
        out = np.zeros((self.H, self.W), np.uint8)
        for y in range(self.H):
            for l in range(self.NL):
                for i in range(self.WB):
                    x = l * self.WB + i
                    if x < self.W and (rows[y][l] >> i) & 1:
                        out[y, x] = 255
        return out

    # ---- carry resolution over lanes: c[j+1] = g[j] | (p[j] & c[j]) through one addition ----
    @staticmethod
    def carries(Gm, Pm, nl):
        a = Gm | Pm
        return ((a + Gm) ^ a ^ Gm) & ((1 << nl) - 1)

    def brev(self, v):
        r = 0
        for i in range(self.WB):
            if (v >> i) & 1:
                r |= 1 << (self.WB - 1 - i)
        return r

    def hfill_dir(self, b, s):
        """fill towards higher bits / lanes: every bit of a b-run at or above a seed"""
        nl, FULL = self.NL, self.FULL
        t = [(b[l] + s[l]) for l in range(nl)]
        Gm = sum(1 << l for l in range(nl) if t[l] > FULL)
        t = [v & FULL for v in t]
        Pm = sum(1 << l for l in range(nl) if t[l] == FULL)
        cm = self.carries(Gm, Pm, nl)
        out = []
        for l in range(nl):
            tt = (t[l] + ((cm >> l) & 1)) & FULL
            out.append((b[l] & ~tt & FULL) | s[l])
        return out

    def hfill(self, b, s):
        r = self.hfill_dir(b, s)
        br = [self.brev(v) for v in reversed(b)]
        sr = [self.brev(v) for v in reversed(s)]
        l = self.hfill_dir(br, sr)
        l = [self.brev(v) for v in reversed(l)]
        return [r[i] | l[i] for i in range(self.NL)]

    # ---- phase A ----
    def flood(self, M):
        H, NL, vm = self.H, self.NL, self.vm
        R = (H + self.strips - 1) // self.strips
        O = [[0] * NL for _ in range(H)]
        fg = [any(M[y]) for y in range(H)]
        border = []
        lastl, lastb = (self.W - 1) // self.WB, (self.W - 1) % self.WB
        for l in range(NL):
            v = 0
            if l == 0:
                v |= 1
            if l == lastl:
                v |= 1 << lastb
            border.append(v)

        def row_o(y):
            return O[y] if fg[y] else list(vm)

        def sweep(y0, y1, direction, prev, first):
            changed = False
            ys = range(y0, y1) if direction > 0 else range(y1 - 1, y0 - 1, -1)
            for y in ys:
                if not fg[y]:
                    prev = list(vm)
                    continue
                b = [~M[y][l] & vm[l] for l in range(NL)]
                o = [0] * NL if first else O[y]
                bm = list(vm) if (y == 0 or y == H - 1) else border
                s = [o[l] | (prev[l] & b[l]) | (bm[l] & b[l]) for l in range(NL)]
                if first or any(s[l] != o[l] for l in range(NL)):
                    r = self.hfill(b, s)
                    if first or any(r[l] != o[l] for l in range(NL)):
                        changed = True
                    O[y] = r
                    prev = r
                else:
                    prev = o
            return changed

        dirty = [True] * self.strips
        it = 0
        while True:
            nxt = [False] * self.strips
            anyc = False
            for s in range(self.strips):          # (on the GPU: concurrently; any order converges)
                y0, y1 = s * R, min(H, (s + 1) * R)
                if y0 >= y1 or not dirty[s]:
                    continue
                if it == 0:
                    c = sweep(y0, y1, +1, [0] * NL, True)
                    c |= sweep(y0, y1, -1, [0] * NL, False)
                    c = True
                else:
                    c = sweep(y0, y1, +1, row_o(y0 - 1) if y0 > 0 else [0] * NL, False)
                    c |= sweep(y0, y1, -1, row_o(y1) if y1 < H else [0] * NL, False)
                if c:
                    anyc = True
                    for k in (s - 1, s, s + 1):
                        if 0 <= k < self.strips:
                            nxt[k] = True
            if not anyc:
                break
            dirty = nxt
            it += 1
        F = [[(vm[l] & ~O[y][l]) if fg[y] else 0 for l in range(NL)] for y in range(H)]
        return F, fg

    # ---- phase A as a union-find over background row runs (second cut of the kernel) ----
    def fill_holes_uf(self, M, SR):
        """F = foreground + background runs that do not reach the outside.  Nodes: background runs of rows with foreground
        (node 0 = outside); rows without foreground are outside as a whole.  Strips of SR rows are swept top-down, the links
        across strip boundaries are made afterwards (the GPU does them after a barrier)."""
        H, NL, vm, WB = self.H, self.NL, self.vm, self.WB
        fg = [any(M[y]) for y in range(H)]
        B = [[~M[y][l] & vm[l] for l in range(NL)] for y in range(H)]
        metas = [None] * H
        base = [0] * H
        P = [0]
        lastl, lastb = (self.W - 1) // WB, (self.W - 1) % WB

        def all_outside(y):
            st, pre, n = metas[y]
            for i in range(n):
                self.union(P, base[y] + i, 0)

        def union4(y):
            c, u = B[y], B[y - 1]
            cst, cpre, _ = metas[y]
            ust, upre, _ = metas[y - 1]
            for l in range(NL):
                cl = (c[l - 1] >> (WB - 1)) & 1 if l > 0 else 0
                ul = (u[l - 1] >> (WB - 1)) & 1 if l > 0 else 0
                I = c[l] & u[l]
                starts = I & ~((I << 1) | (cl & ul)) & self.FULL
                while starts:
                    p = (starts & -starts).bit_length() - 1
                    starts &= starts - 1
                    cn = base[y] + cpre[l] + bin(cst[l] & self.lowmask(p)).count("1") - 1
                    un = base[y - 1] + upre[l] + bin(ust[l] & self.lowmask(p)).count("1") - 1
                    self.union(P, cn, un)

        nstrips = (H + SR - 1) // SR
        order = list(range(nstrips))
        np.random.default_rng(H * 131 + self.W).shuffle(order)         # strips run in any order
        for s in order:
            y0, y1 = s * SR, min(H, (s + 1) * SR)
            have_u = False
            for y in range(y0, y1):
                if not fg[y]:
                    if have_u:
                        all_outside(y - 1)
                    have_u = False
                    continue
                metas[y] = self.meta(B[y])
                n = metas[y][2]
                base[y] = len(P)
                all_out = y == 0 or y == H - 1 or (y > y0 and not have_u)
                for i in range(n):
                    P.append(0 if all_out else base[y] + i)
                if not all_out:
                    if B[y][0] & 1:
                        P[base[y]] = 0                                  # the run that starts at column 0
                    if (B[y][lastl] >> lastb) & 1:
                        st, pre, _ = metas[y]
                        self.union(P, base[y] + pre[lastl] + bin(st[lastl] & self.lowmask(lastb)).count("1") - 1, 0)
                if have_u:
                    union4(y)                                           # also for an all-outside row: its neighbours above reach the outside through it
                have_u = True
        for s in range(1, nstrips):
            yb = s * SR
            up, cu = fg[yb - 1], fg[yb]
            if up and not cu:
                all_outside(yb - 1)
            elif cu and not up:
                all_outside(yb)
            elif up and cu:
                union4(yb)
        F = []
        for y in range(H):
            if not fg[y]:
                F.append([0] * NL)
                continue
            st, pre, _ = metas[y]
            row = []
            for l in range(NL):
                m = B[y][l]
                holes = 0
                while m:
                    run, lo = self.lowest_run(m)
                    m &= ~run
                    node = base[y] + pre[l] + bin(st[l] & self.lowmask(lo)).count("1") - 1
                    if self.find(P, node) != 0:
                        holes |= run
                row.append(M[y][l] | holes)
            F.append(row)
        return F, fg

    # ---- phase B helpers ----
    def meta(self, c):
        """run starts per lane and exclusive prefix of their counts"""
        WB, FULL = self.WB, self.FULL
        starts, pre = [], []
        n = 0
        for l in range(self.NL):
            cl = (c[l - 1] >> (WB - 1)) & 1 if l > 0 else 0
            st = c[l] & ~(((c[l] << 1) | cl)) & FULL
            starts.append(st)
            pre.append(n)
            n += bin(st).count("1")
        return starts, pre, n

    def lowmask(self, p):
        return ((2 << p) - 1) & self.FULL

    @staticmethod
    def lowest_run(m):
        lo = (m & -m).bit_length() - 1
        t = m + (1 << lo)
        return m & ~t, lo

    def find(self, P, x):
        while True:
            p = P[x]
            if p == x:
                return x
            gp = P[p]
            if gp == p:
                return p
            P[x] = gp
            x = gp

    def union(self, P, a, b):
        while True:
            a, b = self.find(P, a), self.find(P, b)
            if a == b:
                return
            if a < b:
                a, b = b, a
            old = P[a]
            P[a] = min(P[a], b)
            if old == a:
                return
            a = old

    def union_row(self, P, c, cst, cpre, cbase, u, ust, upre, ubase):
        WB, FULL, NL = self.WB, self.FULL, self.NL
        for l in range(NL):
            cw, uw = c[l], u[l]
            if cw == 0:
                continue
            cl = (c[l - 1] >> (WB - 1)) & 1 if l > 0 else 0
            ul = (u[l - 1] >> (WB - 1)) & 1 if l > 0 else 0
            ur = u[l + 1] & 1 if l + 1 < NL else 0
            ud = (uw | (uw << 1) | (uw >> 1) | ul | (ur << (WB - 1))) & FULL
            I = cw & ud
            while I:
                run, a = self.lowest_run(I)
                I &= ~run
                b = a + bin(run).count("1") - 1
                cnode = cbase + cpre[l] + bin(cst[l] & self.lowmask(a)).count("1") - 1
                wmask = (run | (run << 1) | (run >> 1)) & FULL
                E = uw & wmask
                skip = a == 0 and cl and (uw & 1) and ul
                if skip:
                    E &= ~(uw & ~(uw + 1))          # the u-run through bit 0: linked by the lane to the left
                if a == 0 and ul and not (uw & 1):
                    self.union(P, cnode, ubase + upre[l] - 1)
                if b == WB - 1 and ur and not (uw >> (WB - 1)) & 1:
                    self.union(P, cnode, ubase + upre[l] + bin(ust[l]).count("1"))
                while E:
                    er, p = self.lowest_run(E)
                    E &= ~er
                    unode = ubase + upre[l] + bin(ust[l] & self.lowmask(p)).count("1") - 1
                    self.union(P, cnode, unode)

    def area_row(self, A, a, ast, apre, abase, b):
        WB, FULL, NL = self.WB, self.FULL, self.NL
        for l in range(NL):
            aw, bw = a[l], b[l]
            if aw == 0:
                continue
            an = a[l + 1] & 1 if l + 1 < NL else 0
            bn = b[l + 1] & 1 if l + 1 < NL else 0
            al = (a[l - 1] >> (WB - 1)) & 1 if l > 0 else 0
            bl = (b[l - 1] >> (WB - 1)) & 1 if l > 0 else 0
            a1 = (aw >> 1) | (an << (WB - 1))
            b1 = (bw >> 1) | (bn << (WB - 1))
            q4 = aw & a1 & bw & b1
            q3a = aw & ((a1 & bw & ~b1) | (a1 & ~bw & b1) | (~a1 & bw & b1)) & FULL
            q3b = ~aw & a1 & bw & b1 & FULL
            carry = 1 if (not al and (aw & 1) and bl and (bw & 1)) else 0
            q3s = ((q3b << 1) | carry) & FULL
            m = aw
            while m:
                run, lo = self.lowest_run(m)
                m &= ~run
                w = 2 * bin(q4 & run).count("1") + bin(q3a & run).count("1") + bin(q3s & run).count("1")
                if w:
                    A[abase + apre[l] + bin(ast[l] & self.lowmask(lo)).count("1") - 1] += w
        return

    def contour_filter(self, img, min_area):
        H, NL = self.H, self.NL
        M = self.pack(img != 0)
        F, fg = self.flood(M)
        F2, fg2 = self.fill_holes_uf(M, max(1, self.H // (2 * self.strips)))
        assert F2 == F and fg2 == fg, "union-find fill differs from the flood"
        metas = [self.meta(F[y]) if fg[y] else ([0] * NL, [0] * NL, 0) for y in range(H)]
        base = [0] * (H + 1)
        for y in range(H):
            base[y + 1] = base[y] + metas[y][2]
        n = base[H]
        P = list(range(n))
        A = [0] * n
        R = (H + self.strips - 1) // self.strips
        # in-strip unions, then seams (order is irrelevant to the result)
        for s in range(self.strips):
            for y in range(s * R + 1, min(H, (s + 1) * R)):
                if fg[y] and fg[y - 1]:
                    self.union_row(P, F[y], metas[y][0], metas[y][1], base[y], F[y - 1], metas[y - 1][0], metas[y - 1][1], base[y - 1])
        for s in range(1, self.strips):
            y = s * R
            if y < H and fg[y] and fg[y - 1]:
                self.union_row(P, F[y], metas[y][0], metas[y][1], base[y], F[y - 1], metas[y - 1][0], metas[y - 1][1], base[y - 1])
        zero = [0] * NL
        for y in range(H):
            if fg[y]:
                self.area_row(A, F[y], metas[y][0], metas[y][1], base[y], F[y + 1] if y + 1 < H else zero)
        for i in range(n):
            r = self.find(P, i)
            if r != i:
                A[r] += A[i]
        thr = int(np.floor(2.0 * min_area))
        out = [[0] * NL for _ in range(H)]
        for y in range(H):
            if not fg[y]:
                continue
            st, pre, _ = metas[y]
            for l in range(NL):
                m = F[y][l]
                keep = 0
                while m:
                    run, lo = self.lowest_run(m)
                    m &= ~run
                    node = base[y] + pre[l] + bin(st[l] & self.lowmask(lo)).count("1") - 1
                    if A[self.find(P, node)] > thr:
                        keep |= run
                out[y][l] = keep
        return self.unpack(out), self.unpack(F)


def main(n_cases=3000, seed=0):
    from oracle import stage_ops as so
    rng = np.random.default_rng(seed)
    for case in range(n_cases):
        H, W = int(rng.integers(1, 24)), int(rng.integers(1, 40))
        WB = int(rng.choice([2, 3, 4, 8]))
        strips = int(rng.choice([1, 2, 3, 5]))
        kind = case % 4
        if kind == 0:
            img = (rng.random((H, W)) < rng.random()).astype(np.uint8) * 255
        elif kind == 1:                                  # rings and blobs
            img = np.zeros((H, W), np.uint8)
            for _ in range(int(rng.integers(1, 5))):
                y0, x0 = int(rng.integers(0, H)), int(rng.integers(0, W))
                y1, x1 = int(rng.integers(y0, H)) + 1, int(rng.integers(x0, W)) + 1
                img[y0:y1, x0:x1] = 255
                if rng.random() < 0.7 and y1 - y0 > 2 and x1 - x0 > 2:
                    img[y0 + 1:y1 - 1, x0 + 1:x1 - 1] = 0
        elif kind == 2:                                  # serpentine walls: long flood paths
            img = np.zeros((H, W), np.uint8)
            for y in range(1, H, 2):
                img[y, :] = 255
                img[y, (W - 1) if (y // 2) % 2 else 0] = 0
            img ^= (rng.random((H, W)) < 0.03).astype(np.uint8) * 255
        else:
            img = (rng.random((H, W)) < 0.6).astype(np.uint8) * 255
        min_area = float(rng.choice([0, 0.5, 1, 2.5, 4, 10]))
        want = so.contour_filter(img, min_area)
        got, _ = Model(H, W, WB, strips).contour_filter(img, min_area)
        if not np.array_equal(want, got):
            print("MISMATCH case", case, "H W WB strips", H, W, WB, strips, "min_area", min_area)
            print(img // 255)
            print(want // 255)
            print(got // 255)
            return 1
    print("ok:", n_cases, "cases")
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 3000))
