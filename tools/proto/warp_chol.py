"""Lane-level emulation (numpy, one array element per lane) of the warp-per-geometry packed-triangular
algorithms of csrc/ic_factor.cu: potrf (right-looking, panels of 8, bordered row = forward solve), back
substitution, row-oriented trtri and lauum, all in place on row-major packed lower storage. Development aid:
validates the index / mask logic against numpy.linalg before the CUDA version is trusted."""
import numpy as np

NB = 8
def tri(i): return i * (i + 1) // 2

def potrf(A, n, border):
    """A: packed, rows 0..n-1 (+ row n = b when border). Returns ok, rdiag[n]"""
    last = n if border else n - 1          # last row index
    rdiag = np.zeros(n)
    ok = True
    lane = np.arange(32)
    for j0 in range(0, n, NB):
        bs = min(NB, n - j0)
        nrows = last - j0 + 1
        nslots = (nrows + 31) // 32
        # ---- panel: registers a[m][c]
        a = np.zeros((nslots, NB, 32))
        for m in range(nslots):
            i = j0 + lane + 32 * m
            for c in range(bs):
                valid = (i <= last) & ((j0 + c <= i) | (i == n))      # stored entries of row i (border row holds all columns < n)
                idx = np.where(valid, tri(np.minimum(i, last)) + j0 + c, 0)
                a[m, c] = np.where(valid, A[idx], 0.0)
        for c in range(bs):
            d = a[0, c, c]                       # shfl from lane c, slot 0
            if not d > 0: ok = False
            rs = 1.0 / np.sqrt(d)
            rdiag[j0 + c] = rs
            for m in range(nslots):
                a[m, c] = a[m, c] * rs           # owner row: d * rs = sqrt(d)
            for c2 in range(c + 1, bs):
                t = a[0, c, c2]                  # L[j0+c2][j0+c]: lane c2, slot 0
                for m in range(nslots):
                    a[m, c2] = a[m, c2] - a[m, c] * t
        for m in range(nslots):
            i = j0 + lane + 32 * m
            for c in range(bs):
                valid = (i <= last) & ((j0 + c <= i) | (i == n))
                idx = tri(np.minimum(i, last)) + j0 + c
                A[idx[valid]] = a[m, c][valid]
        # ---- trailing update: rows i > j0+bs-1, columns j in [j0+bs, min(i, n-1)]
        t0 = j0 + bs
        if t0 >= n: continue
        ncols = n - t0
        cslots = (ncols + 31) // 32
        Lj = np.zeros((cslots, NB, 32))
        for m in range(cslots):
            j = t0 + lane + 32 * m
            for c in range(bs):
                valid = j < n
                Lj[m, c] = np.where(valid, A[np.where(valid, tri(np.minimum(j, n - 1)) + j0 + c, 0)], 0.0)
        for i in range(t0, last + 1):
            Li = [A[tri(i) + j0 + c] for c in range(bs)]           # broadcast loads
            jmax = min(i, n - 1)
            for m in range((jmax - t0) // 32 + 1):
                j = t0 + lane + 32 * m
                valid = j <= jmax
                idx = np.where(valid, tri(i) + j, 0)
                x = A[idx]
                for c in range(bs):
                    x = x - Li[c] * Lj[m, c]
                A[idx[valid]] = x[valid]
    return ok, rdiag

def backsolve(A, n, rdiag):
    """L^T x = y with y = border row n; returns x. Right-looking: after x_k, x_i -= L[k][i] x_k (row k of L)."""
    lane = np.arange(32)
    nslots = (n + 31) // 32
    x = np.zeros((nslots, 32))
    for m in range(nslots):
        i = lane + 32 * m
        x[m] = np.where(i < n, A[np.where(i < n, tri(n) + i, 0)], 0.0)
    for k in range(n - 1, -1, -1):
        xk = x[k // 32, k % 32] * rdiag[k]                 # shfl + mul
        x[k // 32, k % 32] = xk
        for m in range(k // 32 + 1):
            i = lane + 32 * m
            valid = i < k
            Lk = np.where(valid, A[np.where(valid, tri(k) + i, 0)], 0.0)
            x[m] = x[m] - Lk * xk
    out = np.zeros(n)
    for m in range(nslots):
        i = lane + 32 * m
        out[i[i < n]] = x[m][i < n]
    return out

def trtri(A, n, rdiag):
    """in place: L -> W = L^-1 (lower), row panels of 8, top to bottom"""
    lane = np.arange(32)
    for i0 in range(0, n, NB):
        bs = min(NB, n - i0)
        # diagonal block inverse D (bs x bs lower), lanes 0..7 = columns
        Ld = np.zeros((NB, NB))
        for r in range(bs):
            for c in range(r + 1):
                Ld[r, c] = A[tri(i0 + r) + i0 + c]
        D = np.zeros((NB, NB))
        for c in range(bs):                 # lane c
            D[c, c] = rdiag[i0 + c]
            for r in range(c + 1, bs):
                s = 0.0
                for k in range(c, r):
                    s += Ld[r, k] * D[k, c]
                D[r, c] = -s * rdiag[i0 + r]
        # off-diagonal: acc[r][j] = sum_k L[i0+r][k] W[k][j], k < i0, j <= k
        nslots = (i0 + 31) // 32
        acc = np.zeros((nslots, NB, 32))
        for k in range(i0):
            Lr = [A[tri(i0 + r) + k] if r < bs else 0.0 for r in range(NB)]
            for m in range(k // 32 + 1):
                j = lane + 32 * m
                valid = j <= k
                Wk = np.where(valid, A[np.where(valid, tri(k) + j, 0)], 0.0)
                for r in range(bs):
                    acc[m, r] = acc[m, r] + Lr[r] * Wk
        # W[I, 0:i0] = -D acc
        for m in range(nslots):
            j = lane + 32 * m
            valid = j < i0
            for r in range(bs):
                s = np.zeros(32)
                for r2 in range(r + 1):
                    s = s + D[r, r2] * acc[m, r2]
                A[(tri(i0 + r) + j)[valid]] = -s[valid]
        for r in range(bs):
            for c in range(r + 1):
                A[tri(i0 + r) + i0 + c] = D[r, c]

def lauum(A, n):
    """in place: W (lower) -> W^T W (lower triangle), row panels of 8 ascending"""
    lane = np.arange(32)
    for i0 in range(0, n, NB):
        bs = min(NB, n - i0)
        ncols = i0 + bs
        nslots = (ncols + 31) // 32
        acc = np.zeros((nslots, NB, 32))
        for k in range(i0, n):
            Wr = [A[tri(k) + i0 + r] if (r < bs and i0 + r <= k) else 0.0 for r in range(NB)]
            for m in range(nslots):
                j = lane + 32 * m
                valid = (j <= k) & (j < ncols)
                Wk = np.where(valid, A[np.where(valid, tri(k) + j, 0)], 0.0)
                for r in range(bs):
                    acc[m, r] = acc[m, r] + Wr[r] * Wk
        for m in range(nslots):
            j = lane + 32 * m
            for r in range(bs):
                valid = j <= i0 + r
                A[(tri(i0 + r) + j)[valid]] = acc[m, r][valid]

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (3, 8, 12, 33, 84, 100):
        Jm = rng.standard_normal((n, n + 6))
        S = Jm @ Jm.T
        b = rng.standard_normal(n)
        A = np.zeros(tri(n + 1) + n)
        for i in range(n):
            A[tri(i):tri(i) + i + 1] = S[i, :i + 1]
        A[tri(n):tri(n) + n] = b
        ok, rd = potrf(A, n, True)
        L = np.zeros((n, n))
        for i in range(n): L[i, :i + 1] = A[tri(i):tri(i) + i + 1]
        Lref = np.linalg.cholesky(S)
        x = backsolve(A, n, rd)
        e1 = np.abs(L - Lref).max(); e2 = np.abs(x - np.linalg.solve(S, b)).max() / np.abs(x).max()
        trtri(A, n, rd)
        W = np.zeros((n, n))
        for i in range(n): W[i, :i + 1] = A[tri(i):tri(i) + i + 1]
        e3 = np.abs(W - np.linalg.inv(Lref)).max() / np.abs(W).max()
        lauum(A, n)
        Ai = np.zeros((n, n))
        for i in range(n): Ai[i, :i + 1] = A[tri(i):tri(i) + i + 1]
        Ai = np.tril(Ai) + np.tril(Ai, -1).T
        e4 = np.abs(Ai - np.linalg.inv(S)).max() / np.abs(Ai).max()
        print(n, ok, "L %.1e solve %.1e trtri %.1e inv %.1e" % (e1, e2, e3, e4))


# ------------------------------------------------------------------------------------------------------
# DMMA (mma.m8n8k4.f64) emulation: per-lane fragments
#   a : A[row = lane >> 2][col = lane & 3]     b : B[row = lane & 3][col = lane >> 2]
#   c0, c1 : C[row = lane >> 2][col = 2 (lane & 3) + {0, 1}]
# ------------------------------------------------------------------------------------------------------
LANE = np.arange(32)
GID, TIG = LANE >> 2, LANE & 3

def dmma(c0, c1, a, b):
    Am = np.zeros((8, 4)); Bm = np.zeros((4, 8))
    Am[GID, TIG] = a
    Bm[TIG, GID] = b
    D = Am @ Bm
    return c0 + D[GID, 2 * TIG], c1 + D[GID, 2 * TIG + 1]

def ld(A, idx, valid):
    return np.where(valid, A[np.where(valid, idx, 0)], 0.0)

def trailing_dmma(A, n, last, j0, bs):
    """A[i][j] -= sum_c L[i][j0+c] L[j][j0+c] for rows i in [t0, last], columns j in [t0, min(i, n-1)], 8 x 8 tiles"""
    t0 = j0 + bs
    for I in range(t0, last + 1, 8):                       # row tile
        for J in range(t0, min(I + 8, n), 8):              # column tile (J <= I)
            i = I + GID
            jc = J + 2 * TIG
            c0 = np.zeros(32); c1 = np.zeros(32)
            for k0 in range(0, bs, 4):
                va = (i <= last) & (k0 + TIG < bs)
                a = ld(A, tri(np.minimum(i, last)) + j0 + k0 + TIG, va)
                jb = J + GID
                vb = (jb < n) & (k0 + TIG < bs)
                b = ld(A, tri(np.minimum(jb, n - 1)) + j0 + k0 + TIG, vb)
                c0, c1 = dmma(c0, c1, a, b)
            for e, c in ((0, c0), (1, c1)):
                j = jc + e
                valid = (i <= last) & (j <= np.minimum(i, n - 1))
                idx = tri(np.minimum(i, last)) + j
                A[idx[valid]] -= c[valid]

def potrf_dmma(A, n, border):
    last = n if border else n - 1
    rdiag = np.zeros(n)
    ok = True
    lane = np.arange(32)
    for j0 in range(0, n, NB):
        bs = min(NB, n - j0)
        nrows = last - j0 + 1
        nslots = (nrows + 31) // 32
        a = np.zeros((nslots, NB, 32))
        for m in range(nslots):
            i = j0 + lane + 32 * m
            for c in range(bs):
                valid = (i <= last) & ((j0 + c <= i) | (i == n))
                a[m, c] = ld(A, tri(np.minimum(i, last)) + j0 + c, valid)
        for c in range(bs):
            d = a[0, c, c]
            if not d > 0: ok = False
            rs = 1.0 / np.sqrt(d)
            rdiag[j0 + c] = rs
            for m in range(nslots): a[m, c] = a[m, c] * rs
            for c2 in range(c + 1, bs):
                t = a[0, c, c2]
                for m in range(nslots): a[m, c2] = a[m, c2] - a[m, c] * t
        for m in range(nslots):
            i = j0 + lane + 32 * m
            for c in range(bs):
                valid = (i <= last) & ((j0 + c <= i) | (i == n))
                idx = tri(np.minimum(i, last)) + j0 + c
                A[idx[valid]] = a[m, c][valid]
        if j0 + bs < n:
            trailing_dmma(A, n, last, j0, bs)
    return ok, rdiag

def backsolve_blocked(A, n, rdiag):
    """L^T x = y (bordered row), blocks of 8 from the bottom: 8 x 8 triangular solve by shuffles, then a rank-8 update"""
    lane = np.arange(32)
    nslots = (n + 31) // 32
    x = np.zeros((nslots, 32))
    for m in range(nslots):
        i = lane + 32 * m
        x[m] = ld(A, tri(n) + i, i < n)
    nblk = (n + 7) // 8
    for kb in range(nblk - 1, -1, -1):
        k0 = 8 * kb
        bs = min(8, n - k0)
        km, kl0 = k0 // 32, k0 % 32
        xk = np.zeros(8)
        for c in range(bs - 1, -1, -1):
            k = k0 + c
            xk[c] = x[km, kl0 + c] * rdiag[k]          # shfl + mul
            x[km, kl0 + c] = xk[c]
            i = lane + 32 * km                         # only the block's own lanes (i in [k0, k)) update here
            valid = (i >= k0) & (i < k)
            x[km] = x[km] - ld(A, tri(k) + i, valid) * xk[c]
        for m in range(k0 // 32 + 1 if k0 > 0 else 0):
            i = lane + 32 * m
            valid = i < k0
            s = np.zeros(32)
            for c in range(bs):
                s = s + ld(A, tri(k0 + c) + i, valid) * xk[c]
            x[m] = x[m] - s
    out = np.zeros(n)
    for m in range(nslots):
        i = lane + 32 * m
        out[i[i < n]] = x[m][i < n]
    return out

def trtri_dmma(A, n, rdiag, Ds):
    """in place L -> W = L^-1: row panels I of 8 from the top. T = L[I][0:i0] W[0:i0][0:i0] by 8 x 8 tiles (DMMA),
    then W[I][0:i0] = -D T with D = L[I][I]^-1 (DMMA, D from scratch Ds[8][8])."""
    for i0 in range(0, n, NB):
        bs = min(NB, n - i0)
        Ld = np.zeros((NB, NB))
        for r in range(bs):
            for c in range(r + 1): Ld[r, c] = A[tri(i0 + r) + i0 + c]
        D = np.zeros((NB, NB))
        for c in range(bs):
            D[c, c] = rdiag[i0 + c]
            for r in range(c + 1, bs):
                s = 0.0
                for k in range(c, r): s += Ld[r, k] * D[k, c]
                D[r, c] = -s * rdiag[i0 + r]
        Ds[:] = D.reshape(-1)
        # all tiles' T first (reads rows I), then the writes
        tiles = []
        for J in range(0, i0, 8):
            c0 = np.zeros(32); c1 = np.zeros(32)
            for k0 in range(J, i0, 4):                   # W[k][j] = 0 for k < j: start at the tile's first column
                r = i0 + GID
                a = ld(A, tri(np.minimum(r, n - 1)) + k0 + TIG, (r < n))          # L[i0 + gid][k0 + tig]   (k0 + tig < i0 always: i0, J multiples of 8)
                kk = k0 + TIG; jj = J + GID
                b = ld(A, tri(kk) + jj, jj <= kk)                                   # W[k0 + tig][J + gid], zero above the diagonal
                c0, c1 = dmma(c0, c1, a, b)
            tiles.append((J, c0, c1))
        out = []
        for (J, t0_, t1_) in tiles:
            # -D T: A-fragment = D[gid][k], B-fragment = T[k][gid]: T lives in C layout (row gid, cols 2 tig + e) -> needs a
            # transpose through scratch: emulate with a dense 8 x 8
            T = np.zeros((8, 8)); T[GID, 2 * TIG] = t0_; T[GID, 2 * TIG + 1] = t1_
            R = -(D @ T)
            out.append((J, R))
        for (J, R) in out:
            for r in range(bs):
                for c in range(8):
                    A[tri(i0 + r) + J + c] = R[r, c]
        for r in range(bs):
            for c in range(r + 1): A[tri(i0 + r) + i0 + c] = D[r, c]

def lauum_dmma(A, n):
    """in place W -> lower(W^T W): row panels I ascending; tile (I, J), J <= I: sum_k>=i0 W[k][I + r] W[k][J + c]"""
    for i0 in range(0, n, NB):
        bs = min(NB, n - i0)
        tiles = []
        for J in range(0, i0 + 1, 8):
            c0 = np.zeros(32); c1 = np.zeros(32)
            for k0 in range(i0, n, 4):
                kk = k0 + TIG
                ra = i0 + GID                                  # A-fragment: (W^T)[I + gid][k] = W[k][i0 + gid]
                a = ld(A, tri(np.minimum(kk, n - 1)) + ra, (kk < n) & (ra <= kk) & (ra < n))
                cb = J + GID                                   # B-fragment: W[k][J + gid]
                b = ld(A, tri(np.minimum(kk, n - 1)) + cb, (kk < n) & (cb <= kk))
                c0, c1 = dmma(c0, c1, a, b)
            tiles.append((J, c0, c1))
        for (J, c0, c1) in tiles:
            i = i0 + GID
            for e, c in ((0, c0), (1, c1)):
                j = J + 2 * TIG + e
                valid = (i < n) & (j <= i)
                idx = tri(np.minimum(i, n - 1)) + j
                A[idx[valid]] = c[valid]

def test_dmma():
    rng = np.random.default_rng(1)
    for n in (3, 8, 12, 33, 84, 100):
        Jm = rng.standard_normal((n, n + 6))
        S = Jm @ Jm.T
        b = rng.standard_normal(n)
        A = np.zeros(tri(n + 1) + n)
        for i in range(n): A[tri(i):tri(i) + i + 1] = S[i, :i + 1]
        A[tri(n):tri(n) + n] = b
        ok, rd = potrf_dmma(A, n, True)
        L = np.zeros((n, n))
        for i in range(n): L[i, :i + 1] = A[tri(i):tri(i) + i + 1]
        Lref = np.linalg.cholesky(S)
        x = backsolve_blocked(A, n, rd)
        e1 = np.abs(L - Lref).max(); e2 = np.abs(x - np.linalg.solve(S, b)).max() / np.abs(x).max()
        Ds = np.zeros(64)
        trtri_dmma(A, n, rd, Ds)
        W = np.zeros((n, n))
        for i in range(n): W[i, :i + 1] = A[tri(i):tri(i) + i + 1]
        e3 = np.abs(W - np.linalg.inv(Lref)).max() / np.abs(W).max()
        lauum_dmma(A, n)
        Ai = np.zeros((n, n))
        for i in range(n): Ai[i, :i + 1] = A[tri(i):tri(i) + i + 1]
        Ai = np.tril(Ai) + np.tril(Ai, -1).T
        e4 = np.abs(Ai - np.linalg.inv(S)).max() / np.abs(Ai).max()
        print("dmma", n, ok, "L %.1e solve %.1e trtri %.1e inv %.1e" % (e1, e2, e3, e4))

if __name__ == "__main__":
    test_dmma()
