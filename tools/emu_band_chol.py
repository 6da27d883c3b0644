"""Thread-level emulation (numpy) of k_band_chol's index logic: sliding-window banded(+arrow)
Cholesky with 4x4 register blocks per thread and the single-warp axpy back substitution.
Development aid for csrc/ba_band_chol.cuh (not part of the product, not a test oracle)."""
import numpy as np


def build(nb, bw, W, rng):
    n = nb + 3
    B = rng.standard_normal((n, n))
    A = B @ B.T + n * np.eye(n)
    for i in range(nb):
        for j in range(nb):
            if abs(i - j) > bw:
                A[i, j] = 0.0
    A += np.eye(n) * 3 * n          # keep SPD after band truncation
    rhs = rng.standard_normal(n)
    RS = W + 4
    Ab = np.zeros((nb + W + 8, RS))
    for r in range(nb + W + 8):
        if r < nb:
            for k in range(0, min(bw, r) + 1):
                Ab[r, k] = A[r, r - k]
            for a in range(3):
                Ab[r, W + a] = A[nb + a, r]
            Ab[r, W + 3] = rhs[r]
        else:
            Ab[r, 0] = 1.0
    C4 = np.zeros((4, 4))
    C4[:3, :3] = A[nb:, nb:]
    C4[3, :3] = rhs[nb:]
    C4[:3, 3] = rhs[nb:]
    return A, rhs, Ab, C4


def emulate(nb, bw, W, Ab, C4):
    Wb = W // 4
    RS = W + 4
    LS = bw + 1
    threads = [(P, Q) for P in range(Wb + 1) for Q in range(P + 1)]
    v = {}
    for (P, Q) in threads:
        blk = np.zeros((4, 4))
        for a in range(4):
            for b in range(4):
                if P < Wb:
                    rp, rq = 4 * P + a, 4 * Q + b
                    hi, lo = max(rp, rq), min(rp, rq)
                    blk[a, b] = Ab[hi, hi - lo]
                elif Q < Wb:
                    blk[a, b] = Ab[4 * Q + b, W + a]
                else:
                    blk[a, b] = C4[a, b]
        v[(P, Q)] = blk
    Lr = np.zeros((nb, LS))
    La = np.zeros((4, nb))
    dinv = np.zeros(nb)
    nsteps = (nb + 3) // 4 * 4
    for j in range(nsteps):
        pj = j % W
        Pj, ij = pj // 4, pj % 4
        col = np.full(W + 4, np.nan)
        for (P, Q) in threads:                      # phase A
            blk = v[(P, Q)]
            if Q == Pj:
                for a in range(4):
                    col[4 * P + a] = blk[a, ij]
            elif P == Pj:
                for b in range(4):
                    col[4 * Q + b] = blk[ij, b]
        assert not np.isnan(col).any()
        d = col[pj]
        assert d > 0
        invd, rs = 1.0 / d, 1.0 / np.sqrt(d)
        for (P, Q) in threads:                      # phase B
            cp = col[4 * P:4 * P + 4]
            tq = col[4 * Q:4 * Q + 4] * invd
            v[(P, Q)] -= np.outer(cp, tq)
        for p in range(W):
            rp = j + ((p - pj) % W)
            k = rp - j
            if k <= bw and rp < nb and j < nb:
                Lr[rp, k] = col[p] * rs
        if j < nb:
            for a in range(4):
                La[a, j] = col[W + a] * rs
            dinv[j] = rs
        rn = j + W                                  # phase C
        row = Ab[rn]
        for (P, Q) in threads:
            blk = v[(P, Q)]
            if Q == Pj:
                for a in range(4):
                    if P == Wb:
                        val = row[W + a]
                    else:
                        q = 4 * P + a
                        val = row[0] if q == pj else row[W - ((q - pj) % W)]
                    blk[a, ij] = val
                    if P == Pj:
                        blk[ij, a] = val
            elif P == Pj:
                for b in range(4):
                    q = 4 * Q + b
                    blk[ij, b] = row[W - ((q - pj) % W)]
    C = v[(Wb, Wb)]
    xI = np.linalg.solve(C[:3, :3], C[3, :3])
    # back substitution, chunked axpy form with MS slots
    MS = 1 + (bw + 31) // 32
    x = np.zeros(nb + 3)
    x[nb:] = xI
    yy0 = La[3] - xI @ La[:3]
    ctop = (nb + 31) // 32 - 1
    yy = np.zeros((MS, 32))
    def fresh(c, m):
        out = np.zeros(32)
        for l in range(32):
            i = 32 * (c - m) + l
            if 0 <= i < nb:
                out[l] = yy0[i]
        return out
    for m in range(MS):
        yy[m] = fresh(ctop, m)
    for c in range(ctop, -1, -1):
        dl = np.array([dinv[32 * c + l] if 32 * c + l < nb else 0.0 for l in range(32)])
        for jj in range(31, -1, -1):
            j = 32 * c + jj
            xj = yy[0, jj] * dl[jj]
            if j < nb:
                x[j] = xj
            for m in range(MS):
                for l in range(32):
                    i = 32 * (c - m) + l
                    k = jj - l + 32 * m
                    if 1 <= k <= bw and i >= 0 and j < nb:
                        yy[m, l] -= Lr[j, k] * xj
        for m in range(MS - 1):
            yy[m] = yy[m + 1]
        yy[MS - 1] = fresh(c - 1, MS - 1)
    return x


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for nb, bw, W in [(30, 9, 12), (42, 11, 12), (66, 29, 32), (18, 17, 20), (100, 40, 44)]:
        A, rhs, Ab, C4 = build(nb, bw, W, rng)
        x = emulate(nb, bw, W, Ab, C4)
        ref = np.linalg.solve(A, rhs)
        print(nb, bw, W, np.abs(x - ref).max() / np.abs(ref).max())
