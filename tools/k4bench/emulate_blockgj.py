"""Lane-level emulation (NumPy) of the warp-resident blocked Gauss-Jordan inverse used by K4 / the H x H inverses:
verifies the index algebra of the DMMA fragment layouts before any GPU time is spent.  Test infrastructure only.

DMMA m8n8k4 (mma.sync.aligned.m8n8k4.row.col.f64): lane = 4*r + j
  A fragment: lane holds A[r][j]            (8 x 4)
  B fragment: lane holds B[j][r]            (4 x 8)
  C fragment: lane holds C[r][2j], C[r][2j+1]
"""
import numpy as np

R = np.arange(32) >> 2
J = np.arange(32) & 3


def dmma(c, a, b):
    """c: (32, 2) accumulator fragment, a: (32,) A fragment, b: (32,) B fragment -> c += A * B."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    A[R, J] = a
    B[J, R] = b
    C = A @ B
    out = c.copy()
    out[:, 0] += C[R, 2 * J]
    out[:, 1] += C[R, 2 * J + 1]
    return out


def warp_block_gj(S, NT):
    """S: (8NT, 8NT) SPD (already equilibrated).  Returns -inv(S) computed with the per-warp algorithm (NT <= 4)."""
    N = 8 * NT
    lane = np.arange(32)
    # tiles in C layout
    c = np.zeros((NT, NT, 32, 2))
    for ti in range(NT):
        for tj in range(NT):
            c[ti, tj, :, 0] = S[8 * ti + R, 8 * tj + 2 * J]
            c[ti, tj, :, 1] = S[8 * ti + R, 8 * tj + 2 * J + 1]
    Ps = np.zeros(N * 4); Ws = np.zeros(N * 4)
    for s in range(2 * NT):
        tk, half = s >> 1, s & 1
        # (1) publish the column panel from C layout
        for t in range(NT):
            for l in range(32):
                if (J[l] >> 1) == half:
                    Ps[(8 * t + R[l]) * 4 + 2 * (J[l] & 1) + 0] = c[t, tk, l, 0]
                    Ps[(8 * t + R[l]) * 4 + 2 * (J[l] & 1) + 1] = c[t, tk, l, 1]
        # (2) lane = row
        X = np.zeros((32, 4))
        for l in range(min(N, 32)):
            X[l] = Ps[l * 4:l * 4 + 4]
        # (3) four sweeps restricted to the panel
        for cc in range(4):
            kc = 4 * s + cc
            pr = X[kc].copy()              # shuffle broadcast from lane kc
            idv = 1.0 / pr[cc]
            for l in range(min(N, 32)):
                f = (1.0 - idv) if l == kc else X[l, cc] * idv
                for q in range(4):
                    if q != cc:
                        X[l, q] = X[l, q] - f * pr[q]
                X[l, cc] = -idv if l == kc else f
        # (4)/(5) publish -W' (W' = X + I on the pivot entries) and P' = P - I on the pivot entries
        for l in range(min(N, 32)):
            w = X[l].copy()
            if (l >> 2) == s:
                w[l & 3] += 1.0
                Ps[l * 4 + (l & 3)] -= 1.0
            Ws[l * 4:l * 4 + 4] = -w
        # (6) fragments
        pf = np.zeros((NT, 32)); wf = np.zeros((NT, 32))
        for t in range(NT):
            pf[t] = Ps[(8 * t + R) * 4 + J]
            wf[t] = Ws[(8 * t + R) * 4 + J]
        # (7) rank-4 update of every tile
        for ti in range(NT):
            for tj in range(NT):
                c[ti, tj] = dmma(c[ti, tj], wf[ti], pf[tj])
        # (8) -2 on the diagonal of the pivot block
        for l in range(32):
            r = R[l]
            if (r >> 2) == half and J[l] == (r >> 1) & 3 and ((r >> 1) == J[l]):
                c[tk, tk, l, r & 1] -= 2.0
    out = np.zeros((N, N))
    for ti in range(NT):
        for tj in range(NT):
            out[8 * ti + R, 8 * tj + 2 * J] = c[ti, tj, :, 0]
            out[8 * ti + R, 8 * tj + 2 * J + 1] = c[ti, tj, :, 1]
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for NT in (1, 2, 3, 4):
        N = 8 * NT
        Bm = rng.standard_normal((N + 5, N))
        S = Bm.T @ Bm + np.diag(rng.uniform(0.1, 1e4, N))
        sc = 1.0 / np.sqrt(np.diag(S))
        Se = S * sc[:, None] * sc[None, :]
        Minv = -warp_block_gj(Se, NT) * sc[:, None] * sc[None, :]
        ref = np.linalg.inv(S)
        print(NT, np.max(np.abs(Minv - ref)) / np.max(np.abs(ref)), np.linalg.cond(Se))


def warp_block_gj_sym(S, v, NT):
    """Symmetric storage (tiles ti <= tj only) + the right-hand side v riding along in lane = row layout.
    Returns (-inv(S) assembled from the upper tiles, inv(S) @ v)."""
    N = 8 * NT
    c = np.zeros((NT, NT, 32, 2))
    for ti in range(NT):
        for tj in range(ti, NT):
            c[ti, tj, :, 0] = S[8 * ti + R, 8 * tj + 2 * J]
            c[ti, tj, :, 1] = S[8 * ti + R, 8 * tj + 2 * J + 1]
    Ps = np.zeros(N * 4); Ws = np.zeros(N * 4)
    vv = np.zeros(32); vv[:N] = v
    for s in range(2 * NT):
        tk, half = s >> 1, s & 1
        for t in range(NT):
            for l in range(32):
                if t <= tk:
                    if (J[l] >> 1) == half:          # column access into T[t][tk]
                        Ps[(8 * t + R[l]) * 4 + 2 * (J[l] & 1) + 0] = c[t, tk, l, 0]
                        Ps[(8 * t + R[l]) * 4 + 2 * (J[l] & 1) + 1] = c[t, tk, l, 1]
                else:
                    if (R[l] >> 2) == half:          # row access into T[tk][t] (symmetry)
                        Ps[(8 * t + 2 * J[l] + 0) * 4 + (R[l] & 3)] = c[tk, t, l, 0]
                        Ps[(8 * t + 2 * J[l] + 1) * 4 + (R[l] & 3)] = c[tk, t, l, 1]
        X = np.zeros((32, 4))
        for l in range(N):
            X[l] = Ps[l * 4:l * 4 + 4]
        for cc in range(4):
            kc = 4 * s + cc
            pr = X[kc].copy(); prv = vv[kc]
            idv = 1.0 / pr[cc]
            for l in range(N):
                f = (1.0 - idv) if l == kc else X[l, cc] * idv
                for q in range(4):
                    if q != cc:
                        X[l, q] = X[l, q] - f * pr[q]
                vv[l] = vv[l] - f * prv
                X[l, cc] = -idv if l == kc else f
        for l in range(N):
            w = X[l].copy()
            if (l >> 2) == s:
                w[l & 3] += 1.0
                Ps[l * 4 + (l & 3)] -= 1.0
            Ws[l * 4:l * 4 + 4] = -w
        pf = np.zeros((NT, 32)); wf = np.zeros((NT, 32))
        for t in range(NT):
            pf[t] = Ps[(8 * t + R) * 4 + J]
            wf[t] = Ws[(8 * t + R) * 4 + J]
        for ti in range(NT):
            for tj in range(ti, NT):
                c[ti, tj] = dmma(c[ti, tj], wf[ti], pf[tj])
        for l in range(32):
            r = R[l]
            if (r >> 2) == half and J[l] == (r >> 1):
                c[tk, tk, l, r & 1] -= 2.0
    out = np.zeros((N, N))
    for ti in range(NT):
        for tj in range(ti, NT):
            out[8 * ti + R, 8 * tj + 2 * J] = c[ti, tj, :, 0]
            out[8 * ti + R, 8 * tj + 2 * J + 1] = c[ti, tj, :, 1]
            if tj > ti:
                out[8 * tj + 2 * J, 8 * ti + R] = c[ti, tj, :, 0]
                out[8 * tj + 2 * J + 1, 8 * ti + R] = c[ti, tj, :, 1]
    return out, vv[:N]


def warp_block_gj_neg(S, v, NT):
    """Same as warp_block_gj_sym in the sign convention the CUDA code uses: tiles hold T = -S, the panel sweeps run on
    y = -x, the update is T += (Y - E)(Y0 + E)' and +2 on the pivot block's diagonal.  Returns (inv(S), inv(S) @ v)."""
    N = 8 * NT
    c = np.zeros((NT, NT, 32, 2))
    for ti in range(NT):
        for tj in range(ti, NT):
            c[ti, tj, :, 0] = -S[8 * ti + R, 8 * tj + 2 * J]
            c[ti, tj, :, 1] = -S[8 * ti + R, 8 * tj + 2 * J + 1]
    Ps = np.zeros(N * 4); Ws = np.zeros(N * 4)
    vv = np.zeros(32); vv[:N] = v
    for s in range(2 * NT):
        tk, half = s >> 1, s & 1
        for t in range(NT):
            for l in range(32):
                if t <= tk:
                    if (J[l] >> 1) == half:
                        Ps[(8 * t + R[l]) * 4 + 2 * (J[l] & 1) + 0] = c[t, tk, l, 0]
                        Ps[(8 * t + R[l]) * 4 + 2 * (J[l] & 1) + 1] = c[t, tk, l, 1]
                else:
                    if (R[l] >> 2) == half:
                        Ps[(8 * t + 2 * J[l] + 0) * 4 + (R[l] & 3)] = c[tk, t, l, 0]
                        Ps[(8 * t + 2 * J[l] + 1) * 4 + (R[l] & 3)] = c[tk, t, l, 1]
        Yv = np.zeros((32, 4))
        for l in range(N):
            Yv[l] = Ps[l * 4:l * 4 + 4]
        for cc in range(4):
            kc = 4 * s + cc
            pr = Yv[kc].copy(); prv = vv[kc]
            idv = 1.0 / (-pr[cc])
            for l in range(N):
                g = (idv - 1.0) if l == kc else Yv[l, cc] * idv
                for q in range(4):
                    if q != cc:
                        Yv[l, q] = Yv[l, q] + g * pr[q]
                vv[l] = vv[l] + g * prv
                Yv[l, cc] = idv if l == kc else g
        for l in range(N):
            Ws[l * 4:l * 4 + 4] = Yv[l]
        pf = np.zeros((NT, 32)); wf = np.zeros((NT, 32))
        for t in range(NT):
            pf[t] = Ps[(8 * t + R) * 4 + J]
            wf[t] = Ws[(8 * t + R) * 4 + J]
        sel = R == 4 * half + J
        pf[tk][sel] += 1.0
        wf[tk][sel] -= 1.0
        for ti in range(NT):
            for tj in range(ti, NT):
                c[ti, tj] = dmma(c[ti, tj], wf[ti], pf[tj])
        for l in range(32):
            r = R[l]
            if (r >> 2) == half and J[l] == (r >> 1):
                c[tk, tk, l, r & 1] += 2.0
    out = np.zeros((N, N))
    for ti in range(NT):
        for tj in range(ti, NT):
            out[8 * ti + R, 8 * tj + 2 * J] = c[ti, tj, :, 0]
            out[8 * ti + R, 8 * tj + 2 * J + 1] = c[ti, tj, :, 1]
            if tj > ti:
                out[8 * tj + 2 * J, 8 * ti + R] = c[ti, tj, :, 0]
                out[8 * tj + 2 * J + 1, 8 * ti + R] = c[ti, tj, :, 1]
    return out, vv[:N]


def cta_block_gj(S, NTD, nsteps=None):
    """CTA-level variant for the single H x H inverses (H up to 128): full tile grid NTD x NTD of T = -S, every warp forms
    inv(D) of the 4 x 4 pivot block itself and builds its A fragment as Y0[row, :] @ inv(D)[:, j] (+ inv(D)[c][j] on pivot rows);
    B fragment = Y0 + E.  Returns inv(S)."""
    N = 8 * NTD
    c = np.zeros((NTD, NTD, 32, 2))
    for ti in range(NTD):
        for tj in range(NTD):
            c[ti, tj, :, 0] = -S[8 * ti + R, 8 * tj + 2 * J]
            c[ti, tj, :, 1] = -S[8 * ti + R, 8 * tj + 2 * J + 1]
    nsteps = 2 * NTD if nsteps is None else nsteps
    for s in range(nsteps):
        tk, half = s >> 1, s & 1
        Ps = np.zeros(N * 4)
        for t in range(NTD):                    # owners of tile (t, tk) publish their rows of the panel
            for l in range(32):
                if (J[l] >> 1) == half:
                    Ps[(8 * t + R[l]) * 4 + 2 * (J[l] & 1) + 0] = c[t, tk, l, 0]
                    Ps[(8 * t + R[l]) * 4 + 2 * (J[l] & 1) + 1] = c[t, tk, l, 1]
        D = -Ps[4 * (4 * s):4 * (4 * s) + 16].reshape(4, 4)
        a = D.copy()
        for cc in range(4):                     # 4 x 4 in-register sweep: a -> -inv(D)
            idv = 1.0 / a[cc, cc]
            rowc = a[cc].copy()
            for i in range(4):
                if i != cc:
                    f = a[i, cc] * idv
                    for q in range(4):
                        if q != cc:
                            a[i, q] -= f * rowc[q]
                    a[i, cc] = f
            for q in range(4):
                if q != cc:
                    a[cc, q] = rowc[q] * idv
            a[cc, cc] = -idv
        Dinv = -a
        for ti in range(NTD):
            wf = np.zeros(32)
            for l in range(32):
                row = 8 * ti + R[l]
                y0 = Ps[row * 4:row * 4 + 4]
                val = float(y0 @ Dinv[:, J[l]])
                if (row >> 2) == s:
                    val += Dinv[row & 3, J[l]]
                wf[l] = val
            for tj in range(NTD):
                pf = Ps[(8 * tj + R) * 4 + J].copy()
                rows = 8 * tj + R
                pf[rows == 4 * s + J] += 1.0
                c[ti, tj] = dmma(c[ti, tj], wf, pf)
        for l in range(32):
            r = R[l]
            if (r >> 2) == half and J[l] == (r >> 1):
                c[tk, tk, l, r & 1] += 2.0
    out = np.zeros((N, N))
    for ti in range(NTD):
        for tj in range(NTD):
            out[8 * ti + R, 8 * tj + 2 * J] = c[ti, tj, :, 0]
            out[8 * ti + R, 8 * tj + 2 * J + 1] = c[ti, tj, :, 1]
    return out


def warp_block_gj_v3(S, v, NT):
    """Third form (the one the CUDA code implements): as warp_block_gj_neg, but
      * the rank-4 update is assembled from the SEQUENTIAL sweeps' own terms: A fragment = multipliers g_c of sweep c,
        B fragment = column c of the panel right before its sweep (+1 on the pivot entry).  With X_final * P_original the
        update loses accuracy like kappa(D) of the 4 x 4 pivot block (tools/k4bench/acc_variants.py);
      * every lane evolves the 4 x 4 pivot block (and the pivot entries of v) itself from values read once from shared
        memory instead of receiving the current pivot row by shuffle in every sweep;
      * panel buffers are stored column-plane-wise, [col][row] with plane stride 36, so every access pattern is conflict free.
    Returns (inv(S), inv(S) @ v)."""
    N = 8 * NT
    PS = 36
    c = np.zeros((NT, NT, 32, 2))
    for ti in range(NT):
        for tj in range(ti, NT):
            c[ti, tj, :, 0] = -S[8 * ti + R, 8 * tj + 2 * J]
            c[ti, tj, :, 1] = -S[8 * ti + R, 8 * tj + 2 * J + 1]
    Ps = np.zeros(4 * PS); Ws = np.zeros(4 * PS); vs = np.zeros(32)
    vv = np.zeros(32); vv[:N] = v
    for s in range(2 * NT):
        tk, half = s >> 1, s & 1
        for t in range(NT):
            for l in range(32):
                if t <= tk:
                    if (J[l] >> 1) == half:
                        Ps[(2 * (J[l] & 1) + 0) * PS + 8 * t + R[l]] = c[t, tk, l, 0]
                        Ps[(2 * (J[l] & 1) + 1) * PS + 8 * t + R[l]] = c[t, tk, l, 1]
                else:
                    if (R[l] >> 2) == half:
                        Ps[(R[l] & 3) * PS + 8 * t + 2 * J[l] + 0] = c[tk, t, l, 0]
                        Ps[(R[l] & 3) * PS + 8 * t + 2 * J[l] + 1] = c[tk, t, l, 1]
        vs[:] = vv
        # every lane: the 4 x 4 pivot block (rows 4s..4s+3 of the panel) and the pivot entries of v
        a0 = np.array([[Ps[q * PS + 4 * s + i] for q in range(4)] for i in range(4)])
        vk0 = vs[4 * s:4 * s + 4].copy()
        for l in range(N):
            a = a0.copy(); vk = vk0.copy()
            y = np.array([Ps[q * PS + l] for q in range(4)])
            pq = np.zeros(4); gq = np.zeros(4)
            vl = vv[l]
            for cc in range(4):
                kc = 4 * s + cc
                pr = a[cc].copy(); prv = vk[cc]
                idv = 1.0 / (-pr[cc])
                piv = l == kc
                g = (idv - 1.0) if piv else y[cc] * idv
                pq[cc] = y[cc] + (1.0 if piv else 0.0)
                gq[cc] = g
                for q in range(4):
                    if q != cc:
                        y[q] = y[q] + g * pr[q]
                vl = vl + g * prv
                y[cc] = idv if piv else g
                for i in range(cc + 1, 4):            # the later pivot rows, evolved redundantly
                    gi = a[i, cc] * idv
                    for q in range(4):
                        if q != cc:
                            a[i, q] = a[i, q] + gi * pr[q]
                    vk[i] = vk[i] + gi * prv
                    a[i, cc] = gi
            vv[l] = vl
            for q in range(4):
                Ws[q * PS + l] = gq[q]
                Ps[q * PS + l] = pq[q]
        pf = np.zeros((NT, 32)); wf = np.zeros((NT, 32))
        for t in range(NT):
            pf[t] = Ps[J * PS + 8 * t + R]
            wf[t] = Ws[J * PS + 8 * t + R]
        for ti in range(NT):
            for tj in range(ti, NT):
                c[ti, tj] = dmma(c[ti, tj], wf[ti], pf[tj])
        for l in range(32):
            r = R[l]
            if (r >> 2) == half and J[l] == (r >> 1):
                c[tk, tk, l, r & 1] += 2.0
    out = np.zeros((N, N))
    for ti in range(NT):
        for tj in range(ti, NT):
            out[8 * ti + R, 8 * tj + 2 * J] = c[ti, tj, :, 0]
            out[8 * ti + R, 8 * tj + 2 * J + 1] = c[ti, tj, :, 1]
            if tj > ti:
                out[8 * tj + 2 * J, 8 * ti + R] = c[ti, tj, :, 0]
                out[8 * tj + 2 * J + 1, 8 * ti + R] = c[ti, tj, :, 1]
    return out, vv[:N]
