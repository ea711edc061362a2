"""Lane-level NumPy emulation of the register-resident Riccati knot of csrc/riccati_frag.cu.

Test infrastructure (not product code): it restates, lane by lane, what one warp of `k_riccati_frag` does with the
FP64 tensor-core instruction mma.sync.m8n8k4 (fragment ownership: lane L = 4*fr + fc holds A[fr][fc], B[fc][fr],
D[fr][2fc], D[fr][2fc+1]) so that the index algebra of the kernel -- the "physical" ordering of z = [x_e; u], the
k-step classes, the row-0 / column-0 substitutions that carry s, Qz, d and w_d through the same MMAs -- can be
checked on the CPU against a plain dense recursion (tests/test_frag_layout.py).

Physical index p in 0..15:  u_a <-> p = 2a (class (0,0));  x_e <-> p = PHYS_X[e]; the contraction index of k-step ks is
   ks = 0: q = 2fc+1 (tile 0, reg 1)      ks = 1: q = 8+2fc (tile 1, reg 0)      ks = 2: q = 9+2fc (tile 1, reg 1)
with the attitude error e = 3,4,5 on p = 8,10,12 (csrc/frag_layout.cuh).
"""
import numpy as np

NE, M, NZ = 12, 4, 16
PHYS_X = [1, 3, 5, 8, 10, 12, 7, 14, 9, 11, 13, 15]          # logical error-state index e (r, phi, v, omega) -> physical p
PHYS_U = [0, 2, 4, 6]
PHYS = PHYS_X + PHYS_U                                       # logical z index (x_e then u_a) -> physical
KS_TILE = [(0, 1), (1, 0), (1, 1)]                           # k-step -> (column tile, register) of a D fragment
QPHYS = [[2 * fc + 1 for fc in range(4)], [8 + 2 * fc for fc in range(4)], [9 + 2 * fc for fc in range(4)]]


def lanes():
    for L in range(32):
        yield L, L >> 2, L & 3


def dmma(D, A, B):
    """D (2 regs per lane) += A (1 per lane) x B (1 per lane), m8n8k4 fragment semantics."""
    Am = np.zeros((8, 4)); Bm = np.zeros((4, 8))
    for L, fr, fc in lanes():
        Am[fr, fc] = A[L]; Bm[fc, fr] = B[L]
    C = Am @ Bm
    for L, fr, fc in lanes():
        D[L, 0] += C[fr, 2 * fc]; D[L, 1] += C[fr, 2 * fc + 1]


def frag_from_matrix(Mx):
    """16 x 16 physical matrix -> accumulator fragments acc[mi][ni][lane][reg]."""
    acc = np.zeros((2, 2, 32, 2))
    for mi in range(2):
        for ni in range(2):
            for L, fr, fc in lanes():
                for r in range(2):
                    acc[mi, ni, L, r] = Mx[8 * mi + fr, 8 * ni + 2 * fc + r]
    return acc


def matrix_from_frag(acc):
    Mx = np.zeros((16, 16))
    for mi in range(2):
        for ni in range(2):
            for L, fr, fc in lanes():
                for r in range(2):
                    Mx[8 * mi + fr, 8 * ni + 2 * fc + r] = acc[mi, ni, L, r]
    return Mx


def ab_fragments(AB):
    """[A_e B_e] (12 x 16 logical) -> abf[ks][mi][lane] = ABp[q(ks,fc)][8mi+fr] (the record order of the kernel)."""
    ABp = np.zeros((16, 16))                      # rows and columns in physical order
    for e in range(12):
        for j in range(16):
            ABp[PHYS_X[e], PHYS[j]] = AB[e, j]
    abf = np.zeros((3, 2, 32))
    for ks in range(3):
        for mi in range(2):
            for L, fr, fc in lanes():
                abf[ks, mi, L] = ABp[QPHYS[ks][fc], 8 * mi + fr]
    return abf


def phys_vec(v):
    out = np.zeros(16)
    for j in range(16):
        out[PHYS[j]] = v[j]
    return out


def phys_mat(Hm):
    out = np.zeros((16, 16))
    for i in range(16):
        for j in range(16):
            out[PHYS[i], PHYS[j]] = Hm[i, j]
    return out


def terminal_state(HN, gN, junk=0.37):
    """Shat of the terminal knot: S = HN (12 x 12 logical), s = gN in row 0; junk elsewhere in the u rows / columns."""
    Sh = np.full((16, 16), junk)
    for e in range(12):
        for f in range(12):
            Sh[PHYS_X[e], PHYS_X[f]] = HN[e, f]
        Sh[0, PHYS_X[e]] = gN[e]
    return frag_from_matrix(Sh)


def tile_transpose(X):
    """the 4-shuffle transposition of an 8 x 8 tile in D-fragment form (csrc/riccati_frag.cu tile_transpose)"""
    Y = np.zeros((32, 2))
    for L, fr, fc in lanes():
        src = 8 * fc + (fr >> 1)
        a, b = X[src], X[src + 4]
        Y[L, 0] = a[fr & 1]; Y[L, 1] = b[fr & 1]
    return Y


def knot(Sacc, AB, H, g, rho, symmetrise_diag=True):
    """One backward step in fragment form.  Sacc: fragments of Shat_{k+1}; AB 12 x 16, H 16 x 16, g 16 (logical z order).
    Returns (Sacc', K 4 x 12, d 4, dV1 term d'Qu, sum d^2, ok)."""
    abf = ab_fragments(AB)
    gp = phys_vec(g)
    Hacc = frag_from_matrix(phys_mat(H))
    # phase 1: T'[c][j] = sum_q ABp[q][c] Shat[j][q]
    Tacc = np.zeros((2, 2, 32, 2))
    for mi in range(2):
        for nj in range(2):
            for ks in range(3):
                ni, reg = KS_TILE[ks]
                dmma(Tacc[mi, nj], abf[ks, mi], Sacc[nj, ni, :, reg])
    # phase 2: Qhat[i][c] = Hhat + sum_r ABp[r][i] T'[c][r]
    Qacc = Hacc.copy()
    Qacc[0, 1] = np.nan                                # tile (0,1) is never computed nor used
    for mi, nc in ((0, 0), (1, 0), (1, 1)):
        for ks in range(3):
            ni, reg = KS_TILE[ks]
            dmma(Qacc[mi, nc], abf[ks, mi], Tacc[nc, ni, :, reg])
    # Qz in column form: lane (fr, 0) holds (AB's)[8mi+fr] in Tacc[mi][0][.][0]
    qz_col = np.zeros((2, 32))
    for mi in range(2):
        for L, fr, fc in lanes():
            if fc == 0:
                qz_col[mi, L] = gp[8 * mi + fr] + Tacc[mi, 0, L, 0]
    # phase 3: Quu from tile (0,0) reg 0 of lanes (2a, b)
    Quu = np.zeros((4, 4))
    for L, fr, fc in lanes():
        if fr % 2 == 0:
            Quu[fr // 2, fc] = Qacc[0, 0, L, 0]
    Quu = 0.5 * (Quu + Quu.T)
    Mr = Quu + rho * np.eye(4)
    try:
        np.linalg.cholesky(Mr)
        ok = True
    except np.linalg.LinAlgError:
        ok = False
    Minv = np.linalg.inv(Mr)
    bfrag = np.zeros(32)
    for L, fr, fc in lanes():
        a = fr >> 1
        bfrag[L] = ((1.0 if fc == a else 0.0) + rho * Minv[fc, a]) if (fr & 1) else -Minv[fc, a]
    kf = np.zeros((2, 32)); wf = np.zeros((2, 32))
    for mi in range(2):
        afrag = Qacc[mi, 0, :, 0].copy()
        if mi == 0:
            for L, fr, fc in lanes():
                if fr == 0:
                    afrag[L] = qz_col[0, 8 * fc]          # Qu[fc] = Qz[p = 2 fc] from lane (2fc, 0)
        D = np.zeros((32, 2))
        dmma(D, afrag, bfrag)
        kf[mi] = D[:, 0]; wf[mi] = D[:, 1]
    # gains out (logical K[a][e], d[a]) and the dV partials of lanes (0, fc)
    K = np.zeros((4, 12)); d = np.zeros(4); Qu = np.zeros(4)
    for mi in range(2):
        for L, fr, fc in lanes():
            p = 8 * mi + fr
            if p in PHYS_X:
                K[fc, PHYS_X.index(p)] = kf[mi, L]
            if p == 0:
                d[fc] = kf[mi, L]; Qu[fc] = qz_col[0, 8 * fc]
    # phase 4: column 0 <- Qz (column form, in place), row 0 of tile (0,0) <- Qz (two shuffles), S' = Q + W'K on three tiles,
    # tile (0,1) = transpose of tile (1,0) (its row 0 is then s' for p = 8..15), diagonal tiles averaged with their transposes
    for mi in range(2):
        for L, fr, fc in lanes():
            if fc == 0:
                Qacc[mi, 0, L, 0] = qz_col[mi, L]
    for L, fr, fc in lanes():
        if fr == 0:
            for r in range(2):
                Qacc[0, 0, L, r] = qz_col[0, 4 * (2 * fc + r)]           # from lane (2fc + r, 0)
    Snew = Qacc.copy()
    for mi, nj in ((0, 0), (1, 0), (1, 1)):
        dmma(Snew[mi, nj], wf[mi], kf[nj])
    Snew[0, 1] = tile_transpose(Snew[1, 0])
    if symmetrise_diag:
        for t in (0, 1):
            Snew[t, t] = 0.5 * (Snew[t, t] + tile_transpose(Snew[t, t]))
    return Snew, K, d, float(d @ Qu), float(d @ d), ok


def dense_reference(S, s, AB, H, g, rho):
    """Plain recursion (oracle/oracle.hpp backward_pass, one knot) on logical indices."""
    Q = H + AB.T @ S @ AB
    qz = g + AB.T @ s
    Quu = Q[12:, 12:]; Qux = Q[12:, :12]; Qu = qz[12:]
    Mr = Quu + rho * np.eye(4)
    K = -np.linalg.solve(Mr, Qux); d = -np.linalg.solve(Mr, Qu)
    W = Qux - rho * K
    Sn = Q[:12, :12] + W.T @ K
    Sn = 0.5 * (Sn + Sn.T)
    sn = qz[:12] + W.T @ d
    return Sn, sn, K, d, float(d @ Qu), float(0.5 * d @ Quu @ d)
