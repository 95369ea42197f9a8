"""Host-side code constructions needed to generate inputs for the hot path (setup only, numpy).

Bivariate-bicycle CSS codes as in the reference's src/codes_q.py:235-246
(create_bivariate_bicycle_codes) with a small GF(2) toolbox (row reduction, kernel, logicals;
reference: src/utils.py:309-375, src/codes_q.py:61-76).  The logical bases chosen here need not
equal the reference's: every consumer only tests whether ANY logical is flipped, which does not
depend on the basis.
"""
import numpy as np


def gf2_row_reduce(M):
    """Reduced row echelon form over GF(2). Returns (R, pivot_cols, rank, T) with T @ M = R (mod 2)."""
    A = (np.array(M, dtype=np.uint8) & 1).copy()
    m, n = A.shape
    T = np.eye(m, dtype=np.uint8)
    pivots = []
    r = 0
    for c in range(n):
        if r >= m:
            break
        rows = np.nonzero(A[r:, c])[0]
        if rows.size == 0:
            continue
        p = r + rows[0]
        if p != r:
            A[[r, p]] = A[[p, r]]
            T[[r, p]] = T[[p, r]]
        others = np.nonzero(A[:, c])[0]
        others = others[others != r]
        A[others] ^= A[r]
        T[others] ^= T[r]
        pivots.append(c)
        r += 1
    return A, pivots, r, T


def gf2_rank(M):
    return gf2_row_reduce(M)[2]


def gf2_kernel(M):
    """Basis (rows) of {x : M x = 0 mod 2}."""
    M = np.array(M, dtype=np.uint8) & 1
    m, n = M.shape
    R, piv, r, _ = gf2_row_reduce(M)
    free = [c for c in range(n) if c not in set(piv)]
    K = np.zeros((len(free), n), dtype=np.uint8)
    for i, f in enumerate(free):
        K[i, f] = 1
        for row, pc in enumerate(piv):
            if R[row, f]:
                K[i, pc] = 1
    return K


def _independent_extension(base, cand):
    """Rows of `cand` that extend span(base) (greedy, in order)."""
    stack = np.vstack([base, cand]).astype(np.uint8)
    _, piv, _, _ = gf2_row_reduce(stack.T)
    nb = base.shape[0]
    idx = [p - nb for p in piv if p >= nb]
    return cand[idx]


class CssCode:
    """hx, hz and derived objects (subset of the reference's css_code, codes_q.py:7-81)."""

    def __init__(self, hx, hz, name=""):
        self.hx = np.array(hx, dtype=np.int64) % 2
        self.hz = np.array(hz, dtype=np.int64) % 2
        assert self.hx.shape[1] == self.hz.shape[1]
        assert not ((self.hx @ self.hz.T) % 2).any(), "CSS constraint not satisfied"
        self.N = self.hx.shape[1]
        self.hx_perp = gf2_kernel(self.hx).astype(np.int64)     # kernel of hx
        self.hz_perp = gf2_kernel(self.hz).astype(np.int64)
        self.rank_hx = gf2_rank(self.hx)
        self.rank_hz = gf2_rank(self.hz)
        self.K = self.N - self.rank_hx - self.rank_hz
        # logicals: lz in ker(hx) \ rowspace(hz); lx in ker(hz) \ rowspace(hx)
        self.lz = _independent_extension(self.hz.astype(np.uint8), self.hx_perp.astype(np.uint8)).astype(np.int64)
        self.lx = _independent_extension(self.hx.astype(np.uint8), self.hz_perp.astype(np.uint8)).astype(np.int64)
        assert self.lz.shape[0] == self.K and self.lx.shape[0] == self.K
        self.name = name


def _cyclic_shift(l):
    S = np.zeros((l, l), dtype=np.int64)
    for i in range(l):
        S[i, (i + 1) % l] = 1
    return S


def create_bivariate_bicycle_codes(l, m, A_x_pows, A_y_pows, B_x_pows, B_y_pows, name=None):
    """Same signature / return as codes_q.py:235-246: (code, A_list, B_list); x = S_l (x) I_m, y = I_l (x) S_m."""
    x = np.kron(_cyclic_shift(l), np.eye(m, dtype=np.int64))
    y = np.kron(np.eye(l, dtype=np.int64), _cyclic_shift(m))

    def mpow(M, p):
        return np.linalg.matrix_power(M, p) % 2

    A_list = [mpow(x, p) for p in A_x_pows] + [mpow(y, p) for p in A_y_pows]
    B_list = [mpow(y, p) for p in B_y_pows] + [mpow(x, p) for p in B_x_pows]
    A = sum(A_list) % 2
    B = sum(B_list) % 2
    hx = np.hstack((A, B))
    hz = np.hstack((B.T, A.T))
    return CssCode(hx, hz, name=name or f"BB_n{hx.shape[1]}"), A_list, B_list


BB_PARAMS = {           # guessing.py:24-37
    72: (6, 6, [3], [1, 2], [1, 2], [3]),
    90: (15, 3, [9], [1, 2], [2, 7], [0]),
    108: (9, 6, [3], [1, 2], [1, 2], [3]),
    144: (12, 6, [3], [1, 2], [1, 2], [3]),
    288: (12, 12, [3], [2, 7], [1, 2], [3]),
    360: (30, 6, [9], [1, 2], [25, 26], [3]),
    756: (21, 18, [3], [10, 17], [3, 19], [5]),
}


def bb_code(N):
    return create_bivariate_bicycle_codes(*BB_PARAMS[N])
