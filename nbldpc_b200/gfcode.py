"""Host-side GF(q) helpers for synthetic traffic: random valid codewords of a loaded code.

The device Monte-Carlo path takes a pool of codewords (nbl_sim_config.codewords); this module makes one by
Gaussian elimination of H over GF(q) (numpy + the library's multiplication table).  It is traffic generation,
not part of the timed hot path; the reference-compatible systematic encoder (CNBLDPC::Encode, NBLDPC.cpp:477-604)
lives in the C++ harness (nbldpc_b200/host/).
"""
import numpy as np

from . import capi


def parity_matrix(code):
    g = code.graph()
    H = np.zeros((code.M, code.N), np.int64)
    for r in range(code.M):
        for e in range(g["chk_ptr"][r], g["chk_ptr"][r + 1]):
            H[r, g["chk_var"][e]] = g["chk_coef"][e]
    return H


def _row_reduce(H, mul, inv):
    """reduced row echelon form over GF(q); returns (R, pivot_columns)"""
    R = H.copy()
    M, N = R.shape
    piv = []
    r = 0
    for c in range(N - 1, -1, -1):                 # pivots from the right: parity symbols at the end, like the reference
        rows = np.nonzero(R[r:, c])[0]
        if len(rows) == 0:
            continue
        p = r + rows[0]
        R[[r, p]] = R[[p, r]]
        R[r] = mul[inv[R[r, c]], R[r]]
        for rr in range(M):
            if rr != r and R[rr, c]:
                R[rr] ^= mul[R[rr, c], R[r]]
        piv.append(c)
        r += 1
        if r == M:
            break
    return R[:r], piv


def random_codewords(code, n, seed=0):
    """n random codewords [n][N] int32 with H c = 0 (uniform over the code)"""
    mul, inv = capi.gf_tables(code.q)
    mul = mul.astype(np.int64)
    H = parity_matrix(code)
    R, piv = _row_reduce(H, mul, inv)
    free = [c for c in range(code.N) if c not in set(piv)]
    rng = np.random.RandomState(seed)
    cw = np.zeros((n, code.N), np.int64)
    cw[:, free] = rng.randint(0, code.q, size=(n, len(free)))
    for i, c in enumerate(piv):                    # pivot symbol = sum over the free symbols of R[i][f] * x_f
        acc = np.zeros(n, np.int64)
        for f in free:
            if R[i, f]:
                acc ^= mul[R[i, f], cw[:, f]]
        cw[:, c] = acc
    assert not syndromes(code, cw, H, mul).any()
    return cw.astype(np.int32)


def syndromes(code, cw, H=None, mul=None):
    if mul is None:
        mul = capi.gf_tables(code.q)[0].astype(np.int64)
    if H is None:
        H = parity_matrix(code)
    cw = np.asarray(cw, np.int64)
    S = np.zeros((cw.shape[0], code.M), np.int64)
    for r in range(code.M):
        for c in np.nonzero(H[r])[0]:
            S[:, r] ^= mul[H[r, c], cw[:, c]]
    return S
