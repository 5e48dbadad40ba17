"""ctypes binding of oracle/liboracle.so + small text parsers used by the tests (test infrastructure)."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PRIM_POLY = {4: 7, 8: 11, 16: 19, 32: 37, 64: 67, 128: 137, 256: 285, 512: 529}  # Arith.Table headers
_lib = None


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "nbldpc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        _lib.orc_code_create.restype = C.c_void_p
        _lib.orc_code_create.argtypes = [C.c_int] * 4 + [ip, ip, ip, ip, C.c_int, ip, ip, C.c_int]
        _lib.orc_code_destroy.argtypes = [C.c_void_p]
        _lib.orc_decode.restype = C.c_int
        _lib.orc_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                    dp, ip, dp, ip]
        _lib.orc_demap_bpsk.argtypes = [C.c_int, C.c_int, dp, C.c_double, ip, C.c_int, dp]
        _lib.orc_demap_qam.argtypes = [C.c_int, C.c_int, dp, dp, C.c_double, ip, C.c_int, dp]
        _lib.orc_gf_tables.argtypes = [C.c_int, C.c_int, ip, ip]
        _lib.orc_last_tems_leaves.restype = C.c_double
        _lib.orc_set_ems_fresh_sums.argtypes = [C.c_int]
        _lib.orc_osd_create.restype = C.c_void_p
        _lib.orc_osd_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.orc_osd_destroy.argtypes = [C.c_void_p]
        _lib.orc_osd_run.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp, ip]
        _lib.orc_osd_last.argtypes = [C.c_void_p, ip, ip]
        _lib.orc_decode_osd.restype = C.c_int
        _lib.orc_decode_osd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                        dp, ip, dp, ip, C.c_void_p, C.c_int, C.c_double, C.c_int]
    return _lib


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def parse_code_text(text):
    """alist-like code file of the reference (parsed there at NBLDPC.cpp:147-205); indices to 0-based."""
    tok = [int(t) for t in text.split()]
    it = iter(tok)
    N, M, q = next(it), next(it), next(it)
    maxdv, maxdc = next(it), next(it)
    vdeg = [next(it) for _ in range(N)]
    cdeg = [next(it) for _ in range(M)]
    var = [[(next(it) - 1, next(it)) for _ in range(vdeg[i])] for i in range(N)]
    chk = [[(next(it) - 1, next(it)) for _ in range(cdeg[r])] for r in range(M)]
    rest = list(it)
    assert not rest, "leftover tokens in code file"
    return dict(N=N, M=M, q=q, maxdv=maxdv, maxdc=maxdc, vdeg=vdeg, cdeg=cdeg, var=var, chk=chk)


def code_text(code):
    """inverse of parse_code_text: emit the reference's code-file format."""
    out = ["%d %d %d" % (code["N"], code["M"], code["q"]), "%d %d" % (max(code["vdeg"]), max(code["cdeg"]))]
    out.append(" ".join(str(d) for d in code["vdeg"]) + " ")
    out.append(" ".join(str(d) for d in code["cdeg"]) + " ")
    for links in code["var"]:
        out.append(" ".join("%d %d" % (r + 1, g) for r, g in links) + " ")
    for links in code["chk"]:
        out.append(" ".join("%d %d" % (v + 1, g) for v, g in links) + " ")
    return "\n".join(out) + "\n"


def code_from_checks(N, M, q, chk):
    """build both edge lists from the check-side list (variable side ordered by ascending check)."""
    var = [[] for _ in range(N)]
    for r, links in enumerate(chk):
        for v, g in links:
            var[v].append((r, g))
    return dict(N=N, M=M, q=q, maxdv=max(map(len, var)), maxdc=max(map(len, chk)),
                vdeg=[len(l) for l in var], cdeg=[len(l) for l in chk], var=var, chk=[list(l) for l in chk])


class OracleCode:
    def __init__(self, code):
        self.code = code
        N, M, q = code["N"], code["M"], code["q"]
        self.N, self.M, self.q = N, M, q
        sv, sc = max(code["vdeg"]), max(code["cdeg"])
        vl = np.zeros((N, sv), np.int32); vc = np.zeros((N, sv), np.int32)
        cl = np.zeros((M, sc), np.int32); cc = np.zeros((M, sc), np.int32)
        for i, links in enumerate(code["var"]):
            for d, (r, g) in enumerate(links):
                vl[i, d], vc[i, d] = r, g
        for r, links in enumerate(code["chk"]):
            for d, (v, g) in enumerate(links):
                cl[r, d], cc[r, d] = v, g
        vd = np.asarray(code["vdeg"], np.int32); cd = np.asarray(code["cdeg"], np.int32)
        self._h = lib().orc_code_create(q, PRIM_POLY[q], N, M, _ip(vd), _ip(cd), _ip(vl), _ip(vc), sv,
                                        _ip(cl), _ip(cc), sc)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_code_destroy(self._h)
            self._h = None

    def decode(self, method, max_iter, p0, p1, factor, offset, L_ch, osd=None, osd_order=-1, osd_alpha=1.0, osd_flag=0):
        """L_ch [N][q-1] float64 -> (ret, iters, out_sym[N], L_post[N][q-1]); osd = an OracleOsd of this code"""
        L_ch = np.ascontiguousarray(L_ch, np.float64).reshape(self.N, self.q - 1)
        out = np.zeros(self.N, np.int32)
        post = np.zeros_like(L_ch)
        it = C.c_int(0)
        ret = lib().orc_decode_osd(self._h, method, max_iter, p0, p1, float(factor), float(offset),
                                   _dp(L_ch), _ip(out), _dp(post), C.byref(it),
                                   osd._h if osd is not None else None, osd_order, float(osd_alpha), osd_flag)
        return ret, it.value, out, post


class OracleOsd:
    """bit-level OSD post-processor of one code (OSD.h): CRC rows + binary image of H, built once"""

    def __init__(self, oc, crc_len=8, crc_rate=0, fix_matrepr=0):
        self.oc = oc
        self._h = lib().orc_osd_create(oc._h, crc_len, crc_rate, fix_matrepr)
        if not self._h:
            raise ValueError("unsupported CRC configuration")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_osd_destroy(self._h)
            self._h = None

    def run(self, order, L_ch, sym=None, sym_sum=None):
        """stand-alone call: sym/sym_sum None = the reference's flag 1 (method 6)"""
        N, q1 = self.oc.N, self.oc.q - 1
        L_ch = np.ascontiguousarray(L_ch, np.float64).reshape(N, q1)
        flag = sym is None
        out = np.zeros(N, np.int32) if flag else np.ascontiguousarray(sym, np.int32).copy()
        ss = L_ch if flag else np.ascontiguousarray(sym_sum, np.float64).reshape(N, q1)
        lib().orc_osd_run(self._h, order, int(flag), _dp(L_ch), _dp(ss), _ip(out))
        return out

    def last(self):
        p = self.oc.q.bit_length() - 1
        order = np.zeros(self.oc.N * p, np.int32)
        dist = C.c_int(0)
        lib().orc_osd_last(self._h, _ip(order), C.byref(dist))
        return order, dist.value


def set_ems_fresh_sums(on):
    """analysis switch (off = the reference's arithmetic): EMS configuration sums formed afresh, not in the running add / subtract sum"""
    lib().orc_set_ems_fresh_sums(int(bool(on)))


def demap_bpsk(q, N, y_re, sigma, punct=()):
    y = np.ascontiguousarray(y_re, np.float64)
    ps = np.asarray(list(punct), np.int32)
    out = np.zeros((N, q - 1))
    lib().orc_demap_bpsk(q, N, _dp(y), float(sigma), _ip(ps), len(ps), _dp(out))
    return out


def demap_qam(q, N, y, constellation, sigma, punct=()):
    y = np.ascontiguousarray(y, np.float64)
    cs = np.ascontiguousarray(constellation, np.float64)
    ps = np.asarray(list(punct), np.int32)
    out = np.zeros((N, q - 1))
    lib().orc_demap_qam(q, N, _dp(y), _dp(cs), float(sigma), _ip(ps), len(ps), _dp(out))
    return out


def gf_tables(q, poly=None):
    mul = np.zeros((q, q), np.int32); inv = np.zeros(q, np.int32)
    lib().orc_gf_tables(q, poly or PRIM_POLY[q], _ip(mul), _ip(inv))
    return mul, inv


def mat_repr(q, g, poly=None):
    """p x p binary matrix of the field element g as Mat.Repr.GF.q.txt lists it: M[r][c] = bit c of g*x^r."""
    p = q.bit_length() - 1
    poly = poly or PRIM_POLY[q]
    M = np.zeros((p, p), np.int32)
    v = g
    for r in range(p):
        for c in range(p):
            M[r, c] = (v >> c) & 1
        v <<= 1
        if v & q:
            v ^= poly
    return M


def hbit_to_code(rows, q):
    """Binary image (H_bit_GF*.txt; no reference loader exists, layout per SURVEY.md A.6) -> code dict.

    rows > cols means the file holds the transposed image B[p*sym+l][p*chk+k] = M_g[l][k];
    otherwise it holds H_bin[p*chk+k][p*sym+l] = M_g[l][k] (the construction at NBLDPC.cpp:420-430)."""
    p = q.bit_length() - 1
    B = np.array([[int(ch) for ch in r] for r in rows], np.int32)
    if B.shape[0] < B.shape[1]:
        B = B.T                      # now B[p*sym+l][p*chk+k]
    N, M = B.shape[0] // p, B.shape[1] // p
    chk = [[] for _ in range(M)]
    for s in range(N):
        for r in range(M):
            blk = B[s * p:(s + 1) * p, r * p:(r + 1) * p]
            if blk.any():
                g = int(sum(int(blk[0, c]) << c for c in range(p)))
                assert (mat_repr(q, g) == blk).all(), "block is not a field-element matrix"
                chk[r].append((s, g))
    return code_from_checks(N, M, q, chk)
