"""ctypes binding of the C ABI (include/nbldpc_b200.h).  Thin: one Python call per C entry point.

Used by the tests, bench.py and smoke(); a C++ program links libnbldpc_b200.so directly.
The library is required: there is no Python or CPU fallback for any compute entry point.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

METHOD_QSPA, METHOD_EMS, METHOD_TEMS, METHOD_BSTEMS = 1, 2, 4, 7
PREC_F32, PREC_F64 = 0, 1
TIES_AUTO, TIES_FAST, TIES_REFERENCE, TIES_BITEXACT = 0, 1, 2, 3
KERNEL_AUTO, KERNEL_GROUP, KERNEL_WARP_GENERIC = 0, 1, 2
NUM_COUNTERS = 8
COUNTER_NAMES = ["frames", "err_frames", "err_syms", "err_bits", "converged", "undetected", "iterations", "crc_ok"]

_lib = None


class NblError(RuntimeError):
    pass


class CodeInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("q", "p", "N", "M", "E", "max_dv", "max_dc", "prim_poly")]


class DecoderConfig(C.Structure):
    _fields_ = [("method", C.c_int), ("max_iter", C.c_int), ("p0", C.c_int), ("p1", C.c_int), ("factor", C.c_double),
                ("offset", C.c_double), ("precision", C.c_int), ("early_stop", C.c_int), ("max_batch", C.c_int),
                ("device", C.c_int), ("tie_mode", C.c_int), ("kernel", C.c_int)]


class OsdConfig(C.Structure):
    _fields_ = [("order", C.c_int), ("factor", C.c_double), ("flag", C.c_int), ("crc_len", C.c_int),
                ("crc_correct_len", C.c_int), ("fix_matrepr", C.c_int)]


class SimConfig(C.Structure):
    _fields_ = [("ebn0_db", C.c_double), ("modulation", C.c_int), ("constellation", C.c_void_p), ("seed", C.c_uint64),
                ("first_frame", C.c_uint64), ("frames", C.c_uint64), ("codewords", C.c_void_p), ("n_codewords", C.c_int),
                ("random_msg", C.c_int), ("crc_len", C.c_int), ("punct_sym", C.c_void_p), ("n_punct", C.c_int)]


def lib_path():
    return _build.LIB


def lib():
    """Load (building first if the sources are newer) the shared library; raises if it cannot be had."""
    global _lib
    if _lib is None:
        path = os.environ.get("NBLDPC_LIB") or (_build.build() if os.path.isdir(_build.CSRC) else _build.LIB)   # override: A/B builds
        L = C.CDLL(path)
        vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.c_void_p   # data pointers may be device addresses
        L.nbl_last_error.restype = C.c_char_p
        L.nbl_gf_tables.argtypes = [C.c_int, C.c_int, ip, ip]
        L.nbl_gf_check_table_file.argtypes = [C.c_char_p, C.c_int]
        L.nbl_code_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip, ip]
        L.nbl_code_load_file.argtypes = [C.POINTER(vp), C.c_char_p, C.c_int]
        L.nbl_code_load_hbit.argtypes = [C.POINTER(vp), C.c_char_p, C.c_int, C.c_int]
        L.nbl_code_destroy.argtypes = [vp]
        L.nbl_code_destroy.restype = None
        L.nbl_code_get_info.argtypes = [vp, C.POINTER(CodeInfo)]
        L.nbl_code_get_graph.argtypes = [vp] + [ip] * 8
        L.nbl_code_puncture_set.argtypes = [vp, C.c_int, ip]
        L.nbl_decoder_create.argtypes = [C.POINTER(vp), vp, C.POINTER(DecoderConfig), vp]
        L.nbl_decoder_destroy.argtypes = [vp]
        L.nbl_decoder_destroy.restype = None
        L.nbl_decoder_kernel.argtypes = [vp]
        L.nbl_decoder_kernel.restype = C.c_char_p
        L.nbl_decoder_stream.argtypes = [vp]
        L.nbl_decoder_stream.restype = vp
        L.nbl_decoder_sync.argtypes = [vp]
        L.nbl_decode_batch.argtypes = [vp, dp, C.c_int, vp, vp, vp, dp]
        L.nbl_decoder_enable_osd.argtypes = [vp, C.POINTER(OsdConfig)]
        L.nbl_osd_batch.argtypes = [vp, dp, C.c_int, vp]
        L.nbl_demap_bpsk.argtypes = [vp, dp, C.c_int, C.c_double, ip, C.c_int, dp]
        L.nbl_demap_qam.argtypes = [vp, dp, C.c_int, dp, C.c_double, ip, C.c_int, dp]
        L.nbl_demap_decode_batch.argtypes = [vp, C.c_int, dp, C.c_int, dp, C.c_double, ip, C.c_int, vp, vp, vp]
        L.nbl_simulate.argtypes = [vp, C.POINTER(SimConfig), vp]
        L.nbl_simulate_async.argtypes = [vp, C.POINTER(SimConfig), vp]
        L.nbl_decoder_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.nbl_decoder_iter_timing.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
        L.nbl_decoder_iter_active.argtypes = [vp, ip, C.c_int]
        L.nbl_comm_create.argtypes = [C.POINTER(vp), C.c_int, ip]
        L.nbl_comm_unique_id.argtypes = [vp]
        L.nbl_comm_create_rank.argtypes = [C.POINTER(vp), C.c_int, C.c_int, vp, C.c_int]
        L.nbl_comm_destroy.argtypes = [vp]
        L.nbl_comm_destroy.restype = None
        L.nbl_comm_size.argtypes = [vp]
        L.nbl_comm_local_ranks.argtypes = [vp]
        L.nbl_comm_rank.argtypes = [vp, C.c_int]
        L.nbl_comm_device.argtypes = [vp, C.c_int]
        L.nbl_allreduce_counters.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
        L.nbl_simulate_sharded.argtypes = [vp, C.POINTER(vp), C.POINTER(SimConfig), C.c_uint64, C.c_uint64, vp]
        L.nbl_encoder_create.argtypes = [C.POINTER(vp), vp, C.c_int, vp]
        L.nbl_encoder_destroy.argtypes = [vp]
        L.nbl_encoder_destroy.restype = None
        L.nbl_encoder_get_info.argtypes = [vp, ip, ip]
        L.nbl_encode_batch.argtypes = [vp, vp, C.c_int, vp]
        _lib = L
    return _lib


def _check(rc):
    if rc < 0:
        raise NblError("nbldpc_b200 error %d: %s" % (rc, lib().nbl_last_error().decode()))
    return rc


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _ptr(a):
    """numpy array -> host address; int -> raw (device) address; torch tensor -> data_ptr(); None -> NULL"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if isinstance(a, int):
        return a
    return a.data_ptr()


def gf_tables(q, poly=0):
    mul = np.zeros((q, q), np.int32); inv = np.zeros(q, np.int32)
    _check(lib().nbl_gf_tables(q, poly, _ip(mul), _ip(inv)))
    return mul, inv


def gf_check_table_file(path, q):
    return _check(lib().nbl_gf_check_table_file(path.encode(), q))


class Code:
    """parity-check graph + GF(q) tables (host side until a Decoder is made from it)"""

    def __init__(self, handle):
        self._h = handle
        info = CodeInfo()
        _check(lib().nbl_code_get_info(self._h, C.byref(info)))
        for n, _ in CodeInfo._fields_:
            setattr(self, n, getattr(info, n))
        self.K = self.N - self.M

    @classmethod
    def from_file(cls, path, poly=0):
        h = C.c_void_p()
        _check(lib().nbl_code_load_file(C.byref(h), str(path).encode(), poly))
        return cls(h)

    @classmethod
    def from_hbit(cls, path, q, poly=0):
        h = C.c_void_p()
        _check(lib().nbl_code_load_hbit(C.byref(h), str(path).encode(), q, poly))
        return cls(h)

    @classmethod
    def from_checks(cls, q, N, chk, var=None, poly=0):
        """chk: per check row a list of (variable, coefficient); var: optional per variable list of (check, coef)"""
        M = len(chk)
        cptr = np.zeros(M + 1, np.int32)
        for r, l in enumerate(chk):
            cptr[r + 1] = cptr[r] + len(l)
        cvar = np.array([v for l in chk for v, _ in l], np.int32)
        ccoef = np.array([g for l in chk for _, g in l], np.int32)
        vptr = vchk = vcoef = None
        if var is not None:
            vptr = np.zeros(N + 1, np.int32)
            for i, l in enumerate(var):
                vptr[i + 1] = vptr[i] + len(l)
            vchk = np.array([r for l in var for r, _ in l], np.int32)
            vcoef = np.array([g for l in var for _, g in l], np.int32)
        h = C.c_void_p()
        _check(lib().nbl_code_create(C.byref(h), q, poly, N, M, _ip(cptr), _ip(cvar), _ip(ccoef), _ip(vptr), _ip(vchk), _ip(vcoef)))
        return cls(h)

    def graph(self):
        E = self.E
        a = {k: np.zeros(n, np.int32) for k, n in [("chk_ptr", self.M + 1), ("chk_var", E), ("chk_coef", E), ("chk_dv", E),
                                                   ("var_ptr", self.N + 1), ("var_chk", E), ("var_coef", E), ("var_dc", E)]}
        _check(lib().nbl_code_get_graph(self._h, *[_ip(a[k]) for k in ("chk_ptr", "chk_var", "chk_coef", "chk_dv", "var_ptr", "var_chk", "var_coef", "var_dc")]))
        return a

    def puncture_set(self, degree):
        pos = np.zeros(self.N, np.int32)
        n = _check(lib().nbl_code_puncture_set(self._h, degree, _ip(pos)))
        return pos[:n].copy()

    def close(self):
        if self._h:
            lib().nbl_code_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Encoder:
    """nbl_encoder: the reference's systematic encoder (CNBLDPC::InitialEncode / Encode) for batches, on the device"""

    def __init__(self, code, device=0, stream=None):
        self.code = code
        self._h = C.c_void_p()
        _check(lib().nbl_encoder_create(C.byref(self._h), code._h, device, stream))
        k, s = C.c_int(), C.c_int()
        _check(lib().nbl_encoder_get_info(self._h, C.byref(k), C.byref(s)))
        self.K, self.n_swaps = k.value, s.value

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().nbl_encoder_destroy(self._h)
            self._h = C.c_void_p()

    def encode(self, msg):
        """msg [B][K] ints -> code words [B][N] int32 (host arrays)"""
        msg = np.ascontiguousarray(msg, np.int32).reshape(-1, self.K)
        out = np.zeros((msg.shape[0], self.code.N), np.int32)
        _check(lib().nbl_encode_batch(self._h, _ptr(msg), msg.shape[0], _ptr(out)))
        return out


def sim_config(ebn0_db, frames, seed=1, first_frame=0, modulation=2, constellation=None, codewords=None, random_msg=0, crc_len=0, punct=None):
    """nbl_sim_config + the arrays it points into (keep them alive for the duration of the call)"""
    cs = np.ascontiguousarray(constellation, np.float64) if constellation is not None else None
    cw = np.ascontiguousarray(codewords, np.int32) if codewords is not None else None
    ps = np.ascontiguousarray(punct, np.int32) if punct is not None and len(punct) else None
    cfg = SimConfig(float(ebn0_db), modulation, _ptr(cs), seed, first_frame, frames, _ptr(cw), 0 if cw is None else cw.shape[0],
                    int(random_msg), int(crc_len), _ptr(ps), 0 if ps is None else len(ps))
    return cfg, (cs, cw, ps)


def _preload_nccl():
    """A Python process that will also import torch must end up with ONE NCCL: torch's bundled copy (same soname as the system's,
    newer API).  If torch is not loaded yet, map its libnccl.so.2 first so that the library's dlopen finds that copy already
    there; without torch installed the library falls back to the system's NCCL by itself."""
    import importlib.util
    import sys
    if "torch" in sys.modules:
        return
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            path = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(path):
                C.CDLL(path, mode=C.RTLD_GLOBAL)
                return
    except (ImportError, OSError, ValueError):
        pass


class Comm:
    """nbl_comm: the ranks whose Monte-Carlo counters are summed (NCCL).  Comm(n) = one process driving GPUs 0..n-1;
    Comm.from_rank(world, rank, id128, device) = one process per GPU (id128 from Comm.unique_id() on rank 0)."""

    def __init__(self, n_devices=1, devices=None, _handle=None):
        if _handle is not None:
            self._h = _handle
            return
        if n_devices > 1:
            _preload_nccl()
        h = C.c_void_p()
        dv = np.ascontiguousarray(devices, np.int32) if devices is not None else None
        _check(lib().nbl_comm_create(C.byref(h), n_devices, _ip(dv)))
        self._h = h

    @staticmethod
    def unique_id():
        _preload_nccl()
        buf = np.zeros(128, np.uint8)
        _check(lib().nbl_comm_unique_id(_ptr(buf)))
        return buf

    @classmethod
    def from_rank(cls, world, rank, id128, device):
        if world > 1:
            _preload_nccl()
        h = C.c_void_p()
        ident = np.ascontiguousarray(id128, np.uint8) if id128 is not None else None
        _check(lib().nbl_comm_create_rank(C.byref(h), world, rank, _ptr(ident), device))
        return cls(_handle=h)

    size = property(lambda self: lib().nbl_comm_size(self._h))
    local_ranks = property(lambda self: lib().nbl_comm_local_ranks(self._h))

    def rank(self, i=0):
        return lib().nbl_comm_rank(self._h, i)

    def device(self, i=0):
        return lib().nbl_comm_device(self._h, i)

    def allreduce_counters(self, decoders, dev_counters):
        """in-place sum over all ranks of each local rank's device uint64[8] (queued on the decoders' streams)"""
        n = len(decoders)
        hs = (C.c_void_p * n)(*[d._h for d in decoders])
        ps = (C.c_void_p * n)(*[_ptr(c) for c in dev_counters])
        _check(lib().nbl_allreduce_counters(self._h, hs, ps))

    def simulate_sharded(self, decoders, frames_per_rank, round_index, totals, **sim):
        """one round (nbl_simulate_sharded): totals (numpy uint64[8]) += the round's counters summed over all ranks"""
        n = len(decoders)
        hs = (C.c_void_p * n)(*[d._h for d in decoders])
        cfg, keep = sim_config(frames=0, **sim)
        _check(lib().nbl_simulate_sharded(self._h, hs, C.byref(cfg), frames_per_rank, round_index, _ptr(totals)))
        return totals

    def close(self):
        if self._h:
            lib().nbl_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Decoder:
    def __init__(self, code, method, max_iter=10, p0=0, p1=0, factor=1.0, offset=0.0, precision=PREC_F32, early_stop=1,
                 max_batch=1, device=0, stream=None, tie_mode=TIES_AUTO, kernel=KERNEL_AUTO):
        self.code = code
        self.cfg = DecoderConfig(method, max_iter, p0, p1, factor, offset, precision, early_stop, max_batch, device, tie_mode, kernel)
        h = C.c_void_p()
        _check(lib().nbl_decoder_create(C.byref(h), code._h, C.byref(self.cfg), stream))
        self._h = h
        self.max_batch = max_batch

    def decode(self, L_ch, want_post=False):
        """L_ch float64 [B][N][q-1] (numpy) -> dict(sym, converged, iterations[, post])"""
        N, q1 = self.code.N, self.code.q - 1
        L_ch = np.ascontiguousarray(L_ch, np.float64).reshape(-1, N, q1)
        B = L_ch.shape[0]
        sym = np.zeros((B, N), np.int32); conv = np.zeros(B, np.int32); it = np.zeros(B, np.int32)
        post = np.zeros((B, N, q1)) if want_post else None
        _check(lib().nbl_decode_batch(self._h, _ptr(L_ch), B, _ptr(sym), _ptr(conv), _ptr(it), _ptr(post)))
        out = dict(sym=sym, converged=conv, iterations=it)
        if want_post:
            out["post"] = post
        return out

    def kernel(self):
        """name of the iteration kernel this decoder launches ("tems_warp", ..., "group")"""
        return lib().nbl_decoder_kernel(self._h).decode()

    def enable_osd(self, order, factor=1.0, flag=0, crc_len=8, crc_correct_len=0, fix_matrepr=0):
        """OSD post-processing of failed frames in every later decode call (profile keys OSD_order, OSD_factor, OSD_flag,
        crcLen, crc_correctLen); order < 0 switches it off again"""
        cfg = OsdConfig(order, float(factor), flag, crc_len, crc_correct_len, fix_matrepr)
        _check(lib().nbl_decoder_enable_osd(self._h, C.byref(cfg)))

    def osd(self, L_ch):
        """decode method 6: OSD alone on the channel LLRs [B][N][q-1] -> symbols [B][N]"""
        N, q1 = self.code.N, self.code.q - 1
        L_ch = np.ascontiguousarray(L_ch, np.float64).reshape(-1, N, q1)
        sym = np.zeros((L_ch.shape[0], N), np.int32)
        _check(lib().nbl_osd_batch(self._h, _ptr(L_ch), L_ch.shape[0], _ptr(sym)))
        return sym

    def decode_raw(self, L_ch, B, out_sym, converged=None, iterations=None, L_post=None):
        """pointer-level call (numpy arrays, torch tensors or raw addresses; host or device)"""
        _check(lib().nbl_decode_batch(self._h, _ptr(L_ch), B, _ptr(out_sym), _ptr(converged), _ptr(iterations), _ptr(L_post)))

    def demap_bpsk(self, y_re, sigma, punct=None):
        N, q1 = self.code.N, self.code.q - 1
        y = np.ascontiguousarray(y_re, np.float64)
        B = y.shape[0]
        ps = np.asarray(punct if punct is not None else [], np.int32)
        out = np.zeros((B, N, q1))
        _check(lib().nbl_demap_bpsk(self._h, _ptr(y), B, float(sigma), _ip(ps), len(ps), _ptr(out)))
        return out

    def demap_qam(self, y, constellation, sigma, punct=None):
        N, q1 = self.code.N, self.code.q - 1
        y = np.ascontiguousarray(y, np.float64)
        B = y.shape[0]
        cs = np.ascontiguousarray(constellation, np.float64)
        ps = np.asarray(punct if punct is not None else [], np.int32)
        out = np.zeros((B, N, q1))
        _check(lib().nbl_demap_qam(self._h, _ptr(y), B, _ptr(cs), float(sigma), _ip(ps), len(ps), _ptr(out)))
        return out

    def demap_decode(self, modulation, y, sigma, constellation=None, punct=None):
        N = self.code.N
        y = np.ascontiguousarray(y, np.float64)
        B = y.shape[0]
        cs = np.ascontiguousarray(constellation, np.float64) if constellation is not None else None
        ps = np.asarray(punct if punct is not None else [], np.int32)
        sym = np.zeros((B, N), np.int32); conv = np.zeros(B, np.int32); it = np.zeros(B, np.int32)
        _check(lib().nbl_demap_decode_batch(self._h, modulation, _ptr(y), B, _ptr(cs), float(sigma), _ip(ps), len(ps),
                                            _ptr(sym), _ptr(conv), _ptr(it)))
        return dict(sym=sym, converged=conv, iterations=it)

    def demap_decode_raw(self, modulation, y, B, sigma, out_sym, converged=None, iterations=None, constellation=None):
        _check(lib().nbl_demap_decode_batch(self._h, modulation, _ptr(y), B, _ptr(constellation), float(sigma), None, 0,
                                            _ptr(out_sym), _ptr(converged), _ptr(iterations)))

    def simulate(self, ebn0_db, frames, seed=1, first_frame=0, modulation=2, constellation=None, codewords=None, counters=None, random_msg=0, crc_len=0, punct=None):
        cfg, keep = sim_config(ebn0_db, frames, seed, first_frame, modulation, constellation, codewords, random_msg, crc_len, punct)
        cnt = counters if counters is not None else np.zeros(NUM_COUNTERS, np.uint64)
        _check(lib().nbl_simulate(self._h, C.byref(cfg), _ptr(cnt)))
        return cnt

    def simulate_async(self, ebn0_db, frames, dev_counters, seed=1, first_frame=0, modulation=2, constellation=None, codewords=None, random_msg=0, crc_len=0, punct=None):
        """queues the whole call on the decoder's stream and returns; dev_counters (device int64/uint64[8]) is accumulated
        into.  Touch the counters on the same stream (create the Decoder with stream=torch's current stream, as bench.py
        does) or after sync() -- a private decoder stream is not ordered with torch's default stream."""
        cfg, keep = sim_config(ebn0_db, frames, seed, first_frame, modulation, constellation, codewords, random_msg, crc_len, punct)
        _check(lib().nbl_simulate_async(self._h, C.byref(cfg), _ptr(dev_counters)))

    def sync(self):
        _check(lib().nbl_decoder_sync(self._h))

    def stream(self):
        return lib().nbl_decoder_stream(self._h)

    def last_timing(self):
        ms = C.c_float(); n = C.c_int()
        _check(lib().nbl_decoder_last_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def iter_timing(self):
        ms = (C.c_float * 64)()
        n = _check(lib().nbl_decoder_iter_timing(self._h, ms, 64))
        return [ms[i] for i in range(n)]

    def iter_active(self):
        """frames still decoding when each iteration of the last decode ran"""
        a = np.zeros(64, np.int32)
        n = _check(lib().nbl_decoder_iter_active(self._h, _ip(a), 64))
        return a[:n].tolist()

    def close(self):
        if self._h:
            lib().nbl_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count():
    return lib().nbl_device_count()
