"""ctypes binding of the CPU oracle (oracle/librub_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (rub_mimo_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

SC_NULL, SC_PILOT, SC_DATA = 0, 1, 2
DET_ZF, DET_MMSE = 0, 1
EST_FULLBAND, EST_COMB = 0, 1
FLAG_Q1, FLAG_UNBIASED, FLAG_ZF_CHOLESKY = 1, 2, 4


class OrcConfig(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("M", "cp_len", "N", "nac", "D", "q", "detector", "estimator", "P", "flags")] + \
               [("noise_var", C.c_float), ("sctype", C.c_void_p)]


class OrcMseq(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("m", "g", "a", "n", "v", "b")]


class OrcFrameOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("eq", "llr", "bits", "rx_data", "G", "W", "gain", "isig", "counters")]


class OrcSyncResult(C.Structure):
    _fields_ = [("state", C.c_int), ("sync_index", C.c_uint64),
                ("num_samples_processed", C.c_uint64),
                ("plateau_start", C.c_uint64 * 8), ("plateau_end", C.c_uint64 * 8),
                ("corr_indices", C.c_void_p), ("s0_corr_index", C.c_int32 * 8),
                ("window_start", C.c_uint64), ("payload_start", C.c_int64),
                ("symbols_decoded", C.c_uint32)]


def _has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in line + " "
    except OSError:
        pass
    return False


def build(force=False):
    """Compile the oracle with gcc (plain C, no dependencies)."""
    so = os.path.join(_HERE, "librub_oracle.so")
    src = os.path.join(_HERE, "rub_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        name = "librub_oracle.so" if _has_fma() else "librub_oracle_nofma.so"
        L = C.CDLL(os.path.join(_HERE, name))
        L.orc_num_occupied.restype = C.c_uint32
        L.orc_num_training.restype = C.c_uint32
        L.orc_build_info.restype = C.c_char_p
        L.orc_invert_2x2.restype = C.c_float
        L.orc_demodulate.restype = C.c_uint32
        L.orc_mseq_advance.restype = C.c_uint32
        L.orc_mseq_symbol.restype = C.c_uint32
        L.orc_write_sync_words.restype = C.c_uint32
        L.orc_write_comb_words.restype = C.c_uint32
        L.orc_assemble_mimo_packet.restype = C.c_uint32
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Config:
    """Python-side view of orc_config; keeps the sctype array alive."""

    def __init__(self, M, cp_len, N, nac, D, q, detector=DET_ZF, estimator=EST_FULLBAND, P=8,
                 flags=0, noise_var=0.0, sctype=None):
        self.sctype = None if sctype is None else np.ascontiguousarray(sctype, dtype=np.uint8)
        self.c = OrcConfig(M, cp_len, N, nac, D, q, detector, estimator, P, flags,
                           float(noise_var), _p(self.sctype))
        for k in ("M", "cp_len", "N", "nac", "D", "q", "detector", "estimator", "P", "flags"):
            setattr(self, k, getattr(self.c, k))
        self.noise_var = float(noise_var)

    @property
    def L(self):
        return self.M + self.cp_len

    @property
    def Mo(self):
        return int(lib().orc_num_occupied(C.byref(self.c)))

    @property
    def T(self):
        return int(lib().orc_num_training(C.byref(self.c)))

    @property
    def row_bytes(self):
        return (self.Mo * self.q + 7) // 8


# ------------------------------------------------------------------ small wrappers -----
def fft_forward(x):
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty_like(x)
    lib().orc_fft_forward(C.c_uint32(x.size), _p(x), _p(out))
    return out


def fft_backward(x):
    x = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.empty_like(x)
    lib().orc_fft_backward(C.c_uint32(x.size), _p(x), _p(out))
    return out


def fft_twiddles(M):
    tw = np.empty(M, dtype=np.complex64)
    lib().orc_fft_twiddles(C.c_uint32(M), _p(tw))
    return tw


class Mseq:
    def __init__(self, m, g, a=1):
        self.ms = OrcMseq()
        lib().orc_mseq_init(C.byref(self.ms), m, g, a)

    def reset(self):
        lib().orc_mseq_reset(C.byref(self.ms))

    def advance(self):
        return int(lib().orc_mseq_advance(C.byref(self.ms)))

    def symbol(self, bps):
        return int(lib().orc_mseq_symbol(C.byref(self.ms), bps))


def init_default_sctype(M, use_all=True, add_null=True):
    p = np.empty(M, dtype=np.uint8)
    lib().orc_init_default_sctype(_p(p), C.c_uint32(M), int(use_all), int(add_null))
    return p


def validate_sctype(p):
    a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
    p = np.ascontiguousarray(p, dtype=np.uint8)
    rc = lib().orc_validate_sctype(_p(p), C.c_uint32(p.size), C.byref(a), C.byref(b), C.byref(c))
    if rc:
        raise ValueError("invalid subcarrier type")
    return a.value, b.value, c.value


def init_S0(p, M, ms):
    S0 = np.empty(M, np.complex64)
    s0 = np.empty(M, np.complex64)
    p = None if p is None else np.ascontiguousarray(p, dtype=np.uint8)
    rc = lib().orc_init_S0(_p(p), C.c_uint32(M), _p(S0), _p(s0), C.byref(ms.ms))
    if rc:
        raise ValueError("no subcarriers enabled")
    return S0, s0


def init_S1(p, M, nac, ms):
    S1 = np.empty((nac, M), np.complex64)
    s1 = np.empty((nac, M), np.complex64)
    p = None if p is None else np.ascontiguousarray(p, dtype=np.uint8)
    lib().orc_init_S1(_p(p), C.c_uint32(M), C.c_uint32(nac), _p(S1), _p(s1), C.byref(ms.ms))
    return S1, s1


def _ptr_array(rows):
    arr = (C.c_void_p * len(rows))()
    for i, r in enumerate(rows):
        arr[i] = r.ctypes.data
    return arr


def write_sync_words(cfg, s0, s1):
    total = (cfg.nac * cfg.N + 1) * cfg.L
    tx = np.zeros((cfg.N, total), np.complex64)
    s0 = np.ascontiguousarray(s0, np.complex64)
    s1 = np.ascontiguousarray(s1, np.complex64)
    n = lib().orc_write_sync_words(C.byref(cfg.c), _p(s0), _p(s1), _ptr_array(list(tx)))
    assert n == total
    return tx


def write_comb_words(cfg, S1):
    tx = np.zeros((cfg.N, cfg.nac * cfg.L), np.complex64)
    S1 = np.ascontiguousarray(S1, np.complex64)
    lib().orc_write_comb_words(C.byref(cfg.c), _p(S1), _ptr_array(list(tx)))
    return tx


def assemble_mimo_packet(cfg, syms):
    syms = np.ascontiguousarray(syms, np.complex64)  # [N][Mo]
    tx = np.zeros((cfg.N, cfg.L), np.complex64)
    lib().orc_assemble_mimo_packet(C.byref(cfg.c), _ptr_array(list(tx)), _ptr_array(list(syms)))
    return tx


def modulate(q, sym):
    class ocf(C.Structure):
        _fields_ = [("re", C.c_float), ("im", C.c_float)]
    f = lib().orc_modulate
    f.restype = ocf
    r = f(C.c_uint32(q), C.c_uint32(sym))
    return np.complex64(complex(r.re, r.im))


def modulate_table(q):
    return np.array([modulate(q, s) for s in range(1 << q)], dtype=np.complex64)


def demodulate(q, x):
    class ocf(C.Structure):
        _fields_ = [("re", C.c_float), ("im", C.c_float)]
    x = np.complex64(x)
    return int(lib().orc_demodulate(C.c_uint32(q), ocf(float(x.real), float(x.imag))))


def llr(q, x, isig=1.0):
    class ocf(C.Structure):
        _fields_ = [("re", C.c_float), ("im", C.c_float)]
    x = np.complex64(x)
    out = np.empty(q, np.float32)
    lib().orc_llr(C.c_uint32(q), ocf(float(x.real), float(x.imag)), C.c_float(isig), _p(out))
    return out


def invert_2x2(G):
    G = np.ascontiguousarray(G, np.complex64).reshape(4)
    W = np.empty(4, np.complex64)
    g = lib().orc_invert_2x2(_p(W), _p(G))
    return W.reshape(2, 2), float(g)


def weights(cfg, G):
    N = cfg.N
    G = np.ascontiguousarray(G, np.complex64).reshape(N * N)
    W = np.empty(N * N, np.complex64)
    gain = np.empty(N, np.float32)
    isig = np.empty(N, np.float32)
    lib().orc_weights(C.byref(cfg.c), _p(G), _p(W), _p(gain), _p(isig))
    return W.reshape(N, N), gain, isig


# ------------------------------------------------------------------ receive chain -----
def rx_batch(cfg, S1, iq, tx_data=None, first_sample=0, want=("eq", "llr", "bits", "rx_data", "G"),
             n_threads=1):
    """iq [F][N][row] complex64 (dense).  Returns dict of outputs (+ 'counters')."""
    iq = np.ascontiguousarray(iq, np.complex64)
    F = iq.shape[0]
    N, D, q, M, Mo = cfg.N, cfg.D, cfg.q, cfg.M, cfg.Mo
    S1 = np.ascontiguousarray(S1, np.complex64)
    out = {}
    out["eq"] = np.zeros((F, N, D, Mo), np.complex64) if "eq" in want else None
    out["llr"] = np.zeros((F, N, D, Mo, q), np.float32) if "llr" in want else None
    out["bits"] = np.zeros((F, N, D, cfg.row_bytes), np.uint8) if "bits" in want else None
    out["rx_data"] = np.zeros((F, N, D, Mo), np.uint8) if "rx_data" in want else None
    out["G"] = np.zeros((F, N, N, M), np.complex64) if "G" in want else None
    counters = np.zeros((N, 4), np.uint64)
    if tx_data is not None:
        tx_data = np.ascontiguousarray(tx_data, np.uint8)
    rc = lib().orc_rx_batch(C.byref(cfg.c), _p(S1), _p(iq), C.c_uint64(iq.shape[1] * iq.shape[2]),
                            C.c_uint64(iq.shape[2]), C.c_uint64(first_sample), C.c_uint32(F),
                            _p(tx_data), _p(out["eq"]), _p(out["llr"]), _p(out["bits"]),
                            _p(out["rx_data"]), _p(out["G"]), _p(counters), C.c_int(n_threads))
    if rc:
        raise RuntimeError("orc_rx_batch failed (unsupported config)")
    out["counters"] = counters
    return {k: v for k, v in out.items() if v is not None}


def rx_frame(cfg, S1, rows, first_sample=0, timing=None, payload_start=-1, tx_data=None):
    """One frame with optional timing table; rows = list of N complex64 arrays."""
    N, D, q, M, Mo = cfg.N, cfg.D, cfg.q, cfg.M, cfg.Mo
    rows = [np.ascontiguousarray(r, np.complex64) for r in rows]
    S1 = np.ascontiguousarray(S1, np.complex64)
    o = dict(eq=np.zeros((N, D, Mo), np.complex64), llr=np.zeros((N, D, Mo, q), np.float32),
             bits=np.zeros((N, D, cfg.row_bytes), np.uint8), rx_data=np.zeros((N, D, Mo), np.uint8),
             G=np.zeros((N, N, M), np.complex64), W=np.zeros((N, N, M), np.complex64),
             gain=np.zeros((N, M), np.float32), isig=np.zeros((N, M), np.float32),
             counters=np.zeros((N, 4), np.uint64))
    fo = OrcFrameOut(*[_p(o[k]) for k in ("eq", "llr", "bits", "rx_data", "G", "W", "gain", "isig",
                                           "counters")])
    if timing is not None:
        timing = np.ascontiguousarray(timing, np.int32)
    if tx_data is not None:
        tx_data = np.ascontiguousarray(tx_data, np.uint8)
    rc = lib().orc_rx_frame(C.byref(cfg.c), _p(S1), _ptr_array(rows), C.c_uint64(first_sample),
                            _p(timing), C.c_int64(payload_start), _p(tx_data), C.byref(fo))
    if rc:
        raise RuntimeError("orc_rx_frame failed (unsupported config)")
    return o


def sc_metric(M, x):
    x = np.ascontiguousarray(x, np.complex64)
    y = np.empty(x.size, np.float32)
    lib().orc_sc_metric(C.c_uint32(M), _p(x), C.c_uint64(x.size), _p(y))
    return y


def framesync_execute(cfg, S0, S1, capture, threshold=0.95):
    """Faithful state machine on a capture [N][num_samples]."""
    cap = [np.ascontiguousarray(r, np.complex64) for r in capture]
    N, D, M, Mo = cfg.N, cfg.D, cfg.M, cfg.Mo
    corr = np.zeros((N, cfg.nac * N), np.int32)
    res = OrcSyncResult()
    res.corr_indices = corr.ctypes.data
    eq = np.zeros((N, D, Mo), np.complex64)
    G = np.zeros((M, N, N), np.complex64)
    W = np.zeros((M, N, N), np.complex64)
    gain = np.zeros(Mo, np.float32)
    S0 = np.ascontiguousarray(S0, np.complex64)
    S1 = np.ascontiguousarray(S1, np.complex64)
    rc = lib().orc_framesync_execute(C.byref(cfg.c), _p(S0), _p(S1), _ptr_array(cap),
                                     C.c_uint64(cap[0].size), C.c_float(threshold), C.byref(res),
                                     _p(eq), _p(G), _p(W), _p(gain))
    return dict(rc=rc, state=res.state, sync_index=res.sync_index,
                num_samples_processed=res.num_samples_processed,
                plateau_start=list(res.plateau_start)[:N], plateau_end=list(res.plateau_end)[:N],
                corr_indices=corr, s0_corr_index=list(res.s0_corr_index)[:N],
                window_start=res.window_start, payload_start=res.payload_start,
                symbols_decoded=res.symbols_decoded, eq=eq, G=G, W=W, gain=gain)
