"""rub_mimo_b200 — B200-native MIMO-OFDM receive path behind the RUB_MIMO framing API.

The product is the C-ABI shared library ``librubmimo_b200.so`` (CUDA kernels for sm_100a + host
setup code, ``include/rub_mimo/rub_mimo.h``) and the framing.h-compatible C++ facade
(``include/rub_mimo/framing.h``).  This Python package is thin plumbing over the C ABI for
tests and benchmarks: ctypes bindings, numpy for host buffers and torch only for device memory,
streams and torch.distributed.

There is no CPU fallback: importing works anywhere (so the ABI can be inspected), but creating a
``Receiver`` without a CUDA device raises, and a missing library raises at import of ``lib()``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RUB_MIMO_LIB") or os.path.join(_HERE, "librubmimo_b200.so")  # override: A/B builds

# ---- constants mirrored from include/rub_mimo/rub_mimo.h -------------------------------
OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_CUDA, ERR_NOMEM, ERR_NCCL, ERR_IO = range(8)
SCTYPE_NULL, SCTYPE_PILOT, SCTYPE_DATA = 0, 1, 2
MOD_QPSK, MOD_QAM16, MOD_QAM64, MOD_QAM256 = 2, 4, 6, 8
DET_ZF, DET_MMSE = 0, 1
EST_LS_FULLBAND, EST_LS_COMB_INTERP = 0, 1
FLAG_Q1_IDENTITY_INIT, FLAG_MMSE_UNBIASED, FLAG_ZF_CHOLESKY = 1, 2, 4
OUT_EQ, OUT_LLR, OUT_BITS, OUT_RXDATA, OUT_G = 1, 2, 4, 8, 16
PATH_AUTO, PATH_STAGED, PATH_FUSED = 0, 1, 2
SYNC_FIR, SYNC_SCAN = 0, 1
NCCL_UNIQUE_ID_BYTES = 128


class RubError(RuntimeError):
    def __init__(self, status, detail):
        super().__init__(f"rub status {status}: {detail}")
        self.status = status


class rub_config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("M", C.c_uint32), ("cp_len", C.c_uint32),
                ("num_streams", C.c_uint32), ("num_access_codes", C.c_uint32),
                ("num_data_symbols", C.c_uint32), ("modulation", C.c_uint32),
                ("detector", C.c_uint32), ("estimator", C.c_uint32),
                ("pilot_spacing", C.c_uint32), ("flags", C.c_uint32), ("noise_var", C.c_float),
                ("sctype", C.c_void_p)]


class rub_iq_layout(C.Structure):
    _fields_ = [("frame_stride", C.c_uint64), ("rx_stride", C.c_uint64),
                ("first_sample", C.c_uint64)]


class rub_rx_io(C.Structure):
    _fields_ = [("iq", C.c_void_p), ("layout", rub_iq_layout), ("tx_data", C.c_void_p),
                ("timing", C.c_void_p), ("payload_start", C.c_void_p), ("eq", C.c_void_p),
                ("llr", C.c_void_p), ("bits", C.c_void_p), ("rx_data", C.c_void_p),
                ("G", C.c_void_p), ("counters", C.c_void_p), ("out_mask", C.c_uint32)]


class rub_msequence(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("m", "g", "a", "n", "v", "b")]


class rub_frontend_options(C.Structure):
    _fields_ = [("cent_freq", C.c_double), ("samp_rate", C.c_double), ("txgain", C.c_double), ("rxgain", C.c_double),
                ("dsp_gain", C.c_float), ("rx_addr", C.c_char * 64), ("tx_addr", C.c_char * 64),
                ("rx_subdev", C.c_char * 32), ("tx_subdev", C.c_char * 32), ("verbose", C.c_int32),
                ("help", C.c_int32), ("num_nullcarriers", C.c_uint32)]


def config_from_args(argv, cfg=None, fe=None):
    """read_options (mimo/main.cc:174-240) over the C ABI: returns (Config, rub_frontend_options)."""
    cfg = cfg or Config()
    fe = fe or rub_frontend_options(dsp_gain=0.25)
    arr = (C.c_char_p * len(argv))(*[a.encode() for a in argv])
    c = cfg.c
    _check(lib().rub_config_from_args(len(argv), arr, C.byref(c), C.byref(fe)))
    return cfg.replace(M=c.M, cp_len=c.cp_len), fe


def config_from_json(text, cfg=None, fe=None):
    """One GUI device record (Interface/usrp_device.cpp:13-29) -> (Config, rub_frontend_options)."""
    cfg = cfg or Config()
    fe = fe or rub_frontend_options(dsp_gain=0.25)
    c = cfg.c
    _check(lib().rub_config_from_json(text.encode(), C.byref(c), C.byref(fe)))
    return cfg.replace(M=c.M, cp_len=c.cp_len, num_access_codes=c.num_access_codes), fe


class rub_file_job(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_frames", C.c_uint32), ("first_sample", C.c_uint64),
                ("frame_stride", C.c_uint64), ("rx_paths", C.POINTER(C.c_char_p)),
                ("tx_data_paths", C.POINTER(C.c_char_p)), ("eq_paths", C.POINTER(C.c_char_p)),
                ("rx_data_paths", C.POINTER(C.c_char_p)), ("llr_path", C.c_char_p), ("bits_path", C.c_char_p),
                ("chunk_frames", C.c_uint32)]


class rub_synth_params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("first_frame", C.c_uint64), ("n_taps", C.c_uint32),
                ("snr_db", C.c_float), ("baseband_gain", C.c_float), ("fixed_H", C.c_void_p),
                ("include_s0", C.c_uint32), ("lead_zeros", C.c_uint32), ("n_threads", C.c_uint32)]


# every symbol include/rub_mimo/rub_mimo.h declares (tests check the library exports them all)
ABI_SYMBOLS = [
    "rub_config_num_training_symbols", "rub_config_num_occupied", "rub_config_validate",
    "rub_rx_create", "rub_rx_destroy", "rub_rx_process_batch", "rub_rx_process_batch_host",
    "rub_rx_sc_metric_ex", "rub_rx_set_sync_mode", "rub_rx_set_debug_dir", "rub_rx_sync", "rub_rx_set_path", "rub_rx_get_path", "rub_rx_reset_counters", "rub_rx_set_host_chunk",
    "rub_rx_device_counters", "rub_rx_read_counters", "rub_rx_launch_count", "rub_rx_last_timing", "rub_rx_last_kernel",
    "rub_rx_algorithmic_bytes", "rub_rx_sc_metric", "rub_rx_timing_search", "rub_rx_set_S0",
    "rub_framegen_batch_device", "rub_rx_process_files", "rub_config_from_args", "rub_config_from_json", "rub_rx_process_capture",
    "rub_comm_get_unique_id", "rub_comm_init", "rub_allreduce_counters", "rub_rx_read_counters_global",
    "rub_comm_destroy", "rub_shard_range", "rub_msequence_init", "rub_msequence_reset",
    "rub_msequence_advance", "rub_msequence_generate_symbol", "rub_ofdmframe_init_default_sctype",
    "rub_ofdmframe_validate_sctype", "rub_ofdmframe_init_S0", "rub_ofdmframe_init_S1",
    "rub_default_S1", "rub_default_S0", "rub_default_lfsr_poly", "rub_invert_2x2",
    "rub_modem_modulate", "rub_modem_demodulate", "rub_framegen_create", "rub_framegen_destroy",
    "rub_framegen_write_sync_words", "rub_framegen_write_comb_words",
    "rub_framegen_assemble_mimo_packet", "rub_synth_row_samples", "rub_synth_frames",
    "rub_file_read_fc32", "rub_file_write_fc32", "rub_file_write_u32", "rub_strerror",
    "rub_last_error", "rub_abi_version", "rub_device_count",
]


def build(verbose=False):
    """Compile librubmimo_b200.so in-tree (nvcc, sm_100a)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], stdout=out, stderr=out)
    return LIB_PATH


_lib = None


def lib():
    """Load the C-ABI library; raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
                "rub_mimo_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.rub_strerror.restype = C.c_char_p
        L.rub_last_error.restype = C.c_char_p
        L.rub_abi_version.restype = C.c_uint32
        L.rub_config_num_training_symbols.restype = C.c_uint32
        L.rub_config_num_occupied.restype = C.c_uint32
        L.rub_rx_get_path.restype = C.c_uint32
        L.rub_rx_launch_count.restype = C.c_uint64
        L.rub_rx_device_counters.restype = C.c_void_p
        L.rub_rx_algorithmic_bytes.restype = C.c_uint64
        L.rub_rx_algorithmic_bytes.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
        L.rub_synth_row_samples.restype = C.c_uint64
        L.rub_default_lfsr_poly.restype = C.c_uint32
        L.rub_msequence_advance.restype = C.c_uint32
        L.rub_msequence_generate_symbol.restype = C.c_uint32
        L.rub_framegen_write_sync_words.restype = C.c_uint32
        L.rub_framegen_write_comb_words.restype = C.c_uint32
        L.rub_framegen_assemble_mimo_packet.restype = C.c_uint32
        L.rub_rx_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.rub_rx_destroy.argtypes = [C.c_void_p]
        L.rub_rx_process_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.rub_rx_process_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        for f in ("rub_rx_sync", "rub_rx_reset_counters", "rub_allreduce_counters", "rub_comm_destroy"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.rub_rx_set_path.argtypes = [C.c_void_p, C.c_uint32]
        L.rub_rx_set_host_chunk.argtypes = [C.c_void_p, C.c_uint32]
        L.rub_rx_get_path.argtypes = [C.c_void_p]
        L.rub_rx_launch_count.argtypes = [C.c_void_p]
        L.rub_rx_device_counters.argtypes = [C.c_void_p]
        L.rub_rx_read_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.rub_rx_read_counters_global.argtypes = [C.c_void_p, C.c_void_p]
        L.rub_rx_read_counters_global.restype = C.c_int
        L.rub_rx_last_timing.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.rub_rx_last_kernel.argtypes = [C.c_void_p]
        L.rub_rx_last_kernel.restype = C.c_char_p
        L.rub_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.rub_rx_sc_metric.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.rub_rx_sc_metric_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
        L.rub_rx_set_sync_mode.argtypes = [C.c_void_p, C.c_uint32]
        L.rub_rx_set_debug_dir.argtypes = [C.c_void_p, C.c_char_p]
        L.rub_rx_timing_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.rub_framegen_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                                C.c_uint64, C.c_float]
        L.rub_rx_set_S0.argtypes = [C.c_void_p, C.c_void_p]
        L.rub_framegen_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rub_framegen_destroy.argtypes = [C.c_void_p]
        L.rub_framegen_write_sync_words.argtypes = [C.c_void_p, C.c_void_p]
        L.rub_framegen_write_comb_words.argtypes = [C.c_void_p, C.c_void_p]
        L.rub_framegen_assemble_mimo_packet.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _check(st):
    if st != OK:
        L = lib()
        raise RubError(st, f"{L.rub_strerror(st).decode()}: {L.rub_last_error().decode()}")


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Config:
    """Runtime mirror of mimo/config.h (NUM_SUBCARRIERS, CP_LENGTH, NUM_STREAMS,
    NUM_ACCESS_CODES, PID_MAX, ...) -> rub_config."""

    def __init__(self, M=2048, cp_len=152, num_streams=2, num_access_codes=20,
                 num_data_symbols=1000, modulation=MOD_QPSK, detector=DET_ZF,
                 estimator=EST_LS_FULLBAND, pilot_spacing=8, flags=0, noise_var=0.0, sctype=None):
        self.sctype = None if sctype is None else np.ascontiguousarray(sctype, dtype=np.uint8)
        self.c = rub_config(C.sizeof(rub_config), M, cp_len, num_streams, num_access_codes,
                            num_data_symbols, modulation, detector, estimator, pilot_spacing, flags,
                            float(noise_var), _p(self.sctype))
        self.M, self.cp_len, self.N, self.nac, self.D, self.q = (M, cp_len, num_streams,
                                                                 num_access_codes, num_data_symbols,
                                                                 modulation)
        self.detector, self.estimator, self.P, self.flags = detector, estimator, pilot_spacing, flags
        self.noise_var = float(noise_var)

    def with_noise_var(self, nv):
        return Config(self.M, self.cp_len, self.N, self.nac, self.D, self.q, self.detector,
                      self.estimator, self.P, self.flags, nv, self.sctype)

    def replace(self, **kw):
        """A copy with some constructor arguments changed (the allocation is dropped when M changes)."""
        cur = dict(M=self.M, cp_len=self.cp_len, num_streams=self.N, num_access_codes=self.nac,
                   num_data_symbols=self.D, modulation=self.q, detector=self.detector, estimator=self.estimator,
                   pilot_spacing=self.P, flags=self.flags, noise_var=self.noise_var, sctype=self.sctype)
        if kw.get("M", self.M) != self.M and "sctype" not in kw:
            cur["sctype"] = None
        cur.update(kw)
        return Config(**cur)

    def validate(self):
        _check(lib().rub_config_validate(C.byref(self.c)))

    @property
    def L(self):
        return self.M + self.cp_len

    @property
    def T(self):
        return int(lib().rub_config_num_training_symbols(C.byref(self.c)))

    @property
    def Mo(self):
        return int(lib().rub_config_num_occupied(C.byref(self.c)))

    @property
    def row_bytes(self):
        return (self.Mo * self.q + 7) // 8

    @property
    def row_samples(self):
        return (self.T + self.D) * self.L


# ---- named configurations of BASELINE.json -------------------------------
def preset(name, noise_var=0.0, **over):
    name = name.upper()
    if name == "C1":   # 2x2 / 64 / QPSK loopback geometry (cp=16: CP_LENGTH 152 > M, quirk Q7)
        kw = dict(M=64, cp_len=16, num_streams=2, num_access_codes=20, num_data_symbols=1000,
                  modulation=MOD_QPSK, detector=DET_ZF, flags=FLAG_Q1_IDENTITY_INIT)
    elif name == "C2":  # 2x2 / 1024 / 16-QAM / ZF
        kw = dict(M=1024, cp_len=72, num_streams=2, num_access_codes=2, num_data_symbols=14,
                  modulation=MOD_QAM16, detector=DET_ZF)
    elif name in ("C3", "C5"):  # 4x4 / 2048 / 64-QAM / MMSE + LLR
        kw = dict(M=2048, cp_len=152, num_streams=4, num_access_codes=2, num_data_symbols=14,
                  modulation=MOD_QAM64, detector=DET_MMSE, flags=FLAG_MMSE_UNBIASED)
    elif name == "C4":  # 8x8 / 4096 / 256-QAM / MMSE / comb pilots + interpolation
        kw = dict(M=4096, cp_len=288, num_streams=8, num_access_codes=2, num_data_symbols=14,
                  modulation=MOD_QAM256, detector=DET_MMSE, estimator=EST_LS_COMB_INTERP,
                  pilot_spacing=8, flags=FLAG_MMSE_UNBIASED)
    else:
        raise ValueError(name)
    kw.update(over)
    return Config(noise_var=noise_var, **kw)


PRESET_SYNTH = {  # channel / SNR / seed of each named configuration
    "C1": dict(seed=0xC1, n_taps=0, snr_db=30.0, fixed_H=[[1, 0.5], [0.5j, 1]]),
    "C2": dict(seed=0xC2, n_taps=1, snr_db=25.0),
    "C3": dict(seed=0xC3, n_taps=8, snr_db=30.0),
    "C4": dict(seed=0xC4, n_taps=16, snr_db=38.0),
    "C5": dict(seed=0xC3, n_taps=8, snr_db=30.0),
}


# ---- host-only helpers (a8/a9 rows) ------------------------------------------------------
class MSequence:
    """liquid msequence stand-in (mimo/main.cc:1268-1270)."""

    def __init__(self, m, g, a=1):
        self.ms = rub_msequence()
        lib().rub_msequence_init(C.byref(self.ms), m, g, a)

    def reset(self):
        lib().rub_msequence_reset(C.byref(self.ms))

    def advance(self):
        return int(lib().rub_msequence_advance(C.byref(self.ms)))

    def generate_symbol(self, bps):
        return int(lib().rub_msequence_generate_symbol(C.byref(self.ms), bps))


def ofdmframe_init_default_sctype(M, use_all_carriers=True, add_null_carriers=True):
    p = np.empty(M, np.uint8)
    lib().rub_ofdmframe_init_default_sctype(_p(p), C.c_uint32(M), int(use_all_carriers),
                                            int(add_null_carriers))
    return p


def ofdmframe_validate_sctype(p):
    p = np.ascontiguousarray(p, np.uint8)
    a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
    _check(lib().rub_ofdmframe_validate_sctype(_p(p), C.c_uint32(p.size), C.byref(a), C.byref(b),
                                               C.byref(c)))
    return a.value, b.value, c.value


def ofdmframe_init_S0(p, M, ms):
    S0, s0 = np.empty(M, np.complex64), np.empty(M, np.complex64)
    p = None if p is None else np.ascontiguousarray(p, np.uint8)
    _check(lib().rub_ofdmframe_init_S0(_p(p), C.c_uint32(M), _p(S0), _p(s0), C.byref(ms.ms)))
    return S0, s0


def ofdmframe_init_S1(p, M, nac, ms):
    S1, s1 = np.empty((nac, M), np.complex64), np.empty((nac, M), np.complex64)
    p = None if p is None else np.ascontiguousarray(p, np.uint8)
    _check(lib().rub_ofdmframe_init_S1(_p(p), C.c_uint32(M), C.c_uint32(nac), _p(S1), _p(s1),
                                       C.byref(ms.ms)))
    return S1, s1


def default_S1(cfg):
    S1 = np.empty((cfg.N, cfg.nac, cfg.M), np.complex64)
    s1 = np.empty((cfg.N, cfg.nac, cfg.M), np.complex64)
    _check(lib().rub_default_S1(C.byref(cfg.c), _p(S1), _p(s1)))
    return S1, s1


def default_S0(cfg):
    S0, s0 = np.empty(cfg.M, np.complex64), np.empty(cfg.M, np.complex64)
    _check(lib().rub_default_S0(C.byref(cfg.c), _p(S0), _p(s0)))
    return S0, s0


def invert_2x2(G):
    G = np.ascontiguousarray(G, np.complex64).reshape(4)
    W = np.empty(4, np.complex64)
    g = C.c_float()
    _check(lib().rub_invert_2x2(_p(W), _p(G), C.byref(g)))
    return W.reshape(2, 2), g.value


def modem_modulate(q, sym):
    out = (C.c_float * 2)()
    _check(lib().rub_modem_modulate(C.c_uint32(q), C.c_uint32(sym), out))
    return np.complex64(complex(out[0], out[1]))


def modem_demodulate(q, x):
    x = np.complex64(x)
    v = (C.c_float * 2)(float(x.real), float(x.imag))
    s = C.c_uint32()
    _check(lib().rub_modem_demodulate(C.c_uint32(q), v, C.byref(s)))
    return s.value


def _ptrs(rows):
    arr = (C.c_void_p * len(rows))()
    for i, r in enumerate(rows):
        arr[i] = r.ctypes.data
    return arr


class FrameGen:
    """rx_beamforming::framegen (mimo/framing.h:42-103) over the C ABI."""

    def __init__(self, cfg, S0=None, s0=None, s1=None):
        self.cfg = cfg
        self.h = C.c_void_p()
        for a in (S0, s0, s1):
            assert a is None or (a.dtype == np.complex64 and a.flags.c_contiguous)
        _check(lib().rub_framegen_create(C.byref(self.h), C.byref(cfg.c), _p(S0), _p(s0), _p(s1)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().rub_framegen_destroy(self.h)
            self.h = None

    def get_num_streams(self):
        return self.cfg.N

    def write_sync_words(self):
        total = (self.cfg.nac * self.cfg.N + 1) * self.cfg.L
        tx = np.zeros((self.cfg.N, total), np.complex64)
        n = lib().rub_framegen_write_sync_words(self.h, _ptrs(list(tx)))
        assert n == total
        return tx

    def write_comb_words(self):
        tx = np.zeros((self.cfg.N, self.cfg.nac * self.cfg.L), np.complex64)
        lib().rub_framegen_write_comb_words(self.h, _ptrs(list(tx)))
        return tx

    def assemble_mimo_packet(self, in_buff):
        in_buff = np.ascontiguousarray(in_buff, np.complex64)
        tx = np.zeros((self.cfg.N, self.cfg.L), np.complex64)
        n = lib().rub_framegen_assemble_mimo_packet(self.h, _ptrs(list(tx)), _ptrs(list(in_buff)))
        assert n == self.cfg.L
        return tx


def synth_frames(cfg, n_frames, seed, n_taps=8, snr_db=30.0, baseband_gain=0.25, fixed_H=None,
                 first_frame=0, include_s0=False, lead_zeros=0, n_threads=0, S1=None, s1=None,
                 want_tx_data=True):
    """Synthetic IQ source replacing the USRP stream.  Returns (iq [F][N][row] complex64,
    tx_data [F][N][D][Mo] uint8, noise_var)."""
    H = None if fixed_H is None else np.ascontiguousarray(fixed_H, np.complex64)
    sp = rub_synth_params(seed, first_frame, n_taps, snr_db, baseband_gain, _p(H), int(include_s0),
                          lead_zeros, n_threads)
    row = int(lib().rub_synth_row_samples(C.byref(cfg.c), C.byref(sp)))
    iq = np.empty((n_frames, cfg.N, row), np.complex64)
    tx = np.empty((n_frames, cfg.N, cfg.D, cfg.Mo), np.uint8) if want_tx_data else None
    nv = C.c_float()
    _check(lib().rub_synth_frames(C.byref(cfg.c), C.byref(sp), None, _p(S1), _p(s1),
                                  C.c_uint32(n_frames), _p(iq), _p(tx), C.byref(nv)))
    return iq, tx, nv.value


def shard_range(n_frames, rank, world_size):
    b, e = C.c_uint64(), C.c_uint64()
    lib().rub_shard_range(C.c_uint64(n_frames), rank, world_size, C.byref(b), C.byref(e))
    return b.value, e.value


def device_count():
    return int(lib().rub_device_count())


# ---- the receiver ------------------------------------------------------------------------
class Receiver:
    """Batched replacement of rx_beamforming::framesync's decode work (mimo/framing.h:105-213)."""

    def __init__(self, cfg, S1=None, device=None, use_torch_stream=True):
        self.cfg = cfg
        self.h = C.c_void_p()
        stream = None
        dev = -1
        self.tstream = None
        if device_count() > 0:
            import torch
            if device is None:
                device = torch.cuda.current_device()
            dev = int(device)
            torch.cuda.set_device(dev)
            if use_torch_stream:
                # a dedicated torch stream: torch.cuda.Event can time it and torch ops can be
                # ordered against it (process_batch waits for torch's current stream first)
                self.tstream = torch.cuda.Stream(device=dev)
                stream = C.c_void_p(self.tstream.cuda_stream)
        self.device = dev
        S1 = None if S1 is None else np.ascontiguousarray(S1, np.complex64)
        _check(lib().rub_rx_create(C.byref(self.h), C.byref(cfg.c), _p(S1), dev, stream))

    def close(self):
        if getattr(self, "h", None) and lib is not None:   # `lib` is gone during interpreter shutdown
            lib().rub_rx_destroy(self.h)
            self.h = None

    __del__ = close

    def set_path(self, path):
        _check(lib().rub_rx_set_path(self.h, path))

    def set_host_chunk(self, frames):
        """frames per pipeline chunk of process_batch_host (0 = automatic)"""
        _check(lib().rub_rx_set_host_chunk(self.h, frames))

    @property
    def last_path(self):
        return int(lib().rub_rx_get_path(self.h))

    @property
    def launch_count(self):
        return int(lib().rub_rx_launch_count(self.h))

    def sync(self):
        _check(lib().rub_rx_sync(self.h))

    def reset_counters(self):
        _check(lib().rub_rx_reset_counters(self.h))

    def last_kernel(self):
        return lib().rub_rx_last_kernel(self.h).decode()

    def read_counters(self):
        out = np.zeros((self.cfg.N, 4), np.uint64)
        _check(lib().rub_rx_read_counters(self.h, _p(out)))
        return out

    def device_counters_ptr(self):
        return lib().rub_rx_device_counters(self.h)

    def last_timing(self):
        a, b = C.c_float(), C.c_float()
        _check(lib().rub_rx_last_timing(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def algorithmic_bytes(self, n_frames, out_mask, with_tx_data):
        return int(lib().rub_rx_algorithmic_bytes(self.h, n_frames, out_mask, int(with_tx_data)))

    # -- device-resident batch (torch tensors) --
    def alloc_outputs(self, n_frames, out_mask):
        import torch
        c = self.cfg
        dev = torch.device("cuda", self.device)
        o = {}
        if out_mask & OUT_EQ:
            o["eq"] = torch.empty((n_frames, c.N, c.D, c.Mo), dtype=torch.complex64, device=dev)
        if out_mask & OUT_LLR:
            o["llr"] = torch.empty((n_frames, c.N, c.D, c.Mo, c.q), dtype=torch.float32, device=dev)
        if out_mask & OUT_BITS:
            o["bits"] = torch.empty((n_frames, c.N, c.D, c.row_bytes), dtype=torch.uint8, device=dev)
        if out_mask & OUT_RXDATA:
            o["rx_data"] = torch.empty((n_frames, c.N, c.D, c.Mo), dtype=torch.uint8, device=dev)
        if out_mask & OUT_G:
            o["G"] = torch.empty((n_frames, c.N, c.N, c.M), dtype=torch.complex64, device=dev)
        return o

    def process_batch(self, iq, out=None, out_mask=OUT_EQ | OUT_LLR | OUT_BITS, tx_data=None,
                      first_sample=0, timing=None, payload_start=None):
        """iq: torch complex64 CUDA tensor [F][N][row].  Asynchronous on the handle's stream."""
        n_frames = int(iq.shape[0])
        if out is None:
            out = self.alloc_outputs(n_frames, out_mask)
        assert iq.is_cuda and iq.is_contiguous()
        io = rub_rx_io()
        io.iq = iq.data_ptr()
        io.layout = rub_iq_layout(iq.shape[1] * iq.shape[2], iq.shape[2], first_sample)
        io.tx_data = tx_data.data_ptr() if tx_data is not None else None
        io.timing = timing.data_ptr() if timing is not None else None
        io.payload_start = payload_start.data_ptr() if payload_start is not None else None
        for k in ("eq", "llr", "bits", "rx_data", "G"):
            setattr(io, k, out[k].data_ptr() if k in out else None)
        io.counters = None
        io.out_mask = out_mask
        self._keep = (iq, tx_data, timing, payload_start, out)
        if self.tstream is not None:
            import torch
            self.tstream.wait_stream(torch.cuda.current_stream())
        _check(lib().rub_rx_process_batch(self.h, C.byref(io), n_frames))
        return out

    # -- host buffers (numpy) through the pipelined H2D / compute / D2H path --
    def alloc_outputs_host(self, n_frames, out_mask, pinned=False):
        c = self.cfg
        shapes = {}
        if out_mask & OUT_EQ:
            shapes["eq"] = ((n_frames, c.N, c.D, c.Mo), np.complex64)
        if out_mask & OUT_LLR:
            shapes["llr"] = ((n_frames, c.N, c.D, c.Mo, c.q), np.float32)
        if out_mask & OUT_BITS:
            shapes["bits"] = ((n_frames, c.N, c.D, c.row_bytes), np.uint8)
        if out_mask & OUT_RXDATA:
            shapes["rx_data"] = ((n_frames, c.N, c.D, c.Mo), np.uint8)
        if out_mask & OUT_G:
            shapes["G"] = ((n_frames, c.N, c.N, c.M), np.complex64)
        o = {}
        for k, (shp, dt) in shapes.items():
            if pinned:
                import torch
                tdt = {np.complex64: torch.complex64, np.float32: torch.float32, np.uint8: torch.uint8}[dt]
                t = torch.empty(shp, dtype=tdt).pin_memory()
                o[k] = t.numpy()
                o["_pin_" + k] = t
            else:
                o[k] = np.empty(shp, dt)
        return o

    def process_batch_host(self, iq, out=None, out_mask=OUT_EQ | OUT_LLR | OUT_BITS, tx_data=None,
                           first_sample=0, timing=None, payload_start=None, counters=None):
        """iq: numpy complex64 [F][N][row] (pinned or pageable).  Synchronous."""
        assert iq.dtype == np.complex64 and iq.flags.c_contiguous
        n_frames = iq.shape[0]
        if out is None:
            out = self.alloc_outputs_host(n_frames, out_mask)
        io = rub_rx_io()
        io.iq = iq.ctypes.data
        io.layout = rub_iq_layout(iq.shape[1] * iq.shape[2], iq.shape[2], first_sample)
        io.tx_data = tx_data.ctypes.data if tx_data is not None else None
        io.timing = timing.ctypes.data if timing is not None else None
        io.payload_start = payload_start.ctypes.data if payload_start is not None else None
        for k in ("eq", "llr", "bits", "rx_data", "G"):
            setattr(io, k, out[k].ctypes.data if k in out else None)
        io.counters = counters.ctypes.data if counters is not None else None
        io.out_mask = out_mask
        _check(lib().rub_rx_process_batch_host(self.h, C.byref(io), n_frames))
        return out

    # -- multi-burst capture (f1 + f2 + decode) --
    def process_capture(self, capture, max_frames, threshold=0.95, out_mask=OUT_EQ | OUT_RXDATA, tx_data=None, out=None):
        """capture: numpy complex64 [N][n] (pinned memory makes the copies asynchronous).  out: buffers of
        alloc_outputs_host(max_frames, out_mask[, pinned=True]) to reuse across calls.  Returns (n_found,
        sync_index[n_found], outputs dict of numpy arrays sized for the bursts found)."""
        cfg = self.cfg
        capture = np.ascontiguousarray(capture, np.complex64)
        assert capture.shape[0] == cfg.N
        if out is None:
            out = self.alloc_outputs_host(max_frames, out_mask, pinned=False)
        tx_data = None if tx_data is None else np.ascontiguousarray(tx_data, np.uint8)
        cnt = np.zeros((cfg.N, 4), np.uint64)
        io = rub_rx_io()
        io.tx_data = tx_data.ctypes.data if tx_data is not None else None
        for k in ("eq", "llr", "bits", "rx_data", "G"):
            setattr(io, k, out[k].ctypes.data if k in out else None)
        io.counters = cnt.ctypes.data
        io.out_mask = out_mask
        found = C.c_uint32()
        sync = np.zeros(max_frames, np.uint64)
        _check(lib().rub_rx_process_capture(self.h, _p(capture), C.c_uint64(capture.shape[1]), C.c_float(threshold),
                                            C.c_uint32(max_frames), C.byref(io), C.byref(found), _p(sync)))
        n = found.value
        res = {k: v[:n] for k, v in out.items() if not k.startswith("_")}
        res["counters"] = cnt
        return n, sync[:n], res

    # -- offline IQ-file driver (f3) --
    def process_files(self, rx_paths, n_frames, first_sample=0, frame_stride=0, tx_data_paths=None, eq_paths=None,
                      rx_data_paths=None, llr_path=None, bits_path=None, chunk_frames=0):
        """mimo/main.cc:906-918 seam: per-antenna fc32 captures in, the reference's sinks out.
        Returns the number of complete frames processed."""
        def arr(paths):
            if paths is None:
                return None
            assert len(paths) == self.cfg.N
            return (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
        keep = [arr(rx_paths), arr(tx_data_paths), arr(eq_paths), arr(rx_data_paths)]
        job = rub_file_job(C.sizeof(rub_file_job), n_frames, first_sample, frame_stride, keep[0], keep[1], keep[2],
                           keep[3], None if llr_path is None else str(llr_path).encode(),
                           None if bits_path is None else str(bits_path).encode(), chunk_frames)
        done = C.c_uint64()
        _check(lib().rub_rx_process_files(self.h, C.byref(job), C.byref(done)))
        return done.value

    # -- transmit side (f4) --
    def framegen_batch(self, tx_data, baseband_gain=1.0):
        """Batched framegen (mimo/framing.cc:191-235) on the GPU.  tx_data: CUDA uint8 tensor
        [F][N][D][Mo]; returns a CUDA complex64 tensor [F][N][(T+D)*L] (access codes + payload)."""
        import torch
        cfg = self.cfg
        assert tx_data.is_cuda and tx_data.dtype == torch.uint8 and tx_data.is_contiguous()
        F = tx_data.shape[0]
        assert tuple(tx_data.shape) == (F, cfg.N, cfg.D, cfg.Mo)
        row = (cfg.T + cfg.D) * cfg.L
        out = torch.empty((F, cfg.N, row), dtype=torch.complex64, device=tx_data.device)
        if self.tstream is not None:
            self.tstream.wait_stream(torch.cuda.current_stream())
        _check(lib().rub_framegen_batch_device(self.h, C.c_void_p(tx_data.data_ptr()), F, C.c_void_p(out.data_ptr()),
                                               cfg.N * row, row, C.c_float(baseband_gain)))
        if self.tstream is not None:
            torch.cuda.current_stream().wait_stream(self.tstream)
        return out

    # -- synchronisation rows (f1/f2) --
    def sc_metric(self, x, mode=SYNC_FIR):
        """Schmidl & Cox metric of one stream (framing.cc:626-637); x numpy complex64.  mode SYNC_FIR is the
        reference's bit-exact FIR form, SYNC_SCAN the O(1)-per-sample sliding sums."""
        x = np.ascontiguousarray(x, np.complex64)
        y = np.empty(x.size, np.float32)
        _check(lib().rub_rx_sc_metric_ex(self.h, _p(x), x.size, _p(y), mode))
        return y

    def set_sync_mode(self, mode):
        _check(lib().rub_rx_set_sync_mode(self.h, mode))

    def set_debug_dir(self, path):
        """f_sc_%d.dat / corr_%d_%d.dat sinks of process_capture (None = off)"""
        _check(lib().rub_rx_set_debug_dir(self.h, None if path is None else str(path).encode()))

    def set_S0(self, s0):
        s0 = np.ascontiguousarray(s0, np.complex64)
        _check(lib().rub_rx_set_S0(self.h, _p(s0)))

    def timing_search(self, window, want_s0=False):
        """Access-code timing search (framing.cc:702-744); window numpy complex64 [N][Wlen]."""
        window = np.ascontiguousarray(window, np.complex64)
        corr = np.zeros((self.cfg.N, self.cfg.nac * self.cfg.N), np.int32)
        s0i = np.zeros(self.cfg.N, np.int32) if want_s0 else None
        _check(lib().rub_rx_timing_search(self.h, _p(window), window.shape[1], _p(corr), _p(s0i)))
        return (corr, s0i) if want_s0 else corr

    # -- multi-GPU counters --
    def comm_init(self, unique_id, rank, world_size):
        buf = (C.c_uint8 * NCCL_UNIQUE_ID_BYTES).from_buffer_copy(bytes(unique_id))
        _check(lib().rub_comm_init(self.h, buf, rank, world_size))

    def allreduce_counters(self):
        """Snapshot the local cumulative counters and all-reduce them on a side stream (idempotent)."""
        _check(lib().rub_allreduce_counters(self.h))

    def read_counters_global(self):
        """[N][4] uint64: the latest all-reduced counters (sum over ranks)."""
        out = np.zeros((self.cfg.N, 4), np.uint64)
        _check(lib().rub_rx_read_counters_global(self.h, _p(out)))
        return out


def comm_get_unique_id():
    buf = (C.c_uint8 * NCCL_UNIQUE_ID_BYTES)()
    _check(lib().rub_comm_get_unique_id(buf))
    return bytes(buf)
