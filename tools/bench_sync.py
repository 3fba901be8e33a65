#!/usr/bin/env python
"""Rows f1/f2: GPU Schmidl & Cox metric and access-code timing search next to the CPU restatement
of the reference's way (one FIR dot product per sample; one FFT per candidate offset).
usage: python tools/bench_sync.py   (needs a GPU; prints one JSON line per geometry)"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import rub_mimo_b200 as rub
from oracle import orc
from util import to_orc


def run(M, cp, nac, D, reps=5):
    cfg = rub.preset("C1", M=M, cp_len=cp, num_access_codes=nac, num_data_symbols=D)
    S1, s1 = rub.default_S1(cfg)
    S0, s0 = rub.default_S0(cfg)
    lead = (cfg.nac * cfg.N + 1) * cfg.L
    iq, tx, nv = rub.synth_frames(cfg, 1, 0x51, n_taps=0, snr_db=30.0, fixed_H=[[1, 0.5], [0.5j, 1]], include_s0=True,
                                  lead_zeros=lead, S1=S1, s1=s1)
    cap = iq[0]
    t0 = time.perf_counter(); r = orc.framesync_execute(to_orc(cfg), S0, S1, cap); t_all = time.perf_counter() - t0
    assert r["rc"] == 0
    t0 = time.perf_counter(); [orc.sc_metric(cfg.M, cap[s]) for s in range(cfg.N)]; t_sc = time.perf_counter() - t0
    Wlen = cfg.L * (cfg.nac * cfg.N + 4) + cfg.D * cfg.L
    w0 = int(r["window_start"])
    window = np.ascontiguousarray(cap[:, w0:w0 + Wlen])
    rx = rub.Receiver(cfg, S1); rx.set_S0(s0)
    assert np.array_equal(rx.timing_search(window), r["corr_indices"])
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); rx.timing_search(window, want_s0=True); ts.append(time.perf_counter() - t0)
    tm = []
    for _ in range(reps):
        t0 = time.perf_counter(); [rx.sc_metric(cap[s]) for s in range(cfg.N)]; tm.append(time.perf_counter() - t0)
    ffts = cfg.L * cfg.N * (1 + cfg.nac * cfg.N)
    print(json.dumps({"geometry": f"{cfg.N}x{cfg.N} M={M} cp={cp} nac={nac}", "capture_samples": int(cap.shape[1]),
                      "cpu_framesync_total_s": round(t_all, 4), "cpu_sc_metric_s": round(t_sc, 4),
                      "cpu_timing_search_s_upper": round(t_all - t_sc, 4), "reference_ffts": ffts,
                      "gpu_timing_search_ms_host_to_host": round(1e3 * min(ts), 3),
                      "gpu_sc_metric_ms_host_to_host": round(1e3 * min(tm), 3)}))


if __name__ == "__main__":
    run(64, 16, 20, 100)
    run(2048, 152, 20, 14)


def loop_vs_reference(name="ref_m2048_default"):
    """The whole receive loop of mimo/main.cc on one capture (framesync::execute in 4096-sample
    chunks: Schmidl & Cox, access-code buffering, timing search, LS estimate, invert, decode of
    ~1000 OFDM symbols): the SAME caller source (tests/framing_driver.cc) linked against the
    reference's own framing.cc (oracle/_ref, stand-in FFT/VOLK/liquid) and against the facade +
    librubmimo_b200.so."""
    import ctypes as C, os, subprocess, tempfile
    import test_ref_fixtures as t
    root = os.getcwd()
    z, cfg, S0, s0, S1, cap, tx = t._load(name)
    cap = [np.ascontiguousarray(r) for r in cap]
    libs = {}
    ref_so = os.path.join(root, "oracle", "_ref", "libref_framing.so")
    if os.path.exists(ref_so):
        libs["reference framing.cc (1 host thread)"] = C.CDLL(ref_so)
    out = os.path.join(tempfile.mkdtemp(), "libfacade_driver.so")
    libdir = os.path.join(root, "rub_mimo_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-w", "-I", os.path.join(root, "include", "rub_mimo"),
                           "-I", os.path.join(root, "include"), os.path.join(root, "tests", "framing_driver.cc"), "-o", out,
                           "-L", libdir, "-lrubmimo_b200", f"-Wl,-rpath,{libdir}"])
    libs["facade + CUDA library"] = C.CDLL(out)
    res = {"capture": name, "samples_per_stream": int(cap[0].size)}
    for label, lib in libs.items():
        best = None
        for _ in range(2 if "reference" in label else 4):
            r = t._SyncResult()
            G = np.zeros((cfg.M, 2, 2), np.complex64)
            eq = np.zeros((2, cfg.D + 8, cfg.Mo), np.complex64)
            p = np.ascontiguousarray(z["sctype"])
            t0 = time.perf_counter()
            lib.ref_framesync(C.c_uint(cfg.M), C.c_uint(cfg.cp_len), C.c_uint(cfg.nac), t._vp(p), t._vp(cap[0]), t._vp(cap[1]),
                              C.c_uint64(cap[0].size), C.c_uint(4096), C.byref(r), t._vp(G), t._vp(eq), C.c_uint(cfg.D + 8),
                              C.c_uint(cfg.Mo))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            assert r.state == 3 and r.symbols == int(z["symbols"])
        res[label + " [s]"] = round(best, 4)
    print(json.dumps(res))


def capture_throughput(K=64):
    """rub_rx_process_capture on a host capture holding K bursts (2x2 / 1024 / 16-QAM, 14 payload symbols
    each), next to the oracle's receive loop run burst by burst on one host thread."""
    import test_gpu_capture as tc
    cfg = rub.preset("C1", M=1024, cp_len=72, num_access_codes=2, num_data_symbols=14, modulation=rub.MOD_QAM16)
    S0, S1, cap, tx, slices = tc._bursts(cfg, K, seed=0xB5)
    rx = rub.Receiver(cfg, S1)
    rx.process_capture(cap, max_frames=K + 2)
    best = None
    for _ in range(3):
        t0 = time.perf_counter(); n, sync, out = rx.process_capture(cap, max_frames=K + 2, out_mask=rub.OUT_EQ | rub.OUT_RXDATA, tx_data=tx)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert n == K and np.array_equal(out["rx_data"], tx)
    t0 = time.perf_counter()
    for a, b in slices[:4]:
        assert orc.framesync_execute(to_orc(cfg), S0, S1, cap[:, a:b])["rc"] == 0
    t_cpu = (time.perf_counter() - t0) / 4 * K
    print(json.dumps({"capture": f"{K} bursts 2x2 M=1024 16-QAM D=14", "samples_per_stream": int(cap.shape[1]),
                      "gpu_process_capture_s": round(best, 4),
                      "gpu_msamples_per_s": round(cap.size / best / 1e6, 1),
                      "cpu_oracle_loop_s_1_thread": round(t_cpu, 2)}))


if __name__ == "__main__":
    loop_vs_reference()
    capture_throughput()
