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
