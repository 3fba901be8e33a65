#!/usr/bin/env python
"""Regenerates the committed golden fixtures under tests/golden/ (run from the repo root).

The reference has no golden vectors and cannot be built or imported here, so the fixtures are
produced by the CPU oracle and cross-checked against the independent float64 model before being
written.  Inputs come from the product's synthetic source (seeded, counter-based RNG); the IQ
samples themselves are stored so the fixtures do not depend on libm.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rub_mimo_b200 as rub  # noqa: E402
from oracle import oracle_f64 as f64  # noqa: E402
from util import make_case, oracle_run  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

VECTOR_CASES = {
    "g_c1_2x2_m64_qpsk_zf_q1": (dict(M=64, cp_len=16, num_streams=2, num_access_codes=4, num_data_symbols=12,
                                     modulation=2, detector=0, flags=rub.FLAG_Q1_IDENTITY_INIT), 3,
                                dict(n_taps=0, snr_db=12.0, fixed_H=[[1, 0.5], [0.5j, 1]])),
    "g_2x2_m512_16qam_zf": (dict(M=512, cp_len=36, num_streams=2, num_access_codes=2, num_data_symbols=4,
                                 modulation=4, detector=0), 2, dict(n_taps=2, snr_db=22.0)),
    "g_4x4_m512_64qam_mmse": (dict(M=512, cp_len=40, num_streams=4, num_access_codes=2, num_data_symbols=3,
                                   modulation=6, detector=1, flags=rub.FLAG_MMSE_UNBIASED), 2,
                              dict(n_taps=4, snr_db=30.0)),
    "g_8x8_m256_256qam_comb": (dict(M=256, cp_len=24, num_streams=8, num_access_codes=2, num_data_symbols=2,
                                    modulation=8, detector=1, estimator=1, flags=rub.FLAG_MMSE_UNBIASED), 1,
                               dict(n_taps=2, snr_db=38.0)),
}
KAT8 = [
    ("awgn_2x2_qpsk", dict(M=64, cp_len=16, num_streams=2, num_access_codes=20, num_data_symbols=40, modulation=2,
                           detector=0, flags=rub.FLAG_Q1_IDENTITY_INIT), 4, 0xC1,
     dict(n_taps=0, snr_db=9.0, fixed_H=[[1, 0.5], [0.5j, 1]])),
    ("rayleigh_2x2_16qam", dict(M=1024, cp_len=72, num_streams=2, num_access_codes=2, num_data_symbols=14,
                                modulation=4, detector=0), 6, 0xC2, dict(n_taps=1, snr_db=25.0)),
    ("rayleigh_4x4_64qam_mmse", dict(M=2048, cp_len=152, num_streams=4, num_access_codes=2, num_data_symbols=14,
                                     modulation=6, detector=1, flags=rub.FLAG_MMSE_UNBIASED), 3, 0xC3,
     dict(n_taps=8, snr_db=30.0)),
]


def main():
    os.makedirs(GOLD, exist_ok=True)
    for name, (kw, nf, syn) in VECTOR_CASES.items():
        cfg, S1, iq, tx = make_case(rub.Config(**kw), nf, seed=sum(map(ord, name)), **syn)
        ref = oracle_run(cfg, S1, iq, tx)
        m = f64.rx_frame(iq[0], S1, cfg.M, cfg.cp_len, cfg.N, cfg.nac, cfg.D, cfg.q, cfg.detector,
                         cfg.estimator, cfg.P, cfg.flags, cfg.noise_var, cfg.sctype)
        assert np.abs(ref["eq"][0] - m["eq"]).max() < 1e-3, name
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), config=json.dumps(kw), noise_var=np.float32(cfg.noise_var),
                            S1=S1, iq=iq, tx_data=tx, **{"out_" + k: v for k, v in ref.items()})
        print(name, "BER", ref["counters"][:, 0].sum() / ref["counters"][:, 1].sum())
    cases = []
    for name, kw, nf, seed, syn in KAT8:
        cfg, S1, iq, tx = make_case(rub.Config(**kw), nf, seed=seed, **syn)
        ref = oracle_run(cfg, S1, iq, tx)
        syn_j = dict(syn)
        if "fixed_H" in syn_j:  # complex is not JSON: store [re, im] pairs
            H = np.asarray(syn_j.pop("fixed_H"), np.complex64)
            syn_j["fixed_H_ri"] = np.stack([H.real, H.imag], -1).tolist()
        cases.append(dict(name=name, config=kw, frames=nf, seed=seed, synth=syn_j, noise_var=cfg.noise_var,
                          counters=ref["counters"].tolist()))
        print(name, ref["counters"].tolist())
    json.dump(dict(note="regenerate with oracle/make_golden.py", cases=cases),
              open(os.path.join(GOLD, "kat8_counters.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
