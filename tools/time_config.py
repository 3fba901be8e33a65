#!/usr/bin/env python
"""Times rub_rx_process_batch (device buffers) for an arbitrary configuration.
usage: python tools/time_config.py M cp N nac D q frames [zf|mmse]   (RUB_MIMO_LIB selects the build)"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import rub_mimo_b200 as rub

M, cp, N, nac, D, q, F = [int(x) for x in sys.argv[1:8]]
det = rub.DET_MMSE if len(sys.argv) > 8 and sys.argv[8] == "mmse" else rub.DET_ZF
cfg = rub.Config(M=M, cp_len=cp, num_streams=N, num_access_codes=nac, num_data_symbols=D, modulation=q, detector=det)
S1, s1 = rub.default_S1(cfg)
U = 16
iq, tx, nv = rub.synth_frames(cfg, U, 1, n_taps=4, snr_db=28.0, S1=S1, s1=s1)
cfg = cfg.with_noise_var(nv)
rx = rub.Receiver(cfg, S1)
d_iq = torch.from_numpy(iq).repeat((F + U - 1) // U, 1, 1)[:F].contiguous().cuda()
d_tx = torch.from_numpy(tx).repeat((F + U - 1) // U, 1, 1, 1)[:F].contiguous().cuda()
out = rx.alloc_outputs(F, rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS)
ts = []
for i in range(8):
    rx.process_batch(d_iq, out=out, out_mask=rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS, tx_data=d_tx)
    rx.sync()
    ts.append(rx.last_timing()[0])
path = {1: "staged", 2: "fused"}[rx.last_path]
b = rx.algorithmic_bytes(F, rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS, True) if hasattr(rx, "algorithmic_bytes") else 0
print(f"{M}/{cp} {N}x{N} nac{nac} D{D} q{q} F{F} {path}: {min(ts[2:]):.4f} ms  ({b / min(ts[2:]) / 1e6:.0f} GB/s)")
rx.close()
