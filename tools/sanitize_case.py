"""Small fused + staged + host-path invocations for compute-sanitizer (one tool per gpurun call)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rub_mimo_b200 as rub
from util import make_case, gpu_run, oracle_run, assert_parity
for kw, nf, syn, path in [
    # warp-specialised kernel (4x4 / 2048): 64-QAM MMSE and 16-QAM ZF, more frames than one CTA walks in a round
    (dict(M=2048, cp_len=152, num_streams=4, num_access_codes=2, num_data_symbols=3, modulation=6, detector=1, flags=2), 3, dict(n_taps=3, snr_db=28.0), rub.PATH_FUSED),
    (dict(M=2048, cp_len=152, num_streams=4, num_access_codes=2, num_data_symbols=4, modulation=4, detector=0), 2, dict(n_taps=2, snr_db=24.0), rub.PATH_FUSED),
    # staged path with TMA task records (N >= 4, every carrier occupied) and the fused comb LS + weights kernel
    (dict(M=256, cp_len=18, num_streams=4, num_access_codes=2, num_data_symbols=3, modulation=6, detector=1), 3, dict(n_taps=2, snr_db=28.0), rub.PATH_STAGED),
    (dict(M=256, cp_len=18, num_streams=8, num_access_codes=2, num_data_symbols=2, modulation=8, detector=1, estimator=1), 2, dict(n_taps=2, snr_db=30.0), rub.PATH_STAGED),
    (dict(M=512, cp_len=40, num_streams=4, num_access_codes=2, num_data_symbols=2, modulation=6, detector=1, flags=2), 3, dict(n_taps=3, snr_db=28.0), rub.PATH_FUSED),
    (dict(M=512, cp_len=40, num_streams=2, num_access_codes=2, num_data_symbols=3, modulation=4, detector=0), 5, dict(n_taps=2, snr_db=20.0), rub.PATH_FUSED),
    (dict(M=256, cp_len=18, num_streams=3, num_access_codes=2, num_data_symbols=2, modulation=8, detector=1), 2, dict(n_taps=2, snr_db=30.0), rub.PATH_STAGED),
    (dict(M=128, cp_len=9, num_streams=8, num_access_codes=2, num_data_symbols=2, modulation=2, detector=1, estimator=1), 2, dict(n_taps=2, snr_db=15.0), rub.PATH_STAGED),
]:
    cfg, S1, iq, tx = make_case(rub.Config(**kw), nf, seed=9, **syn)
    ref = oracle_run(cfg, S1, iq, tx)
    got = gpu_run(cfg, S1, iq, tx, path=path)
    assert_parity(ref, got, cfg.q)
    rx = rub.Receiver(cfg, S1)
    cnt = np.zeros((cfg.N, 4), np.uint64)
    out = rx.process_batch_host(iq, out_mask=rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS | rub.OUT_RXDATA | rub.OUT_G, tx_data=tx, counters=cnt)
    out["counters"] = cnt
    assert_parity(ref, out, cfg.q)
    rx.close()
print("sanitize cases ok")
