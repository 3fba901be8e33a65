"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs.  Bar: hard bits / symbol indices / error counters bit-exact; equalised
symbols, channel estimates and LLRs bit-exact against the mirror-fp32 oracle (and within 1e-4
of the independent float64 model, see test_gpu_f64.py)."""
import numpy as np
import pytest

import rub_mimo_b200 as rub
from util import assert_parity, gpu_run, make_case, oracle_run

pytestmark = pytest.mark.gpu

CASES = {
    # name: (config kwargs, frames, synth kwargs)
    "c1_aligned": (dict(M=64, cp_len=16, num_streams=2, num_access_codes=20, num_data_symbols=50,
                        modulation=rub.MOD_QPSK, detector=rub.DET_ZF, flags=rub.FLAG_Q1_IDENTITY_INIT),
                   3, dict(n_taps=0, snr_db=30.0, fixed_H=[[1, 0.5], [0.5j, 1]])),
    "c2_small": (dict(M=1024, cp_len=72, num_streams=2, num_access_codes=2, num_data_symbols=14,
                      modulation=rub.MOD_QAM16, detector=rub.DET_ZF), 12, dict(n_taps=1, snr_db=25.0)),
    "c3_small": (dict(M=2048, cp_len=152, num_streams=4, num_access_codes=2, num_data_symbols=14,
                      modulation=rub.MOD_QAM64, detector=rub.DET_MMSE, flags=rub.FLAG_MMSE_UNBIASED),
                 5, dict(n_taps=8, snr_db=30.0)),
    "c3_biased_zf": (dict(M=2048, cp_len=152, num_streams=4, num_access_codes=3, num_data_symbols=5,
                          modulation=rub.MOD_QAM64, detector=rub.DET_ZF), 2, dict(n_taps=8, snr_db=28.0)),
    "c4_small": (dict(M=4096, cp_len=288, num_streams=8, num_access_codes=2, num_data_symbols=4,
                      modulation=rub.MOD_QAM256, detector=rub.DET_MMSE, estimator=rub.EST_LS_COMB_INTERP,
                      flags=rub.FLAG_MMSE_UNBIASED), 2, dict(n_taps=16, snr_db=38.0)),
    "siso": (dict(M=512, cp_len=36, num_streams=1, num_access_codes=4, num_data_symbols=6,
                  modulation=rub.MOD_QAM16, detector=rub.DET_ZF), 4, dict(n_taps=3, snr_db=22.0)),
    "n3_mmse_biased": (dict(M=256, cp_len=18, num_streams=3, num_access_codes=2, num_data_symbols=7,
                            modulation=rub.MOD_QAM64, detector=rub.DET_MMSE), 4, dict(n_taps=2, snr_db=20.0)),
    "n2_mmse_qpsk_m128": (dict(M=128, cp_len=10, num_streams=2, num_access_codes=2, num_data_symbols=9,
                               modulation=rub.MOD_QPSK, detector=rub.DET_MMSE), 6, dict(n_taps=2, snr_db=8.0)),
    "n4_m512_256qam": (dict(M=512, cp_len=40, num_streams=4, num_access_codes=2, num_data_symbols=6,
                            modulation=rub.MOD_QAM256, detector=rub.DET_MMSE, flags=rub.FLAG_MMSE_UNBIASED),
                       7, dict(n_taps=4, snr_db=35.0)),
    "n2_m4096": (dict(M=4096, cp_len=288, num_streams=2, num_access_codes=2, num_data_symbols=3,
                      modulation=rub.MOD_QAM16, detector=rub.DET_ZF, flags=rub.FLAG_ZF_CHOLESKY),
                 3, dict(n_taps=6, snr_db=24.0)),
}


def _run(name, path):
    kw, nf, sk = CASES[name]
    cfg, S1, iq, tx = make_case(rub.Config(**kw), nf, seed=hash(name) & 0xFFFF, **sk)
    ref = oracle_run(cfg, S1, iq, tx)
    got = gpu_run(cfg, S1, iq, tx, path=path)
    assert_parity(ref, got, cfg.q)
    return got


@pytest.mark.parametrize("name", sorted(CASES))
def test_staged_path_matches_oracle(name):
    got = _run(name, rub.PATH_STAGED)
    assert got["path"] == rub.PATH_STAGED


@pytest.mark.parametrize("name", ["c2_small", "c3_small", "c3_biased_zf", "n4_m512_256qam", "n2_m4096"])
def test_fused_path_matches_oracle(name):
    got = _run(name, rub.PATH_FUSED)
    assert got["path"] == rub.PATH_FUSED


def test_ragged_allocation_uses_staged_path():
    p = rub.ofdmframe_init_default_sctype(512, use_all_carriers=False, add_null_carriers=True)
    cfg = rub.Config(M=512, cp_len=36, num_streams=2, num_access_codes=2, num_data_symbols=5,
                     modulation=rub.MOD_QAM64, detector=rub.DET_MMSE, sctype=p)
    cfg, S1, iq, tx = make_case(cfg, 3, seed=77, n_taps=3, snr_db=27.0)
    ref = oracle_run(cfg, S1, iq, tx)
    got = gpu_run(cfg, S1, iq, tx)
    assert got["path"] == rub.PATH_STAGED
    assert_parity(ref, got, cfg.q)
