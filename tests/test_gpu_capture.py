"""Rows f1 + f2 + decode together: rub_rx_process_capture walks a capture holding several bursts
(leading zeros + S0 + access codes + payload, like the reference's transmissions) without any
pre-alignment, and returns what the oracle's faithful receive loop returns burst by burst."""
import numpy as np
import pytest

import rub_mimo_b200 as rub
from oracle import orc
from util import to_orc

pytestmark = pytest.mark.gpu

H2 = [[1, 0.5], [0.5j, 1]]


def _bursts(cfg, K, seed, snr_db=30.0, n_taps=0, H=H2, gaps=None):
    S1, s1 = rub.default_S1(cfg)
    S0, s0 = rub.default_S0(cfg)
    lead = (cfg.nac * cfg.N + 1) * cfg.L
    parts, txs, slices = [], [], []
    pos = 0
    for k in range(K):
        iq, tx, nv = rub.synth_frames(cfg, 1, seed, first_frame=k, n_taps=n_taps, snr_db=snr_db,
                                      fixed_H=None if n_taps else H, include_s0=True, lead_zeros=lead, S1=S1, s1=s1)
        cap = iq[0]
        if gaps:                                   # extra noise-only samples between bursts
            g = gaps[k % len(gaps)]
            cap = np.concatenate([cap, cap[:, :g] * 0 + cap[:, :g][:, ::-1] * (np.abs(cap[:, :g]) < 0.05)], axis=1)
        parts.append(cap)
        txs.append(tx[0])
        slices.append((pos, pos + cap.shape[1]))
        pos += cap.shape[1]
    return S0, S1, np.ascontiguousarray(np.concatenate(parts, axis=1)), np.stack(txs), slices


@pytest.mark.parametrize("geom", [dict(M=64, cp_len=16, num_access_codes=20, num_data_symbols=40),
                                  dict(M=256, cp_len=20, num_access_codes=4, num_data_symbols=25),
                                  dict(M=1024, cp_len=72, num_access_codes=2, num_data_symbols=14)])
def test_every_burst_of_a_capture_is_found_and_decoded_like_the_oracle(geom):
    cfg = rub.preset("C1", **geom)
    K = 5
    S0, S1, cap, tx, slices = _bursts(cfg, K, seed=0xCA + geom["M"], gaps=[0, 37, 500])
    rx = rub.Receiver(cfg, S1)
    n, sync, out = rx.process_capture(cap, max_frames=K + 3, out_mask=rub.OUT_EQ | rub.OUT_RXDATA | rub.OUT_G, tx_data=tx)
    assert n == K
    assert np.array_equal(out["rx_data"], tx)                       # SER 0 at 30 dB
    c = out["counters"]
    assert c[:, 2].sum() == 0 and c[:, 3].sum() == K * cfg.N * cfg.D * cfg.Mo
    for k, (a, b) in enumerate(slices):
        r = orc.framesync_execute(to_orc(cfg), S0, S1, cap[:, a:b])
        assert r["rc"] == 0
        assert abs(int(sync[k]) - (a + int(r["sync_index"]))) <= 1   # the metric history differs at a slice start
        assert np.array_equal(out["eq"][k], r["eq"])                  # bit-exact equalised symbols
        assert np.array_equal(out["G"][k].transpose(2, 0, 1), r["G"])


def test_capture_without_bursts_and_truncated_burst():
    cfg = rub.preset("C1", M=64, cp_len=16, num_access_codes=20, num_data_symbols=40)
    S0, S1, cap, tx, slices = _bursts(cfg, 2, seed=5)
    rx = rub.Receiver(cfg, S1)
    rng = np.random.default_rng(1)
    noise = (0.01 * (rng.standard_normal((2, 30000)) + 1j * rng.standard_normal((2, 30000)))).astype(np.complex64)
    assert rx.process_capture(noise, max_frames=4)[0] == 0
    # second burst cut in the middle of its payload: only the first one is returned
    cut = slices[1][0] + (slices[1][1] - slices[1][0]) * 2 // 3
    n, sync, out = rx.process_capture(cap[:, :cut], max_frames=4)
    assert n == 1 and np.array_equal(out["rx_data"][0], tx[0])
    # max_frames bounds the search
    assert rx.process_capture(cap, max_frames=1)[0] == 1
