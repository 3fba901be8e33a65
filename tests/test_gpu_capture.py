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


@pytest.mark.parametrize("mode", [rub.SYNC_FIR, rub.SYNC_SCAN])
def test_both_metric_forms_find_and_decode_the_same_bursts(mode):
    """process_capture with the bit-exact FIR metric and with the sliding-sum metric (row f2): same bursts, same
    equalised symbols bit for bit (the timing search pins the FFT windows, not the plateau start), plateau starts
    within a sample of the oracle's."""
    cfg = rub.preset("C1", M=256, cp_len=20, num_access_codes=4, num_data_symbols=25)
    K = 4
    S0, S1, cap, tx, slices = _bursts(cfg, K, seed=0xF2, gaps=[11, 0, 250])
    rx = rub.Receiver(cfg, S1)
    rx.set_sync_mode(mode)
    n, sync, out = rx.process_capture(cap, max_frames=K + 2, out_mask=rub.OUT_EQ | rub.OUT_RXDATA, tx_data=tx)
    assert n == K and np.array_equal(out["rx_data"], tx)
    for k, (a, b) in enumerate(slices):
        r = orc.framesync_execute(to_orc(cfg), S0, S1, cap[:, a:b])
        assert abs(int(sync[k]) - (a + int(r["sync_index"]))) <= (1 if mode == rub.SYNC_FIR else 2)
        assert np.array_equal(out["eq"][k], r["eq"])


def test_burst_at_the_very_start_of_a_capture_is_decoded():
    """Round-1 ADVICE: a plateau found before a whole window of samples had been seen was dropped.  The
    reference's window buffer is zero-filled in front of the first sample; so is the device copy now."""
    cfg = rub.preset("C1", M=64, cp_len=16, num_access_codes=20, num_data_symbols=40)
    S0, S1, cap, tx, slices = _bursts(cfg, 1, seed=77)
    lead = (cfg.nac * cfg.N + 1) * cfg.L
    early = np.ascontiguousarray(cap[:, lead - 40:])          # S0 starts 40 samples into the capture
    rx = rub.Receiver(cfg, S1)
    rx.set_sync_mode(rub.SYNC_FIR)
    n, sync, out = rx.process_capture(early, max_frames=2, out_mask=rub.OUT_EQ | rub.OUT_RXDATA, tx_data=tx)
    r = orc.framesync_execute(to_orc(cfg), S0, S1, early)
    assert r["rc"] == 0 and n == 1
    assert int(sync[0]) == int(r["sync_index"])
    assert np.array_equal(out["eq"][0], r["eq"])
    assert np.array_equal(out["rx_data"][0], tx[0])


def test_debug_sinks_have_the_reference_formats(tmp_path):
    """Row f3: f_sc_%d.dat and corr_%d_%d.dat (mimo/framing.cc:598-600, :676-680, :873-883), float32, readable the way
    mimo/apps/plot.py:27-40 reads them, holding the oracle's values."""
    cfg = rub.preset("C1", M=64, cp_len=16, num_access_codes=4, num_data_symbols=12)
    S0, S1, cap, tx, slices = _bursts(cfg, 2, seed=0xF3)
    s0 = rub.default_S0(cfg)[1]
    rx = rub.Receiver(cfg, S1)
    rx.set_S0(s0)
    rx.set_sync_mode(rub.SYNC_FIR)
    rx.set_debug_dir(tmp_path)
    n, sync, out = rx.process_capture(cap, max_frames=4)
    assert n == 2
    L, M, N = cfg.L, cfg.M, cfg.N
    max_ac = cfg.nac * N
    acb_len = L * (max_ac + 4)
    for s in range(N):
        y = np.fromfile(tmp_path / f"f_sc_{s + 1}.dat", np.float32)
        ref = orc.sc_metric(M, cap[s])
        assert y.size == cap.shape[1] and np.array_equal(y.view(np.uint32), ref.view(np.uint32))
    # the last burst's window as the reference holds it when estimate_channel runs
    Wlen = acb_len + cfg.D * L
    i_switch = int(sync[1]) + cfg.D * L + acb_len - L
    win = cap[:, i_switch - Wlen:i_switch]
    for r in range(N):
        for a in range(0, max_ac + 1):
            v = np.fromfile(tmp_path / f"corr_{r + 1}_{a}.dat", np.float32)
            assert v.size == acb_len - M
            tpl = S0 if a == 0 else S1[(a - 1) % N, (a - 1) // N]
            base = 0 if a == 0 else L * a
            nz = np.nonzero(v)[0]
            assert nz.min() >= base and nz.max() < base + L
            for i in (0, 1, L // 2, L - 1):                       # the reference's way: one FFT per candidate offset
                X = orc.fft_forward(win[r, base + i: base + i + M])
                xyz = np.sum(X.astype(np.complex128) * np.conj(tpl.astype(np.complex128)))
                want = abs(xyz) ** 2 / float(M * M)
                assert abs(v[base + i] - want) <= 1e-4 * max(want, np.abs(v).max() * 1e-3)
            assert int(np.argmax(v)) == base + int(np.argmax(v[base:base + L]))
