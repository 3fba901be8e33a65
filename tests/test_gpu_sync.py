"""Rows f1/f2 of SURVEY.md 8 on the GPU: the Schmidl & Cox metric (mimo/framing.cc:626-637) is
bit-exact with the oracle; the access-code timing search (mimo/framing.cc:702-744) returns the
oracle's per-link indices (the oracle evaluates it the reference's way, one FFT per candidate
offset; the kernel correlates in the time domain)."""
import numpy as np
import pytest

import rub_mimo_b200 as rub
from oracle import orc
from util import to_orc

pytestmark = pytest.mark.gpu

H2 = [[1, 0.5], [0.5j, 1]]


def _capture(name, D, seed, snr_db=30.0, fixed_H=H2, n_taps=0, **over):
    cfg = rub.preset(name, num_data_symbols=D, **over)
    S1, s1 = rub.default_S1(cfg)
    S0, s0 = rub.default_S0(cfg)
    lead = (cfg.nac * cfg.N + 1) * cfg.L
    iq, tx, nv = rub.synth_frames(cfg, 1, seed, n_taps=n_taps, snr_db=snr_db, fixed_H=fixed_H, include_s0=True,
                                  lead_zeros=lead, S1=S1, s1=s1)
    return cfg, S0, s0, S1, iq[0], tx[0], lead


@pytest.mark.parametrize("name,D,over", [("C1", 40, {}), ("C1", 12, dict(M=64, cp_len=16, num_access_codes=4)),
                                         ("C1", 8, dict(M=256, cp_len=20, num_access_codes=3))])
def test_sc_metric_is_bit_exact(name, D, over):
    cfg, S0, s0, S1, cap, tx, lead = _capture(name, D, 0x5C + D, **over)
    rx = rub.Receiver(cfg, S1)
    for s in range(cfg.N):
        x = cap[s][: min(cap.shape[1], lead + 6 * cfg.L)]
        y = rx.sc_metric(x)
        ref = orc.sc_metric(cfg.M, x)
        # 0/0 before the first non-zero sample is NaN in both
        assert np.array_equal(y.view(np.uint32), ref.view(np.uint32))
        assert np.nanmax(y) > 0.95
    assert rx.launch_count == cfg.N


def test_scan_metric_tracks_the_fir_metric():
    """Row f2: the sliding-sum form (O(1) per sample) agrees with the bit-exact FIR form to fp32 rounding and
    crosses the plateau threshold within a sample of it."""
    cfg, S0, s0, S1, cap, tx, lead = _capture("C1", 6, 0x5D, M=1024, cp_len=72, num_access_codes=2)
    rx = rub.Receiver(cfg, S1)
    for s in range(cfg.N):
        x = cap[s]
        yf, ys = rx.sc_metric(x, rub.SYNC_FIR), rx.sc_metric(x, rub.SYNC_SCAN)
        ok = np.isfinite(yf) & np.isfinite(ys) & (np.arange(x.size) > lead)
        assert ok.sum() > x.size // 2
        assert np.abs(ys[ok] - yf[ok]).max() <= 2e-4 * max(1.0, np.abs(yf[ok]).max())
        cf_, cs_ = np.nonzero(yf > 0.95)[0], np.nonzero(ys > 0.95)[0]
        assert cf_.size and abs(int(cf_[0]) - int(cs_[0])) <= 1
    # tile boundaries and ragged lengths: every length runs and the defined values agree
    rng = np.random.default_rng(3)
    for n in (1, 255, 2047, 2048, 2049, 5000):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        yf, ys = rx.sc_metric(x, rub.SYNC_FIR), rx.sc_metric(x, rub.SYNC_SCAN)
        assert np.allclose(ys, yf, rtol=2e-4, atol=1e-6), n


def test_sc_metric_ragged_lengths():
    cfg = rub.preset("C1", num_data_symbols=1, M=64, cp_len=16, num_access_codes=2)
    S1, _ = rub.default_S1(cfg)
    rx = rub.Receiver(cfg, S1)
    rng = np.random.default_rng(7)
    for n in (1, 31, 32, 95, 96, 97, 255, 256, 257, 1000):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        assert np.array_equal(rx.sc_metric(x).view(np.uint32), orc.sc_metric(cfg.M, x).view(np.uint32)), n
    assert rx.sc_metric(np.zeros(0, np.complex64)).size == 0


@pytest.mark.parametrize("name,D,over,snr,thr", [
    ("C1", 30, {}, 30.0, 0.95), ("C1", 30, {}, 8.0, 0.3),
    ("C1", 12, dict(M=64, cp_len=16, num_access_codes=4), 20.0, 0.8),
    ("C1", 8, dict(M=256, cp_len=20, num_access_codes=3), 15.0, 0.6),
    ("C1", 4, dict(M=2048, cp_len=152, num_access_codes=2), 20.0, 0.9)])
def test_timing_search_matches_oracle(name, D, over, snr, thr):
    cfg, S0, s0, S1, cap, tx, lead = _capture(name, D, 0x71 + D, snr_db=snr, **over)
    r = orc.framesync_execute(to_orc(cfg), S0, S1, cap, threshold=thr)
    assert r["rc"] == 0
    Wlen = cfg.L * (cfg.nac * cfg.N + 4) + cfg.D * cfg.L
    w0 = int(r["window_start"])
    window = np.ascontiguousarray(cap[:, w0:w0 + Wlen])
    assert window.shape[1] == Wlen
    rx = rub.Receiver(cfg, S1)
    rx.set_S0(s0)
    corr, s0i = rx.timing_search(window, want_s0=True)
    assert np.array_equal(corr, r["corr_indices"])
    assert list(s0i) == list(r["s0_corr_index"])
    assert np.array_equal(rx.timing_search(window), r["corr_indices"])


def test_timing_search_multipath_per_link_offsets():
    """Different links peak at different offsets (quirk Q2: per-link timing)."""
    cfg, S0, s0, S1, cap, tx, lead = _capture("C1", 10, 0xAB, snr_db=25.0, fixed_H=None, n_taps=4,
                                              M=128, cp_len=16, num_access_codes=3)
    r = orc.framesync_execute(to_orc(cfg), S0, S1, cap)
    if r["rc"] != 0:
        pytest.skip("random channel did not produce a plateau on both streams")
    Wlen = cfg.L * (cfg.nac * cfg.N + 4) + cfg.D * cfg.L
    w0 = int(r["window_start"])
    window = np.ascontiguousarray(cap[:, w0:w0 + Wlen])
    if window.shape[1] != Wlen:
        pytest.skip("window clipped")
    rx = rub.Receiver(cfg, S1)
    assert np.array_equal(rx.timing_search(window), r["corr_indices"])


def test_timing_search_rejects_short_window_and_missing_S0():
    cfg = rub.preset("C1", num_data_symbols=1, M=64, cp_len=16, num_access_codes=2)
    S1, _ = rub.default_S1(cfg)
    rx = rub.Receiver(cfg, S1)
    with pytest.raises(rub.RubError):
        rx.timing_search(np.zeros((cfg.N, cfg.L * 2), np.complex64))
    with pytest.raises(rub.RubError):
        rx.timing_search(np.zeros((cfg.N, cfg.L * 8), np.complex64), want_s0=True)
    # an all-zero window keeps the reference's initial index 0 everywhere
    assert not rx.timing_search(np.zeros((cfg.N, cfg.L * 8), np.complex64)).any()
