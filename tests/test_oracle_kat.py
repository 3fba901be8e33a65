"""Known-answer tests that pin the CPU oracle.

The reference ships no tests, fixtures or golden vectors, so these KATs are
derived from its source and from an independent float64 numpy model (oracle/oracle_f64.py)."""
import numpy as np
import pytest

import rub_mimo_b200 as rub
from oracle import oracle_f64 as f64
from oracle import orc
from util import make_case, oracle_run, to_orc


# ---- KAT-1: invert() on fixed 2x2 matrices (mimo/framing.cc:1344-1367) -------------------
@pytest.mark.parametrize("G", [
    [[1, 0], [0, 1]],
    [[1, 0.5], [0.5j, 1]],
    [[0.25 + 0.1j, -0.3], [0.05j, 0.2 - 0.4j]],
])
def test_kat1_invert_2x2(G):
    G = np.array(G, np.complex64)
    W, gain = orc.invert_2x2(G)
    det = G[0, 0] * G[1, 1] - G[0, 1] * G[1, 0]
    adj = np.array([[G[1, 1], -G[0, 1]], [-G[1, 0], G[0, 0]]])
    assert np.allclose(W, np.conj(det) * adj, rtol=1e-6, atol=1e-7)
    assert np.isclose(gain, 1.0 / abs(det) ** 2, rtol=1e-6)
    # gain * W = G^-1: zero forcing with the real division deferred to the gain vector
    assert np.allclose(gain * W @ G, np.eye(2), atol=1e-5)
    # the product library exports the same function
    W2, g2 = rub.invert_2x2(G)
    assert np.array_equal(W, W2) and gain == g2


# ---- KAT-2: subcarrier allocation counts (mimo/framing.cc:949-1030) -----------------------
@pytest.mark.parametrize("M,use_all,expect", [
    (64, True, (0, 0, 64)),
    (2048, True, (0, 0, 2048)),
    # guard G = M/10, DC null, pilot every 8th (offset 4): i in [1, M/2-G)
    (64, False, (64 - 2 * 25, 2 * 3, 2 * 22)),
    (2048, False, (2048 - 2 * 819, 2 * 102, 2 * 717)),
])
def test_kat2_default_sctype_counts(M, use_all, expect):
    p = orc.init_default_sctype(M, use_all, True)
    assert orc.validate_sctype(p) == expect
    assert np.array_equal(p, rub.ofdmframe_init_default_sctype(M, use_all, True))
    if not use_all:
        assert p[0] == orc.SC_NULL and p[M // 2] == orc.SC_NULL  # DC and Nyquist are null
        assert p[4] == orc.SC_PILOT and p[M - 4] == orc.SC_PILOT


def test_kat2_invalid_sctype_rejected():
    p = np.full(64, 2, np.uint8)
    p[5] = 7
    with pytest.raises(ValueError):
        orc.validate_sctype(p)
    with pytest.raises(rub.RubError):
        rub.ofdmframe_validate_sctype(p)


# ---- KAT-3: S0 has two identical halves; S&C metric is 1.0 on clean S0 --------------------
@pytest.mark.parametrize("M,cp", [(64, 16), (2048, 152)])
def test_kat3_s0_periodicity_and_sc_metric(M, cp):
    ms = orc.Mseq(12, 0o10123, 1)
    S0, s0 = orc.init_S0(None, M, ms)
    assert np.all(S0[1::2] == 0) and np.all(np.abs(S0[0::2]) == 1)
    assert np.allclose(s0[:M // 2], s0[M // 2:], atol=1e-6)       # even bins only => periodic
    assert np.isclose(np.mean(np.abs(s0) ** 2), 1.0, rtol=1e-5)   # sqrt(1/M_S0) scaling
    x = np.concatenate([np.zeros(3 * M), s0[M - cp:], s0, np.zeros(2 * M)]).astype(np.complex64)
    y = orc.sc_metric(M, x)
    t0 = 3 * M
    plateau = y[t0 + M - 1: t0 + M + cp]   # cp+1 samples of exact periodicity (appendix C)
    assert np.allclose(plateau, 1.0, atol=1e-4)
    assert y[t0 + M // 4] < 0.5


# ---- KAT-4: noiseless pre-aligned identity channel: zero errors, G = g I + I/(nac sqrt(M)) --
def test_kat4_identity_channel_q1_bias():
    cfg = rub.Config(M=256, cp_len=18, num_streams=2, num_access_codes=5, num_data_symbols=6,
                     modulation=rub.MOD_QAM16, detector=rub.DET_ZF, flags=rub.FLAG_Q1_IDENTITY_INIT)
    cfg, S1, iq, tx = make_case(cfg, 2, seed=4, n_taps=0, snr_db=200.0, fixed_H=[[1, 0], [0, 1]])
    ref = oracle_run(cfg, S1, iq, tx)
    assert ref["counters"][:, 2].sum() == 0 and ref["counters"][:, 3].sum() == 2 * 2 * 6 * 256
    g = 0.25
    expect = (g + 1.0 / (5 * np.sqrt(256))) * np.eye(2)
    assert np.allclose(ref["G"].transpose(0, 3, 1, 2), expect[None, None], atol=2e-5)
    # without the quirk the bias disappears
    cfg2 = rub.Config(M=256, cp_len=18, num_streams=2, num_access_codes=5, num_data_symbols=6,
                      modulation=rub.MOD_QAM16, detector=rub.DET_ZF, noise_var=cfg.noise_var)
    ref2 = oracle_run(cfg2, S1, iq, tx)
    assert np.allclose(ref2["G"].transpose(0, 3, 1, 2), g * np.eye(2)[None, None], atol=2e-5)


# ---- KAT-5: FFT against float64 numpy --------------------------------------------------
@pytest.mark.parametrize("M", [64, 128, 256, 512, 1024, 2048, 4096])
def test_kat5_fft_vs_numpy(M):
    rng = np.random.default_rng(M)
    x = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(np.complex64)
    X = orc.fft_forward(x)
    R = np.fft.fft(x.astype(np.complex128))
    assert np.abs(X - R).max() / np.abs(R).max() < 1e-6
    xb = orc.fft_backward(X) / M
    assert np.abs(xb - x).max() < 5e-6
    # impulse and single-tone known answers
    e = np.zeros(M, np.complex64); e[3] = 1
    assert np.allclose(orc.fft_forward(e), np.exp(-2j * np.pi * 3 * np.arange(M) / M), atol=1e-6)


# ---- KAT-6: noiseless random-H loopback, z = Xd within 1e-4 -------------------------------
@pytest.mark.parametrize("N,M,q,det", [(2, 128, 4, rub.DET_ZF), (4, 256, 6, rub.DET_ZF),
                                       (8, 256, 8, rub.DET_ZF), (4, 512, 6, rub.DET_MMSE)])
def test_kat6_noiseless_loopback(N, M, q, det):
    cfg = rub.Config(M=M, cp_len=M // 8, num_streams=N, num_access_codes=2, num_data_symbols=3,
                     modulation=q, detector=det, flags=rub.FLAG_MMSE_UNBIASED)
    rng = np.random.default_rng(N * 1000 + M)
    # a well-conditioned random flat channel
    H = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))) / np.sqrt(2) + 2 * np.eye(N)
    cfg, S1, iq, tx = make_case(cfg, 2, seed=6, n_taps=0, snr_db=120.0, fixed_H=H)
    ref = oracle_run(cfg, S1, iq, tx)
    pts = orc.modulate_table(q)
    assert np.abs(ref["eq"] - pts[tx]).max() < 1e-4 * 20   # conditioning-limited fp32, |pts| ~ 1
    assert ref["counters"][:, 2].sum() == 0


# ---- KAT-7: demap round trip and LLR / hard-bit consistency -------------------------------
@pytest.mark.parametrize("q", [2, 4, 6, 8])
def test_kat7_modem_roundtrip_and_llr_signs(q):
    tab = orc.modulate_table(q)
    assert np.isclose(np.mean(np.abs(tab) ** 2), 1.0, rtol=1e-5)
    assert np.allclose(tab, f64.constellation(q), atol=1e-7)
    for s in range(1 << q):
        assert orc.demodulate(q, tab[s]) == s
        assert rub.modem_demodulate(q, tab[s]) == s
        assert rub.modem_modulate(q, s) == tab[s]
    rng = np.random.default_rng(q)
    pts = f64.constellation(q)
    bits = (np.arange(1 << q)[:, None] >> (q - 1 - np.arange(q))[None, :]) & 1
    for _ in range(300):
        z = np.complex64(complex(*rng.uniform(-1.6, 1.6, 2)))
        sym = orc.demodulate(q, z)
        llr = orc.llr(q, z, isig=3.0)
        d = np.abs(complex(z) - pts) ** 2
        assert sym == int(np.argmin(d))                              # nearest point
        ref = np.array([d[bits[:, b] == 1].min() - d[bits[:, b] == 0].min() for b in range(q)]) * 3.0
        assert np.allclose(llr, ref, rtol=1e-4, atol=2e-5)           # brute-force max-log
        hard = (sym >> (q - 1 - np.arange(q))) & 1
        assert np.all((llr < 0) <= (hard == 1)) and np.all((llr > 0) <= (hard == 0))
    # ties go to the lower level (liquid compares v > 0)
    assert orc.demodulate(q, np.complex64(0)) == orc.demodulate(q, np.complex64(complex(-1e-9, -1e-9)))


# ---- mirror fp32 oracle against the independent float64 model -----------------------------
@pytest.mark.parametrize("name,kw,syn", [
    ("zf2", dict(M=128, cp_len=10, num_streams=2, num_access_codes=3, num_data_symbols=4, modulation=4,
                 detector=rub.DET_ZF, flags=rub.FLAG_Q1_IDENTITY_INIT), dict(n_taps=2, snr_db=25.0)),
    ("mmse4", dict(M=256, cp_len=20, num_streams=4, num_access_codes=2, num_data_symbols=3, modulation=6,
                   detector=rub.DET_MMSE, flags=rub.FLAG_MMSE_UNBIASED), dict(n_taps=3, snr_db=28.0)),
    ("mmse4_biased", dict(M=256, cp_len=20, num_streams=4, num_access_codes=2, num_data_symbols=3,
                          modulation=4, detector=rub.DET_MMSE), dict(n_taps=3, snr_db=18.0)),
    ("comb8", dict(M=256, cp_len=24, num_streams=8, num_access_codes=2, num_data_symbols=2, modulation=8,
                   detector=rub.DET_MMSE, estimator=rub.EST_LS_COMB_INTERP, flags=rub.FLAG_MMSE_UNBIASED),
     dict(n_taps=2, snr_db=36.0)),
    ("ragged", dict(M=128, cp_len=12, num_streams=2, num_access_codes=2, num_data_symbols=3, modulation=6,
                    detector=rub.DET_MMSE, sctype="default"), dict(n_taps=2, snr_db=26.0)),
])
def test_mirror_oracle_vs_float64_model(name, kw, syn):
    if kw.get("sctype") == "default":
        kw = dict(kw, sctype=rub.ofdmframe_init_default_sctype(kw["M"], False, True))
    cfg, S1, iq, tx = make_case(rub.Config(**kw), 2, seed=len(name) * 7, **syn)
    ref = oracle_run(cfg, S1, iq, tx)
    q = cfg.q
    for f in range(iq.shape[0]):
        m = f64.rx_frame(iq[f], S1, cfg.M, cfg.cp_len, cfg.N, cfg.nac, cfg.D, q, cfg.detector,
                         cfg.estimator, cfg.P, cfg.flags, cfg.noise_var, cfg.sctype)
        scale = np.abs(m["G"]).max()
        assert np.abs(ref["G"][f] - m["G"]).max() < 1e-4 * scale
        # fp32 normal equations lose cond(G^H G + nv I) * eps: the 1e-4 bar applies to
        # well-conditioned carriers, ill-conditioned ones are bounded by 8 * cond * 2^-23
        tol = f64_tolerance(m, cfg)
        err = np.abs(ref["eq"][f] - m["eq"]) / max(1.0, np.abs(m["eq"]).max())
        assert (err <= tol[None, None, :]).all(), (err / tol[None, None, :]).max()
        assert np.median(err) < 1e-4
        # hard decisions agree wherever the float64 decision margin is not razor thin
        safe = m["margin"] > 1e-4 + 20 * tol[None, None, :]
        assert np.array_equal(ref["rx_data"][f][safe], m["sym"][safe])
        assert safe.mean() > 0.3
        # LLRs: relative to the LLR scale of each symbol
        lscale = np.maximum(np.abs(m["llr"]).max(axis=-1, keepdims=True), 1.0)
        ok = np.abs(ref["llr"][f] - m["llr"]) <= 20 * tol[None, None, :, None] * lscale
        assert ok[safe].all()


def f64_tolerance(m, cfg):
    """per occupied carrier: max(1e-4, 8 * cond(G^H G + nv I) * eps_fp32)"""
    occ = np.arange(cfg.M) if cfg.sctype is None else np.nonzero(cfg.sctype != 0)[0]
    nv = cfg.noise_var if cfg.detector == rub.DET_MMSE else 0.0
    kap = np.array([np.linalg.cond(m["G"][:, :, k].conj().T @ m["G"][:, :, k] + nv * np.eye(cfg.N)) for k in occ])
    return np.maximum(1e-4, 8 * kap * 2.0 ** -23)


# ---- KAT-8: seeded AWGN runs with stored counters (tests/golden/kat8_counters.json) --------
def test_kat8_golden_counters():
    import json, os
    path = os.path.join(os.path.dirname(__file__), "golden", "kat8_counters.json")
    gold = json.load(open(path))
    for case in gold["cases"]:
        cfg = rub.Config(**case["config"])
        cfg, S1, iq, tx = make_case(cfg, case["frames"], seed=case["seed"], **case["synth"])
        ref = oracle_run(cfg, S1, iq, tx)
        assert ref["counters"].tolist() == case["counters"], case["name"]
        assert np.isclose(cfg.noise_var, case["noise_var"], rtol=1e-6)
