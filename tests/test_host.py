"""Host-side rows a8/a9 of the product (msequence, allocation, preambles, framegen, modem,
synthetic source) against the oracle's independent restatement: bit-exact."""
import numpy as np
import pytest

import rub_mimo_b200 as rub
from oracle import orc
from util import to_orc


def bits_eq(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


@pytest.mark.parametrize("m,g", [(12, 0o10123), (12, 0o10151), (13, 0o20033), (13, 0o20047)])
def test_msequence_matches_oracle_and_is_maximal(m, g):
    a, b = rub.MSequence(m, g, 1), orc.Mseq(m, g, 1)
    n = (1 << m) - 1
    sa = [a.advance() for _ in range(2 * n)]
    sb = [b.advance() for _ in range(2 * n)]
    assert sa == sb
    assert sa[:n] == sa[n:] and sum(sa[:n]) == (n + 1) // 2   # period 2^m-1, balanced
    a.reset(); b.reset()
    assert [a.generate_symbol(3) for _ in range(50)] == [b.symbol(3) for _ in range(50)]


def test_default_polys_are_distinct_maximal_length():
    polys = [rub.lib().rub_default_lfsr_poly(i) for i in range(8)]
    assert polys[:2] == [0o20033, 0o20047]          # LFSR_LARGE_0/1_GEN_POLY, config.h:74-75
    assert len(set(polys)) == 8
    for g in polys:
        ms = rub.MSequence(13, g, 1)
        v0, n = ms.ms.v, 0
        while True:
            ms.advance(); n += 1
            if ms.ms.v == v0 or n > 8200:
                break
        assert n == 8191


@pytest.mark.parametrize("M,use_all", [(64, True), (256, False), (2048, False), (4096, True)])
def test_preambles_bit_exact(M, use_all):
    p = rub.ofdmframe_init_default_sctype(M, use_all, True)
    assert np.array_equal(p, orc.init_default_sctype(M, use_all, True))
    S0, s0 = rub.ofdmframe_init_S0(p, M, rub.MSequence(12, 0o10123))
    S0o, s0o = orc.init_S0(p, M, orc.Mseq(12, 0o10123))
    assert bits_eq(S0, S0o) and bits_eq(s0, s0o)
    S1, s1 = rub.ofdmframe_init_S1(p, M, 3, rub.MSequence(13, 0o20047))
    S1o, s1o = orc.init_S1(p, M, 3, orc.Mseq(13, 0o20047))
    assert bits_eq(S1, S1o) and bits_eq(s1, s1o)
    assert np.all(S1[:, p == 0] == 0) and np.all(np.abs(S1[:, p != 0]) == 1)


@pytest.mark.parametrize("name", ["C1", "C2", "C3"])
def test_framegen_bit_exact(name):
    cfg = rub.preset(name, num_data_symbols=2)
    oc = to_orc(cfg)
    S1, s1 = rub.default_S1(cfg)
    S0, s0 = rub.default_S0(cfg)
    fg = rub.FrameGen(cfg)
    assert fg.get_num_streams() == cfg.N
    tx = fg.write_sync_words()
    assert bits_eq(tx, orc.write_sync_words(oc, s0, s1))
    # S0 only on stream 0, TDMA afterwards (framing.cc:183-204)
    L = cfg.L
    assert np.all(tx[1:, :L] == 0)
    for ac in range(cfg.nac * cfg.N):
        active = ac % cfg.N
        seg = tx[:, (ac + 1) * L:(ac + 2) * L]
        assert np.all(seg[np.arange(cfg.N) != active] == 0) and np.any(seg[active] != 0)
    rng = np.random.default_rng(5)
    tab = orc.modulate_table(cfg.q)
    syms = tab[rng.integers(0, 1 << cfg.q, (cfg.N, cfg.Mo))]
    pk = fg.assemble_mimo_packet(syms)
    assert bits_eq(pk, orc.assemble_mimo_packet(oc, syms))
    assert bits_eq(pk[:, :cfg.cp_len], pk[:, -cfg.cp_len:])           # cyclic prefix
    assert np.isclose(np.mean(np.abs(pk) ** 2), 1.0, rtol=0.1)        # unit power per tx


def test_comb_words_bit_exact():
    cfg = rub.preset("C4", M=256, cp_len=24)
    S1, _ = rub.default_S1(cfg)
    assert bits_eq(rub.FrameGen(cfg).write_comb_words(), orc.write_comb_words(to_orc(cfg), S1))


def test_synth_is_deterministic_and_shard_independent():
    cfg = rub.preset("C2", M=256, cp_len=18, num_data_symbols=3)
    iq, tx, nv = rub.synth_frames(cfg, 6, seed=11, n_taps=3, snr_db=20.0, n_threads=3)
    iq2, tx2, nv2 = rub.synth_frames(cfg, 6, seed=11, n_taps=3, snr_db=20.0, n_threads=1)
    assert bits_eq(iq, iq2) and np.array_equal(tx, tx2) and nv == nv2
    # frames 2..5 regenerated on their own (a different shard) are identical
    iq3, tx3, _ = rub.synth_frames(cfg, 4, seed=11, n_taps=3, snr_db=20.0, first_frame=2)
    assert bits_eq(iq[2:], iq3) and np.array_equal(tx[2:], tx3)
    iq4, _, _ = rub.synth_frames(cfg, 2, seed=12, n_taps=3, snr_db=20.0)
    assert not bits_eq(iq[:2], iq4)
    assert tx.max() < 16 and len(np.unique(tx)) == 16
    # SNR bookkeeping: noise_var = N g^2 / snr
    assert np.isclose(nv, 2 * 0.25 ** 2 / 10 ** 2.0, rtol=1e-6)
    assert iq.shape == (6, 2, cfg.row_samples)


def test_synth_rejects_bad_arguments():
    cfg = rub.preset("C2")
    with pytest.raises(rub.RubError):
        rub.synth_frames(cfg, 1, seed=1, n_taps=0, fixed_H=None)
    with pytest.raises(rub.RubError):
        rub.synth_frames(cfg, 1, seed=1, n_taps=500)
