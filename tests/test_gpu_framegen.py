"""Row f4 of SURVEY.md 8: the batched GPU framegen (access codes + assemble_mimo_packet,
mimo/framing.cc:191-235) is bit-identical to the host framegen and, through it, to the oracle."""
import numpy as np
import pytest

import rub_mimo_b200 as rub
from oracle import orc
from util import to_orc

pytestmark = pytest.mark.gpu


def _host_waveform(cfg, tx, gain):
    """[F][N][(T+D)*L] from the host framegen, frame by frame."""
    fg = rub.FrameGen(cfg)
    F = tx.shape[0]
    out = np.zeros((F, cfg.N, (cfg.T + cfg.D) * cfg.L), np.complex64)
    train = fg.write_comb_words() if cfg.estimator == rub.EST_LS_COMB_INTERP else fg.write_sync_words()[:, cfg.L:]
    g = np.float32(gain)
    tab = orc.modulate_table(cfg.q)
    for f in range(F):
        out[f, :, : cfg.T * cfg.L] = train
        for d in range(cfg.D):
            syms = tab[tx[f, :, d]]
            out[f, :, (cfg.T + d) * cfg.L:(cfg.T + d + 1) * cfg.L] = fg.assemble_mimo_packet(syms)
    re, im = out.real * g, out.imag * g          # cscale: two separate fp32 multiplies
    return (re + 1j * im).astype(np.complex64)


CASES = {
    "c1": dict(M=64, cp_len=16, num_streams=2, num_access_codes=3, num_data_symbols=4, modulation=rub.MOD_QPSK),
    "n1_m128_16qam": dict(M=128, cp_len=9, num_streams=1, num_access_codes=2, num_data_symbols=3, modulation=rub.MOD_QAM16),
    "c2": dict(M=1024, cp_len=72, num_streams=2, num_access_codes=2, num_data_symbols=3, modulation=rub.MOD_QAM16),
    "c3": dict(M=2048, cp_len=152, num_streams=4, num_access_codes=2, num_data_symbols=2, modulation=rub.MOD_QAM64),
    "c4_comb": dict(M=4096, cp_len=288, num_streams=8, num_access_codes=2, num_data_symbols=1,
                    modulation=rub.MOD_QAM256, estimator=rub.EST_LS_COMB_INTERP),
    "m512_nulls": dict(M=512, cp_len=36, num_streams=2, num_access_codes=2, num_data_symbols=3,
                       modulation=rub.MOD_QAM64, sctype="default_nulls"),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("gain", [1.0, 0.25])
def test_gpu_framegen_is_bit_identical_to_host(name, gain):
    import torch
    kw = dict(CASES[name])
    if kw.get("sctype") == "default_nulls":
        kw["sctype"] = rub.ofdmframe_init_default_sctype(kw["M"], use_all_carriers=False, add_null_carriers=True)
    cfg = rub.Config(**kw)
    F = 3
    rng = np.random.default_rng(0xF4 + len(name))
    tx = rng.integers(0, 1 << cfg.q, size=(F, cfg.N, cfg.D, cfg.Mo), dtype=np.uint8)
    rx = rub.Receiver(cfg)
    got = rx.framegen_batch(torch.from_numpy(tx).cuda(), baseband_gain=gain)
    torch.cuda.synchronize()
    ref = _host_waveform(cfg, tx, gain)
    assert np.array_equal(got.cpu().numpy(), ref)


def _oracle_waveform(cfg, tx, gain):
    """The same waveform from the ORACLE's transmit side (orc_write_sync_words / orc_write_comb_words /
    orc_assemble_mimo_packet, oracle/rub_oracle.c) fed the same access-code tables."""
    S1, s1 = rub.default_S1(cfg)
    S0, s0 = rub.default_S0(cfg)
    oc = to_orc(cfg)
    F = tx.shape[0]
    out = np.zeros((F, cfg.N, (cfg.T + cfg.D) * cfg.L), np.complex64)
    train = orc.write_comb_words(oc, S1) if cfg.estimator == rub.EST_LS_COMB_INTERP else orc.write_sync_words(oc, s0, s1)[:, cfg.L:]
    tab = orc.modulate_table(cfg.q)
    for f in range(F):
        out[f, :, : cfg.T * cfg.L] = train
        for d in range(cfg.D):
            out[f, :, (cfg.T + d) * cfg.L:(cfg.T + d + 1) * cfg.L] = orc.assemble_mimo_packet(oc, tab[tx[f, :, d]])
    g = np.float32(gain)
    return (out.real * g + 1j * (out.imag * g)).astype(np.complex64)


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4_comb", "m512_nulls"])
def test_gpu_framegen_matches_the_oracle(name):
    """GPU framegen against the oracle's framegen directly (value-identical; signed zeros aside)."""
    import torch
    kw = dict(CASES[name])
    if kw.get("sctype") == "default_nulls":
        kw["sctype"] = rub.ofdmframe_init_default_sctype(kw["M"], use_all_carriers=False, add_null_carriers=True)
    cfg = rub.Config(**kw)
    F = 2
    rng = np.random.default_rng(0xF5 + len(name))
    tx = rng.integers(0, 1 << cfg.q, size=(F, cfg.N, cfg.D, cfg.Mo), dtype=np.uint8)
    rx = rub.Receiver(cfg)
    got = rx.framegen_batch(torch.from_numpy(tx).cuda(), baseband_gain=0.25).cpu().numpy()
    ref = _oracle_waveform(cfg, tx, 0.25)
    assert got.shape == ref.shape and bool((got == ref).all())


def test_gpu_framegen_loops_back_through_the_receiver():
    """framegen_batch -> identity channel -> process_batch recovers the symbols (no host framegen involved)."""
    import torch
    cfg = rub.preset("C3", num_data_symbols=4)
    F = 6
    rng = np.random.default_rng(44)
    tx = rng.integers(0, 1 << cfg.q, size=(F, cfg.N, cfg.D, cfg.Mo), dtype=np.uint8)
    rx = rub.Receiver(cfg)
    txd = torch.from_numpy(tx).cuda()
    wave = rx.framegen_batch(txd, baseband_gain=1.0)
    out = rx.process_batch(wave, out_mask=rub.OUT_RXDATA | rub.OUT_EQ, tx_data=txd)
    rx.sync()
    assert np.array_equal(out["rx_data"].cpu().numpy(), tx)
    c = rx.read_counters()
    assert c[:, 0].sum() == 0 and c[:, 2].sum() == 0 and c[:, 3].sum() == F * cfg.N * cfg.D * cfg.Mo
