"""Config C1: the reference's full receive loop (leading zeros + S0 + Schmidl&Cox plateau +
timing search + LS + invert + decode, mimo/framing.cc:471-886) restated in the oracle, and the
CUDA chain fed the same capture through the per-link timing table (quirks Q1, Q2, Q4)."""
import numpy as np
import pytest

import rub_mimo_b200 as rub
from oracle import orc
from util import to_orc

H_C1 = [[1, 0.5], [0.5j, 1]]   # S0 is sent on tx 0 only: an identity channel never syncs on rx 1


def capture(seed=0xC1, snr_db=30.0, D=1000):
    cfg = rub.preset("C1", num_data_symbols=D)
    S1, s1 = rub.default_S1(cfg)
    S0, s0 = rub.default_S0(cfg)
    lead = (cfg.nac * cfg.N + 1) * cfg.L                 # flush burst, mimo/main.cc:941-943
    iq, tx, nv = rub.synth_frames(cfg, 1, seed, n_taps=0, snr_db=snr_db, fixed_H=H_C1, include_s0=True,
                                  lead_zeros=lead, S1=S1, s1=s1)
    return cfg, S0, S1, iq[0], tx[0], lead


def test_faithful_loopback_syncs_and_decodes_without_errors():
    cfg, S0, S1, cap, tx, lead = capture()
    r = orc.framesync_execute(to_orc(cfg), S0, S1, cap)
    assert r["rc"] == 0 and r["state"] == 3                      # STATE_MIMO
    t0 = lead                                                    # first CP sample of S0
    # plateau starts ~ t0 + M - 1 (appendix C) and lasts longer than cp on both streams
    for s in range(2):
        assert abs(int(r["plateau_start"][s]) - (t0 + cfg.M - 1)) <= 3
        assert r["plateau_end"][s] - r["plateau_start"][s] > cfg.cp_len
    assert r["sync_index"] == sum(r["plateau_start"]) // 2
    # timing search finds every access code exactly one symbol apart
    first = lead - r["window_start"] + cfg.L + cfg.cp_len
    expect = first + cfg.L * np.arange(cfg.nac * cfg.N)
    assert np.array_equal(r["corr_indices"], np.stack([expect, expect]))
    assert r["payload_start"] == expect[-1] + cfg.M               # quirk Q4
    assert r["symbols_decoded"] >= cfg.D                          # quirk Q14: callback keeps PID_MAX
    rx = np.array([orc.demodulate(cfg.q, z) for z in r["eq"].reshape(-1)]).reshape(tx.shape)
    assert np.array_equal(rx, tx)                                 # SER 0 at 30 dB
    # G = g*H + I/(nac*sqrt(M)) (quirk Q1 bias), reference layout [k][rx][tx]
    g = 0.25 * np.array(H_C1) + np.eye(2) / (cfg.nac * np.sqrt(cfg.M))
    assert np.abs(r["G"] - g[None]).max() < 0.02
    # normalize_gain * W = G^-1
    Ginv = np.linalg.inv(r["G"][5].astype(np.complex128))
    assert np.allclose(r["gain"][5] * r["W"][5], Ginv, rtol=1e-3, atol=1e-3)


def test_no_sync_is_a_first_class_outcome():
    cfg, S0, S1, cap, tx, lead = capture(snr_db=30.0, D=50)
    noise = cap[:, :lead]   # leading zeros + noise only: no plateau
    r = orc.framesync_execute(to_orc(cfg), S0, S1, noise)
    assert r["rc"] == 1 and r["state"] == 0                      # still STATE_SEEK_PLATEAU


@pytest.mark.gpu
def test_gpu_chain_reproduces_faithful_loopback():
    import torch
    cfg, S0, S1, cap, tx, lead = capture(D=1000)
    r = orc.framesync_execute(to_orc(cfg), S0, S1, cap)
    assert r["rc"] == 0
    Wlen = cfg.L * (cfg.nac * cfg.N + 4) + cfg.D * cfg.L
    w0 = int(r["window_start"])
    window = np.ascontiguousarray(cap[:, w0:w0 + Wlen])[None]     # [1][N][Wlen]
    rx = rub.Receiver(cfg, S1)
    out = rx.process_batch(torch.from_numpy(window).cuda(), out_mask=rub.OUT_EQ | rub.OUT_RXDATA | rub.OUT_G,
                           tx_data=torch.from_numpy(tx[None]).cuda(),
                           timing=torch.from_numpy(r["corr_indices"][None].astype(np.int32)).cuda(),
                           payload_start=torch.tensor([r["payload_start"]], dtype=torch.int32).cuda())
    rx.sync()
    assert rx.last_path == rub.PATH_STAGED                        # timing tables -> staged path
    assert np.array_equal(out["eq"].cpu().numpy()[0], r["eq"])    # bit-exact equalised symbols
    assert np.array_equal(out["rx_data"].cpu().numpy()[0], tx)
    assert np.array_equal(out["G"].cpu().numpy()[0].transpose(2, 0, 1), r["G"])
    c = rx.read_counters()
    assert c[:, 2].sum() == 0 and c[:, 3].sum() == 2 * cfg.D * cfg.Mo
