"""Shared helpers of the test-suite: map a product Config onto the oracle's, generate seeded
inputs with the product's synthetic source, run the oracle."""
import numpy as np

import rub_mimo_b200 as rub
from oracle import orc


def to_orc(cfg):
    return orc.Config(cfg.M, cfg.cp_len, cfg.N, cfg.nac, cfg.D, cfg.q, detector=cfg.detector,
                      estimator=cfg.estimator, P=cfg.P, flags=cfg.flags, noise_var=cfg.noise_var,
                      sctype=cfg.sctype)


def make_case(cfg, n_frames, seed, n_taps=4, snr_db=30.0, fixed_H=None, n_threads=0, fixed_H_ri=None):
    """Returns (cfg with noise_var filled in, S1, iq, tx_data)."""
    if fixed_H_ri is not None:  # JSON form: [re, im] pairs
        a = np.asarray(fixed_H_ri, np.float32)
        fixed_H = a[..., 0] + 1j * a[..., 1]
    S1, s1 = rub.default_S1(cfg)
    iq, tx, nv = rub.synth_frames(cfg, n_frames, seed, n_taps=n_taps, snr_db=snr_db,
                                  fixed_H=fixed_H, S1=S1, s1=s1, n_threads=n_threads)
    return cfg.with_noise_var(nv), S1, iq, tx


def oracle_run(cfg, S1, iq, tx, n_threads=4):
    return orc.rx_batch(to_orc(cfg), S1, iq, tx_data=tx, n_threads=n_threads)


ALL_OUT = rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS | rub.OUT_RXDATA | rub.OUT_G


def gpu_run(cfg, S1, iq, tx, path=rub.PATH_AUTO, out_mask=ALL_OUT):
    import torch
    rx = rub.Receiver(cfg, S1)
    rx.set_path(path)
    d_iq = torch.from_numpy(iq).cuda()
    d_tx = torch.from_numpy(tx).cuda() if tx is not None else None
    out = rx.process_batch(d_iq, out_mask=out_mask, tx_data=d_tx)
    rx.sync()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    res["counters"] = rx.read_counters()
    res["path"] = rx.last_path
    rx.close()
    return res


def assert_parity(ref, got, q, tol_only=False):
    """Bit-exact on integer outputs; float outputs bit-exact against the mirror oracle."""
    assert np.array_equal(ref["rx_data"], got["rx_data"]), "hard decisions differ"
    assert np.array_equal(ref["bits"], got["bits"]), "packed bits differ"
    assert np.array_equal(ref["counters"], got["counters"]), (ref["counters"], got["counters"])
    for k in ("G", "eq", "llr"):
        a, b = ref[k], got[k]
        if not np.array_equal(a, b):
            d = np.abs(a.astype(np.complex128) - b.astype(np.complex128))
            raise AssertionError(f"{k}: max abs diff {d.max():.3e} at {np.unravel_index(d.argmax(), d.shape)} "
                                 f"(ref {a.flat[d.argmax()]}, got {b.flat[d.argmax()]}), "
                                 f"{np.count_nonzero(d)} of {d.size} differ")
