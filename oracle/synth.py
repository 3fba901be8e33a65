"""Synthetic pre-aligned MIMO-OFDM frames built with the ORACLE's own transmit side (test
infrastructure, like everything under oracle/): bench.py's `--impl reference` arm uses it so that the
CPU arm never loads the product library.

Frames follow framegen's layout (mimo/framing.cc:191-235): nac*N TDMA access-code symbols (code c,
transmitter t at symbol c*N + t, the other transmitters silent) followed by D payload symbols from
assemble_mimo_packet; every (rx, tx) link is an n_taps Rayleigh FIR, plus AWGN at snr_db.
"""
import numpy as np

from . import orc

# generator polynomials of the per-stream access-code LFSRs (mimo/config.h:30-35 for the first two)
_POLY13 = [0o20033, 0o20047, 0o20065, 0o20071, 0o20137, 0o20161, 0o20213, 0o20327]


def default_S1(cfg):
    """[N][nac][M] frequency-domain access codes and their time-domain symbols."""
    p = cfg.sctype if cfg.sctype is not None else orc.init_default_sctype(cfg.M, True, True)
    S1 = np.empty((cfg.N, cfg.nac, cfg.M), np.complex64)
    s1 = np.empty((cfg.N, cfg.nac, cfg.M), np.complex64)
    for t in range(cfg.N):
        S1[t], s1[t] = orc.init_S1(p, cfg.M, cfg.nac, orc.Mseq(13, _POLY13[t % len(_POLY13)], 1))
    return S1, s1


def synth_frames(cfg, n_frames, seed, n_taps=8, snr_db=30.0, baseband_gain=0.25):
    """Returns (S1, iq [F][N][(T+D)*L] complex64, tx_data [F][N][D][Mo] uint8, noise_var)."""
    rng = np.random.default_rng(seed)
    N, M, L, D, q, nac = cfg.N, cfg.M, cfg.L, cfg.D, cfg.q, cfg.nac
    Mo = cfg.Mo
    S1, s1 = default_S1(cfg)
    table = orc.modulate_table(q)
    T = nac * N
    row = (T + D) * L
    # training part: transmitter t sends access code c at symbol c*N + t
    train = np.zeros((N, T * L), np.complex64)
    for c in range(nac):
        for t in range(N):
            sym = c * N + t
            x = s1[t, c] * np.float32(baseband_gain)
            train[t, sym * L:sym * L + cfg.cp_len] = x[M - cfg.cp_len:]
            train[t, sym * L + cfg.cp_len:(sym + 1) * L] = x
    iq = np.empty((n_frames, N, row), np.complex64)
    tx_data = rng.integers(0, 1 << q, size=(n_frames, N, D, Mo), dtype=np.uint8)
    sig_pow = 0.0
    taps = max(1, n_taps)
    for f in range(n_frames):
        tx = np.empty((N, row), np.complex64)
        tx[:, :T * L] = train
        for d in range(D):
            tx[:, (T + d) * L:(T + d + 1) * L] = orc.assemble_mimo_packet(cfg, table[tx_data[f, :, d]]) * np.float32(baseband_gain)
        h = (rng.standard_normal((N, N, taps)) + 1j * rng.standard_normal((N, N, taps))) / np.sqrt(2.0 * taps)
        rx = np.zeros((N, row), np.complex128)
        for r in range(N):
            for t in range(N):
                rx[r] += np.convolve(tx[t], h[r, t])[:row]
        sig_pow += float(np.mean(np.abs(rx[:, T * L:]) ** 2))
        iq[f] = rx
    sig_pow /= n_frames
    nv_time = sig_pow / (10.0 ** (snr_db / 10.0))
    noise = rng.standard_normal(iq.shape) + 1j * rng.standard_normal(iq.shape)
    iq += (noise * np.sqrt(nv_time / 2.0)).astype(np.complex64)
    # noise variance seen by the detector on an occupied carrier after the receiver's 1/sqrt(Mo) scaling
    noise_var = float(nv_time * M / Mo)
    return S1, iq, tx_data, noise_var
