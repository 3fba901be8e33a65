#!/usr/bin/env python
"""Runs the REFERENCE's own mimo/framing.cc (oracle/_ref/libref_framing.so, `make -C oracle ref`:
compiled where it lies against the stand-in headers of oracle/shim/) on seeded inputs and commits
what it produced as tests/golden/ref_*.npz.  tests/test_ref_fixtures.py then pins the oracle (and,
on the GPU box, the CUDA path behind the framing.h facade) to these reference outputs.

Only this container has /root/reference; the fixtures travel, the reference does not.

usage: python oracle/make_ref_fixtures.py [--check]     (--check: compare only, write nothing)
"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rub_mimo_b200 as rub  # noqa: E402  (host-side synthetic source only; no GPU involved)
from oracle import orc  # noqa: E402
from util import to_orc  # noqa: E402


class RefSyncResult(C.Structure):
    _fields_ = [("state", C.c_int32), ("sync_index", C.c_uint64), ("num_samples_processed", C.c_uint64),
                ("plateau_start", C.c_uint64 * 2), ("plateau_end", C.c_uint64 * 2), ("symbols", C.c_uint32)]


def ref_lib():
    return C.CDLL(os.path.join(HERE, "_ref", "libref_framing.so"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ref_sctype(M):
    p = np.zeros(M, np.uint8)
    n = [C.c_uint(), C.c_uint(), C.c_uint()]
    ref_lib().ref_default_sctype(C.c_uint(M), _p(p), C.byref(n[0]), C.byref(n[1]), C.byref(n[2]))
    return p, tuple(x.value for x in n)


def ref_framegen(M, cp, nac, p, symbols):
    """symbols [D][2][Mo] complex64 -> tx [2][(nac*2+1+D)*L] complex64 (S0, access codes, packets)."""
    D, N, Mo = symbols.shape
    tx = np.zeros((2, (nac * 2 + 1 + D) * (M + cp)), np.complex64)
    n = ref_lib().ref_framegen(C.c_uint(M), C.c_uint(cp), C.c_uint(nac), _p(p), _p(np.ascontiguousarray(symbols)),
                               C.c_uint(D), C.c_uint(Mo), _p(tx))
    assert n == tx.shape[1]
    return tx


def ref_framesync(M, cp, nac, p, cap, Mo, max_syms, chunk=4096):
    cap = [np.ascontiguousarray(r, np.complex64) for r in cap]
    res = RefSyncResult()
    G = np.zeros((M, 2, 2), np.complex64)
    eq = np.zeros((2, max_syms, Mo), np.complex64)
    ref_lib().ref_framesync(C.c_uint(M), C.c_uint(cp), C.c_uint(nac), _p(p), _p(cap[0]), _p(cap[1]),
                            C.c_uint64(cap[0].size), C.c_uint(chunk), C.byref(res), _p(G), _p(eq),
                            C.c_uint(max_syms), C.c_uint(Mo))
    return dict(state=res.state, sync_index=res.sync_index, num_samples_processed=res.num_samples_processed,
                plateau_start=list(res.plateau_start), plateau_end=list(res.plateau_end), symbols=res.symbols,
                G=G, eq=eq)


CASES = {
    # name: (M, cp, nac, D = PID_MAX of mimo/config.h, seed, snr_db, flat 2x2 channel)
    "ref_c1_m64": (64, 16, 20, 1000, 0xC1, 30.0, [[1, 0.5], [0.5j, 1]]),
    "ref_m256": (256, 20, 4, 1000, 0xC2, 27.0, [[0.9, -0.3j], [0.2 + 0.4j, 1.1]]),
    # 3-tap Rayleigh links: the access codes of different links peak at different offsets (quirk Q2)
    "ref_m512_multipath": (512, 36, 8, 1000, 0xC5, 28.0, None),
    # the reference's default geometry (mimo/config.h:65-66, :94): 2048 carriers, cp 152, 20 access codes
    "ref_m2048_default": (2048, 152, 20, 1000, 0xC6, 30.0, [[1, 0.5], [0.5j, 1]]),
}


def run_case(name):
    M, cp, nac, D, seed, snr, H = CASES[name]
    p, (n_null, n_pilot, n_data) = ref_sctype(M)
    cfg = rub.preset("C1", M=M, cp_len=cp, num_access_codes=nac, num_data_symbols=D, sctype=p)
    assert cfg.Mo == n_pilot + n_data
    S1, s1 = rub.default_S1(cfg)
    S0, s0 = rub.default_S0(cfg)
    lead = (nac * 2 + 1) * cfg.L                       # flush burst, mimo/main.cc:941-943
    iq, tx, nv = rub.synth_frames(cfg, 1, seed, n_taps=0 if H is not None else 3, snr_db=snr, fixed_H=H,
                                  include_s0=True, lead_zeros=lead, S1=S1, s1=s1)
    cap = iq[0]
    # --- transmit side: the reference's framegen on the same symbols
    tab = orc.modulate_table(cfg.q)
    syms = np.ascontiguousarray(tab[tx[0]].transpose(1, 0, 2)[:8 if M <= 512 else 2])  # [D'][2][Mo], first packets
    ref_tx = ref_framegen(M, cp, nac, p, syms)
    # --- receive side: the reference's framesync on the capture
    ref = ref_framesync(M, cp, nac, p, cap, cfg.Mo, D + 8)
    ours = orc.framesync_execute(to_orc(cfg), S0, S1, cap)
    return dict(cfg=cfg, p=p, cap=cap, tx=tx[0], syms=syms, ref_tx=ref_tx, ref=ref, ours=ours, S0=S0, S1=S1, s0=s0, s1=s1,
                lead=lead)


def report(name, r):
    ref, ours, cfg = r["ref"], r["ours"], r["cfg"]
    print(f"== {name}: reference state {ref['state']} sync_index {ref['sync_index']} nsp {ref['num_samples_processed']} "
          f"plateau {ref['plateau_start']} {ref['plateau_end']} symbols {ref['symbols']}")
    print(f"   oracle    state {ours['state']} sync_index {ours['sync_index']} nsp {ours['num_samples_processed']} "
          f"plateau {ours['plateau_start']} {ours['plateau_end']} symbols {ours['symbols_decoded']}")
    n = min(ref["symbols"], ours["symbols_decoded"], ref["eq"].shape[1], ours["eq"].shape[1])
    dG = np.abs(ref["G"] - ours["G"]).max()
    de = np.abs(ref["eq"][:, :n] - ours["eq"][:, :n]).max()
    print(f"   max|G_ref - G_oracle| = {dG:.3e} (bit-equal: {np.array_equal(ref['G'], ours['G'])}); "
          f"max|eq_ref - eq_oracle| over {n} symbols = {de:.3e} (bit-equal: {np.array_equal(ref['eq'][:, :n], ours['eq'][:, :n])})")
    # transmit side against the host framegen of the product (rub_framegen_*)
    fg = rub.FrameGen(cfg)
    mine = np.concatenate([fg.write_sync_words()] + [fg.assemble_mimo_packet(s) for s in r["syms"]], axis=1)
    print(f"   framegen: max|tx_ref - tx_ours| = {np.abs(r['ref_tx'] - mine).max():.3e} (bit-equal: {np.array_equal(r['ref_tx'], mine)})")


def invert_fixture(check):
    """invert() of the reference (mimo/framing.cc:1344-1367) on seeded 2x2 matrices, including
    badly conditioned and tiny ones."""
    rng = np.random.default_rng(0xA3)
    G = (rng.standard_normal((96, 2, 2)) + 1j * rng.standard_normal((96, 2, 2))).astype(np.complex64)
    G[32:48] *= np.float32(1e-3)
    G[48:64, 1] = G[48:64, 0] * np.complex64(1 + 1e-3j)          # nearly singular
    G[64:80] = (np.eye(2) * 0.25 + 1e-3).astype(np.complex64)     # the identity-biased estimate of quirk Q1
    W = np.zeros_like(G)
    gain = np.zeros(96, np.float32)
    ref_lib().ref_invert(_p(G), _p(W), _p(gain), C.c_uint(96))
    mine_W = np.zeros_like(G)
    mine_g = np.zeros(96, np.float32)
    for i in range(96):
        mine_W[i], mine_g[i] = rub.invert_2x2(G[i])
    print(f"== ref_invert: W bit-equal {np.array_equal(W.view(np.uint32), mine_W.view(np.uint32))}, "
          f"gain bit-equal {np.array_equal(gain.view(np.uint32), mine_g.view(np.uint32))}")
    if not check:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_invert.npz"), G=G, W=W, gain=gain)
        print("   wrote tests/golden/ref_invert.npz")


def main():
    check = "--check" in sys.argv
    if "--no-invert" not in sys.argv:
        invert_fixture(check)
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    for name in (only or CASES):
        r = run_case(name)
        report(name, r)
        if check:
            continue
        ref = r["ref"]
        n = ref["symbols"]
        eq = ref["eq"][:, :n]
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", name + ".npz"),
            M=r["cfg"].M, cp_len=r["cfg"].cp_len, nac=r["cfg"].nac, D=r["cfg"].D, sctype=r["p"],
            seed=CASES[name][4], snr_db=CASES[name][5], n_taps=0 if CASES[name][6] is not None else 3,
            H=np.asarray(CASES[name][6] if CASES[name][6] is not None else [[0, 0], [0, 0]], np.complex64), lead=r["lead"],
            state=ref["state"], sync_index=ref["sync_index"], num_samples_processed=ref["num_samples_processed"],
            plateau_start=np.asarray(ref["plateau_start"]), plateau_end=np.asarray(ref["plateau_end"]), symbols=n,
            G=ref["G"], eq_head=eq[:, :16 if r["cfg"].M <= 512 else 4], eq_tail=eq[:, -4:], eq_sha256=hashlib.sha256(eq.tobytes()).hexdigest(),
            eq_sha256_D=hashlib.sha256(np.ascontiguousarray(eq[:, :r["cfg"].D]).tobytes()).hexdigest(),
            syms=r["syms"], ref_tx=r["ref_tx"], cap_sha256=hashlib.sha256(r["cap"].tobytes()).hexdigest())
        print("   wrote tests/golden/" + name + ".npz")


if __name__ == "__main__":
    main()
