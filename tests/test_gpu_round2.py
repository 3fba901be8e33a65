"""Round-2 code paths through the C ABI against the CPU oracle: the warp-specialised fused kernel's two
instances (the usual output set without per-task null checks, and the generic one), the staged path's TMA task
records (N >= 4) and its fused comb-LS + weights kernel, and the capture call with reusable pinned buffers."""
import numpy as np
import pytest

import rub_mimo_b200 as rub
from util import make_case, oracle_run

pytestmark = pytest.mark.gpu

KEYS = {rub.OUT_EQ: "eq", rub.OUT_LLR: "llr", rub.OUT_BITS: "bits", rub.OUT_RXDATA: "rx_data", rub.OUT_G: "G"}


def _run_masks(cfg, S1, iq, tx, path, masks, kernel=None):
    import torch
    ref = oracle_run(cfg, S1, iq, tx)
    d_iq, d_tx = torch.from_numpy(iq).cuda(), torch.from_numpy(tx).cuda()
    for mask, with_tx in masks:
        rx = rub.Receiver(cfg, S1)
        rx.set_path(path)
        out = rx.process_batch(d_iq, out_mask=mask, tx_data=d_tx if with_tx else None)
        rx.sync()
        for bit, k in KEYS.items():
            if mask & bit:
                assert np.array_equal(out[k].cpu().numpy(), ref[k]), (hex(mask), with_tx, k)
            else:
                assert k not in out
        if with_tx:
            assert np.array_equal(rx.read_counters(), ref["counters"]), (hex(mask), ref["counters"])
        if kernel:
            assert rx.last_kernel() == kernel, rx.last_kernel()
        rx.close()


MASKS = [
    (rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS, True),          # the usual set: the null-check-free instance
    (rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS, False),         # no counters
    (rub.OUT_LLR, True),                                      # soft output only
    (rub.OUT_BITS | rub.OUT_RXDATA, True),                    # hard output only
    (rub.OUT_EQ | rub.OUT_G, False),                          # equalised symbols and the channel estimate
]


@pytest.mark.parametrize("mod,det,flags", [(rub.MOD_QAM64, rub.DET_MMSE, rub.FLAG_MMSE_UNBIASED), (rub.MOD_QAM16, rub.DET_ZF, 0),
                                           (rub.MOD_QPSK, rub.DET_MMSE, 0)])
def test_ws_kernel_output_sets(mod, det, flags):
    """4x4 / 2048 runs k_rx_ws for QPSK..64-QAM; every output subset agrees with the oracle bit for bit."""
    cfg = rub.Config(M=2048, cp_len=152, num_streams=4, num_access_codes=2, num_data_symbols=5, modulation=mod,
                     detector=det, flags=flags)
    cfg, S1, iq, tx = make_case(cfg, 4, seed=0x2A + mod, n_taps=6, snr_db=27.0)
    _run_masks(cfg, S1, iq, tx, rub.PATH_FUSED, MASKS, kernel="k_rx_ws")


def test_256qam_at_4x4_2048_takes_the_staged_path():
    """Neither fused kernel holds 256-QAM's LLR staging beside the 4x4 / 2048 rings: forcing the fused path is
    refused, the automatic choice is the staged path (block-mapped detect kernel with TMA task records)."""
    import torch
    cfg = rub.Config(M=2048, cp_len=152, num_streams=4, num_access_codes=2, num_data_symbols=3,
                     modulation=rub.MOD_QAM256, detector=rub.DET_MMSE, flags=rub.FLAG_MMSE_UNBIASED)
    cfg, S1, iq, tx = make_case(cfg, 3, seed=0x256, n_taps=4, snr_db=36.0)
    _run_masks(cfg, S1, iq, tx, rub.PATH_AUTO, MASKS[:1], kernel="k_detect_lean")
    rx = rub.Receiver(cfg, S1)
    rx.set_path(rub.PATH_FUSED)
    with pytest.raises(rub.RubError):
        rx.process_batch(torch.from_numpy(iq).cuda(), out_mask=rub.OUT_EQ)
    rx.close()


@pytest.mark.parametrize("N,M,est", [(4, 256, rub.EST_LS_FULLBAND), (8, 512, rub.EST_LS_FULLBAND), (4, 1024, rub.EST_LS_COMB_INTERP),
                                     (8, 128, rub.EST_LS_COMB_INTERP), (2, 256, rub.EST_LS_COMB_INTERP)])
def test_staged_task_records_and_comb_fusion(N, M, est):
    """Staged path with every carrier occupied: k_detect_lean reads W/gain/isig as TMA task records for N >= 4
    (register prefetch for N = 2); the comb estimator runs k_lscomb_weights with and without the G export."""
    cfg = rub.Config(M=M, cp_len=M // 16 + 2, num_streams=N, num_access_codes=2, num_data_symbols=4,
                     modulation=rub.MOD_QAM64, detector=rub.DET_MMSE, flags=rub.FLAG_MMSE_UNBIASED, estimator=est)
    cfg, S1, iq, tx = make_case(cfg, 3, seed=N * 1000 + M, n_taps=3, snr_db=30.0)
    _run_masks(cfg, S1, iq, tx, rub.PATH_STAGED, MASKS, kernel="k_detect_lean")


def test_comb_pilot_spacing_wider_than_streams():
    """P = 16 > N = 4: tiles hold 128 / 16 + 2 pilots per antenna pair; band edges are held."""
    cfg = rub.Config(M=512, cp_len=36, num_streams=4, num_access_codes=3, num_data_symbols=3, modulation=rub.MOD_QAM16,
                     detector=rub.DET_ZF, estimator=rub.EST_LS_COMB_INTERP, pilot_spacing=16)
    cfg, S1, iq, tx = make_case(cfg, 2, seed=0x16, n_taps=2, snr_db=26.0)
    _run_masks(cfg, S1, iq, tx, rub.PATH_STAGED, [(rub.OUT_EQ | rub.OUT_G | rub.OUT_RXDATA, True), (rub.OUT_EQ, False)])


def test_capture_with_reused_pinned_buffers():
    """rub_rx_process_capture into caller-owned pinned buffers, twice: identical to a call with fresh buffers."""
    import torch
    from test_gpu_capture import _bursts
    cfg = rub.preset("C1", M=256, cp_len=20, num_access_codes=4, num_data_symbols=25)
    S0, S1, cap, tx, slices = _bursts(cfg, 3, seed=0x77, gaps=[0, 211])
    rx = rub.Receiver(cfg, S1)
    mask = rub.OUT_EQ | rub.OUT_RXDATA
    n0, sync0, out0 = rx.process_capture(cap, max_frames=6, out_mask=mask, tx_data=tx)
    assert n0 == 3
    cap_pin = torch.from_numpy(np.ascontiguousarray(cap, np.complex64)).pin_memory()
    buf = rx.alloc_outputs_host(6, mask, pinned=True)
    for _ in range(2):
        n1, sync1, out1 = rx.process_capture(cap_pin.numpy(), max_frames=6, out_mask=mask, tx_data=tx, out=buf)
        assert n1 == n0 and np.array_equal(sync1, sync0)
        assert np.array_equal(out1["eq"], out0["eq"]) and np.array_equal(out1["rx_data"], out0["rx_data"])
    rx.close()


@pytest.mark.parametrize("D,nac,cp,frames", [(1, 1, 152, 3), (2, 3, 152, 149), (3, 5, 0, 2), (14, 2, 152, 301), (6, 1, 64, 1)])
def test_ws_kernel_frame_and_symbol_count_edges(D, nac, cp, frames):
    """k_rx_ws schedules the training symbols of frame f+1 between the payload symbols of frame f from the third
    on: one or two payload symbols per frame, one to five access codes, zero cyclic prefix, one frame for 148
    CTAs, and batches that leave a ragged tail all have to agree with the oracle bit for bit."""
    import torch
    cfg = rub.Config(M=2048, cp_len=cp, num_streams=4, num_access_codes=nac, num_data_symbols=D,
                     modulation=rub.MOD_QAM64, detector=rub.DET_MMSE, flags=rub.FLAG_MMSE_UNBIASED)
    U = min(frames, 3)
    cfg, S1, iq_u, tx_u = make_case(cfg, U, seed=D * 100 + nac, n_taps=4 if cp else 1, snr_db=29.0)
    ref = oracle_run(cfg, S1, iq_u, tx_u)
    idx = np.arange(frames) % U
    rx = rub.Receiver(cfg, S1)
    rx.set_path(rub.PATH_FUSED)
    mask = rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS | rub.OUT_G
    for _ in range(2):   # twice: the second launch reuses the scratch the first one left behind
        out = rx.process_batch(torch.from_numpy(iq_u[idx]).cuda(), out_mask=mask, tx_data=torch.from_numpy(tx_u[idx]).cuda())
        rx.sync()
        assert rx.last_kernel() == "k_rx_ws"
        for k in ("eq", "llr", "bits", "G"):
            assert np.array_equal(out[k].cpu().numpy(), ref[k][idx]), k
    per = np.stack([oracle_run(cfg, S1, iq_u[i:i + 1], tx_u[i:i + 1])["counters"] for i in range(U)])
    assert np.array_equal(rx.read_counters(), 2 * per[idx].sum(axis=0))
    rx.close()


@pytest.mark.parametrize("M,cp", [(1024, 72), (2048, 152)])
def test_two_stream_fused_kernel_variants(M, cp, monkeypatch):
    """2x2 at M = 1024 / 2048 runs the monolithic fused kernel with W/gain/isig as TMA task records; the classic
    variant (register prefetch) stays selectable and both agree with the oracle bit for bit."""
    cfg = rub.Config(M=M, cp_len=cp, num_streams=2, num_access_codes=3, num_data_symbols=5, modulation=rub.MOD_QAM16,
                     detector=rub.DET_ZF)
    cfg, S1, iq, tx = make_case(cfg, 9, seed=M, n_taps=3, snr_db=24.0)
    for no_wtma in (False, True):
        if no_wtma:
            monkeypatch.setenv("RUB_FUSED_NO_WTMA", "1")
        else:
            monkeypatch.delenv("RUB_FUSED_NO_WTMA", raising=False)
        _run_masks(cfg, S1, iq, tx, rub.PATH_FUSED, MASKS[:3], kernel="k_rx_fused")


def _guarded(shape, dtype, guard_bytes=4096):
    """A tensor of `shape` carved out of a larger byte buffer with sentinel guard bands on both sides."""
    import torch
    n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    raw = torch.full((guard_bytes + n + guard_bytes,), 0xA5, dtype=torch.uint8, device="cuda")
    return raw, raw[guard_bytes:guard_bytes + n].view(dtype).view(*shape)


@pytest.mark.parametrize("kw,frames,path,kernel", [
    (dict(M=2048, cp_len=152, num_streams=4, num_access_codes=2, num_data_symbols=3, modulation=rub.MOD_QAM64,
          detector=rub.DET_MMSE, flags=rub.FLAG_MMSE_UNBIASED), 151, rub.PATH_FUSED, "k_rx_ws"),
    (dict(M=1024, cp_len=72, num_streams=2, num_access_codes=2, num_data_symbols=5, modulation=rub.MOD_QAM16,
          detector=rub.DET_ZF), 599, rub.PATH_FUSED, "k_rx_fused"),
    (dict(M=512, cp_len=36, num_streams=8, num_access_codes=2, num_data_symbols=3, modulation=rub.MOD_QAM256,
          detector=rub.DET_MMSE, estimator=rub.EST_LS_COMB_INTERP, flags=rub.FLAG_MMSE_UNBIASED), 5, rub.PATH_STAGED, "k_detect_lean"),
    (dict(M=256, cp_len=18, num_streams=4, num_access_codes=2, num_data_symbols=3, modulation=rub.MOD_QAM64,
          detector=rub.DET_MMSE), 7, rub.PATH_STAGED, "k_detect_lean"),
])
def test_outputs_stay_inside_their_buffers(kw, frames, path, kernel):
    """compute-sanitizer is not available on this pool: every output lives between two 4 KB guard bands filled with
    a sentinel; after a batch that leaves a ragged tail for the persistent kernels (151 / 599 frames on 148 SMs)
    the guards are untouched and the outputs equal those of an ordinary call."""
    import torch
    cfg = rub.Config(**kw)
    U = 3
    cfg, S1, iq_u, tx_u = make_case(cfg, U, seed=frames, n_taps=2, snr_db=28.0)
    idx = np.arange(frames) % U
    d_iq, d_tx = torch.from_numpy(iq_u[idx]).cuda(), torch.from_numpy(tx_u[idx]).cuda()
    mask = rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS | rub.OUT_RXDATA | rub.OUT_G
    rx = rub.Receiver(cfg, S1)
    rx.set_path(path)
    ref = rx.process_batch(d_iq, out_mask=mask, tx_data=d_tx)
    rx.sync()
    assert rx.last_kernel() == kernel
    raws, out = {}, {}
    for k, t in ref.items():
        raws[k], out[k] = _guarded(tuple(t.shape), t.dtype)
    rx.process_batch(d_iq, out=out, out_mask=mask, tx_data=d_tx)
    rx.sync()
    for k, t in ref.items():
        assert torch.equal(out[k], t), k
        raw = raws[k]
        assert bool((raw[:4096] == 0xA5).all()) and bool((raw[-4096:] == 0xA5).all()), f"{k}: guard band overwritten"
    rx.close()
