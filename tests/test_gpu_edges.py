"""Edge cases of the C-ABI receive call on the GPU: empty and odd-sized batches, batches that do
not divide the persistent grid, padded strides, counter accumulation/reset, repeated calls."""
import ctypes as C

import numpy as np
import pytest

import rub_mimo_b200 as rub
from util import assert_parity, make_case, oracle_run

pytestmark = pytest.mark.gpu
MASK = rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS | rub.OUT_RXDATA | rub.OUT_G


def small_cfg(**over):
    kw = dict(M=512, cp_len=40, num_streams=2, num_access_codes=2, num_data_symbols=4,
              modulation=rub.MOD_QAM16, detector=rub.DET_MMSE, flags=rub.FLAG_MMSE_UNBIASED)
    kw.update(over)
    return rub.Config(**kw)


def test_empty_batch_is_a_noop():
    import torch
    cfg, S1, iq, tx = make_case(small_cfg(), 2, seed=1, n_taps=2, snr_db=20.0)
    rx = rub.Receiver(cfg, S1)
    d = torch.from_numpy(iq).cuda()
    out = rx.process_batch(d[:0], out_mask=MASK)
    rx.sync()
    assert out["eq"].shape[0] == 0 and rx.launch_count == 0
    assert rx.read_counters().sum() == 0
    rx.close()


@pytest.mark.parametrize("frames", [1, 149, 301])
@pytest.mark.parametrize("path", [rub.PATH_FUSED, rub.PATH_STAGED])
def test_batches_that_do_not_divide_the_grid(frames, path):
    """148 SMs: 1 frame (most CTAs idle), 149 (one CTA takes two), 301 (ragged tail)."""
    import torch
    cfg, S1, iq_u, tx_u = make_case(small_cfg(), 7, seed=frames, n_taps=2, snr_db=18.0)
    idx = np.arange(frames) % 7
    ref = oracle_run(cfg, S1, iq_u, tx_u)
    rx = rub.Receiver(cfg, S1)
    rx.set_path(path)
    out = rx.process_batch(torch.from_numpy(iq_u[idx]).cuda(), out_mask=MASK, tx_data=torch.from_numpy(tx_u[idx]).cuda())
    rx.sync()
    for k in ("eq", "llr", "bits", "rx_data", "G"):
        assert np.array_equal(out[k].cpu().numpy(), ref[k][idx]), k
    # counters: per-frame counts of the oracle, summed over the repeated frames
    per = np.stack([oracle_run(cfg, S1, iq_u[i:i + 1], tx_u[i:i + 1])["counters"] for i in range(7)])
    assert np.array_equal(rx.read_counters(), per[idx].sum(axis=0))
    rx.close()


@pytest.mark.parametrize("kw", [dict(num_data_symbols=1), dict(num_data_symbols=5, num_access_codes=3),
                                dict(num_streams=4, num_data_symbols=3, modulation=rub.MOD_QPSK),
                                dict(M=1024, cp_len=0, num_data_symbols=2)])
def test_odd_symbol_counts_and_zero_cp(kw):
    from util import gpu_run
    cfg, S1, iq, tx = make_case(small_cfg(**kw), 5, seed=17, n_taps=1, snr_db=15.0)
    ref = oracle_run(cfg, S1, iq, tx)
    for path in (rub.PATH_FUSED, rub.PATH_STAGED):
        assert_parity(ref, gpu_run(cfg, S1, iq, tx, path=path), cfg.q)


def test_padded_strides_and_first_sample():
    """rows longer than a frame (capture buffers): frame_stride / rx_stride / first_sample"""
    import torch
    cfg, S1, iq, tx = make_case(small_cfg(), 4, seed=5, n_taps=2, snr_db=20.0)
    ref = oracle_run(cfg, S1, iq, tx)
    F, N, row = iq.shape
    pad_front, pad_back = 10, 6                       # even: the fused kernel stays eligible
    big = np.zeros((F, N, pad_front + row + pad_back), np.complex64)
    big[:, :, pad_front:pad_front + row] = iq
    rx = rub.Receiver(cfg, S1)
    out = rx.process_batch(torch.from_numpy(big).cuda(), out_mask=MASK, tx_data=torch.from_numpy(tx).cuda(),
                           first_sample=pad_front)
    rx.sync()
    assert rx.last_path == rub.PATH_FUSED
    got = {k: v.cpu().numpy() for k, v in out.items()}
    got["counters"] = rx.read_counters()
    assert_parity(ref, got, cfg.q)
    rx.close()


def test_counters_accumulate_and_reset():
    import torch
    cfg, S1, iq, tx = make_case(small_cfg(), 3, seed=9, n_taps=2, snr_db=10.0)
    ref = oracle_run(cfg, S1, iq, tx)["counters"]
    rx = rub.Receiver(cfg, S1)
    d_iq, d_tx = torch.from_numpy(iq).cuda(), torch.from_numpy(tx).cuda()
    for _ in range(3):
        rx.process_batch(d_iq, out_mask=rub.OUT_RXDATA, tx_data=d_tx)
    assert np.array_equal(rx.read_counters(), 3 * ref)
    rx.reset_counters()
    rx.process_batch(d_iq, out_mask=rub.OUT_RXDATA, tx_data=d_tx)
    assert np.array_equal(rx.read_counters(), ref)
    # user-supplied device counters
    mine = torch.zeros((cfg.N, 4), dtype=torch.int64, device="cuda")
    io_out = rx.alloc_outputs(3, rub.OUT_RXDATA)
    io = rub.rub_rx_io()
    io.iq = d_iq.data_ptr(); io.layout = rub.rub_iq_layout(d_iq.shape[1] * d_iq.shape[2], d_iq.shape[2], 0)
    io.tx_data = d_tx.data_ptr(); io.rx_data = io_out["rx_data"].data_ptr(); io.counters = mine.data_ptr()
    io.out_mask = rub.OUT_RXDATA
    rx.tstream.wait_stream(torch.cuda.current_stream())
    assert rub.lib().rub_rx_process_batch(rx.h, C.byref(io), 3) == 0
    rx.sync()
    assert np.array_equal(mine.cpu().numpy().astype(np.uint64), ref)
    assert np.array_equal(rx.read_counters(), ref)          # the handle's own counters were not touched
    rx.close()


def test_null_arguments_are_rejected():
    cfg = small_cfg()
    rx = rub.Receiver(cfg)
    io = rub.rub_rx_io()
    assert rub.lib().rub_rx_process_batch(rx.h, C.byref(io), 1) == rub.ERR_INVALID_ARG
    assert rub.lib().rub_rx_process_batch(None, C.byref(io), 1) == rub.ERR_INVALID_ARG
    assert rub.lib().rub_rx_process_batch_host(rx.h, C.byref(io), 1) == rub.ERR_INVALID_ARG
    assert rub.lib().rub_rx_set_path(rx.h, 9) == rub.ERR_INVALID_ARG
    rx.close()
    h = C.c_void_p()
    S1 = np.ones((cfg.N, cfg.nac, cfg.M), np.complex64) * 0.5   # not BPSK
    st = rub.lib().rub_rx_create(C.byref(h), C.byref(cfg.c), S1.ctypes.data_as(C.c_void_p), -1, None)
    assert st == rub.ERR_UNSUPPORTED


def test_handle_lifecycle_does_not_leak_device_memory():
    """create -> fused + staged + sync + framegen calls -> destroy, 30 times: free device memory is stable."""
    import torch
    cfg = rub.Config(M=512, cp_len=36, num_streams=2, num_access_codes=2, num_data_symbols=4,
                     modulation=rub.MOD_QAM16, detector=rub.DET_ZF)
    cfg, S1, iq, tx = make_case(cfg, 8, seed=21, n_taps=2, snr_db=25.0)
    d_iq, d_tx = torch.from_numpy(iq).cuda(), torch.from_numpy(tx).cuda()

    def cycle():
        rx = rub.Receiver(cfg, S1)
        for path in (rub.PATH_FUSED, rub.PATH_STAGED):
            rx.set_path(path)
            rx.process_batch(d_iq, out_mask=rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS, tx_data=d_tx)
        rx.sync()
        rx.sc_metric(iq[0, 0])
        rx.timing_search(np.ascontiguousarray(iq[0]))
        rx.framegen_batch(d_tx)
        torch.cuda.synchronize()
        rx.close()

    cycle()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(30):
        cycle()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < (8 << 20), f"leaked {(free0 - free1) >> 20} MiB over 30 handle lifecycles"


@pytest.mark.parametrize("chunk", [1, 2, 3])
def test_host_pipeline_with_many_chunks(chunk):
    """rub_rx_process_batch_host with 3+ pipeline chunks (slot reuse: the waits on the previous compute /
    copy-out of a slot) returns what the oracle computes, and the host counters are the cumulative ones."""
    cfg, S1, iq, tx = make_case(small_cfg(), 7, seed=21, n_taps=2, snr_db=14.0)
    ref = oracle_run(cfg, S1, iq, tx)
    rx = rub.Receiver(cfg, S1)
    rx.set_host_chunk(chunk)
    cnt = np.zeros((cfg.N, 4), np.uint64)
    out = rx.process_batch_host(iq, out_mask=MASK, tx_data=tx, counters=cnt)
    got = {k: v for k, v in out.items() if not k.startswith("_")}
    got["counters"] = cnt
    assert_parity(ref, got, cfg.q)
    rx.process_batch_host(iq, out_mask=MASK, tx_data=tx, counters=cnt)     # cumulative on the second call
    assert np.array_equal(cnt, 2 * ref["counters"])
    rx.close()


def test_allreduce_is_idempotent_on_one_rank():
    """rub_allreduce_counters never modifies the local counters: after any number of calls the global
    counters of a single-rank job equal the local cumulative ones (the in-place reduction
    compounded them)."""
    import torch
    cfg, S1, iq, tx = make_case(small_cfg(), 3, seed=9, n_taps=2, snr_db=10.0)
    ref = oracle_run(cfg, S1, iq, tx)["counters"]
    rx = rub.Receiver(cfg, S1)
    d_iq, d_tx = torch.from_numpy(iq).cuda(), torch.from_numpy(tx).cuda()
    for step in range(1, 5):
        rx.process_batch(d_iq, out_mask=rub.OUT_RXDATA, tx_data=d_tx)
        rx.allreduce_counters()
        rx.allreduce_counters()
        assert np.array_equal(rx.read_counters_global(), step * ref)
        assert np.array_equal(rx.read_counters(), step * ref)
    rx.close()
