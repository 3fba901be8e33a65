"""Row f3 of SURVEY.md 8: the offline IQ-file driver (mimo/main.cc:906-918 source, :1413-1419 sinks)
streams per-antenna captures through pinned double-buffered slots and reproduces the in-memory
batch call bit for bit."""
import numpy as np
import pytest

import rub_mimo_b200 as rub
from util import make_case, oracle_run

pytestmark = pytest.mark.gpu


def _write_capture(tmp_path, cfg, iq, tx, lead, gap):
    """One fc32 file per rx antenna: `lead` junk samples, then the frames `gap` samples apart."""
    F, N, row = iq.shape
    rng = np.random.default_rng(5)
    rx_paths, tx_paths = [], []
    for r in range(N):
        parts = [(rng.standard_normal(lead) + 1j * rng.standard_normal(lead)).astype(np.complex64)]
        for f in range(F):
            parts.append(iq[f, r])
            parts.append(np.zeros(gap, np.complex64))
        p = tmp_path / f"rx{r}.dat"
        np.concatenate(parts).tofile(p)
        rx_paths.append(p)
        t = tmp_path / f"tx_data{r}.dat"
        tx[:, r].astype(np.uint32).tofile(t)
        tx_paths.append(t)
    return rx_paths, tx_paths


@pytest.mark.parametrize("chunk", [0, 2, 5])
def test_file_driver_matches_batch_call(tmp_path, chunk):
    import torch
    cfg = rub.Config(M=256, cp_len=18, num_streams=2, num_access_codes=2, num_data_symbols=6,
                     modulation=rub.MOD_QAM16, detector=rub.DET_MMSE, noise_var=1e-3)
    cfg, S1, iq, tx = make_case(cfg, 7, seed=0xF3, n_taps=3, snr_db=24.0)
    lead, gap = 37, 11
    rx_paths, tx_paths = _write_capture(tmp_path, cfg, iq, tx, lead, gap)
    eq_paths = [tmp_path / f"rx_sig{s}.dat" for s in range(cfg.N)]
    rd_paths = [tmp_path / f"rx_data{s}.dat" for s in range(cfg.N)]
    rx = rub.Receiver(cfg, S1)
    done = rx.process_files(rx_paths, 7, first_sample=lead, frame_stride=iq.shape[2] + gap, tx_data_paths=tx_paths,
                            eq_paths=eq_paths, rx_data_paths=rd_paths, llr_path=tmp_path / "llr.dat",
                            bits_path=tmp_path / "bits.dat", chunk_frames=chunk)
    assert done == 7
    c_files = rx.read_counters()

    ref_rx = rub.Receiver(cfg, S1)
    out = ref_rx.process_batch(torch.from_numpy(iq).cuda(), tx_data=torch.from_numpy(tx).cuda(),
                               out_mask=rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS | rub.OUT_RXDATA)
    ref_rx.sync()
    eq = out["eq"].cpu().numpy()           # [F][N][D][Mo]
    rd = out["rx_data"].cpu().numpy()
    for s in range(cfg.N):
        assert np.array_equal(np.fromfile(eq_paths[s], np.complex64), eq[:, s].reshape(-1))
        assert np.array_equal(np.fromfile(rd_paths[s], np.uint32), rd[:, s].reshape(-1).astype(np.uint32))
    assert np.array_equal(np.fromfile(tmp_path / "llr.dat", np.float32), out["llr"].cpu().numpy().reshape(-1))
    assert np.array_equal(np.fromfile(tmp_path / "bits.dat", np.uint8), out["bits"].cpu().numpy().reshape(-1))
    assert np.array_equal(c_files, ref_rx.read_counters())
    # ... and the files hold what the CPU oracle computes from the same frames
    ref = oracle_run(cfg, S1, iq, tx)
    for s in range(cfg.N):
        assert np.array_equal(np.fromfile(eq_paths[s], np.complex64), ref["eq"][:, s].reshape(-1))
        assert np.array_equal(np.fromfile(rd_paths[s], np.uint32), ref["rx_data"][:, s].reshape(-1).astype(np.uint32))
    assert np.array_equal(np.fromfile(tmp_path / "llr.dat", np.float32), ref["llr"].reshape(-1))
    assert np.array_equal(np.fromfile(tmp_path / "bits.dat", np.uint8), ref["bits"].reshape(-1))
    assert np.array_equal(c_files, ref["counters"])


def test_short_capture_stops_at_last_complete_frame(tmp_path):
    cfg = rub.Config(M=64, cp_len=16, num_streams=2, num_access_codes=2, num_data_symbols=3, modulation=rub.MOD_QPSK)
    cfg, S1, iq, tx = make_case(cfg, 4, seed=9, n_taps=0, snr_db=30.0, fixed_H_ri=[[1, 0], [0.2, 0.1], [0, 0.3], [1, 0]])
    rx_paths, _ = _write_capture(tmp_path, cfg, iq, tx, 0, 0)
    # truncate antenna 1 in the middle of frame 2
    data = np.fromfile(rx_paths[1], np.complex64)
    data[: 2 * iq.shape[2] + 100].tofile(rx_paths[1])
    rd_paths = [tmp_path / f"rx_data{s}.dat" for s in range(cfg.N)]
    rx = rub.Receiver(cfg, S1)
    assert rx.process_files(rx_paths, 10, rx_data_paths=rd_paths, chunk_frames=3) == 2
    assert np.fromfile(rd_paths[0], np.uint32).size == 2 * cfg.D * cfg.Mo


def test_missing_file_is_an_io_error(tmp_path):
    cfg = rub.Config(M=64, cp_len=16, num_streams=2, num_access_codes=2, num_data_symbols=3, modulation=rub.MOD_QPSK)
    rx = rub.Receiver(cfg)
    with pytest.raises(rub.RubError):
        rx.process_files([tmp_path / "nope0.dat", tmp_path / "nope1.dat"], 1)
