"""The C-ABI library loads without a GPU, exports every symbol include/rub_mimo/rub_mimo.h
declares, validates configurations, and refuses to run the receive path without CUDA."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import rub_mimo_b200 as rub

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rub_mimo", "rub_mimo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rub_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = rub.lib()
    syms = header_symbols()
    assert len(syms) >= 45
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(rub.ABI_SYMBOLS) == syms
    assert L.rub_abi_version() == 1


def test_python_structs_match_c_layout():
    # rub_config: 11 x u32 + float + pointer; the library checks struct_size itself
    assert C.sizeof(rub.rub_config) == 56
    cfg = rub.preset("C3")
    cfg.validate()
    bad = rub.preset("C3")
    bad.c.struct_size = 12
    with pytest.raises(rub.RubError) as e:
        bad.validate()
    assert e.value.status == rub.ERR_INVALID_ARG


@pytest.mark.parametrize("kw,status", [
    (dict(M=100), rub.ERR_UNSUPPORTED),               # not a power of two
    (dict(M=8192), rub.ERR_UNSUPPORTED),
    (dict(M=64, cp_len=152), rub.ERR_INVALID_ARG),    # quirk Q7: CP_LENGTH 152 > M underflows framing.cc:184
    (dict(num_streams=9), rub.ERR_UNSUPPORTED),
    (dict(num_streams=0), rub.ERR_UNSUPPORTED),
    (dict(modulation=5), rub.ERR_UNSUPPORTED),        # ARITY 32 (ARB32OPT) has no table here
    (dict(num_access_codes=0), rub.ERR_INVALID_ARG),
    (dict(num_data_symbols=0), rub.ERR_INVALID_ARG),
    (dict(detector=7), rub.ERR_INVALID_ARG),
    (dict(estimator=rub.EST_LS_COMB_INTERP, num_streams=8, pilot_spacing=4), rub.ERR_UNSUPPORTED),
])
def test_config_validation_errors(kw, status):
    base = dict(M=1024, cp_len=72, num_streams=2, num_access_codes=2, num_data_symbols=4, modulation=4)
    base.update(kw)
    with pytest.raises(rub.RubError) as e:
        rub.Config(**base).validate()
    assert e.value.status == status
    assert rub.lib().rub_last_error()


def test_config_derived_quantities():
    c3 = rub.preset("C3")
    assert (c3.T, c3.Mo, c3.L, c3.row_bytes) == (8, 2048, 2200, 1536)
    c4 = rub.preset("C4")
    assert c4.T == 2 and c4.Mo == 4096
    p = rub.ofdmframe_init_default_sctype(2048, False, True)
    g = rub.Config(M=2048, cp_len=152, num_streams=2, num_access_codes=20, num_data_symbols=10, modulation=2, sctype=p)
    assert g.Mo == 1638 and g.row_bytes == (1638 * 2 + 7) // 8


def test_no_cpu_fallback_without_device():
    if rub.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(rub.RubError) as e:
        rub.Receiver(rub.preset("C2"))
    assert e.value.status == rub.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_shard_range_partitions_frames():
    for n, w in [(8192, 8), (1000, 3), (5, 8), (0, 4)]:
        spans = [rub.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1


def test_file_formats_roundtrip(tmp_path):
    """raw fc32 / uint32 files as mimo/main.cc:831-833, :1413-1419 and plot.py:27-40 use."""
    x = (np.arange(20) + 1j * np.arange(20)[::-1]).astype(np.complex64)
    path = str(tmp_path / "rx1.dat").encode()
    assert rub.lib().rub_file_write_fc32(path, x.ctypes.data_as(C.c_void_p), C.c_uint64(x.size)) == 0
    assert np.array_equal(np.fromfile(path.decode(), dtype=np.complex64), x)
    y = np.zeros(32, np.complex64)
    n = C.c_uint64()
    assert rub.lib().rub_file_read_fc32(path, y.ctypes.data_as(C.c_void_p), C.c_uint64(32), C.byref(n)) == 0
    assert n.value == 20 and np.array_equal(y[:20], x)
    d = np.arange(7, dtype=np.uint32)
    p2 = str(tmp_path / "rx_data1.dat").encode()
    assert rub.lib().rub_file_write_u32(p2, d.ctypes.data_as(C.c_void_p), C.c_uint64(7)) == 0
    assert np.array_equal(np.fromfile(p2.decode(), dtype=np.uint32), d)
    assert rub.lib().rub_file_read_fc32(b"/nonexistent/x.dat", y.ctypes.data_as(C.c_void_p), C.c_uint64(1), None) == rub.ERR_IO
