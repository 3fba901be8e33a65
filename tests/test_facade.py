"""The framing.h-compatible C++ facade (include/rub_mimo/framing.h): compiled with g++ against
librubmimo_b200.so and driven the way mimo/main.cc drives the reference."""
import json
import os
import subprocess

import numpy as np
import pytest

import rub_mimo_b200 as rub
from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("facade")
    out = str(d / "facade_main")
    libdir = os.path.join(ROOT, "rub_mimo_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "facade_main.cpp"), "-o", out, "-L", libdir,
                           "-lrubmimo_b200", f"-Wl,-rpath,{libdir}"])
    return out


def run(exe, mode, d):
    r = subprocess.run([exe, mode, str(d)], capture_output=True, text=True)
    return r.returncode, r.stdout, r.stderr


def test_facade_compiles_and_transmit_side_runs_without_gpu(exe, tmp_path):
    rc, out, err = run(exe, "tx", tmp_path)
    assert rc == 0, err
    info = json.loads(out.strip().splitlines()[-1])
    assert info["occupied"] == 64
    cap = np.fromfile(tmp_path / "rx1.dat", dtype=np.complex64)
    assert cap.size == info["samples"]
    L, sync_len = 80, 41 * 80
    assert np.abs(cap[:sync_len]).max() < 0.05            # leading zeros: noise only
    assert np.abs(cap[sync_len:sync_len + L]).mean() > 0.1  # S0 burst on stream 0
    tx = np.fromfile(tmp_path / "tx_data1.dat", dtype=np.uint32)
    assert tx.size == 64 * 200 and tx.max() < 4


def test_facade_receive_fails_loudly_without_gpu(exe, tmp_path):
    if rub.device_count() > 0:
        pytest.skip("a CUDA device is present")
    assert run(exe, "tx", tmp_path)[0] == 0
    rc, out, err = run(exe, "rx", tmp_path)
    assert rc == 1 and "no CPU fallback" in err


@pytest.mark.gpu
def test_facade_receive_matches_oracle(exe, tmp_path):
    assert run(exe, "tx", tmp_path)[0] == 0
    rc, out, err = run(exe, "rx", tmp_path)
    assert rc == 0, err
    info = json.loads(out.strip().splitlines()[-1])
    assert info["state"] == 3 and info["valid_packets"] == 200
    assert info["valid_symbols"] == [info["symbols"]] * 2           # SER 0 at 30 dB
    # the oracle's faithful state machine on the same capture
    cfg = orc.Config(64, 16, 2, 20, 200, 2, flags=orc.FLAG_Q1)
    S0, _ = rub.default_S0(rub.preset("C1"))
    S1, _ = rub.default_S1(rub.preset("C1"))
    cap = [np.fromfile(tmp_path / f"rx{i}.dat", dtype=np.complex64) for i in (1, 2)]
    r = orc.framesync_execute(cfg, S0, S1, cap)
    assert r["rc"] == 0
    for s in range(2):
        eq = np.fromfile(tmp_path / f"rx_sig{s + 1}.dat", dtype=np.complex64)
        assert np.array_equal(eq, r["eq"][s].reshape(-1))            # bit-exact equalised symbols
        assert abs(info["plateau_start"][s] - int(r["plateau_start"][s])) <= 1
    assert np.isclose(info["G00"][0], r["G"][3, 0, 0].real, atol=1e-6)
    assert np.isclose(info["G01"][0], r["G"][3, 0, 1].real, atol=1e-6)
