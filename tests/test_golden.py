"""Committed golden vectors (tests/golden/*.npz, written by oracle/make_golden.py): the oracle
must still reproduce them on any machine, and the CUDA path must reproduce them bit for bit."""
import glob
import json
import os

import numpy as np
import pytest

import rub_mimo_b200 as rub
from util import assert_parity, gpu_run, oracle_run

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "g_*.npz")))


def load(path):
    z = np.load(path)
    kw = json.loads(str(z["config"]))
    cfg = rub.Config(noise_var=float(z["noise_var"]), **kw)
    ref = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    return cfg, z["S1"], z["iq"], z["tx_data"], ref


def test_golden_set_is_present():
    assert len(GOLD) >= 4


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_golden(path):
    cfg, S1, iq, tx, ref = load(path)
    got = oracle_run(cfg, S1, iq, tx)
    assert_parity(ref, got, cfg.q)
    S1d, _ = rub.default_S1(cfg)
    assert np.array_equal(S1, S1d)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
@pytest.mark.parametrize("mode", ["device", "host"])
def test_gpu_reproduces_golden(path, mode):
    cfg, S1, iq, tx, ref = load(path)
    if mode == "device":
        got = gpu_run(cfg, S1, iq, tx)
    else:
        rx = rub.Receiver(cfg, S1)
        cnt = np.zeros((cfg.N, 4), np.uint64)
        mask = rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS | rub.OUT_RXDATA | rub.OUT_G
        got = rx.process_batch_host(iq, out_mask=mask, tx_data=tx, counters=cnt)
        got["counters"] = cnt
        rx.close()
    assert_parity(ref, got, cfg.q)
