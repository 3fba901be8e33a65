"""The N>1 path on real GPUs: two ranks, one GPU each, frame-sharded batches and the library's
ncclAllReduce of the error counters after EVERY step (as bench.py does).  The reduced counters must equal
the single-rank totals however many steps have run (round-1 ADVICE: an in-place reduction compounded them).
Needs two CUDA devices; skipped on a one-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total_frames, steps, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import rub_mimo_b200 as rub
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # plumbing only: the unique id travels over it
    cfg = rub.preset("C3", num_data_symbols=3)
    S1, s1 = rub.default_S1(cfg)
    b, e = rub.shard_range(total_frames, rank, world)
    iq, tx, nv = rub.synth_frames(cfg, e - b, 0xC5, n_taps=4, snr_db=22.0, first_frame=b, S1=S1, s1=s1, n_threads=2)
    cfg = cfg.with_noise_var(nv)
    rx = rub.Receiver(cfg, S1, device=rank)
    uid = [rub.comm_get_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    rx.comm_init(uid[0], rank, world)
    d_iq, d_tx = torch.from_numpy(iq).cuda(), torch.from_numpy(tx).cuda()
    snaps = []
    for _ in range(steps):
        rx.process_batch(d_iq, out_mask=rub.OUT_RXDATA, tx_data=d_tx)
        rx.allreduce_counters()
        snaps.append(rx.read_counters_global().copy())
    ret[f"global{rank}"] = np.stack(snaps)
    ret[f"local{rank}"] = rx.read_counters()
    rx.close()
    dist.destroy_process_group()


def test_nccl_counter_allreduce_over_steps():
    import torch
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import rub_mimo_b200 as rub
    from util import oracle_run
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    total, steps = 9, 3
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, total, steps, ret), nprocs=2, join=True)
        g0, g1 = np.array(ret["global0"]), np.array(ret["global1"])
        l0, l1 = np.array(ret["local0"]), np.array(ret["local1"])
    cfg = rub.preset("C3", num_data_symbols=3)
    S1, s1 = rub.default_S1(cfg)
    iq, tx, nv = rub.synth_frames(cfg, total, 0xC5, n_taps=4, snr_db=22.0, S1=S1, s1=s1)
    ref = oracle_run(cfg.with_noise_var(nv), S1, iq, tx)["counters"]
    assert np.array_equal(g0, g1)
    for k in range(steps):
        assert np.array_equal(g0[k], (k + 1) * ref), k
    assert np.array_equal(l0 + l1, steps * ref)      # the local counters were never overwritten by the reduction
