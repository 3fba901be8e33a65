"""N>1 host logic on CPU: world_size-2 gloo run of the frame sharding + counter reduction that
bench.py uses on GPUs (there the counters are reduced with ncclAllReduce inside the library).
The per-rank compute stand-in is the CPU oracle, which is allowed in tests."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total_frames, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import rub_mimo_b200 as rub
    from util import oracle_run
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = rub.preset("C2", M=256, cp_len=18, num_data_symbols=4)
    S1, s1 = rub.default_S1(cfg)
    b, e = rub.shard_range(total_frames, rank, world)
    # every rank regenerates only its own frames from the global frame index
    iq, tx, nv = rub.synth_frames(cfg, e - b, 0xC5, n_taps=2, snr_db=18.0, first_frame=b, S1=S1, s1=s1, n_threads=1)
    ref = oracle_run(cfg.with_noise_var(nv), S1, iq, tx, n_threads=1)
    t = torch.from_numpy(ref["counters"].astype(np.int64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)          # the one exchange step of the path
    if rank == 0:
        ret["counters"] = t.numpy().copy()
        ret["nv"] = nv
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    sys.path.insert(0, ROOT)
    import rub_mimo_b200 as rub
    from util import oracle_run
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    total = 7
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, total, ret), nprocs=2, join=True)
        got = np.array(ret["counters"])
    cfg = rub.preset("C2", M=256, cp_len=18, num_data_symbols=4)
    S1, s1 = rub.default_S1(cfg)
    iq, tx, nv = rub.synth_frames(cfg, total, 0xC5, n_taps=2, snr_db=18.0, S1=S1, s1=s1)
    ref = oracle_run(cfg.with_noise_var(nv), S1, iq, tx)
    assert np.array_equal(got, ref["counters"].astype(np.int64))
    assert got[:, 3].sum() == total * cfg.N * cfg.D * cfg.Mo
