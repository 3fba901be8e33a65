"""Run-to-run determinism of the fused C3 path at full size: 1024 frames tiled from 32 unique ones, every output
of every tile compared with the first tile's, repeated ITER times (a cross-proxy race shows up as a handful of
frames that differ from their twins, differently in every run)."""
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rub_mimo_b200 as rub
ITER = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cfg = rub.preset("C3")
syn = dict(rub.PRESET_SYNTH["C3"]); seed = syn.pop("seed")
S1, s1 = rub.default_S1(cfg)
U, F = 32, 1024
iq_u, tx_u, nv = rub.synth_frames(cfg, U, seed, S1=S1, s1=s1, **syn)
cfg = cfg.with_noise_var(nv)
d_iq = torch.from_numpy(iq_u).cuda().repeat(F // U, 1, 1)
d_tx = torch.from_numpy(tx_u).cuda().repeat(F // U, 1, 1, 1)
mask = rub.OUT_G | rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS
rx = rub.Receiver(cfg, S1)
out = rx.alloc_outputs(F, mask)
bad = 0
for it in range(ITER):
    rx.process_batch(d_iq, out=out, out_mask=mask, tx_data=d_tx); rx.sync()
    for k in ("G", "eq", "llr", "bits"):
        t = out[k]
        v = t.view(torch.uint8) if t.dtype != torch.uint8 else t
        v = v.reshape(F // U, U, -1)
        n = int((v != v[:1]).any(dim=2).sum().item())
        if n:
            bad += n
            print(f"iteration {it}: {n} frames differ from their twins in {k}")
print(f"{rx.last_kernel()}: {ITER} iterations x {F} frames, frames that differ from their twins: {bad}")
sys.exit(1 if bad else 0)
