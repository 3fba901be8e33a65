"""Development aid: run the fused kernel on the tiled C3 batch several times and report where repetitions of
the same unique frame differ (frame, stream, symbol, carrier range)."""
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rub_mimo_b200 as rub
cfg = rub.preset("C3")
syn = dict(rub.PRESET_SYNTH["C3"]); seed = syn.pop("seed")
S1, s1 = rub.default_S1(cfg)
U, F = 32, 1024
iq_u, tx_u, nv = rub.synth_frames(cfg, U, seed, S1=S1, s1=s1, **syn)
cfg = cfg.with_noise_var(nv)
d_iq = torch.from_numpy(iq_u).cuda().repeat(F // U, 1, 1)
d_tx = torch.from_numpy(tx_u).cuda().repeat(F // U, 1, 1, 1)
mask = rub.OUT_EQ | rub.OUT_G
rx = rub.Receiver(cfg, S1)
out = rx.alloc_outputs(F, mask)
nbad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    rx.process_batch(d_iq, out=out, out_mask=mask, tx_data=d_tx)
    rx.sync()
    for k in ("G", "eq"):
        v = out[k].reshape(F // U, U, *out[k].shape[1:])
        if k == "eq": v = torch.view_as_real(v) if v.is_complex() else v
        ne = (v != v[0:1])
        if bool(ne.any()):
            idx = ne.nonzero()
            nbad += 1
            print(f"iter {it} {k}: {idx.shape[0]} differing values; first {idx[0].tolist()} last {idx[-1].tolist()}")
            rep, fr = idx[:, 0], idx[:, 1]
            frames = torch.unique(rep * U + fr)
            print("   frames:", frames[:12].tolist(), "count", frames.numel(), "-> CTA", [int(f) % 148 for f in frames[:12]], "local frame idx", [int(f) // 148 for f in frames[:12]])
            if k == "G":
                v2 = out["G"].reshape(F // U, U, *out["G"].shape[1:])
                ne2 = (torch.view_as_real(v2) != torch.view_as_real(v2[0:1])).any(-1)   # [rep][U][rx][tx][k]
                ii = ne2.nonzero()
                links = torch.unique(ii[:, 2] * 4 + ii[:, 3]).tolist()
                ks = torch.unique(ii[:, 4])
                print("   G links (rx*4+tx):", links, "carriers", ks.numel(), "min", int(ks.min()), "max", int(ks.max()), "k%128 set", sorted(set((ks % 128).tolist()))[:20], "k//256 set", sorted(set((ks // 256).tolist())))
            elif idx.shape[1] >= 5:
                print("   streams", torch.unique(idx[:, 2]).tolist(), "symbols", torch.unique(idx[:, 3]).tolist(), "carriers", int(idx[:, 4].min()), "..", int(idx[:, 4].max()))
print("iterations with differences:", nbad)
