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
mask = rub.OUT_G | rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS
rx = rub.Receiver(cfg, S1)
out = rx.alloc_outputs(F, mask)
d_tx = torch.from_numpy(tx_u).cuda().repeat(F // U, 1, 1, 1)
for it in range(8):
  rx.process_batch(d_iq, out=out, out_mask=mask, tx_data=d_tx); rx.sync()
  G = out["G"].cpu().numpy()            # [F][rx][tx][k]
  good = G[:U]
  for f in range(F):
      d = G[f] != good[f % U]
      if not d.any(): continue
      rr, tt = np.nonzero(d.any(-1))
      for r_, t_ in zip(rr.tolist(), tt.tolist()):
        ks = np.nonzero(d[r_, t_])[0]
        prev, nxt = f - 148, f + 148
        sp = (G[f, r_, t_, ks] == good[prev % U, r_, t_, ks]).mean() if prev >= 0 else -1
        sn = (G[f, r_, t_, ks] == good[nxt % U, r_, t_, ks]).mean() if nxt < F else -1
        runs = np.split(ks, np.nonzero(np.diff(ks) != 1)[0] + 1)
        print(f"it {it} frame {f} cta {f%148} local {f//148} link rx{r_} tx{t_}: bad carriers {ks.size} runs {[(int(r[0]), int(r.size)) for r in runs[:6]]} == prev frame's value {sp:.2f} == next frame's value {sn:.2f}")
        k0 = ks[0]
        print("     k", k0, "bad", G[f, r_, t_, k0], "good", good[f % U, r_, t_, k0], "prev", good[prev % U, r_, t_, k0] if prev >= 0 else None, "next", good[nxt % U, r_, t_, k0] if nxt < F else None)
