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
mask = rub.OUT_G | rub.OUT_EQ | rub.OUT_LLR | rub.OUT_BITS
rx = rub.Receiver(cfg, S1)
out = rx.alloc_outputs(F, mask)
M, L, cp, N, nac = cfg.M, cfg.L, cfg.cp_len, cfg.N, cfg.nac
def X(frame_u, r, sym):
    w = iq_u[frame_u, r, sym * L + cp: sym * L + cp + M].astype(np.complex128)
    return np.fft.fft(w)
s_ls = (1.0 / np.sqrt(cfg.Mo)) / nac
shown = 0
for it in range(10):
    rx.process_batch(d_iq, out=out, out_mask=mask, tx_data=d_tx); rx.sync()
    G = out["G"].cpu().numpy()
    good = G[:U]
    for f in range(F):
        d = G[f] != good[f % U]
        if not d.any(): continue
        rr, tt = np.nonzero(d.any(-1))
        for r_, t_ in zip(rr.tolist(), tt.tolist()):
            ks = np.nonzero(d[r_, t_])[0]
            k0 = ks[0]
            acc_bad, acc_good = G[f, r_, t_, ks] / s_ls, good[f % U, r_, t_, ks] / s_ls
            fu = f % U
            terms = [X(fu, r_, c * N + t_)[ks] * S1[t_, c, ks].real for c in range(nac)]
            model = terms[0] + terms[1]
            print(f"it {it} frame {f} (unique {fu}) local {f//148} link rx{r_} tx{t_} bad carriers {ks.size}: |good-model| {np.abs(acc_good-model).max():.2e}")
            diff = acc_bad - acc_good
            for name, cand in (("-term0", -terms[0]), ("-term1", -terms[1]), ("-2term0", -2*terms[0]), ("-2term1", -2*terms[1])):
                print(f"      diff vs {name}: {np.abs(diff-cand).max():.3e}", end="")
            print()
            # is bad = term_c(good) + something from another frame / antenna / symbol?
            for c in range(nac):
                rest = acc_bad - terms[1 - c]   # what replaced term c
                best = None
                for fu2 in range(U):
                    for r2 in range(N):
                        for sym2 in range(nac * N):
                            for sg in (1, -1):
                                e = np.abs(rest - sg * X(fu2, r2, sym2)[ks] * S1[t_, c, ks].real).max()
                                if best is None or e < best[0]: best = (e, fu2, r2, sym2, sg)
                print(f"      keeping code {1-c}: the other term best matches unique frame {best[1]} rx {best[2]} symbol {best[3]} sign {best[4]} (err {best[0]:.2e}); expected frame {fu} rx {r_} symbol {c*N+t_}")
            shown += 1
            if shown >= 4: sys.exit(0)
