"""Independent float64 model of the receive chain (numpy: np.fft + np.linalg).  TEST
INFRASTRUCTURE ONLY — shares no code with either the C oracle or the CUDA kernels.

It restates the same mathematics the reference implements (paths relative to /root/reference):
  * payload symbol: drop cp samples, unnormalised forward DFT, scale 1/sqrt(Mo)
    (mimo/framing.cc:557-566),
  * LS estimate G = (I*Q1 + sum_c X_c / S1_c) / (nac*sqrt(Mo))   (mimo/framing.cc:801-824),
  * ZF  W = G^-1 (the reference's adjugate form, mimo/framing.cc:1344-1367, is the same matrix),
    MMSE W = (G^H G + nv I)^-1 G^H, optional unbiasing by 1/mu_s,
  * liquid square-QAM nearest-point demodulation (mimo/main.cc:1405) and brute-force max-log LLRs.
Used to bound the fp32 mirror oracle / CUDA results at 1e-4 and to list near-threshold symbols.
"""
import numpy as np

FLAG_Q1, FLAG_UNBIASED = 1, 2
DET_ZF, DET_MMSE = 0, 1
EST_FULLBAND, EST_COMB = 0, 1


def qam_alpha(q):
    # liquid assigns alpha to a float
    return float(np.float32(1.0 / np.sqrt({2: 2.0, 4: 10.0, 6: 42.0, 8: 170.0}[q])))


def gray_decode(g):
    s = g
    sh = 1
    while sh < 32:
        s ^= s >> sh
        sh <<= 1
    return s


def constellation(q):
    """symbol index -> complex point (liquid modem_modulate_qam)."""
    m = q // 2
    P = 1 << m
    a = qam_alpha(q)
    pts = np.zeros(1 << q, np.complex128)
    for sym in range(1 << q):
        si, sq = gray_decode(sym >> m), gray_decode(sym & (P - 1))
        pts[sym] = (2 * si - P + 1) * a + 1j * (2 * sq - P + 1) * a
    return pts


def rx_frame(rows, S1, M, cp, N, nac, D, q, detector=DET_ZF, estimator=EST_FULLBAND, P=8, flags=0,
             noise_var=0.0, sctype=None, first_sample=0):
    """rows [N][samples] complex; S1 [N][nac][M].  Returns dict(G, W, eq, sym, llr, margin)."""
    rows = np.asarray(rows, np.complex128)
    S1 = np.asarray(S1, np.complex128)
    L = M + cp
    occ = np.arange(M) if sctype is None else np.nonzero(np.asarray(sctype) != 0)[0]
    Mo = len(occ)
    T = nac if estimator == EST_COMB else nac * N
    dn = 1.0 / np.sqrt(Mo)
    G = np.zeros((N, N, M), np.complex128)
    if flags & FLAG_Q1:
        for r in range(N):
            G[r, r, occ] = 1.0
    if estimator == EST_FULLBAND:
        for c in range(nac):
            for t in range(N):
                ac = c * N + t
                st = first_sample + ac * L + cp
                for r in range(N):
                    X = np.fft.fft(rows[r, st:st + M])
                    G[r, t, occ] += X[occ] / S1[t, c, occ]
        G *= dn / nac
    else:
        Gp = np.zeros((N, N, M), np.complex128)
        for c in range(nac):
            st = first_sample + c * L + cp
            for r in range(N):
                X = np.fft.fft(rows[r, st:st + M])
                for t in range(N):
                    k = np.arange(t, M, P)
                    Gp[r, t, k] += X[k] / S1[t, c, k]
        for r in range(N):
            for t in range(N):
                k = np.arange(t, M, P)
                g = (Gp[r, t, k] + (1.0 if (flags & FLAG_Q1) and r == t else 0.0)) * dn / nac
                G[r, t] = np.interp(np.arange(M), k, g.real) + 1j * np.interp(np.arange(M), k, g.imag)
    W = np.zeros((N, N, M), np.complex128)
    gain = np.ones((N, M))
    isig = np.ones((N, M))
    mmse = detector == DET_MMSE and noise_var > 0
    for k in occ:
        Gk = G[:, :, k]
        A = Gk.conj().T @ Gk + (noise_var * np.eye(N) if mmse else 0)
        Ai = np.linalg.inv(A)
        Wk = Ai @ Gk.conj().T
        if noise_var > 0:
            e = noise_var * np.real(np.diag(Ai))
            if mmse and (flags & FLAG_UNBIASED):
                mu = 1 - e
                Wk = Wk / mu[:, None]
                isig[:, k] = mu / e
            else:
                isig[:, k] = 1 / e
        W[:, :, k] = Wk
    pts = constellation(q)
    bits = ((np.arange(1 << q)[:, None] >> (q - 1 - np.arange(q))[None, :]) & 1).astype(bool)
    eq = np.zeros((N, D, Mo), np.complex128)
    sym = np.zeros((N, D, Mo), np.int64)
    llr = np.zeros((N, D, Mo, q))
    margin = np.zeros((N, D, Mo))  # distance gap between best and second-best point
    pay0 = first_sample + T * L
    for d in range(D):
        Y = np.stack([np.fft.fft(rows[r, pay0 + d * L + cp: pay0 + d * L + cp + M]) for r in range(N)]) * dn
        z = np.einsum("srk,rk->sk", W[:, :, occ], Y[:, occ])
        eq[:, d] = z
        dist = np.abs(z[..., None] - pts[None, None, :]) ** 2  # [N][Mo][Q]
        order = np.sort(dist, axis=-1)
        sym[:, d] = np.argmin(dist, axis=-1)
        margin[:, d] = order[..., 1] - order[..., 0]
        for b in range(q):
            d1 = np.min(np.where(bits[:, b][None, None, :], dist, np.inf), axis=-1)
            d0 = np.min(np.where(~bits[:, b][None, None, :], dist, np.inf), axis=-1)
            llr[:, d, :, b] = (d1 - d0) * isig[:, occ]
    return dict(G=G, W=W, eq=eq, sym=sym, llr=llr, margin=margin, isig=isig)
