"""numpy transcription of the per-face algorithm in mst-cfd_b200/csrc/physics.cuh
(roe_contract / ausm_contract).  Lets the CPU-only suite check the closed-form
algebra of the CUDA path against the oracle's literal K |L| K^-1 restatement
without a GPU.  Test helper, not product code."""
import numpy as np


def _prim(q, gm1):
    D = q.shape[-1] - 2
    r = 1.0 / q[..., 0]
    m2 = (q[..., 1:1 + D] ** 2).sum(-1)
    p = (q[..., -1] - 0.5 * m2 * r) * gm1
    ht = (q[..., -1] + p) * r
    u = q[..., 1:1 + D] * r[..., None]
    return r, p, ht, u


def _efix(x, delta):
    return np.where(x > delta, x, (x * x + delta * delta) / (2 * delta))


def roe_contract(A, B, flags, Sd, gamma=1.4, delta=0.125, eor=1e-10):
    D = A.shape[-1] - 2
    U = D + 2
    gm1 = gamma - 1
    ra, pa, hta, ua = _prim(A, gm1)
    rb, pb, htb, ub = _prim(B, gm1)
    w = np.sqrt(np.abs(B[..., 0] * ra))
    iw = 1 / (1 + w)
    uh = (ua + w[..., None] * ub) * iw[..., None]
    q2 = (uh ** 2).sum(-1)
    H = (hta + w * htb) * iw
    g = H - 0.5 * q2
    ah = np.sqrt(np.abs(gm1 * g))
    dq = B - A
    ud = (uh * dq[..., 1:1 + D]).sum(-1)
    theta = (0.5 * q2 * dq[..., 0] - ud + dq[..., -1]) / g
    sh = dq[..., 1:1 + D] - uh * dq[..., :1]
    # central part 1/2 (F_A + F_B) . S through the mass fluxes m.S (physics.cuh: contracted first)
    mna = (Sd * A[..., 1:1 + D]).sum(-1)
    mnb = (Sd * B[..., 1:1 + D]).sum(-1)
    fa = mna / (A[..., 0] + eor)
    fb = mnb / (B[..., 0] + eor)
    phi = np.zeros_like(A)
    phi[..., 0] = 0.5 * (mna + mnb)
    for i in range(D):
        phi[..., i + 1] = 0.5 * (A[..., i + 1] * fa + B[..., i + 1] * fb) + 0.5 * (pa + pb) * Sd[..., i]
    phi[..., -1] = 0.5 * (hta * mna + htb * mnb)
    for d in range(D):
        ss = np.where(flags[..., d] != 0, 0.5, -0.5) * Sd[..., d]
        beta = sh[..., d] / ah
        lm = _efix(np.abs(uh[..., d] - ah), delta)
        le = _efix(np.abs(uh[..., d]), delta)
        lp = _efix(np.abs(uh[..., d] + ah), delta)
        wm = lm * 0.5 * (theta - beta)
        we = le * (dq[..., 0] - theta)
        wp = lp * 0.5 * (theta + beta)
        s = wm + we + wp
        dif = ah * (wp - wm)
        en = H * (wm + wp) + uh[..., d] * dif + 0.5 * q2 * we
        phi[..., 0] -= ss * s
        for i in range(D):
            dis = uh[..., i] * s
            if i == d:
                dis = dis + dif
            else:
                ws = le * sh[..., i]
                dis = dis + ws
                en = en + uh[..., i] * ws
            phi[..., i + 1] -= ss * dis
        phi[..., -1] -= ss * en
    return phi


def ausm_contract(A, B, flags, Sd, gamma=1.4):
    D = A.shape[-1] - 2
    gm1 = gamma - 1
    fac = 2 * gm1 / (gamma + 1)
    ra, pa, hta, ua = _prim(A, gm1)
    rb, pb, htb, ub = _prim(B, gm1)
    asa, asb = np.sqrt(hta * fac), np.sqrt(htb * fac)
    Ua = np.sqrt((A[..., 1:1 + D] ** 2).sum(-1) * ra * ra)
    Ub = np.sqrt((B[..., 1:1 + D] ** 2).sum(-1) * rb * rb)
    ata = asa * asa / np.maximum(asa, Ua)
    atb = asb * asb / np.maximum(asb, Ub)
    iaf = 1 / np.minimum(ata, atb)
    ca, cb = np.sqrt(gamma * pa * ra), np.sqrt(gamma * pb * rb)
    PA, PB = A.copy(), B.copy()
    PA[..., -1] += pa
    PB[..., -1] += pb
    ya, yb = ca[..., None] * PA, cb[..., None] * PB
    ysum, ydif = ya + yb, yb - ya
    phi = np.zeros_like(A)
    for d in range(D):
        fl = flags[..., d] != 0
        ML = np.where(fl, ua[..., d], ub[..., d]) * iaf
        MR = np.where(fl, ub[..., d], ua[..., d]) * iaf
        pL, pR = np.where(fl, pa, pb), np.where(fl, pb, pa)
        t, s = ML * ML - 1, ML + 1
        with np.errstate(divide="ignore", invalid="ignore"):
            Mp = np.where(ML <= 1, 0.25 * s * s + 0.125 * t * t, 0.5 * (ML + np.abs(ML)))
            Pp = np.where(ML <= 1, pL * 0.25 * s * s * (2 - ML) + 0.1875 * ML * t * t,
                          pL * 0.5 * (ML + np.abs(ML)) / ML)
            t, s = MR * MR - 1, MR - 1
            Mm = np.where(MR <= 1, -0.25 * s * s - 0.125 * t * t, 0.5 * (MR - np.abs(MR)))
            Pm = np.where(MR <= 1, pR * 0.25 * s * s * (2 + MR) - 0.1875 * MR * t * t,
                          pR * 0.5 * (MR - np.abs(MR)) / MR)
        Mf, pf = Mm + Mp, Pm + Pp
        aM = np.where(fl, np.abs(Mf), -np.abs(Mf))
        F = 0.5 * (Mf[..., None] * ysum - aM[..., None] * ydif)
        F[..., d + 1] += pf
        phi += Sd[..., d:d + 1] * F
    return phi
