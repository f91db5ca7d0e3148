"""ORACLE / TEST INFRASTRUCTURE -- exact solution of the Sod shock tube
(external known answer used to pin the oracle).

Initial data of the reference (R/time/Time.cpp:26-31 with CONST.h:70-75):
left  (x <= 0.5): rho = 1,     u = 0, p = 1
right (x >  0.5): rho = 0.125, u = 0, p = 0.1          gamma = 1.4
"""
import numpy as np


def sod_exact(x, t, gamma=1.4, x0=0.5, left=(1.0, 0.0, 1.0), right=(0.125, 0.0, 0.1)):
    rl, ul, pl = left
    rr, ur, pr = right
    g = gamma
    al, ar = np.sqrt(g * pl / rl), np.sqrt(g * pr / rr)

    def f(p, rk, pk, ak):
        if p > pk:  # shock
            A, B = 2 / ((g + 1) * rk), (g - 1) / (g + 1) * pk
            return (p - pk) * np.sqrt(A / (p + B))
        return 2 * ak / (g - 1) * ((p / pk) ** ((g - 1) / (2 * g)) - 1)

    lo, hi = 1e-8, 10.0
    for _ in range(200):  # bisection on the pressure function
        mid = 0.5 * (lo + hi)
        if f(mid, rl, pl, al) + f(mid, rr, pr, ar) + ur - ul > 0:
            hi = mid
        else:
            lo = mid
    ps = 0.5 * (lo + hi)
    us = 0.5 * (ul + ur) + 0.5 * (f(ps, rr, pr, ar) - f(ps, rl, pl, al))
    # left rarefaction, right shock (Sod)
    rsl = rl * (ps / pl) ** (1 / g)
    asl = al * (ps / pl) ** ((g - 1) / (2 * g))
    rsr = rr * ((ps / pr + (g - 1) / (g + 1)) / ((g - 1) / (g + 1) * ps / pr + 1))
    S = ur + ar * np.sqrt((g + 1) / (2 * g) * ps / pr + (g - 1) / (2 * g))
    xi = (np.asarray(x) - x0) / t
    rho = np.empty_like(xi); u = np.empty_like(xi); p = np.empty_like(xi)
    head, tail = ul - al, us - asl
    for i, s in enumerate(xi):
        if s < head:
            rho[i], u[i], p[i] = rl, ul, pl
        elif s < tail:
            c = 2 / (g + 1) + (g - 1) / ((g + 1) * al) * (ul - s)
            rho[i], u[i], p[i] = rl * c ** (2 / (g - 1)), 2 / (g + 1) * (al + (g - 1) / 2 * ul + s), pl * c ** (2 * g / (g - 1))
        elif s < us:
            rho[i], u[i], p[i] = rsl, us, ps
        elif s < S:
            rho[i], u[i], p[i] = rsr, us, ps
        else:
            rho[i], u[i], p[i] = rr, ur, pr
    return dict(rho=rho, u=u, p=p, p_star=ps, u_star=us, shock_speed=S,
                x_shock=x0 + S * t, x_contact=x0 + us * t, rho_star_l=rsl, rho_star_r=rsr)
