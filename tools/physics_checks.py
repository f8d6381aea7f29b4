#!/usr/bin/env python
"""KAT-6 physics runs on the GPU: linear Landau damping and the bump-on-tail growth rate, fitted from the
field-energy history of the self-consistent Strang stepper and compared with the kinetic dispersion
relation solved numerically (scipy wofz).  Prints one JSON object; used by tests/test_gpu_physics.py.

Reference context: scripts/bump_on_tail.jl:14-30,64-71 plots W, K, W+K on log axes and eyeballs them;
nothing is asserted upstream.  Landau damping k = 0.5: gamma = -0.1533, omega = 1.4156.
"""
import json
import os
import sys

import numpy as np
from scipy.special import wofz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def plasma_Z(z):
    return 1j * np.sqrt(np.pi) * wofz(z)


def dielectric(omega, k, species):
    """species: list of (density fraction, drift u, thermal sigma)"""
    eps = 1.0 + 0j
    for n, u, s in species:
        zeta = (omega / k - u) / (np.sqrt(2) * s)
        eps += n / (k * s) ** 2 * (1 + zeta * plasma_Z(zeta))
    return eps


def most_unstable_root(k, species, re_range=(0.2, 2.5), im_range=(-0.6, 0.6)):
    best = None
    for wr in np.linspace(*re_range, 24):
        for wi in np.linspace(*im_range, 13):
            w = complex(wr, wi)
            for _ in range(60):
                f = dielectric(w, k, species)
                df = (dielectric(w + 1e-6, k, species) - dielectric(w - 1e-6, k, species)) / 2e-6
                step = f / df
                w -= step
                if abs(step) < 1e-13:
                    break
            if abs(dielectric(w, k, species)) < 1e-9 and w.real > 0.05:
                if best is None or w.imag > best.imag:
                    best = w
    return best


def run_vp(vpm, n, kappa, eps, alpha, sigma, v0, nh, order, dt, nsteps, seed=0x5EED0001):
    L = 2 * np.pi / kappa
    d = vpm.ParticleDistribution(1, 1, n)
    vpm.initialize_(d, vpm.BumpOnTail(eps=eps, kappa=kappa, alpha=alpha, sigma=sigma, v0=v0), seed=seed)
    pot = vpm.Potential(vpm.PeriodicBasisBSplineKit((0.0, L), order, nh))
    m = vpm.SplittingMethod(vpm.VlasovPoisson(d, pot), vpm.tspan_for(nsteps, dt), dt, field="selfconsistent")
    vpm.run_(m, diag_mode=1)
    return m.diagnostics


def fit_envelope_rate(t, W, t0, t1):
    """gamma from the local maxima of W(t) ~ exp(2 gamma t) cos^2(...) inside [t0, t1]"""
    lw = np.log(W)
    idx = [i for i in range(1, len(W) - 1) if lw[i] > lw[i - 1] and lw[i] >= lw[i + 1] and t0 <= t[i] <= t1]
    p = np.polyfit(t[idx], lw[idx], 1)
    return 0.5 * p[0], len(idx)


def fit_growth_rate(t, W, t0, t1):
    sel = (t >= t0) & (t <= t1)
    p = np.polyfit(t[sel], np.log(W[sel]), 1)
    return 0.5 * p[0]


def main(n=int(2e7)):
    import vpm_b200 as vpm
    out = {"particles": n}
    # ---- Landau damping: f = (1 + a cos(k x)) Maxwellian, k = 0.5 (sampler: eps = -a, no beam) ----
    k, a, dt, ns = 0.5, 0.05, 0.05, 500
    dg = run_vp(vpm, n, k, -a, 0.0, 1.0, 0.0, 32, 4, dt, ns)
    landau_hist = dg
    t = dt * np.arange(ns + 1)
    g, npk = fit_envelope_rate(t[1:], dg[1:, 0], 1.0, 16.0)
    w_th = most_unstable_root(k, [(1.0, 0.0, 1.0)], re_range=(1.0, 2.0), im_range=(-0.4, 0.0))
    E = dg[:, 0] + dg[:, 1]
    out["landau"] = {"gamma_fit": g, "gamma_theory": w_th.imag, "omega_theory": w_th.real, "peaks": npk,
                     "energy_drift_rel": float(abs(E[-1] - E[1]) / E[1]), "momentum_drift": float(abs(dg[-1, 2] - dg[0, 2]))}
    # ---- bump-on-tail (scripts/bump_on_tail.jl): kappa 0.3, eps 0.03, alpha 0.1, v0 4.5, sigma 0.5, dt 0.1, T 50 ----
    k, dt, ns = 0.3, 0.1, 500
    dg = run_vp(vpm, n, k, 0.03, 0.1, 0.5, 4.5, 16, 4, dt, ns)
    t = dt * np.arange(ns + 1)
    w_th = most_unstable_root(k, [(0.9, 0.0, 1.0), (0.1, 4.5, 0.5)])
    W = dg[1:, 0]
    g = fit_growth_rate(t[1:], W, 8.0, 20.0)
    E = dg[:, 0] + dg[:, 1]
    out["bump_on_tail"] = {"gamma_fit": g, "gamma_theory": w_th.imag, "omega_theory": w_th.real,
                           "W_first": float(W[0]), "W_max": float(W.max()), "t_sat": float(t[1:][W.argmax()]),
                           "energy_drift_rel": float(np.abs(E[1:] - E[1]).max() / E[1])}
    # growth rate proper: same equilibrium, small perturbation (eps = 1e-3) so that the linear phase spans
    # several e-foldings; fit log W where 50 W(0) < W < W_max / 20 (travelling unstable wave dominates)
    dgs = run_vp(vpm, n, k, 1e-3, 0.1, 0.5, 4.5, 16, 4, dt, 600)
    ts = dt * np.arange(601)
    Ws = dgs[1:, 0]
    lo_i = int(np.argmax(Ws > 50 * Ws[:20].max()))
    hi_i = int(np.argmax(Ws > Ws.max() / 20))
    if hi_i - lo_i < 20:
        lo_i = max(hi_i - 60, 1)
    gs = 0.5 * np.polyfit(ts[1:][lo_i:hi_i], np.log(Ws[lo_i:hi_i]), 1)[0]
    out["bump_on_tail_small_eps"] = {"gamma_fit": float(gs), "gamma_theory": w_th.imag, "fit_window": [float(ts[1:][lo_i]), float(ts[1:][hi_i])],
                                     "W_first": float(Ws[0]), "W_max": float(Ws.max())}
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")) or os.environ.get("GRAFT_REPO_ROOT"):
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez(os.path.join(ROOT, "gpurun_out", "physics_hist.npz"), landau=landau_hist, bot=dg, bot_small=dgs)
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    main(int(float(sys.argv[1])) if len(sys.argv) > 1 else int(2e7))
