"""Generates tests/golden/*.npz from the CPU oracle (oracle/vpm_oracle.c).

The reference has no golden vectors and cannot run here (no Julia), so these fixtures
pin the *oracle* (regression) and give the GPU tests box-independent inputs/outputs.  They are NOT
reference outputs: parity with the Julia package stays "unpinned" (see oracle header).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def vp_case(name, K, nh, n, L, dt, nsteps, chi=1.0, seed=1):
    rng = np.random.default_rng(seed)
    x, v, w = orc.sample_bump_on_tail(n, seed=seed, kappa=2 * np.pi / L)
    x = x + L * rng.integers(-3, 4, n)  # unwrapped positions spanning several periods
    w = w * rng.uniform(0.5, 1.5, n)
    xs = orc.XSpace(0.0, L, K, nh)
    rhs = xs.deposit(x, w)
    phi = xs.poisson_solve(rhs)
    dphi = xs.eval(phi, x, 1)
    x1, v1, diag, phi1 = xs.strang_selfconsistent(x, v, w, dt, nsteps, chi=chi)
    xf, vf, phif = xs.strang_frozen(x, v, x, w, dt, nsteps)
    np.savez_compressed(os.path.join(HERE, name), K=K, nh=nh, L=L, dt=dt, nsteps=nsteps, chi=chi, x=x, v=v, w=w,
                        rhs=rhs, phi=phi, dphi=dphi, x1=x1, v1=v1, diag=diag, phi1=phi1, xf=xf, vf=vf, phif=phif)


def lb_case(name, K, nknots, n, dt, nsteps, seed=2):
    rng = np.random.default_rng(seed)
    v = np.r_[rng.standard_normal(n // 2) + 2.0, rng.standard_normal(n - n // 2) - 2.0]
    v[:6] = [-10.0, 10.0, -9.9, 9.9, -11.0, 12.5]  # boundary cells and out-of-domain particles
    w = rng.uniform(0.5, 1.5, n) / n
    vs = orc.VSpace(-10.0, 10.0, nknots, K)
    rhs = vs.deposit(v, w)
    coef = vs.mass_solve(rhs)
    f, df = vs.eval(coef, v), vs.eval(coef, v, 1)
    m5 = vs.moments(coef, v)
    vdot_lb, _, _ = vs.lb_rhs(v, w, nu=0.7, conservative=False)
    vdot_clb, _, A = vs.lb_rhs(v, w, nu=0.7, conservative=True)
    v_lb, d_lb = vs.rk438(v, w, 0.7, dt, nsteps, conservative=False)
    v_clb, d_clb = vs.rk438(v, w, 0.7, dt, nsteps, conservative=True)
    np.savez_compressed(os.path.join(HERE, name), K=K, nknots=nknots, dt=dt, nsteps=nsteps, nu=0.7, v=v, w=w, rhs=rhs,
                        coef=coef, f=f, df=df, m5=m5, vdot_lb=vdot_lb, vdot_clb=vdot_clb, A=A, v_lb=v_lb, d_lb=d_lb,
                        v_clb=v_clb, d_clb=d_clb, mass=vs.mass())


if __name__ == "__main__":
    vp_case("vp_k4_n16.npz", 4, 16, 4001, 2 * np.pi / 0.3, 0.1, 4)
    vp_case("vp_k3_n16_cfg1.npz", 3, 16, 3000, 1.0, 0.1, 3, seed=3)       # scripts/vlasov_poisson.jl grid
    vp_case("vp_k5_n11_chi.npz", 5, 11, 2500, 7.5, 0.05, 3, chi=1.7, seed=4)
    lb_case("lb_k4_n41.npz", 4, 41, 3001, 0.05, 3)
    lb_case("lb_k5_n12.npz", 5, 12, 2000, 0.02, 2, seed=5)
    print("golden fixtures written to", HERE)
