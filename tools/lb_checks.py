#!/usr/bin/env python
"""KAT-7 at scale: BASELINE config 4 as shipped (scripts/lenard_bernstein_conservative.jl:10-18,46-64) on the GPU.

DoubleMaxwellian(+-2), nknots 41, order 4, v in (-10, 10) Dirichlet, nu = 1, dt = 1e-2, t in (0, 500): 5e4 RK438
steps (2e5 right-hand sides) of the conservative Lenard-Bernstein model, and the same number of steps of the
plain model.  The script upstream prints sum v and sum v^2 before/after (":49-50,64") and eyeballs them; here
the whole history is kept and checked:

* CLB: momentum sum v and energy sum v^2 conserved (relative drift over the run);
* the distribution relaxes to the Maxwellian with the SAME mean and variance: the normalised fourth moment
  E[(v-m)^4]/var^2 goes from 43/25 = 1.72 (two unit Maxwellians at +-2) to 3, the sixth E[(v-m)^6]/var^3 to 15;
* plain LB relaxes to the unit Maxwellian instead (variance 5 -> 1): energy is NOT conserved.

* entropy history (round 2; NON-REFERENCE diagnostic S = -sum w ln max(f_s(v), floor)): CLB must produce
  entropy monotonically (H-theorem of the energy-conserving operator) towards ln sqrt(2 pi e var); plain LB cools the
  ensemble towards the unit Maxwellian, so S falls towards ln sqrt(2 pi e).

Prints one JSON object (profiles/r1_physics_clb_1e7.json, r2_physics_entropy_*.json are copies).
Usage: python tools/lb_checks.py [N] [nsteps] [clb,lb] [entropy]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def moments(v):
    m = v.mean()
    c = v - m
    var = (c * c).mean()
    return {"mean": float(m), "var": float(var), "m4_over_var2": float((c ** 4).mean() / var ** 2),
            "m6_over_var3": float((c ** 6).mean() / var ** 3)}


def run(vpm, n, nsteps, conservative, chunk, entropy=False):
    sd = vpm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    d = vpm.ParticleDistribution(1, 1, n)
    vpm.initialize_(d, vpm.DoubleMaxwellian((-10.0, 10.0), 2.0))
    model = (vpm.ConservativeLenardBernstein if conservative else vpm.LenardBernstein)(d, vpm.CollisionEntropy(sd), nu=1.0)
    dt = 1e-2
    snaps = [dict(t=0.0, **moments(d.get("v")))]
    diags, ents = [], []
    t0 = time.perf_counter()
    done = 0
    step_s = 0.0      # time inside the stepper calls (synchronous: they return with the diagnostics history on the host)
    while done < nsteps:
        k = min(chunk, nsteps - done)
        gi = vpm.GeometricIntegrator(model, vpm.tspan_for(k, dt), dt)
        t1 = time.perf_counter()
        vpm.run_(gi, entropy=entropy)
        step_s += time.perf_counter() - t1
        diags.append(gi.diagnostics if not diags else gi.diagnostics[1:])
        if entropy:
            ents.append(gi.entropy if not ents else gi.entropy[1:])
        done += k
        snaps.append(dict(t=done * dt, **moments(d.get("v"))))
    wall = time.perf_counter() - t0
    dg = np.concatenate(diags)
    return dg, snaps, wall, (np.concatenate(ents) if entropy else None), step_s


def main(n=int(1e7), nsteps=50000, models=("clb", "lb"), entropy=False):
    import vpm_b200 as vpm
    out = {"particles": n, "steps": nsteps, "dt": 1e-2, "nu": 1.0}
    chunk = max(nsteps // 10, 1)
    for cons in [m == "clb" for m in models]:
        dg, snaps, wall, S, step_s = run(vpm, n, nsteps, cons, chunk, entropy)
        key = "clb" if cons else "lb"
        out[key] = {
            "sum_v_first_last": [float(dg[0, 0]), float(dg[-1, 0])],
            "sum_v2_first_last": [float(dg[0, 1]), float(dg[-1, 1])],
            "momentum_drift_over_N": float(np.abs(dg[:, 0] - dg[0, 0]).max() / n),
            "energy_drift_rel": float(np.abs(dg[:, 1] - dg[0, 1]).max() / dg[0, 1]),
            "snapshots": snaps,
            "wall_s": wall,
            "particle_steps_per_s_incl_snapshots": n * nsteps / wall,
            # the snapshots are host work (download + numpy moments of all velocities); the stepping itself:
            "stepper_s": step_s, "ms_per_step": 1e3 * step_s / nsteps, "particle_steps_per_s": n * nsteps / step_s,
        }
        if S is not None:
            var = snaps[-1]["var"]
            dS = np.diff(S)
            out[key]["entropy"] = {
                "definition": "S = -sum_p w_p ln max(f_s(v_p), 1e-14), f_s = spline projection of the state (non-reference diagnostic)",
                "first_last": [float(S[0]), float(S[-1])], "min_step_change": float(dS.min()), "max_step_change": float(dS.max()),
                "monotone_increasing": bool(np.all(dS > 0)), "monotone_decreasing": bool(np.all(dS < 0)),
                "maxwellian_of_final_variance": float(0.5 * np.log(2 * np.pi * np.e * var)),
                "samples": [[float(i * 1e-2), float(S[i])] for i in np.unique(np.linspace(0, len(S) - 1, 41).astype(int))],
            }
        if os.path.isdir(os.path.join(ROOT, "gpurun_out")) or os.environ.get("GRAFT_REPO_ROOT"):
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            np.save(os.path.join(ROOT, "gpurun_out", f"lb_hist_{key}.npy"), dg[:: max(len(dg) // 5000, 1)])
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    main(int(float(sys.argv[1])) if len(sys.argv) > 1 else int(1e7), int(sys.argv[2]) if len(sys.argv) > 2 else 50000,
         tuple(sys.argv[3].split(",")) if len(sys.argv) > 3 else ("clb", "lb"), len(sys.argv) > 4 and sys.argv[4] == "entropy")
