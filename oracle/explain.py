"""Bounded parity for the Autorally pair — TEST INFRASTRUCTURE (used by tests/ and __graft_entry__.smoke()).

ARStandardCost is discontinuous: a point-sampled track map (a 5 cm texel either side of a wheel position), a crash flag that
latches when a texel value crosses `boundary_threshold`, a slip-angle threshold and a roll threshold. Two FP32 evaluations of
the same rollout whose states differ by 1e-6 can therefore differ by a whole crash cost. A quantile bar ("99 % within
1e-3") hides that; this module bounds it instead. For every sample whose trajectory cost is outside the tolerance:

  (i)   the device re-rolls the sample with every step dumped (mppib_sample_trajectories) and that dump must sum to the
        cost K1 stored — it IS K1's trajectory;
  (ii)  the oracle rolls the same constrained controls: the outputs must agree along the whole horizon to `state_tol`
        (the continuous part of the model — network, kinematics, integration — is within tolerance; the cost never feeds
        back into the state);
  (iii) step by step, the ORACLE's cost function evaluated on the DEVICE's outputs with the device's incoming crash flag
        must reproduce the device's per-step cost, unless that step sits on a discontinuity: a wheel within `texel_tol` of a
        texel edge (the reference's device code places the wheels with __sinf / __cosf, its host code with sinf / cosf,
        ar_standard_cost.cu:342-346), the slip angle within `thr_tol` of max_slip_ang, or |roll| within `thr_tol` of pi / 2.

If all three hold the sample's cost difference is a discontinuity crossing and nothing else; otherwise parity fails.
"""
from __future__ import annotations

import math

import numpy as np

from . import binding as O


def _near_texel_edge(cost, x, y, texel_tol):
    p = cost.params
    u = p.r_c1[0] * x + p.r_c2[0] * y + p.trs[0]
    v = p.r_c1[1] * x + p.r_c2[1] * y + p.trs[1]
    w = p.r_c1[2] * x + p.r_c2[2] * y + p.trs[2]
    pu, pv = u / w * p.map_width, v / w * p.map_height  # pixel coordinates of a point-filtered, normalised lookup
    du, dv = abs(pu - round(pu)), abs(pv - round(pv))
    return min(du, dv) < texel_tol


def autorally_outliers_explained(w, e, ref_costs, tol=1e-4, state_tol=2e-3, texel_tol=2e-3, thr_tol=1e-4, max_check=512):
    """w: workloads.autorally(); e: its engine (FLAG_WRITEBACK_CONTROLS) after solve(); ref_costs: the oracle's [N] costs on
    the same noise. Returns a dict of counts; raises AssertionError on an unexplained difference."""
    c = e.get_costs()[0]
    rel = np.abs(c - ref_costs) / np.maximum(np.abs(ref_costs), 1.0)
    bad = np.nonzero(rel > tol)[0]
    info = {"n": int(c.size), "outside_tol": int(bad.size), "max_rel": float(rel.max()), "explained": 0,
            "discontinuity_steps": 0}
    assert bad.size <= 0.03 * c.size, f"{bad.size} of {c.size} samples outside {tol}: not a tail"
    if bad.size == 0:
        return info
    bad = bad[np.argsort(-rel[bad])][:max_check]  # the worst first
    sp = w.sampler.params
    assert all(sp.control_cost_coeff[i] == 0.0 for i in range(2)), "per-step check assumes no likelihood-ratio term"
    outs, costs_dev, crash_dev = e.sample_trajectories(w.x0[0], w.U0[0], bad)
    np.testing.assert_allclose(costs_dev.sum(axis=1), c[bad], rtol=5e-6, err_msg="(i) the dump is not K1's trajectory")
    samples = e.get_samples()[0]
    T = w.T
    p = w.cost.params
    def on_discontinuity(y):
        cs, sn = math.cos(float(y[2])), math.sin(float(y[2]))
        edge = any(_near_texel_edge(w.cost, float(y[0]) + d * cs, float(y[1]) + d * sn, texel_tol) for d in (p.front_d, p.back_d))
        slip = -math.atan(float(y[5]) / abs(float(y[4]))) if abs(float(y[4])) >= 1e-3 else 0.0
        return edge or abs(abs(slip) - p.max_slip_ang) < thr_tol or abs(abs(float(y[3])) - math.pi / 2) < thr_tol

    for k, n in enumerate(bad):
        o_ref, c_ref, crash_ref = O.sampled_trajectory(w.dyn.DYN_ID, w.cost.COST_ID, w.dyn.params, p, sp, w.dyn.nn_theta,
                                                       w.cost.costmap, w.N, T, 0, int(n), False, w.dt, w.lambda_, w.alpha,
                                                       w.x0[0], w.U0[0], samples[n])
        dy = np.abs(outs[k][:, :7] - o_ref[:, :7]).max()
        assert dy < state_tol, f"(ii) sample {n}: outputs differ by {dy} along the horizon"
        events = 0
        crash_in = 0
        for t in range(T):
            y = outs[k][t]
            cst, _, crash_out = O.state_cost(w.cost.COST_ID, p, w.cost.costmap, y, t, crash_in)
            dev = float(costs_dev[k][t]) * T
            if not (abs(dev - cst) <= 1e-5 * max(1.0, abs(cst)) and crash_out == int(crash_dev[k][t])):
                assert on_discontinuity(y), (f"(iii) sample {n} step {t}: device step cost {dev} vs oracle on the same output "
                                             f"{cst}, crash {int(crash_dev[k][t])} vs {crash_out}, and no discontinuity nearby")
                events += 1
            crash_in = int(crash_dev[k][t])  # follow the device's latch: each step is judged on its own decision
        # (iv) where the oracle's OWN trajectory first departs from the device's (flag or step cost), one of the two sits on a
        # discontinuity; after the flags differ the remaining steps follow from the latch
        for t in range(T):
            d_cost = abs(float(c_ref[t]) - float(costs_dev[k][t])) * T
            if int(crash_ref[t]) != int(crash_dev[k][t]) or d_cost > 1e-3 * max(1.0, abs(float(c_ref[t]) * T)):
                assert on_discontinuity(outs[k][t]) or on_discontinuity(o_ref[t]), (
                    f"(iv) sample {n} step {t}: oracle and device part ways (step cost {float(c_ref[t]) * T} vs "
                    f"{float(costs_dev[k][t]) * T}, crash {int(crash_ref[t])} vs {int(crash_dev[k][t])}) away from any discontinuity")
                events += 1
                break
        assert events > 0, f"sample {n}: cost {c[n]} vs oracle {ref_costs[n]} differ with no discontinuity event on the way"
        info["discontinuity_steps"] += events
        info["explained"] += 1
    return info
