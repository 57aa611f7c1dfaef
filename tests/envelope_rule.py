"""SURVEY.md section 4.3, per problem: a solver result is judged against the reference's OWN reproducibility.

The fixtures tests/golden/Envelope_RoboticArm_T*.npz (oracle/make_golden_envelope.py, unmodified reference) hold, for
every problem of the seeded RoboticArm batches, 24 re-solves under rounding-sized perturbations: K, k times
(1 + 1e-15 N), K, k times (1 + 1e-12 N), and the inputs F, c of the backward pass times (1 + 1e-15 N) - 8 seeds each.

    radius S_b   = max over the 24 runs of |J_run - J_ref| / |J_ref|      how far the reference is from itself
    trace-stable = all 24 runs keep the reference's iteration count and accepted-step sequence

Rule for a candidate result (iterations, accepted steps, J):
    * trace-stable problem: the same iteration count and accepted steps, and |J - J_ref| <= max(1e-9, 2 S_b) |J_ref|
      (1e-9 is north_star's FP64 bar; it is what applies wherever the reference reproduces itself that well);
    * any other problem: its trace equals the reference's or one of the 24 perturbed traces, OR its cost is inside
      the envelope, |J - J_ref| <= 2 S_b |J_ref|.
"""
import numpy as np


class Envelope:
    def __init__(self, golden_batch, envelope):
        self.g, self.e = golden_batch, envelope
        self.B = len(golden_batch["batch_iters"])

    def traces(self, b):
        g, e = self.g, self.e
        n = int(g["batch_iters"][b])
        out = [np.asarray(g["batch_alphas"][b, :n], dtype=np.int64)]
        fam, runs = e["iters"].shape[1:]
        for f in range(fam):
            for s in range(runs):
                out.append(np.asarray(e["alphas"][b, f, s, :int(e["iters"][b, f, s])], dtype=np.int64))
        return out

    def radius(self, b):
        Jref = float(self.g["batch_J"][b])
        return float(np.abs(self.e["J"][b] - Jref).max() / abs(Jref))

    def trace_stable(self, b):
        ref = self.traces(b)[0]
        return all(len(t) == len(ref) and np.array_equal(t, ref) for t in self.traces(b)[1:])

    def judge(self, b, iters, alphas, J):
        """-> dict(ok, trace_stable, radius, deviation, trace_match, trace_in_set)"""
        Jref = float(self.g["batch_J"][b])
        S = self.radius(b)
        dev = abs(float(J) - Jref) / abs(Jref)
        mine = np.asarray(alphas[:int(iters)], dtype=np.int64)
        tr = self.traces(b)
        match = len(tr[0]) == len(mine) and np.array_equal(tr[0], mine)
        in_set = any(len(t) == len(mine) and np.array_equal(t, mine) for t in tr)
        stable = self.trace_stable(b)
        if stable:
            ok = match and dev <= max(1e-9, 2 * S)
        else:
            ok = in_set or dev <= 2 * S
        return dict(ok=ok, trace_stable=stable, radius=S, deviation=dev, trace_match=match, trace_in_set=in_set)


def summarise(verdicts):
    n = len(verdicts)
    return dict(problems=n,
                reproducible_1e9=sum(v["trace_stable"] and v["radius"] <= 1e-9 for v in verdicts),
                trace_stable=sum(v["trace_stable"] for v in verdicts),
                trace_match=sum(v["trace_match"] for v in verdicts),
                trace_in_set=sum(v["trace_in_set"] for v in verdicts),
                passed=sum(v["ok"] for v in verdicts),
                worst_deviation_over_radius=max(v["deviation"] / max(v["radius"], 1e-16) for v in verdicts))


# ---------------------------------------------------------------------------------------------- FP32 mode (north_star)
def f32_reproducible(golden_batch, envelope32, tol=1e-4, margin=2.0, criterion=1e-6):
    """Which problems of a seeded batch an FP32 solver can be held to `tol` on the final cost - by the reference's own
    standard (fixtures tests/golden/Envelope32_*.npz, oracle/make_golden_envelope.py):

    stable   the reference keeps its iteration count and moves its final cost by <= tol when float32-sized (6e-8)
             relative noise is put on the outputs (K, k) or on the inputs (F, c) of its backward pass, 8 seeds each;
    clear    no stopping decision of the reference's trace sits within a factor `margin` of the criterion: an FP32
             trajectory carries ~1e-6 relative noise in its cost, so |dJ / J| = 6e-7 against the 1e-6 rule (the
             reference's TwoLink problem 2: it stops there, 44 % above the optimum an FP32 - or a perturbed FP64 - run
             goes on to find) is a coin toss at that resolution, not a property of the solver.
    Returns (stable, stable & clear) boolean masks."""
    g, e = golden_batch, envelope32
    Jref = np.asarray(g["batch_J"], dtype=np.float64)
    stable = (np.all(np.abs(e["J"] - Jref[:, None, None]) <= tol * np.abs(Jref[:, None, None]), axis=(1, 2))
              & np.all(e["iters"] == np.asarray(g["batch_iters"])[:, None, None], axis=(1, 2)))
    clear = np.zeros(len(Jref), dtype=bool)
    for b in range(len(Jref)):
        n = int(g["batch_iters"][b])
        Js, al = g["batch_Js"][b], g["batch_alphas"][b]
        ok = True
        for i in range(1, n):
            r = abs((Js[i] - Js[i - 1]) / Js[i - 1])
            if i == n - 1:
                ok = ok and (al[i] < 0 or r < criterion / margin)       # stopped by a rejected line search, or clearly
            else:
                ok = ok and r > criterion * margin                       # went on, clearly
        clear[b] = ok
    return stable, stable & clear
