"""sympy -> CUDA device models for scenarios without a hand-written model.

`python -m curiousrl_b200.codegen` writes `csrc/ilqr_generated.cuh` from the scenario classes in
`scenario/dynamic_model/vehicles.py`.  The header is committed: the library is built from it without sympy.

What is generated per scenario (a struct `Gen<ClassName>` used by csrc/ilqr_generic.cuh):

    N, M, D, P, lo(i), hi(i)                sizes and the box `constr`
    step(xu, xn)                            x_{t+1} = f(x_t, u_t)              scenario.dynamic_function
    jac(xu, F[N*D])                         transpose(derive_by_array(f, xu))  ilqr_dynamic_model.py:35
    cost(xu, p) / grad(xu, p, c[D])         l, derive_by_array(l, xu)          ilqr_obj_fun.py:46-48
    hess(xu, p, C[D*D])                     d^2 l + 1e-100 I                   ilqr_obj_fun.py:49-51

and a struct `Gen<ClassName>Barrier : Gen<ClassName>` with cost / grad / hess of LogBarrieriLQR's objective
(log_barrier_ilqr.py:84-101: every finite bound adds a (-1/t) log term; parameter row = the scenario's columns, t, t0).

Every expression is differentiated and printed on its own, term order as sympy prints it, WITHOUT common
sub-expression elimination - the reference evaluates the code `sympy.lambdify` prints for exactly these
expressions (under numba), so the operation sequence per entry is the reference's; sharing of repeated
sub-expressions is left to nvcc, which may not change values (`-fmad=false`).  Integer powers are printed as
products (numba lowers `x**2` to `x*x`), everything else goes to the C math library (`sin`, `cos`, `sqrt`, `asin`,
`atan2`, `pow`).
"""
from __future__ import annotations

import os

import numpy as np
import sympy as sp
from sympy.printing.c import C99CodePrinter
from sympy.printing.precedence import PRECEDENCE

HEADER_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "ilqr_generated.cuh")

#: scenario class name -> model id of the C ABI (include/curiousrl_b200.h: CRL_GEN_*)
GENERIC_MODEL_IDS = {"CartPoleSwingUp1": 6, "CartPoleSwingUp2": 7, "VehicleTracking": 8, "CarParking": 9}
#: the same scenarios under LogBarrieriLQR's cost (CRL_GEN_*_BARRIER)
GENERIC_BARRIER_MODEL_IDS = {name: mid + 4 for name, mid in GENERIC_MODEL_IDS.items()}


class _DevicePrinter(C99CodePrinter):
    def __init__(self):
        super().__init__(dict(precision=17))

    def _print_Pow(self, expr):
        base, exp = expr.base, expr.exp
        if exp.is_Integer and 2 <= abs(int(exp)) <= 4:
            prod = "*".join([self.parenthesize(base, PRECEDENCE["Mul"], strict=True)] * abs(int(exp)))
            return "(%s)" % prod if exp > 0 else "(1.0/(%s))" % prod
        return super()._print_Pow(expr)


def _emit_assignments(printer, target, exprs):
    return ["        %s[%d] = %s;" % (target, i, printer.doprint(sp.sympify(e))) for i, e in enumerate(exprs)]


def _aliases(symbols, array):
    """`const double x_u0 = x[0], ...;` - expressions are printed with the scenario's own symbol names, because the
    order in which sympy prints the terms of a sum or product (= the order of the floating-point operations)
    depends on them."""
    if not symbols:
        return []
    return ["        const double %s;" % ", ".join("%s = %s[%d]" % (str(s), array, i) for i, s in enumerate(symbols)),
            "        %s" % " ".join("(void)%s;" % str(s) for s in symbols)]


def _fmt_bound(v):
    if np.isinf(v):
        return "-CRL_GEN_INF" if v < 0 else "CRL_GEN_INF"
    return repr(float(v))


def generate_struct(scenario, name=None) -> str:
    """The `Gen<name>` struct text of one scenario object (name: the scenario's class name unless given)."""
    printer = _DevicePrinter()
    xu = list(scenario.x_u_var)
    n = int(scenario.init_state.shape[0])
    d = len(xu)
    m = d - n
    pv = scenario.add_param_var
    pv = [] if pv is None else list(pv)
    p = max(1, len(pv))                      # no add_param: one unused zero column (ilqr_obj_fun.py:91-92)
    ax, ap = _aliases(xu, "x"), _aliases(pv, "p")
    f = scenario.dynamic_function
    F = sp.transpose(sp.derive_by_array(f, xu))                                    # [n][d]
    l = scenario.obj_fun
    g = sp.derive_by_array(l, xu)
    H = np.asarray(sp.derive_by_array(g, xu)) + 1e-100 * np.eye(d)                 # ilqr_obj_fun.py:51
    constr = np.asarray(scenario.constr, dtype=np.float64)
    name = type(scenario).__name__ if name is None else name
    out = ["// %s  (n = %d, m = %d, p = %d)" % (name, n, m, p),
           "struct Gen%s {" % name,
           "    static constexpr int N = %d, M = %d, D = %d, P = %d;" % (n, m, d, p),
           "    static constexpr int kModelId = %d;" % GENERIC_MODEL_IDS.get(name, -1),
           "    static constexpr bool kBarrier = false;",
           "    CRL_GEN_FN double lo(int i) { constexpr double b[D] = {%s}; return b[i]; }" % ", ".join(_fmt_bound(v) for v in constr[:, 0]),
           "    CRL_GEN_FN double hi(int i) { constexpr double b[D] = {%s}; return b[i]; }" % ", ".join(_fmt_bound(v) for v in constr[:, 1]),
           "    CRL_GEN_FN void step(const double* x, double* xn) {"]
    out += ax + _emit_assignments(printer, "xn", list(f))
    out += ["    }", "    CRL_GEN_FN void jac(const double* x, double* F) {"]
    out += ax + _emit_assignments(printer, "F", [F[i, j] for i in range(n) for j in range(d)])
    out += ["    }", "    CRL_GEN_FN double cost(const double* x, const double* p) {", "        (void)p;"]
    out += ax + ap + ["        return %s;" % printer.doprint(sp.sympify(l))]
    out += ["    }", "    CRL_GEN_FN void grad(const double* x, const double* p, double* c) {", "        (void)p;"]
    out += ax + ap + _emit_assignments(printer, "c", list(g))
    out += ["    }", "    CRL_GEN_FN void hess(const double* x, const double* p, double* C) {", "        (void)p; (void)x;"]
    out += ax + ap + _emit_assignments(printer, "C", [H[i, j] for i in range(d) for j in range(d)])
    out += ["    }", "};", ""]
    return "\n".join(out)


def generate_barrier_struct(scenario, name=None) -> str:
    """`Gen<name>Barrier`: the scenario under LogBarrieriLQR's cost (log_barrier_ilqr.py:84-101) - every finite bound,
    state bounds included, adds (-1/t) log(-(lo - x_u[i])) / (-1/t) log(-(x_u[i] - hi)).  Dynamics and box are the
    plain model's; the parameter row is (the scenario's add_param columns, t, t0) with t0 = the stale barrier
    parameter the first backward pass of an inner loop still sees (ilqr_generic.cuh)."""
    printer = _DevicePrinter()
    xu = list(scenario.x_u_var)
    d = len(xu)
    pv = [] if scenario.add_param_var is None else list(scenario.add_param_var)
    t_var = sp.symbols("t")
    constr = np.asarray(scenario.constr, dtype=np.float64)
    l = scenario.obj_fun
    for i, c in enumerate(constr):                       # same construction order as the reference (:91-95)
        if not np.isinf(c[0]):
            l = l + (-1 / t_var) * sp.log(-(c[0] - xu[i]))
        if not np.isinf(c[1]):
            l = l + (-1 / t_var) * sp.log(-(xu[i] - c[1]))
    g = sp.derive_by_array(l, xu)
    H = np.asarray(sp.derive_by_array(g, xu)) + 1e-100 * np.eye(d)
    name = type(scenario).__name__ if name is None else name
    ax, ap = _aliases(xu, "x"), _aliases(pv + [t_var], "p")
    out = ["// %s under the log-barrier cost  (parameter row: %d scenario columns, t, t0)" % (name, len(pv)),
           "struct Gen%sBarrier : Gen%s {" % (name, name),
           "    static constexpr int P = %d;" % (len(pv) + 2),
           "    static constexpr int kT = %d;                   // index of t in the parameter row; t0 follows" % len(pv),
           "    static constexpr int kModelId = %d;" % GENERIC_BARRIER_MODEL_IDS.get(name, -1),
           "    static constexpr bool kBarrier = true;",
           "    CRL_GEN_FN double cost(const double* x, const double* p) {"]
    out += ax + ap + ["        return %s;" % printer.doprint(sp.sympify(l))]
    out += ["    }", "    CRL_GEN_FN void grad(const double* x, const double* p, double* c) {"]
    out += ax + ap + _emit_assignments(printer, "c", list(g))
    out += ["    }", "    CRL_GEN_FN void hess(const double* x, const double* p, double* C) {"]
    out += ax + ap + _emit_assignments(printer, "C", [H[i, j] for i in range(d) for j in range(d)])
    out += ["    }", "};", ""]
    return "\n".join(out)


HEADER_PREAMBLE = ['#pragma once', '#include <math.h>', '',
                   '#if defined(__CUDACC__)', '#define CRL_GEN_FN __host__ __device__ __forceinline__ static',
                   '#else', '#define CRL_GEN_FN inline static', '#endif',
                   '#define CRL_GEN_INF (__builtin_huge_val())', '', 'namespace crl {', '']


def generate_user_header(scenario) -> str:
    """Header of ONE scenario for the run-time compiled library (jit.py): structs GenUser / GenUserBarrier."""
    head = ['// GENERATED by curiousrl_b200/codegen.py from a %s object at run time (sympy %s).' % (type(scenario).__name__,
                                                                                                  sp.__version__)]
    body = [generate_struct(scenario, "User"), generate_barrier_struct(scenario, "User")]
    return "\n".join(head + HEADER_PREAMBLE + body + ["}  // namespace crl", ""])


def generate_header(scenarios) -> str:
    head = ['// GENERATED by curiousrl_b200/codegen.py from scenario/dynamic_model/vehicles.py - do not edit.',
            '// sympy %s.  Device models of the scenarios without a hand-written model (SURVEY.md 8(f) N2);' % sp.__version__,
            '// expression forms are what sympy.lambdify prints for the reference (no CSE, integer powers as products).',
            '#pragma once', '#include <math.h>', '',
            '#if defined(__CUDACC__)', '#define CRL_GEN_FN __host__ __device__ __forceinline__ static',
            '#else', '#define CRL_GEN_FN inline static', '#endif',
            '#define CRL_GEN_INF (__builtin_huge_val())', '', 'namespace crl {', '']
    body = [generate_struct(s) for s in scenarios] + [generate_barrier_struct(s) for s in scenarios]
    return "\n".join(head + body + ["}  // namespace crl", ""])


def default_scenarios():
    from .scenario.dynamic_model import vehicles
    return [getattr(vehicles, name)() for name in GENERIC_MODEL_IDS]


def main():
    text = generate_header(default_scenarios())
    with open(HEADER_PATH, "w") as fh:
        fh.write(text)
    print("wrote %s (%d lines)" % (HEADER_PATH, text.count("\n")))


if __name__ == "__main__":
    main()
