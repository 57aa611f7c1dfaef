"""The sympy -> CUDA generator (curiousrl_b200/codegen.py) on a scenario it has never seen: a damped pendulum on a cart
with exp / tanh / sqrt terms, one bounded state and a bounded action, time-varying cost parameters.  The generated struct
(plain and log-barrier variant) is compiled for the host with g++ and compared, entry by entry, with what
`sympy.lambdify` evaluates for the same expressions - i.e. with what the reference's evaluators would compute
(ilqr_dynamic_model.py:33-35, ilqr_obj_fun.py:46-51, log_barrier_ilqr.py:91-95).  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import sympy as sp

from curiousrl_b200 import codegen
from curiousrl_b200.scenario.dynamic_model import DynamicModelWrapper


class ToyPendulumCart(DynamicModelWrapper):
    def __init__(self, T=12):
        xu = sp.symbols("x_u:4")                                   # angle, angular velocity, cart position | force
        w = sp.symbols("w:2")
        h = 0.05
        f = sp.Array([xu[0] + h * xu[1],
                      xu[1] + h * (-9.8 * sp.sin(xu[0]) - 0.3 * sp.tanh(xu[1]) + xu[3] * sp.cos(xu[0]) ** 2),
                      xu[2] + h * sp.exp(-xu[2] ** 2) * xu[3]])
        cost = w[0] * sp.sqrt(xu[0] ** 2 + 0.01) + w[1] * (xu[2] - 1) ** 2 + 0.1 * xu[3] ** 2 + xu[1] ** 3 * 1e-3
        table = np.ones((T, 2))
        table[-1] = (50.0, 20.0)
        super().__init__(dynamic_function=f, x_u_var=xu,
                         constr=np.asarray([[-np.inf, np.inf], [-np.inf, np.inf], [-2.0, 2.0], [-5.0, 5.0]]),
                         init_state=np.zeros((3, 1)), init_action=np.zeros((T, 1, 1)), obj_fun=cost, add_param_var=w,
                         add_param=table)


HARNESS = r"""
#include "model.h"
using namespace crl;
extern "C" {
void t_step(const double* x, double* xn) { GenToyPendulumCart::step(x, xn); }
void t_jac(const double* x, double* F) { GenToyPendulumCart::jac(x, F); }
double t_cost(const double* x, const double* p) { return GenToyPendulumCart::cost(x, p); }
void t_grad(const double* x, const double* p, double* c) { GenToyPendulumCart::grad(x, p, c); }
void t_hess(const double* x, const double* p, double* C) { GenToyPendulumCart::hess(x, p, C); }
double b_cost(const double* x, const double* p) { return GenToyPendulumCartBarrier::cost(x, p); }
void b_grad(const double* x, const double* p, double* c) { GenToyPendulumCartBarrier::grad(x, p, c); }
void b_hess(const double* x, const double* p, double* C) { GenToyPendulumCartBarrier::hess(x, p, C); }
void t_bounds(double* lo, double* hi) { for (int i = 0; i < GenToyPendulumCart::D; ++i) { lo[i] = GenToyPendulumCart::lo(i); hi[i] = GenToyPendulumCart::hi(i); } }
int t_sizes() { return GenToyPendulumCart::N * 1000 + GenToyPendulumCart::M * 100 + GenToyPendulumCart::P * 10 + GenToyPendulumCartBarrier::P; }
}
"""


@pytest.fixture(scope="module")
def toy(tmp_path_factory):
    d = tmp_path_factory.mktemp("codegen")
    sc = ToyPendulumCart()
    with open(d / "model.h", "w") as fh:
        fh.write(codegen.generate_header([sc]))
    with open(d / "harness.cpp", "w") as fh:
        fh.write(HARNESS)
    so = str(d / "toy.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", str(d / "harness.cpp"), "-o", so],
                   check=True)
    lib = ctypes.CDLL(so)
    lib.t_cost.restype = lib.b_cost.restype = ctypes.c_double
    return sc, lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_sizes_and_box(toy):
    sc, lib = toy
    assert lib.t_sizes() == 3 * 1000 + 1 * 100 + 2 * 10 + 4            # n, m, p, barrier row = (w0, w1, t, t0)
    lo, hi = np.zeros(4), np.zeros(4)
    lib.t_bounds(_p(lo), _p(hi))
    assert np.array_equal(np.stack([lo, hi], axis=1), sc.constr)


def test_generated_functions_equal_lambdified_expressions(toy):
    sc, lib = toy
    xu, w = sc.x_u_var, sc.add_param_var
    f = sp.lambdify([xu], sc.dynamic_function.tolist(), "math")
    F = sp.lambdify([xu], sp.transpose(sp.derive_by_array(sc.dynamic_function, xu)).tolist(), "math")
    g_sym = sp.derive_by_array(sc.obj_fun, xu)
    H_sym = sp.derive_by_array(g_sym, xu)
    l = sp.lambdify([xu, w], sc.obj_fun, "math")
    g = sp.lambdify([xu, w], g_sym.tolist(), "math")
    H = sp.lambdify([xu, w], H_sym.tolist(), "math")
    rng = np.random.default_rng(0)
    for _ in range(20):
        x = rng.uniform(-1.5, 1.5, 4)
        p = rng.uniform(0.5, 3.0, 2)
        xn, Fj, c, C = np.zeros(3), np.zeros(12), np.zeros(4), np.zeros(16)
        lib.t_step(_p(x), _p(xn)); lib.t_jac(_p(x), _p(Fj)); lib.t_grad(_p(x), _p(p), _p(c)); lib.t_hess(_p(x), _p(p), _p(C))
        assert np.allclose(xn, f(x), rtol=1e-14, atol=1e-15)
        assert np.allclose(Fj.reshape(3, 4), np.asarray(F(x), dtype=float), rtol=1e-13, atol=1e-15)
        assert np.isclose(lib.t_cost(_p(x), _p(p)), l(x, p), rtol=1e-14)
        assert np.allclose(c, g(x, p), rtol=1e-13, atol=1e-15)
        assert np.allclose(C.reshape(4, 4), np.asarray(H(x, p), dtype=float) + 1e-100 * np.eye(4), rtol=1e-13, atol=1e-15)


def test_barrier_variant_adds_the_log_terms_of_every_finite_bound(toy):
    sc, lib = toy
    xu, w = sc.x_u_var, sc.add_param_var
    t = sp.symbols("t")
    expr = sc.obj_fun
    for i, (lo, hi) in enumerate(np.asarray(sc.constr)):
        if np.isfinite(lo):
            expr = expr + (-1 / t) * sp.log(-(lo - xu[i]))
        if np.isfinite(hi):
            expr = expr + (-1 / t) * sp.log(-(xu[i] - hi))
    assert expr.has(sp.log) and len(expr.atoms(sp.log)) == 4            # the bounded STATE has its two terms too
    g_sym = sp.derive_by_array(expr, xu)
    l = sp.lambdify([xu, (*w, t)], expr, "math")
    g = sp.lambdify([xu, (*w, t)], g_sym.tolist(), "math")
    H = sp.lambdify([xu, (*w, t)], sp.derive_by_array(g_sym, xu).tolist(), "math")
    rng = np.random.default_rng(1)
    for _ in range(20):
        x = rng.uniform(-1.5, 1.5, 4)
        p = np.concatenate([rng.uniform(0.5, 3.0, 2), [rng.choice([0.5, 2.0, 100.0])], [7.0]])      # (w, t, t0): t0 is unused here
        c, C = np.zeros(4), np.zeros(16)
        lib.b_grad(_p(x), _p(p), _p(c)); lib.b_hess(_p(x), _p(p), _p(C))
        assert np.isclose(lib.b_cost(_p(x), _p(p)), l(x, p[:3]), rtol=1e-14)
        assert np.allclose(c, g(x, p[:3]), rtol=1e-13, atol=1e-15)
        assert np.allclose(C.reshape(4, 4), np.asarray(H(x, p[:3]), dtype=float) + 1e-100 * np.eye(4), rtol=1e-13, atol=1e-15)
