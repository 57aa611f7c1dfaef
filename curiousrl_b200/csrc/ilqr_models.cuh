// Device models of the three arm scenarios: dynamics, analytic Jacobians, tracking cost.
//
// Hand-written from the scenario formulas (reference files, relative to
// /root/reference/CuriousRL/scenario/dynamic_model/):
//   TwoLink    two_link_planar_manipulator.py:27-44     n=6 m=2
//   ThreeLink  three_link_planar_manipulator.py:30-59   n=8 m=3
//   RoboticArm robotic_arm_tracking.py:28-66            n=6 m=2
// Expression forms (operation order, printed constants) follow what the reference's
// sympy -> lambdify pipeline evaluates, so results agree with it to rounding.
//
// Shared structure exploited by the Riccati kernel (ilqr_core.cuh).  The state splits into
// NS "dynamic" entries s (joint angles/velocities) and NO = 2 "output" entries y (the
// end-effector position) that are functions of the previous s only and never feed back:
//
//        | A(s,u)   0   B(s) |   rows: s'          A: NS x NS, B: NS x M
//    F = |                   |
//        | G(s)     0   0    |   rows: y'          G: NO x NS
//
// Every entry of A, B, G has a compile-time kind: structural zero, one, constant, or
// variable.  Multiplying by structural zeros/ones is exact, so pruning them changes nothing
// but the operation count.  The cost is  sum_i w_i (xu_i - r_i)^2  with the target on y.
#pragma once

#include "ilqr_common.cuh"

namespace crl {

enum Kind : int { KZ = 0, K1 = 1, KC = 2, KV = 3 };

// Euler-integrated joint: theta' = theta + h*omega ; omega' = omega + h*(acceleration)
constexpr double kH = 0.01;

// -----------------------------------------------------------------------------------------
// Kinematic chains (TwoLink J=2, ThreeLink J=3): accelerations are the controls.
template <int J>
struct KinematicArm {
    static constexpr int NS = 2 * J, NO = 2, N = NS + NO, M = J, D = N + M, P = 2;
    static constexpr bool kBarrier = false;
    static constexpr bool kKinematic = true;       // joint states are double integrators: no trigonometry on the rollout's recurrence
    // order in which sympy prints the per-action terms of a sum (symbol names sort as strings: x_u10 < x_u8 < x_u9)
    CRL_HD static constexpr int action_order(int k) { return J == 3 ? (k == 0 ? 2 : k - 1) : k; }

    CRL_HD static constexpr int akind(int i, int j) { return i == j ? K1 : ((i % 2 == 0 && j == i + 1) ? KC : KZ); }
    CRL_HD static constexpr double aconst(int, int) { return kH; }
    CRL_HD static constexpr int bkind(int i, int a) { return (i % 2 == 1 && a == i / 2) ? KC : KZ; }
    CRL_HD static constexpr double bconst(int, int) { return kH; }
    CRL_HD static constexpr int gkind(int, int j) { return (j % 2 == 0) ? KV : KZ; }

    CRL_HD static constexpr double link(int j) { return j == 0 ? 1.0 : 2.0; }       // l1 = 1, l2 = l3 = 2
    CRL_HD static constexpr double weight(int i) {                                  // cost diag
        return i < NS ? ((i % 2) ? 10.0 : 0.0) : (i < N ? 1e4 : 1.0);
    }
    CRL_HD static constexpr double u_lo(int) { return -100.0; }
    CRL_HD static constexpr double u_hi(int) { return 100.0; }

    // per-time-step table of the cooperative backward pass (ilqr_coop.cuh): the variable Jacobian entries
    // G[0][0], G[0][2], ..., G[1][0], G[1][2], ...
    static constexpr int kCoopTab = 2 * J;
    template <class R>
    CRL_HD static void coop_table(const R* xu, R* tab) {
        R ang[J], sn[J], cs[J], G[2][NS];
        angles<R>(xu, ang);
        sincos_batch<J, R>(ang, sn, cs);
        jac_trig<R>(xu, sn, cs, nullptr, nullptr, G);
        CRL_UNROLL for (int i = 0; i < J; ++i) {
            tab[i] = G[0][2 * i];
            tab[J + i] = G[1][2 * i];
        }
    }

    // angles whose sin/cos one time step needs: the cumulative joint angles
    static constexpr int NANG = J;
    template <class R>
    CRL_HD static void angles(const R* xu, R* ang) {
        ang[0] = xu[0];
        CRL_UNROLL for (int j = 1; j < J; ++j) ang[j] = ang[j - 1] + xu[2 * j];
    }

    // x_{t+1} = f(x_t, u_t); xu = (s, y, u); sn/cs = sin/cos of angles(xu)
    template <class R>
    CRL_HD static void step_trig(const R* xu, const R* sn, const R* cs, R* xn) {
        CRL_UNROLL for (int j = 0; j < J; ++j) {
            xn[2 * j] = xu[2 * j] + R(kH) * xu[2 * j + 1];
            xn[2 * j + 1] = xu[2 * j + 1] + R(kH) * xu[N + j];
        }
        R ex = sn[0], ey = cs[0];                            // l1 = 1: sympy prints sin(x0), not 1*sin(x0)
        CRL_UNROLL for (int j = 1; j < J; ++j) {
            ex = ex + R(link(j)) * sn[j];
            ey = ey + R(link(j)) * cs[j];
        }
        xn[NS] = ex;
        xn[NS + 1] = ey;
    }

    // Variable Jacobian entries: G[0][2i] = sum_{j>=i} l_j cos(a_j), G[1][2i] = -sum_{j>=i} l_j sin(a_j)
    template <class R>
    CRL_HD static void jac_trig(const R* xu, const R* sn, const R* cs, R (*A)[NS], R (*B)[M], R (*G)[NS]) {
        (void)xu; (void)A; (void)B;
        CRL_UNROLL for (int i = 0; i < J; ++i) {
            // left-to-right sums as printed: cos(a_i)*l_i + l_{i+1} cos(a_{i+1}) + ...
            R gc = (i == 0) ? cs[0] : R(link(i)) * cs[i];
            R gs = (i == 0) ? -sn[0] : -(R(link(i)) * sn[i]);
            CRL_UNROLL for (int j = i + 1; j < J; ++j) {
                gc = gc + R(link(j)) * cs[j];
                gs = gs - R(link(j)) * sn[j];
            }
            G[0][2 * i] = gc;
            G[1][2 * i] = gs;
        }
    }
};

// step / jac with the trigonometry evaluated in place (single-candidate callers)
template <class Mdl, class R>
CRL_HD void model_step(const R* xu, R* xn) {
    R ang[Mdl::NANG], sn[Mdl::NANG], cs[Mdl::NANG];
    Mdl::template angles<R>(xu, ang);
    sincos_batch<Mdl::NANG, R>(ang, sn, cs);
    Mdl::template step_trig<R>(xu, sn, cs, xn);
}
template <class Mdl, class R>
CRL_HD void model_jac(const R* xu, R (*A)[Mdl::NS], R (*B)[Mdl::M], R (*G)[Mdl::NS]) {
    R ang[Mdl::NANG], sn[Mdl::NANG], cs[Mdl::NANG];
    Mdl::template angles<R>(xu, ang);
    sincos_batch<Mdl::NANG, R>(ang, sn, cs);
    Mdl::template jac_trig<R>(xu, sn, cs, A, B, G);
}

struct TwoLink : KinematicArm<2> {
    static constexpr int ID = 0;
    // 10*x1^2 + 10*x3^2 + 1*x6^2 + 1*x7^2 + (-1e4 p0 + 1e4 x4)(-p0 + x4) + (-1e4 p1 + 1e4 x5)(-p1 + x5)
    CRL_HD static double cost(const double* x, const double* p) {
        double J = 10.0 * (x[1] * x[1]);
        J = J + 10.0 * (x[3] * x[3]);
        J = J + 1.0 * (x[6] * x[6]);
        J = J + 1.0 * (x[7] * x[7]);
        J = J + (-10000.0 * p[0] + 10000.0 * x[4]) * (-p[0] + x[4]);
        J = J + (-10000.0 * p[1] + 10000.0 * x[5]) * (-p[1] + x[5]);
        return J;
    }
};

struct ThreeLink : KinematicArm<3> {
    static constexpr int ID = 1;
    // term order as printed by sympy: x1, x10, x3, x5, x8, x9, then the two target terms
    CRL_HD static double cost(const double* x, const double* p) {
        double J = 10.0 * (x[1] * x[1]);
        J = J + 1.0 * (x[10] * x[10]);
        J = J + 10.0 * (x[3] * x[3]);
        J = J + 10.0 * (x[5] * x[5]);
        J = J + 1.0 * (x[8] * x[8]);
        J = J + 1.0 * (x[9] * x[9]);
        J = J + (-10000.0 * p[0] + 10000.0 * x[6]) * (-p[0] + x[6]);
        J = J + (-10000.0 * p[1] + 10000.0 * x[7]) * (-p[1] + x[7]);
        return J;
    }
};

// -----------------------------------------------------------------------------------------
// Rigid two-link arm with gravity.  delta = th1 - th2, c = cos(delta), s = sin(delta),
//   den = 4 c^2 - 56/9 (= -det H),  R1 = tau1 - 2 w2^2 s + 24.5 sin th1,  R2 = tau2 + 2 w1^2 s + 19.6 sin th2
//   h*ddth1 = (-k1 R1 + 0.02 c R2) / den ,  h*ddth2 = (0.02 c R1 - k2 R2) / den
// with the constants exactly as the reference's generated code prints them.
struct RoboticArm {
    static constexpr int ID = 2;
    static constexpr int NS = 4, NO = 2, N = 6, M = 2, D = 8, P = 2;
    static constexpr bool kBarrier = false;
    static constexpr bool kKinematic = false;
    CRL_HD static constexpr int action_order(int k) { return k; }
    static constexpr double k1 = 0.0266666666666667, k2 = 0.0233333333333333, kden = 6.22222222222222;

    CRL_HD static constexpr int akind(int i, int j) {
        return (i % 2 == 1) ? KV : (i == j ? K1 : (j == i + 1 ? KC : KZ));
    }
    CRL_HD static constexpr double aconst(int, int) { return kH; }
    CRL_HD static constexpr int bkind(int i, int) { return (i % 2 == 1) ? KV : KZ; }
    CRL_HD static constexpr double bconst(int, int) { return 0.0; }
    CRL_HD static constexpr int gkind(int, int j) { return (j % 2 == 0) ? KV : KZ; }
    CRL_HD static constexpr double weight(int i) { return i < NS ? ((i % 2) ? 10.0 : 0.0) : (i < N ? 1e4 : 0.01); }
    // is_with_constraints=True (the default, robotic_arm_tracking.py:59-62); the unconstrained
    // variant is served by passing infinite bounds at run time (see SolveParams::u_lo/u_hi).
    CRL_HD static constexpr double u_lo(int) { return -500.0; }
    CRL_HD static constexpr double u_hi(int) { return 500.0; }

    template <class R>
    struct Terms { R s, c, s1, c1, s2, c2, id, R1, R2, a1, a2; };

    static constexpr int NANG = 3;                           // delta, th1, th2
    template <class R>
    CRL_HD static void angles(const R* xu, R* ang) {
        ang[0] = xu[0] - xu[2];
        ang[1] = xu[0];
        ang[2] = xu[2];
    }

    template <class R>
    CRL_HD static void terms(const R* xu, const R* sn, const R* cs, Terms<R>& t) {
        t.s = sn[0]; t.c = cs[0]; t.s1 = sn[1]; t.c1 = cs[1]; t.s2 = sn[2]; t.c2 = cs[2];
        R den = R(4.0) * (t.c * t.c) - R(kden);
        t.id = R(1.0) / den;
        t.R1 = xu[6] + (R(-2.0) * (xu[3] * xu[3]) * t.s + R(24.5) * t.s1);
        t.R2 = xu[7] + (R(2.0) * (xu[1] * xu[1]) * t.s + R(19.6) * t.s2);
        R hc = R(0.02) * t.c;
        t.a1 = (hc * t.R2 - R(k1) * t.R1) * t.id;
        t.a2 = (hc * t.R1 - R(k2) * t.R2) * t.id;
    }

    template <class R>
    CRL_HD static void step_trig(const R* xu, const R* sn, const R* cs, R* xn) {
        Terms<R> t;
        terms(xu, sn, cs, t);
        xn[0] = xu[0] + R(kH) * xu[1];
        xn[1] = xu[1] + t.a1;
        xn[2] = xu[2] + R(kH) * xu[3];
        xn[3] = xu[3] + t.a2;
        xn[4] = t.s1 + R(2.0) * t.s2;
        xn[5] = t.c1 + R(2.0) * t.c2;
    }

    template <class R>
    CRL_HD static void jac_trig(const R* xu, const R* sn, const R* cs, R (*A)[NS], R (*B)[M], R (*G)[NS]) {
        Terms<R> t;
        terms(xu, sn, cs, t);
        const R w1 = xu[1], w2 = xu[3], s = t.s, c = t.c, id = t.id;
        const R hc = R(0.02) * c, hs = R(0.02) * s;
        // d den / d th1 = -8 c s ; d den / d th2 = +8 c s
        const R dden = R(-8.0) * (c * s);
        // partials of R1, R2
        const R w2sq2c = R(2.0) * (w2 * w2) * c, w1sq2c = R(2.0) * (w1 * w1) * c;
        const R dR1_0 = R(24.5) * t.c1 - w2sq2c, dR1_2 = w2sq2c, dR1_3 = R(-4.0) * (w2 * s);
        const R dR2_0 = w1sq2c, dR2_2 = R(19.6) * t.c2 - w1sq2c, dR2_1 = R(4.0) * (w1 * s);
        // N1 = -k1 R1 + 0.02 c R2 ; N2 = 0.02 c R1 - k2 R2 ; a = N / den ; da = (dN - a dden) / den
        //   d c / d th1 = -s ; d c / d th2 = +s
        const R dN1_0 = (hc * dR2_0 - hs * t.R2) - R(k1) * dR1_0;
        const R dN1_2 = (hc * dR2_2 + hs * t.R2) - R(k1) * dR1_2;
        const R dN2_0 = (hc * dR1_0 - hs * t.R1) - R(k2) * dR2_0;
        const R dN2_2 = (hc * dR1_2 + hs * t.R1) - R(k2) * dR2_2;
        A[1][0] = (dN1_0 - t.a1 * dden) * id;
        A[1][1] = R(1.0) + (hc * dR2_1) * id;
        A[1][2] = (dN1_2 + t.a1 * dden) * id;
        A[1][3] = (R(-k1) * dR1_3) * id;
        A[3][0] = (dN2_0 - t.a2 * dden) * id;
        A[3][1] = (R(-k2) * dR2_1) * id;
        A[3][2] = (dN2_2 + t.a2 * dden) * id;
        A[3][3] = R(1.0) + (hc * dR1_3) * id;
        B[1][0] = R(-k1) * id;
        B[1][1] = hc * id;
        B[3][0] = hc * id;
        B[3][1] = R(-k2) * id;
        G[0][0] = t.c1;
        G[0][2] = R(2.0) * t.c2;
        G[1][0] = -t.s1;
        G[1][2] = R(-2.0) * t.s2;
    }

    // per-time-step table of the cooperative backward pass: A[1][0..3], A[3][0..3], B[1][0..1], B[3][0..1],
    // G[0][0], G[0][2], G[1][0], G[1][2]
    static constexpr int kCoopTab = 16;
    template <class R>
    CRL_HD static void coop_table(const R* xu, R* tab) {
        R ang[NANG], sn[NANG], cs[NANG], A[NS][NS], B[NS][M], G[NO][NS];
        angles<R>(xu, ang);
        sincos_batch<NANG, R>(ang, sn, cs);
        jac_trig<R>(xu, sn, cs, A, B, G);
        CRL_UNROLL for (int j = 0; j < NS; ++j) { tab[j] = A[1][j]; tab[NS + j] = A[3][j]; }
        CRL_UNROLL for (int a = 0; a < M; ++a) { tab[2 * NS + a] = B[1][a]; tab[2 * NS + M + a] = B[3][a]; }
        tab[12] = G[0][0]; tab[13] = G[0][2]; tab[14] = G[1][0]; tab[15] = G[1][2];
    }

    CRL_HD static double cost(const double* x, const double* p) {
        double J = 10.0 * (x[1] * x[1]);
        J = J + 10.0 * (x[3] * x[3]);
        J = J + 0.01 * (x[6] * x[6]);
        J = J + 0.01 * (x[7] * x[7]);
        J = J + (-10000.0 * p[0] + 10000.0 * x[4]) * (-p[0] + x[4]);
        J = J + (-10000.0 * p[1] + 10000.0 * x[5]) * (-p[1] + x[5]);
        return J;
    }
};

// ---- log-barrier variant of a model (LogBarrieriLQR, log_barrier_ilqr.py:86-101) -----------
// Every finite bound of the scenario adds  (-1/t) log(-(lo - u_a))  /  (-1/t) log(-(u_a - hi))  to the
// running cost (in the three arm scenarios only the actions are bounded).  As in the reference the
// barrier parameter t rides in the cost-parameter row: p = (target_x, target_y, t, t0), where t0 is
// the t the reference's FIRST backward pass of a solve still sees - its C, c were evaluated by the
// previous forward pass, before the outer loop moved t on (log_barrier_ilqr.py:121-131).
// Expression forms as sympy prints them for the reference's lambdified cost / gradient / Hessian:
//   l    = l_base - log(hi - u)/t ... - log(u + (-lo))/t ...        (all upper bounds first, actions in
//                                                                    printed order)
//   l_u  = 2 w u - 1/(t (u + (-lo))) + 1/(t (hi - u))
//   l_uu = 2 w + 1/(t (u + (-lo))**2) + 1/(t (hi - u)**2)   (+ 1e-100)
template <class Base>
struct Barrier : Base {
    static constexpr bool kBarrier = true;
    static constexpr int ID = Base::ID + 3;
    static constexpr int P = Base::P + 2;
    static constexpr int kT = Base::P;                       // index of t in the parameter row; t0 follows
    CRL_HD static double cost(const double* x, const double* p) {
        double J = Base::cost(x, p);
        const double t = p[kT];
        double arg[2 * Base::M];
        CRL_UNROLL for (int k = 0; k < Base::M; ++k) {
            const int a = Base::action_order(k);
            arg[k] = Base::u_hi(a) - x[Base::N + a];
            arg[Base::M + k] = x[Base::N + a] + (-Base::u_lo(a));
        }
        // The 2 m terms are evaluated by a loop that STAYS A LOOP: unrolled, every log + division is ~85 instructions of
        // straight-line code, i.e. 2,000 instructions per time step of a four-candidate line-search pass - more than the
        // 32 KB instruction cache behind the schedulers' ~6 KB L0 holds next to the rest of the step (ncu of the
        // LogBarrieriLQR solve: 1.4 `no_instruction` stall cycles per issued instruction, 54 % of the stall samples on
        // these two lines).  The rolled body fits the L0; same operations in the same order, so nothing changes bit-wise.
        CRL_KEEP_ROLLED for (int k = 0; k < 2 * Base::M; ++k) J = J - ::log(arg[k]) / t;
        return J;
    }
};

// ---- cost derivatives shared by all three (diagonal quadratic tracking cost) -------------
// grad_i = 2 w_i x_i                       for untargeted entries (printed as 20.0*x, 2.0*x, 0.02*x, or 0)
// grad_i = -2 w_i p + 2 w_i x_i            for the two end-effector entries (printed -20000.0*p0 + 20000.0*x)
// hess_ii = 2 w_i + 1e-100                 (ilqr_obj_fun.py:51: "+1e-100*eye" keeps every entry float)
// Barrier models add the barrier terms to the action entries (p[kT] = t).
template <class Mdl, class R>
CRL_HD R cost_grad(int i, R xi, const R* p) {
    const R w2 = R(2.0 * Mdl::weight(i));
    if (i >= Mdl::NS && i < Mdl::N) return (-w2) * p[i - Mdl::NS] + w2 * xi;
    if constexpr (Mdl::kBarrier) {
        if (i >= Mdl::N) {
            const int a = i - Mdl::N;
            const R t = p[Mdl::kT];
            return (w2 * xi - R(1) / (t * (xi + R(-Mdl::u_lo(a))))) + R(1) / (t * (R(Mdl::u_hi(a)) - xi));
        }
    }
    return (Mdl::weight(i) == 0.0) ? R(0.0) : w2 * xi;
}
template <class Mdl>
CRL_HD constexpr double cost_hess(int i) { return 2.0 * Mdl::weight(i) + 1e-100; }
// Hessian entry of action a: a constant for the plain models
template <class Mdl, class R>
CRL_HD R cost_hess_u(int a, R u, const R* p) {
    if constexpr (Mdl::kBarrier) {
        const R t = p[Mdl::kT];
        const R lo = u + R(-Mdl::u_lo(a)), hi = R(Mdl::u_hi(a)) - u;
        return ((R(2.0 * Mdl::weight(Mdl::N + a)) + R(1) / (t * (lo * lo))) + R(1) / (t * (hi * hi))) + R(1e-100);
    } else {
        (void)u; (void)p;
        return R(cost_hess<Mdl>(Mdl::N + a));
    }
}

}  // namespace crl
