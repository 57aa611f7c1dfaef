// Warp-cooperative iLQR iteration: one WARP = one trajectory (device only).
//
// The solve kernel's main mode maps one lane to one trajectory: best throughput, but one
// iteration is a ~10^5-instruction dependent chain per lane, so a single slow trajectory
// (the iteration counts are heavy tailed: median 13, p99 ~520, cap 1000) holds a whole warp
// for ~1 ms per iteration.  Once the work queue is empty, warps therefore hand their remaining
// trajectories to a global straggler queue and every warp of the grid becomes a "finisher":
// it takes one straggler at a time and iterates it with all 32 lanes,
//   * backward pass (ilqr_wrapper.py:232-256): every matrix element of the Riccati step is owned
//     by one lane, operands are exchanged through shared memory (4 warp barriers per time step),
//     the 3x3 / 2x2 pivoted LU of Quu is factored by the lanes that then apply it to one
//     right-hand-side column each; several trajectories can be stepped together (coop_backward<NT>);
//   * line search (ilqr_wrapper.py:74-100): one step size per lane, all candidates at once.
// Per element the arithmetic is the SAME operation sequence as the lane-per-trajectory code in
// ilqr_core.cuh (riccati_step / rollout_closed_f), so a trajectory's result does not depend on
// which path finished it (tests/test_gpu_solve_parity.py::test_coop_matches_lane_mode checks
// bit equality).
//
// The element-per-lane pass (coop_backward) exists for the kinematic chains (TwoLink, ThreeLink:
// constant A, B; variable G only); the column-group passes for all three models.
#pragma once

#include "ilqr_core.cuh"

namespace crl {

template <class Mdl> struct CoopSupport { static constexpr bool value = false; };
template <> struct CoopSupport<TwoLink> { static constexpr bool value = true; };
template <> struct CoopSupport<ThreeLink> { static constexpr bool value = true; };
template <> struct CoopSupport<RoboticArm> { static constexpr bool value = true; };
// the log-barrier variants run on the column-group passes (the action entries of the cost gradient / Hessian
// are the only difference: coop_cost_u below); the element-per-lane and pair-group passes are plain-cost only
template <class Base> struct CoopSupport<Barrier<Base>> { static constexpr bool value = CoopSupport<Base>::value; };

// kinematic chain (constant A, B; colgroup_backward) or rigid arm (dense variable rows; colgroup_backward_arm)
template <class Mdl> struct IsChain { static constexpr bool value = true; };
template <> struct IsChain<RoboticArm> { static constexpr bool value = false; };
template <> struct IsChain<Barrier<RoboticArm>> { static constexpr bool value = false; };

// Action entries of the cost gradient and Hessian at one time step: constants for the plain costs; for the
// barrier costs they depend on u and on the barrier parameter in the parameter row (tb = p[kT], or the stale
// p[kT + 1] in the first backward pass of a solve).
template <class Mdl, class R>
__device__ __forceinline__ void coop_cost_u(const R (&u)[Mdl::M], R tb, R (&grad)[Mdl::M], R (&hess)[Mdl::M]) {
    R pp[Mdl::P];
    CRL_UNROLL for (int i = 0; i < Mdl::P; ++i) pp[i] = R(0);
    if constexpr (Mdl::kBarrier) pp[Mdl::kT] = tb;
    CRL_UNROLL for (int a = 0; a < Mdl::M; ++a) {
        grad[a] = cost_grad<Mdl, R>(Mdl::N + a, u[a], pp);
        hess[a] = cost_hess_u<Mdl, R>(a, u[a], pp);
    }
}
template <class Mdl, class R>
__device__ __forceinline__ R coop_barrier_t(const ParamRef<R>& prm, int t, bool first) {
    if constexpr (Mdl::kBarrier) return prm.get(t, first ? Mdl::kT + 1 : Mdl::kT);
    else return R(0);
}

// Shared-memory layout (in reals) of the cooperative backward pass of ONE trajectory.
template <class Mdl>
struct CoopLayout {
    static constexpr int NS = Mdl::NS, M = Mdl::M, J = Mdl::M, NQ = Mdl::NS + 1;
    static constexpr int V = 0;                       // NS x NS value-function Hessian, both triangles
    static constexpr int VS = V + NS * NS;            // NS      value-function gradient
    static constexpr int VY = VS + NS;                // 2       gradient w.r.t. the output entries
    static constexpr int VYY = VY + 2;                // 2       (diagonal) Hessian w.r.t. the output entries
    static constexpr int ZEND = VYY + 2;              // [0, ZEND) is zeroed at the start of a pass
    static constexpr int ONE = ZEND;                  // the constant 1.0
    static constexpr int GCUR = ONE + 1;              // 2 x J   variable Jacobian entries G[y][2i] of the current step
    static constexpr int MA = GCUR + 2 * J;           // NS x NS A^T V
    static constexpr int MB = MA + NS * NS;           // M x NS  B^T V
    static constexpr int QX = MB + M * NS;            // NS x NQ [Qss (upper triangle) | qs]
    static constexpr int QU = QX + NS * NQ;           // M x NQ  [Qus | qu]
    static constexpr int KK = QU + M * NQ;            // M x NQ  [K | k]
    static constexpr int QUU = KK + M * NQ;           // M x M   Quu (before factorisation)
    static constexpr int DUMMY = QUU + M * M;         // where lanes without an element of a stage write
    static constexpr int SZ = (DUMMY + 2) & ~1;       // reals per trajectory
    static constexpr int GROW = 2 * J;                // reals per time step of the global G table
};

// i <= j < n enumerated row by row: element e of the upper triangle of an n-column matrix
__device__ __forceinline__ void coop_pair(int e, int n, int& i, int& j) {
    i = 0;
    int len = n;
    while (e >= len) {
        e -= len;
        ++i;
        --len;
    }
    j = i + e;
}

enum CoopKind : int { CK_PLAIN = 0, CK_HPLUS = 1, CK_HONLY = 2 };

// out = S[o1]            (PLAIN)
//     = S[o0]*h + S[o1]  (HPLUS)
//     = S[o1]*h          (HONLY; o0 == o1)
template <class R>
__device__ __forceinline__ R coop_hsum(const R* S, int o0, int o1, int kind, R h) {
    const R m1 = S[o1];
    const R p = S[o0] * h;
    return kind == CK_PLAIN ? m1 : (kind == CK_HPLUS ? p + m1 : p);
}

// One trajectory of a cooperative backward pass: its current trajectory (rows of D reals, rs reals apart), cost
// parameters, outputs K [T][M*NS], k [T][M] (row-major) and a scratch table [T][2J] for the variable
// Jacobian entries - all global.  on = false: the slot is empty; its pointers must still be valid
// scratch (the pass runs branch-free over all NT slots and ignores what an empty one produces).
template <class R>
struct CoopTraj {
    const R* X;                     // element (t, c) at X[t * rs + c]
    int rs;                         // row stride in reals: D = row-major; GL * DP = a candidate of the engine's line-search buffers
    ParamRef<R> prm;
    R* K;
    R* k;
    R* G;
    bool on;
    bool first;                     // barrier costs: this is the first backward pass of the solve (stale t, see Barrier)
};

// Tables of the variable Jacobian entries of NQ trajectories, all time steps: lanes take time steps
// lane, lane + 32, ...; the rows of the next round are loaded before the current round is evaluated
// (the trajectories sit in DRAM / L2: without the overlap every round waits a full memory latency).
template <class Mdl, class R, int NQ>
__device__ __forceinline__ void coop_gtables(const CoopTraj<R> (&tr)[NQ], int T, int lane) {
    constexpr int J = Mdl::M, D = Mdl::D, NS = Mdl::NS;
    R nxt[NQ][J];
    {
        const int t = lane < T ? lane : T - 1;
        CRL_UNROLL for (int q = 0; q < NQ; ++q)
            CRL_UNROLL for (int j = 0; j < J; ++j) nxt[q][j] = tr[q].X[(size_t)t * tr[q].rs + 2 * j];
    }
    for (int t = lane; t < T; t += 32) {
        R cur[NQ][J];
        CRL_UNROLL for (int q = 0; q < NQ; ++q)
            CRL_UNROLL for (int j = 0; j < J; ++j) cur[q][j] = nxt[q][j];
        {
            const int tn = t + 32 < T ? t + 32 : T - 1;
            CRL_UNROLL for (int q = 0; q < NQ; ++q)
                CRL_UNROLL for (int j = 0; j < J; ++j) nxt[q][j] = tr[q].X[(size_t)tn * tr[q].rs + 2 * j];
        }
        CRL_UNROLL for (int q = 0; q < NQ; ++q) {
            R xu[D], ang[J], sn[J], cs[J], G[2][NS];
            CRL_UNROLL for (int j = 0; j < J; ++j) xu[2 * j] = cur[q][j];
            Mdl::template angles<R>(xu, ang);
            sincos_batch<J, R>(ang, sn, cs);
            Mdl::template jac_trig<R>(xu, sn, cs, nullptr, nullptr, G);
            CRL_UNROLL for (int i = 0; i < J; ++i) {
                tr[q].G[(size_t)t * (2 * J) + i] = G[0][2 * i];
                tr[q].G[(size_t)t * (2 * J) + J + i] = G[1][2 * i];
            }
        }
    }
}

// Backward pass of NT trajectories by one warp, time step by time step for all of them together:
// the NT Riccati recursions are independent dependency chains, so interleaving them fills the issue
// slots a single chain leaves empty (one chain alone uses ~1/4 of them).  S = this warp's
// shared-memory area, NT * CoopLayout::SZ reals.
template <class Mdl, class R, int NT>
__device__ void coop_backward(const CoopTraj<R> (&tr)[NT], int T, R* S, int lane) {
    using L = CoopLayout<Mdl>;
    constexpr int NS = Mdl::NS, M = Mdl::M, J = Mdl::M, N = Mdl::N, NQ = NS + 1;
    constexpr int N1 = NS * NS + M * NS, R1 = (N1 + 31) / 32;                     // stage 1: MA, MB
    constexpr int NP = NS * (NS + 1) / 2;                                          // upper-triangle pairs
    constexpr int N2 = NP + M * NS + NS + M, R2 = (N2 + 31) / 32;                 // stage 2: Qss, Qus, qs, qu
    constexpr int N5 = NP + NS;                                                    // stage 5: V, v
    static_assert(N5 + 2 <= 32 && NT * NQ <= 32 && NT * 2 * J <= 32 && M * NQ <= 32,
                  "element-to-lane mapping needs <= 32 lanes per stage");
    const R h = R(kH);

    // ---- G tables
    static_assert(L::GROW == 2 * J, "table row");
    coop_gtables<Mdl, R, NT>(tr, T, lane);
    CRL_UNROLL for (int q = 0; q < NT; ++q) {
        for (int e = lane; e < L::ZEND; e += 32) S[q * L::SZ + e] = R(0);
        if (lane == 0) S[q * L::SZ + L::ONE] = R(1);
    }
    __syncwarp();

    // ---- per-lane element descriptors (the same for every trajectory)
    int a_o0[R1], a_o1[R1], a_dst[R1], a_kind[R1];
    bool a_ok[R1];
    CRL_UNROLL for (int r = 0; r < R1; ++r) {
        const int e = lane + 32 * r;
        a_ok[r] = e < N1;
        if (e < NS * NS) {
            const int i = e / NS, k = e % NS, odd = i & 1;
            a_o0[r] = L::V + (i - odd) * NS + k;
            a_o1[r] = L::V + i * NS + k;
            a_kind[r] = odd ? CK_HPLUS : CK_PLAIN;
            a_dst[r] = L::MA + e;
        } else {
            const int e2 = a_ok[r] ? e - NS * NS : 0;
            const int a = e2 / NS, k = e2 % NS;
            a_o0[r] = a_o1[r] = L::V + (2 * a + 1) * NS + k;
            a_kind[r] = CK_HONLY;
            a_dst[r] = a_ok[r] ? L::MB + e2 : L::DUMMY;
        }
    }
    int b_o0[R2], b_o1[R2], b_kind[R2], b_dst[R2], b_xcol[R2];
    int b_gai[R2], b_gas[R2], b_gaj[R2], b_gbi[R2], b_gbs[R2], b_gbj[R2];
    bool b_ok[R2], b_g[R2], b_add[R2], b_const[R2];
    R b_cbase[R2], b_cmul[R2];
    CRL_UNROLL for (int r = 0; r < R2; ++r) {
        const int e = lane + 32 * r;
        b_ok[r] = e < N2;
        b_g[r] = false; b_add[r] = false; b_const[r] = true;
        b_cbase[r] = R(0); b_cmul[r] = R(0); b_xcol[r] = -1;
        b_gai[r] = b_gas[r] = b_gaj[r] = b_gbi[r] = b_gbs[r] = b_gbj[r] = L::ONE;
        b_o0[r] = b_o1[r] = L::ONE; b_kind[r] = CK_PLAIN; b_dst[r] = L::DUMMY;
        if (e < NP) {                                             // Qss(i, j), i <= j
            int i, j;
            coop_pair(e, NS, i, j);
            b_kind[r] = (j & 1) ? CK_HPLUS : CK_PLAIN;
            b_o1[r] = L::MA + i * NS + j;
            b_o0[r] = (j & 1) ? b_o1[r] - 1 : b_o1[r];
            b_g[r] = !(i & 1) && !(j & 1);
            b_gai[r] = L::GCUR + i / 2; b_gas[r] = L::VYY; b_gaj[r] = L::GCUR + j / 2;
            b_gbi[r] = L::GCUR + J + i / 2; b_gbs[r] = L::VYY + 1; b_gbj[r] = L::GCUR + J + j / 2;
            b_add[r] = (i == j);
            b_cbase[r] = (i & 1) ? R(cost_hess<Mdl>(1)) : R(cost_hess<Mdl>(0));
            b_dst[r] = L::QX + i * NQ + j;
        } else if (e < NP + M * NS) {                             // Qus(a, j)
            const int e2 = e - NP, a = e2 / NS, j = e2 % NS;
            b_kind[r] = (j & 1) ? CK_HPLUS : CK_PLAIN;
            b_o1[r] = L::MB + a * NS + j;
            b_o0[r] = (j & 1) ? b_o1[r] - 1 : b_o1[r];
            b_dst[r] = L::QU + a * NQ + j;
        } else if (e < NP + M * NS + NS) {                        // qs(i)
            const int i = e - NP - M * NS;
            b_kind[r] = (i & 1) ? CK_HPLUS : CK_PLAIN;
            b_o1[r] = L::VS + i;
            b_o0[r] = (i & 1) ? b_o1[r] - 1 : b_o1[r];
            b_g[r] = !(i & 1);
            b_gai[r] = L::VY; b_gas[r] = L::ONE; b_gaj[r] = L::GCUR + i / 2;
            b_gbi[r] = L::VY + 1; b_gbs[r] = L::ONE; b_gbj[r] = L::GCUR + J + i / 2;
            b_add[r] = true;
            b_const[r] = !(i & 1);                                // weight 0 on the angles: gradient is the literal 0.0
            b_cmul[r] = R(2.0 * Mdl::weight(1));
            b_xcol[r] = (i & 1) ? i : -1;
            b_dst[r] = L::QX + i * NQ + NS;
        } else if (e < N2) {                                      // qu(a)
            const int a = e - NP - M * NS - NS;
            b_kind[r] = CK_HONLY;
            b_o0[r] = b_o1[r] = L::VS + 2 * a + 1;
            b_add[r] = true;
            b_const[r] = false;
            b_cmul[r] = R(2.0 * Mdl::weight(N));
            b_xcol[r] = N + a;
            b_dst[r] = L::QU + a * NQ + NS;
        }
    }
    // stage 5: element (i, j), i <= j <= NS (j == NS: the gradient v)
    int c_i = 0, c_j = 0;
    const bool c_ok = lane < N5;
    if (c_ok) coop_pair(lane, NQ, c_i, c_j);
    const int c_base = L::QX + c_i * NQ + c_j;
    const int vy_y = lane - N5;                                   // lanes N5, N5+1 update vy, vyy
    const bool vy_ok = vy_y >= 0 && vy_y < 2;
    // where this lane's stage-5 results go: V(i,j) and its mirror / v(i); vy and vyy; nothing (dummy)
    const int c_dst1 = c_ok ? ((c_j < NS) ? L::V + c_i * NS + c_j : L::VS + c_i) : (vy_ok ? L::VY + vy_y : L::DUMMY);
    const int c_dst2 = c_ok ? ((c_j < NS) ? L::V + c_j * NS + c_i : c_dst1) : (vy_ok ? L::VYY + vy_y : L::DUMMY);
    // lanes < M*NQ store one entry of K / k per step; the others repeat lane 0's store (same value)
    const bool st_ok = lane < M * NQ;
    const int st_a = st_ok ? lane / NQ : 0, st_j = st_ok ? lane % NQ : 0;
    const size_t st_off = st_j < NS ? (size_t)(st_a * NS + st_j) : (size_t)st_a, st_stride = st_j < NS ? M * NS : M;
    R* st_ptr[NT];
    CRL_UNROLL for (int q = 0; q < NT; ++q) st_ptr[q] = (st_j < NS ? tr[q].K : tr[q].k) + st_off + (size_t)(T - 1) * st_stride;
    // stage 3: lane (q3, r3) solves right-hand-side column r3 of trajectory q3 (and factors its Quu)
    const bool lu_ok = lane < NT * NQ;
    const int q3 = lu_ok ? lane / NQ : NT - 1, r3 = lu_ok ? lane % NQ : NQ - 1;
    R* const S3 = S + q3 * L::SZ;
    // G rows: lane (qg, g) moves entry g of trajectory qg's next row from the global table to GCUR
    const bool gc_ok = lane < NT * 2 * J;
    const int qg = gc_ok ? lane / (2 * J) : 0, gg = gc_ok ? lane % (2 * J) : 0;
    const R* Gsrc = tr[0].G;
    CRL_UNROLL for (int q = 0; q < NT; ++q)
        if (q == qg) Gsrc = tr[q].G;
    Gsrc += gg;
    const int g_dst = gc_ok ? qg * L::SZ + L::GCUR + gg : L::DUMMY;

    // row values needed by this lane (stage 2: x_i / u_a; stage 5: output entry and its target), one step ahead;
    // lanes without such a value read column 0 and ignore it (no branches in the time loop)
    int xcol[R2];
    CRL_UNROLL for (int r = 0; r < R2; ++r) xcol[r] = b_xcol[r] >= 0 ? b_xcol[r] : 0;
    const int vcol = NS + (vy_ok ? vy_y : 0), pcol = vy_ok ? vy_y : 0;
    R xin[NT][R2], xin_n[NT][R2], xv[NT], xv_n[NT], pv[NT], pv_n[NT], g_n;
    CRL_UNROLL for (int q = 0; q < NT; ++q) {
        CRL_UNROLL for (int r = 0; r < R2; ++r) xin_n[q][r] = tr[q].X[(size_t)(T - 1) * tr[q].rs + xcol[r]];
        xv_n[q] = tr[q].X[(size_t)(T - 1) * tr[q].rs + vcol];
        pv_n[q] = tr[q].prm.get(T - 1, pcol);
    }
    g_n = Gsrc[(size_t)(T - 1) * L::GROW];                       // written above by other lanes: after the barrier

    for (int t = T - 1; t >= 0; --t) {
        const R g_cur = g_n;
        CRL_UNROLL for (int q = 0; q < NT; ++q) {
            CRL_UNROLL for (int r = 0; r < R2; ++r) xin[q][r] = xin_n[q][r];
            xv[q] = xv_n[q];
            pv[q] = pv_n[q];
        }
        {
            const int tn = t > 0 ? t - 1 : 0;
            CRL_UNROLL for (int q = 0; q < NT; ++q) {
                CRL_UNROLL for (int r = 0; r < R2; ++r) xin_n[q][r] = tr[q].X[(size_t)tn * tr[q].rs + xcol[r]];
                xv_n[q] = tr[q].X[(size_t)tn * tr[q].rs + vcol];
                pv_n[q] = tr[q].prm.get(tn, pcol);
            }
            g_n = Gsrc[(size_t)tn * L::GROW];
        }
        // Every stage first reads all its operands, then computes, then writes: stores to the shared
        // area would otherwise fence the loads of the next element (the compiler cannot prove that
        // the run-time offsets do not alias) and serialise the independent chains.
        // ---- stage 1: MA = A^T V, MB = B^T V; this step's G row; Quu of the lane's trajectory and its LU factors
        R v1[NT][R1];
        CRL_UNROLL for (int q = 0; q < NT; ++q)
            CRL_UNROLL for (int r = 0; r < R1; ++r) v1[q][r] = coop_hsum<R>(S + q * L::SZ, a_o0[r], a_o1[r], a_kind[r], h);
        R Quu[M][M];
        CRL_UNROLL for (int a = 0; a < M; ++a)
            CRL_UNROLL for (int b = 0; b < M; ++b) {
                const R acc = (S3[L::V + (2 * a + 1) * NS + 2 * b + 1] * h) * h;
                Quu[a][b] = (a == b) ? R(cost_hess<Mdl>(N + a)) + acc : acc;
            }
        CRL_UNROLL for (int q = 0; q < NT; ++q)
            CRL_UNROLL for (int r = 0; r < R1; ++r) S[q * L::SZ + a_dst[r]] = v1[q][r];
        S[g_dst] = g_cur;
        CRL_UNROLL for (int a = 0; a < M; ++a)                        // every lane of the group holds the same Quu
            CRL_UNROLL for (int b = 0; b < M; ++b) S3[L::QUU + a * M + b] = Quu[a][b];
        LuFactors<M, R> lu;
        lu_factor<M, R>(Quu, lu);
        __syncwarp();
        // ---- stage 2: Qss, Qus, qs, qu
        R o2[NT][R2];
        CRL_UNROLL for (int q = 0; q < NT; ++q) {
            const R* Sq = S + q * L::SZ;
            CRL_UNROLL for (int r = 0; r < R2; ++r) {
                R acc = coop_hsum<R>(Sq, b_o0[r], b_o1[r], b_kind[r], h);
                const R mga = Sq[b_gai[r]] * Sq[b_gas[r]];
                const R mgb = Sq[b_gbi[r]] * Sq[b_gbs[r]];
                R accg = fmadd(mga, Sq[b_gaj[r]], acc);
                accg = fmadd(mgb, Sq[b_gbj[r]], accg);
                acc = b_g[r] ? accg : acc;
                const R base = b_const[r] ? b_cbase[r] : b_cmul[r] * xin[q][r];
                o2[q][r] = b_add[r] ? base + acc : acc;
            }
        }
        CRL_UNROLL for (int q = 0; q < NT; ++q)
            CRL_UNROLL for (int r = 0; r < R2; ++r) S[q * L::SZ + b_dst[r]] = o2[q][r];
        __syncwarp();
        // ---- stage 3: gains, one right-hand-side column per lane
        {
            R b[M][1];
            CRL_UNROLL for (int a = 0; a < M; ++a) b[a][0] = S3[L::QU + a * NQ + r3];
            lu_apply<M, 1, R>(lu, b);
            CRL_UNROLL for (int a = 0; a < M; ++a) S3[L::KK + a * NQ + r3] = -b[a][0];   // lanes >= NT*NQ repeat the last column
        }
        __syncwarp();
        // ---- stage 5: V, v (W = K^T Quu recomputed per element), output block, K / k to global
        R o5[NT], kval[NT];
        CRL_UNROLL for (int q = 0; q < NT; ++q) {
            const R* Sq = S + q * L::SZ;
            R Ki[M], Kj[M], Ui[M], Uj[M], W[M];
            CRL_UNROLL for (int a = 0; a < M; ++a) {
                Ki[a] = Sq[L::KK + a * NQ + c_i];
                Kj[a] = Sq[L::KK + a * NQ + c_j];
                Ui[a] = Sq[L::QU + a * NQ + c_i];
                Uj[a] = Sq[L::QU + a * NQ + c_j];
            }
            CRL_UNROLL for (int b = 0; b < M; ++b) {
                R acc = Ki[0] * Sq[L::QUU + b];
                CRL_UNROLL for (int a = 1; a < M; ++a) acc = fmadd(Ki[a], Sq[L::QUU + a * M + b], acc);
                W[b] = acc;
            }
            R sij = Ui[0] * Kj[0], sji = Uj[0] * Ki[0], rr = W[0] * Kj[0];
            CRL_UNROLL for (int a = 1; a < M; ++a) {
                sij = fmadd(Ui[a], Kj[a], sij);
                sji = fmadd(Uj[a], Ki[a], sji);
                rr = fmadd(W[a], Kj[a], rr);
            }
            o5[q] = ((Sq[c_base] + sij) + sji) + rr;
            kval[q] = Sq[L::KK + st_a * NQ + st_j];
        }
        CRL_UNROLL for (int q = 0; q < NT; ++q) {
            R* Sq = S + q * L::SZ;
            const R w2 = R(2.0 * Mdl::weight(NS));
            const R vyv = (-w2) * pv[q] + w2 * xv[q];
            Sq[c_dst1] = c_ok ? o5[q] : vyv;
            Sq[c_dst2] = c_ok ? o5[q] : R(cost_hess<Mdl>(NS));
            *st_ptr[q] = kval[q];
            st_ptr[q] -= st_stride;
        }
        __syncwarp();
    }
}

// -----------------------------------------------------------------------------------------
// Column-group backward pass: 8 lanes = one trajectory, 4 trajectories per warp in true SIMT
// parallel.  Lane c of a group owns COLUMN c of every matrix of the Riccati step in registers
// (c < NS: state column; c == NS: the affine column v / qs / qu / k; the rest idle), so
//   A^T V, B^T V                      are local,
//   (A^T V) A, (B^T V) A              need the left neighbour's column        (shuffle up by one),
//   Quu                               needs B^T V of the odd columns          (9 shuffles),
//   the LU of Quu                     is factored by every lane and applied to its own column,
//   V' = Qss + S + S^T + K^T Quu K    needs K and Qus of every column         (36 shuffles),
//   the mirror of the symmetric V'    goes through a 3 KB shared-memory transpose.
// About 480 instructions and ~130 shuffle / shared-memory wavefronts per time step for four
// trajectories (the element-per-lane pass above: ~300 instructions and ~150 wavefronts per
// trajectory), and the same per-element operation sequence as riccati_step once more.
template <class Mdl>
struct ColGroup {
    static constexpr int GS = 8;                              // lanes per trajectory
    static constexpr int NG = 4;                              // trajectories per warp
    static constexpr int SMEM = 2 * NG * Mdl::NS * GS;        // reals: double-buffered transpose area
    static_assert(Mdl::NS + 1 <= GS, "state columns + the affine column must fit a lane group");
};

template <class Mdl, class R>
__device__ void colgroup_backward(const CoopTraj<R> (&tr)[4], int T, R* S, int lane) {
    using CG = ColGroup<Mdl>;
    constexpr int NS = Mdl::NS, M = Mdl::M, J = Mdl::M, N = Mdl::N, GS = CG::GS, NG = CG::NG;
    const R h = R(kH);
    const int g = lane / GS, c = lane % GS;
    const bool is_col = c < NS, is_aff = c == NS, c_odd = (c & 1) != 0 && is_col, c_even_col = is_col && !(c & 1);
    const int ch = c_even_col ? c / 2 : 0;                    // index of this column's angle (even columns)

    // ---- G tables of the four trajectories
    coop_gtables<Mdl, R, NG>(tr, T, lane);
    __syncwarp();

    // this lane's trajectory
    CoopTraj<R> me = tr[0];
    CRL_UNROLL for (int q = 1; q < NG; ++q)
        if (q == g) me = tr[q];
    // where this lane's gains go: K[a][c] (state column) or k[a] (affine column)
    const int kc = is_col ? c : 0;
    R* kst = (is_aff ? me.k : me.K + kc) + (size_t)(T - 1) * (is_aff ? M : M * NS);
    const int kst_a = is_aff ? 1 : NS, kst_t = is_aff ? M : M * NS;

    // value function: column c of V (affine lane: v); output block in every lane
    R Vc[NS], vy[2], vyy[2];
    CRL_UNROLL for (int i = 0; i < NS; ++i) Vc[i] = R(0);
    vy[0] = vy[1] = vyy[0] = vyy[1] = R(0);

    // row values, one step ahead: G row (2J), x_i of the odd rows (J), u (M), output entries (2), targets (2)
    R Gn[2 * J], xo_n[J], u_n[M], y_n[2], p_n[2], tbar_n;
    {
        const size_t r = (size_t)(T - 1);
        const size_t rs = (size_t)me.rs;
        CRL_UNROLL for (int i = 0; i < 2 * J; ++i) Gn[i] = me.G[r * (2 * J) + i];
        CRL_UNROLL for (int j = 0; j < J; ++j) xo_n[j] = me.X[r * rs + 2 * j + 1];
        CRL_UNROLL for (int a = 0; a < M; ++a) u_n[a] = me.X[r * rs + N + a];
        CRL_UNROLL for (int y = 0; y < 2; ++y) { y_n[y] = me.X[r * rs + NS + y]; p_n[y] = me.prm.get(T - 1, y); }
        tbar_n = coop_barrier_t<Mdl, R>(me.prm, T - 1, me.first);
    }

    for (int t = T - 1; t >= 0; --t) {
        R G[2][J], xo[J], u[M], yv[2], pv[2];
        CRL_UNROLL for (int i = 0; i < J; ++i) { G[0][i] = Gn[i]; G[1][i] = Gn[J + i]; }
        CRL_UNROLL for (int j = 0; j < J; ++j) xo[j] = xo_n[j];
        CRL_UNROLL for (int a = 0; a < M; ++a) u[a] = u_n[a];
        CRL_UNROLL for (int y = 0; y < 2; ++y) { yv[y] = y_n[y]; pv[y] = p_n[y]; }
        R gu[M], hu[M];                                       // barrier costs: action entries of the cost gradient / Hessian
        if constexpr (Mdl::kBarrier) coop_cost_u<Mdl, R>(u, tbar_n, gu, hu);
        {
            const size_t r = (size_t)(t > 0 ? t - 1 : 0);
            const size_t rs = (size_t)me.rs;
            CRL_UNROLL for (int i = 0; i < 2 * J; ++i) Gn[i] = me.G[r * (2 * J) + i];
            CRL_UNROLL for (int j = 0; j < J; ++j) xo_n[j] = me.X[r * rs + 2 * j + 1];
            CRL_UNROLL for (int a = 0; a < M; ++a) u_n[a] = me.X[r * rs + N + a];
            CRL_UNROLL for (int y = 0; y < 2; ++y) { y_n[y] = me.X[r * rs + NS + y]; p_n[y] = me.prm.get((int)r, y); }
            tbar_n = coop_barrier_t<Mdl, R>(me.prm, (int)r, me.first);
        }
        // ---- column c of A^T V and B^T V (affine lane: A^T v, B^T v)
        R MA[NS], MB[M];
        CRL_UNROLL for (int i = 0; i < NS; ++i) MA[i] = (i & 1) ? Vc[i - 1] * h + Vc[i] : Vc[i];
        CRL_UNROLL for (int a = 0; a < M; ++a) MB[a] = Vc[2 * a + 1] * h;
        // ---- Quu from the odd columns, factored by every lane
        R Quu[M][M];
        CRL_UNROLL for (int a = 0; a < M; ++a)
            CRL_UNROLL for (int b = 0; b < M; ++b) {
                const R acc = __shfl_sync(0xffffffffu, MB[a], 2 * b + 1, GS) * h;
                if constexpr (Mdl::kBarrier) Quu[a][b] = (a == b) ? hu[a] + acc : acc;
                else Quu[a][b] = (a == b) ? R(cost_hess<Mdl>(N + a)) + acc : acc;
            }
        LuFactors<M, R> lu;
        lu_factor<M, R>(Quu, lu);
        // ---- column c of [Qss | qs] and [Qus | qu]
        R Gc[2];                                              // G[y][c/2] of an even state column
        CRL_UNROLL for (int y = 0; y < 2; ++y) {
            Gc[y] = G[y][0];
            CRL_UNROLL for (int i = 1; i < J; ++i) Gc[y] = (ch == i) ? G[y][i] : Gc[y];
        }
        R Qx[NS], Qu[M][1];
        CRL_UNROLL for (int i = 0; i < NS; ++i) {
            const R left = __shfl_up_sync(0xffffffffu, MA[i], 1, GS);
            // state column: (MA[i][c-1] h + MA[i][c]) + G terms (+ Hessian on the diagonal)
            R acc = c_odd ? left * h + MA[i] : MA[i];
            // affine column: MA_v[i] + vy G terms, + cost gradient
            R aff = MA[i];
            if (!(i & 1)) {
                R ag = fmadd(G[0][i / 2] * vyy[0], Gc[0], acc);
                ag = fmadd(G[1][i / 2] * vyy[1], Gc[1], ag);
                acc = c_even_col ? ag : acc;
                aff = fmadd(vy[0], G[0][i / 2], aff);
                aff = fmadd(vy[1], G[1][i / 2], aff);
            }
            const R hess = (i & 1) ? R(cost_hess<Mdl>(1)) : R(cost_hess<Mdl>(0));
            const R diag = hess + acc;
            const R grad = (i & 1) ? R(2.0 * Mdl::weight(1)) * xo[i / 2] : R(0.0);
            Qx[i] = is_aff ? grad + aff : ((i == c) ? diag : acc);
        }
        CRL_UNROLL for (int a = 0; a < M; ++a) {
            const R left = __shfl_up_sync(0xffffffffu, MB[a], 1, GS);
            const R col = c_odd ? left * h + MB[a] : MB[a];
            R aff;
            if constexpr (Mdl::kBarrier) aff = gu[a] + MB[a];
            else aff = R(2.0 * Mdl::weight(N)) * u[a] + MB[a];
            Qu[a][0] = is_aff ? aff : col;
        }
        // ---- gains of this column
        R Kc[M][1];
        CRL_UNROLL for (int a = 0; a < M; ++a) Kc[a][0] = Qu[a][0];
        lu_apply<M, 1, R>(lu, Kc);
        CRL_UNROLL for (int a = 0; a < M; ++a) Kc[a][0] = -Kc[a][0];
        if (is_col || is_aff) {
            CRL_UNROLL for (int a = 0; a < M; ++a) kst[a * kst_a] = Kc[a][0];
        }
        kst -= kst_t;
        // ---- new column: rows i <= c of V (affine lane: all of v)
        R Vn[NS];
        CRL_UNROLL for (int i = 0; i < NS; ++i) {
            R Ki[M], Ui[M], W[M];
            CRL_UNROLL for (int a = 0; a < M; ++a) {
                Ki[a] = __shfl_sync(0xffffffffu, Kc[a][0], i, GS);
                Ui[a] = __shfl_sync(0xffffffffu, Qu[a][0], i, GS);
            }
            CRL_UNROLL for (int b = 0; b < M; ++b) {
                R acc = Ki[0] * Quu[0][b];
                CRL_UNROLL for (int a = 1; a < M; ++a) acc = fmadd(Ki[a], Quu[a][b], acc);
                W[b] = acc;
            }
            R sij = Ui[0] * Kc[0][0], sji = Qu[0][0] * Ki[0], rr = W[0] * Kc[0][0];
            CRL_UNROLL for (int a = 1; a < M; ++a) {
                sij = fmadd(Ui[a], Kc[a][0], sij);
                sji = fmadd(Qu[a][0], Ki[a], sji);
                rr = fmadd(W[a], Kc[a][0], rr);
            }
            Vn[i] = ((Qx[i] + sij) + sji) + rr;
        }
        // ---- mirror: V(i, c) for i > c is column i's row c
        R* Sb = S + (t & 1) * (NG * NS * GS) + g * (NS * GS);
        CRL_UNROLL for (int i = 0; i < NS; ++i) Sb[i * GS + c] = Vn[i];
        __syncwarp();
        CRL_UNROLL for (int i = 0; i < NS; ++i) {
            const R low = Sb[(is_col ? c : 0) * GS + i];
            Vc[i] = (is_col && i > c) ? low : Vn[i];
        }
        if (!is_col && !is_aff) {
            CRL_UNROLL for (int i = 0; i < NS; ++i) Vc[i] = R(0);      // idle lanes stay at zero
        }
        // ---- output block
        CRL_UNROLL for (int y = 0; y < 2; ++y) {
            const R w2 = R(2.0 * Mdl::weight(NS));
            vy[y] = (-w2) * pv[y] + w2 * yv[y];
            vyy[y] = R(cost_hess<Mdl>(NS));
        }
    }
    __syncwarp();
}

// -----------------------------------------------------------------------------------------
// Pair-group backward pass (kinematic chains): 4 lanes = one trajectory, 8 trajectories per warp.
// Lane p < J of a group owns the columns (2p, 2p + 1) - joint p's angle and velocity - of every
// matrix of the Riccati step, lane J the affine column (v / qs / qu / k).  Compared with one column
// per lane (colgroup_backward) this
//   * makes (A^T V) A local: the "left neighbour" of the velocity column is the lane's own angle column
//     (A = I + h at (2j, 2j + 1)), so the neighbour shuffles and the odd / even column selects disappear;
//   * amortises Quu, its LU factors, W_i = K_i^T Quu and the row shuffles of the update
//     V' = Qss + S + S^T + K^T Quu K over two columns;
//   * steps eight trajectories per pass for ~0.8x the instructions of the four-trajectory pass
//     (which is issue-bound: 900 instructions per time step, 270 of them FP64).
// Per element the operation sequence is colgroup_backward's (= riccati_step's) once more.
template <class Mdl>
struct PairGroup {
    static constexpr int GS = 4;                              // lanes per trajectory
    static constexpr int NG = 8;                              // trajectories per warp
    static constexpr int SMEM = 2 * NG * Mdl::NS * Mdl::NS;   // reals: double-buffered mirror area [trajectory][column][row]
    static_assert(Mdl::M + 1 <= GS, "joint lanes + the affine lane must fit a lane group");
};

template <class Mdl, class R>
__device__ void pairgroup_backward(const CoopTraj<R> (&tr)[8], int T, R* S, int lane) {
    using PG = PairGroup<Mdl>;
    static_assert(!Mdl::kBarrier, "the pair-group pass is written for the plain costs");
    constexpr int NS = Mdl::NS, M = Mdl::M, J = Mdl::M, N = Mdl::N, GS = PG::GS, NG = PG::NG;
    const R h = R(kH);
    const int g = lane / GS, p = lane % GS;
    const bool is_st = p < J, is_aff = p == J;
    const int pc = is_st ? p : 0;                             // this lane's joint
    const int c0 = 2 * pc, c1 = 2 * pc + 1;                   // its columns

    // ---- G tables of the eight trajectories, four at a time
    {
        CoopTraj<R> half[4];
        CRL_UNROLL for (int q = 0; q < 4; ++q) half[q] = tr[q];
        coop_gtables<Mdl, R, 4>(half, T, lane);
        CRL_UNROLL for (int q = 0; q < 4; ++q) half[q] = tr[4 + q];
        coop_gtables<Mdl, R, 4>(half, T, lane);
    }
    __syncwarp();

    CoopTraj<R> me = tr[0];
    CRL_UNROLL for (int q = 1; q < NG; ++q)
        if (q == g) me = tr[q];
    // where this lane's gains go: K[a][c0], K[a][c1] (joint lane) or k[a] (affine lane)
    R* kst = (is_aff ? me.k : me.K + c0) + (size_t)(T - 1) * (is_aff ? M : M * NS);
    const int kst_a = is_aff ? 1 : NS, kst_t = is_aff ? M : M * NS;

    // value function: columns c0, c1 of V (affine lane: v and nothing); output block in every lane
    R V0[NS], V1[NS], vy[2], vyy[2];
    CRL_UNROLL for (int i = 0; i < NS; ++i) V0[i] = V1[i] = R(0);
    vy[0] = vy[1] = vyy[0] = vyy[1] = R(0);

    // row values, one step ahead: G row (2J) and its entries of this lane's joint (2), x_i of the odd rows (J),
    // u (M), output entries (2), targets (2)
    R Gn[2 * J], Gcn[2], xo_n[J], u_n[M], y_n[2], p_n[2];
    {
        const size_t r = (size_t)(T - 1), rs = (size_t)me.rs;
        CRL_UNROLL for (int i = 0; i < 2 * J; ++i) Gn[i] = me.G[r * (2 * J) + i];
        CRL_UNROLL for (int y = 0; y < 2; ++y) Gcn[y] = me.G[r * (2 * J) + y * J + pc];
        CRL_UNROLL for (int j = 0; j < J; ++j) xo_n[j] = me.X[r * rs + 2 * j + 1];
        CRL_UNROLL for (int a = 0; a < M; ++a) u_n[a] = me.X[r * rs + N + a];
        CRL_UNROLL for (int y = 0; y < 2; ++y) { y_n[y] = me.X[r * rs + NS + y]; p_n[y] = me.prm.get(T - 1, y); }
    }

    for (int t = T - 1; t >= 0; --t) {
        R G[2][J], Gc[2], xo[J], u[M], yv[2], pv[2];
        CRL_UNROLL for (int i = 0; i < J; ++i) { G[0][i] = Gn[i]; G[1][i] = Gn[J + i]; }
        CRL_UNROLL for (int y = 0; y < 2; ++y) Gc[y] = Gcn[y];
        CRL_UNROLL for (int j = 0; j < J; ++j) xo[j] = xo_n[j];
        CRL_UNROLL for (int a = 0; a < M; ++a) u[a] = u_n[a];
        CRL_UNROLL for (int y = 0; y < 2; ++y) { yv[y] = y_n[y]; pv[y] = p_n[y]; }
        {
            const size_t r = (size_t)(t > 0 ? t - 1 : 0), rs = (size_t)me.rs;
            CRL_UNROLL for (int i = 0; i < 2 * J; ++i) Gn[i] = me.G[r * (2 * J) + i];
            CRL_UNROLL for (int y = 0; y < 2; ++y) Gcn[y] = me.G[r * (2 * J) + y * J + pc];
            CRL_UNROLL for (int j = 0; j < J; ++j) xo_n[j] = me.X[r * rs + 2 * j + 1];
            CRL_UNROLL for (int a = 0; a < M; ++a) u_n[a] = me.X[r * rs + N + a];
            CRL_UNROLL for (int y = 0; y < 2; ++y) { y_n[y] = me.X[r * rs + NS + y]; p_n[y] = me.prm.get((int)r, y); }
        }
        // ---- columns c0, c1 of A^T V and B^T V (affine lane: A^T v, B^T v)
        R MA0[NS], MA1[NS], MB0[M], MB1[M];
        CRL_UNROLL for (int i = 0; i < NS; ++i) {
            MA0[i] = (i & 1) ? V0[i - 1] * h + V0[i] : V0[i];
            MA1[i] = (i & 1) ? V1[i - 1] * h + V1[i] : V1[i];
        }
        CRL_UNROLL for (int a = 0; a < M; ++a) { MB0[a] = V0[2 * a + 1] * h; MB1[a] = V1[2 * a + 1] * h; }
        // ---- Quu from the velocity columns (column 2b + 1 = second column of lane b), factored by every lane
        R Quu[M][M];
        CRL_UNROLL for (int a = 0; a < M; ++a)
            CRL_UNROLL for (int b = 0; b < M; ++b) {
                const R acc = __shfl_sync(0xffffffffu, MB1[a], b, GS) * h;
                Quu[a][b] = (a == b) ? R(cost_hess<Mdl>(N + a)) + acc : acc;
            }
        LuFactors<M, R> lu;
        lu_factor<M, R>(Quu, lu);
        // ---- columns c0, c1 of [Qss | qs] and [Qus | qu]
        R Qx0[NS], Qx1[NS], Qu0[M], Qu1[M];
        CRL_UNROLL for (int i = 0; i < NS; ++i) {
            const R hess = (i & 1) ? R(cost_hess<Mdl>(1)) : R(cost_hess<Mdl>(0));
            // angle column c0: MA (no left neighbour) + G terms on the angle rows (+ Hessian on the diagonal);
            // affine column: MA_v + vy G terms, + cost gradient
            R acc0 = MA0[i], aff = MA0[i];
            if (!(i & 1)) {
                R ag = fmadd(G[0][i / 2] * vyy[0], Gc[0], acc0);
                ag = fmadd(G[1][i / 2] * vyy[1], Gc[1], ag);
                acc0 = ag;
                aff = fmadd(vy[0], G[0][i / 2], aff);
                aff = fmadd(vy[1], G[1][i / 2], aff);
            }
            const R diag0 = hess + acc0;
            const R q0 = (!(i & 1) && i == c0) ? diag0 : acc0;          // the diagonal of an angle column is an angle row
            const R grad = (i & 1) ? R(2.0 * Mdl::weight(1)) * xo[i / 2] : R(0.0);
            Qx0[i] = is_aff ? grad + aff : q0;
            // velocity column c1: (MA[i][c0] h + MA[i][c1]) (+ Hessian on the diagonal, a velocity row)
            const R acc1 = MA0[i] * h + MA1[i];
            const R diag1 = hess + acc1;
            Qx1[i] = ((i & 1) && i == c1) ? diag1 : acc1;
        }
        CRL_UNROLL for (int a = 0; a < M; ++a) {
            const R aff = R(2.0 * Mdl::weight(N)) * u[a] + MB0[a];
            Qu0[a] = is_aff ? aff : MB0[a];
            Qu1[a] = MB0[a] * h + MB1[a];
        }
        // ---- gains of the two columns
        R Kc[M][2];
        CRL_UNROLL for (int a = 0; a < M; ++a) { Kc[a][0] = Qu0[a]; Kc[a][1] = Qu1[a]; }
        lu_apply<M, 2, R>(lu, Kc);
        CRL_UNROLL for (int a = 0; a < M; ++a) { Kc[a][0] = -Kc[a][0]; Kc[a][1] = -Kc[a][1]; }
        if (is_st || is_aff) {
            CRL_UNROLL for (int a = 0; a < M; ++a) kst[a * kst_a] = Kc[a][0];
        }
        if (is_st) {
            CRL_UNROLL for (int a = 0; a < M; ++a) kst[a * kst_a + 1] = Kc[a][1];
        }
        kst -= kst_t;
        // ---- new columns (affine lane: all of v); row i's K, Qus and W = K_i^T Quu come from column i's lane
        R Vn0[NS], Vn1[NS];
        CRL_UNROLL for (int i = 0; i < NS; ++i) {
            R Ki[M], Ui[M], W[M];
            CRL_UNROLL for (int a = 0; a < M; ++a) {
                Ki[a] = __shfl_sync(0xffffffffu, Kc[a][i & 1], i / 2, GS);
                Ui[a] = __shfl_sync(0xffffffffu, (i & 1) ? Qu1[a] : Qu0[a], i / 2, GS);
            }
            CRL_UNROLL for (int b = 0; b < M; ++b) {
                R acc = Ki[0] * Quu[0][b];
                CRL_UNROLL for (int a = 1; a < M; ++a) acc = fmadd(Ki[a], Quu[a][b], acc);
                W[b] = acc;
            }
            R sij0 = Ui[0] * Kc[0][0], sji0 = Qu0[0] * Ki[0], rr0 = W[0] * Kc[0][0];
            R sij1 = Ui[0] * Kc[0][1], sji1 = Qu1[0] * Ki[0], rr1 = W[0] * Kc[0][1];
            CRL_UNROLL for (int a = 1; a < M; ++a) {
                sij0 = fmadd(Ui[a], Kc[a][0], sij0);
                sji0 = fmadd(Qu0[a], Ki[a], sji0);
                rr0 = fmadd(W[a], Kc[a][0], rr0);
                sij1 = fmadd(Ui[a], Kc[a][1], sij1);
                sji1 = fmadd(Qu1[a], Ki[a], sji1);
                rr1 = fmadd(W[a], Kc[a][1], rr1);
            }
            Vn0[i] = ((Qx0[i] + sij0) + sji0) + rr0;
            Vn1[i] = ((Qx1[i] + sij1) + sji1) + rr1;
        }
        // ---- mirror: V(i, c) for i > c is column i's row c
        R* Sb = S + (t & 1) * (NG * NS * NS) + g * (NS * NS);
        if (is_st) {
            CRL_UNROLL for (int i = 0; i < NS; ++i) { Sb[c0 * NS + i] = Vn0[i]; Sb[c1 * NS + i] = Vn1[i]; }
        }
        __syncwarp();
        CRL_UNROLL for (int i = 0; i < NS; ++i) {
            R a0 = Vn0[i], a1 = is_st ? Vn1[i] : R(0);             // only joint lanes have a second column
            if (is_st && i > c0) a0 = Sb[i * NS + c0];
            if (is_st && i > c1) a1 = Sb[i * NS + c1];
            V0[i] = a0;
            V1[i] = a1;
        }
        if (!is_st && !is_aff) {
            CRL_UNROLL for (int i = 0; i < NS; ++i) V0[i] = R(0);      // idle lanes stay at zero
        }
        // ---- output block
        CRL_UNROLL for (int y = 0; y < 2; ++y) {
            const R w2 = R(2.0 * Mdl::weight(NS));
            vy[y] = (-w2) * pv[y] + w2 * yv[y];
            vyy[y] = R(cost_hess<Mdl>(NS));
        }
    }
    __syncwarp();
}

// -----------------------------------------------------------------------------------------
// Column-group backward pass of the rigid two-link arm (RoboticArm): the same lane mapping as
// colgroup_backward, but the Jacobian has dense variable rows (rows 1 and 3 of A and B), so
//   A^T V, B^T V        combine rows 0..3 of the lane's own column with the step's A / B entries,
//   (A^T V) A, (B^T V) A need ALL columns of A^T V / B^T V (24 shuffles),
//   Quu                 needs columns 1 and 3 of B^T V.
// The 16 variable entries of a time step (A[1][:], A[3][:], B[1][:], B[3][:], G) are precomputed for
// all time steps in parallel (lanes over t; Mdl::coop_table).  Per element the operation sequence is
// riccati_step's with the structural zeros / ones / h of RoboticArm::akind pruned the same way:
// a product with the literal 1 (x * 1, fma(x, 1, acc)) stands for the plain copy / addition there.
template <class Mdl, class R>
__device__ void colgroup_backward_arm(const CoopTraj<R> (&tr)[4], int T, R* S, int lane) {
    using CG = ColGroup<Mdl>;
    constexpr int NS = Mdl::NS, M = Mdl::M, N = Mdl::N, D = Mdl::D, GS = CG::GS, NG = CG::NG, TAB = Mdl::kCoopTab;
    static_assert(NS == 4 && M == 2 && TAB == 16, "written for the 4-state / 2-action arm");
    const R h = R(kH);
    const int g = lane / GS, c = lane % GS;
    const bool is_col = c < NS, is_aff = c == NS, c_even_col = is_col && !(c & 1);
    const int cc = is_col ? c : 0;

    // ---- tables of the four trajectories: lanes take time steps lane, lane + 32, ...
    for (int t = lane; t < T; t += 32) {
        R xu[NG][D];
        CRL_UNROLL for (int q = 0; q < NG; ++q)
            CRL_UNROLL for (int i = 0; i < D; ++i) xu[q][i] = tr[q].X[(size_t)t * tr[q].rs + i];
        CRL_UNROLL for (int q = 0; q < NG; ++q) {
            R tab[TAB];
            Mdl::template coop_table<R>(xu[q], tab);
            CRL_UNROLL for (int i = 0; i < TAB; ++i) tr[q].G[(size_t)t * TAB + i] = tab[i];
        }
    }
    __syncwarp();

    CoopTraj<R> me = tr[0];
    CRL_UNROLL for (int q = 1; q < NG; ++q)
        if (q == g) me = tr[q];
    R* kst = (is_aff ? me.k : me.K + cc) + (size_t)(T - 1) * (is_aff ? M : M * NS);
    const int kst_a = is_aff ? 1 : NS, kst_t = is_aff ? M : M * NS;

    R Vc[NS], vy[2], vyy[2];
    CRL_UNROLL for (int i = 0; i < NS; ++i) Vc[i] = R(0);
    vy[0] = vy[1] = vyy[0] = vyy[1] = R(0);

    // row values, one step ahead: the table (16), x_1, x_3 (cost gradient), u (2), output entries (2), targets (2)
    R tb_n[TAB], xo_n[2], u_n[M], y_n[2], p_n[2], tbar_n;
    {
        const size_t r = (size_t)(T - 1), rs = (size_t)me.rs;
        CRL_UNROLL for (int i = 0; i < TAB; ++i) tb_n[i] = me.G[r * TAB + i];
        CRL_UNROLL for (int j = 0; j < 2; ++j) xo_n[j] = me.X[r * rs + 2 * j + 1];
        CRL_UNROLL for (int a = 0; a < M; ++a) u_n[a] = me.X[r * rs + N + a];
        CRL_UNROLL for (int y = 0; y < 2; ++y) { y_n[y] = me.X[r * rs + NS + y]; p_n[y] = me.prm.get(T - 1, y); }
        tbar_n = coop_barrier_t<Mdl, R>(me.prm, T - 1, me.first);
    }

    for (int t = T - 1; t >= 0; --t) {
        R A1[NS], A3[NS], B1[M], B3[M], G[2][2], xo[2], u[M], yv[2], pv[2];
        CRL_UNROLL for (int j = 0; j < NS; ++j) { A1[j] = tb_n[j]; A3[j] = tb_n[NS + j]; }
        CRL_UNROLL for (int a = 0; a < M; ++a) { B1[a] = tb_n[2 * NS + a]; B3[a] = tb_n[2 * NS + M + a]; }
        G[0][0] = tb_n[12]; G[0][1] = tb_n[13]; G[1][0] = tb_n[14]; G[1][1] = tb_n[15];   // G[y][i/2] of the even columns
        CRL_UNROLL for (int j = 0; j < 2; ++j) xo[j] = xo_n[j];
        CRL_UNROLL for (int a = 0; a < M; ++a) u[a] = u_n[a];
        CRL_UNROLL for (int y = 0; y < 2; ++y) { yv[y] = y_n[y]; pv[y] = p_n[y]; }
        R gu[M], hu[M];                                       // barrier costs: action entries of the cost gradient / Hessian
        if constexpr (Mdl::kBarrier) coop_cost_u<Mdl, R>(u, tbar_n, gu, hu);
        {
            const size_t r = (size_t)(t > 0 ? t - 1 : 0), rs = (size_t)me.rs;
            CRL_UNROLL for (int i = 0; i < TAB; ++i) tb_n[i] = me.G[r * TAB + i];
            CRL_UNROLL for (int j = 0; j < 2; ++j) xo_n[j] = me.X[r * rs + 2 * j + 1];
            CRL_UNROLL for (int a = 0; a < M; ++a) u_n[a] = me.X[r * rs + N + a];
            CRL_UNROLL for (int y = 0; y < 2; ++y) { y_n[y] = me.X[r * rs + NS + y]; p_n[y] = me.prm.get((int)r, y); }
            tbar_n = coop_barrier_t<Mdl, R>(me.prm, (int)r, me.first);
        }
        // ---- column c of A^T V and B^T V (affine lane: A^T v, B^T v); l ascending over the non-zero A[l][i]
        R MA[NS], MB[M];
        MA[0] = fmadd(Vc[3], A3[0], fmadd(Vc[1], A1[0], Vc[0]));
        MA[1] = fmadd(Vc[3], A3[1], fmadd(Vc[1], A1[1], Vc[0] * h));
        MA[2] = fmadd(Vc[3], A3[2], Vc[1] * A1[2] + Vc[2]);
        MA[3] = fmadd(Vc[3], A3[3], fmadd(Vc[2], h, Vc[1] * A1[3]));
        CRL_UNROLL for (int a = 0; a < M; ++a) MB[a] = fmadd(Vc[3], B3[a], Vc[1] * B1[a]);
        // ---- Quu from columns 1 and 3, factored by every lane
        R Quu[M][M];
        CRL_UNROLL for (int a = 0; a < M; ++a) {
            const R m1 = __shfl_sync(0xffffffffu, MB[a], 1, GS), m3 = __shfl_sync(0xffffffffu, MB[a], 3, GS);
            CRL_UNROLL for (int b = 0; b < M; ++b) {
                const R acc = fmadd(m3, B3[b], m1 * B1[b]);
                if constexpr (Mdl::kBarrier) Quu[a][b] = (a == b) ? hu[a] + acc : acc;
                else Quu[a][b] = (a == b) ? R(cost_hess<Mdl>(N + a)) + acc : acc;
            }
        }
        LuFactors<M, R> lu;
        lu_factor<M, R>(Quu, lu);
        // ---- column c of [Qss | qs] and [Qus | qu]: sum over k of (row)[k] * A[k][c], k ascending over the
        //      non-zero A[k][c]: c < 2: k = 0 (1 or h), 1, 3;  c >= 2: k = 1, 2 (1 or h), 3
        const bool lo = cc < 2;
        R a1c = A1[0], a3c = A3[0];
        CRL_UNROLL for (int j = 1; j < NS; ++j) { a1c = (cc == j) ? A1[j] : a1c; a3c = (cc == j) ? A3[j] : a3c; }
        const R c_first = lo ? ((cc == 0) ? R(1) : h) : a1c;      // coefficient of the first term
        const R c_second = lo ? a1c : ((cc == 2) ? R(1) : h);      // coefficient of the second term
        const R Gc[2] = {(cc == 2) ? G[0][1] : G[0][0], (cc == 2) ? G[1][1] : G[1][0]};
        R Qx[NS], Qu[M][1];
        CRL_UNROLL for (int i = 0; i < NS; ++i) {
            const R r0 = __shfl_sync(0xffffffffu, MA[i], 0, GS), r1 = __shfl_sync(0xffffffffu, MA[i], 1, GS);
            const R r2 = __shfl_sync(0xffffffffu, MA[i], 2, GS), r3 = __shfl_sync(0xffffffffu, MA[i], 3, GS);
            R acc = (lo ? r0 : r1) * c_first;
            acc = fmadd(lo ? r1 : r2, c_second, acc);
            acc = fmadd(r3, a3c, acc);
            R aff = MA[i];
            if (!(i & 1)) {
                R ag = fmadd(G[0][i / 2] * vyy[0], Gc[0], acc);
                ag = fmadd(G[1][i / 2] * vyy[1], Gc[1], ag);
                acc = c_even_col ? ag : acc;
                aff = fmadd(vy[0], G[0][i / 2], aff);
                aff = fmadd(vy[1], G[1][i / 2], aff);
            }
            const R hess = (i & 1) ? R(cost_hess<Mdl>(1)) : R(cost_hess<Mdl>(0));
            const R diag = hess + acc;
            const R grad = (i & 1) ? R(2.0 * Mdl::weight(1)) * xo[i / 2] : R(0.0);
            Qx[i] = is_aff ? grad + aff : ((i == c) ? diag : acc);
        }
        CRL_UNROLL for (int a = 0; a < M; ++a) {
            const R r0 = __shfl_sync(0xffffffffu, MB[a], 0, GS), r1 = __shfl_sync(0xffffffffu, MB[a], 1, GS);
            const R r2 = __shfl_sync(0xffffffffu, MB[a], 2, GS), r3 = __shfl_sync(0xffffffffu, MB[a], 3, GS);
            R col = (lo ? r0 : r1) * c_first;
            col = fmadd(lo ? r1 : r2, c_second, col);
            col = fmadd(r3, a3c, col);
            R aff;
            if constexpr (Mdl::kBarrier) aff = gu[a] + MB[a];
            else aff = R(2.0 * Mdl::weight(N)) * u[a] + MB[a];
            Qu[a][0] = is_aff ? aff : col;
        }
        // ---- gains of this column
        R Kc[M][1];
        CRL_UNROLL for (int a = 0; a < M; ++a) Kc[a][0] = Qu[a][0];
        lu_apply<M, 1, R>(lu, Kc);
        CRL_UNROLL for (int a = 0; a < M; ++a) Kc[a][0] = -Kc[a][0];
        if (is_col || is_aff) {
            CRL_UNROLL for (int a = 0; a < M; ++a) kst[a * kst_a] = Kc[a][0];
        }
        kst -= kst_t;
        // ---- new column: rows i <= c of V (affine lane: all of v)
        R Vn[NS];
        CRL_UNROLL for (int i = 0; i < NS; ++i) {
            R Ki[M], Ui[M], W[M];
            CRL_UNROLL for (int a = 0; a < M; ++a) {
                Ki[a] = __shfl_sync(0xffffffffu, Kc[a][0], i, GS);
                Ui[a] = __shfl_sync(0xffffffffu, Qu[a][0], i, GS);
            }
            CRL_UNROLL for (int b = 0; b < M; ++b) {
                R acc = Ki[0] * Quu[0][b];
                CRL_UNROLL for (int a = 1; a < M; ++a) acc = fmadd(Ki[a], Quu[a][b], acc);
                W[b] = acc;
            }
            R sij = Ui[0] * Kc[0][0], sji = Qu[0][0] * Ki[0], rr = W[0] * Kc[0][0];
            CRL_UNROLL for (int a = 1; a < M; ++a) {
                sij = fmadd(Ui[a], Kc[a][0], sij);
                sji = fmadd(Qu[a][0], Ki[a], sji);
                rr = fmadd(W[a], Kc[a][0], rr);
            }
            Vn[i] = ((Qx[i] + sij) + sji) + rr;
        }
        // ---- mirror: V(i, c) for i > c is column i's row c
        R* Sb = S + (t & 1) * (NG * NS * GS) + g * (NS * GS);
        CRL_UNROLL for (int i = 0; i < NS; ++i) Sb[i * GS + c] = Vn[i];
        __syncwarp();
        CRL_UNROLL for (int i = 0; i < NS; ++i) {
            const R low = Sb[cc * GS + i];
            Vc[i] = (is_col && i > c) ? low : Vn[i];
        }
        if (!is_col && !is_aff) {
            CRL_UNROLL for (int i = 0; i < NS; ++i) Vc[i] = R(0);
        }
        // ---- output block
        CRL_UNROLL for (int y = 0; y < 2; ++y) {
            const R w2 = R(2.0 * Mdl::weight(NS));
            vy[y] = (-w2) * pv[y] + w2 * yv[y];
            vyy[y] = R(cost_hess<Mdl>(NS));
        }
    }
    __syncwarp();
}

// Feed of a closed-loop rollout over single-trajectory arrays (old trajectory with rows RS reals
// apart, gains row-major) that loads step t+1's inputs while step t computes (one warp alone on an SM
// partition has nothing else to hide the load latency with) and pulls the rows kAhead steps further
// on into the L2 (the engine's scratch lives in DRAM: one step of compute is shorter than a DRAM access).
template <class Mdl, class R, int RS>
struct RowPrefetchFeed {
    static constexpr int kKN = Mdl::NS;
    static constexpr int kAhead = 6;
    const R* X;
    const R* K;
    const R* k;
    int nsteps;
    R nx[Mdl::NS], nu[Mdl::M], nK[Mdl::M * Mdl::NS], nk[Mdl::M];
    __device__ RowPrefetchFeed(const R* X_, const R* K_, const R* k_) : X(X_), K(K_), k(k_), nsteps(0) {}
    __device__ __forceinline__ static void pf(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
    __device__ __forceinline__ void prefetch(int t) const {
        const char* xr = reinterpret_cast<const char*>(X + (size_t)t * RS);
        const char* Kr = reinterpret_cast<const char*>(K + (size_t)t * (Mdl::M * Mdl::NS));
        pf(xr);
        pf(xr + (Mdl::D - 1) * sizeof(R));
        pf(Kr);
        pf(Kr + (Mdl::M * Mdl::NS - 1) * sizeof(R));
        pf(k + (size_t)t * Mdl::M);
    }
    __device__ __forceinline__ void load(int t) {
        const R* xr = X + (size_t)t * RS;
        const R* Kr = K + (size_t)t * (Mdl::M * Mdl::NS);
        const R* kr = k + (size_t)t * Mdl::M;
        CRL_UNROLL for (int j = 0; j < Mdl::NS; ++j) nx[j] = xr[j];
        CRL_UNROLL for (int a = 0; a < Mdl::M; ++a) nu[a] = xr[Mdl::N + a];
        CRL_UNROLL for (int j = 0; j < Mdl::M * Mdl::NS; ++j) nK[j] = Kr[j];
        CRL_UNROLL for (int a = 0; a < Mdl::M; ++a) nk[a] = kr[a];
    }
    __device__ void begin(int n) {
        nsteps = n;
        for (int t = 1; t <= kAhead && t < n; ++t) prefetch(t);
        if (n > 0) load(0);
    }
    __device__ void end() {}
    __device__ void row0(R* xu) const {
        CRL_UNROLL for (int i = 0; i < Mdl::D; ++i) xu[i] = X[i];
    }
    __device__ void get(int t, R* xo, R* uo, R* Kc, R* kc) {
        CRL_UNROLL for (int j = 0; j < Mdl::NS; ++j) xo[j] = nx[j];
        CRL_UNROLL for (int a = 0; a < Mdl::M; ++a) uo[a] = nu[a];
        CRL_UNROLL for (int j = 0; j < Mdl::M * Mdl::NS; ++j) Kc[j] = nK[j];
        CRL_UNROLL for (int a = 0; a < Mdl::M; ++a) kc[a] = nk[a];
        if (t + 1 + kAhead < nsteps) prefetch(t + 1 + kAhead);
        if (t + 1 < nsteps) load(t + 1);
    }
};

}  // namespace crl
