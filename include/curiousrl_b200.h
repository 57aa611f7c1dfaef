/*
 * curiousrl_b200 - C ABI of the B200 batched-iLQR path (libcuriousrl_b200.so).
 *
 * Drop-in boundary for the hot path of cheng-zilong/CuriousRL.  The reference has no native
 * code; its "operator interface" for this path is the set of Python methods the solver looks
 * up through `self` (paths relative to /root/reference/CuriousRL/algorithm/ilqr_solver/).
 * Each entry point below names the reference method(s) it replaces.  Python binds these with
 * ctypes (curiousrl_b200/_native.py); INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; arrays are dense,
 *    row-major, element type given by `dtype` (J / traces are always double);
 *  - shapes use B = batch, T = horizon, n/m/p = state/action/parameter sizes of the model,
 *    d = n + m;
 *  - no allocation, no global state, no exceptions: work is enqueued on `stream`
 *    (a cudaStream_t, 0 = default stream) and the call returns a status at once;
 *  - status 0 = ok; < 0 = error (crl_status_string()); NaN/Inf in data are not errors, they
 *    flow through comparisons exactly like the reference's float64 code.
 */
#ifndef CURIOUSRL_B200_H
#define CURIOUSRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRL_ABI_VERSION 8

typedef void* crl_stream_t; /* cudaStream_t */

enum crl_model {              /* scenario.name -> precompiled device model                     */
    CRL_TWO_LINK = 0,         /* TwoLinkPlanarManipulator   two_link_planar_manipulator.py:14  */
    CRL_THREE_LINK = 1,       /* ThreeLinkPlanarManipulator three_link_planar_manipulator.py:15 */
    CRL_ROBOTIC_ARM = 2,      /* RoboticArmTracking         robotic_arm_tracking.py:15         */
    /* The same scenarios under LogBarrieriLQR's cost (log_barrier_ilqr.py:86-101): every bounded action adds
       (-1/t) log(-(lo - u)) + (-1/t) log(-(u - hi)).  FP64 only.  The parameter row grows to p = 4:
       (target_x, target_y, t, t0), t0 = the barrier parameter the first backward pass of a solve still
       sees (the reference evaluates its cost derivatives before the outer loop moves t on).            */
    CRL_TWO_LINK_BARRIER = 3,
    CRL_THREE_LINK_BARRIER = 4,
    CRL_ROBOTIC_ARM_BARRIER = 5,
    /* Scenarios whose device model is GENERATED from the scenario's sympy expressions (curiousrl_b200/codegen.py ->
       csrc/ilqr_generated.cuh): dense Jacobians, state-dependent cost Hessians, time-varying cost parameters,
       per-component bounds on states and actions compiled into the model (u_lo / u_hi of crl_problem are ignored;
       crl_model_bounds reports them).  FP64 only.  Every entry point below accepts them.                  */
    CRL_GEN_CART_POLE_1 = 6,       /* CartPoleSwingUp1  cart_pole_swingup1.py:16   n=4 m=1 p=5 */
    CRL_GEN_CART_POLE_2 = 7,       /* CartPoleSwingUp2  cart_pole_swingup2.py:16   n=5 m=1 p=6 */
    CRL_GEN_VEHICLE_TRACKING = 8,  /* VehicleTracking   vehicle_tracking.py:15     n=4 m=2 p=1 (one unused zero column) */
    CRL_GEN_CAR_PARKING = 9,       /* CarParking        car_parking.py:23          n=4 m=2 p=2 */
    /* ... under LogBarrieriLQR's cost: here the barrier also sits on the finite STATE bounds (log_barrier_ilqr.py:91-95
       runs over every row of constr).  Parameter row = (the scenario's add_param columns, t, t0): p + 2 entries
       (VehicleTracking, which has no add_param: (t, t0)).                                                            */
    CRL_GEN_CART_POLE_1_BARRIER = 10,
    CRL_GEN_CART_POLE_2_BARRIER = 11,
    CRL_GEN_VEHICLE_TRACKING_BARRIER = 12,
    CRL_GEN_CAR_PARKING_BARRIER = 13,
    /* A scenario compiled at run time (curiousrl_b200/jit.py: the same generator, nvcc at `init`, a library of its own
       that exports this ABI for exactly these two ids - the way BasiciLQR.init takes ANY DynamicModelWrapper,
       basic_ilqr.py:50-66).  The main library answers CRL_ERR_UNSUPPORTED_MODEL for them.              */
    CRL_GEN_USER = 100,
    CRL_GEN_USER_BARRIER = 101
};
enum crl_dtype { CRL_F64 = 0, CRL_F32 = 1 };
enum crl_line_search { CRL_LS_VANILLA = 0, CRL_LS_FEASIBILITY = 1, CRL_LS_NONE = 2 }; /* ilqr_wrapper.py:74-146 */
enum crl_stopping { CRL_STOP_RELATIVE = 0, CRL_STOP_VANILLA = 1 };                    /* ilqr_wrapper.py:148-174 */

enum crl_status {
    CRL_OK = 0,
    CRL_ERR_BAD_ARGUMENT = -1,
    CRL_ERR_UNSUPPORTED_MODEL = -2,
    CRL_ERR_WORKSPACE_TOO_SMALL = -3,
    CRL_ERR_NO_DEVICE = -4,
    CRL_ERR_CUDA = -1000      /* -(1000 + cudaError_t) */
};

/* One batch of problems: which model, how many trajectories, the per-step cost parameters
 * (`add_param` of the reference, ilqr_obj_fun.py:43-52) and the action box (`constr`).   */
typedef struct crl_problem {
    int32_t model;            /* enum crl_model                                               */
    int32_t dtype;            /* enum crl_dtype                                               */
    int64_t batch;            /* B >= 0                                                       */
    int32_t horizon;          /* T >= 1                                                       */
    int32_t params_stride_t;  /* elements between time steps of `params`: p, or 0 = constant  */
    int64_t params_stride_b;  /* elements between trajectories of `params` (0 = shared)       */
    const void* params;       /* [B][T][p] (or broadcast per the strides)                     */
    double u_lo, u_hi;        /* action bounds, same for every action; +-inf = unconstrained  */
    const double* bounds_host; /* generated models (CRL_GEN_*) only, HOST memory, read during the call: the box
                                  `constr` as [d][2] rows of (lo, hi), states first - e.g. all +-inf for
                                  is_with_constraints=False; NULL = the scenario's default box (crl_model_bounds).
                                  The barrier variants are generated for the default box: NULL or equal to it.   */
} crl_problem_t;

/* BasiciLQR constructor arguments (basic_ilqr.py:13-20).                                      */
typedef struct crl_solver_opts {
    int32_t max_iter;
    int32_t is_check_stop;
    double stopping_criterion;
    int32_t max_line_search;
    double gamma;
    int32_t line_search_method; /* enum crl_line_search */
    int32_t stopping_method;    /* enum crl_stopping    */
    int32_t grid_blocks;        /* 0 = auto: resident CTAs per SM x SM count                   */
    int32_t reserved;           /* launch configuration of the solve kernel, 0 = chosen by     */
                                /* batch size (results do not depend on it).  Generated models: */
                                /* 2 = one independent solve per thread (the checker of the     */
                                /* default lock-step kernel with line-search helper lanes)      */
    int32_t coop_threshold;     /* once the batch queue is empty, a warp with at most this many
                                   unfinished trajectories hands them to the warp-cooperative
                                   engine (four trajectories per warp): 0 = default (12), < 0 = never.
                                   Scheduling only - results do not depend on it.                */
    int32_t reserved2;          /* must be 0                                                   */
} crl_solver_opts_t;

int crl_abi_version(void);
const char* crl_status_string(int status);
/* n, m, p of a model; CRL_ERR_UNSUPPORTED_MODEL for anything else (no fallback).               */
int crl_model_dims(int model, int* n, int* m, int* p);
/* Default action box of the scenario class (u in [-100,100] / [-500,500]).                      */
int crl_model_default_bounds(int model, double* u_lo, double* u_hi);
/* The box `constr` of a scenario, component by component: lo / hi receive d = n + m entries (states first).
 * For the arm models the states are unbounded and every action has the default box.               */
int crl_model_bounds(int model, double* lo, double* hi);

/* iLQRDynamicModel.eval_traj + iLQRObjectiveFunction.eval_obj_fun
 * (ilqr_dynamic_model.py:39-54,96-109; ilqr_obj_fun.py:54-62,86-95).
 * x0 [B][n]; u0 [T][m] (u0_stride_b = 0), [B][T][m] (u0_stride_b = T*m) or NULL = zeros;
 * out: X [B][T][d], J [B] (NULL to skip).                                                       */
int crl_rollout(const crl_problem_t* prob, const void* x0, const void* u0, int64_t u0_stride_b,
                void* X, double* J, crl_stream_t stream);

/* eval_obj_fun alone: J [B] of stored trajectories X [B][T][d].                                 */
int crl_cost(const crl_problem_t* prob, const void* X, double* J, crl_stream_t stream);

/* eval_grad_dynamic_model / eval_grad_obj_fun / eval_hessian_obj_fun
 * (ilqr_dynamic_model.py:73-81,136-147; ilqr_obj_fun.py:64-84,97-119).
 * out: F [B][T][n][d], c [B][T][d], C [B][T][d][d]; any may be NULL.                           */
int crl_derivs(const crl_problem_t* prob, const void* X, void* F, void* c, void* C, crl_stream_t stream);

/* iLQRWrapper.backward_pass with the derivatives recomputed on chip from X
 * (ilqr_wrapper.py:216-256).  out: K [B][T][m][n], k [B][T][m].                                 */
int crl_backward(const crl_problem_t* prob, const void* X, void* K, void* k, crl_stream_t stream);

/* iLQRWrapper.backward_pass(C, c, F) on caller-supplied dense matrices (same file:lines);
 * (n, m) must be one of the supported model shapes.  F [B][T][n][d], c [B][T][d], C [B][T][d][d]. */
int crl_backward_dense(int32_t n, int32_t m, int32_t dtype, int64_t batch, int32_t horizon,
                       const void* F, const void* c, const void* C, void* K, void* k, crl_stream_t stream);

/* iLQRDynamicModel.update_traj + eval_obj_fun for one step size per call
 * (ilqr_dynamic_model.py:56-71,111-134).  out: X_new [B][T][d], J_new [B] (NULL to skip).       */
int crl_update_traj(const crl_problem_t* prob, const void* X, const void* K, const void* k, double alpha,
                    void* X_new, double* J_new, crl_stream_t stream);

/* iLQRWrapper.forward_pass without the derivative re-evaluation: line search + stopping test
 * (ilqr_wrapper.py:74-214).  in: J_last [B]; out: X_new [B][T][d] (the old trajectory when no
 * step is accepted), J_new [B], alpha_index [B] (-1 = none accepted), stop [B] (0/1).            */
int crl_forward_pass(const crl_problem_t* prob, const crl_solver_opts_t* opts, const void* X, const void* K,
                     const void* k, const double* J_last, void* X_new, double* J_new, int32_t* alpha_index,
                     uint8_t* stop, crl_stream_t stream);

/* BasiciLQR.solve for the whole batch in one persistent kernel (basic_ilqr.py:68-97).
 *  in : x0 [B][n]; u0 as in crl_rollout; J_init [B] or NULL (= +inf, i.e. set_obj_fun_value(inf));
 *  out: X [B][T][d] final trajectories (NULL to skip), J [B], iters [B],
 *       converged [B] (1 = stopped by the criterion, 0 = ran max_iter),
 *       J_trace [B][max_iter] / alpha_trace [B][max_iter] per-iteration cost and accepted step
 *       index (-1 = rejected; unused slots untouched) - both may be NULL;
 *  workspace: at least crl_solve_workspace_bytes() bytes, 256-byte aligned.                     */
size_t crl_solve_workspace_bytes(const crl_problem_t* prob, const crl_solver_opts_t* opts);
int crl_solve(const crl_problem_t* prob, const crl_solver_opts_t* opts, const void* x0, const void* u0,
              int64_t u0_stride_b, const double* J_init, void* X, double* J, int32_t* iters,
              uint8_t* converged, double* J_trace, int8_t* alpha_trace, void* workspace,
              size_t workspace_bytes, crl_stream_t stream);

/* LogBarrieriLQR.solve for the whole batch in ONE persistent kernel (log_barrier_ilqr.py:108-149): the outer
 * loop over the barrier parameter t runs per trajectory inside the kernel.  Trajectory b executes one inner
 * iLQR loop per stage s = 0 .. n_stages-1 with the parameter row params[b][s] (prob->params is
 * [B][n_stages][p], params_stride_b = n_stages * p or 0 = shared, params_stride_t must be 0); between stages
 * it keeps its trajectory, restarts the iteration counter (opts->max_iter is per inner loop, :126) and resets
 * the stored cost to +inf if the loop ended by the stopping criterion (:140-141).  Barrier models only
 * (CRL_*_BARRIER): row s is (target_x, target_y, t_s, t_{s-1}), t_{-1} = t_0 (the stale value the first
 * backward pass of a loop still sees).  n_stages = 1 is crl_solve without traces.
 * Generated barrier models (CRL_GEN_*_BARRIER) may also carry time-varying rows: params [B][n_stages][T][p] with
 * params_stride_t = p (stage stride T * p, params_stride_b = n_stages * T * p or 0 = shared).
 *  out: as crl_solve (iters = total over the stages; J, converged = the last stage's) +
 *       stage_iters [B][n_stages] iterations of every inner loop.
 *  workspace: crl_solve_workspace_bytes() of the same prob / opts.                                    */
int crl_solve_stages(const crl_problem_t* prob, const crl_solver_opts_t* opts, int32_t n_stages, const void* x0,
                     const void* u0, int64_t u0_stride_b, const double* J_init, void* X, double* J, int32_t* iters,
                     uint8_t* converged, int32_t* stage_iters, void* workspace, size_t workspace_bytes,
                     crl_stream_t stream);

/* NNiLQRDynamicModel.eval_grad_dynamic_model (nn_ilqr.py:265-274) for a batch of points: value and Jacobian of the
 * learned dynamics  y = W3 relu(W2 relu(W1 x + b1) + b2) + b3  (SmallNetwork / LargeNetwork, nn_ilqr.py:66-100, in eval
 * mode, BatchNorm folded into W, b by the caller) at n_points rows x [n_points][d_in] - all FP32, like the reference's
 * `.float().cuda()` tensors.
 *  weights: w1 [h1][d_in], b1 [h1], w2 [h2][h1] and its transpose w2t [h1][h2], b2 [h2], w3 [n_out][h2], b3 [n_out];
 *           h1, h2 multiples of 32 (pad with zero rows / columns), h2 <= 512, d_in <= 16, n_out <= 16;
 *  out    : y [n_points][n_out], jac [n_points][n_out][d_in] (= F[t] of the reference, one [n][n+m] block per point);
 *  impl   : 1 = tensor cores (tcgen05.mma kind::tf32, FP32 accumulation in tensor memory; the product path),
 *           0 = FP32 CUDA cores (the checker of impl 1; a handful of points);
 *  workspace: crl_mlp_jacobian_workspace_bytes() bytes, 256-byte aligned (sign-bit masks of the two hidden layers;
 *           impl 0: also its activations).                                                                      */
size_t crl_mlp_jacobian_workspace_bytes(int32_t d_in, int32_t h1, int32_t h2, int32_t n_out, int64_t n_points, int32_t impl);
int crl_mlp_jacobian(int32_t d_in, int32_t h1, int32_t h2, int32_t n_out, int64_t n_points, const float* x, const float* w1,
                     const float* b1, const float* w2, const float* w2t, const float* b2, const float* w3, const float* b3,
                     float* y, float* jac, void* workspace, size_t workspace_bytes, int32_t impl, crl_stream_t stream);

/* Linear filter along the time axis of [batch][horizon][width] FP64 arrays: out[b][t][:] = sum_t' G[t][t'] in[b][t'][:],
 * G [horizon][horizon] row-major on the device.  NNiLQR.solve smooths the learned Jacobians with
 * scipy.ndimage.gaussian_filter1d(F, sigma, axis=0) (nn_ilqr.py:379-380); G is that filter's matrix.  in != out.  */
int crl_smooth_time(int64_t batch, int32_t horizon, int32_t width, const double* G, const double* in, double* out,
                    crl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CURIOUSRL_B200_H */
