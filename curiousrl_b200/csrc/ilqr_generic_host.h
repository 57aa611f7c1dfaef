// Internal interface between the C ABI (ilqr_kernels.cu) and the generated-model kernels (ilqr_generic.cu).
// The ABI functions validate arguments and select the device; these only size and launch.  Every launcher
// returns the cudaError_t of the launch as an int.
#pragma once

#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

#include "../../include/curiousrl_b200.h"

namespace crl {

bool gen_is_model(int model);                                    // CRL_GEN_* ids
bool gen_is_barrier(int model);                                  // CRL_GEN_*_BARRIER ids
int gen_dims(int model, int* n, int* m, int* p);                 // 0 = ok
int gen_bounds(int model, double* lo, double* hi);               // d entries each

int gen_launch_rollout(const crl_problem_t* prob, const double* x0, const double* u0, long long u0_sb, double* X,
                       double* J, cudaStream_t s);
int gen_launch_cost(const crl_problem_t* prob, const double* X, double* J, cudaStream_t s);
int gen_launch_derivs(const crl_problem_t* prob, const double* X, double* F, double* c, double* C, cudaStream_t s);
int gen_launch_backward(const crl_problem_t* prob, const double* X, double* K, double* k, cudaStream_t s);
int gen_launch_update(const crl_problem_t* prob, const double* X, const double* K, const double* k, double alpha,
                      double* Xn, double* Jn, cudaStream_t s);
int gen_launch_forward(const crl_problem_t* prob, const crl_solver_opts_t* opts, const double* X, const double* K,
                       const double* k, const double* J_last, double* Xn, double* J_new, int32_t* alpha_index,
                       uint8_t* stop, cudaStream_t s);
size_t gen_solve_workspace_bytes(const crl_problem_t* prob);
int gen_launch_solve(const crl_problem_t* prob, const crl_solver_opts_t* opts, const double* x0, const double* u0,
                     long long u0_sb, const double* J_init, double* X, double* J, int32_t* iters, uint8_t* converged,
                     double* J_trace, int8_t* alpha_trace, void* workspace, cudaStream_t s, int n_stage,
                     int32_t* stage_iters);

}  // namespace crl
