// The C ABI of include/curiousrl_b200.h for ONE scenario compiled at run time (curiousrl_b200/jit.py).
//
// The reference's BasiciLQR.init takes any DynamicModelWrapper (basic_ilqr.py:50-66, ilqr_dynamic_model.py:25-37: it
// lambdifies the scenario's sympy expressions and JIT-compiles them with numba on first use).  The counterpart here: the
// generator that writes csrc/ilqr_generated.cuh prints the scenario as structs GenUser / GenUserBarrier, nvcc compiles
// this unit (the general-form kernels of ilqr_generic.cu over that header) into a library of its own, and the Python
// session binds it with the same ctypes declarations as the main library.  Model ids: CRL_GEN_USER, CRL_GEN_USER_BARRIER.
//
//   nvcc ... -DCRL_USER_HEADER="\"/path/model.cuh\"" -shared -o model.so ilqr_user_abi.cu
#define CRL_USER_MODEL 1
#ifndef CRL_GEN_MAX_D
#define CRL_GEN_MAX_D 16
#endif
#include "ilqr_generic.cu"

namespace crl {
namespace user_abi {

int cuda_status(cudaError_t e) { return e == cudaSuccess ? CRL_OK : CRL_ERR_CUDA - (int)e; }

int check_problem(const crl_problem_t* p) {
    if (!p) return CRL_ERR_BAD_ARGUMENT;
    if (!gen_is_model(p->model)) return CRL_ERR_UNSUPPORTED_MODEL;
    int n = 0, m = 0, pw = 0;
    gen_dims(p->model, &n, &m, &pw);
    if (p->dtype != CRL_F64) return CRL_ERR_UNSUPPORTED_MODEL;
    if (p->batch < 0 || p->horizon < 1 || (p->batch > 0 && !p->params)) return CRL_ERR_BAD_ARGUMENT;
    if ((p->params_stride_t != 0 && p->params_stride_t != pw) || p->params_stride_b < 0) return CRL_ERR_BAD_ARGUMENT;
    if (p->bounds_host) {
        double lo[CRL_GEN_MAX_D], hi[CRL_GEN_MAX_D];
        gen_bounds(p->model, lo, hi);
        for (int i = 0; i < n + m; ++i) {
            const double bl = p->bounds_host[2 * i], bh = p->bounds_host[2 * i + 1];
            if (!(bl <= bh)) return CRL_ERR_BAD_ARGUMENT;
            if (gen_is_barrier(p->model) && (bl != lo[i] || bh != hi[i])) return CRL_ERR_UNSUPPORTED_MODEL;
        }
    }
    return CRL_OK;
}

int check_opts(const crl_solver_opts_t* o) {
    if (!o) return CRL_ERR_BAD_ARGUMENT;
    if (o->max_iter < 0 || o->max_line_search < 0 || o->reserved2 != 0) return CRL_ERR_BAD_ARGUMENT;
    if (o->line_search_method < CRL_LS_VANILLA || o->line_search_method > CRL_LS_NONE) return CRL_ERR_BAD_ARGUMENT;
    if (o->stopping_method < CRL_STOP_RELATIVE || o->stopping_method > CRL_STOP_VANILLA) return CRL_ERR_BAD_ARGUMENT;
    return CRL_OK;
}

// makes the device that owns `ptr` current for the call and rejects host pointers (as the main library does)
struct DeviceScope {
    int prev = -1, status = CRL_OK;
    bool changed = false;
    explicit DeviceScope(const void* ptr) {
        cudaPointerAttributes at;
        cudaError_t e = cudaPointerGetAttributes(&at, ptr);
        if (e != cudaSuccess) {
            cudaGetLastError();
            status = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? CRL_ERR_NO_DEVICE : cuda_status(e);
            return;
        }
        if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) {
            status = CRL_ERR_BAD_ARGUMENT;
            return;
        }
        if (cudaGetDevice(&prev) == cudaSuccess && prev != at.device) {
            status = cuda_status(cudaSetDevice(at.device));
            changed = (status == CRL_OK);
        }
    }
    ~DeviceScope() {
        if (changed) cudaSetDevice(prev);
    }
};

__global__ void __launch_bounds__(128) user_backward_dense_kernel(long long B, int T, const double* F, const double* c,
                                                                  const double* C, double* K, double* k) {
    constexpr int N = GenUser::N, M = GenUser::M, D = N + M;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    backward_dense<N, M, double>(F + b * T * N * D, c + b * T * D, C + b * T * D * D, T, K + b * T * M * N, k + b * T * M);
}

}  // namespace user_abi
}  // namespace crl

using namespace crl;
using namespace crl::user_abi;

#define CRL_USER_CALL(first_ptr, launch)                           \
    int st = check_problem(prob);                                  \
    if (st) return st;                                             \
    if (prob->batch == 0) return CRL_OK;                           \
    DeviceScope scope(first_ptr);                                  \
    if (scope.status) return scope.status;                         \
    return cuda_status((cudaError_t)(launch))

extern "C" {

int crl_abi_version(void) { return CRL_ABI_VERSION; }

const char* crl_status_string(int status) {
    switch (status) {
        case CRL_OK: return "ok";
        case CRL_ERR_BAD_ARGUMENT: return "bad argument";
        case CRL_ERR_UNSUPPORTED_MODEL: return "unsupported model (this library holds one run-time compiled scenario)";
        case CRL_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
        case CRL_ERR_NO_DEVICE: return "no CUDA device";
        default: break;
    }
    if (status <= CRL_ERR_CUDA) return cudaGetErrorString((cudaError_t)(CRL_ERR_CUDA - status));
    return "unknown status";
}

int crl_model_dims(int model, int* n, int* m, int* p) { return gen_dims(model, n, m, p) == 0 ? CRL_OK : CRL_ERR_UNSUPPORTED_MODEL; }
int crl_model_default_bounds(int, double*, double*) { return CRL_ERR_UNSUPPORTED_MODEL; }
int crl_model_bounds(int model, double* lo, double* hi) { return gen_bounds(model, lo, hi) == 0 ? CRL_OK : CRL_ERR_UNSUPPORTED_MODEL; }

int crl_rollout(const crl_problem_t* prob, const void* x0, const void* u0, int64_t u0_stride_b, void* X, double* J,
                crl_stream_t stream) {
    if (prob && prob->batch > 0 && (!x0 || !X)) return CRL_ERR_BAD_ARGUMENT;
    CRL_USER_CALL(x0, gen_launch_rollout(prob, (const double*)x0, (const double*)u0, u0_stride_b, (double*)X, J, (cudaStream_t)stream));
}
int crl_cost(const crl_problem_t* prob, const void* X, double* J, crl_stream_t stream) {
    if (prob && prob->batch > 0 && (!X || !J)) return CRL_ERR_BAD_ARGUMENT;
    CRL_USER_CALL(X, gen_launch_cost(prob, (const double*)X, J, (cudaStream_t)stream));
}
int crl_derivs(const crl_problem_t* prob, const void* X, void* F, void* c, void* C, crl_stream_t stream) {
    if (prob && prob->batch > 0 && !X) return CRL_ERR_BAD_ARGUMENT;
    CRL_USER_CALL(X, gen_launch_derivs(prob, (const double*)X, (double*)F, (double*)c, (double*)C, (cudaStream_t)stream));
}
int crl_backward(const crl_problem_t* prob, const void* X, void* K, void* k, crl_stream_t stream) {
    if (prob && prob->batch > 0 && (!X || !K || !k)) return CRL_ERR_BAD_ARGUMENT;
    CRL_USER_CALL(X, gen_launch_backward(prob, (const double*)X, (double*)K, (double*)k, (cudaStream_t)stream));
}
int crl_backward_dense(int32_t n, int32_t m, int32_t dtype, int64_t batch, int32_t horizon, const void* F, const void* c,
                       const void* C, void* K, void* k, crl_stream_t stream) {
    if (batch < 0 || horizon < 1) return CRL_ERR_BAD_ARGUMENT;
    if (n != GenUser::N || m != GenUser::M || dtype != CRL_F64) return CRL_ERR_UNSUPPORTED_MODEL;
    if (batch == 0) return CRL_OK;
    if (!F || !c || !C || !K || !k) return CRL_ERR_BAD_ARGUMENT;
    DeviceScope scope(F);
    if (scope.status) return scope.status;
    user_backward_dense_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        batch, horizon, (const double*)F, (const double*)c, (const double*)C, (double*)K, (double*)k);
    return cuda_status(cudaGetLastError());
}
int crl_update_traj(const crl_problem_t* prob, const void* X, const void* K, const void* k, double alpha, void* X_new,
                    double* J_new, crl_stream_t stream) {
    if (prob && prob->batch > 0 && (!X || !K || !k || !X_new)) return CRL_ERR_BAD_ARGUMENT;
    CRL_USER_CALL(X, gen_launch_update(prob, (const double*)X, (const double*)K, (const double*)k, alpha, (double*)X_new, J_new,
                                       (cudaStream_t)stream));
}
int crl_forward_pass(const crl_problem_t* prob, const crl_solver_opts_t* opts, const void* X, const void* K,
                     const void* k, const double* J_last, void* X_new, double* J_new, int32_t* alpha_index,
                     uint8_t* stop, crl_stream_t stream) {
    const int so = check_opts(opts);
    if (so) return so;
    if (prob && prob->batch > 0 && (!X || !K || !k || !J_last || !X_new || !J_new || !alpha_index || !stop)) return CRL_ERR_BAD_ARGUMENT;
    CRL_USER_CALL(X, gen_launch_forward(prob, opts, (const double*)X, (const double*)K, (const double*)k, J_last, (double*)X_new,
                                        J_new, alpha_index, stop, (cudaStream_t)stream));
}

size_t crl_solve_workspace_bytes(const crl_problem_t* prob, const crl_solver_opts_t* opts) {
    if (check_problem(prob) || check_opts(opts)) return 0;
    return gen_solve_workspace_bytes(prob);
}

static int solve_impl(const crl_problem_t* prob, const crl_solver_opts_t* opts, int n_stage, int32_t* stage_iters, const void* x0,
                      const void* u0, int64_t u0_stride_b, const double* J_init, void* X, double* J, int32_t* iters,
                      uint8_t* converged, double* J_trace, int8_t* alpha_trace, void* workspace, size_t workspace_bytes,
                      crl_stream_t stream) {
    int st = check_problem(prob);
    if (st) return st;
    st = check_opts(opts);
    if (st) return st;
    if (n_stage > 1) {
        if (!gen_is_barrier(prob->model)) return CRL_ERR_UNSUPPORTED_MODEL;
        if (n_stage >= 128 || J_trace || alpha_trace) return CRL_ERR_BAD_ARGUMENT;
        if (prob->batch > 0 && !stage_iters) return CRL_ERR_BAD_ARGUMENT;
    }
    if (prob->batch == 0) return CRL_OK;
    if (!x0 || !J || !iters || !converged || !workspace || opts->max_iter < 1) return CRL_ERR_BAD_ARGUMENT;
    if (((uintptr_t)workspace & 255u) != 0) return CRL_ERR_BAD_ARGUMENT;
    if (workspace_bytes < gen_solve_workspace_bytes(prob)) return CRL_ERR_WORKSPACE_TOO_SMALL;
    DeviceScope scope(x0);
    if (scope.status) return scope.status;
    return cuda_status((cudaError_t)gen_launch_solve(prob, opts, (const double*)x0, (const double*)u0, u0_stride_b, J_init,
                                                     (double*)X, J, iters, converged, J_trace, alpha_trace, workspace,
                                                     (cudaStream_t)stream, n_stage, stage_iters));
}

int crl_solve(const crl_problem_t* prob, const crl_solver_opts_t* opts, const void* x0, const void* u0, int64_t u0_stride_b,
              const double* J_init, void* X, double* J, int32_t* iters, uint8_t* converged, double* J_trace,
              int8_t* alpha_trace, void* workspace, size_t workspace_bytes, crl_stream_t stream) {
    return solve_impl(prob, opts, 1, nullptr, x0, u0, u0_stride_b, J_init, X, J, iters, converged, J_trace, alpha_trace,
                      workspace, workspace_bytes, stream);
}

int crl_solve_stages(const crl_problem_t* prob, const crl_solver_opts_t* opts, int32_t n_stages, const void* x0, const void* u0,
                     int64_t u0_stride_b, const double* J_init, void* X, double* J, int32_t* iters, uint8_t* converged,
                     int32_t* stage_iters, void* workspace, size_t workspace_bytes, crl_stream_t stream) {
    if (n_stages < 1) return CRL_ERR_BAD_ARGUMENT;
    return solve_impl(prob, opts, n_stages, stage_iters, x0, u0, u0_stride_b, J_init, X, J, iters, converged, nullptr, nullptr,
                      workspace, workspace_bytes, stream);
}

}  // extern "C"
