// TEST-ONLY host build of the generic (generated-model) device math: curiousrl_b200/csrc/ilqr_generic.cuh +
// ilqr_generated.cuh compiled with g++, so that the `-m "not gpu"` tests can check the generated models and the
// general-form iteration against the oracle / the reference's golden vectors without a GPU.  Not product code.
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../curiousrl_b200/csrc/ilqr_generic.cuh"

using namespace crl;

namespace {
bool g_box_on = false;
GenBox g_box;
template <class G>
GenParams hp(const double* params, int stride_t) {
    return GenParams{params, stride_t, g_box_on ? g_box : gen_default_box<G>()};
}
}  // namespace
// the box of the following calls: n_comp (lo, hi) rows, or n_comp = 0 for the model's default
extern "C" void hg_set_box(const double* rows, int n_comp) {
    g_box_on = n_comp > 0;
    for (int i = 0; i < 8; ++i) { g_box.lo[i] = i < n_comp ? rows[2 * i] : 0.0; g_box.hi[i] = i < n_comp ? rows[2 * i + 1] : 0.0; }
}

#define HG_DISPATCH(model, ...)                                                     \
    switch (model) {                                                                \
        case 6: { using G = GenCartPoleSwingUp1; __VA_ARGS__; } break;              \
        case 7: { using G = GenCartPoleSwingUp2; __VA_ARGS__; } break;              \
        case 8: { using G = GenVehicleTracking; __VA_ARGS__; } break;               \
        case 9: { using G = GenCarParking; __VA_ARGS__; } break;                    \
        case 10: { using G = GenCartPoleSwingUp1Barrier; __VA_ARGS__; } break;      \
        case 11: { using G = GenCartPoleSwingUp2Barrier; __VA_ARGS__; } break;      \
        case 12: { using G = GenVehicleTrackingBarrier; __VA_ARGS__; } break;       \
        case 13: { using G = GenCarParkingBarrier; __VA_ARGS__; } break;            \
        default: return -2;                                                         \
    }

extern "C" int hg_dims(int model, int* n, int* m, int* p) {
    HG_DISPATCH(model, (*n = G::N, *m = G::M, *p = G::P));
    return 0;
}
extern "C" int hg_bounds(int model, double* lo, double* hi) {
    HG_DISPATCH(model, { for (int i = 0; i < G::D; ++i) { lo[i] = G::lo(i); hi[i] = G::hi(i); } });
    return 0;
}
extern "C" int hg_rollout(int model, const double* x0, const double* u0, const double* params, int stride_t, int T,
                          double* X, double* J) {
    HG_DISPATCH(model, *J = (gen_rollout_open<G>(x0, u0, G::M, hp<G>(params, stride_t), T, View<double, G::D, 1>{X})));
    return 0;
}
extern "C" int hg_derivs(int model, const double* X, const double* params, int stride_t, int T, double* F, double* c,
                         double* C) {
    HG_DISPATCH(model, {
        for (int t = 0; t < T; ++t) {
            double p[G::P];
            hp<G>(params, stride_t).load(t, p);
            gen_derivs<G>(X + (size_t)t * G::D, p, F + (size_t)t * G::N * G::D, c + (size_t)t * G::D,
                          C + (size_t)t * G::D * G::D);
        }
    });
    return 0;
}
extern "C" int hg_backward(int model, const double* X, const double* params, int stride_t, int T, double* K, double* k) {
    HG_DISPATCH(model, (gen_backward<G>(View<double, G::D, 1>{const_cast<double*>(X)}, hp<G>(params, stride_t), T,
                                        View<double, G::M * G::N, 1>{K}, View<double, G::M, 1>{k})));
    return 0;
}
extern "C" int hg_update(int model, const double* X, const double* K, const double* k, double alpha, const double* params,
                         int stride_t, int T, double* Xn, double* J) {
    HG_DISPATCH(model, *J = (gen_rollout_closed<G>(View<double, G::D, 1>{const_cast<double*>(X)},
                                                   View<double, G::M * G::N, 1>{const_cast<double*>(K)},
                                                   View<double, G::M, 1>{const_cast<double*>(k)}, alpha,
                                                   hp<G>(params, stride_t), T, View<double, G::D, 1>{Xn})));
    return 0;
}
extern "C" int hg_solve(int model, const double* x0, const double* u0, const double* params, int stride_t, int T,
                        int max_iter, int check_stop, double criterion, int max_ls, double gamma, int ls_method,
                        int stop_method, double J_init, double* X_out, double* J_out, int* iters_out, int* conv_out,
                        double* J_trace, signed char* alpha_trace, int n_stage, long long stage_stride, int* stage_iters) {
    HG_DISPATCH(model, {
        std::vector<double> xa((size_t)T * G::D), xb((size_t)T * G::D), Kb((size_t)T * G::M * G::N), kb((size_t)T * G::M);
        GenOpts op{max_iter, check_stop, max_ls, ls_method, stop_method, criterion, gamma};
        bool in_a = true;
        gen_solve_one<G>(x0, u0, G::M, hp<G>(params, stride_t), T, op, J_init, View<double, G::D, 1>{xa.data()},
                         View<double, G::D, 1>{xb.data()}, View<double, G::M * G::N, 1>{Kb.data()},
                         View<double, G::M, 1>{kb.data()}, J_out, iters_out, conv_out, &in_a, J_trace, alpha_trace, n_stage,
                         stage_stride, stage_iters);
        const std::vector<double>& fin = in_a ? xa : xb;
        for (size_t i = 0; i < fin.size(); ++i) X_out[i] = fin[i];
    });
    return 0;
}
