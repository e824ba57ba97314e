// Host build of the product's residual/Jacobian header (eventcalib_b200/csrc/ecb_residual.h) so the CPU test
// suite can compare the closed-form Jacobian with the dual-number oracle without a GPU.
#include "../../eventcalib_b200/csrc/ecb_residual_so3.h"
extern "C" double host_residual(const double *intr, const double *Q, const double *T, const double *b, const double *obs,
                                const double *lm, double radius, double huber, double *J33, double *cost, double *raw) {
    EcbResidualOut o = ecb_residual<true>(intr, Q, T, b, obs[0], obs[1], lm[0], lm[1], lm[2], radius, huber, J33);
    *cost = o.cost;
    *raw = o.raw;
    return o.res;
}
// CalibReprojectionError_SO3 variant (eventcalib_b200/csrc/ecb_residual_so3.h, ecb_so3.h)
extern "C" double host_residual_so3(const double *intr, const double *Q, const double *T, const double *b, const double *obs,
                                    const double *lm, double radius, double huber, double *J33, double *cost, double *raw) {
    EcbResidualOut o = ecb_residual_so3<true>(intr, Q, T, b, obs[0], obs[1], lm[0], lm[1], lm[2], radius, huber, J33);
    *cost = o.cost;
    *raw = o.raw;
    return o.res;
}
extern "C" void host_so3_plus(const double *x, const double *d, double *out) { ecb_so3::plus(x, d, out); }
