// cost evaluation kernels (placeholder until the residual/Jacobian path lands)
#include "ecb_common.cuh"
extern "C" void ecb_cost_free(ecb_ctx *ctx) { (void) ctx; }
