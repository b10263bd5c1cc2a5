#define FC_KS 12
#define FC_JAC true
#define FC_FUNC dmma2d_dispatch_ks12_jac
#include "eval_dmma2d_inst.cuh"
