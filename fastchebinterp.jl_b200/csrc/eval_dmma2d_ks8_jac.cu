#define FC_KS 8
#define FC_JAC true
#define FC_FUNC dmma2d_dispatch_ks8_jac
#include "eval_dmma2d_inst.cuh"
