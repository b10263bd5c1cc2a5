#define FC_KS 16
#define FC_JAC true
#define FC_FUNC dmma_dispatch_ks16_jac
#include "eval_dmma_inst.cuh"
