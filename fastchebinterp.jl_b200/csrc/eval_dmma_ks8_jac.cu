#define FC_KS 8
#define FC_JAC true
#define FC_FUNC dmma_dispatch_ks8_jac
#include "eval_dmma_inst.cuh"
