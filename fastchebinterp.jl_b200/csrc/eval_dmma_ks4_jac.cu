#define FC_KS 4
#define FC_JAC true
#define FC_FUNC dmma_dispatch_ks4_jac
#include "eval_dmma_inst.cuh"
