#define FC_KS 12
#define FC_JAC true
#define FC_FUNC dmma_dispatch_ks12_jac
#include "eval_dmma_inst.cuh"
