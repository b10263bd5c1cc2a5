#define FC_KS 12
#define FC_JAC false
#define FC_FUNC dmma_dispatch_ks12_val
#include "eval_dmma_inst.cuh"
