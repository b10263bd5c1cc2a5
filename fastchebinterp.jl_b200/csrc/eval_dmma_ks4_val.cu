#define FC_KS 4
#define FC_JAC false
#define FC_FUNC dmma_dispatch_ks4_val
#include "eval_dmma_inst.cuh"
