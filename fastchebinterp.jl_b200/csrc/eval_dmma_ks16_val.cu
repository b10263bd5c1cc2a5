#define FC_KS 16
#define FC_JAC false
#define FC_FUNC dmma_dispatch_ks16_val
#include "eval_dmma_inst.cuh"
