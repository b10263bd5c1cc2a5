#define FC_KS 8
#define FC_FUNC dmma_rows_dispatch_ks8
#include "eval_dmma_rows_inst.cuh"
