#define FC_KS 4
#define FC_FUNC dmma_rows_dispatch_ks4
#include "eval_dmma_rows_inst.cuh"
