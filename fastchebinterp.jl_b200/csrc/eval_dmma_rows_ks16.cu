#define FC_KS 16
#define FC_FUNC dmma_rows_dispatch_ks16
#include "eval_dmma_rows_inst.cuh"
