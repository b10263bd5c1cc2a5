#define FC_REAL float
#define FC_JAC false
#define FC_FUNC clenshaw_f32_val
#include "eval_clenshaw_inst.cuh"
