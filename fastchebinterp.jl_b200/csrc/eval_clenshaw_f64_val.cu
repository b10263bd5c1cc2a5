#define FC_REAL double
#define FC_JAC false
#define FC_FUNC clenshaw_f64_val
#include "eval_clenshaw_inst.cuh"
