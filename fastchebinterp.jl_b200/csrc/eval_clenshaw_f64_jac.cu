#define FC_REAL double
#define FC_JAC true
#define FC_FUNC clenshaw_f64_jac
#include "eval_clenshaw_inst.cuh"
