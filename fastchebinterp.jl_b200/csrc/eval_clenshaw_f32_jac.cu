#define FC_REAL float
#define FC_JAC true
#define FC_FUNC clenshaw_f32_jac
#include "eval_clenshaw_inst.cuh"
