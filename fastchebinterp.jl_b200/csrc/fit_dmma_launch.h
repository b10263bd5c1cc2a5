// Host-side entry points of the tensor-core DCT-I path for batches of short 1-D fits (fit_dmma.cuh).
#pragma once
#include "host_util.h"

namespace fastcheb {
// true when `nlines` lines of length n are better served by the folded-matrix DMMA transform
bool fit_dmma_applicable(int n, int E, long long howmany);
// lines are the (howmany x E) real sequences of length n stored with element stride E inside each fit
int fit_dmma_1d(const DeviceInfo* di, int dev, int n, long long nlines, int dtype, const void* vals, void* coefs, int E,
                cudaStream_t st);
}  // namespace fastcheb
