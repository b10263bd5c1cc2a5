/* libfastcheb_cuda — C ABI of the B200 (sm_100a) hot path of FastChebInterp.jl.
 *
 * This is the drop-in boundary: what a Julia `src/cuda.jl` would `ccall` (INTEGRATION.md shows the
 * binding).  The reference has no FFI layer of its own — its boundary is Julia's method table — so
 * every entry point below names the reference method it replaces (file:line under the reference
 * tree, FastChebInterp.jl v1.3.1).
 *
 * Conventions
 *  - All functions return an `int` status (FASTCHEB_OK == 0).  A message for the last failure on the
 *    calling thread is available from fastcheb_last_error(); the index of the first offending point
 *    (EDOMAIN) or value (ENONFINITE) from fastcheb_error_index().
 *  - Memory layouts are Julia's: coefficient arrays are column-major with
 *    E = ncomp*(is_complex?2:1) reals stored inline per element (component index fastest, re/im
 *    fastest of all); points are AoS `npts x ndim` reals (`Vector{SVector{N,T}}`); values are AoS
 *    `npts x E` reals; Jacobians are `npts x (E x ndim)` with each E x ndim block column-major
 *    (`SMatrix{K,N}` layout, J[i,j] at i + K*j).
 *  - `dtype` selects the real scalar type of coefficients, points and outputs alike (promotion of
 *    mixed precisions is the host wrapper's job, as it is Julia's in the reference).
 *  - `mem` says where the caller's buffers live: FASTCHEB_MEM_HOST (pageable or pinned host memory;
 *    the library stages through the GPU and returns after the results are back) or
 *    FASTCHEB_MEM_DEVICE (device pointers on the handle's device; work is enqueued on `stream`).
 *  - `stream` is a `cudaStream_t` passed as `void*` (NULL = the legacy default stream).
 *  - The caller owns every buffer it passes; the library copies coefficients at create time and never
 *    retains caller pointers after a call returns (for *_async: after the stream work completes).
 *  - There is no CPU fallback: every compute entry point fails with FASTCHEB_ECUDA if no sm_100
 *    device is usable.
 */
#ifndef FASTCHEB_H
#define FASTCHEB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FASTCHEB_VERSION 100 /* 0.1.0 */

/* real scalar type */
#define FASTCHEB_F32 0
#define FASTCHEB_F64 1

/* where caller buffers live */
#define FASTCHEB_MEM_HOST 0
#define FASTCHEB_MEM_DEVICE 1

/* status codes (-> the Julia exception the wrapper re-throws) */
#define FASTCHEB_OK 0
#define FASTCHEB_EDOMAIN 1      /* point outside [lb,ub] or NaN   -> ArgumentError  (eval.jl:64,148) */
#define FASTCHEB_ENONFINITE 2   /* non-finite value in `vals`     -> DomainError    (interp.jl:40)   */
#define FASTCHEB_EDIMMISMATCH 3 /* gradient of a vector-valued c  -> DimensionMismatch (eval.jl:170) */
#define FASTCHEB_EINVAL 4       /* bad argument (null, rank, dtype, negative size) -> ArgumentError  */
#define FASTCHEB_ECUDA 5        /* CUDA runtime failure / no usable device                          */
#define FASTCHEB_ENOMEM 6       /* host or device allocation failed                                 */
#define FASTCHEB_EUNSUPPORTED 7 /* a shape the kernels do not cover (ndim > FASTCHEB_MAX_NDIM ...)   */

#define FASTCHEB_MAX_NDIM 6

/* evaluation algorithm selector for fastcheb_set_algo (diagnostics / benchmarking only) */
#define FASTCHEB_ALGO_AUTO 0
#define FASTCHEB_ALGO_CLENSHAW 1 /* CUDA-core Clenshaw, one thread per point            */
#define FASTCHEB_ALGO_DMMA 2     /* FP64 tensor-core first-dimension contraction (F64)   */
#define FASTCHEB_ALGO_DMMA_RING 3 /* same, but never the shared-memory-resident 2-D kernel */
#define FASTCHEB_ALGO_PRODUCT 4  /* 1-D, 40 <= n <= 512: product-basis contraction, one DFMA per coefficient       */

typedef struct fastcheb_poly fastcheb_poly; /* device-resident ChebPoly{N,T,Td} (FastChebInterp.jl:37-41) */

/* ---- library / device ------------------------------------------------------------------------ */
int fastcheb_version(void);
/* kernels launched by the library in this process so far (diagnostics; bench.py's gpu_launches) */
int64_t fastcheb_launch_count(void);
/* number of usable CUDA devices (0 if none; never fails) */
int fastcheb_device_count(void);
/* copy the calling thread's last error message (NUL-terminated, truncated to n) ; returns its length */
size_t fastcheb_last_error(char* buf, size_t n);
/* index (0-based) of the first offending point / value of the calling thread's last EDOMAIN / ENONFINITE */
int fastcheb_error_index(int64_t* idx);

/* ---- the ChebPoly handle ----------------------------------------------------------------------
 * Mirrors `struct ChebPoly{N,T,Td}` (FastChebInterp.jl:37-41): coefs::Array{T,N}, lb, ub.
 * dims[d] = size(coefs, d+1) >= 1.  `coefs` is read according to `coefs_mem`.  lb/ub travel as double
 * (Td may be an integer type in the reference, runtests.jl:170).  The handle is immutable after
 * creation; any number of host threads may evaluate it concurrently. */
int fastcheb_create(fastcheb_poly** out, int ndim, const int64_t* dims, int dtype, int is_complex,
                    int ncomp, const void* coefs, int coefs_mem, const double* lb, const double* ub,
                    int device);
int fastcheb_destroy(fastcheb_poly* p); /* NULL is a no-op; safe from a finalizer thread */

/* handle queries */
int fastcheb_ndim(const fastcheb_poly* p);
int fastcheb_dims(const fastcheb_poly* p, int64_t* dims_out);
int fastcheb_dtype(const fastcheb_poly* p);
int fastcheb_is_complex(const fastcheb_poly* p);
int fastcheb_ncomp(const fastcheb_poly* p);
int fastcheb_device(const fastcheb_poly* p);
/* copy the coefficients back in Julia layout (dst on host or device per dst_mem) */
int fastcheb_get_coefs(const fastcheb_poly* p, void* dst, int dst_mem);
/* force an evaluation algorithm (default AUTO); returns EUNSUPPORTED if the handle cannot run it.  A diagnostic /
 * benchmarking switch: it selects among kernels that agree within the stated tolerance and may be called at any time
 * (the selection is an atomic), but results of evaluations racing with it may come from either kernel. */
int fastcheb_set_algo(fastcheb_poly* p, int algo);
/* which algorithm AUTO resolves to for value (jac=0) / Jacobian (jac=1) calls */
int fastcheb_get_algo(const fastcheb_poly* p, int jac);
/* The 1-D product-basis path (FASTCHEB_ALGO_PRODUCT) is taken by AUTO only after a plan-time self-check against the
 * bit-exact Clenshaw kernel on 6144 probe points (uniform, Chebyshev-clustered, and 2^-1..2^-53 of the span from both
 * end points): this returns the check's worst |dv| / (16 eps sum|c|) and worst |dv'| / (16 eps sum|c_k| k^2 2/(ub-lb));
 * AUTO accepts values (derivatives) at <= 0.25. */
int fastcheb_product_check(const fastcheb_poly* p, double* dv_over_tol, double* dJ_over_tol);

/* ---- batched evaluation -----------------------------------------------------------------------
 * fastcheb_eval      : c.(ys)                       — (::ChebPoly{N})(x) eval.jl:63-70 per point
 * fastcheb_eval_jac  : chebjacobian.(Ref(c), ys)    — eval.jl:147-155; out_v npts x E, out_J npts x E x ndim
 * fastcheb_eval_grad : chebgradient.(Ref(c), ys)    — eval.jl:168-177; requires ncomp == 1 (else EDIMMISMATCH)
 * The *_tuple variants write one interleaved record per point, `E` reals of v followed by `E*ndim`
 * reals of J — byte-identical to Julia's `Vector{Tuple{V,SMatrix{K,N}}}` / `Vector{Tuple{T,SVector{N}}}`.
 * These calls are synchronous with respect to the result status: they return after the domain check
 * of every point is known (FASTCHEB_EDOMAIN + fastcheb_error_index on failure; outputs are then
 * unspecified — the reference's broadcast aborts on the first throw). */
int fastcheb_eval(const fastcheb_poly* p, const void* pts, int64_t npts, void* out, int mem, void* stream);
int fastcheb_eval_jac(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, void* out_J,
                      int mem, void* stream);
int fastcheb_eval_grad(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_v, void* out_g,
                       int mem, void* stream);
int fastcheb_eval_jac_tuple(const fastcheb_poly* p, const void* pts, int64_t npts, void* out_vJ, int mem,
                            void* stream);

/* Asynchronous device-memory variants: enqueue on `stream` and return immediately.  `status_dev`
 * (device pointer to one int64, may be NULL) receives the index of the first out-of-domain point, or
 * INT64_MAX if all points were inside; the caller initialises nothing and reads it after syncing. */
int fastcheb_eval_async(const fastcheb_poly* p, const void* pts_dev, int64_t npts, void* out_dev,
                        int64_t* status_dev, void* stream);
int fastcheb_eval_jac_async(const fastcheb_poly* p, const void* pts_dev, int64_t npts, void* out_v_dev,
                            void* out_J_dev, int64_t* status_dev, void* stream);

/* Batched AD rules of c(x) (chainrules.jl:4-39) with the Jacobian contracted inside the kernel — it is never written
 * to HBM (24 B instead of 96 B of output per point for the headline polynomial):
 * fastcheb_eval_vjp : rrule pullback, out_xbar[p, :] = real(J_p' * dy[p, :])   (chainrules.jl:7,14); dy is npts x E
 *                     reals (complex cotangents as re,im pairs), out_xbar npts x ndim.
 * fastcheb_eval_jvp : frule,          out_ydot[p, :] = J_p * dx[p, :]          (chainrules.jl:24); dx is npts x ndim,
 *                     out_ydot npts x E.  (The reference's extra terms for a perturbed ChebPoly, chainrules.jl:27-37,
 *                     are one more jvp with a shifted dx plus one fastcheb_eval of the coefficient tangent.)
 * out_v (npts x E) receives the values.  Synchronous; EDOMAIN as for fastcheb_eval. */
int fastcheb_eval_vjp(const fastcheb_poly* p, const void* pts, int64_t npts, const void* dy, void* out_v, void* out_xbar,
                      int mem, void* stream);
int fastcheb_eval_jvp(const fastcheb_poly* p, const void* pts, int64_t npts, const void* dx, void* out_v, void* out_ydot,
                      int mem, void* stream);

/* fastcheb_eval_many : many independent 1-D polynomials, each evaluated at its own points, in one launch — the
 * reference spells this `[c.(xs) for (c, xs) in zip(cs, xss)]` with (::ChebPoly{1})(x) eval.jl:63-70 per point (and
 * chebgradient eval.jl:174-177 when out_d != NULL: d/dx of every component).  `coefs` holds `howmany` coefficient
 * lines of n elements (E inline reals each) back to back — exactly what fastcheb_fit writes for a batch of 1-D fits;
 * pts is howmany x M reals, out_v / out_d are howmany x M x E.  lb/ub: one pair (bounds_per_poly = 0) or `howmany`
 * pairs.  mem = HOST: synchronous, EDOMAIN + fastcheb_error_index (flat point index) on a point outside its box.
 * mem = DEVICE: every pointer (lb/ub included) is a device pointer, the call only enqueues, and status_dev (may be
 * NULL) receives the first offending flat point index or INT64_MAX. */
int fastcheb_eval_many(int64_t howmany, int64_t n, int dtype, int is_complex, int ncomp, const void* coefs,
                       const double* lb, const double* ub, int bounds_per_poly, const void* pts, int64_t M,
                       void* out_v, void* out_d, int64_t* status_dev, int mem, int device, void* stream);

/* fastcheb_eval_many_shared : the same batch when every polynomial is evaluated at the SAME M points and shares one
 * box — `[c.(xs) for c in cs]`: V[f, m] = sum_k C[f, k] T_k(z_m) is then a (howmany x n) x (n x M) matrix product and runs
 * on the FP64 tensor cores (n <= 65).  pts: M reals; lb / ub: one pair (host doubles for mem = HOST, device doubles for
 * mem = DEVICE); out_v / out_d: howmany x M x E.  Status and errors as fastcheb_eval_many (the index is the point index).
 * Values agree with the per-polynomial Clenshaw evaluation within 16 eps sum|c| (another summation order), not bit
 * for bit. */
int fastcheb_eval_many_shared(int64_t howmany, int64_t n, int dtype, int is_complex, int ncomp, const void* coefs,
                              const double* lb, const double* ub, const void* pts, int64_t M, void* out_v, void* out_d,
                              int64_t* status_dev, int mem, int device, void* stream);

/* ---- the fit ----------------------------------------------------------------------------------
 * fastcheb_fit : chebcoefs(vals) (interp.jl:39-71) for `howmany` independent arrays stored back to
 * back, each `dims[0] x ... x dims[ndim-1]` elements of E inline reals (so `chebinterp_v1`'s
 * K x n1 x ... input, interp.jl:155-156, is ncomp = K).  Computes the N-d DCT-I (FFTW REDFT00 along
 * every dim with n > 1, identity along n == 1, interp.jl:43-44), the 1/prod 2(n-1) scaling (:49) and the
 * interior doubling (:50-54), every real component independently.  `vals` and `coefs` may alias.
 * maxabs (may be NULL; `howmany` doubles, same memory kind as coefs): max_k infnorm(C_k) of each fit
 * (interp.jl:74-78; complex modulus for complex elements), the quantity droptol scales `tol` by.
 * Non-finite input -> FASTCHEB_ENONFINITE (+ flat real index via fastcheb_error_index). */
int fastcheb_fit(int ndim, const int64_t* dims, int dtype, int is_complex, int ncomp, int64_t howmany,
                 const void* vals, void* coefs, double* maxabs, int mem, int device, void* stream);

/* Asynchronous device-memory variant: enqueue on `stream` and return.  maxabs_dev (device, `howmany`
 * doubles) and status_dev (device, one int64: flat index of the first non-finite real of `vals`, or
 * INT64_MAX) may each be NULL; the caller reads them after synchronising. */
int fastcheb_fit_async(int ndim, const int64_t* dims, int dtype, int is_complex, int ncomp, int64_t howmany,
                       const void* vals_dev, void* coefs_dev, double* maxabs_dev, int64_t* status_dev, int device,
                       void* stream);

/* droptol(coefs, tol) shape decision (interp.jl:77-93) for one coefficient array: writes the trimmed
 * sizes to newdims (== dims when nothing is dropped, including tol == 0).  The caller copies the
 * leading block (interp.jl:92). */
int fastcheb_droptol_shape(int ndim, const int64_t* dims, int dtype, int is_complex, int ncomp,
                           const void* coefs, double tol, int64_t* newdims, int mem, int device,
                           void* stream);

/* chebinterp(vals, lb, ub; tol) (interp.jl:95-99,140-141) end to end on the device: fit, droptol,
 * and creation of the handle holding the trimmed coefficients.  tol < 0 selects the reference's
 * default eps(real type of vals) (interp.jl:102). */
int fastcheb_interp(fastcheb_poly** out, int ndim, const int64_t* dims, int dtype, int is_complex,
                    int ncomp, const void* vals, int mem, const double* lb, const double* ub, double tol,
                    int device, void* stream);

/* chebinterp(c, order, lb, ub; tol) where the interpolated function is itself a device-resident ChebPoly `p`
 * (interp.jl:163-170 applied to a ChebPoly, the loop body of roots, roots.jl:105-107): Chebyshev points of the new
 * box [lb,ub] (inside p's box; interp.jl:4-36) -> batched evaluation of p -> N-d fit -> droptol -> new handle, all
 * on p's device.  order[d] >= 0 per dimension; tol < 0 selects eps of the real type. */
int fastcheb_reinterp(fastcheb_poly** out, const fastcheb_poly* p, const int64_t* order, const double* lb,
                      const double* ub, double tol, void* stream);

/* chebvandermonde(x, lb, ub, order) (regression.jl:5-46): the m x P Chebyshev-Vandermonde matrix of least-squares
 * regression, P = prod dims, A[j + lda*i] = prod_d T_{k_d}(z_d(x_j)) with i the column-major multi-index (k_1
 * fastest) — Julia's Array{T,2} layout.  pts: m x ndim reals AoS.  One pass, O(m P) (the reference's generic routine
 * is O(m P^2)).  Synchronous; a point whose normalised coordinate leaves [-1,1] (or is NaN) -> FASTCHEB_EDOMAIN +
 * fastcheb_error_index (regression.jl:33).  The least-squares solve itself (`A \ Y`, regression.jl:67) is a dense
 * QR and stays a library call on the caller's side, as it is LAPACK in the reference. */
int fastcheb_vandermonde(int ndim, const int64_t* dims, int dtype, const void* pts, int64_t m, const double* lb,
                         const double* ub, void* A, int64_t lda, int mem, int device, void* stream);

/* ---- multi-GPU from ONE host process --------------------------------------------------------------
 * What `c.(ys)` on a multi-GPU node needs (the reference's user broadcasts in one process: eval.jl:63-70,
 * FastChebInterp.jl:37): a fastcheb_multi replicates the polynomial on `ngpus` devices of the node
 * (devices == NULL: 0..ngpus-1; coefficients uploaded from the host pointer once per device, or — coefs_mem ==
 * FASTCHEB_MEM_DEVICE, coefficients on devices[0] — copied peer to peer over NVLink) and every batch is
 * block-partitioned over them: one worker thread and one stream per device, results written straight into the caller's
 * single array.  mem == FASTCHEB_MEM_HOST: host arrays (each device stages its block).  mem == FASTCHEB_MEM_DEVICE:
 * points and results live on devices[0]; the other devices' kernels load their block of points from, and store their
 * results into, devices[0]'s HBM through peer access (fused evaluation + gather over NVLink/NVSwitch; staged peer
 * copies when peer access is unavailable).  Status and errors are those of the single-GPU calls; EDOMAIN reports the
 * first offending GLOBAL point index (the minimum over the devices).  Calls on one handle are serialised. */
typedef struct fastcheb_multi fastcheb_multi;
int fastcheb_multi_create(fastcheb_multi** out, int ngpus, const int* devices, int ndim, const int64_t* dims, int dtype,
                          int is_complex, int ncomp, const void* coefs, int coefs_mem, const double* lb, const double* ub);
int fastcheb_multi_destroy(fastcheb_multi* m); /* NULL is a no-op */
int fastcheb_multi_ngpus(const fastcheb_multi* m);
int fastcheb_multi_device(const fastcheb_multi* m, int i);               /* CUDA device of replica i */
const fastcheb_poly* fastcheb_multi_poly(const fastcheb_multi* m, int i); /* replica i (queries, fastcheb_set_algo) */
/* c.(ys), chebjacobian.(Ref(c), ys), chebgradient.(Ref(c), ys) and the tuple layout — as the single-GPU entry points */
int fastcheb_multi_eval(fastcheb_multi* m, const void* pts, int64_t npts, void* out, int mem);
int fastcheb_multi_eval_jac(fastcheb_multi* m, const void* pts, int64_t npts, void* out_v, void* out_J, int mem);
int fastcheb_multi_eval_grad(fastcheb_multi* m, const void* pts, int64_t npts, void* out_v, void* out_g, int mem);
int fastcheb_multi_eval_jac_tuple(fastcheb_multi* m, const void* pts, int64_t npts, void* out_vJ, int mem);
/* fastcheb_fit for a batch of `howmany` independent fits, block-partitioned by fit over the devices (a single fit is
 * never split: interp.jl:39-57 per array).  MEM_DEVICE: vals / coefs / maxabs live on devices[0]. */
int fastcheb_multi_fit(int ngpus, const int* devices, int ndim, const int64_t* dims, int dtype, int is_complex, int ncomp,
                       int64_t howmany, const void* vals, void* coefs, double* maxabs, int mem);

/* ---- multi-GPU, one process per GPU: fused evaluation + gather over NVLink ---------------------------
 * One process per GPU.  The reference has no counterpart (it has no parallelism at all); these
 * entry points exist so that `c.(ys)` sharded over the GPUs of a node returns ONE result array: the gathering
 * rank allocates the buffer (peer_alloc) and ships the 64-byte IPC handle; the other ranks map it (peer_open) and
 * pass `mapped + slice offset` as the output pointer of fastcheb_eval_async / fastcheb_eval_jac_async — the
 * kernel's result stores go over NVLink/NVSwitch directly into the gatherer's HBM, no separate collective. */
#define FASTCHEB_IPC_HANDLE_BYTES 64
int fastcheb_peer_alloc(void** dev_ptr, size_t bytes, void* ipc_handle_out, int device);
int fastcheb_peer_free(void* dev_ptr, int device);
int fastcheb_peer_open(void** mapped_ptr, const void* ipc_handle, int device);
int fastcheb_peer_close(void* mapped_ptr, int device);
/* synchronous device -> host copy of (part of) a peer buffer */
int fastcheb_peer_read(void* host_dst, const void* dev_src, size_t bytes, int device);

#ifdef __cplusplus
}
#endif
#endif /* FASTCHEB_H */
