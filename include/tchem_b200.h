/*
 * tchem_b200.h - C ABI of the B200-native internal-coordinate hot path.
 *
 * This is the only boundary between host code (the tchem:: C++ classes in include/tchem/,
 * the Python binding in torch_chemistry_b200/) and the CUDA kernels. Plain pointers and
 * sizes, no C++ or torch types. Every entry point names the reference interface it replaces
 * (paths relative to the Torch-Chemistry tree).
 *
 * Conventions
 *  - All floating-point data is IEEE float64. "Device" pointers are device memory on the
 *    table's device; "host" entry points (suffix _host) take host pointers and perform the
 *    host<->device copies themselves (pinned staging, chunked and overlapped with compute).
 *  - Batched superset of the reference: the reference takes ONE geometry r[cartdim] per call
 *    (source/IC/IntCoordSet.cpp:140-141); here every tensor has a leading batch axis of
 *    length B and is C-contiguous: r[B][cartdim], q[B][intdim], J[B][intdim][cartdim],
 *    K[B][intdim][cartdim][cartdim], g_q[B][intdim], H_q[B][intdim][intdim], ...
 *  - Every function returns 0 on success or a TCHEM_E* code; tchem_last_error() returns a
 *    thread-local message. Kernels are launched on the caller's stream and are not
 *    synchronised. There is no CPU fallback: without a CUDA device every compute call fails
 *    with TCHEM_ECUDA.
 */
#ifndef TCHEM_B200_H
#define TCHEM_B200_H

#include <stdint.h>

#define TCHEM_API __attribute__((visibility("default")))

#ifdef __cplusplus
extern "C" {
#endif

#define TCHEM_OK        0
#define TCHEM_EINVAL    1   /* -> std::invalid_argument in the C++ classes */
#define TCHEM_EDOMAIN   2   /* -> std::domain_error (unimplemented type)     */
#define TCHEM_EFILE     3   /* -> CL::utility::file_error                    */
#define TCHEM_ECUDA     4   /* -> std::runtime_error                         */
#define TCHEM_ENOMEM    5

/* tchem::IC::InvDisp_type, include/tchem/IC/InvDisp.hpp:9-44 (same order, same values) */
enum tchem_invdisp_type {
    TCHEM_DUMMY = 0, TCHEM_STRETCHING, TCHEM_BENDING, TCHEM_COSBENDING,
    TCHEM_TORSION, TCHEM_SINTORSION, TCHEM_COSTORSION,
    TCHEM_TORSION2, TCHEM_SINTORSION2, TCHEM_COSTORSION2, TCHEM_PXTORSION2, TCHEM_PYTORSION2,
    TCHEM_OUTOFPLANE, TCHEM_SINOOP, TCHEM_NTYPES
};

/* One (coefficient, InvDisp) pair of one IntCoord:
 * include/tchem/IC/InvDisp.hpp:54-63 + include/tchem/IC/IntCoord.hpp:11-12, flattened. */
typedef struct tchem_invdisp {
    int32_t ic;        /* index of the internal coordinate (row of q / J / K)       */
    int32_t type;      /* enum tchem_invdisp_type                                    */
    int32_t natoms;    /* 0, 2, 3, 4 or 5                                            */
    int32_t atoms[5];  /* 0-based atom indices, unused entries ignored               */
    double  coeff;     /* linear-combination coefficient, already normalised         */
    double  min;       /* torsion / torsion2 range start, default -pi                */
} tchem_invdisp_t;

typedef struct tchem_ic_table  tchem_ic_table;   /* opaque: compiled IntCoordSet      */
typedef struct tchem_sap_table tchem_sap_table;  /* opaque: compiled SAPSet           */
typedef void * tchem_stream_t;                   /* cudaStream_t                      */

TCHEM_API const char * tchem_last_error(void);
/* number of CUDA devices visible (0 when there is no driver / device) */
TCHEM_API int tchem_device_count(void);

/* ---- definition files -------------------------------------------------------------- */
/* IntCoordSet::IntCoordSet(format, file), source/IC/IntCoordSet.cpp:11-120 (both formats,
 * including IntCoord::normalize, source/IC/IntCoord.cpp:26-31). Pure host code.
 * Writes up to `capacity` terms to `terms`, always stores the required count in *n_terms
 * and the number of internal coordinates in *n_ic. */
TCHEM_API int tchem_ic_parse_file(const char * format, const char * file,
                        tchem_invdisp_t * terms, int capacity, int * n_terms, int * n_ic);
TCHEM_API int tchem_ic_parse_text(const char * format, const char * text,
                        tchem_invdisp_t * terms, int capacity, int * n_terms, int * n_ic);

/* SAPSet::SAPSet(sapoly_file, irreducible, dimensions), source/polynomial/SAPSet.cpp:89-109 and
 * SAP::SAP(strs), source/polynomial/SAP.cpp:42-76. Flattened as CSR: monomial i owns
 * coords[term_ptr[i] .. term_ptr[i+1]) as (irreducible, index) pairs, 0-based, sorted
 * descending like SAP::coords_. term_ptr needs capacity_terms + 1 entries. */
TCHEM_API int tchem_sap_parse_file(const char * file, int32_t * term_ptr, int capacity_terms,
                         int32_t * coords /* [capacity_coords][2] */, int capacity_coords,
                         int * n_terms, int * n_coords);
TCHEM_API int tchem_sap_parse_text(const char * text, int32_t * term_ptr, int capacity_terms,
                         int32_t * coords, int capacity_coords, int * n_terms, int * n_coords);

/* ---- compiled tables ---------------------------------------------------------------- */
/* Flattens an IntCoordSet (std::vector<IntCoord>, include/tchem/IC/IntCoordSet.hpp:11-12)
 * into the device-resident program the kernels interpret. n_atoms fixes cartdim = 3*n_atoms;
 * pass 0 to take 1 + the largest atom index used. */
TCHEM_API int tchem_ic_table_create(const tchem_invdisp_t * terms, int n_terms, int n_ic, int n_atoms,
                          int device, tchem_ic_table ** out);
TCHEM_API void tchem_ic_table_destroy(tchem_ic_table * t);
TCHEM_API int tchem_ic_table_intdim(const tchem_ic_table * t);
TCHEM_API int tchem_ic_table_cartdim(const tchem_ic_table * t);
TCHEM_API int tchem_ic_table_device(const tchem_ic_table * t);

/* SAPSet (include/tchem/polynomial/SAPSet.hpp:9-31): irreducible of the set, dimensions per
 * irreducible, monomials in CSR form as produced by tchem_sap_parse_*. */
TCHEM_API int tchem_sap_table_create(const int32_t * term_ptr, const int32_t * coords, int n_terms,
                           int irreducible, const int32_t * dims, int n_irreducibles,
                           int device, tchem_sap_table ** out);
TCHEM_API void tchem_sap_table_destroy(tchem_sap_table * t);
TCHEM_API int tchem_sap_table_nterms(const tchem_sap_table * t);
TCHEM_API int tchem_sap_table_sumdim(const tchem_sap_table * t);

/* ---- q / J / K ------------------------------------------------------------------------ */
/* Replaces, over a batch:
 *   value_path != 0 (J == K == NULL): IntCoordSet::operator(),   source/IC/IntCoordSet.cpp:139-145
 *                                     (the clamped formulas of InvDisp::operator(), InvDisp.cpp:64-254)
 *   K == NULL:                         IntCoordSet::compute_IC_J,  source/IC/IntCoordSet.cpp:147-159
 *   K != NULL:                         IntCoordSet::compute_IC_J_K, source/IC/IntCoordSet.cpp:161-175
 * q may be NULL when only J (and K) are wanted. */
TCHEM_API int tchem_ic_eval(const tchem_ic_table * t, const double * r, int64_t B,
                  double * q, double * J, double * K, int value_path, tchem_stream_t stream);
/* same with host buffers; copies and kernels are pipelined in chunks on internal streams and
 * the call returns after the results have landed in the host buffers. */
TCHEM_API int tchem_ic_eval_host(const tchem_ic_table * t, const double * r, int64_t B,
                       double * q, double * J, double * K, int value_path);

/* ---- gradient transforms -------------------------------------------------------------- */
/* IntCoordSet::gradient_int2cart, source/IC/IntCoordSet.cpp:207-218:  g_r = J^T g_q */
TCHEM_API int tchem_ic_grad_int2cart(const tchem_ic_table * t, const double * r, const double * intgrad,
                           int64_t B, double * cartgrad, tchem_stream_t stream);
/* IntCoordSet::gradient_cart2int, :190-205:  g_q = (J J^T)^-1 J g_r (Cholesky + explicit inverse) */
TCHEM_API int tchem_ic_grad_cart2int(const tchem_ic_table * t, const double * r, const double * cartgrad,
                           int64_t B, double * intgrad, tchem_stream_t stream);
/* IntCoordSet::gradient_cart2int_matrix, :178-188:  M[B][intdim][cartdim] = (J J^T)^-1 J */
TCHEM_API int tchem_ic_grad_cart2int_matrix(const tchem_ic_table * t, const double * r, int64_t B,
                                  double * M, tchem_stream_t stream);

/* ---- Hessian transforms --------------------------------------------------------------- */
/* IntCoordSet::Hessian_int2cart, :248-266:  H_r = J^T H_q J + sum_k g_q[k] K[k] */
TCHEM_API int tchem_ic_hess_int2cart(const tchem_ic_table * t, const double * r, const double * intgrad,
                           const double * intHess, int64_t B, double * cartHess,
                           tchem_stream_t stream);
/* IntCoordSet::Hessian_cart2int, :221-246:
 *   A^T = (J J^T)^-1 J,  H_q = A^T H_r A - sum_k (A^T g_r)[k] A^T K[k] A */
TCHEM_API int tchem_ic_hess_cart2int(const tchem_ic_table * t, const double * r, const double * cartgrad,
                           const double * cartHess, int64_t B, double * intHess,
                           tchem_stream_t stream);

/* The cart2int entry points factor J J^T per geometry. Returns 1 (and clears the flag) when any
 * factorisation since the last call met a non-positive pivot - redundant or singular internal
 * coordinates, where the reference's at::cholesky throws - else 0. Synchronises the device. */
TCHEM_API int tchem_ic_factor_status(const tchem_ic_table * t);

/* ---- SAPSet ---------------------------------------------------------------------------- */
/* x[B][sumdim] holds the per-irreducible vectors xs[0], xs[1], ... concatenated.
 * y[B][nterms]          = SAPSet::operator()(xs),      source/polynomial/SAPSet.cpp:134-144
 * J[B][nterms][sumdim]  = SAPSet::Jacobian_(xs, J),    source/polynomial/SAPSet.cpp:178-201
 * Either output may be NULL. */
TCHEM_API int tchem_sap_eval(const tchem_sap_table * t, const double * x, int64_t B,
                   double * y, double * J, tchem_stream_t stream);
/* Second-order Jacobian of the set, concatenated and symmetric: K[B][nterms][sumdim][sumdim]. Replaces,
 * batched, SAPSet::Jacobian2nd_(xs, K) (source/polynomial/SAPSet.cpp:243-276; per monomial
 * SAP::Hessian_, source/polynomial/SAP.cpp:226-263). x[B][sumdim] and K are device pointers. */
TCHEM_API int tchem_sap_jacobian2nd(const tchem_sap_table * t, const double * x, int64_t B, double * K,
                                    tchem_stream_t stream);
/* chained path: r -> q (compute_IC_J formulas) -> xs = q split by `dims` -> SAPSet value and
 * Jacobian, one kernel, q never leaves the chip unless q_out != NULL. Requires
 * intdim == sumdim. Callers of the reference do this by hand (test/normal_mode/main.cpp:56-61). */
TCHEM_API int tchem_ic_sap_eval(const tchem_ic_table * ic, const tchem_sap_table * sap, const double * r,
                      int64_t B, double * q_out, double * y, double * J, tchem_stream_t stream);
TCHEM_API int tchem_ic_sap_eval_host(const tchem_ic_table * ic, const tchem_sap_table * sap, const double * r,
                           int64_t B, double * q_out, double * y, double * J);
/* chained path, CARTESIAN derivatives:
 *   dydr  [B][nterms][cartdim]            = (dSAP/dq) J
 *   d2ydr2[B][nterms][cartdim][cartdim]   = J^T (d2SAP/dq2) J + sum_k (dSAP/dq_k) K[k]       (may be NULL)
 * what callers of the reference compose by hand from SAPSet::Jacobian_ / Jacobian2nd_ (source/polynomial/
 * SAPSet.cpp:178-201, :243-276) and IntCoordSet::compute_IC_J / compute_IC_J_K (source/IC/IntCoordSet.cpp:147-175),
 * cf. test/normal_mode/main.cpp:56-61. y[B][nterms] may be NULL. Device pointers. */
TCHEM_API int tchem_ic_sap_eval_cart(const tchem_ic_table * ic, const tchem_sap_table * sap, const double * r, int64_t B,
                                     double * y, double * dydr, double * d2ydr2, tchem_stream_t stream);

/* ---- introspection (host only, no device needed; used by the CPU tests) ------------------------ */
/* Compiles the definition for `mode` (0 value path, 1 q+J, 2 q+J+K) exactly as
 * tchem_ic_table_create does and reports the sizes of the resulting program:
 * counts[0..8] (room for 16) = primitives, chunks, q-map words, q sources, element-map entries,
 * packed element sources (uint32 words, 4 per quad), compact capacity, zero runs, distinct
 * coefficients. With non-NULL pointers the arrays themselves are copied out (layouts:
 * torch_chemistry_b200/csrc/ic_program.h: PrimDesc, ChunkDesc, uint32, Src, ElemMap, uint32,
 * ZeroRun, double). */
TCHEM_API int tchem_ic_program_export(const tchem_invdisp_t * terms, int n_terms, int n_ic, int cartdim, int mode,
                                      int64_t * counts, void * prims, void * chunks, uint32_t * map, void * qsrcs,
                                      void * emap, void * srcs, void * runs, double * coefs);

/* Large batches of q + J on tables without 5-atom terms are served by a kernel SPECIALISED to the
 * table: the definition is emitted as straight-line CUDA C++ over the same device functions and
 * compiled for sm_100a with NVRTC on first use (environment: TCHEM_JIT=0 disables it,
 * TCHEM_JIT_MIN_BATCH sets the batch threshold). These two entry points expose the generator for
 * inspection and for the CPU-side tests: the source for a definition (returns the buffer size
 * needed, 0 when the table is not eligible) and an NVRTC compile check of a source (no device
 * needed; 0 = compiled, the compiler / ptxas log is copied to `log`). */
TCHEM_API int64_t tchem_ic_jit_source(const tchem_invdisp_t * terms, int n_terms, int n_ic, int cartdim, int dcap,
                                      char * buf, int64_t capacity);
/* the variant whose write-out goes through the TMA unit (cp.async.bulk stores from geometry-major
 * staging rows; TCHEM_JIT_BULK=0 disables it); 0 when the shape does not give 16-byte aligned pieces */
TCHEM_API int64_t tchem_ic_jit_source_bulk(const tchem_invdisp_t * terms, int n_terms, int n_ic, int cartdim, int dcap,
                                           char * buf, int64_t capacity);
/* the group variant: G warps of a block share one tile and ONE whole-geometry TMA box (contiguous tile stores) */
TCHEM_API int64_t tchem_ic_jit_source_group(const tchem_invdisp_t * terms, int n_terms, int n_ic, int cartdim, int G,
                                            char * buf, int64_t capacity);
/* source of the kernel specialised to a SAPSet for tchem_sap_jacobian2nd (large batches; whole-geometry
 * staging, one 1-D TMA bulk store per tile); definition arrays as for tchem_sap_table_create */
TCHEM_API int64_t tchem_sap_jit_hess_source(const int32_t * term_ptr, const int32_t * coords, int n_terms,
                                            const int32_t * dims, int n_irreducibles, char * buf, int64_t capacity);
/* source of the kernel specialised to a SAPSet for tchem_sap_eval (bulk != 0: whole-tile TMA bulk stores) */
TCHEM_API int64_t tchem_sap_jit_source(const int32_t * term_ptr, const int32_t * coords, int n_terms,
                                       const int32_t * dims, int n_irreducibles, int bulk, char * buf, int64_t capacity);
/* source of the kernel specialised to an (IntCoordSet, SAPSet) pair with the Jacobian in Cartesian coordinates
 * (tchem_ic_sap_eval_cart, first order); 0 when the pair is not eligible */
TCHEM_API int64_t tchem_ic_sap_jit_cart_source(const tchem_invdisp_t * terms, int n_terms, int n_ic, int cartdim,
                                               const int32_t * term_ptr, const int32_t * coords, int n_sap,
                                               const int32_t * dims, int n_irreducibles, char * buf, int64_t capacity);
/* the same pair kernel with whole-tile TMA bulk stores (tchem_jit_chain_cart_bulk); 0 when the pair is not eligible (odd row lengths) */
TCHEM_API int64_t tchem_ic_sap_jit_cart_bulk_source(const tchem_invdisp_t * terms, int n_terms, int n_ic, int cartdim,
                                                    const int32_t * term_ptr, const int32_t * coords, int n_sap,
                                                    const int32_t * dims, int n_irreducibles, char * buf, int64_t capacity);
TCHEM_API int tchem_ic_jit_compile_check(const char * source, char * log, int64_t capacity);
/* 1 when a q + J call on B geometries is served by the specialised kernel, 0 = interpreter */
TCHEM_API int tchem_ic_uses_specialised(const tchem_ic_table * t, int64_t B);
/* the same question for tchem_ic_sap_eval (pair-specialised chained kernel); ic == NULL: tchem_sap_eval */
TCHEM_API int tchem_ic_sap_uses_specialised(const tchem_sap_table * sap, const tchem_ic_table * ic, int64_t B);

/* NVRTC bookkeeping of this process: out4 = {kernels compiled, kernels taken from the on-disk cubin cache,
 * kernels that failed to build, microseconds spent inside NVRTC}. The cache lives in $TCHEM_JIT_CACHE
 * (empty or "0": disabled), else $XDG_CACHE_HOME/tchem_b200 or $HOME/.cache/tchem_b200, keyed by a hash of the
 * generated source, the embedded math header and the CUDA runtime version. A specialised kernel that cannot
 * be built is reported once on stderr and the interpreter kernel serves the table; TCHEM_JIT_STRICT=1 makes
 * the call fail with TCHEM_ECUDA instead. */
TCHEM_API int tchem_jit_stats(int64_t * out4);

/* ---- measurement helpers (bench.py only) ------------------------------------------------ */
/* number of kernels this library has launched in this process so far */
TCHEM_API int64_t tchem_launch_count(void);
/* "kernel name=launch count;" for every kernel this library has launched in this process (what was actually
 * launched, not what the environment asked for). Returns the buffer size needed, copies at most `capacity`. */
TCHEM_API int64_t tchem_kernel_log(char * buf, int64_t capacity);
/* average device time in ms (CUDA events on `stream`) of `iters` launches of tchem_ic_eval */
TCHEM_API int tchem_ic_eval_timed(const tchem_ic_table * t, const double * r, int64_t B, double * q,
                        double * J, double * K, int value_path, int iters,
                        tchem_stream_t stream, float * ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* TCHEM_B200_H */
