/* volinterp_b200.h — C ABI of libvolinterp_b200.so
 *
 * B200-native (sm_100a CUDA) implementation of the two data-parallel hot paths of
 * amisr/volumetricinterp: the per-record regularised least-squares fit and the
 * Estimate evaluation.  The reference is pure Python and has NO FFI today; this
 * header is the boundary a maintainer would bind with ctypes (INTEGRATION.md
 * shows the stubs).  Each entry point names the reference code it replaces
 * (file:line under /root/reference/volumetricinterp).
 *
 * Conventions
 *   - every function returns 0 on success or a negative VI_E* code; the message
 *     is available from vi_last_error() (thread local).  Nothing throws.
 *   - pointers are DEVICE pointers to caller-owned buffers unless the function
 *     name ends in _host (then they are host pointers and the call performs the
 *     host<->device copies itself, synchronously).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls
 *     are asynchronous with respect to the host except vi_fit_batched (which
 *     synchronises internally between search rounds) and the _host variants.
 *   - all arithmetic is IEEE binary64; matrices are row-major.
 *   - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef VOLINTERP_B200_H
#define VOLINTERP_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VI_MAXL_MAX 16
#define VI_MAXK_MAX 16
#define VI_NALPHA 102 /* alpha = 0,-1,...,-101: decade walk of interpolate.py:185-207 */
#define VI_NMAX_SMEM 160 /* largest nbasis whose per-record system fits one SM's shared memory */

/* error codes */
#define VI_OK 0
#define VI_EINVAL (-1)    /* bad argument */
#define VI_ECUDA (-2)     /* CUDA runtime error (message has the cudaError string) */
#define VI_EWORKSPACE (-3)/* workspace too small */
#define VI_EUNSUPPORTED (-4)

/* per-record status written by vi_fit_batched (NaN-record convention of
 * interpolate.py:142-145 and :558-563) */
#define VI_ST_OK 0          /* chi^2 = nu root bracketed and refined (brentq) */
#define VI_ST_TOO_SMOOTH 1  /* chi2(lambda=1) < nu: lambda = 0 (interpolate.py:188-191) */
#define VI_ST_NO_ROOT 2     /* no sign change down to 1e-101 -> NaN record (interpolate.py:210-211) */
#define VI_ST_NONFINITE 3   /* NaN/inf in the system: lstsq(check_finite) raises -> NaN record */
#define VI_ST_NOCONV 4      /* iteration limit in brentq / eigen-solver -> NaN record */
#define VI_ST_EMPTY 5       /* record without a single valid gate (reference crashes; NaN record here) */

/* regularisation-parameter method (REGULARIZATION_METHOD) */
#define VI_METHOD_NONE 0    /* empty REGULARIZATION_LIST: plain lstsq per record */
#define VI_METHOD_CHI2 1    /* interpolate.py:152-218 */
#define VI_METHOD_GCV 2     /* interpolate.py:263-351 (leave-one-gate-out residual sum, Nelder-Mead) */

/* normal-equation modes */
#define VI_NE_STRICT 0      /* reference summation order, bit-identical to np.einsum (interpolate.py:456,458) */
#define VI_NE_FAST 1        /* tiled FP64 contraction, any order */

/* Host-precomputed description of one sphharmlag model (models/sphharmlag.py:57-75). */
typedef struct vi_shl_params {
  int32_t maxk, maxl;
  double ct0, st0;                        /* cos/sin of centre colatitude (sphharmlag.py:346) */
  double kx, ky;                          /* rotation axis (sphharmlag.py:349) */
  double nu[VI_MAXL_MAX];                 /* degree nu(l) (sphharmlag.py:114) */
  double kvm[VI_MAXL_MAX][VI_MAXL_MAX];   /* K_{nu(l),|m|} (sphharmlag.py:318-320) */
  double g1[VI_MAXL_MAX][VI_MAXL_MAX];    /* Gamma(nu-|m|+1) */
  double g2[VI_MAXL_MAX][VI_MAXL_MAX];    /* Gamma(nu+|m|+1) (may be +inf) */
} vi_shl_params;

const char* vi_version(void);
const char* vi_last_error(void);

/* models/sphharmlag.py:118-145 (basis) + :324-359 (transform_coord).
 * A: npts x N row-major (may be NULL), At: N x npts (may be NULL). params: HOST pointer. */
int vi_basis_sphharmlag(const double* lat, const double* lon, const double* alt, int64_t npts,
                        const vi_shl_params* params, double* A, double* At, void* stream);

/* models/sphharmlag.py:148-184 (grad_basis): gradient of every basis function along (z-hat, theta-hat,
 * phi-hat) of the model coordinates.  out: npts x 3 x N row-major, the shape the reference returns
 * (np.array(Ag).T).  Not called on the reference's fit / Estimate path. */
int vi_grad_basis_sphharmlag(const double* lat, const double* lon, const double* alt, int64_t npts,
                             const vi_shl_params* params, double* out, void* stream);

/* models/radbasfun.py:83-112 (+ transform_coords :232-256). centers: N x 3 ECEF metres (device). */
int vi_basis_radbasfun(const double* lat, const double* lon, const double* alt, int64_t npts,
                       const double* centers, int32_t N, double eps, double* A, double* At, void* stream);

/* interpolate.py:516-524 (gate mask, W = error**-2, b) and :456-458 (A^T W A, A^T W b) for R
 * records at once.  value/error: R x P with NaN = invalid gate.  weight (optional, may be NULL):
 * R x P caller-computed error**-2; if NULL the kernel uses the correctly rounded 1/error^2.
 * Outputs: G R x N x N, y R x N, sWbb R (sum W b^2), npts R (valid gates), and optionally the
 * masked weights / data Wm, bm (R x P, zero at invalid gates; may be NULL). */
int vi_normal_eq_batched(const double* A, const double* value, const double* error, const double* weight,
                         int32_t R, int32_t P, int32_t N, int32_t mode,
                         double* G, double* y, double* sWbb, int32_t* npts, double* Wm, double* bm, void* stream);

/* Scratch size needed by vi_fit_batched / vi_solve_batched for the given shape. `systems` is the
 * number of simultaneous eigen-systems the caller is willing to hold (0 = library default: as many
 * as 32 GiB hold, rounded down to whole waves of the QL kernel, at most what the batch needs;
 * about 1 MB per system at N = 144). */
int vi_fit_workspace_bytes(int32_t R, int32_t P, int32_t N, int32_t nreg, int64_t systems, int64_t* bytes);

/* interpolate.py:462 for S independent systems: C_s = lstsq(sym(G[rec_s]) + sum_i lam[s][i] Reg_i, y[rec_s])
 * (minimum-norm solution over |eigenvalue| > rcond * max|eigenvalue|).  rec: S int32 record index
 * (NULL = identity).  lam: S x nreg.  Outputs C S x N, rank S, status S. */
int vi_solve_batched(const double* G, const double* y, const int32_t* rec, const double* regmats,
                     const double* lam, int64_t S, int32_t N, int32_t nreg, double rcond,
                     double* C, int32_t* rank, int32_t* status,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* interpolate.py:462-467: vi_solve_batched plus the covariance dC_s = H X0 H, H = pinv(X_s) (cut-off
 * N eps max|eigenvalue| = scipy.linalg.pinv's rule), X0 = G[rec_s] unregularised.  dC: S x N x N device (NULL: as
 * vi_solve_batched).  The workspace must hold the covariance scratch too (vi_fit_workspace_bytes with R = S). */
int vi_solve_cov_batched(const double* G, const double* y, const int32_t* rec, const double* regmats,
                         const double* lam, int64_t S, int32_t N, int32_t nreg, double rcond,
                         double* C, double* dC, int32_t* rank, int32_t* status,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* interpolate.py:555-569 for R records: find_reg_param (:97-147, chi2 :152-218, chi2objfunct
 * :220-261), NaN-record rule (:558-563), final eval_C (:566) and chi^2 (:569).
 * At: N x P (transposed design matrix); A: P x N (row-major; needed by VI_METHOD_GCV only, else may be
 * NULL); Wm/bm: masked weights/data from vi_normal_eq_batched.
 * regmats: nreg x N x N.  Outputs: C R x N, dC R x N x N (may be NULL), chi2 R, lam R x nreg,
 * rank R, status R, nsolve (optional, 1 int64: number of eigen-systems solved).
 * dC may also be a PINNED HOST pointer (cudaHostAlloc / cudaHostRegister): the covariance is then produced in chunks
 * of 512 records into a device ring and copied out on a side stream while the next chunk is computed (the R x N x N
 * block, 1.66 GB at R = 10 k and N = 144, never sits in HBM); `stream` is made to wait for the last copy. */
int vi_fit_batched(const double* At, const double* A, const double* Wm, const double* bm,
                   const double* G, const double* y, const int32_t* npts,
                   int32_t R, int32_t P, int32_t N,
                   const double* regmats, int32_t nreg, int32_t method,
                   double* C, double* dC, double* chi2, double* lam, int32_t* rank, int32_t* status,
                   int64_t* nsolve, void* workspace, int64_t workspace_bytes, void* stream);

/* Diagnostics of the last VI_METHOD_CHI2 search that ran in `workspace` (same R, P, nreg as that call):
 * table U x VI_NALPHA = chi2(10^-k) as the decade walk of interpolate.py:180-207 saw it (entries the walk
 * never read are NaN), nu U = len(b) * scale factor of the bracket (interpolate.py:175,181), k_lo U = decade
 * of the bracket's lower end (alpha = -k_lo; -1 when none), kdone U = distinct table entries evaluated.
 * U = R * nreg; any output may be NULL.  Device pointers. */
int vi_fit_search_trace(const void* workspace, int64_t workspace_bytes, int32_t R, int32_t P, int32_t nreg,
                        double* table, double* nu, int32_t* k_lo, int32_t* kdone, void* stream);

/* estimate.py:113-121: out[r][p] = sum_n basis(p)[n] * C[r][n], NaN outside the hull.
 * C: Rsel x N (device).  hull_eq: F x 4 facet equations [n|d] of ConvexHull(hull_vert) (inside iff
 * n.x + d <= 0 for all facets; equivalent of estimate.py:153-178), NULL/F=0 = no check. */
int vi_estimate_sphharmlag(const double* lat, const double* lon, const double* alt, int64_t npts,
                           const vi_shl_params* params, const double* C, int32_t Rsel,
                           const double* hull_eq, int32_t F, double* out, void* stream);
int vi_estimate_radbasfun(const double* lat, const double* lon, const double* alt, int64_t npts,
                          const double* centers, int32_t N, double eps, const double* C, int32_t Rsel,
                          const double* hull_eq, int32_t F, double* out, void* stream);

/* estimate.py:113-121 for MANY records of one coefficient file (BASELINE configs[3]: 1 k records): the points are
 * tested against the hull and compacted first, the basis rows of the in-hull points are evaluated once (one thread
 * per point) and out = rows . C^T runs as an FP64 tensor-core GEMM; points outside the hull cost one NaN store per
 * record and nothing else.  Same outputs as vi_estimate_* (which stay the entry points for a few records).
 * workspace: vi_estimate_workspace_bytes(npts, N, Rsel) bytes of device scratch; npts < 2^31 per call. */
int vi_estimate_workspace_bytes(int64_t npts, int32_t N, int32_t Rsel, int64_t* bytes);
int vi_estimate_sphharmlag_many(const double* lat, const double* lon, const double* alt, int64_t npts,
                                const vi_shl_params* params, const double* C, int32_t Rsel,
                                const double* hull_eq, int32_t F, double* out,
                                void* workspace, int64_t workspace_bytes, void* stream);
int vi_estimate_radbasfun_many(const double* lat, const double* lon, const double* alt, int64_t npts,
                               const double* centers, int32_t N, double eps, const double* C, int32_t Rsel,
                               const double* hull_eq, int32_t F, double* out,
                               void* workspace, int64_t workspace_bytes, void* stream);

/* End-to-end convenience entry points on HOST buffers (the calls the Python `Interpolate` /
 * `Estimate` classes make when handed numpy arrays): H2D, kernels, D2H, synchronous. */
int vi_fit_host(const double* A /*P x N*/, const double* value, const double* error, const double* weight,
                int32_t R, int32_t P, int32_t N, const double* regmats, int32_t nreg, int32_t method,
                int32_t ne_mode, double* C, double* dC, double* chi2, double* lam, int32_t* rank,
                int32_t* status);
int vi_estimate_sphharmlag_host(const double* lat, const double* lon, const double* alt, int64_t npts,
                                const vi_shl_params* params, const double* C, int32_t Rsel,
                                const double* hull_eq, int32_t F, double* out);

/* FP64 peak probes used by bench.py for the roofline denominators: mode 0 = dependent-free DFMA
 * chains, mode 1 = mma.sync m8n8k4 f64 (DMMA).  Writes achieved TFLOP/s to *tflops (host). */
int vi_fp64_peak_probe(int32_t mode, int32_t iters, double* tflops, void* stream);

/* Launch accounting (diagnostics; not thread safe).  Kinds, in order: basis, normal_eq, tridiag (band
 * reduction / one-stage Householder), tql, apply, chi2, covariance, estimate, misc, chase (band -> tridiagonal).  vi_profile_read returns, per kind, the device milliseconds
 * (CUDA events on the launching stream; only while enabled) and the number of kernel launches (always
 * counted) since the last vi_profile_reset. */
#define VI_PROFILE_KINDS 10
int vi_profile_enable(int32_t on);
int vi_profile_reset(void);
int vi_profile_read(double* ms_by_kind, int64_t* launches_by_kind, int32_t nkinds);
/* Givens rotations generated by the tridiagonal eigen-solves (= entries of the rotation tapes written; each is read
 * twice by the replay) and eigen-systems solved by vi_fit_batched since the last vi_profile_reset. */
int vi_profile_counters(int64_t* rotations, int64_t* systems);

#ifdef __cplusplus
}
#endif
#endif
