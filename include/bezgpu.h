/*
 * bezgpu.h -- C ABI of libbezgpu.so: the B200 (sm_100a) implementation of the
 * Bezier constraint-evaluation hot path of
 * caslabuiowa/OptimalBezierTrajectoryGeneration.
 *
 * Every entry point is what a binding for the reference would call in place of
 * the cited reference code (paths relative to the upstream repository root).
 * Conventions
 *   - all pointers named d_* are DEVICE pointers owned by the caller (PyTorch
 *     tensors in the Python host layer); h_* are HOST pointers.  The library
 *     allocates only plan-resident constant tables.
 *   - every launch is asynchronous on the caller's stream (a cudaStream_t
 *     passed as void*; NULL = legacy default stream).
 *   - return value: 0 = OK, > 0 = a cudaError_t, < 0 = argument error
 *     (BEZ_E*).  bez_last_error() returns a thread-local description.
 *   - all arithmetic is IEEE fp64.  Geometry entry points (split, extrema,
 *     GJK, minDist, collCheck) are compiled without FMA contraction so that
 *     they reproduce the reference's numba/numpy rounding.
 *
 * Control-point layout in HBM ("vehicle rows"):   cpts[b][v][S]
 *     b = evaluation point (base or finite-difference perturbed x)   0..B-1
 *     v = curve index (vehicles first, then point obstacles)         0..N-1
 *     S = dim*(n+1) rounded up to an even count (16-byte aligned rows); entry
 *         d*(n+1)+k is control point k of dimension d, the pad entry is 0.
 * A warp that walks over 32 consecutive pairs (i, j..j+31) reads one row
 * (broadcast) plus one contiguous 32*S*8-byte span with 128-bit loads.
 */
#ifndef BEZGPU_H
#define BEZGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BEZ_OK            0
#define BEZ_EINVAL       (-1)   /* bad argument                        */
#define BEZ_ENOMEM       (-2)   /* host allocation failed              */
#define BEZ_EUNSUPPORTED (-3)   /* degree / dimension out of range     */

#define BEZ_MAX_DEGREE    24    /* n   <= 24 for the constraint kernels */
#define BEZ_MAX_GEOM_PTS  32    /* n+1 <= 32 for split/GJK/minDist      */

typedef struct bez_plan bez_plan;

const char *bez_last_error(void);
int bez_version(void);

/* ---- plan: device-resident constant tables for one (n, dim, elev) ------
 * Replaces the class-level caches BezierParams.elevationMatrixCache /
 * productMatrixCache (bezier.py:48-52) and the table builders
 * elevMatrix (bezier.py:1127-1147), prodMatrix (bezier.py:1151-1176).
 * The caller builds the tables ON THE HOST with scipy.special.binom exactly
 * as the reference does and passes them in:
 *   h_prodW [(n+1)*(n+1)]   W[i][j] = C(n,i)C(n,j)/C(2n,i+j)
 *   h_elevT [(2n+1)*L]      T[j][i] = C(2n,j)C(E,i-j)/C(2n+E,i),  L = 2n+E+1
 *   h_elev1 [n*(n+1)]       elevMatrix(n-1,1) used by Bezier.diff (bezier.py:519)
 */
int bez_plan_create(int n, int dim, int elev, int device,
                    const double *h_prodW, const double *h_elevT,
                    const double *h_elev1, bez_plan **out);
int bez_plan_destroy(bez_plan *plan);
int bez_plan_info(const bez_plan *plan, int *n, int *dim, int *elev, int *L);

/* ---- A0: BezOptimization.reshapeVector (optimization.py:242-285) + the
 * point-obstacle stacking of temporalSeparationConstraints
 * (optimization.py:86-94), for B optimisation vectors at once.
 *   d_x        [B][nvar]
 *   fixed_ends 0/1 : initPoints/finalPoints given (columns 0 and n fixed)
 *   dubins     0/1 : initSpeeds given (columns 1 and n-1 from speed*tf/n*(cos,sin); dim must be 2)
 *   timeopt    0/1 : tf = x[nvar-1] else tf_fixed
 *   d_init/d_final [numVeh][dim]; d_ispeed/d_fspeed/d_icos/d_isin/d_fcos/d_fsin [numVeh]
 *   d_obst     [nObs][dim] (may be NULL when nObs == 0)
 *   d_cpts     [B][numVeh+nObs][S]          (output)
 *   d_tf       [B]                          (output; the tf each x implies)
 */
int bez_assemble_cpts(const bez_plan *plan, const double *d_x, int B, int nvar,
                      int numVeh, int nObs, int fixed_ends, int dubins, int timeopt,
                      double tf_fixed,
                      const double *d_init, const double *d_final,
                      const double *d_ispeed, const double *d_fspeed,
                      const double *d_icos, const double *d_isin,
                      const double *d_fcos, const double *d_fsin,
                      const double *d_obst,
                      double *d_cpts, double *d_tf, void *stream);

/* ---- A1-A4: _temporalSeparationConstraints (optimization.py:311-346) =
 * Bezier.sub (bezier.py:347-374) -> normSquare/_normSquare (bezier.py:869-889,
 * 1724-1756; incl. the dim/2 factor) -> elev(E) (bezier.py:469-495)
 * -> minus maxSep^2, fused, for the pairs [pair_begin, pair_begin+npairs) of
 * the lexicographic i<j list over N curves, for B evaluation points.
 *   d_out     [B][npairs][L]
 *   d_pairmin [B][npairs] or NULL: min over the L values of each pair (the
 *             quantity whose sign is the active-pair flag).
 */
int bez_pair_sepsq_elev(const bez_plan *plan, const double *d_cpts, int B, int N,
                        int64_t pair_begin, int64_t npairs, double maxSep2,
                        double *d_out, double *d_pairmin, void *stream);

/* ---- A5: _maxSpeedConstraints / _minSpeedConstraints (optimization.py:349-422)
 * = Bezier.diff (bezier.py:497-519, same degree) -> normSquare ->
 * elev(E) -> alpha*value + beta, for vehicles [veh_begin, veh_begin+nveh).
 *   max speed: alpha=-1, beta=maxSpeed^2 ; min speed: alpha=+1, beta=-minSpeed^2
 *   d_tf  [B]      d_out [B][nveh][L]
 */
int bez_speed_sq_elev(const bez_plan *plan, const double *d_cpts, const double *d_tf,
                      int B, int N, int veh_begin, int nveh,
                      double alpha, double beta, double *d_out, void *stream);

/* ---- A6: _maxAngularRateConstraints -> _angularRateSqr (optimization.py:425-459,
 * 578-611) with Bezier.mul / multiplyBezCurves (bezier.py:376-432, 1211-1246):
 * 2-D vehicles only (the reference raises ValueError otherwise, optimization.py:590).
 * Tables (host, scipy.special.binom, m = n + elev):
 *   h_Tpos   [(n+1)*(m+1)]  elevMatrix(n, elev)
 *   h_elev1m [m*(m+1)]      elevMatrix(m-1, 1)
 *   h_Cm [m+1] = C(m,.)     h_C2m [2m+1] = C(2m,.)
 *   d_out [B][nveh][4m+1] = alpha * (num^2/den^2) + beta   (max rate: alpha=-1, beta=maxAngRate^2)
 *   row_stride = S of the control-point rows (see top of file)
 */
typedef struct bez_angrate_tables bez_angrate_tables;
int bez_angrate_tables_create(int n, int elev, int device, const double *h_Tpos,
                              const double *h_elev1m, const double *h_Cm, const double *h_C2m,
                              bez_angrate_tables **out);
int bez_angrate_tables_destroy(bez_angrate_tables *tables);
int bez_angrate_sq(const bez_angrate_tables *tables, const double *d_cpts, const double *d_tf,
                   int B, int N, int row_stride, int veh_begin, int nveh,
                   double alpha, double beta, double *d_out, void *stream);

/* ---- A7: the finite-difference Jacobian SciPy's SLSQP forms from nvar+1 calls
 * of the constraint callables (scipy/optimize/_slsqp_py.py:349-367 ->
 * _numdiff.py:585-596, 683-712):  J[:,k] = (f(x+h_k e_k) - f(x)) / dx_k,
 * dx_k = (x_k+h_k) - x_k.  Both constraints are quadratic in the control points,
 * which are affine in x, so the quotient is evaluated in closed form
 *      J[:,k] = elev( scale * B(2a + dx_k*delta, delta) ),  delta = d a / d x_k
 * (no cancellation; only rows that depend on x_k are touched).
 *   d_cpts [N][S]   control points of the base x (bez_assemble_cpts with B = 1)
 *   ncols, offset   free control points per row of x and index of the first one
 *   d_dx   [nvar]   SciPy's divisors
 *   d_dir  [N][S]   d(control points)/d x_kdir for a variable that moves every
 *                   curve (tf of time-optimal problems), kdir = its index or -1
 *   dense = 1: d_out is J^T [nvar][ld] (zero-filled here, ld >= rows of the block;
 *              row k holds column k of J, so stores stay coalesced)
 *   dense = 0: sweep layout, [numVeh*dim*ncols][N-1][L] (partner curves in
 *              ascending order, the vehicle itself skipped) followed, when
 *              kdir >= 0, by [P][L] for the kdir column.
 */
int bez_jac_sepsq_elev(const bez_plan *plan, const double *d_cpts, int N, int numVeh,
                       int ncols, int offset, const double *d_dx, const double *d_dir,
                       int kdir, int dense, double *d_out, int64_t ld, void *stream);
/* speed rows: alpha as in bez_speed_sq_elev; sweep layout [numVeh*dim*ncols][L]
 * (+ [numVeh][L] for kdir). */
int bez_jac_speed_sq_elev(const bez_plan *plan, const double *d_cpts, int N, int numVeh,
                          int ncols, int offset, double tf, double alpha,
                          const double *d_dx, const double *d_dir, int kdir, int dense,
                          double *d_out, int64_t ld, void *stream);

/* Literal 2-point quotient (scipy/optimize/_numdiff.py:709-711) for constraint
 * blocks without a closed form (angular rate): d_F [nvar+1][m] holds f(x0) in row
 * 0 and f(x0 + h_k e_k) in row k+1; d_JT [nvar][m] = (F[k+1] - F[0]) / dx[k]. */
int bez_fd_quotient(const double *d_F, const double *d_dx, int nvar, int64_t m,
                    double *d_JT, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BEZGPU_H */
