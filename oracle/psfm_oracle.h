/*
 * psfm_oracle.h — CPU ORACLE for the two hot paths.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (particle-sfm_b200/) never links, imports or
 * falls back to it.
 *
 * PARITY UNPINNED: the reference ships no tests / golden vectors for these paths and
 * cannot be compiled here (Ceres 2.0.0, Eigen, COLMAP bd84ad6 absent — SURVEY.md §8c).
 * The oracle is therefore a restatement of the *published algorithm* of Ceres 2.0.0
 * (TrustRegionMinimizer, LevenbergMarquardtStrategy, DoglegStrategy, Corrector,
 * SchurEliminator, ConjugateGradientsSolver, SchurJacobiPreconditioner) anchored on the
 * reference's own call sites, and is pinned only by the known-answer tests in tests/
 * (closed-form minimisers, finite differences, scipy.optimize.least_squares optima).
 *
 * It reuses the problem/option/summary structs of include/psfm_b200.h so that the
 * parity tests feed both sides byte-identical inputs.
 */
#ifndef PSFM_ORACLE_H_
#define PSFM_ORACLE_H_

#include "../include/psfm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* HP1: restates optimize_location (trajectory_optimize.cpp:30-96) with
   PathConsistencyError (path_consistency_cost.h:42-59), BiLinearInterpolator
   (linear_interpolation.h:97-123), ceres::Grid2D clamping, and Ceres' trust-region
   loop with TRADITIONAL_DOGLEG + per-block 4x4 Cholesky (== SPARSE_NORMAL_CHOLESKY on a
   block-diagonal normal matrix).
   reduction_mode 0: canonical chunked tree (bit-reproducible, shared definition with
                     the CUDA kernel — see DESIGN.md "canonical sum");
   reduction_mode 1: plain left-to-right sums (used to audit decision margins).
   num_threads <= 0 -> omp default. */
int psfm_oracle_traj_optimize(const double* uv12, const double* ref1, const double* ref2,
                              const double* scale, const float* flow12, int32_t n,
                              int32_t w, int32_t h, const psfm_traj_options* opts,
                              double* out_uv12, psfm_traj_summary* summary,
                              int32_t num_threads, int32_t reduction_mode);

/* Residuals (6n) and the dense 6x4 Jacobian blocks (24n, row-major per block) at uv12. */
int psfm_oracle_traj_evaluate(const double* uv12, const double* ref1, const double* ref2,
                              const double* scale, const float* flow12, int32_t n,
                              int32_t w, int32_t h, double* residuals, double* jacobians);

/* The bilinear interpolator alone: value (2), d/dr (2), d/dc (2) at (r, c). */
void psfm_oracle_bilinear(const float* flow, int32_t w, int32_t h, double r, double c,
                          double* f, double* dfdr, double* dfdc);

/* HP2: restates BundleAdjuster::Solve (bundle_adjustment.cc:259-320) on the flattened
   problem: SIMPLE_PINHOLE reprojection (COLMAP cost_functions.h), loss corrector,
   quaternion/subset parameterisations, LM, Schur elimination of the points and either
   a dense Cholesky of the reduced camera system (DENSE_/SPARSE_SCHUR) or CG with the
   Schur-Jacobi preconditioner (ITERATIVE_SCHUR).
   opts->linear_solver: AUTO follows the reference rule; EXACT_SCHUR = Cholesky. */
int psfm_oracle_ba_solve(psfm_ba_problem* problem, const psfm_ba_options* opts,
                         psfm_ba_summary* summary, int32_t num_threads);

/* Cost, loss-corrected residuals [2M], tangent gradient (camera side [6F+3C] in slot
   layout, inactive slots 0; point side [3P]) at the problem's current state. */
int psfm_oracle_ba_evaluate(const psfm_ba_problem* problem, const psfm_ba_options* opts,
                            double* cost, double* residuals, double* gradient_cam,
                            double* gradient_pts, int32_t num_threads);

/* Per-observation loss-corrected, UNscaled Jacobian blocks for finite-difference tests:
   jc [M*12] (2x6: rot3,t3), jp [M*6] (2x3), jk [M*6] (2x3: f,cx,cy). */
int psfm_oracle_ba_jacobians(const psfm_ba_problem* problem, const psfm_ba_options* opts,
                             double* jc, double* jp, double* jk);

/* One linear solve at the current state with LM diagonal from `radius`:
   writes the (scaled-space) step for camera slots [6F+3C] and points [3P].
   Used to check GPU PCG / exact steps against the Cholesky step. */
int psfm_oracle_ba_linear_step(const psfm_ba_problem* problem, const psfm_ba_options* opts,
                               double radius, int32_t solver, double* step_cam,
                               double* step_pts, int32_t* num_linear_iterations,
                               int32_t num_threads);

int psfm_oracle_num_threads(void);

/* Option defaults restated independently of the product library (the oracle must not
   link it): bundle_adjustment.h:48-102, controllers/global_mapper.cc:41-71,
   trajectory_optimize.cpp:74-79 + Ceres 2.0.0 defaults. */
void psfm_oracle_ba_default_options(psfm_ba_options* o);
void psfm_oracle_ba_global_options(psfm_ba_options* o);
void psfm_oracle_traj_default_options(psfm_traj_options* o);

#ifdef __cplusplus
}
#endif
#endif
