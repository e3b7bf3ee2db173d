/*
 * hydrob200.h -- C-ABI of libhydrob200.so, the B200 (sm_100a) batched solver
 * for the 1-D Richards equation time-stepping core of vrettasm/HydroModel.
 *
 * The reference has no FFI: the seam this library replaces is the Python call
 *     RichardsPDE.solve(t_span, y0, args_i)          code/src/richards_pde.py:478-537
 * made once per forcing row by
 *     Simulation.run                                 code/src/simulation.py:576-649
 *     Simulation.initial_conditions                  code/src/simulation.py:444-480
 * together with the companion diagnostic call h_model(y_i, z, args_i)
 * (simulation.py:564,623) and the arg_out side channel (simulation.py:629-630).
 * Here one context holds an ENSEMBLE of independent columns and the whole
 * window loop runs on the device.
 *
 * Conventions: plain C types; every pointer is a HOST pointer unless the name
 * ends in _dev; caller owns all buffers; functions return 0 on success and a
 * negative code on error (text via hb_last_error()); no exceptions cross the
 * boundary.  One context per GPU; calls on one context must be serialised by
 * the caller; kernels are asynchronous on the context's stream, hb_sync() or
 * any hb_get_* waits for them.  Array layouts: include/hb_layout.h.
 */
#ifndef HYDROB200_H
#define HYDROB200_H

#include <stdint.h>
#include "hb_layout.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_ctx hb_ctx;

/* Library / device management. */
const char *hb_last_error(void);
int  hb_version(void);
/* nx: nodes per column (4 <= nx <= hb_max_nx()); n_cols: columns on this GPU. */
hb_ctx *hb_create(int device, int nx, int n_cols);
int  hb_destroy(hb_ctx *ctx);
int  hb_max_nx(void);
/* Launch on an existing CUDA stream (cudaStream_t passed as void*; NULL = the
 * context's own stream).  Lets a caller time the kernels with its own events. */
int  hb_set_stream(hb_ctx *ctx, void *cuda_stream);
int  hb_sync(hb_ctx *ctx);

/* Grid shared by all columns of the context (RichardsPDE.__init__,
 * richards_pde.py:48-79) and the finite-difference colouring scipy's BDF uses
 * for the tridiagonal jac_sparsity (richards_pde.py:499-503; scipy
 * bdf.py:263-268, optimize/_numdiff.py:227-284). */
int  hb_set_grid(hb_ctx *ctx, const double *x, const int32_t *groups, int n_groups);
/* solve_ivp(rtol, atol); the reference hard-codes 1e-3 / 1e-3 (richards_pde.py:496). */
int  hb_set_tolerances(hb_ctx *ctx, double rtol, double atol);

/* Static description of columns [col0, col0+ncol): par[ncol][HB_NPARAM],
 * stat[ncol][HB_NSTATIC][nx].  Replaces the object graph Simulation.setupModel
 * hands to RichardsPDE through mData (simulation.py:146-231,355). */
int  hb_set_columns(hb_ctx *ctx, int col0, int ncol, const double *par, const double *stat);

/* Pressure-head state psi[ncol][nx] (y0 of simulation.py:514 / IC_Filename). */
int  hb_set_state(hb_ctx *ctx, int col0, int ncol, const double *psi);
int  hb_get_state(hb_ctx *ctx, int col0, int ncol, double *psi);

/* Forcing: n_series tables of T rows (args_i of simulation.py:591-596):
 * precip[n_series][T], wtd_idx[n_series][T] (node index, -1 = skip window),
 * daylight[n_series][T], atm[n_series][T].  series_of_col[n_cols] maps each
 * column to its table (NULL: all columns use table 0). */
int  hb_set_forcing(hb_ctx *ctx, int n_series, int T, const double *precip,
                    const int32_t *wtd_idx, const int32_t *daylight, const double *atm,
                    const int32_t *series_of_col);

/* Noise fields n_rnd.  TAPE: tape[n_cols_given][n_slots][nx] standard normals
 * drawn on the host in the reference's order (slot 0 = the run's first draw,
 * simulation.py:561; slot k = k-th refresh, simulation.py:599-602).
 * PHILOX: generated on the device from (seed, global column id, window). */
int  hb_set_noise_tape(hb_ctx *ctx, int col0, int ncol, int n_slots, const double *tape);
int  hb_set_noise_philox(hb_ctx *ctx, uint64_t seed, int64_t first_global_column);

/* Output plan: per-window scalars are always kept ([T][HB_NWSCAL] per column);
 * the four fields (psi, theta, K, K_bkg) are kept for rows i with
 * i % out_stride == 0 and for the last row; out_stride <= 0 keeps none.
 * Must be called after hb_set_forcing. */
int  hb_set_output(hb_ctx *ctx, int out_stride);
int  hb_n_output_rows(hb_ctx *ctx);

/* Simulation.initial_conditions (simulation.py:389-493) for every column:
 * repeats the unit window (0,1) with forcing row 0 under SPINUP flags until
 * |wtd_obs - wtd_est| <= 2 dz and mean((y_j - y_{j-1})^2) <= 0.01 or max_iter
 * (reference: 1500).  nrnd0[n_cols][nx] is the spin-up noise draw
 * (simulation.py:426) in TAPE mode, NULL in PHILOX mode.  zwtd0_cm[n_cols] is
 * the first snapped water-table observation.  iters_out[n_cols] may be NULL. */
int  hb_spinup(hb_ctx *ctx, int max_iter, const double *nrnd0, const double *zwtd0_cm,
               int32_t *iters_out);

/* Simulation.run window loop (simulation.py:576-649) for rows [w0, w1) of
 * every column; row 0 only records the initial diagnostics.  Asynchronous. */
int  hb_run(hb_ctx *ctx, int w0, int w1);

/* Fast mode of hb_run: flags = OR of HB_FAST_* (hb_layout.h), 0 (default) = the reference's
 * arithmetic.  Departs from scipy common.py:63-134,268-455 and bdf.py:198-241, and with
 * HB_FAST_F32 from the double-precision closure of vrettas_fung.py:108-249 inside the Newton
 * iteration, as the flag comments say; any flag also turns the divisions of the element-wise loops
 * into multiplications by a refined reciprocal (1 ulp).  hb_spinup and the hooks always run the
 * parity path. */
int  hb_set_fast_mode(hb_ctx *ctx, int flags);

/* Scheduling of hb_run.  Columns are independent, so the schedule never changes a
 * result, only the length of the launch: with adaptive != 0 (default) the columns
 * the previous hb_run on this context found expensive (RHS evaluations well above
 * the mean) are started first and integrated back to back, the others are time
 * sliced in chunks of rows; adaptive == 0 keeps the plain chunk-major order. */
int  hb_set_schedule(hb_ctx *ctx, int adaptive);

/* Results.  wscal[ncol][T][HB_NWSCAL]; fields[ncol][n_out][HB_NFIELD][nx];
 * counters[ncol][HB_NCOUNTER].  The _rows variants read a slice of the rows
 * (wscal[ncol][nrow][HB_NWSCAL], fields[ncol][nout][HB_NFIELD][nx]) -- the form a
 * long run integrated slice by slice (hb_run(w0, w1)) reads its results in.  A
 * getter fails if a kernel raised an error (e.g. a noise tape that is too short). */
int  hb_get_wscal(hb_ctx *ctx, int col0, int ncol, double *wscal);
int  hb_get_wscal_rows(hb_ctx *ctx, int col0, int ncol, int row0, int nrow, double *wscal);
int  hb_get_fields(hb_ctx *ctx, int col0, int ncol, double *fields);
int  hb_get_fields_rows(hb_ctx *ctx, int col0, int ncol, int out0, int nout, double *fields);
int  hb_get_counters(hb_ctx *ctx, int col0, int ncol, int64_t *counters);
int  hb_reset_counters(hb_ctx *ctx);
/* Ensemble statistics over this context's columns, written to DEVICE buffers (so
 * that the caller can all-reduce them with NCCL; sums do not depend on the
 * schedule): per row, the number of columns that contributed (skipped windows do
 * not), the sum and the sum of squares of the estimated water-table node index
 * (count_dev, wtd_sum_dev, wtd_sumsq_dev: [T] each; count_dev may be NULL); and,
 * when theta_sum_dev / theta_sumsq_dev are given ([n_out][nx] each), the same two
 * moments of the water content theta at every output row and node. */
int  hb_moments_dev(hb_ctx *ctx, double *count_dev, double *wtd_sum_dev, double *wtd_sumsq_dev,
                    double *theta_sum_dev, double *theta_sumsq_dev);
/* Number of this library's kernel launches since hb_create. */
int64_t hb_launch_count(hb_ctx *ctx);
/* Diagnostic builds (-DHB_PHASE_TIMING=1) only: cycles thread 0 spent in each
 * phase of the integrator loop, summed over the columns since the last call;
 * all zero in production builds. */
int  hb_get_phase_cycles(hb_ctx *ctx, int64_t *out16);
/* FP64 roofline denominator: TFLOP/s of a register-resident DFMA loop on this
 * GPU (burst figure, best of 3). */
double hb_measure_fp64_peak(hb_ctx *ctx);

/* ---- test hooks: the pieces of the path exposed one by one --------------- */
/* RichardsPDE.__call__ (richards_pde.py:82-160) for column col at state psi. */
int  hb_rhs(hb_ctx *ctx, int col, const double *psi, const double *nrnd, double precip,
            int wtd, int daylight, int spinup, double atm, double *dydt, double *lateral_flow,
            double *transpiration);
/* h_model(psi, z, args) on the nodes: out[4][nx] = theta, K, C, K_bkg. */
int  hb_closure_nodes(hb_ctx *ctx, int col, const double *psi, const double *nrnd,
                      double *out);
/* scipy num_jac on the tridiagonal pattern (common.py:268-455) at psi:
 * row storage dl/d/du[nx] and the updated step factors. */
int  hb_num_jac(hb_ctx *ctx, int col, const double *psi, const double *nrnd, double precip,
                int wtd, int daylight, int spinup, const double *fac_in,
                double *dl, double *d, double *du, double *fac_out);
/* Tridiagonal solve (I - c J) x = b with the in-kernel partition solver. */
int  hb_tri_solve(hb_ctx *ctx, const double *dl, const double *d, const double *du,
                  double c, const double *b, double *x);
/* find_wtd(psi >= psi_sat) of column col on the device (utilities.py:56-99). */
int  hb_find_wtd(hb_ctx *ctx, int col, const double *psi, int32_t *wtd_node);
/* The n standard normals the PHILOX mode draws for (seed, global column, window):
 * window >= 1 is the refresh field of that forcing row, 0 the run's base field,
 * -1 the spin-up field. */
int  hb_philox_normals(hb_ctx *ctx, uint64_t seed, int64_t global_column, int window, int n,
                       double *out);
/* One RichardsPDE.solve (richards_pde.py:478-537) for column col: y0 -> y_end,
 * nrnd is damped in place on failure; stats[HB_NCOUNTER]; t_accepted (may be NULL,
 * else room for nx values) receives the accepted step times sol.t[1:] of the last
 * solve_ivp attempt. */
int  hb_solve_window(hb_ctx *ctx, int col, double t0, double t1, const double *y0,
                     double *nrnd, double precip, int wtd, int daylight, int spinup, double atm,
                     double *y_end, int64_t *stats, double *lateral_flow, double *transpiration,
                     int *status, double *t_accepted);

#ifdef __cplusplus
}
#endif
#endif /* HYDROB200_H */
