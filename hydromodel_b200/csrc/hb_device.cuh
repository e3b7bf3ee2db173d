// hb_device.cuh -- device code of the batched Richards/BDF solver (sm_100a):
// this file holds the context, the memory layouts and the constants; the device
// functions (hb_device_impl.cuh) and the kernels (hb_kernels.cuh) are included at
// the bottom, once per instance namespace (hbr / hbf, see there).
//
// Mapping: one CTA integrates one soil column.  Cell i of the column belongs to
// thread (i % NT); the per-column vectors of the Newton iteration (tridiagonal
// Jacobian and its factors, Newton iterate, RHS scratch, noise field, static
// profiles) live in shared memory in natural order, so neighbour access is a
// conflict-free shared load and every element-wise update is private to the
// owning thread; vectors touched once per step or only by their owner (BDF
// difference table, step factors, ET vectors) live in a per-CTA global scratch.
// Scalars of the step controller (h, order, norms ...) are computed redundantly
// by every thread from block-wide reductions, which keeps control flow
// CTA-uniform without broadcasts.  The tridiagonal Newton systems are solved by
// warp 0 with a partition (chunked Thomas) / PCR hybrid, the back-substitution by
// all threads.
//
// The integrator is written as ONE loop around ONE inlined right-hand-side
// evaluation (a state machine): the instruction footprint of the hot loop is
// what bounds this kernel, not arithmetic (profiles/).
//
// Reference arithmetic reproduced (paths relative to /root/reference/code/src;
// SCIPY = scipy/integrate/_ivp of scipy 1.18.1, the package the reference calls):
//   closure()            models/vrettas_fung.py:108-249, models/vanGenuchten.py:71-117,
//                        utilities.py:4-54
//   rhs_eval()           richards_pde.py:82-160, :198-204, :306-376, :414-476, :542-595
//   jacobian states      SCIPY common.py:268-341,385-455 (num_jac, grouped columns)
//   tri_factor/solve()   in place of splu / LU.solve, SCIPY bdf.py:222-227
//   change_D()           SCIPY bdf.py:18-33
//   bdf_attempt()        SCIPY bdf.py:36-69,198-256,314-455; common.py:63-134;
//                        base.py:159-197; ivp.py:659-666
//   solve_window()       richards_pde.py:478-537
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/hb_layout.h"

#define HB_EPS 2.220446049250313e-16
#define HB_ERR_TAPE_SHORT 1   // a column refreshed its noise field more often than the tape has slots
#define HB_FULL 0xffffffffu

// Everything a kernel needs, passed by value as a __grid_constant__ parameter
// (lives in the constant bank: reading a field costs no register).
struct DevCtx {
    int nx, nxp, n_cols, T, ng, n_out, out_stride, n_slots, noise_mode, n_series;
    double rtol, atol;
    const double *x, *dx, *zx, *xz;
    const int *grp;
    const double *par, *stat;
    double *psi, *noise_base;
    const double *tape;
    int *slot_next;
    const double *precip, *atm;
    const int *wtd_idx, *daylight, *series;
    double *wscal, *fields;
    long long *counters;
    int *queue;
    long long *phase_cycles; // [16] summed phase timings (HB_PHASE_TIMING builds), else unused
    double *dbuf;            // [resident CTAs][G_ROWS][nxp] per-CTA global scratch (difference table, ...)
    int *done;               // [n_cols] finished row-chunks per column in this launch
    int *credit;             // [n_cols] chunks granted but not yet integrated (run_kernel scheduler)
    int *order;              // [n_cols] launch plan: hot columns first (plan_kernel), or null
    int *n_hot;              // [1] number of hot columns in `order`
    long long *cost;         // [n_cols] RHS evaluations per column in the current launch (next plan)
    double *t_acc;           // test hook only: accepted step times of the current solve ([nx]), else null
    int *err_flags;          // [1] HB_ERR_* bits raised by kernels, checked by hb_sync / hb_get_*
    int chunk;               // windows per work item of run_kernel
    int fast;                // HB_FAST_* flags of the fast build of run_kernel (0 in parity mode)
    int jac_fuse;            // 1: grouped forward differences of the Jacobian in one pass (jac_fused); 0: one evaluation per group
    int dbg;                 // HB_DUP_PROBE builds: bit mask of phases executed twice
    double sqrt_nx;          // sqrt((double)nx), the denominator of the RMS norm
    double newton_tol;       // max(10 eps / rtol, min(0.03, sqrt(rtol)))  (bdf.py:220)
    unsigned long long seed;
    long long col_offset;
};

#ifndef HB_PHASE_TIMING
#define HB_PHASE_TIMING 0
#endif
// Phase timing (diagnostic builds only): thread 0 accumulates clock64() deltas
// between the marks of the integrator loop; the marks include barrier waits.
#if HB_PHASE_TIMING
#define HB_TICK(c, NV, id)                                                             \
    do {                                                                               \
        if (threadIdx.x == 0) {                                                        \
            long long *tt_ = reinterpret_cast<long long *>(HB_XA(c, NV, X_TIME));      \
            const long long now_ = clock64();                                          \
            tt_[id] += now_ - tt_[15];                                                 \
            tt_[15] = now_;                                                            \
        }                                                                              \
    } while (0)
#else
#define HB_TICK(c, NV, id) do { } while (0)
#endif
// Diagnostic builds (-DHB_DUP_PROBE=1): phases selected by the bit mask DevCtx.dbg
// are executed twice (idempotently), so the change in throughput is the in-situ
// cost of that phase.  Production builds compile none of it.
#ifndef HB_DUP_PROBE
#define HB_DUP_PROBE 0
#endif
#if HB_DUP_PROBE
#define HB_DUP(c, bit) (((c).dbg & (bit)) ? 2 : 1)
#else
#define HB_DUP(c, bit) 1
#endif

// The cells of this thread (i = tid, tid + NT, ...) with a compile-time trip count, used in
// the per-step and per-Newton-iteration loops: the iterations are unrolled so that the loads of
// all cells are issued before the first dependent arithmetic (+3 % measured; converting the
// remaining, colder loops as well made the kernel smaller but 2.5 % slower -- left alone).
#define HB_FOR_CELLS(i) _Pragma("unroll") for (int r_ = 0; r_ < CPT; ++r_) if (const int i = tid + r_ * NT; i < nx)

namespace hb {

extern __shared__ double smem[];

// ---- shared-memory layout: vectors of nxp doubles, then a small area --------
enum {
    V_JL = 0, V_JD, V_JU,   // Jacobian rows: lower, diagonal, upper
    V_TW, V_TIP, V_TVL, V_TVR,   // tridiagonal factors (Jacobian scratch as well)
    V_Y,                    // Newton iterate / RHS input
    V_CM, V_FM,             // RHS scratch: C and flux at the midpoints
    V_B,                    // linear-system rhs / solution, RHS output
    V_NR,                   // noise field n_rnd of the current window
    V_PSI, V_SCL, V_DD,     // Newton: psi, scale, accumulated correction d
    NVEC_BASE,              // = 15
    V_PHI = NVEC_BASE, V_KM, V_CF,   // static profiles at the midpoints   (FULL only)
    NVEC_FULL               // = 18
};
// Shared memory is kept to 18 vectors (64 KiB at nx = 401) on purpose: two resident
// CTAs then fit the 132 KiB carve-out and leave 124 KiB of L1 for the difference
// tables and the register-spill stack -- the kernel is measurably sensitive to L1
// capacity (28 / 60 / 92 KiB of L1: 424 K / 463 K / 511 K column-windows/s).
// Vectors that are touched rarely or only by their owning thread live in a
// per-CTA global scratch instead (rows of DevCtx.dbuf):
enum {
    G_D = 0,                // rows 0..7: BDF difference table D[0..7]
    G_FAC = 8,              // num_jac step factors (Jacobian build only)
    G_TH = 9, G_SK = 10,    // theta and root-uptake sink at the midpoints (ET rows only)
    G_ROWS = 11
};
#define HB_GROW(c, k) ((c).dbuf + ((size_t)blockIdx.x * hb::G_ROWS + (k)) * HB_NXP(c))
// The BDF difference table D[0..7] (touched once per step, never inside a
// Newton iteration) lives in global memory: 8*nxp doubles per resident CTA,
// L2-resident, each element only ever touched by its owning thread.
enum {                      // offsets (doubles) inside the small area
    X_RED = 0,              // [2][32] block-reduction slots
    X_RU = 64,              // [36] change_D matrix
    X_PCR = 100,            // [14][32] reduced-system coefficients + exchange slots
    X_PAR = 548,            // [HB_NPARAM] column scalars
    X_MISC = 548 + HB_NPARAM,    // [8]  0: p_left  1: transpiration of the last RHS  2..: ET scalars
    X_LATW = X_MISC + 8,    // [32] per-warp lateral-flow partial sums
    X_WMAX = X_LATW + 32,   // [32 ints] deepest unsaturated midpoint per warp
    X_CNT = X_WMAX + 16,    // [8 long long] solver counters (thread 0 writes)
    X_SCAN = X_CNT + 8,     // [32] per-warp totals of the block prefix sum (ET)
    X_TIME = X_SCAN + 32,   // [16 long long] phase cycle counters (HB_PHASE_TIMING builds)
    X_ARGS = X_TIME + 16,   // [8] arguments of the current window: precip, atm, t0, t_bound, wtd, mode
    X_END = X_ARGS + 8
};
enum { K_NFEV = 0, K_NJEV, K_NLU, K_NSTEPS, K_NREJECT, K_NFAIL, K_NRHS, K_NNEWTON };

__host__ __device__ inline size_t smem_doubles(int nxp, bool full)
{
    return (size_t)(full ? NVEC_FULL : NVEC_BASE) * nxp + X_END;
}

#define HB_VEC(c, k) (hb::smem + (size_t)(k) * HB_NXP(c))
#define HB_XA(c, NV, off) (hb::smem + (size_t)(NV) * HB_NXP(c) + (off))

// args_i of simulation.py:591-596 in device form; mode: 0 night, 1 daylight
// (surface evaporation may act), 2 SPINUP.
struct Args {
    double precip;          // already scaled by the column's multiplier
    int wtd;                // observed water-table node index
    int mode;
    double atm;             // atmospheric demand of the row (ET)
};

// BDF coefficient tables (bdf.py:244-247), filled by the host in double.
__constant__ double k_gamma[6];
__constant__ double k_alpha[6];
__constant__ double k_error_const[6];
// U of change_D (bdf.py:18-33): U[k][j] = prod_{r=1..k} ((r-1) - j) / r, U[k][0] = (k == 0); row stride 6.
__constant__ double k_U[36];

}  // namespace hb

// ---------------------------------------------------------------------------
// The device functions and kernels are compiled twice (hb_device_impl.cuh,
// hb_kernels.cuh):
//   namespace hbr  padded vector length nxp read from DevCtx at run time (any grid);
//   namespace hbf  nxp fixed at 416 = 32 x 13 at compile time (385..416 nodes: the
//                  reference's 401-node column).  Every shared-memory vector base
//                  then is an immediate offset instead of a live register:
//                  fewer spills, fewer address instructions, +8 % measured.
// ---------------------------------------------------------------------------
#define HB_FIXED_NXP 416
#define HB_IMPL_NS hbr
#define HB_NXP(c) ((c).nxp)
#include "hb_device_impl.cuh"
#include "hb_kernels.cuh"
#undef HB_IMPL_NS
#undef HB_NXP
#define HB_IMPL_NS hbf
#define HB_NXP(c) HB_FIXED_NXP
#include "hb_device_impl.cuh"
#include "hb_kernels.cuh"
#undef HB_IMPL_NS
#undef HB_NXP
