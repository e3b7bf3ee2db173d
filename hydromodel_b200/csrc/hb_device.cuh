// hb_device.cuh -- device code of the batched Richards/BDF solver (sm_100a).
//
// Mapping: one CTA integrates one soil column.  Cell i of the column belongs to
// thread (i % NT); all per-column vectors (BDF difference table, tridiagonal
// Jacobian and its factors, Newton iterate, noise field) live in shared memory
// in natural order, so neighbour access is a conflict-free shared load and
// every element-wise update is private to the owning thread.  Scalars of the
// step controller (h, order, norms ...) are computed redundantly by every
// thread from block-wide reductions, which keeps control flow CTA-uniform
// without broadcasts.  The tridiagonal Newton systems are solved by warp 0 with
// a partition (chunked Thomas) / warp-shuffle PCR hybrid.
//
// Reference arithmetic reproduced (paths relative to /root/reference/code/src;
// SCIPY = scipy/integrate/_ivp of scipy 1.18.1, the package the reference calls):
//   closure()            models/vrettas_fung.py:108-249, models/vanGenuchten.py:71-117,
//                        utilities.py:4-54
//   rhs_eval()           richards_pde.py:82-160, :198-204, :306-376, :414-476, :542-595
//   num_jac()            SCIPY common.py:268-341,385-455
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
    unsigned long long seed;
    long long col_offset;
};

namespace hb {

extern __shared__ double smem[];

// ---- shared-memory layout: NVEC vectors of nxp doubles, then a small area ---
enum {
    V_D = 0,                // D[0..7]: BDF difference table
    V_JL = 8, V_JD, V_JU,   // Jacobian rows: lower, diagonal, upper
    V_TW, V_TIP, V_TVL, V_TVR,   // tridiagonal factors (num_jac scratch as well)
    V_Y,                    // Newton iterate / RHS input
    V_CM, V_FM,             // RHS scratch: C and flux at the midpoints
    V_B,                    // linear-system rhs / solution, RHS output
    V_FAC,                  // num_jac step factors
    V_NR,                   // noise field n_rnd of the current window
    NVEC
};
enum {                      // offsets (doubles) inside the small area
    X_RED = 0,              // [2][32] block-reduction slots
    X_RU = 64,              // [36] change_D matrix
    X_PCR = 100,            // [13][32] reduced-system coefficients
    X_PAR = 516,            // [HB_NPARAM] column scalars
    X_MISC = 516 + HB_NPARAM,    // [8]  0: p_left
    X_LATW = X_MISC + 8,    // [32] per-warp lateral-flow partial sums
    X_WMAX = X_LATW + 32,   // [32 ints] deepest unsaturated midpoint per warp
    X_END = X_WMAX + 16
};

__host__ __device__ inline size_t smem_doubles(int nxp) { return (size_t)NVEC * nxp + X_END; }

#define HB_VEC(c, k) (hb::smem + (size_t)(k) * (c).nxp)
#define HB_X(c, off) (hb::smem + (size_t)hb::NVEC * (c).nxp + (off))

struct Counters {
    long long nfev, njev, nlu, nsteps, nreject, nfail, nrhs, nnewton;
};

// args_i of simulation.py:591-596 in device form; bit 0 of `mode`: daylight,
// bit 1: SPINUP.
struct Args {
    double precip;          // already scaled by the column's multiplier
    int wtd;                // observed water-table node index
    int mode;
};

// BDF coefficient tables (bdf.py:244-247), filled by the host in double.
__constant__ double k_gamma[6];
__constant__ double k_alpha[6];
__constant__ double k_error_const[6];

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(HB_FULL, v, o);
    return v;
}

// Sum over the CTA, identical in every thread.  Contains one __syncthreads().
// `parity` alternates between two slot sets so that back-to-back reductions
// need no second barrier.
template <int NT>
__device__ __noinline__ double block_sum(const DevCtx &c, int &parity, double v)
{
    v = warp_sum(v);
    double *r = HB_X(c, X_RED) + (parity << 5);
    parity ^= 1;
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
    __syncthreads();
    double acc = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) acc += r[w];
    return acc;
}

template <int NT>
__device__ __forceinline__ int block_max_int(const DevCtx &c, int &parity, int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(HB_FULL, v, o));
    int *r = reinterpret_cast<int *>(HB_X(c, X_RED) + (parity << 5));
    parity ^= 1;
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
    __syncthreads();
    int acc = r[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) acc = max(acc, r[w]);
    return acc;
}

// RMS norm (SCIPY common.py:63-65) from this thread's partial sum of squares.
template <int NT>
__device__ __noinline__ double rms(const DevCtx &c, int &parity, double local_sq)
{
    return sqrt(block_sum<NT>(c, parity, local_sq)) / sqrt((double)c.nx);
}

// rate ** n for the small positive integers of the Newton convergence test
// (python float ** int; bdf.py:56).
__device__ __forceinline__ double ipow(double x, int n)
{
    double r = x;
    for (int i = 1; i < n; ++i) r *= x;
    return r;
}

struct Clos { double theta, K, C, kb; };

// One closure evaluation.  Same map as the reference's VrettasFung /
// vanGenuchten __call__, with the power functions composed from one shared
// log/exp pair; the cancellation-prone pieces (1 - Se, log(v/M^2 + 1),
// (q - theta_r)/delta_s) keep the reference's literal form.  Measured on the
// golden tapes this is as close to the reference as a literal pow() version.
__device__ __forceinline__ Clos closure(const double *par, double psi, double phi,
                                        double kmean, double coef, double e)
{
    const double psi_sat = par[HB_P_PSI_SAT];
    const double ksoil = par[HB_P_KSAT_SOIL];
    const bool vg = par[HB_P_MODEL] != 0.0;
    Clos r;
    if (psi >= psi_sat) {                       // saturated cell
        r.theta = phi;
        r.C = par[HB_P_EPS];
        r.kb = (!vg && coef >= 0.0) ? kmean : ksoil;
        r.K = r.kb;
        return r;
    }
    const double n = par[HB_P_N], m = par[HB_P_M], alpha = par[HB_P_ALPHA];
    const double th_r = par[HB_P_THETA_RES], lam = par[HB_P_LAMBDA];
    const double delta_s = phi - th_r;
    const double ax = alpha * fabs(psi);
    double xn, base, X;
    if (n == 2.0) {
        xn = ax * ax; base = 1.0 + xn; X = 1.0 / sqrt(base);
    } else {
        xn = exp(n * log(ax)); base = 1.0 + xn; X = exp(-m * log(base));
    }
    const double q = th_r + delta_s * X;
    const double s_eff = fmin(fmax((q - th_r) / delta_s, 0.0), 1.0);
    double K, kb;
    if (!vg) {
        const double selam = (lam == 1.0) ? s_eff : (lam == 2.0 ? s_eff * s_eff : pow(s_eff, lam));
        kb = ksoil;
        if (coef >= 0.0) {
            const double mean = (kmean == 0.0) ? 1.0e-7 : kmean;
            const double var = par[HB_P_SIGMA] * (1.0 - s_eff);
            const double l1 = log(var / (mean * mean) + 1.0);
            kb = mean * exp(-0.5 * l1 + sqrt(l1) * (coef * e));
        }
        K = selam * kb;
    } else {
        const double mth = 1.0 / base;
        const double u = 1.0 - ((n == 2.0) ? sqrt(1.0 - mth) : pow(1.0 - mth, m));
        kb = ksoil;
        K = kb * sqrt(s_eff) * ((n == 2.0) ? u * u : pow(u, n));
    }
    K = fmin(K, kb);
    double C = (m * n) * alpha * delta_s * (s_eff / base) * (xn / ax);
    if (C < par[HB_P_EPS] || !isfinite(C)) C = par[HB_P_EPS];
    r.theta = q; r.K = K; r.C = C; r.kb = kb;
    return r;
}

// Out-of-line copy for the cold call sites (boundary closure, diagnostics).
__device__ __noinline__ void closure_cold(const double *par, double psi, double phi, double kmean,
                                          double coef, double e, Clos *out)
{
    *out = closure(par, psi, phi, kmean, coef, e);
}

// dpsi/dt at the state held in vector V_Y (must be visible: caller
// synchronises before the call).  The result is written to vector `out`
// (each thread writes, and may afterwards read, its own cells).  Contains one
// __syncthreads(); writes V_CM, V_FM and the small slots WMAX, MISC, LATW.
template <int NT, int CPT>
__device__ __noinline__ void rhs_eval(const DevCtx &c, int col, double precip, int wtd,
                                      int mode, int out)
{
    const int nx = c.nx, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *par = HB_X(c, X_PAR);
    const double *y = HB_VEC(c, V_Y), *nr = HB_VEC(c, V_NR);
    double *cm = HB_VEC(c, V_CM), *fm = HB_VEC(c, V_FM);
    const double psi_sat = par[HB_P_PSI_SAT];
    const double *st = c.stat + (size_t)col * HB_NSTATIC * nx;
    int deepest = -1;                    // deepest unsaturated interior midpoint of this warp

    // ---- phase A: closures and fluxes at the midpoints ----------------------
#pragma unroll 1
    for (int r = 0; r < CPT; ++r) {
        const int k = tid + r * NT;
        bool unsat_interior = false;
        if (k < nx - 1) {
            const double y0 = y[k], y1 = y[k + 1];
            const double ym = 0.5 * (y0 + y1);
            const double dy = (y1 - y0) / c.dx[k];
            // noise index is local to the array handed to the closure (quirk Q1)
            const double e = nr[k == 0 ? 0 : k - 1];
            Clos cl = closure(par, ym, st[(size_t)HB_S_PHI_MID * nx + k],
                              st[(size_t)HB_S_KMEAN_MID * nx + k],
                              st[(size_t)HB_S_COEF_MID * nx + k], e);
            cm[k] = cl.C;
            fm[k] = cl.K * (dy - 1.0);
            unsat_interior = (k >= 1) && !(ym >= psi_sat);
        }
        const unsigned mask = __ballot_sync(HB_FULL, unsat_interior);
        if (mask) deepest = max(deepest, (warp << 5) + r * NT + (31 - __clz(mask)));
    }
    int *wmax = reinterpret_cast<int *>(HB_X(c, X_WMAX));
    if (lane == 0) wmax[warp] = deepest;

    // top boundary (bc_fun): closure at (y[0], z = x[0]) on the last thread,
    // which has no midpoint in the last round
    if (tid == NT - 1) {
        const double y0 = y[0];
        Clos b;
        closure_cold(par, y0, par[HB_P_PHI_TOP], par[HB_P_KMEAN_TOP], par[HB_P_COEF_TOP], nr[0], &b);
        const double q_inf_max = fmin(2.0 * (par[HB_P_PHI_TOP] - b.theta) * par[HB_P_DZ], b.kb);
        double p_left = 0.0;
        if (y0 < psi_sat) p_left = fmin((1.0 - par[HB_P_INTERCEPT]) * fabs(precip), q_inf_max);
        if (mode == 1) {                 // daylight and not SPINUP
            // np.all(theta_left > theta_res) over the broadcast profile (quirk Q2)
            const double th_r = par[HB_P_THETA_RES];
            const double ax = par[HB_P_ALPHA] * fabs(y0);
            const double X = (par[HB_P_N] == 2.0) ? 1.0 / sqrt(1.0 + ax * ax)
                                                   : pow(1.0 + pow(ax, par[HB_P_N]), -par[HB_P_M]);
            const bool sat0 = y0 >= psi_sat;
            const double dsm = sat0 ? par[HB_P_DSMIN_EXCL0] : par[HB_P_DSMIN_ALL];
            if ((th_r + dsm * X > th_r) && (!sat0 || par[HB_P_PHI_TOP] > th_r))
                p_left -= par[HB_P_SURF_EVAP];
        }
        HB_X(c, X_MISC)[0] = p_left;
    }
    __syncthreads();

    // ---- phase B: lateral-flow sink (monitoring mode) and assembly ----------
    int lo = 1, hi = 0;                  // active sink range in midpoint index: [lo, hi)
    if (((int)par[HB_P_FLAGS]) & HB_FLAG_LF) {
        int kmax = wmax[0];
#pragma unroll
        for (int w = 1; w < NT / 32; ++w) kmax = max(kmax, wmax[w]);
        const int dim_d = nx - 2;
        const int wtd_obs = min(wtd, dim_d - 1);
        const int wtd_est = (kmax < 0) ? 0 : min(kmax, dim_d - 1);
        if (wtd_est < wtd_obs) { lo = wtd_est + 1; hi = wtd_obs + 1; }
    }
    double lat = 0.0;
    double *fo = HB_VEC(c, out);
#pragma unroll 1
    for (int r = 0; r < CPT; ++r) {
        const int i = tid + r * NT;
        if (i < nx) {
            const double yi = y[i];
            double s_i = 0.0, s_m = 0.0, res;
            if (i >= lo && i < hi) {          // own midpoint i
                s_i = fmin(-2.5e-4 * (0.5 * (yi + y[i + 1])), 0.0);
                lat += fabs(s_i);
            }
            if (i - 1 >= lo && i - 1 < hi)    // midpoint i-1
                s_m = fmin(-2.5e-4 * (0.5 * (y[i - 1] + yi)), 0.0);
            if (i == 0) {
                const double zx = c.zx[0];
                double den = zx * cm[0];
                if (den == 0.0) den = 1.0;
                res = (HB_X(c, X_MISC)[0] + (fm[0] + zx * s_i)) / den;
            } else if (i < nx - 1) {
                const double zx = c.zx[i], xz = c.xz[i];
                double den = zx * cm[i] + xz * cm[i - 1];
                if (den == 0.0) den = 1.0;
                res = (fm[i] - fm[i - 1] + (zx * s_i + xz * s_m)) / den;
            } else {
                const double xz = c.xz[i];
                double den = -xz * cm[i - 1];
                if (den == 0.0) den = 1.0;
                res = (0.0 + (fm[i - 1] - xz * s_m)) / den;
            }
            fo[i] = res;
        }
    }
    lat = warp_sum(lat);
    if (lane == 0) HB_X(c, X_LATW)[warp] = lat;
}

// Lateral flow of the last RHS evaluation (arg_out, quirk Q6); thread 0 only.
template <int NT>
__device__ inline double lateral_flow(const DevCtx &c)
{
    double lat = 0.0;
    for (int w = 0; w < NT / 32; ++w) lat += HB_X(c, X_LATW)[w];
    return lat * HB_X(c, X_PAR)[HB_P_DZ];
}

// ---------------------------------------------------------------------------
// Tridiagonal partition solver, executed by warp 0 only.
// Matrix rows: a_i = -c*jl[i], b_i = 1 - c*jd[i], c_i = -c*ju[i].
// Lane p owns rows [p*m, min((p+1)*m, nx)); its last row is an interface
// unknown, the others are eliminated locally (Thomas, no pivoting) and their
// coupling to the two neighbouring interfaces is kept as spike vectors.  The 32
// interface unknowns form a tridiagonal system solved by PCR with shuffles.
// ---------------------------------------------------------------------------
__device__ __noinline__ void tri_factor(const DevCtx &c, double cc)
{
    const int lane = threadIdx.x & 31;
    const int nx = c.nx;
    const double *jl = HB_VEC(c, V_JL), *jd = HB_VEC(c, V_JD), *ju = HB_VEC(c, V_JU);
    double *tw = HB_VEC(c, V_TW), *tip = HB_VEC(c, V_TIP), *tvl = HB_VEC(c, V_TVL),
           *tvr = HB_VEC(c, V_TVR), *pcr = HB_X(c, X_PCR);
    const int m = (nx + 31) >> 5;
    const int s = lane * m;
    const int e = min(s + m, nx) - 1;           // inclusive; e < s: lane has no rows
    const int cnt = e - s + 1;
    double vl_last = 0.0, vr_last = 0.0, vl_first = 0.0, vr_first = 0.0;
    if (cnt >= 2) {
        // interior rows s..e-1
        double piv = 1.0 - cc * jd[s];
        double ip = 1.0 / piv;
        tw[s] = 0.0; tip[s] = ip;
        double g = -cc * jl[s];                 // rhs of the left spike, forward swept
        tvl[s] = g;
        for (int r = s + 1; r < e; ++r) {
            const double w = (-cc * jl[r]) * ip;
            piv = (1.0 - cc * jd[r]) - w * (-cc * ju[r - 1]);
            ip = 1.0 / piv;
            tw[r] = w; tip[r] = ip;
            g = -w * g;
            tvl[r] = g;
        }
        // back substitution for both spikes
        double zl = tvl[e - 1] * tip[e - 1];
        double zr = (-cc * ju[e - 1]) * tip[e - 1];
        tvl[e - 1] = zl; tvr[e - 1] = zr;
        for (int r = e - 2; r >= s; --r) {
            const double cu = (-cc * ju[r]) * tip[r];
            zl = tvl[r] * tip[r] - cu * zl;
            zr = -cu * zr;
            tvl[r] = zl; tvr[r] = zr;
        }
        vl_last = tvl[e - 1]; vr_last = tvr[e - 1];
        vl_first = tvl[s]; vr_first = tvr[s];
    }
    // interface row e
    double ae = 0.0, be = 1.0, ce = 0.0;
    if (cnt >= 1) { ae = -cc * jl[e]; be = 1.0 - cc * jd[e]; ce = (e < nx - 1) ? -cc * ju[e] : 0.0; }
    if (s == 0 && cnt == 1) ae = 0.0;
    // neighbour (lane+1) data
    const int cnt_n = __shfl_down_sync(HB_FULL, cnt, 1);
    const double vlf_n = __shfl_down_sync(HB_FULL, vl_first, 1);
    const double vrf_n = __shfl_down_sync(HB_FULL, vr_first, 1);
    const bool has_n = (lane < 31) && (cnt_n >= 1);
    double A, B, C;
    if (cnt >= 2) { A = -ae * vl_last; B = be - ae * vr_last; }
    else { A = ae; B = be; }
    if (!has_n) { C = 0.0; ce = 0.0; }
    else if (cnt_n >= 2) { B -= ce * vlf_n; C = -ce * vrf_n; }
    else C = ce;
    if (lane == 0) A = 0.0;
    pcr[11 * 32 + lane] = (cnt >= 2) ? ae : 0.0;             // multiplies z_last in rho
    pcr[12 * 32 + lane] = (has_n && cnt_n >= 2) ? ce : 0.0;  // multiplies z'_first in rho
#pragma unroll
    for (int st = 0; st < 5; ++st) {
        const int d = 1 << st;
        double Am = __shfl_up_sync(HB_FULL, A, d), Bm = __shfl_up_sync(HB_FULL, B, d),
               Cm = __shfl_up_sync(HB_FULL, C, d);
        double Ap = __shfl_down_sync(HB_FULL, A, d), Bp = __shfl_down_sync(HB_FULL, B, d),
               Cp = __shfl_down_sync(HB_FULL, C, d);
        if (lane < d) { Am = 0.0; Bm = 1.0; Cm = 0.0; }
        if (lane + d > 31) { Ap = 0.0; Bp = 1.0; Cp = 0.0; }
        const double k1 = A / Bm, k2 = C / Bp;
        pcr[(2 * st) * 32 + lane] = k1;
        pcr[(2 * st + 1) * 32 + lane] = k2;
        B = B - k1 * Cm - k2 * Ap;
        A = -k1 * Am;
        C = -k2 * Cp;
    }
    pcr[10 * 32 + lane] = 1.0 / B;
    __syncwarp();
}

// Solves in place: V_B holds the rhs on entry, the solution on exit.
__device__ __noinline__ void tri_solve(const DevCtx &c, double cc)
{
    const int lane = threadIdx.x & 31;
    const int nx = c.nx;
    const double *ju = HB_VEC(c, V_JU), *tw = HB_VEC(c, V_TW), *tip = HB_VEC(c, V_TIP),
                 *tvl = HB_VEC(c, V_TVL), *tvr = HB_VEC(c, V_TVR), *pcr = HB_X(c, X_PCR);
    double *b = HB_VEC(c, V_B);
    const int m = (nx + 31) >> 5;
    const int s = lane * m;
    const int e = min(s + m, nx) - 1;
    const int cnt = e - s + 1;
    double z_first = 0.0, z_last = 0.0;
    if (cnt >= 2) {
        double g = b[s];
        for (int r = s + 1; r < e; ++r) {       // forward sweep
            g = fma(-tw[r], g, b[r]);
            b[r] = g;
        }
        double z = g * tip[e - 1];
        b[e - 1] = z;
        for (int r = e - 2; r >= s; --r) {      // backward sweep
            const double ip = tip[r];
            const double cu = (-cc * ju[r]) * ip;
            z = fma(-cu, z, b[r] * ip);
            b[r] = z;
        }
        z_first = z; z_last = b[e - 1];
    }
    double rho = (cnt >= 1) ? b[e] : 0.0;
    const double zf_n = __shfl_down_sync(HB_FULL, z_first, 1);
    rho -= pcr[11 * 32 + lane] * z_last;
    rho -= pcr[12 * 32 + lane] * ((lane < 31) ? zf_n : 0.0);
#pragma unroll
    for (int st = 0; st < 5; ++st) {
        const int d = 1 << st;
        double rm = __shfl_up_sync(HB_FULL, rho, d);
        double rp = __shfl_down_sync(HB_FULL, rho, d);
        if (lane < d) rm = 0.0;
        if (lane + d > 31) rp = 0.0;
        rho = rho - pcr[(2 * st) * 32 + lane] * rm - pcr[(2 * st + 1) * 32 + lane] * rp;
    }
    const double xi = rho * pcr[10 * 32 + lane];
    double x_left = __shfl_up_sync(HB_FULL, xi, 1);
    if (lane == 0) x_left = 0.0;
    if (cnt >= 1) b[e] = xi;
    for (int r = s; r < e; ++r)
        b[r] = b[r] - x_left * tvl[r] - xi * tvr[r];
    __syncwarp();
}

// ---------------------------------------------------------------------------
// change_D (SCIPY bdf.py:18-33): D[0..order] <- (R U)^T D[0..order].
// ---------------------------------------------------------------------------
template <int NT>
__device__ __noinline__ void change_D(const DevCtx &c, int order, double factor)
{
    const int tid = threadIdx.x;
    const int q = order + 1;
    double *ru = HB_X(c, X_RU);
    __syncthreads();                       // previous readers of RU are done
    if (tid < q * q) {
        const int i = tid / q, j = tid % q;
        double acc = 0.0;
        for (int k = 0; k < q; ++k) {
            // R[i][k] with `factor`, U[k][j] with 1 (cumulative products down rows)
            double Rik = 1.0, Ukj = 1.0;
            for (int rr = 1; rr <= i; ++rr)
                Rik *= (k == 0) ? 0.0 : ((double)(rr - 1) - factor * (double)k) / (double)rr;
            for (int rr = 1; rr <= k; ++rr)
                Ukj *= (j == 0) ? 0.0 : ((double)(rr - 1) - 1.0 * (double)j) / (double)rr;
            acc += Rik * Ukj;
        }
        ru[i * 6 + j] = acc;
    }
    __syncthreads();
    double *D = HB_VEC(c, V_D);
    const int nxp = c.nxp;
    for (int p = tid; p < c.nx; p += NT) {
        double col[6], out[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) col[j] = (j < q) ? D[j * nxp + p] : 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j) if (j < q) acc += ru[j * 6 + i] * col[j];
            out[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) if (i < q) D[i * nxp + p] = out[i];
    }
}

// ---------------------------------------------------------------------------
// Forward-difference Jacobian on the tridiagonal pattern with scipy's column
// grouping and step-factor adaptation (SCIPY common.py:268-341,385-455).
// The evaluation point is taken from vector `src` (V_D: D[0] at window start;
// V_B: the predictor).  Uses V_TW (f at the base point), V_TIP (f at the
// perturbed point), V_TVL (copy of the base point), V_TVR (steps h) -- the
// tridiagonal factors are invalid after a Jacobian refresh anyway -- and V_Y as
// the RHS input.  Result: row storage in V_JL / V_JD / V_JU, factors in V_FAC.
// ---------------------------------------------------------------------------
template <int NT, int CPT>
__device__ __noinline__ void num_jac(const DevCtx &c, int col, double precip, int wtd, int mode,
                                     int src, bool first, Counters &cnt)
{
    const int nx = c.nx, tid = threadIdx.x;
    double *y = HB_VEC(c, V_Y), *fb = HB_VEC(c, V_TW), *fn = HB_VEC(c, V_TIP),
           *yb = HB_VEC(c, V_TVL), *hv = HB_VEC(c, V_TVR), *fac = HB_VEC(c, V_FAC);
    const double *ysrc = HB_VEC(c, src);
    double ysc[CPT], cl[CPT], cd[CPT], cu[CPT], maxd[CPT], scl[CPT];
    bool small_[CPT];

    __syncthreads();
    for (int i = tid; i < nx; i += NT) { const double v = ysrc[i]; yb[i] = v; y[i] = v; }
    __syncthreads();
    rhs_eval<NT, CPT>(c, col, precip, wtd, mode, V_TW); cnt.nrhs++;   // jac_wrapped: f = fun_single(t, y)
#pragma unroll
    for (int r = 0; r < CPT; ++r) {
        const int i = tid + r * NT;
        ysc[r] = 0.0; cl[r] = cd[r] = cu[r] = 0.0; maxd[r] = 0.0; scl[r] = 0.0;
        small_[r] = false;
        if (i < nx) {
            double fc = first ? sqrt(HB_EPS) : fac[i];
            const double fs = (fb[i] >= 0.0) ? 1.0 : -1.0;
            const double ybi = yb[i];
            ysc[r] = fs * fmax(c.atol, fabs(ybi));
            double h = (ybi + fc * ysc[r]) - ybi;
            while (h == 0.0) { fc *= 10.0; h = (ybi + fc * ysc[r]) - ybi; }
            fac[i] = fc; hv[i] = h;
        }
    }
    const double REJECT = pow(HB_EPS, 0.875), SMALL = pow(HB_EPS, 0.75), BIG = pow(HB_EPS, 0.25);

    // column statistics with scipy's sparse-argmax semantics (_data.py:249-275)
    auto column_stats = [&](int j, double &dl_, double &dd_, double &du_, double &md, double &sc) {
        const int lo = j > 0 ? j - 1 : 0, hi = j < nx - 1 ? j + 1 : nx - 1;
        int best = lo; double bv = fabs(fn[lo] - fb[lo]);
        for (int i = lo + 1; i <= hi; ++i) {
            const double v = fabs(fn[i] - fb[i]);
            if (v > bv) { bv = v; best = i; }
        }
        int mi;
        if (bv > 0.0) { mi = best; md = bv; }
        else { mi = 0; md = (lo == 0) ? fabs(fn[0] - fb[0]) : 0.0; }
        sc = fmax(fabs(fb[mi]), fabs(fn[mi]));
        dl_ = (j > 0) ? fn[j - 1] - fb[j - 1] : 0.0;
        dd_ = fn[j] - fb[j];
        du_ = (j < nx - 1) ? fn[j + 1] - fb[j + 1] : 0.0;
    };

    for (int g = 0; g < c.ng; ++g) {
        for (int i = tid; i < nx; i += NT) y[i] = yb[i] + ((c.grp[i] == g) ? hv[i] : 0.0);
        __syncthreads();
        rhs_eval<NT, CPT>(c, col, precip, wtd, mode, V_TIP); cnt.nrhs++;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < CPT; ++r) {
            const int j = tid + r * NT;
            if (j < nx && c.grp[j] == g) {
                column_stats(j, cl[r], cd[r], cu[r], maxd[r], scl[r]);
                small_[r] = maxd[r] < REJECT * scl[r];
            }
        }
    }
    bool any_small = false;
#pragma unroll
    for (int r = 0; r < CPT; ++r) any_small |= small_[r];
    if (__syncthreads_or(any_small)) {
        double hn[CPT];
#pragma unroll
        for (int r = 0; r < CPT; ++r) {
            const int i = tid + r * NT;
            hn[r] = 0.0;
            if (i < nx && small_[r]) hn[r] = (yb[i] + 10.0 * fac[i] * ysc[r]) - yb[i];
        }
        for (int g = 0; g < c.ng; ++g) {
            bool mine = false;
#pragma unroll
            for (int r = 0; r < CPT; ++r) {
                const int i = tid + r * NT;
                mine |= (i < nx) && small_[r] && (c.grp[i] == g);
            }
            if (!__syncthreads_or(mine)) continue;
#pragma unroll
            for (int r = 0; r < CPT; ++r) {
                const int i = tid + r * NT;
                if (i < nx) y[i] = yb[i] + ((c.grp[i] == g) ? hn[r] : 0.0);
            }
            __syncthreads();
            rhs_eval<NT, CPT>(c, col, precip, wtd, mode, V_TIP); cnt.nrhs++;
            __syncthreads();
#pragma unroll
            for (int r = 0; r < CPT; ++r) {
                const int j = tid + r * NT;
                if (j < nx && small_[r] && c.grp[j] == g) {
                    double nl, nd, nu, md, sn;
                    column_stats(j, nl, nd, nu, md, sn);
                    if (maxd[r] * sn < md * scl[r]) {
                        fac[j] = 10.0 * fac[j];
                        hv[j] = hn[r];
                        cl[r] = nl; cd[r] = nd; cu[r] = nu;
                        scl[r] = sn; maxd[r] = md;
                    }
                }
            }
        }
    }
    __syncthreads();                       // all reads of fb / fn are done
    // column storage: cl = J[j-1][j], cd = J[j][j], cu = J[j+1][j]
    double *jl = HB_VEC(c, V_JL), *jd = HB_VEC(c, V_JD), *ju = HB_VEC(c, V_JU);
#pragma unroll
    for (int r = 0; r < CPT; ++r) {
        const int j = tid + r * NT;
        if (j < nx) {
            const double h = hv[j];
            fb[j] = cl[r] / h;             // upper entry of row j-1
            jd[j] = cd[r] / h;
            fn[j] = cu[r] / h;             // lower entry of row j+1
            double fc = fac[j];
            if (maxd[r] < SMALL * scl[r]) fc *= 10.0;
            if (maxd[r] > BIG * scl[r]) fc *= 0.1;
            fac[j] = fmax(fc, 1e3 * HB_EPS);
        }
    }
    __syncthreads();
    for (int i = tid; i < nx; i += NT) {
        jl[i] = (i > 0) ? fn[i - 1] : 0.0;
        ju[i] = (i < nx - 1) ? fb[i + 1] : 0.0;
    }
    __syncthreads();
    cnt.njev++;
}

// ---------------------------------------------------------------------------
// One solve_ivp(method='BDF') call from t0 to t_bound.  y0 is read from
// global memory (y_glob); on exit V_Y holds solver.y, the last accepted state
// (ivp.py:659-666).  Returns 0 (finished) or -1 (step size underflow).
// ---------------------------------------------------------------------------
template <int NT, int CPT>
__device__ __noinline__ int bdf_attempt(const DevCtx &c, int col, double precip, int wtd, int mode,
                                        double t0, double t_bound, const double *y_glob,
                                        Counters &cnt)
{
    const int nx = c.nx, nxp = c.nxp, tid = threadIdx.x;
    const double rtol = c.rtol, atol = c.atol;
    const double newton_tol = fmax(10.0 * HB_EPS / rtol, fmin(0.03, sqrt(rtol)));
    double *D = HB_VEC(c, V_D), *y = HB_VEC(c, V_Y), *b = HB_VEC(c, V_B);
    double d_[CPT], psi_[CPT], scl_[CPT];
    int par_ = 0;                               // block-reduction slot parity

    // ---- BDF.__init__ (bdf.py:198-256) ---------------------------------------
    __syncthreads();
    for (int i = tid; i < nx; i += NT) { const double v = y_glob[i]; y[i] = v; D[i] = v; }
    __syncthreads();
    rhs_eval<NT, CPT>(c, col, precip, wtd, mode, V_D + 1); cnt.nrhs++; cnt.nfev++;   // f0 -> D[1]
    double h_abs;
    {   // select_initial_step (common.py:68-134), order 1
        const double interval = fabs(t_bound - t0);
        double l0 = 0.0, l1 = 0.0;
#pragma unroll
        for (int r = 0; r < CPT; ++r) {
            const int i = tid + r * NT;
            scl_[r] = 1.0;
            if (i < nx) {
                const double y0 = D[i];
                scl_[r] = atol + fabs(y0) * rtol;
                const double q0 = y0 / scl_[r], q1 = D[nxp + i] / scl_[r];
                l0 += q0 * q0; l1 += q1 * q1;
            }
        }
        const double d0 = rms<NT>(c, par_, l0), d1 = rms<NT>(c, par_, l1);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, interval);
        for (int i = tid; i < nx; i += NT) y[i] = D[i] + h0 * D[nxp + i];
        __syncthreads();
        rhs_eval<NT, CPT>(c, col, precip, wtd, mode, V_B); cnt.nrhs++; cnt.nfev++;
        double l2 = 0.0;
#pragma unroll
        for (int r = 0; r < CPT; ++r) {
            const int i = tid + r * NT;
            if (i < nx) { const double q2 = (b[i] - D[nxp + i]) / scl_[r]; l2 += q2 * q2; }
        }
        const double d2 = rms<NT>(c, par_, l2) / h0;
        const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3)
                                                       : sqrt(0.01 / fmax(d1, d2));
        h_abs = fmin(fmin(100.0 * h0, h1), interval);
    }
    num_jac<NT, CPT>(c, col, precip, wtd, mode, V_D, true, cnt);
    for (int i = tid; i < nx; i += NT) D[nxp + i] = D[nxp + i] * h_abs;
    int order = 1, n_equal = 0;
    bool lu_valid = false;
    double t = t0;
    double c_lu = 0.0;

    // ---- driver loop (ivp.py:659, base.py:159-197) ----------------------------
    while (t != t_bound) {
        const double min_step = 10.0 * fabs(nextafter(t, (double)INFINITY) - t);
        double h_loc;
        if (h_abs < min_step) {
            h_loc = min_step;
            change_D<NT>(c, order, min_step / h_abs);
            n_equal = 0;
        } else h_loc = h_abs;
        bool current_jac = false, accepted = false;
        int n_iter = 0;
        double t_new = t, safety = 0.0, error_norm = 0.0;
        while (!accepted) {
            if (h_loc < min_step) {
                // failure: solver.y is the last accepted state == D[0]
                __syncthreads();
                for (int i = tid; i < nx; i += NT) y[i] = D[i];
                __syncthreads();
                return -1;
            }
            double h = h_loc;
            t_new = t + h;
            if (t_new - t_bound > 0.0) {
                t_new = t_bound;
                change_D<NT>(c, order, fabs(t_new - t) / h_loc);
                n_equal = 0; lu_valid = false;
            }
            h = t_new - t; h_loc = fabs(h);
            // predictor, scale and psi (bdf.py:360-363); predictor kept in V_B
            __syncthreads();
#pragma unroll
            for (int r = 0; r < CPT; ++r) {
                const int i = tid + r * NT;
                psi_[r] = 0.0; scl_[r] = 1.0;
                if (i < nx) {
                    double acc = D[i], ps = 0.0;
                    for (int k = 1; k <= order; ++k) {
                        const double dk = D[k * nxp + i];
                        acc += dk; ps += dk * k_gamma[k];
                    }
                    scl_[r] = atol + rtol * fabs(acc);
                    psi_[r] = ps / k_alpha[order];
                    y[i] = acc;
                }
            }
            bool converged = false;
            const double cc = h / k_alpha[order];
            while (!converged) {
                if (!lu_valid) {
                    __syncthreads();
                    if (tid < 32) tri_factor(c, cc);
                    c_lu = cc; lu_valid = true; cnt.nlu++;
                }
                // solve_bdf_system (bdf.py:36-69); V_Y holds y_predict
#pragma unroll
                for (int r = 0; r < CPT; ++r) d_[r] = 0.0;
                __syncthreads();
                double dy_norm_old = 0.0; bool have_old = false; int k;
                for (k = 0; k < HB_NEWTON_MAXITER; ++k) {
                    rhs_eval<NT, CPT>(c, col, precip, wtd, mode, V_B); cnt.nrhs++; cnt.nfev++;
                    bool finite = true;
#pragma unroll
                    for (int r = 0; r < CPT; ++r) {
                        const int i = tid + r * NT;
                        if (i < nx) {
                            const double f = b[i];
                            finite &= isfinite(f);
                            b[i] = cc * f - psi_[r] - d_[r];
                        }
                    }
                    if (!__syncthreads_and(finite)) break;
                    if (tid < 32) tri_solve(c, c_lu);
                    cnt.nnewton++;
                    __syncthreads();
                    double loc = 0.0;
#pragma unroll
                    for (int r = 0; r < CPT; ++r) {
                        const int i = tid + r * NT;
                        if (i < nx) {
                            const double dy = b[i];
                            const double q = dy / scl_[r];
                            loc += q * q;
                            y[i] += dy;            // tentative: discarded if we break
                            d_[r] += dy;
                        }
                    }
                    const double dy_norm = rms<NT>(c, par_, loc);
                    double rate = 0.0;
                    if (have_old) {
                        rate = dy_norm / dy_norm_old;
                        if (rate >= 1.0 ||
                            ipow(rate, HB_NEWTON_MAXITER - k) / (1.0 - rate) * dy_norm > newton_tol)
                            break;
                    }
                    if (dy_norm == 0.0 || (have_old && rate / (1.0 - rate) * dy_norm < newton_tol)) {
                        converged = true; break;
                    }
                    dy_norm_old = dy_norm; have_old = true;
                }
                n_iter = (k < HB_NEWTON_MAXITER) ? k + 1 : HB_NEWTON_MAXITER;
                if (!converged) {
                    // rebuild the predictor (V_Y was advanced tentatively)
                    __syncthreads();
                    for (int i = tid; i < nx; i += NT) {
                        double acc = D[i];
                        for (int kk = 1; kk <= order; ++kk) acc += D[kk * nxp + i];
                        y[i] = acc; b[i] = acc;
                    }
                    if (current_jac) break;
                    num_jac<NT, CPT>(c, col, precip, wtd, mode, V_B, false, cnt);
                    // num_jac used V_Y as its work vector: restore the predictor
                    for (int i = tid; i < nx; i += NT) y[i] = HB_VEC(c, V_TVL)[i];
                    lu_valid = false; current_jac = true;
                }
            }
            if (!converged) {
                h_loc *= 0.5;
                change_D<NT>(c, order, 0.5);
                n_equal = 0; lu_valid = false; cnt.nreject++;
                continue;
            }
            safety = 0.9 * (2 * HB_NEWTON_MAXITER + 1) / (double)(2 * HB_NEWTON_MAXITER + n_iter);
            double le = 0.0;
#pragma unroll
            for (int r = 0; r < CPT; ++r) {
                const int i = tid + r * NT;
                if (i < nx) {
                    scl_[r] = atol + rtol * fabs(y[i]);
                    const double q = (k_error_const[order] * d_[r]) / scl_[r];
                    le += q * q;
                }
            }
            error_norm = rms<NT>(c, par_, le);
            if (error_norm > 1.0) {
                const double factor = fmax(HB_MIN_FACTOR,
                                           safety * pow(error_norm, -1.0 / (order + 1)));
                h_loc *= factor;
                change_D<NT>(c, order, factor);
                n_equal = 0; cnt.nreject++;
            } else accepted = true;
        }
        n_equal++;
        t = t_new;
        h_abs = h_loc;
        cnt.nsteps++;
        double lm = 0.0, lp = 0.0;
        const bool decide = (n_equal >= order + 1);
#pragma unroll
        for (int r = 0; r < CPT; ++r) {
            const int i = tid + r * NT;
            if (i < nx) {
                D[(order + 2) * nxp + i] = d_[r] - D[(order + 1) * nxp + i];
                D[(order + 1) * nxp + i] = d_[r];
                for (int k = order; k >= 0; --k) D[k * nxp + i] += D[(k + 1) * nxp + i];
                if (decide) {
                    if (order > 1) {
                        const double q = (k_error_const[order - 1] * D[order * nxp + i]) / scl_[r];
                        lm += q * q;
                    }
                    if (order < HB_MAX_ORDER) {
                        const double q = (k_error_const[order + 1] * D[(order + 2) * nxp + i]) / scl_[r];
                        lp += q * q;
                    }
                }
            }
        }
        if (!decide) continue;
        double em = INFINITY, ep = INFINITY;
        if (order > 1) em = rms<NT>(c, par_, lm);
        if (order < HB_MAX_ORDER) ep = rms<NT>(c, par_, lp);
        const double fm_ = pow(em, -1.0 / order), fe_ = pow(error_norm, -1.0 / (order + 1)),
                     fp_ = pow(ep, -1.0 / (order + 2));
        int delta = -1; double best = fm_;
        if (fe_ > best) { best = fe_; delta = 0; }
        if (fp_ > best) { best = fp_; delta = 1; }
        order += delta;
        const double factor = fmin(HB_MAX_FACTOR, safety * best);
        h_abs *= factor;
        change_D<NT>(c, order, factor);
        n_equal = 0; lu_valid = false;
    }
    __syncthreads();
    return 0;
}

// RichardsPDE.solve (richards_pde.py:478-537): up to 5 attempts, the noise
// field damped in place by 0.8 after each failure (quirk Q4); the state of the
// last attempt is left in V_Y (quirk Q5).  Returns 0 ok / 1 failed.
template <int NT, int CPT>
__device__ inline int solve_window(const DevCtx &c, int col, const Args &a, double t0, double t1,
                                   const double *y_glob, Counters &cnt, bool &noise_damped)
{
    int trials = HB_N_TRIALS, status = 0;
    while (trials > 0) {
        status = bdf_attempt<NT, CPT>(c, col, a.precip, a.wtd, a.mode, t0, t1, y_glob, cnt);
        if (status == 0) break;
        for (int i = threadIdx.x; i < c.nx; i += NT) HB_VEC(c, V_NR)[i] *= HB_NOISE_DAMP;
        noise_damped = true;
        trials--; cnt.nfail++;
        __syncthreads();
    }
    return status == 0 ? 0 : 1;
}

// find_wtd(y >= psi_sat) on the nodes (utilities.py:56-99).  One sync inside.
template <int NT>
__device__ inline int find_wtd_nodes(const DevCtx &c, int &parity, const double *y)
{
    const double psi_sat = HB_X(c, X_PAR)[HB_P_PSI_SAT];
    int deepest = -1;
    for (int i = threadIdx.x; i < c.nx; i += NT)
        if (!(y[i] >= psi_sat)) deepest = i;
    deepest = block_max_int<NT>(c, parity, deepest);
    return (deepest < 0) ? 0 : min(deepest + 1, c.nx - 1);
}

// Philox4x32-10 + Box-Muller: standard normal number `idx` of stream
// (seed, column, window).
__device__ inline void philox4x32_10(uint32_t ctr[4], const uint32_t key[2])
{
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
        const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n1 = lo1, n2 = hi0 ^ ctr[3] ^ k1, n3 = lo0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__device__ inline double philox_normal(uint64_t seed, int64_t column, int window, int idx)
{
    uint32_t ctr[4] = {(uint32_t)(idx >> 1), (uint32_t)window, (uint32_t)(column & 0xffffffff),
                       (uint32_t)((uint64_t)column >> 32)};
    const uint32_t key[2] = {(uint32_t)(seed & 0xffffffff), (uint32_t)(seed >> 32)};
    philox4x32_10(ctr, key);
    // two 53-bit uniforms in (0,1]
    const uint64_t a = (((uint64_t)ctr[0] << 32) | ctr[1]) >> 11, b = (((uint64_t)ctr[2] << 32) | ctr[3]) >> 11;
    const double u1 = ((double)a + 1.0) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)b + 0.5) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    return (idx & 1) ? rad * sn : rad * cs;
}

}  // namespace hb
