// hb_device_impl.cuh -- device functions of the solver; included by hb_device.cuh once
// per instance namespace (HB_IMPL_NS, HB_NXP).  No include guard on purpose.
namespace HB_IMPL_NS {
using namespace hb;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(HB_FULL, v, o);
    return v;
}

// Sum over the CTA, identical in every thread.  Contains one __syncthreads().
// `parity` alternates between two slot sets so that back-to-back reductions
// need no second barrier.
template <int NT, int NV>
__device__ __forceinline__ double block_sum(const DevCtx &c, int &parity, double v)
{
    v = warp_sum(v);
    double *r = HB_XA(c, NV, X_RED) + (parity << 5);
    parity ^= 1;
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
    __syncthreads();
    // pairwise tree over the per-warp partial sums (depth log2 instead of a serial
    // chain of NT/32 dependent additions)
    constexpr int NW = NT / 32;
    double t[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) t[w] = r[w];
#pragma unroll
    for (int s = 1; s < NW; s <<= 1)
#pragma unroll
        for (int w = 0; w + s < NW; w += 2 * s) t[w] += t[w + s];
    return t[0];
}

template <int NT, int NV>
__device__ __noinline__ int block_max_int(const DevCtx &c, int &parity, int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(HB_FULL, v, o));
    int *r = reinterpret_cast<int *>(HB_XA(c, NV, X_RED) + (parity << 5));
    parity ^= 1;
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
    __syncthreads();
    int acc = r[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) acc = max(acc, r[w]);
    return acc;
}

// RMS norm (SCIPY common.py:63-65) from this thread's partial sum of squares.
template <int NT, int NV>
__device__ __noinline__ double rms(const DevCtx &c, int &parity, double local_sq)
{
    return sqrt(block_sum<NT, NV>(c, parity, local_sq)) / c.sqrt_nx;
}

// rate ** n for the small positive integers of the Newton convergence test
// (python float ** int; bdf.py:56).
__device__ __forceinline__ double ipow(double x, int n)
{
    double r = x;
    for (int i = 1; i < n; ++i) r *= x;
    return r;
}

// x ** (-1/q) for the step-size factors (bdf.py:399,444), out of line: cold.
__device__ __noinline__ double inv_root(double x, int q)
{
    return pow(x, -1.0 / q);
}

// Out-of-line IEEE double helpers.  The hot loop of this kernel is bound by
// its instruction footprint (profiles/): every inlined FP64 division, sqrt, exp
// or log costs 25-70 SASS instructions per call site, a call costs one.
__device__ __noinline__ double hb_div(double a, double b) { return a / b; }
__device__ __noinline__ double hb_sqrt(double a) { return sqrt(a); }
__device__ __noinline__ double hb_exp(double a) { return exp(a); }
__device__ __noinline__ double hb_log(double a) { return log(a); }

// a / b of the hot element-wise loops.  Parity mode: IEEE division (out of line).  Fast mode
// (never the default): a times a Newton-refined reciprocal, inline -- one ulp, a sixth of the instructions.
template <bool FAST>
__device__ __forceinline__ double div_hot(double a, double b)
{
    if (FAST) {
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
        double e = fma(-b, r, 1.0); r = fma(r, e, r);
        e = fma(-b, r, 1.0); r = fma(r, e, r);
        return a * r;
    }
    return hb_div(a, b);
}

struct Clos { double theta, K, C, kb; };

// Rare branches of the closure, out of line: van Genuchten-Mualem conductivity
// (vanGenuchten.py:91-104) and a general lambda exponent (vrettas_fung.py:229).
__device__ __noinline__ double closure_rare_K(const double *par, double s_eff, double base,
                                              double kb, bool vg)
{
    const double n = par[HB_P_N], m = par[HB_P_M];
    if (!vg) return pow(s_eff, par[HB_P_LAMBDA]) * kb;
    const double mth = 1.0 / base;
    const double u = 1.0 - ((n == 2.0) ? sqrt(1.0 - mth) : pow(1.0 - mth, m));
    return kb * sqrt(s_eff) * ((n == 2.0) ? u * u : pow(u, n));
}

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
        xn = ax * ax; base = 1.0 + xn; X = hb_div(1.0, hb_sqrt(base));
    } else {
        xn = hb_exp(n * hb_log(ax)); base = 1.0 + xn; X = hb_exp(-m * hb_log(base));
    }
    const double q = th_r + delta_s * X;
    const double s_eff = fmin(fmax(hb_div(q - th_r, delta_s), 0.0), 1.0);
    double K, kb = ksoil;
    if (!vg && coef >= 0.0) {
        const double mean = (kmean == 0.0) ? 1.0e-7 : kmean;
        const double var = par[HB_P_SIGMA] * (1.0 - s_eff);
        const double l1 = hb_log(hb_div(var, mean * mean) + 1.0);
        kb = mean * hb_exp(-0.5 * l1 + hb_sqrt(l1) * (coef * e));
    }
    if (vg || (lam != 1.0 && lam != 2.0)) K = closure_rare_K(par, s_eff, base, kb, vg);
    else K = ((lam == 1.0) ? s_eff : s_eff * s_eff) * kb;
    K = fmin(K, kb);
    double C = (m * n) * alpha * delta_s * hb_div(s_eff, base) * hb_div(xn, ax);
    if (C < par[HB_P_EPS] || !isfinite(C)) C = par[HB_P_EPS];
    r.theta = q; r.K = K; r.C = C; r.kb = kb;
    return r;
}

// Fast mode only (HB_FAST_F32, never the default): the same map in single precision with the
// hardware exp2 / log2, for the evaluations inside the Newton iteration.  1 - Se and the noise
// term's log(1 + x) use expm1 / log1p so that nothing cancels at float precision near saturation.
// Saturated cells and the column-uniform rare cases are left to the double-precision code
// (returns false).
__device__ __forceinline__ bool closure_f32(const double *par, double psi_d, double phi_d, double kmean_d,
                                            double coef_d, double e_d, Clos *out)
{
    const bool vg = par[HB_P_MODEL] != 0.0;
    const double lam_d = par[HB_P_LAMBDA];
    if (psi_d >= par[HB_P_PSI_SAT] || vg || (lam_d != 1.0 && lam_d != 2.0)) return false;
    const float n = (float)par[HB_P_N], m = (float)par[HB_P_M], alpha = (float)par[HB_P_ALPHA];
    const float th_r = (float)par[HB_P_THETA_RES];
    const float delta_s = (float)phi_d - th_r;
    const float ax = alpha * fabsf((float)psi_d);
    const float xn = __expf(n * __logf(ax));
    const float base = 1.0f + xn;
    const float t = -m * log1pf(xn);
    const float X = __expf(t);
    const float one_minus_X = -expm1f(t);
    const float s_eff = fminf(fmaxf(X, 0.0f), 1.0f);
    float kb = (float)par[HB_P_KSAT_SOIL];
    if (coef_d >= 0.0) {
        const float mean = (kmean_d == 0.0) ? 1.0e-7f : (float)kmean_d;
        const float var = (float)par[HB_P_SIGMA] * fminf(fmaxf(one_minus_X, 0.0f), 1.0f);
        const float l1 = log1pf(var / (mean * mean));
        kb = mean * __expf(-0.5f * l1 + sqrtf(l1) * ((float)coef_d * (float)e_d));
    }
    float K = ((lam_d == 1.0) ? s_eff : s_eff * s_eff) * kb;
    K = fminf(K, kb);
    float C = (m * n) * alpha * delta_s * (s_eff / base) * (xn / ax);
    const float eps = (float)par[HB_P_EPS];
    if (C < eps || !isfinite(C)) C = eps;
    out->theta = (double)(th_r + delta_s * X); out->K = (double)K; out->C = (double)C; out->kb = (double)kb;
    return true;
}

// Out-of-line copy for the diagnostics on the nodes (cold).
__device__ __noinline__ void closure_cold(const double *par, double psi, double phi, double kmean,
                                          double coef, double e, Clos *out)
{
    *out = closure(par, psi, phi, kmean, coef, e);
}

// Arguments of the current window (args_i of simulation.py:591-596 and the t_span of
// RichardsPDE.solve) in the small shared area.  All threads must call it; it synchronises.
enum { WA_PRECIP = 0, WA_ATM, WA_T0, WA_TBOUND, WA_WTD, WA_MODE };
template <int NV>
__device__ inline void set_window_args(const DevCtx &c, const Args &a, double t0, double t1)
{
    __syncthreads();                       // readers of the previous window are done
    if (threadIdx.x == 0) {
        double *wa = HB_XA(c, NV, X_ARGS);
        wa[WA_PRECIP] = a.precip; wa[WA_ATM] = a.atm; wa[WA_T0] = t0; wa[WA_TBOUND] = t1;
        wa[WA_WTD] = (double)a.wtd; wa[WA_MODE] = (double)a.mode;
    }
    __syncthreads();
}

// Pointers to the static per-cell inputs of the right-hand side: shared memory
// in the FULL layout, global memory (L1/L2) otherwise.
struct Statics {
    const double *phi, *km, *cf;      // (grid metrics dx, zx, xz: read through DevCtx, shared by all columns)
};

template <bool FULL>
__device__ __forceinline__ Statics statics_of(const DevCtx &c, int col)
{
    Statics s;
    if (FULL) {
        s.phi = HB_VEC(c, V_PHI); s.km = HB_VEC(c, V_KM); s.cf = HB_VEC(c, V_CF);
    } else {
        const double *st = c.stat + (size_t)col * HB_NSTATIC * c.nx;
        s.phi = st + (size_t)HB_S_PHI_MID * c.nx; s.km = st + (size_t)HB_S_KMEAN_MID * c.nx;
        s.cf = st + (size_t)HB_S_COEF_MID * c.nx;
    }
    return s;
}

// Stage one column: scalars, and in the FULL layout the static profiles and
// grid metrics.  Caller synchronises afterwards.
template <int NT, bool FULL>
__device__ inline void bind_column(const DevCtx &c, int col)
{
    constexpr int NV = FULL ? NVEC_FULL : NVEC_BASE;
    for (int i = threadIdx.x; i < HB_NPARAM; i += NT)
        HB_XA(c, NV, X_PAR)[i] = c.par[(size_t)col * HB_NPARAM + i];
    if (threadIdx.x < 8) reinterpret_cast<long long *>(HB_XA(c, NV, X_CNT))[threadIdx.x] = 0;
#if HB_PHASE_TIMING
    if (threadIdx.x < 16) reinterpret_cast<long long *>(HB_XA(c, NV, X_TIME))[threadIdx.x] = (threadIdx.x == 15) ? clock64() : 0;
#endif
    // identity padding rows of the tridiagonal system (see tri_solve_fixed)
    for (int i = c.nx + threadIdx.x; i < HB_NXP(c); i += NT) {
        HB_VEC(c, V_JL)[i] = 0.0; HB_VEC(c, V_JD)[i] = 0.0; HB_VEC(c, V_JU)[i] = 0.0;
        HB_VEC(c, V_B)[i] = 0.0;
    }
    if (FULL) {
        const int nx = c.nx;
        const double *st = c.stat + (size_t)col * HB_NSTATIC * nx;
        for (int i = threadIdx.x; i < nx; i += NT) {
            HB_VEC(c, V_PHI)[i] = st[(size_t)HB_S_PHI_MID * nx + i];
            HB_VEC(c, V_KM)[i] = st[(size_t)HB_S_KMEAN_MID * nx + i];
            HB_VEC(c, V_CF)[i] = st[(size_t)HB_S_COEF_MID * nx + i];
        }
    }
}

// Hydraulic redistribution term of the flux, night rows only
// (richards_pde.py:234-254): f += 0.5 * c_hr * (dpsi/dz * dz) in the root zone.
__device__ __noinline__ double hydraulic_lift(const DevCtx &c, const double *par, int col, int k,
                                              double ym, double dy, double flux)
{
    const double rootk = c.stat[((size_t)col * HB_NSTATIC + HB_S_ROOT_MID) * c.nx + k];
    if (!(rootk > 0.0)) return flux;
    const double one = 1.0 - par[HB_P_IPSI50] * ym;
    const double c_hr = (1800 * par[HB_P_LAI]) * (one * one) * rootk;
    return flux + 0.5 * c_hr * (dy * par[HB_P_DZ]);
}

// alpha_01 * alpha_02 of TreeRoots.efficiency for one cell (tree_roots.py:234-265),
// before the "all cells flooded" damping; a2 is returned separately.
__device__ __forceinline__ double root_alphas(double th, double phi, double fcap, double wilt,
                                              double cum_over_total, double *a2_out)
{
    double d01 = phi - wilt;
    if (d01 == 0.0) d01 = 1.0;
    const double a1 = fmax(th / d01, cum_over_total);
    double a2 = 0.0;
    if (wilt < th && th <= fcap) {
        double d02 = fcap - wilt;
        if (d02 == 0.0) d02 = 1.0;
        a2 = (th - fcap) / d02;
    }
    if (th > fcap) a2 = 1.0;
    *a2_out = fmin(fmax(a2, 0.0), 1.0);
    return a1;
}

// Root-water-uptake sink (richards_pde.py:258-302, tree_roots.py:179-292), daylight
// rows only.  The reference evaluates it once for the top half-cell (one midpoint,
// its own normalisation) and once for the interior midpoints 1..nx-2; both are
// reproduced.  theta at the midpoints is in V_TH; the sink goes to V_SK and the
// integrated transpiration of the interior call to MISC[1].  Block reductions and
// a block prefix sum inside: all threads must call it.  Cold code, out of line.
template <int NT, int CPT, int NV>
__device__ __noinline__ void et_sink(const DevCtx &c, int col, double atm, int &parity)
{
    const int nx = c.nx, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *par = HB_XA(c, NV, X_PAR);
    const double dz = par[HB_P_DZ];
    const double *st = c.stat + (size_t)col * HB_NSTATIC * nx;
    const double *root = st + (size_t)HB_S_ROOT_MID * nx, *phi = st + (size_t)HB_S_PHI_MID * nx,
                 *fcap = st + (size_t)HB_S_FCAP_MID * nx, *wilt = st + (size_t)HB_S_WILT_MID * nx;
    const double *th = HB_GROW(c, G_TH);
    double *sk = HB_GROW(c, G_SK), *scan = HB_XA(c, NV, X_SCAN), *misc = HB_XA(c, NV, X_MISC);

    // ---- top half-cell call: midpoint 0 alone (thread 0) -----------------------
    if (tid == 0) {
        double s0 = 0.0;
        if (root[0] > 0.0) {
            double water = (th[0] - wilt[0]) * dz, rho = 0.0;
            if (water > 0.0) {
                double a2;
                // cumsum / total of a single cell: 1, or 0/1 = 0 when theta is zero
                const double ratio = (th[0] * dz == 0.0) ? 0.0 : 1.0;
                const double a1s = root_alphas(th[0], phi[0], fcap[0], wilt[0], ratio, &a2);
                if (a2 == 1.0) a2 *= 0.1;
                rho = fabs(a1s * a2);
                double tot = rho * dz;
                if (tot == 0.0) tot = 1.0;
                rho /= tot;
            } else water = 0.0;
            double x = rho * root[0], totx = x * dz;
            if (totx > 1.0) { x /= totx; totx = x * dz; }
            if (totx > 0.0) s0 = -(fmin(atm, water) / totx * x);
        }
        sk[0] = s0;
    }

    // ---- interior call: midpoints 1 .. nx-2 -------------------------------------
    double thv[CPT], rt[CPT];
    double lw = 0.0, lt = 0.0;
#pragma unroll
    for (int r = 0; r < CPT; ++r) {
        const int k = tid + r * NT;
        thv[r] = 0.0; rt[r] = 0.0;
        if (k >= 1 && k <= nx - 2) {
            rt[r] = root[k];
            if (rt[r] > 0.0) { thv[r] = th[k]; lw += thv[r] - wilt[k]; lt += thv[r]; }
        }
    }
    double water = block_sum<NT, NV>(c, parity, lw) * dz;
    double total = block_sum<NT, NV>(c, parity, lt) * dz;
    double xv[CPT];
    if (water > 0.0) {
        if (total == 0.0) total = 1.0;
        // inclusive prefix sum of theta over the root cells in depth order
        double carry = 0.0, a1v[CPT], a2v[CPT], nall = 0.0;
#pragma unroll
        for (int r = 0; r < CPT; ++r) {
            const int k = tid + r * NT;
            double v = (rt[r] > 0.0) ? thv[r] : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(HB_FULL, v, o);
                if (lane >= o) v += u;
            }
            __syncthreads();
            if (lane == 31) scan[warp] = v;
            __syncthreads();
            double before = carry, round_total = 0.0;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) {
                const double t = scan[w];
                if (w < warp) before += t;
                round_total += t;
            }
            carry += round_total;
            a1v[r] = 0.0; a2v[r] = 1.0;
            if (rt[r] > 0.0) {
                const double cum = (before + v) * dz;
                a1v[r] = root_alphas(thv[r], phi[k], fcap[k], wilt[k], cum / total, &a2v[r]);
                if (a2v[r] != 1.0) nall += 1.0;
            }
        }
        const bool all_one = block_sum<NT, NV>(c, parity, nall) == 0.0;
        double lr = 0.0;
#pragma unroll
        for (int r = 0; r < CPT; ++r) {
            xv[r] = 0.0;
            if (rt[r] > 0.0) { xv[r] = fabs(a1v[r] * (all_one ? a2v[r] * 0.1 : a2v[r])); lr += xv[r]; }
        }
        double tot = block_sum<NT, NV>(c, parity, lr) * dz;
        if (tot == 0.0) tot = 1.0;
#pragma unroll
        for (int r = 0; r < CPT; ++r) xv[r] = (rt[r] > 0.0) ? (xv[r] / tot) * rt[r] : 0.0;
    } else {
        water = 0.0;
#pragma unroll
        for (int r = 0; r < CPT; ++r) xv[r] = 0.0;
    }
    double lx = 0.0;
#pragma unroll
    for (int r = 0; r < CPT; ++r) lx += xv[r];
    double totx = block_sum<NT, NV>(c, parity, lx) * dz;
    if (totx > 1.0) {
        lx = 0.0;
#pragma unroll
        for (int r = 0; r < CPT; ++r) { xv[r] /= totx; lx += xv[r]; }
        totx = block_sum<NT, NV>(c, parity, lx) * dz;
    }
    double lu = 0.0;
    const double tr_pot = (totx > 0.0) ? fmin(atm, water) / totx : 0.0;
#pragma unroll
    for (int r = 0; r < CPT; ++r) {
        const int k = tid + r * NT;
        if (k >= 1 && k < nx) {
            const double up = (totx > 0.0 && rt[r] > 0.0) ? tr_pot * xv[r] : 0.0;
            sk[k] = -up;
            lu += up;
        }
    }
    const double transp = block_sum<NT, NV>(c, parity, lu) * dz;
    if (tid == 0) misc[1] = transp;
    __syncthreads();
}

// Surface evaporation of bc_fun (richards_pde.py:455-468), daylight only: it
// acts when np.all(theta_left > theta_res) over the profile the porosity call
// broadcasts to (quirk Q2).  Cold: daily forcing never reaches it.
__device__ __noinline__ double evaporation(const double *par, double y0)
{
    const double th_r = par[HB_P_THETA_RES];
    const double ax = par[HB_P_ALPHA] * fabs(y0);
    const double X = (par[HB_P_N] == 2.0) ? 1.0 / sqrt(1.0 + ax * ax)
                                           : pow(1.0 + pow(ax, par[HB_P_N]), -par[HB_P_M]);
    const bool sat0 = y0 >= par[HB_P_PSI_SAT];
    const double dsm = sat0 ? par[HB_P_DSMIN_EXCL0] : par[HB_P_DSMIN_ALL];
    if ((th_r + dsm * X > th_r) && (!sat0 || par[HB_P_PHI_TOP] > th_r)) return par[HB_P_SURF_EVAP];
    return 0.0;
}

// dpsi/dt at the state held in vector V_Y (must be visible: caller
// synchronises before the call).  The result is written to vector `out`
// (each thread writes, and may afterwards read, its own cells).  Contains one
// __syncthreads(); writes V_CM, V_FM and the small slots WMAX, MISC, LATW.
template <int NT, int CPT, int NV, bool FAST = false>
__device__ __forceinline__ void rhs_eval(const DevCtx &c, const Statics &S, int col, int out, int &parity,
                                         bool f32 = false)
{
    // the row's arguments (precipitation, demand, water-table observation, mode) are read from
    // shared memory where they are needed instead of living in registers across the integrator loop
    const double *wa = HB_XA(c, NV, X_ARGS);
    const int mode = (int)wa[WA_MODE];
    const int nx = c.nx, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *par = HB_XA(c, NV, X_PAR);
    const double *y = HB_VEC(c, V_Y), *nr = HB_VEC(c, V_NR);
    double *cm = HB_VEC(c, V_CM), *fm = HB_VEC(c, V_FM), *thm = HB_GROW(c, G_TH);
    const double psi_sat = par[HB_P_PSI_SAT];
    const int flags = (int)par[HB_P_FLAGS];
    const bool et_on = (mode == 1) && (flags & HB_FLAG_ET);        // daylight, not SPINUP
    const bool hl_on = (mode == 0) && (flags & HB_FLAG_HLIFT);     // night, not SPINUP
    int deepest = -1;                    // deepest unsaturated interior midpoint of this warp

    // ---- phase A: closures and fluxes at the midpoints ----------------------
    // The top-boundary closure (bc_fun: (y[0], z = x[0])) rides along as one
    // extra cell of the last thread in the last round, where that thread has no
    // midpoint of its own (nx <= NT*CPT), so the closure code exists only once.
#pragma unroll 1
    for (int r = 0; r < CPT * HB_DUP(c, 4); ++r) {
        const int k = tid + (r % CPT) * NT;
        const bool bnd = (r % CPT == CPT - 1) && (tid == NT - 1);
        bool unsat_interior = false;
        if (k < nx - 1 || bnd) {
            double ym, dy = 0.0, e, phi, km, cf;
            if (bnd) {
                ym = y[0]; e = nr[0];
                phi = par[HB_P_PHI_TOP]; km = par[HB_P_KMEAN_TOP]; cf = par[HB_P_COEF_TOP];
            } else {
                const double y0 = y[k], y1 = y[k + 1];
                ym = 0.5 * (y0 + y1);
                dy = div_hot<FAST>(y1 - y0, c.dx[k]);
                // noise index is local to the array handed to the closure (quirk Q1)
                e = nr[k == 0 ? 0 : k - 1];
                phi = S.phi[k]; km = S.km[k]; cf = S.cf[k];
            }
            Clos cl;
            bool have = false;
            if (FAST && f32) have = closure_f32(par, ym, phi, km, cf, e, &cl);      // (fast mode, Newton evaluations only)
            if (!have) cl = closure(par, ym, phi, km, cf, e);
            if (!bnd) {
                cm[k] = cl.C;
                double fl = cl.K * (dy - 1.0);
                if (hl_on) fl = hydraulic_lift(c, par, col, k, ym, dy, fl);
                fm[k] = fl;
                if (et_on) thm[k] = cl.theta;
                unsat_interior = (k >= 1) && !(ym >= psi_sat);
            } else {
                const double q_inf_max = fmin(2.0 * (phi - cl.theta) * par[HB_P_DZ], cl.kb);
                double p_left = 0.0;
                if (ym < psi_sat) p_left = fmin((1.0 - par[HB_P_INTERCEPT]) * fabs(wa[WA_PRECIP]), q_inf_max);
                if (mode == 1) p_left -= evaporation(par, ym);
                HB_XA(c, NV, X_MISC)[0] = p_left;
            }
        }
        const unsigned mask = __ballot_sync(HB_FULL, unsat_interior);
        if (mask) deepest = max(deepest, (warp << 5) + (r % CPT) * NT + (31 - __clz(mask)));
    }
    int *wmax = reinterpret_cast<int *>(HB_XA(c, NV, X_WMAX));
    if (lane == 0) wmax[warp] = deepest;
    __syncthreads();
    HB_TICK(c, NV, 1);                   // closure phase incl. its barrier
    if (et_on) et_sink<NT, CPT, NV>(c, col, wa[WA_ATM], parity);
    else if (tid == 0) HB_XA(c, NV, X_MISC)[1] = 0.0;
    const double *sk = HB_GROW(c, G_SK);

    // ---- phase B: lateral-flow sink (monitoring mode) and assembly ----------
    int lo = 1, hi = 0;                  // active sink range in midpoint index: [lo, hi)
    if (flags & HB_FLAG_LF) {
        int kmax = wmax[0];
#pragma unroll
        for (int w = 1; w < NT / 32; ++w) kmax = max(kmax, wmax[w]);
        const int dim_d = nx - 2;
        const int wtd_obs = min((int)wa[WA_WTD], dim_d - 1);
        const int wtd_est = (kmax < 0) ? 0 : min(kmax, dim_d - 1);
        if (wtd_est < wtd_obs) { lo = wtd_est + 1; hi = wtd_obs + 1; }
    }
    double lat = 0.0;
    double *fo = HB_VEC(c, out);
#pragma unroll
    for (int r = 0; r < CPT * HB_DUP(c, 8); ++r) {
        const int i = tid + (r % CPT) * NT;
        if (HB_DUP_PROBE && r == CPT) lat = 0.0;
        if (i < nx) {
            const double yi = y[i];
            // sink at midpoints i and i-1: root uptake (<= 0, daylight) and, inside
            // the lateral-flow range, min(alpha_lat * psi_mid, sink)
            double s_i = (et_on && i < nx - 1) ? sk[i] : 0.0, s_m = (et_on && i >= 1) ? sk[i - 1] : 0.0;
            if (i >= lo && i < hi) {          // own midpoint i
                s_i = fmin(-2.5e-4 * (0.5 * (yi + y[i + 1])), s_i);
                lat += fabs(s_i);
            }
            if (i - 1 >= lo && i - 1 < hi)    // midpoint i-1
                s_m = fmin(-2.5e-4 * (0.5 * (y[i - 1] + yi)), s_m);
            // the three node types of richards_pde.py:109-156:
            //   top     (p_left + (f_0 + zx*s_0)) / (zx*C_0)
            //   inner   (f_i - f_{i-1} + (zx*s_i + xz*s_{i-1})) / (zx*C_i + xz*C_{i-1})
            //   bottom  (0 + (f_{n-2} - xz*s_{n-2})) / (-xz*C_{n-2})
            double num, den;
            if (i == 0) {
                const double zx = c.zx[0];
                den = zx * cm[0];
                num = HB_XA(c, NV, X_MISC)[0] + (fm[0] + zx * s_i);
            } else if (i < nx - 1) {
                const double zx = c.zx[i], xz = c.xz[i];
                den = zx * cm[i] + xz * cm[i - 1];
                num = fm[i] - fm[i - 1] + (zx * s_i + xz * s_m);
            } else {
                const double xz = c.xz[i];
                den = -xz * cm[i - 1];
                num = 0.0 + (fm[i - 1] - xz * s_m);
            }
            if (den == 0.0) den = 1.0;
            fo[i] = div_hot<FAST>(num, den);
        }
    }
    lat = warp_sum(lat);
    if (lane == 0) HB_XA(c, NV, X_LATW)[warp] = lat;
}

// Lateral flow of the last RHS evaluation (arg_out, quirk Q6); thread 0 only.
template <int NT, int NV>
__device__ inline double lateral_flow(const DevCtx &c)
{
    double lat = 0.0;
    for (int w = 0; w < NT / 32; ++w) lat += HB_XA(c, NV, X_LATW)[w];
    return lat * HB_XA(c, NV, X_PAR)[HB_P_DZ];
}

// ---------------------------------------------------------------------------
// Tridiagonal partition solver, executed by warp 0 only.
// Matrix rows: a_i = -c*jl[i], b_i = 1 - c*jd[i], c_i = -c*ju[i].
// Lane p owns rows [p*m, min((p+1)*m, nx)); its last row is an interface
// unknown, the others are eliminated locally (Thomas, no pivoting) and their
// coupling to the two neighbouring interfaces is kept as spike vectors.  The 32
// interface unknowns form a tridiagonal system solved by PCR with shuffles.
// Loops are deliberately rolled: measured on B200, a fully unrolled,
// register-resident variant (3x more SASS) made the whole kernel 20 % slower --
// the instruction footprint of the hot loop matters more than the length of the
// dependent chain.  (This solver is ours, not
// a restatement of reference arithmetic, so FMA contraction and a
// Newton-refined reciprocal are used.)
// ---------------------------------------------------------------------------
__device__ __noinline__ double fast_rcp(double x)
{
    const double ax = fabs(x);
    if (!(ax > 1e-30 && ax < 1e30)) return 1.0 / x;
    double r = (double)__frcp_rn((float)x);
    double e = fma(-x, r, 1.0); r = fma(r, e, r);
    e = fma(-x, r, 1.0); r = fma(r, e, r);
    e = fma(-x, r, 1.0); r = fma(r, e, r);
    return r;
}

// Inline reciprocal for the unrolled factorisation: hardware seed (about 20
// bits, exponent range of a double) and two Newton steps.
__device__ __forceinline__ double rcp_nr(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0); r = fma(r, e, r);
    e = fma(-x, r, 1.0); r = fma(r, e, r);
    return r;
}

template <int NV>
__device__ __noinline__ void tri_factor(const DevCtx &c, double cc)
{
    const int lane = threadIdx.x & 31;
    const int nx = c.nx;
    const double *jl = HB_VEC(c, V_JL), *jd = HB_VEC(c, V_JD), *ju = HB_VEC(c, V_JU);
    double *tw = HB_VEC(c, V_TW), *tip = HB_VEC(c, V_TIP), *tvl = HB_VEC(c, V_TVL),
           *tvr = HB_VEC(c, V_TVR), *pcr = HB_XA(c, NV, X_PCR);
    const int m = (nx + 31) >> 5;
    const int s = lane * m;
    const int e = min(s + m, nx) - 1;           // inclusive; e < s: lane has no rows
    const int cnt = e - s + 1;
    double vl_last = 0.0, vr_last = 0.0, vl_first = 0.0, vr_first = 0.0;
    if (cnt >= 2) {
        // interior rows s..e-1
        double ip = fast_rcp(1.0 - cc * jd[s]);
        tw[s] = 0.0; tip[s] = ip;
        double g = -cc * jl[s];                 // rhs of the left spike, forward swept
        tvl[s] = g;
#pragma unroll 1
        for (int r = s + 1; r < e; ++r) {
            const double w = (-cc * jl[r]) * ip;
            ip = fast_rcp(fma(-w, -cc * ju[r - 1], 1.0 - cc * jd[r]));
            tw[r] = w; tip[r] = ip;
            g = -w * g;
            tvl[r] = g;
        }
        // back substitution for both spikes
        double zl = tvl[e - 1] * ip;
        double zr = (-cc * ju[e - 1]) * ip;
        tvl[e - 1] = zl; tvr[e - 1] = zr;
        vl_last = zl; vr_last = zr;
#pragma unroll 1
        for (int r = e - 2; r >= s; --r) {
            const double ipr = tip[r];
            const double cu = (-cc * ju[r]) * ipr;
            zl = fma(-cu, zl, tvl[r] * ipr);
            zr = -cu * zr;
            tvl[r] = zl; tvr[r] = zr;
        }
        vl_first = zl; vr_first = zr;
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
#pragma unroll 1
    for (int st = 0; st < 5; ++st) {
        const int d = 1 << st;
        double Am = __shfl_up_sync(HB_FULL, A, d), Bm = __shfl_up_sync(HB_FULL, B, d),
               Cm = __shfl_up_sync(HB_FULL, C, d);
        double Ap = __shfl_down_sync(HB_FULL, A, d), Bp = __shfl_down_sync(HB_FULL, B, d),
               Cp = __shfl_down_sync(HB_FULL, C, d);
        if (lane < d) { Am = 0.0; Bm = 1.0; Cm = 0.0; }
        if (lane + d > 31) { Ap = 0.0; Bp = 1.0; Cp = 0.0; }
        const double k1 = A * fast_rcp(Bm), k2 = C * fast_rcp(Bp);
        pcr[(2 * st) * 32 + lane] = k1;
        pcr[(2 * st + 1) * 32 + lane] = k2;
        B = B - k1 * Cm - k2 * Ap;
        A = -k1 * Am;
        C = -k2 * Cp;
    }
    pcr[10 * 32 + lane] = fast_rcp(B);
    __syncwarp();
}

// Fully unrolled variants for a compile-time chunk length M.  The vectors are
// padded to 32*M rows with identity rows (J = 0, rhs = 0), so every lane owns
// exactly M rows and the code has no predicates or loops: the serial part of
// the Newton iteration then is one dependent FMA per row.
template <int NV, int M>
__device__ __forceinline__ void tri_factor_fixed(const DevCtx &c, double cc)
{
    const int lane = threadIdx.x & 31;
    const int s = lane * M;
    const double *jl = HB_VEC(c, V_JL) + s, *jd = HB_VEC(c, V_JD) + s, *ju = HB_VEC(c, V_JU) + s;
    double *tw = HB_VEC(c, V_TW) + s, *tip = HB_VEC(c, V_TIP) + s, *tvl = HB_VEC(c, V_TVL) + s,
           *tvr = HB_VEC(c, V_TVR) + s, *pcr = HB_XA(c, NV, X_PCR);
    double av[M], bv[M], cv[M];
#pragma unroll
    for (int r = 0; r < M; ++r) { av[r] = -cc * jl[r]; bv[r] = 1.0 - cc * jd[r]; cv[r] = -cc * ju[r]; }
    double ip = rcp_nr(bv[0]);
    double g = av[0];
    bv[0] = ip; tw[0] = 0.0;
#pragma unroll
    for (int r = 1; r < M - 1; ++r) {
        const double w = av[r] * ip;
        ip = rcp_nr(fma(-w, cv[r - 1], bv[r]));
        tw[r] = w;
        g = -w * g;
        bv[r] = ip; av[r] = g;
    }
    double zl = av[M - 2] * bv[M - 2], zr = cv[M - 2] * bv[M - 2];
    const double vl_last = zl, vr_last = zr;
    tip[M - 2] = bv[M - 2]; tvl[M - 2] = zl; tvr[M - 2] = zr;
#pragma unroll
    for (int r = M - 3; r >= 0; --r) {
        const double cu = cv[r] * bv[r];
        zl = fma(-cu, zl, av[r] * bv[r]);
        zr = -cu * zr;
        tip[r] = bv[r]; tvl[r] = zl; tvr[r] = zr;
    }
    const double ae = av[M - 1], be = bv[M - 1];
    double ce = cv[M - 1];
    const double vlf_n = __shfl_down_sync(HB_FULL, zl, 1), vrf_n = __shfl_down_sync(HB_FULL, zr, 1);
    double A = -ae * vl_last, B = be - ae * vr_last, C;
    if (lane == 31) { C = 0.0; ce = 0.0; }
    else { B -= ce * vlf_n; C = -ce * vrf_n; }
    if (lane == 0) A = 0.0;
    pcr[11 * 32 + lane] = ae;
    pcr[12 * 32 + lane] = ce;
#pragma unroll 1
    for (int st = 0; st < 5; ++st) {
        const int d = 1 << st;
        double Am = __shfl_up_sync(HB_FULL, A, d), Bm = __shfl_up_sync(HB_FULL, B, d),
               Cm = __shfl_up_sync(HB_FULL, C, d);
        double Ap = __shfl_down_sync(HB_FULL, A, d), Bp = __shfl_down_sync(HB_FULL, B, d),
               Cp = __shfl_down_sync(HB_FULL, C, d);
        if (lane < d) { Am = 0.0; Bm = 1.0; Cm = 0.0; }
        if (lane + d > 31) { Ap = 0.0; Bp = 1.0; Cp = 0.0; }
        const double k1 = A * rcp_nr(Bm), k2 = C * rcp_nr(Bp);
        pcr[(2 * st) * 32 + lane] = k1;
        pcr[(2 * st + 1) * 32 + lane] = k2;
        B = B - k1 * Cm - k2 * Ap;
        A = -k1 * Am;
        C = -k2 * Cp;
    }
    pcr[10 * 32 + lane] = rcp_nr(B);
    __syncwarp();
}

// Warp 0's part of the solve: local sweeps and the reduced system.  On exit
// V_B holds the locally eliminated values z of the interior rows and the 32
// interface unknowns are in the exchange slots; every thread then finishes its
// own rows with tri_x_fixed (two FMAs per row), so the serial section carries
// no back-substitution.
template <int NV, int M>
__device__ __forceinline__ void tri_solve_fixed(const DevCtx &c, double cc)
{
    const int lane = threadIdx.x & 31;
    const int s = lane * M;
    const double *tw = HB_VEC(c, V_TW) + s, *tip = HB_VEC(c, V_TIP) + s, *ju = HB_VEC(c, V_JU) + s,
                 *pcr = HB_XA(c, NV, X_PCR);
    double *b = HB_VEC(c, V_B) + s;
    double gv[M], xv[M];
#pragma unroll
    for (int r = 0; r < M; ++r) gv[r] = b[r];
#pragma unroll
    for (int r = 1; r < M - 1; ++r) xv[r] = tw[r];
    HB_TICK(c, NV, 8);
#pragma unroll
    for (int r = 1; r < M - 1; ++r) gv[r] = fma(-xv[r], gv[r - 1], gv[r]);      // forward sweep
    HB_TICK(c, NV, 9);
#pragma unroll
    for (int r = 0; r < M - 1; ++r) { const double ip = tip[r]; xv[r] = (-cc * ju[r]) * ip; gv[r] = gv[r] * ip; }
#pragma unroll
    for (int r = M - 3; r >= 0; --r) gv[r] = fma(-xv[r], gv[r + 1], gv[r]);     // backward sweep
#pragma unroll
    for (int r = 0; r < M - 1; ++r) b[r] = gv[r];
    HB_TICK(c, NV, 10);
    double rho = gv[M - 1];
    // neighbour exchange through shared memory (measured on B200: cheaper than
    // the shuffle sequence in this single-warp section)
    double *ex = HB_XA(c, NV, X_PCR) + 13 * 32;
    ex[lane] = gv[0];
    __syncwarp();
    const double zf_n = (lane < 31) ? ex[lane + 1] : 0.0;
    __syncwarp();
    rho = fma(-pcr[11 * 32 + lane], gv[M - 2], rho);
    rho = fma(-pcr[12 * 32 + lane], zf_n, rho);          // pcr[12] is 0 on lane 31
#pragma unroll
    for (int st = 0; st < 5; ++st) {
        const int d = 1 << st;
        ex[lane] = rho;
        __syncwarp();
        const double rm = (lane >= d) ? ex[lane - d] : 0.0;
        const double rp = (lane + d <= 31) ? ex[lane + d] : 0.0;
        __syncwarp();
        rho = fma(-pcr[(2 * st + 1) * 32 + lane], rp, fma(-pcr[(2 * st) * 32 + lane], rm, rho));
    }
    HB_TICK(c, NV, 11);
    ex[lane] = rho * pcr[10 * 32 + lane];
    __syncwarp();
    HB_TICK(c, NV, 12);
}

// Row i of the solution after tri_solve_fixed (any thread, after a barrier).
template <int NV, int M>
__device__ __forceinline__ double tri_x_fixed(const DevCtx &c, int i)
{
    const int l = i / M, r = i - l * M;
    const double *xs = HB_XA(c, NV, X_PCR) + 13 * 32;
    const double xi = xs[l];
    if (r == M - 1) return xi;
    const double x_left = (l == 0) ? 0.0 : xs[l - 1];
    return fma(-xi, HB_VEC(c, V_TVR)[i], fma(-x_left, HB_VEC(c, V_TVL)[i], HB_VEC(c, V_B)[i]));
}

template <int NV>
__device__ __noinline__ void tri_factor13(const DevCtx &c, double cc) { tri_factor_fixed<NV, 13>(c, cc); }
template <int NV>
__device__ __noinline__ void tri_solve13(const DevCtx &c, double cc) { tri_solve_fixed<NV, 13>(c, cc); }

// Solves in place: V_B holds the rhs on entry, the solution on exit.
template <int NV>
__device__ __noinline__ void tri_solve(const DevCtx &c, double cc)
{
    const int lane = threadIdx.x & 31;
    const int nx = c.nx;
    const double *ju = HB_VEC(c, V_JU), *tw = HB_VEC(c, V_TW), *tip = HB_VEC(c, V_TIP),
                 *tvl = HB_VEC(c, V_TVL), *tvr = HB_VEC(c, V_TVR), *pcr = HB_XA(c, NV, X_PCR);
    double *b = HB_VEC(c, V_B);
    const int m = (nx + 31) >> 5;
    const int s = lane * m;
    const int e = min(s + m, nx) - 1;
    const int cnt = e - s + 1;
    double z_first = 0.0, z_last = 0.0;
    if (cnt >= 2) {
        // forward sweep; the operands of the next row are fetched before the
        // dependent FMA of the current one
        double g = b[s];
        int rn = (s + 1 < e) ? s + 1 : s;
        double wn = tw[rn], bn = b[rn];
#pragma unroll 1
        for (int r = s + 1; r < e; ++r) {
            const double w = wn, br = bn;
            rn = (r + 1 < e) ? r + 1 : r;
            wn = tw[rn]; bn = b[rn];
            g = fma(-w, g, br);
            b[r] = g;
        }
        double z = g * tip[e - 1];
        b[e - 1] = z;
        z_last = z;
        // backward sweep
        int rq = (e - 2 >= s) ? e - 2 : s;
        double ipn = tip[rq], jun = ju[rq], bq = b[rq];
#pragma unroll 1
        for (int r = e - 2; r >= s; --r) {
            const double cu = (-cc * jun) * ipn, t0 = bq * ipn;
            rq = (r - 1 >= s) ? r - 1 : s;
            ipn = tip[rq]; jun = ju[rq]; bq = b[rq];
            z = fma(-cu, z, t0);
            b[r] = z;
        }
        z_first = z;
    }
    double rho = (cnt >= 1) ? b[e] : 0.0;
    const double zf_n = __shfl_down_sync(HB_FULL, z_first, 1);
    rho -= pcr[11 * 32 + lane] * z_last;
    rho -= pcr[12 * 32 + lane] * ((lane < 31) ? zf_n : 0.0);
#pragma unroll 1
    for (int st = 0; st < 5; ++st) {
        const int d = 1 << st;
        double rm = __shfl_up_sync(HB_FULL, rho, d);
        double rp = __shfl_down_sync(HB_FULL, rho, d);
        if (lane < d) rm = 0.0;
        if (lane + d > 31) rp = 0.0;
        rho = fma(-pcr[(2 * st + 1) * 32 + lane], rp, fma(-pcr[(2 * st) * 32 + lane], rm, rho));
    }
    const double xi = rho * pcr[10 * 32 + lane];
    double x_left = __shfl_up_sync(HB_FULL, xi, 1);
    if (lane == 0) x_left = 0.0;
    if (cnt >= 1) b[e] = xi;
#pragma unroll 1
    for (int r = s; r < e; ++r)
        b[r] = fma(-xi, tvr[r], fma(-x_left, tvl[r], b[r]));
    __syncwarp();
}

// Dispatch: the unrolled 13-row variant needs the padded 416-row layout (FULL
// layout, 385..416 nodes); everything else takes the generic solver (which also
// does its own back-substitution).
template <int NV, bool FULL>
__device__ __forceinline__ bool tri_is13(const DevCtx &c) { return HB_NXP(c) == 32 * 13; }
template <int NV, bool FULL>
__device__ __forceinline__ void tri_factor_warp0(const DevCtx &c, double cc)
{
    if (HB_NXP(c) == 32 * 13) tri_factor13<NV>(c, cc); else tri_factor<NV>(c, cc);
}
template <int NV, bool FULL>
__device__ __forceinline__ void tri_solve_warp0(const DevCtx &c, double cc)
{
    if (HB_NXP(c) == 32 * 13) tri_solve13<NV>(c, cc); else tri_solve<NV>(c, cc);
}
template <int NV, bool FULL>
__device__ __forceinline__ double tri_x(const DevCtx &c, int i)
{
    if (HB_NXP(c) == 32 * 13) return tri_x_fixed<NV, 13>(c, i);
    return HB_VEC(c, V_B)[i];
}

// ---------------------------------------------------------------------------
// change_D (SCIPY bdf.py:18-33): D[0..order] <- (R U)^T D[0..order].
// ---------------------------------------------------------------------------
template <int NT, int CPT, int Q>
__device__ __forceinline__ void transform_D(double *D, int nxp, int nx, const double *ru)
{
    const int tid = threadIdx.x;
    HB_FOR_CELLS(p) {
        double col[Q], out[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) col[j] = D[j * nxp + p];
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < Q; ++j) acc += ru[j * 6 + i] * col[j];
            out[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < Q; ++i) D[i * nxp + p] = out[i];
    }
}

template <int NT, int CPT, int NV>
__device__ __noinline__ void change_D(const DevCtx &c, int order, double factor)
{
    const int tid = threadIdx.x;
    const int q = order + 1;
    double *ru = HB_XA(c, NV, X_RU), *rm = HB_XA(c, NV, X_RED);
    __syncthreads();                       // previous readers of RU and of the reduction slots are done
    // R[i][k] with `factor` (cumulative products down the rows, bdf.py:18-27), one entry per thread;
    // U = R(factor = 1) does not depend on anything: host-made table k_U (same operations in the
    // same order, hb_create).  Each entry of RU used to recompute its whole row of R and column of U.
    for (int e_ = tid; e_ < q * q; e_ += NT) {     // (NT may be smaller than 36: narrow A/B shapes)
        const int i = e_ / q, k = e_ % q;
        double Rik = 1.0;
        for (int rr = 1; rr <= i; ++rr)
            Rik *= (k == 0) ? 0.0 : ((double)(rr - 1) - factor * (double)k) / (double)rr;
        rm[i * 6 + k] = Rik;
    }
    __syncthreads();
    for (int e_ = tid; e_ < q * q; e_ += NT) {
        const int i = e_ / q, j = e_ % q;
        double acc = 0.0;
        for (int k = 0; k < q; ++k) acc += rm[i * 6 + k] * k_U[k * 6 + j];
        ru[i * 6 + j] = acc;
    }
    __syncthreads();
    const int nxp = HB_NXP(c);
    double *D = HB_GROW(c, G_D);
    switch (q) {                           // D[0..order] <- (R U)^T D[0..order], unrolled per order
    case 2: transform_D<NT, CPT, 2>(D, nxp, c.nx, ru); break;
    case 3: transform_D<NT, CPT, 3>(D, nxp, c.nx, ru); break;
    case 4: transform_D<NT, CPT, 4>(D, nxp, c.nx, ru); break;
    case 5: transform_D<NT, CPT, 5>(D, nxp, c.nx, ru); break;
    default: transform_D<NT, CPT, 6>(D, nxp, c.nx, ru); break;
    }
}

// Column statistics of the grouped forward differences with scipy's
// sparse-argmax semantics (common.py:391-396, sparse/_data.py:249-275): first
// maximum among the stored rows j-1..j+1, row 0 when the column is all zero.
struct ColStat { double dl, dd, du, maxd, scl; };

__device__ __forceinline__ void column_stats(const double *fb, const double *fn, int j, int nx,
                                          ColStat *o)
{
    const int lo = j > 0 ? j - 1 : 0, hi = j < nx - 1 ? j + 1 : nx - 1;
    int best = lo; double bv = fabs(fn[lo] - fb[lo]);
    for (int i = lo + 1; i <= hi; ++i) {
        const double v = fabs(fn[i] - fb[i]);
        if (v > bv) { bv = v; best = i; }
    }
    int mi;
    if (bv > 0.0) { mi = best; o->maxd = bv; }
    else { mi = 0; o->maxd = (lo == 0) ? fabs(fn[0] - fb[0]) : 0.0; }
    o->scl = fmax(fabs(fb[mi]), fabs(fn[mi]));
    o->dl = (j > 0) ? fn[j - 1] - fb[j - 1] : 0.0;
    o->dd = fn[j] - fb[j];
    o->du = (j < nx - 1) ? fn[j + 1] - fb[j + 1] : 0.0;
}

// ---------------------------------------------------------------------------
// One solve_ivp(method='BDF') call from t0 to t_bound, as a state machine
// around a single right-hand-side evaluation.  y0 is read from global memory
// (y_glob); on exit V_Y holds solver.y, the last accepted state
// (ivp.py:659-666).  Returns 0 (finished) or -1 (step size underflow).
//
// Jacobian scratch while a forward-difference Jacobian is being built (the
// tridiagonal factors and the Newton vectors are dead then):
//   V_TW  f at the base point        V_TIP f at the perturbed point
//   V_TVL the base point             V_TVR step h
//   V_JL/V_JD/V_JU un-normalised column differences J[j-1][j], J[j][j], J[j+1][j]
//   V_PSI max |difference|           V_SCL its scale       V_DD "too small" flag
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// num_jac's grouped forward differences (SCIPY common.py:385-443, first pass over the
// groups) in ONE pass instead of one right-hand-side evaluation per group.
//
// scipy evaluates f(y + h_g) for every colour group g and reads, for each column j of the
// group, the rows j-1..j+1 of f(y + h_g) - f(y).  Row i of the right-hand side depends on
// y[i-1..i+1] only (through the midpoints i-1 and i), and same-group columns are at least
// three apart, so those three rows see exactly ONE perturbed entry, y[j] + h[j].  Midpoint k
// therefore takes only three different values over all groups: the base one, the one with
// y[k] perturbed ("L", used by column k) and the one with y[k+1] perturbed ("R", used by
// column k+1).  The pass evaluates L and R for every midpoint -- two closure evaluations
// instead of one per group -- and assembles the three rows of every column from them with
// the same operations rhs_eval uses, so every number is bit-identical to the evaluation per
// group (tests/test_gpu_round2.py::test_fused_jacobian_is_bit_identical).  The one
// non-local input, the estimated water table of the lateral-flow sink (deepest unsaturated
// interior midpoint of the whole perturbed state), is reduced per group.  Rows with the
// root-water-uptake sink (daylight, ET on: global sums over the profile) keep the
// evaluation per group, and so does the retry pass with the 10x steps.
//
// In:  V_TVL base point, V_TVR steps h, V_TW f(base); C, flux and boundary flux of the base
//      evaluation: V_CM / V_FM / MISC[0] when the S_JB state has just made it, else (initial
//      Jacobian) the copies the S_F0 state left in the scratch rows G_TH / G_SK and MISC[2]
//      (free: the root-uptake sink is off in every window that comes here); which of the two
//      is flagged in the parked controller state (X_RU).
// Out: V_JL/V_JD/V_JU, V_PSI, V_SCL, V_DD exactly as the S_JG state leaves them after its last group.
// Scratch (all dead while a Jacobian is built): V_Y / V_B = C / flux of the L variants,
//      V_TIP / V_DD = C / flux of the R variants (V_DD until the statistics overwrite it),
//      V_JL/V_JD/V_JU = the three rows of f(y + h_g) per column before they become differences,
//      X_PCR = per-group water-table reduction, f(y + h_g)[0] per group, perturbed boundary flux.
// ---------------------------------------------------------------------------
#define HB_JAC_MAXG 8
template <int NT, int CPT, bool FULL>
__device__ __noinline__ void jac_fused(const DevCtx &c, int col)
{
    constexpr int NV = FULL ? NVEC_FULL : NVEC_BASE;
    __syncthreads();                                 // the caller's parked controller state is visible
    // base evaluation: the one at the start of the attempt (C, flux, boundary flux saved by the S_F0
    // state) for the initial Jacobian, else the one just made by the S_JB state
    const bool initial = (reinterpret_cast<const int *>(HB_XA(c, NV, X_RU) + 10)[4] & 64) != 0;
    const double *cm0 = initial ? HB_GROW(c, G_TH) : HB_VEC(c, V_CM);
    const double *fm0 = initial ? HB_GROW(c, G_SK) : HB_VEC(c, V_FM);
    const double p_left0 = HB_XA(c, NV, X_MISC)[initial ? 2 : 0];
    double *fr = HB_VEC(c, V_DD);
    const Statics S = statics_of<FULL>(c, col);
    const int nx = c.nx, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, ng = c.ng;
    constexpr int NW = NT / 32;
    const double *wa = HB_XA(c, NV, X_ARGS);
    const int mode = (int)wa[WA_MODE];
    const double *par = HB_XA(c, NV, X_PAR);
    const double psi_sat = par[HB_P_PSI_SAT];
    const int flags = (int)par[HB_P_FLAGS];
    const bool hl_on = (mode == 0) && (flags & HB_FLAG_HLIFT);
    const double *ybv = HB_VEC(c, V_TVL), *hv = HB_VEC(c, V_TVR), *fbv = HB_VEC(c, V_TW), *nr = HB_VEC(c, V_NR);
    double *cl_ = HB_VEC(c, V_Y), *fl_ = HB_VEC(c, V_B), *cr = HB_VEC(c, V_TIP);
    double *jl = HB_VEC(c, V_JL), *jd = HB_VEC(c, V_JD), *ju = HB_VEC(c, V_JU);
    double *vpsi = HB_VEC(c, V_PSI), *vscl = HB_VEC(c, V_SCL), *vd = HB_VEC(c, V_DD);
    int *wmaxg = reinterpret_cast<int *>(HB_XA(c, NV, X_PCR));           // [HB_JAC_MAXG][NW <= 16]
    double *fn0 = HB_XA(c, NV, X_PCR) + 64;                              // f(y + h_g)[0] per group
    double *plp = HB_XA(c, NV, X_PCR) + 72;                              // boundary flux with y[0] perturbed

    __syncthreads();
    if (tid < HB_JAC_MAXG) fn0[tid] = fbv[0];
    // ---- phase 1: the two perturbed evaluations of every midpoint -----------------------
    int deep[HB_JAC_MAXG];
#pragma unroll
    for (int g = 0; g < HB_JAC_MAXG; ++g) deep[g] = -1;
#pragma unroll 1
    for (int r = 0; r < CPT; ++r) {
        const int k = tid + r * NT;
        bool u0 = false, uL = false, uR = false;
        int gk = -1, gk1 = -1;
        if (k < nx - 1) {
            const double y0 = ybv[k], y1 = ybv[k + 1];
            const double y0p = y0 + hv[k], y1p = y1 + hv[k + 1];
            const double e = nr[k == 0 ? 0 : k - 1];
            const double phi = S.phi[k], km = S.km[k], cf = S.cf[k], dxk = c.dx[k];
            {   // L: y[k] perturbed
                const double ym = 0.5 * (y0p + y1), dy = hb_div(y1 - y0p, dxk);
                const Clos cl = closure(par, ym, phi, km, cf, e);
                double fl = cl.K * (dy - 1.0);
                if (hl_on) fl = hydraulic_lift(c, par, col, k, ym, dy, fl);
                cl_[k] = cl.C; fl_[k] = fl;
                uL = (k >= 1) && !(ym >= psi_sat);
            }
            {   // R: y[k+1] perturbed
                const double ym = 0.5 * (y0 + y1p), dy = hb_div(y1p - y0, dxk);
                const Clos cl = closure(par, ym, phi, km, cf, e);
                double fl = cl.K * (dy - 1.0);
                if (hl_on) fl = hydraulic_lift(c, par, col, k, ym, dy, fl);
                cr[k] = cl.C; fr[k] = fl;
                uR = (k >= 1) && !(ym >= psi_sat);
            }
            u0 = (k >= 1) && !(0.5 * (y0 + y1) >= psi_sat);
            gk = c.grp[k]; gk1 = c.grp[k + 1];
        }
        // deepest unsaturated interior midpoint of the state perturbed by group g
#pragma unroll
        for (int g = 0; g < HB_JAC_MAXG; ++g) {
            if (g < ng) {
                const bool u = (gk == g) ? uL : (gk1 == g) ? uR : u0;
                const unsigned mask = __ballot_sync(HB_FULL, u);
                if (mask) deep[g] = max(deep[g], (warp << 5) + r * NT + (31 - __clz(mask)));
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int g = 0; g < HB_JAC_MAXG; ++g) wmaxg[g * 16 + warp] = deep[g];
    }
    if (tid == NT - 1) {
        // boundary closure with y[0] perturbed (bc_fun, as in rhs_eval)
        const double ym = ybv[0] + hv[0];
        const double phi = par[HB_P_PHI_TOP];
        const Clos cl = closure(par, ym, phi, par[HB_P_KMEAN_TOP], par[HB_P_COEF_TOP], nr[0]);
        const double q_inf_max = fmin(2.0 * (phi - cl.theta) * par[HB_P_DZ], cl.kb);
        double p_left = 0.0;
        if (ym < psi_sat) p_left = fmin((1.0 - par[HB_P_INTERCEPT]) * fabs(wa[WA_PRECIP]), q_inf_max);
        if (mode == 1) p_left -= evaporation(par, ym);
        plp[0] = p_left;
    }
    __syncthreads();

    // ---- phase 2a: rows j-1, j, j+1 of f(y + h_g) for every column j (g = its group) ------
#pragma unroll 1
    for (int r = 0; r < CPT; ++r) {
        const int j = tid + r * NT;
        if (j >= nx) continue;
        const int g = c.grp[j];
        int lo = 1, hi = 0;                  // active range of the lateral-flow sink: [lo, hi)
        if (flags & HB_FLAG_LF) {
            int kmax = wmaxg[g * 16];
#pragma unroll
            for (int w = 1; w < NW; ++w) kmax = max(kmax, wmaxg[g * 16 + w]);
            const int dim_d = nx - 2;
            const int wtd_obs = min((int)wa[WA_WTD], dim_d - 1);
            const int wtd_est = (kmax < 0) ? 0 : min(kmax, dim_d - 1);
            if (wtd_est < wtd_obs) { lo = wtd_est + 1; hi = wtd_obs + 1; }
        }
        const double yjp = ybv[j] + hv[j];
        double f3[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int i = j - 1 + q;
            if (i < 0 || i >= nx) continue;
            // state, C and flux as row i sees them: y[j], midpoint j (L) and midpoint j-1 (R) perturbed
            const double yi = (q == 1) ? yjp : ybv[i];
            double s_i = 0.0, s_m = 0.0;
            if (i >= lo && i < hi) {
                const double ynext = (q == 0) ? yjp : ybv[i + 1];
                s_i = fmin(-2.5e-4 * (0.5 * (yi + ynext)), s_i);
            }
            if (i - 1 >= lo && i - 1 < hi) {
                const double yprev = (q == 2) ? yjp : ybv[i - 1];
                s_m = fmin(-2.5e-4 * (0.5 * (yprev + yi)), s_m);
            }
            // midpoint i:   q == 0 -> j-1 (R), q == 1 -> j (L), q == 2 -> j+1 (base)
            // midpoint i-1: q == 0 -> j-2 (base), q == 1 -> j-1 (R), q == 2 -> j (L)
            double num, den;
            if (i == 0) {
                const double zx = c.zx[0];
                const double cmi = (q == 0) ? cr[0] : cl_[0], fmi = (q == 0) ? fr[0] : fl_[0];
                den = zx * cmi;
                num = ((j == 0) ? plp[0] : p_left0) + (fmi + zx * s_i);
            } else if (i < nx - 1) {
                const double zx = c.zx[i], xz = c.xz[i];
                const double cmi = (q == 0) ? cr[i] : (q == 1) ? cl_[i] : cm0[i];
                const double fmi = (q == 0) ? fr[i] : (q == 1) ? fl_[i] : fm0[i];
                const double cmm = (q == 0) ? cm0[i - 1] : (q == 1) ? cr[i - 1] : cl_[i - 1];
                const double fmm = (q == 0) ? fm0[i - 1] : (q == 1) ? fr[i - 1] : fl_[i - 1];
                den = zx * cmi + xz * cmm;
                num = fmi - fmm + (zx * s_i + xz * s_m);
            } else {
                const double xz = c.xz[i];
                const double cmm = (q == 1) ? cr[i - 1] : cl_[i - 1];
                const double fmm = (q == 1) ? fr[i - 1] : fl_[i - 1];
                den = -xz * cmm;
                num = 0.0 + (fmm - xz * s_m);
            }
            if (den == 0.0) den = 1.0;
            f3[q] = hb_div(num, den);
        }
        jl[j] = f3[0]; jd[j] = f3[1]; ju[j] = f3[2];
        if (j == 0) fn0[g] = f3[1];
        if (j == 1) fn0[g] = f3[0];
    }
    __syncthreads();

    // ---- phase 2b: column statistics with scipy's sparse-argmax semantics (column_stats) ---
    const double REJECT = 2.0097183471152322e-14;     // EPS ** 0.875
#pragma unroll 1
    for (int r = 0; r < CPT; ++r) {
        const int j = tid + r * NT;
        if (j >= nx) continue;
        const int lo_r = j > 0 ? j - 1 : 0, hi_r = j < nx - 1 ? j + 1 : nx - 1;
        const double fnv3[3] = {jl[j], jd[j], ju[j]};
        double dv[3] = {0.0, 0.0, 0.0};
        int best = lo_r; double bv = -1.0;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int i = j - 1 + q;
            if (i < lo_r || i > hi_r) continue;
            dv[q] = fnv3[q] - fbv[i];
            const double v = fabs(dv[q]);
            if (i == lo_r || v > bv) { bv = v; best = i; }
        }
        double maxd, fb_mi, fn_mi;
        if (bv > 0.0) { maxd = bv; fb_mi = fbv[best]; fn_mi = (best == j - 1) ? fnv3[0] : (best == j) ? fnv3[1] : fnv3[2]; }
        else { fb_mi = fbv[0]; fn_mi = fn0[c.grp[j]]; maxd = (lo_r == 0) ? fabs(fn_mi - fb_mi) : 0.0; }
        const double scl = fmax(fabs(fb_mi), fabs(fn_mi));
        jl[j] = (j > 0) ? dv[0] : 0.0; jd[j] = dv[1]; ju[j] = (j < nx - 1) ? dv[2] : 0.0;
        vpsi[j] = maxd; vscl[j] = scl;
        vd[j] = (maxd < REJECT * scl) ? 1.0 : 0.0;
    }
    __syncthreads();
}

// Fast mode only (never the default): what one window hands to the next
// one of the same column inside a chunk of rows.
struct Warm {
    double h;        // step size the controller proposed after the last accepted step
    int order, n_equal;
    bool valid;      // the previous window of this column ended successfully in this CTA: difference
                     // table (scratch rows), Jacobian rows and step factors are those of its last step
    bool cont;       // set by the caller: the right-hand side of this window is the same function as in
                     // the previous one (same forcing row values, same noise field) -> continue the
                     // multistep integration instead of restarting it
};

template <int NT, int CPT, bool FULL, bool FAST = false>
__device__ __noinline__ int bdf_attempt(const DevCtx &c, int col,
                                        const double *y_glob, int jac_only = 0, Warm *warm = nullptr,
                                        bool after_jac = false)
{
    // jac_only (test hook): 1 = return after the initial Jacobian, 2 = the same
    // but starting from the step factors already in the G_FAC scratch row.
    // Return value 2 = "grouped differences wanted": the caller runs jac_fused() and calls again with
    // after_jac = true (bdf_solve below).  The pass is called from out there and not from inside the
    // integrator loop because a call of a large function in here makes the register allocator keep half
    // of the controller state in local memory for the WHOLE loop (measured: spill traffic doubled, -8 %);
    // the controller state crosses the gap parked in shared memory.
    constexpr int NV = FULL ? NVEC_FULL : NVEC_BASE;
    enum { S_F0, S_F1, S_JB, S_JG, S_JR, S_NW };
    enum { GO_NONE, GO_JAC_END, GO_STEP_BEGIN, GO_STEP_RETRY, GO_PREDICT, GO_NEWTON_BEGIN,
           GO_JAC_BEGIN };
    const int nx = c.nx, nxp = HB_NXP(c), tid = threadIdx.x;
    const double rtol = c.rtol, atol = c.atol;
    const double newton_tol = c.newton_tol;
    const double *wa = HB_XA(c, NV, X_ARGS);          // t0, t_bound of the window (set_window_args)
    double *D = HB_GROW(c, G_D), *y = HB_VEC(c, V_Y), *b = HB_VEC(c, V_B);
    double *vpsi = HB_VEC(c, V_PSI), *vscl = HB_VEC(c, V_SCL), *vd = HB_VEC(c, V_DD);
    double *fbv = HB_VEC(c, V_TW), *fnv = HB_VEC(c, V_TIP), *ybv = HB_VEC(c, V_TVL),
           *hv = HB_VEC(c, V_TVR), *fac = HB_GROW(c, G_FAC);
    double *jl = HB_VEC(c, V_JL), *jd = HB_VEC(c, V_JD), *ju = HB_VEC(c, V_JU);
    long long *cnt = reinterpret_cast<long long *>(HB_XA(c, NV, X_CNT));
    const Statics S = statics_of<FULL>(c, col);
    int par_ = 0;                                   // block-reduction slot parity

    // controller state (identical in every thread)
    int state = S_F0, order = 1, n_equal = 0, k_newton = 0, jac_g = 0;
    bool lu_valid = false, current_jac = false, have_old = false, jac_first = true, jac_init = true;
    bool any_accept = false;                        // a step of THIS attempt has been accepted
    // grouped forward differences of the Jacobian in one pass (jac_fused) unless the root-uptake
    // sink is active in this window (its global sums need the evaluation per group); the fast-mode
    // instance keeps its three evaluations (measured: the pass saves little there and costs spills)
    // (recomputed where it is needed: nothing of it stays live across the integrator loop)
#define HB_JAC_FUSE(c, NV, wa) (!FAST && (c).jac_fuse && !((int)(wa)[WA_MODE] == 1 && ((int)HB_XA(c, NV, X_PAR)[HB_P_FLAGS] & HB_FLAG_ET)))
    double t = wa[WA_T0], t_new = t, h_abs = 0.0, h_loc = 0.0, min_step = 0.0, cc = 0.0, c_lu = 0.0;
    double dy_norm_old = 0.0, d1_init = 0.0, h0_init = 0.0;

    double *ctl = HB_XA(c, NV, X_RU);               // parked controller state (change_D's matrix is dead then)
    int *cti = reinterpret_cast<int *>(ctl + 10);
    int go = GO_NONE;
    bool jb_now = false;
    __syncthreads();
    if (!FAST && after_jac) {
        // ---- back from jac_fused: the differences of all first-pass groups are in place -----------
        t = ctl[0]; t_new = ctl[1]; h_abs = ctl[2]; h_loc = ctl[3]; min_step = ctl[4]; cc = ctl[5];
        c_lu = ctl[6]; dy_norm_old = ctl[7]; d1_init = ctl[8]; h0_init = ctl[9];
        order = cti[0]; n_equal = cti[1]; k_newton = cti[2]; par_ = cti[3];
        { const int fl = cti[4]; lu_valid = fl & 1; current_jac = fl & 2; have_old = fl & 4; jac_first = fl & 8;
          jac_init = fl & 16; any_accept = fl & 32; }
        // what the S_JG branch does after its last group: the retry pass over the groups that hold a
        // column whose difference was too small (10x step), one evaluation per group
        int want = c.ng;
        for (int j = tid; j < nx; j += NT)
            if (vd[j] != 0.0) want = min(want, c.grp[j]);
        want = -block_max_int<NT, NV>(c, par_, -want);
        state = S_JG;
        if (want >= c.ng) go = GO_JAC_END;
        else {
            HB_FOR_CELLS(i) {
                double step = 0.0;
                if (c.grp[i] == want && vd[i] != 0.0) {
                    const double ybi = ybv[i];
                    const double ysc = ((fbv[i] >= 0.0) ? 1.0 : -1.0) * fmax(atol, fabs(ybi));
                    step = (ybi + 10.0 * fac[i] * ysc) - ybi;
                }
                y[i] = ybv[i] + step;
            }
            jac_g = want; state = S_JR;
        }
    } else
    // __ldcg: the state may have been written by another SM (time-sliced columns)
    if (!(FAST && warm->valid && warm->cont))
        for (int i = tid; i < nx; i += NT) { const double v = __ldcg(y_glob + i); y[i] = v; D[i] = v; }

    // Fast mode, continuation: the previous window of this column ended at t0 with the same
    // right-hand-side function; the BDF history (order, step size, difference table) and the
    // Jacobian are still those of its last step, so the integration simply goes on to the new
    // t_bound -- no select_initial_step, no new Jacobian, no restart at order 1 (what scipy does
    // at every solve_ivp call, bdf.py:198-241, ivp.py:582-660).
    bool resume = false;
    if (FAST && !after_jac) {
        resume = warm->valid && warm->cont;
        if (resume) { order = warm->order; n_equal = warm->n_equal; h_abs = warm->h; t = wa[WA_T0]; }
        warm->valid = false;
    }

    // (re-entry after jac_fused: into the loop behind the evaluation, with `go` as set above)
    if (!FAST && after_jac) goto behind_eval;
    for (;;) {
        go = GO_NONE;
        jb_now = false;
        if (FAST && resume) { resume = false; go = GO_STEP_BEGIN; }
        else {
        // ======== the single right-hand-side evaluation ========================
        const int out = (state == S_JB) ? V_TW : (state == S_JG || state == S_JR) ? V_TIP : V_B;
        __syncthreads();
#if HB_DUP_PROBE
        if (c.dbg & 8192) { __syncthreads(); __syncthreads(); }
#endif
        HB_TICK(c, NV, 0);                   // controller glue since the last mark + loop-top barrier
        rhs_eval<NT, CPT, NV, FAST>(c, S, col, out, par_, FAST && state == S_NW && (c.fast & HB_FAST_F32));
        HB_TICK(c, NV, 2);                   // assembly phase
        if (tid == 0) { cnt[K_NRHS]++; if (state == S_F0 || state == S_F1 || state == S_NW) cnt[K_NFEV]++; }

        if (state == S_NW) {
            // ---- one Newton iteration (solve_bdf_system, bdf.py:41-67) --------
            bool finite = true;
            HB_FOR_CELLS(i) {
                const double f = b[i];
                finite &= isfinite(f);
                b[i] = cc * f - vpsi[i] - vd[i];
            }
#if HB_DUP_PROBE
            if (c.dbg & 128) { double *s2 = HB_GROW(c, G_SK); for (int i = tid; i < nx; i += NT) s2[i] = cc * b[i] - vpsi[i] - vd[i]; }
#endif
            bool converged = false, stop = false;
            if (!__syncthreads_and(finite)) stop = true;
            else {
                HB_TICK(c, NV, 3);           // residual + barrier
#if HB_DUP_PROBE
                if (c.dbg & 3) {             // bit 0: a second solve; bit 1: only its copies and barriers
                    double *s1 = HB_GROW(c, G_TH), *s2 = HB_GROW(c, G_SK);
                    for (int i = tid; i < nx; i += NT) s1[i] = b[i];
                    __syncthreads();
                    if ((c.dbg & 1) && tid < 32) tri_solve_warp0<NV, FULL>(c, c_lu);
                    __syncthreads();
                    for (int i = tid; i < nx; i += NT) { s2[i] = b[i]; b[i] = s1[i]; }
                    __syncthreads();
                }
#endif
                if (tid < 32) tri_solve_warp0<NV, FULL>(c, c_lu);
                if (tid == 0) cnt[K_NNEWTON]++;
                __syncthreads();
                HB_TICK(c, NV, 4);           // tridiagonal solve + barrier
                double loc = 0.0;
#if HB_DUP_PROBE
                if (c.dbg & 256) { double *s1 = HB_GROW(c, G_TH), *s2 = HB_GROW(c, G_SK);
                    for (int i = tid; i < nx; i += NT) { const double dy = b[i]; const double q = hb_div(dy, vscl[i]); loc += q * q; s1[i] = y[i] + dy; s2[i] = vd[i] + dy; }
                    loc = 0.0; }
#endif
                HB_FOR_CELLS(i) {
                    const double dy = tri_x<NV, FULL>(c, i);
                    const double q = div_hot<FAST>(dy, vscl[i]);
                    loc += q * q;
                    y[i] += dy;                    // tentative: discarded if we stop
                    vd[i] += dy;
                }
#if HB_DUP_PROBE
                if (c.dbg & 16) (void)rms<NT, NV>(c, par_, loc);
#endif
                const double dy_norm = rms<NT, NV>(c, par_, loc);
                HB_TICK(c, NV, 5);           // update + norm
                double rate = 0.0;
                if (have_old) {
                    rate = hb_div(dy_norm, dy_norm_old);
                    if (rate >= 1.0 ||
                        hb_div(ipow(rate, HB_NEWTON_MAXITER - k_newton), 1.0 - rate) * dy_norm > newton_tol)
                        stop = true;
                }
                if (!stop && (dy_norm == 0.0 ||
                              (have_old && hb_div(rate, 1.0 - rate) * dy_norm < newton_tol))) {
                    converged = true; stop = true;
                }
                dy_norm_old = dy_norm; have_old = true;
            }
            const int n_iter = k_newton + 1;
            k_newton++;
            if (!stop && k_newton < HB_NEWTON_MAXITER) continue;      // next Newton iteration
            if (!converged) {
                if (!current_jac) {
                    // refresh the Jacobian at (t_new, y_predict) and try again (bdf.py:375-380)
                    __syncthreads();
                    for (int i = tid; i < nx; i += NT) {
                        double acc = D[i];
                        for (int kk = 1; kk <= order; ++kk) acc += D[kk * nxp + i];
                        b[i] = acc;
                    }
                    jac_first = false; jac_init = false;
                    go = GO_JAC_BEGIN;
                } else {
                    h_loc *= 0.5;
                    change_D<NT, CPT, NV>(c, order, 0.5);
                    n_equal = 0; lu_valid = false;
                    if (tid == 0) cnt[K_NREJECT]++;
                    go = GO_STEP_RETRY;
                }
            } else {
                // ---- error test (bdf.py:390-406) ------------------------------
                const double safety = 0.9 * (2 * HB_NEWTON_MAXITER + 1) /
                                      (double)(2 * HB_NEWTON_MAXITER + n_iter);
                double le = 0.0;
                for (int rp_ = 0; rp_ < HB_DUP(c, 512); ++rp_) { le = 0.0;
                HB_FOR_CELLS(i) {
                    const double sc = atol + rtol * fabs(y[i]);
                    vscl[i] = sc;
                    const double q = div_hot<FAST>(k_error_const[order] * vd[i], sc);
                    le += q * q;
                } }
#if HB_DUP_PROBE
                if (c.dbg & 512) (void)rms<NT, NV>(c, par_, le);
#endif
                const double error_norm = rms<NT, NV>(c, par_, le);
                if (error_norm > 1.0) {
                    const double factor = fmax(HB_MIN_FACTOR, safety * inv_root(error_norm, order + 1));
                    h_loc *= factor;
                    change_D<NT, CPT, NV>(c, order, factor);
                    n_equal = 0;
                    if (tid == 0) cnt[K_NREJECT]++;
                    go = GO_STEP_RETRY;
                } else {
                    // ---- accepted: update differences (bdf.py:408-424) --------
                    n_equal++;
                    t = t_new;
                    h_abs = h_loc;
                    any_accept = true;
                    if (tid == 0) {
                        if (c.t_acc && cnt[K_NSTEPS] < nx) c.t_acc[cnt[K_NSTEPS]] = t_new;     // hb_solve_window hook
                        cnt[K_NSTEPS]++;
                    }
                    const bool decide = (n_equal >= order + 1);
#if HB_DUP_PROBE
                    if (c.dbg & 1024) { double *s2 = HB_GROW(c, G_SK);
                        for (int i = tid; i < nx; i += NT) { double a = vd[i] - D[(order + 1) * nxp + i];
                            for (int kk = order; kk >= 0; --kk) a += D[kk * nxp + i] + D[(kk + 1) * nxp + i];
                            s2[i] = a; D[7 * nxp + i] = a; } }
                    if (c.dbg & 2048) change_D<NT, CPT, NV>(c, order, 1.0);
#endif
                    double lm = 0.0, lp = 0.0;
                    HB_FOR_CELLS(i) {
                        const double di = vd[i];
                        D[(order + 2) * nxp + i] = di - D[(order + 1) * nxp + i];
                        D[(order + 1) * nxp + i] = di;
                        for (int kk = order; kk >= 0; --kk) D[kk * nxp + i] += D[(kk + 1) * nxp + i];
                        if (decide) {
                            const double sc = vscl[i];
                            if (order > 1) {
                                const double q = div_hot<FAST>(k_error_const[order - 1] * D[order * nxp + i], sc);
                                lm += q * q;
                            }
                            if (order < HB_MAX_ORDER) {
                                const double q = div_hot<FAST>(k_error_const[order + 1] * D[(order + 2) * nxp + i], sc);
                                lp += q * q;
                            }
                        }
                    }
                    if (decide) {
                        // ---- order / step selection (bdf.py:426-453) ----------
                        double em = INFINITY, ep = INFINITY;
                        if (order > 1) em = rms<NT, NV>(c, par_, lm);
                        if (order < HB_MAX_ORDER) ep = rms<NT, NV>(c, par_, lp);
                        // the three factors are one pow() per warp: lane l % 3 evaluates its own
                        // argument pair, lanes 0..2 broadcast (same function, same bits; every
                        // warp used to evaluate all three one after the other)
                        const int l3 = (tid & 31) % 3;
                        const double r3 = inv_root((l3 == 0) ? em : (l3 == 1) ? error_norm : ep, order + l3);
                        const double fm_ = __shfl_sync(HB_FULL, r3, 0), fe_ = __shfl_sync(HB_FULL, r3, 1),
                                     fp_ = __shfl_sync(HB_FULL, r3, 2);
                        int delta = -1; double best = fm_;
                        if (fe_ > best) { best = fe_; delta = 0; }
                        if (fp_ > best) { best = fp_; delta = 1; }
                        order += delta;
                        const double factor = fmin(HB_MAX_FACTOR, safety * best);
                        h_abs *= factor;
                        change_D<NT, CPT, NV>(c, order, factor);
                        n_equal = 0; lu_valid = false;
                    }
                    go = GO_STEP_BEGIN;
                }
            }
        } else if (state == S_F0) {
            // ---- select_initial_step, first half (common.py:105-115) ----------
            double l0 = 0.0, l1 = 0.0;
            for (int i = tid; i < nx; i += NT) {
                const double y0 = D[i], f0 = b[i];
                D[nxp + i] = f0;                                   // D[1] <- f0 (scaled by h later)
                const double sc = atol + fabs(y0) * rtol;
                vscl[i] = sc;
                const double q0 = y0 / sc, q1 = f0 / sc;
                l0 += q0 * q0; l1 += q1 * q1;
            }
            const double d0 = rms<NT, NV>(c, par_, l0);
            d1_init = rms<NT, NV>(c, par_, l1);
            const double h0 = (d0 < 1e-5 || d1_init < 1e-5) ? 1e-6 : 0.01 * d0 / d1_init;
            h0_init = fmin(h0, fabs(wa[WA_TBOUND] - wa[WA_T0]));
            for (int i = tid; i < nx; i += NT) y[i] = D[i] + h0_init * D[nxp + i];
            if (HB_JAC_FUSE(c, NV, wa)) {
                // C, flux and boundary flux of this evaluation at y0: the base values of the
                // initial Jacobian (the evaluation of the S_F1 state overwrites V_CM / V_FM)
                double *cm_s = HB_GROW(c, G_TH), *fm_s = HB_GROW(c, G_SK);
                for (int i = tid; i < nx - 1; i += NT) { cm_s[i] = HB_VEC(c, V_CM)[i]; fm_s[i] = HB_VEC(c, V_FM)[i]; }
                if (tid == 0) HB_XA(c, NV, X_MISC)[2] = HB_XA(c, NV, X_MISC)[0];
            }
            state = S_F1;
        } else if (state == S_F1) {
            // ---- select_initial_step, second half (common.py:116-134) ---------
            double l2 = 0.0;
            for (int i = tid; i < nx; i += NT) { const double q2 = (b[i] - D[nxp + i]) / vscl[i]; l2 += q2 * q2; }
            const double d2 = rms<NT, NV>(c, par_, l2) / h0_init;
            const double h1 = (d1_init <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0_init * 1e-3)
                                                                 : sqrt(0.01 / fmax(d1_init, d2));
            h_abs = fmin(fmin(100.0 * h0_init, h1), fabs(wa[WA_TBOUND] - wa[WA_T0]));
            // Jacobian at (t0, y0): jac_wrapped's f = fun_single(t0, y0) repeats the
            // evaluation that produced f0 (same inputs, deterministic), so the base
            // value is taken from D[1] = f0 instead of a sixth RHS evaluation
            for (int i = tid; i < nx; i += NT) { const double v = D[i]; ybv[i] = v; fbv[i] = D[nxp + i]; }
            jac_first = (jac_only != 2); jac_init = true;
            jb_now = true;
        } else if (state == S_JB) {
            jb_now = true;
        } else {
            // ---- num_jac: differences of group jac_g (common.py:385-443) ------
            const double REJECT = 2.0097183471152322e-14;     // EPS ** 0.875
            __syncthreads();
            for (int rp_ = 0; rp_ < ((state == S_JG) ? HB_DUP(c, 4096) : 1); ++rp_)
            for (int j = tid; j < nx; j += NT) {
                if (c.grp[j] != jac_g) continue;
                if (state != S_JG && vd[j] == 0.0) continue;       // retry pass: flagged columns only
                ColStat cs;                                        // (one inlined call site: stays in registers)
                column_stats(fbv, fnv, j, nx, &cs);
                if (state == S_JG) {
                    jl[j] = cs.dl; jd[j] = cs.dd; ju[j] = cs.du;
                    vpsi[j] = cs.maxd; vscl[j] = cs.scl;
                    vd[j] = (cs.maxd < REJECT * cs.scl) ? 1.0 : 0.0;
                } else {
                    if (vpsi[j] * cs.scl < cs.maxd * vscl[j]) {
                        fac[j] = 10.0 * fac[j];
                        const double ybi = ybv[j];
                        const double ysc = ((fbv[j] >= 0.0) ? 1.0 : -1.0) * fmax(atol, fabs(ybi));
                        hv[j] = (ybi + fac[j] * ysc) - ybi;
                        jl[j] = cs.dl; jd[j] = cs.dd; ju[j] = cs.du;
                        vscl[j] = cs.scl; vpsi[j] = cs.maxd;
                    }
                    vd[j] = 0.0;                   // handled
                }
            }
            // next evaluation: remaining first-pass groups, then the groups that
            // hold a column whose difference was too small (retry, 10x step)
            int next_state = state, next_g = jac_g + 1;
            bool done = false;
            if (state == S_JG && next_g >= c.ng) { next_state = S_JR; next_g = 0; }
            if (FAST && next_state == S_JR) done = true;       // fast mode: no 10x-step retry pass
            else if (next_state == S_JR) {
                int want = c.ng;
                for (int j = tid; j < nx; j += NT)
                    if (vd[j] != 0.0 && c.grp[j] >= next_g) want = min(want, c.grp[j]);
                want = -block_max_int<NT, NV>(c, par_, -want);
                if (want >= c.ng) done = true; else next_g = want;
            }
            if (done) go = GO_JAC_END;
            else {
                HB_FOR_CELLS(i) {
                    double step = 0.0;
                    if (c.grp[i] == next_g) {
                        if (next_state == S_JG) step = hv[i];
                        else if (vd[i] != 0.0) {
                            const double ybi = ybv[i];
                            const double ysc = ((fbv[i] >= 0.0) ? 1.0 : -1.0) * fmax(atol, fabs(ybi));
                            step = (ybi + 10.0 * fac[i] * ysc) - ybi;
                        }
                    }
                    y[i] = ybv[i] + step;
                }
                jac_g = next_g; state = next_state;
            }
        }

        }   // (not resuming)
behind_eval:

        if (jb_now) {
            // ---- num_jac: steps from the base value (common.py:320-341) -------
            HB_FOR_CELLS(i) {
                double fc = jac_first ? sqrt(HB_EPS) : fac[i];
                const double ybi = ybv[i];
                const double ysc = ((fbv[i] >= 0.0) ? 1.0 : -1.0) * fmax(atol, fabs(ybi));
                double h = (ybi + fc * ysc) - ybi;
                while (h == 0.0) { fc *= 10.0; h = (ybi + fc * ysc) - ybi; }
                fac[i] = fc; hv[i] = h;
                vd[i] = 0.0;
                y[i] = ybi + ((c.grp[i] == 0) ? h : 0.0);
            }
            if (HB_JAC_FUSE(c, NV, wa)) {
                // park the controller state and hand over to jac_fused (see the head of this function)
                const bool initial = (state == S_F1);          // base evaluation: the S_F0 one (saved) or the S_JB one
                __syncthreads();
                if (tid == 0) {
                    ctl[0] = t; ctl[1] = t_new; ctl[2] = h_abs; ctl[3] = h_loc; ctl[4] = min_step; ctl[5] = cc;
                    ctl[6] = c_lu; ctl[7] = dy_norm_old; ctl[8] = d1_init; ctl[9] = h0_init;
                    cti[0] = order; cti[1] = n_equal; cti[2] = k_newton; cti[3] = par_;
                    cti[4] = (lu_valid ? 1 : 0) | (current_jac ? 2 : 0) | (have_old ? 4 : 0) | (jac_first ? 8 : 0) |
                             (jac_init ? 16 : 0) | (any_accept ? 32 : 0) | (initial ? 64 : 0);
                    cnt[K_NRHS] += c.ng;                       // the evaluations scipy makes (one per group)
                }
                return 2;
            } else { jac_g = 0; state = S_JG; }
        }

        if (go == GO_JAC_END) {
            // ---- num_jac: normalise, adapt factors, column -> row storage ------
            const double SMALL = 1.8189894035458565e-12;      // EPS ** 0.75
            const double BIG = 1.220703125e-04;               // EPS ** 0.25
            __syncthreads();
            for (int j = tid; j < nx; j += NT) {
                const double h = hv[j];
                fbv[j] = hb_div(jl[j], h);     // J[j-1][j]: upper entry of row j-1
                jd[j] = hb_div(jd[j], h);
                fnv[j] = hb_div(ju[j], h);     // J[j+1][j]: lower entry of row j+1
                double fc = fac[j];
                if (vpsi[j] < SMALL * vscl[j]) fc *= 10.0;
                if (vpsi[j] > BIG * vscl[j]) fc *= 0.1;
                fac[j] = fmax(fc, 1e3 * HB_EPS);
            }
            __syncthreads();
            HB_FOR_CELLS(i) {
                jl[i] = (i > 0) ? fnv[i - 1] : 0.0;
                ju[i] = (i < nx - 1) ? fbv[i + 1] : 0.0;
            }
            if (tid == 0) cnt[K_NJEV]++;
            lu_valid = false;
            if (jac_init) {
                if (jac_only) { __syncthreads(); return 0; }
                for (int i = tid; i < nx; i += NT) D[nxp + i] = D[nxp + i] * h_abs;   // D[1] = f0 * h
                order = 1; n_equal = 0; t = wa[WA_T0];
                go = GO_STEP_BEGIN;
            } else {
                current_jac = true;
                go = GO_PREDICT;               // same t_new, h: rebuild predictor, psi, scale
            }
        }
        if (go == GO_STEP_BEGIN) {
            // ---- driver loop / top of _step_impl (ivp.py:659, bdf.py:318-333) --
            if (t == wa[WA_TBOUND]) {
                if (FAST) { warm->h = h_abs; warm->order = order; warm->n_equal = n_equal; warm->valid = true; }
                __syncthreads();
                return 0;
            }
            min_step = 10.0 * fabs(nextafter(t, (double)INFINITY) - t);
            if (h_abs < min_step) {
                h_loc = min_step;
                change_D<NT, CPT, NV>(c, order, min_step / h_abs);
                n_equal = 0;
            } else h_loc = h_abs;
            current_jac = false;
            go = GO_STEP_RETRY;
        }
        if (go == GO_STEP_RETRY) {
            // ---- top of the `while not step_accepted` loop (bdf.py:345-358) ---
            if (h_loc < min_step) {
                // failure: solver.y is the last accepted state (ivp.py:659-666).  After an accepted
                // step that is D[0]; before the first one it is the attempt's y0 -- D[0] may have
                // been destroyed by then (a right-hand side that is not finite at t0 gives h = 0 in
                // select_initial_step and change_D(min_step / 0) fills the table with NaN)
                __syncthreads();
                if (any_accept) { for (int i = tid; i < nx; i += NT) y[i] = D[i]; }
                else { for (int i = tid; i < nx; i += NT) y[i] = __ldcg(y_glob + i); }
                __syncthreads();
                return -1;
            }
            t_new = t + h_loc;
            if (t_new - wa[WA_TBOUND] > 0.0) {
                t_new = wa[WA_TBOUND];
                change_D<NT, CPT, NV>(c, order, fabs(t_new - t) / h_loc);
                n_equal = 0; lu_valid = false;
            }
            h_loc = fabs(t_new - t);
            go = GO_PREDICT;
        }
        if (go == GO_PREDICT) {
            // ---- predictor, scale, psi (bdf.py:360-366) ----------------------
            cc = hb_div(t_new - t, k_alpha[order]);
            __syncthreads();
            for (int rp_ = 0; rp_ < HB_DUP(c, 32); ++rp_)
            HB_FOR_CELLS(i) {
                double acc = D[i], ps = 0.0;
                for (int kk = 1; kk <= order; ++kk) {
                    const double dk = D[kk * nxp + i];
                    acc += dk; ps += dk * k_gamma[kk];
                }
                y[i] = acc;
                vscl[i] = atol + rtol * fabs(acc);
                vpsi[i] = div_hot<FAST>(ps, k_alpha[order]);
                vd[i] = 0.0;
            }
            go = GO_NEWTON_BEGIN;
        }
        if (go == GO_NEWTON_BEGIN) {
            if (!lu_valid) {
                __syncthreads();
                HB_TICK(c, NV, 6);           // step control + predictor
                for (int rp_ = 0; rp_ < HB_DUP(c, 64); ++rp_)
                if (tid < 32) tri_factor_warp0<NV, FULL>(c, cc);
                c_lu = cc; lu_valid = true;
                if (tid == 0) cnt[K_NLU]++;
                HB_TICK(c, NV, 7);           // tridiagonal factorisation (warp 0)
            }
            k_newton = 0; have_old = false; dy_norm_old = 0.0;
            state = S_NW;
        }
        if (go == GO_JAC_BEGIN) {
            // base point is in V_B: copy to V_TVL and V_Y
            __syncthreads();
            for (int i = tid; i < nx; i += NT) { const double v = b[i]; ybv[i] = v; y[i] = v; }
            state = S_JB;
        }
    }
}

// One solve_ivp call: bdf_attempt with the grouped Jacobian differences made out here.
template <int NT, int CPT, bool FULL, bool FAST = false>
__device__ __forceinline__ int bdf_solve(const DevCtx &c, int col, const double *y_glob, int jac_only = 0,
                                         Warm *warm = nullptr)
{
    int st = bdf_attempt<NT, CPT, FULL, FAST>(c, col, y_glob, jac_only, warm, false);
    while (st == 2) {
        jac_fused<NT, CPT, FULL>(c, col);
        st = bdf_attempt<NT, CPT, FULL, FAST>(c, col, y_glob, jac_only, warm, true);
    }
    return st;
}

// RichardsPDE.solve (richards_pde.py:478-537): up to 5 attempts, the noise
// field damped in place by 0.8 after each failure (quirk Q4); the state of the
// last attempt is left in V_Y (quirk Q5).  Returns 0 ok / 1 failed.
template <int NT, int CPT, bool FULL, bool FAST = false>
__device__ inline int solve_window(const DevCtx &c, int col, const Args &a, double t0, double t1,
                                   const double *y_glob, bool &noise_damped, Warm *warm = nullptr)
{
    constexpr int NV = FULL ? NVEC_FULL : NVEC_BASE;
    int trials = HB_N_TRIALS, status = 0;
    set_window_args<NV>(c, a, t0, t1);
    while (trials > 0) {
        status = bdf_solve<NT, CPT, FULL, FAST>(c, col, y_glob, 0, warm);
        if (status == 0) break;
        for (int i = threadIdx.x; i < c.nx; i += NT) HB_VEC(c, V_NR)[i] *= HB_NOISE_DAMP;
        noise_damped = true;
        trials--;
        if (threadIdx.x == 0) reinterpret_cast<long long *>(HB_XA(c, NV, X_CNT))[K_NFAIL]++;
        __syncthreads();
    }
    return status == 0 ? 0 : 1;
}

// find_wtd(y >= psi_sat) on the nodes (utilities.py:56-99).  One sync inside.
template <int NT, int NV>
__device__ inline int find_wtd_nodes(const DevCtx &c, int &parity, const double *y)
{
    const double psi_sat = HB_XA(c, NV, X_PAR)[HB_P_PSI_SAT];
    int deepest = -1;
    for (int i = threadIdx.x; i < c.nx; i += NT)
        if (!(y[i] >= psi_sat)) deepest = i;
    deepest = block_max_int<NT, NV>(c, parity, deepest);
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


}  // namespace HB_IMPL_NS
