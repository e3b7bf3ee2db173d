// hb_capi.cu -- kernels and the C-ABI (include/hydrob200.h) of libhydrob200.so.
//
// Kernels (all FP64, no tensor cores -- nothing here is a dense contraction):
//   run_kernel      persistent CTAs pull columns from an atomic queue and run
//                   the window loop of Simulation.run (simulation.py:576-649)
//   spinup_kernel   Simulation.initial_conditions (simulation.py:444-480)
//   hook_kernel     one CTA exposing RHS / closure / Jacobian / tridiagonal
//                   solve / single window for the parity tests
//   philox_base_kernel, moments_kernel   noise initialisation, ensemble stats
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/hydrob200.h"
#ifndef HB_MIN_BLOCKS
#define HB_MIN_BLOCKS 2
#endif
// ---- test hooks ---------------------------------------------------------------
struct HookArgs {
    int mode, col;
    double precip, t0, t1, c, atm;
    int wtd, daylight, spinup, first;
    double *io;          // device scratch: [8][nx]
    long long *stats;    // [HB_NCOUNTER + 2]
};
enum { HK_RHS = 0, HK_NODES, HK_JAC, HK_TRI, HK_WINDOW, HK_WTD };
#include "hb_device.cuh"

using namespace hb;

__global__ void philox_base_kernel(DevCtx c)
{
    const size_t n = (size_t)c.n_cols * c.nx;
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < n;
         g += (size_t)gridDim.x * blockDim.x) {
        const int col = (int)(g / c.nx), i = (int)(g % c.nx);
        c.noise_base[g] = hbr::philox_normal(c.seed, c.col_offset + col, 0, i);
    }
}

// Test hook: the standard normals run_kernel / spinup_kernel draw for (column, window).
__global__ void philox_probe_kernel(unsigned long long seed, long long column, int window, int n, double *out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = hbr::philox_normal(seed, column, window, i);
}

// Register-resident DFMA loop: the FP64 roofline denominator (MEASURED_PEAKS.json
// has no FP64 figure).  8 independent chains per thread, 2 flop per FMA.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0,
           a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// Ensemble statistics, two passes so that the sums do not depend on the schedule:
// every thread owns one element e (a row, or a (output row, node) pair) and adds up
// a slice of the columns -- neighbouring threads read neighbouring addresses -- into
// part[slice][3][M]; moments_final_kernel adds the slices in a fixed order.
// WTD: element = row, value = estimated water-table node index, skipped rows
// (status 2) do not count.  THETA: element = (output row, node), value = theta.
template <bool THETA>
__global__ void __launch_bounds__(256) moments_partial_kernel(DevCtx c, int M, int cols_per_slice, double *part)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M) return;
    const int c0 = blockIdx.y * cols_per_slice, c1 = min(c0 + cols_per_slice, c.n_cols);
    double n = 0.0, s = 0.0, q = 0.0;
    if (THETA) {
        const int o = e / c.nx, k = e - o * c.nx;
        const size_t per_col = (size_t)c.n_out * HB_NFIELD * c.nx;
        const double *f = c.fields + ((size_t)o * HB_NFIELD + HB_F_THETA) * c.nx + k;
        for (int col = c0; col < c1; ++col) { const double v = f[col * per_col]; n += 1.0; s += v; q += v * v; }
    } else {
        for (int col = c0; col < c1; ++col) {
            const double2 *ws = reinterpret_cast<const double2 *>(c.wscal + ((size_t)col * c.T + e) * HB_NWSCAL);
            const double2 a = ws[0], b = ws[1];          // (wtd_est, lateral), (transp, status)
            if (b.y != 2.0) { n += 1.0; s += a.x; q += a.x * a.x; }
        }
    }
    double *dst = part + (size_t)blockIdx.y * 3 * M;
    dst[e] = n; dst[M + e] = s; dst[2 * (size_t)M + e] = q;
}

__global__ void __launch_bounds__(256) moments_final_kernel(const double *part, int M, int n_slices, double *count,
                                                            double *sum, double *sumsq)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M) return;
    double n = 0.0, s = 0.0, q = 0.0;
    for (int k = 0; k < n_slices; ++k) {
        const double *src = part + (size_t)k * 3 * M;
        n += src[e]; s += src[M + e]; q += src[2 * (size_t)M + e];
    }
    if (count) count[e] = n;
    sum[e] = s; sumsq[e] = q;
}

// Launch plan of run_kernel from the cost (RHS evaluations) each column had in the
// previous launch: columns above `factor` x the mean go first and run unpaced
// ("hot"), the others follow in chunk-major order.  One CTA.  Without a history
// (all costs zero) the plan is the identity with no hot column.
__global__ void __launch_bounds__(1024) plan_kernel(const DevCtx c, int cap)
{
    __shared__ double s_part[32];
    __shared__ double s_mean;
    __shared__ int s_cnt, s_hot, s_reg;
    const int tid = threadIdx.x, n = c.n_cols;
    double loc = 0.0;
    for (int i = tid; i < n; i += blockDim.x) loc += (double)c.cost[i];
    for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(HB_FULL, loc, o);
    if ((tid & 31) == 0) s_part[tid >> 5] = loc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
        s_mean = t / (double)n;
    }
    __syncthreads();
    const double mean = s_mean;
    double thr = 1.5 * mean;
    for (int it = 0; it < 12; ++it) {          // raise the bar until the hot set fits `cap`
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        int m = 0;
        for (int i = tid; i < n; i += blockDim.x) m += ((double)c.cost[i] > thr);
        if (m) atomicAdd(&s_cnt, m);
        __syncthreads();
        const int cnt = s_cnt;
        __syncthreads();
        if (cnt <= cap) break;
        thr *= 1.5;
    }
    if (!(mean > 0.0)) thr = INFINITY;
    if (tid == 0) { s_hot = 0; s_reg = 0; }
    __syncthreads();
    // two passes keep the order of both lists deterministic only up to the atomics; the
    // results of the run do not depend on the order (columns are independent)
    for (int i = tid; i < n; i += blockDim.x)
        if ((double)c.cost[i] > thr) c.order[atomicAdd(&s_hot, 1)] = i;
    __syncthreads();
    const int n_hot = s_hot;
    for (int i = tid; i < n; i += blockDim.x)
        if (!((double)c.cost[i] > thr)) c.order[n_hot + atomicAdd(&s_reg, 1)] = i;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) c.cost[i] = 0;
    if (tid == 0) *c.n_hot = n_hot;
}

// ==================================================================================
// Host side
// ==================================================================================
static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return -1; }
#define CK(call)                                                                        \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess)                                                          \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)

struct hb_ctx {
    int device = 0;
    DevCtx d{};
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::vector<void *> allocs;
    struct SlotT { void *p = nullptr; size_t bytes = 0; };
    SlotT s_precip, s_atm, s_wtd, s_day, s_series, s_wscal, s_fields, s_tape, s_nr0, s_zw, s_it, s_mom;
    double *hook_io = nullptr;
    long long *hook_stats = nullptr;
    int64_t launches = 0;
    int sms = 0;
    bool grid_set = false, cols_set = false, forcing_set = false;
    std::vector<int> h_flags;            // Simulation_Flags bitmask per column (host copy)
    std::vector<int> h_daylight, h_series;  // forcing daylight flags / series map (host copies)
    int variant = 0;        // kernel variant, see HB_FOR_VARIANT
    int shape = 0;          // experimental shape of the fixed-length instance (HB_SHAPE), 0 = shipped
    bool fixed = false;     // the compile-time nxp = 416 instance (namespace hbf) applies
    size_t smem_bytes = 0;
    int blocks_per_sm = 1;
    bool chunk_fixed = false;  // HB_CHUNK given: no automatic chunk length
    int wscal_T = 0;        // table length the per-row result buffer was laid out for
    bool adaptive = true;   // run_kernel launch plan from the previous launch's per-column cost
    bool have_cost = false; // a previous hb_run left per-column costs on the device
    int fast_flags = 0;     // HB_FAST_* (fast mode, never the default)
    int *grp3 = nullptr;    // colouring j mod 3 of the tridiagonal pattern (fast mode)
};

template <typename T>
static int dev_alloc(hb_ctx *ctx, T **p, size_t n)
{
    void *q = nullptr;
    CK(cudaMalloc(&q, n * sizeof(T) + 16));
    CK(cudaMemset(q, 0, n * sizeof(T) + 16));
    ctx->allocs.push_back(q);
    *p = static_cast<T *>(q);
    return 0;
}

// (Re)allocate a named buffer only when it has to grow: repeated hb_set_* calls
// (every benchmark step re-uploads its inputs) must not leak device memory.
// Buffers the caller overwrites completely right away pass zero = false; result
// buffers are cleared when they are (re)allocated or when their shape changes.
template <typename T, typename SlotT>
static int slot_alloc(hb_ctx *ctx, SlotT &s, T **p, size_t n, bool zero)
{
    const size_t need = n * sizeof(T) + 16;
    if (s.bytes < need) {
        if (s.p) { cudaFree(s.p); s.p = nullptr; s.bytes = 0; }
        CK(cudaMalloc(&s.p, need));
        s.bytes = need;
        zero = true;
    }
    if (zero) CK(cudaMemsetAsync(s.p, 0, need, ctx->stream));
    *p = static_cast<T *>(s.p);
    return 0;
}

// Tuning knobs read from the environment (A/B experiments only): every active one
// is reported on stderr so that a run never changes code paths silently.
static const char *tuning_env(const char *name)
{
    const char *v = getenv(name);
    if (v && *v) {
        static bool told[10] = {false};
        static const char *names[10] = {"HB_CARVEOUT", "HB_VARIANT", "HB_NO_FIXED", "HB_NO_TRI13", "HB_CHUNK",
                                        "HB_DBG", "HB_NO_ADAPTIVE", "HB_TRI_DEBUG", "HB_SHAPE", "HB_JAC_SEQ"};
        for (int k = 0; k < 10; ++k)
            if (!strcmp(names[k], name) && !told[k]) {
                told[k] = true;
                fprintf(stderr, "[libhydrob200] tuning knob %s=%s is active\n", name, v);
            }
        return v;
    }
    return nullptr;
}

template <typename K>
static int prep_kernel(hb_ctx *ctx, K kernel, int nt)
{
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_bytes));
    if (const char *v = tuning_env("HB_CARVEOUT"))            // tuning experiments: shared-memory carve-out in percent
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(v)));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, nt, ctx->smem_bytes));
    if (nb < 1) return fail("kernel does not fit on an SM");
    ctx->blocks_per_sm = nb;
    return 0;
}

// Kernel variants: <threads per column, cells per thread, min CTAs per SM,
// FULL = static midpoint profiles staged in shared memory>.  Measured on B200
// with one binary (1,024 columns x 120 windows, nx = 401), column-windows/s:
//   <256,2,2,true>  515 K   (default: 19 shared vectors = 66 KB, 128 registers, 2 CTAs/SM)
//   <256,2,3,true>  441 K   (3 CTAs/SM, 80 registers: spills)
//   <192,3,3,true>  432 K   (3 CTAs/SM, 96 registers)
//   <128,4,3,true>  478 K   (3 CTAs/SM, 168 registers)
// i.e. a third resident column per SM does not pay; only the winner is built.
//   0: <256,2,2,true>   nx <= 448
//   1: <256,2,2,false>  nx <= 512
//   2: <512,2,1,false>  nx <= 1024
//   3: <512,3,1,false>  nx <= 1536
// STMT sees NT_, CPT_, MB_, FL_ and the instance namespace alias K (hbf: nxp fixed at 416
// at compile time, variant 0 only; hbr: nxp at run time).
// (HB_V0_*: shape of variant 0 in A/B builds, tools/build_alt.py; the shipped library uses 256 / 2 / 2)
#ifndef HB_SHAPES
#define HB_SHAPES 0
#endif
#ifndef HB_V0_NT
#define HB_V0_NT 256
#define HB_V0_CPT 2
#define HB_V0_MB 2
#endif
#ifndef HB_V0_FULL
#define HB_V0_FULL true
#endif
// -DHB_LEAN=1 (A/B builds only): compile the fixed-length instance alone (401-node grids, parity mode)
#ifndef HB_LEAN
#define HB_LEAN 0
#endif
// Experimental shapes of the fixed-length instance (HB_SHAPE=n, A/B builds with -DHB_SHAPES=1):
//   1: <32,13,4,false>  one warp per column, 4 columns per SM
//   2: <64,7,4,false>   two warps per column
//   3: <128,4,4,false>  four warps per column, 4 columns per SM
//   4: <32,13,3,true>   one warp per column, statics in shared memory, 3 columns per SM
#if HB_SHAPES
#define HB_EXTRA_SHAPES(STMT) \
            case 1: { constexpr int NT_ = 32, CPT_ = 13, MB_ = 4; constexpr bool FL_ = false; STMT; } break; \
            case 2: { constexpr int NT_ = 64, CPT_ = 7, MB_ = 4; constexpr bool FL_ = false; STMT; } break; \
            case 3: { constexpr int NT_ = 128, CPT_ = 4, MB_ = 4; constexpr bool FL_ = false; STMT; } break; \
            case 4: { constexpr int NT_ = 32, CPT_ = 13, MB_ = 3; constexpr bool FL_ = true; STMT; } break;
#else
#define HB_EXTRA_SHAPES(STMT)
#endif
#define HB_FOR_VARIANT(ctx, STMT)                                   \
    do {                                                            \
        if ((ctx)->fixed) {                                         \
            namespace K = hbf;                                      \
            switch ((ctx)->shape) {                                 \
            HB_EXTRA_SHAPES(STMT)                                   \
            default: { constexpr int NT_ = HB_V0_NT, CPT_ = HB_V0_CPT, MB_ = HB_V0_MB; constexpr bool FL_ = HB_V0_FULL; STMT; } break; \
            }                                                       \
            break;                                                  \
        }                                                           \
        HB_OTHER_VARIANTS(ctx, STMT);                               \
    } while (0)
#if HB_LEAN
#define HB_OTHER_VARIANTS(ctx, STMT) do { fprintf(stderr, "[libhydrob200] lean A/B build: 385..416-node grids only\n"); abort(); } while (0)
#else
#define HB_OTHER_VARIANTS(ctx, STMT)                                \
    do {                                                            \
        namespace K = hbr;                                          \
        switch ((ctx)->variant) {                                   \
        case 0: { constexpr int NT_ = HB_V0_NT, CPT_ = HB_V0_CPT, MB_ = HB_V0_MB; constexpr bool FL_ = true; STMT; } break;  \
        case 1: { constexpr int NT_ = 256, CPT_ = 2, MB_ = 2; constexpr bool FL_ = false; STMT; } break; \
        case 2: { constexpr int NT_ = 512, CPT_ = 2, MB_ = 1; constexpr bool FL_ = false; STMT; } break; \
        default: { constexpr int NT_ = 512, CPT_ = 3, MB_ = 1; constexpr bool FL_ = false; STMT; } break; \
        }                                                           \
    } while (0)
#endif

#define DISPATCH_RUN(ctx, FAST_, grid, ...)                                                   \
    do {                                                                                      \
        HB_FOR_VARIANT(ctx, (K::run_kernel<NT_, CPT_, MB_, FL_, (FAST_) && !HB_LEAN><<<grid, NT_, (ctx)->smem_bytes, (ctx)->stream>>>(__VA_ARGS__))); \
        (ctx)->launches++;                                                                    \
    } while (0)

#define DISPATCH(ctx, NAME, grid, ...)                                                        \
    do {                                                                                      \
        HB_FOR_VARIANT(ctx, (K::NAME<NT_, CPT_, MB_, FL_><<<grid, NT_, (ctx)->smem_bytes, (ctx)->stream>>>(__VA_ARGS__))); \
        (ctx)->launches++;                                                                    \
    } while (0)

extern "C" {

const char *hb_last_error(void) { return g_err.c_str(); }
int hb_version(void) { return 200; }
int hb_max_nx(void) { return 1536; }

hb_ctx *hb_create(int device, int nx, int n_cols)
{
    if (nx < 4 || nx > hb_max_nx() || n_cols < 1) { fail("hb_create: bad nx / n_cols"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fail("hb_create: no CUDA device (this library has no CPU fallback)"); return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { fail("hb_create: cudaSetDevice failed"); return nullptr; }
    hb_ctx *ctx = new hb_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->sms = prop.multiProcessorCount;
    DevCtx &d = ctx->d;
    d.nx = nx; d.nxp = (nx + 1) & ~1; d.n_cols = n_cols;
    // 385..416 nodes (the reference's 401-node default): pad every vector to
    // 32 x 13 rows so that the tridiagonal solver runs its unrolled variant
    if (nx > 32 * 12 && nx <= 32 * 13 && !tuning_env("HB_NO_TRI13")) d.nxp = 32 * 13;
    d.rtol = 1e-3; d.atol = 1e-3; d.sqrt_nx = sqrt((double)nx);
    d.newton_tol = fmax(10.0 * HB_EPS / d.rtol, fmin(0.03, sqrt(d.rtol)));
    ctx->variant = nx <= 448 ? 0 : (nx <= 512 ? 1 : (nx <= 1024 ? 2 : 3));
    if (const char *v = tuning_env("HB_VARIANT")) {          // tuning experiments
        const int want = atoi(v);
        if (nx <= 512 && want == 1) ctx->variant = 1;
    }
    ctx->fixed = ctx->variant == 0 && d.nxp == HB_FIXED_NXP && !tuning_env("HB_NO_FIXED");
    bool full = ctx->variant == 0 && (!ctx->fixed || HB_V0_FULL);
#if HB_SHAPES
    if (ctx->fixed) if (const char *v = tuning_env("HB_SHAPE")) { ctx->shape = atoi(v); full = ctx->shape == 0 || ctx->shape == 4; }
#endif
    ctx->smem_bytes = smem_doubles(d.nxp, full) * sizeof(double);
    {   // BDF coefficient tables exactly as scipy builds them (bdf.py:244-247)
        const double kappa[6] = {0.0, -0.1850, -1.0 / 9.0, -0.0823, -0.0415, 0.0};
        double gam[6], alp[6], erc[6];
        gam[0] = 0.0;
        for (int k = 1; k <= 5; ++k) gam[k] = gam[k - 1] + 1.0 / k;
        for (int k = 0; k <= 5; ++k) { alp[k] = (1.0 - kappa[k]) * gam[k]; erc[k] = kappa[k] * gam[k] + 1.0 / (k + 1); }
        // U of change_D, the operations and their order exactly as the device used to make them
        // (and scipy's compute_R(order, 1), bdf.py:18-27, does)
        double umat[36];
        for (int k = 0; k < 6; ++k)
            for (int j = 0; j < 6; ++j) {
                volatile double ukj = 1.0;
                for (int rr = 1; rr <= k; ++rr) {
                    volatile double num = (double)(rr - 1) - 1.0 * (double)j;
                    volatile double quo = num / (double)rr;
                    ukj = (j == 0) ? ukj * 0.0 : ukj * quo;
                }
                umat[k * 6 + j] = ukj;
            }
        if (cudaMemcpyToSymbol(hb::k_U, umat, sizeof(umat)) != cudaSuccess ||
            cudaMemcpyToSymbol(hb::k_gamma, gam, sizeof(gam)) != cudaSuccess ||
            cudaMemcpyToSymbol(hb::k_alpha, alp, sizeof(alp)) != cudaSuccess ||
            cudaMemcpyToSymbol(hb::k_error_const, erc, sizeof(erc)) != cudaSuccess) {
            fail("hb_create: constant upload failed"); delete ctx; return nullptr;
        }
    }
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        fail("hb_create: stream"); delete ctx; return nullptr;
    }
    ctx->stream = ctx->own_stream;
    int rc = 0;
    double *x, *dx, *zx, *xz, *par, *stat, *psi, *nb; int *grp, *queue, *slot, *done, *credit, *order, *nhot, *errf;
    long long *cnt, *cost;
    rc |= dev_alloc(ctx, &x, nx); rc |= dev_alloc(ctx, &dx, nx); rc |= dev_alloc(ctx, &zx, nx);
    rc |= dev_alloc(ctx, &xz, nx); rc |= dev_alloc(ctx, &grp, nx);
    rc |= dev_alloc(ctx, &par, (size_t)n_cols * HB_NPARAM);
    rc |= dev_alloc(ctx, &stat, (size_t)n_cols * HB_NSTATIC * nx);
    rc |= dev_alloc(ctx, &psi, (size_t)n_cols * nx);
    rc |= dev_alloc(ctx, &nb, (size_t)n_cols * nx);
    rc |= dev_alloc(ctx, &queue, 4); rc |= dev_alloc(ctx, &slot, n_cols); rc |= dev_alloc(ctx, &done, n_cols);
    rc |= dev_alloc(ctx, &credit, n_cols); rc |= dev_alloc(ctx, &order, n_cols); rc |= dev_alloc(ctx, &nhot, 4);
    rc |= dev_alloc(ctx, &errf, 4); rc |= dev_alloc(ctx, &cost, n_cols);
    rc |= dev_alloc(ctx, &cnt, (size_t)n_cols * HB_NCOUNTER);
    rc |= dev_alloc(ctx, &ctx->hook_io, (size_t)8 * nx + 8);
    rc |= dev_alloc(ctx, &ctx->hook_stats, HB_NCOUNTER + 2);
    if (rc) { hb_destroy(ctx); return nullptr; }
    d.x = x; d.dx = dx; d.zx = zx; d.xz = xz; d.grp = grp; d.par = par; d.stat = stat;
    d.psi = psi; d.noise_base = nb; d.queue = queue; d.slot_next = slot; d.counters = cnt;
    d.done = done; d.chunk = 16;
    d.credit = credit; d.order = order; d.n_hot = nhot; d.err_flags = errf; d.cost = cost;
    d.t_acc = nullptr;
    {
        std::vector<int> g3(nx);
        for (int k = 0; k < nx; ++k) g3[k] = k % 3;
        if (dev_alloc(ctx, &ctx->grp3, nx) ||
            cudaMemcpy(ctx->grp3, g3.data(), nx * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
            fail("hb_create: colouring upload failed"); hb_destroy(ctx); return nullptr;
        }
    }
    if (tuning_env("HB_NO_ADAPTIVE")) ctx->adaptive = false;
    { long long *pc; if (dev_alloc(ctx, &pc, 16)) { hb_destroy(ctx); return nullptr; } d.phase_cycles = pc; }
    if (const char *v = tuning_env("HB_CHUNK")) { const int w = atoi(v); if (w > 0) { d.chunk = w; ctx->chunk_fixed = true; } }
    d.dbg = 0; d.fast = 0;
    d.jac_fuse = tuning_env("HB_JAC_SEQ") ? 0 : 1;   // (A/B and the bit-identity test: one evaluation per group)
#if HB_DUP_PROBE
    if (const char *v = tuning_env("HB_DBG")) d.dbg = atoi(v);
#endif
    d.noise_mode = HB_NOISE_TAPE; d.n_slots = 0; d.tape = nullptr;
    HB_FOR_VARIANT(ctx, {
        rc = prep_kernel(ctx, K::hook_kernel<NT_, CPT_, MB_, FL_>, NT_);
        if (!rc) rc = prep_kernel(ctx, K::spinup_kernel<NT_, CPT_, MB_, FL_>, NT_);
        if (!rc) rc = prep_kernel(ctx, K::run_kernel<NT_, CPT_, MB_, FL_, !HB_LEAN>, NT_);
        if (!rc) rc = prep_kernel(ctx, K::run_kernel<NT_, CPT_, MB_, FL_, false>, NT_);
    });
    if (rc) { hb_destroy(ctx); return nullptr; }
    {   // per-CTA global scratch of the resident CTAs (difference table, step factors, ET vectors)
        double *db;
        if (dev_alloc(ctx, &db, (size_t)ctx->sms * ctx->blocks_per_sm * hb::G_ROWS * d.nxp)) { hb_destroy(ctx); return nullptr; }
        d.dbuf = db;
    }
    return ctx;
}

int hb_destroy(hb_ctx *ctx)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (void *p : ctx->allocs) cudaFree(p);
    for (auto *sl : {&ctx->s_precip, &ctx->s_atm, &ctx->s_wtd, &ctx->s_day, &ctx->s_series, &ctx->s_wscal,
                     &ctx->s_fields, &ctx->s_tape, &ctx->s_nr0, &ctx->s_zw, &ctx->s_it, &ctx->s_mom})
        if (sl->p) cudaFree(sl->p);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return 0;
}

int hb_set_stream(hb_ctx *ctx, void *s)
{
    if (!ctx) return fail("null ctx");
    ctx->stream = s ? static_cast<cudaStream_t>(s) : ctx->own_stream;
    return 0;
}

// Wait for the stream and turn error bits raised by kernels into a failure.
static int finish(hb_ctx *ctx)
{
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    int flags = 0;
    CK(cudaMemcpyAsync(&flags, ctx->d.err_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (flags & HB_ERR_TAPE_SHORT) {
        CK(cudaMemsetAsync(ctx->d.err_flags, 0, sizeof(int), ctx->stream));
        return fail("noise tape too short: a column refreshed its field more often than the tape has slots "
                    "(hb_set_noise_tape n_slots; count the refresh rows with the column's precipitation multiplier)");
    }
    return 0;
}

int hb_sync(hb_ctx *ctx)
{
    if (!ctx) return fail("null ctx");
    CK(cudaSetDevice(ctx->device));
    return finish(ctx);
}

int hb_set_grid(hb_ctx *ctx, const double *x, const int32_t *groups, int n_groups)
{
    if (!ctx || !x || !groups) return fail("hb_set_grid: null argument");
    const int nx = ctx->d.nx;
    std::vector<double> dx(nx, 1.0), zx(nx, 0.0), xz(nx, 0.0), xm(nx, 0.0);
    for (int k = 0; k + 1 < nx; ++k) {
        dx[k] = x[k + 1] - x[k];
        if (!(dx[k] > 0.0)) return fail("hb_set_grid: Space domain is not increasing.");
        xm[k] = x[k] + 0.5 * dx[k];            // richards_pde.py:67
        zx[k] = xm[k] - x[k];                  // richards_pde.py:79
    }
    for (int k = 1; k < nx; ++k) xz[k] = x[k] - xm[k - 1];   // richards_pde.py:76
    if (n_groups < 1 || n_groups > 8) return fail("hb_set_grid: n_groups out of range");
    for (int k = 0; k < nx; ++k) {
        if (groups[k] < 0 || groups[k] >= n_groups) return fail("hb_set_grid: bad group id");
        if (k > 0 && groups[k] == groups[k - 1]) return fail("hb_set_grid: colouring is not valid");
        if (k > 1 && groups[k] == groups[k - 2]) return fail("hb_set_grid: colouring is not valid");
    }
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync((void *)ctx->d.x, x, nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync((void *)ctx->d.dx, dx.data(), nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync((void *)ctx->d.zx, zx.data(), nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync((void *)ctx->d.xz, xz.data(), nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync((void *)ctx->d.grp, groups, nx * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->d.ng = n_groups;
    ctx->grid_set = true;
    return 0;
}

int hb_set_tolerances(hb_ctx *ctx, double rtol, double atol)
{
    if (!ctx || !(rtol > 0.0) || !(atol >= 0.0)) return fail("hb_set_tolerances: bad argument");
    ctx->d.rtol = rtol; ctx->d.atol = atol;
    ctx->d.newton_tol = fmax(10.0 * HB_EPS / rtol, fmin(0.03, sqrt(rtol)));
    return 0;
}

static int range_ok(hb_ctx *ctx, int col0, int ncol)
{
    if (!ctx) return fail("null ctx");
    if (col0 < 0 || ncol < 0 || col0 + ncol > ctx->d.n_cols) return fail("column range out of bounds");
    return 0;
}

int hb_set_columns(hb_ctx *ctx, int col0, int ncol, const double *par, const double *stat)
{
    if (range_ok(ctx, col0, ncol)) return -1;
    if (!par || !stat) return fail("hb_set_columns: null argument");
    const int nx = ctx->d.nx;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync((void *)(ctx->d.par + (size_t)col0 * HB_NPARAM), par,
                       (size_t)ncol * HB_NPARAM * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync((void *)(ctx->d.stat + (size_t)col0 * HB_NSTATIC * nx), stat,
                       (size_t)ncol * HB_NSTATIC * nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if ((int)ctx->h_flags.size() != ctx->d.n_cols) ctx->h_flags.assign(ctx->d.n_cols, 0);
    for (int k = 0; k < ncol; ++k) ctx->h_flags[col0 + k] = (int)par[(size_t)k * HB_NPARAM + HB_P_FLAGS];
    ctx->cols_set = true;
    return 0;
}

int hb_set_state(hb_ctx *ctx, int col0, int ncol, const double *psi)
{
    if (range_ok(ctx, col0, ncol)) return -1;
    if (!psi) return fail("hb_set_state: null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->d.psi + (size_t)col0 * ctx->d.nx, psi, (size_t)ncol * ctx->d.nx * sizeof(double),
                       cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int hb_get_state(hb_ctx *ctx, int col0, int ncol, double *psi)
{
    if (range_ok(ctx, col0, ncol)) return -1;
    if (!psi) return fail("hb_get_state: null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(psi, ctx->d.psi + (size_t)col0 * ctx->d.nx, (size_t)ncol * ctx->d.nx * sizeof(double),
                       cudaMemcpyDeviceToHost, ctx->stream));
    return finish(ctx);
}

int hb_set_forcing(hb_ctx *ctx, int n_series, int T, const double *precip, const int32_t *wtd_idx,
                   const int32_t *daylight, const double *atm, const int32_t *series_of_col)
{
    if (!ctx || !precip || !wtd_idx || !daylight) return fail("hb_set_forcing: null argument");
    if (n_series < 1 || T < 1) return fail("hb_set_forcing: empty forcing");
    DevCtx &d = ctx->d;
    const size_t n = (size_t)n_series * T;
    for (size_t k = 0; k < n; ++k)
        if (wtd_idx[k] >= d.nx) return fail("hb_set_forcing: water-table index outside the grid");
    if (series_of_col)
        for (int k = 0; k < d.n_cols; ++k)
            if (series_of_col[k] < 0 || series_of_col[k] >= n_series) return fail("hb_set_forcing: bad series id");
    CK(cudaSetDevice(ctx->device));
    double *p, *a; int *w, *dl, *s = nullptr;
    if (slot_alloc(ctx, ctx->s_precip, &p, n, false) || slot_alloc(ctx, ctx->s_atm, &a, n, !atm) ||
        slot_alloc(ctx, ctx->s_wtd, &w, n, false) || slot_alloc(ctx, ctx->s_day, &dl, n, false)) return -1;
    CK(cudaMemcpyAsync(p, precip, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (atm) CK(cudaMemcpyAsync(a, atm, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(w, wtd_idx, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dl, daylight, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (series_of_col) {
        if (slot_alloc(ctx, ctx->s_series, &s, (size_t)d.n_cols, false)) return -1;
        CK(cudaMemcpyAsync(s, series_of_col, d.n_cols * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    d.precip = p; d.atm = a; d.wtd_idx = w; d.daylight = dl; d.series = s;
    d.n_series = n_series; d.T = T;
    ctx->h_daylight.assign(daylight, daylight + n);
    if (series_of_col) ctx->h_series.assign(series_of_col, series_of_col + d.n_cols); else ctx->h_series.clear();
    // per-row results: cleared only when the table length changes (a re-upload of the same
    // forcing between launches keeps the rows already written)
    double *ws;
    const bool new_shape = !ctx->forcing_set || ctx->wscal_T != T;
    if (slot_alloc(ctx, ctx->s_wscal, &ws, (size_t)d.n_cols * T * HB_NWSCAL, new_shape)) return -1;
    ctx->wscal_T = T;
    d.wscal = ws;
    if (new_shape) { d.fields = nullptr; d.n_out = 0; d.out_stride = 0; }
    ctx->forcing_set = true;
    return 0;
}

int hb_set_noise_tape(hb_ctx *ctx, int col0, int ncol, int n_slots, const double *tape)
{
    if (range_ok(ctx, col0, ncol)) return -1;
    if (n_slots < 1 || !tape) return fail("hb_set_noise_tape: bad argument");
    DevCtx &d = ctx->d;
    CK(cudaSetDevice(ctx->device));
    if (d.n_slots != n_slots || !d.tape) {
        double *t;
        if (slot_alloc(ctx, ctx->s_tape, &t, (size_t)d.n_cols * n_slots * d.nx, false)) return -1;
        d.tape = t; d.n_slots = n_slots;
    }
    CK(cudaMemcpyAsync((void *)(d.tape + (size_t)col0 * n_slots * d.nx), tape,
                       (size_t)ncol * n_slots * d.nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // slot 0 is the run's first draw: it becomes the (mutable) base field
    CK(cudaMemcpy2DAsync(d.noise_base + (size_t)col0 * d.nx, d.nx * sizeof(double),
                         d.tape + (size_t)col0 * n_slots * d.nx, (size_t)n_slots * d.nx * sizeof(double),
                         d.nx * sizeof(double), ncol, cudaMemcpyDeviceToDevice, ctx->stream));
    std::vector<int> ones(ncol, 1);
    CK(cudaMemcpyAsync(d.slot_next + col0, ones.data(), ncol * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    d.noise_mode = HB_NOISE_TAPE;
    return 0;
}

int hb_set_noise_philox(hb_ctx *ctx, uint64_t seed, int64_t first_global_column)
{
    if (!ctx) return fail("null ctx");
    DevCtx &d = ctx->d;
    d.noise_mode = HB_NOISE_PHILOX; d.seed = seed; d.col_offset = first_global_column;
    CK(cudaSetDevice(ctx->device));
    philox_base_kernel<<<ctx->sms * 4, 256, 0, ctx->stream>>>(d);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

int hb_set_output(hb_ctx *ctx, int out_stride)
{
    if (!ctx || !ctx->forcing_set) return fail("hb_set_output: call hb_set_forcing first");
    DevCtx &d = ctx->d;
    if (out_stride <= 0) { d.out_stride = 0; d.n_out = 0; d.fields = nullptr; return 0; }
    const int n_out = (d.T - 1) / out_stride + 1 + (((d.T - 1) % out_stride) != 0);
    const bool new_shape = !d.fields || d.out_stride != out_stride || d.n_out != n_out;
    d.out_stride = out_stride;
    d.n_out = n_out;
    double *f;
    CK(cudaSetDevice(ctx->device));
    if (slot_alloc(ctx, ctx->s_fields, &f, (size_t)d.n_cols * d.n_out * HB_NFIELD * d.nx, new_shape)) return -1;
    d.fields = f;
    return 0;
}

int hb_n_output_rows(hb_ctx *ctx) { return ctx ? ctx->d.n_out : -1; }

static int ready(hb_ctx *ctx)
{
    if (!ctx) return fail("null ctx");
    if (!ctx->grid_set) return fail("grid not set (hb_set_grid)");
    if (!ctx->cols_set) return fail("columns not set (hb_set_columns)");
    return 0;
}

// Modes that are not implemented on the device must fail loudly, never be
// skipped: PREDICT changes the lateral-flow sink (richards_pde.py:312-351; it
// raises in the reference itself on NumPy 2).
static int sinks_supported(hb_ctx *ctx, int w0, int w1)
{
    (void)w0; (void)w1;
    for (int col = 0; col < ctx->d.n_cols; ++col) {
        const int fl = ctx->h_flags.empty() ? 0 : ctx->h_flags[col];
        if ((fl & HB_FLAG_PREDICT) && (fl & HB_FLAG_LF))
            return fail("PREDICT lateral-flow mode is not implemented (the reference itself raises on NumPy 2)");
    }
    return 0;
}

static int persistent_grid(hb_ctx *ctx)
{
    const int cap = ctx->sms * ctx->blocks_per_sm;
    return ctx->d.n_cols < cap ? ctx->d.n_cols : cap;
}

int hb_spinup(hb_ctx *ctx, int max_iter, const double *nrnd0, const double *zwtd0_cm, int32_t *iters_out)
{
    if (ready(ctx)) return -1;
    if (!ctx->forcing_set) return fail("hb_spinup: forcing not set");
    if (!zwtd0_cm) return fail("hb_spinup: null zwtd0_cm");
    DevCtx &d = ctx->d;
    if (d.noise_mode == HB_NOISE_TAPE && !nrnd0) return fail("hb_spinup: TAPE mode needs nrnd0");
    CK(cudaSetDevice(ctx->device));
    double *nr = nullptr, *zw; int *it;
    if (nrnd0) {
        if (slot_alloc(ctx, ctx->s_nr0, &nr, (size_t)d.n_cols * d.nx, false)) return -1;
        CK(cudaMemcpyAsync(nr, nrnd0, (size_t)d.n_cols * d.nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (slot_alloc(ctx, ctx->s_zw, &zw, (size_t)d.n_cols, false) || slot_alloc(ctx, ctx->s_it, &it, (size_t)d.n_cols, true)) return -1;
    CK(cudaMemcpyAsync(zw, zwtd0_cm, d.n_cols * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(d.queue, 0, 2 * sizeof(int), ctx->stream));
    const int grid = persistent_grid(ctx);
    DISPATCH(ctx, spinup_kernel, grid, d, max_iter, nr, zw, it);
    CK(cudaGetLastError());
    if (iters_out) {
        CK(cudaMemcpyAsync(iters_out, it, d.n_cols * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
    }
    return 0;
}

int hb_run(hb_ctx *ctx, int w0, int w1)
{
    if (ready(ctx)) return -1;
    if (!ctx->forcing_set) return fail("hb_run: forcing not set");
    DevCtx &d = ctx->d;
    if (w0 < 0 || w1 > d.T || w0 >= w1) return fail("hb_run: bad window range");
    if (d.noise_mode == HB_NOISE_TAPE && !d.tape) return fail("hb_run: noise not set");
    if (sinks_supported(ctx, w0, w1)) return -1;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(d.queue, 0, 2 * sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(d.done, 0, (size_t)d.n_cols * sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(d.credit, 0, (size_t)d.n_cols * sizeof(int), ctx->stream));
    const int grid = persistent_grid(ctx);
    DevCtx k = d;
    if (!ctx->chunk_fixed) {
        // Rows per work item.  Every chunk re-stages the column (state, noise field, statics:
        // ~22 KB of HBM traffic at nx = 401), so chunks should be long; the tail of the launch is
        // one chunk, so they should be short.  Keep the tail below ~1 % of the launch:
        // chunk <= 0.01 * (columns per resident CTA) * rows, at least 16 rows
        // (65,536 columns x 120 rows: the whole slice in one piece, DRAM traffic of the launch
        // 11.9 GB -> measured in profiles/; 8,192 columns: 33 rows; 1,024 columns x 365 rows: 16).
        const long long rows = w1 - w0;
        long long len = (long long)(0.01 * (double)d.n_cols / (double)grid * (double)rows);
        if (len < 16) len = 16;
        if (len > rows) len = rows;
        k.chunk = (int)len;
    }
    if (ctx->adaptive && ctx->have_cost && d.n_cols > grid) {
        // columns the previous launch found expensive start first and run unpaced
        plan_kernel<<<1, 1024, 0, ctx->stream>>>(d, grid / 2);
        ctx->launches++;
    } else {
        k.order = nullptr;
        CK(cudaMemsetAsync(d.cost, 0, (size_t)d.n_cols * sizeof(long long), ctx->stream));
    }
    if (ctx->fast_flags) {
        // fast mode (never the default): continuation across windows with unchanged forcing,
        // 3-colour Jacobian without retry pass
        k.fast = ctx->fast_flags;
        if (ctx->fast_flags & HB_FAST_JAC3) { k.grp = ctx->grp3; k.ng = 3; }
        DISPATCH_RUN(ctx, true, grid, k, w0, w1);
    } else {
        DISPATCH_RUN(ctx, false, grid, k, w0, w1);
    }
    CK(cudaGetLastError());
    ctx->have_cost = true;
    return 0;
}

int hb_get_wscal_rows(hb_ctx *ctx, int col0, int ncol, int row0, int nrow, double *wscal)
{
    if (range_ok(ctx, col0, ncol)) return -1;
    if (!ctx->forcing_set) return fail("hb_get_wscal: forcing not set");
    if (!wscal) return fail("hb_get_wscal: null argument");
    if (row0 < 0 || nrow < 0 || row0 + nrow > ctx->d.T) return fail("hb_get_wscal: row range out of bounds");
    CK(cudaSetDevice(ctx->device));
    if (ncol > 0 && nrow > 0) {
        const size_t pitch = (size_t)ctx->d.T * HB_NWSCAL * sizeof(double), width = (size_t)nrow * HB_NWSCAL * sizeof(double);
        CK(cudaMemcpy2DAsync(wscal, width, ctx->d.wscal + ((size_t)col0 * ctx->d.T + row0) * HB_NWSCAL, pitch, width,
                             (size_t)ncol, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return finish(ctx);
}

int hb_get_wscal(hb_ctx *ctx, int col0, int ncol, double *wscal)
{
    if (!ctx) return fail("null ctx");
    return hb_get_wscal_rows(ctx, col0, ncol, 0, ctx->d.T, wscal);
}

int hb_get_fields_rows(hb_ctx *ctx, int col0, int ncol, int out0, int nout, double *fields)
{
    if (range_ok(ctx, col0, ncol)) return -1;
    if (!ctx->d.fields) return fail("hb_get_fields: no field output planned (hb_set_output)");
    if (!fields) return fail("hb_get_fields: null argument");
    if (out0 < 0 || nout < 0 || out0 + nout > ctx->d.n_out) return fail("hb_get_fields: output-row range out of bounds");
    CK(cudaSetDevice(ctx->device));
    if (ncol > 0 && nout > 0) {
        const size_t row_bytes = (size_t)HB_NFIELD * ctx->d.nx * sizeof(double);
        const size_t pitch = (size_t)ctx->d.n_out * row_bytes, width = (size_t)nout * row_bytes;
        CK(cudaMemcpy2DAsync(fields, width,
                             ctx->d.fields + ((size_t)col0 * ctx->d.n_out + out0) * HB_NFIELD * ctx->d.nx, pitch, width,
                             (size_t)ncol, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return finish(ctx);
}

int hb_get_fields(hb_ctx *ctx, int col0, int ncol, double *fields)
{
    if (!ctx) return fail("null ctx");
    return hb_get_fields_rows(ctx, col0, ncol, 0, ctx->d.n_out, fields);
}

int hb_get_counters(hb_ctx *ctx, int col0, int ncol, int64_t *counters)
{
    if (range_ok(ctx, col0, ncol)) return -1;
    if (!counters) return fail("hb_get_counters: null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(counters, ctx->d.counters + (size_t)col0 * HB_NCOUNTER,
                       (size_t)ncol * HB_NCOUNTER * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    return finish(ctx);
}

int hb_reset_counters(hb_ctx *ctx)
{
    if (!ctx) return fail("null ctx");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d.counters, 0, (size_t)ctx->d.n_cols * HB_NCOUNTER * sizeof(long long), ctx->stream));
    return 0;
}

int hb_set_fast_mode(hb_ctx *ctx, int flags)
{
    if (!ctx) return fail("null ctx");
    if (flags & ~(HB_FAST_CONT | HB_FAST_JAC3 | HB_FAST_F32)) return fail("hb_set_fast_mode: unknown flag");
    ctx->fast_flags = flags;
    return 0;
}

int hb_set_schedule(hb_ctx *ctx, int adaptive)
{
    if (!ctx) return fail("null ctx");
    ctx->adaptive = adaptive != 0;
    ctx->have_cost = false;
    return 0;
}

int hb_moments_dev(hb_ctx *ctx, double *count_dev, double *wtd_sum_dev, double *wtd_sumsq_dev,
                   double *theta_sum_dev, double *theta_sumsq_dev)
{
    if (!ctx || !ctx->forcing_set) return fail("hb_moments_dev: forcing not set");
    if (!wtd_sum_dev || !wtd_sumsq_dev) return fail("hb_moments_dev: null argument");
    if ((theta_sum_dev == nullptr) != (theta_sumsq_dev == nullptr)) return fail("hb_moments_dev: give both theta buffers or none");
    if (theta_sum_dev && !ctx->d.fields) return fail("hb_moments_dev: no field output planned (hb_set_output)");
    const DevCtx &d = ctx->d;
    CK(cudaSetDevice(ctx->device));
    const int M_theta = theta_sum_dev ? d.n_out * d.nx : 0, M_max = d.T > M_theta ? d.T : M_theta;
    // enough column slices to fill the GPU, at least 64 columns each
    int n_slices = (4 * ctx->sms * 256 + M_max - 1) / M_max;
    const int max_slices = (d.n_cols + 63) / 64;
    if (n_slices > max_slices) n_slices = max_slices;
    if (n_slices < 1) n_slices = 1;
    const int per = (d.n_cols + n_slices - 1) / n_slices;
    n_slices = (d.n_cols + per - 1) / per;
    double *part;
    if (slot_alloc(ctx, ctx->s_mom, &part, (size_t)n_slices * 3 * M_max, false)) return -1;
    {
        const dim3 grid((d.T + 255) / 256, n_slices);
        moments_partial_kernel<false><<<grid, 256, 0, ctx->stream>>>(d, d.T, per, part);
        moments_final_kernel<<<(d.T + 255) / 256, 256, 0, ctx->stream>>>(part, d.T, n_slices, count_dev, wtd_sum_dev, wtd_sumsq_dev);
        ctx->launches += 2;
    }
    if (theta_sum_dev) {
        const dim3 grid((M_theta + 255) / 256, n_slices);
        moments_partial_kernel<true><<<grid, 256, 0, ctx->stream>>>(d, M_theta, per, part);
        moments_final_kernel<<<(M_theta + 255) / 256, 256, 0, ctx->stream>>>(part, M_theta, n_slices, nullptr, theta_sum_dev, theta_sumsq_dev);
        ctx->launches += 2;
    }
    CK(cudaGetLastError());
    return 0;
}

int64_t hb_launch_count(hb_ctx *ctx) { return ctx ? ctx->launches : -1; }

int hb_get_phase_cycles(hb_ctx *ctx, int64_t *out16)
{
    if (!ctx || !out16) return fail("hb_get_phase_cycles: null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out16, ctx->d.phase_cycles, 16 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemsetAsync(ctx->d.phase_cycles, 0, 16 * sizeof(long long), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

double hb_measure_fp64_peak(hb_ctx *ctx)
{
    if (!ctx) { fail("null ctx"); return -1.0; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) { fail("cudaSetDevice"); return -1.0; }
    const int blocks = ctx->sms * 8, threads = 256, iters = 1 << 16;
    double *out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) { fail("cudaMalloc"); return -1.0; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, ctx->stream);
        fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
        ctx->launches++;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) { fail("fp64 peak kernel failed"); return -1.0; }
    return best;
}

// ---- hooks -----------------------------------------------------------------------
static int run_hook(hb_ctx *ctx, HookArgs h)
{
    h.io = ctx->hook_io; h.stats = ctx->hook_stats;
    DevCtx k = ctx->d;
    if (h.mode == HK_WINDOW) k.t_acc = ctx->hook_io + (size_t)7 * ctx->d.nx;     // accepted step times
    DISPATCH(ctx, hook_kernel, 1, k, h);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    return 0;
}

static int hook_in(hb_ctx *ctx, int slot, const double *src)
{
    CK(cudaMemcpyAsync(ctx->hook_io + (size_t)slot * ctx->d.nx, src, ctx->d.nx * sizeof(double),
                       cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

static int hook_out(hb_ctx *ctx, int slot, double *dst, int n)
{
    CK(cudaMemcpyAsync(dst, ctx->hook_io + (size_t)slot * ctx->d.nx, n * sizeof(double),
                       cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int hb_rhs(hb_ctx *ctx, int col, const double *psi, const double *nrnd, double precip, int wtd,
           int daylight, int spinup, double atm, double *dydt, double *lateral_flow,
           double *transpiration)
{
    if (ready(ctx) || range_ok(ctx, col, 1)) return -1;
    CK(cudaSetDevice(ctx->device));
    if (hook_in(ctx, 0, psi) || hook_in(ctx, 1, nrnd)) return -1;
    HookArgs h{}; h.mode = HK_RHS; h.col = col; h.atm = atm; h.precip = precip; h.wtd = wtd; h.daylight = daylight; h.spinup = spinup;
    if (run_hook(ctx, h)) return -1;
    if (hook_out(ctx, 2, dydt, ctx->d.nx)) return -1;
    double two[2];
    if (hook_out(ctx, 3, two, 2)) return -1;
    if (lateral_flow) *lateral_flow = two[0];
    if (transpiration) *transpiration = two[1];
    return 0;
}

int hb_closure_nodes(hb_ctx *ctx, int col, const double *psi, const double *nrnd, double *out)
{
    if (ready(ctx) || range_ok(ctx, col, 1)) return -1;
    CK(cudaSetDevice(ctx->device));
    if (hook_in(ctx, 0, psi) || hook_in(ctx, 1, nrnd)) return -1;
    HookArgs h{}; h.mode = HK_NODES; h.col = col;
    if (run_hook(ctx, h)) return -1;
    return hook_out(ctx, 2, out, 4 * ctx->d.nx);
}

int hb_num_jac(hb_ctx *ctx, int col, const double *psi, const double *nrnd, double precip, int wtd,
               int daylight, int spinup, const double *fac_in, double *dl, double *d, double *du,
               double *fac_out)
{
    if (ready(ctx) || range_ok(ctx, col, 1)) return -1;
    CK(cudaSetDevice(ctx->device));
    if (hook_in(ctx, 0, psi) || hook_in(ctx, 1, nrnd)) return -1;
    if (fac_in && hook_in(ctx, 2, fac_in)) return -1;
    HookArgs h{}; h.mode = HK_JAC; h.col = col; h.precip = precip; h.wtd = wtd; h.daylight = daylight;
    h.spinup = spinup; h.first = fac_in ? 0 : 1;
    if (run_hook(ctx, h)) return -1;
    if (hook_out(ctx, 3, dl, ctx->d.nx) || hook_out(ctx, 4, d, ctx->d.nx) ||
        hook_out(ctx, 5, du, ctx->d.nx) || hook_out(ctx, 6, fac_out, ctx->d.nx)) return -1;
    return 0;
}

int hb_tri_solve(hb_ctx *ctx, const double *dl, const double *d, const double *du, double c,
                 const double *b, double *x)
{
    if (ready(ctx)) return -1;
    CK(cudaSetDevice(ctx->device));
    std::vector<double> zero(ctx->d.nx, 0.0);
    if (hook_in(ctx, 0, zero.data()) || hook_in(ctx, 1, zero.data()) || hook_in(ctx, 2, dl) ||
        hook_in(ctx, 3, d) || hook_in(ctx, 4, du) || hook_in(ctx, 5, b)) return -1;
    HookArgs h{}; h.mode = HK_TRI; h.col = 0; h.c = c;
    if (run_hook(ctx, h)) return -1;
    if (tuning_env("HB_TRI_DEBUG")) {            // cycles: factor, then four repeated solves
        double t[5];
        CK(cudaMemcpy(t, ctx->hook_io + (size_t)7 * ctx->d.nx, sizeof(t), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[hb_tri_solve] cycles: factor %.0f, solve %.0f %.0f %.0f %.0f\n", t[0], t[1], t[2], t[3], t[4]);
    }
    return hook_out(ctx, 6, x, ctx->d.nx);
}

int hb_find_wtd(hb_ctx *ctx, int col, const double *psi, int32_t *wtd_node)
{
    if (ready(ctx) || range_ok(ctx, col, 1)) return -1;
    if (!psi || !wtd_node) return fail("hb_find_wtd: null argument");
    CK(cudaSetDevice(ctx->device));
    if (hook_in(ctx, 0, psi) || hook_in(ctx, 1, psi)) return -1;
    HookArgs h{}; h.mode = HK_WTD; h.col = col;
    if (run_hook(ctx, h)) return -1;
    long long w = 0;
    CK(cudaMemcpy(&w, ctx->hook_stats, sizeof(w), cudaMemcpyDeviceToHost));
    *wtd_node = (int32_t)w;
    return 0;
}

int hb_philox_normals(hb_ctx *ctx, uint64_t seed, int64_t global_column, int window, int n, double *out)
{
    if (!ctx || !out || n < 1) return fail("hb_philox_normals: bad argument");
    CK(cudaSetDevice(ctx->device));
    double *d = nullptr;
    CK(cudaMalloc(&d, (size_t)n * sizeof(double)));
    philox_probe_kernel<<<ctx->sms * 4, 256, 0, ctx->stream>>>(seed, global_column, window, n, d);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(out, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(std::string("hb_philox_normals: ") + cudaGetErrorString(e));
    return 0;
}

int hb_solve_window(hb_ctx *ctx, int col, double t0, double t1, const double *y0, double *nrnd,
                    double precip, int wtd, int daylight, int spinup, double atm, double *y_end,
                    int64_t *stats, double *lateral_flow, double *transpiration, int *status,
                    double *t_accepted)
{
    if (ready(ctx) || range_ok(ctx, col, 1)) return -1;
    CK(cudaSetDevice(ctx->device));
    if (hook_in(ctx, 0, y0) || hook_in(ctx, 1, nrnd)) return -1;
    {
        const int fl = ctx->h_flags.empty() ? 0 : ctx->h_flags[col];
        if ((fl & HB_FLAG_PREDICT) && (fl & HB_FLAG_LF)) return fail("PREDICT lateral-flow mode is not implemented");
    }
    HookArgs h{}; h.mode = HK_WINDOW; h.col = col; h.atm = atm; h.precip = precip; h.wtd = wtd; h.daylight = daylight;
    h.spinup = spinup; h.t0 = t0; h.t1 = t1;
    if (run_hook(ctx, h)) return -1;
    if (hook_out(ctx, 2, y_end, ctx->d.nx) || hook_out(ctx, 1, nrnd, ctx->d.nx)) return -1;
    double two[2];
    if (hook_out(ctx, 3, two, 2)) return -1;
    if (lateral_flow) *lateral_flow = two[0];
    if (transpiration) *transpiration = two[1];
    long long st[HB_NCOUNTER + 2];
    CK(cudaMemcpy(st, ctx->hook_stats, sizeof(st), cudaMemcpyDeviceToHost));
    if (stats) for (int k = 0; k < HB_NCOUNTER; ++k) stats[k] = st[k];
    if (status) *status = (int)st[HB_NCOUNTER];
    if (t_accepted) {                       // accepted step times of the LAST attempt (sol.t[1:])
        const long long n = st[HB_C_NSTEPS] < ctx->d.nx ? st[HB_C_NSTEPS] : ctx->d.nx;
        if (n > 0 && hook_out(ctx, 7, t_accepted, (int)n)) return -1;
    }
    return 0;
}

}  // extern "C"
