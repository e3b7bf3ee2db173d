// hb_kernels.cuh -- the kernels of the path; included by hb_device.cuh once per instance
// namespace, right after hb_device_impl.cuh.  No include guard on purpose.
namespace HB_IMPL_NS {
using namespace hb;

__device__ inline void add_counters(long long *dst, const long long *k, long long windows, bool atomic = true)
{
    const long long v[HB_NCOUNTER] = {k[K_NFEV], k[K_NJEV], k[K_NLU], k[K_NSTEPS], k[K_NREJECT],
                                      k[K_NFAIL], k[K_NRHS], k[K_NNEWTON], windows};
    for (int i = 0; i < HB_NCOUNTER; ++i) {
        if (atomic) atomicAdd(reinterpret_cast<unsigned long long *>(dst) + i, (unsigned long long)v[i]);
        else dst[i] += v[i];
    }
}

// Closure on the nodes at the state in V_Y with the noise in V_NR (the
// diagnostic call of simulation.py:564,623: node k uses n_rnd[k]).
template <int NT, int NV>
__device__ inline void nodes_diag(const DevCtx &c, int col, double *th, double *K, double *C, double *kb)
{
    const int nx = c.nx;
    const double *st = c.stat + (size_t)col * HB_NSTATIC * nx;
    for (int k = threadIdx.x; k < nx; k += NT) {
        Clos r;
        closure_cold(HB_XA(c, NV, X_PAR), HB_VEC(c, V_Y)[k], st[(size_t)HB_S_PHI_NODE * nx + k],
                     st[(size_t)HB_S_KMEAN_NODE * nx + k], st[(size_t)HB_S_COEF_NODE * nx + k],
                     HB_VEC(c, V_NR)[k], &r);
        th[k] = r.theta; K[k] = r.K; kb[k] = r.kb;
        if (C) C[k] = r.C;
    }
}

// Diagnostics of one row: estimated water table, lateral flow of the last RHS
// evaluation (arg_out, quirk Q6) and, on output rows, the closure on the nodes
// (simulation.py:612-630).  State is in V_Y.
template <int NT, int NV>
__device__ inline void write_row(const DevCtx &c, int &parity, int col, int row, int status, bool solved)
{
    const int nx = c.nx, tid = threadIdx.x;
    double *ws = c.wscal + ((size_t)col * c.T + row) * HB_NWSCAL;
    if (status == 2) {
        if (tid == 0) { ws[HB_W_WTD_EST] = 0.0; ws[HB_W_LATERAL] = 0.0; ws[HB_W_TRANSP] = 0.0; ws[HB_W_STATUS] = 2.0; }
        return;
    }
    const int wtd_est = find_wtd_nodes<NT, NV>(c, parity, HB_VEC(c, V_Y));
    if (tid == 0) {
        ws[HB_W_WTD_EST] = (double)wtd_est;
        ws[HB_W_LATERAL] = solved ? lateral_flow<NT, NV>(c) : 0.0;
        ws[HB_W_TRANSP] = solved ? HB_XA(c, NV, X_MISC)[1] : 0.0;
        ws[HB_W_STATUS] = (double)status;
    }
    if (c.fields && c.out_stride > 0 && (row % c.out_stride == 0 || row == c.T - 1)) {
        const int o = (row % c.out_stride == 0) ? row / c.out_stride : c.n_out - 1;
        double *f = c.fields + ((size_t)col * c.n_out + o) * HB_NFIELD * nx;
        for (int k = tid; k < nx; k += NT) f[(size_t)HB_F_PSI * nx + k] = HB_VEC(c, V_Y)[k];
        nodes_diag<NT, NV>(c, col, f + (size_t)HB_F_THETA * nx, f + (size_t)HB_F_K * nx, nullptr,
                           f + (size_t)HB_F_KBKG * nx);
    }
}

// Work item = (column, chunk of c.chunk consecutive rows).  Scheduling is a credit
// ("mailbox") scheme without spin-waits:
//   * tickets come from one atomic counter.  The first c.n_hot tickets name the
//     columns the previous launch found expensive (plan_kernel) and carry ALL
//     their chunks, so that such a column starts at once and is never paced; the
//     other tickets enumerate the remaining columns in chunk-major order, one
//     chunk each -- a persistent CTA that finishes early keeps pulling work and
//     the tail of the launch is one chunk long;
//   * a ticket is a credit deposited on its column (atomicAdd).  Whoever finds
//     the column idle (old credit 0) becomes its owner and integrates chunks
//     until the credit is used up; a ticket for a busy column only leaves its
//     credit behind -- the owner will run that chunk right after the current one
//     (a column that falls behind therefore runs back to back, and nobody waits).
// The chunks of a column are integrated in row order by one owner at a time; state,
// base noise field and tape position are handed over through global memory
// (release: __threadfence + atomic on the credit; acquire: the atomic + L2 loads).
template <int NT, int CPT, int MB, bool FULL, bool FAST = false>
__global__ void __launch_bounds__(NT, MB) run_kernel(const __grid_constant__ DevCtx c, int w0, int w1)
{
    constexpr int NV = FULL ? NVEC_FULL : NVEC_BASE;
    __shared__ int s_msg[2];
    const int nx = c.nx, tid = threadIdx.x;
    double *nr = HB_VEC(c, V_NR), *y = HB_VEC(c, V_Y);
    const int n_chunks = (w1 - w0 + c.chunk - 1) / c.chunk;
    const int n_hot = c.order ? __ldcg(c.n_hot) : 0, n_reg = c.n_cols - n_hot;
    const long long n_tickets = (long long)n_hot + (long long)n_reg * n_chunks;
    int parity = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            int col = -1;
            const long long ticket = (long long)atomicAdd(reinterpret_cast<unsigned long long *>(c.queue), 1ull);
            if (ticket >= n_tickets) col = -2;
            else {
                int add = 1, pos;
                if (ticket < n_hot) { pos = (int)ticket; add = n_chunks; }
                else pos = n_hot + (int)((ticket - n_hot) % n_reg);
                const int cand = c.order ? __ldcg(c.order + pos) : pos;
                if (atomicAdd(c.credit + cand, add) == 0) { __threadfence(); col = cand; }
            }
            s_msg[0] = col;
        }
        __syncthreads();
        const int col = s_msg[0];
        if (col == -2) break;
        if (col < 0) continue;                 // column busy: its owner takes the chunk
        bind_column<NT, FULL>(c, col);
        double *psi_g = c.psi + (size_t)col * nx;
        double *nb = c.noise_base + (size_t)col * nx;
        const int ser = c.series ? c.series[col] : 0;
        bool more = true;
        while (more) {
            const int chunk = __ldcg(c.done + col);
            const int r0 = w0 + chunk * c.chunk, r1 = min(r0 + c.chunk, w1);
            __syncthreads();
            if (tid < 8) reinterpret_cast<long long *>(HB_XA(c, NV, X_CNT))[tid] = 0;
            // L2 loads: another SM may have written the state and (after a failed solve)
            // the damped base field
            for (int i = tid; i < nx; i += NT) { nr[i] = __ldcg(nb + i); y[i] = __ldcg(psi_g + i); }
            bool nr_is_base = true;
            long long windows = 0;
            Warm warm = {0.0, 1, 0, false, false};  // fast mode: nothing is carried into a chunk
            double prev_p = 0.0, prev_atm = 0.0;
            int prev_mode = -1;
            __syncthreads();
            for (int row = r0; row < r1; ++row) {
                int status = 0;
                bool solved = false;
                if (row > 0) {
                    const int wtd = c.wtd_idx[(size_t)ser * c.T + row];
                    if (wtd < 0) { status = 2; if (FAST) warm.valid = false; }
                    else {
                        const double p = HB_XA(c, NV, X_PAR)[HB_P_PRECIP_MULT] * c.precip[(size_t)ser * c.T + row];
                        const bool refresh = (p > 0.5) || (row % 48 == 0);   // simulation.py:599
                        __syncthreads();
                        if (refresh) {
                            if (c.noise_mode == HB_NOISE_TAPE) {
                                const int slot = __ldcg(c.slot_next + col);          // uniform read
                                if (slot >= c.n_slots && tid == 0) atomicOr(c.err_flags, HB_ERR_TAPE_SHORT);
                                const double *src = c.tape + ((size_t)col * c.n_slots + min(slot, c.n_slots - 1)) * nx;
                                for (int i = tid; i < nx; i += NT) nr[i] = src[i];
                                __syncthreads();
                                if (tid == 0) __stcg(c.slot_next + col, slot + 1);
                            } else {
                                for (int i = tid; i < nx; i += NT)
                                    nr[i] = philox_normal(c.seed, c.col_offset + col, row, i);
                            }
                            nr_is_base = false;
                        } else if (!nr_is_base) {
                            for (int i = tid; i < nx; i += NT) nr[i] = __ldcg(nb + i);
                            nr_is_base = true;
                        }
                        __syncthreads();
                        Args a;
                        a.precip = p; a.wtd = wtd;
                        a.mode = c.daylight[(size_t)ser * c.T + row] ? 1 : 0;
                        a.atm = c.atm ? c.atm[(size_t)ser * c.T + row] : 0.0;
                        bool damped = false;
                        if (FAST) {
                            // same function of (t, psi) as in the previous window: same row values and the
                            // same noise field (both windows on the base field, no refresh in between)
                            // (a new water-table observation only moves the end of the lateral-flow sink
                            // range, a term of size 2.5e-4 |psi|: the error control takes care of it)
                            warm.cont = !refresh && nr_is_base && p == prev_p && a.mode == prev_mode &&
                                        a.atm == prev_atm && (c.fast & HB_FAST_CONT);
                            prev_p = refresh ? -1.0 : p; prev_mode = a.mode; prev_atm = a.atm;
                        }
                        status = solve_window<NT, CPT, FULL, FAST>(c, col, a, (double)(row - 1), (double)row, psi_g,
                                                                   damped, &warm);
                        if (FAST && (status != 0 || damped)) warm.valid = false;
                        if (damped && nr_is_base)
                            for (int i = tid; i < nx; i += NT) __stcg(nb + i, nr[i]);
                        for (int i = tid; i < nx; i += NT) __stcg(psi_g + i, y[i]);
                        windows++; solved = true;
                    }
                }
                write_row<NT, NV>(c, parity, col, row, status, solved);
            }
            __syncthreads();
            if (tid == 0) {
                const long long *k = reinterpret_cast<long long *>(HB_XA(c, NV, X_CNT));
                add_counters(c.counters + (size_t)col * HB_NCOUNTER, k, windows);
                if (c.cost) c.cost[col] += k[K_NRHS];          // owner-only: plain update
#if HB_PHASE_TIMING
                for (int q = 0; q < 15; ++q)
                    atomicAdd(reinterpret_cast<unsigned long long *>(c.phase_cycles) + q,
                              (unsigned long long)reinterpret_cast<long long *>(HB_XA(c, NV, X_TIME))[q]);
#endif
                __stcg(c.done + col, chunk + 1);
                __threadfence();                   // state, noise, counters, done before the hand-over
                s_msg[1] = atomicSub(c.credit + col, 1) > 1;     // credit left: next chunk, same owner
            }
            __syncthreads();
            more = s_msg[1] != 0;
        }
    }
}

template <int NT, int CPT, int MB, bool FULL>
__global__ void __launch_bounds__(NT, MB) spinup_kernel(const __grid_constant__ DevCtx c, int max_iter,
                                                         const double *nrnd0, const double *zwtd0,
                                                         int *iters_out)
{
    constexpr int NV = FULL ? NVEC_FULL : NVEC_BASE;
    __shared__ int s_col;
    const int nx = c.nx, tid = threadIdx.x;
    double *nr = HB_VEC(c, V_NR), *y = HB_VEC(c, V_Y);
    int parity = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_col = atomicAdd(c.queue, 1);
        __syncthreads();
        const int col = s_col;
        if (col >= c.n_cols) break;
        bind_column<NT, FULL>(c, col);
        double *psi_g = c.psi + (size_t)col * nx;
        for (int i = tid; i < nx; i += NT)
            nr[i] = nrnd0 ? nrnd0[(size_t)col * nx + i]
                          : philox_normal(c.seed, c.col_offset + col, -1, i);
        const int ser = c.series ? c.series[col] : 0;
        __syncthreads();
        Args a;
        a.precip = HB_XA(c, NV, X_PAR)[HB_P_PRECIP_MULT] * c.precip[(size_t)ser * c.T];
        a.wtd = c.wtd_idx[(size_t)ser * c.T]; a.mode = 2; a.atm = 0.0;
        const double dz = HB_XA(c, NV, X_PAR)[HB_P_DZ];
        int it = 0;
        for (it = 0; it < max_iter; ++it) {
            bool damped = false;
            solve_window<NT, CPT, FULL>(c, col, a, 0.0, 1.0, psi_g, damped);
            const int wtd_est = find_wtd_nodes<NT, NV>(c, parity, y);
            const double abs_err = fabs(zwtd0[col] - c.x[wtd_est]);
            double loc = 0.0;
            for (int i = tid; i < nx; i += NT) { const double dl = y[i] - psi_g[i]; loc += dl * dl; }
            const double mse = block_sum<NT, NV>(c, parity, loc) / (double)nx;
            for (int i = tid; i < nx; i += NT) psi_g[i] = y[i];
            if (abs_err <= 2.0 * dz && mse <= 0.01) { it++; break; }
        }
        __syncthreads();
        if (tid == 0) {
            add_counters(c.counters + (size_t)col * HB_NCOUNTER,
                         reinterpret_cast<long long *>(HB_XA(c, NV, X_CNT)), it);
            if (iters_out) iters_out[col] = it;
        }
    }
}



template <int NT, int CPT, int MB, bool FULL>
__global__ void __launch_bounds__(NT, MB) hook_kernel(const __grid_constant__ DevCtx c, HookArgs h)
{
    constexpr int NV = FULL ? NVEC_FULL : NVEC_BASE;
    const int nx = c.nx, tid = threadIdx.x;
    bind_column<NT, FULL>(c, h.col);
    double *io = h.io;
    double *y = HB_VEC(c, V_Y), *nr = HB_VEC(c, V_NR), *b = HB_VEC(c, V_B);
    Args a;
    a.precip = h.precip; a.wtd = h.wtd; a.mode = h.spinup ? 2 : (h.daylight ? 1 : 0); a.atm = h.atm;
    int parity = 0;
    for (int i = tid; i < nx; i += NT) { y[i] = io[i]; nr[i] = io[nx + i]; }
    __syncthreads();
    if (h.mode == HK_RHS) {
        const Statics S = statics_of<FULL>(c, h.col);
        set_window_args<NV>(c, a, 0.0, 1.0);
        rhs_eval<NT, CPT, NV>(c, S, h.col, V_B, parity);
        for (int i = tid; i < nx; i += NT) io[2 * nx + i] = b[i];
        __syncthreads();
        if (tid == 0) { io[3 * nx] = lateral_flow<NT, NV>(c); io[3 * nx + 1] = HB_XA(c, NV, X_MISC)[1]; }
    } else if (h.mode == HK_NODES) {
        nodes_diag<NT, NV>(c, h.col, io + 2 * nx, io + 3 * nx, io + 4 * nx, io + 5 * nx);
    } else if (h.mode == HK_JAC) {
        // the production state machine, stopped right after the initial Jacobian
        if (!h.first) for (int i = tid; i < nx; i += NT) HB_GROW(c, G_FAC)[i] = io[2 * nx + i];
        set_window_args<NV>(c, a, 0.0, 1.0);
        bdf_solve<NT, CPT, FULL>(c, h.col, io, h.first ? 1 : 2);
        for (int i = tid; i < nx; i += NT) {
            io[3 * nx + i] = HB_VEC(c, V_JL)[i]; io[4 * nx + i] = HB_VEC(c, V_JD)[i];
            io[5 * nx + i] = HB_VEC(c, V_JU)[i]; io[6 * nx + i] = HB_GROW(c, G_FAC)[i];
        }
    } else if (h.mode == HK_TRI) {
        for (int i = tid; i < nx; i += NT) {
            HB_VEC(c, V_JL)[i] = io[2 * nx + i]; HB_VEC(c, V_JD)[i] = io[3 * nx + i];
            HB_VEC(c, V_JU)[i] = io[4 * nx + i]; b[i] = io[5 * nx + i];
        }
        __syncthreads();
        // factor once, solve; the solve is repeated (same input) to expose cold-code
        // effects: cycles of repetition r go to io[7*nx + r] (diagnostic)
        long long t_f0 = clock64();
        if (tid < 32) {
            tri_factor_warp0<NV, FULL>(c, h.c);
        }
        long long t_f1 = clock64();
        for (int rep = 0; rep < 4; ++rep) {
            __syncthreads();
            for (int i = tid; i < nx; i += NT) b[i] = io[5 * nx + i];
            __syncthreads();
            const long long t0 = clock64();
            if (tid < 32) {
                tri_solve_warp0<NV, FULL>(c, h.c);
            }
            const long long t1 = clock64();
            if (tid == 0) io[7 * nx + 1 + rep] = (double)(t1 - t0);
        }
        if (tid == 0) io[7 * nx] = (double)(t_f1 - t_f0);
        __syncthreads();
        for (int i = tid; i < nx; i += NT) io[6 * nx + i] = tri_x<NV, FULL>(c, i);
    } else if (h.mode == HK_WTD) {
        // find_wtd(y >= psi_sat) on the nodes (utilities.py:56-99) for the state in io[0..nx)
        const int w = find_wtd_nodes<NT, NV>(c, parity, y);
        if (tid == 0) h.stats[0] = w;
    } else if (h.mode == HK_WINDOW) {
        if (tid < 8) reinterpret_cast<long long *>(HB_XA(c, NV, X_CNT))[tid] = 0;
        __syncthreads();
        bool damped = false;
        const int st = solve_window<NT, CPT, FULL>(c, h.col, a, h.t0, h.t1, io, damped);
        for (int i = tid; i < nx; i += NT) { io[2 * nx + i] = y[i]; io[nx + i] = nr[i]; }
        __syncthreads();
        if (tid == 0) {
            io[3 * nx] = lateral_flow<NT, NV>(c);
            io[3 * nx + 1] = HB_XA(c, NV, X_MISC)[1];
            long long z[HB_NCOUNTER] = {0};
            add_counters(z, reinterpret_cast<long long *>(HB_XA(c, NV, X_CNT)), 1, false);
            for (int k = 0; k < HB_NCOUNTER; ++k) h.stats[k] = z[k];
            h.stats[HB_NCOUNTER] = st;
        }
    }
}

}  // namespace HB_IMPL_NS
