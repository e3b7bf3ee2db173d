/*
 * hb_layout.h -- flat data layout shared by the CUDA kernels, the C-ABI
 * (hydrob200.h), the Python host layer (hydromodel_b200/pack.py) and the CPU
 * oracle (oracle/hb_oracle.c).
 *
 * One "column" is one 1-D soil column of the reference (one Simulation object,
 * code/src/simulation.py:79-387).  Everything the reference recomputes on every
 * right-hand-side call from static objects (porosity interpolants,
 * porosity.py:184-206; layer means, vrettas_fung.py:22-49,132-190) is packed
 * once on the host into per-cell arrays.
 */
#ifndef HB_LAYOUT_H
#define HB_LAYOUT_H

/* ---- per-column scalar block: double[HB_NPARAM] ------------------------- */
enum {
    HB_P_N = 0,         /* van Genuchten n            soil_properties.py:27    */
    HB_P_M,             /* m = 1 - 1/n                soil_properties.py:82-90 */
    HB_P_ALPHA,         /* alpha (a0)                 soil_properties.py:36    */
    HB_P_PSI_SAT,       /* min(psi_sat, 0)            soil_properties.py:45    */
    HB_P_EPS,           /* max(epsilon, 1e-8)         hydrological_model.py:32 */
    HB_P_THETA_RES,     /* residual water content                              */
    HB_P_LAMBDA,        /* K = Se^lambda * K_bkg      vrettas_fung.py:229      */
    HB_P_SIGMA,         /* noise amplitude            vrettas_fung.py:160      */
    HB_P_KSAT_SOIL,     /* saturated K of the soil layer                       */
    HB_P_DZ,            /* model dz used by q_inf_max vrettas_fung.py:254      */
    HB_P_INTERCEPT,     /* Interception_pct           richards_pde.py:449      */
    HB_P_SURF_EVAP,     /* surface evaporation        simulation.py:352        */
    HB_P_MODEL,         /* 0 = vrettas_fung, 1 = vanGenuchten                  */
    HB_P_PRECIP_MULT,   /* ensemble precipitation multiplier (1 = reference)   */
    HB_P_PHI_TOP,       /* porosity at node 0 (boundary closure)               */
    HB_P_KMEAN_TOP,     /* layer-mean K_bkg at z = x[0]                        */
    HB_P_COEF_TOP,      /* noise coefficient of the layer of z = x[0]          */
    HB_P_DSMIN_ALL,     /* min_j (phi_j - theta_res), broadcast quirk Q2       */
    HB_P_DSMIN_EXCL0,   /* same without node 0                                 */
    HB_P_FLAGS,         /* HB_FLAG_* bitmask stored as a double                */
    HB_P_ROOT_CELLS_TOP,/* 1 if the top half-cell midpoint is in the root zone */
    HB_P_LAI,           /* leaf area index            simulation.py:343        */
    HB_P_IPSI50,        /* 1/psi_50                   simulation.py:339        */
    HB_P_RESERVED0,
    HB_NPARAM
};

/* Simulation_Flags (input_parameters.json:61-67).  SPINUP is a run mode, not a
 * stored flag: it is toggled by the caller (simulation.py:398,489). */
#define HB_FLAG_LF       1
#define HB_FLAG_ET       2
#define HB_FLAG_HLIFT    4
#define HB_FLAG_PREDICT  8

/* ---- per-column per-cell static arrays: double[HB_NSTATIC][nx] ----------
 * "mid" arrays are indexed by midpoint k = 0..nx-2 (between nodes k, k+1),
 * entry nx-1 is padding.  coef = 0.05 / 0.10 / 1.0 for soil / saprolite /
 * weathered bedrock (vrettas_fung.py:154,172,190); coef < 0 marks a cell
 * outside every layer, where K_bkg stays Ksat_soil (vrettas_fung.py:143). */
enum {
    HB_S_PHI_MID = 0,
    HB_S_KMEAN_MID,
    HB_S_COEF_MID,
    HB_S_PHI_NODE,
    HB_S_KMEAN_NODE,
    HB_S_COEF_NODE,
    HB_S_ROOT_MID,      /* root pdf at midpoints, 0 outside the root zone */
    HB_S_FCAP_MID,      /* field capacity at midpoints                    */
    HB_S_WILT_MID,      /* wilting point at midpoints                     */
    HB_NSTATIC
};

/* ---- per-window forcing: one record per CSV row ------------------------- */
/* precip[T] (cm), wtd_idx[T] (node index of the observed water table, -1 if
 * it is not on the grid: window skipped, simulation.py:582-588),
 * daylight[T] (1 if hour in [6,18), richards_pde.py:230), atm[T]. */

/* ---- per-window outputs -------------------------------------------------- */
enum {                       /* columns of the per-window scalar record     */
    HB_W_WTD_EST = 0,        /* find_wtd(psi >= psi_sat)  simulation.py:612  */
    HB_W_LATERAL,            /* arg_out["lateral_flow"]   simulation.py:629  */
    HB_W_TRANSP,             /* arg_out["transpiration"]  simulation.py:630  */
    HB_W_STATUS,             /* 0 ok, 1 solver failed 5x (Q5), 2 skipped     */
    HB_NWSCAL
};

enum {                       /* per-column counters (int64)                 */
    HB_C_NFEV = 0,           /* scipy nfev (RHS evals outside num_jac)       */
    HB_C_NJEV,
    HB_C_NLU,
    HB_C_NSTEPS,             /* accepted steps                               */
    HB_C_NREJECT,            /* error-test + convergence rejections          */
    HB_C_NFAIL,              /* failed solve_ivp attempts                    */
    HB_C_NRHS,               /* every RHS evaluation incl. FD Jacobian       */
    HB_C_NNEWTON,            /* Newton iterations (= tridiagonal solves)     */
    HB_C_NWINDOWS,           /* windows integrated                           */
    HB_NCOUNTER
};

enum {                       /* output field ids for hb_get_field           */
    HB_F_PSI = 0, HB_F_THETA, HB_F_K, HB_F_KBKG, HB_NFIELD
};

/* noise modes */
#define HB_NOISE_TAPE    0   /* host-generated N(0,1) vectors (parity mode)  */
#define HB_NOISE_PHILOX  1   /* on-device Philox4x32-10 + Box-Muller         */

/* Fast mode (hb_set_fast_mode; never the default -- results then differ from the
 * reference's arithmetic within the integrator's own tolerance).
 * Any flag selects the fast build of the window loop. */
#define HB_FAST_CONT     1   /* a window whose forcing row (precipitation, day/night, demand -- the
                              * water-table observation may differ) and noise field equal the previous window's is
                              * integrated as the CONTINUATION of that window: order, step size,
                              * difference table and Jacobian carry on and the steps are clipped at
                              * the window ends only.  scipy restarts at every solve_ivp call
                              * (select_initial_step common.py:63-134, order 1 and a new Jacobian
                              * bdf.py:198-241); windows with new forcing still restart that way */
#define HB_FAST_JAC3     2   /* finite-difference Jacobian with the 3 colours j mod 3 instead of
                              * scipy's 5 groups, and without the 10x-step retry pass
                              * (common.py:268-455, _numdiff.py:227-284) */
#define HB_FAST_F32      4   /* the closure of the right-hand-side evaluations INSIDE the Newton iteration
                              * (bdf.py:41-67; 38 of the 46 evaluations of a window) in single precision
                              * with the hardware exp2 / log2: theta, K, C carry a relative error of ~1e-6,
                              * far below the Newton tolerance of 0.03 in units of atol + rtol |psi|.  The
                              * evaluations that feed differences (select_initial_step, the finite-difference
                              * Jacobian) stay in double precision (vrettas_fung.py:108-249) */

/* integrator constants, scipy/integrate/_ivp/bdf.py:12-15,244-247 */
#define HB_MAX_ORDER       5
#define HB_NEWTON_MAXITER  4
#define HB_MIN_FACTOR      0.2
#define HB_MAX_FACTOR      10.0
#define HB_N_TRIALS        5      /* richards_pde.py:493 */
#define HB_NOISE_DAMP      0.8    /* richards_pde.py:522 */
#define HB_BURN_IN         1500   /* simulation.py:420   */

#endif /* HB_LAYOUT_H */
