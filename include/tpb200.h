/* tpb200.h - C ABI of libtpb200.so, the B200-native hot path of thermalporous.
 *
 * The reference (tlroy/thermalporous) is pure Python on Firedrake/PETSc; its hot path is
 * everything that runs inside one `self.solver.solve()` (thermalporous/thermalmodel.py:165):
 * residual + Jacobian assembly of the DG0/TPFA forms (singlephase.py:60-273,
 * twophase.py:67-411), the PC set-up (preconditioners.py:875-878, 1545-1548) and the
 * CPR/CPTR-preconditioned (F)GMRES solve driven by SNES newtonls.  Each entry point below
 * names the reference interface it stands in for.  A maintainer of the reference binds
 * these with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - all pointers are raw addresses; `on_device` arguments say whether they are host or
 *     CUDA device addresses (device pointers must belong to the handle's device).
 *   - fp64 everywhere.  cell index c = i + nx*(j + ny*k) over the LOCAL slab.
 *   - state / vectors are field-major SoA: v[f*ncell + c], f in (p, T) or (p, T, S_o)
 *     (the field order of twophase.py:99 and of Firedrake's mixed-space aij matrix).
 *   - Jacobian values use the block-stencil layout J[((s*nf + r)*nf + c)*ncell + cell],
 *     s in (diag, x-, x+, y-, y+, z-, z+); entries that would leave the domain are 0.
 *   - every function returns 0 on success or a negative TPB_ERR_*; solver outcomes are
 *     reported in `reason` fields using PETSc's numbering (KSP/SNESConvergedReason).
 *   - one host thread per handle; all work is queued on the handle's CUDA stream;
 *     functions that return scalars synchronise that stream.
 */
#ifndef TPB200_H
#define TPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPB_OK 0
#define TPB_ERR_ARG (-1)
#define TPB_ERR_CUDA (-2)
#define TPB_ERR_STATE (-3)
#define TPB_ERR_UNSUPPORTED (-4)
#define TPB_ERR_NCCL (-5)

typedef struct tpb_handle_s* tpb_handle;

/* structured grid of the local slab: rectanglegeo.py:28-34 (2-D), boxgeo.py:31-44 (3-D).
 * dim == 2 => nz must be 1.  Slabs are cut along the slowest axis (z in 3-D, y in 2-D). */
typedef struct {
    int dim;
    int nx, ny, nz;      /* local owned cells */
    double dx, dy, dz;
    int has_lo, has_hi;  /* 1 if another rank owns the plane below / above this slab */
} tpb_grid;

/* physicalparameters.py:9-35 (only what the forms read) */
typedef struct {
    double ko, kw, kr;
    double c_v_w, c_v_o, c_r, rho_r;
    double T_inj, T_prod;
    double API;
    double g;
    double S_o;          /* initial saturation: enters p_weight/o_weight, twophase.py:144-147 */
    double U;            /* heater coefficient, physicalparameters.py:29 */
    int gravity;         /* 1: 3-D forms carry g (always, twophase.py:317); 0 for tests */
} tpb_params;

enum tpb_field { TPB_PHI = 0, TPB_KX = 1, TPB_KY = 2, TPB_KZ = 3, TPB_KT = 4 };
enum tpb_source_kind { TPB_PROD = 0, TPB_INJ = 1, TPB_HEATER = 2 };

/* one (cell, kind) source entry = one non-zero of a reference delta Function times the cell
 * volume (wellcase.py:110-169, heatercase.py:80-118, sourceterms.py:86-153).
 * Rates follow wellcase.py:171-266 / sourceterms.py:155-269. */
typedef struct {
    int64_t cell;        /* local cell index */
    int32_t kind;        /* tpb_source_kind */
    int32_t const_rate;  /* constant_rate=True: rate = max_rate */
    double weight;       /* V_cell * delta(cell) */
    double bhp;
    double max_rate;     /* < 0 for producers (wellcase.py:100) */
} tpb_source;

/* ---- solver option surface: singlephase.py:275-444, twophase.py:413-1002 ---------------- */
enum tpb_ksp_type { TPB_KSP_GMRES = 0, TPB_KSP_FGMRES = 1 };
enum tpb_stage1 {
    TPB_S1_NONE = 0,          /* pc_ilu / pc_bilu: second stage only */
    TPB_S1_CPR = 1,           /* CPRStage1PC, preconditioners.py:335-906 */
    TPB_S1_CPTR = 2,          /* CPTRStage1PC, preconditioners.py:1243-1571 (two-phase) */
    TPB_S1_FIELDSPLIT = 3     /* single-phase pc_fieldsplit_*: Schur FULL on (p | T), no stage 2 */
};
enum tpb_decoup { TPB_DECOUP_NO = 0, TPB_DECOUP_QI = 1, TPB_DECOUP_TI = 2,
                  TPB_DECOUP_QI_TEMP = 3, TPB_DECOUP_TI_TEMP = 4 };
enum tpb_schur_pre { TPB_SCHUR_CONVDIFF = 0,  /* ConvDiffSchur(TwoPhases)PC, preconditioners.py:11-333 */
                     TPB_SCHUR_A11 = 1,       /* pc_fieldsplit_schur_precondition a11 */
                     TPB_SCHUR_DIAG = 2,      /* pc_fieldsplit_diag: additive, no coupling */
                     TPB_SCHUR_SELFP = 3 };   /* pc_fieldsplit_schur_precondition selfp (singlephase.py:322-329):
                                                 A11 - A10 diag(A00)^-1 A01 collapsed onto the 5|7-point stencil
                                                 (products that leave the stencil are lumped into the diagonal) */
enum tpb_stage2 { TPB_S2_NONE = 0, TPB_S2_ILU0 = 1, TPB_S2_BJACOBI = 2 /* per-cell block Jacobi */ };

typedef struct {
    /* SNES newtonls */
    int snes_max_it;          /* 15 single-phase (singlephase.py:293), 25 two-phase (twophase.py:424) */
    double snes_rtol, snes_atol, snes_stol;  /* PETSc defaults 1e-8, 1e-50, 1e-8 */
    int linesearch;           /* 0 basic (full step; Firedrake default), 1 backtracking */
    /* KSP */
    int ksp_type;             /* tpb_ksp_type */
    int ksp_max_it, ksp_restart;   /* 200, 200 */
    double ksp_rtol, ksp_atol;     /* 1e-5 (gmres default) | 1e-8 (twophase.py:432) */
    /* PC tree */
    int stage1, decoup, schur_pre, stage2;
    /* pressure / temperature multigrid V-cycle standing in for BoomerAMG (v_cycle dict) */
    int mg_pre, mg_post;      /* smoothing sweeps (red-black Gauss-Seidel) */
    int mg_coarse_sweeps;     /* red-black sweeps on the coarsest level */
    int mg_min_cells;         /* stop coarsening at or below this many cells */
    double mg_overcorrection; /* scaling of the piecewise-constant coarse correction */
    int mg_cycles;            /* V-cycles per application (pc_hypre_boomeramg_max_iter) */
    double mg_semi_theta;     /* an axis is coarsened on a level only if its mean coupling is at least
                                 theta * the strongest axis' (0 = always coarsen every axis) */
    int mg_full_below;        /* levels with at most this many cells coarsen every axis (0 = never) */
    double mg_dd_stop;        /* a level whose rows all satisfy sum|off-diagonals| <= mg_dd_stop * |diagonal| is the
                                 last one: it is solved by a few Gauss-Seidel sweeps (enough for a 1e-3 contraction, at
                                 most mg_coarse_sweeps) instead of being coarsened - what BoomerAMG's max_row_sum does
                                 to diagonally dominant rows.  The temperature Schur block is such a matrix while dt
                                 is small.  0 = never */
    double mg_coarse_scale;   /* Galerkin coarse operators of piecewise-constant aggregates are too stiff by a factor 2
                                 along every coarsened axis (the cell-centred multigrid scaling defect: the coarse
                                 centres are 2h apart, the summed face couplings still act over h): the couplings
                                 along a coarsened axis are multiplied by this factor, row sums are kept.  0.5 = the
                                 exact factor for smooth coefficients (default); 1 = plain Galerkin */
    int mg_smoother;          /* tpb_mg_smoother: 0 red-black point Gauss-Seidel, 1 zebra z-line Gauss-Seidel (3-D:
                                 columns coloured by (i+j)&1, every column solved exactly by the Thomas algorithm; z is
                                 then never coarsened while x or y can be) - what thin reservoir layers (Dz << Dx, Dy)
                                 need.  Default 1 in 3-D; 2-D grids always use 0 */
    int mg_tile_sweeps;       /* the z-line smoother works on tiles of columns (8 x 3 for nz <= 85) that live in one thread
                                 block's shared memory: Gauss-Seidel inside a tile, the columns around it frozen (block
                                 Jacobi between tiles, as hypre's hybrid smoother is between processes).  A tile does this
                                 many sweeps before the tiles exchange their rims (one kernel launch per exchange);
                                 0 = all sweeps of a smoothing step on frozen rims (default: on the 60x220x85 SPE10
                                 case it costs 8 % more Krylov iterations than 1 and a third less time per iteration) */
    /* second stage: block ILU(0) of the nf x nf block stencil in red-black ordering, one block per
     * rank as PETSc bjacobi+ilu (the slab couplings to other ranks are dropped).  The triangular solves read an
     * fp32 colour-separated copy of the factor (fp64 arithmetic): a fixed preconditioner, converged fields are
     * unaffected */
    int verbose;
} tpb_solver_opts;
enum tpb_mg_smoother { TPB_MG_RBGS = 0, TPB_MG_ZLINE = 1 };

typedef struct {
    int nits;                 /* snes.getIterationNumber(), thermalmodel.py:327 */
    int lits;                 /* snes.getLinearSolveIterations(), thermalmodel.py:328 */
    int reason;               /* SNESConvergedReason numbering: >0 converged, <0 diverged */
    int nfev;                 /* residual evaluations (line search included) */
    double fnorm0, fnorm;
    double t_assemble_ms, t_pcsetup_ms, t_ksp_ms, t_total_ms;  /* host clock around stream-synchronised phases */
} tpb_stats;

/* ---- life cycle ------------------------------------------------------------------------- */
/* replaces: model construction on a geo (singlephase.py:7-50, twophase.py:8-55). nphase 1|2. */
int tpb_create(const tpb_grid* grid, int nphase, const tpb_params* prm, int device, tpb_handle* out);
int tpb_destroy(tpb_handle h);
const char* tpb_last_error(tpb_handle h);
int tpb_version(void);

/* replaces: geo.phi/K_x/K_y/K_z/kT Functions (SPE10model3D.py:26-72, homogeneousboxgeo.py:10-19).
 * `data` has ncell doubles.  lo/hi: the neighbour ranks' boundary planes (may be NULL when
 * has_lo/has_hi is 0; tpb_exchange_static fills them over NCCL instead). */
int tpb_set_field(tpb_handle h, int field, const double* data, int on_device);
int tpb_set_field_ghost(tpb_handle h, int field, const double* lo, const double* hi, int on_device);
/* replaces: case.prod_wells / inj_wells / heaters / deltas_* (wellcase.py, heatercase.py, sourceterms.py) */
int tpb_set_sources(tpb_handle h, int n, const tpb_source* src);

/* ---- assembly (K1/K2): replaces assemble(F), assemble(J) inside SNES (thermalmodel.py:36,165) -- */
/* u, u_old: nf*ncell; F: nf*ncell; J (optional, may be NULL): ns*nf*nf*ncell, all device. */
int tpb_assemble(tpb_handle h, const double* u, const double* u_old, double dt, double* F, double* J);
/* ghost planes of the state for multi-rank slabs (nf*nplane each, device); NULL keeps previous */
int tpb_set_state_ghost(tpb_handle h, const double* u_lo, const double* u_hi);
size_t tpb_jacobian_size(tpb_handle h);   /* number of doubles in J */
int tpb_nstencil(tpb_handle h);

/* ---- SpMV (K9): replaces PETSc MatMult on the aij Jacobian ------------------------------- */
int tpb_spmv(tpb_handle h, const double* J, const double* x, double* y);

/* ---- preconditioner (K3-K8): replaces PCSetUp / PCApply of the composite tree -------------- */
int tpb_solver_defaults(int nphase, tpb_solver_opts* o);
int tpb_set_solver_opts(tpb_handle h, const tpb_solver_opts* o);
/* J: Jacobian to precondition; u: state it was assembled at (frozen coefficients of the
 * ConvDiff Schur operator, appctx["state"], preconditioners.py:24,178); dt as in assembly */
int tpb_pc_setup(tpb_handle h, const double* J, const double* u, double dt);
int tpb_pc_apply(tpb_handle h, const double* x, double* y);

/* ---- Krylov (K10): replaces KSPSolve (gmres right-PC | fgmres) ----------------------------- */
int tpb_ksp_solve(tpb_handle h, const double* J, const double* b, double* x, int* its, int* reason,
                  double* rnorm);

/* ---- Newton (K11 + A9): replaces NonlinearVariationalSolver.solve() (thermalmodel.py:165) -- */
/* u: in = initial guess, out = solution (device, nf*ncell); u_old device. */
int tpb_newton_solve(tpb_handle h, double* u, const double* u_old, double dt, tpb_stats* stats);
/* same, but u / u_old are HOST buffers: copies in, solves, copies the solution back */
int tpb_newton_solve_host(tpb_handle h, double* u_host, const double* u_old_host, double dt, tpb_stats* stats);

/* ---- small reductions used by the time loop (thermalmodel.py:190-229) ---------------------- */
/* out[0]=min S, out[1]=max S over field f of u */
int tpb_field_minmax(tpb_handle h, const double* u, int f, double* out);
int tpb_clip_field(tpb_handle h, double* u, int f, double lo, double hi);
/* two-phase: total oil mass in the reservoir, sum over all ranks' cells of V phi S_o rho_o(p, T) - replaces
 * assemble(phi*S_o*oil_rho(p,T)*dx), thermalmodel.py:190 */
int tpb_oil_mass(tpb_handle h, const double* u, double* out);
int tpb_dot(tpb_handle h, const double* x, const double* y, size_t n, double* out);

/* ---- multi-GPU: slab partition; NCCL is loaded at run time (dlopen of libnccl.so.2) -------- */
/* nccl_unique_id: the 128-byte ncclUniqueId made by rank 0 and broadcast by the caller */
int tpb_comm_init(tpb_handle h, const void* nccl_unique_id, int rank, int nranks);
int tpb_comm_unique_id(void* out128);
int tpb_exchange_static(tpb_handle h);   /* ghost planes of phi,K*,kT after tpb_set_field */
/* Which exchanges go through the peer-memory mailboxes (CUDA IPC over NVLink) instead of NCCL: bit 0 Krylov
   all-reduce, bit 1 halo planes, bit 2 multigrid gather level, bit 3 halo exchange fused into the SpMV kernel (one
   launch pushes the boundary planes, multiplies, and reads the neighbours' planes); 0 = NCCL only (single rank, IPC not available,
   or TPB_P2P=0 in the environment). */
int tpb_comm_peer_mode(tpb_handle h);

/* ---- preconditioner introspection (component-level parity tests; also handy when porting) ---- */
/* which: 0 = pressure hierarchy, 1 = temperature (Schur) hierarchy */
int tpb_pc_mg_nlevels(tpb_handle h, int which);
/* dims6 = nx,ny,nz of level l and its coarsening factors cx,cy,cz; op_out (device, ns*n_l doubles, may be NULL) */
int tpb_pc_mg_level(tpb_handle h, int which, int l, int* dims6, double* op_out);
int tpb_pc_mg_apply(tpb_handle h, int which, const double* b, double* y);      /* y = V-cycle(b), device */
int tpb_pc_stage2_apply(tpb_handle h, const double* r, double* z);              /* z = ILU(0)^-1 r, device */
int tpb_pc_get_weights(tpb_handle h, int f, double* out);                      /* decoupling weights w_f, device */

/* ---- instrumentation ------------------------------------------------------------------------ */
/* number of kernels this handle has launched since creation (bench.py "gpu_launches") */
int64_t tpb_launch_count(tpb_handle h);
/* timing of the dominant kernels with CUDA events on the handle's stream: runs `reps` launches
 * of kernel `which` (0 assemble F+J [property pre-pass + flux kernel + sources], 1 assemble F, 2 spmv,
 * 3 one colour pass of the fine-level pressure smoother; needs tpb_pc_setup) and returns the mean ms per launch */
int tpb_time_kernel(tpb_handle h, int which, const double* u, const double* u_old, double dt,
                    double* F, double* J, const double* x, double* y, int reps, double* ms);
void* tpb_stream(tpb_handle h);
/* block the calling host thread until everything queued on the handle's stream is done */
int tpb_sync(tpb_handle h);

#ifdef __cplusplus
}
#endif
#endif
