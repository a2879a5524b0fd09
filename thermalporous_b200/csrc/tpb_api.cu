// tpb_api.cu - extern "C" surface of libtpb200.so (declared in include/tpb200.h).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>

#include "tpb_internal.cuh"

void tpb_minmax_impl(tpb_handle_s* h, size_t n, const double* x, double* out2);
void tpb_oil_mass_impl(tpb_handle_s* h, const double* u);
void tpb_clip_impl(tpb_handle_s* h, size_t n, double* x, double lo, double hi);
void tpb_comm_init_impl(tpb_handle_s* h, const void* id128, int rank, int nranks);
void tpb_comm_unique_id_impl(void* out128);
void tpb_allreduce_max(tpb_handle_s* h, double* dev_buf, int count);
int tpb_pc_mg_nlevels_impl(tpb_handle_s* h, int which);
int tpb_pc_mg_level_impl(tpb_handle_s* h, int which, int l, int* dims6, double* op_out);
void tpb_pc_mg_apply_impl(tpb_handle_s* h, int which, const double* b, double* y);
void tpb_pc_stage2_apply_impl(tpb_handle_s* h, const double* r, double* z);
const double* tpb_pc_weights_impl(tpb_handle_s* h, int f);
long long tpb_pc_rbgs_pass_impl(tpb_handle_s* h, int col);

static thread_local std::string g_err;

bool tpb_pdl_enabled() {
    static const bool on = !(getenv("TPB_PDL") && atoi(getenv("TPB_PDL")) == 0);
    return on;
}

#define TPB_TRY(h)  \
    try {           \
        if (h) TPB_CUDA(cudaSetDevice((h)->device));
#define TPB_CATCH(h)                         \
    }                                        \
    catch (const tpb_exception& e) {         \
        if (h) (h)->err = e.msg;             \
        g_err = e.msg;                       \
        return e.code;                       \
    }                                        \
    catch (const std::exception& e) {        \
        if (h) (h)->err = e.what();          \
        g_err = e.what();                    \
        return TPB_ERR_STATE;                \
    }                                        \
    return TPB_OK;

extern "C" {

int tpb_version(void) { return 100; }

const char* tpb_last_error(tpb_handle h) { return h ? h->err.c_str() : g_err.c_str(); }

int tpb_create(const tpb_grid* grid, int nphase, const tpb_params* prm, int device, tpb_handle* out) {
    tpb_handle_s* h = nullptr;
    try {
        TPB_REQUIRE(grid && prm && out, TPB_ERR_ARG, "null argument");
        TPB_REQUIRE(nphase == 1 || nphase == 2, TPB_ERR_ARG, "nphase must be 1 or 2");
        TPB_REQUIRE(grid->dim == 2 || grid->dim == 3, TPB_ERR_ARG, "dim must be 2 or 3");
        TPB_REQUIRE(grid->nx > 0 && grid->ny > 0 && grid->nz > 0, TPB_ERR_ARG, "empty grid");
        TPB_REQUIRE(grid->dim == 3 || grid->nz == 1, TPB_ERR_ARG, "2-D grids need nz == 1");
        TPB_REQUIRE((double)grid->nx * grid->ny * ((double)grid->nz + 2.0) < 2147483647.0, TPB_ERR_ARG,
                    "a slab must hold fewer than 2^31 cells (kernels index cells in 32 bits); use more slabs");
        int ndev = 0;
        TPB_CUDA(cudaGetDeviceCount(&ndev));
        TPB_REQUIRE(device >= 0 && device < ndev, TPB_ERR_CUDA, "no such CUDA device (libtpb200 has no CPU fallback)");
        TPB_CUDA(cudaSetDevice(device));
        h = new tpb_handle_s();
        h->device = device;
        h->nphase = nphase;
        h->nf = nphase == 1 ? 2 : 3;
        h->ns = tpb_ns(grid->dim);
        Geom& g = h->g;
        g.dim = grid->dim;
        g.nx = grid->nx;
        g.ny = grid->ny;
        g.nz = grid->nz;
        g.n = (long long)g.nx * g.ny * g.nz;
        g.np = g.dim == 3 ? g.nx * g.ny : g.nx;
        g.nl = g.dim == 3 ? g.nz : g.ny;
        g.has_lo = grid->has_lo ? 1 : 0;
        g.has_hi = grid->has_hi ? 1 : 0;
        g.h[0] = grid->dx;
        g.h[1] = grid->dy;
        g.h[2] = g.dim == 3 ? grid->dz : 1.0;
        for (int a = 0; a < 3; a++) g.ih[a] = 1.0 / g.h[a];
        if (g.dim == 3) {
            g.area[0] = grid->dy * grid->dz;
            g.area[1] = grid->dx * grid->dz;
            g.area[2] = grid->dx * grid->dy;
            g.vol = grid->dx * grid->dy * grid->dz;
        } else {
            g.area[0] = grid->dy;
            g.area[1] = grid->dx;
            g.area[2] = 0.0;
            g.vol = grid->dx * grid->dy;
        }
        h->prm = *prm;
        DevParams& d = h->dp;
        d.ko = prm->ko;
        d.kw = prm->kw;
        d.kr = prm->kr;
        d.c_v_w = prm->c_v_w;
        d.c_v_o = prm->c_v_o;
        d.c_r = prm->c_r;
        d.rho_r = prm->rho_r;
        d.T_inj = prm->T_inj;
        d.T_prod = prm->T_prod;
        d.U = prm->U;
        d.g = (g.dim == 3 && prm->gravity) ? prm->g : 0.0;
        if (nphase == 2) {
            d.Wp = prm->T_prod;                                                         // twophase.py:144
            d.Wo = prm->T_prod * (prm->c_v_w * (1.0 - prm->S_o) + prm->c_v_o * prm->S_o);  // :147
        } else {
            d.Wp = 1.0;  // scaled_eqns = False, singlephase.py:26,112-115
            d.Wo = 1.0;
        }
        d.rho_ref = 141.5 / (prm->API + 131.5) * 999.0;                                 // physicalparameters.py:39-40
        d.mu_o_pref = 1e-3 * pow(10.0, -0.8021 * prm->API + 23.8765);                   // :52-57
        d.mu_o_exp = 0.31458 * prm->API - 9.21592;
        TPB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        TPB_CUDA(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
        TPB_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        TPB_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        h->src_index = tpb_dalloc<int>(g.n);
        TPB_CUDA(cudaMemsetAsync(h->src_index, 0xff, g.n * sizeof(int), h->stream));
        for (int f = 0; f < 5; f++) {
            h->fld[f] = tpb_dalloc<double>(g.n);
            h->fld_lo[f] = tpb_dalloc<double>(g.np);
            h->fld_hi[f] = tpb_dalloc<double>(g.np);
            TPB_CUDA(cudaMemsetAsync(h->fld[f], 0, g.n * sizeof(double), h->stream));
            TPB_CUDA(cudaMemsetAsync(h->fld_lo[f], 0, g.np * sizeof(double), h->stream));
            TPB_CUDA(cudaMemsetAsync(h->fld_hi[f], 0, g.np * sizeof(double), h->stream));
        }
        h->u_lo = tpb_dalloc<double>((size_t)h->nf * g.np);
        h->u_hi = tpb_dalloc<double>((size_t)h->nf * g.np);
        h->x_lo = tpb_dalloc<double>((size_t)h->nf * g.np);
        h->x_hi = tpb_dalloc<double>((size_t)h->nf * g.np);
        TPB_CUDA(cudaMemsetAsync(h->u_lo, 0, (size_t)h->nf * g.np * sizeof(double), h->stream));
        TPB_CUDA(cudaMemsetAsync(h->u_hi, 0, (size_t)h->nf * g.np * sizeof(double), h->stream));
        TPB_CUDA(cudaMemsetAsync(h->x_lo, 0, (size_t)h->nf * g.np * sizeof(double), h->stream));
        TPB_CUDA(cudaMemsetAsync(h->x_hi, 0, (size_t)h->nf * g.np * sizeof(double), h->stream));
        tpb_solver_defaults(nphase, &h->opts);
        TPB_CUDA(cudaStreamSynchronize(h->stream));
        *out = h;
    } catch (const tpb_exception& e) {
        g_err = e.msg;
        delete h;
        return e.code;
    }
    return TPB_OK;
}

int tpb_destroy(tpb_handle h) {
    if (!h) return TPB_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    tpb_pc_free(h);
    tpb_ksp_free(h);
    tpb_comm_free(h);
    for (int f = 0; f < 5; f++) {
        tpb_dfree(h->fld[f]);
        tpb_dfree(h->fld_lo[f]);
        tpb_dfree(h->fld_hi[f]);
    }
    tpb_dfree(h->u_lo);
    tpb_dfree(h->u_hi);
    tpb_dfree(h->x_lo);
    tpb_dfree(h->x_hi);
    tpb_dfree(h->scr);
    for (int a = 0; a < 3; a++) tpb_dfree(h->trans[a]);
    tpb_dfree(h->trans_lo);
    tpb_dfree(h->src_cell);
    tpb_dfree(h->src_off);
    tpb_dfree(h->src_ent);
    tpb_dfree(h->src_index);
    tpb_dfree(h->src_acc);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    tpb_dfree(h->red_partial);
    tpb_dfree(h->red_counter);
    tpb_dfree(h->red_out);
    if (h->red_host) cudaFreeHost(h->red_host);
    tpb_dfree(h->nw_F);
    tpb_dfree(h->nw_J);
    tpb_dfree(h->nw_du);
    tpb_dfree(h->nw_utrial);
    tpb_dfree(h->nw_Ftrial);
    tpb_dfree(h->nw_uold);
    tpb_dfree(h->nw_u);
    cudaStreamDestroy(h->stream);
    delete h;
    return TPB_OK;
}

int tpb_set_field(tpb_handle h, int field, const double* data, int on_device) {
    TPB_TRY(h)
    TPB_REQUIRE(h && data, TPB_ERR_ARG, "null argument");
    TPB_REQUIRE(field >= 0 && field < 5, TPB_ERR_ARG, "unknown field id");
    TPB_CUDA(cudaMemcpyAsync(h->fld[field], data, h->g.n * sizeof(double),
                             on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    h->fld_set[field] = true;
    h->trans_dirty = true;
    TPB_CATCH(h)
}

int tpb_set_field_ghost(tpb_handle h, int field, const double* lo, const double* hi, int on_device) {
    TPB_TRY(h)
    TPB_REQUIRE(h, TPB_ERR_ARG, "null handle");
    TPB_REQUIRE(field >= 0 && field < 5, TPB_ERR_ARG, "unknown field id");
    cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (lo) TPB_CUDA(cudaMemcpyAsync(h->fld_lo[field], lo, h->g.np * sizeof(double), kind, h->stream));
    if (hi) TPB_CUDA(cudaMemcpyAsync(h->fld_hi[field], hi, h->g.np * sizeof(double), kind, h->stream));
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    h->trans_dirty = true;
    TPB_CATCH(h)
}

int tpb_set_sources(tpb_handle h, int n, const tpb_source* src) {
    TPB_TRY(h)
    TPB_REQUIRE(h && (n == 0 || src), TPB_ERR_ARG, "null argument");
    tpb_dfree(h->src_cell);
    tpb_dfree(h->src_off);
    tpb_dfree(h->src_ent);
    tpb_dfree(h->src_acc);
    h->src_acc = nullptr;
    TPB_CUDA(cudaMemsetAsync(h->src_index, 0xff, h->g.n * sizeof(int), h->stream));
    h->src_cell = nullptr;
    h->src_off = nullptr;
    h->src_ent = nullptr;
    h->nsrc_cells = 0;
    if (n > 0) {
        std::vector<tpb_source> ent(src, src + n);
        for (auto& s : ent) {
            TPB_REQUIRE(s.cell >= 0 && s.cell < h->g.n, TPB_ERR_ARG, "source cell outside the local slab");
            TPB_REQUIRE(s.kind >= 0 && s.kind <= 2, TPB_ERR_ARG, "unknown source kind");
        }
        // group by cell, keeping the caller's order inside a cell (deterministic summation)
        std::stable_sort(ent.begin(), ent.end(), [](const tpb_source& a, const tpb_source& b) { return a.cell < b.cell; });
        std::vector<int64_t> cells;
        std::vector<int> off;
        for (int i = 0; i < n; i++) {
            if (i == 0 || ent[i].cell != ent[i - 1].cell) {
                cells.push_back(ent[i].cell);
                off.push_back(i);
            }
        }
        off.push_back(n);
        h->nsrc_cells = (int)cells.size();
        h->src_cell = tpb_dalloc<int64_t>(cells.size());
        h->src_off = tpb_dalloc<int>(off.size());
        h->src_ent = tpb_dalloc<tpb_source>(n);
        TPB_CUDA(cudaMemcpy(h->src_cell, cells.data(), cells.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        TPB_CUDA(cudaMemcpy(h->src_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
        TPB_CUDA(cudaMemcpy(h->src_ent, ent.data(), (size_t)n * sizeof(tpb_source), cudaMemcpyHostToDevice));
        h->src_acc = tpb_dalloc<double>(cells.size() * (size_t)(h->nf + h->nf * h->nf));
        TPB_CUDA(cudaMemsetAsync(h->src_acc, 0, cells.size() * (size_t)(h->nf + h->nf * h->nf) * sizeof(double), h->stream));
        std::vector<int> idx(h->g.n, -1);
        for (size_t q = 0; q < cells.size(); q++) idx[cells[q]] = (int)q;
        TPB_CUDA(cudaMemcpyAsync(h->src_index, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        TPB_CUDA(cudaStreamSynchronize(h->stream));
    }
    TPB_CATCH(h)
}

size_t tpb_jacobian_size(tpb_handle h) { return h ? (size_t)h->ns * h->nf * h->nf * h->g.n : 0; }
int tpb_nstencil(tpb_handle h) { return h ? h->ns : 0; }

static void check_fields(tpb_handle_s* h) {
    TPB_REQUIRE(h->fld_set[TPB_PHI] && h->fld_set[TPB_KX] && h->fld_set[TPB_KY], TPB_ERR_STATE,
                "phi, K_x and K_y must be set before assembly");
    TPB_REQUIRE(h->g.dim == 2 || h->fld_set[TPB_KZ], TPB_ERR_STATE, "K_z must be set for 3-D grids");
    TPB_REQUIRE(h->nphase == 2 || h->fld_set[TPB_KT], TPB_ERR_STATE, "kT must be set for single-phase models");
}

int tpb_set_state_ghost(tpb_handle h, const double* u_lo, const double* u_hi) {
    TPB_TRY(h)
    TPB_REQUIRE(h, TPB_ERR_ARG, "null handle");
    size_t bytes = (size_t)h->nf * h->g.np * sizeof(double);
    if (u_lo) TPB_CUDA(cudaMemcpyAsync(h->u_lo, u_lo, bytes, cudaMemcpyDeviceToDevice, h->stream));
    if (u_hi) TPB_CUDA(cudaMemcpyAsync(h->u_hi, u_hi, bytes, cudaMemcpyDeviceToDevice, h->stream));
    TPB_CATCH(h)
}

int tpb_assemble(tpb_handle h, const double* u, const double* u_old, double dt, double* F, double* J) {
    TPB_TRY(h)
    TPB_REQUIRE(h && u && u_old && F, TPB_ERR_ARG, "null argument");
    TPB_REQUIRE(dt > 0.0, TPB_ERR_ARG, "dt must be positive");
    check_fields(h);
    if (h->comm && (h->g.has_lo || h->g.has_hi)) tpb_halo_vector(h, u, h->nf, h->u_lo, h->u_hi);
    tpb_launch_assemble(h, u, u_old, dt, F, J);
    TPB_CATCH(h)
}

int tpb_spmv(tpb_handle h, const double* J, const double* x, double* y) {
    TPB_TRY(h)
    TPB_REQUIRE(h && J && x && y, TPB_ERR_ARG, "null argument");
    tpb_launch_spmv(h, J, x, y);
    TPB_CATCH(h)
}

int tpb_solver_defaults(int nphase, tpb_solver_opts* o) {
    if (!o) return TPB_ERR_ARG;
    memset(o, 0, sizeof(*o));
    o->snes_max_it = nphase == 1 ? 15 : 25;       // singlephase.py:293, twophase.py:424
    o->snes_rtol = 1e-8;                          // PETSc SNES defaults (the dicts set none)
    o->snes_atol = 1e-50;
    o->snes_stol = 1e-8;
    o->linesearch = 0;                            // Firedrake's NonlinearVariationalSolver default: basic
    o->ksp_type = nphase == 1 ? TPB_KSP_GMRES : TPB_KSP_FGMRES;   // singlephase.py:295, twophase.py:426
    o->ksp_max_it = 200;
    o->ksp_restart = 200;
    o->ksp_rtol = nphase == 1 ? 1e-5 : 1e-8;      // KSP default | twophase.py:432
    o->ksp_atol = 1e-50;
    o->stage1 = nphase == 1 ? TPB_S1_CPR : TPB_S1_CPTR;
    o->decoup = TPB_DECOUP_NO;
    o->schur_pre = TPB_SCHUR_CONVDIFF;
    o->stage2 = TPB_S2_ILU0;
    o->mg_pre = 2;
    o->mg_post = 2;
    o->mg_coarse_sweeps = 4;
    o->mg_min_cells = 8;
    o->mg_overcorrection = 1.0;
    o->mg_cycles = 1;
    o->mg_semi_theta = 0.5;
    o->mg_full_below = 0;
    o->mg_dd_stop = 0.1;
    o->mg_coarse_scale = 0.5;
    o->mg_smoother = TPB_MG_ZLINE;
    o->mg_tile_sweeps = 0;
    o->verbose = 0;
    return TPB_OK;
}

int tpb_set_solver_opts(tpb_handle h, const tpb_solver_opts* o) {
    TPB_TRY(h)
    TPB_REQUIRE(h && o, TPB_ERR_ARG, "null argument");
    TPB_REQUIRE(o->ksp_restart > 0 && o->ksp_restart <= 250 && o->ksp_max_it >= 0 && o->snes_max_it >= 0, TPB_ERR_ARG,
                "bad iteration limits (ksp_restart must be in [1, 250])");
    TPB_REQUIRE(o->ksp_type == TPB_KSP_GMRES || o->ksp_type == TPB_KSP_FGMRES, TPB_ERR_ARG, "ksp_type must be gmres (0) or fgmres (1)");
    TPB_REQUIRE(o->stage1 >= 0 && o->stage1 <= 3 && o->stage2 >= 0 && o->stage2 <= 2, TPB_ERR_ARG, "bad PC stage");
    TPB_REQUIRE(o->schur_pre >= 0 && o->schur_pre <= 3 && o->decoup >= 0 && o->decoup <= 4, TPB_ERR_ARG, "bad PC option");
    TPB_REQUIRE(!(o->stage1 == TPB_S1_CPTR && h->nphase == 1), TPB_ERR_UNSUPPORTED,
                "CPTR needs the two-phase model (preconditioners.py:1258-1267)");
    TPB_REQUIRE(!(o->stage1 == TPB_S1_FIELDSPLIT && h->nphase == 2), TPB_ERR_UNSUPPORTED,
                "pc_fieldsplit_* option sets exist for the single-phase model only (singlephase.py:309-338)");
    TPB_REQUIRE(!((o->decoup == TPB_DECOUP_QI_TEMP || o->decoup == TPB_DECOUP_TI_TEMP) &&
                  !(h->nphase == 2 && o->stage1 == TPB_S1_CPR)),
                TPB_ERR_UNSUPPORTED, "QI_temp/TI_temp decouple (T,S) from p: two-phase CPR only (preconditioners.py:442)");
    tpb_pc_free(h);
    tpb_ksp_free(h);
    h->opts = *o;
    TPB_CATCH(h)
}

int tpb_pc_setup(tpb_handle h, const double* J, const double* u, double dt) {
    TPB_TRY(h)
    TPB_REQUIRE(h && J && u, TPB_ERR_ARG, "null argument");
    tpb_pc_setup_impl(h, J, u, dt);
    TPB_CATCH(h)
}

int tpb_pc_apply(tpb_handle h, const double* x, double* y) {
    TPB_TRY(h)
    TPB_REQUIRE(h && x && y, TPB_ERR_ARG, "null argument");
    tpb_pc_apply_impl(h, x, y);
    TPB_CATCH(h)
}

int tpb_ksp_solve(tpb_handle h, const double* J, const double* b, double* x, int* its, int* reason, double* rnorm) {
    TPB_TRY(h)
    TPB_REQUIRE(h && J && b && x, TPB_ERR_ARG, "null argument");
    int it = 0, rs = 0;
    double rn = 0.0;
    tpb_ksp_solve_impl(h, J, b, x, &it, &rs, &rn);
    if (its) *its = it;
    if (reason) *reason = rs;
    if (rnorm) *rnorm = rn;
    TPB_CATCH(h)
}

int tpb_newton_solve(tpb_handle h, double* u, const double* u_old, double dt, tpb_stats* stats) {
    TPB_TRY(h)
    TPB_REQUIRE(h && u && u_old && stats, TPB_ERR_ARG, "null argument");
    TPB_REQUIRE(dt > 0.0, TPB_ERR_ARG, "dt must be positive");
    check_fields(h);
    tpb_newton_impl(h, u, u_old, dt, stats);
    TPB_CATCH(h)
}

int tpb_newton_solve_host(tpb_handle h, double* u_host, const double* u_old_host, double dt, tpb_stats* stats) {
    TPB_TRY(h)
    TPB_REQUIRE(h && u_host && u_old_host && stats, TPB_ERR_ARG, "null argument");
    TPB_REQUIRE(dt > 0.0, TPB_ERR_ARG, "dt must be positive");
    check_fields(h);
    size_t nd = (size_t)h->nf * h->g.n;
    if (!h->nw_u) h->nw_u = tpb_dalloc<double>(nd);
    if (!h->nw_uold) h->nw_uold = tpb_dalloc<double>(nd);
    TPB_CUDA(cudaMemcpyAsync(h->nw_u, u_host, nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    TPB_CUDA(cudaMemcpyAsync(h->nw_uold, u_old_host, nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    tpb_newton_impl(h, h->nw_u, h->nw_uold, dt, stats);
    TPB_CUDA(cudaMemcpyAsync(u_host, h->nw_u, nd * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    TPB_CATCH(h)
}

int tpb_oil_mass(tpb_handle h, const double* u, double* out) {
    TPB_TRY(h)
    TPB_REQUIRE(h && u && out, TPB_ERR_ARG, "null argument");
    TPB_REQUIRE(h->nphase == 2, TPB_ERR_UNSUPPORTED, "oil mass is a two-phase diagnostic (thermalmodel.py:184-192)");
    TPB_REQUIRE(h->fld_set[TPB_PHI], TPB_ERR_STATE, "porosity not set");
    tpb_oil_mass_impl(h, u);
    if (h->comm) tpb_allreduce_sum(h, h->red_out, 1);
    TPB_CUDA(cudaMemcpyAsync(out, h->red_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    TPB_CATCH(h)
}

int tpb_field_minmax(tpb_handle h, const double* u, int f, double* out) {
    TPB_TRY(h)
    TPB_REQUIRE(h && u && out && f >= 0 && f < h->nf, TPB_ERR_ARG, "bad argument");
    tpb_minmax_impl(h, h->g.n, u + (size_t)f * h->g.n, out);
    if (h->comm) {
        // global min/max over the slabs: max-reduce (-min, max)
        double v[2] = {-out[0], out[1]};
        TPB_CUDA(cudaMemcpyAsync(h->red_out, v, 2 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        tpb_allreduce_max(h, h->red_out, 2);
        TPB_CUDA(cudaMemcpyAsync(v, h->red_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        TPB_CUDA(cudaStreamSynchronize(h->stream));
        out[0] = -v[0];
        out[1] = v[1];
    }
    TPB_CATCH(h)
}

int tpb_clip_field(tpb_handle h, double* u, int f, double lo, double hi) {
    TPB_TRY(h)
    TPB_REQUIRE(h && u && f >= 0 && f < h->nf, TPB_ERR_ARG, "bad argument");
    tpb_clip_impl(h, h->g.n, u + (size_t)f * h->g.n, lo, hi);
    TPB_CATCH(h)
}

int tpb_dot(tpb_handle h, const double* x, const double* y, size_t n, double* out) {
    TPB_TRY(h)
    TPB_REQUIRE(h && x && y && out, TPB_ERR_ARG, "null argument");
    *out = tpb_dot_sync(h, n, x, y);
    TPB_CATCH(h)
}

int tpb_comm_init(tpb_handle h, const void* nccl_unique_id, int rank, int nranks) {
    TPB_TRY(h)
    TPB_REQUIRE(h && (nranks == 1 || nccl_unique_id), TPB_ERR_ARG, "null argument");
    tpb_comm_init_impl(h, nccl_unique_id, rank, nranks);
    TPB_CATCH(h)
}

int tpb_comm_unique_id(void* out128) {
    tpb_handle_s* h = nullptr;
    TPB_TRY(h)
    TPB_REQUIRE(out128, TPB_ERR_ARG, "null argument");
    tpb_comm_unique_id_impl(out128);
    TPB_CATCH(h)
}

int tpb_exchange_static(tpb_handle h) {
    TPB_TRY(h)
    TPB_REQUIRE(h, TPB_ERR_ARG, "null handle");
    for (int f = 0; f < 5; f++)
        if (h->fld_set[f]) tpb_halo_vector(h, h->fld[f], 1, h->fld_lo[f], h->fld_hi[f]);
    h->trans_dirty = true;
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    TPB_CATCH(h)
}

int tpb_pc_mg_nlevels(tpb_handle h, int which) { return h ? tpb_pc_mg_nlevels_impl(h, which) : 0; }
int tpb_pc_mg_level(tpb_handle h, int which, int l, int* dims6, double* op_out) {
    TPB_TRY(h)
    TPB_REQUIRE(h, TPB_ERR_ARG, "null handle");
    TPB_REQUIRE(tpb_pc_mg_level_impl(h, which, l, dims6, op_out) == 0, TPB_ERR_ARG, "no such multigrid level");
    TPB_CATCH(h)
}
int tpb_pc_mg_apply(tpb_handle h, int which, const double* b, double* y) {
    TPB_TRY(h)
    TPB_REQUIRE(h && b && y, TPB_ERR_ARG, "null argument");
    tpb_pc_mg_apply_impl(h, which, b, y);
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    TPB_CATCH(h)
}
int tpb_pc_stage2_apply(tpb_handle h, const double* r, double* z) {
    TPB_TRY(h)
    TPB_REQUIRE(h && r && z, TPB_ERR_ARG, "null argument");
    tpb_pc_stage2_apply_impl(h, r, z);
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    TPB_CATCH(h)
}
int tpb_pc_get_weights(tpb_handle h, int f, double* out) {
    TPB_TRY(h)
    TPB_REQUIRE(h && out && f >= 0 && f < h->nf, TPB_ERR_ARG, "bad argument");
    const double* w = tpb_pc_weights_impl(h, f);
    TPB_REQUIRE(w != nullptr, TPB_ERR_STATE, "no decoupling weights (stage 1 not set up)");
    TPB_CUDA(cudaMemcpyAsync(out, w, h->g.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    TPB_CATCH(h)
}

int64_t tpb_launch_count(tpb_handle h) { return h ? h->launches : 0; }
int tpb_comm_peer_mode(tpb_handle h) { return h ? tpb_comm_peer_mode_impl(h) : 0; }
void* tpb_stream(tpb_handle h) { return h ? (void*)h->stream : nullptr; }
int tpb_sync(tpb_handle h) {
    TPB_TRY(h)
    TPB_REQUIRE(h, TPB_ERR_ARG, "null handle");
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    TPB_CATCH(h)
}

int tpb_time_kernel(tpb_handle h, int which, const double* u, const double* u_old, double dt, double* F, double* J,
                    const double* x, double* y, int reps, double* ms) {
    TPB_TRY(h)
    TPB_REQUIRE(h && ms && reps > 0, TPB_ERR_ARG, "bad argument");
    cudaEvent_t e0, e1;
    TPB_CUDA(cudaEventCreate(&e0));
    TPB_CUDA(cudaEventCreate(&e1));
    auto once = [&]() {
        if (which == 0)
            tpb_launch_assemble(h, u, u_old, dt, F, J);
        else if (which == 1)
            tpb_launch_assemble(h, u, u_old, dt, F, nullptr);
        else if (which == 3)
            tpb_pc_rbgs_pass_impl(h, 1);
        else
            tpb_launch_spmv(h, J, x, y);
    };
    once();
    TPB_CUDA(cudaEventRecord(e0, h->stream));
    for (int r = 0; r < reps; r++) once();
    TPB_CUDA(cudaEventRecord(e1, h->stream));
    TPB_CUDA(cudaEventSynchronize(e1));
    float t = 0.f;
    TPB_CUDA(cudaEventElapsedTime(&t, e0, e1));
    *ms = (double)t / reps;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    TPB_CATCH(h)
}

}  // extern "C"
