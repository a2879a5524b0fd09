// temporary stubs until tpb_pc.cu / tpb_solver.cu land
#include "tpb_internal.cuh"
void tpb_pc_free(tpb_handle_s*) {}
void tpb_ksp_free(tpb_handle_s*) {}
void tpb_pc_setup_impl(tpb_handle_s*, const double*, const double*, double) { throw tpb_exception{TPB_ERR_UNSUPPORTED, "pc not built"}; }
void tpb_pc_apply_impl(tpb_handle_s*, const double*, double*) { throw tpb_exception{TPB_ERR_UNSUPPORTED, "pc not built"}; }
void tpb_ksp_solve_impl(tpb_handle_s*, const double*, const double*, double*, int*, int*, double*) { throw tpb_exception{TPB_ERR_UNSUPPORTED, "ksp not built"}; }
void tpb_newton_impl(tpb_handle_s*, double*, const double*, double, tpb_stats*) { throw tpb_exception{TPB_ERR_UNSUPPORTED, "newton not built"}; }
