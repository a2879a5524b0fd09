// tpb_spmv.cu - K9: y = J x on the block-stencil layout J[s][r][c][cell].
//
// Stands in for PETSc MatMult on the assembled aij Jacobian (one per Krylov iteration and one
// inside the multiplicative composite PC: singlephase.py:341-343, twophase.py:531-533).
// No column indices are stored; the 5|7 neighbour cells are implied by the structured grid, so
// the compulsory traffic is the ns*nf*nf matrix values plus x and y (552 B/cell for 3-D two-phase).
// One thread per cell, x fastest: every J[s][r][c][:] read is a coalesced row; the neighbour
// reads of x are served by L1/L2 after the first touch.
#include "tpb_internal.cuh"

namespace {

template <int NF, int DIM>
__global__ void __launch_bounds__(256) spmv_kernel(const double* __restrict__ J, const double* __restrict__ x,
                                                   const double* __restrict__ x_lo, const double* __restrict__ x_hi,
                                                   double* __restrict__ y, Geom g) {
    const long long n = g.n;
    long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    const int nx = g.nx, ny = g.ny, np = g.np;
    int i = 0, j = 0, k = 0;
    if (cell < n) tpb_ijk(cell, nx, ny, i, j, k);
    pdl_wait();   // x (and its ghost planes) come from the predecessor
    if (cell >= n) return;

    double acc[NF];
#pragma unroll
    for (int r = 0; r < NF; r++) acc[r] = 0.0;

#pragma unroll
    for (int s = 0; s < 2 * DIM + 1; s++) {
        bool exists = true;
        long long nb = cell;
        if (s > 0) {
            const int axis = (s - 1) >> 1;
            const bool hi_side = ((s - 1) & 1) != 0;
            if (axis == 0) {
                exists = hi_side ? (i < nx - 1) : (i > 0);
                nb = cell + (hi_side ? 1 : -1);
            } else if (axis == 1) {
                if (DIM == 2)
                    exists = hi_side ? (j < ny - 1 || g.has_hi) : (j > 0 || g.has_lo);
                else
                    exists = hi_side ? (j < ny - 1) : (j > 0);
                nb = cell + (hi_side ? nx : -nx);
            } else {
                exists = hi_side ? (k < g.nz - 1 || g.has_hi) : (k > 0 || g.has_lo);
                nb = cell + (hi_side ? (long long)nx * ny : -(long long)nx * ny);
            }
        }
        if (!exists) continue;
        double xv[NF];
        if (nb < 0) {
#pragma unroll
            for (int c = 0; c < NF; c++) xv[c] = x_lo[(long long)c * np + nb + np];
        } else if (nb >= n) {
#pragma unroll
            for (int c = 0; c < NF; c++) xv[c] = x_hi[(long long)c * np + nb - n];
        } else {
#pragma unroll
            for (int c = 0; c < NF; c++) xv[c] = x[(long long)c * n + nb];
        }
#pragma unroll
        for (int r = 0; r < NF; r++)
#pragma unroll
            for (int c = 0; c < NF; c++)
                acc[r] = fma(__ldcs(&J[((long long)(s * NF + r) * NF + c) * n + cell]), xv[c], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < NF; r++) y[(long long)r * n + cell] = acc[r];
}

template <int NF, int DIM>
void launch_t(tpb_handle_s* h, const double* J, const double* x, double* y) {
    const long long n = h->g.n;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    launch_pdl(spmv_kernel<NF, DIM>, blocks, (unsigned)threads, h->stream, J, x, (const double*)h->x_lo, (const double*)h->x_hi, y, h->g);
    h->launches++;
    TPB_CUDA(cudaGetLastError());
}

}  // namespace

// x must already have its ghost planes in h->x_lo / h->x_hi when the slab has neighbours
void tpb_launch_spmv(tpb_handle_s* h, const double* J, const double* x, double* y) {
    if (h->g.has_lo || h->g.has_hi) {
        P2PView pv;
        if (tpb_p2p_view(h, 2, &pv))
            tpb_halo_vector(h, x, h->nf, h->x_lo, h->x_hi);   // peer-memory push/pull kernels: fine inside a graph capture
        else
            tpb_comm_op(h, [h, x]() { tpb_halo_vector(h, x, h->nf, h->x_lo, h->x_hi); });
    }
    if (h->nf == 2) {
        if (h->g.dim == 2)
            launch_t<2, 2>(h, J, x, y);
        else
            launch_t<2, 3>(h, J, x, y);
    } else {
        if (h->g.dim == 2)
            launch_t<3, 2>(h, J, x, y);
        else
            launch_t<3, 3>(h, J, x, y);
    }
}
