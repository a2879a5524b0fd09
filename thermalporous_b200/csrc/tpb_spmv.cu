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


// ---- fused halo exchange + SpMV over peer memory ------------------------------------------------------------
// Multi-rank slabs with the mailboxes up (tpb_internal.cuh): ONE kernel pushes this slab's two boundary planes of x
// straight into the neighbours' HBM over NVLink, multiplies, and reads the neighbours' planes out of its own
// mailbox - instead of pack/send/recv (or push and pull kernels) followed by the multiply.  Block order is
// remapped so that the blocks holding the first AND the last plane run first: their values are on the wire while
// the interior is processed, and only those blocks ever wait, just before their two slab-axis slots (the last two
// of the stencil), with everything else of their rows already accumulated.  Protocol: message number e =
// epoch + 1; data into the parity-(e&1) buffer, system fence, the last pushing block of a plane publishes e in the
// neighbour's flag word; the last block to have read the epoch advances it (tickets wrap to zero by atomicInc).
template <int NF, int DIM>
__global__ void __launch_bounds__(256) spmv_halo_kernel(const double* __restrict__ J, const double* __restrict__ x,
                                                        double* __restrict__ y, Geom g, P2PView v,
                                                        unsigned int* __restrict__ tickets, unsigned Bl, unsigned Bh) {
    const long long n = g.n;
    pdl_launch_dependents();
    const int nx = g.nx, ny = g.ny, np = g.np;
    // chunk order: [chunks of the first plane][chunks of the last plane, from the end][the rest]
    const unsigned b = blockIdx.x, nb_all = gridDim.x;
    const unsigned chunk = b < Bl ? b : (b < Bl + Bh ? nb_all - 1 - (b - Bl) : b - Bh);
    const long long cell = (long long)chunk * blockDim.x + threadIdx.x;
    const bool valid = cell < n;
    int i = 0, j = 0, k = 0;
    if (valid) tpb_ijk(cell, nx, ny, i, j, k);
    const bool in_lo = valid && cell < np, in_hi = valid && cell >= n - np;
    const bool blk_lo = g.has_lo && (long long)chunk * blockDim.x < np;                 // block touches the first plane
    const bool blk_hi = g.has_hi && (long long)(chunk + 1) * blockDim.x > n - np;       // ... the last plane
    pdl_wait();   // x comes from the predecessor; the epoch from the previous exchange

    __shared__ unsigned long long s_e;
    if (threadIdx.x == 0) s_e = *reinterpret_cast<volatile unsigned long long*>(&v.epoch[P2P_SLOT_HALO_LO]) + 1;
    __syncthreads();
    const unsigned long long e = s_e;
    const int par = (int)(e & 1);
    if (blk_lo || blk_hi) {
        if (g.has_lo && in_lo) {   // my first plane = upper ghost of the rank below
            double* dst = p2p_halo_area(v, v.rank - 1, false, par);
#pragma unroll
            for (int c = 0; c < NF; c++) dst[(long long)c * np + cell] = x[(long long)c * n + cell];
        }
        if (g.has_hi && in_hi) {   // my last plane = lower ghost of the rank above
            double* dst = p2p_halo_area(v, v.rank + 1, true, par);
#pragma unroll
            for (int c = 0; c < NF; c++) dst[(long long)c * np + (cell - (n - np))] = x[(long long)c * n + cell];
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            if (blk_lo && atomicInc(&tickets[0], Bl - 1) == Bl - 1) {
                __threadfence_system();
                p2p_store_flag(p2p_flag(v, v.rank - 1, P2P_SLOT_HALO_HI, v.rank), e);
            }
            if (blk_hi && atomicInc(&tickets[1], Bh - 1) == Bh - 1) {
                __threadfence_system();
                p2p_store_flag(p2p_flag(v, v.rank + 1, P2P_SLOT_HALO_LO, v.rank), e);
            }
        }
    }
    // every block has read the epoch by now: the last one to say so advances it
    if (threadIdx.x == 0 && atomicInc(&tickets[2], nb_all - 1) == nb_all - 1)
        *reinterpret_cast<volatile unsigned long long*>(&v.epoch[P2P_SLOT_HALO_LO]) = e;

    double acc[NF];
#pragma unroll
    for (int r = 0; r < NF; r++) acc[r] = 0.0;
    constexpr int NS = 2 * DIM + 1;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        if (s == NS - 2 && (blk_lo || blk_hi)) {
            // the neighbours' planes are needed from here on (block-uniform branch)
            if (threadIdx.x == 0) {
                if (blk_lo) p2p_wait(v, P2P_SLOT_HALO_LO, v.rank - 1, e);
                if (blk_hi) p2p_wait(v, P2P_SLOT_HALO_HI, v.rank + 1, e);
            }
            __syncthreads();
        }
        bool exists = valid;
        long long nb = cell;
        if (s > 0) {
            const int axis = (s - 1) >> 1;
            const bool hi_side = ((s - 1) & 1) != 0;
            if (axis == 0) {
                exists = valid && (hi_side ? (i < nx - 1) : (i > 0));
                nb = cell + (hi_side ? 1 : -1);
            } else if (axis == 1 && DIM == 3) {
                exists = valid && (hi_side ? (j < ny - 1) : (j > 0));
                nb = cell + (hi_side ? nx : -nx);
            } else {
                exists = valid && (hi_side ? (!in_hi || g.has_hi) : (!in_lo || g.has_lo));
                nb = cell + (hi_side ? (long long)np : -(long long)np);
            }
        }
        if (!exists) continue;
        double xv[NF];
        if (nb < 0) {
            const double* src = p2p_halo_area(v, v.rank, true, par);
#pragma unroll
            for (int c = 0; c < NF; c++) xv[c] = __ldcg(src + (long long)c * np + nb + np);
        } else if (nb >= n) {
            const double* src = p2p_halo_area(v, v.rank, false, par);
#pragma unroll
            for (int c = 0; c < NF; c++) xv[c] = __ldcg(src + (long long)c * np + nb - n);
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
    if (valid) {
#pragma unroll
        for (int r = 0; r < NF; r++) y[(long long)r * n + cell] = acc[r];
    }
}

template <int NF, int DIM>
bool launch_fused_t(tpb_handle_s* h, const double* J, const double* x, double* y) {
    P2PView pv;
    unsigned int* tickets = nullptr;
    if (!tpb_p2p_halo(h, &pv, &tickets)) return false;
    const long long n = h->g.n;
    const int np = h->g.np;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    const unsigned Bl = h->g.has_lo ? (unsigned)((np + threads - 1) / threads) : 0;
    const unsigned Bh = h->g.has_hi ? blocks - (unsigned)((n - np) / threads) : 0;
    if ((long long)np * NF > pv.halo_cap || Bl + Bh > blocks) return false;   // slabs of one or two planes: unfused path
    launch_pdl(spmv_halo_kernel<NF, DIM>, blocks, (unsigned)threads, h->stream, J, x, y, h->g, pv, tickets, Bl, Bh);
    h->launches++;
    TPB_CUDA(cudaGetLastError());
    return true;
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
        bool fused = false;
        if (h->nf == 2)
            fused = h->g.dim == 2 ? launch_fused_t<2, 2>(h, J, x, y) : launch_fused_t<2, 3>(h, J, x, y);
        else
            fused = h->g.dim == 2 ? launch_fused_t<3, 2>(h, J, x, y) : launch_fused_t<3, 3>(h, J, x, y);
        if (fused) return;
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
