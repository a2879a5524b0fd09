// tpb_blas.cu - K10/K11 vector kernels: fused multi-dot, norms, axpy family.
//
// Stand in for PETSc VecMDot / VecNorm / VecMAXPY / VecAXPY used by KSPGMRES (classical
// Gram-Schmidt: all k dots of one Arnoldi step are fused into one pass and one reduction) and by
// SNES (norms).  Reductions: warp shuffles -> one partial per block -> the last block to finish
// (atomic ticket) folds the partials in a fixed order, so results are deterministic; multi-rank
// runs then sum the k results with a single ncclAllReduce.
#include "tpb_internal.cuh"

namespace {

constexpr int RB = 256;      // threads per reduction block
constexpr int KC = 8;        // vectors per multi-dot / maxpy pass
constexpr int MAXBLK = 1184; // 148 SMs x 8

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

template <int K>
__global__ void __launch_bounds__(RB) mdot_kernel(size_t n, const double* __restrict__ x, const double* __restrict__ Y,
                                                  size_t ldy, int kact, double* __restrict__ partial,
                                                  unsigned int* __restrict__ counter, double* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    double acc[K];
#pragma unroll
    for (int j = 0; j < K; j++) acc[j] = 0.0;
    for (size_t i = (size_t)blockIdx.x * RB + threadIdx.x; i < n; i += (size_t)gridDim.x * RB) {
        double xv = x[i];
#pragma unroll
        for (int j = 0; j < K; j++)
            if (j < kact) acc[j] = fma(xv, Y[(size_t)j * ldy + i], acc[j]);
    }
    __shared__ double sm[K][RB / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < K; j++) {
        double v = warp_sum(acc[j]);
        if (lane == 0) sm[j][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double v = 0.0;
        for (int w = 0; w < RB / 32; w++) v += sm[threadIdx.x][w];
        partial[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int ticket = atomicInc(counter, gridDim.x - 1);
        last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        // warp j folds the partials of vector j in a fixed order
        for (int j = wid; j < kact; j += RB / 32) {
            double v = 0.0;
            for (unsigned b = lane; b < gridDim.x; b += 32) v += partial[(size_t)j * gridDim.x + b];
            v = warp_sum(v);
            if (lane == 0) out[j] = v;
        }
    }
}

struct Coef {
    double c[KC];
};

template <int K>
__global__ void __launch_bounds__(256) maxpy_kernel(size_t n, double* __restrict__ y, const double* __restrict__ V,
                                                    size_t ldv, int kact, Coef cf, double scale) {
    pdl_launch_dependents();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        double a = y[i];
#pragma unroll
        for (int j = 0; j < K; j++)
            if (j < kact) a = fma(-cf.c[j], V[(size_t)j * ldv + i], a);
        y[i] = a * scale;
    }
}

__global__ void axpby_kernel(size_t n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
    pdl_launch_dependents();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = (b == 0.0) ? a * x[i] : fma(a, x[i], b * y[i]);
}
__global__ void waxpy_kernel(size_t n, double a, const double* __restrict__ x, const double* __restrict__ y,
                             double* __restrict__ w) {
    pdl_launch_dependents();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        w[i] = fma(a, x[i], y[i]);
}
__global__ void scale_kernel(size_t n, double a, double* __restrict__ x) {
    pdl_launch_dependents();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] *= a;
}
__global__ void minmax_kernel(size_t n, const double* __restrict__ x, double* __restrict__ partial,
                              unsigned int* __restrict__ counter, double* __restrict__ out) {
    double mn = 1e300, mx = -1e300;
    for (size_t i = (size_t)blockIdx.x * RB + threadIdx.x; i < n; i += (size_t)gridDim.x * RB) {
        double v = x[i];
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    __shared__ double smn[RB / 32], smx[RB / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
        smn[wid] = mn;
        smx[wid] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < RB / 32; w++) {
            mn = fmin(mn, smn[w]);
            mx = fmax(mx, smx[w]);
        }
        partial[blockIdx.x] = mn;
        partial[gridDim.x + blockIdx.x] = mx;
        __threadfence();
        unsigned int ticket = atomicInc(counter, gridDim.x - 1);
        last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double a = 1e300, b = -1e300;
        for (unsigned i = 0; i < gridDim.x; i++) {
            a = fmin(a, partial[i]);
            b = fmax(b, partial[gridDim.x + i]);
        }
        out[0] = a;
        out[1] = b;
    }
}
// total oil mass of the slab: sum over cells of vol * phi * S_o * rho_o(p, T)  (thermalmodel.py:190); same
// two-stage deterministic reduction as the dots
__global__ void __launch_bounds__(RB) oil_mass_kernel(size_t n, const double* __restrict__ u, const double* __restrict__ phi,
                                                      DevParams P, double vol, double* __restrict__ partial,
                                                      unsigned int* __restrict__ counter, double* __restrict__ out) {
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * RB + threadIdx.x; i < n; i += (size_t)gridDim.x * RB)
        acc += phi[i] * u[2 * n + i] * oil_rho_v(P, u[i], u[n + i]);
    __shared__ double sm[RB / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    acc = warp_sum(acc);
    if (lane == 0) sm[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int w = 0; w < RB / 32; w++) v += sm[w];
        partial[blockIdx.x] = v;
        __threadfence();
        unsigned int ticket = atomicInc(counter, gridDim.x - 1);
        last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        __threadfence();
        double v = 0.0;
        for (unsigned b = lane; b < gridDim.x; b += 32) v += partial[b];
        v = warp_sum(v);
        if (lane == 0) out[0] = v * vol;
    }
}
__global__ void clip_kernel(size_t n, double* __restrict__ x, double lo, double hi) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] = fmin(fmax(x[i], lo), hi);
}

inline unsigned grid_for(size_t n, int threads) {
    size_t b = (n + threads - 1) / threads;
    if (b > MAXBLK) b = MAXBLK;
    if (b < 1) b = 1;
    return (unsigned)b;
}

void ensure_red(tpb_handle_s* h) {
    if (h->red_partial) return;
    h->red_cap = 256;
    h->red_partial = tpb_dalloc<double>((size_t)KC * MAXBLK);
    h->red_counter = tpb_dalloc<unsigned int>(1);
    TPB_CUDA(cudaMemsetAsync(h->red_counter, 0, sizeof(unsigned int), h->stream));
    h->red_out = tpb_dalloc<double>(h->red_cap + 8);
    TPB_CUDA(cudaMallocHost(&h->red_host, (h->red_cap + 8) * sizeof(double)));
}

}  // namespace

void tpb_axpy(tpb_handle_s* h, size_t n, double a, const double* x, double* y) { tpb_axpby(h, n, a, x, 1.0, y); }
void tpb_axpby(tpb_handle_s* h, size_t n, double a, const double* x, double b, double* y) {
    launch_pdl(axpby_kernel, grid_for(n, 256), 256, h->stream, n, a, x, b, y);
    h->launches++;
}
void tpb_waxpy(tpb_handle_s* h, size_t n, double a, const double* x, const double* y, double* w) {
    launch_pdl(waxpy_kernel, grid_for(n, 256), 256, h->stream, n, a, x, y, w);
    h->launches++;
}
void tpb_scale(tpb_handle_s* h, size_t n, double a, double* x) {
    launch_pdl(scale_kernel, grid_for(n, 256), 256, h->stream, n, a, x);
    h->launches++;
}
void tpb_copy(tpb_handle_s* h, size_t n, const double* x, double* y) {
    TPB_CUDA(cudaMemcpyAsync(y, x, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
}
void tpb_zero(tpb_handle_s* h, size_t n, double* x) { TPB_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), h->stream)); }

// k dots <x, Y_j> queued on the stream; results land in h->red_out[out_off + j] (device)
void tpb_mdot_dev(tpb_handle_s* h, size_t n, const double* x, const double* Y, size_t ldy, int k, int out_off) {
    ensure_red(h);
    TPB_REQUIRE(out_off + k <= h->red_cap, TPB_ERR_ARG, "mdot: too many vectors");
    unsigned blocks = grid_for(n, RB);
    for (int j0 = 0; j0 < k; j0 += KC) {
        int kact = k - j0 < KC ? k - j0 : KC;
        launch_pdl(mdot_kernel<KC>, blocks, RB, h->stream, n, x, Y + (size_t)j0 * ldy, ldy, kact, h->red_partial,
                   h->red_counter, h->red_out + out_off + j0);
        h->launches++;
    }
}

// sum the first `count` queued results over the ranks, copy them to the host and synchronise
void tpb_red_get(tpb_handle_s* h, int count, double* host_out) {
    ensure_red(h);
    if (!tpb_allreduce_sum_hot(h, h->red_out, count, h->red_host))
        TPB_CUDA(cudaMemcpyAsync(h->red_host, h->red_out, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    for (int j = 0; j < count; j++) host_out[j] = h->red_host[j];
}

void tpb_mdot(tpb_handle_s* h, size_t n, const double* x, const double* Y, size_t ldy, int k, double* host_out) {
    tpb_mdot_dev(h, n, x, Y, ldy, k, 0);
    tpb_red_get(h, k, host_out);
}

double tpb_dot_sync(tpb_handle_s* h, size_t n, const double* x, const double* y) {
    double r;
    tpb_mdot(h, n, x, y, 0, 1, &r);
    return r;
}
double tpb_norm2(tpb_handle_s* h, size_t n, const double* x) { return sqrt(tpb_dot_sync(h, n, x, x)); }

void tpb_maxpy(tpb_handle_s* h, size_t n, double* y, const double* V, size_t ldv, int k, const double* c) {
    for (int j0 = 0; j0 < k; j0 += KC) {
        int kact = k - j0 < KC ? k - j0 : KC;
        Coef cf;
        for (int j = 0; j < KC; j++) cf.c[j] = j < kact ? c[j0 + j] : 0.0;
        launch_pdl(maxpy_kernel<KC>, grid_for(n, 256), 256, h->stream, n, y, V + (size_t)j0 * ldv, ldv, kact, cf, 1.0);
        h->launches++;
    }
}

// y = scale * (y - sum_j c[j] V_j) for one group of <= KC contiguous vectors
void tpb_maxpy_scale(tpb_handle_s* h, size_t n, double* y, const double* V, size_t ldv, int k, const double* c,
                     double scale) {
    TPB_REQUIRE(k <= KC, TPB_ERR_ARG, "maxpy_scale: group too large");
    Coef cf;
    for (int j = 0; j < KC; j++) cf.c[j] = j < k ? c[j] : 0.0;
    launch_pdl(maxpy_kernel<KC>, grid_for(n, 256), 256, h->stream, n, y, V, ldv, k, cf, scale);
    h->launches++;
}

void tpb_minmax_impl(tpb_handle_s* h, size_t n, const double* x, double* out2) {
    ensure_red(h);
    unsigned blocks = grid_for(n, RB);
    minmax_kernel<<<blocks, RB, 0, h->stream>>>(n, x, h->red_partial, h->red_counter, h->red_out);
    h->launches++;
    TPB_CUDA(cudaMemcpyAsync(h->red_host, h->red_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    out2[0] = h->red_host[0];
    out2[1] = h->red_host[1];
}
// this slab's oil mass -> h->red_out[0] (device); the caller reduces over the ranks and copies out
void tpb_oil_mass_impl(tpb_handle_s* h, const double* u) {
    ensure_red(h);
    unsigned blocks = grid_for(h->g.n, RB);
    oil_mass_kernel<<<blocks, RB, 0, h->stream>>>((size_t)h->g.n, u, h->fld[TPB_PHI], h->dp, h->g.vol, h->red_partial,
                                                  h->red_counter, h->red_out);
    h->launches++;
}
void tpb_clip_impl(tpb_handle_s* h, size_t n, double* x, double lo, double hi) {
    clip_kernel<<<grid_for(n, 256), 256, 0, h->stream>>>(n, x, lo, hi);
    h->launches++;
}
