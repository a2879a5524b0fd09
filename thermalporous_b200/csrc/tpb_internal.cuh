// tpb_internal.cuh - shared declarations of libtpb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <functional>
#include <string>
#include <vector>

#include "tpb200.h"

#define TPB_MAXF 3

struct tpb_exception {
    int code;
    std::string msg;
};

#define TPB_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            throw tpb_exception{TPB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)}; \
        }                                                                                \
    } while (0)

#define TPB_REQUIRE(cond, code, text)            \
    do {                                         \
        if (!(cond)) throw tpb_exception{code, text}; \
    } while (0)

// geometry of the local slab, passed by value to kernels
struct Geom {
    int dim, nx, ny, nz;
    int np;          // cells per plane of the slab axis (nx*ny in 3-D, nx in 2-D)
    int nl;          // planes in the slab (nz in 3-D, ny in 2-D)
    int has_lo, has_hi;
    long long n;     // owned cells
    double h[3];     // centre distances Dx, Dy, Dz
    double ih[3];    // 1 / h
    double area[3];  // facet measure per axis
    double vol;
};

// physical constants in device form
struct DevParams {
    double ko, kw, kr, c_v_w, c_v_o, c_r, rho_r, T_inj, T_prod, U;
    double g;            // 0 when gravity is off or dim == 2
    double Wp, Wo;       // p_weight, o_weight (twophase.py:144-147)
    // oil_rho / oil_mu coefficients (physicalparameters.py:37-57)
    double rho_ref, mu_o_pref, mu_o_exp;
};

// a cell field plus the two ghost planes of the slab axis
struct GField {
    const double* v;
    const double* lo;
    const double* hi;
};

__host__ __device__ inline int tpb_ns(int dim) { return dim == 3 ? 7 : 5; }

// (i, j, k) of a linear index in 32-bit arithmetic: a slab holds fewer than 2^31 cells (checked in tpb_create), and
// 64-bit integer division is emulated on the GPU - two of them cost more than the rest of a small stencil kernel
__device__ __forceinline__ void tpb_ijk(long long c, int nx, int ny, int& i, int& j, int& k) {
    const unsigned cu = (unsigned)c;
    const unsigned t = cu / (unsigned)nx;
    i = (int)(cu - t * (unsigned)nx);
    const unsigned kk = t / (unsigned)ny;
    j = (int)(t - kk * (unsigned)ny);
    k = (int)kk;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): the ~190 small kernels of one PC application depend on each other in a
// chain, and most of a small level's pass is launch latency plus two dependent rounds of loads.  Kernels launched
// through launch_pdl() may start while their predecessor drains: they decode their cell and load what does not
// change during a solve (their row of the operator, decoupling weights) first, and only then wait for the
// predecessor's results (pdl_wait), so one round of loads and the launch overlap the previous kernel.  Rules:
// every thread of a kernel launched this way executes pdl_wait before it exits (a kernel that finished without
// waiting would let ITS successor overtake the predecessor), and nothing a predecessor writes is touched before it.
// pdl_launch_dependents at the top of a kernel lets the successor start as soon as all blocks are resident.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// launch with programmatic stream serialisation; TPB_PDL=0 launches normally
bool tpb_pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_pdl_smem(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                            Args... args);
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, Args... args) {
    launch_pdl_smem(kernel, grid, block, 0, st, args...);
}
// the same with dynamic shared memory (kernels that need more than 48 KB opt in once with cudaFuncSetAttribute)
template <typename... KArgs, typename... Args>
inline void launch_pdl_smem(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                            Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = tpb_pdl_enabled() ? 1 : 0;
    TPB_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

// ---------------------------------------------------------------------------------------------
// Peer-memory mailboxes (tpb_comm.cu): every rank owns one cudaMalloc'ed mailbox, mapped into every other rank's
// process with CUDA IPC, so a kernel can store straight into a peer's HBM over NVLink/NVSwitch.  A message is
// data stores + fence.sys + a release store of the slot's epoch number into the receiver's flag word; the receiver
// spins on its own (local) flag with acquire loads and reads the data with L1-bypassing loads.  Data areas are
// double-buffered on the epoch's parity: a sender can only be one message ahead of a receiver, because message
// e+1 of any slot is sent after the sender has consumed the receiver's message e.  Spins are bounded; a time-out
// raises the error word, which the host turns into TPB_ERR_NCCL at the next synchronisation.
// ---------------------------------------------------------------------------------------------
constexpr int P2P_MAXR = 16;        // ranks
constexpr int P2P_AR_MAX = 256;     // doubles per all-reduce
enum { P2P_SLOT_AR = 0, P2P_SLOT_HALO_LO = 1, P2P_SLOT_HALO_HI = 2, P2P_SLOT_MG = 3 /* + hierarchy (0|1) */, P2P_NSLOT = 8 };
constexpr long long P2P_SPIN_MAX = 4000000;   // with the back-off in p2p_wait: tens of seconds

struct P2PView {
    char* box[P2P_MAXR];            // box[r]: rank r's mailbox as mapped in this process
    unsigned long long* epoch;      // this rank's per-slot message counters (device)
    int* err;                       // time-out word (device)
    int rank, nranks;
    long long off_ar, off_halo_lo, off_halo_hi, off_mg;   // byte offsets of the data areas inside a mailbox
    long long halo_cap, mg_cap;     // doubles per parity buffer
};

__device__ __forceinline__ unsigned long long* p2p_flag(const P2PView& v, int r, int slot, int src) {
    return reinterpret_cast<unsigned long long*>(v.box[r]) + slot * P2P_MAXR + src;
}
__device__ __forceinline__ void p2p_store_flag(unsigned long long* f, unsigned long long e) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(e) : "memory");
}
__device__ __forceinline__ unsigned long long p2p_load_flag(const unsigned long long* f) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
    return v;
}
// wait until rank `src` has delivered message `e` of `slot` into THIS rank's mailbox
__device__ __forceinline__ bool p2p_wait(const P2PView& v, int slot, int src, unsigned long long e) {
    const unsigned long long* f = p2p_flag(v, v.rank, slot, src);
    for (long long spin = 0; spin < P2P_SPIN_MAX; spin++) {
        if (p2p_load_flag(f) >= e) return true;
        __nanosleep(spin < 4096 ? 32 : 4000);   // a peer that is merely late (allocation, graph instantiation) gets ~ 20 s
    }
    atomicExch(v.err, 1);
    return false;
}
__device__ __forceinline__ double* p2p_ar_area(const P2PView& v, int r, int par, int src) {
    return reinterpret_cast<double*>(v.box[r] + v.off_ar) + ((long long)par * P2P_MAXR + src) * P2P_AR_MAX;
}
__device__ __forceinline__ double* p2p_halo_area(const P2PView& v, int r, bool from_lo, int par) {
    return reinterpret_cast<double*>(v.box[r] + (from_lo ? v.off_halo_lo : v.off_halo_hi)) + (long long)par * v.halo_cap;
}
__device__ __forceinline__ double* p2p_mg_area(const P2PView& v, int r, int hier, int par) {
    return reinterpret_cast<double*>(v.box[r] + v.off_mg) + ((long long)hier * 2 + par) * v.mg_cap;
}
// In-place all-gather of a small vector by ONE thread block (the multigrid gather level, <= mg_cap doubles): this
// rank's section [my_off, my_off + my_cnt) of buf goes into every peer's mailbox, the flags are exchanged, and the
// peers' sections are copied out of the local mailbox.  Ends with a block barrier: buf is complete for every thread.
__device__ __forceinline__ void p2p_gather_block(const P2PView& v, int hier, double* buf, long long my_off,
                                                 long long my_cnt, long long total) {
    __shared__ unsigned long long s_ep;
    if (threadIdx.x == 0) s_ep = ++v.epoch[P2P_SLOT_MG + hier];
    __syncthreads();
    const unsigned long long e = s_ep;
    const int par = (int)(e & 1);
    for (int r = 0; r < v.nranks; r++) {
        if (r == v.rank) continue;
        double* dst = p2p_mg_area(v, r, hier, par) + my_off;
        for (long long i = threadIdx.x; i < my_cnt; i += blockDim.x) dst[i] = buf[my_off + i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < v.nranks && threadIdx.x != v.rank) {
        p2p_store_flag(p2p_flag(v, threadIdx.x, P2P_SLOT_MG + hier, v.rank), e);
        p2p_wait(v, P2P_SLOT_MG + hier, threadIdx.x, e);
    }
    __syncthreads();
    const double* src = p2p_mg_area(v, v.rank, hier, par);
    for (long long i = threadIdx.x; i < total; i += blockDim.x)
        if (i < my_off || i >= my_off + my_cnt) buf[i] = __ldcg(src + i);
    __syncthreads();
}
// true (and *v filled) when the handle has working mailboxes; `want` = bit of TPB_P2P (1 all-reduce, 2 halo, 4 multigrid)
bool tpb_p2p_view(tpb_handle_s* h, int want, P2PView* v);
// mailboxes + the three ticket counters of the fused halo kernels (TPB_P2P bit 8; needs bit 2); false = unfused path
bool tpb_p2p_halo(tpb_handle_s* h, P2PView* v, unsigned int** tickets);
bool tpb_p2p_gather(tpb_handle_s* h, int hier, double* buf, long long my_off, long long my_cnt, long long total);
void tpb_p2p_check(tpb_handle_s* h);
int tpb_comm_peer_mode_impl(tpb_handle_s* h);   // throws if a peer-memory wait timed out

// ---------------------------------------------------------------------------------------------
// forward-mode dual numbers (value + N partials); everything is unrolled into registers
// ---------------------------------------------------------------------------------------------
template <int N>
struct Dual {
    double v;
    double d[N > 0 ? N : 1];
};

template <int N>
__device__ __forceinline__ Dual<N> dconst(double a) {
    Dual<N> r;
    r.v = a;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = 0.0;
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> dvar(double a, int slot) {
    Dual<N> r;
    r.v = a;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = (i == slot) ? 1.0 : 0.0;
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r;
    r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r;
    r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(const Dual<N>& a) {
    Dual<N> r;
    r.v = -a.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = -a.d[i];
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r;
    r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator*(double s, const Dual<N>& a) {
    Dual<N> r;
    r.v = s * a.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = s * a.d[i];
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, double s) { return s * a; }
template <int N>
__device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, double s) {
    Dual<N> r = a;
    r.v += s;
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, double s) {
    Dual<N> r = a;
    r.v -= s;
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(double s, const Dual<N>& a) {
    Dual<N> r;
    r.v = s - a.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = -a.d[i];
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r;
    double ib = 1.0 / b.v;
    r.v = a.v * ib;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator/(double s, const Dual<N>& b) {
    Dual<N> r;
    double ib = 1.0 / b.v;
    r.v = s * ib;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = -r.v * b.d[i] * ib;
    return r;
}
// place a Dual<M> into slots [OFF, OFF+M) of a Dual<N>
template <int N, int OFF, int M>
__device__ __forceinline__ Dual<N> dembed(const Dual<M>& a) {
    Dual<N> r;
    r.v = a.v;
#pragma unroll
    for (int i = 0; i < N; i++) r.d[i] = (i >= OFF && i < OFF + M) ? a.d[(i - OFF) < M ? (i - OFF) : 0] : 0.0;
    return r;
}

// ---------------------------------------------------------------------------------------------
// properties with partials w.r.t. (p, T); physicalparameters.py:37-98
// ---------------------------------------------------------------------------------------------
// oil_rho = rho_ref * e^{5.5e-5 (10 p - 1.01325)} * e^{-2.5e-4 (T - 288.7056)}   (:37-46)
__device__ __forceinline__ void oil_rho_d(const DevParams& P, double p, double T, double& r, double& r_p,
                                          double& r_T) {
    r = P.rho_ref * exp(5.5e-5 * (p * 10.0 - 1.01325) - 2.5e-4 * (T - (15.5556 + 273.15)));
    r_p = 5.5e-4 * r;
    r_T = -2.5e-4 * r;
}
__device__ __forceinline__ double oil_rho_v(const DevParams& P, double p, double T) {
    return P.rho_ref * exp(5.5e-5 * (p * 10.0 - 1.01325) - 2.5e-4 * (T - (15.5556 + 273.15)));
}
// 1/oil_mu, oil_mu = 1e-3 * 10^{A1 API + A2} * Tf^{A3 API + A4}, Tf = 1.8 (T - 273.15) + 32   (:48-57)
__device__ __forceinline__ void oil_imu_d(const DevParams& P, double T, double& im, double& im_T) {
    double Tf = 1.8 * (T - 273.15) + 32.0;
    // Tf^e as exp(e log Tf): |e log Tf| < 40, so the relative error is < 1e-14 (parity tolerance 1e-12) at half
    // the cost of the correctly rounded pow()
    im = exp(-P.mu_o_exp * log(Tf)) / P.mu_o_pref;
    im_T = -im * P.mu_o_exp * 1.8 / Tf;  // d(1/mu)/dT = -(1/mu) * exp * Tf'/Tf
}
// water_rho: Trangenstein/Kell (:69-82), Tc = T - 272.15
__device__ __forceinline__ void water_rho_d(double p, double T, double& r, double& r_p, double& r_T) {
    const double E0 = 999.83952, E1 = 16.955176, E2 = -7.987e-3, E3 = -46.170461e-6, E4 = 105.56302e-9,
                 E5 = -280.54353e-12, E6 = 16.87985e-3, E7 = 10.2, Cw = 3.98854e-4;
    double Tc = T - 272.15;
    double poly = E0 + Tc * (E1 + Tc * (E2 + Tc * (E3 + Tc * (E4 + Tc * E5))));
    double dpoly = E1 + Tc * (2.0 * E2 + Tc * (3.0 * E3 + Tc * (4.0 * E4 + Tc * 5.0 * E5)));
    double den = 1.0 / (1.0 + E6 * Tc);
    double ex = exp(Cw * (p - E7));
    r = poly * ex * den;
    r_p = Cw * r;
    r_T = (dpoly - poly * E6 * den) * ex * den;
}
__device__ __forceinline__ double water_rho_v(double p, double T) {
    double r, a, b;
    water_rho_d(p, T, r, a, b);
    return r;
}
// 1/water_mu, Grabowski (:84-90), Tf = 1.8 (T - 272.15) + 32
__device__ __forceinline__ void water_imu_d(double T, double& im, double& im_T) {
    const double Aw = 2.1850, Bw = 0.04012, Cw = 5.1547e-6;
    double Tf = 1.8 * (T - 272.15) + 32.0;
    double q = -1.0 + Bw * Tf + Cw * Tf * Tf;
    im = q / (1e-3 * Aw);
    im_T = (Bw + 2.0 * Cw * Tf) * 1.8 / (1e-3 * Aw);
}

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
struct ScalarStencil;  // tpb_pc.cu
struct PcState;
struct KspState;
struct CommState;

struct tpb_handle_s {
    int device = 0;
    int nphase = 1, nf = 2, ns = 5;
    Geom g{};
    DevParams dp{};
    tpb_params prm{};
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;

    // static fields: owned + ghost planes
    double* fld[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double* fld_lo[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double* fld_hi[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool fld_set[5] = {false, false, false, false, false};
    // state ghost planes (nf * np each)
    double* u_lo = nullptr;
    double* u_hi = nullptr;
    // per-cell property scratch of the assembly (owned + ghost planes), tpb_assemble.cu
    double* scr = nullptr;
    // static face transmissibilities area*K_facet per axis (+ the slab's bottom faces), rebuilt when a field changes
    double* trans[3] = {nullptr, nullptr, nullptr};
    double* trans_lo = nullptr;
    bool trans_dirty = true;
    // generic vector ghost planes for SpMV inputs (nf * np each)
    double* x_lo = nullptr;
    double* x_hi = nullptr;

    // sources grouped by cell
    int nsrc_cells = 0;
    int64_t* src_cell = nullptr;  // unique cells
    int* src_off = nullptr;       // CSR offsets into src_ent
    tpb_source* src_ent = nullptr;
    int* src_index = nullptr;     // per cell: index into src_cell / src_acc, or -1
    double* src_acc = nullptr;    // nsrc_cells * (nf + nf*nf) contributions of the current assembly
    cudaStream_t stream2 = nullptr;   // side stream (source cells run beside the property pre-pass)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    // reductions
    double* red_partial = nullptr;   // per-block partials
    unsigned int* red_counter = nullptr;
    double* red_out = nullptr;       // device results
    double* red_host = nullptr;      // pinned host mirror
    int red_cap = 0;

    tpb_solver_opts opts{};
    PcState* pc = nullptr;
    KspState* ksp = nullptr;
    CommState* comm = nullptr;

    // Newton workspace
    double* nw_F = nullptr;
    double* nw_J = nullptr;
    double* nw_du = nullptr;
    double* nw_utrial = nullptr;
    double* nw_Ftrial = nullptr;
    double* nw_uold = nullptr;
    double* nw_u = nullptr;
};

// ---- internal entry points shared between translation units -----------------------------------
void tpb_launch_assemble(tpb_handle_s* h, const double* u, const double* u_old, double dt, double* F, double* J);
void tpb_launch_spmv(tpb_handle_s* h, const double* J, const double* x, double* y);
void tpb_halo_vector(tpb_handle_s* h, const double* x, int nfields, double* lo, double* hi);
void tpb_allreduce_sum(tpb_handle_s* h, double* dev_buf, int count);
bool tpb_allreduce_sum_hot(tpb_handle_s* h, double* dev_buf, int count, double* host_out);
void tpb_allreduce_max(tpb_handle_s* h, double* dev_buf, int count);
int tpb_comm_rank(tpb_handle_s* h);
int tpb_comm_size(tpb_handle_s* h);
const std::vector<int>& tpb_comm_planes(tpb_handle_s* h);
void tpb_allgatherv(tpb_handle_s* h, double* buf, const long long* off, const long long* cnt, int nrep, long long stride);
// A communication step (NCCL) inside code that may be under CUDA-graph capture (the PC application): run now when
// nothing is being captured; otherwise the capture is cut into two graphs around it (tpb_pc.cu).
void tpb_comm_op(tpb_handle_s* h, const std::function<void()>& op);

// blas-1 (tpb_blas.cu)
void tpb_axpy(tpb_handle_s* h, size_t n, double a, const double* x, double* y);          // y += a x
void tpb_axpby(tpb_handle_s* h, size_t n, double a, const double* x, double b, double* y);  // y = a x + b y
void tpb_waxpy(tpb_handle_s* h, size_t n, double a, const double* x, const double* y, double* w);  // w = a x + y
void tpb_scale(tpb_handle_s* h, size_t n, double a, double* x);
void tpb_copy(tpb_handle_s* h, size_t n, const double* x, double* y);
void tpb_zero(tpb_handle_s* h, size_t n, double* x);
// k dots <x, Y_j> (Y_j = Y + j*ldy), results in h->red_host[0..k) after sync (global over ranks)
void tpb_mdot(tpb_handle_s* h, size_t n, const double* x, const double* Y, size_t ldy, int k, double* host_out);
void tpb_mdot_dev(tpb_handle_s* h, size_t n, const double* x, const double* Y, size_t ldy, int k, int out_off);
void tpb_red_get(tpb_handle_s* h, int count, double* host_out);
void tpb_maxpy_scale(tpb_handle_s* h, size_t n, double* y, const double* V, size_t ldv, int k, const double* c,
                     double scale);
double tpb_norm2(tpb_handle_s* h, size_t n, const double* x);
double tpb_dot_sync(tpb_handle_s* h, size_t n, const double* x, const double* y);
// y -= sum_j c[j] V_j  (c on host)
void tpb_maxpy(tpb_handle_s* h, size_t n, double* y, const double* V, size_t ldv, int k, const double* c);

// pc / ksp (tpb_pc.cu, tpb_solver.cu)
void tpb_pc_free(tpb_handle_s* h);
void tpb_ksp_free(tpb_handle_s* h);
void tpb_comm_free(tpb_handle_s* h);
void tpb_pc_setup_impl(tpb_handle_s* h, const double* J, const double* u, double dt);
void tpb_pc_apply_impl(tpb_handle_s* h, const double* x, double* y);
void tpb_ksp_solve_impl(tpb_handle_s* h, const double* J, const double* b, double* x, int* its, int* reason,
                        double* rnorm);
void tpb_newton_impl(tpb_handle_s* h, double* u, const double* u_old, double dt, tpb_stats* st);

template <typename T>
inline T* tpb_dalloc(size_t count) {
    T* p = nullptr;
    if (count == 0) count = 1;
    TPB_CUDA(cudaMalloc(&p, count * sizeof(T)));
    return p;
}
inline void tpb_dfree(void* p) {
    if (p) cudaFree(p);
}
