// tpb_comm.cu - slab-partition plumbing: ghost-plane exchange, scalar all-reduces and the multigrid gather.
//
// Stands in for the MPI traffic PETSc/PyOP2 generate under the reference (halo update before
// every assembly, MatMult ghost scatter, MPI_Allreduce per Krylov dot).  The path
// has three exchange patterns: one boundary plane per neighbour, a sum of k doubles over all ranks, and an
// all-gather of the multigrid gather level.  Two transports:
//   * peer-memory mailboxes (default on one NVLink/NVSwitch node): each rank's mailbox is mapped into every peer
//     with CUDA IPC and kernels store straight into the receiver's HBM (protocol: tpb_internal.cuh).  Kernels here:
//     halo_push/halo_pull, p2p_allreduce, mg_gather; the fused forms live with their compute (spmv_halo_kernel in
//     tpb_spmv.cu, tail_dist_kernel in tpb_pc.cu).
//   * NCCL (ncclSend/ncclRecv grouped, ncclAllReduce, ncclAllGather): the set-up traffic always, every exchange
//     when the mailboxes cannot be mapped or TPB_P2P=0.  NCCL is resolved at run time with dlopen("libnccl.so.2")
//     - the copy torch already mapped into the process - so the single-GPU library has no NCCL link dependency.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "tpb_internal.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi& api() {
    static NcclApi a;
    if (a.lib) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !a.lib; i++) a.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    TPB_REQUIRE(a.lib != nullptr, TPB_ERR_NCCL, "libnccl.so.2 not found (import torch first, or set LD_LIBRARY_PATH)");
#define L(field, sym)                                   \
    a.field = (decltype(a.field))dlsym(a.lib, sym);     \
    TPB_REQUIRE(a.field != nullptr, TPB_ERR_NCCL, "missing NCCL symbol " sym)
    L(GetUniqueId, "ncclGetUniqueId");
    L(CommInitRank, "ncclCommInitRank");
    L(CommDestroy, "ncclCommDestroy");
    L(AllReduce, "ncclAllReduce");
    L(Send, "ncclSend");
    L(Recv, "ncclRecv");
    L(Broadcast, "ncclBroadcast");
    L(AllGather, "ncclAllGather");
    L(GroupStart, "ncclGroupStart");
    L(GroupEnd, "ncclGroupEnd");
    L(GetErrorString, "ncclGetErrorString");
#undef L
    return a;
}

#define TPB_NCCL(call)                                                                       \
    do {                                                                                     \
        int r_ = (call);                                                                     \
        if (r_ != ncclSuccess)                                                               \
            throw tpb_exception{TPB_ERR_NCCL, std::string(#call) + ": " + api().GetErrorString(r_)}; \
    } while (0)

__global__ void pack_planes_kernel(const double* __restrict__ x, long long n, int np, int nfields,
                                   double* __restrict__ first, double* __restrict__ last) {
    // first[f*np + q] = x[f*n + q] ; last[f*np + q] = x[f*n + n - np + q]
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)np * nfields) return;
    int f = (int)(t / np);
    int q = (int)(t % np);
    first[t] = x[(long long)f * n + q];
    last[t] = x[(long long)f * n + n - np + q];
}

// ---- peer-memory kernels ---------------------------------------------------------------------------
// sum of `count` doubles over the ranks, in rank order (identical bits on every rank); one CTA
// host_out (may be null): pinned host memory that also receives the sums, saving the copy engine's round trip
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(P2PView v, double* buf, int count, double* host_out) {
    __shared__ unsigned long long ep;
    if (threadIdx.x == 0) ep = ++v.epoch[P2P_SLOT_AR];
    __syncthreads();
    const unsigned long long e = ep;
    const int par = (int)(e & 1);
    for (int r = 0; r < v.nranks; r++) {
        double* dst = p2p_ar_area(v, r, par, v.rank);
        for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = buf[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < v.nranks) {
        p2p_store_flag(p2p_flag(v, threadIdx.x, P2P_SLOT_AR, v.rank), e);
        p2p_wait(v, P2P_SLOT_AR, threadIdx.x, e);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        double sum = 0.0;
        for (int r = 0; r < v.nranks; r++) sum += __ldcg(p2p_ar_area(v, v.rank, par, r) + i);
        buf[i] = sum;
        if (host_out) host_out[i] = sum;
    }
    if (host_out) __threadfence_system();
}

// boundary planes of x straight into the neighbours' mailboxes; the last block to finish publishes the epoch
__global__ void __launch_bounds__(256) halo_push_kernel(P2PView v, const double* __restrict__ x, long long n, int np,
                                                        int nfields, int has_lo, int has_hi, unsigned int* ticket) {
    const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(&v.epoch[P2P_SLOT_HALO_LO]) + 1;
    const int par = (int)(e & 1);
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)np * nfields) {
        const int f = (int)(t / np), q = (int)(t % np);
        // my first plane is the upper ghost of the rank below; my last plane the lower ghost of the rank above
        if (has_lo) p2p_halo_area(v, v.rank - 1, false, par)[t] = x[(long long)f * n + q];
        if (has_hi) p2p_halo_area(v, v.rank + 1, true, par)[t] = x[(long long)f * n + n - np + q];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(&v.epoch[P2P_SLOT_HALO_LO]) = e;
        if (has_lo) p2p_store_flag(p2p_flag(v, v.rank - 1, P2P_SLOT_HALO_HI, v.rank), e);
        if (has_hi) p2p_store_flag(p2p_flag(v, v.rank + 1, P2P_SLOT_HALO_LO, v.rank), e);
    }
}

// wait for the neighbours' planes of the current epoch and copy them out of the mailbox
__global__ void __launch_bounds__(256) halo_pull_kernel(P2PView v, int np, int nfields, int has_lo, int has_hi,
                                                        double* __restrict__ lo, double* __restrict__ hi) {
    const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(&v.epoch[P2P_SLOT_HALO_LO]);
    const int par = (int)(e & 1);
    if (threadIdx.x == 0) {
        if (has_lo) p2p_wait(v, P2P_SLOT_HALO_LO, v.rank - 1, e);
        if (has_hi) p2p_wait(v, P2P_SLOT_HALO_HI, v.rank + 1, e);
    }
    __syncthreads();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)np * nfields) {
        if (has_lo) lo[t] = __ldcg(p2p_halo_area(v, v.rank, true, par) + t);
        if (has_hi) hi[t] = __ldcg(p2p_halo_area(v, v.rank, false, par) + t);
    }
}

// all-gather of the multigrid gather level's right-hand side: this rank's section goes into every rank's mailbox,
// then the whole level is copied out of the local mailbox once every section of the epoch has arrived; one CTA
__global__ void __launch_bounds__(1024) mg_gather_kernel(P2PView v, int hier, double* buf, long long my_off,
                                                         long long my_cnt, long long total) {
    p2p_gather_block(v, hier, buf, my_off, my_cnt, total);
}

}  // namespace

struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    double* send_first = nullptr;  // packed boundary planes (nf_max * np)
    double* send_last = nullptr;
    size_t cap = 0;
    std::vector<int> planes;       // owned planes of every rank along the slab axis (filled on first use)
    double* scratch = nullptr;     // nranks doubles
    // peer-memory mailboxes (tpb_internal.cuh)
    bool p2p_ok = false;
    int p2p_mask = 0;              // TPB_P2P bits: 1 all-reduce, 2 halo, 4 multigrid gather, 8 halo fused into SpMV
    char* my_box = nullptr;
    std::vector<char*> peer_box;   // opened IPC mappings (nullptr for this rank)
    unsigned long long* epoch = nullptr;
    int* err = nullptr;
    unsigned int* ticket = nullptr;
    P2PView view{};
};

void tpb_comm_free(tpb_handle_s* h) {
    if (!h->comm) return;
    if (h->comm->comm) api().CommDestroy(h->comm->comm);
    tpb_dfree(h->comm->send_first);
    tpb_dfree(h->comm->send_last);
    tpb_dfree(h->comm->scratch);
    for (char* q : h->comm->peer_box)
        if (q) cudaIpcCloseMemHandle(q);
    tpb_dfree(h->comm->my_box);
    tpb_dfree(h->comm->epoch);
    tpb_dfree(h->comm->err);
    tpb_dfree(h->comm->ticket);
    delete h->comm;
    h->comm = nullptr;
}

// exchange the first/last owned planes of an nfields-field vector with the slab neighbours:
// lo <- neighbour below's last plane, hi <- neighbour above's first plane
void tpb_halo_vector(tpb_handle_s* h, const double* x, int nfields, double* lo, double* hi) {
    if (!(h->g.has_lo || h->g.has_hi)) return;
    TPB_REQUIRE(h->comm && h->comm->comm, TPB_ERR_STATE, "slab has neighbours but tpb_comm_init was not called");
    CommState* c = h->comm;
    const int np = h->g.np;
    size_t cnt = (size_t)np * nfields;
    if (c->p2p_ok && (c->p2p_mask & 2) && (long long)cnt <= c->view.halo_cap) {
        const unsigned blocks = (unsigned)((cnt + 255) / 256);
        halo_push_kernel<<<blocks, 256, 0, h->stream>>>(c->view, x, h->g.n, np, nfields, h->g.has_lo, h->g.has_hi, c->ticket);
        halo_pull_kernel<<<blocks, 256, 0, h->stream>>>(c->view, np, nfields, h->g.has_lo, h->g.has_hi, lo, hi);
        h->launches += 2;
        return;
    }
    if (cnt > c->cap) {
        tpb_dfree(c->send_first);
        tpb_dfree(c->send_last);
        c->send_first = tpb_dalloc<double>(cnt);
        c->send_last = tpb_dalloc<double>(cnt);
        c->cap = cnt;
    }
    pack_planes_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, h->stream>>>(x, h->g.n, np, nfields, c->send_first,
                                                                            c->send_last);
    h->launches++;
    NcclApi& a = api();
    TPB_NCCL(a.GroupStart());
    if (h->g.has_lo) {
        TPB_NCCL(a.Send(c->send_first, cnt, ncclFloat64, c->rank - 1, c->comm, h->stream));
        TPB_NCCL(a.Recv(lo, cnt, ncclFloat64, c->rank - 1, c->comm, h->stream));
    }
    if (h->g.has_hi) {
        TPB_NCCL(a.Send(c->send_last, cnt, ncclFloat64, c->rank + 1, c->comm, h->stream));
        TPB_NCCL(a.Recv(hi, cnt, ncclFloat64, c->rank + 1, c->comm, h->stream));
    }
    TPB_NCCL(a.GroupEnd());
}

void tpb_allreduce_sum(tpb_handle_s* h, double* dev_buf, int count) {
    if (!h->comm || h->comm->nranks == 1) return;
    TPB_NCCL(api().AllReduce(dev_buf, dev_buf, (size_t)count, ncclFloat64, ncclSum, h->comm->comm, h->stream));
}

// the Krylov / Newton reductions: peer-memory all-reduce when the mailboxes are up, NCCL otherwise
// returns true when the sums were also stored to host_out (pinned) by the kernel itself
bool tpb_allreduce_sum_hot(tpb_handle_s* h, double* dev_buf, int count, double* host_out) {
    if (!h->comm || h->comm->nranks == 1) return false;
    CommState* c = h->comm;
    if (c->p2p_ok && (c->p2p_mask & 1) && count <= P2P_AR_MAX) {
        p2p_allreduce_kernel<<<1, 256, 0, h->stream>>>(c->view, dev_buf, count, host_out);
        h->launches++;
        return host_out != nullptr;
    }
    tpb_allreduce_sum(h, dev_buf, count);
    return false;
}

// in-place all-gather of one small vector through the mailboxes (the multigrid gather level, once per V-cycle);
// false when the mailboxes are not available or the level does not fit - the caller then goes through NCCL
bool tpb_p2p_gather(tpb_handle_s* h, int hier, double* buf, long long my_off, long long my_cnt, long long total) {
    if (!h->comm || !h->comm->p2p_ok || !(h->comm->p2p_mask & 4)) return false;
    CommState* c = h->comm;
    if (total > c->view.mg_cap || hier < 0 || hier > 1) return false;
    mg_gather_kernel<<<1, 1024, 0, h->stream>>>(c->view, hier, buf, my_off, my_cnt, total);
    h->launches++;
    return true;
}

bool tpb_p2p_halo(tpb_handle_s* h, P2PView* v, unsigned int** tickets) {
    if (!h->comm || !h->comm->p2p_ok || (h->comm->p2p_mask & 10) != 10) return false;
    *v = h->comm->view;
    *tickets = h->comm->ticket + 1;
    return true;
}

bool tpb_p2p_view(tpb_handle_s* h, int want, P2PView* v) {
    if (!h->comm || !h->comm->p2p_ok || !(h->comm->p2p_mask & want)) return false;
    *v = h->comm->view;
    return true;
}

void tpb_p2p_check(tpb_handle_s* h) {
    if (!h->comm || !h->comm->p2p_ok) return;
    int e = 0;
    TPB_CUDA(cudaMemcpyAsync(&e, h->comm->err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    TPB_REQUIRE(e == 0, TPB_ERR_NCCL, "a peer-memory exchange timed out (a rank stopped or left the common call sequence)");
}

// Set up the mailboxes: allocate, exchange the CUDA IPC handles over NCCL, map the peers' boxes.  Any failure (IPC
// not permitted in this container, no peer access) leaves the NCCL paths in charge.  TPB_P2P=0 disables, default 15
// (1 Krylov all-reduce, 2 halo planes, 4 multigrid gather, 8 halo fused into the SpMV kernel).
static void p2p_init(tpb_handle_s* h) {
    CommState* c = h->comm;
    const int mask = getenv("TPB_P2P") ? atoi(getenv("TPB_P2P")) : 15;
    if (mask == 0 || c->nranks < 2 || c->nranks > P2P_MAXR) return;
    NcclApi& a = api();
    P2PView& v = c->view;
    memset(&v, 0, sizeof(v));
    v.rank = c->rank;
    v.nranks = c->nranks;
    v.halo_cap = (long long)h->g.np * TPB_MAXF;
    v.mg_cap = 8192;
    long long off = 4096;   // flags: P2P_NSLOT * P2P_MAXR words
    v.off_ar = off;
    off += 2LL * P2P_MAXR * P2P_AR_MAX * 8;
    v.off_halo_lo = off;
    off += 2 * v.halo_cap * 8;
    v.off_halo_hi = off;
    off += 2 * v.halo_cap * 8;
    v.off_mg = off;
    off += 4 * v.mg_cap * 8;
    const size_t bytes = (size_t)off;
    // zeroed on the handle's stream (every consumer runs there; it is non-blocking, so the legacy default stream
    // would not be ordered against it) and complete before any peer can learn the handle: the all-gather below
    // synchronises the stream first
    bool fine = cudaMalloc(&c->my_box, bytes) == cudaSuccess && cudaMemsetAsync(c->my_box, 0, bytes, h->stream) == cudaSuccess &&
                cudaStreamSynchronize(h->stream) == cudaSuccess;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    fine = fine && cudaIpcGetMemHandle(&mine, c->my_box) == cudaSuccess;
    if (!fine) cudaGetLastError();
    // every rank takes part in the exchange whatever happened locally: [ok flag | 64-byte handle] as 9 doubles
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    const int W = 9;
    std::vector<double> all((size_t)W * c->nranks, 0.0), me(W, 0.0);
    me[0] = fine ? 1.0 : 0.0;
    memcpy(&me[1], &mine, 64);
    struct DevBuf {   // freed on every way out, including a throwing TPB_CUDA / TPB_NCCL
        double* p;
        ~DevBuf() { tpb_dfree(p); }
    } devbuf{tpb_dalloc<double>((size_t)W * c->nranks)};
    double* dev = devbuf.p;
    TPB_CUDA(cudaMemcpyAsync(dev + (size_t)W * c->rank, me.data(), W * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    TPB_NCCL(a.AllGather(dev + (size_t)W * c->rank, dev, (size_t)W, ncclFloat64, c->comm, h->stream));
    TPB_CUDA(cudaMemcpyAsync(all.data(), dev, all.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    c->peer_box.assign(c->nranks, nullptr);
    for (int r = 0; r < c->nranks; r++) fine = fine && all[(size_t)W * r] == 1.0;
    for (int r = 0; r < c->nranks && fine; r++) {
        if (r == c->rank) {
            v.box[r] = c->my_box;
            continue;
        }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, &all[(size_t)W * r + 1], 64);
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            fine = false;
            break;
        }
        c->peer_box[r] = (char*)q;
        v.box[r] = (char*)q;
    }
    // agree on the outcome: everybody or nobody
    me[0] = fine ? 0.0 : 1.0;
    TPB_CUDA(cudaMemcpyAsync(dev, me.data(), sizeof(double), cudaMemcpyHostToDevice, h->stream));
    TPB_NCCL(a.AllReduce(dev, dev, 1, ncclFloat64, ncclSum, c->comm, h->stream));
    TPB_CUDA(cudaMemcpyAsync(me.data(), dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    if (me[0] != 0.0) {
        // some rank could not map a peer: NCCL paths stay in charge; give the mailbox and the mappings back
        for (int r = 0; r < c->nranks; r++)
            if (c->peer_box[r]) {
                cudaIpcCloseMemHandle(c->peer_box[r]);
                c->peer_box[r] = nullptr;
            }
        tpb_dfree(c->my_box);
        c->my_box = nullptr;
        cudaGetLastError();
        return;
    }
    c->epoch = tpb_dalloc<unsigned long long>(P2P_NSLOT);
    c->err = tpb_dalloc<int>(1);
    c->ticket = tpb_dalloc<unsigned int>(4);   // [0] push/pull kernels, [1..3] fused halo kernels
    TPB_CUDA(cudaMemsetAsync(c->epoch, 0, P2P_NSLOT * sizeof(unsigned long long), h->stream));
    TPB_CUDA(cudaMemsetAsync(c->err, 0, sizeof(int), h->stream));
    TPB_CUDA(cudaMemsetAsync(c->ticket, 0, 4 * sizeof(unsigned int), h->stream));
    TPB_CUDA(cudaStreamSynchronize(h->stream));
    v.epoch = c->epoch;
    v.err = c->err;
    c->p2p_mask = mask;
    c->p2p_ok = true;
}

void tpb_allreduce_max(tpb_handle_s* h, double* dev_buf, int count) {
    if (!h->comm || h->comm->nranks == 1) return;
    TPB_NCCL(api().AllReduce(dev_buf, dev_buf, (size_t)count, ncclFloat64, ncclMax, h->comm->comm, h->stream));
}

// owned planes (along the slab axis) of every rank; one synchronising all-reduce on first use
const std::vector<int>& tpb_comm_planes(tpb_handle_s* h) {
    TPB_REQUIRE(h->comm != nullptr, TPB_ERR_STATE, "no communicator");
    CommState* c = h->comm;
    if ((int)c->planes.size() == c->nranks) return c->planes;
    std::vector<double> v(c->nranks, 0.0);
    v[c->rank] = (double)h->g.nl;
    if (c->nranks > 1) {
        if (!c->scratch) c->scratch = tpb_dalloc<double>(c->nranks);
        TPB_CUDA(cudaMemcpyAsync(c->scratch, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        tpb_allreduce_sum(h, c->scratch, c->nranks);
        TPB_CUDA(cudaMemcpyAsync(v.data(), c->scratch, v.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        TPB_CUDA(cudaStreamSynchronize(h->stream));
    }
    c->planes.resize(c->nranks);
    for (int r = 0; r < c->nranks; r++) c->planes[r] = (int)(v[r] + 0.5);
    return c->planes;
}

// in-place all-gather of pieces of different lengths: rank r's piece lives at buf + off[r] (cnt[r] doubles) on
// every rank, `nrep` such buffers `stride` doubles apart, all in one NCCL group
void tpb_allgatherv(tpb_handle_s* h, double* buf, const long long* off, const long long* cnt, int nrep, long long stride) {
    if (!h->comm || h->comm->nranks == 1) return;
    CommState* c = h->comm;
    NcclApi& a = api();
    bool equal = true;
    for (int r = 1; r < c->nranks; r++) equal = equal && cnt[r] == cnt[0] && off[r] == off[0] + (long long)r * cnt[0];
    TPB_NCCL(a.GroupStart());
    for (int q = 0; q < nrep; q++) {
        double* b = buf + (long long)q * stride;
        if (equal) {
            TPB_NCCL(a.AllGather(b + off[c->rank], b + off[0], (size_t)cnt[0], ncclFloat64, c->comm, h->stream));
        } else {
            for (int r = 0; r < c->nranks; r++)
                TPB_NCCL(a.Broadcast(b + off[r], b + off[r], (size_t)cnt[r], ncclFloat64, r, c->comm, h->stream));
        }
    }
    TPB_NCCL(a.GroupEnd());
}

int tpb_comm_peer_mode_impl(tpb_handle_s* h) { return (h->comm && h->comm->p2p_ok) ? h->comm->p2p_mask : 0; }
int tpb_comm_rank(tpb_handle_s* h) { return h->comm ? h->comm->rank : 0; }
int tpb_comm_size(tpb_handle_s* h) { return h->comm ? h->comm->nranks : 1; }

void tpb_comm_init_impl(tpb_handle_s* h, const void* id128, int rank, int nranks) {
    TPB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, TPB_ERR_ARG, "bad rank/nranks");
    tpb_comm_free(h);
    h->comm = new CommState();
    h->comm->rank = rank;
    h->comm->nranks = nranks;
    if (nranks == 1) return;
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    TPB_NCCL(api().CommInitRank(&h->comm->comm, nranks, id, rank));
    p2p_init(h);
}

void tpb_comm_unique_id_impl(void* out128) {
    ncclUniqueId id;
    TPB_NCCL(api().GetUniqueId(&id));
    memcpy(out128, id.internal, 128);
}
