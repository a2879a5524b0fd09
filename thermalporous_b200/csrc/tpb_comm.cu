// tpb_comm.cu - slab-partition plumbing: ghost-plane exchange and scalar all-reduces over NCCL.
//
// Stands in for the MPI traffic PETSc/PyOP2 generate under the reference (halo update before
// every assembly, MatMult ghost scatter, MPI_Allreduce per Krylov dot).  The path
// has exactly two exchange patterns: one boundary plane per neighbour (ncclSend/ncclRecv grouped)
// and a sum of k doubles.  NCCL is resolved at run time with dlopen("libnccl.so.2") - the copy
// torch already mapped into the process - so the single-GPU library has no NCCL link dependency.
#include <dlfcn.h>

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

}  // namespace

struct CommState {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    double* send_first = nullptr;  // packed boundary planes (nf_max * np)
    double* send_last = nullptr;
    size_t cap = 0;
    std::vector<int> planes;       // owned planes of every rank along the slab axis (filled on first use)
    double* scratch = nullptr;     // nranks doubles
};

void tpb_comm_free(tpb_handle_s* h) {
    if (!h->comm) return;
    if (h->comm->comm) api().CommDestroy(h->comm->comm);
    tpb_dfree(h->comm->send_first);
    tpb_dfree(h->comm->send_last);
    tpb_dfree(h->comm->scratch);
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
}

void tpb_comm_unique_id_impl(void* out128) {
    ncclUniqueId id;
    TPB_NCCL(api().GetUniqueId(&id));
    memcpy(out128, id.internal, 128);
}
