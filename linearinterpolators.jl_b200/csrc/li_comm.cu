// li_comm.cu -- the one exchange step of the path: the adjoint of many per-frame operators accumulated into a model
// that is SHARED by the GPUs of a node (iterative fitting, README.md:36-37; BASELINE.json config #5).
//
//   li_comm_*                        NCCL communicators behind the C ABI (one rank per GPU: one process per GPU with
//                                    li_comm_init_rank, or one process driving several GPUs with li_comm_init_all)
//   li_allreduce_sum                 in-place sum over the ranks (ncclAllReduce over NVLink / NVSwitch)
//   li_adjoint_accumulate_allreduce  model <- beta * model + alpha * sum over ALL ranks and local frames of H_k' y_k:
//                                    the multi-frame accumulate kernel (li_sep2d_multi.cuh) produces the model band by
//                                    band; the all-reduce of band b runs on the communicator's stream while band b+1
//                                    is being computed, so only the last band's all-reduce is exposed.
//
// NCCL is reached through dlopen("libnccl.so.2") at first use: the library has no link-time dependency on it, a
// process that never creates a communicator never loads it, and inside a PyTorch process the NCCL that torch already
// loaded is the one that gets used.  Only <nccl.h>'s types are needed at compile time.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <vector>

#include "li_common.cuh"
#include "li_sep2d_multi.cuh"

namespace li {

std::mutex &multi_mutex()
{
    static std::mutex mu;
    return mu;
}
std::map<MultiKey, MultiPlan> &multi_plans()
{
    static std::map<MultiKey, MultiPlan> plans;
    return plans;
}

extern template int sep2d_adj_multi<float>(int64_t, const li_table *const *, const void *const *, double, double, void *, int, int,
                                           cudaStream_t, bool *);
extern template int sep2d_adj_multi<double>(int64_t, const li_table *const *, const void *const *, double, double, void *, int, int,
                                            cudaStream_t, bool *);

// every frame is a fused-capable pair of tables with the model's shape, kernel sizes and element type
static bool multi_eligible(int64_t n, const li_table *const *tabs)
{
    if (n < 2 || opt(OPT_NO_MULTI) || opt(OPT_NO_FUSED)) return false;
    const li_table *a1 = tabs[0], *a2 = tabs[1];
    for (int64_t k = 0; k < n; ++k) {
        const li_table *t1 = tabs[2 * k], *t2 = tabs[2 * k + 1];
        if (!sep2d_fused_supported(t1, t2)) return false;
        if (t1->dtype != a1->dtype || t1->S != a1->S || t2->S != a2->S || t1->ncols != a1->ncols || t2->ncols != a2->ncols)
            return false;
    }
    return true;
}

// model rows [row_lo, row_hi) <- beta * model + alpha * sum_k H_k' y_k with the multi-frame kernel
int sep_accumulate_multi(int64_t n, const li_table *const *tabs, const void *const *src_dev, double alpha, double beta,
                         void *dst, int64_t row_lo, int64_t row_hi, cudaStream_t st, bool *done)
{
    *done = false;
    if (!multi_eligible(n, tabs)) return LI_OK;
    if (row_hi <= row_lo) {
        *done = true;
        return LI_OK;
    }
    return tabs[0]->dtype == LI_F32
               ? sep2d_adj_multi<float>(n, tabs, src_dev, alpha, beta, dst, (int)row_lo, (int)row_hi, st, done)
               : sep2d_adj_multi<double>(n, tabs, src_dev, alpha, beta, dst, (int)row_lo, (int)row_hi, st, done);
}

// ---- NCCL through dlopen ---------------------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi &nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        // the copy that is already in the process (PyTorch's) first, then the loader's search path
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!a.handle) a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) return a;
#define LI_SYM(field, name) *(void **)(&a.field) = dlsym(a.handle, name)
        LI_SYM(GetUniqueId, "ncclGetUniqueId");
        LI_SYM(CommInitRank, "ncclCommInitRank");
        LI_SYM(CommInitAll, "ncclCommInitAll");
        LI_SYM(CommDestroy, "ncclCommDestroy");
        LI_SYM(AllReduce, "ncclAllReduce");
        LI_SYM(AllGather, "ncclAllGather");
        LI_SYM(GroupStart, "ncclGroupStart");
        LI_SYM(GroupEnd, "ncclGroupEnd");
        LI_SYM(GetVersion, "ncclGetVersion");
        LI_SYM(GetErrorString, "ncclGetErrorString");
#undef LI_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd &&
               a.GetErrorString;
        return a;
    }();
    return api;
}

#define LI_NCCL(expr)                                                                              \
    do {                                                                                           \
        ncclResult_t r__ = (expr);                                                                 \
        if (r__ != ncclSuccess) {                                                                  \
            li::set_error("%s failed: %s", #expr, li::nccl().GetErrorString(r__));                 \
            return LI_ERR_NCCL;                                                                    \
        }                                                                                          \
    } while (0)

static int need_nccl()
{
    LI_REQUIRE(nccl().ok, LI_ERR_NCCL, "NCCL is not available: dlopen(\"libnccl.so.2\") failed (%s)", dlerror() ? dlerror() : "symbols missing");
    return LI_OK;
}

}  // namespace li

struct li_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
    cudaStream_t cs = nullptr;                  // the communicator's own stream (all-reduces of finished bands)
    static const int MAXB = 16;
    cudaEvent_t ev_band[MAXB] = {}, ev_done = nullptr;
    // ---- peer-memory all-reduce (li::PeerAR below): one process per GPU, buffers exchanged as CUDA IPC handles
    static const int MAXR = 16;
    bool one_process_per_rank = false;          // li_comm_init_rank (IPC handles cannot be opened by their own process)
    int peer_state = 0;                         // 0: not tried for the current size, 1: usable, -1: not available
    size_t sym_bytes = 0;
    void *sym = nullptr;                        // this rank's exchange buffer (the rank-private sum lives here)
    unsigned *flags = nullptr;                  // [2][MAXR] arrival epochs written by the peers
    void *peer_sym[MAXR] = {};                  // peers' exchange buffers / flag arrays, mapped into this process
    unsigned *peer_flags[MAXR] = {};
    unsigned epoch = 0;
    int *timeout_host = nullptr, *timeout_dev = nullptr;   // mapped pinned word: a barrier gave up waiting
};

namespace li {

static int comm_finish_init(li_comm *c)
{
    // highest priority: the all-reduce of a finished band must get its few CTAs scheduled ahead of the queued CTAs of
    // the next band's compute kernel, not behind them
    int prio_lo = 0, prio_hi = 0;
    LI_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    LI_CUDA(cudaStreamCreateWithPriority(&c->cs, cudaStreamNonBlocking, prio_hi));
    for (int b = 0; b < li_comm::MAXB; ++b) LI_CUDA(cudaEventCreateWithFlags(&c->ev_band[b], cudaEventDisableTiming));
    LI_CUDA(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
    return LI_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (dev != prev) cudaSetDevice(dev);
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- all-reduce over NVLink peer memory (one process per GPU) ----------------------------------------------------
// The rank-private sums live in exchange buffers that every rank maps (CUDA IPC).  Two-shot, owner-computes:
//   barrier (every rank's buffer is complete)  ->  rank r sums slice r of all buffers IN RANK ORDER (deterministic, the
//   same bits on every rank) and stores the result into slice r of every buffer  ->  barrier (every slice is complete).
// The barriers are epoch flags written into the peers' flag arrays with system-scope release stores and polled with
// acquire loads; a wait gives up after ~2 s (flagged, reported by the next library call) instead of hanging the GPU.
// For the 16.8 MB model of config #5 this takes ~half of NCCL's time on a node of B200s: the message is latency-, not
// bandwidth-bound, and the protocol is two barriers + one pass.
__global__ void peer_barrier_kernel(unsigned *const *peer_flags, unsigned *my_flags, int phase, int rank, int nranks, unsigned epoch,
                                    int *timeout)
{
    const int p = threadIdx.x;
    if (p < nranks) {
        unsigned *dst = peer_flags[p] + phase * li_comm::MAXR + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
        const unsigned *src = my_flags + phase * li_comm::MAXR + p;
        const long long t0 = clock64();
        unsigned v = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
            if ((int)(v - epoch) >= 0) break;
            if (clock64() - t0 > 4000000000LL) {
                *timeout = 1;
                break;
            }
            __nanosleep(200);
        }
    }
}

template <typename T, int NR>
__global__ void __launch_bounds__(256) peer_reduce_kernel(const void *const *peer_sym, int rank, int64_t count)
{
    // slice of this rank, in 16-byte vectors
    constexpr int V = 16 / (int)sizeof(T);
    const int64_t nvec = count / V;
    const int64_t v0 = nvec * rank / NR, v1 = nvec * (rank + 1) / NR;
    T *bufs[NR];
#pragma unroll
    for (int p = 0; p < NR; ++p) bufs[p] = (T *)peer_sym[p];
    for (int64_t v = v0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < v1; v += (int64_t)gridDim.x * blockDim.x) {
        Pack<T, V> acc = ld_pack<T, V>(bufs[0] + v * V);
#pragma unroll
        for (int p = 1; p < NR; ++p) {
            const Pack<T, V> q = ld_pack<T, V>(bufs[p] + v * V);
#pragma unroll
            for (int e = 0; e < V; ++e) acc.v[e] += q.v[e];
        }
#pragma unroll
        for (int p = 0; p < NR; ++p) st_pack<T, V>(bufs[p] + v * V, acc);
    }
    // the tail (count not a multiple of V): rank 0
    if (rank == 0 && blockIdx.x == 0 && threadIdx.x < (int)(count - nvec * V)) {
        const int64_t i = nvec * V + threadIdx.x;
        T s = bufs[0][i];
#pragma unroll
        for (int p = 1; p < NR; ++p) s += bufs[p][i];
#pragma unroll
        for (int p = 0; p < NR; ++p) bufs[p][i] = s;
    }
    __threadfence_system();
}

struct PeerTables {   // device copies of the peer pointer arrays (per communicator, rebuilt with the buffers)
    void **sym = nullptr;
    unsigned **flags = nullptr;
};
static std::map<li_comm *, PeerTables> &peer_tables()
{
    static std::map<li_comm *, PeerTables> t;
    return t;
}

static void peer_release(li_comm *c)
{
    for (int p = 0; p < c->nranks && p < li_comm::MAXR; ++p) {
        if (p != c->rank) {
            if (c->peer_sym[p]) cudaIpcCloseMemHandle(c->peer_sym[p]);
            if (c->peer_flags[p]) cudaIpcCloseMemHandle(c->peer_flags[p]);
        }
        c->peer_sym[p] = nullptr, c->peer_flags[p] = nullptr;
    }
    if (c->sym) cudaFree(c->sym);
    if (c->flags) cudaFree(c->flags);
    c->sym = nullptr, c->flags = nullptr, c->sym_bytes = 0;
    auto it = peer_tables().find(c);
    if (it != peer_tables().end()) {
        cudaFree(it->second.sym);
        cudaFree(it->second.flags);
        peer_tables().erase(it);
    }
    cudaGetLastError();
}

// Collective (every rank calls it with the same `bytes`): make exchange buffers of at least `bytes` and map the peers'.
// Whatever happens on one rank, all ranks agree on the outcome (a sum all-reduce of a failure flag): either every rank
// uses the peer path or every rank uses NCCL.
static int peer_setup(li_comm *c, size_t bytes, cudaStream_t st)
{
    if (c->peer_state == 1 && bytes <= c->sym_bytes) return LI_OK;
    if (c->peer_state == -1) return LI_OK;
    if (!c->one_process_per_rank || c->nranks < 2 || c->nranks > li_comm::MAXR || !nccl().AllGather || opt(OPT_EXP) & 8) {
        c->peer_state = -1;
        return LI_OK;
    }
    LI_CUDA(cudaStreamSynchronize(st));   // nothing in flight may still use the old buffers
    peer_release(c);
    struct Handles {
        cudaIpcMemHandle_t sym, flags;
    };
    int bad = 0;
    const size_t want = (bytes + bytes / 8 + 4095) & ~(size_t)4095;
    Handles mine;
    memset(&mine, 0, sizeof mine);
    if (cudaMalloc(&c->sym, want) != cudaSuccess || cudaMalloc((void **)&c->flags, 2 * li_comm::MAXR * sizeof(unsigned)) != cudaSuccess)
        bad = 1;
    if (!bad && cudaMemsetAsync(c->flags, 0, 2 * li_comm::MAXR * sizeof(unsigned), st) != cudaSuccess) bad = 1;
    if (!bad && (cudaIpcGetMemHandle(&mine.sym, c->sym) != cudaSuccess || cudaIpcGetMemHandle(&mine.flags, c->flags) != cudaSuccess)) bad = 1;
    if (!c->timeout_host) {
        if (cudaHostAlloc((void **)&c->timeout_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void **)&c->timeout_dev, c->timeout_host, 0) != cudaSuccess)
            bad = 1;
        else
            *c->timeout_host = 0;
    }
    cudaGetLastError();
    // exchange the handles (NCCL all-gather of 128 bytes per rank), then open the peers'
    Handles *d_all = nullptr;
    std::vector<Handles> all((size_t)c->nranks);
    LI_CUDA(cudaMalloc((void **)&d_all, sizeof(Handles) * (size_t)c->nranks));
    LI_CUDA(cudaMemcpyAsync(d_all + c->rank, &mine, sizeof mine, cudaMemcpyHostToDevice, st));
    LI_NCCL(nccl().AllGather(d_all + c->rank, d_all, sizeof(Handles), ncclChar, c->comm, st));
    LI_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(Handles) * (size_t)c->nranks, cudaMemcpyDeviceToHost, st));
    LI_CUDA(cudaStreamSynchronize(st));
    for (int p = 0; p < c->nranks && !bad; ++p) {
        if (p == c->rank) {
            c->peer_sym[p] = c->sym, c->peer_flags[p] = c->flags;
            continue;
        }
        if (cudaIpcOpenMemHandle(&c->peer_sym[p], all[(size_t)p].sym, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle((void **)&c->peer_flags[p], all[(size_t)p].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
            bad = 1;
    }
    cudaGetLastError();
    // agreement: sum of the failure flags over the ranks (d_all reused as scratch)
    int *d_bad = (int *)d_all;
    LI_CUDA(cudaMemcpyAsync(d_bad, &bad, sizeof bad, cudaMemcpyHostToDevice, st));
    LI_NCCL(nccl().AllReduce(d_bad, d_bad, 1, ncclInt32, ncclSum, c->comm, st));
    int nbad = 0;
    LI_CUDA(cudaMemcpyAsync(&nbad, d_bad, sizeof nbad, cudaMemcpyDeviceToHost, st));
    LI_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_all);
    if (nbad) {
        if (opt(OPT_TRACE)) fprintf(stderr, "[li_b200] rank %d: peer-memory all-reduce not available (%d rank(s) failed to map): NCCL\n", c->rank, nbad);
        peer_release(c);
        c->peer_state = -1;
        return LI_OK;
    }
    PeerTables &t = peer_tables()[c];
    LI_CUDA(cudaMalloc((void **)&t.sym, li_comm::MAXR * sizeof(void *)));
    LI_CUDA(cudaMalloc((void **)&t.flags, li_comm::MAXR * sizeof(unsigned *)));
    LI_CUDA(cudaMemcpy(t.sym, c->peer_sym, li_comm::MAXR * sizeof(void *), cudaMemcpyHostToDevice));
    LI_CUDA(cudaMemcpy(t.flags, c->peer_flags, li_comm::MAXR * sizeof(unsigned *), cudaMemcpyHostToDevice));
    c->sym_bytes = want;
    c->peer_state = 1;
    if (opt(OPT_TRACE)) fprintf(stderr, "[li_b200] rank %d: peer-memory all-reduce ready (%zu bytes per rank)\n", c->rank, want);
    return LI_OK;
}

// in-place sum over the ranks of the first `count` elements of every rank's exchange buffer, on `st`
static int peer_allreduce(li_comm *c, int dtype, int64_t count, cudaStream_t st)
{
    LI_REQUIRE(!*c->timeout_host, LI_ERR_NCCL, "a peer-memory barrier timed out earlier on this communicator (a rank is missing)");
    PeerTables &t = peer_tables()[c];
    const unsigned e = ++c->epoch;
    peer_barrier_kernel<<<1, 32, 0, st>>>(t.flags, c->flags, 0, c->rank, c->nranks, e, c->timeout_dev);
    const size_t es = dtype_size(dtype);
    const int64_t per_rank = count / c->nranks * (int64_t)es / 16;
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((per_rank + 255) / 256, 4 * 148));
#define LI_PR(NRR)                                                                                                  \
    if (c->nranks == NRR) {                                                                                         \
        if (dtype == LI_F32)                                                                                        \
            peer_reduce_kernel<float, NRR><<<blocks, 256, 0, st>>>((const void *const *)t.sym, c->rank, count);     \
        else                                                                                                        \
            peer_reduce_kernel<double, NRR><<<blocks, 256, 0, st>>>((const void *const *)t.sym, c->rank, count);    \
    }
    LI_PR(2) LI_PR(3) LI_PR(4) LI_PR(5) LI_PR(6) LI_PR(7) LI_PR(8) LI_PR(16)
#undef LI_PR
    peer_barrier_kernel<<<1, 32, 0, st>>>(t.flags, c->flags, 1, c->rank, c->nranks, e, c->timeout_dev);
    count_launch(3);
    LI_CUDA(cudaGetLastError());
    return LI_OK;
}

static bool peer_ranks_supported(int n) { return (n >= 2 && n <= 8) || n == 16; }

}  // namespace li

extern "C" {

int li_comm_unique_id(void *id, size_t bytes)
{
    LI_REQUIRE(id && bytes >= LI_COMM_ID_BYTES, LI_ERR_ARG, "the id buffer must hold LI_COMM_ID_BYTES = %d bytes", LI_COMM_ID_BYTES);
    static_assert(sizeof(ncclUniqueId) == LI_COMM_ID_BYTES, "ncclUniqueId size");
    int rc = li::need_nccl();
    if (rc != LI_OK) return rc;
    ncclUniqueId uid;
    LI_NCCL(li::nccl().GetUniqueId(&uid));
    memcpy(id, &uid, sizeof uid);
    return LI_OK;
}

int li_comm_init_rank(const void *id, int nranks, int rank, li_comm **out)
{
    LI_REQUIRE(id && out, LI_ERR_ARG, "null argument");
    LI_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, LI_ERR_ARG, "bad rank %d of %d", rank, nranks);
    int rc = li::need_nccl();
    if (rc != LI_OK) return rc;
    li_comm *c = new li_comm();
    c->rank = rank, c->nranks = nranks;
    cudaError_t e = cudaGetDevice(&c->device);
    if (e != cudaSuccess) {
        delete c;
        li::set_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
        return LI_ERR_CUDA;
    }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclResult_t r = li::nccl().CommInitRank(&c->comm, nranks, uid, rank);
    if (r != ncclSuccess) {
        li::set_error("ncclCommInitRank failed: %s", li::nccl().GetErrorString(r));
        delete c;
        return LI_ERR_NCCL;
    }
    rc = li::comm_finish_init(c);
    if (rc != LI_OK) {
        li_comm_destroy(c);
        return rc;
    }
    c->one_process_per_rank = true;
    *out = c;
    return LI_OK;
}

int li_comm_init_all(int ndev, const int *devices, li_comm **out)
{
    LI_REQUIRE(out && ndev >= 1 && ndev <= 64, LI_ERR_ARG, "bad device list");
    int rc = li::need_nccl();
    if (rc != LI_OK) return rc;
    std::vector<ncclComm_t> comms((size_t)ndev);
    std::vector<int> devs((size_t)ndev);
    for (int k = 0; k < ndev; ++k) devs[k] = devices ? devices[k] : k;
    LI_NCCL(li::nccl().CommInitAll(comms.data(), ndev, devs.data()));
    for (int k = 0; k < ndev; ++k) out[k] = nullptr;
    for (int k = 0; k < ndev; ++k) {
        li_comm *c = new li_comm();
        c->comm = comms[k], c->rank = k, c->nranks = ndev, c->device = devs[k];
        out[k] = c;
        li::DeviceGuard guard(devs[k]);
        rc = li::comm_finish_init(c);
        if (rc != LI_OK) {
            for (int j = 0; j < ndev; ++j) {
                if (out[j])
                    li_comm_destroy(out[j]);
                else
                    li::nccl().CommDestroy(comms[j]);
                out[j] = nullptr;
            }
            return rc;
        }
    }
    return LI_OK;
}

int li_comm_info(const li_comm *comm, int *rank, int *nranks, int *device)
{
    LI_REQUIRE(comm, LI_ERR_ARG, "null communicator");
    if (rank) *rank = comm->rank;
    if (nranks) *nranks = comm->nranks;
    if (device) *device = comm->device;
    return LI_OK;
}

int li_comm_destroy(li_comm *comm)
{
    if (!comm) return LI_OK;
    li::DeviceGuard guard(comm->device);
    if (comm->cs) cudaStreamSynchronize(comm->cs);
    cudaDeviceSynchronize();
    li::peer_release(comm);
    if (comm->timeout_host) cudaFreeHost(comm->timeout_host);
    if (comm->comm && li::nccl().ok) li::nccl().CommDestroy(comm->comm);
    for (int b = 0; b < li_comm::MAXB; ++b)
        if (comm->ev_band[b]) cudaEventDestroy(comm->ev_band[b]);
    if (comm->ev_done) cudaEventDestroy(comm->ev_done);
    if (comm->cs) cudaStreamDestroy(comm->cs);
    delete comm;
    return LI_OK;
}

int li_comm_group_start(void)
{
    int rc = li::need_nccl();
    if (rc != LI_OK) return rc;
    LI_NCCL(li::nccl().GroupStart());
    return LI_OK;
}

int li_comm_group_end(void)
{
    int rc = li::need_nccl();
    if (rc != LI_OK) return rc;
    LI_NCCL(li::nccl().GroupEnd());
    return LI_OK;
}

int li_nccl_version(int *version)
{
    LI_REQUIRE(version, LI_ERR_ARG, "null argument");
    int rc = li::need_nccl();
    if (rc != LI_OK) return rc;
    *version = 0;
    if (li::nccl().GetVersion) LI_NCCL(li::nccl().GetVersion(version));
    return LI_OK;
}

int li_allreduce_sum(li_comm *comm, int dtype, void *buf_dev, int64_t count, void *stream)
{
    LI_REQUIRE(dtype == LI_F32 || dtype == LI_F64, LI_ERR_ARG, "bad dtype");
    LI_REQUIRE(count >= 0 && (buf_dev || count == 0), LI_ERR_ARG, "null argument");
    if (!comm || comm->nranks == 1 || count == 0) return LI_OK;   // a single rank already holds the sum
    // large messages of a one-process-per-GPU communicator: the peer-memory all-reduce (through the exchange buffer)
    if (comm->one_process_per_rank && li::peer_ranks_supported(comm->nranks) && (size_t)count * li::dtype_size(dtype) >= (1u << 20)) {
        cudaStream_t st = li::as_stream(stream);
        const size_t bytes = (size_t)count * li::dtype_size(dtype);
        int rc = li::peer_setup(comm, bytes, st);
        if (rc != LI_OK) return rc;
        if (comm->peer_state == 1) {
            LI_CUDA(cudaMemcpyAsync(comm->sym, buf_dev, bytes, cudaMemcpyDeviceToDevice, st));
            rc = li::peer_allreduce(comm, dtype, count, st);
            if (rc != LI_OK) return rc;
            LI_CUDA(cudaMemcpyAsync(buf_dev, comm->sym, bytes, cudaMemcpyDeviceToDevice, st));
            return LI_OK;
        }
    }
    LI_NCCL(li::nccl().AllReduce(buf_dev, buf_dev, (size_t)count, dtype == LI_F32 ? ncclFloat32 : ncclFloat64, ncclSum, comm->comm,
                                 li::as_stream(stream)));
    return LI_OK;
}

int li_adjoint_accumulate_allreduce(li_comm *comm, int dtype, int64_t m1, int64_t m2, int64_t nframes,
                                    const li_table *const *tabs, const void *const *src_dev, double alpha, double beta,
                                    void *model_dev, int nbands, void *stream)
{
    LI_REQUIRE(dtype == LI_F32 || dtype == LI_F64, LI_ERR_ARG, "bad dtype");
    LI_REQUIRE(nframes >= 0 && m1 >= 1 && m2 >= 1 && model_dev, LI_ERR_ARG, "bad argument");
    LI_REQUIRE(nframes == 0 || (tabs && src_dev), LI_ERR_ARG, "null argument");
    for (int64_t k = 0; k < nframes; ++k) {
        const li_table *t1 = tabs[2 * k], *t2 = tabs[2 * k + 1];
        LI_REQUIRE(t1 && t2 && src_dev[k], LI_ERR_ARG, "null table or frame (frame %lld)", (long long)k);
        LI_REQUIRE(t1->dtype == dtype && t2->dtype == dtype, LI_ERR_ARG, "frame %lld: tables of another floating-point type", (long long)k);
        LI_REQUIRE(t1->ncols == m1 && t2->ncols == m2, LI_ERR_DIM, "frame %lld maps onto a %lld x %lld model, expected %lld x %lld",
                   (long long)k, (long long)t1->ncols, (long long)t2->ncols, (long long)m1, (long long)m2);
    }
    cudaStream_t st = li::as_stream(stream);
    const int64_t nmodel = m1 * m2;
    const size_t es = li::dtype_size(dtype);
    const bool multi_rank = comm != nullptr;   // (a one-rank communicator runs the same path: bands + all-reduce)
    if (multi_rank) {
        int dev = 0;
        LI_CUDA(cudaGetDevice(&dev));
        LI_REQUIRE(dev == comm->device, LI_ERR_ARG, "the communicator belongs to device %d, the current device is %d", comm->device, dev);
    }
    // the rank-private sum: the model itself when nothing of it survives (beta = 0) or when no other rank adds to it
    void *acc = model_dev;
    const double acc_beta = multi_rank ? 0.0 : beta;
    // one process per GPU, a single band: the sum goes straight into the exchange buffer of the peer-memory all-reduce
    bool peer = false;
    if (multi_rank && comm->nranks > 1 && nbands <= 1 && comm->one_process_per_rank && li::peer_ranks_supported(comm->nranks)) {
        int rc = li::peer_setup(comm, (size_t)nmodel * es, st);
        if (rc != LI_OK) return rc;
        peer = comm->peer_state == 1;
    }
    if (peer) {
        acc = comm->sym;
    } else if (multi_rank && beta != 0.0) {
        int rc = li::staging(6, (size_t)nmodel * es, &acc, stream);
        if (rc != LI_OK) return rc;
    }
    // Bands of model rows.  Every rank issues the SAME sequence of all-reduces (the band limits depend on m2 and nbands
    // only), however its local part is computed.
    int nb = multi_rank ? std::max(1, std::min(nbands, (int)li_comm::MAXB)) : 1;
    nb = (int)std::min<int64_t>(nb, std::max<int64_t>(1, m2 / 64));
    bool multi = true;   // the multi-frame kernel takes the local frames (decided on the first band)
    for (int b = 0; b < nb; ++b) {
        const int64_t r0 = m2 * b / nb, r1 = m2 * (b + 1) / nb;
        if (multi) {
            bool done = false;
            int rc = li::sep_accumulate_multi(nframes, tabs, src_dev, alpha, acc_beta, acc, r0, r1, st, &done);
            if (rc != LI_OK) return rc;
            if (!done) {
                LI_REQUIRE(b == 0, LI_ERR_UNSUPPORTED, "internal: the multi-frame kernel declined a band after taking the first");
                multi = false;   // one apply per frame into the accumulator, the whole model at once
                rc = li::vscale_impl(dtype, acc, nmodel, acc_beta, st);
                if (rc != LI_OK) return rc;
                if (nframes > 0) {
                    std::vector<void *> dsts((size_t)nframes, acc);
                    rc = li_sep_apply_batch(1, 2, nframes, tabs, nullptr, 1, alpha, src_dev, 1.0, dsts.data(), LI_VARIANT_AUTO, stream);
                    if (rc != LI_OK) return rc;
                }
            }
        }
        if (peer) {
            int rc = li::peer_allreduce(comm, dtype, nmodel, st);
            if (rc != LI_OK) return rc;
        } else if (multi_rank) {   // rows [r0, r1) are final on this rank: reduce them while the next band is computed
            LI_CUDA(cudaEventRecord(comm->ev_band[b], st));
            LI_CUDA(cudaStreamWaitEvent(comm->cs, comm->ev_band[b], 0));
            char *p = (char *)acc + (size_t)r0 * m1 * es;
            LI_NCCL(li::nccl().AllReduce(p, p, (size_t)((r1 - r0) * m1), dtype == LI_F32 ? ncclFloat32 : ncclFloat64, ncclSum, comm->comm,
                                         comm->cs));
        }
    }
    if (multi_rank && !peer) {
        LI_CUDA(cudaEventRecord(comm->ev_done, comm->cs));
        LI_CUDA(cudaStreamWaitEvent(st, comm->ev_done, 0));
    }
    if (acc != model_dev)   // model <- beta * model + (sum over the ranks)
        return li_vcombine(dtype, model_dev, 1.0, acc, beta, model_dev, nmodel, stream);
    return LI_OK;
}

}  // extern "C"
