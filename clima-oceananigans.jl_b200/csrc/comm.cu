// comm.cu -- NCCL plumbing for the slab-decomposed path (reference: src/Distributed, which uses
// MPI.Isend/Irecv for halos and PencilFFTs all-to-all transposes; here NCCL over NVLink/NVSwitch).
// libnccl.so.2 is bound lazily with dlopen so that single-GPU use has no NCCL dependency; in a
// process that already loaded torch's bundled NCCL the same library instance is reused.
#include "internal.h"
#include <dlfcn.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace ob {
namespace comm {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct Api {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
};
static Api api;
static bool loaded = false;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_size = 1;

static void load() {
    if (loaded) return;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw Error(std::string("cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* n) {
        void* p = dlsym(h, n);
        if (!p) throw Error(std::string("libnccl.so.2 lacks ") + n);
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    loaded = true;
}
static void ck(ncclResult_t r, const char* what) {
    if (r != 0) throw Error(std::string(what) + ": " + api.GetErrorString(r));
}

void unique_id(char out[128]) {
    load();
    ncclUniqueId id;
    ck(api.GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out, id.internal, 128);
}
void init(int nranks, int rank, const char id_bytes[128]) {
    load();
    if (g_comm) throw Error("communicator already initialised");
    ncclUniqueId id;
    memcpy(id.internal, id_bytes, 128);
    ck(api.CommInitRank(&g_comm, nranks, id, rank), "ncclCommInitRank");
    g_rank = rank; g_size = nranks;
}
static void link_release();
static void peer_barrier_release();
void barrier();
void destroy() {
    if (g_comm) { cudaStreamSynchronize(stream()); barrier(); cudaStreamSynchronize(stream()); peer_barrier_release(); }
    if (g_comm) link_release();
    if (g_comm) { api.CommDestroy(g_comm); g_comm = nullptr; g_size = 1; g_rank = 0; }
}
int rank() { return g_rank; }
int size() { return g_size; }
bool active() { return g_comm != nullptr; }

void group_start() { ck(api.GroupStart(), "ncclGroupStart"); }
void group_end() { ck(api.GroupEnd(), "ncclGroupEnd"); count_launch(1); }
void send(const void* buf, size_t bytes, int peer) { ck(api.Send(buf, bytes, ncclInt8, peer, g_comm, stream()), "ncclSend"); }
void recv(void* buf, size_t bytes, int peer) { ck(api.Recv(buf, bytes, ncclInt8, peer, g_comm, stream()), "ncclRecv"); }
void allreduce_f64(double* buf, size_t n, bool max) {
    ck(api.AllReduce(buf, buf, n, ncclFloat64, max ? ncclMax : ncclSum, g_comm, stream()), "ncclAllReduce");
    count_launch(1);
}

void allgather_bytes(const void* send_dev, void* recv_dev, size_t bytes_per_rank) {
    ck(api.AllGather(send_dev, recv_dev, bytes_per_rank, ncclInt8, g_comm, stream()), "ncclAllGather");
    count_launch(1);
}
// stream-ordered barrier over all ranks: every rank's preceding work on the library stream has completed
// (and its peer-memory stores are visible) before any rank's following work starts
void barrier() {
    static double* tok = nullptr;
    if (!tok) { OB_CUDA(cudaMalloc(&tok, 8)); OB_CUDA(cudaMemset(tok, 0, 8)); }
    ck(api.AllReduce(tok, tok, 1, ncclFloat64, ncclSum, g_comm, stream()), "ncclAllReduce(barrier)");
    count_launch(1);
}

// ---------------------------------------------------------------------------------------------------------------
// Can the ranks reach each other's memory?  Collective, evaluated once: all ranks on one host (hash of the host name),
// and every rank can open a CUDA IPC handle of every other rank's probe allocation (fails on GPUs without peer access, on
// multi-node communicators and in containers with IPC disabled).  If any rank says no, ALL ranks use the NCCL paths
// (grouped send / recv halo exchange, all_to_all transposes) -- the decision is global so that the two sides of an
// exchange never disagree.  OB200_NO_P2P=1 forces the NCCL paths.
// ---------------------------------------------------------------------------------------------------------------
bool peer_access_ok() {
    static int state = -1;
    if (state >= 0) return state != 0;
    if (!active() || getenv("OB200_NO_P2P") != nullptr) { state = 0; return false; }
    struct Probe { unsigned long long host; int ok; int pad; cudaIpcMemHandle_t h; };
    Probe mine{};
    char name[256] = {0};
    gethostname(name, sizeof(name) - 1);
    unsigned long long hsh = 1469598103934665603ULL;
    for (const char* q = name; *q; ++q) hsh = (hsh ^ (unsigned char)*q) * 1099511628211ULL;
    mine.host = hsh;
    void* probe = nullptr;
    mine.ok = cudaMalloc(&probe, 4096) == cudaSuccess && cudaIpcGetMemHandle(&mine.h, probe) == cudaSuccess;
    cudaGetLastError();
    Probe *dsend = nullptr, *drecv = nullptr;
    OB_CUDA(cudaMalloc(&dsend, sizeof(Probe)));
    OB_CUDA(cudaMalloc(&drecv, sizeof(Probe) * g_size));
    OB_CUDA(cudaMemcpyAsync(dsend, &mine, sizeof(Probe), cudaMemcpyHostToDevice, stream()));
    allgather_bytes(dsend, drecv, sizeof(Probe));
    std::vector<Probe> all(g_size);
    OB_CUDA(cudaMemcpyAsync(all.data(), drecv, sizeof(Probe) * g_size, cudaMemcpyDeviceToHost, stream()));
    OB_CUDA(cudaStreamSynchronize(stream()));
    int ok = mine.ok;
    for (int r = 0; r < g_size && ok; ++r) {
        if (r == g_rank) continue;
        if (!all[r].ok || all[r].host != mine.host) { ok = 0; break; }
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
        cudaIpcCloseMemHandle(q);
    }
    double bad = ok ? 0.0 : 1.0;
    double* dflag = reinterpret_cast<double*>(dsend);
    OB_CUDA(cudaMemcpyAsync(dflag, &bad, sizeof(double), cudaMemcpyHostToDevice, stream()));
    allreduce_f64(dflag, 1, true);              // max of the failure flags; also: nobody frees its probe before all have closed it
    OB_CUDA(cudaMemcpyAsync(&bad, dflag, sizeof(double), cudaMemcpyDeviceToHost, stream()));
    OB_CUDA(cudaStreamSynchronize(stream()));
    cudaFree(dsend); cudaFree(drecv);
    if (probe) cudaFree(probe);
    state = bad == 0.0 ? 1 : 0;
    if (!state && g_rank == 0 && getenv("OB200_QUIET") == nullptr)
        fprintf(stderr, "libocean_b200: no peer memory access between the ranks (other host, no P2P or CUDA IPC disabled): "
                        "using the NCCL send/recv and all-to-all paths\n");
    return state != 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Stream-ordered barrier over all ranks through peer memory: every rank owns R flag words that all ranks map through
// CUDA IPC; one R-thread kernel stores the epoch number into its own word of every rank (system-scope fence first: the
// peer-memory stores of the preceding kernels are visible before the flag) and then spins until all R of its own
// words show the epoch.  A few microseconds (one NVLink round trip) against 20-40 us for the NCCL all-reduce it
// replaces (two per distributed pressure solve, six per time step).  Falls back to the NCCL barrier without peer access.
// ---------------------------------------------------------------------------------------------------------------
struct PeerBarrier {
    unsigned long long* mine = nullptr;            // R words
    unsigned long long* peer[64] = {};             // the same block of every rank
    unsigned long long epoch = 0;
    int* err = nullptr;
    bool ready = false;
};
static PeerBarrier g_pb;
static unsigned long long link_timeout_ns();
__global__ void peer_barrier_kernel(unsigned long long* const* peers, const unsigned long long* mine, int rank, int R,
                                    unsigned long long epoch, int* err, unsigned long long timeout_ns) {
    __shared__ unsigned long long* sp[64];
    const int t = threadIdx.x;
    if (t < R) sp[t] = peers[t];
    __syncthreads();
    if (t >= R) return;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(sp[t] + rank) = epoch;
    const volatile unsigned long long* f = mine + t;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*f < epoch) {
        __nanosleep(32);
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) { *err = 2; __threadfence_system(); __trap(); }
    }
    __threadfence_system();
}
static unsigned long long** g_pb_dev_peers = nullptr;
static void peer_barrier_setup() {
    PeerBarrier& B = g_pb;
    if (B.ready) return;
    if (g_size > 64) throw Error("peer barrier: more than 64 ranks");
    OB_CUDA(cudaMalloc(&B.mine, 64 * sizeof(unsigned long long)));
    OB_CUDA(cudaMemsetAsync(B.mine, 0, 64 * sizeof(unsigned long long), stream()));
    OB_CUDA(cudaMalloc(&B.err, sizeof(int)));
    OB_CUDA(cudaMemsetAsync(B.err, 0, sizeof(int), stream()));
    cudaIpcMemHandle_t h;
    OB_CUDA(cudaIpcGetMemHandle(&h, B.mine));
    cudaIpcMemHandle_t *dsend, *drecv;
    OB_CUDA(cudaMalloc(&dsend, sizeof(h)));
    OB_CUDA(cudaMalloc(&drecv, sizeof(h) * g_size));
    OB_CUDA(cudaMemcpyAsync(dsend, &h, sizeof(h), cudaMemcpyHostToDevice, stream()));
    allgather_bytes(dsend, drecv, sizeof(h));
    std::vector<cudaIpcMemHandle_t> all(g_size);
    OB_CUDA(cudaMemcpyAsync(all.data(), drecv, sizeof(h) * g_size, cudaMemcpyDeviceToHost, stream()));
    OB_CUDA(cudaStreamSynchronize(stream()));
    cudaFree(dsend); cudaFree(drecv);
    for (int r = 0; r < g_size; ++r) {
        if (r == g_rank) { B.peer[r] = B.mine; continue; }
        void* q = nullptr;
        OB_CUDA(cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess));
        B.peer[r] = (unsigned long long*)q;
    }
    OB_CUDA(cudaMalloc(&g_pb_dev_peers, 64 * sizeof(unsigned long long*)));
    OB_CUDA(cudaMemcpyAsync(g_pb_dev_peers, B.peer, 64 * sizeof(unsigned long long*), cudaMemcpyHostToDevice, stream()));
    barrier();                 // every rank has zeroed its words before anybody publishes an epoch
    OB_CUDA(cudaStreamSynchronize(stream()));
    B.ready = true;
}
static void peer_barrier_release() {
    PeerBarrier& B = g_pb;
    if (!B.ready) return;
    for (int r = 0; r < g_size; ++r) if (r != g_rank && B.peer[r]) cudaIpcCloseMemHandle(B.peer[r]);
    cudaFree(B.mine); cudaFree(B.err); cudaFree(g_pb_dev_peers);
    B = PeerBarrier();
    g_pb_dev_peers = nullptr;
}
// barrier used inside the distributed pressure solve
void fast_barrier() {
    static const bool nccl_only = getenv("OB200_NCCL_BARRIER") != nullptr;
    if (nccl_only || !peer_access_ok()) { barrier(); return; }
    peer_barrier_setup();
    PeerBarrier& B = g_pb;
    B.epoch += 1;
    peer_barrier_kernel<<<1, 64, 0, stream()>>>(g_pb_dev_peers, B.mine, g_rank, g_size, B.epoch, B.err, link_timeout_ns());
    count_launch(1);
}

// ---------------------------------------------------------------------------------------------------------------
// Peer-memory halo link (ring neighbours of the slab decomposition).
// Reference: Distributed/halo_communication.jl:62-183 posts MPI.Isend / Irecv per field and side.  Here every rank
// owns receive buffers that its two ring neighbours map through CUDA IPC: the pack kernel of the SENDER stores the
// boundary planes straight into the receiver's buffer over NVLink, a one-thread kernel then publishes an epoch
// number in the receiver's flag word (system-scope fence + store), and the receiver's unpack kernel is preceded by a
// one-thread kernel that spins on its own flag words.  No collective call and no host synchronisation is involved;
// buffers and flags are double-buffered by epoch parity, which makes reuse safe: a rank packs exchange e+2 only
// after it has unpacked exchange e+1, i.e. after its neighbour packed e+1, which the neighbour did after unpacking e.
// ---------------------------------------------------------------------------------------------------------------
struct PeerLink {
    size_t cap = 0;                       // bytes per (parity, side) buffer
    unsigned char* block = nullptr;       // mine: [256 B flags][2 parities][2 sides][cap]
    unsigned char* above = nullptr;       // the same block of the rank above / below (IPC mappings; may coincide)
    unsigned char* below = nullptr;
    unsigned long long epoch = 0;
    int* err = nullptr;                   // device flag: a wait timed out
};
static PeerLink g_link;
static const size_t LINK_HDR = 256;

static void link_release() {
    PeerLink& L = g_link;
    if (!L.block) return;
    cudaStreamSynchronize(stream());
    barrier();
    cudaStreamSynchronize(stream());
    if (L.above && L.above != L.block) cudaIpcCloseMemHandle(L.above);
    if (L.below && L.below != L.block && L.below != L.above) cudaIpcCloseMemHandle(L.below);
    cudaFree(L.block);
    L.block = L.above = L.below = nullptr;
    L.cap = 0;
}

// collective: every rank calls it with the same size
void peer_halo_prepare(size_t bytes_per_side) {
    PeerLink& L = g_link;
    if (!active()) throw Error("peer halo link without an initialised communicator");
    if (bytes_per_side <= L.cap) return;
    link_release();
    size_t cap = ((bytes_per_side * 5 / 4 + 4095) / 4096) * 4096;       // head room: the field list may grow
    size_t total = LINK_HDR + 4 * cap;
    OB_CUDA(cudaMalloc(&L.block, total));
    OB_CUDA(cudaMemsetAsync(L.block, 0, LINK_HDR, stream()));
    if (!L.err) { OB_CUDA(cudaMalloc(&L.err, sizeof(int))); OB_CUDA(cudaMemsetAsync(L.err, 0, sizeof(int), stream())); }
    cudaIpcMemHandle_t mine;
    OB_CUDA(cudaIpcGetMemHandle(&mine, L.block));
    cudaIpcMemHandle_t *dsend, *drecv;
    OB_CUDA(cudaMalloc(&dsend, sizeof(mine)));
    OB_CUDA(cudaMalloc(&drecv, sizeof(mine) * g_size));
    OB_CUDA(cudaMemcpyAsync(dsend, &mine, sizeof(mine), cudaMemcpyHostToDevice, stream()));
    allgather_bytes(dsend, drecv, sizeof(mine));
    std::vector<cudaIpcMemHandle_t> all(g_size);
    OB_CUDA(cudaMemcpyAsync(all.data(), drecv, sizeof(mine) * g_size, cudaMemcpyDeviceToHost, stream()));
    OB_CUDA(cudaStreamSynchronize(stream()));
    cudaFree(dsend); cudaFree(drecv);
    const int up = (g_rank + 1) % g_size, dn = (g_rank - 1 + g_size) % g_size;
    auto open = [&](int r) -> unsigned char* {
        if (r == g_rank) return L.block;
        void* q = nullptr;
        OB_CUDA(cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess));
        return (unsigned char*)q;
    };
    L.above = open(up);
    L.below = dn == up ? L.above : open(dn);
    L.cap = cap;
    L.epoch = 0;
    barrier();                 // every rank has zeroed its flags before anybody publishes
    OB_CUDA(cudaStreamSynchronize(stream()));
}

static unsigned char* link_buf(unsigned char* block, size_t cap, int parity, int side) {
    return block + LINK_HDR + (size_t)(2 * parity + side) * cap;
}
static unsigned long long* link_flag(unsigned char* block, int parity, int side) {
    return reinterpret_cast<unsigned long long*>(block) + (2 * parity + side);
}

__global__ void link_signal_kernel(unsigned long long* f0, unsigned long long* f1, unsigned long long epoch) {
    __threadfence_system();
    unsigned long long* f = threadIdx.x == 0 ? f0 : f1;
    if (f) *reinterpret_cast<volatile unsigned long long*>(f) = epoch;
}
// The wait gives up only after `timeout_ns` (OB200_LINK_TIMEOUT_S, default 600 s: a rank may legitimately be seconds
// late -- output, plan creation, a debugger) and then it is FATAL: the error word is set and the kernel traps, so the
// unpack kernel behind it never copies stale halo planes into a field and every later call on this context fails
// (cudaErrorLaunchFailure -> status != 0 from every entry point that touches the stream) instead of carrying on.
__global__ void link_wait_kernel(const unsigned long long* f0, const unsigned long long* f1, unsigned long long epoch, int* err,
                                 unsigned long long timeout_ns) {
    const volatile unsigned long long* f = threadIdx.x == 0 ? f0 : f1;
    if (f) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*f < epoch) {
            __nanosleep(128);
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {
                *err = 1;
                __threadfence_system();
                __trap();
            }
        }
    }
    __threadfence_system();
}
static unsigned long long link_timeout_ns() {
    static unsigned long long ns = 0;
    if (!ns) {
        const char* e = getenv("OB200_LINK_TIMEOUT_S");
        double s = e ? atof(e) : 600.0;
        if (!(s > 0)) s = 600.0;
        ns = (unsigned long long)(s * 1e9);
    }
    return ns;
}

// side 0 = low halo, 1 = high halo.  begin(): next epoch; send_ptr(side): where MY boundary planes for the neighbour's
// halo `side` go (side 1: my bottom planes -> the rank below's high halo; side 0: my top planes -> the rank above's low halo);
// recv_ptr(side): my own buffer for my halo `side`.
void peer_halo_begin() { g_link.epoch += 1; }
// The parity double buffering is safe as long as a rank never packs exchange e+2 into a neighbour that has not unpacked
// exchange e.  Two-sided exchanges wait on both neighbours, so that holds by construction; ONE-sided exchanges (the single
// planes around the pressure solve) are only safe in alternation with a collective in between (the distributed solve), which
// is how model_pressure_step issues them.  Two one-sided exchanges of the same side in a row would break it: refuse them.
static int g_last_one_sided = -1;
void peer_halo_check_pattern(bool need_lo, bool need_hi) {
    const int kind = (need_lo && need_hi) ? -1 : (need_lo ? 0 : 1);
    if (kind >= 0 && kind == g_last_one_sided)
        throw Error("peer halo link: two one-sided exchanges of the same side in a row (unsafe buffer reuse)");
    g_last_one_sided = kind;
}
void* peer_halo_send_ptr(int side) {
    PeerLink& L = g_link;
    return link_buf(side == 1 ? L.below : L.above, L.cap, (int)(L.epoch & 1), side);
}
void* peer_halo_recv_ptr(int side) {
    PeerLink& L = g_link;
    return link_buf(L.block, L.cap, (int)(L.epoch & 1), side);
}
void peer_halo_signal(bool lo, bool hi) {
    PeerLink& L = g_link;
    const int par = (int)(L.epoch & 1);
    link_signal_kernel<<<1, 2, 0, stream()>>>(lo ? link_flag(L.above, par, 0) : nullptr, hi ? link_flag(L.below, par, 1) : nullptr, L.epoch);
    count_launch(1);
}
void peer_halo_wait(bool lo, bool hi) {
    PeerLink& L = g_link;
    const int par = (int)(L.epoch & 1);
    link_wait_kernel<<<1, 2, 0, stream()>>>(lo ? link_flag(L.block, par, 0) : nullptr, hi ? link_flag(L.block, par, 1) : nullptr, L.epoch, L.err, link_timeout_ns());
    count_launch(1);
}
bool peer_halo_error() {
    if (!g_link.err) return false;
    int e = 0;
    cudaMemcpy(&e, g_link.err, sizeof(int), cudaMemcpyDeviceToHost);
    return e != 0;
}
}  // namespace comm
}  // namespace ob
