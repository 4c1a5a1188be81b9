// comm.cu -- NCCL plumbing for the slab-decomposed path (reference: src/Distributed, which uses
// MPI.Isend/Irecv for halos and PencilFFTs all-to-all transposes; here NCCL over NVLink/NVSwitch).
// libnccl.so.2 is bound lazily with dlopen so that single-GPU use has no NCCL dependency; in a
// process that already loaded torch's bundled NCCL the same library instance is reused.
#include "internal.h"
#include <dlfcn.h>
#include <cstring>

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
void destroy() {
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
}  // namespace comm
}  // namespace ob
