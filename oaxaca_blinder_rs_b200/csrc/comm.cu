// csrc/comm.cu -- collectives for row sharding across the GPUs of one box (mode N).
//
// The data path needs exactly three exchanges per panel batch: the replicate column sums of the Poisson body
// (int64 all-reduce, a few KB), the per-rank Gram sums (all-gather, [2][slots][P] doubles per rank; combined in
// fixed tree order by gram_combine so the result is bit-identical to a single GPU), and the status flags.
// Everything else of the bootstrap is local to a GPU.
//
//  * NcclComm: one process per GPU.  libnccl.so.2 is opened at run time (the one torch already mapped when the
//    host is Python, else the system library), so libobboot carries no link-time NCCL dependency.
//  * LocalComm: several contexts inside one process (threads).  Ranks publish their device pointers in a shared
//    table, meet at a host barrier and pull peers' buffers with cudaMemcpyAsync (peer copies when the contexts
//    sit on different GPUs, plain device copies when they share one).
#include "internal.h"

#include <condition_variable>
#include <cstring>
#include <dlfcn.h>
#include <mutex>

namespace ob {

namespace {

[[noreturn]] void comm_fail(const std::string& m) { throw StatusError{OB_ERR_NCCL, m}; }

// ------------------------------------------------------------------ NCCL (dlopen)
// The handful of prototypes used, restated from the public nccl.h (stable C ABI since NCCL 2.0).
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueIdT { char internal[128]; };
enum { NCCL_SUCCESS = 0 };
enum { NCCL_INT32 = 2, NCCL_INT64 = 4, NCCL_FLOAT64 = 8, NCCL_INT8 = 0 };
enum { NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueIdT*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueIdT, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("OBBOOT_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.handle, "ncclAllReduce"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(api.handle, "ncclBroadcast"));
        api.Send = reinterpret_cast<decltype(api.Send)>(dlsym(api.handle, "ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(dlsym(api.handle, "ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(api.handle, "ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(api.handle, "ncclGroupEnd"));
    });
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.AllGather || !api.Broadcast ||
        !api.GroupStart || !api.GroupEnd || !api.Send || !api.Recv)
        comm_fail("libnccl.so.2 could not be loaded (set OBBOOT_NCCL_LIB to its path)");
    return api;
}

void nccl_check(int rc, const char* what) {
    if (rc == NCCL_SUCCESS) return;
    NcclApi& a = nccl_api();
    comm_fail(std::string("NCCL error in ") + what + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "?"));
}

struct NcclComm final : Comm {
    ncclComm_t comm = nullptr;
    ~NcclComm() override { if (comm) nccl_api().CommDestroy(comm); }
    void allreduce(void* buf, size_t count, CommDType dt, CommOp op, cudaStream_t st) override {
        const int t = dt == CommDType::I32 ? NCCL_INT32 : dt == CommDType::I64 ? NCCL_INT64 : NCCL_FLOAT64;
        const int o = op == CommOp::SUM ? NCCL_SUM : op == CommOp::MAX ? NCCL_MAX : NCCL_MIN;
        nccl_check(nccl_api().AllReduce(buf, buf, count, t, o, comm, st), "ncclAllReduce");
    }
    void allgather(const void* send, void* recv, size_t bytes, cudaStream_t st) override {
        nccl_check(nccl_api().AllGather(send, recv, bytes, NCCL_INT8, comm, st), "ncclAllGather");
    }
    void allgatherv(const void* send, void* recv, const size_t* offsets, const size_t* sizes, cudaStream_t st) override {
        // one broadcast per root, fused into a single NCCL group: every block travels once over NVLink into place
        NcclApi& a = nccl_api();
        nccl_check(a.GroupStart(), "ncclGroupStart");
        for (int r = 0; r < world; ++r) {
            if (sizes[r] == 0) continue;
            char* dst = static_cast<char*>(recv) + offsets[r];
            nccl_check(a.Broadcast(r == rank ? send : dst, dst, sizes[r], NCCL_INT8, r, comm, st), "ncclBroadcast");
        }
        nccl_check(a.GroupEnd(), "ncclGroupEnd");
    }
    void alltoallv(const void* send, const size_t* send_off, void* recv, const size_t* recv_off, const size_t* bytes,
                   cudaStream_t st) override {
        // point-to-point sends and receives of one NCCL group: every block crosses NVLink once, straight into place
        NcclApi& a = nccl_api();
        const size_t* mine = bytes + (size_t)rank * world;       // bytes[src * world + dst]
        if (mine[rank])
            OB_CUDA(cudaMemcpyAsync(static_cast<char*>(recv) + recv_off[rank], static_cast<const char*>(send) + send_off[rank], mine[rank],
                                    cudaMemcpyDeviceToDevice, st));
        nccl_check(a.GroupStart(), "ncclGroupStart");
        for (int r = 0; r < world; ++r) {
            if (r == rank) continue;
            if (mine[r]) nccl_check(a.Send(static_cast<const char*>(send) + send_off[r], mine[r], NCCL_INT8, r, comm, st), "ncclSend");
            const size_t in = bytes[(size_t)r * world + rank];
            if (in) nccl_check(a.Recv(static_cast<char*>(recv) + recv_off[r], in, NCCL_INT8, r, comm, st), "ncclRecv");
        }
        nccl_check(a.GroupEnd(), "ncclGroupEnd");
    }
};

// ------------------------------------------------------------------ in-process transport
template <typename T>
__global__ void local_reduce_kernel(const T* __restrict__ stage, int world, size_t count, int op, T* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        T acc = stage[i];
        for (int r = 1; r < world; ++r) {   // ascending rank order on every rank -> identical results everywhere
            const T v = stage[(size_t)r * count + i];
            acc = op == 0 ? acc + v : (op == 1 ? (v > acc ? v : acc) : (v < acc ? v : acc));
        }
        out[i] = acc;
    }
}

}  // namespace

struct LocalGroup {
    int world;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long long generation = 0;
    std::vector<const void*> ptrs;
    std::vector<const size_t*> offs;     // alltoallv: every rank's send offsets
    explicit LocalGroup(int w) : world(w), ptrs((size_t)w, nullptr), offs((size_t)w, nullptr) {}
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long long gen = generation;
        if (++arrived == world) { arrived = 0; ++generation; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != gen; });
    }
};

namespace {

struct LocalComm final : Comm {
    LocalGroup* grp = nullptr;
    int device = 0;
    void* stage = nullptr; size_t stage_bytes = 0;
    ~LocalComm() override { if (stage) cudaFree(stage); }
    void need_stage(size_t b) {
        if (b <= stage_bytes) return;
        if (stage) cudaFree(stage);
        OB_CUDA(cudaMalloc(&stage, b)); stage_bytes = b;
    }
    // every rank pulls every rank's `bytes` into dst [world][bytes]
    void pull_all(const void* mine, void* dst, size_t bytes, cudaStream_t st) {
        OB_CUDA(cudaStreamSynchronize(st));          // my buffer is final before peers read it
        grp->ptrs[(size_t)rank] = mine;
        grp->barrier();
        for (int r = 0; r < world; ++r)
            OB_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + (size_t)r * bytes, grp->ptrs[(size_t)r], bytes, cudaMemcpyDefault, st));
        OB_CUDA(cudaStreamSynchronize(st));
        grp->barrier();                              // nobody overwrites a buffer a peer is still reading
    }
    void allreduce(void* buf, size_t count, CommDType dt, CommOp op, cudaStream_t st) override {
        const size_t es = dt == CommDType::I32 ? 4 : 8;
        need_stage(es * count * (size_t)world);
        pull_all(buf, stage, es * count, st);
        const int o = op == CommOp::SUM ? 0 : op == CommOp::MAX ? 1 : 2;
        const unsigned blocks = (unsigned)std::min<size_t>((count + 255) / 256, 1024);
        if (dt == CommDType::I32) local_reduce_kernel<int><<<blocks, 256, 0, st>>>((const int*)stage, world, count, o, (int*)buf);
        else if (dt == CommDType::I64) local_reduce_kernel<long long><<<blocks, 256, 0, st>>>((const long long*)stage, world, count, o, (long long*)buf);
        else local_reduce_kernel<double><<<blocks, 256, 0, st>>>((const double*)stage, world, count, o, (double*)buf);
        OB_CUDA(cudaGetLastError());
    }
    void allgather(const void* send, void* recv, size_t bytes, cudaStream_t st) override { pull_all(send, recv, bytes, st); }
    void allgatherv(const void* send, void* recv, const size_t* offsets, const size_t* sizes, cudaStream_t st) override {
        OB_CUDA(cudaStreamSynchronize(st));
        grp->ptrs[(size_t)rank] = send;
        grp->barrier();
        for (int r = 0; r < world; ++r)
            if (sizes[r])
                OB_CUDA(cudaMemcpyAsync(static_cast<char*>(recv) + offsets[r], grp->ptrs[(size_t)r], sizes[r], cudaMemcpyDefault, st));
        OB_CUDA(cudaStreamSynchronize(st));
        grp->barrier();
    }
    void alltoallv(const void* send, const size_t* send_off, void* recv, const size_t* recv_off, const size_t* bytes,
                   cudaStream_t st) override {
        OB_CUDA(cudaStreamSynchronize(st));
        grp->ptrs[(size_t)rank] = send;
        grp->offs[(size_t)rank] = send_off;
        grp->barrier();
        for (int r = 0; r < world; ++r) {      // pull what rank r holds for me
            const size_t in = bytes[(size_t)r * world + rank];
            if (in)
                OB_CUDA(cudaMemcpyAsync(static_cast<char*>(recv) + recv_off[r],
                                        static_cast<const char*>(grp->ptrs[(size_t)r]) + grp->offs[(size_t)r][rank], in, cudaMemcpyDefault, st));
        }
        OB_CUDA(cudaStreamSynchronize(st));
        grp->barrier();
    }
};

}  // namespace

LocalGroup* local_group_create(int world) { return new LocalGroup(world); }
void local_group_destroy(LocalGroup* g) { delete g; }

Comm* comm_create_local(LocalGroup* g, int rank, int device) {
    auto* c = new LocalComm;
    c->grp = g; c->rank = rank; c->world = g->world; c->device = device;
    return c;
}

void nccl_unique_id(uint8_t* id128) {
    ncclUniqueIdT id;
    nccl_check(nccl_api().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id128, id.internal, 128);
}

Comm* comm_create_nccl(const uint8_t* id128, int rank, int world) {
    ncclUniqueIdT id;
    memcpy(id.internal, id128, 128);
    auto c = std::make_unique<NcclComm>();
    c->rank = rank; c->world = world;
    nccl_check(nccl_api().CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank");
    return c.release();
}

}  // namespace ob
