// tests/emu/emu_nccl.cpp — TEST INFRASTRUCTURE: in-process stand-in for NCCL (see tests/emu/nccl.h). One OS thread per rank;
// point-to-point messages go through per-(source, destination) mailboxes, collectives through a generation barrier.
#include "nccl.h"

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <vector>

namespace {
struct World {
    int nranks = 0, joined = 0, refs = 0;
    std::mutex m;
    std::condition_variable cv;
    std::map<std::pair<int, int>, std::deque<std::vector<char>>> mail;   // (src, dst) -> messages in order
    // barrier
    int waiting = 0;
    unsigned long long gen = 0;
    std::vector<const void*> deposit;
};
std::mutex g_worldsLock;
std::map<unsigned long long, World*> g_worlds;
unsigned long long g_nextId = 1;

void barrier(World& w, std::unique_lock<std::mutex>& lk) {
    const unsigned long long g = w.gen;
    if (++w.waiting == w.nranks) { w.waiting = 0; w.gen++; w.cv.notify_all(); }
    else w.cv.wait(lk, [&] { return w.gen != g; });
}
size_t dtSize(ncclDataType_t dt) { return dt == ncclUint8 ? 1 : (dt == ncclUint32 ? 4 : 8); }

struct Op { bool send; const void* sbuf; void* rbuf; size_t bytes; int peer; EmuNcclComm* comm; };
thread_local int t_group = 0;
thread_local std::vector<Op> t_ops;
}  // namespace

struct EmuNcclComm { World* world; int rank; };

namespace {
ncclResult_t flush() {
    std::vector<Op> ops;
    ops.swap(t_ops);
    for (const Op& o : ops) {                       // all sends first: nobody blocks before its messages are out
        if (!o.send) continue;
        World& w = *o.comm->world;
        std::lock_guard<std::mutex> lk(w.m);
        w.mail[{o.comm->rank, o.peer}].emplace_back((const char*)o.sbuf, (const char*)o.sbuf + o.bytes);
        w.cv.notify_all();
    }
    for (const Op& o : ops) {
        if (o.send) continue;
        World& w = *o.comm->world;
        std::unique_lock<std::mutex> lk(w.m);
        auto& q = w.mail[{o.peer, o.comm->rank}];
        w.cv.wait(lk, [&] { return !q.empty(); });
        std::vector<char> msg = std::move(q.front());
        q.pop_front();
        if (msg.size() != o.bytes) return ncclInvalidArgument;
        std::memcpy(o.rbuf, msg.data(), o.bytes);
    }
    return ncclSuccess;
}
}  // namespace

extern "C" {
ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
    std::lock_guard<std::mutex> lk(g_worldsLock);
    std::memset(id, 0, sizeof(*id));
    const unsigned long long v = g_nextId++;
    std::memcpy(id->internal, &v, sizeof(v));
    return ncclSuccess;
}
ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank) {
    unsigned long long v;
    std::memcpy(&v, id.internal, sizeof(v));
    World* w;
    {
        std::lock_guard<std::mutex> lk(g_worldsLock);
        World*& slot = g_worlds[v];
        if (!slot) { slot = new World(); slot->nranks = nranks; slot->deposit.assign(nranks, nullptr); }
        w = slot;
        w->refs++;
    }
    if (w->nranks != nranks) return ncclInvalidArgument;
    *comm = new EmuNcclComm{w, rank};
    std::unique_lock<std::mutex> lk(w->m);
    barrier(*w, lk);
    return ncclSuccess;
}
ncclResult_t ncclCommDestroy(ncclComm_t comm) {
    if (!comm) return ncclSuccess;
    {
        std::lock_guard<std::mutex> lk(g_worldsLock);
        comm->world->refs--;   // worlds are tiny and ids are never reused: they are left to the end of the process
    }
    delete comm;
    return ncclSuccess;
}
ncclResult_t ncclGroupStart() { t_group++; return ncclSuccess; }
ncclResult_t ncclGroupEnd() {
    if (--t_group > 0) return ncclSuccess;
    return flush();
}
ncclResult_t ncclSend(const void* buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, cudaStream_t) {
    t_ops.push_back(Op{true, buf, nullptr, count * dtSize(dt), peer, comm});
    return t_group ? ncclSuccess : flush();
}
ncclResult_t ncclRecv(void* buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t comm, cudaStream_t) {
    t_ops.push_back(Op{false, nullptr, buf, count * dtSize(dt), peer, comm});
    return t_group ? ncclSuccess : flush();
}
ncclResult_t ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t dt, ncclRedOp_t op, ncclComm_t comm, cudaStream_t) {
    World& w = *comm->world;
    std::unique_lock<std::mutex> lk(w.m);
    w.deposit[comm->rank] = send;
    barrier(w, lk);
    std::vector<char> out(count * dtSize(dt));
    auto reduce = [&](auto zero) {
        using T = decltype(zero);
        T* o = reinterpret_cast<T*>(out.data());
        for (size_t k = 0; k < count; k++) {
            T acc = reinterpret_cast<const T*>(w.deposit[0])[k];
            for (int r = 1; r < w.nranks; r++) {
                const T v = reinterpret_cast<const T*>(w.deposit[r])[k];
                acc = op == ncclSum ? (T)(acc + v) : (v < acc ? v : acc);
            }
            o[k] = acc;
        }
    };
    if (dt == ncclUint32) reduce((uint32_t)0); else if (dt == ncclUint64) reduce((unsigned long long)0); else reduce((unsigned char)0);
    barrier(w, lk);                        // everyone has read every deposit before anyone overwrites an in-place buffer
    std::memcpy(recv, out.data(), out.size());
    barrier(w, lk);
    return ncclSuccess;
}
const char* ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "emulated NCCL error"; }
}
