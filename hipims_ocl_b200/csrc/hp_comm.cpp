// hp_comm.cpp -- see hp_comm.h
#include "hp_comm.h"

#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <string>

#include <nccl.h>

namespace hp {

namespace {

struct Api {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

Api g_api;
thread_local std::string g_err;

const char* load_api() {
    if (g_api.lib) return nullptr;
    // HIPIMS_NCCL_LIB lets the launcher point at the NCCL build torch.distributed already loaded
    const char* names[] = {getenv("HIPIMS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) { g_err = std::string("cannot load NCCL: ") + dlerror(); return g_err.c_str(); }
#define HP_SYM(field, name)                                                           \
    g_api.field = reinterpret_cast<decltype(g_api.field)>(dlsym(lib, name));          \
    if (!g_api.field) { g_err = std::string("NCCL symbol missing: ") + name; return g_err.c_str(); }
    HP_SYM(GetUniqueId, "ncclGetUniqueId")
    HP_SYM(CommInitRank, "ncclCommInitRank")
    HP_SYM(CommDestroy, "ncclCommDestroy")
    HP_SYM(GroupStart, "ncclGroupStart")
    HP_SYM(GroupEnd, "ncclGroupEnd")
    HP_SYM(Send, "ncclSend")
    HP_SYM(Recv, "ncclRecv")
    HP_SYM(AllReduce, "ncclAllReduce")
    HP_SYM(GetErrorString, "ncclGetErrorString")
#undef HP_SYM
    g_api.lib = lib;
    return nullptr;
}

const char* check(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return nullptr;
    g_err = std::string(what) + ": " + g_api.GetErrorString(r);
    return g_err.c_str();
}

}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    bool broken = false;
};

static_assert(sizeof(ncclUniqueId) == 128, "HP_COMM_ID_BYTES");

const char* comm_unique_id(void* id_out_128) {
    if (const char* e = load_api()) return e;
    return check(g_api.GetUniqueId(static_cast<ncclUniqueId*>(id_out_128)), "ncclGetUniqueId");
}

const char* comm_create(Comm** out, const void* id_128, int rank, int world_size) {
    if (const char* e = load_api()) return e;
    Comm* c = new Comm();
    c->rank = rank; c->world = world_size;
    ncclUniqueId id = *static_cast<const ncclUniqueId*>(id_128);
    if (const char* e = check(g_api.CommInitRank(&c->comm, world_size, id, rank), "ncclCommInitRank")) { delete c; return e; }
    *out = c;
    return nullptr;
}

void comm_destroy(Comm* c) {
    if (!c) return;
    if (c->comm && g_api.CommDestroy) g_api.CommDestroy(c->comm);
    delete c;
}

static const char* exchange_calls(Comm* c, const Planes& p, const Grid& g, int halo, size_t rb, cudaStream_t st);

// A failed call inside a group must still close the group: an open ncclGroupStart would silently defer every later
// NCCL call of this thread (torch.distributed's included).  The first error is the one reported; the communicator
// is marked broken and refuses further work (its state after a failed group is undefined).
static const char* close_group(Comm* c, const char* first) {
    const ncclResult_t r = g_api.GroupEnd();
    if (first) { c->broken = true; return first; }
    if (const char* e = check(r, "ncclGroupEnd")) { c->broken = true; return e; }
    return nullptr;
}
static const char* broken_comm() { g_err = "the NCCL communicator failed earlier and cannot be used"; return g_err.c_str(); }

const char* comm_exchange_and_allreduce(Comm* c, const Planes& p, const Grid& g, int halo, size_t rb, unsigned long long* value, cudaStream_t st) {
    if (c->broken) return broken_comm();
    if (const char* e = check(g_api.GroupStart(), "ncclGroupStart")) return e;
    const char* e = exchange_calls(c, p, g, halo, rb, st);
    std::string keep;
    if (e) keep = e;
    else if (const char* e2 = check(g_api.AllReduce(value, value, 1, ncclUint64, ncclMax, c->comm, st), "ncclAllReduce")) keep = e2;
    if (!keep.empty()) { close_group(c, keep.c_str()); g_err = keep; return g_err.c_str(); }
    return close_group(c, nullptr);
}

const char* comm_exchange_halos(Comm* c, const Planes& p, const Grid& g, int halo, size_t rb, cudaStream_t st) {
    if (c->broken) return broken_comm();
    if (const char* e = check(g_api.GroupStart(), "ncclGroupStart")) return e;
    if (const char* e = exchange_calls(c, p, g, halo, rb, st)) { const std::string keep = e; close_group(c, keep.c_str()); g_err = keep; return g_err.c_str(); }
    return close_group(c, nullptr);
}

static const char* exchange_calls(Comm* c, const Planes& p, const Grid& g, int halo, size_t rb, cudaStream_t st) {
    const bool south = c->rank > 0, north = c->rank + 1 < c->world;
    const size_t row = static_cast<size_t>(g.pitch) * rb, bytes = row * halo;
    char* planes[4] = {static_cast<char*>(p.eta), static_cast<char*>(p.emax), static_cast<char*>(p.qx), static_cast<char*>(p.qy)};
    for (char* base : planes) {
        if (south) {   // my lowest owned rows -> southern neighbour's northern halo; its top rows -> my southern halo
            if (const char* e = check(g_api.Send(base + row * g.own_y0, bytes, ncclUint8, c->rank - 1, c->comm, st), "ncclSend")) return e;
            if (const char* e = check(g_api.Recv(base + row * (g.own_y0 - halo), bytes, ncclUint8, c->rank - 1, c->comm, st), "ncclRecv")) return e;
        }
        if (north) {
            if (const char* e = check(g_api.Send(base + row * (g.own_y1 - halo), bytes, ncclUint8, c->rank + 1, c->comm, st), "ncclSend")) return e;
            if (const char* e = check(g_api.Recv(base + row * g.own_y1, bytes, ncclUint8, c->rank + 1, c->comm, st), "ncclRecv")) return e;
        }
    }
    return nullptr;
}

const char* comm_allreduce_max(Comm* c, unsigned long long* value, cudaStream_t st) {
    if (c->broken) return broken_comm();
    return check(g_api.AllReduce(value, value, 1, ncclUint64, ncclMax, c->comm, st), "ncclAllReduce");
}

}  // namespace hp
