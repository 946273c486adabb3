// transpose.cu - pencil transposes zTOx / xTOz (mpi_transpose.f90:50-117) across GPUs.
//
// The z-pass and x-pass kernels already write one contiguous block per destination rank
// ("pack") and read one block per source rank ("unpack"), so a transpose is a pure block
// exchange: rank r sends block j to rank j and receives block q from rank q.  It is issued as
// one grouped ncclSend/ncclRecv all-to-all on the handle's stream for a whole chunk of y-planes
// and all 3 (zTOx) or 6 (xTOz) components at once (the reference issues 9 MPI_Alltoall per plane,
// mpi_transpose.f90:74,109).  NCCL is loaded with dlopen so that single-GPU use has no NCCL
// dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <string>

#include "chb_internal.h"

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) {
        chb_set_error(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
        return false;
    }
#define LOAD(field, sym)                                                   \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, sym);                     \
    if (!g_nccl.field) {                                                   \
        chb_set_error(std::string("libnccl is missing symbol ") + sym);    \
        return false;                                                      \
    }
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(GroupStart, "ncclGroupStart");
    LOAD(GroupEnd, "ncclGroupEnd");
    LOAD(Send, "ncclSend");
    LOAD(Recv, "ncclRecv");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(Broadcast, "ncclBroadcast");
    LOAD(AllGather, "ncclAllGather");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    return true;
}
}  // namespace

#define NCCL_OK(call)                                                                         \
    do {                                                                                      \
        ncclResult_t r__ = (call);                                                            \
        if (r__ != ncclSuccess) {                                                             \
            chb_set_error(std::string(#call) + ": " + g_nccl.GetErrorString(r__));            \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

int chb_nccl_unique_id(char* id) {
    if (!load_nccl()) return 1;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId u;
    NCCL_OK(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return 0;
}

int chb_nccl_init(chb_handle_s* h, const char* id) {
    if (!load_nccl()) return 1;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t comm;
    NCCL_OK(g_nccl.CommInitRank(&comm, h->g.nranks, u, h->g.rank));
    h->nccl_comm = comm;
    return 0;
}

void chb_nccl_destroy(chb_handle_s* h) {
    if (h->nccl_comm) {
        g_nccl.CommDestroy((ncclComm_t)h->nccl_comm);
        h->nccl_comm = nullptr;
    }
}

// all-to-all of `count` complex elements in total (count/nranks per peer)
int chb_alltoall(chb_handle_s* h, const cplx* send, cplx* recv, size_t count) {
    const int P = h->g.nranks;
    const size_t per = count / P;
    ncclComm_t comm = (ncclComm_t)h->nccl_comm;
    ScopedKernelTimer tm(h, "alltoall", h->cstream);
    NCCL_OK(g_nccl.GroupStart());
    for (int p = 0; p < P; ++p) {
        NCCL_OK(g_nccl.Send(send + (size_t)p * per, per * 2, ncclDouble, p, comm, h->cstream));
        NCCL_OK(g_nccl.Recv(recv + (size_t)p * per, per * 2, ncclDouble, p, comm, h->cstream));
    }
    NCCL_OK(g_nccl.GroupEnd());
    return 0;
}

// cfl is a non-negative double stored as its bit pattern: max over uint64 == max over doubles
int chb_allreduce_max_cfl(chb_handle_s* h) {
    ncclComm_t comm = (ncclComm_t)h->nccl_comm;
    NCCL_OK(g_nccl.AllReduce(&h->sc->cfl_bits, &h->sc->cfl_bits, 1, ncclUint64, ncclMax, comm, h->stream));
    return 0;
}

// mean-mode scalars live on the rank with nx0==0 (rank 0, has_average mpi_transpose.f90:217)
int chb_bcast_scalars(chb_handle_s* h) {
    ncclComm_t comm = (ncclComm_t)h->nccl_comm;
    char* base = reinterpret_cast<char*>(h->sc) + sizeof(unsigned long long);
    const size_t bytes = sizeof(DevScalars) - sizeof(unsigned long long);
    NCCL_OK(g_nccl.Broadcast(base, base, bytes, ncclChar, 0, comm, h->stream));
    return 0;
}

// agreement on values every rank derives from its own state (free memory, environment): element-wise minimum
int chb_allreduce_min_i64(chb_handle_s* h, long long* v, int n) {
    ncclComm_t comm = (ncclComm_t)h->nccl_comm;
    long long* d = nullptr;
    CHB_CUDA_OK(cudaMalloc((void**)&d, sizeof(long long) * n));
    CHB_CUDA_OK(cudaMemcpy(d, v, sizeof(long long) * n, cudaMemcpyHostToDevice));
    NCCL_OK(g_nccl.AllReduce(d, d, (size_t)n, ncclInt64, ncclMin, comm, h->stream));
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    CHB_CUDA_OK(cudaMemcpy(v, d, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

// ---- direct NVLink mode ---------------------------------------------------------------------
// Every rank maps every peer's work arena (receive buffers Ar, Br and barrier flags of every lane, same offsets on
// every rank) through CUDA IPC.  The pack side of zTOx / xTOz then IS the transpose: zfwd / xpass store each element
// straight into the owner's HBM over NVLink while they compute (PeerPtrs, chb_internal.h).  What is left of the
// collective is a flag barrier between the producing and the consuming kernel.
int chb_p2p_setup(chb_handle_s* h) {
    const int P = h->g.nranks, r = h->g.rank;
    ncclComm_t comm = (ncclComm_t)h->nccl_comm;
    h->n_ipc_opened = 0;
    cudaIpcMemHandle_t mine;
    CHB_CUDA_OK(cudaIpcGetMemHandle(&mine, h->arena));
    cudaIpcMemHandle_t* d_all = nullptr;
    CHB_CUDA_OK(cudaMalloc((void**)&d_all, sizeof(mine) * P));
    CHB_CUDA_OK(cudaMemcpy(d_all + r, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    NCCL_OK(g_nccl.AllGather(d_all + r, d_all, sizeof(mine), ncclChar, comm, h->stream));
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    cudaIpcMemHandle_t all[CHB_MAX_RANKS];
    CHB_CUDA_OK(cudaMemcpy(all, d_all, sizeof(mine) * P, cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    char* base[CHB_MAX_RANKS];
    for (int q = 0; q < P; ++q) {
        if (q == r) {
            base[q] = h->arena;
            continue;
        }
        void* pa = nullptr;
        CHB_CUDA_OK(cudaIpcOpenMemHandle(&pa, all[q], cudaIpcMemLazyEnablePeerAccess));
        h->ipc_opened[h->n_ipc_opened++] = pa;
        base[q] = (char*)pa;
    }
    for (int L = 0; L < h->nlanes; ++L) {
        Lane& ln = h->lane[L];
        const ptrdiff_t oa = (char*)ln.Ar - h->arena, ob = (char*)ln.Br - h->arena, of = (char*)ln.flags - h->arena;
        for (int q = 0; q < P; ++q) {
            ln.Aw.p[q] = (cplx*)(base[q] + oa);
            ln.Bw.p[q] = (cplx*)(base[q] + ob);
            ln.peer_flags[q] = (unsigned long long*)(base[q] + of);
        }
    }
    chb_select_lane(h, 0);
    return 0;
}

void chb_p2p_teardown(chb_handle_s* h) {
    for (int i = 0; i < h->n_ipc_opened; ++i) cudaIpcCloseMemHandle(h->ipc_opened[i]);
    h->n_ipc_opened = 0;
}

struct FlagPtrs {
    unsigned long long* p[CHB_MAX_RANKS];
};

// thread q: publish "rank `me` has finished epoch e" in peer q's flags, then wait until peer q has
// published the same in mine.  The producing kernel precedes this one on the stream, so its
// (remote) stores are complete; the fences order them with the flag for the other GPUs.
// A peer that never arrives (it failed between two collective calls) must not hang the others forever: after
// `timeout_ns` the kernel gives up, records the fact in *err and the host turns it into an error return
// (chb_p2p_check), the library's form of the reference's STOP.
__global__ void p2p_barrier_kernel(FlagPtrs peers, volatile unsigned long long* mine, int me, int P, unsigned long long e,
                                   unsigned long long* err, unsigned long long timeout_ns) {
    const int q = threadIdx.x;
    if (q >= P) return;
    __threadfence_system();
    *((volatile unsigned long long*)(peers.p[q] + me)) = e;
    __threadfence_system();
#ifndef CHB_HOST_EMUL
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#endif
    while (mine[q] < e) {
        __nanosleep(200);
#ifndef CHB_HOST_EMUL
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) {
            atomicExch(err, ((unsigned long long)(q + 1) << 32) | (e & 0xffffffffull));
            break;
        }
#endif
    }
    __threadfence_system();
}

int chb_p2p_check(chb_handle_s* h) {
    if (!h->p2p) return 0;
    unsigned long long v = 0;
    CHB_CUDA_OK(cudaMemcpyAsync(&v, h->p2p_error, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
    CHB_CUDA_OK(cudaStreamSynchronize(h->stream));
    if (v) {
        chb_set_error("pencil transpose: rank " + std::to_string((int)(v >> 32) - 1) + " did not reach barrier " +
                      std::to_string((unsigned)(v & 0xffffffffull)) + " within the timeout (CHB_BARRIER_TIMEOUT_S); the field is invalid");
        return 6;
    }
    return 0;
}

int chb_exchange(chb_handle_s* h, bool a_side) {
    const Geometry& g = h->g;
    if (g.nranks == 1) return 0;
    if (h->p2p) {
        static const unsigned long long timeout_ns = []() {
            const char* e = getenv("CHB_BARRIER_TIMEOUT_S");
            return (unsigned long long)((e ? atof(e) : 60.0) * 1e9);
        }();
        FlagPtrs fp;
        for (int q = 0; q < g.nranks; ++q) fp.p[q] = h->peer_flags[q];
        ScopedKernelTimer tm(h, "p2p_barrier", h->cstream);
        CHB_LAUNCH(1, 32, 0, h->cstream, p2p_barrier_kernel)(fp, h->flags, g.rank, g.nranks, ++h->lane[h->cur_lane].epoch,
                                                             h->p2p_error, timeout_ns);
        h->launches++;
        return 0;
    }
    const size_t np = h->chunk_planes;
    if (a_side) return chb_alltoall(h, h->A, h->Ar, (size_t)3 * np * g.nzd * g.nxB);   // mpi_transpose.f90:74
    return chb_alltoall(h, h->B, h->Br, (size_t)6 * np * g.nzd * g.nxB);               // mpi_transpose.f90:109
}
