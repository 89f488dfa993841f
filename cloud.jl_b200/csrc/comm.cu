// comm.cu — the element-partitioned residual behind the C ABI (SURVEY.md §8b "Threading", §8e).
//
// The reference has no distributed path; the only cross-element coupling of semi_discrete_residual! is the gather of
// neighbour facet states through mesh.mapP between its two element loops (Solvers.jl:505-511; BR1 adds a second gather,
// standard_form_second_order.jl:20, 55-64).  One handle = one GPU = one NCCL rank holding a contiguous element range ordered
// interior-first; ghost facet slots follow the owned ones in u_f / q_f.  Per residual:
//
//     pass A (all elements) -> pack cut faces -> [comm stream] ncclSend / ncclRecv with every neighbour rank
//                           -> pass B on the interior elements (overlaps the transfer)
//                           -> unpack ghosts -> pass B on the halo-adjacent elements
//
// Two ways to form the communicator, both ending in the same per-handle state:
//   * one process per GPU  : sse_comm_unique_id on rank 0, the host broadcasts the 128 bytes, sse_comm_init on every rank;
//   * one process, N GPUs  : sse_comm_init_all over N handles (ncclCommInitAll), then sse_rhs_multi / sse_step_ck54_multi drive
//                            all of them from the single host thread the Julia caller is (ODEProblem is single-threaded).
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already loaded by the host process if there is one), so the
// library has no link-time dependency on it and single-GPU users never load it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>       // types and prototypes only; every call goes through the table below

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "handle.h"

using namespace sse;

// ------------------------------------------------------------------------------ NCCL binding
namespace {
struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};
NcclApi g_nccl;

const NcclApi* nccl_api() {
    if (g_nccl.lib) return &g_nccl;
    const char* names[] = {getenv("SSE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    }
    if (!lib) { fail(SSE_ERR_COMM, "cannot load NCCL (libnccl.so.2; set SSE_NCCL_LIB): %s", dlerror()); return nullptr; }
    NcclApi a;
    a.lib = lib;
#define BIND(f)                                                                                         \
    a.f = (decltype(a.f))dlsym(lib, "nccl" #f);                                                         \
    if (!a.f) { fail(SSE_ERR_COMM, "NCCL symbol nccl" #f " missing"); return nullptr; }
    BIND(GetUniqueId) BIND(CommInitRank) BIND(CommInitAll) BIND(CommDestroy) BIND(Send) BIND(Recv) BIND(AllReduce)
    BIND(GroupStart) BIND(GroupEnd) BIND(GetErrorString) BIND(GetVersion)
#undef BIND
    g_nccl = a;
    return &g_nccl;
}
}  // namespace

#define NC_(x)                                                                                                      \
    do {                                                                                                            \
        ncclResult_t r_ = (x);                                                                                      \
        if (r_ != ncclSuccess) return fail(SSE_ERR_COMM, "%s failed: %s (%s:%d)", #x, api->GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------ communicator lifetime
static int32_t comm_streams(sse_handle* h) {
    sse_comm& c = h->comm;
    CU(cudaSetDevice(h->device));
    if (!c.s_comm) CU(cudaStreamCreateWithFlags(&c.s_comm, cudaStreamNonBlocking));
    if (!c.e_packed) CU(cudaEventCreateWithFlags(&c.e_packed, cudaEventDisableTiming));
    if (!c.e_done) CU(cudaEventCreateWithFlags(&c.e_done, cudaEventDisableTiming));
    if (!c.e_sent) CU(cudaEventCreateWithFlags(&c.e_sent, cudaEventDisableTiming));
    return SSE_OK;
}

extern "C" int32_t sse_comm_unique_id(uint8_t* id128) {
    if (!id128) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    const NcclApi* api = nccl_api();
    if (!api) return SSE_ERR_COMM;
    static_assert(sizeof(ncclUniqueId) == 128, "sse_comm_unique_id hands out 128 bytes");
    ncclUniqueId id;
    NC_(api->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return SSE_OK;
}

extern "C" int32_t sse_comm_init(sse_handle* h, const uint8_t* id128, int32_t rank, int32_t world) {
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return fail(SSE_ERR_BAD_ARGUMENT, "bad communicator arguments");
    if (h->comm.nccl) return fail(SSE_ERR_BAD_ARGUMENT, "handle already has a communicator");
    const NcclApi* api = nccl_api();
    if (!api) return SSE_ERR_COMM;
    int32_t rc;
    if ((rc = comm_streams(h))) return rc;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    NC_(api->CommInitRank(&comm, world, id, rank));
    h->comm.nccl = comm; h->comm.rank = rank; h->comm.world = world;
    return SSE_OK;
}

extern "C" int32_t sse_comm_init_all(sse_handle* const* hs, int32_t n) {
    if (!hs || n < 1) return fail(SSE_ERR_BAD_ARGUMENT, "bad communicator arguments");
    for (int i = 0; i < n; i++)
        if (!hs[i] || hs[i]->comm.nccl) return fail(SSE_ERR_BAD_ARGUMENT, "handle %d is null or already has a communicator", i);
    const NcclApi* api = nccl_api();
    if (!api) return SSE_ERR_COMM;
    std::vector<int> devs((size_t)n);
    std::vector<ncclComm_t> comms((size_t)n, nullptr);
    for (int i = 0; i < n; i++) devs[(size_t)i] = hs[i]->device;
    NC_(api->CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; i++) {
        int32_t rc;
        if ((rc = comm_streams(hs[i]))) return rc;
        hs[i]->comm.nccl = comms[(size_t)i]; hs[i]->comm.rank = i; hs[i]->comm.world = n;
    }
    return SSE_OK;
}

// One process, several handles, no NCCL: the halos move by peer-to-peer copies (cudaMemcpyPeerAsync: NVLink when peer access
// is available, a device-local copy when two partitions share a GPU).  Only sse_rhs_multi / sse_step_ck54_multi drive such handles.
extern "C" int32_t sse_comm_init_local(sse_handle* const* hs, int32_t n) {
    if (!hs || n < 1) return fail(SSE_ERR_BAD_ARGUMENT, "bad communicator arguments");
    for (int i = 0; i < n; i++)
        if (!hs[i] || hs[i]->comm.nccl || !hs[i]->comm.peers.empty()) return fail(SSE_ERR_BAD_ARGUMENT, "handle %d is null or already has a communicator", i);
    for (int i = 0; i < n; i++) {
        int32_t rc;
        if ((rc = comm_streams(hs[i]))) return rc;
        hs[i]->comm.rank = i; hs[i]->comm.world = n;
        hs[i]->comm.peers.assign(hs, hs + n);
        for (int j = 0; j < n; j++) {
            if (hs[j]->device == hs[i]->device) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, hs[i]->device, hs[j]->device) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(hs[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(SSE_ERR_CUDA, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
                cudaGetLastError();
            }
        }
    }
    return SSE_OK;
}

extern "C" int32_t sse_comm_info(const sse_handle* h, int32_t* rank, int32_t* world, int32_t* nccl_version) {
    if (!h) return fail(SSE_ERR_BAD_ARGUMENT, "null handle");
    if (rank) *rank = h->comm.rank;
    if (world) *world = h->comm.world;
    if (nccl_version) {
        *nccl_version = 0;
        if (h->comm.nccl) { const NcclApi* api = nccl_api(); int v = 0; if (api && api->GetVersion(&v) == ncclSuccess) *nccl_version = v; }
    }
    return SSE_OK;
}

void sse::comm_release(sse_handle* h) {
    sse_comm& c = h->comm;
    if (c.nccl && g_nccl.lib) g_nccl.CommDestroy((ncclComm_t)c.nccl);
    c.nccl = nullptr;
    if (c.s_comm) cudaStreamDestroy(c.s_comm);
    if (c.e_packed) cudaEventDestroy(c.e_packed);
    if (c.e_done) cudaEventDestroy(c.e_done);
    if (c.e_sent) cudaEventDestroy(c.e_sent);
    c.e_sent = nullptr; c.peers.clear();
    c.s_comm = nullptr; c.e_packed = c.e_done = nullptr;
}

// The halo plan of this rank: neighbour ranks, facet nodes sent to / received from each (segments of the packed buffers, in this
// order; the receive segments are the ghost slots in order), the 1-based send list (as sse_halo_configure takes it) and the
// number of interior elements -- elements [0, n_interior) must not read a ghost slot, which is verified here against the
// per-element flags sse_create derived from mapP.  A neighbour equal to this rank (periodic wrap inside one partition) is a
// device-local copy.
extern "C" int32_t sse_halo_plan(sse_handle* h, int32_t n_nbr, const int32_t* nbr_rank, const int64_t* send_count, const int64_t* recv_count,
                                 const int64_t* send_index, int64_t n_interior) {
    if (!h || n_nbr < 0 || (n_nbr > 0 && (!nbr_rank || !send_count || !recv_count))) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    if (n_interior < 0 || n_interior > h->cfg.N_e) return fail(SSE_ERR_BAD_ARGUMENT, "n_interior out of range");
    long long ns = 0, nr = 0;
    for (int i = 0; i < n_nbr; i++) {
        if (send_count[i] < 0 || recv_count[i] < 0) return fail(SSE_ERR_BAD_ARGUMENT, "negative halo count");
        if (nbr_rank[i] < 0 || nbr_rank[i] >= h->comm.world) return fail(SSE_ERR_BAD_ARGUMENT, "neighbour rank %d outside the communicator (%d ranks)", nbr_rank[i], h->comm.world);
        if (nbr_rank[i] == h->comm.rank && send_count[i] != recv_count[i]) return fail(SSE_ERR_BAD_ARGUMENT, "self halo must send what it receives");
        ns += send_count[i]; nr += recv_count[i];
    }
    if (nr != h->cfg.N_ghost) return fail(SSE_ERR_BAD_ARGUMENT, "receive counts sum to %lld, the handle has %lld ghost facet nodes", nr, (long long)h->cfg.N_ghost);
    // interior elements must not read ghosts: verified on the neighbour lists built from mapP at sse_create
    for (long long k = 0; k < n_interior; k++)
        if (h->nbr_ghost.size() > (size_t)k && h->nbr_ghost[(size_t)k]) return fail(SSE_ERR_BAD_ARGUMENT, "element %lld reads a ghost facet but lies in the interior range [0, %lld)", k, (long long)n_interior);
    int32_t rc;
    if ((rc = sse_halo_configure(h, send_index, ns))) return rc;
    if ((rc = comm_streams(h))) return rc;
    sse_comm& c = h->comm;
    c.nbr_rank.assign(nbr_rank, nbr_rank + n_nbr);
    c.send_count.assign(send_count, send_count + n_nbr);
    c.recv_count.assign(recv_count, recv_count + n_nbr);
    c.n_interior = n_interior;
    c.planned = true;
    return SSE_OK;
}

// ------------------------------------------------------------------------------ halo exchange
// pack on the handle's stream, then hand over to the comm stream
static int32_t exchange_pack(sse_handle* h, int which) {
    int32_t rc;
    CU(cudaSetDevice(h->device));      // h->stream may be the default stream: it names a stream of the CURRENT device
    if ((rc = sse_halo_pack(h, which))) return rc;
    CU(cudaEventRecord(h->comm.e_packed, h->stream));
    CU(cudaStreamWaitEvent(h->comm.s_comm, h->comm.e_packed, 0));
    return SSE_OK;
}
// the sends / receives of one handle (inside an NCCL group opened by the caller)
static int32_t exchange_post(sse_handle* h, int which, const NcclApi* api) {
    sse_comm& c = h->comm;
    CU(cudaSetDevice(h->device));
    const int nv = h->cfg.N_c * (which ? h->cfg.d : 1);
    long long so = 0, ro = 0;
    for (size_t i = 0; i < c.nbr_rank.size(); i++) {
        double* sb = h->d_send + so * nv;
        double* rb = h->d_recv + ro * nv;
        const size_t ns = (size_t)c.send_count[i] * nv, nr = (size_t)c.recv_count[i] * nv;
        so += c.send_count[i]; ro += c.recv_count[i];
        if (c.nbr_rank[i] == c.rank) {
            CU(cudaMemcpyAsync(rb, sb, ns * sizeof(double), cudaMemcpyDeviceToDevice, c.s_comm));
            continue;
        }
        if (!c.peers.empty()) {        // push into the neighbour's receive segment that is fed by this rank
            sse_handle* p = c.peers[(size_t)c.nbr_rank[i]];
            long long off = 0;
            size_t j = 0;
            for (; j < p->comm.nbr_rank.size(); j++) { if (p->comm.nbr_rank[j] == c.rank) break; off += p->comm.recv_count[j]; }
            if (j == p->comm.nbr_rank.size() || p->comm.recv_count[j] != c.send_count[i])
                return fail(SSE_ERR_COMM, "halo plans of ranks %d and %d do not match", c.rank, c.nbr_rank[i]);
            CU(cudaMemcpyPeerAsync(p->d_recv + off * nv, p->device, sb, h->device, ns * sizeof(double), c.s_comm));
            continue;
        }
        if (!c.nccl) return fail(SSE_ERR_COMM, "halo neighbour %d but no communicator: call sse_comm_init first", c.nbr_rank[i]);
        if (ns) NC_(api->Send(sb, ns, ncclDouble, c.nbr_rank[i], (ncclComm_t)c.nccl, c.s_comm));
        if (nr) NC_(api->Recv(rb, nr, ncclDouble, c.nbr_rank[i], (ncclComm_t)c.nccl, c.s_comm));
    }
    return SSE_OK;
}
static bool needs_nccl(const sse_handle* h) {
    if (!h->comm.peers.empty()) return false;
    for (int r : h->comm.nbr_rank) if (r != h->comm.rank) return true;
    return false;
}
// one rank per process: the whole exchange of this handle
static int32_t exchange_start(sse_handle* h, int which) {
    int32_t rc;
    if (!h->comm.peers.empty()) return fail(SSE_ERR_COMM, "handles joined by sse_comm_init_local are driven by sse_rhs_multi / sse_step_ck54_multi");
    if ((rc = exchange_pack(h, which))) return rc;
    const NcclApi* api = nullptr;
    if (needs_nccl(h) && !(api = nccl_api())) return SSE_ERR_COMM;
    if (api) NC_(api->GroupStart());
    rc = exchange_post(h, which, api);
    if (api) { ncclResult_t r = api->GroupEnd(); if (rc == SSE_OK && r != ncclSuccess) rc = fail(SSE_ERR_COMM, "ncclGroupEnd failed: %s", api->GetErrorString(r)); }
    if (rc) return rc;
    CU(cudaEventRecord(h->comm.e_done, h->comm.s_comm));
    return SSE_OK;
}
static int32_t exchange_finish(sse_handle* h, int which) {
    CU(cudaSetDevice(h->device));
    CU(cudaStreamWaitEvent(h->stream, h->comm.e_done, 0));
    return sse_halo_unpack(h, which);
}

static int32_t require_plan(const sse_handle* h) {
    if (!h->comm.planned)
        return fail(SSE_ERR_COMM, "handle has ghost facets but no halo plan: call sse_comm_init / sse_halo_plan, or drive pass_a / halo exchange / pass_b explicitly");
    return SSE_OK;
}

// stages of the partitioned residual between the exchanges; `stage` 0: after pass A, 1: after the first exchange ... (see dist_rhs)
static int32_t interior_then(sse_handle* h, double* d_dudt, bool aux, RkStage rk) {
    const long long ni = h->comm.n_interior;
    return aux ? sse_rhs_pass_aux(h, d_dudt, 0, ni) : pass_b_stage(h, d_dudt, 0, ni, rk);
}
static int32_t boundary_then(sse_handle* h, double* d_dudt, bool aux, RkStage rk) {
    const long long ni = h->comm.n_interior, nb = h->cfg.N_e - ni;
    return aux ? sse_rhs_pass_aux(h, d_dudt, ni, nb) : pass_b_stage(h, d_dudt, ni, nb, rk);
}

int32_t sse::dist_rhs(sse_handle* h, const double* d_u, double* d_dudt, RkStage rk) {
    if (!h || !d_u || !d_dudt) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    int32_t rc;
    if ((rc = require_plan(h))) return rc;
    if ((rc = sse_rhs_pass_a(h, d_u))) return rc;
    if ((rc = exchange_start(h, 0))) return rc;
    if (h->second_order) {                                     // BR1: u_f halo -> auxiliary variable -> q_f halo -> time derivative
        if ((rc = interior_then(h, d_dudt, true, rk))) return rc;
        if ((rc = exchange_finish(h, 0))) return rc;
        if ((rc = boundary_then(h, d_dudt, true, rk))) return rc;
        if ((rc = exchange_start(h, 1))) return rc;
        if ((rc = interior_then(h, d_dudt, false, rk))) return rc;
        if ((rc = exchange_finish(h, 1))) return rc;
        return boundary_then(h, d_dudt, false, rk);
    }
    if ((rc = interior_then(h, d_dudt, false, rk))) return rc;    // overlaps the transfer
    if ((rc = exchange_finish(h, 0))) return rc;
    return boundary_then(h, d_dudt, false, rk);
}

// ------------------------------------------------------------------------------ one process, several GPUs
static int32_t exchange_all(sse_handle* const* hs, int n, int which) {
    int32_t rc;
    for (int i = 0; i < n; i++) { CU(cudaSetDevice(hs[i]->device)); if ((rc = exchange_pack(hs[i], which))) return rc; }
    if (!hs[0]->comm.peers.empty()) {
        // peer copies: a sender may write a neighbour's receive buffer only after that neighbour is past its own pack (all
        // its earlier unpacks are then done), and a receiver's data is complete when every neighbour has sent
        for (int i = 0; i < n; i++) {
            sse_comm& c = hs[i]->comm;
            CU(cudaSetDevice(hs[i]->device));
            for (int r : c.nbr_rank) CU(cudaStreamWaitEvent(c.s_comm, c.peers[(size_t)r]->comm.e_packed, 0));
            if ((rc = exchange_post(hs[i], which, nullptr))) return rc;
            CU(cudaEventRecord(c.e_sent, c.s_comm));
        }
        for (int i = 0; i < n; i++) {
            sse_comm& c = hs[i]->comm;
            CU(cudaSetDevice(hs[i]->device));
            for (int r : c.nbr_rank) CU(cudaStreamWaitEvent(c.s_comm, c.peers[(size_t)r]->comm.e_sent, 0));
            CU(cudaEventRecord(c.e_done, c.s_comm));
        }
        return SSE_OK;
    }
    const NcclApi* api = nullptr;
    bool any = false;
    for (int i = 0; i < n; i++) any = any || needs_nccl(hs[i]);
    if (any && !(api = nccl_api())) return SSE_ERR_COMM;
    if (api) NC_(api->GroupStart());
    rc = SSE_OK;
    for (int i = 0; i < n && rc == SSE_OK; i++) { cudaSetDevice(hs[i]->device); rc = exchange_post(hs[i], which, api); }
    if (api) { ncclResult_t r = api->GroupEnd(); if (rc == SSE_OK && r != ncclSuccess) rc = fail(SSE_ERR_COMM, "ncclGroupEnd failed: %s", api->GetErrorString(r)); }
    if (rc) return rc;
    for (int i = 0; i < n; i++) { CU(cudaSetDevice(hs[i]->device)); CU(cudaEventRecord(hs[i]->comm.e_done, hs[i]->comm.s_comm)); }
    return SSE_OK;
}
static int32_t multi_rhs(sse_handle* const* hs, int n, double* const* d_u, double* const* d_dudt, const RkStage* rks) {
    int32_t rc;
    for (int i = 0; i < n; i++) {
        if (!hs[i] || !d_u[i] || !d_dudt[i]) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
        if (hs[i]->cfg.N_ghost == 0 && n > 1) return fail(SSE_ERR_BAD_ARGUMENT, "handle %d has no ghost facets", i);
        if ((rc = require_plan(hs[i]))) return rc;
        if (hs[i]->second_order != hs[0]->second_order) return fail(SSE_ERR_BAD_ARGUMENT, "handles of one residual must share the conservation law");
    }
    const bool second = hs[0]->second_order;
    for (int i = 0; i < n; i++) if ((rc = sse_rhs_pass_a(hs[i], d_u[i]))) return rc;
    if ((rc = exchange_all(hs, n, 0))) return rc;
    if (second) {
        for (int i = 0; i < n; i++) if ((rc = interior_then(hs[i], d_dudt[i], true, rks[i]))) return rc;
        for (int i = 0; i < n; i++) if ((rc = exchange_finish(hs[i], 0)) || (rc = boundary_then(hs[i], d_dudt[i], true, rks[i]))) return rc;
        if ((rc = exchange_all(hs, n, 1))) return rc;
    }
    for (int i = 0; i < n; i++) if ((rc = interior_then(hs[i], d_dudt[i], false, rks[i]))) return rc;
    for (int i = 0; i < n; i++) if ((rc = exchange_finish(hs[i], second ? 1 : 0)) || (rc = boundary_then(hs[i], d_dudt[i], false, rks[i]))) return rc;
    return SSE_OK;
}
extern "C" int32_t sse_rhs_multi(sse_handle* const* hs, int32_t n, double* const* d_u, double* const* d_dudt, double t) {
    (void)t;
    if (!hs || !d_u || !d_dudt || n < 1) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    std::vector<RkStage> rks((size_t)n);
    return multi_rhs(hs, n, d_u, d_dudt, rks.data());
}
// Carpenter & Kennedy (1994) 2N-storage RK4(5) (the coefficients of sse_step_ck54)
static const double MCK_A[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                                -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
static const double MCK_B[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                                1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                                2277821191437.0 / 14882151754819.0};
extern "C" int32_t sse_step_ck54_multi(sse_handle* const* hs, int32_t n, double* const* d_u, double* const* d_tmp, double* const* d_dudt,
                                       double t, double dt) {
    (void)t;
    if (!hs || !d_u || !d_tmp || !d_dudt || n < 1) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    int32_t rc;
    std::vector<RkStage> rks((size_t)n);
    for (int s = 0; s < 5; s++) {
        for (int i = 0; i < n; i++) {
            if (!hs[i] || !d_tmp[i]) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
            RkStage rk;
            if (hs[i]->variant == 1 && hs[i]->ct.ok) { rk.u = d_u[i]; rk.tmp = d_tmp[i]; rk.A = MCK_A[s]; rk.B = MCK_B[s]; rk.dt = dt; }
            rks[(size_t)i] = rk;
        }
        if ((rc = multi_rhs(hs, n, d_u, d_dudt, rks.data()))) return rc;
        for (int i = 0; i < n; i++)
            if (!rks[(size_t)i].u && (rc = sse_lsrk_stage(hs[i], d_u[i], d_tmp[i], d_dudt[i], MCK_A[s], MCK_B[s], dt))) return rc;
    }
    return SSE_OK;
}

// ------------------------------------------------------------------------------ reductions
int32_t sse::dist_allreduce_sum(sse_handle* h, double* d_buf, int n) {
    if (!h->comm.nccl || h->comm.world == 1) return SSE_OK;
    const NcclApi* api = nccl_api();
    if (!api) return SSE_ERR_COMM;
    NC_(api->AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)h->comm.nccl, h->stream));
    return SSE_OK;
}

// ------------------------------------------------------------------------------ host-buffer residual of one rank
// semi_discrete_residual! on this rank's HOST arrays: the halo-adjacent elements are uploaded and put through pass A first so
// that the facet halos travel while the interior ranges are still being uploaded; pass B of an interior range starts as soon
// as pass A has covered its face neighbours (mapP), and its slice of dudt goes back while later ranges arrive; the
// halo-adjacent elements finish after the ghosts have been unpacked.  Synchronous.
int32_t sse::dist_rhs_host(sse_handle* h, const double* h_u, double* h_dudt, double* d_u, double* d_du, int32_t chunks) {
    int32_t rc;
    const long long ne = h->cfg.N_e, ni = h->comm.n_interior;
    const size_t per = (size_t)h->cfg.N_p * h->cfg.N_c;
    if (chunks <= 0) chunks = std::max(6, 48 / std::max(1, h->comm.world));
    if (ni < 4LL * chunks) chunks = (int)std::max<long long>(1, ni / 4);
    std::vector<std::pair<long long, long long>> ranges;
    if (ne > ni) ranges.push_back({ni, ne});
    for (int c = 0; c < chunks && ni > 0; c++) ranges.push_back({ni * c / chunks, ni * (c + 1) / chunks});
    const int n = (int)ranges.size();
    if (h->plan_chunks != -chunks) {                      // cached per range count (negative: the partitioned plan)
        make_range_plan_general(ranges, ne, h->cfg.N_fac, h->nbr, h->nbr_hi, h->plan_ready);
        h->plan_chunks = -chunks;
    }
    const std::vector<int>& ready = h->plan_ready;
    while (h->events.size() < (size_t)(2 * n + 2)) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->events.push_back(e);
    }
    cudaEvent_t e_start = h->events[(size_t)(2 * n)], e_done = h->events[(size_t)(2 * n + 1)];
    CU(cudaEventRecord(e_start, h->stream));
    CU(cudaStreamWaitEvent(h->s_in, e_start, 0));
    CU(cudaStreamWaitEvent(h->s_out, e_start, 0));
    const int first_interior = (ne > ni) ? 1 : 0;
    auto pass_b_download = [&](int k) -> int32_t {
        const long long a = ranges[(size_t)k].first, b = ranges[(size_t)k].second;
        int32_t r;
        if ((r = sse_rhs_pass_b(h, d_du, a, b - a))) return r;
        CU(cudaEventRecord(h->events[(size_t)(n + k)], h->stream));
        CU(cudaStreamWaitEvent(h->s_out, h->events[(size_t)(n + k)], 0));
        CU(cudaMemcpyAsync(h_dudt + per * (size_t)a, d_du + per * (size_t)a, per * (size_t)(b - a) * sizeof(double), cudaMemcpyDeviceToHost, h->s_out));
        return SSE_OK;
    };
    for (int i = 0; i < n; i++) {
        const long long a = ranges[(size_t)i].first, b = ranges[(size_t)i].second;
        CU(cudaMemcpyAsync(d_u + per * (size_t)a, h_u + per * (size_t)a, per * (size_t)(b - a) * sizeof(double), cudaMemcpyHostToDevice, h->s_in));
        CU(cudaEventRecord(h->events[(size_t)i], h->s_in));
        CU(cudaStreamWaitEvent(h->stream, h->events[(size_t)i], 0));
        if ((rc = sse_rhs_pass_a_range(h, d_u, a, b - a))) return rc;
        if (i == 0 && first_interior && (rc = exchange_start(h, 0))) return rc;     // cut faces are complete: halos travel from here on
        for (int k = first_interior; k < n; k++)
            if (ready[(size_t)k] == i && (rc = pass_b_download(k))) return rc;
    }
    if (first_interior) {
        if ((rc = exchange_finish(h, 0))) return rc;
        if ((rc = pass_b_download(0))) return rc;
    }
    CU(cudaEventRecord(e_done, h->s_out));
    CU(cudaStreamWaitEvent(h->stream, e_done, 0));
    CU(cudaStreamSynchronize(h->stream));
    return check_flag(h);
}
