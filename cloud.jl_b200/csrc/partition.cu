// partition.cu — host-only helper: the local view of one rank of an element partition (SURVEY.md §8e "Partitioning").
//
// The reference has one mesh and one connectivity array, mesh.mapP (N_f x N_e, 1-based linear indices into the (N_f, N_e)
// facet array; Solvers.jl:207, flux_differencing_form.jl:313).  Given that array and the owner rank of every element -- slabs,
// blocks, or the output of any graph partitioner run on the face adjacency mapP implies -- sse_partition_create builds what a
// rank needs to create its handle and its halo plan:
//   * its elements, interior first (no ghost reads), halo-adjacent last, each group in ascending global order;
//   * mapP rewritten to local + ghost numbering: owned facet nodes keep (local element, node), remote ones point at ghost
//     slots N_f n_local + g;
//   * ghost slots grouped by neighbour rank (ascending), inside a rank ordered by ascending GLOBAL facet-node index -- the
//     order the owner uses for its send list, so both sides agree without exchanging anything;
//   * the send list per neighbour: the owned facet nodes some element of that neighbour reads, in the same global order.
// No device is touched; the Julia side calls it once per rank after (or instead of) its own partitioner.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "handle.h"

using namespace sse;

struct sse_partition {
    int32_t N_f = 0;
    std::vector<int64_t> elem_gid;        // 1-based global element ids, interior first
    int64_t n_interior = 0, n_ghost = 0;
    std::vector<int64_t> mapP_local;      // N_f x n_local, 1-based
    std::vector<int32_t> nbr_rank;
    std::vector<int64_t> send_count, recv_count, send_index;
};

extern "C" int32_t sse_partition_create(const int64_t* mapP, int64_t N_e, int32_t N_f, const int32_t* owner, int32_t n_parts, int32_t rank,
                                        sse_partition** out) {
    if (!mapP || !owner || !out || N_e < 1 || N_f < 1 || n_parts < 1 || rank < 0 || rank >= n_parts) return fail(SSE_ERR_BAD_ARGUMENT, "bad partition arguments");
    *out = nullptr;
    const int64_t NF = (int64_t)N_f;
    for (int64_t k = 0; k < N_e; k++)
        if (owner[k] < 0 || owner[k] >= n_parts) return fail(SSE_ERR_BAD_ARGUMENT, "owner[%lld] = %d outside [0, %d)", (long long)k, owner[k], n_parts);
    for (int64_t t = 0; t < NF * N_e; t++)
        if (mapP[t] < 1 || mapP[t] > NF * N_e) return fail(SSE_ERR_BAD_ARGUMENT, "mapP[%lld] = %lld out of range (BoundsError)", (long long)t, (long long)mapP[t]);
    sse_partition* p = new sse_partition();
    p->N_f = N_f;
    // local elements: interior first
    std::vector<int64_t> interior, boundary;
    for (int64_t k = 0; k < N_e; k++) {
        if (owner[k] != rank) continue;
        bool halo = false;
        for (int64_t j = 0; j < NF && !halo; j++) halo = owner[(mapP[k * NF + j] - 1) / NF] != rank;
        (halo ? boundary : interior).push_back(k);
    }
    p->n_interior = (int64_t)interior.size();
    std::vector<int64_t> loc(interior);
    loc.insert(loc.end(), boundary.begin(), boundary.end());
    const int64_t nl = (int64_t)loc.size();
    if (nl == 0) { delete p; return fail(SSE_ERR_BAD_ARGUMENT, "rank %d owns no element", rank); }
    std::vector<int64_t> g2l((size_t)N_e, -1);
    for (int64_t i = 0; i < nl; i++) g2l[(size_t)loc[(size_t)i]] = i;
    p->elem_gid.resize((size_t)nl);
    for (int64_t i = 0; i < nl; i++) p->elem_gid[(size_t)i] = loc[(size_t)i] + 1;
    // remote facet nodes this rank reads (0-based global linear ids), per owner rank, sorted and unique
    std::vector<std::vector<int64_t>> need((size_t)n_parts), give((size_t)n_parts);
    for (int64_t b : boundary)
        for (int64_t j = 0; j < NF; j++) {
            const int64_t gq = mapP[b * NF + j] - 1;
            const int32_t o = owner[gq / NF];
            if (o != rank) need[(size_t)o].push_back(gq);
        }
    // owned facet nodes read by elements of other ranks: scan the elements adjacent to ours (they are exactly the elements our
    // halo-adjacent elements point to when the mesh is conforming; a full scan keeps non-symmetric connectivities correct)
    for (int64_t k = 0; k < N_e; k++) {
        const int32_t o = owner[k];
        if (o == rank) continue;
        for (int64_t j = 0; j < NF; j++) {
            const int64_t gq = mapP[k * NF + j] - 1;
            if (owner[gq / NF] == rank) give[(size_t)o].push_back(gq);
        }
    }
    std::vector<int64_t> ghost_base((size_t)n_parts, -1);
    int64_t ghost = 0;
    for (int32_t r = 0; r < n_parts; r++) {
        auto& nd = need[(size_t)r];
        auto& gv = give[(size_t)r];
        std::sort(nd.begin(), nd.end()); nd.erase(std::unique(nd.begin(), nd.end()), nd.end());
        std::sort(gv.begin(), gv.end()); gv.erase(std::unique(gv.begin(), gv.end()), gv.end());
        if (nd.empty() && gv.empty()) continue;
        p->nbr_rank.push_back(r);
        p->recv_count.push_back((int64_t)nd.size());
        p->send_count.push_back((int64_t)gv.size());
        ghost_base[(size_t)r] = ghost;
        ghost += (int64_t)nd.size();
        for (int64_t gq : gv) p->send_index.push_back(g2l[(size_t)(gq / NF)] * NF + gq % NF + 1);
    }
    p->n_ghost = ghost;
    // mapP in local + ghost numbering
    p->mapP_local.resize((size_t)(NF * nl));
    for (int64_t i = 0; i < nl; i++)
        for (int64_t j = 0; j < NF; j++) {
            const int64_t gq = mapP[loc[(size_t)i] * NF + j] - 1;
            const int32_t o = owner[gq / NF];
            int64_t v;
            if (o == rank) v = g2l[(size_t)(gq / NF)] * NF + gq % NF;
            else {
                const auto& nd = need[(size_t)o];
                v = NF * nl + ghost_base[(size_t)o] + (int64_t)(std::lower_bound(nd.begin(), nd.end(), gq) - nd.begin());
            }
            p->mapP_local[(size_t)(i * NF + j)] = v + 1;
        }
    *out = p;
    return SSE_OK;
}

extern "C" int32_t sse_partition_sizes(const sse_partition* p, int64_t* n_local, int64_t* n_interior, int64_t* n_ghost, int32_t* n_nbr, int64_t* n_send) {
    if (!p) return fail(SSE_ERR_BAD_ARGUMENT, "null partition");
    if (n_local) *n_local = (int64_t)p->elem_gid.size();
    if (n_interior) *n_interior = p->n_interior;
    if (n_ghost) *n_ghost = p->n_ghost;
    if (n_nbr) *n_nbr = (int32_t)p->nbr_rank.size();
    if (n_send) *n_send = (int64_t)p->send_index.size();
    return SSE_OK;
}

extern "C" int32_t sse_partition_fill(const sse_partition* p, int64_t* elem_gid, int64_t* mapP_local, int32_t* nbr_rank, int64_t* send_count,
                                      int64_t* recv_count, int64_t* send_index) {
    if (!p) return fail(SSE_ERR_BAD_ARGUMENT, "null partition");
    if (elem_gid) std::copy(p->elem_gid.begin(), p->elem_gid.end(), elem_gid);
    if (mapP_local) std::copy(p->mapP_local.begin(), p->mapP_local.end(), mapP_local);
    if (nbr_rank) std::copy(p->nbr_rank.begin(), p->nbr_rank.end(), nbr_rank);
    if (send_count) std::copy(p->send_count.begin(), p->send_count.end(), send_count);
    if (recv_count) std::copy(p->recv_count.begin(), p->recv_count.end(), recv_count);
    if (send_index) std::copy(p->send_index.begin(), p->send_index.end(), send_index);
    return SSE_OK;
}

extern "C" int32_t sse_partition_destroy(sse_partition* p) {
    delete p;
    return SSE_OK;
}
