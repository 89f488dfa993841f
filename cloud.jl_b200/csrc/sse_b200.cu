// sse_b200.cu — C ABI of libsse_b200.so (see include/sse_b200.h).
//
// Host side: turns the reference `Solver` image into device tables once (sse_create), then
// every sse_rhs is two (first-order) or three (second-order) kernel launches that never touch
// the host.  No CPU fallback exists: if no CUDA device is usable every entry point reports
// SSE_ERR_CUDA.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sse_b200.h"
#include "common.cuh"
#include "kernels_generic.cuh"
#include "kernels_tensor.cuh"
#include "kernels_geometry.cuh"
#include "ct_api.h"

using namespace sse;

#include "handle.h"

// ------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
int32_t sse::fail(int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

template <class T>
static int32_t upload(sse_handle* h, const std::vector<T>& v, const T** out) {
    void* p = nullptr;
    size_t n = std::max<size_t>(v.size(), 1) * sizeof(T);
    CU(cudaMalloc(&p, n));
    h->owned.push_back(p);
    if (!v.empty()) CU(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)p;
    return SSE_OK;
}
template <class T>
static int32_t upload_raw(sse_handle* h, const T* src, size_t count, const T** out) {
    void* p = nullptr;
    CU(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    h->owned.push_back(p);
    if (count) CU(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)p;
    return SSE_OK;
}
static int32_t dalloc(sse_handle* h, size_t count, double** out) {
    void* p = nullptr;
    CU(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(double)));
    h->owned.push_back(p);
    *out = (double*)p;
    return SSE_OK;
}

struct HostSp {
    std::vector<int> ptr, idx;
    std::vector<double> val;
};
// rows of a column-major dense nrow x ncol matrix (transpose = true: rows of A^T)
static HostSp compress(const double* A, int nrow, int ncol, bool transpose) {
    HostSp s;
    int nr = transpose ? ncol : nrow, nc = transpose ? nrow : ncol;
    s.ptr.assign(nr + 1, 0);
    for (int r = 0; r < nr; r++) {
        s.ptr[r] = (int)s.idx.size();
        for (int c = 0; c < nc; c++) {
            double v = transpose ? A[c + (size_t)nrow * r] : A[r + (size_t)nrow * c];
            if (v != 0.0) { s.idx.push_back(c); s.val.push_back(v); }
        }
    }
    s.ptr[nr] = (int)s.idx.size();
    return s;
}
static int32_t upload_sp(sse_handle* h, const HostSp& s, SpMat* out) {
    int32_t rc;
    if ((rc = upload(h, s.ptr, &out->ptr))) return rc;
    if ((rc = upload(h, s.idx, &out->idx))) return rc;
    if ((rc = upload(h, s.val, &out->val))) return rc;
    const int rows = (int)s.ptr.size() - 1;
    int w = 0;
    for (int r = 0; r < rows; r++) w = std::max(w, s.ptr[(size_t)r + 1] - s.ptr[(size_t)r]);
    std::vector<int> ei((size_t)std::max(w, 1) * std::max(rows, 1), -1);
    std::vector<double> ev(ei.size(), 0.0);
    const int sq = w > 16 ? rows : 1, sr = w > 16 ? 1 : std::max(w, 1);
    for (int r = 0; r < rows; r++)
        for (int q = s.ptr[(size_t)r]; q < s.ptr[(size_t)r + 1]; q++) {
            ei[(size_t)(q - s.ptr[(size_t)r]) * sq + (size_t)r * sr] = s.idx[(size_t)q];
            ev[(size_t)(q - s.ptr[(size_t)r]) * sq + (size_t)r * sr] = s.val[(size_t)q];
        }
    out->rows = rows; out->w = w; out->sq = sq; out->sr = sr;
    if ((rc = upload(h, ei, &out->ei))) return rc;
    return upload(h, ev, &out->ev);
}

// ------------------------------------------------------------------------------ dispatch
#define DISPATCH_DNC(h, CALL)                                                  \
    do {                                                                       \
        const int d_ = (h)->cfg.d, nc_ = (h)->cfg.N_c;                         \
        if (d_ == 1 && nc_ == 1) { CALL(1, 1); }                               \
        else if (d_ == 1 && nc_ == 3) { CALL(1, 3); }                          \
        else if (d_ == 2 && nc_ == 1) { CALL(2, 1); }                          \
        else if (d_ == 2 && nc_ == 4) { CALL(2, 4); }                          \
        else if (d_ == 3 && nc_ == 1) { CALL(3, 1); }                          \
        else if (d_ == 3 && nc_ == 5) { CALL(3, 5); }                          \
        else return fail(SSE_ERR_UNSUPPORTED, "unsupported (d, N_c) = (%d, %d)", d_, nc_); \
    } while (0)

static size_t smem_nodal_bytes(const Ops& o) {
    int NC = o.NC;
    return sizeof(double) * (size_t)(o.Np * NC + 2 * o.Nq * NC + o.Nf * NC + warp_z_size(o, NC) + warp_w_size(o, NC));
}
static size_t smem_time_bytes(const sse_handle* h) {
    const Ops& o = h->ops;
    int NC = o.NC, D = o.d;
    size_t n = 0;
    if (h->cfg.form == SSE_FORM_FLUX_DIFFERENCING)
        n = 2 * o.Nq * NC + 3 * o.Nf * NC + D * o.Nf + o.Nq * D * D + o.Np * NC;
    else if (h->cfg.form == SSE_FORM_STANDARD_REFERENCE)
        n = 2 * o.Nq * NC + o.Nq * NC * D + 3 * o.Nf * NC + D * o.Nf + o.Nq * D * D + o.Np * NC;
    else
        n = o.Nq * NC + 2 * o.Nq * NC * D + 3 * o.Nf * NC + D * o.Nf;
    n += warp_z_size(o, NC) + warp_w_size(o, NC);
    if (h->cfg.form == SSE_FORM_FLUX_DIFFERENCING) n += (size_t)D * o.Nf;                                    // halfnJf
    if (h->cfg.form == SSE_FORM_FLUX_DIFFERENCING && NC == D + 2) n += (size_t)(D + 3) * (o.Nq + o.Nf);     // primitive tables of the fast EC path
    return n * sizeof(double);
}
static size_t smem_aux_bytes(const Ops& o) {
    int NC = o.NC, D = o.d;
    return sizeof(double) * (size_t)(2 * o.Nq * NC + 2 * o.Nf * NC + D * o.Nf + o.Np * NC + warp_z_size(o, NC) + warp_w_size(o, NC));
}

// The attribute is a property of the kernel template, not of a handle: a later handle with smaller elements must not lower the
// cap under an earlier one, so every runtime-sized kernel is opted in to the device maximum once (launches pass their own size).
static const int SMEM_OPT_IN_MAX = 227 * 1024;
template <int D, int NC>
static int32_t set_attrs(sse_handle* h) {
    CU(cudaFuncSetAttribute(k_nodal_generic<D, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN_MAX));
    CU(cudaFuncSetAttribute(k_time_fluxdiff_generic<D, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN_MAX));
    CU(cudaFuncSetAttribute(k_time_standard_reference<D, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN_MAX));
    CU(cudaFuncSetAttribute(k_time_physical<D, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN_MAX));
    CU(cudaFuncSetAttribute(k_aux_physical<D, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN_MAX));
    return tensor_set_attrs<D, NC>(h->tp) == cudaSuccess ? SSE_OK : fail(SSE_ERR_CUDA, "cudaFuncSetAttribute (tensor kernels) failed");
}

// ------------------------------------------------------------------------------ create / destroy
extern "C" int32_t sse_abi_version(void) { return SSE_ABI_VERSION; }
extern "C" const char* sse_last_error_string(void) { return g_err.c_str(); }

// ---- schedule of the pipelined host-buffer residual (sse_rhs_host); pure host logic, exported as sse_host_range_plan so
//      that it is testable without a device
// local face neighbours of every element from mapP: the highest one (nbr_hi) and, when no element has more than N_fac
// distinct ones, the exact list (nbr, -1 padded; cleared otherwise).  Ghost slots lie beyond the local elements.
static void collect_neighbours(const int64_t* mapP, long long Ne, int Nf, int Nfac, std::vector<long long>& nbr_hi, std::vector<long long>& nbr) {
    nbr_hi.assign((size_t)Ne, 0);
    nbr.assign((size_t)Ne * Nfac, -1);
    bool few = true;
    for (long long k = 0; k < Ne; k++) {
        long long hi = k;
        int cnt = 0;
        for (int j = 0; j < Nf; j++) {
            const long long nb = (mapP[(size_t)k * Nf + j] - 1) / Nf;
            if (nb >= Ne) continue;
            if (nb > hi) hi = nb;
            if (!few || nb == k) continue;
            bool seen = false;
            for (int q = 0; q < cnt; q++) seen = seen || nbr[(size_t)k * Nfac + q] == nb;
            if (!seen) { if (cnt < Nfac) nbr[(size_t)k * Nfac + cnt++] = nb; else few = false; }
        }
        nbr_hi[(size_t)k] = hi;
    }
    if (!few) nbr.clear();
}
// ranges c = [ne c / chunks, ne (c + 1) / chunks): upload order and, for every range, the upload position after which its
// pass B may run (all ranges holding one of its face neighbours are through pass A).  With exact neighbour lists the last
// range goes first: on a periodic slab-ordered mesh range 0 otherwise waits for the wrap-around neighbour until the very end.
static void make_range_plan(long long ne, int chunks, int nfac, const std::vector<long long>& nbr, const std::vector<long long>& nbr_hi,
                            std::vector<int>& order, std::vector<int>& ready) {
    std::vector<long long> bounds((size_t)chunks + 1);
    for (int c = 0; c <= chunks; c++) bounds[(size_t)c] = ne * c / chunks;
    auto owner = [&](long long k) { int c = (int)((k * chunks) / ne); while (k >= bounds[(size_t)c + 1]) c++; while (k < bounds[(size_t)c]) c--; return c; };
    const bool exact = !nbr.empty();
    order.resize((size_t)chunks);
    std::vector<int> pos((size_t)chunks);
    for (int i = 0; i < chunks; i++) order[(size_t)i] = exact ? (i == 0 ? chunks - 1 : i - 1) : i;
    for (int i = 0; i < chunks; i++) pos[(size_t)order[(size_t)i]] = i;
    ready.assign((size_t)chunks, 0);
    for (int c = 0; c < chunks; c++) {
        int r = pos[(size_t)c];
        if (exact) {
            for (long long k = bounds[(size_t)c]; k < bounds[(size_t)c + 1]; k++)
                for (int q = 0; q < nfac; q++) {
                    const long long nb = nbr[(size_t)k * nfac + q];
                    if (nb >= 0) r = std::max(r, pos[(size_t)owner(nb)]);
                }
        } else {
            long long hi = bounds[(size_t)c];
            for (long long k = bounds[(size_t)c]; k < bounds[(size_t)c + 1]; k++) hi = std::max(hi, nbr_hi[(size_t)k]);
            r = std::max(r, owner(hi));
        }
        ready[(size_t)c] = r;
    }
}
// the same for an arbitrary list of contiguous ranges given in upload order (the multi-GPU host-buffer residual uploads the
// halo-adjacent elements first): ready[i] = upload position after which pass B of range i may run
void sse::make_range_plan_general(const std::vector<std::pair<long long, long long>>& ranges, long long ne, int nfac,
                                  const std::vector<long long>& nbr, const std::vector<long long>& nbr_hi, std::vector<int>& ready) {
    (void)nbr_hi;
    const int n = (int)ranges.size();
    std::vector<int> pos_of((size_t)ne, 0);
    for (int i = 0; i < n; i++)
        for (long long k = ranges[(size_t)i].first; k < ranges[(size_t)i].second; k++) pos_of[(size_t)k] = i;
    ready.assign((size_t)n, n - 1);                      // without exact neighbour lists: after the last upload
    if (nbr.empty()) return;
    for (int i = 0; i < n; i++) {
        int r = i;
        for (long long k = ranges[(size_t)i].first; k < ranges[(size_t)i].second; k++)
            for (int q = 0; q < nfac; q++) {
                const long long nb = nbr[(size_t)k * nfac + q];
                if (nb >= 0) r = std::max(r, pos_of[(size_t)nb]);
            }
        ready[(size_t)i] = r;
    }
}
extern "C" int32_t sse_host_range_plan(const int64_t* mapP, int64_t N_e, int32_t N_f, int32_t N_fac, int32_t chunks, int32_t* order, int32_t* ready) {
    if (!mapP || !order || !ready || N_e <= 0 || N_f <= 0 || N_fac <= 0 || chunks <= 0 || N_e < chunks) return fail(SSE_ERR_BAD_ARGUMENT, "bad argument");
    std::vector<long long> hi, nbr;
    collect_neighbours(mapP, N_e, N_f, N_fac, hi, nbr);
    std::vector<int> o, r;
    make_range_plan(N_e, chunks, N_fac, nbr, hi, o, r);
    for (int c = 0; c < chunks; c++) { order[c] = o[(size_t)c]; ready[c] = r[(size_t)c]; }
    return SSE_OK;
}

// W / J_q per volume node, with the very expression the projection kernel used to evaluate in its prologue (bit-identical)
static __global__ void k_ijw(long long n, int Nq, const double* __restrict__ W, const double* __restrict__ J, double* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = W[i % Nq] * rcp_fast(J[i]);
}

// C tensor of the collapsed tet in the order of the compile-time kernels: C3[l * N + a3], l = canonical modal index
static int32_t upload_c3(sse_handle* h, const sse_arrays* a, int N) {
    std::vector<double> c3;
    for (int b1 = 0; b1 < N; b1++)
        for (int b2 = 0; b1 + b2 < N; b2++)
            for (int b3 = 0; b1 + b2 + b3 < N; b3++)
                for (int a3 = 0; a3 < N; a3++) c3.push_back(a->C[a3 + N * (b1 + N * (b2 + N * b3))]);
    return upload(h, c3, &h->ct.dev.C3);
}

static int32_t build(sse_handle* h, const sse_config* cfg, const sse_arrays* a) {
    const int d = cfg->d, NC = cfg->N_c, Np = cfg->N_p, Nq = cfg->N_q, Nf = cfg->N_f, Nfac = cfg->N_fac;
    const long long Ne = cfg->N_e;
    int32_t rc;
    Ops& o = h->ops;
    memset(&o, 0, sizeof(o));
    o.d = d; o.NC = NC; o.Np = Np; o.Nq = Nq; o.Nf = Nf; o.Nfac = Nfac; o.npf = Nf / Nfac;
    o.v_kind = cfg->v_kind;
    if (!a->R || !a->W || !a->Bf || !a->J_q || !a->J_f || !a->nJf || !a->mapP)
        return fail(SSE_ERR_BAD_ARGUMENT, "R, W, Bf, J_q, J_f, nJf and mapP are required");
    // ---- V
    if (cfg->v_kind == SSE_V_IDENTITY) {
        if (Np != Nq) return fail(SSE_ERR_BAD_ARGUMENT, "identity V needs N_p == N_q (DimensionMismatch)");
    } else if (cfg->v_kind == SSE_V_DENSE) {
        if (!a->V) return fail(SSE_ERR_BAD_ARGUMENT, "dense V missing");
        if ((rc = upload_raw(h, a->V, (size_t)Nq * Np, &o.Vd))) return rc;
    } else if (cfg->v_kind == SSE_V_WARPED) {
        if (d < 2 || !a->A || !a->B || (d == 3 && !a->C) || !a->sigma_i || !a->sigma_o)
            return fail(SSE_ERR_BAD_ARGUMENT, "warped V needs A, B, (C), sigma_i, sigma_o and d >= 2");
        const int P1 = cfg->p + 1;
        if (P1 > 8) return fail(SSE_ERR_UNSUPPORTED, "p + 1 > 8 not supported by the warped kernels");
        o.P1 = P1; o.M1 = cfg->M1d[0]; o.M2 = cfg->M1d[1]; o.M3 = (d == 3) ? cfg->M1d[2] : 1;
        if (o.M1 * o.M2 * o.M3 != Nq) return fail(SSE_ERR_BAD_ARGUMENT, "warped V: M1*M2*M3 != N_q");
        std::vector<double> C((size_t)o.M3 * P1 * P1 * P1, 0.0);
        std::vector<int> si((size_t)P1 * P1 * P1, -1), so((size_t)Nq, 0);
        if (d == 3) {
            std::copy(a->C, a->C + C.size(), C.begin());
            for (size_t t = 0; t < si.size(); t++) si[t] = (int)a->sigma_i[t] - 1;
        } else {
            for (int b1 = 0; b1 < P1; b1++)
                for (int b2 = 0; b2 < P1; b2++) {
                    C[0 + 1 * (b1 + P1 * (b2 + P1 * 0))] = 1.0;
                    si[b1 + P1 * (b2 + P1 * 0)] = (int)a->sigma_i[b1 + P1 * b2] - 1;
                }
        }
        for (int t = 0; t < Nq; t++) so[t] = (int)a->sigma_o[t] - 1;
        int count = 0;
        for (int b1 = 0; b1 < P1; b1++) {
            int n2 = 0;
            for (int b2 = 0; b2 < P1; b2++) {
                int n3 = 0;
                for (int b3 = 0; b3 < P1; b3++) n3 += si[b1 + P1 * (b2 + P1 * b3)] >= 0;
                o.N3[b1 * 8 + b2] = n3;
                n2 += n3 > 0;
                count += n3;
            }
            o.N2[b1] = n2;
        }
        if (count != Np) return fail(SSE_ERR_BAD_ARGUMENT, "warped V: count(sigma_i > 0) != N_p");
        if ((rc = upload_raw(h, a->A, (size_t)o.M1 * P1, &o.A))) return rc;
        if ((rc = upload_raw(h, a->B, (size_t)o.M2 * P1 * P1, &o.B))) return rc;
        if ((rc = upload(h, C, &o.C))) return rc;
        if ((rc = upload(h, si, &o.sig_i))) return rc;
        if ((rc = upload(h, so, &o.sig_o))) return rc;
        if (a->V) { if ((rc = upload_raw(h, a->V, (size_t)Nq * Np, &o.Vd))) return rc; }
        // Small elements (triangles and low-order tetrahedra of the generic / tensor-line kernels): the N_q x N_p matrix of the
        // warped product is a few kB, and one dense row per thread -- fixed trip count, coalesced, no index arithmetic -- beats
        // the three sum-factorised phases with their table-driven bounds (Euler on p = 4 triangles: 5 + 3 applications of
        // V or V' per residual were 2/3 of the time).  V[sigma_o(a), sigma_i(b)] = A[a1,b1] B[a2,b1,b2] C[a3,b1,b2,b3].
        const char* vs = getenv("SSE_V_SMALL");
        if ((size_t)Nq * Np <= 2048 && !(vs && atoi(vs) == 0)) {
            std::vector<double> Vs((size_t)Nq * Np, 0.0);
            for (int b1 = 0; b1 < P1; b1++)
                for (int b2 = 0; b2 < o.N2[b1]; b2++)
                    for (int b3 = 0; b3 < o.N3[b1 * 8 + b2]; b3++) {
                        const int j = si[b1 + P1 * (b2 + P1 * b3)];
                        if (j < 0) return fail(SSE_ERR_BAD_ARGUMENT, "warped V: the valid modes of a (b1, b2) fibre must be contiguous");
                        for (int a1 = 0; a1 < o.M1; a1++)
                            for (int a2 = 0; a2 < o.M2; a2++)
                                for (int a3 = 0; a3 < o.M3; a3++)
                                    Vs[so[a1 + o.M1 * (a2 + o.M2 * a3)] + (size_t)Nq * j] =
                                        a->A[a1 + o.M1 * b1] * a->B[a2 + o.M2 * (b1 + P1 * b2)] * C[a3 + o.M3 * (b1 + P1 * (b2 + P1 * b3))];
                    }
            if ((rc = upload(h, Vs, &o.Vd))) return rc;
            o.v_small = 1;
        }
    } else
        return fail(SSE_ERR_BAD_ARGUMENT, "unknown v_kind %d", cfg->v_kind);
    // ---- R, W, B
    if ((rc = upload_sp(h, compress(a->R, Nf, Nq, false), &o.R))) return rc;
    if ((rc = upload_sp(h, compress(a->R, Nf, Nq, true), &o.Rt))) return rc;
    if ((rc = upload_raw(h, a->W, Nq, &o.W))) return rc;
    if ((rc = upload_raw(h, a->Bf, Nf, &o.Bf))) return rc;
    std::vector<double> nref((size_t)d * Nfac, 0.0);
    if (a->nref) std::copy(a->nref, a->nref + nref.size(), nref.begin());
    if ((rc = upload(h, nref, &o.nref))) return rc;
    // ---- form-specific operators
    Geo& g = h->geo;
    memset(&g, 0, sizeof(g));
    g.Ne = Ne; g.NFT = (long long)Nf * Ne + cfg->N_ghost; g.mass_solver = cfg->mass_solver;
    {
        int* hf = nullptr;
        CU(cudaHostAlloc((void**)&hf, sizeof(int), cudaHostAllocMapped));
        *hf = 0;
        h->h_flag = hf;
        CU(cudaHostGetDevicePointer((void**)&g.flag, hf, 0));
    }
    if (cfg->mass_solver != SSE_MASS_WEIGHT_ADJUSTED && cfg->mass_solver != SSE_MASS_DIAGONAL && cfg->mass_solver != SSE_MASS_CHOLESKY)
        return fail(SSE_ERR_BAD_ARGUMENT, "unknown mass_solver %d", cfg->mass_solver);
    // CholeskySolver with V = I is the DiagonalSolver (mass_matrix.jl:26-28)
    if (cfg->mass_solver == SSE_MASS_CHOLESKY && cfg->v_kind == SSE_V_IDENTITY) g.mass_solver = SSE_MASS_DIAGONAL;
    if (cfg->form == SSE_FORM_STANDARD_REFERENCE) {
        if (!a->Lambda_q) return fail(SSE_ERR_BAD_ARGUMENT, "Lambda_q required");
        for (int m = 0; m < d; m++) {
            if (!a->D[m]) return fail(SSE_ERR_BAD_ARGUMENT, "D[%d] required for StandardForm+ReferenceOperator", m);
            if ((rc = upload_sp(h, compress(a->D[m], Nq, Nq, false), &o.D[m]))) return rc;
            if ((rc = upload_sp(h, compress(a->D[m], Nq, Nq, true), &o.Dt[m]))) return rc;
        }
    } else if (cfg->form == SSE_FORM_FLUX_DIFFERENCING) {
        if (!a->Lambda_q) return fail(SSE_ERR_BAD_ARGUMENT, "Lambda_q required");
        std::vector<int> vptr(Nq + 1, 0), vj;
        std::vector<double> vS;
        for (int m = 0; m < d; m++)
            if (!a->S[m]) return fail(SSE_ERR_BAD_ARGUMENT, "S[%d] required for FluxDifferencingForm", m);
        // the reference only visits i < j (flux_differencing_form.jl:10-11, 52); row j receives -S[i,j]
        for (int i = 0; i < Nq; i++) {
            vptr[i] = (int)vj.size();
            for (int j = 0; j < Nq; j++) {
                if (i == j) continue;
                int lo = std::min(i, j), hi = std::max(i, j);
                bool any = false;
                for (int m = 0; m < d; m++) any |= a->S[m][lo + (size_t)Nq * hi] != 0.0;
                if (!any) continue;
                vj.push_back(j);
                for (int m = 0; m < d; m++) vS.push_back((i < j ? 1.0 : -1.0) * a->S[m][lo + (size_t)Nq * hi]);
            }
        }
        vptr[Nq] = (int)vj.size();
        if ((rc = upload(h, vptr, &o.vol_ptr))) return rc;
        if ((rc = upload(h, vj, &o.vol_j))) return rc;
        if ((rc = upload(h, vS, &o.vol_S))) return rc;
        // slot-major copies for the generic pair kernel (coalesced row reads)
        auto ell = [&](const std::vector<int>& ptr, const std::vector<int>& idx, const std::vector<double>& val, int rows, int nval,
                       int* width, const int** d_idx, const double** d_val) -> int32_t {
            int w = 0;
            for (int r = 0; r < rows; r++) w = std::max(w, ptr[(size_t)r + 1] - ptr[(size_t)r]);
            std::vector<int> ei((size_t)std::max(w, 1) * rows, -1);
            std::vector<double> ev((size_t)std::max(w, 1) * rows * nval, 0.0);
            // wide rows slot-major (coalesced across the threads of a warp), narrow rows row-major (one cache line per thread)
            const bool wide = w > 16;
            for (int r = 0; r < rows; r++)
                for (int q = ptr[(size_t)r]; q < ptr[(size_t)r + 1]; q++) {
                    const int s_ = q - ptr[(size_t)r];
                    ei[wide ? (size_t)s_ * rows + r : (size_t)r * w + s_] = idx[(size_t)q];
                    for (int m = 0; m < nval; m++)
                        ev[wide ? ((size_t)s_ * nval + m) * rows + r : ((size_t)r * w + s_) * nval + m] = val[(size_t)q * nval + m];
                }
            *width = wide ? w : -w;               // sign = layout
            int32_t rc_;
            if ((rc_ = upload(h, ei, d_idx))) return rc_;
            return upload(h, ev, d_val);
        };
        if ((rc = ell(vptr, vj, vS, Nq, d, &o.vol_w, &o.vol_je, &o.vol_Se))) return rc;
        o.has_C = a->Cfd != nullptr;
        if (o.has_C) {
            if (!a->nJq && !a->nref) return fail(SSE_ERR_BAD_ARGUMENT, "facet correction needs nJq or nref");
            const HostSp cq = compress(a->Cfd, Nq, Nf, false), cf = compress(a->Cfd, Nq, Nf, true);
            if ((rc = upload_sp(h, cq, &o.Cq))) return rc;
            if ((rc = upload_sp(h, cf, &o.Cf))) return rc;
            if ((rc = ell(cq.ptr, cq.idx, cq.val, Nq, 1, &o.cq_w, &o.cq_je, &o.cq_ve))) return rc;
            if ((rc = ell(cf.ptr, cf.idx, cf.val, Nf, 1, &o.cf_w, &o.cf_ie, &o.cf_ve))) return rc;
        }
        // dense operators (multidimensional schemes): all-pairs tables for k_time_fluxdiff_dense
        h->dense.ok = 0;
        if (cfg->pde == SSE_PDE_EULER && cfg->two_point_flux == SSE_TWO_POINT_ENTROPY_CONSERVATIVE && NC == d + 2 && d >= 2 && o.has_C &&
            std::abs(o.vol_w) > 16 && Nq <= 128 && Nf <= 128) {
            std::vector<double> S4((size_t)Nq * d * Nq, 0.0), C4((size_t)Nf * Nq, 0.0);
            for (int j = 0; j < Nq; j++)
                for (int i = 0; i < Nq; i++) {
                    if (i == j) continue;
                    const int lo = std::min(i, j), hi = std::max(i, j);
                    for (int m = 0; m < d; m++) S4[((size_t)j * d + m) * Nq + i] = 0.25 * (i < j ? 1.0 : -1.0) * a->S[m][lo + (size_t)Nq * hi];
                }
            for (int j = 0; j < Nf; j++)
                for (int i = 0; i < Nq; i++) C4[(size_t)j * Nq + i] = 0.25 * a->Cfd[i + (size_t)Nq * j];
            std::vector<double> RT((size_t)Nf * Nq);
            for (int j = 0; j < Nf; j++)
                for (int i = 0; i < Nq; i++) RT[(size_t)j * Nq + i] = a->R[j + (size_t)Nf * i];
            if ((rc = upload(h, S4, &h->dense.S4)) || (rc = upload(h, C4, &h->dense.C4)) || (rc = upload(h, RT, &h->dense.RT))) return rc;
            h->dense.ok = 1;
        }
    } else if (cfg->form == SSE_FORM_STANDARD_PHYSICAL) {
        if (!a->VOL || !a->FAC) return fail(SSE_ERR_BAD_ARGUMENT, "VOL and FAC required for PhysicalOperators");
        if ((rc = upload_raw(h, a->VOL, (size_t)Np * Nq * d * Ne, &g.VOL))) return rc;
        if ((rc = upload_raw(h, a->FAC, (size_t)Np * Nf * Ne, &g.FAC))) return rc;
    } else
        return fail(SSE_ERR_BAD_ARGUMENT, "unknown form %d", cfg->form);
    // ---- geometry
    if ((rc = upload_raw(h, a->J_q, (size_t)Nq * Ne, &g.J_q))) return rc;
    if ((rc = upload_raw(h, a->J_f, (size_t)Nf * Ne, &g.J_f))) return rc;
    if ((rc = upload_raw(h, a->nJf, (size_t)d * Nf * Ne, &g.nJf))) return rc;
    if (a->Lambda_q) { if ((rc = upload_raw(h, a->Lambda_q, (size_t)Nq * d * d * Ne, &g.Lambda_q))) return rc; }
    if (a->nJq && cfg->form == SSE_FORM_FLUX_DIFFERENCING) { if ((rc = upload_raw(h, a->nJq, (size_t)d * Nfac * Nq * Ne, &g.nJq))) return rc; }
    {
        const long long lim = g.NFT;
        for (size_t t = 0; t < (size_t)Nf * Ne; t++)
            if (a->mapP[t] < 1 || a->mapP[t] > lim) return fail(SSE_ERR_BAD_ARGUMENT, "mapP[%zu] = %lld out of range (BoundsError)", t, (long long)a->mapP[t]);
        collect_neighbours(a->mapP, Ne, Nf, Nfac, h->nbr_hi, h->nbr);
        h->nbr_ghost.assign((size_t)Ne, 0);
        if (cfg->N_ghost)
            for (long long k = 0; k < Ne; k++)
                for (int j = 0; j < Nf; j++)
                    if (a->mapP[(size_t)k * Nf + j] > (long long)Nf * Ne) { h->nbr_ghost[(size_t)k] = 1; break; }
        const long long* mp = nullptr;
        if ((rc = upload_raw(h, (const long long*)a->mapP, (size_t)Nf * Ne, &mp))) return rc;
        g.mapP = mp;
    }
    if (g.mass_solver == SSE_MASS_CHOLESKY) {
        // CholeskySolver(J_q, V, W) (mass_matrix.jl:30-39): one factorisation per element, on the device
        double* chol = nullptr;
        if ((rc = dalloc(h, (size_t)Np * Np * Ne, &chol))) return rc;
        int* bad = nullptr;
        CU(cudaMalloc((void**)&bad, sizeof(int)));
        h->owned.push_back(bad);
        CU(cudaMemset(bad, 0, sizeof(int)));
        const size_t smem = sizeof(double) * (size_t)(Np * Np + Np + Nq + warp_z_size(o, 1) + warp_w_size(o, 1));
        if (smem > 227 * 1024) return fail(SSE_ERR_UNSUPPORTED, "element mass matrix exceeds 227 KB of shared memory");
        CU(cudaFuncSetAttribute(k_cholesky_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        k_cholesky_factor<<<(unsigned)Ne, 128, smem, h->stream>>>(o, g, chol, bad);
        int hbad = 0;
        CU(cudaMemcpy(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost));
        CU(cudaGetLastError());
        if (hbad) return fail(SSE_ERR_BAD_ARGUMENT, "mass matrix V' WJ V is not positive definite (PosDefException, Solvers.jl:411-412)");
        g.chol = chol;
    }
    // ---- law
    Law& L = h->law;
    // the kernels see the viscous Burgers law as Burgers (two-point flux, wave speed: burgers.jl:45, 103-143) + the BR1 terms
    L.pde = cfg->pde == SSE_PDE_VISCOUS_BURGERS ? SSE_PDE_BURGERS : cfg->pde; L.two_point = cfg->two_point_flux; L.inviscid = cfg->inviscid_flux;
    L.half_lambda = cfg->half_lambda; L.b = cfg->b;
    L.viscous = (cfg->pde == SSE_PDE_ADVECTION_DIFFUSION || cfg->pde == SSE_PDE_VISCOUS_BURGERS);
    for (int m = 0; m < 3; m++) L.a[m] = cfg->a[m];
    L.gamma = cfg->gamma; L.gm1 = cfg->gamma - 1.0; L.igm1 = 1.0 / (cfg->gamma - 1.0); L.log_gm1 = std::log(cfg->gamma - 1.0);
    L.lmq[0] = -1.0 / 3.0; L.lmq[1] = -4.0 / 45.0; L.lmq[2] = -44.0 / 945.0; L.cc2 = 2.0 / (105.0 * (cfg->gamma - 1.0));
    h->second_order = (cfg->pde == SSE_PDE_ADVECTION_DIFFUSION || cfg->pde == SSE_PDE_VISCOUS_BURGERS);
    if (h->second_order && cfg->form != SSE_FORM_STANDARD_PHYSICAL)
        return fail(SSE_ERR_UNSUPPORTED, "second-order laws are only implemented with PhysicalOperators (Solvers.jl:357-376)");
    if (cfg->pde == SSE_PDE_EULER && NC != d + 2) return fail(SSE_ERR_BAD_ARGUMENT, "Euler needs N_c = d + 2");
    if (cfg->pde != SSE_PDE_EULER && NC != 1) return fail(SSE_ERR_BAD_ARGUMENT, "scalar law needs N_c = 1");
    // entropy-projection variant (flux_differencing_form.jl:171-292)
    h->project = 0;
    if (cfg->form == SSE_FORM_FLUX_DIFFERENCING && NC > 1) {
        if (cfg->v_kind == SSE_V_IDENTITY) h->project = o.has_C ? 1 : 0;
        else h->project = 2;
    }
    // ---- scratch
    if ((rc = dalloc(h, (size_t)Nq * NC * Ne, &h->u_q))) return rc;
    if ((rc = dalloc(h, (size_t)g.NFT * NC, &h->u_f))) return rc;
    if (h->second_order) {
        if ((rc = dalloc(h, (size_t)Nq * NC * d * Ne, &h->q_q))) return rc;
        if ((rc = dalloc(h, (size_t)g.NFT * NC * d, &h->q_f))) return rc;
    }
    CU(cudaMemsetAsync(h->u_f, 0, sizeof(double) * (size_t)g.NFT * NC, h->stream));
    // ---- tensor-line specialisation
    tensor_plan_build(h->tp, *cfg, *a, o);
    if (h->tp.ok) {
        if ((rc = tensor_plan_upload(h->tp, [&](const void* src, size_t bytes, const void** out) -> int32_t {
                 void* p = nullptr;
                 if (cudaMalloc(&p, std::max<size_t>(bytes, 8)) != cudaSuccess) return SSE_ERR_CUDA;
                 h->owned.push_back(p);
                 if (bytes && cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return SSE_ERR_CUDA;
                 *out = p;
                 return SSE_OK;
             })))
            return fail(rc, "uploading the tensor-line tables failed");
    }
    {
        int N = 0;
        if (ct_eligible(*cfg, *a, h->tp, &N) && ct_schedule_matches(h->tp, N) && ct_facet_factors(*cfg, *a, N, h->ct.facetR)) {
            h->ct.N = N;
            h->ct.A.assign(a->A, a->A + N * N);
            h->ct.B.assign(a->B, a->B + N * N * N);
            h->ct.dev.C = o.C; h->ct.dev.W = o.W; h->ct.dev.R = o.R; h->ct.dev.Rt = o.Rt; h->ct.dev.Ne = Ne;
            h->ct.dev.Bf = o.Bf;
            if ((rc = upload_c3(h, a, N))) return rc;
            if ((rc = upload(h, h->ct.facetR, &h->ct.dev.facR))) return rc;
            {   // power-of-two scalings of the pair weights (exact): see ec_finish_scaled in kernels_ct.cuh
                std::vector<double> vS(h->tp.v_S), fC(h->tp.f_C);
                for (double& x : vS) x *= 0.25;
                for (double& x : fC) x *= 0.125;
                if ((rc = upload(h, vS, &h->ct.dev.vS)) || (rc = upload(h, fC, &h->ct.dev.fC))) return rc;
            }
            for (int i = 0; i < 12; i++) h->ct.dev.nref[i] = (i < d * Nfac && a->nref) ? a->nref[i] : 0.0;
            h->ct.fR.assign((size_t)h->tp.dev.n_frounds * Nq, 0.0);
            for (int fr = 0; fr < h->tp.dev.n_frounds; fr++)
                for (int i = 0; i < Nq; i++) h->ct.fR[(size_t)fr * Nq + i] = a->R[h->tp.f_partner[(size_t)fr * Nq + i] + (size_t)Nf * i];
            if ((rc = upload(h, h->ct.fR, &h->ct.dev.fR))) return rc;
            if (const char* mb = getenv("SSE_FD_MINB")) h->ct.minb = atoi(mb);
            if (const char* mb = getenv("SSE_PROJ_MINB")) h->ct.proj_minb = atoi(mb);
            if (const char* mb = getenv("SSE_FD_DUAL")) h->ct.dual = atoi(mb);
            if (ct_set_attrs(N) != cudaSuccess) return fail(SSE_ERR_CUDA, "cudaFuncSetAttribute (compile-time kernels) failed");
            {   // the projection kernel copies W / J straight into its tile (cp.async): 1 kB per element, read instead of J_q
                double* ijw = nullptr;
                if ((rc = dalloc(h, (size_t)Nq * Ne, &ijw))) return rc;
                const long long n = (long long)Nq * Ne;
                k_ijw<<<(unsigned)std::min<long long>((n + 255) / 256, 8LL * h->sm_count), 256, 0, h->stream>>>(n, Nq, o.W, g.J_q, ijw);
                CU(cudaGetLastError());
                g.iJW = ijw;
            }
            h->ct.ok = 1;
        }
    }
    {
        int N = 0;
        std::vector<double> D1;
        if (!h->ct.ok && ct_eligible_standard(*cfg, *a, &N, D1, h->ct.fR) && ct_facet_factors(*cfg, *a, N, h->ct.facetR)) {
            h->ct.N = N; h->ct.kind = 1; h->ct.D1 = D1;
            h->ct.A.assign(a->A, a->A + N * N);
            h->ct.B.assign(a->B, a->B + N * N * N);
            h->ct.dev.C = o.C; h->ct.dev.W = o.W; h->ct.dev.R = o.R; h->ct.dev.Rt = o.Rt; h->ct.dev.Ne = Ne; h->ct.dev.Bf = o.Bf;
            if ((rc = upload_c3(h, a, N))) return rc;
            if ((rc = upload(h, h->ct.fR, &h->ct.dev.fR))) return rc;
            if (ct_set_attrs(N) != cudaSuccess) return fail(SSE_ERR_CUDA, "cudaFuncSetAttribute (compile-time kernels) failed");
            // fused two-kernel path (kernels_adv.cuh): derived coefficient tables in the kernels' own order; falls back to the
            // three-kernel path when the tables cannot be built (memory) or are switched off (SSE_ADV_FUSED=0)
            ct_adv_build(h->ct, g, h->law, o.W, o.Bf, Ne, g.NFT, h->stream, h->owned);
            cudaGetLastError();
            h->ct.ok = 1;
        }
    }
    {   // 2-D Euler flux differencing on collapsed triangles: warp-per-element kernels (kernels_tri.cuh)
        if (!h->ct.ok && tri_eligible(*cfg, *a, h->tp, h->ct)) {
            CtPlan& c = h->ct;
            c.kind = 2; c.sms = h->sm_count > 0 ? h->sm_count : 148;
            if ((rc = upload(h, c.triV, &c.tri.V)) || (rc = upload(h, c.trivS, &c.tri.vS)) || (rc = upload(h, c.trifC, &c.tri.fC)) ||
                (rc = upload(h, c.trifR, &c.tri.fR)) || (rc = upload(h, c.triRfac, &c.tri.rfac)))
                return rc;
            c.tri.W = o.W; c.tri.Bf = o.Bf; c.tri.A = o.A; c.tri.B = o.B;
            if (tri_set_attrs(c.N) != cudaSuccess) return fail(SSE_ERR_CUDA, "cudaFuncSetAttribute (triangle kernels) failed");
            double* ijw = nullptr;
            if ((rc = dalloc(h, (size_t)Nq * Ne, &ijw))) return rc;
            const long long n = (long long)Nq * Ne;
            k_ijw<<<(unsigned)std::min<long long>((n + 255) / 256, 8LL * std::max(h->sm_count, 1)), 256, 0, h->stream>>>(n, Nq, o.W, g.J_q, ijw);
            CU(cudaGetLastError());
            g.iJW = ijw;
            c.ok = 1;
        }
    }
    {   // 2-D linear advection, StandardForm + ReferenceOperators on collapsed triangles: warp-per-element kernels (kernels_tri.cuh)
        if (!h->ct.ok && tri_adv_eligible(*cfg, *a, h->ct)) {
            CtPlan& c = h->ct;
            c.kind = 3; c.sms = h->sm_count > 0 ? h->sm_count : 148;
            if ((rc = upload(h, c.triV, &c.tri.V)) || (rc = upload(h, c.trifR, &c.tri.fR)) || (rc = upload(h, c.triD1, &c.tri.D1)) ||
                (rc = upload(h, c.triRV, &c.tri.RV)))
                return rc;
            c.tri.W = o.W; c.tri.Bf = o.Bf;
            double* ijw = nullptr;
            if ((rc = dalloc(h, (size_t)Nq * Ne, &ijw))) return rc;
            const long long n = (long long)Nq * Ne;
            k_ijw<<<(unsigned)std::min<long long>((n + 255) / 256, 8LL * std::max(h->sm_count, 1)), 256, 0, h->stream>>>(n, Nq, o.W, g.J_q, ijw);
            CU(cudaGetLastError());
            g.iJW = ijw;
            c.ok = 1;
        }
    }
    // threads per element of the generic (one CTA per element) kernels: enough for one volume / facet node per thread, at most 128
    {   // threads per element of the packed kernels; SSE_THREADS_MIN=64 restores the round-1 minimum (A/B)
        const char* tm = getenv("SSE_THREADS_MIN");
        h->threads = std::min(128, std::max(tm ? atoi(tm) : 32, (std::max(Nq, Nf) + 31) / 32 * 32));
    }
    h->smem_nodal = smem_nodal_bytes(o);
    h->smem_time = smem_time_bytes(h);
    h->smem_dense = h->smem_time + sizeof(double) * (size_t)Nfac * d * Nq;
    if (h->dense.ok && (h->tp.ok || h->ct.ok || h->smem_dense > 227 * 1024)) h->dense.ok = 0;        // structured operators have better kernels
    if (h->dense.ok) {
        CU(cudaFuncSetAttribute(k_time_fluxdiff_dense<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN_MAX));
        CU(cudaFuncSetAttribute(k_time_fluxdiff_dense<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN_MAX));
    }
    h->smem_aux = smem_aux_bytes(o);
    if (std::max(h->smem_nodal, std::max(h->smem_time, h->smem_aux)) > 227 * 1024)
        return fail(SSE_ERR_UNSUPPORTED, "element tiles exceed 227 KB of shared memory");
#define SETA(D_, NC_) do { if ((rc = set_attrs<D_, NC_>(h))) return rc; } while (0)
    DISPATCH_DNC(h, SETA);
#undef SETA
    return SSE_OK;
}

extern "C" int32_t sse_create(const sse_config* cfg, const sse_arrays* arr, int32_t device, sse_handle** out) {
    if (!cfg || !arr || !out) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    *out = nullptr;
    if (cfg->abi_version != SSE_ABI_VERSION) return fail(SSE_ERR_BAD_ARGUMENT, "ABI version mismatch: header %d, library %d", cfg->abi_version, SSE_ABI_VERSION);
    if (cfg->d < 1 || cfg->d > 3 || cfg->N_e < 1 || cfg->N_fac < 1 || cfg->N_f % cfg->N_fac != 0)
        return fail(SSE_ERR_BAD_ARGUMENT, "bad sizes (d=%d, N_e=%lld, N_f=%d, N_fac=%d)", cfg->d, (long long)cfg->N_e, cfg->N_f, cfg->N_fac);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(SSE_ERR_CUDA, "no CUDA device available (%s); libsse_b200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(SSE_ERR_BAD_ARGUMENT, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    sse_handle* h = new sse_handle();
    h->cfg = *cfg;
    h->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    int32_t rc = build(h, cfg, arr);
    if (rc == SSE_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(SSE_ERR_CUDA, "device error during setup: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc != SSE_OK) {
        std::string keep = g_err;
        sse_destroy(h);
        g_err = keep;
        return rc;
    }
    *out = h;
    return SSE_OK;
}

extern "C" int32_t sse_destroy(sse_handle* h) {
    if (!h) return SSE_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    comm_release(h);
    if (h->graph.exec) cudaGraphExecDestroy(h->graph.exec);
    if (h->graph.stream) cudaStreamDestroy(h->graph.stream);
    if (h->graph.e_in) cudaEventDestroy(h->graph.e_in);
    if (h->graph.e_out) cudaEventDestroy(h->graph.e_out);
    for (void* p : h->owned) cudaFree(p);
    if (h->h_flag) cudaFreeHost((void*)h->h_flag);
    for (cudaEvent_t e : h->events) cudaEventDestroy(e);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    delete h;
    return SSE_OK;
}

extern "C" int32_t sse_set_stream(sse_handle* h, void* s) {
    if (!h) return fail(SSE_ERR_BAD_ARGUMENT, "null handle");
    h->stream = (cudaStream_t)s;
    return SSE_OK;
}
extern "C" int32_t sse_set_kernel_variant(sse_handle* h, int32_t v) {
    if (!h || v < 0 || v > 1) return fail(SSE_ERR_BAD_ARGUMENT, "bad kernel variant");
    h->variant = v;
    return SSE_OK;
}
extern "C" int32_t sse_get_kernel_variant(const sse_handle* h, int32_t* v) {
    if (!h || !v) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    *v = (h->variant == 1 && h->ct.ok) ? 2 : ((h->variant == 1 && h->dense.ok) ? 3 : ((h->variant == 1 && h->tp.ok && h->tp.has_fluxdiff) ? 1 : 0));
    return SSE_OK;
}

// Non-physical states: the reference raises a DomainError from log / sqrt.  The kernels that write dudt set a
// flag in mapped host memory when a value is not finite (log_nobranch returns NaN for a non-positive argument, so a
// negative density or pressure always ends up there); the blocking entry points report it once and clear it.
int32_t sse::check_flag(sse_handle* h) {
    if (h->h_flag && *h->h_flag) {
        *h->h_flag = 0;
        return fail(SSE_ERR_NONFINITE, "non-finite residual: the state left the physical domain (DomainError in the reference: log / sqrt of a "
                                       "negative density or pressure)");
    }
    return SSE_OK;
}

// ------------------------------------------------------------------------------ state vectors
static size_t state_len(const sse_handle* h) { return (size_t)h->cfg.N_p * h->cfg.N_c * h->cfg.N_e; }
extern "C" int32_t sse_state_alloc(sse_handle* h, double** out) {
    if (!h || !out) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaMalloc((void**)out, state_len(h) * sizeof(double)));
    CU(cudaMemsetAsync(*out, 0, state_len(h) * sizeof(double), h->stream));
    return SSE_OK;
}
extern "C" int32_t sse_state_free(sse_handle* h, double* p) {
    if (!h) return fail(SSE_ERR_BAD_ARGUMENT, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaFree(p));
    return SSE_OK;
}
extern "C" int32_t sse_state_fill(sse_handle* h, double* d_x, double value) {
    if (!h || !d_x) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    CU(cudaSetDevice(h->device));
    if (value == 0.0) { CU(cudaMemsetAsync(d_x, 0, state_len(h) * sizeof(double), h->stream)); return SSE_OK; }
    const long long n = (long long)state_len(h);
    k_fill<<<(unsigned)std::min<long long>((n + 255) / 256, 8 * h->sm_count), 256, 0, h->stream>>>(n, value, d_x);
    h->launches += 1;
    CU(cudaGetLastError());
    return SSE_OK;
}
extern "C" int32_t sse_state_upload(sse_handle* h, double* d_dst, const double* h_src) {
    if (!h || !d_dst || !h_src) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(d_dst, h_src, state_len(h) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return SSE_OK;
}
extern "C" int32_t sse_state_download(sse_handle* h, double* h_dst, const double* d_src) {
    if (!h || !h_dst || !d_src) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h_dst, d_src, state_len(h) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return check_flag(h);
}
extern "C" int32_t sse_synchronize(sse_handle* h) {
    if (!h) return fail(SSE_ERR_BAD_ARGUMENT, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return check_flag(h);
}

// ------------------------------------------------------------------------------ the hot path
// Element packing of the one-CTA-per-element kernels (common.cuh: sse_element / sse_row_smem / sse_sync): rows of T threads, one
// element per row, as many rows as fit 128 threads and 96 kB of shared memory; the remainder of the range goes to a second
// launch with one row per CTA.  SSE_PACK_ROWS=1 keeps one element per CTA (A/B).  Returns the number of launches.
static int pack_rows(int T, size_t smem_el) {
    static const int cap = [] { const char* e = getenv("SSE_PACK_ROWS"); return e ? std::max(1, atoi(e)) : 8; }();
    int r = std::max(1, std::min(cap, 128 / T));
    while (r > 1 && r * smem_el > 96 * 1024) r--;
    return r;
}
template <class F>
static int launch_rows(int T, size_t smem_el, long long first, long long count, F&& launch) {
    const int R = pack_rows(T, smem_el);
    const long long full = count / R, rest = count - full * R;
    if (full) launch((unsigned)full, dim3(T, R), R * smem_el, first);
    if (rest) launch((unsigned)rest, dim3(T, 1), smem_el, first + full * R);
    return (full ? 1 : 0) + (rest ? 1 : 0);
}
static bool use_tensor(const sse_handle* h) { return h->variant == 1 && h->tp.ok && h->tp.has_fluxdiff; }

extern "C" int32_t sse_rhs_pass_a_range(sse_handle* h, const double* d_u, int64_t first, int64_t count) {
    if (!h || !d_u) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    if (count <= 0) return SSE_OK;
    if (first < 0 || first + count > h->cfg.N_e) return fail(SSE_ERR_BAD_ARGUMENT, "element range out of bounds");
    CU(cudaSetDevice(h->device));
    int nl = 1;
    if (h->variant == 1 && h->ct.ok) {
        ct_nodal(h->ct, h->geo, h->law, first, count, d_u, h->u_q, h->u_f, h->stream);
    } else {
#define LA(D_, NC_)                                                                                                                  \
    nl = launch_rows(h->threads, h->smem_nodal, first, count, [&](unsigned grid, dim3 block, size_t smem, long long f0) {            \
        k_nodal_generic<D_, NC_><<<grid, block, smem, h->stream>>>(h->ops, h->geo, h->law, h->project, f0, d_u, h->u_q, h->u_f);     \
    })
        DISPATCH_DNC(h, LA);
#undef LA
    }
    h->launches += nl;
    CU(cudaGetLastError());
    return SSE_OK;
}
extern "C" int32_t sse_rhs_pass_a(sse_handle* h, const double* d_u) {
    if (!h) return fail(SSE_ERR_BAD_ARGUMENT, "null handle");
    return sse_rhs_pass_a_range(h, d_u, 0, h->cfg.N_e);
}

extern "C" int32_t sse_rhs_pass_aux(sse_handle* h, double* d_dudt, int64_t first, int64_t count) {
    (void)d_dudt;
    if (!h) return fail(SSE_ERR_BAD_ARGUMENT, "null handle");
    if (!h->second_order || count <= 0) return SSE_OK;
    if (first < 0 || first + count > h->cfg.N_e) return fail(SSE_ERR_BAD_ARGUMENT, "element range out of bounds");
    CU(cudaSetDevice(h->device));
    int nl = 1;
#define LA(D_, NC_)                                                                                                                  \
    nl = launch_rows(h->threads, h->smem_aux, first, count, [&](unsigned grid, dim3 block, size_t smem, long long f0) {              \
        k_aux_physical<D_, NC_><<<grid, block, smem, h->stream>>>(h->ops, h->geo, h->law, f0, h->u_q, h->u_f, h->q_q, h->q_f);       \
    })
    DISPATCH_DNC(h, LA);
#undef LA
    h->launches += nl;
    CU(cudaGetLastError());
    return SSE_OK;
}

extern "C" int32_t sse_rhs_pass_b(sse_handle* h, double* d_dudt, int64_t first, int64_t count) {
    return pass_b_stage(h, d_dudt, first, count, RkStage());
}
// rk.u != nullptr: the caller asks for the 2N-storage stage update to be fused; *fused reports whether it was
int32_t sse::pass_b_stage(sse_handle* h, double* d_dudt, int64_t first, int64_t count, RkStage rk, cudaEvent_t mid) {
    if (!h || !d_dudt) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    if (count <= 0) return SSE_OK;
    if (first < 0 || first + count > h->cfg.N_e) return fail(SSE_ERR_BAD_ARGUMENT, "element range out of bounds");
    CU(cudaSetDevice(h->device));
    const unsigned n = (unsigned)count;
    int nl = 1;
    if (h->cfg.form == SSE_FORM_FLUX_DIFFERENCING) {
        if (h->variant == 1 && h->ct.ok) {
            ct_fluxdiff(h->ct, h->tp, h->ops, h->geo, h->law, first, count, h->u_q, h->u_f, d_dudt, h->stream, rk, mid);
            h->launches += h->ct.kind == 2 ? 0 : 1;       // pair kernel + projection kernel (one fused kernel on triangles)
        } else if (h->variant == 1 && h->dense.ok) {
            const int T = std::min(128, (std::max(h->cfg.N_q, h->cfg.N_f) + 31) / 32 * 32);
            if (h->cfg.d == 2) k_time_fluxdiff_dense<2><<<n, T, h->smem_dense, h->stream>>>(h->ops, h->geo, h->law, h->dense, first, h->u_q, h->u_f, d_dudt);
            else k_time_fluxdiff_dense<3><<<n, T, h->smem_dense, h->stream>>>(h->ops, h->geo, h->law, h->dense, first, h->u_q, h->u_f, d_dudt);
        } else if (use_tensor(h) && h->tp.has_fluxdiff) {
#define LA(D_, NC_)                                                                                                                  \
    nl = launch_rows(h->tp.threads, h->tp.smem_fluxdiff, first, count, [&](unsigned grid, dim3 block, size_t smem, long long f0) {   \
        tensor_launch_fluxdiff<D_, NC_>(h->tp, h->ops, h->geo, h->law, grid, block, smem, f0, h->u_q, h->u_f, d_dudt, h->stream);    \
    })
            DISPATCH_DNC(h, LA);
#undef LA
        } else {
#define LA(D_, NC_)                                                                                                                  \
    nl = launch_rows(h->threads, h->smem_time, first, count, [&](unsigned grid, dim3 block, size_t smem, long long f0) {             \
        k_time_fluxdiff_generic<D_, NC_><<<grid, block, smem, h->stream>>>(h->ops, h->geo, h->law, f0, h->u_q, h->u_f, d_dudt);      \
    })
            DISPATCH_DNC(h, LA);
#undef LA
        }
    } else if (h->cfg.form == SSE_FORM_STANDARD_REFERENCE && h->variant == 1 && h->ct.ok && (h->ct.kind == 1 || h->ct.kind == 3)) {
        ct_standard(h->ct, h->geo, h->law, first, count, h->u_q, h->u_f, d_dudt, h->stream, rk, mid);
        h->launches += (h->ct.adv_ok || h->ct.kind == 3) ? 0 : 1;           // derivative kernel + projection kernel, or the fused kernel alone
    } else if (h->cfg.form == SSE_FORM_STANDARD_REFERENCE) {
#define LA(D_, NC_)                                                                                                                  \
    nl = launch_rows(h->threads, h->smem_time, first, count, [&](unsigned grid, dim3 block, size_t smem, long long f0) {             \
        k_time_standard_reference<D_, NC_><<<grid, block, smem, h->stream>>>(h->ops, h->geo, h->law, f0, h->u_q, h->u_f, d_dudt);    \
    })
        DISPATCH_DNC(h, LA);
#undef LA
    } else {
#define LA(D_, NC_)                                                                                                                  \
    nl = launch_rows(h->threads, h->smem_time, first, count, [&](unsigned grid, dim3 block, size_t smem, long long f0) {             \
        k_time_physical<D_, NC_><<<grid, block, smem, h->stream>>>(h->ops, h->geo, h->law, f0, h->second_order, h->u_q, h->u_f,      \
                                                                    h->q_q, h->q_f, d_dudt);                                          \
    })
        DISPATCH_DNC(h, LA);
#undef LA
    }
    h->launches += nl;
    CU(cudaGetLastError());
    return SSE_OK;
}

extern "C" int32_t sse_rhs(sse_handle* h, const double* d_u, double* d_dudt, double t) {
    (void)t;   // no method of the reference uses t (Solvers.jl:474-564)
    if (!h) return fail(SSE_ERR_BAD_ARGUMENT, "null handle");
    if (h->cfg.N_ghost != 0) return dist_rhs(h, d_u, d_dudt, RkStage());      // element-partitioned handle: comm.cu
    int32_t rc;
    if ((rc = sse_rhs_pass_a(h, d_u))) return rc;
    if ((rc = sse_rhs_pass_aux(h, d_dudt, 0, h->cfg.N_e))) return rc;
    return sse_rhs_pass_b(h, d_dudt, 0, h->cfg.N_e);
}

// semi_discrete_residual!(dudt::Array, u::Array, solver, t) with HOST arrays (Solvers.jl:474-564 as OrdinaryDiffEq calls it on
// CPU state): upload of u, pass A, pass B and download of dudt are pipelined over `chunks` contiguous element ranges on
// three streams.  Pass A of a range starts when its slice of u has arrived; pass B of a range starts once pass A has
// covered the range holding its highest face neighbour (mapP); its slice of dudt goes back while later uploads are still
// in flight (full-duplex PCIe).  Host buffers should be page-locked (sse_host_pin) or the copies serialise.
// Synchronous: dudt is complete on return.
extern "C" int32_t sse_rhs_host(sse_handle* h, const double* h_u, double* h_dudt, double t, int32_t chunks) {
    if (!h || !h_u || !h_dudt) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    CU(cudaSetDevice(h->device));
    int32_t rc;
    const long long ne = h->cfg.N_e;
    const size_t per = (size_t)h->cfg.N_p * h->cfg.N_c, total = per * (size_t)ne;
    if (!h->h2d_u) {
        if ((rc = dalloc(h, total, &h->h2d_u)) || (rc = dalloc(h, total, &h->d2h_du))) return rc;
        CU(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    }
    double *d_u = h->h2d_u, *d_du = h->d2h_du;
    if (h->cfg.N_ghost != 0 && !h->second_order && h->comm.planned) return dist_rhs_host(h, h_u, h_dudt, d_u, d_du, chunks);
    if (chunks <= 0) chunks = 48;        // measured on B200 + PCIe 5: 35.4 ms at 48 ranges against 39.8 (16) and 37.3 (64 and up) for 1 053 696 elements
    if (h->second_order || h->cfg.N_ghost || chunks == 1 || ne < 4 * (long long)chunks) {
        CU(cudaMemcpyAsync(d_u, h_u, total * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        if ((rc = sse_rhs(h, d_u, d_du, t))) return rc;
        CU(cudaMemcpyAsync(h_dudt, d_du, total * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return check_flag(h);
    }
    while (h->events.size() < (size_t)(2 * chunks + 2)) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->events.push_back(e);
    }
    std::vector<long long> bounds((size_t)chunks + 1);
    for (int c = 0; c <= chunks; c++) bounds[(size_t)c] = ne * c / chunks;
    if (h->plan_chunks != chunks) {
        make_range_plan(ne, chunks, h->cfg.N_fac, h->nbr, h->nbr_hi, h->plan_order, h->plan_ready);
        h->plan_chunks = chunks;
    }
    const std::vector<int>&order = h->plan_order, &ready = h->plan_ready;
    cudaEvent_t e_start = h->events[(size_t)(2 * chunks)], e_done = h->events[(size_t)(2 * chunks + 1)];
    CU(cudaEventRecord(e_start, h->stream));
    CU(cudaStreamWaitEvent(h->s_in, e_start, 0));
    CU(cudaStreamWaitEvent(h->s_out, e_start, 0));
    for (int i = 0; i < chunks; i++) {
        const long long a = bounds[(size_t)order[(size_t)i]], b = bounds[(size_t)order[(size_t)i] + 1];
        CU(cudaMemcpyAsync(d_u + per * (size_t)a, h_u + per * (size_t)a, per * (size_t)(b - a) * sizeof(double), cudaMemcpyHostToDevice, h->s_in));
        CU(cudaEventRecord(h->events[(size_t)i], h->s_in));
        CU(cudaStreamWaitEvent(h->stream, h->events[(size_t)i], 0));
        if ((rc = sse_rhs_pass_a_range(h, d_u, a, b - a))) return rc;
        for (int k = 0; k < chunks; k++) {
            if (ready[(size_t)k] != i) continue;
            const long long ka = bounds[(size_t)k], kb = bounds[(size_t)k + 1];
            if ((rc = sse_rhs_pass_b(h, d_du, ka, kb - ka))) return rc;
            CU(cudaEventRecord(h->events[(size_t)(chunks + k)], h->stream));
            CU(cudaStreamWaitEvent(h->s_out, h->events[(size_t)(chunks + k)], 0));
            CU(cudaMemcpyAsync(h_dudt + per * (size_t)ka, d_du + per * (size_t)ka, per * (size_t)(kb - ka) * sizeof(double), cudaMemcpyDeviceToHost, h->s_out));
        }
    }
    CU(cudaEventRecord(e_done, h->s_out));
    CU(cudaStreamWaitEvent(h->stream, e_done, 0));
    CU(cudaStreamSynchronize(h->stream));
    return check_flag(h);
}
// page-lock / release a host array for the asynchronous copies of sse_rhs_host (cudaHostRegister)
extern "C" int32_t sse_host_pin(void* p, int64_t bytes) {
    if (!p || bytes <= 0) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    CU(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault));
    return SSE_OK;
}
extern "C" int32_t sse_host_unpin(void* p) {
    if (!p) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    CU(cudaHostUnregister(p));
    return SSE_OK;
}

// ------------------------------------------------------------------------------ halo
extern "C" int32_t sse_halo_configure(sse_handle* h, const int64_t* send_index, int64_t n_send) {
    if (!h || (n_send > 0 && !send_index)) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    CU(cudaSetDevice(h->device));
    const long long owned = (long long)h->cfg.N_f * h->cfg.N_e;
    for (int64_t s = 0; s < n_send; s++)
        if (send_index[s] < 1 || send_index[s] > owned) return fail(SSE_ERR_BAD_ARGUMENT, "send_index[%lld] out of range", (long long)s);
    h->n_send = n_send;
    h->halo_vars = h->cfg.N_c * (h->second_order ? h->cfg.d : 1);
    const long long* p = nullptr;
    int32_t rc;
    if ((rc = upload_raw(h, (const long long*)send_index, (size_t)n_send, &p))) return rc;
    h->d_send_idx = (long long*)p;
    if ((rc = dalloc(h, (size_t)n_send * h->halo_vars, &h->d_send))) return rc;
    if ((rc = dalloc(h, (size_t)h->cfg.N_ghost * h->halo_vars, &h->d_recv))) return rc;
    return SSE_OK;
}
static int halo_nvar(const sse_handle* h, int which) { return which == 0 ? h->cfg.N_c : h->cfg.N_c * h->cfg.d; }
extern "C" int32_t sse_halo_pack(sse_handle* h, int32_t which) {
    if (!h || which < 0 || which > 1 || (which == 1 && !h->second_order)) return fail(SSE_ERR_BAD_ARGUMENT, "bad halo selector");
    if (h->n_send == 0) return SSE_OK;
    CU(cudaSetDevice(h->device));
    const int nv = halo_nvar(h, which);
    const long long n = h->n_send * nv;
    k_halo_pack<<<(unsigned)std::min<long long>((n + 255) / 256, 4096), 256, 0, h->stream>>>(h->n_send, nv, h->geo.NFT, h->d_send_idx, which ? h->q_f : h->u_f, h->d_send);
    h->launches += 1;
    CU(cudaGetLastError());
    return SSE_OK;
}
extern "C" int32_t sse_halo_unpack(sse_handle* h, int32_t which) {
    if (!h || which < 0 || which > 1 || (which == 1 && !h->second_order)) return fail(SSE_ERR_BAD_ARGUMENT, "bad halo selector");
    if (h->cfg.N_ghost == 0) return SSE_OK;
    CU(cudaSetDevice(h->device));
    const int nv = halo_nvar(h, which);
    const long long n = h->cfg.N_ghost * nv;
    k_halo_unpack<<<(unsigned)std::min<long long>((n + 255) / 256, 4096), 256, 0, h->stream>>>(h->cfg.N_ghost, nv, h->geo.NFT, (long long)h->cfg.N_f * h->cfg.N_e, h->d_recv, which ? h->q_f : h->u_f);
    h->launches += 1;
    CU(cudaGetLastError());
    return SSE_OK;
}
extern "C" int32_t sse_halo_send_buffer(sse_handle* h, double** d_buf, int64_t* n) {
    if (!h || !d_buf || !n) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    *d_buf = h->d_send; *n = h->n_send * h->halo_vars;
    return SSE_OK;
}
extern "C" int32_t sse_halo_recv_buffer(sse_handle* h, int32_t which, double** d_buf, int64_t* n) {
    if (!h || !d_buf || !n) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    *d_buf = h->d_recv; *n = h->cfg.N_ghost * halo_nvar(h, which);
    return SSE_OK;
}

// ------------------------------------------------------------------------------ callers either side
extern "C" int32_t sse_axpby(sse_handle* h, double a, const double* d_x, double b, double* d_y) {
    if (!h || !d_x || !d_y) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    CU(cudaSetDevice(h->device));
    const long long n = (long long)state_len(h);
    k_axpby<<<(unsigned)std::min<long long>((n + 255) / 256, 8 * h->sm_count), 256, 0, h->stream>>>(n, a, d_x, b, d_y);
    h->launches += 1;
    CU(cudaGetLastError());
    return SSE_OK;
}
extern "C" int32_t sse_lsrk_stage(sse_handle* h, double* d_u, double* d_tmp, const double* d_dudt, double A, double B, double dt) {
    if (!h || !d_u || !d_tmp || !d_dudt) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    CU(cudaSetDevice(h->device));
    const long long n = (long long)state_len(h);
    k_lsrk_stage<<<(unsigned)std::min<long long>((n + 255) / 256, 8 * h->sm_count), 256, 0, h->stream>>>(n, d_u, d_tmp, d_dudt, A, B, dt);
    h->launches += 1;
    CU(cudaGetLastError());
    return SSE_OK;
}
// Carpenter & Kennedy (1994) 2N-storage RK4(5) coefficients (OrdinaryDiffEq's CarpenterKennedy2N54)
static const double CK_A[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                               -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
static const double CK_B[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                               1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                               2277821191437.0 / 14882151754819.0};
static const double CK_C[5] = {0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
                               2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0};
// semi_discrete_residual! followed by one 2N-storage stage; on the compile-time path the stage update rides in the
// epilogue of the projection kernel (no extra launch, dudt is not re-read)
extern "C" int32_t sse_rhs_lsrk(sse_handle* h, double* d_u, double* d_tmp, double* d_dudt, double A, double B, double dt, double t) {
    (void)t;
    if (!h || !d_u || !d_tmp || !d_dudt) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    int32_t rc;
    const bool fused = h->variant == 1 && h->ct.ok;
    RkStage rk;
    if (fused) { rk.u = d_u; rk.tmp = d_tmp; rk.A = A; rk.B = B; rk.dt = dt; }
    if (h->cfg.N_ghost != 0) {
        if ((rc = dist_rhs(h, d_u, d_dudt, rk))) return rc;
    } else {
        if ((rc = sse_rhs_pass_a(h, d_u))) return rc;
        if ((rc = sse_rhs_pass_aux(h, d_dudt, 0, h->cfg.N_e))) return rc;
        if ((rc = pass_b_stage(h, d_dudt, 0, h->cfg.N_e, rk))) return rc;
    }
    if (!fused) return sse_lsrk_stage(h, d_u, d_tmp, d_dudt, A, B, dt);
    return SSE_OK;
}
static int32_t step_ck54_launches(sse_handle* h, double* d_u, double* d_tmp, double* d_dudt, double t, double dt);

// CUDA-graph replay of the step (sse_set_graph_mode): the 11 - 20 launches of one CarpenterKennedy2N54 step are captured once per
// (u, tmp, dudt, dt) and replayed with one cudaGraphLaunch -- on meshes of a few thousand elements the step is bound by launch
// latency, not by the kernels.  The capture runs on an internal stream (the caller's stream may be the legacy default stream,
// which cannot be captured); the replay is ordered after the caller's stream and the caller's stream after the replay.
static int32_t step_ck54_graph(sse_handle* h, double* d_u, double* d_tmp, double* d_dudt, double t, double dt) {
    CU(cudaSetDevice(h->device));
    sse_handle::Graph& G = h->graph;
    if (!G.stream) {
        CU(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&G.e_in, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&G.e_out, cudaEventDisableTiming));
    }
    if (!G.exec || G.u != d_u || G.tmp != d_tmp || G.dudt != d_dudt || G.dt != dt || G.variant != h->variant) {
        if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
        cudaStream_t user = h->stream;
        const long long l0 = h->launches;
        h->stream = G.stream;
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamBeginCapture(G.stream, cudaStreamCaptureModeThreadLocal);
        int32_t rc = e == cudaSuccess ? step_ck54_launches(h, d_u, d_tmp, d_dudt, t, dt) : SSE_ERR_CUDA;
        const cudaError_t e2 = cudaStreamEndCapture(G.stream, &graph);
        h->stream = user;
        G.launches = h->launches - l0;
        h->launches = l0;
        if (e != cudaSuccess || e2 != cudaSuccess || rc != SSE_OK || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            return rc != SSE_OK ? rc : fail(SSE_ERR_CUDA, "capturing the Runge-Kutta step failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        }
        e = cudaGraphInstantiate(&G.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { G.exec = nullptr; return fail(SSE_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
        G.u = d_u; G.tmp = d_tmp; G.dudt = d_dudt; G.dt = dt; G.variant = h->variant;
    }
    CU(cudaEventRecord(G.e_in, h->stream));
    CU(cudaStreamWaitEvent(G.stream, G.e_in, 0));
    CU(cudaGraphLaunch(G.exec, G.stream));
    CU(cudaEventRecord(G.e_out, G.stream));
    CU(cudaStreamWaitEvent(h->stream, G.e_out, 0));
    h->launches += G.launches;
    return SSE_OK;
}
extern "C" int32_t sse_set_graph_mode(sse_handle* h, int32_t on) {
    if (!h) return fail(SSE_ERR_BAD_ARGUMENT, "null handle");
    h->graph.on = on != 0;
    return SSE_OK;
}

extern "C" int32_t sse_step_ck54(sse_handle* h, double* d_u, double* d_tmp, double* d_dudt, double t, double dt) {
    if (!h || !d_u || !d_tmp || !d_dudt) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    if (h->graph.on && h->cfg.N_ghost == 0) return step_ck54_graph(h, d_u, d_tmp, d_dudt, t, dt);
    return step_ck54_launches(h, d_u, d_tmp, d_dudt, t, dt);
}
static int32_t step_ck54_launches(sse_handle* h, double* d_u, double* d_tmp, double* d_dudt, double t, double dt) {
    int32_t rc;
    static const bool fuse_stages = [] { const char* e = getenv("SSE_CK54_FUSED"); return !e || atoi(e) != 0; }();
    if (fuse_stages && h->variant == 1 && h->ct.ok && h->ct.kind == 0 && h->cfg.N_ghost == 0) {
        // compile-time Euler path on one GPU: 11 launches instead of 15 -- pass A once, then per stage the pair kernel and ONE
        // kernel that finishes the stage (projection, mass solve, 2N-storage update) and starts the next (entropy projection of
        // the updated state); the last stage ends with the plain projection kernel
        CU(cudaSetDevice(h->device));
        const long long ne = h->cfg.N_e;
        if ((rc = sse_rhs_pass_a(h, d_u))) return rc;
        for (int s = 0; s < 5; s++) {
            RkStage rk;
            rk.u = d_u; rk.tmp = d_tmp; rk.A = CK_A[s]; rk.B = CK_B[s]; rk.dt = dt;
            ct_pair(h->ct, h->geo, h->law, 0, ne, h->u_q, h->u_f, h->stream);
            if (s < 4) ct_project_nodal(h->ct, h->geo, h->law, 0, ne, h->u_q, h->u_f, d_dudt, h->stream, rk);
            else ct_project(h->ct, h->geo, 0, ne, h->u_q, d_dudt, h->stream, rk);
            h->launches += 2;
        }
        CU(cudaGetLastError());
        return SSE_OK;
    }
    for (int s = 0; s < 5; s++)
        if ((rc = sse_rhs_lsrk(h, d_u, d_tmp, d_dudt, CK_A[s], CK_B[s], dt, t + CK_C[s] * dt))) return rc;
    return SSE_OK;
}

// conservation / energy / entropy residual reductions (Analysis/conservation.jl:145-189)
//   out[e < NC] = sum_k 1' WJ_k V dudt_k[:, e]                      (:145-152)
//   out[NC]     = sum_k sum_e u_k[:, e]' M_k dudt_k[:, e]           (:154-167), M_k = mass_matrix(mass_solver, k):
//                 diag(W J_k) for the DiagonalSolver, V' diag(W J_k) V for the CholeskySolver; for the WeightAdjustedSolver M_k = (V' diag(W / J_k) V)^-1
//                 (mass_matrix.jl:140-153), applied by conjugate gradients on the SPD operator the residual itself
//                 applies (its condition number is max J / min J over the element, so a few iterations suffice)
//   out[NC + 1] = sum_k (V' WJ_k w(V u_k))' dudt_k                   (:169-189, M_k symmetric)
template <int D, int NC>
__global__ void k_functionals(Ops o, Geo g, Law L, const double* __restrict__ u, const double* __restrict__ dudt, double* __restrict__ out) {
    extern __shared__ double sm[];
    const int Nq = o.Nq, Np = o.Np;
    double* s_u = sm;                 // Np x NC
    double* s_d = s_u + Np * NC;      // Np x NC   dudt, then the CG residual
    double* s_y = s_d + Np * NC;      // Np x NC   CG iterate
    double* s_p = s_y + Np * NC;      // Np x NC   CG direction
    double* s_ap = s_p + Np * NC;     // Np x NC
    double* s_uq = s_ap + Np * NC;    // Nq x NC
    double* s_dq = s_uq + Nq * NC;    // Nq x NC
    double* s_z = s_dq + Nq * NC;
    double* s_w = s_z + warp_z_size(o, NC);
    __shared__ double acc[NC + 2], rr[NC], pap[NC], rr0[NC];
    __shared__ int go;
    if (threadIdx.x < NC + 2) acc[threadIdx.x] = 0.0;
    for (long long k = blockIdx.x; k < g.Ne; k += gridDim.x) {
        __syncthreads();
        SSE_FOR(t, Np * NC) { s_u[t] = u[(size_t)Np * NC * k + t]; s_d[t] = dudt[(size_t)Np * NC * k + t]; }
        __syncthreads();
        apply_V<NC>(o, s_u, s_uq, s_z, s_w);
        apply_V<NC>(o, s_d, s_dq, s_z, s_w);
        const double* J = g.J_q + (size_t)Nq * k;
        SSE_FOR(i, Nq) {
            const double wj = o.W[i] * J[i];
            double ui[NC], wi[NC];
#pragma unroll
            for (int e = 0; e < NC; e++) { ui[e] = s_uq[i + Nq * e]; atomicAdd(&acc[e], wj * s_dq[i + Nq * e]); }
            if (g.mass_solver != SSE_MASS_WEIGHT_ADJUSTED) {      // M = V' WJ V (V = I for the DiagonalSolver)
#pragma unroll
                for (int e = 0; e < NC; e++) atomicAdd(&acc[NC], wj * ui[e] * s_dq[i + Nq * e]);
            }
            if (L.pde == SSE_PDE_EULER) {
                cons_to_entropy<D, NC>(L, ui, wi);
                double s = 0.0;
#pragma unroll
                for (int e = 0; e < NC; e++) s += wi[e] * s_dq[i + Nq * e];
                atomicAdd(&acc[NC + 1], wj * s);
            }
        }
        if (g.mass_solver == SSE_MASS_WEIGHT_ADJUSTED) {
            __syncthreads();
            if (threadIdx.x < NC) rr[threadIdx.x] = 0.0;
            SSE_FOR(t, Np * NC) { s_y[t] = 0.0; s_p[t] = s_d[t]; }
            __syncthreads();
            SSE_FOR(t, Np * NC) atomicAdd(&rr[t / Np], s_d[t] * s_d[t]);
            __syncthreads();
            if (threadIdx.x < NC) rr0[threadIdx.x] = rr[threadIdx.x];
            for (int it = 0; it < 60; it++) {
                SSE_FOR(t, Np * NC) s_ap[t] = s_p[t];
                if (threadIdx.x < NC) pap[threadIdx.x] = 0.0;
                if (threadIdx.x == 0) go = 0;
                __syncthreads();
                mass_solve<NC>(o, g, k, s_ap, s_dq, s_z, s_w);                   // A p = V' diag(W/J) V p
                __syncthreads();
                SSE_FOR(t, Np * NC) atomicAdd(&pap[t / Np], s_p[t] * s_ap[t]);
                __syncthreads();
                SSE_FOR(t, Np * NC) {
                    const int e = t / Np;
                    const double al = pap[e] > 0.0 ? rr[e] / pap[e] : 0.0;
                    s_y[t] = fma(al, s_p[t], s_y[t]);
                    s_d[t] = fma(-al, s_ap[t], s_d[t]);
                }
                __syncthreads();
                if (threadIdx.x < NC) pap[threadIdx.x] = 0.0;                    // reused for the new <r, r>
                __syncthreads();
                SSE_FOR(t, Np * NC) atomicAdd(&pap[t / Np], s_d[t] * s_d[t]);
                __syncthreads();
                SSE_FOR(t, Np * NC) {
                    const int e = t / Np;
                    const double be = rr[e] > 0.0 ? pap[e] / rr[e] : 0.0;
                    s_p[t] = fma(be, s_p[t], s_d[t]);
                }
                if (threadIdx.x < NC && pap[threadIdx.x] > 1e-30 * rr0[threadIdx.x]) go = 1;
                __syncthreads();
                if (threadIdx.x < NC) rr[threadIdx.x] = pap[threadIdx.x];
                const int cont = go;
                __syncthreads();
                if (!cont) break;
            }
            SSE_FOR(t, Np * NC) atomicAdd(&acc[NC], s_u[t] * s_y[t]);
        }
    }
    __syncthreads();
    if (threadIdx.x < NC + 2) atomicAdd(&out[threadIdx.x], acc[threadIdx.x]);
}

extern "C" int32_t sse_functionals(sse_handle* h, const double* d_u, const double* d_dudt, double* out) {
    if (!h || !d_u || !d_dudt || !out) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    CU(cudaSetDevice(h->device));
    const int NC = h->cfg.N_c;
    double* d_out = nullptr;
    CU(cudaMalloc((void**)&d_out, sizeof(double) * (NC + 2)));
    CU(cudaMemsetAsync(d_out, 0, sizeof(double) * (NC + 2), h->stream));
    const Ops& o = h->ops;
    size_t smem = sizeof(double) * (size_t)(5 * o.Np * NC + 2 * o.Nq * NC + warp_z_size(o, NC) + warp_w_size(o, NC));
    unsigned grid = (unsigned)std::min<long long>(h->cfg.N_e, 4LL * h->sm_count);
#define LA(D_, NC_)                                                                                                  \
    do {                                                                                                             \
        /* the kernel also holds a few static shared scalars: the opt-in limit covers static + dynamic */          \
        cudaFuncAttributes fa;                                                                                       \
        e_l = cudaFuncGetAttributes(&fa, k_functionals<D_, NC_>);                                                    \
        if (e_l == cudaSuccess)                                                                                      \
            e_l = cudaFuncSetAttribute(k_functionals<D_, NC_>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                       SMEM_OPT_IN_MAX - (int)fa.sharedSizeBytes);                                   \
        if (e_l == cudaSuccess) {                                                                                    \
            k_functionals<D_, NC_><<<grid, 128, smem, h->stream>>>(h->ops, h->geo, h->law, d_u, d_dudt, d_out);       \
            e_l = cudaGetLastError();                                                                                \
        }                                                                                                            \
    } while (0)
    cudaError_t e_l = cudaSuccess;
    DISPATCH_DNC(h, LA);
#undef LA
    if (e_l != cudaSuccess) { cudaFree(d_out); return fail(SSE_ERR_CUDA, "functionals launch failed: %s", cudaGetErrorString(e_l)); }
    h->launches += 1;
    // element-partitioned handles: the functionals are sums over all ranks (every rank calls, every rank gets the total)
    const int32_t rc_red = dist_allreduce_sum(h, d_out, NC + 2);
    if (rc_red) { cudaFree(d_out); return rc_red; }
    cudaError_t e = cudaMemcpyAsync(out, d_out, sizeof(double) * (NC + 2), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(SSE_ERR_CUDA, "functionals failed: %s", cudaGetErrorString(e));
    return SSE_OK;
}

// GeometricFactors for a whole mesh (mesh.jl:229-506): host in, host out, element chunks staged through the device
extern "C" int32_t sse_geometric_factors(const sse_geom_config* c, const sse_geom_ops* op, int32_t device, const double* const xyz[3],
                                         double* J_q, double* Lambda_q, double* J_f, double* nJf) {
    if (!c || !op || !xyz || !J_q || !Lambda_q || !J_f || !nJf) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    const int d = c->d, Nm = c->N_map, N1 = c->N1, Nq = c->N_q, Nf = c->N_f;
    if (d < 1 || d > 3 || Nm < 2 || Nq < 1 || Nf < 1 || c->N_e < 0) return fail(SSE_ERR_BAD_ARGUMENT, "bad geometry sizes");
    const bool curl3 = c->metric == SSE_METRIC_CURL && d == 3;
    if (!op->Vq || !op->Vf || !op->nrstJ) return fail(SSE_ERR_BAD_ARGUMENT, "Vq, Vf and nrstJ are required");
    for (int m = 0; m < d; m++) if (!op->Drst[m] || !xyz[m]) return fail(SSE_ERR_BAD_ARGUMENT, "Drst / xyz missing");
    if (curl3) {
        if (!op->Vq1 || !op->Vf1 || !op->D1[0] || !op->D1[1] || !op->D1[2]) return fail(SSE_ERR_BAD_ARGUMENT, "3-D curl metrics need D1, Vq1, Vf1");
        if ((N1 != Nm) != (op->up != nullptr)) return fail(SSE_ERR_BAD_ARGUMENT, "up must be given exactly when N1 != N_map");
    } else if (N1 != Nm) return fail(SSE_ERR_BAD_ARGUMENT, "N1 != N_map only for the 3-D curl metrics");
    if (cudaSetDevice(device) != cudaSuccess) return fail(SSE_ERR_CUDA, "no usable CUDA device %d (libsse_b200 has no CPU fallback)", device);
    std::vector<void*> owned;
    auto cleanup = [&]() { for (void* p : owned) cudaFree(p); };
    auto up_ = [&](const double* src, size_t n, const double** out) -> bool {
        void* p = nullptr;
        if (cudaMalloc(&p, sizeof(double) * std::max<size_t>(n, 1)) != cudaSuccess) return false;
        owned.push_back(p);
        if (n && cudaMemcpy(p, src, sizeof(double) * n, cudaMemcpyHostToDevice) != cudaSuccess) return false;
        *out = (const double*)p;
        return true;
    };
    GeomDev g;
    memset(&g, 0, sizeof(g));
    g.d = d; g.Nmap = Nm; g.N1 = N1; g.Nq = Nq; g.Nf = Nf; g.metric = c->metric;
    bool ok = up_(op->Vq, (size_t)Nq * Nm, &g.Vq) && up_(op->Vf, (size_t)Nf * Nm, &g.Vf) && up_(op->nrstJ, (size_t)Nf * d, &g.nrstJ);
    for (int m = 0; m < d && ok; m++) ok = up_(op->Drst[m], (size_t)Nm * Nm, &g.Drst[m]);
    if (ok && curl3) {
        ok = up_(op->Vq1, (size_t)Nq * N1, &g.Vq1) && up_(op->Vf1, (size_t)Nf * N1, &g.Vf1);
        for (int m = 0; m < 3 && ok; m++) ok = up_(op->D1[m], (size_t)N1 * N1, &g.D1[m]);
        if (ok && op->up) ok = up_(op->up, (size_t)N1 * Nm, &g.up);
    }
    const long long CH = 65536;
    const long long nch = std::min<long long>(CH, std::max<long long>(c->N_e, 1));
    double *dx[3] = {nullptr, nullptr, nullptr}, *dJq = nullptr, *dL = nullptr, *dJf = nullptr, *dn = nullptr;
    auto dal = [&](size_t n, double** out) -> bool {
        void* p = nullptr;
        if (cudaMalloc(&p, sizeof(double) * std::max<size_t>(n, 1)) != cudaSuccess) return false;
        owned.push_back(p);
        *out = (double*)p;
        return true;
    };
    for (int m = 0; m < d && ok; m++) ok = dal((size_t)Nm * nch, &dx[m]);
    ok = ok && dal((size_t)Nq * nch, &dJq) && dal((size_t)Nq * d * d * nch, &dL) && dal((size_t)Nf * nch, &dJf) && dal((size_t)Nf * d * nch, &dn);
    if (!ok) { cleanup(); return fail(SSE_ERR_CUDA, "device allocation / upload failed in sse_geometric_factors"); }
    const int NM = std::max(N1, Nm);
    const size_t smem = sizeof(double) * (size_t)(d * NM + d * d * NM + 3 * NM + Nm);
    if (smem > 227 * 1024) { cleanup(); return fail(SSE_ERR_UNSUPPORTED, "mapping element too large for the geometry kernel"); }
    cudaFuncSetAttribute(k_geometry<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_geometry<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_geometry<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e = cudaSuccess;
    for (long long k0 = 0; k0 < c->N_e && e == cudaSuccess; k0 += nch) {
        const long long n = std::min<long long>(nch, c->N_e - k0);
        for (int m = 0; m < d && e == cudaSuccess; m++) {
            e = cudaMemcpy(dx[m], xyz[m] + (size_t)Nm * k0, sizeof(double) * (size_t)Nm * n, cudaMemcpyHostToDevice);
            g.xyz[m] = dx[m];
        }
        if (e != cudaSuccess) break;
        g.Ne = n; g.J_q = dJq; g.Lambda_q = dL; g.J_f = dJf; g.nJf = dn;
        if (d == 1) k_geometry<1><<<(unsigned)n, 128, smem>>>(g);
        else if (d == 2) k_geometry<2><<<(unsigned)n, 128, smem>>>(g);
        else k_geometry<3><<<(unsigned)n, 128, smem>>>(g);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if ((e = cudaMemcpy(J_q + (size_t)Nq * k0, dJq, sizeof(double) * (size_t)Nq * n, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if ((e = cudaMemcpy(Lambda_q + (size_t)Nq * d * d * k0, dL, sizeof(double) * (size_t)Nq * d * d * n, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if ((e = cudaMemcpy(J_f + (size_t)Nf * k0, dJf, sizeof(double) * (size_t)Nf * n, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        e = cudaMemcpy(nJf + (size_t)Nf * d * k0, dn, sizeof(double) * (size_t)Nf * d * n, cudaMemcpyDeviceToHost);
    }
    cleanup();
    if (e != cudaSuccess) return fail(SSE_ERR_CUDA, "sse_geometric_factors failed: %s", cudaGetErrorString(e));
    return SSE_OK;
}

// One residual with a CUDA event between every kernel, `reps` times: ms[0] pass A, ms[1] auxiliary pass (BR1), ms[2] first
// kernel of pass B (on the compile-time paths the pair / derivative kernel), ms[3] second kernel of pass B (the projection
// kernel; 0 where pass B is one kernel).  Averages over the repetitions; blocking.  bench.py's per-kernel roofline uses it.
extern "C" int32_t sse_profile_rhs(sse_handle* h, const double* d_u, double* d_dudt, int32_t reps, double* ms) {
    if (!h || !d_u || !d_dudt || !ms || reps < 1) return fail(SSE_ERR_BAD_ARGUMENT, "bad argument");
    if (h->cfg.N_ghost != 0) return fail(SSE_ERR_UNSUPPORTED, "sse_profile_rhs times the kernels of a single-GPU handle");
    CU(cudaSetDevice(h->device));
    cudaEvent_t e[5];
    for (int i = 0; i < 5; i++) CU(cudaEventCreate(&e[i]));
    for (int i = 0; i < 4; i++) ms[i] = 0.0;
    int32_t rc = SSE_OK;
    const bool two = h->variant == 1 && h->ct.ok;
    for (int r = 0; r < reps && rc == SSE_OK; r++) {
        cudaEventRecord(e[0], h->stream);
        if ((rc = sse_rhs_pass_a(h, d_u))) break;
        cudaEventRecord(e[1], h->stream);
        if ((rc = sse_rhs_pass_aux(h, d_dudt, 0, h->cfg.N_e))) break;
        cudaEventRecord(e[2], h->stream);
        if ((rc = pass_b_stage(h, d_dudt, 0, h->cfg.N_e, RkStage(), two ? e[3] : nullptr))) break;
        if (!two) cudaEventRecord(e[3], h->stream);
        cudaEventRecord(e[4], h->stream);
        if (cudaEventSynchronize(e[4]) != cudaSuccess) { rc = fail(SSE_ERR_CUDA, "profile run failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        for (int i = 0; i < 4; i++) { float t = 0; cudaEventElapsedTime(&t, e[i], e[i + 1]); ms[i] += t / reps; }
    }
    for (int i = 0; i < 5; i++) cudaEventDestroy(e[i]);
    return rc;
}
extern "C" int32_t sse_launch_count(const sse_handle* h, int64_t* n) {
    if (!h || !n) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    *n = h->launches;
    return SSE_OK;
}
extern "C" int32_t sse_debug_views(sse_handle* h, double** d_u_q, double** d_u_f) {
    if (!h) return fail(SSE_ERR_BAD_ARGUMENT, "null handle");
    if (d_u_q) *d_u_q = h->u_q;
    if (d_u_f) *d_u_f = h->u_f;
    return SSE_OK;
}

// Host-only diagnostic: builds the tensor-line schedule for (cfg, arr) without touching a device and
// replays it against S and C.  info[0..7] = {specialised?, threads/CTA, volume rounds, facet sub-rounds,
// reducer items max, reducer sources max, shared memory bytes, two-point fluxes per element}.
extern "C" int32_t sse_plan_selfcheck(const sse_config* cfg, const sse_arrays* arr, int32_t* info, double* max_err) {
    if (!cfg || !arr || !info || !max_err) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    Ops o;
    memset(&o, 0, sizeof(o));
    o.d = cfg->d; o.NC = cfg->N_c; o.Np = cfg->N_p; o.Nq = cfg->N_q; o.Nf = cfg->N_f; o.Nfac = cfg->N_fac;
    o.npf = cfg->N_f / cfg->N_fac; o.v_kind = cfg->v_kind;
    if (cfg->v_kind == SSE_V_WARPED) { o.P1 = cfg->p + 1; o.M1 = cfg->M1d[0]; o.M2 = cfg->M1d[1]; o.M3 = cfg->d == 3 ? cfg->M1d[2] : 1; }
    TensorPlan tp;
    tensor_plan_build(tp, *cfg, *arr, o);
    for (int i = 0; i < 8; i++) info[i] = 0;
    *max_err = 0.0;
    info[0] = tp.ok;
    {   // 2: compile-time flux-differencing kernels, 3: compile-time advection StandardForm kernels
        int N = 0;
        std::vector<double> D1, fR, fac;
        if (tp.ok && ct_eligible(*cfg, *arr, tp, &N) && ct_schedule_matches(tp, N) && ct_facet_factors(*cfg, *arr, N, fac)) info[0] = 2;
        else if (ct_eligible_standard(*cfg, *arr, &N, D1, fR) && ct_facet_factors(*cfg, *arr, N, fac)) { info[0] = 3; info[1] = 128; return SSE_OK; }
        else if (tp.ok) { CtPlan tri; if (tri_eligible(*cfg, *arr, tp, tri)) info[0] = 4; }      // 4: warp-per-element triangle kernels
        else { CtPlan tri; if (tri_adv_eligible(*cfg, *arr, tri)) { info[0] = 5; info[1] = 128; return SSE_OK; } }   // 5: ... for 2-D advection
    }
    if (!tp.ok) return SSE_OK;
    if (info[0] != 2 && !tp.has_fluxdiff) { info[0] = 0; return SSE_OK; }      // schedule exists, but no kernel for this size
    info[1] = tp.threads; info[2] = tp.dev.n_vrounds; info[3] = tp.dev.n_frounds;
    info[4] = tp.dev.red_items_max; info[5] = tp.dev.red_max; info[6] = (int32_t)tp.smem_fluxdiff;
    int evals = 0;
    for (int x : tp.v_partner) evals += x >= 0;
    for (int x : tp.f_partner) evals += x >= 0;
    info[7] = evals + cfg->N_f;
    *max_err = tensor_plan_selfcheck(tp, *cfg, *arr);
    return SSE_OK;
}

extern "C" int32_t sse_fp64_peak(int32_t device, double* flops) {
    if (!flops) return fail(SSE_ERR_BAD_ARGUMENT, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(SSE_ERR_CUDA, "no such CUDA device");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double* d = nullptr;
    CU(cudaMalloc((void**)&d, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    k_fp64_peak<<<blocks, threads>>>(d, 1 << 10);
    double best = 0.0;
    for (int r = 0; r < 5; r++) {
        CU(cudaEventRecord(e0));
        k_fp64_peak<<<blocks, threads>>>(d, iters);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        double f = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3);
        best = std::max(best, f);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *flops = best;
    return SSE_OK;
}
