// ct_api.h — host-visible interface of the compile-time-sized kernels (instantiated in ct_kernels.cu)
#pragma once
#include <vector>
#include "common.cuh"
#include "kernels_tensor.cuh"

namespace sse {

struct CtDev {
    const double* C;        // C[a3 + N*(b1 + N*(b2 + N*b3))]
    const double* C3;       // the same in the kernels' order: C3[l * N + a3], l the canonical modal index (tet_l)
    const double* W;        // volume quadrature weights
    SpMat R, Rt;
    long long Ne;
    // flux-differencing schedule weights (partners are closed-form in the kernel)
    const double* vS;       // [volume round][m][Nq]  skew-extended S_m[i, partner] / 4   (scaled pair flux, kernels_ct.cuh)
    const double* fC;       // [facet sub-round][Nq]  C[i, partner] / 8
    const double* fR;       // [facet sub-round][Nq]  R[partner, i]
    const double* Bf;       // [Nf]
    const double* facR;     // 1-D factors of R: r0, r1, r2, r3 (N each), I3[y + N b] (FacetR order); Euler path only
    double nref[12];        // d x N_fac reference normals
};

// optional 2N-storage Runge-Kutta stage fused into the projection epilogue: tmp = A tmp + dt dudt; u += B tmp
struct RkStage {
    double* u = nullptr;
    double* tmp = nullptr;
    double A = 0.0, B = 0.0, dt = 0.0;
};

// device tables of the fused 3-D advection path (kernels_adv.cuh), warp-interleaved: [task of 32/N elements][item][32 lanes]
struct AdvDev {
    const double* C = nullptr;      // [task][m][a1][a2][32]   c_m = sum_n (W Lambda_mn / 2) a_n at node (a1, a2, lane's a3)
    const double* iJW = nullptr;    // [task][a1][a2][32]      W / J
    const double* F = nullptr;      // [task][2][4 N][32]      fa = BJf (a.n)/2, fl = BJf halflambda |a.n| at the lane's facet nodes
    const int* map = nullptr;       // [task][4 N][32]         0-based index of the neighbour's facet node in u_f
    double* um = nullptr;           // [N_e][N_p]              modal coefficients handed from pass A to pass B
};

// device tables of the warp-per-element 2-D Euler path (kernels_tri.cuh); all of them element independent
struct TriDev {
    const double* A = nullptr;      // 1-D tensors of the warped V: A[a1 + N b1], B[a2 + N (b1 + N b2)]   (tensor_simplex.jl:84-140)
    const double* B = nullptr;
    const double* V = nullptr;      // dense warped V, column-major N_q x N_p (node a1 N + a2, canonical modal index)
    const double* vS = nullptr;     // [round][2][N_q]  skew-extended S_m[i, partner] / 4, round = line direction * (N/2) + shift - 1
    const double* fC = nullptr;     // [3][N_q]         C[i, partner] / 8
    const double* fR = nullptr;     // [3][N_q]         R[partner, i]
    const double* rfac = nullptr;   // [N_f][N]         the N non-zeros of row j of R, along the facet node's tensor line
    const double* W = nullptr;      // [N_q]
    const double* Bf = nullptr;     // [N_f]
    double nref[6] = {0, 0, 0, 0, 0, 0};   // 2 x 3 reference normals, column-major
    const double* D1 = nullptr;     // kind 3: the two 1-D derivative matrices, [m][row + N col]
    const double* RV = nullptr;     // kind 3: R V, column-major N_f x N_p (pass A is u_f = (R V) u)
};

struct CtPlan {
    int ok = 0, N = 0;
    int kind = 0;                   // 0: 3-D Euler flux differencing; 1: 3-D linear advection, StandardForm + ReferenceOperators;
                                    // 2: 2-D Euler flux differencing on triangles (kernels_tri.cuh)
                                    // 3: 2-D linear advection, StandardForm + ReferenceOperators on triangles (kernels_tri.cuh)
    TriDev tri;                     // kind 2
    std::vector<double> triV, trivS, trifC, trifR, triRfac;     // host images of the kind-2 tables
    std::vector<double> triD1, triRV;                           // ... and of the kind-3 tables (with triV, trifR)
    int sms = 148;                  // SMs of the device (grid of the persistent kind-2 / kind-3 kernels)
    std::vector<double> D1;         // kind 1: the three 1-D derivative matrices, [m][t + N*s]
    CtDev dev{};
    std::vector<double> A, B;       // host copies of the 1-D tensors handed to the kernels by value
    std::vector<double> fR;         // host image of dev.fR
    std::vector<double> facetR;     // 1-D factors of R: r0, r1, r2, r3 (N each), I3 (N x N)   (FacetR in kernels_ct.cuh)
    int minb = 4;                   // resident CTAs per SM requested for k_fluxdiff_ct (tuning knob)
    int dual = 1;                   // two pairs per thread and round in k_fluxdiff_ct (N = 5; SSE_FD_DUAL=0 disables)
    int proj_minb = 3;              // same for k_nodal_ct / k_project_ct
    AdvDev adv;                     // kind 1: tables of the fused two-kernel path; adv_ok = 0 -> the three-kernel path
    int adv_ok = 0;
};

bool ct_eligible(const sse_config& cfg, const sse_arrays& a, const TensorPlan& tp, int* Nout);
// advection + StandardForm + ReferenceOperators on ModalTensor tets; fills D1 and fR (host images) on success
bool ct_eligible_standard(const sse_config& cfg, const sse_arrays& a, int* Nout, std::vector<double>& D1, std::vector<double>& fR);
void ct_standard(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, double* u_q, const double* u_f,
                 double* dudt, cudaStream_t s, RkStage rk = RkStage(), cudaEvent_t mid = nullptr);
// kind 2: d = 2, Euler + EC two-point flux, flux differencing, warped V with M1 = M2 = p + 1 <= 5, weight-adjusted mass solver,
// R and the pair schedule in the closed form k_tri_* hard-code; fills the host images of the tables on success
bool tri_eligible(const sse_config& cfg, const sse_arrays& a, const TensorPlan& tp, CtPlan& p);
cudaError_t tri_set_attrs(int N);
// kind 3: d = 2, linear advection, StandardForm + ReferenceOperators, ModalTensor triangles with p + 1 <= 5, weight-adjusted mass solver
bool tri_adv_eligible(const sse_config& cfg, const sse_arrays& a, CtPlan& p);
bool ct_facet_factors(const sse_config& cfg, const sse_arrays& a, int N, std::vector<double>& out);
// true when the generic tables of tp equal the closed-form schedule k_fluxdiff_ct hard-codes
bool ct_schedule_matches(const TensorPlan& tp, int N);
cudaError_t ct_set_attrs(int N);
// builds the tables of the fused advection path on the device (allocations are appended to `owned`); false on failure
bool ct_adv_build(CtPlan& p, const Geo& g, const Law& L, const double* W, const double* Bf, long long Ne, long long NFT,
                  cudaStream_t s, std::vector<void*>& owned);
void ct_nodal(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, const double* u, double* u_q, double* u_f,
              cudaStream_t s);
void ct_fluxdiff(const CtPlan& p, const TensorPlan& tp, const Ops& o, const Geo& g, const Law& L, long long first, long long count,
                 double* u_q, const double* u_f, double* dudt, cudaStream_t s, RkStage rk = RkStage(), cudaEvent_t mid = nullptr);
// mid: recorded between the two kernels of pass B (sse_profile_rhs times them separately)
// pair kernel of pass B alone: r_q stays in the u_q scratch for ct_project_nodal / the projection kernel
void ct_pair(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, double* u_q, const double* u_f, cudaStream_t s);
// projection kernel alone (pass B-2): dudt = M^-1 V' r_q with the optional fused stage update
void ct_project(const CtPlan& p, const Geo& g, long long first, long long count, const double* r_q, double* dudt, cudaStream_t s, RkStage rk);
// pass B-2 of one Runge-Kutta stage fused with pass A of the next (k_nodal_ct<..., FUSED>): dudt = M^-1 V' r_q, the 2N-storage
// update of rk.u / rk.tmp, then u_q / u_f of the updated state
void ct_project_nodal(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, double* u_q, double* u_f,
                      double* dudt, cudaStream_t s, RkStage rk);

}  // namespace sse
