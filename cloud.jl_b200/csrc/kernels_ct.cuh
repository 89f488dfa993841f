// kernels_ct.cuh — compile-time-sized kernels for the headline path: 3-D Euler, flux differencing,
// ModalTensor(p) on collapsed tetrahedra with N = p+1 nodes per direction (N = 5 at p = 4).
//
//   k_nodal_ct    pass A  : entropy projection  (flux_differencing_form.jl:214-292)
//   k_fluxdiff_ct pass B-1: interface flux, volume flux differencing, facet correction, lift
//                           (flux_differencing_form.jl:294-342); leaves r_q in the u_q scratch, as the
//                           reference itself reuses u_q[:,:,k] (flux_differencing_form.jl:341-346)
//   k_project_ct  pass B-2: dudt = M^-1 V' r_q  (flux_differencing_form.jl:345-346, mass_matrix.jl:185-196)
//
// Sum factorisation (warped_product_3d.jl:47-136) is mapped as one thread per (element, variable, a3):
// the thread owns the N x N slab y[a1][a2] of its eta_3 index in registers, the A and B tensors arrive as
// kernel parameters (constant bank, free FMA operands after full unrolling), the C tensor sits in a shared
// table [modal index][a3], and the only cross-thread step of V' (the sum over a3) goes through shared memory
// in a bank-conflict-free layout.  V -> diag(W/J) -> V' of the weight-adjusted mass solve never leaves registers.
// A warp holds floor(32/N) groups of N lanes; a CTA of NC warps processes EPB = floor(32/N) elements.  Every tile
// between two pointwise (entropy-variable) phases belongs to the N lanes of one group, so those phases are
// ordered by __syncwarp and only the pointwise phases need CTA barriers.
#pragma once
#include "common.cuh"
#include "kernels_tensor.cuh"
#include "ct_api.h"

// independent entropy-variable transforms per thread and trip in k_nodal_ct: volume-node loops / facet-node loop.
// Measured at 82 944 elements once the maps were branch-free: (1,1) 0.643, (1,2) 0.639, (2,2) 0.646, (3,3) 0.682, (5,4) 0.78 ms
// (beyond two the idle slots of the last trip and the register pressure cost more than the interleaving hides)
#ifndef SSE_FD_FF_PAD
#define SSE_FD_FF_PAD 1
#endif
#ifndef SSE_FD_NREF_STATIC
#define SSE_FD_NREF_STATIC 1
#endif
#ifndef SSE_FD_RED3
#define SSE_FD_RED3 1
#endif
#ifndef SSE_PROJ_PREFETCH
#define SSE_PROJ_PREFETCH 150
#endif
#ifndef SSE_NODAL_ILP_Q
#define SSE_NODAL_ILP_Q 1
#endif
#ifndef SSE_NODAL_ILP_F
#define SSE_NODAL_ILP_F 1
#endif

// C-tensor symmetry in the applies of k_nodal_ct (15 instead of 35 table loads per application; measured with spills in round 1)
#ifndef SSE_NODAL_SYM
#define SSE_NODAL_SYM 0
#endif
#define SSE_NODAL_SYMB (SSE_NODAL_SYM != 0)

namespace sse {

// schedule-weight tables (vS, fC, fR: 34 x Nq doubles, the same for every element): kept in L1 against the streaming
// element data with an evict-last hint (pass B 1.976 -> 1.935 ms at 82 944 elements)
__device__ __forceinline__ double ld_tab(const double* p) {
    double v;
    asm("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

template <int N> struct SFCoef {       // A[a1 + N*b1], B[a2 + N*(b1 + N*b2)]  (reference column-major)
    double A[N * N];
    double B[N * N * N];
};

template <int N> struct Tet {
    static constexpr int Nq = N * N * N;
    static constexpr int Np = N * (N + 1) * (N + 2) / 6;
    static constexpr int npf = N * N;
    static constexpr int Nf = 4 * N * N;
    static constexpr int NT = ((Nq > Nf ? Nq : Nf) + 31) / 32 * 32;   // threads of the node-per-thread kernels (N = 3: N_f > N_q)
    static constexpr int GPW = 32 / N;          // groups (of N lanes) per warp
    static constexpr int EPB = GPW;             // elements per CTA in the projection kernels
    static constexpr int LPT = (Np + N - 1) / N;  // modal outputs per lane in the a3-reduction
    // V' partials of one (element, variable): partial (l, a3) sits at l * RED_SL + a3 * RED_SA.  For N = 5 the pair (5, 3)
    // with a group stride of 15 (mod 16) makes both the stores (lanes = (group, a3), fixed l) and the reducer loads
    // (lane a3 sums its LPT consecutive l) free of bank conflicts; the plain (N, 1) layout costs the loads a 2-way conflict
    static constexpr int RED_SL = N, RED_SA = (N == 5) ? 3 : 1;
    static constexpr int RED_SPAN = (Np - 1) * RED_SL + (N - 1) * RED_SA + 1;
};

// canonical modal ordering of warped_product (tensor_simplex.jl:113-131): i slowest, k fastest, i+j+k <= p
template <int N> __host__ __device__ constexpr int tet_l(int b1, int b2, int b3) {
    int l = 0;
    for (int i = 0; i < b1; i++) { int m = N - i; l += m * (m + 1) / 2; }
    for (int j = 0; j < b2; j++) l += N - b1 - j;
    return l + b3;
}

// y[a1][a2] (fixed a3) = sum A[a1,b1] B[a2,b1,b2] C[a3,b1,b2,b3] x[l(b1,b2,b3)]      warped_product_3d.jl:47-84
// c3[l * CS]: CS = 1 for a register array, CS = N for the shared table [l][a3] (pointer offset by a3)
// The collapsed-tet C tensor, 2 (1 - eta_3)^(i+j) P_k^(2i+2j+2,0)(eta_3) (tensor_simplex.jl:123-130), depends on (i + j, k) only
// (ct_eligible verifies C[a3,i,j,k] == C[a3,0,i+j,k] bit for bit), so a thread needs N (N + 1) / 2 values of it per
// application instead of N_p: they are read once into registers (c3_sym) ahead of the b1 slabs.
// SYM selects it per call site: it pays in k_project_ct (pass B 1.796 -> 1.782 ms at 82 944 elements) and loses in k_nodal_ct,
// whose transforms leave no registers for the N (N + 1) / 2 values (pass A 0.646 -> 0.707 ms with spills).
template <int N> __host__ __device__ constexpr int c3_sym_idx(int s, int k) { return s * N - s * (s - 1) / 2 + k; }   // k <= N - 1 - s
// C3G: c3 points into the global table (read through L1 with the evict-last hint) instead of a shared copy
template <int N, int CS, bool C3G = false>
__device__ __forceinline__ void load_c3_sym(const double* c3, double (&cs)[N * (N + 1) / 2]) {
#pragma unroll
    for (int sidx = 0; sidx < N; sidx++)
#pragma unroll
        for (int k = 0; k < N - sidx; k++) {
            if constexpr (C3G) cs[c3_sym_idx<N>(sidx, k)] = ld_tab(c3 + tet_l<N>(0, sidx, k) * CS);
            else cs[c3_sym_idx<N>(sidx, k)] = c3[tet_l<N>(0, sidx, k) * CS];
        }
}
template <int N, int CS = 1, bool SYM = false, bool C3G = false>
__device__ __forceinline__ void sf3_fwd(const SFCoef<N>& cf, const double* c3, const double* __restrict__ xs, double (&y)[N][N]) {
    static_assert(!C3G || SYM, "the global C table is read through its (i + j, k) symmetry");
    double cs[SYM ? N * (N + 1) / 2 : 1];
    if constexpr (SYM) load_c3_sym<N, CS, C3G>(c3, cs);
#pragma unroll
    for (int a1 = 0; a1 < N; a1++)
#pragma unroll
        for (int a2 = 0; a2 < N; a2++) y[a1][a2] = 0.0;
#pragma unroll
    for (int b1 = 0; b1 < N; b1++) {
        asm volatile("" ::: "memory");             // keep the loads of each b1-slab next to their use (register pressure)
        double w[N];
#pragma unroll
        for (int a2 = 0; a2 < N; a2++) w[a2] = 0.0;
#pragma unroll
        for (int b2 = 0; b2 < N - b1; b2++) {
            double z = 0.0;
#pragma unroll
            for (int b3 = 0; b3 < N - b1 - b2; b3++) {
                const int l = tet_l<N>(b1, b2, b3);
                if constexpr (SYM) z = fma(cs[c3_sym_idx<N>(b1 + b2, b3)], xs[l], z);
                else z = fma(c3[l * CS], xs[l], z);
            }
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) w[a2] = fma(cf.B[a2 + N * (b1 + N * b2)], z, w[a2]);
        }
#pragma unroll
        for (int a1 = 0; a1 < N; a1++)
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) y[a1][a2] = fma(cf.A[a1 + N * b1], w[a2], y[a1][a2]);
    }
}

// partial[l][a3] = C[a3,l] * sum_{a2} B[a2,b1,b2] sum_{a1} A[a1,b1] x[a1][a2]        warped_product_3d.jl:94-136
// written to red[l * RED_SL] (the caller passes red already offset by group and a3 * RED_SA)
template <int N, int CS = 1, bool SYM = false, bool C3G = false>
__device__ __forceinline__ void sf3_bwd_partials(const SFCoef<N>& cf, const double* c3, const double (&x)[N][N], double* __restrict__ red) {
    static_assert(!C3G || SYM, "the global C table is read through its (i + j, k) symmetry");
    double cs[SYM ? N * (N + 1) / 2 : 1];
    if constexpr (SYM) load_c3_sym<N, CS, C3G>(c3, cs);
#pragma unroll
    for (int b1 = 0; b1 < N; b1++) {
        asm volatile("" ::: "memory");
        double wt[N];
#pragma unroll
        for (int a2 = 0; a2 < N; a2++) {
            double s = 0.0;
#pragma unroll
            for (int a1 = 0; a1 < N; a1++) s = fma(cf.A[a1 + N * b1], x[a1][a2], s);
            wt[a2] = s;
        }
#pragma unroll
        for (int b2 = 0; b2 < N - b1; b2++) {
            double z = 0.0;
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) z = fma(cf.B[a2 + N * (b1 + N * b2)], wt[a2], z);
#pragma unroll
            for (int b3 = 0; b3 < N - b1 - b2; b3++) {
                const int l = tet_l<N>(b1, b2, b3);
                if constexpr (SYM) red[l * Tet<N>::RED_SL] = cs[c3_sym_idx<N>(b1 + b2, b3)] * z;
                else red[l * Tet<N>::RED_SL] = c3[l * CS] * z;
            }
        }
    }
}

// lane a3 of a group sums the partials of its LPT modal outputs; red is offset by group
template <int N>
__device__ __forceinline__ void sf3_bwd_reduce(const double* __restrict__ red, int a3, double (&out)[Tet<N>::LPT]) {
#pragma unroll
    for (int q = 0; q < Tet<N>::LPT; q++) {
        const int l = a3 * Tet<N>::LPT + q;
        double s = 0.0;
        if (l < Tet<N>::Np) {
#pragma unroll
            for (int a = 0; a < N; a++) s += red[l * Tet<N>::RED_SL + a * Tet<N>::RED_SA];
        }
        out[q] = s;
    }
}

// shared table s_c3[l * N + a3] = C[a3, b1, b2, b3] with l the canonical modal index: a straight copy of the table the host
// laid out in this order (every CTA starts with it, so it must not cost index arithmetic or dependent loads)
template <int N>
__device__ __forceinline__ void load_c3_shared(const CtDev& t, double* s_c3) {
    for (int i = threadIdx.x; i < Tet<N>::Np * N; i += blockDim.x) s_c3[i] = t.C3[i];
}
// 1-D factors of the facet extrapolation R on the collapsed tet (tensor_simplex.jl:265-268), extracted from Matrix(R)
// and verified entry by entry on the host (ct_facet_factors):
//   face 0 (eta_2 = -1), node (a1, a3): sum_a2 r0[a2] q[a1,a2,a3]      face 1 / 2 (eta_1 = +1 / -1), node (a2, a3): sum_a1 r1|r2[a1] q
//   face 3 (eta_3 = -1), node (a1, b):  sum_a2 I3[b + N a2] sum_a3 r3[a3] q[a1,a2,a3]
template <int N> struct FacetR { double r0[N], r1[N], r2[N], r3[N], I3[N * N]; };

// faces 0..2 from the slab y[a1][a2] of one eta_3 index: no shared-memory reads.  wf is the facet tile of (element, variable)
template <int N>
__device__ __forceinline__ void facet_rows_slab(const FacetR<N>& fr, const double (&y)[N][N], int a3, double* __restrict__ wf) {
    constexpr int NN = N * N;
#pragma unroll
    for (int a1 = 0; a1 < N; a1++) {
        double s = 0.0;
#pragma unroll
        for (int a2 = 0; a2 < N; a2++) s = fma(fr.r0[a2], y[a1][a2], s);
        wf[a1 * N + a3] = s;
    }
#pragma unroll
    for (int a2 = 0; a2 < N; a2++) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int a1 = 0; a1 < N; a1++) { s1 = fma(fr.r1[a1], y[a1][a2], s1); s2 = fma(fr.r2[a1], y[a1][a2], s2); }
        wf[NN + a2 * N + a3] = s1;
        wf[2 * NN + a2 * N + a3] = s2;
    }
}
// face 3, the N nodes (a1, .) of one a1, from the nodal tile q of (element, variable)
template <int N>
__device__ __forceinline__ void facet_rows_face3(const FacetR<N>& fr, const double* __restrict__ q, int a1, double* __restrict__ wf) {
    double tt[N];
#pragma unroll
    for (int a2 = 0; a2 < N; a2++) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < N; c++) s = fma(fr.r3[c], q[(a1 * N + a2) * N + c], s);
        tt[a2] = s;
    }
#pragma unroll
    for (int b = 0; b < N; b++) {
        double s = 0.0;
#pragma unroll
        for (int a2 = 0; a2 < N; a2++) s = fma(fr.I3[b + N * a2], tt[a2], s);
        wf[3 * N * N + a1 * N + b] = s;
    }
}

// shared-memory plan of the projection kernels (doubles).  FACETS: room for the facet tile wf of pass A, which aliases
// x and wij (both dead once the last V has been applied)
template <int N, int NC, bool FACETS = false> struct ProjSmem {
    using T = Tet<N>;
    static constexpr int WARPS = (NC == 1) ? 4 : NC;               // CTA = WARPS warps
    static constexpr int NG = WARPS * T::GPW;                      // groups (element, variable) per CTA
    static constexpr int EPB = NG / NC;                            // elements per CTA
    static_assert(NG % NC == 0, "groups must split evenly into elements");
    // group strides of the nodal tile and of the V' partials, padded to N (mod 16) doubles so that the N-lane groups of
    // a warp fall on disjoint bank ranges (an un-padded 125 / 175 makes neighbouring groups overlap: 2-way conflicts)
    static constexpr int QS = T::Nq + ((N - T::Nq % 16) % 16 + 16) % 16;
    static constexpr int RSM = (N == 5) ? 15 : N;                  // group stride of the partials modulo 16 (see Tet::RED_SA)
    static constexpr int RS = T::RED_SPAN + ((RSM - T::RED_SPAN % 16) % 16 + 16) % 16;
    static constexpr int WS = T::Nf + ((N - T::Nf % 16) % 16 + 16) % 16;   // group stride of the facet tile
    static constexpr int x = 0;                                    // [NG][Np]
    static constexpr int wij = x + NG * T::Np;                     // [EPB][QS]  W / J (systems only: the scalar kernels, 24
                                                                   //            elements per CTA, read W and J_q directly)
    static constexpr int wf = 0;                                   // [NG][WS]   facet tile (pass A), over x | wij
    static constexpr int lo_a = wij + (NC == 1 ? 0 : EPB * QS);
    static constexpr int lo_b = FACETS ? NG * WS : 0;
    static constexpr int big = lo_a > lo_b ? lo_a : lo_b;          // union: q [NG][Nq]  |  red [NG][Np][N]
    static constexpr int big_sz = (NG * QS > NG * RS) ? NG * QS : NG * RS;
    static constexpr int c3 = big + big_sz;                        // [Np][N]    C tensor, a3 fastest
    static constexpr int total = c3 + T::Np * N;
};

// ---------------------------------------------------------------------------------------------------------
// pass A — nodal_values! with the general (modal) entropy projection.
// FUSED (device-resident CarpenterKennedy2N54, sse_step_ck54): the kernel starts with pass B-2 of the PREVIOUS Runge-Kutta stage
// on the same elements -- dudt = M^-1 V' r_q (k_project_ct), the 2N-storage update u += B (tmp = A tmp + dt dudt) -- and carries
// the new modal coefficients straight into pass A of the next stage: no second launch, no re-read of u, one W / J tile.
template <int N, int NC, int MINB, bool PROJECT, bool FUSED = false>
__global__ void __launch_bounds__(ProjSmem<N, NC>::WARPS * 32, MINB)
k_nodal_ct(SFCoef<N> cf, FacetR<N> fr, CtDev t, Geo g, Law L, long long first, long long count, const double* __restrict__ u,
           double* __restrict__ u_q, double* __restrict__ u_f, double* __restrict__ dudt = nullptr, RkStage rk = RkStage()) {
    constexpr int D = 3;
    static_assert(!FUSED || PROJECT, "the fused stage kernel is the Euler kernel");
    static_assert(!PROJECT || NC == D + 2, "the entropy projection of this kernel is written for the Euler equations");
    using T = Tet<N>;
    using S = ProjSmem<N, NC, true>;
    constexpr int Nq = T::Nq, Np = T::Np, Nf = T::Nf, EPB = S::EPB;
    extern __shared__ double sm[];
    double* s_x = sm + S::x;
    double* s_q = sm + S::big;
    double* s_red = sm + S::big;
    double* s_wij = sm + S::wij;
    double* s_wf = sm + S::wf;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int gl = lane / N, a3 = lane - gl * N;
    const int grp = warp * T::GPW + gl;                 // (element slot, variable) = (grp / NC, grp % NC)
    const long long e0 = first + (long long)blockIdx.x * EPB;   // first element of this CTA
    const int nel = (int)((first + count - e0 < EPB) ? (first + count - e0) : EPB);
    const bool act = gl < T::GPW && (grp / NC) < nel;

    double y[N][N];
    const double* c3 = sm + S::c3 + a3;
    if constexpr (FUSED) {
        // ---- pass B-2 of the previous stage (the body of k_project_ct): r_q sits in the u_q scratch of these elements
        for (int i = tid; i < Np * N; i += NT) cp_async8(sm + S::c3 + i, t.C3 + i);
        for (int it = tid; it < nel * Nq; it += NT) {
            const int el = it / Nq, i = it - el * Nq;
            cp_async8(s_wij + el * S::QS + i, g.iJW + (size_t)(e0 + el) * Nq + i);
        }
        cp_async_commit();
        // J_q is needed much later (entropy-variable phase): pull its lines into L2 now
        for (int i = tid; i < (nel * Nq * 8 + 127) / 128; i += NT)
            asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)(g.J_q + (size_t)e0 * Nq) + (size_t)i * 128));
        if (act) {
            const double* src = u_q + (size_t)e0 * NC * Nq + (size_t)grp * Nq + a3;
#pragma unroll
            for (int a1 = 0; a1 < N; a1++)
#pragma unroll
                for (int a2 = 0; a2 < N; a2++) y[a1][a2] = src[(a1 * N + a2) * N];
        }
        const double* c3g = t.C3 + a3;
        double out[T::LPT];
        if (act) sf3_bwd_partials<N, N, true, true>(cf, c3g, y, s_red + grp * S::RS + a3 * T::RED_SA);
        __syncwarp();
        if (act) {
            sf3_bwd_reduce<N>(s_red + grp * S::RS, a3, out);
#pragma unroll
            for (int q = 0; q < T::LPT; q++) { const int l = a3 * T::LPT + q; if (l < Np) s_x[grp * Np + l] = out[q]; }
        }
        __syncwarp();
        if (act) sf3_fwd<N, N, true, true>(cf, c3g, s_x + grp * Np, y);
        cp_async_wait<0>();
        __syncthreads();                                   // W / J tile and the shared C table (filled by all warps)
        if (act) {
            const double* wij = s_wij + (grp / NC) * S::QS;
#pragma unroll
            for (int a1 = 0; a1 < N; a1++)
#pragma unroll
                for (int a2 = 0; a2 < N; a2++) y[a1][a2] *= wij[(a1 * N + a2) * N + a3];
            sf3_bwd_partials<N, N, true, true>(cf, c3g, y, s_red + grp * S::RS + a3 * T::RED_SA);
        }
        __syncwarp();
        if (act) {
            sf3_bwd_reduce<N>(s_red + grp * S::RS, a3, out);
#pragma unroll
            for (int q = 0; q < T::LPT; q++) {
                const int l = a3 * T::LPT + q;
                if (l < Np) {
                    const size_t idx = (size_t)e0 * NC * Np + grp * Np + l;
                    dudt[idx] = out[q];
                    flag_nonfinite(g.flag, out[q]);
                    const double tm = fma(rk.A, rk.tmp[idx], rk.dt * out[q]);      // 2N-storage stage (Carpenter & Kennedy 1994)
                    rk.tmp[idx] = tm;
                    const double un = fma(rk.B, tm, rk.u[idx]);
                    rk.u[idx] = un;
                    s_x[grp * Np + l] = un;                // ... and straight into pass A of the next stage
                }
            }
        }
        __syncthreads();                                   // the nodal tiles written next lie over the partials of other groups
    } else {
    // every global load of the prologue as an asynchronous copy: the C table and u (needed by the first V) form the first
    // group, J_q (needed by the entropy-variable phase only) the second one, which stays in flight behind the first V
    for (int i = tid; i < Np * N; i += NT) cp_async8(sm + S::c3 + i, t.C3 + i);
    for (int i = tid; i < nel * NC * Np; i += NT) cp_async8(s_x + i, u + (size_t)e0 * NC * Np + i);
    cp_async_commit();
    if constexpr (PROJECT) {
        for (int it = tid; it < nel * Nq; it += NT) {
            const int el = it / Nq, i = it - el * Nq;
            cp_async8(s_wij + el * S::QS + i, g.J_q + (size_t)(e0 + el) * Nq + i);
        }
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    }

    // u_q = V u
    if (act) sf3_fwd<N, N, SSE_NODAL_SYMB>(cf, c3, s_x + grp * Np, y);
    if constexpr (PROJECT) {
    if (act) {
#pragma unroll
        for (int a1 = 0; a1 < N; a1++)
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) s_q[grp * S::QS + (a1 * N + a2) * N + a3] = y[a1][a2];
    }
    cp_async_wait<0>();
    __syncthreads();
    // w_q = WJ * w(u_q)                                      flux_differencing_form.jl:230-235
    // UT independent nodes per thread and trip: the branch-free maps of physics.cuh make the trip one basic block, so
    // the ~75-deep dependent FP64 chains of the nodes interleave (the kernel runs 15 warps per SM)
    {
        constexpr int UT = SSE_NODAL_ILP_Q;
        const int total = nel * Nq;
        for (int it0 = tid; it0 < total; it0 += UT * NT) {
            double ui[UT][NC], wi[UT][NC], J[UT], W[UT];
            int sq[UT], sw[UT];
            bool ok[UT];
#pragma unroll
            for (int k = 0; k < UT; k++) {
                int it = it0 + k * NT;
                ok[k] = it < total;
                if (!ok[k]) it = it0;                          // idle slot: recompute a valid node, store nothing
                const int el = it / Nq, i = it - el * Nq;
                sq[k] = el * NC * S::QS + i;
                sw[k] = el * S::QS + i;
                if constexpr (FUSED) J[k] = g.J_q[(size_t)(e0 + el) * Nq + i];       // the tile already holds W / J
                else J[k] = s_wij[sw[k]];
                W[k] = t.W[i];
#pragma unroll
                for (int e = 0; e < NC; e++) ui[k][e] = s_q[sq[k] + e * S::QS];
            }
#pragma unroll
            for (int k = 0; k < UT; k++) euler_cons_to_entropy_nb<D>(L.gamma, L.gm1, L.igm1, ui[k], wi[k]);
#pragma unroll
            for (int k = 0; k < UT; k++) {
                if (ok[k]) {
                    const double wj = W[k] * J[k];
                    if constexpr (!FUSED) s_wij[sw[k]] = W[k] * rcp_fast(J[k]);
#pragma unroll
                    for (int e = 0; e < NC; e++) s_q[sq[k] + e * S::QS] = wi[k][e] * wj;
                }
            }
        }
    }
    __syncthreads();
    // From here to the last V every tile a thread touches (its group's partials, modal coefficients and nodal slab)
    // belongs to the N lanes of its own group, which sit in one warp: __syncwarp orders them, not a CTA barrier.
    // w = V' w_q
    if (act) {
#pragma unroll
        for (int a1 = 0; a1 < N; a1++)
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) y[a1][a2] = s_q[grp * S::QS + (a1 * N + a2) * N + a3];
    }
    __syncthreads();                                   // the partials of a group lie over the nodal tiles of other groups
    double out[T::LPT];
    if (act) sf3_bwd_partials<N, N, SSE_NODAL_SYMB>(cf, c3, y, s_red + grp * S::RS + a3 * T::RED_SA);
    __syncwarp();
    if (act) {
        sf3_bwd_reduce<N>(s_red + grp * S::RS, a3, out);
#pragma unroll
        for (int q = 0; q < T::LPT; q++) { const int l = a3 * T::LPT + q; if (l < Np) s_x[grp * Np + l] = out[q]; }
    }
    __syncwarp();
    // w = M \ w : V, diag(W/J), V'                           mass_matrix.jl:185-196
    if (act) {
        sf3_fwd<N, N, SSE_NODAL_SYMB>(cf, c3, s_x + grp * Np, y);
        const double* wij = s_wij + (grp / NC) * S::QS;
#pragma unroll
        for (int a1 = 0; a1 < N; a1++)
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) y[a1][a2] *= wij[(a1 * N + a2) * N + a3];
        sf3_bwd_partials<N, N, SSE_NODAL_SYMB>(cf, c3, y, s_red + grp * S::RS + a3 * T::RED_SA);
    }
    __syncwarp();
    if (act) {
        sf3_bwd_reduce<N>(s_red + grp * S::RS, a3, out);
#pragma unroll
        for (int q = 0; q < T::LPT; q++) { const int l = a3 * T::LPT + q; if (l < Np) s_x[grp * Np + l] = out[q]; }
    }
    __syncwarp();
    // w_q = V w
    if (act) sf3_fwd<N, N, SSE_NODAL_SYMB>(cf, c3, s_x + grp * Np, y);
    }
    // Every warp is past its partials, modal coefficients and W / J: the nodal tile (over the partials of other groups) and
    // the facet tile (over x | wij) may be written.  Facet values R w_q (R u_q for scalar laws) come from the 1-D factors
    // of R: faces 0..2 from the slab still in registers, face 3 from the group's own nodal tile.
    __syncthreads();
    if (act) {
#pragma unroll
        for (int a1 = 0; a1 < N; a1++)
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) s_q[grp * S::QS + (a1 * N + a2) * N + a3] = y[a1][a2];
        facet_rows_slab<N>(fr, y, a3, s_wf + grp * S::WS);
    }
    __syncwarp();
    if (act) facet_rows_face3<N>(fr, s_q + grp * S::QS, a3, s_wf + grp * S::WS);
    __syncthreads();
    // u_q = u(w_q), u_f = u(R w_q)                           flux_differencing_form.jl:240-249
    if constexpr (PROJECT && NC == D + 2) {
        constexpr int UT = SSE_NODAL_ILP_Q, UF = SSE_NODAL_ILP_F;
        // volume nodes
        const int totq = nel * Nq;
        for (int it0 = tid; it0 < totq; it0 += UT * NT) {
            double wi[UT][NC], ui[UT][NC];
            size_t dst[UT];
            bool ok[UT];
#pragma unroll
            for (int k = 0; k < UT; k++) {
                int it = it0 + k * NT;
                ok[k] = it < totq;
                if (!ok[k]) it = it0;
                const int el = it / Nq, i = it - el * Nq;
                dst[k] = (size_t)(e0 + el) * NC * Nq + i;
#pragma unroll
                for (int e = 0; e < NC; e++) wi[k][e] = s_q[(el * NC + e) * S::QS + i];
            }
#pragma unroll
            for (int k = 0; k < UT; k++) euler_entropy_to_cons_nb<D>(L.gamma, L.gm1, L.igm1, L.log_gm1, wi[k], ui[k]);
#pragma unroll
            for (int k = 0; k < UT; k++) {
                if (ok[k]) {
#pragma unroll
                    for (int e = 0; e < NC; e++) u_q[dst[k] + (size_t)e * Nq] = ui[k][e];
                }
            }
        }
        // facet nodes
        const int totf = nel * Nf;
        for (int it0 = tid; it0 < totf; it0 += UF * NT) {
            double wi[UF][NC], ui[UF][NC];
            size_t dst[UF];
            bool ok[UF];
#pragma unroll
            for (int k = 0; k < UF; k++) {
                int it = it0 + k * NT;
                ok[k] = it < totf;
                if (!ok[k]) it = it0;
                const int el = it / Nf, j = it - el * Nf;
                dst[k] = (size_t)(e0 + el) * Nf + j;
#pragma unroll
                for (int e = 0; e < NC; e++) wi[k][e] = s_wf[(el * NC + e) * S::WS + j];
            }
#pragma unroll
            for (int k = 0; k < UF; k++) euler_entropy_to_cons_nb<D>(L.gamma, L.gm1, L.igm1, L.log_gm1, wi[k], ui[k]);
#pragma unroll
            for (int k = 0; k < UF; k++) {
                if (ok[k]) {
#pragma unroll
                    for (int e = 0; e < NC; e++) u_f[dst[k] + (size_t)g.NFT * e] = ui[k][e];
                }
            }
        }
    } else {
        for (int it = tid; it < nel * (Nq + Nf); it += NT) {
            const int el = it / (Nq + Nf), i = it - el * (Nq + Nf);
            if (i < Nq) {
#pragma unroll
                for (int e = 0; e < NC; e++) u_q[((size_t)(e0 + el) * NC + e) * Nq + i] = s_q[(el * NC + e) * S::QS + i];
            } else {
                const int j = i - Nq;
#pragma unroll
                for (int e = 0; e < NC; e++) u_f[(size_t)(e0 + el) * Nf + j + (size_t)g.NFT * e] = s_wf[(el * NC + e) * S::WS + j];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// pass B-2 — dudt = M^-1 V' r_q     (r_q sits in the u_q scratch)
template <int N, int NC, int MINB>
__global__ void __launch_bounds__(ProjSmem<N, NC>::WARPS * 32, MINB)
k_project_ct(SFCoef<N> cf, CtDev t, Geo g, long long first, long long count, const double* __restrict__ r_q, double* __restrict__ dudt,
             RkStage rk) {
    using T = Tet<N>;
    using S = ProjSmem<N, NC>;
    constexpr int Nq = T::Nq, Np = T::Np, EPB = S::EPB;
    extern __shared__ double sm[];
    double* s_x = sm + S::x;
    double* s_q = sm + S::big;
    double* s_red = sm + S::big;
    double* s_wij = sm + S::wij;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int gl = lane / N, a3 = lane - gl * N;
    const int grp = warp * T::GPW + gl;
    const long long e0 = first + (long long)blockIdx.x * EPB;
    const long long rem = first + count - e0;
    const int nel = (int)(rem < EPB ? rem : EPB);
    const bool act = gl < T::GPW && (grp / NC) < nel;

    double y[N][N], out[T::LPT];
    // r_q comes from DRAM (the pair kernel wrote 5 kB per element since): pull the tiles of the CTA that starts SSE_PROJ_PREFETCH
    // CTAs later into L2, so that its slab loads and its W / J copy pay an L2 round trip instead of a DRAM one
    if constexpr (SSE_PROJ_PREFETCH > 0 && NC > 1) {
        const long long ep = e0 + (long long)SSE_PROJ_PREFETCH * EPB;
        if (ep + EPB <= first + count) {
            const char* pr = (const char*)(r_q + (size_t)ep * NC * Nq);
            for (int i = tid; i < (EPB * NC * Nq * 8 + 127) / 128; i += NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + (size_t)i * 128));
            const char* pw = (const char*)(g.iJW + (size_t)ep * Nq);
            for (int i = tid; i < (EPB * Nq * 8 + 127) / 128; i += NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(pw + (size_t)i * 128));
        }
    }
    // Prologue without a barrier: the W / J tile arrives by asynchronous copies (precomputed at sse_create, so no reciprocal
    // sits between the load and the tile) and is waited for where it is first used; the C tensor is read from its global
    // table through L1 (15 values per application by the (i + j, k) symmetry) instead of being staged in shared memory.
    if constexpr (NC > 1) {
        for (int it = tid; it < nel * Nq; it += NT) {
            const int el = it / Nq, i = it - el * Nq;
            cp_async8(s_wij + el * S::QS + i, g.iJW + (size_t)(e0 + el) * Nq + i);
        }
        cp_async_commit();
    }
    if (act) {
        const double* src = r_q + (size_t)e0 * NC * Nq + (size_t)grp * Nq + a3;
#pragma unroll
        for (int a1 = 0; a1 < N; a1++)
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) y[a1][a2] = src[(a1 * N + a2) * N];
    }
    const double* c3 = t.C3 + a3;
    if (act) sf3_bwd_partials<N, N, true, true>(cf, c3, y, s_red + grp * S::RS + a3 * T::RED_SA);
    __syncwarp();
    if (act) {
        sf3_bwd_reduce<N>(s_red + grp * S::RS, a3, out);
#pragma unroll
        for (int q = 0; q < T::LPT; q++) { const int l = a3 * T::LPT + q; if (l < Np) s_x[grp * Np + l] = out[q]; }
    }
    __syncwarp();
    if (act) sf3_fwd<N, N, true, true>(cf, c3, s_x + grp * Np, y);
    if constexpr (NC > 1) {
        cp_async_wait<0>();
        __syncthreads();                                   // the tile of an element is filled by threads of several warps
    }
    if (act) {
        if constexpr (NC > 1) {
            const double* wij = s_wij + (grp / NC) * S::QS;
#pragma unroll
            for (int a1 = 0; a1 < N; a1++)
#pragma unroll
                for (int a2 = 0; a2 < N; a2++) y[a1][a2] *= wij[(a1 * N + a2) * N + a3];
        } else {
            const double* Jq = g.J_q + (size_t)(e0 + grp) * Nq + a3;
#pragma unroll
            for (int a1 = 0; a1 < N; a1++)
#pragma unroll
                for (int a2 = 0; a2 < N; a2++) y[a1][a2] *= t.W[(a1 * N + a2) * N + a3] * rcp_fast(Jq[(a1 * N + a2) * N]);
        }
        sf3_bwd_partials<N, N, true, true>(cf, c3, y, s_red + grp * S::RS + a3 * T::RED_SA);
    }
    __syncwarp();
    if (act) {
        sf3_bwd_reduce<N>(s_red + grp * S::RS, a3, out);
#pragma unroll
        for (int q = 0; q < T::LPT; q++) {
            const int l = a3 * T::LPT + q;
            if (l < Np) {
                const size_t idx = (size_t)e0 * NC * Np + grp * Np + l;
                dudt[idx] = out[q];
                flag_nonfinite(g.flag, out[q]);
                if (rk.u) {                        // fused 2N-storage RK stage (Carpenter & Kennedy 1994)
                    const double tm = fma(rk.A, rk.tmp[idx], rk.dt * out[q]);
                    rk.tmp[idx] = tm;
                    rk.u[idx] = fma(rk.B, tm, rk.u[idx]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// pass B-1 — one CTA per element, one thread per volume node (same schedule tables as k_fluxdiff_tensor)
template <int N, bool DUAL = false> struct FdSmem {
    using T = Tet<N>;
    static constexpr int NP = 6, NC = 5, D = 3;
    static constexpr int prim = 0;                         // [NP][Nq]
    static constexpr int lam = prim + NP * T::Nq;          // [D*D][Nq]
    static constexpr int fprim = lam + D * D * T::Nq;      // [NP][Nf]
    static constexpr int hnf = fprim + NP * T::Nf;         // [D][Nf]
    // rows of the facet residual FFS apart with FFS = N^2 (mod 16): reducer t = (variable, facet node of one face) then
    // updates bank t (mod 16) -- 2 instead of 3 wavefronts per access of its read-modify-write and of the lift
    static constexpr int FFS = SSE_FD_FF_PAD ? T::Nf + ((N * N - T::Nf) % 16 + 16) % 16 : T::Nf;
    static constexpr int ff = hnf + D * T::Nf;             // [NC][FFS]
    static constexpr int stage = ff + NC * FFS;            // [2][NC][Nq]
    // DUAL: in the facet sub-rounds a staged vector of volume node (a, b, c) sits at a * PS + b * N + c with the planes
    // padded to PS = N (mod 16) and the variables FS = N * PS apart, which puts the (facet node, variable) reducer of
    // thread t on bank t (mod 16) for every face: no conflicts on its N loads (they cost 3.5 - 4 wavefronts each unpadded).
    // The staging stores pay for it (3 instead of 2 wavefronts); the layout with the fewest wavefronts in total (unpadded
    // planes, FS = N^3 + 12: 320 + 560 instead of 480 + 480 per element, tools/bank_sim.py) is 1.3 % SLOWER -- the reducer
    // loads sit between two barriers, the stores do not.
    static constexpr int PS = N * N + ((N - (N * N) % 16) % 16 + 16) % 16;
    static constexpr int FS = N * PS;
    static constexpr int buf = NC * T::Nq;                 // doubles per stage buffer of the volume rounds
    static constexpr int total = stage + (DUAL ? 4 : 2) * buf;
    static_assert(!DUAL || 2 * NC * FS <= 4 * buf, "the two padded facet stages reuse the four volume stages");
};

// closed-form facet partner of volume node (a,b,c) in facet sub-round fr (the rotation schedule of
// tensor_plan_build; ct_schedule_matches() verifies it against the generic tables on the host)
template <int N>
__device__ __forceinline__ int facet_partner(int fr, int ca, int cb, int cc) {
    constexpr int NN = N * N;
    if (fr == 0) return ca * N + cc;
    if (fr == 1) return NN + cb * N + cc;
    if (fr == 2) return 2 * NN + cb * N + cc;
    int bp = (fr - 3) - cc;
    if (bp < 0) bp += N;
    return 3 * NN + ca * N + bp;
}

// 2 halfnJq[n, f, i] = sum_l Lambda[i, l, n] nref[l, f] (mesh.jl:262-269) for the reference normals of the collapsed tet,
// (0,-1,0), (1,1,1), (-1,0,0), (0,0,-1) (ct_eligible checks them): a row of Lambda with its sign flipped, or the sum of the
// three rows in the order of the reference's loop -- the same bits as the three multiply-adds with 0 / +-1, none of them
// issued.  f is a compile-time constant once the facet loop is unrolled.
__device__ __forceinline__ double tet_face_h(int f, const double (&lam)[3][3], int n) {
    return f == 0 ? -lam[1][n] : (f == 1 ? (lam[0][n] + lam[1][n]) + lam[2][n] : (f == 2 ? -lam[0][n] : -lam[2][n]));
}

template <int N, int MINB, bool DUAL>
__global__ void __launch_bounds__(Tet<N>::NT, MINB)
k_fluxdiff_ct(CtDev t, Geo g, Law L, long long first, double* __restrict__ u_q, const double* __restrict__ u_f) {
    constexpr int NC = 5, D = 3, NP = 6;
    using T = Tet<N>;
    using S = FdSmem<N, DUAL>;
    static_assert(!DUAL || (N % 2 == 1 && N / 2 == 2), "the two-pairs-per-round variant is written for N = 5");
    constexpr int Nq = T::Nq, Nf = T::Nf, NN = N * N, NSH = N / 2, NVR = D * NSH, NFR = 3 + N;
    extern __shared__ double sm[];
    double* s_prim = sm + S::prim;
    double* s_lam = sm + S::lam;
    double* s_fprim = sm + S::fprim;
    double* s_hnf = sm + S::hnf;
    double* s_ff = sm + S::ff;
    double* s_stage = sm + S::stage;
    const int tid = threadIdx.x;
    const long long k = first + blockIdx.x;
    const bool node = tid < Nq;
    const int ca = tid / NN, cb = (tid / N) % N, cc = tid % N;

    // ---- prologue: every global load of the element is issued before any arithmetic.  Idle lanes load a clamped
    //      (valid) address so that the loads need no branch: node data, own facet data and the mapP-dependent
    //      neighbour gather are all in flight together instead of one dependent round trip after another.
    static_assert(Nf <= T::NT && Nq <= T::NT, "one volume node and one facet node per thread");
    const bool fac = tid < Nf;
    const int tn = node ? tid : Nq - 1, tj = fac ? tid : Nf - 1;
    double qi[NP], lam[D][D], r[NC], sw[D];
    const size_t jo = (size_t)(g.mapP[(size_t)k * Nf + tj] - 1);
    double un[NC], ui[NC], uo[NC], njf[D];
#pragma unroll
    for (int e = 0; e < NC; e++) un[e] = __ldcs(u_q + ((size_t)k * NC + e) * Nq + tn);
#pragma unroll
    for (int e = 0; e < NC; e++) ui[e] = __ldcs(u_f + (size_t)k * Nf + tj + (size_t)g.NFT * e);
    const double jf = __ldcs(g.J_f + (size_t)k * Nf + tj);
#pragma unroll
    for (int m = 0; m < D; m++) njf[m] = __ldcs(g.nJf + m + D * ((size_t)k * Nf + tj));
#pragma unroll
    for (int n = 0; n < D; n++)
#pragma unroll
        for (int m = 0; m < D; m++) lam[m][n] = __ldcs(g.Lambda_q + ((size_t)k * D * D + (m + D * n)) * Nq + tn);
#pragma unroll
    for (int e = 0; e < NC; e++) uo[e] = __ldcs(u_f + jo + (size_t)g.NFT * e);
#pragma unroll
    for (int m = 0; m < D; m++) sw[m] = ld_tab(t.vS + ((0 * D + m) * Nq + tn));            // weights of round 0
    const double bf = t.Bf[tj];
#pragma unroll
    for (int e = 0; e < NC; e++) r[e] = 0.0;
    to_prim_fast<D>(L, un, qi);
    if (node) {
#pragma unroll
        for (int c = 0; c < NP; c++) s_prim[c * Nq + tid] = qi[c];
#pragma unroll
        for (int n = 0; n < D; n++)
#pragma unroll
            for (int m = 0; m < D; m++) s_lam[(m + D * n) * Nq + tid] = lam[m][n];
    }
    {
        double qa[NP], qb[NP], nf[D], nfq[D], phi[NC];
        const double ijf = rcp_fast(jf);
#pragma unroll
        for (int m = 0; m < D; m++) { nf[m] = njf[m] * ijf; nfq[m] = 0.25 * nf[m]; }     // n_f = nJf / J_f   operators.jl:59
        const double ira = to_prim_fast<D>(L, ui, qa);
        const double irb = to_prim_fast<D>(L, uo, qb);
        ec_contract_scaled<D>(L, qa, qb, nfq, phi);            // F#(u-, u+) . n    ConservationLaws.jl:75-128
        if (L.inviscid == SSE_FLUX_LAX_FRIEDRICHS) {
            double vni = 0.0, vno = 0.0;
#pragma unroll
            for (int m = 0; m < D; m++) { vni = fma(qa[1 + m], nf[m], vni); vno = fma(qb[1 + m], nf[m], vno); }
            // max(c_i, c_o) = sqrt(max(c_i^2, c_o^2)) bit for bit (sqrt is monotone and correctly rounded): one square root
            const double cm = sqrt(fmax(L.gamma * (0.5 * qa[D + 1]) * ira, L.gamma * (0.5 * qb[D + 1]) * irb));
            const double a = L.half_lambda * (fmax(fabs(vni), fabs(vno)) + cm);
#pragma unroll
            for (int e = 0; e < NC; e++) phi[e] = fma(a, ui[e] - uo[e], phi[e]);
        }
        const double bj = bf * jf;                             // BJf               operators.jl:58
        if (fac) {
#pragma unroll
            for (int m = 0; m < D; m++) s_hnf[m * Nf + tid] = njf[m];       // 2 halfnJf (operators.jl:78); the 1/2 lives in fC
#pragma unroll
            for (int c = 0; c < NP; c++) s_fprim[c * Nf + tid] = qa[c];
#pragma unroll
            for (int e = 0; e < NC; e++) s_ff[e * S::FFS + tid] = bj * phi[e];
        }
    }
    __syncthreads();

    // ---- two pairs per thread and round: both shifts of a line direction / two facet sub-rounds are evaluated
    //      back to back so their dependent FP64 chains interleave (ILP), with one barrier per two pairs
    if constexpr (DUAL) {
        int buf = 0;
        double swb[D];
#pragma unroll
        for (int m = 0; m < D; m++) swb[m] = ld_tab(t.vS + ((1 * D + m) * Nq + tn));
#pragma unroll               // the three line directions as straight-line code: static strides and m >= l tests, no weight
                             // hand-over moves (pass B 1.936 -> 1.820 ms at 82 944 elements; unrolling the facet loop as well loses)
        for (int l = 0; l < D; l++, buf ^= 1) {
            double* stA = s_stage + (2 * buf) * S::buf;
            double* stB = stA + S::buf;
            const int cl = (l == 0) ? ca : ((l == 1) ? cb : cc);
            const int stride = (l == 0) ? NN : ((l == 1) ? N : 1);
            int c1 = cl + 1; if (c1 >= N) c1 -= N;
            int c2 = cl + 2; if (c2 >= N) c2 -= N;
            const int jA = tid + (c1 - cl) * stride, jB = tid + (c2 - cl) * stride;
            double swnA[D], swnB[D];
#pragma unroll
            for (int m = 0; m < D; m++) {
                swnA[m] = (l + 1 < D) ? ld_tab(t.vS + (((2 * l + 2) * D + m) * Nq + tn)) : 0.0;
                swnB[m] = (l + 1 < D) ? ld_tab(t.vS + (((2 * l + 3) * D + m) * Nq + tn)) : 0.0;
            }
            if (node) {
                double gA[D], gB[D], qA[NP], qB[NP], pA[NC], pB[NC];
#pragma unroll
                for (int n = 0; n < D; n++) { gA[n] = 0.0; gB[n] = 0.0; }
#pragma unroll
                for (int m = 0; m < D; m++) {
                    if (m >= l) {
#pragma unroll
                        for (int n = 0; n < D; n++) {
                            gA[n] = fma(sw[m], lam[m][n] + s_lam[(m + D * n) * Nq + jA], gA[n]);
                            gB[n] = fma(swb[m], lam[m][n] + s_lam[(m + D * n) * Nq + jB], gB[n]);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < NP; c++) { qA[c] = s_prim[c * Nq + jA]; qB[c] = s_prim[c * Nq + jB]; }
                ec_contract_scaled2<D>(L, qi, qA, qB, gA, gB, pA, pB);
#pragma unroll
                for (int e = 0; e < NC; e++) { r[e] -= pA[e] + pB[e]; stA[e * Nq + jA] = pA[e]; stB[e * Nq + jB] = pB[e]; }
            }
            __syncthreads();
            if (node) {
#pragma unroll
                for (int e = 0; e < NC; e++) r[e] += stA[e * Nq + tid] + stB[e * Nq + tid];
            }
#pragma unroll
            for (int m = 0; m < D; m++) { sw[m] = swnA[m]; swb[m] = swnB[m]; }
        }
        double cwA = ld_tab(t.fC + (tn)), cwB = ld_tab(t.fC + (Nq + tn));
        static_assert(NN * NC == Nq, "reducer items = volume nodes");
        constexpr int PS = S::PS, FS = S::FS;
        const int fslot = tid + (PS - NN) * ca;               // padded slot of this thread's volume node (FdSmem)
        const int fslot3 = fslot + (cc == 0 ? N - 1 : -1);    // ... one column to the left (cyclic), see SSE_FD_RED3
        // one pair of padded stages for all facet sub-rounds (the barrier that closes a sub-round already separates its
        // reducer loads from the next stores); they lie over the volume stages, hence the barrier here
        double* stA = s_stage;
        double* stB = s_stage + NC * FS;
        __syncthreads();
        const int re = tid / NN, rjj = tid - re * NN, rx = rjj / N, ry = rjj - rx * N;
        int rc3 = ry ? N - ry : 0;          // volume column (fr - 3 - y) mod N feeding facet node (x, y) of face 4, sub-round 3
#ifndef SSE_FD_UNROLL_F
#define SSE_FD_UNROLL_F 4
#endif
        constexpr int FUN = SSE_FD_UNROLL_F;
#pragma unroll FUN
        for (int fr = 0; fr < NFR; fr += 2) {
            const int fA = fr < 3 ? fr : 3, fB = fr + 1 < 3 ? fr + 1 : 3;
            const double cwnA = (fr + 2 < NFR) ? ld_tab(t.fC + ((fr + 2) * Nq + tn)) : 0.0;
            const double cwnB = (fr + 3 < NFR) ? ld_tab(t.fC + ((fr + 3) * Nq + tn)) : 0.0;
            if (node) {
                double hA[D], hB[D];
#pragma unroll
                for (int n = 0; n < D; n++) {
                    if (g.nJq) {
                        hA[n] = g.nJq[n + D * (fA + (size_t)4 * (tid + (size_t)Nq * k))];
                        hB[n] = g.nJq[n + D * (fB + (size_t)4 * (tid + (size_t)Nq * k))];
                    } else {
#if SSE_FD_NREF_STATIC
                        hA[n] = tet_face_h(fA, lam, n); hB[n] = tet_face_h(fB, lam, n);
#else
                        double sa = 0.0, sb = 0.0;
#pragma unroll
                        for (int l = 0; l < D; l++) { sa = fma(lam[l][n], t.nref[l + D * fA], sa); sb = fma(lam[l][n], t.nref[l + D * fB], sb); }
                        hA[n] = sa; hB[n] = sb;
#endif
                    }
                }
                const int jA = facet_partner<N>(fr, ca, cb, cc), jB = facet_partner<N>(fr + 1, ca, cb, cc);
                double gA[D], gB[D], qA[NP], qB[NP], pA[NC], pB[NC];
#pragma unroll
                for (int n = 0; n < D; n++) { gA[n] = cwA * (s_hnf[n * Nf + jA] + hA[n]); gB[n] = cwB * (s_hnf[n * Nf + jB] + hB[n]); }
#pragma unroll
                for (int c = 0; c < NP; c++) { qA[c] = s_fprim[c * Nf + jA]; qB[c] = s_fprim[c * Nf + jB]; }
                ec_contract_scaled2<D>(L, qi, qA, qB, gA, gB, pA, pB);
                // SSE_FD_RED3: when both sub-rounds pair with the slanted face, the second stage is stored one column to the left
                // (c -> c - 1), so that its reducer reads the same column as the reducer of the first stage
                const int fslotB = (SSE_FD_RED3 && fA == fB) ? fslot3 : fslot;
#pragma unroll
                for (int e = 0; e < NC; e++) { r[e] -= pA[e] + pB[e]; stA[e * FS + fslot] = pA[e]; stB[e * FS + fslotB] = pB[e]; }
            }
            cwA = cwnA; cwB = cwnB;
            __syncthreads();
            // (facet node, variable) reducers: thread = (re, rjj) sums the N staged vectors of its facet node, for the
            // face of either sub-round; with N = N_c = 5 there are exactly N^3 such items per face
            if (node) {
                double sA = 0.0, sB = 0.0;
                if (fA != fB) {             // two different faces: independent targets
                    int bA, dA, bB, dB;
                    if (fr == 0) { bA = rx * PS + ry; dA = N; bB = rjj; dB = PS; }
                    else { bA = rjj; dA = PS; bB = rx * PS + rc3; dB = N; rc3 = rc3 + 1 == N ? 0 : rc3 + 1; }
                    int tB = rjj;
#if SSE_FD_RED3
                    if (fr != 0) { tB = rx * N + (bB - rx * PS); bB = rx * PS + ry; }     // slanted face: own column, permuted target
#endif
#pragma unroll
                    for (int i = 0; i < N; i++) { sA += stA[re * FS + bA + i * dA]; sB += stB[re * FS + bB + i * dB]; }
                    s_ff[re * S::FFS + fA * NN + rjj] -= sA;
                    s_ff[re * S::FFS + fB * NN + tB] -= sB;
                } else {                    // both sub-rounds feed face 4: one reducer sums both stages
                    const int cA = rc3, cB = cA + 1 == N ? 0 : cA + 1;
                    rc3 = cB + 1 == N ? 0 : cB + 1;
#if SSE_FD_RED3
                    // The column c = (fr - 3 - y) mod N that feeds facet node (x, y) is an involution of y: this thread sums
                    // column ry of both stages (its own bank: conflict-free loads, they sit between two barriers) and hands
                    // the sum to facet node (rx, cA) -- the same additions in the same order for every facet node
#pragma unroll
                    for (int i = 0; i < N; i++) { sA += stA[re * FS + rx * PS + ry + i * N]; sB += stB[re * FS + rx * PS + ry + i * N]; }
                    s_ff[re * S::FFS + 3 * NN + rx * N + cA] -= sA + sB;
#else
#pragma unroll
                    for (int i = 0; i < N; i++) { sA += stA[re * FS + rx * PS + cA + i * N]; sB += stB[re * FS + rx * PS + cB + i * N]; }
                    s_ff[re * S::FFS + 3 * NN + rjj] -= sA + sB;
#endif
                }
            }
            __syncthreads();      // two reducers of one iteration may hit the same facet node only across iterations
        }
    } else {
    // ---- volume term: NVR rounds along the tensor lines (flux_difference!, flux_differencing_form.jl:37-75)
    int buf = 0;
#pragma unroll 1
    for (int rd = 0; rd < NVR; rd++, buf ^= 1) {
        double* st = s_stage + buf * NC * Nq;
        const int l = rd / NSH, sh = rd - l * NSH + 1;
        const int cl = (l == 0) ? ca : ((l == 1) ? cb : cc);
        const int stride = (l == 0) ? NN : ((l == 1) ? N : 1);
        int cj = cl + sh; if (cj >= N) cj -= N;
        int cs = cl - sh; if (cs < 0) cs += N;
        const bool half = (2 * sh == N);
        const bool active = node && !(half && cl >= sh);
        const bool recv = node && !(half && cs >= sh);
        const int j = tid + (cj - cl) * stride;
        double swn[D];                                         // prefetch the next round's weights
#pragma unroll
        for (int m = 0; m < D; m++) swn[m] = (rd + 1 < NVR) ? ld_tab(t.vS + (((rd + 1) * D + m) * Nq + tn)) : 0.0;
        if (active) {
            double gv[D], qj[NP], phi[NC];
#pragma unroll
            for (int n = 0; n < D; n++) gv[n] = 0.0;
#pragma unroll
            for (int m = 0; m < D; m++) {
                if (m >= l) {                                  // S_m couples eta_l-lines only for m >= l
#pragma unroll
                    for (int n = 0; n < D; n++) gv[n] = fma(sw[m], lam[m][n] + s_lam[(m + D * n) * Nq + j], gv[n]);
                }
            }
#pragma unroll
            for (int c = 0; c < NP; c++) qj[c] = s_prim[c * Nq + j];
            ec_contract_scaled<D>(L, qi, qj, gv, phi);
#pragma unroll
            for (int e = 0; e < NC; e++) { r[e] -= phi[e]; st[e * Nq + j] = phi[e]; }
        }
        __syncthreads();
        if (recv) {
#pragma unroll
            for (int e = 0; e < NC; e++) r[e] += st[e * Nq + tid];
        }
#pragma unroll
        for (int m = 0; m < D; m++) sw[m] = swn[m];
    }

    // ---- facet correction: NFR sub-rounds (facet_correction!, flux_differencing_form.jl:126-168)
    double cw = ld_tab(t.fC + (tn));
    double hq[D];
#pragma unroll
    for (int n = 0; n < D; n++) hq[n] = 0.0;
#pragma unroll 1
    for (int fr = 0; fr < NFR; fr++, buf ^= 1) {
        double* st = s_stage + buf * NC * Nq;
        const int f = fr < 3 ? fr : 3;
        const double cwn = (fr + 1 < NFR) ? ld_tab(t.fC + ((fr + 1) * Nq + tn)) : 0.0;
        if (node) {
            if (fr <= 3) {                         // 2 halfnJq[:, f, i] = sum_l Lambda[i,l,:] nref[l,f]   mesh.jl:262-269
                if (g.nJq) {
#pragma unroll
                    for (int n = 0; n < D; n++) hq[n] = g.nJq[n + D * (f + (size_t)4 * (tid + (size_t)Nq * k))];
                } else {
#pragma unroll
                    for (int n = 0; n < D; n++) {
                        double s = 0.0;
#pragma unroll
                        for (int l = 0; l < D; l++) s = fma(lam[l][n], t.nref[l + D * f], s);
                        hq[n] = s;
                    }
                }
            }
            const int j = facet_partner<N>(fr, ca, cb, cc);
            double gv[D], qj[NP], phi[NC];
#pragma unroll
            for (int n = 0; n < D; n++) gv[n] = cw * (s_hnf[n * Nf + j] + hq[n]);
#pragma unroll
            for (int c = 0; c < NP; c++) qj[c] = s_fprim[c * Nf + j];
            ec_contract_scaled<D>(L, qi, qj, gv, phi);
#pragma unroll
            for (int e = 0; e < NC; e++) { r[e] -= phi[e]; st[e * Nq + tid] = phi[e]; }
        }
        cw = cwn;
        __syncthreads();
        // (facet node, variable) reducers: every facet node of the sub-round's face sums its N staged vectors
        for (int q = tid; q < NN * NC; q += blockDim.x) {
            const int e = q / NN, jj = q - e * NN, x = jj / N, y = jj - x * N;
            int base, stride;
            if (fr == 0) { base = x * NN + y; stride = N; }
            else if (fr < 3) { base = jj; stride = NN; }
            else { int c = (fr - 3) - y; if (c < 0) c += N; base = x * NN + c; stride = N; }
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < N; i++) s += st[e * Nq + base + i * stride];
            s_ff[e * S::FFS + f * NN + jj] -= s;
        }
    }
    }
    __syncthreads();
    // ---- lift: r_q -= R' f_f (flux_differencing_form.jl:341-342); r_q goes to k_project_ct through the u_q scratch
#ifndef SSE_FD_LIFT_PLAIN
    if constexpr (DUAL) {
        // The slanted face through its 1-D factors, R[(a, y), (a, b, c)] = I3[y, b] r3[c] (FacetR): the N^2 N_c sums
        // G[e][a][b] = sum_y I3[y, b] f_f[e][(a, y)] are shared by the N nodes c of a line, one per thread; a node then
        // needs 5 + 15 facet values instead of 40.  The stages are free by now and hold G.
        double* s_G = s_stage;
        const double* fac = t.facR;
        if (node) {
            const int re = tid / NN, rjj = tid - re * NN, rx = rjj / N, ry = rjj - rx * N;
            double gs = 0.0;
#pragma unroll
            for (int y = 0; y < N; y++) gs = fma(ld_tab(fac + 4 * N + y + N * ry), s_ff[re * S::FFS + 3 * NN + rx * N + y], gs);
            s_G[tid] = gs;
        }
        double rw[3];
#pragma unroll
        for (int fr = 0; fr < 3; fr++) rw[fr] = ld_tab(t.fR + (fr * Nq + tn));
        const double r3c = ld_tab(fac + 3 * N + cc % N);
        __syncthreads();
        if (node) {
#pragma unroll
            for (int fr = 0; fr < 3; fr++) {
                const int j = facet_partner<N>(fr, ca, cb, cc);
#pragma unroll
                for (int e = 0; e < NC; e++) r[e] = fma(-rw[fr], s_ff[e * S::FFS + j], r[e]);
            }
#pragma unroll
            for (int e = 0; e < NC; e++) r[e] = fma(-r3c, s_G[e * NN + ca * N + cb], r[e]);
#pragma unroll
            for (int e = 0; e < NC; e++) u_q[((size_t)k * NC + e) * Nq + tid] = r[e];
        }
        return;
    }
#endif
    if (node) {
        double rw[NFR];
#pragma unroll
        for (int fr = 0; fr < NFR; fr++) rw[fr] = ld_tab(t.fR + (fr * Nq + tid));
#pragma unroll
        for (int fr = 0; fr < NFR; fr++) {
            const int j = facet_partner<N>(fr, ca, cb, cc);
#pragma unroll
            for (int e = 0; e < NC; e++) r[e] = fma(-rw[fr], s_ff[e * S::FFS + j], r[e]);
        }
#pragma unroll
        for (int e = 0; e < NC; e++) u_q[((size_t)k * NC + e) * Nq + tid] = r[e];
    }
}


// ---------------------------------------------------------------------------------------------------------
// pass B-1 for LinearAdvectionEquation + StandardForm + ReferenceOperators on collapsed tets
// (standard_form_first_order.jl:16-63).  The flux is linear, f_n = a_n u, so the d^2 (D_m, D_m') pairs of the
// reference collapse to d pairs acting on  g_m = (sum_n halfWLambda_mn a_n) u  and on u, and the facet
// difference  sum_n halfN_n R f_n  is  0.5 (a.n) u_f  with the u_f = R u_q pass A already wrote.
// HBM-bound: ~16 kB of metric/state traffic per element against ~30 kflop.
template <int N> struct AdvTabs { double D1[3][N * N]; };     // D_1D[m][t + N*s]: row t, column s

template <int N, int MINB>
__global__ void __launch_bounds__(Tet<N>::NT, MINB)
k_standard_adv_ct(AdvTabs<N> a, CtDev t, Geo g, Law L, long long first, double* __restrict__ u_q, const double* __restrict__ u_f) {
    constexpr int D = 3;
    using T = Tet<N>;
    constexpr int Nq = T::Nq, Nf = T::Nf, NN = N * N, NFR = 3 + N;
    __shared__ double s_u[Nq], s_g[D][Nq], s_ff[Nf], s_D[D][N * N];
    const int tid = threadIdx.x;
    const long long k = first + blockIdx.x;
    const bool node = tid < Nq;
    const int ca = tid / NN, cb = (tid / N) % N, cc = tid % N;
    for (int i = tid; i < D * N * N; i += blockDim.x) s_D[i / (N * N)][i % (N * N)] = a.D1[i / (N * N)][i % (N * N)];
    // every global load of the element is issued before any arithmetic (idle lanes load a clamped, valid address): the
    // kernel is a few hundred instructions long, so its time is the latency of these loads unless they overlap
    static_assert(Nf <= T::NT && Nq <= T::NT, "one volume node and one facet node per thread");
    const bool fac = tid < Nf;
    const int tn = node ? tid : Nq - 1, tj = fac ? tid : Nf - 1;
    const size_t jo = (size_t)(g.mapP[(size_t)k * Nf + tj] - 1);
    const double u = __ldcs(u_q + (size_t)k * Nq + tn);
    double lam[D][D], njf[D], rw[NFR];
#pragma unroll
    for (int m = 0; m < D; m++)
#pragma unroll
        for (int n = 0; n < D; n++) lam[m][n] = __ldcs(g.Lambda_q + ((size_t)k * D * D + (m + D * n)) * Nq + tn);
    const double ui = __ldcs(u_f + (size_t)k * Nf + tj);
    const double jf = __ldcs(g.J_f + (size_t)k * Nf + tj);
#pragma unroll
    for (int m = 0; m < D; m++) njf[m] = __ldcs(g.nJf + m + D * ((size_t)k * Nf + tj));
    const double uo = __ldcs(u_f + jo);
#pragma unroll
    for (int fr = 0; fr < NFR; fr++) rw[fr] = t.fR[fr * Nq + tn];
    const double hw = 0.5 * t.W[tn], bf = t.Bf[tj];
    double c[D];
#pragma unroll
    for (int m = 0; m < D; m++) {
        double s = 0.0;
#pragma unroll
        for (int n = 0; n < D; n++) s = fma(hw * lam[m][n], L.a[n], s);      // halfWLambda_mn a_n
        c[m] = s;
        if (node) s_g[m][tid] = s * u;
    }
    if (node) s_u[tid] = u;
    {
        const double ijf = rcp_fast(jf);
        double an = 0.0;
#pragma unroll
        for (int m = 0; m < D; m++) an = fma(L.a[m], njf[m] * ijf, an);
        double fs = (0.5 * (ui + uo)) * an;                                   // F#.n        ConservationLaws.jl:75-128
        if (L.inviscid == SSE_FLUX_LAX_FRIEDRICHS) fs = fma(L.half_lambda * fabs(an), ui - uo, fs);
        if (fac) s_ff[tid] = (bf * jf) * (fs - 0.5 * an * ui);                // BJf (f* - sum_n halfN_n R f_n)
    }
    __syncthreads();
    if (node) {
        double r = 0.0;
#pragma unroll
        for (int m = 0; m < D; m++) {
            const int cm = (m == 0) ? ca : ((m == 1) ? cb : cc);
            const int stride = (m == 0) ? NN : ((m == 1) ? N : 1);
            const int base = tid - cm * stride;
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int q = 0; q < N; q++) {
                s1 = fma(s_D[m][q + N * cm], s_g[m][base + q * stride], s1);     // D_m' (halfWLambda f)
                s2 = fma(s_D[m][cm + N * q], s_u[base + q * stride], s2);       // D_m f
            }
            r += s1 - c[m] * s2;
        }
#pragma unroll
        for (int fr = 0; fr < NFR; fr++) r = fma(-rw[fr], s_ff[facet_partner<N>(fr, ca, cb, cc)], r);   // - R' f_f
        u_q[(size_t)k * Nq + tid] = r;
    }
}

}  // namespace sse
