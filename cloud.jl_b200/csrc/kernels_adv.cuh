// kernels_adv.cuh — BASELINE config 4: LinearAdvectionEquation + StandardForm + ReferenceOperators on collapsed tetrahedra
// (standard_form_first_order.jl:16-63) as TWO kernels that stream every element once:
//
//   k_adv_facets_ct  pass A : u_f = R V u  (nodal_values!, Solvers.jl:505-507); keeps the modal coefficients in the scratch
//   k_adv_fused_ct   pass B : V u -> volume terms (D_m, D_m') -> interface flux, lift -> V' -> M^-1 (V, W/J, V') -> dudt
//                             (time_derivative!, standard_form_first_order.jl:16-63, mass_matrix.jl:185-196); u_q, the
//                             fluxes and r_q never leave the SM
//
// HBM-bound path (2-3 flop/B), so the design is about bytes and about keeping them in flight:
//   * one thread per (element, eta_3 index) owns the N x N slab of its element in registers (the mapping of the projection
//     kernels): derivatives along eta_1 / eta_2 are register-only with constant-bank coefficients, the eta_3 direction and the
//     slanted face go through a per-element tile in shared memory; every exchange stays inside the N lanes of one element,
//     so the kernels contain no CTA barrier at all (__syncwarp only);
//   * the flux is linear, f_n = a_n u, so the d^2 (D, D') pairs of the reference collapse to d pairs acting on
//     c_m u with c_m = sum_n (W Lambda_mn / 2) a_n -- per node 3 coefficients instead of the 9 metric terms -- and the facet
//     term BJf (f* - sum_n halfN_n R f_n) collapses to fa u+ + fl (u- - u+) with fa = BJf (a.n)/2, fl = BJf halflambda |a.n|;
//   * these derived coefficients (and W / J, and mapP as 0-based int32) are built once at sse_create in the order the
//     kernel reads them: [task of 32/N elements][item][32 lanes], so that every load instruction of a warp is one contiguous,
//     aligned 256-byte read.  Per element and residual: 8.6 kB of HBM traffic against 14.6 kB algorithmic.
#pragma once
#include "kernels_ct.cuh"

namespace sse {

template <int N> struct AdvTabs2 { double D1[3][N * N]; };     // D_1D[m][row + N * col]

template <int N> struct AdvSmem {
    using T = Tet<N>;
    static constexpr int GPW = 32 / N;
    static constexpr int QS = T::Nq + ((N - T::Nq % 16) % 16 + 16) % 16;            // nodal tile stride (= N mod 16)
    static constexpr int RSM = (N == 5) ? 15 : N;
    static constexpr int RS = T::RED_SPAN + ((RSM - T::RED_SPAN % 16) % 16 + 16) % 16;
    static constexpr int T2a = RS > QS ? RS : QS;
    static constexpr int T2 = T2a > T::Nf ? T2a : T::Nf;                            // second tile | V' partials | facet tile (pass A) | u+ gather
    static constexpr int GSLOTS = GPW + (32 % N != 0);   // groups per warp in shared memory: the idle lanes of a warp (32 mod N)
                                                         // compute along on a dummy slot instead of being masked everywhere
    static constexpr int tile = 0;                       // [QS]   u_q of the element
    static constexpr int x = 0;                          // [Np]   modal coefficients: over the tile (read before u_q is stored,
                                                         //        rewritten after its last use)
    static constexpr int tile2 = tile + QS;              // [T2]   c_3 u_q, later the V' partials
    static constexpr int f3 = tile2 + T2;                // [N N]  f_f of the slanted face
    static constexpr int group = f3 + N * N + ((N - (f3 + N * N) % 16) % 16 + 16) % 16;   // per-element stride (= N mod 16)
};

// which facet node does lane a3 own as item (face, i)?  faces 0..2: the node whose second facet coordinate is a3;
// face 3: row a1 = a3 of the face, column i
template <int N> __host__ __device__ constexpr int adv_facet_node(int face, int i, int a3) {
    return face < 3 ? face * N * N + i * N + a3 : 3 * N * N + a3 * N + i;
}

// ---------------------------------------------------------------------------------------------------------
// pass A: facet values of every element (the neighbours read them in pass B) and a copy of the modal coefficients
template <int N, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_adv_facets_ct(SFCoef<N> cf, FacetR<N> fr, CtDev t, AdvDev ad, long long first, long long count, const double* __restrict__ u,
                double* __restrict__ u_f) {
    using T = Tet<N>;
    using S = AdvSmem<N>;
    constexpr int Np = T::Np, Nf = T::Nf, GPW = S::GPW;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane / N, a3 = lane - gl * N;
    const long long task = first / GPW + (long long)blockIdx.x * WARPS + warp;
    if (task > (first + count - 1) / GPW) return;              // a spare warp of the last CTA (no CTA barrier below: safe)
    const long long k = task * GPW + gl;
    const bool act = gl < GPW && k >= first && k < first + count;
    double* s_x = sm + (warp * S::GSLOTS + gl) * S::group + S::x;
    double* s_t = s_x - S::x + S::tile;
    double* s_wf = s_x - S::x + S::tile2;                       // facet tile (4 N N <= T2)
    static_assert(Nf <= S::T2, "facet tile fits the second tile");
    if (act) {
        for (int l = a3; l < Np; l += N) cp_async8(s_x + l, u + (size_t)k * Np + l);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    if (act) {
        for (int l = a3; l < Np; l += N) ad.um[(size_t)k * Np + l] = s_x[l];
    }
    double y[N][N];
    sf3_fwd<N, N, true, true>(cf, t.C3 + a3, s_x, y);
    __syncwarp();                                               // the tile lies over the modal coefficients
    if (act) {
#pragma unroll
        for (int a1 = 0; a1 < N; a1++)
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) s_t[(a1 * N + a2) * N + a3] = y[a1][a2];
        facet_rows_slab<N>(fr, y, a3, s_wf);
    }
    __syncwarp();
    if (act) facet_rows_face3<N>(fr, s_t, a3, s_wf);
    __syncwarp();
    if (act) {
        for (int j = a3; j < Nf; j += N) u_f[(size_t)k * Nf + j] = s_wf[j];
    }
}

// ---------------------------------------------------------------------------------------------------------
// pass B, fused
#ifdef SSE_ADV_MAXREG
template <int N, int WARPS, int MINB>
__global__ void __maxnreg__(SSE_ADV_MAXREG)
#else
template <int N, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
#endif
k_adv_fused_ct(SFCoef<N> cf, FacetR<N> fr, AdvTabs2<N> tb, CtDev t, AdvDev ad, Geo g, long long first, long long count,
               const double* __restrict__ u_f, double* __restrict__ dudt, RkStage rk) {
    using T = Tet<N>;
    using S = AdvSmem<N>;
    constexpr int Np = T::Np, NN = N * N, GPW = S::GPW, NI = 4 * N;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane / N, a3 = lane - gl * N;
    const long long task = first / GPW + (long long)blockIdx.x * WARPS + warp;
    if (task > (first + count - 1) / GPW) return;              // a spare warp of the last CTA (no CTA barrier below: safe)
    const long long k = task * GPW + gl;
    const bool act = gl < GPW && k >= first && k < first + count;
    double* s_x = sm + (warp * S::GSLOTS + gl) * S::group + S::x;
    double* s_t = s_x - S::x + S::tile;
    double* s_t2 = s_x - S::x + S::tile2;
    double* s_f3 = s_x - S::x + S::f3;
    // warp-interleaved tables of this task: entry (item) of this lane
    const double* pC = ad.C + (size_t)task * (3 * NN) * 32 + lane;
    const double* pW = ad.iJW + (size_t)task * NN * 32 + lane;
    const double* pF = ad.F + (size_t)task * (2 * NI) * 32 + lane;
    const int* pM = ad.map + (size_t)task * NI * 32 + lane;

#ifndef SSE_ADV_PREFETCH
#define SSE_ADV_PREFETCH 150       // tasks ahead; measured at 196 608 elements: 0 -> 0.482, 32..300 -> 0.429-0.438, 600 -> 0.474, 1200 -> 0.500 ms
#endif
#if SSE_ADV_PREFETCH > 0
    // The kernel is bound by the latency of its table reads at 8 warps per SM: pull the tables of a task that a later wave of
    // CTAs will process from DRAM into L2 now (one prefetch per 128-byte line, spread over the lanes)
    {
        const long long pt = task + SSE_ADV_PREFETCH;
        if (pt <= (first + count - 1) / GPW) {
            const char* pc = (const char*)(ad.C + (size_t)pt * (3 * NN) * 32);
            for (int i = lane; i < 3 * NN * 2; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + (size_t)i * 128));
            const char* pw = (const char*)(ad.iJW + (size_t)pt * NN * 32);
            for (int i = lane; i < NN * 2; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pw + (size_t)i * 128));
            const char* pf = (const char*)(ad.F + (size_t)pt * (2 * NI) * 32);
            for (int i = lane; i < 2 * NI * 2; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (size_t)i * 128));
            const char* pm = (const char*)(ad.map + (size_t)pt * NI * 32);
            for (int i = lane; i < NI; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pm + (size_t)i * 128));
            const char* pu = (const char*)(ad.um + (size_t)pt * GPW * Np);
            for (int i = lane; i < (GPW * Np * 8 + 127) / 128; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pu + (size_t)i * 128));
        }
    }
#endif
    // neighbour indices first: the gather of u+ depends on them (two round trips), everything else is one
    int jo[NI];
#pragma unroll
    for (int i = 0; i < NI; i++) jo[i] = __ldcs(pM + i * 32);
    if (act) {
        for (int l = a3; l < Np; l += N) cp_async8(s_x + l, ad.um + (size_t)k * Np + l);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();

    double y[N][N], r[N][N];
    sf3_fwd<N, N, true, true>(cf, t.C3 + a3, s_x, y);                 // u_q = V u
    __syncwarp();                                              // the tile lies over the modal coefficients

    // ---- volume terms: r = sum_m D_m' (c_m u) - c_m (D_m u)        standard_form_first_order.jl:38-46
    // eta_3 lines (m = 2) first: they cross the N lanes of the element, u_q and c_3 u_q go through the two tiles
#pragma unroll
    for (int a1 = 0; a1 < N; a1++)
#pragma unroll
        for (int a2 = 0; a2 < N; a2++) {
            s_t[(a1 * N + a2) * N + a3] = y[a1][a2];
            s_t2[(a1 * N + a2) * N + a3] = __ldg(pC + (2 * NN + a1 * N + a2) * 32) * y[a1][a2];
        }
    __syncwarp();
    {
        double d3r[N], d3c[N];                                 // eta_3 row / column of the 1-D derivative matrix of this lane
#pragma unroll                                                 // (select chains: a3 is not a compile-time index of the parameter bank)
        for (int q = 0; q < N; q++) {
            d3c[q] = tb.D1[2][q]; d3r[q] = tb.D1[2][N * q];
#pragma unroll
            for (int c = 1; c < N; c++) { d3c[q] = (a3 == c) ? tb.D1[2][q + N * c] : d3c[q]; d3r[q] = (a3 == c) ? tb.D1[2][c + N * q] : d3r[q]; }
        }
#pragma unroll
        for (int a1 = 0; a1 < N; a1++)
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) {
                double s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int q = 0; q < N; q++) {
                    s1 = fma(d3c[q], s_t2[(a1 * N + a2) * N + q], s1);
                    s2 = fma(d3r[q], s_t[(a1 * N + a2) * N + q], s2);
                }
                r[a1][a2] = s1 - __ldg(pC + (2 * NN + a1 * N + a2) * 32) * s2;     // second read of c_3: an L1 hit, not 50 registers
            }
    }
    __syncwarp();
    // the second tile is free until the V' partials: the gather of the neighbours' facet values u+ lands there as asynchronous
    // copies (no registers while in flight) and is hidden behind the register-only eta_1 / eta_2 terms
    if (act) {
#pragma unroll
        for (int i = 0; i < NI; i++) cp_async8(s_t2 + i * N + a3, u_f + jo[i]);
    }
    cp_async_commit();
    // eta_1 lines (m = 0) and eta_2 lines (m = 1) live in this thread's registers
#pragma unroll
    for (int a2 = 0; a2 < N; a2++) {
        double c[N], gq[N];
#pragma unroll
        for (int q = 0; q < N; q++) { c[q] = __ldcs(pC + (0 * NN + q * N + a2) * 32); gq[q] = c[q] * y[q][a2]; }
#pragma unroll
        for (int a1 = 0; a1 < N; a1++) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int q = 0; q < N; q++) { s1 = fma(tb.D1[0][q + N * a1], gq[q], s1); s2 = fma(tb.D1[0][a1 + N * q], y[q][a2], s2); }
            r[a1][a2] += s1 - c[a1] * s2;
        }
    }
#pragma unroll
    for (int a1 = 0; a1 < N; a1++) {
        double c[N], gq[N];
#pragma unroll
        for (int q = 0; q < N; q++) { c[q] = __ldcs(pC + (1 * NN + a1 * N + q) * 32); gq[q] = c[q] * y[a1][q]; }
#pragma unroll
        for (int a2 = 0; a2 < N; a2++) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int q = 0; q < N; q++) { s1 = fma(tb.D1[1][q + N * a2], gq[q], s1); s2 = fma(tb.D1[1][a2 + N * q], y[a1][q], s2); }
            r[a1][a2] += s1 - c[a2] * s2;
        }
    }
    cp_async_wait<0>();                                        // every lane reads back only what it gathered itself
    const double* uo = s_t2 + a3;                              // u+ of item i at uo[i * N]

    // ---- facet terms: f_f = BJf (f* - sum_n halfN_n R f_n) = fa u+ + fl (u- - u+), lifted with the 1-D factors of R
    //      (standard_form_first_order.jl:48-57; ConservationLaws.jl:75-128)
    {
        double ff[N];
        // face 0 (eta_2 = -1): nodes (a1, a3), u- = sum_a2 r0[a2] u_q
#pragma unroll
        for (int i = 0; i < N; i++) {
            double ui = 0.0;
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) ui = fma(fr.r0[a2], y[i][a2], ui);
            const double fa = __ldcs(pF + (0 * NI + 0 * N + i) * 32), fl = __ldcs(pF + (1 * NI + 0 * N + i) * 32);
            ff[i] = fma(fa, uo[(0 * N + i) * N], fl * (ui - uo[(0 * N + i) * N]));
        }
#pragma unroll
        for (int a1 = 0; a1 < N; a1++)
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) r[a1][a2] = fma(-fr.r0[a2], ff[a1], r[a1][a2]);
        // faces 1, 2 (eta_1 = +1, -1): nodes (a2, a3)
#pragma unroll
        for (int f = 1; f <= 2; f++) {
            const double* rf = f == 1 ? fr.r1 : fr.r2;
#pragma unroll
            for (int i = 0; i < N; i++) {
                double ui = 0.0;
#pragma unroll
                for (int a1 = 0; a1 < N; a1++) ui = fma(rf[a1], y[a1][i], ui);
                const double fa = __ldcs(pF + (0 * NI + f * N + i) * 32), fl = __ldcs(pF + (1 * NI + f * N + i) * 32);
                ff[i] = fma(fa, uo[(f * N + i) * N], fl * (ui - uo[(f * N + i) * N]));
            }
#pragma unroll
            for (int a1 = 0; a1 < N; a1++)
#pragma unroll
                for (int a2 = 0; a2 < N; a2++) r[a1][a2] = fma(-rf[a1], ff[a2], r[a1][a2]);
        }
        // face 3 (eta_3 = -1): this lane owns row a1 = a3: u- = sum_a2 I3[b, a2] sum_c r3[c] u_q[a3, a2, c]
        {
            double tt[N];
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < N; c++) s = fma(fr.r3[c], s_t[(a3 * N + a2) * N + c], s);
                tt[a2] = s;
            }
#pragma unroll
            for (int b = 0; b < N; b++) {
                double ui = 0.0;
#pragma unroll
                for (int a2 = 0; a2 < N; a2++) ui = fma(fr.I3[b + N * a2], tt[a2], ui);
                const double fa = __ldcs(pF + (0 * NI + 3 * N + b) * 32), fl = __ldcs(pF + (1 * NI + 3 * N + b) * 32);
                s_f3[a3 * N + b] = fma(fa, uo[(3 * N + b) * N], fl * (ui - uo[(3 * N + b) * N]));
            }
        }
        __syncwarp();
        double r3c = fr.r3[0];
#pragma unroll
        for (int c = 1; c < N; c++) r3c = (a3 == c) ? fr.r3[c] : r3c;
#pragma unroll
        for (int a1 = 0; a1 < N; a1++) {
            double f3[N];
#pragma unroll
            for (int yb = 0; yb < N; yb++) f3[yb] = s_f3[a1 * N + yb];
#pragma unroll
            for (int a2 = 0; a2 < N; a2++) {
                double G = 0.0;
#pragma unroll
                for (int yb = 0; yb < N; yb++) G = fma(fr.I3[yb + N * a2], f3[yb], G);
                r[a1][a2] = fma(-r3c, G, r[a1][a2]);
            }
        }
    }

    // ---- dudt = M^-1 V' r_q      (standard_form_first_order.jl:59-62, mass_matrix.jl:185-196)
    double out[T::LPT];
    const double* c3 = t.C3 + a3;
    __syncwarp();                                              // the partials lie over the second tile
    sf3_bwd_partials<N, N, true, true>(cf, c3, r, s_t2 + a3 * T::RED_SA);
    __syncwarp();
    sf3_bwd_reduce<N>(s_t2, a3, out);
#pragma unroll
    for (int q = 0; q < T::LPT; q++) { const int l = a3 * T::LPT + q; if (l < Np) s_x[l] = out[q]; }
    __syncwarp();
    sf3_fwd<N, N, true, true>(cf, c3, s_x, r);
#pragma unroll
    for (int a1 = 0; a1 < N; a1++)
#pragma unroll
        for (int a2 = 0; a2 < N; a2++) r[a1][a2] *= __ldcs(pW + (a1 * N + a2) * 32);
    sf3_bwd_partials<N, N, true, true>(cf, c3, r, s_t2 + a3 * T::RED_SA);
    __syncwarp();
    sf3_bwd_reduce<N>(s_t2, a3, out);
    if (act) {
#pragma unroll
        for (int q = 0; q < T::LPT; q++) {
            const int l = a3 * T::LPT + q;
            if (l < Np) {
                const size_t idx = (size_t)k * Np + l;
                dudt[idx] = out[q];
                flag_nonfinite(g.flag, out[q]);
                if (rk.u) {
                    const double tm = fma(rk.A, rk.tmp[idx], rk.dt * out[q]);
                    rk.tmp[idx] = tm;
                    rk.u[idx] = fma(rk.B, tm, rk.u[idx]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// builders of the warp-interleaved tables (run once at sse_create)
template <int N>
__global__ void k_adv_build(long long Ne, const double* __restrict__ W, const double* __restrict__ Lambda_q, const double* __restrict__ J_q,
                            const double* __restrict__ J_f, const double* __restrict__ nJf, const double* __restrict__ Bf,
                            const long long* __restrict__ mapP, Law L, double* __restrict__ C, double* __restrict__ iJW,
                            double* __restrict__ F, int* __restrict__ map) {
    using T = Tet<N>;
    constexpr int Nq = T::Nq, Nf = T::Nf, NN = N * N, GPW = 32 / N, NI = 4 * N, D = 3;
    const long long k = blockIdx.x;
    const long long task = k / GPW;
    const int gl = (int)(k - task * GPW);
    for (int i = threadIdx.x; i < Nq; i += blockDim.x) {
        const int a1 = i / NN, a2 = (i / N) % N, a3 = i % N, lane = gl * N + a3;
        const double hw = 0.5 * W[i];
#pragma unroll
        for (int m = 0; m < D; m++) {
            double s = 0.0;
#pragma unroll
            for (int n = 0; n < D; n++) s = fma(hw * Lambda_q[((size_t)k * D * D + (m + D * n)) * Nq + i], L.a[n], s);   // halfWLambda_mn a_n
            C[((size_t)task * (3 * NN) + m * NN + a1 * N + a2) * 32 + lane] = s;
        }
        iJW[((size_t)task * NN + a1 * N + a2) * 32 + lane] = W[i] * rcp_fast(J_q[(size_t)k * Nq + i]);
    }
    for (int j = threadIdx.x; j < Nf; j += blockDim.x) {
        const int f = j / NN, x = (j % NN) / N, yv = j % N;
        const int a3 = f < 3 ? yv : x, item = f * N + (f < 3 ? x : yv), lane = gl * N + a3;
        const double jf = J_f[(size_t)k * Nf + j], ijf = rcp_fast(jf);
        double an = 0.0;
#pragma unroll
        for (int m = 0; m < D; m++) an = fma(L.a[m], nJf[m + D * ((size_t)k * Nf + j)] * ijf, an);     // a . n_f, n_f = nJf / J_f
        const double bj = Bf[j] * jf;
        F[((size_t)task * (2 * NI) + item) * 32 + lane] = bj * (0.5 * an);
        F[((size_t)task * (2 * NI) + NI + item) * 32 + lane] = (L.inviscid == SSE_FLUX_LAX_FRIEDRICHS) ? bj * (L.half_lambda * fabs(an)) : 0.0;
        map[((size_t)task * NI + item) * 32 + lane] = (int)(mapP[(size_t)k * Nf + j] - 1);
    }
}

}  // namespace sse
