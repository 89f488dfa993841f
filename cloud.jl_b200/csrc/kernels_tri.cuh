// kernels_tri.cuh — compile-time-sized kernels for 2-D Euler, flux differencing, ModalTensor(p) on collapsed triangles
// (BASELINE config 2: test/euler_vortex_2d_modal.jl), N = p + 1 = 3 .. 5 nodes per direction.
//
//   k_tri_nodal    pass A: entropy projection                      (flux_differencing_form.jl:214-292)
//   k_tri_fluxdiff pass B: interface flux, volume flux differencing, facet correction, lift, V', mass solve
//                          in ONE launch (flux_differencing_form.jl:294-347, mass_matrix.jl:185-196)
//
// A p = 4 triangle has 25 volume nodes, 15 facet nodes and 15 modes per variable: one WARP owns one element (lane = volume
// node (a1, a2) = a1 N + a2; lanes 0 .. 3N-1 double as the facet nodes), a CTA is a set of independent warps that walk over
// the elements with a grid stride.  Nothing in the element loop is a CTA barrier; the phases of an element are ordered by
// __syncwarp.  What the one-CTA-per-element kernels re-read for every element lives here for the lifetime of the warp:
//   * the pair weights of the lane's node (S_m[i, partner] / 4 of both line directions, C[i, partner] / 8 and R[partner, i]
//     of the three faces) in registers -- every partner is closed-form (tri_schedule_matches verifies the tables);
//   * the warped V as a dense 25 x 15 matrix (tensor_simplex.jl:84-140 collapses to A[a1,b1] B[a2,b1,b2] in 2-D) in shared
//     memory in the two orders its applications read it: V[:, l] across the lanes (V x: lane = node) and V[i, :] across
//     16 lanes (V' t: lane = (mode, variable pair), so all 32 lanes carry 2 of the 4 x 15 sums).
// Volume term: both nodes of a pair sit in the same warp, so the +phi of a unique pair goes to the partner by warp shuffles
// along the tensor line (the cyclic schedule of kernels_tensor.cuh: node c evaluates (c, c + s mod N), s = 1 .. N/2, and
// receives from c - s) -- no staging buffer and no barrier between the two halves of a pair.  The facet correction stages
// its three -phi per node once and the facet lanes sum their N entries.
// The loads of the next element (state, metrics, facet data, mapP) are issued before the arithmetic of the current one;
// the mapP-dependent neighbour gather is issued at the top of an element and consumed after its volume term.
#pragma once
#include "common.cuh"
#include "ct_api.h"

namespace sse {

// sum-factorised applies in pass A / pass B (tri_V_sf, tri_Vt_sf below) instead of the dense products: measured and rejected,
// 0.311 -> 0.380 ms and 0.515 -> 0.564 ms at 131 072 triangles (see the comment at tri_V_sf)
#ifndef SSE_TRI_SF_A
#define SSE_TRI_SF_A 0
#endif
#ifndef SSE_TRI_SF_B
#define SSE_TRI_SF_B 0
#endif
// row of V in registers for the three V applications of pass A (15 doubles per lane; 90 of the 600 shared-memory wavefronts
// per element were these table reads): pass A 0.311 -> 0.271 ms at 131 072 triangles with 150 registers / 12 warps per SM
#ifndef SSE_TRI_VREG_A
#define SSE_TRI_VREG_A 1
#endif
template <int N> struct TriT {
    static constexpr int D = 2, NC = 4, NP = 5;
    static constexpr int Nq = N * N, Np = N * (N + 1) / 2, Nf = 3 * N, NSH = N / 2;
    static_assert(Nq <= 32 && Np <= 16 && Nf <= 32, "one volume node per lane, one (mode, variable pair) per lane");
    // CTA-wide tables (doubles)
    static constexpr int vt = 0;                      // V[i, l] at l * 32 + i      (lane = node)
    static constexpr int vc = vt + Np * 32;           // V[i, l] at i * 16 + l      (lane & 15 = mode)
    static constexpr int tables = vc + Nq * 16;
    // per-warp tiles (doubles); every tile is [item][slots] with the variables of an item adjacent (128-bit accesses)
    static constexpr int PS = 6;                      // primitives (rho, V1, V2, 2p, rho/p) padded to 6
    static constexpr int x = 0;                       // [16][NC]   modal coefficients
    static constexpr int t = x + 16 * NC;             // [32][NC]   nodal values
    static constexpr int z = t + 32 * NC;             // [32][NC]   stage tile of the sum-factorised applies (only with SSE_TRI_SF_*)
    static constexpr int a_warp = z + ((SSE_TRI_SF_A || SSE_TRI_SF_B) ? 32 * NC : 0);        // pass A ends here
    static constexpr int prim = a_warp;               // [32][PS]
    static constexpr int lam = prim + 32 * PS;        // [32][4]    Lambda[m][n] at m + 2 n
    static constexpr int fprim = lam + 32 * 4;        // [32][PS]   facet nodes
    static constexpr int hnf = fprim + 32 * PS;       // [32][2]
    static constexpr int ff = hnf + 32 * 2;           // [32][NC]
    static constexpr int stage = ff + 32 * NC;        // [3][32][NC]
    static constexpr int b_warp = stage + 3 * 32 * NC;
    template <bool PASS_B> static constexpr int smem_doubles(int warps) { return tables + warps * (PASS_B ? b_warp : a_warp); }
};

// element-independent weights read through L1 with an evict-last hint (the streaming element data must not push them out)
__device__ __forceinline__ double ld_tab2(const double* p) {
    double v;
    asm("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void ld4(const double* p, double (&v)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void st4(double* p, const double (&v)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}

// copy of the dense V into the two shared orders (all threads of the CTA; followed by __syncthreads)
template <int N>
__device__ __forceinline__ void tri_fill_tables(const TriDev& t, double* sm) {
    using T = TriT<N>;
    for (int i = threadIdx.x; i < T::Np * 32; i += blockDim.x) { const int l = i >> 5, n = i & 31; sm[T::vt + i] = n < T::Nq ? t.V[n + T::Nq * l] : 0.0; }
    for (int i = threadIdx.x; i < T::Nq * 16; i += blockDim.x) { const int n = i >> 4, l = i & 15; sm[T::vc + i] = l < T::Np ? t.V[n + T::Nq * l] : 0.0; }
}
// y[e] = sum_l V[lane, l] x[l][e]            (lanes >= N_q read zeros)
template <int N>
__device__ __forceinline__ void tri_V(const double* s_vt, const double* s_x, int lane, double (&y)[4]) {
#pragma unroll
    for (int e = 0; e < 4; e++) y[e] = 0.0;
#pragma unroll
    for (int l = 0; l < TriT<N>::Np; l++) {
        const double v = s_vt[l * 32 + lane];
        double xv[4];
        ld4(s_x + l * 4, xv);
#pragma unroll
        for (int e = 0; e < 4; e++) y[e] = fma(v, xv[e], y[e]);
    }
}
// the same with the lane's row of V in registers (pass A: three applications per element)
template <int N>
__device__ __forceinline__ void tri_V_reg(const double (&vr)[TriT<N>::Np], const double* s_x, double (&y)[4]) {
#pragma unroll
    for (int e = 0; e < 4; e++) y[e] = 0.0;
#pragma unroll
    for (int l = 0; l < TriT<N>::Np; l++) {
        double xv[4];
        ld4(s_x + l * 4, xv);
#pragma unroll
        for (int e = 0; e < 4; e++) y[e] = fma(vr[l], xv[e], y[e]);
    }
}
// lane = (l, h): c[q] = sum_i V[i, l] t[i][2 h + q]            (lanes with l >= N_p read zeros)
template <int N>
__device__ __forceinline__ void tri_Vt(const double* s_vc, const double* s_t, int lane, double (&c)[2]) {
    const int l = lane & 15, h = lane >> 4;
    c[0] = 0.0; c[1] = 0.0;
#pragma unroll
    for (int i = 0; i < TriT<N>::Nq; i++) {
        const double v = s_vc[i * 16 + l];
        const double2 tv = *reinterpret_cast<const double2*>(s_t + i * 4 + 2 * h);
        c[0] = fma(v, tv.x, c[0]);
        c[1] = fma(v, tv.y, c[1]);
    }
}


// ---- sum-factorised applies (tensor_simplex.jl:84-140 in 2-D: V[(a1,a2), (b1,b2)] = A[a1,b1] B[a2,b1,b2], b1 + b2 <= p).
// MEASURED AND REJECTED (kept behind SSE_TRI_SF_A / SSE_TRI_SF_B for the A/B).  The dense products above keep pass A at 94 % of
// the shared-memory pipe with the FP64 pipe at 39 % (ncu, profiles/r2_s4_ncu_tri.csv): a broadcast LDS.128 costs 2 wavefronts
// and every multiply-add needs 10 bytes of it.  Two stages through a [32][4] tile with the 4 N coefficients of the lane in
// registers need 2/3 of the multiply-adds and half of the load instructions -- but their loads are N-address multicasts
// (the N lanes of a tensor line share an address), which cost 6 wavefronts per LDS.128 instead of 2: 720 instead of 600
// wavefronts per element, plus a second dependent round trip per apply.  Pass A 0.311 -> 0.380 ms, pass B 0.515 -> 0.564 ms.
//   V x:  z[a2][b1] = sum_b2 B[a2,b1,b2] x[b1,b2]   (lane = (a2, b1)),
//                                             y[a1][a2] = sum_b1 A[a1,b1] z[a2][b1]    (lane = node (a1, a2));
//                                      V' t: w[a2][b1] = sum_a1 A[a1,b1] t[a1][a2]     (lane = (a2, b1)),
//                                             c[b1,b2]  = sum_a2 B[a2,b1,b2] w[a2][b1] (lane = (mode, variable pair)).
template <int N> __host__ __device__ constexpr int tri_l(int b1, int b2) { return b1 * N - b1 * (b1 - 1) / 2 + b2; }
template <int N> struct TriSF {
    double Bc[N], Ac[N], At[N], Bt[N];
    int lbase, nb2, zrow, a2p, mb1;
};
template <int N>
__device__ __forceinline__ void tri_sf_init(const TriDev& t, int tn, int ml, TriSF<N>& c) {
    const int hi = tn / N, lo = tn - hi * N;          // stage-1 lane (a2', b1') = (hi, lo); node lane (a1, a2) = (hi, lo)
    c.lbase = tri_l<N>(lo, 0); c.nb2 = N - lo; c.zrow = lo * N; c.a2p = hi;
#pragma unroll
    for (int q = 0; q < N; q++) {
        c.Bc[q] = (lo + q < N) ? t.B[hi + N * (lo + N * q)] : 0.0;
        c.Ac[q] = t.A[hi + N * q];
        c.At[q] = t.A[q + N * lo];
    }
    int mb1 = 0, mb2 = 0;
    bool valid = false;
#pragma unroll
    for (int b1 = 0; b1 < N; b1++)
#pragma unroll
        for (int b2 = 0; b2 < N - b1; b2++)
            if (tri_l<N>(b1, b2) == ml) { mb1 = b1; mb2 = b2; valid = true; }
    c.mb1 = mb1;
#pragma unroll
    for (int q = 0; q < N; q++) c.Bt[q] = valid ? t.B[q + N * (mb1 + N * mb2)] : 0.0;
}
// y[e] = sum_l V[node, l] x[l][e]   (s_z: the warp's [32][4] stage tile)
template <int N>
__device__ __forceinline__ void tri_V_sf(const TriSF<N>& c, const double* s_x, double* s_z, int lane, bool node, double (&y)[4]) {
    double z[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int b2 = 0; b2 < N; b2++) {
        double xv[4];
        ld4(s_x + (c.lbase + (b2 < c.nb2 ? b2 : 0)) * 4, xv);
#pragma unroll
        for (int e = 0; e < 4; e++) z[e] = fma(c.Bc[b2], xv[e], z[e]);
    }
    if (node) st4(s_z + lane * 4, z);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 4; e++) y[e] = 0.0;
#pragma unroll
    for (int b1 = 0; b1 < N; b1++) {
        double zv[4];
        ld4(s_z + (c.zrow + b1) * 4, zv);
#pragma unroll
        for (int e = 0; e < 4; e++) y[e] = fma(c.Ac[b1], zv[e], y[e]);
    }
}
// lane = (l, h): c2[q] = sum_i V[i, l] t[i][2 h + q]
template <int N>
__device__ __forceinline__ void tri_Vt_sf(const TriSF<N>& c, const double* s_t, double* s_z, int lane, bool node, double (&c2)[2]) {
    double w[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int a1 = 0; a1 < N; a1++) {
        double tv[4];
        ld4(s_t + (a1 * N + c.a2p) * 4, tv);
#pragma unroll
        for (int e = 0; e < 4; e++) w[e] = fma(c.At[a1], tv[e], w[e]);
    }
    if (node) st4(s_z + lane * 4, w);
    __syncwarp();
    const int h = lane >> 4;
    c2[0] = 0.0; c2[1] = 0.0;
#pragma unroll
    for (int a2 = 0; a2 < N; a2++) {
        const double2 wv = *reinterpret_cast<const double2*>(s_z + (a2 * N + c.mb1) * 4 + 2 * h);
        c2[0] = fma(c.Bt[a2], wv.x, c2[0]);
        c2[1] = fma(c.Bt[a2], wv.y, c2[1]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// pass A — nodal_values! with the modal entropy projection: u_q = u(V M^-1 V' WJ w(V u)), u_f = u(R V ...)
template <int N, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_tri_nodal(TriDev t, Geo g, Law L, long long first, long long count, const double* __restrict__ u, double* __restrict__ u_q,
            double* __restrict__ u_f) {
    using T = TriT<N>;
    constexpr int D = 2, NC = 4, Nq = T::Nq, Np = T::Np, Nf = T::Nf;
    extern __shared__ __align__(16) double sm[];
    tri_fill_tables<N>(t, sm);
    __syncthreads();
    const double* s_vt = sm + T::vt;
    const double* s_vc = sm + T::vc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* s_x = sm + T::tables + warp * T::a_warp + T::x;
    double* s_t = sm + T::tables + warp * T::a_warp + T::t;
    double* s_z = sm + T::tables + warp * T::a_warp + T::z;
    const bool node = lane < Nq, fac = lane < Nf;
    const int tn = node ? lane : Nq - 1, tj = fac ? lane : Nf - 1;
    const int ml = lane & 15, mh = lane >> 4;
    const bool mode = ml < Np;
    // facet node tj = (face f, position q) extrapolates the N volume nodes fb + c fs (tensor_simplex.jl:265-268 in 2-D)
    const int ff_ = tj / N, fq = tj - ff_ * N;
    const int fb = ff_ == 0 ? fq * N : fq, fs = ff_ == 0 ? 1 : N;
    double rf[N];
#pragma unroll
    for (int c = 0; c < N; c++) rf[c] = t.rfac[tj * N + c];
    const double Wn = t.W[tn];
#if SSE_TRI_SF_A
    TriSF<N> sf;
    tri_sf_init<N>(t, tn, ml, sf);
#define TRI_V_(Y) tri_V_sf<N>(sf, s_x, s_z, lane, node, Y)
#define TRI_VT_(C) tri_Vt_sf<N>(sf, s_t, s_z, lane, node, C)
#elif SSE_TRI_VREG_A
    (void)s_z;
    double vrow[Np];
#pragma unroll
    for (int l = 0; l < Np; l++) vrow[l] = s_vt[l * 32 + lane];
#define TRI_V_(Y) tri_V_reg<N>(vrow, s_x, Y)
#define TRI_VT_(C) tri_Vt<N>(s_vc, s_t, lane, C)
#else
    (void)s_z;
#define TRI_V_(Y) tri_V<N>(s_vt, s_x, lane, Y)
#define TRI_VT_(C) tri_Vt<N>(s_vc, s_t, lane, C)
#endif
    // modal coefficients of an element: NC * Np values, two per lane, (variable e, mode l) at x = e Np + l
    const int xa = lane, xb = lane + 32;
    const bool ha = xa < NC * Np, hb = xb < NC * Np;
    const int xac = ha ? xa : 0, xbc = hb ? xb : 0;
    const int sa = (xac % Np) * 4 + xac / Np, sb = (xbc % Np) * 4 + xbc / Np;

    const long long stride = (long long)gridDim.x * WARPS, end = first + count;
    long long k = first + (long long)blockIdx.x * WARPS + warp;
    if (k >= end) return;
    double ua = u[(size_t)k * NC * Np + xac], ub = u[(size_t)k * NC * Np + xbc];
    double J = g.J_q[(size_t)k * Nq + tn], ijw = g.iJW[(size_t)k * Nq + tn];
    for (;;) {
        const long long kn = k + stride, kl = kn < end ? kn : k;
        if (ha) s_x[sa] = ua;
        if (hb) s_x[sb] = ub;
        const double Jc = J, ijwc = ijw;
        ua = u[(size_t)kl * NC * Np + xac]; ub = u[(size_t)kl * NC * Np + xbc];
        J = g.J_q[(size_t)kl * Nq + tn]; ijw = g.iJW[(size_t)kl * Nq + tn];
        __syncwarp();
        double y[NC], w[NC], c2[2];
        TRI_V_(y);                                     // u_q = V u
        if (!node) { y[0] = 1.0; y[1] = 0.0; y[2] = 0.0; y[3] = 1.0; }      // idle lanes: any physical state
        euler_cons_to_entropy_nb<D>(L.gamma, L.gm1, L.igm1, y, w);        // w_q = WJ w(u_q)   flux_differencing_form.jl:230-235
        {
            const double wj = Wn * Jc;
#pragma unroll
            for (int e = 0; e < NC; e++) w[e] *= wj;
        }
        if (node) st4(s_t + lane * 4, w);
        __syncwarp();
        TRI_VT_(c2);                                   // w = V' w_q
        if (mode) *reinterpret_cast<double2*>(s_x + ml * 4 + 2 * mh) = make_double2(c2[0], c2[1]);
        __syncwarp();
        TRI_V_(y);                                     // w = M \ w: V, diag(W / J), V'   mass_matrix.jl:185-196
#pragma unroll
        for (int e = 0; e < NC; e++) y[e] *= ijwc;
        if (node) st4(s_t + lane * 4, y);
        __syncwarp();
        TRI_VT_(c2);
        if (mode) *reinterpret_cast<double2*>(s_x + ml * 4 + 2 * mh) = make_double2(c2[0], c2[1]);
        __syncwarp();
        TRI_V_(y);                                     // w_q = V w
        if (node) st4(s_t + lane * 4, y);
        __syncwarp();
        double wf[NC], uq[NC], uf[NC];                                    // w_f = R w_q
#pragma unroll
        for (int e = 0; e < NC; e++) wf[e] = 0.0;
#pragma unroll
        for (int c = 0; c < N; c++) {
            double tv[4];
            ld4(s_t + (fb + c * fs) * 4, tv);
#pragma unroll
            for (int e = 0; e < NC; e++) wf[e] = fma(rf[c], tv[e], wf[e]);
        }
        if (!node) {
#pragma unroll
            for (int e = 0; e < NC; e++) y[e] = wf[e];                    // idle lanes: a valid entropy vector
        }
        // u_q = u(w_q), u_f = u(R w_q)                                   flux_differencing_form.jl:240-249
        euler_entropy_to_cons_nb<D>(L.gamma, L.gm1, L.igm1, L.log_gm1, y, uq);
        euler_entropy_to_cons_nb<D>(L.gamma, L.gm1, L.igm1, L.log_gm1, wf, uf);
        if (node) {
#pragma unroll
            for (int e = 0; e < NC; e++) u_q[((size_t)k * NC + e) * Nq + lane] = uq[e];
        }
        if (fac) {
#pragma unroll
            for (int e = 0; e < NC; e++) u_f[(size_t)k * Nf + lane + (size_t)g.NFT * e] = uf[e];
        }
        if (kn >= end) break;
        k = kn;
    }
}
#undef TRI_V_
#undef TRI_VT_

// ---------------------------------------------------------------------------------------------------------
// pass B — everything between the facet states of pass A and dudt of one element, in one warp
template <int N, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_tri_fluxdiff(TriDev t, Geo g, Law L, long long first, long long count, const double* __restrict__ u_q,
               const double* __restrict__ u_f, double* __restrict__ dudt, RkStage rk) {
    using T = TriT<N>;
    constexpr int D = 2, NC = 4, NP = 5, Nq = T::Nq, Np = T::Np, Nf = T::Nf, NSH = T::NSH, PS = T::PS;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) double sm[];
    tri_fill_tables<N>(t, sm);
    __syncthreads();
    const double* s_vt = sm + T::vt;
    const double* s_vc = sm + T::vc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sw_ = sm + T::tables + warp * T::b_warp;
    double* s_x = sw_ + T::x;
    double* s_t = sw_ + T::t;
    double* s_z = sw_ + T::z;
    double* s_prim = sw_ + T::prim;
    double* s_lam = sw_ + T::lam;
    double* s_fprim = sw_ + T::fprim;
    double* s_hnf = sw_ + T::hnf;
    double* s_ff = sw_ + T::ff;
    double* s_stage = sw_ + T::stage;
    const bool node = lane < Nq, fac = lane < Nf;
    const int tn = node ? lane : Nq - 1, tj = fac ? lane : Nf - 1;
    const int a1 = tn / N, a2 = tn - a1 * N;
    const int ml = lane & 15, mh = lane >> 4;
    const bool mode = ml < Np;

    // ---- per-lane constants of the schedule (the same for every element)
#ifndef SSE_TRI_WREG
#define SSE_TRI_WREG 1
#endif
#if SSE_TRI_WREG
    double sw0[NSH][D], sw1[NSH], cw[3], rw[3];
#define SW0_(s_, m_) sw0[s_][m_]
#define SW1_(s_) sw1[s_]
#define CW_(f_) cw[f_]
#define RW_(f_) rw[f_]
#else       // read where they are used: 12 L1-resident loads per element instead of 24 registers held for the lifetime of the warp
#define SW0_(s_, m_) ld_tab2(t.vS + ((size_t)(0 * NSH + (s_)) * D + (m_)) * Nq + tn)
#define SW1_(s_) ld_tab2(t.vS + ((size_t)(1 * NSH + (s_)) * D + 1) * Nq + tn)
#define CW_(f_) ld_tab2(t.fC + (f_) * Nq + tn)
#define RW_(f_) ld_tab2(t.fR + (f_) * Nq + tn)
#endif
    int jp0[NSH], js0[NSH], jp1[NSH], js1[NSH];
    bool ac0[NSH], ac1[NSH];
#pragma unroll
    for (int s = 0; s < NSH; s++) {
        const int sh = s + 1;
        const bool half = 2 * sh == N;
#if SSE_TRI_WREG
#pragma unroll
        for (int m = 0; m < D; m++) sw0[s][m] = t.vS[((size_t)(0 * NSH + s) * D + m) * Nq + tn];
        sw1[s] = t.vS[((size_t)(1 * NSH + s) * D + 1) * Nq + tn];
#endif
        int cj = a1 + sh; if (cj >= N) cj -= N;
        int cs = a1 - sh; if (cs < 0) cs += N;
        jp0[s] = tn + (cj - a1) * N; js0[s] = tn + (cs - a1) * N; ac0[s] = !(half && a1 >= sh);
        cj = a2 + sh; if (cj >= N) cj -= N;
        cs = a2 - sh; if (cs < 0) cs += N;
        jp1[s] = tn + (cj - a2); js1[s] = tn + (cs - a2); ac1[s] = !(half && a2 >= sh);
    }
    const int jf_[3] = {a1, N + a2, 2 * N + a2};              // facet partner of the lane's volume node on each face
#if SSE_TRI_WREG
#pragma unroll
    for (int f = 0; f < 3; f++) { cw[f] = t.fC[f * Nq + tn]; rw[f] = t.fR[f * Nq + tn]; }
#endif
    const double bf = t.Bf[tj];
#if SSE_TRI_SF_B
    TriSF<N> sf;
    tri_sf_init<N>(t, tn, ml, sf);
#define TRI_V_(Y) tri_V_sf<N>(sf, s_x, s_z, lane, node, Y)
#define TRI_VT_(C) tri_Vt_sf<N>(sf, s_t, s_z, lane, node, C)
#else
    (void)s_z;
#define TRI_V_(Y) tri_V<N>(s_vt, s_x, lane, Y)
#define TRI_VT_(C) tri_Vt<N>(s_vc, s_t, lane, C)
#endif
    const int ff_ = tj / N, fq = tj - ff_ * N;                // facet lane: its N volume nodes fb + c fs
    const int fb = ff_ == 0 ? fq * N : fq, fs = ff_ == 0 ? 1 : N;

    const long long stride = (long long)gridDim.x * WARPS, end = first + count;
    long long k = first + (long long)blockIdx.x * WARPS + warp;
    if (k >= end) return;

    // ---- loads of the first element
    double un[NC], lam[D][D], ui[NC], jf, njf[D], ijw;
    long long mp;
#define SSE_TRI_LOAD(K_)                                                                                                  \
    do {                                                                                                                  \
        _Pragma("unroll") for (int e = 0; e < NC; e++) un[e] = __ldcs(u_q + ((size_t)(K_) * NC + e) * Nq + tn);           \
        _Pragma("unroll") for (int n = 0; n < D; n++)                                                                     \
            _Pragma("unroll") for (int m = 0; m < D; m++) lam[m][n] = __ldcs(g.Lambda_q + ((size_t)(K_) * D * D + (m + D * n)) * Nq + tn); \
        _Pragma("unroll") for (int e = 0; e < NC; e++) ui[e] = u_f[(size_t)(K_) * Nf + tj + (size_t)g.NFT * e];           \
        jf = __ldcs(g.J_f + (size_t)(K_) * Nf + tj);                                                                      \
        _Pragma("unroll") for (int m = 0; m < D; m++) njf[m] = __ldcs(g.nJf + m + D * ((size_t)(K_) * Nf + tj));          \
        mp = g.mapP[(size_t)(K_) * Nf + tj];                                                                              \
        ijw = __ldcs(g.iJW + (size_t)(K_) * Nq + tn);                                                                     \
    } while (0)
    SSE_TRI_LOAD(k);

    for (;;) {
        const long long kn = k + stride, kl = kn < end ? kn : k;
        // ---- neighbour gather of this element, then the current values move aside and the next element's loads go out
        double uo[NC];
        {
            const size_t jo = (size_t)(mp - 1);
#pragma unroll
            for (int e = 0; e < NC; e++) uo[e] = u_f[jo + (size_t)g.NFT * e];
        }
        double cu[NC], cl[D][D], ci[NC], cnj[D];
#pragma unroll
        for (int e = 0; e < NC; e++) { cu[e] = un[e]; ci[e] = ui[e]; }
#pragma unroll
        for (int m = 0; m < D; m++) { cnj[m] = njf[m]; cl[m][0] = lam[m][0]; cl[m][1] = lam[m][1]; }
        const double cjf = jf, cijw = ijw;
        SSE_TRI_LOAD(kl);

        // ---- node primitives and metrics into the warp's tiles
        double qi[NP], r[NC];
        to_prim_fast<D>(L, cu, qi);
        if (node) {
            double* p = s_prim + lane * PS;
            *reinterpret_cast<double2*>(p) = make_double2(qi[0], qi[1]);
            *reinterpret_cast<double2*>(p + 2) = make_double2(qi[2], qi[3]);
            p[4] = qi[4];
            double* q = s_lam + lane * 4;
            *reinterpret_cast<double2*>(q) = make_double2(cl[0][0], cl[1][0]);
            *reinterpret_cast<double2*>(q + 2) = make_double2(cl[0][1], cl[1][1]);
        }
#pragma unroll
        for (int e = 0; e < NC; e++) r[e] = 0.0;
        __syncwarp();

        // ---- volume term (flux_difference!, flux_differencing_form.jl:37-75): line direction 0 (a1, both S_m), then 1 (a2, S_2 only)
#pragma unroll
        for (int l = 0; l < D; l++) {
            double gv[NSH][D], qj[NSH][NP], ph[NSH][NC];
#pragma unroll
            for (int s = 0; s < NSH; s++) {
                const int j = l == 0 ? jp0[s] : jp1[s];
                const double* p = s_prim + j * PS;
                const double2 p01 = *reinterpret_cast<const double2*>(p), p23 = *reinterpret_cast<const double2*>(p + 2);
                qj[s][0] = p01.x; qj[s][1] = p01.y; qj[s][2] = p23.x; qj[s][3] = p23.y; qj[s][4] = p[4];
                if (l == 0) {
                    double lj[4];
                    ld4(s_lam + j * 4, lj);
#pragma unroll
                    for (int n = 0; n < D; n++)
                        gv[s][n] = fma(SW0_(s, 1), cl[1][n] + lj[1 + 2 * n], SW0_(s, 0) * (cl[0][n] + lj[0 + 2 * n]));
                } else {
                    const double l10 = s_lam[j * 4 + 1], l11 = s_lam[j * 4 + 3];
                    const double w1 = SW1_(s);
                    gv[s][0] = w1 * (cl[1][0] + l10);
                    gv[s][1] = w1 * (cl[1][1] + l11);
                }
            }
            if constexpr (NSH == 2) ec_contract_scaled2<D>(L, qi, qj[0], qj[1], gv[0], gv[1], ph[0], ph[1]);
            else ec_contract_scaled<D>(L, qi, qj[0], gv[0], ph[0]);
#pragma unroll
            for (int s = 0; s < NSH; s++) {
                const bool ac = l == 0 ? ac0[s] : ac1[s];
                const int src = l == 0 ? js0[s] : js1[s];
#pragma unroll
                for (int e = 0; e < NC; e++) {
                    const double pz = ac ? ph[s][e] : 0.0;                // even N: the half-way pairs once only
                    r[e] += __shfl_sync(FULL, pz, src) - pz;               // -phi stays, +phi goes to the partner along the line
                }
            }
        }

        // ---- interface flux (numerical_flux!, ConservationLaws.jl:75-128) on the facet lanes
        double ffv[NC];
        {
            double qa[NP], qb[NP], nf[D], nfq[D], phi[NC];
            const double ijf = rcp_fast(cjf);
#pragma unroll
            for (int m = 0; m < D; m++) { nf[m] = cnj[m] * ijf; nfq[m] = 0.25 * nf[m]; }        // n_f = nJf / J_f   operators.jl:59
            const double ira = to_prim_fast<D>(L, ci, qa);
            const double irb = to_prim_fast<D>(L, uo, qb);
            ec_contract_scaled<D>(L, qa, qb, nfq, phi);
            if (L.inviscid == SSE_FLUX_LAX_FRIEDRICHS) {
                double vni = 0.0, vno = 0.0;
#pragma unroll
                for (int m = 0; m < D; m++) { vni = fma(qa[1 + m], nf[m], vni); vno = fma(qb[1 + m], nf[m], vno); }
                const double cm = sqrt(fmax(L.gamma * (0.5 * qa[D + 1]) * ira, L.gamma * (0.5 * qb[D + 1]) * irb));
                const double a = L.half_lambda * (fmax(fabs(vni), fabs(vno)) + cm);
#pragma unroll
                for (int e = 0; e < NC; e++) phi[e] = fma(a, ci[e] - uo[e], phi[e]);
            }
            const double bj = bf * cjf;                                   // BJf               operators.jl:58
#pragma unroll
            for (int e = 0; e < NC; e++) ffv[e] = bj * phi[e];
            if (fac) {
                double* p = s_fprim + lane * PS;
                *reinterpret_cast<double2*>(p) = make_double2(qa[0], qa[1]);
                *reinterpret_cast<double2*>(p + 2) = make_double2(qa[2], qa[3]);
                p[4] = qa[4];
                *reinterpret_cast<double2*>(s_hnf + lane * 2) = make_double2(cnj[0], cnj[1]);      // 2 halfnJf (operators.jl:78); the 1/2 lives in fC
            }
        }
        __syncwarp();

        // ---- facet correction (facet_correction!, flux_differencing_form.jl:126-168): one pair per node and face
        {
            double gv[3][D], qj[3][NP], ph[3][NC];
#pragma unroll
            for (int f = 0; f < 3; f++) {
                const int j = jf_[f];
                double hq[D];
                if (g.nJq) {
#pragma unroll
                    for (int n = 0; n < D; n++) hq[n] = g.nJq[n + D * (f + (size_t)3 * (tn + (size_t)Nq * k))];
                } else {
#pragma unroll
                    for (int n = 0; n < D; n++) hq[n] = fma(cl[1][n], t.nref[1 + D * f], cl[0][n] * t.nref[0 + D * f]);   // mesh.jl:262-269
                }
                const double2 hn = *reinterpret_cast<const double2*>(s_hnf + j * 2);
                const double cwf = CW_(f);
                gv[f][0] = cwf * (hn.x + hq[0]);
                gv[f][1] = cwf * (hn.y + hq[1]);
                const double* p = s_fprim + j * PS;
                const double2 p01 = *reinterpret_cast<const double2*>(p), p23 = *reinterpret_cast<const double2*>(p + 2);
                qj[f][0] = p01.x; qj[f][1] = p01.y; qj[f][2] = p23.x; qj[f][3] = p23.y; qj[f][4] = p[4];
            }
            ec_contract_scaled2<D>(L, qi, qj[0], qj[1], gv[0], gv[1], ph[0], ph[1]);
            ec_contract_scaled<D>(L, qi, qj[2], gv[2], ph[2]);
#pragma unroll
            for (int f = 0; f < 3; f++) {
#pragma unroll
                for (int e = 0; e < NC; e++) r[e] -= ph[f][e];
                if (node) st4(s_stage + (f * 32 + lane) * 4, ph[f]);
            }
        }
        __syncwarp();
        {   // facet lanes: f_f = BJf f* - sum of the staged vectors of the facet node's N volume nodes
            double s[NC];
#pragma unroll
            for (int e = 0; e < NC; e++) s[e] = 0.0;
#pragma unroll
            for (int c = 0; c < N; c++) {
                double v[4];
                ld4(s_stage + (ff_ * 32 + fb + c * fs) * 4, v);
#pragma unroll
                for (int e = 0; e < NC; e++) s[e] += v[e];
            }
#pragma unroll
            for (int e = 0; e < NC; e++) ffv[e] -= s[e];
            if (fac) st4(s_ff + lane * 4, ffv);
        }
        __syncwarp();
        // ---- lift: r_q -= R' f_f (flux_differencing_form.jl:341-342)
#pragma unroll
        for (int f = 0; f < 3; f++) {
            double v[4];
            ld4(s_ff + jf_[f] * 4, v);
            const double rwf = RW_(f);
#pragma unroll
            for (int e = 0; e < NC; e++) r[e] = fma(-rwf, v[e], r[e]);
        }
        // ---- dudt = M^-1 V' r_q: V', V, diag(W / J), V' (flux_differencing_form.jl:345-346, mass_matrix.jl:185-196)
        if (node) st4(s_t + lane * 4, r);
        __syncwarp();
        double c2[2], y[NC];
        TRI_VT_(c2);
        if (mode) *reinterpret_cast<double2*>(s_x + ml * 4 + 2 * mh) = make_double2(c2[0], c2[1]);
        __syncwarp();
        TRI_V_(y);
#pragma unroll
        for (int e = 0; e < NC; e++) y[e] *= cijw;
        if (node) st4(s_t + lane * 4, y);
        __syncwarp();
        TRI_VT_(c2);
        if (mode) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const size_t idx = ((size_t)k * NC + 2 * mh + q) * Np + ml;
                dudt[idx] = c2[q];
                flag_nonfinite(g.flag, c2[q]);
                if (rk.u) {                        // fused 2N-storage RK stage (Carpenter & Kennedy 1994)
                    const double tm = fma(rk.A, rk.tmp[idx], rk.dt * c2[q]);
                    rk.tmp[idx] = tm;
                    rk.u[idx] = fma(rk.B, tm, rk.u[idx]);
                }
            }
        }
        if (kn >= end) break;
        k = kn;
        __syncwarp();
    }
#undef SSE_TRI_LOAD
#undef SW0_
#undef SW1_
#undef CW_
#undef RW_
#undef TRI_V_
#undef TRI_VT_
}

// ---------------------------------------------------------------------------------------------------------
// pass B, second form: the same arithmetic with every global load of the element loop issued as an asynchronous copy into
// the warp's own tiles (LDGSTS: no registers are held while the data of the NEXT element is in flight) and the pair
// weights read from their L1-resident tables where they are used.  k_tri_fluxdiff keeps 17 doubles of prefetched data and
// 12 weights per lane in registers and spills at the 128-register cap of 16 warps per SM; this form does not.
template <int N> struct TriB {
    using T = TriT<N>;
    static constexpr int NS = 8;                          // node slot: u_q[0..3], Lambda[m][n] at 4 + m + 2 n
    static constexpr int FSL = 8;                         // facet slot: u_f[0..3], J_f, nJf[0..1], mapP (raw 64 bits)
    static constexpr int nb = 0;                          // [2][32][NS]
    static constexpr int fbuf = nb + 2 * 32 * NS;         // [2][16][FSL]
    static constexpr int gath = fbuf + 2 * 16 * FSL;      // [16][4]      neighbour states
    static constexpr int fprim = gath + 16 * 4;           // [16][PS]
    static constexpr int ff = fprim + 16 * T::PS;         // [16][4]
    static constexpr int un = ff + 16 * 4;                // union: prim [32][PS] (volume term) | stage [3][32][4] (facet term) |
    static constexpr int prim = un;                       //        t [32][4] + x [16][4] (projection)
    static constexpr int stage = un;
    static constexpr int t = un;
    static constexpr int x = un + 32 * 4;
    static constexpr int warp = un + 3 * 32 * 4;
    static_assert(T::Nf <= 16, "facet tiles hold 16 nodes");
    static constexpr int smem_doubles(int warps) { return T::tables + warps * warp; }
};

template <int N, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_tri_fluxdiff_async(TriDev t, Geo g, Law L, long long first, long long count, const double* __restrict__ u_q,
                     const double* __restrict__ u_f, double* __restrict__ dudt, RkStage rk) {
    using T = TriT<N>;
    using B = TriB<N>;
    constexpr int D = 2, NC = 4, NP = 5, Nq = T::Nq, Np = T::Np, Nf = T::Nf, NSH = T::NSH, PS = T::PS, NS = B::NS, FSL = B::FSL;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) double sm[];
    tri_fill_tables<N>(t, sm);
    __syncthreads();
    const double* s_vt = sm + T::vt;
    const double* s_vc = sm + T::vc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sw_ = sm + T::tables + warp * B::warp;
    double* s_nb = sw_ + B::nb;
    double* s_fb = sw_ + B::fbuf;
    double* s_gath = sw_ + B::gath;
    double* s_fprim = sw_ + B::fprim;
    double* s_ff = sw_ + B::ff;
    double* s_prim = sw_ + B::prim;
    double* s_stage = sw_ + B::stage;
    double* s_t = sw_ + B::t;
    double* s_x = sw_ + B::x;
    const bool node = lane < Nq, fac = lane < Nf;
    const int tn = node ? lane : Nq - 1, tj = fac ? lane : Nf - 1;
    const int a1 = tn / N, a2 = tn - a1 * N;
    const int ml = lane & 15, mh = lane >> 4;
    const bool mode = ml < Np;
    int jp0[NSH], js0[NSH], jp1[NSH], js1[NSH];
    bool ac0[NSH], ac1[NSH];
#pragma unroll
    for (int s = 0; s < NSH; s++) {
        const int sh = s + 1;
        const bool half = 2 * sh == N;
        int cj = a1 + sh; if (cj >= N) cj -= N;
        int cs = a1 - sh; if (cs < 0) cs += N;
        jp0[s] = tn + (cj - a1) * N; js0[s] = tn + (cs - a1) * N; ac0[s] = !(half && a1 >= sh);
        cj = a2 + sh; if (cj >= N) cj -= N;
        cs = a2 - sh; if (cs < 0) cs += N;
        jp1[s] = tn + (cj - a2); js1[s] = tn + (cs - a2); ac1[s] = !(half && a2 >= sh);
    }
    const int jf_[3] = {a1, N + a2, 2 * N + a2};
    const double bf = t.Bf[tj];
    const int ff_ = tj / N, fq = tj - ff_ * N;
    const int fb = ff_ == 0 ? fq * N : fq, fs = ff_ == 0 ? 1 : N;

    const long long stride = (long long)gridDim.x * WARPS, end = first + count;
    long long k = first + (long long)blockIdx.x * WARPS + warp;
    if (k >= end) return;

    auto issue = [&](long long K, int bb) {               // node and facet data of element K into buffer bb
        if (node) {
            double* d = s_nb + (bb * 32 + lane) * NS;
#pragma unroll
            for (int e = 0; e < NC; e++) cp_async8(d + e, u_q + ((size_t)K * NC + e) * Nq + lane);
#pragma unroll
            for (int c = 0; c < D * D; c++) cp_async8(d + 4 + c, g.Lambda_q + ((size_t)K * D * D + c) * Nq + lane);
        }
        if (fac) {
            double* d = s_fb + (bb * 16 + lane) * FSL;
#pragma unroll
            for (int e = 0; e < NC; e++) cp_async8(d + e, u_f + (size_t)K * Nf + lane + (size_t)g.NFT * e);
            cp_async8(d + 4, g.J_f + (size_t)K * Nf + lane);
#pragma unroll
            for (int m = 0; m < D; m++) cp_async8(d + 5 + m, g.nJf + m + D * ((size_t)K * Nf + lane));
            cp_async8(d + 7, reinterpret_cast<const double*>(g.mapP + (size_t)K * Nf + lane));
        }
        cp_async_commit();
    };
    issue(k, 0);
    int b = 0;
    for (;;) {
        const long long kn = k + stride, kl = kn < end ? kn : k;
        cp_async_wait<0>();
        __syncwarp();
        if (fac) {                                       // neighbour gather of this element (consumed after the volume term)
            const size_t jo = (size_t)(__double_as_longlong(s_fb[(b * 16 + lane) * FSL + 7]) - 1);
#pragma unroll
            for (int e = 0; e < NC; e++) cp_async8(s_gath + lane * 4 + e, u_f + jo + (size_t)g.NFT * e);
        }
        cp_async_commit();
        issue(kl, b ^ 1);
        const double* nbc = s_nb + b * 32 * NS;
        const double* fbc = s_fb + b * 16 * FSL;
        const double cijw = __ldcs(g.iJW + (size_t)k * Nq + tn);

        double cu[NC], cl4[4], qi[NP], r[NC];
        ld4(nbc + tn * NS, cu);
        ld4(nbc + tn * NS + 4, cl4);                      // cl4[m + 2 n] = Lambda[m][n]
        to_prim_fast<D>(L, cu, qi);
        if (node) {
            double* p = s_prim + lane * PS;
            *reinterpret_cast<double2*>(p) = make_double2(qi[0], qi[1]);
            *reinterpret_cast<double2*>(p + 2) = make_double2(qi[2], qi[3]);
            p[4] = qi[4];
        }
#pragma unroll
        for (int e = 0; e < NC; e++) r[e] = 0.0;
        __syncwarp();

        // ---- volume term
#pragma unroll
        for (int l = 0; l < D; l++) {
            double gv[NSH][D], qj[NSH][NP], ph[NSH][NC];
#pragma unroll
            for (int s = 0; s < NSH; s++) {
                const int j = l == 0 ? jp0[s] : jp1[s];
                const double* p = s_prim + j * PS;
                const double2 p01 = *reinterpret_cast<const double2*>(p), p23 = *reinterpret_cast<const double2*>(p + 2);
                qj[s][0] = p01.x; qj[s][1] = p01.y; qj[s][2] = p23.x; qj[s][3] = p23.y; qj[s][4] = p[4];
                double lj[4];
                ld4(nbc + j * NS + 4, lj);
                if (l == 0) {
                    const double w0 = ld_tab2(t.vS + ((size_t)(0 * NSH + s) * D + 0) * Nq + tn), w1 = ld_tab2(t.vS + ((size_t)(0 * NSH + s) * D + 1) * Nq + tn);
#pragma unroll
                    for (int n = 0; n < D; n++) gv[s][n] = fma(w1, cl4[1 + 2 * n] + lj[1 + 2 * n], w0 * (cl4[0 + 2 * n] + lj[0 + 2 * n]));
                } else {
                    const double w1 = ld_tab2(t.vS + ((size_t)(1 * NSH + s) * D + 1) * Nq + tn);
                    gv[s][0] = w1 * (cl4[1] + lj[1]);
                    gv[s][1] = w1 * (cl4[3] + lj[3]);
                }
            }
            if constexpr (NSH == 2) ec_contract_scaled2<D>(L, qi, qj[0], qj[1], gv[0], gv[1], ph[0], ph[1]);
            else ec_contract_scaled<D>(L, qi, qj[0], gv[0], ph[0]);
#pragma unroll
            for (int s = 0; s < NSH; s++) {
                const bool ac = l == 0 ? ac0[s] : ac1[s];
                const int src = l == 0 ? js0[s] : js1[s];
#pragma unroll
                for (int e = 0; e < NC; e++) {
                    const double pz = ac ? ph[s][e] : 0.0;
                    r[e] += __shfl_sync(FULL, pz, src) - pz;
                }
            }
        }

        // ---- interface flux on the facet lanes
        cp_async_wait<1>();                               // the gather has landed (the next element's data may still be in flight)
        __syncwarp();
        double ffv[NC];
        {
            double ci[NC], uo[NC], fm[4], qa[NP], qb[NP], nf[D], nfq[D], phi[NC];
            ld4(fbc + tj * FSL, ci);
            ld4(fbc + tj * FSL + 4, fm);                  // J_f, nJf[0], nJf[1], mapP
            ld4(s_gath + tj * 4, uo);
            const double ijf = rcp_fast(fm[0]);
            nf[0] = fm[1] * ijf; nf[1] = fm[2] * ijf;
#pragma unroll
            for (int m = 0; m < D; m++) nfq[m] = 0.25 * nf[m];
            const double ira = to_prim_fast<D>(L, ci, qa);
            const double irb = to_prim_fast<D>(L, uo, qb);
            ec_contract_scaled<D>(L, qa, qb, nfq, phi);
            if (L.inviscid == SSE_FLUX_LAX_FRIEDRICHS) {
                double vni = 0.0, vno = 0.0;
#pragma unroll
                for (int m = 0; m < D; m++) { vni = fma(qa[1 + m], nf[m], vni); vno = fma(qb[1 + m], nf[m], vno); }
                const double cm = sqrt(fmax(L.gamma * (0.5 * qa[D + 1]) * ira, L.gamma * (0.5 * qb[D + 1]) * irb));
                const double a = L.half_lambda * (fmax(fabs(vni), fabs(vno)) + cm);
#pragma unroll
                for (int e = 0; e < NC; e++) phi[e] = fma(a, ci[e] - uo[e], phi[e]);
            }
            const double bj = bf * fm[0];
#pragma unroll
            for (int e = 0; e < NC; e++) ffv[e] = bj * phi[e];
            if (fac) {
                double* p = s_fprim + lane * PS;
                *reinterpret_cast<double2*>(p) = make_double2(qa[0], qa[1]);
                *reinterpret_cast<double2*>(p + 2) = make_double2(qa[2], qa[3]);
                p[4] = qa[4];
            }
        }
        __syncwarp();

        // ---- facet correction: one pair per node and face
        {
            double gv[3][D], qj[3][NP], ph[3][NC];
#pragma unroll
            for (int f = 0; f < 3; f++) {
                const int j = jf_[f];
                double hq[D];
                if (g.nJq) {
#pragma unroll
                    for (int n = 0; n < D; n++) hq[n] = g.nJq[n + D * (f + (size_t)3 * (tn + (size_t)Nq * k))];
                } else {
#pragma unroll
                    for (int n = 0; n < D; n++) hq[n] = fma(cl4[1 + 2 * n], t.nref[1 + D * f], cl4[0 + 2 * n] * t.nref[0 + D * f]);
                }
                const double cwf = ld_tab2(t.fC + f * Nq + tn);
                gv[f][0] = cwf * (fbc[j * FSL + 5] + hq[0]);
                gv[f][1] = cwf * (fbc[j * FSL + 6] + hq[1]);
                const double* p = s_fprim + j * PS;
                const double2 p01 = *reinterpret_cast<const double2*>(p), p23 = *reinterpret_cast<const double2*>(p + 2);
                qj[f][0] = p01.x; qj[f][1] = p01.y; qj[f][2] = p23.x; qj[f][3] = p23.y; qj[f][4] = p[4];
            }
            ec_contract_scaled2<D>(L, qi, qj[0], qj[1], gv[0], gv[1], ph[0], ph[1]);
            ec_contract_scaled<D>(L, qi, qj[2], gv[2], ph[2]);
#pragma unroll
            for (int f = 0; f < 3; f++) {
#pragma unroll
                for (int e = 0; e < NC; e++) r[e] -= ph[f][e];
                if (node) st4(s_stage + (f * 32 + lane) * 4, ph[f]);      // over prim: every lane is past the volume term
            }
        }
        __syncwarp();
        {
            double s[NC];
#pragma unroll
            for (int e = 0; e < NC; e++) s[e] = 0.0;
#pragma unroll
            for (int c = 0; c < N; c++) {
                double v[4];
                ld4(s_stage + (ff_ * 32 + fb + c * fs) * 4, v);
#pragma unroll
                for (int e = 0; e < NC; e++) s[e] += v[e];
            }
#pragma unroll
            for (int e = 0; e < NC; e++) ffv[e] -= s[e];
            if (fac) st4(s_ff + lane * 4, ffv);
        }
        __syncwarp();
#pragma unroll
        for (int f = 0; f < 3; f++) {
            double v[4];
            ld4(s_ff + jf_[f] * 4, v);
            const double rwf = ld_tab2(t.fR + f * Nq + tn);
#pragma unroll
            for (int e = 0; e < NC; e++) r[e] = fma(-rwf, v[e], r[e]);
        }
        // ---- dudt = M^-1 V' r_q  (t and x lie over the stages: every reducer is past them)
        if (node) st4(s_t + lane * 4, r);
        __syncwarp();
        double c2[2], y[NC];
        tri_Vt<N>(s_vc, s_t, lane, c2);
        if (mode) *reinterpret_cast<double2*>(s_x + ml * 4 + 2 * mh) = make_double2(c2[0], c2[1]);
        __syncwarp();
        tri_V<N>(s_vt, s_x, lane, y);
#pragma unroll
        for (int e = 0; e < NC; e++) y[e] *= cijw;
        if (node) st4(s_t + lane * 4, y);
        __syncwarp();
        tri_Vt<N>(s_vc, s_t, lane, c2);
        if (mode) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const size_t idx = ((size_t)k * NC + 2 * mh + q) * Np + ml;
                dudt[idx] = c2[q];
                flag_nonfinite(g.flag, c2[q]);
                if (rk.u) {
                    const double tm = fma(rk.A, rk.tmp[idx], rk.dt * c2[q]);
                    rk.tmp[idx] = tm;
                    rk.u[idx] = fma(rk.B, tm, rk.u[idx]);
                }
            }
        }
        if (kn >= end) break;
        k = kn;
        b ^= 1;
        __syncwarp();
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------------------
// BASELINE config 1: 2-D linear advection, StandardForm + ReferenceOperators on collapsed triangles
// (standard_form_first_order.jl:16-63), the same warp-per-element mapping with one variable.  The flux is linear, so the
// d^2 (D_m, D_m') pairs collapse to d pairs on g_m = c_m u, c_m = sum_n (W Lambda_mn / 2) a_n, and the facet difference
// sum_n halfN_n R f_n is (a.n)/2 u_f (the collapse of kernels_adv.cuh for the tetrahedron).
//   k_tri_adv_facets  pass A: u_f = (R V) u -- one 15-term dot product per facet lane; the modal coefficients are copied into
//                             the u_q scratch (pass B must not depend on the caller's state staying alive and unchanged)
//   k_tri_adv         pass B: u_q = V u (recomputed from the lane's row of V), volume terms along the two tensor lines with
//                             the 1-D derivative rows / columns of the lane in registers, interface flux, lift, V', mass solve
// Per lane and for the lifetime of the warp: row of V (N_p), a half column of V (V' sums split over the two half-warps and
// joined by one shuffle), 4 N derivative coefficients, three lift weights.
template <int N> struct TriA {
    using T = TriT<N>;
    static constexpr int HC = (T::Nq + 1) / 2;            // nodes per half-warp in the V' sums
    static constexpr int x = 0;                           // [16]  modal coefficients
    static constexpr int u = x + 16;                      // [32]  nodal values
    static constexpr int g = u + 32;                      // [2][32]
    static constexpr int ff = g + 64;                     // [16]
    static constexpr int warp = ff + 16;
};

template <int N, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_tri_adv_facets(TriDev t, Geo g, long long first, long long count, const double* __restrict__ u, double* __restrict__ u_f,
                 double* __restrict__ um) {
    using T = TriT<N>;
    constexpr int Np = T::Np, Nf = T::Nf;
    __shared__ double s_x[WARPS][16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool fac = lane < Nf, md = lane < Np;
    const int tj = fac ? lane : Nf - 1;
    double rv[Np];
#pragma unroll
    for (int l = 0; l < Np; l++) rv[l] = t.RV[tj + Nf * l];
    const long long stride = (long long)gridDim.x * WARPS, end = first + count;
    for (long long k = first + (long long)blockIdx.x * WARPS + warp; k < end; k += stride) {
        if (md) {
            const double x = u[(size_t)k * Np + lane];
            s_x[warp][lane] = x;
            um[(size_t)k * Np + lane] = x;             // pass B reads the coefficients from the handle's scratch, not from the caller's state
        }
        __syncwarp();
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < Np; l++) s = fma(rv[l], s_x[warp][l], s);
        if (fac) u_f[(size_t)k * Nf + lane] = s;
        __syncwarp();
    }
    (void)g;
}

template <int N, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_tri_adv(TriDev t, Geo g, Law L, long long first, long long count, const double* u, const double* __restrict__ u_f,
          double* __restrict__ dudt, RkStage rk) {
    using T = TriT<N>;
    using A = TriA<N>;
    constexpr int D = 2, Nq = T::Nq, Np = T::Np, Nf = T::Nf, HC = A::HC;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sw_ = sm + warp * A::warp;
    double* s_x = sw_ + A::x;
    double* s_u = sw_ + A::u;
    double* s_g = sw_ + A::g;
    double* s_ff = sw_ + A::ff;
    const bool node = lane < Nq, fac = lane < Nf;
    const int tn = node ? lane : Nq - 1, tj = fac ? lane : Nf - 1;
    const int a1 = tn / N, a2 = tn - a1 * N;
    const int ml = lane & 15, mh = lane >> 4;
    const bool mode = ml < Np;
    // ---- per-lane constants
    double vrow[Np], vcol[HC], dr[D][N], dc[D][N], rw[3];
#pragma unroll
    for (int l = 0; l < Np; l++) vrow[l] = t.V[tn + Nq * l];
#pragma unroll
    for (int q = 0; q < HC; q++) { const int i = mh * HC + q; vcol[q] = (mode && i < Nq) ? t.V[i + Nq * ml] : 0.0; }
#pragma unroll
    for (int q = 0; q < N; q++) {
        dr[0][q] = t.D1[0 * N * N + a1 + N * q]; dc[0][q] = t.D1[0 * N * N + q + N * a1];     // D_1D[row, col] at row + N col
        dr[1][q] = t.D1[1 * N * N + a2 + N * q]; dc[1][q] = t.D1[1 * N * N + q + N * a2];
    }
    const int jf_[3] = {a1, N + a2, 2 * N + a2};
#pragma unroll
    for (int f = 0; f < 3; f++) rw[f] = t.fR[f * Nq + tn];
    const double hw = 0.5 * t.W[tn], bf = t.Bf[tj];
    const int base0 = a2, base1 = a1 * N;                  // first node of the lane's a1-line (stride N) / a2-line (stride 1)

    const long long stride = (long long)gridDim.x * WARPS, end = first + count;
    long long k = first + (long long)blockIdx.x * WARPS + warp;
    if (k >= end) return;
    double xm, lam[D][D], ui, jf, njf[D], ijw;
    long long mp;
#define SSE_TRIA_LOAD(K_)                                                                                                 \
    do {                                                                                                                  \
        xm = u[(size_t)(K_) * Np + (mode ? ml : 0)];                                                                      \
        _Pragma("unroll") for (int n = 0; n < D; n++)                                                                     \
            _Pragma("unroll") for (int m = 0; m < D; m++) lam[m][n] = __ldcs(g.Lambda_q + ((size_t)(K_) * D * D + (m + D * n)) * Nq + tn); \
        ui = u_f[(size_t)(K_) * Nf + tj];                                                                                 \
        jf = __ldcs(g.J_f + (size_t)(K_) * Nf + tj);                                                                      \
        _Pragma("unroll") for (int m = 0; m < D; m++) njf[m] = __ldcs(g.nJf + m + D * ((size_t)(K_) * Nf + tj));          \
        mp = g.mapP[(size_t)(K_) * Nf + tj];                                                                              \
        ijw = __ldcs(g.iJW + (size_t)(K_) * Nq + tn);                                                                     \
    } while (0)
    SSE_TRIA_LOAD(k);
    for (;;) {
        const long long kn = k + stride, kl = kn < end ? kn : k;
        const double uo = u_f[(size_t)(mp - 1)];
        if (lane < 16) s_x[lane] = mode ? xm : 0.0;
        double c[D];
#pragma unroll
        for (int m = 0; m < D; m++) c[m] = fma(hw * lam[m][1], L.a[1], (hw * lam[m][0]) * L.a[0]);      // halfWLambda_mn a_n
        const double ci = ui, cjf = jf, cn0 = njf[0], cn1 = njf[1], cijw = ijw;
        SSE_TRIA_LOAD(kl);
        __syncwarp();
        // u_q = V u
        double uq = 0.0;
#pragma unroll
        for (int l = 0; l < Np; l++) uq = fma(vrow[l], s_x[l], uq);
        if (node) { s_u[lane] = uq; s_g[lane] = c[0] * uq; s_g[32 + lane] = c[1] * uq; }
        {   // interface flux, f_f = BJf (f* - (a.n)/2 u_f)           standard_form_first_order.jl:48-58
            const double ijf = rcp_fast(cjf);
            const double an = fma(L.a[1], cn1 * ijf, L.a[0] * (cn0 * ijf));
            double fs = (0.5 * (ci + uo)) * an;
            if (L.inviscid == SSE_FLUX_LAX_FRIEDRICHS) fs = fma(L.half_lambda * fabs(an), ci - uo, fs);
            if (fac) s_ff[lane] = (bf * cjf) * (fs - 0.5 * an * ci);
        }
        __syncwarp();
        // volume terms: r = sum_m D_m' (c_m u) - c_m (D_m u)            standard_form_first_order.jl:38-46
        double r = 0.0;
        {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int q = 0; q < N; q++) { s1 = fma(dc[0][q], s_g[base0 + q * N], s1); s2 = fma(dr[0][q], s_u[base0 + q * N], s2); }
            r += s1 - c[0] * s2;
            s1 = 0.0; s2 = 0.0;
#pragma unroll
            for (int q = 0; q < N; q++) { s1 = fma(dc[1][q], s_g[32 + base1 + q], s1); s2 = fma(dr[1][q], s_u[base1 + q], s2); }
            r += s1 - c[1] * s2;
        }
#pragma unroll
        for (int f = 0; f < 3; f++) r = fma(-rw[f], s_ff[jf_[f]], r);                               // - R' f_f
        __syncwarp();                                      // every lane is past s_u / s_g
        if (node) s_u[lane] = r;
        __syncwarp();
        // dudt = M^-1 V' r: V' (two half-warp partial sums joined by a shuffle), V, diag(W / J), V'
        double p = 0.0;
#pragma unroll
        for (int q = 0; q < HC; q++) { const int i = mh * HC + q; p = fma(vcol[q], s_u[i < Nq ? i : Nq - 1], p); }
        p += __shfl_xor_sync(FULL, p, 16);
        if (lane < 16) s_x[lane] = mode ? p : 0.0;
        __syncwarp();
        double y = 0.0;
#pragma unroll
        for (int l = 0; l < Np; l++) y = fma(vrow[l], s_x[l], y);
        y *= cijw;
        __syncwarp();                                      // the first V' has read s_u everywhere (the shuffle ordered the warp)
        if (node) s_u[lane] = y;
        __syncwarp();
        p = 0.0;
#pragma unroll
        for (int q = 0; q < HC; q++) { const int i = mh * HC + q; p = fma(vcol[q], s_u[i < Nq ? i : Nq - 1], p); }
        p += __shfl_xor_sync(FULL, p, 16);
        if (lane < Np) {
            const size_t idx = (size_t)k * Np + lane;
            dudt[idx] = p;
            flag_nonfinite(g.flag, p);
            if (rk.u) {
                const double tm = fma(rk.A, rk.tmp[idx], rk.dt * p);
                rk.tmp[idx] = tm;
                rk.u[idx] = fma(rk.B, tm, rk.u[idx]);
            }
        }
        if (kn >= end) break;
        k = kn;
        __syncwarp();
    }
#undef SSE_TRIA_LOAD
}

}  // namespace sse
