// ct_kernels.cu — instantiations of the compile-time-sized kernels (N = p + 1 = 3 .. 8)
#include "kernels_ct.cuh"
#include "kernels_adv.cuh"
#include "kernels_tri.cuh"

namespace sse {

// polynomial degrees with compile-time kernels: N = p + 1 = 3 .. 8 (p = 7: the reference's examples/advection_3d.ipynb)
#define SSE_CT_DISPATCH(N_, CALL)                                                                   \
    switch (N_) { case 3: CALL(3); break; case 4: CALL(4); break; case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; case 8: CALL(8); break; default: break; }
static bool ct_size_ok(int N) {                                // SSE_CT_NMAX: A/B switch (larger N fall back to the run-time kernels)
    int hi = 8;
    if (const char* e = getenv("SSE_CT_NMAX")) hi = atoi(e);
    return N >= 3 && N <= hi;
}
static int tet_l_rt(int N, int b1, int b2, int b3) {          // tet_l<N> for a run-time N
    int l = 0;
    for (int i = 0; i < b1; i++) { const int m = N - i; l += m * (m + 1) / 2; }
    for (int j = 0; j < b2; j++) l += N - b1 - j;
    return l + b3;
}

// C[a3, i, j, k] depends on (i + j, k) only (kernels_ct.cuh: c3_sym)
static bool c_tensor_symmetric(const sse_arrays& a, int N) {
    if (!a.C) return false;
    for (int b1 = 0; b1 < N; b1++) for (int b2 = 0; b1 + b2 < N; b2++) for (int b3 = 0; b1 + b2 + b3 < N; b3++) for (int a3 = 0; a3 < N; a3++)
        if (a.C[a3 + N * (b1 + N * (b2 + N * b3))] != a.C[a3 + N * (0 + N * ((b1 + b2) + N * b3))]) return false;
    return true;
}

// the compile-time path needs: d = 3, Euler + EC two-point flux, flux differencing, warped V with
// M1 = M2 = M3 = p + 1 in the canonical orderings, weight-adjusted mass solver, and a tensor plan
bool ct_eligible(const sse_config& cfg, const sse_arrays& a, const TensorPlan& tp, int* Nout) {
    if (!tp.ok || cfg.d != 3 || cfg.N_c != 5 || cfg.pde != SSE_PDE_EULER) return false;
    if (cfg.form != SSE_FORM_FLUX_DIFFERENCING || cfg.two_point_flux != SSE_TWO_POINT_ENTROPY_CONSERVATIVE) return false;
    if (cfg.v_kind != SSE_V_WARPED || cfg.mass_solver != SSE_MASS_WEIGHT_ADJUSTED || !a.Cfd) return false;
    const int N = cfg.p + 1;
    if (!ct_size_ok(N)) return false;
    if (cfg.M1d[0] != N || cfg.M1d[1] != N || cfg.M1d[2] != N) return false;
    if (cfg.N_q != N * N * N || cfg.N_p != N * (N + 1) * (N + 2) / 6 || cfg.N_f != 4 * N * N || cfg.N_fac != 4) return false;
    for (int t = 0; t < N * N * N; t++) {                  // canonical orderings (tensor_simplex.jl:111-131)
        const int a1 = t % N, a2 = (t / N) % N, a3 = t / (N * N);
        if (a.sigma_o[t] - 1 != (a1 * N + a2) * N + a3) return false;
        const long long want = (a1 + a2 + a3 <= N - 1) ? tet_l_rt(N, a1, a2, a3) + 1 : 0;
        if (a.sigma_i[t] != want) return false;
    }
    if (!c_tensor_symmetric(a, N)) return false;
    if (!a.nJq) {                                          // the pair kernel forms 2 halfnJq from these normals without multiplying
        static const double want[12] = {0, -1, 0, 1, 1, 1, -1, 0, 0, 0, 0, -1};
        if (!a.nref) return false;
        for (int i = 0; i < 12; i++) if (a.nref[i] != want[i]) return false;
    }
    *Nout = N;
    return true;
}

bool ct_schedule_matches(const TensorPlan& tp, int N) {
    const int NN = N * N, Nq = N * NN, NSH = N / 2;
    if (tp.dev.n_vrounds != 3 * NSH || tp.dev.n_frounds != 3 + N) return false;
    for (int i = 0; i < Nq; i++) {
        const int c[3] = {i / NN, (i / N) % N, i % N};
        const int stride[3] = {NN, N, 1};
        for (int rd = 0; rd < 3 * NSH; rd++) {
            const int l = rd / NSH, sh = rd % NSH + 1;
            const bool half = 2 * sh == N;
            const int cj = (c[l] + sh) % N, cs = (c[l] - sh + N) % N;
            const int want_p = (half && c[l] >= sh) ? -1 : i + (cj - c[l]) * stride[l];
            const int want_s = (half && cs >= sh) ? -1 : i + (cs - c[l]) * stride[l];
            if (tp.v_partner[(size_t)rd * Nq + i] != want_p || tp.v_source[(size_t)rd * Nq + i] != want_s) return false;
            if (tp.v_mlo[rd] < l) return false;              // the kernel skips m < l
        }
        for (int fr = 0; fr < 3 + N; fr++) {
            int j;
            if (fr == 0) j = c[0] * N + c[2];
            else if (fr == 1) j = NN + c[1] * N + c[2];
            else if (fr == 2) j = 2 * NN + c[1] * N + c[2];
            else j = 3 * NN + c[0] * N + ((fr - 3 - c[2]) % N + N) % N;
            if (tp.f_partner[(size_t)fr * Nq + i] != j) return false;
            if (tp.f_face[fr] != (fr < 3 ? fr : 3)) return false;
        }
    }
    return true;
}

// 1-D factors of R (FacetR in kernels_ct.cuh) from the dense Matrix(R) (N_f x N_q, column-major); false unless every
// entry of R equals its structured value to 1e-14 (zeros exactly)
bool ct_facet_factors(const sse_config& cfg, const sse_arrays& a, int N, std::vector<double>& out) {
    const int NN = N * N, Nq = N * NN, Nf = 4 * NN;
    if (!a.R || cfg.N_q != Nq || cfg.N_f != Nf) return false;
    auto R = [&](int j, int a1, int a2, int a3) { return a.R[j + (size_t)Nf * ((a1 * N + a2) * N + a3)]; };
    out.assign(4 * N + NN, 0.0);
    double *r0 = out.data(), *r1 = r0 + N, *r2 = r1 + N, *r3 = r2 + N, *I3 = r3 + N;
    for (int c = 0; c < N; c++) { r0[c] = R(0, 0, c, 0); r1[c] = R(NN, c, 0, 0); r2[c] = R(2 * NN, c, 0, 0); }
    // face 3: R[(a1, b), (a1, a2, a3)] = I3[b, a2] r3[a3], normalised to I3[0, 0] = 1; split at the largest entry of the
    // (b, a2) = (0, 0) fibre
    int cm = 0;
    for (int c = 0; c < N; c++) { r3[c] = R(3 * NN, 0, 0, c); if (std::fabs(r3[c]) > std::fabs(r3[cm])) cm = c; }
    if (r3[cm] == 0.0) return false;
    for (int b = 0; b < N; b++) for (int a2 = 0; a2 < N; a2++) I3[b + N * a2] = R(3 * NN + b, 0, a2, cm) / r3[cm];
    double scale = 0.0;
    for (size_t i = 0; i < (size_t)Nf * Nq; i++) scale = std::max(scale, std::fabs(a.R[i]));
    for (int a1 = 0; a1 < N; a1++) for (int a2 = 0; a2 < N; a2++) for (int a3 = 0; a3 < N; a3++)
        for (int j = 0; j < Nf; j++) {
            const int f = j / NN, x = (j % NN) / N, y = j % N;
            double want = 0.0;
            if (f == 0) want = (x == a1 && y == a3) ? r0[a2] : 0.0;
            else if (f == 1) want = (x == a2 && y == a3) ? r1[a1] : 0.0;
            else if (f == 2) want = (x == a2 && y == a3) ? r2[a1] : 0.0;
            else want = (x == a1) ? I3[y + N * a2] * r3[a3] : 0.0;
            const double got = R(j, a1, a2, a3);
            if (want == 0.0 ? got != 0.0 : std::fabs(got - want) > 1e-14 * scale) return false;
        }
    return true;
}


// ---------------------------------------------------------------------------------------------------------
// kind 2: 2-D Euler flux differencing on collapsed triangles (kernels_tri.cuh)
#ifndef SSE_TRI_WARPS
#define SSE_TRI_WARPS 4
#endif
// resident CTAs per SM requested from ptxas: 3 = 168 registers, 12 warps per SM.  Both kernels want more registers than
// the 128 of 16 warps per SM (pass B spills 300 bytes there, and its spill traffic competes for the shared-memory pipe):
// measured at 131 072 triangles, pass A 0.276 (4) / 0.271 ms (3) with the V row in registers, pass B 0.515 (4) / 0.508 ms (3)
#ifndef SSE_TRI_MINB
#define SSE_TRI_MINB 3
#endif
// pass B with asynchronous copies into the warp's tiles (k_tri_fluxdiff_async) or with register prefetch (k_tri_fluxdiff)
#ifndef SSE_TRI_ASYNC
#define SSE_TRI_ASYNC 0
#endif
#ifndef SSE_TRI_MINB_A
#define SSE_TRI_MINB_A SSE_TRI_MINB
#endif
constexpr int TRI_WARPS = SSE_TRI_WARPS, TRI_MINB = SSE_TRI_MINB, TRI_MINB_A = SSE_TRI_MINB_A;
#define SSE_TRI_DISPATCH(N_, CALL) \
    switch (N_) { case 3: CALL(3); break; case 4: CALL(4); break; case 5: CALL(5); break; default: break; }

bool tri_eligible(const sse_config& cfg, const sse_arrays& a, const TensorPlan& tp, CtPlan& p) {
    if (const char* e = getenv("SSE_TRI_CT")) if (atoi(e) == 0) return false;
    if (!tp.ok || cfg.d != 2 || cfg.N_c != 4 || cfg.pde != SSE_PDE_EULER) return false;
    if (cfg.form != SSE_FORM_FLUX_DIFFERENCING || cfg.two_point_flux != SSE_TWO_POINT_ENTROPY_CONSERVATIVE) return false;
    if (cfg.v_kind != SSE_V_WARPED || cfg.mass_solver != SSE_MASS_WEIGHT_ADJUSTED || !a.Cfd || !a.A || !a.B || !a.R) return false;
    if (!a.sigma_i || !a.sigma_o || !a.W || !a.Bf) return false;
    const int N = cfg.p + 1;
    if (N < 3 || N > 5) return false;
    if (cfg.M1d[0] != N || cfg.M1d[1] != N) return false;
    const int Nq = N * N, Np = N * (N + 1) / 2, Nf = 3 * N, NSH = N / 2;
    if (cfg.N_q != Nq || cfg.N_p != Np || cfg.N_f != Nf || cfg.N_fac != 3) return false;
    if (!a.nJq && !a.nref) return false;
    // node ordering: tensor index (a1, a2) -> node a1 N + a2 (tensor_simplex.jl:111-112); every mode used exactly once
    std::vector<char> seen(Np, 0);
    for (int t = 0; t < Nq; t++) {
        const int a1 = t % N, a2 = t / N;
        if (a.sigma_o[t] - 1 != a1 * N + a2) return false;
        const long long l = a.sigma_i[t];
        if (l < 0 || l > Np) return false;
        if ((l > 0) != (a1 + a2 <= N - 1)) return false;
        if (l > 0 && l - 1 != a1 * N - a1 * (a1 - 1) / 2 + a2) return false;      // canonical modal index (tri_l in kernels_tri.cuh)
        if (l > 0) { if (seen[l - 1]) return false; seen[l - 1] = 1; }
    }
    // dense V[node, l] = A[a1, b1] B[a2, b1, b2] (warped_product_2d; the expression of sse_create's small-element V)
    p.triV.assign((size_t)Nq * Np, 0.0);
    for (int b1 = 0; b1 < N; b1++)
        for (int b2 = 0; b1 + b2 < N; b2++) {
            const int l = (int)a.sigma_i[b1 + N * b2] - 1;
            for (int a1 = 0; a1 < N; a1++)
                for (int a2 = 0; a2 < N; a2++) p.triV[(a1 * N + a2) + (size_t)Nq * l] = a.A[a1 + N * b1] * a.B[a2 + N * (b1 + N * b2)];
        }
    // R: facet node (0, a1) <- the a2-line of a1; facet nodes (1, a2), (2, a2) <- the a1-line of a2; nothing else
    p.triRfac.assign((size_t)Nf * N, 0.0);
    for (int j = 0; j < Nf; j++) {
        const int f = j / N, q = j % N;
        for (int i = 0; i < Nq; i++) {
            const int a1 = i / N, a2 = i % N;
            const bool on = f == 0 ? a1 == q : a2 == q;
            const double v = a.R[j + (size_t)Nf * i];
            if (on) p.triRfac[(size_t)j * N + (f == 0 ? a2 : a1)] = v;
            else if (v != 0.0) return false;
        }
    }
    // pair schedule: the closed form of the kernel against the generic tables
    if (tp.dev.n_vrounds != 2 * NSH || tp.dev.n_frounds != 3) return false;
    for (int fr = 0; fr < 3; fr++) if (tp.f_face[fr] != fr) return false;
    for (int i = 0; i < Nq; i++) {
        const int c[2] = {i / N, i % N}, stride[2] = {N, 1};
        for (int rd = 0; rd < 2 * NSH; rd++) {
            const int l = rd / NSH, sh = rd % NSH + 1;
            const bool half = 2 * sh == N;
            const int cj = (c[l] + sh) % N, cs = (c[l] - sh + N) % N;
            const int want_p = (half && c[l] >= sh) ? -1 : i + (cj - c[l]) * stride[l];
            const int want_s = (half && cs >= sh) ? -1 : i + (cs - c[l]) * stride[l];
            if (tp.v_partner[(size_t)rd * Nq + i] != want_p || tp.v_source[(size_t)rd * Nq + i] != want_s) return false;
            if (l == 1 && tp.v_S[((size_t)rd * 2 + 0) * Nq + i] != 0.0) return false;      // a2-lines carry S_2 only
        }
        const int want_f[3] = {c[0], N + c[1], 2 * N + c[1]};
        for (int fr = 0; fr < 3; fr++) if (tp.f_partner[(size_t)fr * Nq + i] != want_f[fr]) return false;
    }
    // power-of-two scalings of the pair weights (exact): ec_finish_scaled in physics.cuh
    p.trivS = tp.v_S; for (double& x : p.trivS) x *= 0.25;
    p.trifC = tp.f_C; for (double& x : p.trifC) x *= 0.125;
    p.trifR.assign((size_t)3 * Nq, 0.0);
    for (int fr = 0; fr < 3; fr++)
        for (int i = 0; i < Nq; i++) p.trifR[(size_t)fr * Nq + i] = a.R[tp.f_partner[(size_t)fr * Nq + i] + (size_t)Nf * i];
    for (int i = 0; i < 6; i++) p.tri.nref[i] = a.nref ? a.nref[i] : 0.0;
    p.N = N;
    return true;
}

template <int N> static cudaError_t tri_set_attrs_n() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_tri_nodal<N, TRI_WARPS, TRI_MINB_A>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(double) * TriT<N>::template smem_doubles<false>(TRI_WARPS))))) return e;
    if ((e = cudaFuncSetAttribute(k_tri_fluxdiff_async<N, TRI_WARPS, TRI_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(double) * TriB<N>::smem_doubles(TRI_WARPS))))) return e;
    return cudaFuncSetAttribute(k_tri_fluxdiff<N, TRI_WARPS, TRI_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(sizeof(double) * TriT<N>::template smem_doubles<true>(TRI_WARPS)));
}
cudaError_t tri_set_attrs(int N) {
    cudaError_t e = cudaErrorInvalidValue;
#define CALL_(N_) e = tri_set_attrs_n<N_>()
    SSE_TRI_DISPATCH(N, CALL_);
#undef CALL_
    return e;
}
// persistent warps: one resident wave of CTAs, each warp strides over the elements of the range
static unsigned tri_grid(const CtPlan& p, long long count, int minb = TRI_MINB) {
    const long long want = (count + TRI_WARPS - 1) / TRI_WARPS, wave = (long long)p.sms * minb;
    return (unsigned)std::max<long long>(1, std::min(want, wave));
}
template <int N>
static void tri_nodal_n(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, const double* u, double* u_q,
                        double* u_f, cudaStream_t s) {
    k_tri_nodal<N, TRI_WARPS, TRI_MINB_A><<<tri_grid(p, count, TRI_MINB_A), TRI_WARPS * 32, sizeof(double) * TriT<N>::template smem_doubles<false>(TRI_WARPS), s>>>(
        p.tri, g, L, first, count, u, u_q, u_f);
}
template <int N>
static void tri_fluxdiff_n(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, const double* u_q,
                           const double* u_f, double* dudt, cudaStream_t s, RkStage rk) {
#if SSE_TRI_ASYNC
    k_tri_fluxdiff_async<N, TRI_WARPS, TRI_MINB><<<tri_grid(p, count), TRI_WARPS * 32, sizeof(double) * TriB<N>::smem_doubles(TRI_WARPS), s>>>(
        p.tri, g, L, first, count, u_q, u_f, dudt, rk);
#else
    k_tri_fluxdiff<N, TRI_WARPS, TRI_MINB><<<tri_grid(p, count), TRI_WARPS * 32, sizeof(double) * TriT<N>::template smem_doubles<true>(TRI_WARPS), s>>>(
        p.tri, g, L, first, count, u_q, u_f, dudt, rk);
#endif
}


// kind 3: 2-D linear advection, StandardForm + ReferenceOperators on collapsed triangles (k_tri_adv_facets / k_tri_adv)
bool tri_adv_eligible(const sse_config& cfg, const sse_arrays& a, CtPlan& p) {
    if (const char* e = getenv("SSE_TRI_CT")) if (atoi(e) == 0) return false;
    if (cfg.d != 2 || cfg.N_c != 1 || cfg.pde != SSE_PDE_ADVECTION || cfg.form != SSE_FORM_STANDARD_REFERENCE) return false;
    if (cfg.v_kind != SSE_V_WARPED || cfg.mass_solver != SSE_MASS_WEIGHT_ADJUSTED || !a.A || !a.B || !a.R || !a.D[0] || !a.D[1]) return false;
    if (!a.sigma_i || !a.sigma_o || !a.W || !a.Bf) return false;
    const int N = cfg.p + 1;
    if (N < 3 || N > 5) return false;
    if (cfg.M1d[0] != N || cfg.M1d[1] != N) return false;
    const int Nq = N * N, Np = N * (N + 1) / 2, Nf = 3 * N;
    if (cfg.N_q != Nq || cfg.N_p != Np || cfg.N_f != Nf || cfg.N_fac != 3) return false;
    for (int t = 0; t < Nq; t++) {
        const int a1 = t % N, a2 = t / N;
        if (a.sigma_o[t] - 1 != a1 * N + a2) return false;
        const long long want = (a1 + a2 <= N - 1) ? a1 * N - a1 * (a1 - 1) / 2 + a2 + 1 : 0;
        if (a.sigma_i[t] != want) return false;
    }
    p.triV.assign((size_t)Nq * Np, 0.0);
    for (int b1 = 0; b1 < N; b1++)
        for (int b2 = 0; b1 + b2 < N; b2++) {
            const int l = (int)a.sigma_i[b1 + N * b2] - 1;
            for (int a1 = 0; a1 < N; a1++)
                for (int a2 = 0; a2 < N; a2++) p.triV[(a1 * N + a2) + (size_t)Nq * l] = a.A[a1 + N * b1] * a.B[a2 + N * (b1 + N * b2)];
        }
    // R: one tensor line per facet node (as for kind 2); lift weights R[partner_f, i]
    p.trifR.assign((size_t)3 * Nq, 0.0);
    for (int j = 0; j < Nf; j++) {
        const int f = j / N, q = j % N;
        for (int i = 0; i < Nq; i++) {
            const int a1 = i / N, a2 = i % N;
            const bool on = f == 0 ? a1 == q : a2 == q;
            const double v = a.R[j + (size_t)Nf * i];
            if (on) p.trifR[(size_t)f * Nq + i] = v;
            else if (v != 0.0) return false;
        }
    }
    p.triRV.assign((size_t)Nf * Np, 0.0);
    for (int l = 0; l < Np; l++)
        for (int j = 0; j < Nf; j++) {
            double s = 0.0;
            for (int i = 0; i < Nq; i++) s = std::fma(a.R[j + (size_t)Nf * i], p.triV[i + (size_t)Nq * l], s);
            p.triRV[j + (size_t)Nf * l] = s;
        }
    // D[m] must be the Kronecker product of a 1-D matrix along direction m (a1: stride N, a2: stride 1) with the identity
    const int stride[2] = {N, 1};
    p.triD1.assign((size_t)2 * N * N, 0.0);
    for (int m = 0; m < 2; m++) {
        for (int t = 0; t < N; t++) for (int s = 0; s < N; s++) p.triD1[(size_t)m * N * N + t + N * s] = a.D[m][t * stride[m] + (size_t)Nq * (s * stride[m])];
        for (int i = 0; i < Nq; i++)
            for (int j = 0; j < Nq; j++) {
                const int ci[2] = {i / N, i % N}, cj[2] = {j / N, j % N};
                const bool line = ci[1 - m] == cj[1 - m];
                const double want = line ? p.triD1[(size_t)m * N * N + ci[m] + N * cj[m]] : 0.0;
                if (a.D[m][i + (size_t)Nq * j] != want) return false;
            }
    }
    p.N = N;
    return true;
}
template <int N> static void tri_adv_facets_n(const CtPlan& p, const Geo& g, long long first, long long count, const double* u, double* u_f, double* um,
                                              cudaStream_t s) {
    const long long want = (count + TRI_WARPS - 1) / TRI_WARPS;
    k_tri_adv_facets<N, TRI_WARPS><<<(unsigned)std::max<long long>(1, std::min<long long>(want, (long long)p.sms * 8)), TRI_WARPS * 32, 0, s>>>(p.tri, g, first, count, u, u_f, um);
}
template <int N> static void tri_adv_n(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, const double* u, const double* u_f,
                                       double* dudt, cudaStream_t s, RkStage rk) {
    k_tri_adv<N, TRI_WARPS, TRI_MINB><<<tri_grid(p, count), TRI_WARPS * 32, sizeof(double) * TRI_WARPS * TriA<N>::warp, s>>>(p.tri, g, L, first, count, u, u_f, dudt, rk);
}

template <int N> static FacetR<N> make_facet(const CtPlan& p) {
    FacetR<N> f;
    const double* s = p.facetR.data();
    for (int i = 0; i < N; i++) { f.r0[i] = s[i]; f.r1[i] = s[N + i]; f.r2[i] = s[2 * N + i]; f.r3[i] = s[3 * N + i]; }
    for (int i = 0; i < N * N; i++) f.I3[i] = s[4 * N + i];
    return f;
}

// resident CTAs per SM requested for the pair kernel (register cap = 65536 / (128 MINB))
#ifndef SSE_FD_MINB_CT
#define SSE_FD_MINB_CT 4
#endif
// Launch shapes per degree (resident CTAs per SM requested from ptxas = its register cap).  N <= 5: the measured values above.
// N = 6, 7 (216 / 343 nodes, slabs of 36 / 49 doubles per thread in the projection kernels): two CTAs per SM -- the pair kernel
// gets 128 / 80 registers (four CTAs left it 72 with spills at N = 6), the projection kernels 200 without spills.  Measured on
// Euler EC at 6 000 elements (profiles/r2_s5_highp_ab.log): p = 5 0.597 (4 / 3 CTAs) -> 0.519 ms (2 / 2), 0.541 with 3 / 3;
// p = 6 1.082 (1 / 1) -> 1.027 ms (2 / 2).  N = 8 (512 nodes, 64-double slabs): one CTA per SM is all the shared memory allows
// (131 kB pair tiles, 127 kB projection tiles): 128 registers for the 512-thread pair kernel, 255 for the projection kernels.
#ifndef SSE_FD_MINB_N6
#define SSE_FD_MINB_N6 2
#endif
#ifndef SSE_FD_MINB_N7
#define SSE_FD_MINB_N7 2
#endif
template <int N> constexpr int fd_minb() { return N <= 5 ? SSE_FD_MINB_CT : (N == 6 ? SSE_FD_MINB_N6 : (N == 7 ? SSE_FD_MINB_N7 : 1)); }
template <int N> constexpr int nodal1_minb() { return N <= 6 ? 4 : 2; }       // scalar laws (128 threads)
template <int N> constexpr int proj1_minb() { return N <= 6 ? 3 : 2; }
template <int N> constexpr int sadv_minb() { return N <= 6 ? 8 : 2; }         // k_standard_adv_ct (N_q threads)
template <int N> static SFCoef<N> make_coef(const CtPlan& p) {
    SFCoef<N> c;
    for (int i = 0; i < N * N; i++) c.A[i] = p.A[i];
    for (int i = 0; i < N * N * N; i++) c.B[i] = p.B[i];
    return c;
}

// resident CTAs per SM requested for the Euler projection kernels (3: 128 registers, 15 warps per SM)
#ifndef SSE_NODAL_MINB_CT
#define SSE_NODAL_MINB_CT 3
#endif
#ifndef SSE_PROJ_MINB_CT
#define SSE_PROJ_MINB_CT 3
#endif
#ifndef SSE_PROJ_MINB_N6
#define SSE_PROJ_MINB_N6 2
#endif
#ifndef SSE_PROJ_MINB_N7
#define SSE_PROJ_MINB_N7 2
#endif
template <int N> constexpr int nodal_minb() { return N <= 5 ? SSE_NODAL_MINB_CT : (N == 6 ? SSE_PROJ_MINB_N6 : (N == 7 ? SSE_PROJ_MINB_N7 : 1)); }
template <int N> constexpr int proj_minb() { return N <= 5 ? SSE_PROJ_MINB_CT : (N == 6 ? SSE_PROJ_MINB_N6 : (N == 7 ? SSE_PROJ_MINB_N7 : 1)); }
// fused advection path: warps per CTA and resident CTAs per SM requested from ptxas (register cap 65536 / (32 WARPS MINB))
#ifndef SSE_ADV_WARPS
#define SSE_ADV_WARPS 2
#endif
#ifndef SSE_ADV_MINB
#define SSE_ADV_MINB 4
#endif
constexpr int ADV_WARPS = SSE_ADV_WARPS, ADV_MINB = SSE_ADV_MINB;
template <int N> static constexpr int adv_smem() { return (int)(sizeof(double) * ADV_WARPS * AdvSmem<N>::GSLOTS * AdvSmem<N>::group); }

template <int N> static bool adv_build_n(CtPlan& p, const Geo& g, const Law& L, const double* W, const double* Bf, long long Ne,
                                         cudaStream_t s, std::vector<void*>& owned) {
    constexpr int GPW = 32 / N, NN = N * N, NI = 4 * N, Np = Tet<N>::Np;
    const long long tasks = (Ne + GPW - 1) / GPW;
    double *C = nullptr, *iJW = nullptr, *F = nullptr, *um = nullptr;
    int* map = nullptr;
    auto al = [&](void** q, size_t bytes) { if (cudaMalloc(q, bytes) != cudaSuccess) return false; owned.push_back(*q); return cudaMemsetAsync(*q, 0, bytes, s) == cudaSuccess; };
    if (!al((void**)&C, sizeof(double) * tasks * 3 * NN * 32) || !al((void**)&iJW, sizeof(double) * tasks * NN * 32) ||
        !al((void**)&F, sizeof(double) * tasks * 2 * NI * 32) || !al((void**)&map, sizeof(int) * tasks * NI * 32) ||
        !al((void**)&um, sizeof(double) * Ne * Np))
        return false;
    k_adv_build<N><<<(unsigned)Ne, 128, 0, s>>>(Ne, W, g.Lambda_q, g.J_q, g.J_f, g.nJf, Bf, g.mapP, L, C, iJW, F, map);
    if (cudaGetLastError() != cudaSuccess) return false;
    p.adv.C = C; p.adv.iJW = iJW; p.adv.F = F; p.adv.map = map; p.adv.um = um;
    p.adv_ok = 1;
    return true;
}
bool ct_adv_build(CtPlan& p, const Geo& g, const Law& L, const double* W, const double* Bf, long long Ne, long long NFT,
                  cudaStream_t s, std::vector<void*>& owned) {
    if (p.kind != 1 || NFT >= 2147483647LL) return false;             // neighbour indices are stored as int32
    if (const char* e = getenv("SSE_ADV_FUSED")) if (atoi(e) == 0) return false;
    bool ok = false;
#define CALL_(N_) ok = adv_build_n<N_>(p, g, L, W, Bf, Ne, s, owned)
    SSE_CT_DISPATCH(p.N, CALL_);
#undef CALL_
    return ok;
}

template <int N> static cudaError_t set_attrs_n() {
    cudaError_t e;
    const int proj5 = (int)(sizeof(double) * ProjSmem<N, 5>::total), proj1 = (int)(sizeof(double) * ProjSmem<N, 1>::total);
    const int nod5 = (int)(sizeof(double) * ProjSmem<N, 5, true>::total), nod1 = (int)(sizeof(double) * ProjSmem<N, 1, true>::total);
    constexpr size_t cap = 232448;                           // 227 KB of dynamic shared memory per CTA on sm_100
    static_assert(sizeof(double) * ProjSmem<N, 5, true>::total <= cap && sizeof(double) * ProjSmem<N, 5>::total <= cap, "projection tiles fit one SM");
    static_assert(sizeof(double) * FdSmem<N>::total <= cap && (size_t)adv_smem<N>() <= cap, "pair-kernel tiles fit one SM");
    if ((e = cudaFuncSetAttribute(k_nodal_ct<N, 5, nodal_minb<N>(), true>, cudaFuncAttributeMaxDynamicSharedMemorySize, nod5))) return e;
    if ((e = cudaFuncSetAttribute(k_nodal_ct<N, 5, nodal_minb<N>(), true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, nod5))) return e;
    if ((e = cudaFuncSetAttribute(k_project_ct<N, 5, proj_minb<N>()>, cudaFuncAttributeMaxDynamicSharedMemorySize, proj5))) return e;
    if ((e = cudaFuncSetAttribute(k_nodal_ct<N, 1, nodal1_minb<N>(), false>, cudaFuncAttributeMaxDynamicSharedMemorySize, nod1))) return e;
    if ((e = cudaFuncSetAttribute(k_project_ct<N, 1, proj1_minb<N>()>, cudaFuncAttributeMaxDynamicSharedMemorySize, proj1))) return e;
    if ((e = cudaFuncSetAttribute(k_fluxdiff_ct<N, fd_minb<N>(), false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * FdSmem<N>::total)))) return e;
    if ((e = cudaFuncSetAttribute(k_adv_facets_ct<N, ADV_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, adv_smem<N>()))) return e;
    if ((e = cudaFuncSetAttribute(k_adv_fused_ct<N, ADV_WARPS, ADV_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, adv_smem<N>()))) return e;
    if constexpr (N == 5) {
        if ((e = cudaFuncSetAttribute(k_fluxdiff_ct<N, fd_minb<N>(), true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * FdSmem<N, true>::total)))) return e;
    }
    return cudaSuccess;
}
cudaError_t ct_set_attrs(int N) {
    cudaError_t e = cudaErrorInvalidValue;
#define CALL_(N_) e = set_attrs_n<N_>()
    SSE_CT_DISPATCH(N, CALL_);
#undef CALL_
    return e;
}

template <int N>
static void nodal_n(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, const double* u, double* u_q, double* u_f,
                    cudaStream_t s) {
    if (p.kind == 0) {
        const unsigned grid = (unsigned)((count + ProjSmem<N, 5>::EPB - 1) / ProjSmem<N, 5>::EPB);
        k_nodal_ct<N, 5, nodal_minb<N>(), true><<<grid, 160, sizeof(double) * ProjSmem<N, 5, true>::total, s>>>(make_coef<N>(p), make_facet<N>(p), p.dev, g, L, first, count, u, u_q, u_f);
    } else if (p.adv_ok) {
        constexpr int GPW = 32 / N;
        const long long tasks = (first + count - 1) / GPW - first / GPW + 1;
        k_adv_facets_ct<N, ADV_WARPS><<<(unsigned)((tasks + ADV_WARPS - 1) / ADV_WARPS), ADV_WARPS * 32, adv_smem<N>(), s>>>(
            make_coef<N>(p), make_facet<N>(p), p.dev, p.adv, first, count, u, u_f);
    } else {
        const unsigned grid = (unsigned)((count + ProjSmem<N, 1>::EPB - 1) / ProjSmem<N, 1>::EPB);
        k_nodal_ct<N, 1, nodal1_minb<N>(), false><<<grid, 128, sizeof(double) * ProjSmem<N, 1, true>::total, s>>>(make_coef<N>(p), make_facet<N>(p), p.dev, g, L, first, count, u, u_q, u_f);
    }
}
void ct_nodal(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, const double* u, double* u_q, double* u_f,
              cudaStream_t s) {
    if (p.kind == 2) {
#define CALL_(N_) tri_nodal_n<N_>(p, g, L, first, count, u, u_q, u_f, s)
        SSE_TRI_DISPATCH(p.N, CALL_);
#undef CALL_
        return;
    }
    if (p.kind == 3) {
#define CALL_(N_) tri_adv_facets_n<N_>(p, g, first, count, u, u_f, u_q, s)
        SSE_TRI_DISPATCH(p.N, CALL_);
#undef CALL_
        return;
    }
#define CALL_(N_) nodal_n<N_>(p, g, L, first, count, u, u_q, u_f, s)
    SSE_CT_DISPATCH(p.N, CALL_);
#undef CALL_
}

template <int N>
static void pair_n(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, double* u_q, const double* u_f, cudaStream_t s) {
    constexpr int NT = Tet<N>::NT;
    if constexpr (N == 5) {
        if (p.dual) { k_fluxdiff_ct<N, fd_minb<N>(), true><<<(unsigned)count, NT, sizeof(double) * FdSmem<N, true>::total, s>>>(p.dev, g, L, first, u_q, u_f); return; }
    }
    k_fluxdiff_ct<N, fd_minb<N>(), false><<<(unsigned)count, NT, sizeof(double) * FdSmem<N>::total, s>>>(p.dev, g, L, first, u_q, u_f);
}
void ct_pair(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, double* u_q, const double* u_f, cudaStream_t s) {
#define CALL_(N_) pair_n<N_>(p, g, L, first, count, u_q, u_f, s)
    SSE_CT_DISPATCH(p.N, CALL_);
#undef CALL_
}
template <int N>
static void project_n(const CtPlan& p, const Geo& g, long long first, long long count, const double* r_q, double* dudt, cudaStream_t s, RkStage rk) {
    const unsigned grid = (unsigned)((count + ProjSmem<N, 5>::EPB - 1) / ProjSmem<N, 5>::EPB);
    k_project_ct<N, 5, proj_minb<N>()><<<grid, 160, sizeof(double) * ProjSmem<N, 5>::total, s>>>(make_coef<N>(p), p.dev, g, first, count, r_q, dudt, rk);
}
void ct_project(const CtPlan& p, const Geo& g, long long first, long long count, const double* r_q, double* dudt, cudaStream_t s, RkStage rk) {
#define CALL_(N_) project_n<N_>(p, g, first, count, r_q, dudt, s, rk)
    SSE_CT_DISPATCH(p.N, CALL_);
#undef CALL_
}
template <int N>
static void project_nodal_n(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, double* u_q, double* u_f,
                            double* dudt, cudaStream_t s, RkStage rk) {
    const unsigned grid = (unsigned)((count + ProjSmem<N, 5>::EPB - 1) / ProjSmem<N, 5>::EPB);
    k_nodal_ct<N, 5, nodal_minb<N>(), true, true><<<grid, 160, sizeof(double) * ProjSmem<N, 5, true>::total, s>>>(
        make_coef<N>(p), make_facet<N>(p), p.dev, g, L, first, count, nullptr, u_q, u_f, dudt, rk);
}
void ct_project_nodal(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, double* u_q, double* u_f,
                      double* dudt, cudaStream_t s, RkStage rk) {
#define CALL_(N_) project_nodal_n<N_>(p, g, L, first, count, u_q, u_f, dudt, s, rk)
    SSE_CT_DISPATCH(p.N, CALL_);
#undef CALL_
}

template <int N>
static void fluxdiff_n(const CtPlan& p, const TensorPlan& tp, const Ops& o, const Geo& g, const Law& L, long long first, long long count,
                       double* u_q, const double* u_f, double* dudt, cudaStream_t s, RkStage rk, cudaEvent_t mid) {
    constexpr int NT = Tet<N>::NT;
    (void)tp; (void)o;
    bool done = false;
    if constexpr (N == 5) {
        if (p.dual) { k_fluxdiff_ct<N, fd_minb<N>(), true><<<(unsigned)count, NT, sizeof(double) * FdSmem<N, true>::total, s>>>(p.dev, g, L, first, u_q, u_f); done = true; }
    }
    if (!done) k_fluxdiff_ct<N, fd_minb<N>(), false><<<(unsigned)count, NT, sizeof(double) * FdSmem<N>::total, s>>>(p.dev, g, L, first, u_q, u_f);
    if (mid) cudaEventRecord(mid, s);
    const unsigned grid = (unsigned)((count + ProjSmem<N, 5>::EPB - 1) / ProjSmem<N, 5>::EPB);
    k_project_ct<N, 5, proj_minb<N>()><<<grid, 160, sizeof(double) * ProjSmem<N, 5>::total, s>>>(make_coef<N>(p), p.dev, g, first, count, u_q, dudt, rk);
}
void ct_fluxdiff(const CtPlan& p, const TensorPlan& tp, const Ops& o, const Geo& g, const Law& L, long long first, long long count,
                 double* u_q, const double* u_f, double* dudt, cudaStream_t s, RkStage rk, cudaEvent_t mid) {
    if (p.kind == 2) {                                // all of pass B in one launch
        if (mid) cudaEventRecord(mid, s);
#define CALL_(N_) tri_fluxdiff_n<N_>(p, g, L, first, count, u_q, u_f, dudt, s, rk)
        SSE_TRI_DISPATCH(p.N, CALL_);
#undef CALL_
        return;
    }
#define CALL_(N_) fluxdiff_n<N_>(p, tp, o, g, L, first, count, u_q, u_f, dudt, s, rk, mid)
    SSE_CT_DISPATCH(p.N, CALL_);
#undef CALL_
}


bool ct_eligible_standard(const sse_config& cfg, const sse_arrays& a, int* Nout, std::vector<double>& D1, std::vector<double>& fR) {
    if (cfg.d != 3 || cfg.N_c != 1 || cfg.pde != SSE_PDE_ADVECTION || cfg.form != SSE_FORM_STANDARD_REFERENCE) return false;
    if (cfg.v_kind != SSE_V_WARPED || cfg.mass_solver != SSE_MASS_WEIGHT_ADJUSTED) return false;
    const int N = cfg.p + 1;
    if (!ct_size_ok(N)) return false;
    if (cfg.M1d[0] != N || cfg.M1d[1] != N || cfg.M1d[2] != N) return false;
    const int NN = N * N, Nq = N * NN, Nf = 4 * NN;
    if (cfg.N_q != Nq || cfg.N_p != N * (N + 1) * (N + 2) / 6 || cfg.N_f != Nf || cfg.N_fac != 4) return false;
    for (int t = 0; t < Nq; t++) {
        const int a1 = t % N, a2 = (t / N) % N, a3 = t / NN;
        if (a.sigma_o[t] - 1 != (a1 * N + a2) * N + a3) return false;
        const long long want = (a1 + a2 + a3 <= N - 1) ? tet_l_rt(N, a1, a2, a3) + 1 : 0;
        if (a.sigma_i[t] != want) return false;
    }
    // D[m] must be the Kronecker product I (x) D_1D (x) I along direction m
    const int stride[3] = {NN, N, 1};
    D1.assign(3 * NN, 0.0);
    for (int m = 0; m < 3; m++) {
        if (!a.D[m]) return false;
        for (int t = 0; t < N; t++) for (int s = 0; s < N; s++) D1[m * NN + t + N * s] = a.D[m][t * stride[m] + (size_t)Nq * (s * stride[m])];
        for (int i = 0; i < Nq; i++)
            for (int j = 0; j < Nq; j++) {
                const int ci[3] = {i / NN, (i / N) % N, i % N}, cj[3] = {j / NN, (j / N) % N, j % N};
                bool line = true;
                for (int q = 0; q < 3; q++) if (q != m && ci[q] != cj[q]) line = false;
                const double want = line ? D1[m * NN + ci[m] + N * cj[m]] : 0.0;
                if (a.D[m][i + (size_t)Nq * j] != want) return false;
            }
    }
    // R must have exactly the closed-form facet partners of k_fluxdiff_ct
    fR.assign((size_t)(3 + N) * Nq, 0.0);
    std::vector<char> seen((size_t)Nf * Nq, 0);
    for (int i = 0; i < Nq; i++) {
        const int c[3] = {i / NN, (i / N) % N, i % N};
        for (int fr = 0; fr < 3 + N; fr++) {
            int j;
            if (fr == 0) j = c[0] * N + c[2];
            else if (fr == 1) j = NN + c[1] * N + c[2];
            else if (fr == 2) j = 2 * NN + c[1] * N + c[2];
            else j = 3 * NN + c[0] * N + ((fr - 3 - c[2]) % N + N) % N;
            fR[(size_t)fr * Nq + i] = a.R[j + (size_t)Nf * i];
            seen[j + (size_t)Nf * i] = 1;
        }
    }
    for (size_t x = 0; x < seen.size(); x++) if (!seen[x] && a.R[x] != 0.0) return false;
    if (!c_tensor_symmetric(a, N)) return false;
    if (!a.nJq) {                                          // the pair kernel forms 2 halfnJq from these normals without multiplying
        static const double want[12] = {0, -1, 0, 1, 1, 1, -1, 0, 0, 0, 0, -1};
        if (!a.nref) return false;
        for (int i = 0; i < 12; i++) if (a.nref[i] != want[i]) return false;
    }
    *Nout = N;
    return true;
}

template <int N>
static void standard_n(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, double* u_q, const double* u_f,
                       double* dudt, cudaStream_t s, RkStage rk, cudaEvent_t mid) {
    if (p.adv_ok) {                                   // two-kernel path: everything of pass B in one launch
        constexpr int GPW = 32 / N;
        AdvTabs2<N> tb;
        for (int m = 0; m < 3; m++) for (int i = 0; i < N * N; i++) tb.D1[m][i] = p.D1[m * N * N + i];
        const long long tasks = (first + count - 1) / GPW - first / GPW + 1;
        if (mid) cudaEventRecord(mid, s);
        k_adv_fused_ct<N, ADV_WARPS, ADV_MINB><<<(unsigned)((tasks + ADV_WARPS - 1) / ADV_WARPS), ADV_WARPS * 32, adv_smem<N>(), s>>>(
            make_coef<N>(p), make_facet<N>(p), tb, p.dev, p.adv, g, first, count, u_f, dudt, rk);
        return;
    }
    constexpr int NT = Tet<N>::NT;
    AdvTabs<N> tabs;
    for (int m = 0; m < 3; m++) for (int i = 0; i < N * N; i++) tabs.D1[m][i] = p.D1[m * N * N + i];
    k_standard_adv_ct<N, sadv_minb<N>()><<<(unsigned)count, NT, 0, s>>>(tabs, p.dev, g, L, first, u_q, u_f);
    if (mid) cudaEventRecord(mid, s);
    const unsigned grid = (unsigned)((count + ProjSmem<N, 1>::EPB - 1) / ProjSmem<N, 1>::EPB);
    k_project_ct<N, 1, proj1_minb<N>()><<<grid, 128, sizeof(double) * ProjSmem<N, 1>::total, s>>>(make_coef<N>(p), p.dev, g, first, count, u_q, dudt, rk);
}
void ct_standard(const CtPlan& p, const Geo& g, const Law& L, long long first, long long count, double* u_q, const double* u_f,
                 double* dudt, cudaStream_t s, RkStage rk, cudaEvent_t mid) {
    if (p.kind == 3) {                                // triangles: all of pass B in one launch, from the modal coefficients of pass A
        if (mid) cudaEventRecord(mid, s);
#define CALL_(N_) tri_adv_n<N_>(p, g, L, first, count, u_q, u_f, dudt, s, rk)
        SSE_TRI_DISPATCH(p.N, CALL_);
#undef CALL_
        return;
    }
#define CALL_(N_) standard_n<N_>(p, g, L, first, count, u_q, u_f, dudt, s, rk, mid)
    SSE_CT_DISPATCH(p.N, CALL_);
#undef CALL_
}

}  // namespace sse
