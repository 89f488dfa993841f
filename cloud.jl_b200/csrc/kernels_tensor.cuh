// kernels_tensor.cuh — flux-differencing pass B specialised for collapsed tensor-product operators
// (ModalTensor / NodalTensor on Tri and Tet; the BASELINE headline path).
//
// Structure exploited (verified on the host by tensor_plan_build, otherwise the generic kernels run):
//   * volume nodes form an N1^d tensor grid and every non-zero of S_m couples two nodes of one
//     tensor line (operators.jl:191-199 with D = I (x) D_1D (x) I, tensor_simplex.jl:298-300);
//   * every volume node couples to the same number of facet nodes per face through C = R'B
//     (tensor_simplex.jl:265-268).
//
// Mapping: one CTA per element, one thread per volume node (125 of 128 lanes at p = 4).
//   volume term  : d*floor(N1/2) rounds; in a round every thread evaluates ONE two-point flux with the
//                  node `shift` places further along its line and keeps -phi; +phi is handed to the
//                  partner through a double-buffered shared-memory stage, so each unique pair is
//                  evaluated exactly once (750 instead of the reference's 1500 evaluations per p=4 tet,
//                  flux_differencing_form.jl:37-75) without atomics.
//   facet correct: deg sub-rounds (8 at p = 4); each thread evaluates one (volume node, facet node)
//                  pair per sub-round, keeps -phi and stages it; (facet node, variable) reducer threads
//                  sum the staged vectors into f_f (flux_differencing_form.jl:126-168).
//   The two-point flux is evaluated already contracted with the pair's metric vector
//   g = sum_m S_m[i,j] (Λ_i + Λ_j)[m,:]  resp.  C_ij (halfnJf_j + halfnJq_i), from per-node primitive
//   variables (rho, V, p, rho/p) computed once per node instead of once per pair.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#include "common.cuh"
#include "kernels_generic.cuh"

namespace sse {

struct TensorDev {                 // device tables (element independent)
    int N1, n_vrounds, n_frounds, red_items_max, red_max;
    const int* v_partner;          // [round][Nq]  partner node or -1
    const int* v_source;           // [round][Nq]  node whose phi this node receives, or -1
    const double* v_S;             // [round][D][Nq] skew-extended S_m[i, partner]
    const int* v_mlo;              // [round] first m with a non-zero weight
    const int* f_partner;          // [sub-round][Nq] facet node
    const double* f_C;             // [sub-round][Nq] C_ij
    const int* f_face;             // [sub-round] face of all partners of this sub-round
    const int* red_n;              // [sub-round] number of reducer items
    const int* red_dst;            // [sub-round][items_max] facet node
    const int* red_cnt;            // [sub-round][items_max]
    const int* red_src;            // [sub-round][items_max][red_max] staged volume nodes
};

struct TensorPlan {
    int ok = 0, has_nodal = 0, has_fluxdiff = 0;
    int threads = 0;
    size_t smem_fluxdiff = 0;
    TensorDev dev{};
    // host images
    std::vector<int> v_partner, v_source, v_mlo, f_partner, f_face, red_n, red_dst, red_cnt, red_src;
    std::vector<double> v_S, f_C;
};

template <int D, int NC> struct PrimCount { static constexpr int value = (NC == D + 2) ? D + 3 : 1; };

// shared-memory layout of k_fluxdiff_tensor (doubles)
struct FdLayout {
    int prim, lam, fprim, hnf, ff, stage, total;
    int post_r, post_m, post_tq, post_z, post_w;
};
__host__ __device__ inline FdLayout fd_layout(const Ops& o, int D, int NC, int NPRIM) {
    FdLayout l;
    l.prim = 0;
    l.lam = l.prim + NPRIM * o.Nq;
    l.fprim = l.lam + D * D * o.Nq;
    l.hnf = l.fprim + NPRIM * o.Nf;
    l.ff = l.hnf + D * o.Nf;
    l.stage = l.ff + NC * o.Nf;
    int fd_total = l.stage + 2 * NC * o.Nq;
    // after the pair phases: r lives in stage buffer 0, everything before `ff` is scratch
    l.post_r = l.stage;
    l.post_m = 0;
    l.post_tq = l.post_m + o.Np * NC;
    l.post_z = l.post_tq + o.Nq * NC;
    l.post_w = l.post_z + warp_z_size(o, NC);
    int post_end = l.post_w + warp_w_size(o, NC);
    // scratch may run over ff/hnf (dead by then) but must stay clear of stage buffer 0
    l.total = fd_total;
    if (post_end > l.stage) {      // not enough dead space in front of the stage: append
        l.post_r = l.stage;        // keep r where it is and move the scratch behind the stage buffers
        l.post_m = fd_total;
        l.post_tq = l.post_m + o.Np * NC;
        l.post_z = l.post_tq + o.Nq * NC;
        l.post_w = l.post_z + warp_z_size(o, NC);
        l.total = l.post_w + warp_w_size(o, NC);
    }
    return l;
}

// ---------------------------------------------------------------------------------------------
// contracted two-point fluxes: phi[e] = sum_n g[n] F[e][n](a, b)
// Euler: Ranocha's EC flux (euler_navierstokes.jl:171-195) from primitives (rho, V, p, beta = rho/p)
template <int D>
__device__ __forceinline__ void ec_flux_contract(const double* a, const double* b, const double* g, double igm1, double* phi) {
    const double rho_hat = logmean(a[0], b[0]);
    const double ilm = inv_logmean(a[D + 2], b[D + 2]);
    double dot = 0.0, ga = 0.0, gb = 0.0;
#pragma unroll
    for (int m = 0; m < D; m++) { dot = fma(a[1 + m], b[1 + m], dot); ga = fma(g[m], a[1 + m], ga); gb = fma(g[m], b[1 + m], gb); }
    const double Cc = fma(igm1, ilm, 0.5 * dot);
    const double mf = rho_hat * (0.5 * (ga + gb));
    const double p_avg = 0.5 * (a[D + 1] + b[D + 1]);
    phi[0] = mf;
#pragma unroll
    for (int m = 0; m < D; m++) phi[1 + m] = fma(mf, 0.5 * (a[1 + m] + b[1 + m]), p_avg * g[m]);
    phi[D + 1] = fma(mf, Cc, 0.5 * fma(a[D + 1], gb, b[D + 1] * ga));
}

// Euler: the power-of-two-scaled form of the compile-time kernels (physics.cuh: primitives (rho, V, 2p, rho/p), branch-free
// log-means and reciprocals); PAIR_G_SCALE is folded into the weight g by the caller
template <int D, int NC> struct PairScale { static constexpr double g = (NC == D + 2) ? 0.25 : 1.0; };
template <int D, int NC>
__device__ __forceinline__ void pair_flux(const Law& L, const double* a, const double* b, const double* g, double* phi) {
    if constexpr (NC == D + 2) {
        ec_contract_scaled<D>(L, a, b, g, phi);
    } else {   // linear advection (linear_advection_diffusion.jl:113-119)
        double ag = 0.0;
#pragma unroll
        for (int m = 0; m < D; m++) ag = fma(L.a[m], g[m], ag);
        phi[0] = ag * (0.5 * (a[0] + b[0]));
    }
}

// conservative state -> primitives used by the pair kernels
template <int D, int NC>
__device__ __forceinline__ double to_prim(const Law& L, const double* u, double* q) {
    if constexpr (NC == D + 2) {
        return to_prim_fast<D>(L, u, q);           // (rho, V, 2p, rho/p); returns 1/rho
    } else {
        q[0] = u[0];
        return 1.0;
    }
}

// ---------------------------------------------------------------------------------------------
template <int D, int NC>
__global__ void __launch_bounds__(128, 4)
k_fluxdiff_tensor(TensorDev t, Ops o, Geo g, Law L, long long first, const double* __restrict__ u_q,
                  const double* __restrict__ u_f, double* __restrict__ dudt) {
    constexpr int NP = PrimCount<D, NC>::value;
    extern __shared__ double sm_cta[];
    double* sm = sse_row_smem(sm_cta);          // element packing: common.cuh
    const int tid = threadIdx.x;
    const long long k = sse_element(first);
    const int Nq = o.Nq, Nf = o.Nf, Np = o.Np;
    const FdLayout lay = fd_layout(o, D, NC, NP);
    double* s_prim = sm + lay.prim;
    double* s_lam = sm + lay.lam;
    double* s_fprim = sm + lay.fprim;
    double* s_hnf = sm + lay.hnf;
    double* s_ff = sm + lay.ff;
    double* s_stage = sm + lay.stage;

    // ---- phase 0: loads, primitives, interface numerical flux
    double qi[NP], lam[D][D], r[NC];
#pragma unroll
    for (int e = 0; e < NC; e++) r[e] = 0.0;
    if (tid < Nq) {
        double ui[NC];
#pragma unroll
        for (int e = 0; e < NC; e++) ui[e] = u_q[(size_t)Nq * NC * k + tid + Nq * e];
        to_prim<D, NC>(L, ui, qi);
#pragma unroll
        for (int c = 0; c < NP; c++) s_prim[c * Nq + tid] = qi[c];
#pragma unroll
        for (int n = 0; n < D; n++)
#pragma unroll
            for (int m = 0; m < D; m++) {
                lam[m][n] = g.Lambda_q[(size_t)Nq * D * D * k + tid + Nq * (m + D * n)];
                s_lam[(m + D * n) * Nq + tid] = lam[m][n];
            }
    }
    for (int j = tid; j < Nf; j += blockDim.x) {
        double ui[NC], uo[NC], qa[NP], qb[NP], nf[D], hn[D], phi[NC];
        const size_t jo = (size_t)(g.mapP[(size_t)Nf * k + j] - 1);
#pragma unroll
        for (int e = 0; e < NC; e++) { ui[e] = u_f[(size_t)Nf * k + j + (size_t)g.NFT * e]; uo[e] = u_f[jo + (size_t)g.NFT * e]; }
        const double jf = g.J_f[(size_t)Nf * k + j];
#pragma unroll
        for (int m = 0; m < D; m++) {
            const double nj = g.nJf[m + D * ((size_t)Nf * k + j)];
            nf[m] = nj / jf;                       // operators.jl:59
            hn[m] = 0.5 * nj;                      // halfnJf, operators.jl:78
            s_hnf[m * Nf + j] = hn[m];
        }
        const double ira = to_prim<D, NC>(L, ui, qa);
        const double irb = to_prim<D, NC>(L, uo, qb);
#pragma unroll
        for (int c = 0; c < NP; c++) s_fprim[c * Nf + j] = qa[c];
        double nfs[D];
#pragma unroll
        for (int m = 0; m < D; m++) nfs[m] = PairScale<D, NC>::g * nf[m];
        pair_flux<D, NC>(L, qa, qb, nfs, phi);     // F#(u-, u+) . n   (ConservationLaws.jl:75-128)
        if (L.inviscid == SSE_FLUX_LAX_FRIEDRICHS) {
            double a;
            if constexpr (NC == D + 2) {
                double vni = 0.0, vno = 0.0;
#pragma unroll
                for (int m = 0; m < D; m++) { vni = fma(qa[1 + m], nf[m], vni); vno = fma(qb[1 + m], nf[m], vno); }
                const double ci = sqrt(L.gamma * (0.5 * qa[D + 1]) * ira), co = sqrt(L.gamma * (0.5 * qb[D + 1]) * irb);   // 2p stored
                a = L.half_lambda * (fmax(fabs(vni), fabs(vno)) + fmax(ci, co));
            } else {
                double s = 0.0;
#pragma unroll
                for (int m = 0; m < D; m++) s = fma(L.a[m], nf[m], s);
                a = L.half_lambda * fabs(s);
            }
#pragma unroll
            for (int e = 0; e < NC; e++) phi[e] = fma(a, ui[e] - uo[e], phi[e]);
        }
        const double bj = o.Bf[j] * jf;            // BJf, operators.jl:58
#pragma unroll
        for (int e = 0; e < NC; e++) s_ff[e * Nf + j] = bj * phi[e];
    }
    sse_sync();

    // ---- phase 1: volume flux differencing along tensor lines
    int buf = 0;
    for (int rd = 0; rd < t.n_vrounds; rd++, buf ^= 1) {
        double* st = s_stage + buf * NC * Nq;
        if (tid < Nq) {
            const int j = t.v_partner[rd * Nq + tid];
            if (j >= 0) {
                double gv[D], qj[NP], phi[NC];
#pragma unroll
                for (int n = 0; n < D; n++) gv[n] = 0.0;
                const int mlo = t.v_mlo[rd];
#pragma unroll
                for (int m = 0; m < D; m++) {
                    if (m >= mlo) {                // uniform over the CTA: Λ_ref is upper triangular (tensor_simplex.jl:66-75)
                        const double s = PairScale<D, NC>::g * t.v_S[(rd * D + m) * Nq + tid];
#pragma unroll
                        for (int n = 0; n < D; n++) gv[n] = fma(s, lam[m][n] + s_lam[(m + D * n) * Nq + j], gv[n]);
                    }
                }
#pragma unroll
                for (int c = 0; c < NP; c++) qj[c] = s_prim[c * Nq + j];
                pair_flux<D, NC>(L, qi, qj, gv, phi);
#pragma unroll
                for (int e = 0; e < NC; e++) { r[e] -= phi[e]; st[e * Nq + j] = phi[e]; }
            }
        }
        sse_sync();
        if (tid < Nq && t.v_source[rd * Nq + tid] >= 0) {
#pragma unroll
            for (int e = 0; e < NC; e++) r[e] += st[e * Nq + tid];
        }
    }

    // ---- phase 2: facet correction
    int face_prev = -1;
    double hq[D];
#pragma unroll
    for (int n = 0; n < D; n++) hq[n] = 0.0;
    for (int fr = 0; fr < t.n_frounds; fr++, buf ^= 1) {
        double* st = s_stage + buf * NC * Nq;
        const int f = t.f_face[fr];
        if (tid < Nq) {
            if (f != face_prev) {                  // halfnJq[:, f, i] = 0.5 sum_l Λ[i,l,:] nref[l,f]  (mesh.jl:262-269)
                if (g.nJq) {
#pragma unroll
                    for (int n = 0; n < D; n++) hq[n] = 0.5 * g.nJq[n + D * (f + (size_t)o.Nfac * (tid + (size_t)Nq * k))];
                } else {
#pragma unroll
                    for (int n = 0; n < D; n++) {
                        double s = 0.0;
#pragma unroll
                        for (int l = 0; l < D; l++) s += lam[l][n] * o.nref[l + D * f];
                        hq[n] = 0.5 * s;
                    }
                }
            }
            const int j = t.f_partner[fr * Nq + tid];
            const double c = PairScale<D, NC>::g * t.f_C[fr * Nq + tid];
            double gv[D], qj[NP], phi[NC];
#pragma unroll
            for (int n = 0; n < D; n++) gv[n] = c * (s_hnf[n * Nf + j] + hq[n]);
#pragma unroll
            for (int cc = 0; cc < NP; cc++) qj[cc] = s_fprim[cc * Nf + j];
            pair_flux<D, NC>(L, qi, qj, gv, phi);
#pragma unroll
            for (int e = 0; e < NC; e++) { r[e] -= phi[e]; st[e * Nq + tid] = phi[e]; }
        }
        face_prev = f;
        sse_sync();
        const int nred = t.red_n[fr] * NC;
        for (int q = tid; q < nred; q += blockDim.x) {
            const int item = q / NC, e = q - item * NC;
            const int base = fr * t.red_items_max + item;
            const int* src = t.red_src + (size_t)base * t.red_max;
            const int cnt = t.red_cnt[base];
            double s = 0.0;
            for (int c = 0; c < cnt; c++) s += st[e * Nq + src[c]];
            s_ff[e * Nf + t.red_dst[base]] -= s;
        }
    }
    sse_sync();

    // ---- phase 3: lift, project, mass solve (flux_differencing_form.jl:340-346)
    double* s_r = sm + lay.post_r;
    if (tid < Nq) {
        SSE_ROW_FOR(o.Rt, tid, j, rv) {
#pragma unroll
            for (int e = 0; e < NC; e++) r[e] = fma(-rv, s_ff[e * Nf + j], r[e]);
        }
    }
    sse_sync();                               // every read of s_ff / stage done before s_r and scratch are written
    if (tid < Nq) {
#pragma unroll
        for (int e = 0; e < NC; e++) s_r[e * Nq + tid] = r[e];
    }
    sse_sync();
    double* s_m = sm + lay.post_m;
    apply_Vt<NC>(o, s_r, s_m, sm + lay.post_z, sm + lay.post_w);
    mass_solve<NC>(o, g, k, s_m, sm + lay.post_tq, sm + lay.post_z, sm + lay.post_w);
    SSE_FOR(x, Np * NC) { dudt[(size_t)Np * NC * k + x] = s_m[x]; flag_nonfinite(g.flag, s_m[x]); }
}

// ---------------------------------------------------------------------------------------------
// host: structure detection and table construction
inline void tensor_plan_build(TensorPlan& tp, const sse_config& cfg, const sse_arrays& a, const Ops& o) {
    tp.ok = 0;
    const int d = cfg.d, Nq = cfg.N_q, Nf = cfg.N_f, Nfac = cfg.N_fac;
    if (cfg.form != SSE_FORM_FLUX_DIFFERENCING || d < 2) return;
    if (cfg.pde == SSE_PDE_EULER && cfg.two_point_flux != SSE_TWO_POINT_ENTROPY_CONSERVATIVE) return;
    if (cfg.pde == SSE_PDE_ADVECTION_DIFFUSION || cfg.pde == SSE_PDE_BURGERS || cfg.pde == SSE_PDE_VISCOUS_BURGERS) return;     // scalar path is linear advection
    int N1 = (int)std::lround(std::pow((double)Nq, 1.0 / d));
    int chk = 1;
    for (int m = 0; m < d; m++) chk *= N1;
    if (chk != Nq || N1 < 2 || N1 > 8) return;
    const int threads = ((std::max(Nq, 1) + 31) / 32) * 32;
    const bool runtime_kernel = threads <= 128;      // k_fluxdiff_tensor is compiled for <= 128 threads per element; the schedule
                                                     // tables are still built beyond that (the compile-time kernels take them)
    int stride[3] = {1, 1, 1};
    for (int m = 0; m < d; m++) { stride[m] = 1; for (int mm = m + 1; mm < d; mm++) stride[m] *= N1; }
    auto coord = [&](int i, int m) { return (i / stride[m]) % N1; };
    // --- volume: every S entry must couple two nodes of one tensor line
    for (int i = 0; i < Nq; i++)
        for (int j = i + 1; j < Nq; j++) {
            bool any = false;
            for (int m = 0; m < d; m++) any |= a.S[m][i + (size_t)Nq * j] != 0.0;
            if (!any) continue;
            int ndiff = 0;
            for (int m = 0; m < d; m++) ndiff += coord(i, m) != coord(j, m);
            if (ndiff != 1) return;
        }
    auto Sext = [&](int m, int i, int j) { return i < j ? a.S[m][i + (size_t)Nq * j] : -a.S[m][j + (size_t)Nq * i]; };
    const int nsh = N1 / 2;
    const int nvr = d * nsh;
    tp.v_partner.assign((size_t)nvr * Nq, -1);
    tp.v_source.assign((size_t)nvr * Nq, -1);
    tp.v_S.assign((size_t)nvr * d * Nq, 0.0);
    tp.v_mlo.assign(nvr, 0);
    for (int l = 0; l < d; l++)
        for (int sh = 1; sh <= nsh; sh++) {
            const int rd = l * nsh + (sh - 1);
            int mlo = d;
            for (int i = 0; i < Nq; i++) {
                const int ci = coord(i, l);
                const bool active = !(2 * sh == N1 && ci >= sh);      // even N1: the half-way pairs once only
                if (!active) continue;
                const int cj = (ci + sh) % N1;
                const int j = i + (cj - ci) * stride[l];
                tp.v_partner[(size_t)rd * Nq + i] = j;
                tp.v_source[(size_t)rd * Nq + j] = i;
                for (int m = 0; m < d; m++) {
                    const double s = Sext(m, i, j);
                    tp.v_S[((size_t)rd * d + m) * Nq + i] = s;
                    if (s != 0.0) mlo = std::min(mlo, m);
                }
            }
            tp.v_mlo[rd] = std::min(mlo, d - 1);
        }
    // --- facet correction
    int deg = 0;
    if (a.Cfd) {
        const int npf = Nf / Nfac;
        std::vector<std::vector<int>> J(Nq), I(Nf);
        for (int j = 0; j < Nf; j++)
            for (int i = 0; i < Nq; i++)
                if (a.Cfd[i + (size_t)Nq * j] != 0.0) { J[i].push_back(j); I[j].push_back(i); }
        // group sizes per face must be the same for every volume node
        std::vector<int> gsz(Nfac, 0), gbase(Nfac, 0);
        for (int j : J[0]) gsz[j / npf]++;
        for (int i = 0; i < Nq; i++) {
            std::vector<int> gi(Nfac, 0);
            for (int j : J[i]) gi[j / npf]++;
            if (gi != gsz) return;
        }
        for (int f = 0; f < Nfac; f++) { gbase[f] = deg; deg += gsz[f]; }
        if (deg == 0) return;
        tp.f_partner.assign((size_t)deg * Nq, -1);
        tp.f_C.assign((size_t)deg * Nq, 0.0);
        tp.f_face.assign(deg, 0);
        for (int f = 0; f < Nfac; f++) for (int q = 0; q < gsz[f]; q++) tp.f_face[gbase[f] + q] = f;
        for (int rot = 1; rot >= 0; rot--) {          // try the balanced rotation first, plain ranks otherwise
            std::fill(tp.f_partner.begin(), tp.f_partner.end(), -1);
            bool valid = true;
            for (int i = 0; i < Nq && valid; i++) {
                std::vector<int> rank_in_group(Nfac, 0);
                for (int j : J[i]) {
                    const int f = j / npf;
                    const int pos = (int)(std::find(I[j].begin(), I[j].end(), i) - I[j].begin());
                    const int shift = rot ? pos % gsz[f] : 0;
                    const int fr = gbase[f] + (rank_in_group[f] + shift) % gsz[f];
                    rank_in_group[f]++;
                    if (tp.f_partner[(size_t)fr * Nq + i] >= 0) { valid = false; break; }
                    tp.f_partner[(size_t)fr * Nq + i] = j;
                    tp.f_C[(size_t)fr * Nq + i] = a.Cfd[i + (size_t)Nq * j];
                }
            }
            if (valid) break;
            if (!rot) return;
        }
        for (size_t x = 0; x < tp.f_partner.size(); x++) if (tp.f_partner[x] < 0) return;
        // reducers
        std::vector<std::vector<std::vector<int>>> src(deg, std::vector<std::vector<int>>(Nf));
        for (int fr = 0; fr < deg; fr++)
            for (int i = 0; i < Nq; i++) src[fr][tp.f_partner[(size_t)fr * Nq + i]].push_back(i);
        int items_max = 0, red_max = 0;
        for (int fr = 0; fr < deg; fr++) {
            int n = 0;
            for (int j = 0; j < Nf; j++) if (!src[fr][j].empty()) { n++; red_max = std::max(red_max, (int)src[fr][j].size()); }
            items_max = std::max(items_max, n);
        }
        tp.red_n.assign(deg, 0);
        tp.red_dst.assign((size_t)deg * items_max, 0);
        tp.red_cnt.assign((size_t)deg * items_max, 0);
        tp.red_src.assign((size_t)deg * items_max * red_max, 0);
        for (int fr = 0; fr < deg; fr++) {
            int n = 0;
            for (int j = 0; j < Nf; j++) {
                if (src[fr][j].empty()) continue;
                const size_t base = (size_t)fr * items_max + n;
                tp.red_dst[base] = j;
                tp.red_cnt[base] = (int)src[fr][j].size();
                for (size_t c = 0; c < src[fr][j].size(); c++) tp.red_src[base * red_max + c] = src[fr][j][c];
                n++;
            }
            tp.red_n[fr] = n;
        }
        tp.dev.red_items_max = items_max;
        tp.dev.red_max = red_max;
    }
    tp.dev.N1 = N1;
    tp.dev.n_vrounds = nvr;
    tp.dev.n_frounds = deg;
    tp.threads = threads;
    const int NP = (cfg.N_c == d + 2) ? d + 3 : 1;
    tp.smem_fluxdiff = sizeof(double) * (size_t)fd_layout(o, d, cfg.N_c, NP).total;
    if (tp.smem_fluxdiff > 227 * 1024) return;
    tp.has_fluxdiff = runtime_kernel ? 1 : 0;
    tp.ok = 1;
}

// Host-only emulation of the pair schedule: rebuilds S_m and C from the tables and returns the largest
// deviation from the operators handed in (also checks that every pair is visited exactly once).
inline double tensor_plan_selfcheck(const TensorPlan& tp, const sse_config& cfg, const sse_arrays& a) {
    const int d = cfg.d, Nq = cfg.N_q, Nf = cfg.N_f;
    double err = 0.0;
    std::vector<double> S((size_t)d * Nq * Nq, 0.0);
    std::vector<int> hits((size_t)Nq * Nq, 0);
    for (int rd = 0; rd < tp.dev.n_vrounds; rd++)
        for (int i = 0; i < Nq; i++) {
            const int j = tp.v_partner[(size_t)rd * Nq + i];
            if (j < 0) continue;
            if (tp.v_source[(size_t)rd * Nq + j] != i) err = 1e300;
            hits[std::min(i, j) + (size_t)Nq * std::max(i, j)]++;
            for (int m = 0; m < d; m++) {
                const double s = tp.v_S[((size_t)rd * d + m) * Nq + i];
                if (m < tp.v_mlo[rd] && s != 0.0) err = 1e300;
                S[(size_t)m * Nq * Nq + i + (size_t)Nq * j] += s;      // r_i -= s * (...)
                S[(size_t)m * Nq * Nq + j + (size_t)Nq * i] -= s;      // r_j += s * (...)
            }
        }
    for (int i = 0; i < Nq; i++)
        for (int j = i + 1; j < Nq; j++) {
            bool any = false;
            for (int m = 0; m < d; m++) {
                const double ref = a.S[m][i + (size_t)Nq * j];
                any |= ref != 0.0;
                err = std::max(err, std::fabs(S[(size_t)m * Nq * Nq + i + (size_t)Nq * j] - ref));
                err = std::max(err, std::fabs(S[(size_t)m * Nq * Nq + j + (size_t)Nq * i] + ref));
            }
            if (hits[i + (size_t)Nq * j] > 1 || (any && hits[i + (size_t)Nq * j] != 1)) err = 1e300;
        }
    if (a.Cfd) {
        std::vector<double> C((size_t)Nq * Nf, 0.0), Cred((size_t)Nq * Nf, 0.0);
        for (int fr = 0; fr < tp.dev.n_frounds; fr++) {
            for (int i = 0; i < Nq; i++) {
                const int j = tp.f_partner[(size_t)fr * Nq + i];
                if (j < 0 || j / (Nf / cfg.N_fac) != tp.f_face[fr]) { err = 1e300; continue; }
                C[i + (size_t)Nq * j] += tp.f_C[(size_t)fr * Nq + i];
            }
            for (int it = 0; it < tp.red_n[fr]; it++) {
                const size_t base = (size_t)fr * tp.dev.red_items_max + it;
                for (int c = 0; c < tp.red_cnt[base]; c++) {
                    const int i = tp.red_src[base * tp.dev.red_max + c];
                    if (tp.f_partner[(size_t)fr * Nq + i] != tp.red_dst[base]) err = 1e300;
                    Cred[i + (size_t)Nq * tp.red_dst[base]] += tp.f_C[(size_t)fr * Nq + i];
                }
            }
        }
        for (size_t x = 0; x < C.size(); x++) {
            err = std::max(err, std::fabs(C[x] - a.Cfd[x]));
            err = std::max(err, std::fabs(Cred[x] - a.Cfd[x]));
        }
    }
    return err;
}

template <class F>
inline int32_t tensor_plan_upload(TensorPlan& tp, F up) {
    int32_t rc;
    const void* p;
#define UPV(vec, field, T)                                                                     \
    do {                                                                                       \
        if ((rc = up(tp.vec.data(), tp.vec.size() * sizeof(T), &p))) return rc;                \
        tp.dev.field = (const T*)p;                                                            \
    } while (0)
    UPV(v_partner, v_partner, int);
    UPV(v_source, v_source, int);
    UPV(v_S, v_S, double);
    UPV(v_mlo, v_mlo, int);
    UPV(f_partner, f_partner, int);
    UPV(f_C, f_C, double);
    UPV(f_face, f_face, int);
    UPV(red_n, red_n, int);
    UPV(red_dst, red_dst, int);
    UPV(red_cnt, red_cnt, int);
    UPV(red_src, red_src, int);
#undef UPV
    return SSE_OK;
}

template <int D, int NC>
inline cudaError_t tensor_set_attrs(const TensorPlan& tp) {
    if (!tp.ok) return cudaSuccess;
    if constexpr (D >= 2) {
        return cudaFuncSetAttribute(k_fluxdiff_tensor<D, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);   // per kernel, never lowered by a later handle
    }
    return cudaSuccess;
}

template <int D, int NC>
inline void tensor_launch_nodal(const TensorPlan&, const Ops&, const Geo&, const Law&, int, const double*, double*, double*,
                                long long, int, cudaStream_t) {}

template <int D, int NC>
inline void tensor_launch_fluxdiff(const TensorPlan& tp, const Ops& o, const Geo& g, const Law& L, unsigned grid, dim3 block, size_t smem,
                                   long long first, const double* u_q, const double* u_f, double* dudt, cudaStream_t s) {
    if constexpr (D >= 2) {
        k_fluxdiff_tensor<D, NC><<<grid, block, smem, s>>>(tp.dev, o, g, L, first, u_q, u_f, dudt);     // block = (tp.threads, rows)
    }
}

}  // namespace sse
