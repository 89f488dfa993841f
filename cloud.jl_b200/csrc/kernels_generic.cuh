// kernels_generic.cuh — element-per-CTA kernels for ANY operator set the reference accepts
// (dense / sparse / warped operators, d = 1..3).  They are the correctness backbone and the
// fallback for operator families without tensor-line structure; the tensor-product
// simplex paths of the BASELINE configs run the specialised kernels in kernels_tensor.cuh.
//
// One CTA owns one element; all per-element tiles live in shared memory; threads stride
// over (node, variable) items.  Pair contributions are evaluated "row-wise" (each thread
// accumulates only its own node), so no atomics are needed.
#pragma once
#include "common.cuh"

namespace sse {

// ------------------------------------------------------------------ V, V^T on shared tiles
// y (Nq x NC) = V x (Np x NC).  zb, wb: scratch of NC*P1*P1*M3 and NC*P1*M2*M3 doubles.
// warped_product_3d.jl:47-84 (2-D: warped_product_2d.jl:31-55 embedded with M3 = 1)
template <int NC>
__device__ void apply_V(const Ops& o, const double* x, double* y, double* zb, double* wb) {
    const int Nq = o.Nq, Np = o.Np;
    if (o.v_kind == SSE_V_IDENTITY) {
        SSE_FOR(t, Nq * NC) y[t] = x[t];
        sse_sync();
        return;
    }
    if (o.v_kind == SSE_V_DENSE || o.v_small) {
        // one node per thread, all variables at once: the entry of V is loaded once (coalesced over the nodes) and feeds NC
        // independent accumulators; same summation order per (node, variable) as a plain dot product
        SSE_FOR(i, Nq) {
            double acc[NC];
#pragma unroll
            for (int e = 0; e < NC; e++) acc[e] = 0.0;
#pragma unroll 4
            for (int j = 0; j < Np; j++) {
                const double v = o.Vd[i + (size_t)Nq * j];
#pragma unroll
                for (int e = 0; e < NC; e++) acc[e] = fma(v, x[j + Np * e], acc[e]);
            }
#pragma unroll
            for (int e = 0; e < NC; e++) y[i + Nq * e] = acc[e];
        }
        sse_sync();
        return;
    }
    const int P1 = o.P1, M1 = o.M1, M2 = o.M2, M3 = o.M3;
    SSE_FOR(t, NC * P1 * P1 * M3) {
        int a3 = t % M3, b2 = (t / M3) % P1, b1 = (t / (M3 * P1)) % P1, e = t / (M3 * P1 * P1);
        if (b2 < o.N2[b1]) {
            double s = 0.0;
            for (int b3 = 0; b3 < o.N3[b1 * 8 + b2]; b3++)
                s = fma(o.C[a3 + M3 * (b1 + P1 * (b2 + P1 * b3))], x[o.sig_i[b1 + P1 * (b2 + P1 * b3)] + Np * e], s);
            zb[t] = s;
        }
    }
    sse_sync();
    SSE_FOR(t, NC * P1 * M2 * M3) {
        int a3 = t % M3, a2 = (t / M3) % M2, b1 = (t / (M3 * M2)) % P1, e = t / (M3 * M2 * P1);
        double s = 0.0;
        for (int b2 = 0; b2 < o.N2[b1]; b2++)
            s = fma(o.B[a2 + M2 * (b1 + P1 * b2)], zb[((e * P1 + b1) * P1 + b2) * M3 + a3], s);
        wb[t] = s;
    }
    sse_sync();
    SSE_FOR(t, NC * M1 * M2 * M3) {
        int a3 = t % M3, a2 = (t / M3) % M2, a1 = (t / (M3 * M2)) % M1, e = t / (M3 * M2 * M1);
        double s = 0.0;
        for (int b1 = 0; b1 < P1; b1++) s = fma(o.A[a1 + M1 * b1], wb[((e * P1 + b1) * M2 + a2) * M3 + a3], s);
        y[o.sig_o[a1 + M1 * (a2 + M2 * a3)] + Nq * e] = s;
    }
    sse_sync();
}

// y (Np x NC) = V^T x (Nq x NC).  warped_product_3d.jl:94-136 / warped_product_2d.jl:61-89
template <int NC>
__device__ void apply_Vt(const Ops& o, const double* x, double* y, double* zb, double* wb) {
    const int Nq = o.Nq, Np = o.Np;
    if (o.v_kind == SSE_V_IDENTITY) {
        SSE_FOR(t, Nq * NC) y[t] = x[t];
        sse_sync();
        return;
    }
    if (o.v_kind == SSE_V_DENSE || o.v_small) {
        // one mode per thread, all variables at once (NC independent accumulators, fixed trip count: the loads pipeline)
        SSE_FOR(j, Np) {
            double acc[NC];
#pragma unroll
            for (int e = 0; e < NC; e++) acc[e] = 0.0;
            const double* vj = o.Vd + (size_t)Nq * j;
#pragma unroll 4
            for (int i = 0; i < Nq; i++) {
                const double v = vj[i];
#pragma unroll
                for (int e = 0; e < NC; e++) acc[e] = fma(v, x[i + Nq * e], acc[e]);
            }
#pragma unroll
            for (int e = 0; e < NC; e++) y[j + Np * e] = acc[e];
        }
        sse_sync();
        return;
    }
    const int P1 = o.P1, M1 = o.M1, M2 = o.M2, M3 = o.M3;
    SSE_FOR(t, NC * P1 * M2 * M3) {
        int a3 = t % M3, a2 = (t / M3) % M2, b1 = (t / (M3 * M2)) % P1, e = t / (M3 * M2 * P1);
        double s = 0.0;
        for (int a1 = 0; a1 < M1; a1++) s = fma(o.A[a1 + M1 * b1], x[o.sig_o[a1 + M1 * (a2 + M2 * a3)] + Nq * e], s);
        wb[t] = s;
    }
    sse_sync();
    SSE_FOR(t, NC * P1 * P1 * M3) {
        int a3 = t % M3, b2 = (t / M3) % P1, b1 = (t / (M3 * P1)) % P1, e = t / (M3 * P1 * P1);
        if (b2 < o.N2[b1]) {
            double s = 0.0;
            for (int a2 = 0; a2 < M2; a2++) s = fma(o.B[a2 + M2 * (b1 + P1 * b2)], wb[((e * P1 + b1) * M2 + a2) * M3 + a3], s);
            zb[t] = s;
        }
    }
    sse_sync();
    SSE_FOR(t, NC * P1 * P1 * P1) {
        int b3 = t % P1, b2 = (t / P1) % P1, b1 = (t / (P1 * P1)) % P1, e = t / (P1 * P1 * P1);
        if (b2 < o.N2[b1] && b3 < o.N3[b1 * 8 + b2]) {
            double s = 0.0;
            for (int a3 = 0; a3 < M3; a3++) s = fma(o.C[a3 + M3 * (b1 + P1 * (b2 + P1 * b3))], zb[((e * P1 + b1) * P1 + b2) * M3 + a3], s);
            y[o.sig_i[b1 + P1 * (b2 + P1 * b3)] + Np * e] = s;
        }
    }
    sse_sync();
}

// y (nrow x NC) = A x, rows of A compressed; ldx / ldy are the leading dimensions of the tiles
template <int NC>
__device__ void apply_sp(const SpMat& A, int nrow, const double* x, int ldx, double* y, int ldy) {
    SSE_FOR(t, nrow * NC) {
        int i = t % nrow, e = t / nrow;
        double s = 0.0;
        SSE_ROW_FOR(A, i, c_, v_) s = fma(v_, x[c_ + ldx * e], s);
        y[i + ldy * e] = s;
    }
    sse_sync();
}

__host__ __device__ inline int warp_z_size(const Ops& o, int NC) { return o.v_kind == SSE_V_WARPED ? NC * o.P1 * o.P1 * o.M3 : 0; }
__host__ __device__ inline int warp_w_size(const Ops& o, int NC) { return o.v_kind == SSE_V_WARPED ? NC * o.P1 * o.M2 * o.M3 : 0; }

// mass_matrix_solve! (mass_matrix.jl:169-196) on a shared Np x NC tile; tq: Nq x NC scratch
template <int NC>
__device__ void mass_solve(const Ops& o, const Geo& g, long long k, double* rhs, double* tq, double* zb, double* wb) {
    const double* J = g.J_q + (size_t)o.Nq * k;
    if (g.mass_solver == SSE_MASS_DIAGONAL) {
        SSE_FOR(t, o.Np * NC) { int i = t % o.Np; rhs[t] *= 1.0 / (o.W[i] * J[i]); }
        sse_sync();
        return;
    }
    if (g.mass_solver == SSE_MASS_CHOLESKY) {
        // ldiv!(cholesky(Symmetric(V' WJ_k V)), rhs) (mass_matrix.jl:169-175): U' y = rhs, then U x = y, column sweeps
        // over the shared tile; the (i, variable) updates of one column run in parallel
        const int Np = o.Np;
        const double* U = g.chol + (size_t)Np * Np * k;
        for (int j = 0; j < Np; j++) {
            if (threadIdx.x < NC) rhs[j + Np * threadIdx.x] /= U[j + (size_t)Np * j];
            sse_sync();
            SSE_FOR(t, (Np - 1 - j) * NC) {
                const int i = j + 1 + t % (Np - 1 - j), e = t / (Np - 1 - j);
                rhs[i + Np * e] = fma(-U[j + (size_t)Np * i], rhs[j + Np * e], rhs[i + Np * e]);
            }
            sse_sync();
        }
        for (int j = Np - 1; j >= 0; j--) {
            if (threadIdx.x < NC) rhs[j + Np * threadIdx.x] /= U[j + (size_t)Np * j];
            sse_sync();
            SSE_FOR(t, j * NC) {
                const int i = t % j, e = t / j;
                rhs[i + Np * e] = fma(-U[i + (size_t)Np * j], rhs[j + Np * e], rhs[i + Np * e]);
            }
            sse_sync();
        }
        return;
    }
    apply_V<NC>(o, rhs, tq, zb, wb);
    SSE_FOR(i, o.Nq) {                           // one quotient per node (the same W / J every variable was multiplied by)
        const double wj = o.W[i] / J[i];
#pragma unroll
        for (int e = 0; e < NC; e++) tq[i + o.Nq * e] *= wj;
    }
    sse_sync();
    apply_Vt<NC>(o, tq, rhs, zb, wb);
}

// CholeskySolver constructor (mass_matrix.jl:30-39) on the device: M_k = V' diag(W J_k) V column by column through the
// matrix-free V, then its upper Cholesky factor in place (right-looking); one CTA per element.
// shared: M (Np x Np) | e (Np) | q (Nq) | z | w
static __global__ void k_cholesky_factor(Ops o, Geo g, double* __restrict__ chol, int* __restrict__ bad) {
    extern __shared__ double sm[];
    const int Np = o.Np, Nq = o.Nq;
    const long long k = blockIdx.x;
    double* s_M = sm;
    double* s_e = s_M + Np * Np;
    double* s_q = s_e + Np;
    double* s_z = s_q + Nq;
    double* s_w = s_z + warp_z_size(o, 1);
    const double* J = g.J_q + (size_t)Nq * k;
    for (int c = 0; c < Np; c++) {
        SSE_FOR(t, Np) s_e[t] = (t == c) ? 1.0 : 0.0;
        sse_sync();
        apply_V<1>(o, s_e, s_q, s_z, s_w);
        SSE_FOR(i, Nq) s_q[i] *= o.W[i] * J[i];
        sse_sync();
        apply_Vt<1>(o, s_q, s_M + Np * c, s_z, s_w);
    }
    for (int j = 0; j < Np; j++) {
        const double piv = s_M[j + Np * j];
        sse_sync();
        if (!(piv > 0.0)) { if (threadIdx.x == 0) *bad = 1; return; }      // PosDefException in the reference (Solvers.jl:411-412)
        const double ujj = sqrt(piv);
        SSE_FOR(i, Np - j) s_M[j + Np * (j + i)] = (i == 0) ? ujj : s_M[j + Np * (j + i)] / ujj;
        sse_sync();
        const int n = Np - 1 - j;
        SSE_FOR(t, n * n) {
            const int i = j + 1 + t % n, l = j + 1 + t / n;
            if (i <= l) s_M[i + Np * l] = fma(-s_M[j + Np * i], s_M[j + Np * l], s_M[i + Np * l]);
        }
        sse_sync();
    }
    SSE_FOR(t, Np * Np) chol[(size_t)Np * Np * k + t] = (t % Np <= t / Np) ? s_M[t] : 0.0;
}


// ------------------------------------------------------------------ pass A: nodal_values!
// standard_form_first_order.jl:1-14; flux_differencing_form.jl:171-292.
// project: 0 = none, 1 = nodal (w_f = R w(u_q)), 2 = general/modal entropy projection.
// NV = number of "variables" moved per node = NC.
template <int D, int NC>
__global__ void k_nodal_generic(Ops o, Geo g, Law L, int project, long long first, const double* __restrict__ u,
                                double* __restrict__ u_q, double* __restrict__ u_f) {
    extern __shared__ double sm_cta[];
    double* sm = sse_row_smem(sm_cta);
    const long long k = sse_element(first);
    const int Nq = o.Nq, Np = o.Np, Nf = o.Nf;
    double* s_u = sm;                       // Np x NC
    double* s_a = s_u + Np * NC;            // Nq x NC
    double* s_b = s_a + Nq * NC;            // Nq x NC
    double* s_f = s_b + Nq * NC;            // Nf x NC
    double* s_z = s_f + Nf * NC;
    double* s_w = s_z + warp_z_size(o, NC);

    SSE_FOR(t, Np * NC) s_u[t] = u[(size_t)Np * NC * k + t];
    sse_sync();
    apply_V<NC>(o, s_u, s_a, s_z, s_w);                           // u_q = V u
    if (project == 0) {
        apply_sp<NC>(o.R, Nf, s_a, Nq, s_f, Nf);                  // u_f = R u_q
    } else if (project == 1) {
        SSE_FOR(i, Nq) {
            double ui[NC], wi[NC];
#pragma unroll
            for (int e = 0; e < NC; e++) ui[e] = s_a[i + Nq * e];
            cons_to_entropy<D, NC>(L, ui, wi);
#pragma unroll
            for (int e = 0; e < NC; e++) s_b[i + Nq * e] = wi[e];
        }
        sse_sync();
        apply_sp<NC>(o.R, Nf, s_b, Nq, s_f, Nf);                  // w_f = R w_q
        SSE_FOR(i, Nf) {
            double wi[NC], ui[NC];
#pragma unroll
            for (int e = 0; e < NC; e++) wi[e] = s_f[i + Nf * e];
            entropy_to_cons<D, NC>(L, wi, ui);
#pragma unroll
            for (int e = 0; e < NC; e++) s_f[i + Nf * e] = ui[e];
        }
        sse_sync();
    } else {
        const double* J = g.J_q + (size_t)Nq * k;
        SSE_FOR(i, Nq) {                                           // w_q = WJ * w(u_q)
            double ui[NC], wi[NC];
#pragma unroll
            for (int e = 0; e < NC; e++) ui[e] = s_a[i + Nq * e];
            cons_to_entropy<D, NC>(L, ui, wi);
            double wj = o.W[i] * J[i];
#pragma unroll
            for (int e = 0; e < NC; e++) s_b[i + Nq * e] = wi[e] * wj;
        }
        sse_sync();
        apply_Vt<NC>(o, s_b, s_u, s_z, s_w);                       // w = V' w_q
        mass_solve<NC>(o, g, k, s_u, s_b, s_z, s_w);               // w = M \ w
        apply_V<NC>(o, s_u, s_b, s_z, s_w);                        // w_q = V w
        apply_sp<NC>(o.R, Nf, s_b, Nq, s_f, Nf);                   // w_f = R w_q
        SSE_FOR(i, Nq + Nf) {
            double wi[NC], ui[NC];
            if (i < Nq) {
#pragma unroll
                for (int e = 0; e < NC; e++) wi[e] = s_b[i + Nq * e];
                entropy_to_cons<D, NC>(L, wi, ui);
#pragma unroll
                for (int e = 0; e < NC; e++) s_a[i + Nq * e] = ui[e];
            } else {
                int j = i - Nq;
#pragma unroll
                for (int e = 0; e < NC; e++) wi[e] = s_f[j + Nf * e];
                entropy_to_cons<D, NC>(L, wi, ui);
#pragma unroll
                for (int e = 0; e < NC; e++) s_f[j + Nf * e] = ui[e];
            }
        }
        sse_sync();
    }
    SSE_FOR(t, Nq * NC) u_q[(size_t)Nq * NC * k + t] = s_a[t];
    SSE_FOR(t, Nf * NC) { int i = t % Nf, e = t / Nf; u_f[(size_t)Nf * k + i + (size_t)g.NFT * e] = s_f[t]; }
}

// ------------------------------------------------------------------ shared pieces of pass B
// interior/exterior facet states and unit normals of element k into shared tiles
template <int D, int NC>
__device__ void load_facets(const Ops& o, const Geo& g, long long k, const double* __restrict__ u_f,
                            double* s_in, double* s_out, double* s_nf) {
    const int Nf = o.Nf;
    SSE_FOR(t, Nf * NC) {
        int i = t % Nf, e = t / Nf;
        s_in[t] = u_f[(size_t)Nf * k + i + (size_t)g.NFT * e];
        s_out[t] = u_f[(size_t)(g.mapP[(size_t)Nf * k + i] - 1) + (size_t)g.NFT * e];
    }
    SSE_FOR(i, Nf) {
        double jf = g.J_f[(size_t)Nf * k + i];
#pragma unroll
        for (int m = 0; m < D; m++) s_nf[m + D * i] = g.nJf[m + D * ((size_t)Nf * k + i)] / jf;   // operators.jl:19,59
    }
    sse_sync();
}

// ------------------------------------------------------------------ pass B: flux differencing
// time_derivative! flux_differencing_form.jl:294-347 with flux_difference! (:37-75) and
// facet_correction! (:126-168) evaluated row-wise.
template <int D, int NC>
__global__ void k_time_fluxdiff_generic(Ops o, Geo g, Law L, long long first, const double* __restrict__ u_q,
                                        const double* __restrict__ u_f, double* __restrict__ dudt) {
    extern __shared__ double sm_cta[];
    double* sm = sse_row_smem(sm_cta);
    const long long k = sse_element(first);
    const int Nq = o.Nq, Np = o.Np, Nf = o.Nf;
    double* s_uq = sm;                       // Nq x NC
    double* s_r = s_uq + Nq * NC;            // Nq x NC
    double* s_in = s_r + Nq * NC;            // Nf x NC
    double* s_out = s_in + Nf * NC;          // Nf x NC
    double* s_ff = s_out + Nf * NC;          // Nf x NC
    double* s_nf = s_ff + Nf * NC;           // D x Nf
    double* s_lam = s_nf + D * Nf;           // Nq x D x D
    double* s_m = s_lam + Nq * D * D;        // Np x NC
    double* s_z = s_m + Np * NC;
    double* s_w = s_z + warp_z_size(o, NC);
    // Euler with the entropy-conservative two-point flux: every pair is evaluated in the contracted, scaled form of the
    // compile-time kernels (ec_contract_scaled: primitives (rho, V, 2p, rho/p) per node, one shared reciprocal for both
    // log-means, no flux tensor, no division per pair) -- about a third of the instructions of two_point_flux + contraction.
    // Dense operators (ModalMulti / NodalMulti, multidimensional.jl) evaluate N_q^2 + 2 N_q N_f pairs per element, so this is
    // what decides their speed.
    constexpr int NPR = D + 3;
    const bool fast_ec = (NC == D + 2) && L.pde == SSE_PDE_EULER && L.two_point == SSE_TWO_POINT_ENTROPY_CONSERVATIVE;
    double* s_prim = s_w + warp_w_size(o, NC);   // NPR x Nq   (fast_ec only)
    double* s_fprim = s_prim + NPR * Nq;         // NPR x Nf
    double* s_hnf = (NC == D + 2) ? s_fprim + NPR * Nf : s_prim;   // D x Nf   halfnJf (operators.jl:78)

    SSE_FOR(t, Nq * NC) s_uq[t] = u_q[(size_t)Nq * NC * k + t];
    SSE_FOR(t, Nq * D * D) s_lam[t] = g.Lambda_q[(size_t)Nq * D * D * k + t];
    load_facets<D, NC>(o, g, k, u_f, s_in, s_out, s_nf);
    SSE_FOR(t, D * Nf) s_hnf[t] = 0.5 * g.nJf[(size_t)D * Nf * k + t];
    if constexpr (NC == D + 2) {
        if (fast_ec) {
            SSE_FOR(i, Nq) {
                double ui[NC], q[NPR];
#pragma unroll
                for (int e = 0; e < NC; e++) ui[e] = s_uq[i + Nq * e];
                to_prim_fast<D>(L, ui, q);
#pragma unroll
                for (int c = 0; c < NPR; c++) s_prim[i + Nq * c] = q[c];
            }
            SSE_FOR(j, Nf) {
                double ui[NC], q[NPR];
#pragma unroll
                for (int e = 0; e < NC; e++) ui[e] = s_in[j + Nf * e];
                to_prim_fast<D>(L, ui, q);
#pragma unroll
                for (int c = 0; c < NPR; c++) s_fprim[j + Nf * c] = q[c];
            }
        }
    }
    sse_sync();

    // facet side: f_f = BJf f* - sum_i C_ij (F(u_i, u_fj) . nJ_ij)
    SSE_FOR(j, Nf) {
        double ui[NC], uo[NC], fs[NC];
#pragma unroll
        for (int e = 0; e < NC; e++) { ui[e] = s_in[j + Nf * e]; uo[e] = s_out[j + Nf * e]; }
        numerical_flux<D, NC>(L, L.two_point, ui, uo, s_nf + D * j, fs);
        double bj = o.Bf[j] * g.J_f[(size_t)Nf * k + j];
#pragma unroll
        for (int e = 0; e < NC; e++) fs[e] *= bj;
        if (o.has_C) {
            const int f = j / o.npf;
            double hf[D], nr[D];
#pragma unroll
            for (int m = 0; m < D; m++) { hf[m] = s_hnf[m + D * j]; nr[m] = o.nref[m + D * f]; }
            double qf[NPR];
            if constexpr (NC == D + 2) {
                if (fast_ec) {
#pragma unroll
                    for (int c = 0; c < NPR; c++) qf[c] = s_fprim[j + Nf * c];
                }
            }
            const int cfw = o.cf_w < 0 ? -o.cf_w : o.cf_w, cfq = o.cf_w < 0 ? 1 : Nf, cfr = o.cf_w < 0 ? cfw : 1;
            for (int q = 0; q < cfw; q++) {
                const int i = o.cf_ie[q * cfq + j * cfr];
                if (i < 0) break;
                double uq[NC], F[NC][D], nJ[D];
                if (!fast_ec) {
#pragma unroll
                    for (int e = 0; e < NC; e++) uq[e] = s_uq[i + Nq * e];
                    two_point_flux<D, NC>(L, L.two_point, uq, ui, F);
                }
#pragma unroll
                for (int m = 0; m < D; m++) {
                    double hq;
                    if (g.nJq) hq = 0.5 * g.nJq[m + D * (f + (size_t)o.Nfac * (i + (size_t)Nq * k))];
                    else {
                        double t = 0.0;
#pragma unroll
                        for (int l = 0; l < D; l++) t += s_lam[i + Nq * (l + D * m)] * nr[l];
                        hq = 0.5 * t;
                    }
                    nJ[m] = hf[m] + hq;
                }
                const double cij = o.cf_ve[q * cfq + j * cfr];
                if constexpr (NC == D + 2) {
                    if (fast_ec) {
                        double qi[NPR], gq[D], phi[NC];
#pragma unroll
                        for (int c = 0; c < NPR; c++) qi[c] = s_prim[i + Nq * c];
#pragma unroll
                        for (int m = 0; m < D; m++) gq[m] = (0.25 * cij) * nJ[m];
                        ec_contract_scaled<D>(L, qi, qf, gq, phi);
#pragma unroll
                        for (int e = 0; e < NC; e++) fs[e] -= phi[e];
                        continue;
                    }
                }
#pragma unroll
                for (int e = 0; e < NC; e++) {
                    double Fn = 0.0;
#pragma unroll
                    for (int m = 0; m < D; m++) Fn = fma(nJ[m], F[e][m], Fn);
                    fs[e] -= cij * Fn;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < NC; e++) s_ff[j + Nf * e] = fs[e];
    }
    // volume side: r_i = -sum_j sum_m S_m[i,j] (Lam_i + Lam_j)[m,:] . F(u_i,u_j) - sum_j C_ij F(u_i,u_fj).nJ_ij
    SSE_FOR(i, Nq) {
        double ui[NC], r[NC];
#pragma unroll
        for (int e = 0; e < NC; e++) { ui[e] = s_uq[i + Nq * e]; r[e] = 0.0; }
        double qi[NPR];
        if constexpr (NC == D + 2) {
            if (fast_ec) {
#pragma unroll
                for (int c = 0; c < NPR; c++) qi[c] = s_prim[i + Nq * c];
            }
        }
        const int vw = o.vol_w < 0 ? -o.vol_w : o.vol_w, vq = o.vol_w < 0 ? 1 : Nq, vr = o.vol_w < 0 ? vw : 1;
        for (int q = 0; q < vw; q++) {
            const int j = o.vol_je[q * vq + i * vr];
            if (j < 0) break;
            if constexpr (NC == D + 2) {
                if (fast_ec) {
                    double qj[NPR], gq[D], phi[NC];
#pragma unroll
                    for (int c = 0; c < NPR; c++) qj[c] = s_prim[j + Nq * c];
#pragma unroll
                    for (int n = 0; n < D; n++) gq[n] = 0.0;
#pragma unroll
                    for (int m = 0; m < D; m++) {
                        const double Sm = 0.25 * o.vol_Se[o.vol_w < 0 ? (i * vw + q) * D + m : (q * D + m) * Nq + i];
#pragma unroll
                        for (int n = 0; n < D; n++) gq[n] = fma(Sm, s_lam[i + Nq * (m + D * n)] + s_lam[j + Nq * (m + D * n)], gq[n]);
                    }
                    ec_contract_scaled<D>(L, qi, qj, gq, phi);
#pragma unroll
                    for (int e = 0; e < NC; e++) r[e] -= phi[e];
                    continue;
                }
            }
            double uj[NC], F[NC][D];
#pragma unroll
            for (int e = 0; e < NC; e++) uj[e] = s_uq[j + Nq * e];
            two_point_flux<D, NC>(L, L.two_point, ui, uj, F);
#pragma unroll
            for (int m = 0; m < D; m++) {
                const double Sm = o.vol_Se[o.vol_w < 0 ? (i * vw + q) * D + m : (q * D + m) * Nq + i];
                if (Sm != 0.0) {
#pragma unroll
                    for (int e = 0; e < NC; e++) {
                        double Fm = 0.0;
#pragma unroll
                        for (int n = 0; n < D; n++) Fm = fma(s_lam[i + Nq * (m + D * n)] + s_lam[j + Nq * (m + D * n)], F[e][n], Fm);
                        r[e] -= Sm * Fm;
                    }
                }
            }
        }
        if (o.has_C) {
            int f_prev = -1;
            double hqv[D];
#pragma unroll
            for (int m = 0; m < D; m++) hqv[m] = 0.0;
            const int cqw = o.cq_w < 0 ? -o.cq_w : o.cq_w, cqq = o.cq_w < 0 ? 1 : Nq, cqr = o.cq_w < 0 ? cqw : 1;
            for (int q = 0; q < cqw; q++) {
                const int j = o.cq_je[q * cqq + i * cqr];
                if (j < 0) break;
                const int f = j / o.npf;
                double uj[NC], F[NC][D], nJ[D];
                if (!fast_ec) {
#pragma unroll
                    for (int e = 0; e < NC; e++) uj[e] = s_in[j + Nf * e];
                    two_point_flux<D, NC>(L, L.two_point, ui, uj, F);
                }
                if (f != f_prev) {                     // halfnJq depends on (node, face) only: mesh.jl:262-269
                    f_prev = f;
#pragma unroll
                    for (int m = 0; m < D; m++) {
                        if (g.nJq) hqv[m] = 0.5 * g.nJq[m + D * (f + (size_t)o.Nfac * (i + (size_t)Nq * k))];
                        else {
                            double t = 0.0;
#pragma unroll
                            for (int l = 0; l < D; l++) t += s_lam[i + Nq * (l + D * m)] * o.nref[l + D * f];
                            hqv[m] = 0.5 * t;
                        }
                    }
                }
#pragma unroll
                for (int m = 0; m < D; m++) nJ[m] = s_hnf[m + D * j] + hqv[m];
                const double cij = o.cq_ve[q * cqq + i * cqr];
                if constexpr (NC == D + 2) {
                    if (fast_ec) {
                        double qj[NPR], gq[D], phi[NC];
#pragma unroll
                        for (int c = 0; c < NPR; c++) qj[c] = s_fprim[j + Nf * c];
#pragma unroll
                        for (int m = 0; m < D; m++) gq[m] = (0.25 * cij) * nJ[m];
                        ec_contract_scaled<D>(L, qi, qj, gq, phi);
#pragma unroll
                        for (int e = 0; e < NC; e++) r[e] -= phi[e];
                        continue;
                    }
                }
#pragma unroll
                for (int e = 0; e < NC; e++) {
                    double Fn = 0.0;
#pragma unroll
                    for (int m = 0; m < D; m++) Fn = fma(nJ[m], F[e][m], Fn);
                    r[e] -= cij * Fn;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < NC; e++) s_r[i + Nq * e] = r[e];
    }
    sse_sync();
    // r_q -= R' f_f
    SSE_FOR(t, Nq * NC) {
        int i = t % Nq, e = t / Nq;
        double s = 0.0;
        SSE_ROW_FOR(o.Rt, i, c_, v_) s = fma(v_, s_ff[c_ + Nf * e], s);
        s_r[t] -= s;
    }
    sse_sync();
    apply_Vt<NC>(o, s_r, s_m, s_z, s_w);
    mass_solve<NC>(o, g, k, s_m, s_uq, s_z, s_w);
    SSE_FOR(t, Np * NC) { dudt[(size_t)Np * NC * k + t] = s_m[t]; flag_nonfinite(g.flag, s_m[t]); }
}

// physical_flux! into a shared Nq x NC x D tile
// ---------------------------------------------------------------------------------------------------------
// Dense flux differencing for the multidimensional schemes (ModalMulti / NodalMulti, multidimensional.jl:1-75): S_m and C are
// dense, so an element evaluates N_q^2 volume pairs and 2 N_q N_f facet pairs of the entropy-conservative two-point flux
// (flux_differencing_form.jl:1-35, 91-124: the dense methods of flux_difference! / facet_correction!).  All-pairs tile, like an
// N-body kernel: thread i keeps its node (primitives, metric terms, residual) in registers and sweeps j, whose data every lane
// reads from the same shared address (broadcast, no bank conflicts, nothing handed back: each side of a pair evaluates it for
// itself, as the reference does when it visits (i, j) and (j, i)); two partners per trip so their dependent chains interleave.
// Tables are dense and slot-major ([j][m][i]: coalesced over the threads), skew-extended and pre-scaled by 1/4 (scaled pair
// flux, physics.cuh).  Euler + EC two-point flux only; everything else of pass B (interface flux, lift, V', mass solve) is the
// generic code.
template <int D>
__global__ void __launch_bounds__(128, 3)
k_time_fluxdiff_dense(Ops o, Geo g, Law L, DenseDev dd, long long first, const double* __restrict__ u_q, const double* __restrict__ u_f,
                      double* __restrict__ dudt) {
    constexpr int NC = D + 2, NPR = D + 3;
    extern __shared__ double sm[];
    const long long k = first + blockIdx.x;
    const int Nq = o.Nq, Np = o.Np, Nf = o.Nf, Nfac = o.Nfac, npf = o.npf;
    double* s_uq = sm;                       // Nq x NC  (conservative; scratch of the mass solve later)
    double* s_r = s_uq + Nq * NC;            // Nq x NC
    double* s_in = s_r + Nq * NC;            // Nf x NC
    double* s_out = s_in + Nf * NC;          // Nf x NC
    double* s_ff = s_out + Nf * NC;          // Nf x NC
    double* s_nf = s_ff + Nf * NC;           // D x Nf
    double* s_lam = s_nf + D * Nf;           // Nq x D x D
    double* s_m = s_lam + Nq * D * D;        // Np x NC
    double* s_z = s_m + Np * NC;
    double* s_w = s_z + warp_z_size(o, NC);
    double* s_prim = s_w + warp_w_size(o, NC);   // NPR x Nq
    double* s_fprim = s_prim + NPR * Nq;         // NPR x Nf
    double* s_hnf = s_fprim + NPR * Nf;          // D x Nf       halfnJf / 4 ... kept unscaled, the 1/4 sits in C4
    double* s_hq = s_hnf + D * Nf;               // Nfac x D x Nq  halfnJq (mesh.jl:262-269)

    SSE_FOR(t, Nq * NC) s_uq[t] = u_q[(size_t)Nq * NC * k + t];
    SSE_FOR(t, Nq * D * D) s_lam[t] = g.Lambda_q[(size_t)Nq * D * D * k + t];
    load_facets<D, NC>(o, g, k, u_f, s_in, s_out, s_nf);
    SSE_FOR(t, D * Nf) s_hnf[t] = 0.5 * g.nJf[(size_t)D * Nf * k + t];
    SSE_FOR(i, Nq) {
        double ui[NC], q[NPR];
#pragma unroll
        for (int e = 0; e < NC; e++) ui[e] = s_uq[i + Nq * e];
        to_prim_fast<D>(L, ui, q);
#pragma unroll
        for (int c = 0; c < NPR; c++) s_prim[i + Nq * c] = q[c];
        for (int f = 0; f < Nfac; f++)
#pragma unroll
            for (int m = 0; m < D; m++) {
                double hq;
                if (g.nJq) hq = 0.5 * g.nJq[m + D * (f + (size_t)Nfac * (i + (size_t)Nq * k))];
                else {
                    double t = 0.0;
#pragma unroll
                    for (int l = 0; l < D; l++) t += s_lam[i + Nq * (l + D * m)] * o.nref[l + D * f];
                    hq = 0.5 * t;
                }
                s_hq[i + Nq * (m + D * f)] = hq;
            }
    }
    SSE_FOR(j, Nf) {
        double ui[NC], q[NPR];
#pragma unroll
        for (int e = 0; e < NC; e++) ui[e] = s_in[j + Nf * e];
        to_prim_fast<D>(L, ui, q);
#pragma unroll
        for (int c = 0; c < NPR; c++) s_fprim[j + Nf * c] = q[c];
    }
    sse_sync();

    // ---- volume side: r_i = - sum_j F(u_i, u_j) . g_ij,  g_ij = sum_m S_m[i,j] (Lambda_i + Lambda_j)[m, :]
    //                        - sum_j F(u_i, u_fj) . C_ij (halfnJf_j + halfnJq_i,f(j))
    SSE_FOR(i, Nq) {
        double qi[NPR], li[D][D], r[NC];
#pragma unroll
        for (int c = 0; c < NPR; c++) qi[c] = s_prim[i + Nq * c];
#pragma unroll
        for (int m = 0; m < D; m++)
#pragma unroll
            for (int n = 0; n < D; n++) li[m][n] = s_lam[i + Nq * (m + D * n)];
#pragma unroll
        for (int e = 0; e < NC; e++) r[e] = 0.0;
        // the S table (N_q^2 d doubles) does not stay in L1 next to the tiles of six resident CTAs: the weights of the NEXT trip are
        // fetched (L2) while the current one computes
        double sA[D], sB[D];
#pragma unroll
        for (int m = 0; m < D; m++) { sA[m] = dd.S4[(size_t)m * Nq + i]; sB[m] = dd.S4[((size_t)(Nq > 1 ? 1 : 0) * D + m) * Nq + i]; }
        for (int j0 = 0; j0 < Nq; j0 += 2) {
            const int jA = j0, jB = (j0 + 1 < Nq) ? j0 + 1 : j0;       // odd N_q: the last trip repeats a partner with weight 0
            const double wB = (j0 + 1 < Nq) ? 1.0 : 0.0;
            const int nA = (j0 + 2 < Nq) ? j0 + 2 : jA, nB = (j0 + 3 < Nq) ? j0 + 3 : jB;
            double tA[D], tB[D];
#pragma unroll
            for (int m = 0; m < D; m++) { tA[m] = dd.S4[((size_t)nA * D + m) * Nq + i]; tB[m] = dd.S4[((size_t)nB * D + m) * Nq + i]; }
            double qA[NPR], qB[NPR], gA[D], gB[D], pA[NC], pB[NC];
#pragma unroll
            for (int c = 0; c < NPR; c++) { qA[c] = s_prim[jA + Nq * c]; qB[c] = s_prim[jB + Nq * c]; }
#pragma unroll
            for (int n = 0; n < D; n++) { gA[n] = 0.0; gB[n] = 0.0; }
#pragma unroll
            for (int m = 0; m < D; m++) {
                const double wsB = wB * sB[m];
#pragma unroll
                for (int n = 0; n < D; n++) {
                    gA[n] = fma(sA[m], li[m][n] + s_lam[jA + Nq * (m + D * n)], gA[n]);
                    gB[n] = fma(wsB, li[m][n] + s_lam[jB + Nq * (m + D * n)], gB[n]);
                }
            }
            ec_contract_scaled2<D>(L, qi, qA, qB, gA, gB, pA, pB);
#pragma unroll
            for (int e = 0; e < NC; e++) r[e] -= pA[e] + pB[e];
#pragma unroll
            for (int m = 0; m < D; m++) { sA[m] = tA[m]; sB[m] = tB[m]; }
        }
        if (o.has_C) {
            for (int f = 0; f < Nfac; f++) {
                double hq[D];
#pragma unroll
                for (int n = 0; n < D; n++) hq[n] = s_hq[i + Nq * (n + D * f)];
                double cnA = dd.C4[(size_t)(f * npf) * Nq + i], cnB = dd.C4[(size_t)(f * npf + (npf > 1 ? 1 : 0)) * Nq + i];
                for (int j0 = f * npf; j0 < (f + 1) * npf; j0 += 2) {
                    const int jA = j0, jB = (j0 + 1 < (f + 1) * npf) ? j0 + 1 : j0;
                    const double cA = cnA, cB = (jB != jA) ? cnB : 0.0;
                    cnA = dd.C4[(size_t)((j0 + 2 < (f + 1) * npf) ? j0 + 2 : jA) * Nq + i];      // next trip's weights
                    cnB = dd.C4[(size_t)((j0 + 3 < (f + 1) * npf) ? j0 + 3 : jB) * Nq + i];
                    double qA[NPR], qB[NPR], gA[D], gB[D], pA[NC], pB[NC];
#pragma unroll
                    for (int c = 0; c < NPR; c++) { qA[c] = s_fprim[jA + Nf * c]; qB[c] = s_fprim[jB + Nf * c]; }
#pragma unroll
                    for (int n = 0; n < D; n++) { gA[n] = cA * (s_hnf[n + D * jA] + hq[n]); gB[n] = cB * (s_hnf[n + D * jB] + hq[n]); }
                    ec_contract_scaled2<D>(L, qi, qA, qB, gA, gB, pA, pB);
#pragma unroll
                    for (int e = 0; e < NC; e++) r[e] -= pA[e] + pB[e];
                }
            }
        }
#pragma unroll
        for (int e = 0; e < NC; e++) s_r[i + Nq * e] = r[e];
    }
    // ---- facet side: f_f[j] = BJf f*(u-, u+) - sum_i F(u_i, u_fj) . C_ij (halfnJf_j + halfnJq_i,f(j))
    SSE_FOR(j, Nf) {
        double ui[NC], uo[NC], fs[NC], qj[NPR], hf[D];
#pragma unroll
        for (int e = 0; e < NC; e++) { ui[e] = s_in[j + Nf * e]; uo[e] = s_out[j + Nf * e]; }
        numerical_flux<D, NC>(L, L.two_point, ui, uo, s_nf + D * j, fs);
        const double bj = o.Bf[j] * g.J_f[(size_t)Nf * k + j];
#pragma unroll
        for (int e = 0; e < NC; e++) fs[e] *= bj;
        if (o.has_C) {
            const int f = j / npf;
#pragma unroll
            for (int c = 0; c < NPR; c++) qj[c] = s_fprim[j + Nf * c];
#pragma unroll
            for (int n = 0; n < D; n++) hf[n] = s_hnf[n + D * j];
            double cnA = dd.C4[(size_t)j * Nq], cnB = dd.C4[(size_t)j * Nq + (Nq > 1 ? 1 : 0)];
            for (int i0 = 0; i0 < Nq; i0 += 2) {
                const int iA = i0, iB = (i0 + 1 < Nq) ? i0 + 1 : i0;
                const double cA = cnA, cB = (iB != iA) ? cnB : 0.0;
                cnA = dd.C4[(size_t)j * Nq + ((i0 + 2 < Nq) ? i0 + 2 : iA)];                    // next trip's weights
                cnB = dd.C4[(size_t)j * Nq + ((i0 + 3 < Nq) ? i0 + 3 : iB)];
                double qA[NPR], qB[NPR], gA[D], gB[D], pA[NC], pB[NC];
#pragma unroll
                for (int c = 0; c < NPR; c++) { qA[c] = s_prim[iA + Nq * c]; qB[c] = s_prim[iB + Nq * c]; }
#pragma unroll
                for (int n = 0; n < D; n++) { gA[n] = cA * (hf[n] + s_hq[iA + Nq * (n + D * f)]); gB[n] = cB * (hf[n] + s_hq[iB + Nq * (n + D * f)]); }
                ec_contract_scaled2<D>(L, qj, qA, qB, gA, gB, pA, pB);       // the flux is symmetric in its two states
#pragma unroll
                for (int e = 0; e < NC; e++) fs[e] -= pA[e] + pB[e];
            }
        }
#pragma unroll
        for (int e = 0; e < NC; e++) s_ff[j + Nf * e] = fs[e];
    }
    sse_sync();
    // ---- r_q -= R' f_f as a dense product (no index indirection, no data-dependent loop exit: the loads pipeline);
    //      dudt = M^-1 V' r_q         flux_differencing_form.jl:341-346
    SSE_FOR(i, Nq) {
        double acc[NC];
#pragma unroll
        for (int e = 0; e < NC; e++) acc[e] = 0.0;
#pragma unroll 4
        for (int j = 0; j < Nf; j++) {
            const double v = dd.RT[(size_t)j * Nq + i];
#pragma unroll
            for (int e = 0; e < NC; e++) acc[e] = fma(v, s_ff[j + Nf * e], acc[e]);
        }
#pragma unroll
        for (int e = 0; e < NC; e++) s_r[i + Nq * e] -= acc[e];
    }
    sse_sync();
    apply_Vt<NC>(o, s_r, s_m, s_z, s_w);
    mass_solve<NC>(o, g, k, s_m, s_uq, s_z, s_w);
    SSE_FOR(t, Np * NC) { dudt[(size_t)Np * NC * k + t] = s_m[t]; flag_nonfinite(g.flag, s_m[t]); }
}

template <int D, int NC>
__device__ void physical_flux_tile(const Ops& o, const Law& L, const double* s_uq, const double* s_qq, double* s_fq) {
    const int Nq = o.Nq;
    SSE_FOR(i, Nq) {
        if (L.pde == SSE_PDE_EULER) {
            if constexpr (NC == D + 2) {
                double ui[NC], F[NC][D];
#pragma unroll
                for (int e = 0; e < NC; e++) ui[e] = s_uq[i + Nq * e];
                euler_physical_flux<D>(L, ui, F);
#pragma unroll
                for (int e = 0; e < NC; e++)
#pragma unroll
                    for (int m = 0; m < D; m++) s_fq[i + Nq * (e + NC * m)] = F[e][m];
            }
        } else {
#pragma unroll
            for (int m = 0; m < D; m++) {
                double f = L.a[m] * s_uq[i];
                if (L.pde == SSE_PDE_BURGERS) f = 0.5 * L.a[m] * s_uq[i] * s_uq[i];        // burgers.jl:52-58
                if (L.viscous) f -= L.b * s_qq[i + Nq * NC * m];     // linear_advection_diffusion.jl:64-71, burgers.jl:60-70
                s_fq[i + Nq * NC * m] = f;
            }
        }
    }
    sse_sync();
}

// ------------------------------------------------------------------ pass B: StandardForm + ReferenceOperators
// standard_form_first_order.jl:16-63
template <int D, int NC>
__global__ void k_time_standard_reference(Ops o, Geo g, Law L, long long first, const double* __restrict__ u_q,
                                          const double* __restrict__ u_f, double* __restrict__ dudt) {
    extern __shared__ double sm_cta[];
    double* sm = sse_row_smem(sm_cta);
    const long long k = sse_element(first);
    const int Nq = o.Nq, Np = o.Np, Nf = o.Nf;
    double* s_uq = sm;                        // Nq x NC (later scratch)
    double* s_r = s_uq + Nq * NC;             // Nq x NC
    double* s_fq = s_r + Nq * NC;             // Nq x NC x D
    double* s_in = s_fq + Nq * NC * D;        // Nf x NC
    double* s_out = s_in + Nf * NC;
    double* s_ff = s_out + Nf * NC;
    double* s_nf = s_ff + Nf * NC;            // D x Nf
    double* s_lam = s_nf + D * Nf;            // Nq x D x D, premultiplied by 0.5 W (halfWΛ, operators.jl:16-18)
    double* s_m = s_lam + Nq * D * D;         // Np x NC
    double* s_z = s_m + Np * NC;
    double* s_w = s_z + warp_z_size(o, NC);

    SSE_FOR(t, Nq * NC) s_uq[t] = u_q[(size_t)Nq * NC * k + t];
    SSE_FOR(t, Nq * D * D) { int i = t % Nq; s_lam[t] = (0.5 * o.W[i]) * g.Lambda_q[(size_t)Nq * D * D * k + t]; }
    load_facets<D, NC>(o, g, k, u_f, s_in, s_out, s_nf);
    physical_flux_tile<D, NC>(o, L, s_uq, nullptr, s_fq);
    // volume terms: r = sum_{m,n} D_m' (hWL_mn f_n) - hWL_mn (D_m f_n)
    SSE_FOR(t, Nq * NC) {
        int i = t % Nq, e = t / Nq;
        double r = 0.0;
#pragma unroll
        for (int n = 0; n < D; n++) {
            const double* fn = s_fq + Nq * (e + NC * n);
#pragma unroll
            for (int m = 0; m < D; m++) {
                const double* hl = s_lam + Nq * (m + D * n);
                double s1 = 0.0, s2 = 0.0;
                SSE_ROW_FOR(o.Dt[m], i, j, v_) s1 = fma(v_, hl[j] * fn[j], s1);
                SSE_ROW_FOR(o.D[m], i, c_, v_) s2 = fma(v_, fn[c_], s2);
                r += s1;
                r -= hl[i] * s2;
            }
        }
        s_r[t] = r;
    }
    // facet terms: f_f = BJf (f* - sum_n halfN_n R f_n)
    SSE_FOR(j, Nf) {
        double ui[NC], uo[NC], fs[NC];
#pragma unroll
        for (int e = 0; e < NC; e++) { ui[e] = s_in[j + Nf * e]; uo[e] = s_out[j + Nf * e]; }
        numerical_flux<D, NC>(L, SSE_TWO_POINT_CONSERVATIVE, ui, uo, s_nf + D * j, fs);
#pragma unroll
        for (int n = 0; n < D; n++) {
            const double hn = 0.5 * s_nf[n + D * j];
#pragma unroll
            for (int e = 0; e < NC; e++) {
                double s = 0.0;
                SSE_ROW_FOR(o.R, j, c_, v_) s = fma(v_, s_fq[c_ + Nq * (e + NC * n)], s);
                fs[e] -= hn * s;
            }
        }
        double bj = o.Bf[j] * g.J_f[(size_t)Nf * k + j];
#pragma unroll
        for (int e = 0; e < NC; e++) s_ff[j + Nf * e] = bj * fs[e];
    }
    sse_sync();
    SSE_FOR(t, Nq * NC) {
        int i = t % Nq, e = t / Nq;
        double s = 0.0;
        SSE_ROW_FOR(o.Rt, i, c_, v_) s = fma(v_, s_ff[c_ + Nf * e], s);
        s_r[t] -= s;
    }
    sse_sync();
    apply_Vt<NC>(o, s_r, s_m, s_z, s_w);
    mass_solve<NC>(o, g, k, s_m, s_uq, s_z, s_w);
    SSE_FOR(t, Np * NC) { dudt[(size_t)Np * NC * k + t] = s_m[t]; flag_nonfinite(g.flag, s_m[t]); }
}

// ------------------------------------------------------------------ PhysicalOperators paths
// BR1 auxiliary variable: standard_form_second_order.jl:3-34, linear_advection_diffusion.jl:75-86
// resident CTAs per SM requested for the scalar 2-D instantiation (BASELINE config 3: rows of one warp, 4 per CTA): without
// bounds ptxas assumes 1 024 threads per CTA, i.e. 64 registers = 8 CTAs of 128 threads per SM; the kernels are bound by the
// chain of global round trips of a warp (profiles/r2_s5_ncu_config3.csv), so resident warps are what hides it
// Measured at 131 072 triangles (profiles/r2_s5_config3_bounds_ab.log), aux / pass B: unbounded (96 / 71 registers) 0.675 / 0.414 ms,
// 10 CTAs (48 registers) 0.512 / 0.328, 12 CTAs (40) 0.458 / 0.302, 16 CTAs (32) 0.511 / 0.280.  The other instantiations are held
// at 64 registers (1 024-thread bound), which is what they used before the two-partial-sum form of the wide rows.
#ifndef SSE_PHYS_MINB
#define SSE_PHYS_MINB 12
#endif
#define SSE_PHYS_BOUNDS __launch_bounds__((D == 2 && NC == 1 && SSE_PHYS_MINB > 1) ? 128 : 1024, (D == 2 && NC == 1) ? SSE_PHYS_MINB : 1)
template <int D, int NC>
__global__ void SSE_PHYS_BOUNDS k_aux_physical(Ops o, Geo g, Law L, long long first, const double* __restrict__ u_q,
                               const double* __restrict__ u_f, double* __restrict__ q_q, double* __restrict__ q_f) {
    extern __shared__ double sm_cta[];
    double* sm = sse_row_smem(sm_cta);
    const long long k = sse_element(first);
    const int Nq = o.Nq, Np = o.Np, Nf = o.Nf;
    double* s_uq = sm;                        // Nq x NC
    double* s_in = s_uq + Nq * NC;            // Nf x NC
    double* s_out = s_in + Nf * NC;
    double* s_nf = s_out + Nf * NC;           // D x Nf
    double* s_m = s_nf + D * Nf;              // Np x NC
    double* s_qq = s_m + Np * NC;             // Nq x NC
    double* s_z = s_qq + Nq * NC;
    double* s_w = s_z + warp_z_size(o, NC);
    SSE_FOR(t, Nq * NC) s_uq[t] = u_q[(size_t)Nq * NC * k + t];
    load_facets<D, NC>(o, g, k, u_f, s_in, s_out, s_nf);
    const double* FAC = g.FAC + (size_t)Np * Nf * k;
    for (int m = 0; m < D; m++) {
        const double* VOL = g.VOL + (size_t)Np * Nq * (m + (size_t)D * k);
        if (blockDim.x == 32 && Np * NC <= 16) {
            // small elements, one warp per element: the per-element operators stream from DRAM, so every dot product is split
            // over the two half-warps (all 32 lanes load, twice as many independent loads in flight) and joined by a shuffle
            const int t = threadIdx.x & 15, h = threadIdx.x >> 4;
            const int tc = t < Np * NC ? t : 0, a = tc % Np, e = tc / Np;
            const int q0 = h ? (Nq + 1) / 2 : 0, q1 = h ? Nq : (Nq + 1) / 2, f0 = h ? (Nf + 1) / 2 : 0, f1 = h ? Nf : (Nf + 1) / 2;
            double s = 0.0, s2 = 0.0;
#pragma unroll 4
            for (int q = q0; q < q1; q++) s = fma(__ldcs(VOL + a + (size_t)Np * q), s_uq[q + Nq * e], s);
#pragma unroll 4
            for (int f = f0; f < f1; f++) {
                double un = 0.5 * (s_in[f + Nf * e] + s_out[f + Nf * e]) * s_nf[m + D * f];
                s2 = fma(__ldcs(FAC + a + (size_t)Np * f), un, s2);
            }
            double r = -s - s2;
            r += __shfl_xor_sync(0xffffffffu, r, 16);
            if (h == 0 && t < Np * NC) s_m[t] = r;
        } else {
        SSE_FOR(t, Np * NC) {                      // the same two partial sums per entry, so the bits do not depend on the launch shape
            int a = t % Np, e = t / Np;
            double r[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int q0 = h ? (Nq + 1) / 2 : 0, q1 = h ? Nq : (Nq + 1) / 2, f0 = h ? (Nf + 1) / 2 : 0, f1 = h ? Nf : (Nf + 1) / 2;
                double s = 0.0, s2 = 0.0;
#pragma unroll 4
                for (int q = q0; q < q1; q++) s = fma(VOL[a + (size_t)Np * q], s_uq[q + Nq * e], s);
#pragma unroll 4
                for (int f = f0; f < f1; f++) {
                    double un = 0.5 * (s_in[f + Nf * e] + s_out[f + Nf * e]) * s_nf[m + D * f];
                    s2 = fma(FAC[a + (size_t)Np * f], un, s2);
                }
                r[h] = -s - s2;
            }
            s_m[t] = r[0] + r[1];
        }
        }
        sse_sync();
        apply_V<NC>(o, s_m, s_qq, s_z, s_w);
        SSE_FOR(t, Nq * NC) q_q[(size_t)Nq * NC * (m + (size_t)D * k) + t] = s_qq[t];
        SSE_FOR(t, Nf * NC) {
            int j = t % Nf, e = t / Nf;
            double s = 0.0;
            SSE_ROW_FOR(o.R, j, c_, v_) s = fma(v_, s_qq[c_ + Nq * e], s);
            q_f[(size_t)Nf * k + j + (size_t)g.NFT * (e + NC * m)] = s;
        }
        sse_sync();
    }
}

// first-order (standard_form_first_order.jl:65-94) and second-order (standard_form_second_order.jl:38-75)
// time derivative with per-element VOL / FAC; viscous flux linear_advection_diffusion.jl:90-102
template <int D, int NC>
__global__ void SSE_PHYS_BOUNDS k_time_physical(Ops o, Geo g, Law L, long long first, int second_order, const double* __restrict__ u_q,
                                const double* __restrict__ u_f, const double* __restrict__ q_q,
                                const double* __restrict__ q_f, double* __restrict__ dudt) {
    extern __shared__ double sm_cta[];
    double* sm = sse_row_smem(sm_cta);
    const long long k = sse_element(first);
    const int Nq = o.Nq, Np = o.Np, Nf = o.Nf;
    double* s_uq = sm;                        // Nq x NC
    double* s_qq = s_uq + Nq * NC;            // Nq x NC x D
    double* s_fq = s_qq + Nq * NC * D;        // Nq x NC x D
    double* s_in = s_fq + Nq * NC * D;
    double* s_out = s_in + Nf * NC;
    double* s_ff = s_out + Nf * NC;
    double* s_nf = s_ff + Nf * NC;
    SSE_FOR(t, Nq * NC) s_uq[t] = u_q[(size_t)Nq * NC * k + t];
    if (second_order) SSE_FOR(t, Nq * NC * D) s_qq[t] = q_q[(size_t)Nq * NC * D * k + t];
    load_facets<D, NC>(o, g, k, u_f, s_in, s_out, s_nf);
    physical_flux_tile<D, NC>(o, L, s_uq, s_qq, s_fq);
    SSE_FOR(j, Nf) {
        double ui[NC], uo[NC], fs[NC];
#pragma unroll
        for (int e = 0; e < NC; e++) { ui[e] = s_in[j + Nf * e]; uo[e] = s_out[j + Nf * e]; }
        numerical_flux<D, NC>(L, SSE_TWO_POINT_CONSERVATIVE, ui, uo, s_nf + D * j, fs);
        if (second_order) {
            const size_t jo = (size_t)(g.mapP[(size_t)Nf * k + j] - 1);
#pragma unroll
            for (int e = 0; e < NC; e++) {
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < D; m++) {
                    double qi = q_f[(size_t)Nf * k + j + (size_t)g.NFT * (e + NC * m)];
                    double qo = q_f[jo + (size_t)g.NFT * (e + NC * m)];
                    acc += L.b * (-0.5 * (qi + qo)) * s_nf[m + D * j];
                }
                fs[e] += acc;
            }
        }
#pragma unroll
        for (int e = 0; e < NC; e++) s_ff[j + Nf * e] = fs[e];
    }
    sse_sync();
    const double* FAC = g.FAC + (size_t)Np * Nf * k;
    if (blockDim.x == 32 && Np * NC <= 16) {             // as in k_aux_physical: half-warp partial sums, one shuffle
        const int t = threadIdx.x & 15, h = threadIdx.x >> 4;
        const int tc = t < Np * NC ? t : 0, a = tc % Np, e = tc / Np;
        const int q0 = h ? (Nq + 1) / 2 : 0, q1 = h ? Nq : (Nq + 1) / 2, f0 = h ? (Nf + 1) / 2 : 0, f1 = h ? Nf : (Nf + 1) / 2;
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < D; m++) {
            const double* VOL = g.VOL + (size_t)Np * Nq * (m + (size_t)D * k);
            double sv = 0.0;
#pragma unroll 4
            for (int q = q0; q < q1; q++) sv = fma(__ldcs(VOL + a + (size_t)Np * q), s_fq[q + Nq * (e + NC * m)], sv);
            s += sv;
        }
        double sf = 0.0;
#pragma unroll 4
        for (int f = f0; f < f1; f++) sf = fma(__ldcs(FAC + a + (size_t)Np * f), s_ff[f + Nf * e], sf);
        double r = s + sf;
        r += __shfl_xor_sync(0xffffffffu, r, 16);
        if (h == 0 && t < Np * NC) {
            dudt[(size_t)Np * NC * k + t] = r;
            flag_nonfinite(g.flag, r);
        }
    } else {
    SSE_FOR(t, Np * NC) {                          // the same two partial sums per entry as the half-warp form above
        int a = t % Np, e = t / Np;
        double r[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int q0 = h ? (Nq + 1) / 2 : 0, q1 = h ? Nq : (Nq + 1) / 2, f0 = h ? (Nf + 1) / 2 : 0, f1 = h ? Nf : (Nf + 1) / 2;
            double s = 0.0;
#pragma unroll
            for (int m = 0; m < D; m++) {
                const double* VOL = g.VOL + (size_t)Np * Nq * (m + (size_t)D * k);
                double sv = 0.0;
#pragma unroll 4
                for (int q = q0; q < q1; q++) sv = fma(VOL[a + (size_t)Np * q], s_fq[q + Nq * (e + NC * m)], sv);
                s += sv;
            }
            double sf = 0.0;
#pragma unroll 4
            for (int f = f0; f < f1; f++) sf = fma(FAC[a + (size_t)Np * f], s_ff[f + Nf * e], sf);
            r[h] = s + sf;
        }
        const double v = r[0] + r[1];
        dudt[(size_t)Np * NC * k + t] = v;
        flag_nonfinite(g.flag, v);
    }
    }
}

// ------------------------------------------------------------------ small streaming kernels
static __global__ void k_axpby(long long n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = (b == 0.0) ? a * x[i] : fma(a, x[i], b * y[i]);
}
static __global__ void k_fill(long long n, double v, double* __restrict__ y) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = v;
}
// 2N-storage RK stage (Carpenter & Kennedy 1994): tmp = A tmp + dt dudt ; u += B tmp
static __global__ void k_lsrk_stage(long long n, double* __restrict__ u, double* __restrict__ tmp, const double* __restrict__ dudt,
                             double A, double B, double dt) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double t = fma(A, tmp[i], dt * dudt[i]);
        tmp[i] = t;
        u[i] = fma(B, t, u[i]);
    }
}
// halo pack / unpack: buffers are variable-fastest [slot][var]
static __global__ void k_halo_pack(long long n_send, int nvar, long long NFT, const long long* __restrict__ send_idx,
                            const double* __restrict__ facet, double* __restrict__ buf) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_send * nvar; t += (long long)gridDim.x * blockDim.x) {
        long long s = t / nvar; int e = (int)(t % nvar);
        buf[t] = facet[(send_idx[s] - 1) + NFT * e];
    }
}
static __global__ void k_halo_unpack(long long n_ghost, int nvar, long long NFT, long long owned, const double* __restrict__ buf,
                              double* __restrict__ facet) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_ghost * nvar; t += (long long)gridDim.x * blockDim.x) {
        long long s = t / nvar; int e = (int)(t % nvar);
        facet[owned + s + NFT * e] = buf[t];
    }
}

// register-resident DFMA peak (8 independent chains per thread)
static __global__ void k_fp64_peak(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace sse
