/*
 * sse_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference algorithm for the hot path
 * `semi_discrete_residual!` of StableSpectralElements.jl v0.2.13, written loop by
 * loop after the Julia sources (every function cites the file:line it follows,
 * relative to /root/reference).  It takes exactly the inputs of the C ABI
 * (include/sse_b200.h: sse_config + sse_arrays, reference layouts) and is used
 *   - by tests/ as the parity checker for the CUDA library,
 *   - by __graft_entry__.smoke() as the checker,
 *   - by bench.py's cpu_baseline / `--impl reference` leg (OpenMP over elements,
 *     like the reference's Threads.@threads, Solvers.jl:495-514).
 * Nothing under cloud.jl_b200/ may import, link or call it.
 *
 * Pinning (see oracle/README.md): the reference cannot run in this image (no
 * Julia).  The oracle reproduces the reference's own golden vectors of
 * test/runtests.jl to round-off: 1-D (:14-36 advection-diffusion BR1, :89-96 Euler
 * Gauss collocation), 2-D (:38-60 ModalTensor Tri advection, :62-80 NodalTensor
 * Quad flux differencing, :111-121 ModalTensor Tri Euler vortex: flux differencing,
 * entropy projection, facet correction, weight-adjusted mass solve) and 3-D
 * (:131-144 NodalTensor Hex Euler), plus the reference's invariant assertions
 * (conservation / energy / entropy, runtests.jl:35-142); see tests/test_oracle_*.py.
 * Not reproducible here: :98-109 (tabulated SBP nodes) and :123-129 (tetrahedral
 * mesh split, 3-D warp-and-blend nodes and Jaskowiec-Sukumar error quadrature of the
 * un-vendored StartUpDG/NodesAndModes): for the tetrahedral golden parity is unpinned.
 *
 * Third-party arithmetic restated here: LinearMaps.jl "3" (Kronecker/Block/
 * Transpose mul!, column-by-column application) and Octavian.jl "0.3"
 * (matmul_serial!) are plain linear algebra; they are applied as sparse
 * matrix-vector products over the non-zeros of Matrix(map), which for Kronecker
 * products with identity factors is the same operation count as the reference's
 * sum-factorised application.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/sse_b200.h"

#define MAXD 3
#define MAXC 5

/* ---------------------------------------------------------------- sparse helpers */
typedef struct {
    int nrow, ncol;
    int *ptr, *idx;
    double *val;
} sp_t; /* CSR (ptr over rows) or CSC (ptr over cols), by use */

static sp_t dense_to_csr(const double *A, int nrow, int ncol) { /* A column-major */
    sp_t s = {nrow, ncol, calloc(nrow + 1, sizeof(int)), NULL, NULL};
    int nnz = 0;
    for (int i = 0; i < nrow; i++)
        for (int j = 0; j < ncol; j++)
            if (A[i + (size_t)nrow * j] != 0.0) nnz++;
    s.idx = malloc(sizeof(int) * (nnz ? nnz : 1));
    s.val = malloc(sizeof(double) * (nnz ? nnz : 1));
    nnz = 0;
    for (int i = 0; i < nrow; i++) {
        s.ptr[i] = nnz;
        for (int j = 0; j < ncol; j++)
            if (A[i + (size_t)nrow * j] != 0.0) { s.idx[nnz] = j; s.val[nnz++] = A[i + (size_t)nrow * j]; }
    }
    s.ptr[nrow] = nnz;
    return s;
}
/* Julia's sparse(Matrix): CSC keeping every entry != 0 */
static sp_t dense_to_csc(const double *A, int nrow, int ncol) {
    sp_t s = {nrow, ncol, calloc(ncol + 1, sizeof(int)), NULL, NULL};
    int nnz = 0;
    for (size_t t = 0; t < (size_t)nrow * ncol; t++) if (A[t] != 0.0) nnz++;
    s.idx = malloc(sizeof(int) * (nnz ? nnz : 1));
    s.val = malloc(sizeof(double) * (nnz ? nnz : 1));
    nnz = 0;
    for (int j = 0; j < ncol; j++) {
        s.ptr[j] = nnz;
        for (int i = 0; i < nrow; i++)
            if (A[i + (size_t)nrow * j] != 0.0) { s.idx[nnz] = i; s.val[nnz++] = A[i + (size_t)nrow * j]; }
    }
    s.ptr[ncol] = nnz;
    return s;
}
static void sp_free(sp_t *s) { free(s->ptr); free(s->idx); free(s->val); memset(s, 0, sizeof(*s)); }

/* y(nrow x nc) = A x(ncol x nc), column by column (LinearMaps matrix fallback) */
static void csr_mul(const sp_t *A, const double *x, int ldx, double *y, int ldy, int nc) {
    for (int e = 0; e < nc; e++)
        for (int i = 0; i < A->nrow; i++) {
            double t = 0.0;
            for (int q = A->ptr[i]; q < A->ptr[i + 1]; q++) t = fma(A->val[q], x[A->idx[q] + (size_t)ldx * e], t);
            y[i + (size_t)ldy * e] = t;
        }
}
/* y(ncol x nc) = A' x(nrow x nc) */
static void csr_mul_t(const sp_t *A, const double *x, int ldx, double *y, int ldy, int nc) {
    for (int e = 0; e < nc; e++) {
        for (int j = 0; j < A->ncol; j++) y[j + (size_t)ldy * e] = 0.0;
        for (int i = 0; i < A->nrow; i++) {
            double xi = x[i + (size_t)ldx * e];
            for (int q = A->ptr[i]; q < A->ptr[i + 1]; q++) y[A->idx[q] + (size_t)ldy * e] = fma(A->val[q], xi, y[A->idx[q] + (size_t)ldy * e]);
        }
    }
}

/* ---------------------------------------------------------------- the solver image */
typedef struct {
    sse_config c;
    const sse_arrays *a;
    int P1;
    sp_t Vcsr, Rcsr, Dcsr[MAXD], Scsc[MAXD], Ccsc;
    int has_C;
    int64_t *sig_i, *sig_o; /* 0-based copies */
    int N2[8], N3[8][8];
    double *chol;           /* CholeskySolver: upper factors U_k (N_p x N_p column-major per element), mass_matrix.jl:30-39 */
} ora_t;

/* ---------------------------------------------------------------- V, V' */
/* warped_product_2d.jl:31-55 */
static void warped2d_mul(const ora_t *o, const double *x, double *y) {
    const int P1 = o->P1, M1 = o->c.M1d[0], M2 = o->c.M1d[1];
    const double *A = o->a->A, *B = o->a->B;
    double Z[8][8];
    for (int a2 = 0; a2 < M2; a2++)
        for (int b1 = 0; b1 < P1; b1++) {
            double t = 0.0;
            for (int b2 = 0; b2 < o->N2[b1]; b2++)
                t = fma(B[a2 + M2 * (b1 + P1 * b2)], x[o->sig_i[b1 + P1 * b2]], t);
            Z[b1][a2] = t;
        }
    for (int a1 = 0; a1 < M1; a1++)
        for (int a2 = 0; a2 < M2; a2++) {
            double t = 0.0;
            for (int b1 = 0; b1 < P1; b1++) t = fma(A[a1 + M1 * b1], Z[b1][a2], t);
            y[o->sig_o[a1 + M1 * a2]] = t;
        }
}
/* warped_product_2d.jl:61-89 */
static void warped2d_mul_t(const ora_t *o, const double *x, double *y) {
    const int P1 = o->P1, M1 = o->c.M1d[0], M2 = o->c.M1d[1];
    const double *A = o->a->A, *B = o->a->B;
    double Z[8][8];
    for (int b1 = 0; b1 < P1; b1++)
        for (int a2 = 0; a2 < M2; a2++) {
            double t = 0.0;
            for (int a1 = 0; a1 < M1; a1++) t = fma(A[a1 + M1 * b1], x[o->sig_o[a1 + M1 * a2]], t);
            Z[b1][a2] = t;
        }
    for (int b1 = 0; b1 < P1; b1++)
        for (int b2 = 0; b2 < o->N2[b1]; b2++) {
            double t = 0.0;
            for (int a2 = 0; a2 < M2; a2++) t = fma(B[a2 + M2 * (b1 + P1 * b2)], Z[b1][a2], t);
            y[o->sig_i[b1 + P1 * b2]] = t;
        }
}
/* warped_product_3d.jl:47-84 */
static void warped3d_mul(const ora_t *o, const double *x, double *y) {
    const int P1 = o->P1, M1 = o->c.M1d[0], M2 = o->c.M1d[1], M3 = o->c.M1d[2];
    const double *A = o->a->A, *B = o->a->B, *C = o->a->C;
    double Z[8][8][8], Wt[8][8][8];
    for (int b1 = 0; b1 < P1; b1++)
        for (int b2 = 0; b2 < o->N2[b1]; b2++)
            for (int a3 = 0; a3 < M3; a3++) {
                double t = 0.0;
                for (int b3 = 0; b3 < o->N3[b1][b2]; b3++)
                    t = fma(C[a3 + M3 * (b1 + P1 * (b2 + P1 * b3))], x[o->sig_i[b1 + P1 * (b2 + P1 * b3)]], t);
                Z[b1][b2][a3] = t;
            }
    for (int b1 = 0; b1 < P1; b1++)
        for (int a2 = 0; a2 < M2; a2++)
            for (int a3 = 0; a3 < M3; a3++) {
                double t = 0.0;
                for (int b2 = 0; b2 < o->N2[b1]; b2++) t = fma(B[a2 + M2 * (b1 + P1 * b2)], Z[b1][b2][a3], t);
                Wt[b1][a2][a3] = t;
            }
    for (int a1 = 0; a1 < M1; a1++)
        for (int a2 = 0; a2 < M2; a2++)
            for (int a3 = 0; a3 < M3; a3++) {
                double t = 0.0;
                for (int b1 = 0; b1 < P1; b1++) t = fma(A[a1 + M1 * b1], Wt[b1][a2][a3], t);
                y[o->sig_o[a1 + M1 * (a2 + M2 * a3)]] = t;
            }
}
/* warped_product_3d.jl:94-136 */
static void warped3d_mul_t(const ora_t *o, const double *x, double *y) {
    const int P1 = o->P1, M1 = o->c.M1d[0], M2 = o->c.M1d[1], M3 = o->c.M1d[2];
    const double *A = o->a->A, *B = o->a->B, *C = o->a->C;
    double Z[8][8][8], Wt[8][8][8];
    for (int b1 = 0; b1 < P1; b1++)
        for (int a2 = 0; a2 < M2; a2++)
            for (int a3 = 0; a3 < M3; a3++) {
                double t = 0.0;
                for (int a1 = 0; a1 < M1; a1++) t = fma(A[a1 + M1 * b1], x[o->sig_o[a1 + M1 * (a2 + M2 * a3)]], t);
                Wt[b1][a2][a3] = t;
            }
    for (int b1 = 0; b1 < P1; b1++)
        for (int b2 = 0; b2 < o->N2[b1]; b2++)
            for (int a3 = 0; a3 < M3; a3++) {
                double t = 0.0;
                for (int a2 = 0; a2 < M2; a2++) t = fma(B[a2 + M2 * (b1 + P1 * b2)], Wt[b1][a2][a3], t);
                Z[b1][b2][a3] = t;
            }
    for (int b1 = 0; b1 < P1; b1++)
        for (int b2 = 0; b2 < o->N2[b1]; b2++)
            for (int b3 = 0; b3 < o->N3[b1][b2]; b3++) {
                double t = 0.0;
                for (int a3 = 0; a3 < M3; a3++) t = fma(C[a3 + M3 * (b1 + P1 * (b2 + P1 * b3))], Z[b1][b2][a3], t);
                y[o->sig_i[b1 + P1 * (b2 + P1 * b3)]] = t;
            }
}

/* y (N_q x nc) = V x (N_p x nc) */
static void V_mul(const ora_t *o, const double *x, double *y, int nc) {
    const int Nq = o->c.N_q, Np = o->c.N_p;
    if (o->c.v_kind == SSE_V_IDENTITY) { memcpy(y, x, sizeof(double) * Nq * nc); return; }
    if (o->c.v_kind == SSE_V_DENSE) { csr_mul(&o->Vcsr, x, Np, y, Nq, nc); return; }
    for (int e = 0; e < nc; e++) {
        if (o->c.d == 2) warped2d_mul(o, x + (size_t)Np * e, y + (size_t)Nq * e);
        else warped3d_mul(o, x + (size_t)Np * e, y + (size_t)Nq * e);
    }
}
static void Vt_mul(const ora_t *o, const double *x, double *y, int nc) {
    const int Nq = o->c.N_q, Np = o->c.N_p;
    if (o->c.v_kind == SSE_V_IDENTITY) { memcpy(y, x, sizeof(double) * Nq * nc); return; }
    if (o->c.v_kind == SSE_V_DENSE) { csr_mul_t(&o->Vcsr, x, Nq, y, Np, nc); return; }
    for (int e = 0; e < nc; e++) {
        if (o->c.d == 2) warped2d_mul_t(o, x + (size_t)Nq * e, y + (size_t)Np * e);
        else warped3d_mul_t(o, x + (size_t)Nq * e, y + (size_t)Np * e);
    }
}

/* mass_matrix_solve! (mass_matrix.jl:169-196); rhs is N_p x nc, temp N_q x nc */
static void mass_solve(const ora_t *o, int64_t k, double *rhs, double *temp, int nc) {
    const int Nq = o->c.N_q, Np = o->c.N_p;
    const double *W = o->a->W, *J = o->a->J_q + (size_t)Nq * k;
    if (o->c.mass_solver == SSE_MASS_DIAGONAL) { /* :177-183, WJ^-1 = inv(Diagonal(W .* J)) */
        for (int e = 0; e < nc; e++)
            for (int i = 0; i < Np; i++) rhs[i + (size_t)Np * e] *= 1.0 / (W[i] * J[i]);
        return;
    }
    if (o->c.mass_solver == SSE_MASS_CHOLESKY) { /* :169-175, ldiv!(cholesky(Symmetric(V' WJ V)), rhs): U' \ then U \ */
        const double *U = o->chol + (size_t)Np * Np * k;
        for (int e = 0; e < nc; e++) {
            double *b = rhs + (size_t)Np * e;
            for (int i = 0; i < Np; i++) {
                double t = b[i];
                for (int j = 0; j < i; j++) t -= U[j + (size_t)Np * i] * b[j];
                b[i] = t / U[i + (size_t)Np * i];
            }
            for (int i = Np - 1; i >= 0; i--) {
                double t = b[i];
                for (int j = i + 1; j < Np; j++) t -= U[i + (size_t)Np * j] * b[j];
                b[i] = t / U[i + (size_t)Np * i];
            }
        }
        return;
    }
    /* WeightAdjusted, M^-1 = I (:185-196; ctor :59-75) */
    V_mul(o, rhs, temp, nc);
    for (int e = 0; e < nc; e++)
        for (int i = 0; i < Nq; i++) temp[i + (size_t)Nq * e] *= W[i] / J[i];
    Vt_mul(o, temp, rhs, nc);
}

/* ---------------------------------------------------------------- physics */
/* ConservationLaws.jl:132-145 */
static inline double logmean(double x, double y) {
    double f2 = (x * (x - 2 * y) + y * y) / (x * (x + 2 * y) + y * y);
    if (f2 < 1.0e-4) return (x + y) * 105 / (210 + f2 * (70 + f2 * (42 + f2 * 30)));
    return (y - x) / log(y / x);
}
/* ConservationLaws.jl:147-156 */
static inline double inv_logmean(double x, double y) {
    double f2 = (x * (x - 2 * y) + y * y) / (x * (x + 2 * y) + y * y);
    if (f2 < 1.0e-4) return (210 + f2 * (70 + f2 * (42 + f2 * 30))) / ((x + y) * 105);
    return log(y / x) / (y - x);
}
/* euler_navierstokes.jl:58-68 */
static void euler_physical_flux(const sse_config *c, const double *u, double F[MAXC][MAXD]) {
    const int d = c->d;
    double V[MAXD], s = 0.0;
    for (int m = 0; m < d; m++) { V[m] = u[m + 1] / u[0]; s += u[m + 1] * V[m]; }
    double p = (c->gamma - 1) * (u[d + 1] - 0.5 * s), ht = u[d + 1] + p;
    for (int n = 0; n < d; n++) {
        F[0][n] = u[n + 1];
        for (int m = 0; m < d; m++) F[m + 1][n] = u[m + 1] * V[n] + (m == n ? p : 0.0);
        F[d + 1][n] = ht * V[n];
    }
}
/* compute_two_point_flux: advection linear_advection_diffusion.jl:113-119; Euler conservative
   euler_navierstokes.jl:152-158; Euler EC (Ranocha) euler_navierstokes.jl:171-195 */
static void two_point_flux(const sse_config *c, int tp, const double *uL, const double *uR, double F[MAXC][MAXD]) {
    const int d = c->d;
    if (c->pde != SSE_PDE_EULER) {
        double f1 = 0.5 * (uL[0] + uR[0]);
        if (c->pde == SSE_PDE_BURGERS || c->pde == SSE_PDE_VISCOUS_BURGERS)   /* BurgersType, burgers.jl:45, 111-143 */
            f1 = (tp == SSE_TWO_POINT_ENTROPY_CONSERVATIVE) ? (uL[0] * uL[0] + uL[0] * uR[0] + uR[0] * uR[0]) / 6
                                                            : (uL[0] * uL[0] + uR[0] * uR[0]) * 0.25;
        for (int m = 0; m < d; m++) F[0][m] = c->a[m] * f1;
        return;
    }
    if (tp == SSE_TWO_POINT_CONSERVATIVE) {
        double FL[MAXC][MAXD], FR[MAXC][MAXD];
        euler_physical_flux(c, uL, FL); euler_physical_flux(c, uR, FR);
        for (int e = 0; e < d + 2; e++) for (int m = 0; m < d; m++) F[e][m] = 0.5 * (FL[e][m] + FR[e][m]);
        return;
    }
    const double gm1 = c->gamma - 1, igm1 = 1 / (c->gamma - 1);
    double VL[MAXD], VR[MAXD], sL = 0, sR = 0, dot = 0;
    for (int m = 0; m < d; m++) { VL[m] = uL[m + 1] / uL[0]; VR[m] = uR[m + 1] / uR[0]; }
    for (int m = 0; m < d; m++) { sL += VL[m] * VL[m]; sR += VR[m] * VR[m]; dot += VL[m] * VR[m]; }
    double pL = gm1 * (uL[d + 1] - 0.5 * uL[0] * sL), pR = gm1 * (uR[d + 1] - 0.5 * uR[0] * sR);
    double rho_avg = logmean(uL[0], uR[0]);
    double Vavg[MAXD];
    for (int m = 0; m < d; m++) Vavg[m] = 0.5 * (VL[m] + VR[m]);
    double p_avg = 0.5 * (pL + pR);
    double Cc = 0.5 * dot + igm1 * inv_logmean(uL[0] / pL, uR[0] / pR);
    for (int n = 0; n < d; n++) {
        double frho = rho_avg * Vavg[n];
        F[0][n] = frho;
        for (int m = 0; m < d; m++) F[m + 1][n] = rho_avg * Vavg[m] * Vavg[n] + (m == n ? p_avg : 0.0);
        F[d + 1][n] = frho * Cc + 0.5 * (pL * VR[n] + pR * VL[n]);
    }
}
/* wave_speed: Euler euler_navierstokes.jl:133-150; advection linear_advection_diffusion.jl:106-111 */
static double wave_speed(const sse_config *c, const double *ui, const double *uo, const double *n) {
    const int d = c->d;
    if (c->pde != SSE_PDE_EULER) {
        double s = 0; for (int m = 0; m < d; m++) s += c->a[m] * n[m];
        if (c->pde == SSE_PDE_BURGERS || c->pde == SSE_PDE_VISCOUS_BURGERS) return fmax(fabs(s * ui[0]), fabs(s * uo[0]));    /* burgers.jl:103-109 */
        return fabs(s);
    }
    double si = 0, so = 0, vni = 0, vno = 0;
    for (int m = 0; m < d; m++) { si += ui[m + 1] * ui[m + 1]; so += uo[m + 1] * uo[m + 1]; }
    double pi_ = (c->gamma - 1) * (ui[d + 1] - (0.5 / ui[0]) * si), po = (c->gamma - 1) * (uo[d + 1] - (0.5 / uo[0]) * so);
    for (int m = 0; m < d; m++) { vni += ui[m + 1] / ui[0] * n[m]; vno += uo[m + 1] / uo[0] * n[m]; }
    double ci = sqrt(c->gamma * pi_ / ui[0]), co = sqrt(c->gamma * po / uo[0]);
    return fmax(fabs(vni), fabs(vno)) + fmax(ci, co);
}
/* euler_navierstokes.jl:100-113 ; generic identity ConservationLaws.jl:178-190 */
static void cons_to_entropy(const sse_config *c, const double *u, double *w) {
    const int d = c->d;
    if (c->pde != SSE_PDE_EULER) { w[0] = u[0]; return; }
    const double g = c->gamma, gm1 = g - 1, igm1 = 1 / gm1;
    double s = 0; for (int m = 0; m < d; m++) s += u[m + 1] * u[m + 1];
    double kk = (0.5 / u[0]) * s, p = gm1 * (u[d + 1] - kk), ip = 1.0 / p;
    w[0] = igm1 * (g - log(p / pow(u[0], g))) - kk * ip;
    for (int m = 0; m < d; m++) w[m + 1] = u[m + 1] * ip;
    w[d + 1] = -u[0] * ip;
}
/* euler_navierstokes.jl:115-131 */
static void entropy_to_cons(const sse_config *c, const double *win, double *u) {
    const int d = c->d;
    if (c->pde != SSE_PDE_EULER) { u[0] = win[0]; return; }
    const double g = c->gamma, gm1 = g - 1, igm1 = 1 / gm1;
    double w[MAXC];
    for (int e = 0; e < d + 2; e++) w[e] = win[e] * gm1;
    double s2 = 0; for (int m = 0; m < d; m++) s2 += w[m + 1] * w[m + 1];
    double kk = s2 / (2 * w[d + 1]);
    double s = g - w[0] + kk;
    double rho_e = pow(gm1 / pow(-w[d + 1], g), igm1) * exp(-s * igm1);
    u[0] = -w[d + 1] * rho_e;
    for (int m = 0; m < d; m++) u[m + 1] = w[m + 1] * rho_e;
    u[d + 1] = rho_e * (1 - kk);
}

/* numerical_flux! ConservationLaws.jl:75-101 (LF) and :103-128 (central / EC).
   u_in, u_out, f_star: N_f x N_c (ld = N_f); n_f: d x N_f */
static void numerical_flux(const ora_t *o, int tp, const double *u_in, const double *u_out, const double *n_f, double *f_star) {
    const sse_config *c = &o->c;
    const int d = c->d, Nc = c->N_c, Nf = c->N_f;
    for (int i = 0; i < Nf; i++) {
        double ui[MAXC], uo[MAXC], F[MAXC][MAXD];
        for (int e = 0; e < Nc; e++) { ui[e] = u_in[i + (size_t)Nf * e]; uo[e] = u_out[i + (size_t)Nf * e]; }
        two_point_flux(c, tp, ui, uo, F);
        if (c->inviscid_flux == SSE_FLUX_LAX_FRIEDRICHS) {
            double a = c->half_lambda * wave_speed(c, ui, uo, n_f + (size_t)d * i);
            for (int e = 0; e < Nc; e++) {
                double avg = 0.0;
                for (int m = 0; m < d; m++) avg = fma(F[e][m], n_f[m + (size_t)d * i], avg);
                f_star[i + (size_t)Nf * e] = fma(a, ui[e] - uo[e], avg);
            }
        } else {
            for (int e = 0; e < Nc; e++) {
                double t = 0.0;
                for (int m = 0; m < d; m++) t = fma(F[e][m], n_f[m + (size_t)d * i], t);
                f_star[i + (size_t)Nf * e] = t;
            }
        }
    }
}

/* ---------------------------------------------------------------- scratch */
typedef struct {
    double *f_q, *f_f, *f_n, *r_q, *temp, *u_in, *u_out, *u_n, *w, *nf, *aux;
} scr_t;
static scr_t scr_alloc(const sse_config *c) {
    size_t Nq = c->N_q, Nf = c->N_f, Nc = c->N_c, d = c->d, Np = c->N_p;
    size_t big = (Nq > Nf ? Nq : Nf);
    if (Np > big) big = Np;
    scr_t s;
    s.f_q = malloc(sizeof(double) * Nq * Nc * d);
    s.f_f = malloc(sizeof(double) * Nf * Nc);
    s.f_n = malloc(sizeof(double) * Nf * Nc);
    s.r_q = malloc(sizeof(double) * Nq * Nc);
    s.temp = malloc(sizeof(double) * big * Nc);
    s.u_in = malloc(sizeof(double) * Nf * Nc * (d + 1));
    s.u_out = malloc(sizeof(double) * Nf * Nc * (d + 1));
    s.u_n = malloc(sizeof(double) * Nf * Nc * d);
    s.w = malloc(sizeof(double) * big * Nc);
    s.nf = malloc(sizeof(double) * d * Nf);
    s.aux = malloc(sizeof(double) * big * Nc);
    return s;
}
static void scr_free(scr_t *s) {
    free(s->f_q); free(s->f_f); free(s->f_n); free(s->r_q); free(s->temp); free(s->u_in);
    free(s->u_out); free(s->u_n); free(s->w); free(s->nf); free(s->aux);
}

/* gather u_f[:, k, :] and u_f[CI[connectivity[:, k]], :]  (flux_differencing_form.jl:312-313).
   u_f layout (N_f, N_e, N_c) "switched order" (Solvers.jl:205) */
static void gather_facets(const ora_t *o, const double *u_f, int64_t k, double *u_in, double *u_out) {
    const int Nf = o->c.N_f, Nc = o->c.N_c;
    const int64_t Ne = o->c.N_e;
    const int64_t *mapP = o->a->mapP + (size_t)Nf * k;
    for (int e = 0; e < Nc; e++)
        for (int i = 0; i < Nf; i++) {
            u_in[i + (size_t)Nf * e] = u_f[i + (size_t)Nf * (k + Ne * e)];
            u_out[i + (size_t)Nf * e] = u_f[(mapP[i] - 1) + (size_t)Nf * Ne * e];
        }
}
/* n_f[m,:,k] = nJf[m,:,k] ./ J_f[:,k]  (operators.jl:19,59,115) */
static void normals(const ora_t *o, int64_t k, double *nf) {
    const int d = o->c.d, Nf = o->c.N_f;
    for (int i = 0; i < Nf; i++)
        for (int m = 0; m < d; m++) nf[m + d * i] = o->a->nJf[m + (size_t)d * (i + (size_t)Nf * k)] / o->a->J_f[i + (size_t)Nf * k];
}

/* ---------------------------------------------------------------- pass A: nodal_values! */
static void nodal_values(const ora_t *o, scr_t *s, const double *u, double *u_q, double *u_f, int64_t k) {
    const sse_config *c = &o->c;
    const int Nq = c->N_q, Nf = c->N_f, Nc = c->N_c, Np = c->N_p;
    const int64_t Ne = c->N_e;
    const double *uk = u + (size_t)Np * Nc * k;
    double *uqk = u_q + (size_t)Nq * Nc * k;
    double *ufk = s->f_f; /* N_f x N_c staging, then scattered into (N_f, N_e, N_c) */
    const int project = (c->form == SSE_FORM_FLUX_DIFFERENCING) && Nc > 1;
    if (!project) {
        /* standard_form_first_order.jl:1-14 ; flux_differencing_form.jl:253-265 */
        V_mul(o, uk, uqk, Nc);
        csr_mul(&o->Rcsr, uqk, Nq, ufk, Nf, Nc);
    } else if (c->v_kind == SSE_V_IDENTITY && !o->has_C) {
        /* flux_differencing_form.jl:171-187 (nodal, diagonal-E) */
        V_mul(o, uk, uqk, Nc);
        csr_mul(&o->Rcsr, uqk, Nq, ufk, Nf, Nc);
    } else if (c->v_kind == SSE_V_IDENTITY) {
        /* flux_differencing_form.jl:190-211 (nodal, general R) */
        double *w_q = s->r_q, *w_f = s->f_n;
        V_mul(o, uk, uqk, Nc);
        for (int i = 0; i < Nq; i++) {
            double ui[MAXC], wi[MAXC];
            for (int e = 0; e < Nc; e++) ui[e] = uqk[i + (size_t)Nq * e];
            cons_to_entropy(c, ui, wi);
            for (int e = 0; e < Nc; e++) w_q[i + (size_t)Nq * e] = wi[e];
        }
        csr_mul(&o->Rcsr, w_q, Nq, w_f, Nf, Nc);
        for (int i = 0; i < Nf; i++) {
            double wi[MAXC], ui[MAXC];
            for (int e = 0; e < Nc; e++) wi[e] = w_f[i + (size_t)Nf * e];
            entropy_to_cons(c, wi, ui);
            for (int e = 0; e < Nc; e++) ufk[i + (size_t)Nf * e] = ui[e];
        }
    } else {
        /* flux_differencing_form.jl:214-250 (general / modal) */
        double *w_q = s->r_q, *w_f = s->f_n, *w = s->w;
        V_mul(o, uk, uqk, Nc);
        for (int i = 0; i < Nq; i++) {
            double ui[MAXC], wi[MAXC];
            for (int e = 0; e < Nc; e++) ui[e] = uqk[i + (size_t)Nq * e];
            cons_to_entropy(c, ui, wi);
            for (int e = 0; e < Nc; e++) w_q[i + (size_t)Nq * e] = wi[e];
        }
        for (int e = 0; e < Nc; e++) /* lmul!(WJ, w_q) */
            for (int i = 0; i < Nq; i++) w_q[i + (size_t)Nq * e] *= o->a->W[i] * o->a->J_q[i + (size_t)Nq * k];
        Vt_mul(o, w_q, w, Nc);
        mass_solve(o, k, w, w_q, Nc);
        V_mul(o, w, w_q, Nc);
        csr_mul(&o->Rcsr, w_q, Nq, w_f, Nf, Nc);
        for (int i = 0; i < Nq; i++) {
            double wi[MAXC], ui[MAXC];
            for (int e = 0; e < Nc; e++) wi[e] = w_q[i + (size_t)Nq * e];
            entropy_to_cons(c, wi, ui);
            for (int e = 0; e < Nc; e++) uqk[i + (size_t)Nq * e] = ui[e];
        }
        for (int i = 0; i < Nf; i++) {
            double wi[MAXC], ui[MAXC];
            for (int e = 0; e < Nc; e++) wi[e] = w_f[i + (size_t)Nf * e];
            entropy_to_cons(c, wi, ui);
            for (int e = 0; e < Nc; e++) ufk[i + (size_t)Nf * e] = ui[e];
        }
    }
    for (int e = 0; e < Nc; e++)
        for (int i = 0; i < Nf; i++) u_f[i + (size_t)Nf * (k + Ne * e)] = ufk[i + (size_t)Nf * e];
}

/* ---------------------------------------------------------------- pass B variants */
/* physical_flux!: advection linear_advection_diffusion.jl:54-61, adv-diff :64-71, Euler euler_navierstokes.jl:85-91 */
static void physical_flux(const ora_t *o, const double *u_q, const double *q_q, double *f_q) {
    const sse_config *c = &o->c;
    const int d = c->d, Nq = c->N_q, Nc = c->N_c;
    if (c->pde == SSE_PDE_EULER) {
        for (int i = 0; i < Nq; i++) {
            double ui[MAXC], F[MAXC][MAXD];
            for (int e = 0; e < Nc; e++) ui[e] = u_q[i + (size_t)Nq * e];
            euler_physical_flux(c, ui, F);
            for (int e = 0; e < Nc; e++) for (int m = 0; m < d; m++) f_q[i + (size_t)Nq * (e + Nc * m)] = F[e][m];
        }
        return;
    }
    for (int m = 0; m < d; m++)
        for (int i = 0; i < Nq; i++) {
            double f = c->a[m] * u_q[i];
            if (c->pde == SSE_PDE_BURGERS) f = 0.5 * c->a[m] * u_q[i] * u_q[i];           /* burgers.jl:52-58 */
            if (c->pde == SSE_PDE_ADVECTION_DIFFUSION) f = c->a[m] * u_q[i] - c->b * q_q[i + (size_t)Nq * Nc * m];
            if (c->pde == SSE_PDE_VISCOUS_BURGERS) f = 0.5 * c->a[m] * u_q[i] * u_q[i] - c->b * q_q[i + (size_t)Nq * Nc * m];   /* burgers.jl:60-70 */
            f_q[i + (size_t)Nq * Nc * m] = f;
        }
}

/* standard_form_first_order.jl:16-63 */
static void time_derivative_standard_reference(const ora_t *o, scr_t *s, double *u_q, const double *u_f, double *dudt, int64_t k) {
    const sse_config *c = &o->c;
    const int d = c->d, Nq = c->N_q, Nf = c->N_f, Nc = c->N_c, Np = c->N_p;
    double *uqk = u_q + (size_t)Nq * Nc * k, *dk = dudt + (size_t)Np * Nc * k;
    const double *W = o->a->W, *Lam = o->a->Lambda_q + (size_t)Nq * d * d * k;
    physical_flux(o, uqk, NULL, s->f_q);
    gather_facets(o, u_f, k, s->u_in, s->u_out);
    normals(o, k, s->nf);
    numerical_flux(o, SSE_TWO_POINT_CONSERVATIVE, s->u_in, s->u_out, s->nf, s->f_f);
    memset(s->r_q, 0, sizeof(double) * Nq * Nc);
    for (int n = 0; n < d; n++) {
        const double *fqn = s->f_q + (size_t)Nq * Nc * n;
        for (int m = 0; m < d; m++) {
            const double *L = Lam + (size_t)Nq * (m + d * n); /* Λ_q[:, m, n, k] */
            for (int e = 0; e < Nc; e++)
                for (int i = 0; i < Nq; i++) s->temp[i + (size_t)Nq * e] = (0.5 * W[i] * L[i]) * fqn[i + (size_t)Nq * e];
            csr_mul_t(&o->Dcsr[m], s->temp, Nq, uqk, Nq, Nc);
            for (int t = 0; t < Nq * Nc; t++) s->r_q[t] += uqk[t];
            csr_mul(&o->Dcsr[m], fqn, Nq, uqk, Nq, Nc);
            for (int e = 0; e < Nc; e++)
                for (int i = 0; i < Nq; i++) uqk[i + (size_t)Nq * e] *= (0.5 * W[i] * L[i]);
            for (int t = 0; t < Nq * Nc; t++) s->r_q[t] -= uqk[t];
        }
        csr_mul(&o->Rcsr, fqn, Nq, s->f_n, Nf, Nc);
        for (int e = 0; e < Nc; e++)
            for (int i = 0; i < Nf; i++) s->f_f[i + (size_t)Nf * e] -= (0.5 * s->nf[n + d * i]) * s->f_n[i + (size_t)Nf * e];
    }
    for (int e = 0; e < Nc; e++)
        for (int i = 0; i < Nf; i++) s->f_f[i + (size_t)Nf * e] *= o->a->Bf[i] * o->a->J_f[i + (size_t)Nf * k];
    csr_mul_t(&o->Rcsr, s->f_f, Nf, s->temp, Nq, Nc);
    for (int t = 0; t < Nq * Nc; t++) s->r_q[t] -= s->temp[t];
    Vt_mul(o, s->r_q, dk, Nc);
    mass_solve(o, k, dk, s->temp, Nc);
}

/* dense per-element GEMV:  y (N_p x nc) (+)= A (N_p x n) x (n x nc) */
static void gemv_acc(const double *A, int Np, int n, const double *x, int nc, double sign, double *y) {
    for (int e = 0; e < nc; e++)
        for (int i = 0; i < Np; i++) {
            double t = 0.0;
            for (int j = 0; j < n; j++) t = fma(A[i + (size_t)Np * j], x[j + (size_t)n * e], t);
            y[i + (size_t)Np * e] += sign * t;
        }
}

/* standard_form_first_order.jl:65-94 */
static void time_derivative_standard_physical(const ora_t *o, scr_t *s, const double *u_q, const double *u_f, double *dudt, int64_t k) {
    const sse_config *c = &o->c;
    const int d = c->d, Nq = c->N_q, Nf = c->N_f, Nc = c->N_c, Np = c->N_p;
    const double *uqk = u_q + (size_t)Nq * Nc * k;
    double *dk = dudt + (size_t)Np * Nc * k;
    physical_flux(o, uqk, NULL, s->f_q);
    gather_facets(o, u_f, k, s->u_in, s->u_out);
    normals(o, k, s->nf);
    numerical_flux(o, SSE_TWO_POINT_CONSERVATIVE, s->u_in, s->u_out, s->nf, s->f_f);
    memset(dk, 0, sizeof(double) * Np * Nc);
    for (int m = 0; m < d; m++)
        gemv_acc(o->a->VOL + (size_t)Np * Nq * (m + (size_t)d * k), Np, Nq, s->f_q + (size_t)Nq * Nc * m, Nc, 1.0, dk);
    gemv_acc(o->a->FAC + (size_t)Np * Nf * k, Np, Nf, s->f_f, Nc, 1.0, dk);
}

/* standard_form_second_order.jl:3-34 ; BR1 solution flux linear_advection_diffusion.jl:75-86.
   q_q: (N_q, N_c, d, N_e); q_f: (N_f, N_e, N_c, d) */
static void auxiliary_variable(const ora_t *o, scr_t *s, const double *u_q, const double *u_f, double *q_q, double *q_f, double *dudt, int64_t k) {
    const sse_config *c = &o->c;
    const int d = c->d, Nq = c->N_q, Nf = c->N_f, Nc = c->N_c, Np = c->N_p;
    const int64_t Ne = c->N_e;
    const double *uqk = u_q + (size_t)Nq * Nc * k;
    double *dk = dudt + (size_t)Np * Nc * k;
    gather_facets(o, u_f, k, s->u_in, s->u_out);
    normals(o, k, s->nf);
    for (int m = 0; m < d; m++)
        for (int e = 0; e < Nc; e++)
            for (int i = 0; i < Nf; i++)
                s->u_n[i + (size_t)Nf * (e + Nc * m)] = 0.5 * (s->u_in[i + (size_t)Nf * e] + s->u_out[i + (size_t)Nf * e]) * s->nf[m + d * i];
    for (int m = 0; m < d; m++) {
        memset(dk, 0, sizeof(double) * Np * Nc);
        gemv_acc(o->a->VOL + (size_t)Np * Nq * (m + (size_t)d * k), Np, Nq, uqk, Nc, -1.0, dk);
        gemv_acc(o->a->FAC + (size_t)Np * Nf * k, Np, Nf, s->u_n + (size_t)Nf * Nc * m, Nc, -1.0, dk);
        double *qqm = q_q + (size_t)Nq * Nc * (m + (size_t)d * k);
        V_mul(o, dk, qqm, Nc);
        csr_mul(&o->Rcsr, qqm, Nq, s->f_n, Nf, Nc);
        for (int e = 0; e < Nc; e++)
            for (int i = 0; i < Nf; i++) q_f[i + (size_t)Nf * (k + Ne * (e + (size_t)Nc * m))] = s->f_n[i + (size_t)Nf * e];
    }
}

/* standard_form_second_order.jl:38-75 ; BR1 viscous flux linear_advection_diffusion.jl:90-102 */
static void time_derivative_second_order(const ora_t *o, scr_t *s, const double *u_q, const double *u_f, const double *q_q, const double *q_f, double *dudt, int64_t k) {
    const sse_config *c = &o->c;
    const int d = c->d, Nq = c->N_q, Nf = c->N_f, Nc = c->N_c, Np = c->N_p;
    const int64_t Ne = c->N_e;
    const double *uqk = u_q + (size_t)Nq * Nc * k;
    const int64_t *mapP = o->a->mapP + (size_t)Nf * k;
    double *dk = dudt + (size_t)Np * Nc * k;
    physical_flux(o, uqk, q_q + (size_t)Nq * Nc * d * k, s->f_q);
    gather_facets(o, u_f, k, s->u_in, s->u_out);
    normals(o, k, s->nf);
    numerical_flux(o, SSE_TWO_POINT_CONSERVATIVE, s->u_in, s->u_out, s->nf, s->f_f);
    for (int e = 0; e < Nc; e++)
        for (int i = 0; i < Nf; i++) {
            double acc = 0.0;
            for (int m = 0; m < d; m++) {
                double qi = q_f[i + (size_t)Nf * (k + Ne * (e + (size_t)Nc * m))];
                double qo = q_f[(mapP[i] - 1) + (size_t)Nf * Ne * (e + (size_t)Nc * m)];
                double minus_q_avg = -0.5 * (qi + qo);
                acc += c->b * minus_q_avg * s->nf[m + d * i];
            }
            s->f_f[i + (size_t)Nf * e] += acc;
        }
    memset(dk, 0, sizeof(double) * Np * Nc);
    for (int m = 0; m < d; m++)
        gemv_acc(o->a->VOL + (size_t)Np * Nq * (m + (size_t)d * k), Np, Nq, s->f_q + (size_t)Nq * Nc * m, Nc, 1.0, dk);
    gemv_acc(o->a->FAC + (size_t)Np * Nf * k, Np, Nf, s->f_f, Nc, 1.0, dk);
}

/* flux_difference! sparse (flux_differencing_form.jl:37-75; the dense method :1-35 visits the same
   pairs).  Lam: Λ_q[:,:,:,k] */
static void flux_difference(const ora_t *o, const double *Lam, const double *u_q, double *r_q) {
    const sse_config *c = &o->c;
    const int d = c->d, Nq = c->N_q, Nc = c->N_c;
    memset(r_q, 0, sizeof(double) * Nq * Nc);
    for (int m = 0; m < d; m++) {
        const sp_t *S = &o->Scsc[m];
        for (int j = 0; j < Nq; j++)
            for (int ii = S->ptr[j]; ii < S->ptr[j + 1]; ii++) {
                int i = S->idx[ii];
                if (i < j) {
                    double ui[MAXC], uj[MAXC], F[MAXC][MAXD];
                    for (int e = 0; e < Nc; e++) { ui[e] = u_q[i + (size_t)Nq * e]; uj[e] = u_q[j + (size_t)Nq * e]; }
                    two_point_flux(c, c->two_point_flux, ui, uj, F);
                    double Sm = S->val[ii];
                    for (int e = 0; e < Nc; e++) {
                        double Fm = 0.0;
                        for (int n = 0; n < d; n++) {
                            double Lij = Lam[i + (size_t)Nq * (m + d * n)] + Lam[j + (size_t)Nq * (m + d * n)];
                            Fm = fma(Lij, F[e][n], Fm);
                        }
                        double diff = Sm * Fm;
                        r_q[i + (size_t)Nq * e] -= diff;
                        r_q[j + (size_t)Nq * e] += diff;
                    }
                }
            }
    }
}

/* facet_correction! sparse (flux_differencing_form.jl:126-168; dense :91-124) */
static void facet_correction(const ora_t *o, int64_t k, const double *u_q, const double *u_f_in, double *r_q, double *f_f) {
    const sse_config *c = &o->c;
    const int d = c->d, Nq = c->N_q, Nf = c->N_f, Nc = c->N_c, Nfac = c->N_fac;
    const int npf = Nf / Nfac;
    const sp_t *C = &o->Ccsc;
    const double *Lam = o->a->Lambda_q + (size_t)Nq * d * d * k;
    for (int j = 0; j < Nf; j++)
        for (int ii = C->ptr[j]; ii < C->ptr[j + 1]; ii++) {
            int i = C->idx[ii];
            double ui[MAXC], uj[MAXC], F[MAXC][MAXD];
            for (int e = 0; e < Nc; e++) { ui[e] = u_q[i + (size_t)Nq * e]; uj[e] = u_f_in[j + (size_t)Nf * e]; }
            two_point_flux(c, c->two_point_flux, ui, uj, F);
            double Cij = C->val[ii];
            int f = j / npf;
            double nJ[MAXD];
            for (int m = 0; m < d; m++) {
                double hq;
                if (o->a->nJq) hq = 0.5 * o->a->nJq[m + (size_t)d * (f + (size_t)Nfac * (i + (size_t)Nq * k))];
                else { /* mesh.jl:262-269 */
                    double t = 0.0;
                    for (int l = 0; l < d; l++) t += Lam[i + (size_t)Nq * (l + d * m)] * o->a->nref[l + d * f];
                    hq = 0.5 * t;
                }
                nJ[m] = 0.5 * o->a->nJf[m + (size_t)d * (j + (size_t)Nf * k)] + hq;
            }
            for (int e = 0; e < Nc; e++) {
                double Fn = 0.0;
                for (int m = 0; m < d; m++) Fn = fma(nJ[m], F[e][m], Fn);
                double diff = Cij * Fn;
                r_q[i + (size_t)Nq * e] -= diff;
                f_f[j + (size_t)Nf * e] -= diff;
            }
        }
}

/* flux_differencing_form.jl:294-347 */
static void time_derivative_flux_differencing(const ora_t *o, scr_t *s, double *u_q, const double *u_f, double *dudt, int64_t k) {
    const sse_config *c = &o->c;
    const int d = c->d, Nq = c->N_q, Nf = c->N_f, Nc = c->N_c, Np = c->N_p;
    double *uqk = u_q + (size_t)Nq * Nc * k, *dk = dudt + (size_t)Np * Nc * k;
    gather_facets(o, u_f, k, s->u_in, s->u_out);
    normals(o, k, s->nf);
    numerical_flux(o, c->two_point_flux, s->u_in, s->u_out, s->nf, s->f_f);
    for (int e = 0; e < Nc; e++)
        for (int i = 0; i < Nf; i++) s->f_f[i + (size_t)Nf * e] *= o->a->Bf[i] * o->a->J_f[i + (size_t)Nf * k];
    flux_difference(o, o->a->Lambda_q + (size_t)Nq * d * d * k, uqk, s->r_q);
    if (o->has_C) facet_correction(o, k, uqk, s->u_in, s->r_q, s->f_f);
    csr_mul_t(&o->Rcsr, s->f_f, Nf, uqk, Nq, Nc);
    for (int t = 0; t < Nq * Nc; t++) s->r_q[t] -= uqk[t];
    Vt_mul(o, s->r_q, dk, Nc);
    mass_solve(o, k, dk, uqk, Nc);
}

/* ---------------------------------------------------------------- driver */
static int ora_init(ora_t *o, const sse_config *cfg, const sse_arrays *arr) {
    memset(o, 0, sizeof(*o));
    o->c = *cfg; o->a = arr; o->P1 = cfg->p + 1;
    const int Nq = cfg->N_q, Nf = cfg->N_f, Np = cfg->N_p, d = cfg->d;
    if (cfg->N_c > MAXC || d > MAXD || o->P1 > 8) return SSE_ERR_UNSUPPORTED;
    if (cfg->N_ghost != 0) return SSE_ERR_UNSUPPORTED;
    if (cfg->v_kind == SSE_V_DENSE) o->Vcsr = dense_to_csr(arr->V, Nq, Np);
    o->Rcsr = dense_to_csr(arr->R, Nf, Nq);
    for (int m = 0; m < d; m++) {
        if (arr->D[m]) o->Dcsr[m] = dense_to_csr(arr->D[m], Nq, Nq);
        if (arr->S[m]) o->Scsc[m] = dense_to_csc(arr->S[m], Nq, Nq);
    }
    o->has_C = arr->Cfd != NULL;
    if (o->has_C) o->Ccsc = dense_to_csc(arr->Cfd, Nq, Nf);
    if (cfg->v_kind == SSE_V_WARPED) {
        int P1 = o->P1, n = 1;
        for (int m = 0; m < d; m++) n *= P1;
        o->sig_i = malloc(sizeof(int64_t) * n);
        for (int t = 0; t < n; t++) o->sig_i[t] = arr->sigma_i[t] - 1;
        int no = 1;
        for (int m = 0; m < d; m++) no *= cfg->M1d[m];
        o->sig_o = malloc(sizeof(int64_t) * no);
        for (int t = 0; t < no; t++) o->sig_o[t] = arr->sigma_o[t] - 1;
        /* N2, N3 trip counts (warped_product_3d.jl:24-30, warped_product_2d.jl:15) */
        for (int b1 = 0; b1 < P1; b1++) {
            int cnt = 0;
            if (d == 2) { for (int b2 = 0; b2 < P1; b2++) cnt += arr->sigma_i[b1 + P1 * b2] > 0; }
            else { for (int b2 = 0; b2 < P1; b2++) cnt += arr->sigma_i[b1 + P1 * (b2 + P1 * 0)] > 0; }
            o->N2[b1] = cnt;
            if (d == 3)
                for (int b2 = 0; b2 < P1; b2++) {
                    int c3 = 0;
                    for (int b3 = 0; b3 < P1; b3++) c3 += arr->sigma_i[b1 + P1 * (b2 + P1 * b3)] > 0;
                    o->N3[b1][b2] = c3;
                }
        }
    }
    if (cfg->mass_solver == SSE_MASS_CHOLESKY) {
        /* CholeskySolver (mass_matrix.jl:26-39): with V = I it is the DiagonalSolver; otherwise
           M_k = V' diag(W J_k) V column by column and its upper Cholesky factor (LAPACK potrf 'U' order) */
        if (cfg->v_kind == SSE_V_IDENTITY) { o->c.mass_solver = SSE_MASS_DIAGONAL; return SSE_OK; }
        o->chol = malloc(sizeof(double) * (size_t)Np * Np * cfg->N_e);
        double *e_c = calloc(Np, sizeof(double)), *vq = malloc(sizeof(double) * Nq);
        for (int64_t k = 0; k < cfg->N_e; k++) {
            double *M = o->chol + (size_t)Np * Np * k;
            const double *J = arr->J_q + (size_t)Nq * k;
            for (int c = 0; c < Np; c++) {
                e_c[c] = 1.0;
                V_mul(o, e_c, vq, 1);
                for (int i = 0; i < Nq; i++) vq[i] *= arr->W[i] * J[i];
                Vt_mul(o, vq, M + (size_t)Np * c, 1);
                e_c[c] = 0.0;
            }
            for (int j = 0; j < Np; j++) {          /* U'U = M, column by column, upper triangle in place */
                for (int i = 0; i <= j; i++) {
                    double t = M[i + (size_t)Np * j];
                    for (int l = 0; l < i; l++) t -= M[l + (size_t)Np * i] * M[l + (size_t)Np * j];
                    if (i < j) M[i + (size_t)Np * j] = t / M[i + (size_t)Np * i];
                    else { if (t <= 0.0) { free(e_c); free(vq); return SSE_ERR_BAD_ARGUMENT; } M[i + (size_t)Np * j] = sqrt(t); }
                }
            }
        }
        free(e_c); free(vq);
    }
    return SSE_OK;
}
static void ora_free(ora_t *o) {
    sp_free(&o->Vcsr); sp_free(&o->Rcsr); sp_free(&o->Ccsc);
    for (int m = 0; m < MAXD; m++) { sp_free(&o->Dcsr[m]); sp_free(&o->Scsc[m]); }
    free(o->sig_i); free(o->sig_o); free(o->chol);
}

/* semi_discrete_residual! (Solvers.jl:474-564), Threaded variant.  Scratch u_q (N_q,N_c,N_e) and
   u_f (N_f,N_e,N_c) are returned to the caller when the pointers are non-NULL (state after pass A). */
int32_t sse_oracle_rhs(const sse_config *cfg, const sse_arrays *arr, const double *u, double *dudt,
                       int32_t nthreads, double *u_q_out, double *u_f_out) {
    ora_t o;
    int rc = ora_init(&o, cfg, arr);
    if (rc) return rc;
    const sse_config *c = &o.c;
    const int64_t Ne = c->N_e;
    const size_t Nq = c->N_q, Nf = c->N_f, Nc = c->N_c, d = c->d;
    const int second = (c->pde == SSE_PDE_ADVECTION_DIFFUSION || c->pde == SSE_PDE_VISCOUS_BURGERS);
    double *u_q = malloc(sizeof(double) * Nq * Nc * Ne);
    double *u_f = malloc(sizeof(double) * Nf * Nc * Ne);
    double *q_q = second ? malloc(sizeof(double) * Nq * Nc * d * Ne) : NULL;
    double *q_f = second ? malloc(sizeof(double) * Nf * Nc * d * Ne) : NULL;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        scr_t s = scr_alloc(c);
#pragma omp for schedule(static)
        for (int64_t k = 0; k < Ne; k++) nodal_values(&o, &s, u, u_q, u_f, k);
        /* implicit barrier: end of the first Threads.@threads loop (Solvers.jl:505-507) */
#pragma omp single
        {
            if (u_q_out) memcpy(u_q_out, u_q, sizeof(double) * Nq * Nc * Ne);
            if (u_f_out) memcpy(u_f_out, u_f, sizeof(double) * Nf * Nc * Ne);
        }
        if (second) {
#pragma omp for schedule(static)
            for (int64_t k = 0; k < Ne; k++) auxiliary_variable(&o, &s, u_q, u_f, q_q, q_f, dudt, k);
        }
#pragma omp for schedule(static)
        for (int64_t k = 0; k < Ne; k++) {
            if (second) time_derivative_second_order(&o, &s, u_q, u_f, q_q, q_f, dudt, k);
            else if (c->form == SSE_FORM_FLUX_DIFFERENCING) time_derivative_flux_differencing(&o, &s, u_q, u_f, dudt, k);
            else if (c->form == SSE_FORM_STANDARD_REFERENCE) time_derivative_standard_reference(&o, &s, u_q, u_f, dudt, k);
            else time_derivative_standard_physical(&o, &s, u_q, u_f, dudt, k);
        }
        scr_free(&s);
    }
    free(u_q); free(u_f); free(q_q); free(q_f);
    ora_free(&o);
    return SSE_OK;
}

/* Repeated evaluation for CPU-baseline timing: returns seconds per RHS (best of `reps`), keeping the
   operator setup and scratch allocation outside the timed region like the reference's preallocated
   Solver does. */
double sse_oracle_time_rhs(const sse_config *cfg, const sse_arrays *arr, const double *u, double *dudt,
                           int32_t nthreads, int32_t reps, int32_t *used_threads) {
    ora_t o;
    if (ora_init(&o, cfg, arr)) return -1.0;
    const sse_config *c = &o.c;
    const int64_t Ne = c->N_e;
    const size_t Nq = c->N_q, Nf = c->N_f, Nc = c->N_c, d = c->d;
    const int second = (c->pde == SSE_PDE_ADVECTION_DIFFUSION || c->pde == SSE_PDE_VISCOUS_BURGERS);
    double *u_q = malloc(sizeof(double) * Nq * Nc * Ne);
    double *u_f = malloc(sizeof(double) * Nf * Nc * Ne);
    double *q_q = second ? malloc(sizeof(double) * Nq * Nc * d * Ne) : NULL;
    double *q_f = second ? malloc(sizeof(double) * Nf * Nc * d * Ne) : NULL;
    double best = 1e300;
    int nt = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    nt = omp_get_max_threads();
#endif
    if (used_threads) *used_threads = nt;
    scr_t *S = malloc(sizeof(scr_t) * nt);
    for (int t = 0; t < nt; t++) S[t] = scr_alloc(c);
    for (int r = 0; r < reps; r++) {
        double t0 = 0, t1 = 0;
#ifdef _OPENMP
        t0 = omp_get_wtime();
#endif
#pragma omp parallel
        {
            int tid = 0;
#ifdef _OPENMP
            tid = omp_get_thread_num();
#endif
            scr_t *s = &S[tid];
#pragma omp for schedule(static)
            for (int64_t k = 0; k < Ne; k++) nodal_values(&o, s, u, u_q, u_f, k);
            if (second) {
#pragma omp for schedule(static)
                for (int64_t k = 0; k < Ne; k++) auxiliary_variable(&o, s, u_q, u_f, q_q, q_f, dudt, k);
            }
#pragma omp for schedule(static)
            for (int64_t k = 0; k < Ne; k++) {
                if (second) time_derivative_second_order(&o, s, u_q, u_f, q_q, q_f, dudt, k);
                else if (c->form == SSE_FORM_FLUX_DIFFERENCING) time_derivative_flux_differencing(&o, s, u_q, u_f, dudt, k);
                else if (c->form == SSE_FORM_STANDARD_REFERENCE) time_derivative_standard_reference(&o, s, u_q, u_f, dudt, k);
                else time_derivative_standard_physical(&o, s, u_q, u_f, dudt, k);
            }
        }
#ifdef _OPENMP
        t1 = omp_get_wtime();
#endif
        if (t1 - t0 < best) best = t1 - t0;
    }
    for (int t = 0; t < nt; t++) scr_free(&S[t]);
    free(S); free(u_q); free(u_f); free(q_q); free(q_f);
    ora_free(&o);
    return best;
}

/* pointwise helpers exported for unit tests of the physics */
double sse_oracle_logmean(double x, double y) { return logmean(x, y); }
double sse_oracle_inv_logmean(double x, double y) { return inv_logmean(x, y); }
void sse_oracle_two_point_flux(const sse_config *cfg, int32_t tp, const double *uL, const double *uR, double *F /* N_c x d col-major */) {
    double Ft[MAXC][MAXD];
    two_point_flux(cfg, tp, uL, uR, Ft);
    for (int e = 0; e < cfg->N_c; e++) for (int m = 0; m < cfg->d; m++) F[e + cfg->N_c * m] = Ft[e][m];
}
void sse_oracle_cons_to_entropy(const sse_config *cfg, const double *u, double *w) { cons_to_entropy(cfg, u, w); }
void sse_oracle_entropy_to_cons(const sse_config *cfg, const double *w, double *u) { entropy_to_cons(cfg, w, u); }
