// kernels_geometry.cuh — GeometricFactors on the device.
//
// Restates src/SpatialDiscretizations/mesh.jl of the reference for whole meshes at once:
//   ExactMetrics             mesh.jl:229-282   dx/dr at the mapping nodes, interpolated to the volume / facet quadrature
//                                              nodes, then metrics() (mesh.jl:181-227) point by point
//   ConservativeCurl / ChanWilcox metrics
//     2-D                    mesh.jl:284-339   StartUpDG geometric_factors(x, y, Dr, Ds) at the mapping nodes, interpolated
//     3-D Hex                mesh.jl:341-408   curl form collocated at the mapping nodes (N1 = N_map, no `up`)
//     3-D Tet                mesh.jl:410-506   curl argument formed on the degree N+1 nodes (`up`, D1), metrics of degree N
// and the normals  nJf[m,i] = sum_l Lambda_f[i,l,m] nrstJ[l][i],  J_f = |nJf|.
// One CTA per element; every operator is a small dense matrix (column-major, as Julia stores it) read through L1/L2.
#pragma once
#include <cuda_runtime.h>
#include "../../include/sse_b200.h"

namespace sse {

struct GeomDev {
    int d, Nmap, N1, Nq, Nf, metric;
    long long Ne;
    const double* Drst[3];
    const double* D1[3];
    const double *Vq, *Vf, *up, *Vq1, *Vf1, *nrstJ;
    const double* xyz[3];        // (N_map, N_e) each
    double *J_q, *Lambda_q, *J_f, *nJf;
};

// y[i] = sum_j M[i + rows*j] v[j], i < rows, for the threads of the CTA
__device__ __forceinline__ void geo_matvec(const double* __restrict__ M, int rows, int cols, const double* v, double* y) {
    for (int i = threadIdx.x; i < rows; i += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < cols; j++) s = fma(M[i + (size_t)rows * j], v[j], s);
        y[i] = s;
    }
}

// J and Lambda = J inv(dx/dr) (rows l = reference, columns m = physical) from a[m][n] = dx_m / dr_n   mesh.jl:181-227
template <int D>
__device__ __forceinline__ void geo_metrics(const double a[3][3], double& J, double L[3][3]) {
    if constexpr (D == 1) { J = a[0][0]; L[0][0] = 1.0; }
    if constexpr (D == 2) {
        J = a[0][0] * a[1][1] - a[0][1] * a[1][0];
        L[0][0] = a[1][1]; L[0][1] = -a[0][1]; L[1][0] = -a[1][0]; L[1][1] = a[0][0];
    }
    if constexpr (D == 3) {
        L[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1]; L[0][1] = a[0][2] * a[2][1] - a[0][1] * a[2][2]; L[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
        L[1][0] = a[1][2] * a[2][0] - a[1][0] * a[2][2]; L[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0]; L[1][2] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
        L[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0]; L[2][1] = a[0][1] * a[2][0] - a[0][0] * a[2][1]; L[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
        J = a[0][0] * L[0][0] + a[0][1] * L[1][0] + a[0][2] * L[2][0];
    }
}

// shared layout (doubles): X[D][N1] | G[D*D][N1] (dx/dr or the metric at the nodes) | T[3][N1] | Jn[Nmap]
template <int D>
__global__ void __launch_bounds__(128) k_geometry(GeomDev g) {
    extern __shared__ double sm[];
    const int Nmap = g.Nmap, N1 = g.N1, Nq = g.Nq, Nf = g.Nf;
    const int NM = N1 > Nmap ? N1 : Nmap;
    double* X = sm;                    // [D][NM]
    double* G = X + D * NM;            // [D*D][NM]   entry (l, m) at l + D*m
    double* T = G + D * D * NM;        // [3][NM]
    double* Jn = T + 3 * NM;           // [Nmap]
    const long long k = blockIdx.x;
    for (int t = threadIdx.x; t < D * Nmap; t += blockDim.x) { const int m = t / Nmap, j = t - m * Nmap; X[m * NM + j] = g.xyz[m][(size_t)Nmap * k + j]; }
    __syncthreads();
    const bool curl = g.metric == SSE_METRIC_CURL && D > 1;
    // dx_m/dr_n at the mapping nodes: G[(m + D*n)]
    for (int m = 0; m < D; m++)
        for (int n = 0; n < D; n++) geo_matvec(g.Drst[n], Nmap, Nmap, X + m * NM, G + (m + D * n) * NM);
    __syncthreads();
    if (curl) {
        // Jacobian as a degree-N polynomial at the mapping nodes
        for (int i = threadIdx.x; i < Nmap; i += blockDim.x) {
            double a[3][3], J, L[3][3];
            for (int m = 0; m < D; m++) for (int n = 0; n < D; n++) a[m][n] = G[(m + D * n) * NM + i];
            geo_metrics<D>(a, J, L);
            Jn[i] = J;
        }
        __syncthreads();
        if constexpr (D == 2) {
            // rxJ = ys, sxJ = -yr, ryJ = -xs, syJ = xr at the nodes (mesh.jl:305), entry (l, m) at (l + D*m): overwrites dx/dr
            for (int i = threadIdx.x; i < Nmap; i += blockDim.x) {
                const double xr = G[(0 + D * 0) * NM + i], xs = G[(0 + D * 1) * NM + i], yr = G[(1 + D * 0) * NM + i], ys = G[(1 + D * 1) * NM + i];
                G[(0 + D * 0) * NM + i] = ys; G[(1 + D * 0) * NM + i] = -yr; G[(0 + D * 1) * NM + i] = -xs; G[(1 + D * 1) * NM + i] = xr;
            }
            __syncthreads();
        }
        if constexpr (D == 3) {
            // coordinates on the N1 nodes
            if (g.up) {
                double* Y = G;          // reuse G as a temporary for the lifted coordinates (dx/dr is already folded into Jn)
                for (int m = 0; m < D; m++) geo_matvec(g.up, N1, Nmap, X + m * NM, Y + m * NM);
                __syncthreads();
                for (int t = threadIdx.x; t < D * N1; t += blockDim.x) { const int m = t / N1, j = t - m * N1; X[m * NM + j] = Y[m * NM + j]; }
                __syncthreads();
            }
            // Lambda(l, m) = sgn_m * curl_xi( b grad_xi a )_l with (a, b, sgn) = (y, z, +), (x, z, -), (y, x, -)
            const int ia[3] = {1, 0, 1}, ib[3] = {2, 2, 0};
            const double sg[3] = {1.0, -1.0, -1.0};
            for (int m = 0; m < 3; m++) {
                const double* a = X + ia[m] * NM;
                const double* b = X + ib[m] * NM;
                for (int n = 0; n < 3; n++) geo_matvec(g.D1[n], N1, N1, a, T + n * NM);
                __syncthreads();
                for (int t = threadIdx.x; t < 3 * N1; t += blockDim.x) { const int n = t / N1, j = t - n * N1; T[n * NM + j] *= b[j]; }
                __syncthreads();
                // comp_r = Dt Fs - Ds Ft, comp_s = Dr Ft - Dt Fr, comp_t = Ds Fr - Dr Fs
                for (int i = threadIdx.x; i < N1; i += blockDim.x) {
                    double cr = 0.0, cs = 0.0, ct = 0.0;
                    for (int j = 0; j < N1; j++) {
                        const double dr = g.D1[0][i + (size_t)N1 * j], ds = g.D1[1][i + (size_t)N1 * j], dt = g.D1[2][i + (size_t)N1 * j];
                        const double Fr = T[0 * NM + j], Fs = T[1 * NM + j], Ft = T[2 * NM + j];
                        cr = fma(dt, Fs, fma(-ds, Ft, cr));
                        cs = fma(dr, Ft, fma(-dt, Fr, cs));
                        ct = fma(ds, Fr, fma(-dr, Fs, ct));
                    }
                    G[(0 + D * m) * NM + i] = sg[m] * cr;
                    G[(1 + D * m) * NM + i] = sg[m] * cs;
                    G[(2 + D * m) * NM + i] = sg[m] * ct;
                }
                __syncthreads();
            }
        }
        // interpolate the nodal metric and Jacobian to the quadrature nodes
        const double* Vq1 = (D == 3) ? g.Vq1 : g.Vq;
        const double* Vf1 = (D == 3) ? g.Vf1 : g.Vf;
        const int Nn = (D == 3) ? N1 : Nmap;
        for (int i = threadIdx.x; i < Nq; i += blockDim.x) {
            double J = 0.0;
            for (int j = 0; j < Nmap; j++) J = fma(g.Vq[i + (size_t)Nq * j], Jn[j], J);
            g.J_q[(size_t)Nq * k + i] = J;
            for (int e = 0; e < D * D; e++) {
                double s = 0.0;
                for (int j = 0; j < Nn; j++) s = fma(Vq1[i + (size_t)Nq * j], G[e * NM + j], s);
                g.Lambda_q[((size_t)k * D * D + e) * Nq + i] = s;
            }
        }
        for (int i = threadIdx.x; i < Nf; i += blockDim.x) {
            double L[3][3];
            for (int e = 0; e < D * D; e++) {
                double s = 0.0;
                for (int j = 0; j < Nn; j++) s = fma(Vf1[i + (size_t)Nf * j], G[e * NM + j], s);
                L[e % D][e / D] = s;
            }
            double n2 = 0.0;
            for (int m = 0; m < D; m++) {
                double s = 0.0;
                for (int l = 0; l < D; l++) s = fma(L[l][m], g.nrstJ[i + (size_t)Nf * l], s);
                g.nJf[m + D * ((size_t)Nf * k + i)] = s;
                n2 = fma(s, s, n2);
            }
            g.J_f[(size_t)Nf * k + i] = sqrt(n2);
        }
        return;
    }
    // exact metrics: interpolate dx/dr, then metrics() at every quadrature node
    for (int i = threadIdx.x; i < Nq + Nf; i += blockDim.x) {
        const bool vol = i < Nq;
        const int ii = vol ? i : i - Nq, rows = vol ? Nq : Nf;
        const double* V = vol ? g.Vq : g.Vf;
        double a[3][3], J, L[3][3];
        for (int m = 0; m < D; m++)
            for (int n = 0; n < D; n++) {
                double s = 0.0;
                for (int j = 0; j < Nmap; j++) s = fma(V[ii + (size_t)rows * j], G[(m + D * n) * NM + j], s);
                a[m][n] = s;
            }
        geo_metrics<D>(a, J, L);
        if (vol) {
            g.J_q[(size_t)Nq * k + ii] = J;
            for (int m = 0; m < D; m++) for (int l = 0; l < D; l++) g.Lambda_q[((size_t)k * D * D + (l + D * m)) * Nq + ii] = L[l][m];
        } else {
            double n2 = 0.0;
            for (int m = 0; m < D; m++) {
                double s = 0.0;
                for (int l = 0; l < D; l++) s = fma(L[l][m], g.nrstJ[ii + (size_t)Nf * l], s);
                g.nJf[m + D * ((size_t)Nf * k + ii)] = s;
                n2 = fma(s, s, n2);
            }
            g.J_f[(size_t)Nf * k + ii] = sqrt(n2);
        }
    }
}

}  // namespace sse
