// common.cuh — device-side views of the Solver image shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "physics.cuh"

namespace sse {

struct SpMat {              // compressed rows: row r owns entries [ptr[r], ptr[r+1])
    const int* ptr;
    const int* idx;
    const double* val;
    // the same rows slot-major, padded to the longest row with column -1: entry q of row r at [q * rows + r].  Kernels map one
    // row to one thread, so a warp reads consecutive addresses (with CSR every lane reads its own cache line: a dense R or D
    // of the multidimensional schemes costs 32 L1 wavefronts per load).  Entries keep their CSR order: same sums, same bits.
    // Narrow rows (w <= 16: Kronecker D, selection-like R) stay row-major (sq = 1, sr = w): a thread's row is one or two cache
    // lines fetched by its first entry, which is what small, latency-bound meshes want.
    int rows, w, sq, sr;        // entry q of row r at [q * sq + r * sr]
    const int* ei;
    const double* ev;
};
// s += sum_q A[r, c_q] x[c_q * 1 + off]  over the stored entries of row r
#define SSE_ROW_FOR(A, r, c, v)                                                              \
    for (int q_ = 0, c = 0; q_ < (A).w && (c = (A).ei[q_ * (A).sq + (r) * (A).sr]) >= 0; q_++) \
        if (const double v = (A).ev[q_ * (A).sq + (r) * (A).sr]; true)

// Reference-element operators (element independent), device pointers.
struct Ops {
    int d, NC, Np, Nq, Nf, Nfac, npf;
    int v_kind;
    int P1, M1, M2, M3;            // warped tensors (2-D embedded as M3 = 1)
    const double* Vd;              // dense V, column-major Nq x Np
    int v_small;                   // warped V of a small element, applied through its dense matrix Vd (sse_create)
    const double *A, *B, *C;       // warped_product_3d.jl:2-35
    const int *sig_i, *sig_o;      // 0-based, -1 = unused
    int N2[8];
    int N3[64];                    // [b1 * 8 + b2]
    SpMat R, Rt;                   // R by facet rows; R^T by volume-node rows
    SpMat D[3], Dt[3];             // D_m by rows; D_m^T by rows
    // flux differencing: for volume node i the partners j != i with the d values S_m[i,j]
    const int* vol_ptr;
    const int* vol_j;
    const double* vol_S;           // [entry * d + m]
    SpMat Cq;                      // C by volume-node rows  (j, C_ij)
    SpMat Cf;                      // C by facet-node rows   (i, C_ij)
    int has_C;
    // the same three pair lists slot-major (padded to the longest row, -1 = no entry): entry q of row r at [q * rows + r], so
    // that the threads of a warp (one row each) read consecutive addresses -- with dense operators a row has N_q - 1 / N_f entries
    int vol_w, cq_w, cf_w;
    const int *vol_je, *cq_je, *cf_ie;
    const double *vol_Se;          // [(q * d + m) * Nq + i]
    const double *cq_ve, *cf_ve;
    const double *W, *Bf, *nref;   // nref: d x Nfac column-major
};

// dense all-pairs tables of k_time_fluxdiff_dense (kernels_generic.cuh)
struct DenseDev {
    const double* S4;      // [(j * D + m) * Nq + i]  S_m[i, j] / 4 (skew-extended: -S_m[j, i] below the diagonal)
    const double* C4;      // [j * Nq + i]            C[i, j] / 4
    const double* RT;      // [j * Nq + i]            R[j, i]: the lift r_q -= R' f_f as a dense product, coalesced over i
    int ok;
};

// Per-element geometry and scratch, device pointers (reference layouts).
struct Geo {
    const double* J_q;       // Nq x Ne
    const double* Lambda_q;  // Nq x d x d x Ne
    const double* J_f;       // Nf x Ne
    const double* nJf;       // d x Nf x Ne
    const double* nJq;       // d x Nfac x Nq x Ne or nullptr
    const double* VOL;       // Np x Nq x d x Ne
    const double* FAC;       // Np x Nf x Ne
    const long long* mapP;   // Nf x Ne, 1-based linear index into the (Nf, Ne [+ghost]) facet array
    long long Ne;
    long long NFT;           // Nf*Ne + N_ghost: stride between variables of u_f / q_f
    int mass_solver;
    const double* iJW;       // W / J_q per volume node (Nq x Ne), built at sse_create for the compile-time Euler path; else nullptr
    int* flag;               // set to 1 by the kernels that write dudt when a value is not finite (SSE_ERR_NONFINITE)
    const double* chol;      // CholeskySolver: upper factors U_k, Np x Np column-major per element (mass_matrix.jl:30-39)
};

// a residual entry that is NaN or infinite raises the handle's flag (one predicated store, never taken on physical states)
__device__ __forceinline__ void flag_nonfinite(int* flag, double v) {
    if (!(fabs(v) <= 1.7976931348623157e308)) *flag = 1;
}

// asynchronous global -> shared copies (LDGSTS): the data never passes through registers, so a CTA prologue can put all of
// its loads in flight at once and wait for them where they are first needed
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

#define SSE_FOR(t, n) for (int t = threadIdx.x; t < (n); t += blockDim.x)

// Element packing of the one-CTA-per-element kernels (generic and tensor-line): blockDim = (threads per element, elements per
// CTA).  Row threadIdx.y works on element first + blockIdx.x * blockDim.y + threadIdx.y in its own slice of the dynamic shared
// memory; the host launches full CTAs only (a second launch with one row per CTA takes the remainder), so no row is idle.
// With 32 threads per element a row is one warp and its phases are ordered by warp barriers: the rows of a CTA never wait for
// each other.  (Small elements -- 1-D, triangles and quadrilaterals up to 32 nodes -- used 25 of 64 lanes of a CTA of their own.)
__device__ __forceinline__ long long sse_element(long long first) { return first + (long long)blockIdx.x * blockDim.y + threadIdx.y; }
__device__ __forceinline__ double* sse_row_smem(double* sm) {
    unsigned bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(bytes));
    return sm + (size_t)threadIdx.y * (bytes / (8u * blockDim.y));
}
__device__ __forceinline__ void sse_sync() {
    if (blockDim.x == 32) __syncwarp();
    else __syncthreads();
}

}  // namespace sse
