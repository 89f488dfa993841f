/*
 * sse_b200.h — C ABI of libsse_b200.so, the B200-native drop-in for
 * StableSpectralElements.jl's semi-discrete residual
 *
 *     semi_discrete_residual!(dudt, u, solver::Solver, t)       src/Solvers/Solvers.jl:474-564
 *
 * One handle == one reference `Solver` (src/Solvers/Solvers.jl:259-272) living on
 * one GPU.  The Julia side keeps `Solver`, `ConservationLaws`,
 * `SpatialDiscretization`, `semidiscretize` and the OrdinaryDiffEq integration; a
 * new `AbstractParallelism` subtype (Solvers.jl:72,79-80) carries the handle and its
 * `semi_discrete_residual!` method is one `ccall` of sse_rhs.
 *
 * Conventions
 *  - All arrays are the reference's arrays verbatim: Float64, column-major (first
 *    index fastest), sizes as documented per field.  Integer index arrays are
 *    Int64 and 1-based, exactly as Julia holds them.
 *  - sse_create copies everything it needs to the device; the caller keeps
 *    ownership of the host arrays.  Device state vectors (u, dudt, ...) are plain
 *    device pointers to N_p*N_c*N_e doubles in the reference layout
 *    (N_p, N_c, N_e); they may come from sse_state_alloc or from any other CUDA
 *    allocator in the same process (CUDA.jl, torch, cudaMalloc).
 *  - Every entry point returns an int32 status (0 = SSE_OK).  Nothing throws
 *    across the boundary; sse_last_error_string() describes the last failure of
 *    the calling thread.  There is no CPU fallback: without a CUDA device every
 *    compute entry point fails with SSE_ERR_CUDA.
 *  - Calls on one handle are asynchronous with respect to the host (enqueued on
 *    the handle's stream) except download/synchronize/functionals.
 */
#ifndef SSE_B200_H
#define SSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSE_ABI_VERSION 1

/* status codes */
enum {
    SSE_OK = 0,
    SSE_ERR_BAD_ARGUMENT = 1, /* reference: DimensionMismatch from LinearMaps.check_dim_mul etc. */
    SSE_ERR_UNSUPPORTED = 2,  /* reference: MethodError (combination has no method)           */
    SSE_ERR_CUDA = 3,
    SSE_ERR_NONFINITE = 4,    /* reference: DomainError from log/sqrt of a non-physical state  */
    SSE_ERR_COMM = 5
};

/* AbstractConservationLaw subtypes on the path (src/ConservationLaws) */
enum { SSE_PDE_ADVECTION = 0,            /* LinearAdvectionEquation{d}           linear_advection_diffusion.jl:10-20 */
       SSE_PDE_ADVECTION_DIFFUSION = 1,  /* LinearAdvectionDiffusionEquation{d}  linear_advection_diffusion.jl:30-42 */
       SSE_PDE_EULER = 2,                /* EulerEquations{d}                    euler_navierstokes.jl:23-38        */
       SSE_PDE_BURGERS = 3,              /* InviscidBurgersEquation{d} (a in sse_config.a)  burgers.jl:1-21          */
       SSE_PDE_VISCOUS_BURGERS = 4 };    /* ViscousBurgersEquation{d} (a, b): F = a u^2/2 - b q, BR1   burgers.jl:23-49, 59-99 */

/* residual form x operator strategy (Solvers.jl:75-115, constructors :287-376) */
enum { SSE_FORM_STANDARD_REFERENCE = 0,  /* StandardForm + ReferenceOperator  standard_form_first_order.jl:16-63 */
       SSE_FORM_STANDARD_PHYSICAL = 1,   /* StandardForm + PhysicalOperator   standard_form_first_order.jl:65-94,
                                            and every SecondOrder law          standard_form_second_order.jl:3-75 */
       SSE_FORM_FLUX_DIFFERENCING = 2 }; /* FluxDifferencingForm              flux_differencing_form.jl:294-347  */

/* AbstractInviscidNumericalFlux (ConservationLaws.jl:52-63) */
enum { SSE_FLUX_LAX_FRIEDRICHS = 0, SSE_FLUX_CENTRAL = 1, SSE_FLUX_ENTROPY_CONSERVATIVE = 2 };
/* AbstractViscousNumericalFlux (ConservationLaws.jl:65-68) */
enum { SSE_VISCOUS_NONE = 0, SSE_VISCOUS_BR1 = 1 };
/* AbstractTwoPointFlux (ConservationLaws.jl:70-73) */
enum { SSE_TWO_POINT_CONSERVATIVE = 0, SSE_TWO_POINT_ENTROPY_CONSERVATIVE = 1 };
/* AbstractMassMatrixSolver (mass_matrix.jl:1-17) */
enum { SSE_MASS_WEIGHT_ADJUSTED = 0,     /* M^-1 = I assumed (assume_orthonormal=true default, mass_matrix.jl:59-75) */
       SSE_MASS_DIAGONAL = 1,
       SSE_MASS_CHOLESKY = 2 };          /* CholeskySolver: ldiv!(cholesky(Symmetric(V' WJ_k V)), rhs)  mass_matrix.jl:26-39, 169-175;
                                            the factors are computed on the device at sse_create (N_p^2 doubles per element) */
/* type of the generalized Vandermonde V (MatrixFreeOperators) */
enum { SSE_V_IDENTITY = 0,               /* LinearMaps.UniformScalingMap (nodal schemes)               */
       SSE_V_DENSE = 1,                  /* OctavianMap / WrappedMap                                   */
       SSE_V_WARPED = 2 };               /* WarpedTensorProductMap2D/3D  warped_product_{2d,3d}.jl     */

typedef struct sse_config {
    int32_t abi_version;     /* SSE_ABI_VERSION */
    int32_t d;               /* spatial dimension 1..3 */
    int32_t N_c, N_p, N_q, N_f, N_fac;
    int32_t p;               /* polynomial degree (size of warped tensors = p+1) */
    int64_t N_e;             /* elements owned by this handle */
    int64_t N_ghost;         /* ghost facet nodes appended to u_f after the N_f*N_e owned ones
                                (multi-GPU halo); 0 for a single GPU */
    int32_t pde, form, inviscid_flux, viscous_flux, two_point_flux, mass_solver, v_kind;
    int32_t M1d[3];          /* 1-D node counts of the warped tensors (size(A,1), size(B,1), size(C,1)) */
    double  half_lambda;     /* LaxFriedrichsNumericalFlux.halfλ (ConservationLaws.jl:54-60) */
    double  a[3];            /* advection velocity */
    double  b;               /* diffusion coefficient */
    double  gamma;           /* EulerEquations.γ */
} sse_config;

/* Host arrays of the reference Solver.  Unused pointers are NULL. */
typedef struct sse_arrays {
    /* --- reference-element operators (ReferenceApproximation, SpatialDiscretizations.jl:191-251) */
    const double*  V;        /* N_q x N_p  Matrix(V)            (v_kind = DENSE)                      */
    const double*  A;        /* M1 x (p+1)                      (v_kind = WARPED; warped_product_3d.jl:2-35) */
    const double*  B;        /* M2 x (p+1) x (p+1)                                                     */
    const double*  C;        /* M3 x (p+1)^3 (d = 3 only)                                              */
    const int64_t* sigma_i;  /* (p+1)^d, 1-based modal index, 0 = unused                               */
    const int64_t* sigma_o;  /* M1 x M2 [x M3], 1-based nodal index                                    */
    const double*  R;        /* N_f x N_q  Matrix(R)                                                   */
    const double*  W;        /* N_q   diag(W)                                                          */
    const double*  Bf;       /* N_f   diag(B)                                                          */
    const double*  D[3];     /* N_q x N_q Matrix(D[m])  (StandardForm+ReferenceOperator)               */
    const double*  S[3];     /* N_q x N_q Matrix(S[m])  (FluxDifferencingOperators.S, Solvers.jl:166)  */
    const double*  Cfd;      /* N_q x N_f Matrix(C), NULL when C === nothing (diag-E; Solvers.jl:167)  */
    /* --- geometric factors (GeometricFactors, SpatialDiscretizations.jl:283-289), as passed to the
           Solver constructors (Solvers.jl:287-376).  For StandardForm+ReferenceOperator, Lambda_q is the
           composite metric returned by apply_reference_mapping (SpatialDiscretizations.jl:398-411). */
    const double*  J_q;      /* N_q x N_e                                                              */
    const double*  Lambda_q; /* N_q x d x d x N_e                                                      */
    const double*  J_f;      /* N_f x N_e                                                              */
    const double*  nJf;      /* d x N_f x N_e                                                          */
    const double*  nJq;      /* d x N_fac x N_q x N_e, or NULL: recomputed from Lambda_q and nref
                                as in mesh.jl:262-269                                                  */
    const double*  nref;     /* d x N_fac reference normal of each face (first node of the face)       */
    /* --- PhysicalOperators (Solvers.jl:147-154, operators.jl:83-160) */
    const double*  VOL;      /* N_p x N_q x d x N_e   VOL[k][m]                                        */
    const double*  FAC;      /* N_p x N_f x N_e       FAC[k]                                           */
    /* --- connectivity: mesh.mapP (Solvers.jl:268,305), N_f x N_e, 1-based linear index into the
           (N_f, N_e [+ghost]) facet array */
    const int64_t* mapP;
} sse_arrays;

typedef struct sse_handle sse_handle;

/* -- lifetime ------------------------------------------------------------------------------- */
/* replaces Solver(conservation_law, spatial_discretization, form, strategy, alg, mass_solver,
   parallelism)  (Solvers.jl:287-376) */
int32_t sse_create(const sse_config* cfg, const sse_arrays* arr, int32_t device, sse_handle** out);
int32_t sse_destroy(sse_handle* h);
/* run all work of this handle on a caller-owned CUDA stream (cudaStream_t); NULL = default */
int32_t sse_set_stream(sse_handle* h, void* cuda_stream);
/* 0: generic kernels only; 1 (default): use the tensor-line specialised kernels when the
   operators have the collapsed tensor-product structure */
int32_t sse_set_kernel_variant(sse_handle* h, int32_t variant);
/* reports the family in use: 0 generic, 1 tensor-line, 2 compile-time (p = 2..7 tets; p = 2..4 triangles: 2-D Euler flux
   differencing and 2-D advection StandardForm), 3 dense all-pairs (multidimensional schemes) */
int32_t sse_get_kernel_variant(const sse_handle* h, int32_t* variant);

/* -- state vectors (u, dudt of Solvers.jl:474-483; layout (N_p, N_c, N_e)) --------------------- */
int32_t sse_state_alloc(sse_handle* h, double** d_out);
int32_t sse_state_free(sse_handle* h, double* d_ptr);
int32_t sse_state_fill(sse_handle* h, double* d_x, double value);      /* fill!(x, value), zero(x) */
int32_t sse_state_upload(sse_handle* h, double* d_dst, const double* h_src);
int32_t sse_state_download(sse_handle* h, double* h_dst, const double* d_src);

/* -- the hot path ------------------------------------------------------------------------------ */
/* semi_discrete_residual!(dudt, u, solver, t)  (Solvers.jl:474-564): all passes.  On an element-partitioned handle
   (N_ghost > 0, after sse_comm_init + sse_halo_plan) the same call also runs the facet-trace halo exchange of its rank
   (NCCL send/recv on a side stream, overlapped with pass B of the interior elements); every rank calls it.
   Non-physical states: a residual entry that is NaN / Inf raises a device flag; the next blocking call on the handle
   (sse_synchronize, sse_state_download, sse_rhs_host) returns SSE_ERR_NONFINITE once -- the reference's DomainError. */
int32_t sse_rhs(sse_handle* h, const double* d_u, double* d_dudt, double t);

/* The same call on HOST arrays (`Array{Float64,3}` arguments, as OrdinaryDiffEq passes CPU state to
 * semi_discrete_residual!, Solvers.jl:474-483): upload, both passes and download are pipelined over `chunks` element
 * ranges (<= 0: default) on three streams.  Synchronous; page-lock the arrays with sse_host_pin for full overlap. */
int32_t sse_rhs_host(sse_handle* h, const double* h_u, double* h_dudt, double t, int32_t chunks);
/* the schedule sse_rhs_host follows, from mapP alone (no device needed): order[i] = range uploaded i-th, ready[c] = upload
 * position after which pass B of range c may run (every range holding one of its face neighbours is through pass A) */
int32_t sse_host_range_plan(const int64_t* mapP, int64_t N_e, int32_t N_f, int32_t N_fac, int32_t chunks, int32_t* order, int32_t* ready);
int32_t sse_host_pin(void* p, int64_t bytes);
int32_t sse_host_unpin(void* p);
/* Split form (what sse_rhs is made of; also lets a caller run its own exchange).  Pass A (nodal_values!, Solvers.jl:505-507)
   fills u_q and the owned part of u_f and packs the halo send buffer; the caller exchanges halos
   (NCCL) into sse_halo_recv_buffer; pass B (time_derivative!, Solvers.jl:509-511) runs on the
   element range [first, first+count).  Second-order laws have an extra aux pass + exchange. */
/* Contract of the split form: pass B of an element range consumes the u_q scratch pass A wrote for it (the compile-time
   kernels overwrite it with r_q, as the reference reuses u_q[:,:,k], flux_differencing_form.jl:341-346).  Pass B is therefore
   NOT idempotent on the 3-D compile-time Euler path: run pass A again before repeating pass B on a range; after pass B
   sse_debug_views shows r_q, not u_q.  The warp-per-element triangle kernels (2-D Euler flux differencing, 2-D advection) do
   all of pass B in one launch and leave the scratch untouched: there pass B is a pure function of the scratch (on the 2-D
   advection path the u_q scratch holds the modal coefficients pass A copied, not nodal values).  No path reads the caller's
   state again after pass A. */
int32_t sse_rhs_pass_a(sse_handle* h, const double* d_u);
/* pass A on the element range [first, first+count): lets a host-buffer caller overlap the upload of u with pass A */
int32_t sse_rhs_pass_a_range(sse_handle* h, const double* d_u, int64_t first, int64_t count);
int32_t sse_rhs_pass_aux(sse_handle* h, double* d_dudt, int64_t first, int64_t count);
int32_t sse_rhs_pass_b(sse_handle* h, double* d_dudt, int64_t first, int64_t count);
/* halo plumbing: the send list holds 1-based linear indices into the owned (N_f, N_e) facet array */
int32_t sse_halo_configure(sse_handle* h, const int64_t* send_index, int64_t n_send);
int32_t sse_halo_pack(sse_handle* h, int32_t which /*0: u_f, 1: q_f*/);
int32_t sse_halo_send_buffer(sse_handle* h, double** d_buf, int64_t* n_doubles);
int32_t sse_halo_recv_buffer(sse_handle* h, int32_t which, double** d_buf, int64_t* n_doubles);
int32_t sse_halo_unpack(sse_handle* h, int32_t which);

/* -- multi-GPU (SURVEY.md §8b "Threading", §8e): elements partitioned over GPUs, facet-trace halos only --------------------
   The reference's plug-in point is the `parallelism` kwarg of semidiscretize (Solvers.jl:438) and the element loops of
   Solvers.jl:495-514; a handle built from a partition (N_ghost > 0, elements ordered interior first, mapP rewritten to
   local + ghost numbering) joins a communicator and then behaves like a single-GPU handle: sse_rhs, sse_rhs_lsrk,
   sse_step_ck54, sse_rhs_host and sse_functionals (which all-reduces) are collective over the ranks.
   (a) one process per GPU: rank 0 calls sse_comm_unique_id, the host broadcasts the 128 bytes (MPI.jl, Distributed.jl,
       torch.distributed ...), every rank calls sse_comm_init;
   (b) one process driving N GPUs (the Julia host stays one process): sse_comm_init_all over the N handles, then
       sse_rhs_multi / sse_step_ck54_multi advance all of them from the single host thread.
   NCCL is loaded at run time (libnccl.so.2, or $SSE_NCCL_LIB); a partition onto one rank needs no NCCL at all. */
int32_t sse_comm_unique_id(uint8_t* id128);
int32_t sse_comm_init(sse_handle* h, const uint8_t* id128, int32_t rank, int32_t world);
int32_t sse_comm_init_all(sse_handle* const* handles, int32_t n);
/* (b) without NCCL: the handles of one process exchange their halos by peer-to-peer copies (two partitions may share a GPU) */
int32_t sse_comm_init_local(sse_handle* const* handles, int32_t n);
int32_t sse_comm_info(const sse_handle* h, int32_t* rank, int32_t* world, int32_t* nccl_version);
/* Halo plan of this rank: neighbour ranks, facet nodes sent to / received from each (the receive segments are the ghost slots
   in order, their sum is N_ghost), the concatenated send list (1-based linear indices into the owned (N_f, N_e) facet array)
   and the number of interior elements: elements [0, n_interior) read no ghost slot (verified against mapP). */
int32_t sse_halo_plan(sse_handle* h, int32_t n_nbr, const int32_t* nbr_rank, const int64_t* send_count, const int64_t* recv_count,
                      const int64_t* send_index, int64_t n_interior);
/* (b): the residual / one CarpenterKennedy2N54 step on all handles of one process; d_u[i], d_tmp[i], d_dudt[i] live on the
   device of handles[i].  The halo exchanges of all handles form one NCCL group. */
int32_t sse_rhs_multi(sse_handle* const* handles, int32_t n, double* const* d_u, double* const* d_dudt, double t);
int32_t sse_step_ck54_multi(sse_handle* const* handles, int32_t n, double* const* d_u, double* const* d_tmp, double* const* d_dudt,
                            double t, double dt);

/* Partition helper (host only, no device): the local view of `rank` of an element partition, from the reference's global
   connectivity mesh.mapP (N_f x N_e, 1-based; Solvers.jl:207) and the owner rank of every element (slabs, blocks, or any graph
   partition of the face adjacency).  Local elements are ordered interior first; ghost slots are grouped by neighbour rank
   (ascending) and ordered by global facet-node index inside a rank, which is also the order of the owner's send list -- both
   sides agree without communicating.  sse_partition_fill returns exactly what sse_create (mapP_local, N_e = n_local,
   N_ghost = n_ghost; the operator / metric arrays of the elements elem_gid) and sse_halo_plan take. */
typedef struct sse_partition sse_partition;
int32_t sse_partition_create(const int64_t* mapP, int64_t N_e, int32_t N_f, const int32_t* owner, int32_t n_parts, int32_t rank,
                             sse_partition** out);
int32_t sse_partition_sizes(const sse_partition* p, int64_t* n_local, int64_t* n_interior, int64_t* n_ghost, int32_t* n_nbr, int64_t* n_send);
/* elem_gid[n_local] (1-based global element ids), mapP_local[N_f x n_local] (1-based, local + ghost numbering), nbr_rank[n_nbr],
   send_count[n_nbr], recv_count[n_nbr], send_index[n_send] (1-based into the owned (N_f, n_local) facet array); NULL skips */
int32_t sse_partition_fill(const sse_partition* p, int64_t* elem_gid, int64_t* mapP_local, int32_t* nbr_rank, int64_t* send_count,
                           int64_t* recv_count, int64_t* send_index);
int32_t sse_partition_destroy(sse_partition* p);

/* -- callers either side of the path ----------------------------------------- */
/* y = a*x + b*y on state vectors (OrdinaryDiffEq broadcast updates) */
int32_t sse_axpby(sse_handle* h, double a, const double* d_x, double b, double* d_y);
/* 2N low-storage RK stage: tmp = A*tmp + dt*dudt ; u += B*tmp   (CarpenterKennedy2N54 stage) */
int32_t sse_lsrk_stage(sse_handle* h, double* d_u, double* d_tmp, const double* d_dudt,
                       double A, double B, double dt);
/* semi_discrete_residual! + one 2N-storage stage in one call; on the compile-time kernel path the stage update is fused
   into the epilogue of the projection kernel, so the state is updated where dudt is produced */
int32_t sse_rhs_lsrk(sse_handle* h, double* d_u, double* d_tmp, double* d_dudt, double A, double B, double dt, double t);
/* one full CarpenterKennedy2N54 step on device (collective on element-partitioned handles, like sse_rhs) */
int32_t sse_step_ck54(sse_handle* h, double* d_u, double* d_tmp, double* d_dudt, double t, double dt);
/* on != 0: sse_step_ck54 of a single-GPU handle captures its launches into a CUDA graph once per (u, tmp, dudt, dt) and replays
   it with one cudaGraphLaunch -- for meshes small enough that the step is bound by launch latency (the reference's 2-D examples) */
int32_t sse_set_graph_mode(sse_handle* h, int32_t on);
/* conservation / energy / entropy residuals of Analysis/conservation.jl:145-189.
   out[0..N_c-1] = sum_k 1' WJ_k V dudt[:,e,k];  out[N_c] = sum_k u_k' M_k dudt_k (energy);
   out[N_c+1] = sum_k (P_k w(V u_k))' M_k dudt_k (entropy; Euler only, else 0).  Blocking.  With a communicator the sums
   run over all ranks (NCCL all-reduce; every rank calls and receives the totals). */
int32_t sse_functionals(sse_handle* h, const double* d_u, const double* d_dudt, double* out);

/* -- misc -------------------------------------------------------------------------------------- */
int32_t sse_synchronize(sse_handle* h);
const char* sse_last_error_string(void);
int32_t sse_abi_version(void);
/* one residual with a CUDA event between every kernel, averaged over `reps` runs: ms[0] pass A, ms[1] auxiliary pass (BR1),
   ms[2] first kernel of pass B (pair / derivative kernel), ms[3] second kernel of pass B (projection; 0 if pass B is one kernel) */
int32_t sse_profile_rhs(sse_handle* h, const double* d_u, double* d_dudt, int32_t reps, double* ms);
/* kernels launched through this handle so far (bench.py's gpu_launches) */
int32_t sse_launch_count(const sse_handle* h, int64_t* n);
/* scratch views for testing/analysis: u_q (N_q,N_c,N_e) and u_f (N_f,N_e+ghost,N_c) device pointers */
int32_t sse_debug_views(sse_handle* h, double** d_u_q, double** d_u_f);
/* host-only diagnostic: build the tensor-line pair schedule for (cfg, arr) and replay it against S and C.
   info[0..7] = {kernel family (0 generic, 1 tensor-line, 2 compile-time flux differencing, 3 compile-time advection
   StandardForm, 4 warp-per-element 2-D Euler flux differencing on triangles, 5 warp-per-element 2-D advection StandardForm
   on triangles), threads per CTA, volume rounds, facet sub-rounds, reducer items max, reducer
   sources max, shared-memory bytes, two-point flux evaluations per element}; max_err = largest deviation of
   the replayed S_m / C from the operators passed in (0 when every pair is visited exactly once). */
int32_t sse_plan_selfcheck(const sse_config* cfg, const sse_arrays* arr, int32_t* info, double* max_err);
/* register-resident DFMA microbenchmark: achieved FP64 FLOP/s on the handle's device (FMA = 2) */
int32_t sse_fp64_peak(int32_t device, double* flops_per_s);

/* -- device-side geometry --------------------------------------------------------
   GeometricFactors(mesh, reference_element, metric_type) (mesh.jl:229-506) for a whole mesh: from the mapping-node
   coordinates xyz[m] (N_map, N_e) to J_q (N_q,N_e), Λ_q (N_q,d,d,N_e), J_f (N_f,N_e), nJf (d,N_f,N_e) — exactly the arrays
   sse_arrays takes.  All matrices are column-major as Julia stores them; all pointers are HOST pointers (the library
   stages element chunks through the device).  Operators of the reference element (StartUpDG RefElemData fields):
   Drst = (Dr, Ds, Dt) (N_map,N_map), Vq (N_q,N_map), Vf (N_f,N_map), nrstJ[m] concatenated as (N_f, d).
   SSE_METRIC_EXACT: mesh.jl:229-282.  SSE_METRIC_CURL: 2-D mesh.jl:284-339; 3-D conservative curl form evaluated on N1
   nodes with derivative matrices D1 (N1,N1) and interpolations Vq1 (N_q,N1), Vf1 (N_f,N1): for Hex N1 = N_map and
   up = NULL (mesh.jl:341-408); for Tet the degree N+1 nodes with up = N_to_Nplus1 (N1,N_map), Vq1 = Vq*Nplus1_to_N
   (mesh.jl:410-506). */
enum { SSE_METRIC_EXACT = 0, SSE_METRIC_CURL = 1 };
typedef struct {
    int32_t d, N_map, N1, N_q, N_f, metric;
    int64_t N_e;
} sse_geom_config;
typedef struct {
    const double* Drst[3];
    const double* Vq;
    const double* Vf;
    const double* nrstJ;
    const double* up;        /* NULL when N1 == N_map */
    const double* D1[3];     /* 3-D curl only */
    const double* Vq1;
    const double* Vf1;
} sse_geom_ops;
int32_t sse_geometric_factors(const sse_geom_config* cfg, const sse_geom_ops* ops, int32_t device, const double* const xyz[3],
                              double* J_q, double* Lambda_q, double* J_f, double* nJf);

#ifdef __cplusplus
}
#endif
#endif /* SSE_B200_H */
