/*
 * jfem_b200.h -- C ABI of libjfem_b200.so: the B200 (sm_100a) implementation of JuliaFEM's
 * 3D-elasticity hot path (element integration -> global scatter -> matrix-free / assembled
 * K.u -> CG / Newton-Krylov).  This is the drop-in boundary: a Julia `ccall` shim
 * (juliafem.jl_b200/julia/JuliaFEMB200.jl), the Python ctypes host (juliafem.jl_b200/_lib.py) and
 * any other FFI bind exactly these symbols.  Plain pointers and sizes only; no torch types.
 *
 * Each entry point names the reference interface it replaces (paths relative to JuliaFEM.jl):
 *   ext = ext/JuliaFEMCUDAExt.jl,  cpu = src/backend/cpu.jl,  eas = src/element_assembly_structures.jl,
 *   pe  = src/problems_elasticity.jl, sp = src/sparse/sparse.jl.
 *
 * Conventions
 *   - every function returns 0 on success, a JFEM_E* code otherwise; jfem_last_error() gives the
 *     message (the reference throws Julia exceptions; the shim turns non-zero into error(msg)).
 *     CG / Newton non-convergence is NOT an error (ext:630-631 only warns): status 0, iters = max.
 *   - ids are the reference's: nodes 1..n_nodes, dof = 3*(node-1)+c, c=1..3 (src/assembly/problems.jl:476).
 *     `index_base` (0 or 1) says what the caller's arrays use; Julia passes 1.
 *   - all floating point is IEEE double.  Vectors are n_dofs = 3*n_nodes long, reference dof order.
 *   - `on_device` != 0: vector pointers are CUDA device pointers on the handle's device, the call is
 *     asynchronous on the handle's stream; == 0: host pointers, the call copies in/out and synchronises.
 *   - host arrays stay owned by the caller; the library copies what it keeps.
 *   - a handle is not thread-safe; different handles are independent (no global mutable state).
 *   - there is no CPU fallback: without a usable CUDA device jfem_create fails with JFEM_ENODEV.
 */
#ifndef JFEM_B200_H
#define JFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JFEM_ABI_VERSION 2

/* status codes */
#define JFEM_OK 0
#define JFEM_EINVAL 1   /* bad argument (unsupported element type, id out of range, ...) */
#define JFEM_ENODEV 2   /* no CUDA device / device not sm_100 capable */
#define JFEM_ECUDA 3    /* CUDA runtime error */
#define JFEM_ESTATE 4   /* call out of order (e.g. matvec before material is set) */
#define JFEM_EDOMAIN 5  /* invalid deformation, J <= 0 (DomainError, src/materials/neo_hookean.jl:137) */
#define JFEM_ENCCL 6    /* NCCL error */

/* element types = nodes per element (Tet4, Hex8, Tet10; src/topology/tetrahedra.jl:73-105, hexahedra.jl:14-18) */
#define JFEM_TET4 4
#define JFEM_HEX8 8
#define JFEM_TET10 10

/* material kinds (src/materials/linear_elastic.jl, neo_hookean.jl, perfect_plasticity.jl) */
#define JFEM_MAT_LINEAR_ELASTIC 0 /* params: E, nu */
#define JFEM_MAT_NEO_HOOKEAN 1    /* params: E, nu  (mu, lambda derived as neo_hookean.jl:87-100); implies finite strain */
#define JFEM_MAT_PERFECT_PLASTICITY 2 /* params: E, nu, sigma_y, H ; 13 state doubles per Gauss point */
#define JFEM_MAT_STVK 3           /* params: E, nu.  St. Venant-Kirchhoff: Hooke's D on the Green-Lagrange strain = the classic
                                     Problem(Elasticity) path with props.finite_strain = true (pe:255-332); the geometric
                                     stiffness Kg (pe:378-404) enters the tangent only with option "geometric_stiffness" = 1 */

/* surface element types of jfem_surface_load = nodes per face (Tri3, Quad4, Tri6; the Elasticity3DSurfaceElements of pe:454-458) */
#define JFEM_TRI3 3
#define JFEM_QUAD4 4
#define JFEM_TRI6 6

/* jfem_nodal_recover fields (pe:520-545) */
#define JFEM_FIELD_STRAIN 0 /* 11, 22, 33, 12, 23, 13 (tensor shear components) */
#define JFEM_FIELD_STRESS 1 /* linear-elastic stress of the small strain, same order */

/* jfem_matvec / jfem_cg flags */
#define JFEM_PROJECT 1      /* zero Dirichlet rows of the result (apply_dirichlet_kernel!, ext:423-435) */
#define JFEM_TANGENT 2      /* operator = tangent K(u_lin) incl. geometric stiffness, not the linear-elastic K */
#define JFEM_USE_CSR 4      /* CG / matvec through the assembled CSR matrix (cpu:221-254) instead of matrix-free */
#define JFEM_JACOBI 8       /* opt-in 3x3 block-Jacobi preconditioned CG (not in the reference, which lists it as the first missing piece) */

typedef struct jfem_handle jfem_handle;

typedef struct jfem_info {
    int32_t abi_version, device, elem_type, n_ranks;
    int64_t n_nodes, n_elems, n_dofs, n_fixed;
    int64_t n_patches, n_interface_nodes, n_affine_elems;
    int64_t patch_elems, patch_max_nodes;
    int64_t device_bytes;        /* device memory held by the handle */
    int64_t matvec_launches;     /* kernel launches issued by the last jfem_matvec */
    int64_t total_launches;      /* kernel launches issued since creation */
    int64_t smem_bytes;          /* dynamic shared memory per block of the last patch kernel */
    int64_t blocks_per_sm;       /* resident blocks per SM of the last patch kernel */
    double setup_seconds;        /* host time spent building patches */
} jfem_info;

/* --- library ------------------------------------------------------------------------------ */
int jfem_abi_version(void);
const char *jfem_last_error(void);                 /* thread-local message of the last failing call */
int jfem_device_count(int *count);

/* --- problem setup: replaces initialize_gpu_data! (ext:86-217) ------------------------------ */
/* coords: 3 x n_nodes column-major (x,y,z of node 1, then node 2, ...) as ext:114-123;
 * conn: nnpe x n_elems column-major int32 (nodes of element 1, then element 2, ...) as ext:126-132. */
int jfem_create(jfem_handle **out, int device, int elem_type, int64_t n_nodes, int64_t n_elems,
                const double *coords, const int32_t *conn, int index_base);
int jfem_destroy(jfem_handle *h);
/* tuning knobs, before the first operator call: "patch_elems" (elements per thread-block patch),
 * "deterministic" (1: ordered interface reduction [default], 0: fp64 atomics), "affine_fast_path" (1/0),
 * "warp_specialised" (1/0), "lane_window" (candidates per lane of the bank-aware lane assignment, 0 = off),
 * "fused_halo" (1: halo exchange inside the patch kernel when peer mappings exist [default]),
 * "fused_interface" (1: cooperative-launch in-kernel interface reduction; default 0, measured slower),
 * "geometric_stiffness" (props.geometric_stiffness of pe:36-46 for JFEM_MAT_STVK; default 0),
 * "assembly_kernel" (1: one warp per element with the Gauss-point geometry shared by the columns [default], 0: one thread per
 * (element, column)),
 * profiling aids "debug_timing", "debug_skip" */
int jfem_set_option(jfem_handle *h, const char *key, double value);
/* homogeneous material (per_element == 0: params has n_params entries, 2 <= n_params <= 4) or per-element
 * (per_element != 0: params is n_params x n_elems column-major), like E_vec/nu_vec of ext:135-141 */
int jfem_set_material(jfem_handle *h, int kind, const double *params, int n_params, int per_element);
/* is_fixed / prescribed of ext:144-158.  dofs use index_base given at creation. */
int jfem_set_dirichlet(jfem_handle *h, const int64_t *dofs, const double *values, int64_t n);
int jfem_get_info(jfem_handle *h, jfem_info *info);
/* run all later work of this handle on an existing CUDA stream (cudaStream_t passed as void*) */
int jfem_set_stream(jfem_handle *h, void *cuda_stream);
int jfem_synchronize(jfem_handle *h);

/* --- matrix-free operator: replaces stiffness_operator_gpu / tangent_operator_gpu (ext:488-511),
 *     matrix_vector_product (eas:307-309).  y = K x (pure K.v: fixes the f_ext defect at ext:468). */
int jfem_matvec(jfem_handle *h, const double *x, double *y, int flags, int on_device);
/* Partitioned handles (n_ranks > 1): x and y are LOCAL vectors (owned entries first, then ghosts).  The forward halo of x happens
 * inside the call; the GHOST entries of a device-resident x are scratch: the NCCL and the unfused peer paths overwrite them with
 * the neighbours' values, the fused path reads the landing buffer and leaves them as they are.  Only owned entries of y are valid. */
/* r = f_int(u) (compute_residual_gpu! ext:442-478 without the f_ext subtraction); for plasticity the
 * trial state is kept aside until jfem_commit_state (src/materials/abstract_material.jl:203-207). */
int jfem_internal_force(jfem_handle *h, const double *u, double *f_int, int flags, int on_device);
int jfem_set_linearization(jfem_handle *h, const double *u, int on_device); /* u_current of ext:531-537 */
int jfem_commit_state(jfem_handle *h);
int jfem_get_state(jfem_handle *h, double *state /* 13 x ngp x n_elems */, int committed);
int jfem_set_state(jfem_handle *h, const double *state);

/* --- element matrices and assembled operator: replaces assemble_element! (pe:203-451),
 *     add!(COO) + sparse() (sp:53-55,121-132) ------------------------------------------------- */
/* Ke: ndof x ndof column-major per element, fe: ndof per element, elements [e0, e0+ne) in caller order (0-based e0) */
int jfem_element_matrices(jfem_handle *h, const double *u /* may be NULL */, int64_t e0, int64_t ne, double *Ke, double *fe);
int jfem_csr_size(jfem_handle *h, int64_t *n_rows, int64_t *nnz);
/* rowptr (n_rows+1) and colind (nnz), index_base-based like the reference's colptr/rowval; host pointers */
int jfem_csr_pattern(jfem_handle *h, int64_t *rowptr, int32_t *colind);
/* assemble K(u) (+Kg for finite strain) into the handle's device CSR; vals/f_int (host or device, may be NULL) receive copies;
 * symmetrise != 0 applies K <- (K+K')/2 (src/solvers.jl:289-292) */
int jfem_assemble_csr(jfem_handle *h, const double *u, double *vals, double *f_int, int symmetrise, int on_device);
int jfem_spmv(jfem_handle *h, const double *x, double *y, int flags, int on_device);

/* Penalty Dirichlet conditions on the assembled operator, the CPU backend's apply_dirichlet_bc! (eas:237-252): for every
 * fixed dof i  K[i,i] += penalty, r[i] = penalty * prescribed[i]  with penalty = scale * max|K| (the reference uses
 * scale = 1e10).  Works on the handle's device CSR (after jfem_assemble_csr); rhs: n_dofs vector, modified in place
 * (may be NULL); *penalty_out returns the value used. */
int jfem_csr_penalty_bc(jfem_handle *h, double scale, double *rhs, double *penalty_out, int on_device);

/* --- external loads: replaces the f_ext part of assemble_element! (pe:412-426) and the surface assembly
 *     assemble!(..., Elasticity3DSurfaceElements) (pe:454-502; GPU precedent apply_surface_traction_kernel!, ext:368-416).
 *     Consistent integration with the reference's default rules (GLTET4 / GLHEX8 / GLTET1; GLTRI1 / GLTRI3 / GLQUAD4),
 *     one thread per node summing its incident elements / faces in ascending order: deterministic, no atomics. */
/* f (+)= sum_e sum_gp w detJ N b.  b: 3 values ("displacement load 1..3"), or 3 x n_elems column-major if per_element */
int jfem_body_load(jfem_handle *h, const double *b, int per_element, int accumulate, double *f, int on_device);
/* faces: face_type x n_faces column-major node ids of the VOLUME mesh (index_base-based), reference node order.
 * traction: 3 x n_faces ("displacement traction force", may be NULL); pressure: n_faces ("surface pressure": positive
 * pressure acts against the face normal n = dX/dxi1 x dX/dxi2, pe:485-491; may be NULL).  host arrays. */
int jfem_surface_load(jfem_handle *h, int face_type, int64_t n_faces, const int32_t *faces, const double *traction,
                      const double *pressure, int accumulate, double *f, int on_device);
/* reaction forces of the constrained dofs, la = f_int(u) - f_ext there and 0 elsewhere: what solve! returns as the
 * Lagrange multipliers of the eliminated boundary rows (src/solvers.jl:205-216) */
int jfem_reactions(jfem_handle *h, const double *u, const double *f_ext, double *la, int on_device);

/* --- post-processing: replaces postprocess!(problem, time, Val{:strain|:stress}) -> lsq_fit (pe:547-594):
 *     least-squares fit of the Gauss-point field to the nodes, M x = b with M = sum w detJ N N', b_i = sum w detJ f_i N,
 *     M assembled on the node adjacency pattern (coloured, deterministic) and solved by Jacobi-PCG to 1e-14 (the reference
 *     factorises M).  out: 6 x n_nodes column-major. */
int jfem_nodal_recover(jfem_handle *h, const double *u, int field, double *out, int on_device);

/* --- solvers: replaces cg_solve_matfree_gpu! (ext:531-577) / cg_solve (cpu:221-254) and
 *     solve_newton_krylov_gpu! (ext:685-854) --------------------------------------------------- */
/* x: initial guess in, solution out.  Stop: sqrt(r.r) < tol (absolute, as the reference) or, if
 * tol_is_relative, sqrt(r.r) <= tol*||b|| (projected b).  Fixed dofs of r and A.p are zeroed each iteration. */
int jfem_cg(jfem_handle *h, const double *b, double *x, double tol, int tol_is_relative, int max_iter, int flags,
            int *iters, double *resid, int on_device);
/* history: 3 doubles per Newton step (cg_iters, ||R||, eta) like ElasticitySolution.history
 * (src/backend/abstract.jl:138-145); eta = min(forcing_max, ||R||^forcing_power) (ext:819-820). */
int jfem_newton_krylov(jfem_handle *h, const double *f_ext, double *u, double newton_tol, int max_newton,
                       int max_cg_per_newton, double forcing_power, double forcing_max, int flags,
                       int *newton_iters, int *cg_iters, double *resid, double *history, int history_cap, int on_device);

/* --- multi-GPU (one process per GPU): replaces partition/halo of benchmarks/multigpu_mpi_benchmark.jl:120-360
 *     and the MPI.Allreduce dots of demos/krylov_mpi_gpu_demo.jl:231-277 ------------------------- */
int jfem_comm_unique_id(char *id128);  /* 128 bytes, to be broadcast by the host launcher */
/* the handle was created on the LOCAL mesh (owned nodes first, then ghosts); n_owned nodes are owned. */
int jfem_comm_init(jfem_handle *h, int n_ranks, int rank, const char *id128, int64_t n_owned_nodes);
/* per neighbour: local node ids (index_base-based) to send / to receive into, both ascending global id */
int jfem_comm_set_halo(jfem_handle *h, int n_neighbours, const int32_t *neighbour_rank,
                       const int64_t *send_ptr, const int32_t *send_nodes,
                       const int64_t *recv_ptr, const int32_t *recv_nodes);
/* optional: replace the NCCL send/recv halo by direct peer-memory stores over NVLink (CUDA IPC, same node).
 * export: returns 128 bytes (two cudaIpcMemHandle_t: landing buffer, flags); the host all-gathers them.
 * import: all_handles = n_ranks x 128 bytes; recv_offsets[r*n_ranks+s] = node offset of rank s's segment inside rank r's
 * landing buffer (= recv_ptr of r for neighbour s, -1 if none); halves[r] = landing-buffer half size of rank r in doubles
 * (3*total_recv_nodes+8). */
int jfem_comm_p2p_export(jfem_handle *h, char *handles128);
int jfem_comm_p2p_import(jfem_handle *h, const char *all_handles, const int64_t *recv_offsets, const int64_t *halves);
/* Halo sequence number of the peer-memory exchange (it must agree on all ranks; every exchange advances it by one).
 * *seq returns the current value; set_to >= 0 sets it first.  Only needed to re-synchronise ranks after some of them
 * abandoned enqueued work (e.g. a failed CUDA-graph capture); no counterpart in the reference (its MPI halo is stateless,
 * benchmarks/multigpu_mpi_benchmark.jl:302-360). */
int jfem_comm_p2p_seq(jfem_handle *h, int64_t set_to, int64_t *seq);
int jfem_comm_destroy(jfem_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* JFEM_B200_H */
