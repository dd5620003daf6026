// handle.h -- the opaque jfem_handle and the internal entry points shared by the .cu files.
#pragma once
#include <unordered_map>

#include "common.h"

struct ncclComm;

struct CGScalars {        // lives on the device; read by every CG kernel
    double rr, pAp, alpha, beta, rr_new, bnorm2, thr, rz, rz_new;
    int iters, done, max_iter, rel;
    int pcg, pad_[3];
    unsigned int ticket[4];
};

struct jfem_handle {
    int device = 0, n_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int index_base = 1;
    MeshHost mesh;
    // options
    int patch_elems = 256;
    bool deterministic = true, affine = true, warp_specialised = true;
    bool asm_warp = true;       // assembled path: one warp per element (shared geometry); false = one thread per (element, column)
    bool geometric_stiffness = false;   // option "geometric_stiffness": Kg in the St. Venant-Kirchhoff tangent (pe:378-404)
    int debug_skip = 0;                 // profiling aid (option "debug_skip"): phases of the ws kernel to leave out
    int lane_window = 48;               // candidates examined per lane by the bank-aware lane assignment (0 = off)
    // material
    int mat_kind = -1;
    double mat[4] = {0, 0, 0, 0};
    std::vector<double> mat_per_elem;   // n_params x n_elems (caller order), empty when homogeneous
    int mat_nparams = 0;
    DevBuf<double> matp;                // 4 x n_elems SoA in internal element order
    // patches
    bool built = false;
    double setup_seconds = 0;
    double pattern_seconds = 0;  // host + device time of csr_build (adjacency, colouring, pattern expansion)
    PatchSetHost hsets[N_CLASSES];
    InterfaceHost hif;
    PatchSetDev dsets[N_CLASSES];
    DevBuf<uint32_t> inodes;            // interface node ids, then the nodes no element touches
    DevBuf<uint32_t> slot_node;         // partial slot -> node id
    DevBuf<int32_t> ibase;              // first partial slot of each of those (+ end sentinel)
    DevBuf<double> ipart;               // interface partials: 3 doubles per (interface node, touching patch)
    DevBuf<unsigned int> gbar;          // grid-barrier arrival counter of the fused interface reduction
    unsigned int gbar_count = 0;        // host mirror: counter value after the last launch
    bool fused_iface = false, coop_ok = false;   // fused (cooperative-launch tail) variant measured slower than the separate kernel
    DevBuf<double> coords;
    DevBuf<uint8_t> fixed;              // per dof
    DevBuf<double> prescribed;          // per dof (values of fixed dofs)
    int64_t n_fixed = 0;
    // nonlinear state
    DevBuf<double> ulin;                // linearisation point
    bool has_lin = false;
    DevBuf<double> st_old, st_new;      // 13 x n_gp SoA (internal element order)
    DevBuf<long long> timing;           // debug phase timing (option "debug_timing")
    DevBuf<int> dflags;                 // [0] = fail flag (invalid deformation)
    // work vectors
    DevBuf<double> wx, wy;              // staging for host-pointer calls
    DevBuf<double> xal;                 // 16-byte aligned copy of an operand that arrived misaligned (128-bit gathers)
    DevBuf<double> cg_r, cg_p, cg_Ap, cg_z, cg_dinv, nk_R, nk_du, nk_f;
    DevBuf<double> red_partials;
    DevBuf<CGScalars> cg_s;
    // CSR
    bool csr_built = false;
    std::vector<int64_t> h_nadj_ptr;    // node adjacency (host)
    std::vector<int32_t> h_nadj;
    DevBuf<int64_t> nadj_ptr, rowptr;
    DevBuf<int32_t> nadj, colind;
    DevBuf<double> vals;
    DevBuf<uint16_t> eblk;              // per element nnpe x nnpe : position of node l in adjacency of node k
    DevBuf<int32_t> dconn;              // device connectivity, caller order, 0-based
    DevBuf<int32_t> colour_elems;       // elements sorted by colour
    std::vector<int64_t> colour_ptr;
    DevBuf<int64_t> e2i;                // caller element -> internal element index (state lookup)
    DevBuf<long long> n2e_ptr, n2e_inc; // node -> (element * nnpe + local node) incidences, ascending (loads.cu)
    bool vals_valid = false;
    // comm
    ncclComm *comm = nullptr;
    int n_ranks = 1, rank = 0;
    int64_t n_owned_nodes = -1;
    std::vector<int> nb_rank;
    std::vector<int64_t> send_ptr, recv_ptr;
    DevBuf<int32_t> send_nodes, recv_nodes;
    DevBuf<double> send_buf, recv_buf;
    // peer-to-peer halo (CUDA IPC)
    bool p2p_ready = false;
    unsigned long long p2p_seq = 0;
    size_t p2p_half = 0;
    DevBuf<double> p2p_land;
    DevBuf<unsigned long long> p2p_flags;
    DevBuf<unsigned int> p2p_ticket;
    std::vector<void *> p2p_opened;     // every pointer returned by cudaIpcOpenMemHandle (closed by jfem_comm_destroy)
    std::vector<double *> p2p_peer_land;
    std::vector<unsigned long long *> p2p_peer_flag, p2p_ctrl;
    unsigned long long p2p_ar_seq = 0;
    bool fused_halo = true;             // option "fused_halo": exchange inside the patch kernel when the ws kernel runs
    bool recv_contiguous = false;       // ghost node g receives into landing slot g - n_owned
    bool halo_armed = false;            // halo_exchange() deferred the exchange to the next op_apply on halo_x
    const double *halo_x = nullptr;
    bool halo_in_kernel = false;        // the last exchange ran inside the patch kernel (statistics)
    std::vector<int64_t> p2p_peer_off;
    std::vector<size_t> p2p_peer_half;
    std::unordered_map<const void *, size_t> smem_attr;   // kernels already opted into this much dynamic shared memory (on this device)
    // stats
    int64_t matvec_launches = 0, total_launches = 0, last_smem = 0, last_blocks_per_sm = 0;

    int64_t n_dofs() const { return 3 * mesh.n_nodes; }
    int64_t n_owned_dofs() const { return 3 * (n_owned_nodes >= 0 ? n_owned_nodes : mesh.n_nodes); }
    int ngp() const { return mesh.nnpe == 10 ? 4 : (mesh.nnpe == 8 ? 8 : 1); }
};


// Arguments of the halo exchange fused into the patch kernel (peer-memory stores over NVLink, see comm.cu / matvec.cu)
struct HaloFused {
    int n_nb;
    double *peer_land[8];               // neighbour's landing buffer (mapped), offset to MY segment and this exchange's half
    unsigned long long *peer_flag[8];   // neighbour's flag word for me
    const unsigned long long *my_flag[8];
    long long send_off[9];              // node offsets of the per-neighbour send segments
    const int32_t *send_nodes;
    const double *land;                 // my landing half of this exchange: 3 doubles per ghost node, in ghost order
    unsigned long long seq;
    unsigned int *ticket;
};

int ensure_built(jfem_handle *h);
bool ws_halo_capable(jfem_handle *h);
void halo_fill_fused(jfem_handle *h, HaloFused &f);
int op_apply(jfem_handle *h, int mode, const double *x_dev, double *y_dev, int flags, const int *done_flag);
// fused_ok: the caller applies the patch operator to x next, so the exchange may be carried by that kernel
int halo_exchange(jfem_handle *h, double *x_dev, bool fused_ok = false);
int comm_allreduce_sum(jfem_handle *h, double *buf_dev, int count);
int upload_fixed(jfem_handle *h);
int upload_material(jfem_handle *h);

int cg_solve(jfem_handle *h, const double *b_dev, double *x_dev, double tol, int rel, int max_iter, int flags, int *iters, double *resid);
int newton_krylov(jfem_handle *h, const double *fext_dev, double *u_dev, double newton_tol, int max_newton, int max_cg,
                  double forcing_power, double forcing_max, int flags, int *newton_iters, int *cg_iters, double *resid,
                  double *history, int history_cap);
int vec_dot(jfem_handle *h, const double *a, const double *b, double *out_host);

int ensure_colouring(jfem_handle *h);
int csr_build(jfem_handle *h);
int csr_assemble(jfem_handle *h, const double *u_dev, int symmetrise);
int csr_fint(jfem_handle *h, const double *u_dev, double *f_dev);
int csr_spmv(jfem_handle *h, const double *x_dev, double *y_dev, int flags, const int *done_flag);
int element_matrices(jfem_handle *h, const double *u_dev, int64_t e0, int64_t ne, double *Ke_dev, double *fe_dev);
int jacobi_build(jfem_handle *h, int flags);
