// abi.cu -- the extern "C" surface declared in include/jfem_b200.h.
#include <stdarg.h>
#include <string.h>

#include "handle.h"

static thread_local char g_err[1024] = "";

void jfem_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

#define CHECK_H(h)                                             \
    do {                                                       \
        if (!(h)) { jfem_set_error("null handle"); return JFEM_EINVAL; } \
        JFEM_CUDA(cudaSetDevice((h)->device));                 \
    } while (0)

// stage a host vector on the device (or pass a device pointer through)
static int in_vec(jfem_handle *h, const double *p, int on_device, DevBuf<double> &stage, const double **out) {
    if (!p) { *out = nullptr; return JFEM_OK; }
    if (on_device) { *out = p; return JFEM_OK; }
    size_t n = (size_t)h->n_dofs();
    if (stage.n != n) JFEM_TRY(stage.alloc(n));
    JFEM_CUDA(cudaMemcpyAsync(stage.p, p, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    *out = stage.p;
    return JFEM_OK;
}
static int out_vec_begin(jfem_handle *h, double *p, int on_device, DevBuf<double> &stage, double **out) {
    if (on_device) { *out = p; return JFEM_OK; }
    size_t n = (size_t)h->n_dofs();
    if (stage.n != n) JFEM_TRY(stage.alloc(n));
    *out = stage.p;
    return JFEM_OK;
}
static int out_vec_end(jfem_handle *h, double *p, int on_device, const double *dev) {
    if (on_device) return JFEM_OK;
    JFEM_CUDA(cudaMemcpyAsync(p, dev, (size_t)h->n_dofs() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    return JFEM_OK;
}

extern "C" {

int jfem_abi_version(void) { return JFEM_ABI_VERSION; }
const char *jfem_last_error(void) { return g_err; }

int jfem_device_count(int *count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; jfem_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); cudaGetLastError(); return JFEM_ENODEV; }
    *count = n;
    return JFEM_OK;
}

int jfem_create(jfem_handle **out, int device, int elem_type, int64_t n_nodes, int64_t n_elems, const double *coords,
                const int32_t *conn, int index_base) {
    if (!out) { jfem_set_error("null output pointer"); return JFEM_EINVAL; }
    *out = nullptr;
    if (elem_type != JFEM_TET4 && elem_type != JFEM_HEX8 && elem_type != JFEM_TET10) {
        // same refusal as the reference for non-volume elements in a 3D problem (src/problems_elasticity.jl:510-518)
        jfem_set_error("unsupported element type %d for 3D continuum elasticity (supported: Tet4=4, Hex8=8, Tet10=10)", elem_type);
        return JFEM_EINVAL;
    }
    if (n_nodes <= 0 || n_elems < 0 || !coords || (!conn && n_elems) || (index_base != 0 && index_base != 1)) {
        jfem_set_error("jfem_create: bad arguments"); return JFEM_EINVAL;
    }
    if (n_nodes >= JFEM_MAX_NODES) { jfem_set_error("at most %d nodes per handle", JFEM_MAX_NODES - 1); return JFEM_EINVAL; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        jfem_set_error("no CUDA device available (this library has no CPU fallback)");
        return JFEM_ENODEV;
    }
    if (device < 0 || device >= ndev) { jfem_set_error("device %d out of range (%d devices)", device, ndev); return JFEM_ENODEV; }
    cudaDeviceProp prop;
    JFEM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { jfem_set_error("device %d is sm_%d%d; this build targets sm_100a (B200) only", device, prop.major, prop.minor); return JFEM_ENODEV; }
    JFEM_CUDA(cudaSetDevice(device));
    jfem_handle *h = new jfem_handle();
    h->device = device; h->index_base = index_base; h->n_sms = prop.multiProcessorCount;
    h->mesh.nnpe = elem_type; h->mesh.n_nodes = n_nodes; h->mesh.n_elems = n_elems;
    h->mesh.coords.assign(coords, coords + 3 * n_nodes);
    h->mesh.conn.resize((size_t)n_elems * elem_type);
    for (int64_t i = 0; i < n_elems * elem_type; i++) {
        int64_t v = (int64_t)conn[i] - index_base;
        if (v < 0 || v >= n_nodes) { delete h; jfem_set_error("connectivity entry %lld out of range", (long long)i); return JFEM_EINVAL; }
        h->mesh.conn[i] = (int32_t)v;
    }
    h->mesh.fixed.assign(3 * n_nodes, 0);
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; jfem_set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); return JFEM_ECUDA; }
    h->own_stream = true;
    *out = h;
    return JFEM_OK;
}

int jfem_destroy(jfem_handle *h) {
    if (!h) return JFEM_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    jfem_comm_destroy(h);
    for (int c = 0; c < N_CLASSES; c++) h->dsets[c].release();
    h->matp.release(); h->inodes.release(); h->ibase.release(); h->ipart.release(); h->gbar.release(); h->slot_node.release(); h->coords.release(); h->fixed.release();
    h->prescribed.release(); h->ulin.release(); h->st_old.release(); h->st_new.release(); h->dflags.release(); h->wx.release(); h->wy.release();
    h->cg_r.release(); h->cg_p.release(); h->cg_Ap.release(); h->cg_z.release(); h->cg_dinv.release(); h->nk_R.release(); h->nk_du.release();
    h->nk_f.release(); h->red_partials.release(); h->cg_s.release(); h->nadj_ptr.release(); h->rowptr.release(); h->nadj.release();
    h->colind.release(); h->vals.release(); h->eblk.release(); h->dconn.release(); h->colour_elems.release(); h->e2i.release();
    h->send_nodes.release(); h->recv_nodes.release(); h->send_buf.release(); h->recv_buf.release();
    h->timing.release(); h->xal.release(); h->cg_s.release(); h->n2e_ptr.release(); h->n2e_inc.release();
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return JFEM_OK;
}

int jfem_set_option(jfem_handle *h, const char *key, double value) {
    CHECK_H(h);
    if (!key) { jfem_set_error("null option key"); return JFEM_EINVAL; }
    if (!strcmp(key, "patch_elems")) {
        int v = (int)value;
        if (v != 128 && v != 256 && v != 512) { jfem_set_error("patch_elems must be 128, 256 or 512"); return JFEM_EINVAL; }
        if (v != h->patch_elems) { h->patch_elems = v; h->built = false; }
    } else if (!strcmp(key, "debug_timing")) {
        if (value != 0) { JFEM_TRY(h->timing.alloc(128)); JFEM_CUDA(cudaMemset(h->timing.p, 0, 128 * sizeof(long long))); }
        else h->timing.release();
    } else if (!strcmp(key, "warp_specialised")) {
        h->warp_specialised = value != 0;
    } else if (!strcmp(key, "assembly_kernel")) {
        // 1 (default): one warp per element, geometry shared by the columns; 0: one thread per (element, column)
        h->asm_warp = value != 0;
    } else if (!strcmp(key, "async_gather")) {
        // accepted for compatibility: the gather is always asynchronous (LDGSTS.128)
    } else if (!strcmp(key, "fused_halo")) {
        h->fused_halo = value != 0;
    } else if (!strcmp(key, "fused_interface")) {
        h->fused_iface = value != 0;
    } else if (!strcmp(key, "geometric_stiffness")) {
        h->geometric_stiffness = value != 0;
    } else if (!strcmp(key, "debug_skip")) {
        h->debug_skip = (int)value;
    } else if (!strcmp(key, "lane_window")) {
        int v = (int)value;
        if (v < 0 || v > 4096) { jfem_set_error("lane_window must be in [0, 4096]"); return JFEM_EINVAL; }
        if (v != h->lane_window) { h->lane_window = v; h->built = false; }
    } else if (!strcmp(key, "deterministic")) {
        h->deterministic = value != 0;
    } else if (!strcmp(key, "affine_fast_path")) {
        // a request only: the closed forms exist for the linear-elastic operator alone, ensure_built() decides
        bool v = value != 0;
        if (v != h->affine) { h->affine = v; if (h->mat_kind == JFEM_MAT_LINEAR_ELASTIC || h->mat_kind < 0) h->built = false; }
    } else {
        jfem_set_error("unknown option '%s'", key);
        return JFEM_EINVAL;
    }
    return JFEM_OK;
}

int jfem_set_material(jfem_handle *h, int kind, const double *params, int n_params, int per_element) {
    CHECK_H(h);
    const int need = kind == JFEM_MAT_PERFECT_PLASTICITY ? 4 : 2;
    if (kind < 0 || kind > JFEM_MAT_STVK || !params || n_params < need || n_params > 4) { jfem_set_error("jfem_set_material: bad kind/params"); return JFEM_EINVAL; }
    const int64_t nsets = per_element ? h->mesh.n_elems : 1;
    for (int64_t e = 0; e < nsets; e++) {
        const double *q = params + e * n_params;
        const double E = q[0], nu = q[1];
        // parameter validation of the reference constructors (linear_elastic.jl:51-56, perfect_plasticity.jl:176-181)
        if (!(E > 0.0)) { jfem_set_error("Young's modulus must be positive, got E = %g", E); return JFEM_EINVAL; }
        if (!(nu > -1.0 && nu < 0.5)) { jfem_set_error("Poisson's ratio must be in (-1, 0.5), got nu = %g", nu); return JFEM_EINVAL; }
        if (kind == JFEM_MAT_PERFECT_PLASTICITY && !(q[2] > 0.0 && q[3] >= 0.0)) {
            jfem_set_error("yield stress must be positive and hardening modulus non-negative"); return JFEM_EINVAL;
        }
    }
    if (per_element) { h->mat_per_elem.assign(params, params + (size_t)n_params * h->mesh.n_elems); h->mat_nparams = n_params; }
    else { h->mat_per_elem.clear(); h->mat_nparams = 0; h->matp.release(); }
    // the affine closed forms are only valid for the linear-elastic operator: the element classes (and with them the patch
    // sets) depend on the material kind, see ensure_built()
    const bool affine_prev = h->affine && (h->mat_kind == JFEM_MAT_LINEAR_ELASTIC || h->mat_kind < 0);
    h->mat_kind = kind;
    for (int i = 0; i < 4; i++) h->mat[i] = i < n_params ? params[i] : 0.0;
    if ((h->affine && kind == JFEM_MAT_LINEAR_ELASTIC) != affine_prev) h->built = false;
    if (kind == JFEM_MAT_PERFECT_PLASTICITY && h->built && h->st_old.n == 0) h->built = false;
    if (h->built) return upload_material(h);
    return JFEM_OK;
}

int jfem_set_dirichlet(jfem_handle *h, const int64_t *dofs, const double *values, int64_t n) {
    CHECK_H(h);
    if (n < 0 || (n && !dofs)) { jfem_set_error("jfem_set_dirichlet: bad arguments"); return JFEM_EINVAL; }
    std::vector<double> pres(h->n_dofs(), 0.0);
    std::fill(h->mesh.fixed.begin(), h->mesh.fixed.end(), 0);
    for (int64_t i = 0; i < n; i++) {
        int64_t d = dofs[i] - h->index_base;
        if (d < 0 || d >= h->n_dofs()) { jfem_set_error("Dirichlet dof %lld out of range", (long long)dofs[i]); return JFEM_EINVAL; }
        h->mesh.fixed[d] = 1;
        pres[d] = values ? values[i] : 0.0;
    }
    h->n_fixed = 0;
    for (uint8_t f : h->mesh.fixed) h->n_fixed += f;
    JFEM_TRY(h->prescribed.upload(pres));
    return upload_fixed(h);
}

int jfem_get_info(jfem_handle *h, jfem_info *info) {
    CHECK_H(h);
    memset(info, 0, sizeof *info);
    info->abi_version = JFEM_ABI_VERSION; info->device = h->device; info->elem_type = h->mesh.nnpe; info->n_ranks = h->n_ranks;
    info->n_nodes = h->mesh.n_nodes; info->n_elems = h->mesh.n_elems; info->n_dofs = h->n_dofs(); info->n_fixed = h->n_fixed;
    info->patch_elems = h->patch_elems;
    if (h->built) {
        for (int c = 0; c < N_CLASSES; c++) {
            info->n_patches += h->dsets[c].n_patches;
            if (h->dsets[c].max_nodes > info->patch_max_nodes) info->patch_max_nodes = h->dsets[c].max_nodes;
            info->device_bytes += (int64_t)h->dsets[c].bytes();
        }
        info->n_affine_elems = h->dsets[CLASS_AFFINE].n_elems;
        info->n_interface_nodes = (int64_t)h->hif.inodes.size();
        info->device_bytes += (int64_t)(h->inodes.bytes() + h->ibase.bytes() + h->ipart.bytes() + h->coords.bytes() + h->fixed.bytes() +
                                        h->st_old.bytes() + h->st_new.bytes() + h->ulin.bytes() + h->cg_r.bytes() * 3 + h->rowptr.bytes() + h->colind.bytes() +
                                        h->vals.bytes() + h->eblk.bytes() + h->nadj.bytes() + h->nadj_ptr.bytes() + h->dconn.bytes() + h->e2i.bytes());
    }
    info->matvec_launches = h->matvec_launches; info->total_launches = h->total_launches; info->setup_seconds = h->setup_seconds;
    info->smem_bytes = h->last_smem; info->blocks_per_sm = h->last_blocks_per_sm;
    return JFEM_OK;
}

int jfem_set_stream(jfem_handle *h, void *cuda_stream) {
    CHECK_H(h);
    cudaStreamSynchronize(h->stream);
    if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
    h->stream = (cudaStream_t)cuda_stream;
    return JFEM_OK;
}

int jfem_synchronize(jfem_handle *h) {
    CHECK_H(h);
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    return JFEM_OK;
}

int jfem_debug_timing(jfem_handle *h, long long *out64) {   /* not part of the public header: debugging aid */
    CHECK_H(h);
    if (!h->timing.p) { jfem_set_error("debug_timing option is off"); return JFEM_ESTATE; }
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    JFEM_CUDA(cudaMemcpy(out64, h->timing.p, 128 * sizeof(long long), cudaMemcpyDeviceToHost));
    return JFEM_OK;
}

static int check_domain(jfem_handle *h, int on_device, const char *what) {
    if (on_device) return JFEM_OK;   // asynchronous call: the flag is checked by the next synchronous call
    int fail = 0;
    JFEM_CUDA(cudaMemcpyAsync(&fail, h->dflags.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    if (fail) {
        JFEM_CUDA(cudaMemset(h->dflags.p, 0, sizeof(int)));
        jfem_set_error("Jacobian J = sqrt(det(C)) must be positive (invalid deformation in %s)", what);
        return JFEM_EDOMAIN;
    }
    return JFEM_OK;
}

int jfem_matvec(jfem_handle *h, const double *x, double *y, int flags, int on_device) {
    CHECK_H(h);
    if (!x || !y) { jfem_set_error("jfem_matvec: null vector"); return JFEM_EINVAL; }
    JFEM_TRY(ensure_built(h));
    const double *dx; double *dy;
    JFEM_TRY(in_vec(h, x, on_device, h->wx, &dx));
    JFEM_TRY(out_vec_begin(h, y, on_device, h->wy, &dy));
    if (h->n_ranks > 1) JFEM_TRY(halo_exchange(h, (double *)dx, !(flags & JFEM_USE_CSR)));
    if (flags & JFEM_USE_CSR) JFEM_TRY(csr_spmv(h, dx, dy, flags, nullptr));
    else JFEM_TRY(op_apply(h, (flags & JFEM_TANGENT) ? OP_TANGENT : OP_LINEAR, dx, dy, flags, nullptr));
    JFEM_TRY(out_vec_end(h, y, on_device, dy));
    return check_domain(h, on_device, "matvec");
}

int jfem_internal_force(jfem_handle *h, const double *u, double *f, int flags, int on_device) {
    CHECK_H(h);
    if (!u || !f) { jfem_set_error("jfem_internal_force: null vector"); return JFEM_EINVAL; }
    JFEM_TRY(ensure_built(h));
    const double *du; double *df;
    JFEM_TRY(in_vec(h, u, on_device, h->wx, &du));
    JFEM_TRY(out_vec_begin(h, f, on_device, h->wy, &df));
    if (h->n_ranks > 1) JFEM_TRY(halo_exchange(h, (double *)du, true));
    JFEM_TRY(op_apply(h, OP_RESIDUAL, du, df, flags, nullptr));
    JFEM_TRY(out_vec_end(h, f, on_device, df));
    return check_domain(h, on_device, "internal force");
}

int jfem_set_linearization(jfem_handle *h, const double *u, int on_device) {
    CHECK_H(h);
    size_t n = (size_t)h->n_dofs();
    if (h->ulin.n != n) JFEM_TRY(h->ulin.alloc(n));
    JFEM_CUDA(cudaMemcpyAsync(h->ulin.p, u, n * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    if (!on_device) JFEM_CUDA(cudaStreamSynchronize(h->stream));
    h->has_lin = true;
    return JFEM_OK;
}

int jfem_commit_state(jfem_handle *h) {
    CHECK_H(h);
    if (h->st_old.n) JFEM_CUDA(cudaMemcpyAsync(h->st_old.p, h->st_new.p, h->st_old.bytes(), cudaMemcpyDeviceToDevice, h->stream));
    return JFEM_OK;
}

// state at the boundary: 13 x ngp x n_elems column-major in CALLER element order; device: SoA [13][n_gp] in internal order
int jfem_get_state(jfem_handle *h, double *state, int committed) {
    CHECK_H(h);
    JFEM_TRY(ensure_built(h));
    if (!h->st_old.n) { jfem_set_error("material has no integration-point state"); return JFEM_ESTATE; }
    const int ng = h->ngp();
    const int64_t n_gp = h->mesh.n_elems * ng;
    std::vector<double> tmp((size_t)JFEM_NSTATE * n_gp);
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    JFEM_CUDA(cudaMemcpy(tmp.data(), committed ? h->st_old.p : h->st_new.p, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    int64_t off = 0;
    for (int c = 0; c < N_CLASSES; c++) {
        const PatchSetHost &S = h->hsets[c];
        for (int64_t i = 0; i < S.n_elems; i++)
            for (int g = 0; g < ng; g++)
                for (int s = 0; s < JFEM_NSTATE; s++) state[(S.elem_perm[i] * ng + g) * JFEM_NSTATE + s] = tmp[(size_t)s * n_gp + (off + i) * ng + g];
        off += S.n_elems;
    }
    return JFEM_OK;
}

int jfem_set_state(jfem_handle *h, const double *state) {
    CHECK_H(h);
    JFEM_TRY(ensure_built(h));
    if (!h->st_old.n) { jfem_set_error("material has no integration-point state"); return JFEM_ESTATE; }
    const int ng = h->ngp();
    const int64_t n_gp = h->mesh.n_elems * ng;
    std::vector<double> tmp((size_t)JFEM_NSTATE * n_gp);
    int64_t off = 0;
    for (int c = 0; c < N_CLASSES; c++) {
        const PatchSetHost &S = h->hsets[c];
        for (int64_t i = 0; i < S.n_elems; i++)
            for (int g = 0; g < ng; g++)
                for (int s = 0; s < JFEM_NSTATE; s++) tmp[(size_t)s * n_gp + (off + i) * ng + g] = state[(S.elem_perm[i] * ng + g) * JFEM_NSTATE + s];
        off += S.n_elems;
    }
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    JFEM_CUDA(cudaMemcpy(h->st_old.p, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice));
    JFEM_CUDA(cudaMemcpy(h->st_new.p, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice));
    return JFEM_OK;
}

int jfem_element_matrices(jfem_handle *h, const double *u, int64_t e0, int64_t ne, double *Ke, double *fe) {
    CHECK_H(h);
    if (e0 < 0 || ne < 0 || e0 + ne > h->mesh.n_elems) { jfem_set_error("element range out of bounds"); return JFEM_EINVAL; }
    JFEM_TRY(ensure_built(h));
    const int nd = 3 * h->mesh.nnpe;
    const double *du;
    JFEM_TRY(in_vec(h, u, 0, h->wx, &du));
    DevBuf<double> dK, df;
    if (Ke) JFEM_TRY(dK.alloc((size_t)ne * nd * nd));
    if (fe) JFEM_TRY(df.alloc((size_t)ne * nd));
    int rc = element_matrices(h, du, e0, ne, dK.p, df.p);
    if (rc == JFEM_OK) {
        if (Ke && cudaMemcpy(Ke, dK.p, dK.bytes(), cudaMemcpyDeviceToHost) != cudaSuccess) rc = JFEM_ECUDA;
        if (fe && cudaMemcpy(fe, df.p, df.bytes(), cudaMemcpyDeviceToHost) != cudaSuccess) rc = JFEM_ECUDA;
    }
    dK.release(); df.release();
    return rc;
}

int jfem_csr_size(jfem_handle *h, int64_t *n_rows, int64_t *nnz) {
    CHECK_H(h);
    JFEM_TRY(csr_build(h));
    if (n_rows) *n_rows = h->n_dofs();
    if (nnz) *nnz = 9 * h->h_nadj_ptr[h->mesh.n_nodes];
    return JFEM_OK;
}

int jfem_csr_pattern(jfem_handle *h, int64_t *rowptr, int32_t *colind) {
    CHECK_H(h);
    JFEM_TRY(csr_build(h));
    const int64_t nr = h->n_dofs(), nnz = 9 * h->h_nadj_ptr[h->mesh.n_nodes];
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    if (rowptr) {
        JFEM_CUDA(cudaMemcpy(rowptr, h->rowptr.p, (nr + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
        if (h->index_base) for (int64_t i = 0; i <= nr; i++) rowptr[i] += h->index_base;
    }
    if (colind) {
        JFEM_CUDA(cudaMemcpy(colind, h->colind.p, nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (h->index_base) for (int64_t i = 0; i < nnz; i++) colind[i] += h->index_base;
    }
    return JFEM_OK;
}

int jfem_assemble_csr(jfem_handle *h, const double *u, double *vals, double *f_int, int symmetrise, int on_device) {
    CHECK_H(h);
    JFEM_TRY(csr_build(h));
    const double *du;
    JFEM_TRY(in_vec(h, u, on_device, h->wx, &du));
    JFEM_TRY(csr_assemble(h, du, symmetrise));
    if (vals) JFEM_CUDA(cudaMemcpyAsync(vals, h->vals.p, h->vals.bytes(), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    if (f_int) {
        double *df;
        JFEM_TRY(out_vec_begin(h, f_int, on_device, h->wy, &df));
        if (du) JFEM_TRY(op_apply(h, OP_RESIDUAL, du, df, 0, nullptr));
        else JFEM_CUDA(cudaMemsetAsync(df, 0, (size_t)h->n_dofs() * sizeof(double), h->stream));
        JFEM_TRY(out_vec_end(h, f_int, on_device, df));
    }
    if (!on_device) JFEM_CUDA(cudaStreamSynchronize(h->stream));
    return JFEM_OK;
}

int jfem_spmv(jfem_handle *h, const double *x, double *y, int flags, int on_device) {
    return jfem_matvec(h, x, y, flags | JFEM_USE_CSR, on_device);
}

int jfem_cg(jfem_handle *h, const double *b, double *x, double tol, int tol_is_relative, int max_iter, int flags, int *iters,
            double *resid, int on_device) {
    CHECK_H(h);
    if (!b || !x) { jfem_set_error("jfem_cg: null vector"); return JFEM_EINVAL; }
    JFEM_TRY(ensure_built(h));
    const double *db; double *dx;
    JFEM_TRY(in_vec(h, b, on_device, h->wy, &db));
    if (on_device) dx = x;
    else {
        const double *t;
        JFEM_TRY(in_vec(h, x, 0, h->wx, &t));
        dx = (double *)t;
    }
    JFEM_TRY(cg_solve(h, db, dx, tol, tol_is_relative, max_iter, flags, iters, resid));
    return out_vec_end(h, x, on_device, dx);
}

int jfem_newton_krylov(jfem_handle *h, const double *f_ext, double *u, double newton_tol, int max_newton, int max_cg_per_newton,
                       double forcing_power, double forcing_max, int flags, int *newton_iters, int *cg_iters, double *resid,
                       double *history, int history_cap, int on_device) {
    CHECK_H(h);
    if (!f_ext || !u) { jfem_set_error("jfem_newton_krylov: null vector"); return JFEM_EINVAL; }
    JFEM_TRY(ensure_built(h));
    size_t n = (size_t)h->n_dofs();
    if (h->ulin.n != n) JFEM_TRY(h->ulin.alloc(n));
    const double *df; double *du;
    JFEM_TRY(in_vec(h, f_ext, on_device, h->wy, &df));
    if (on_device) du = u;
    else {
        const double *t;
        JFEM_TRY(in_vec(h, u, 0, h->wx, &t));
        du = (double *)t;
    }
    JFEM_TRY(newton_krylov(h, df, du, newton_tol, max_newton, max_cg_per_newton, forcing_power, forcing_max, flags, newton_iters, cg_iters,
                           resid, history, history_cap));
    return out_vec_end(h, u, on_device, du);
}

}  // extern "C"
