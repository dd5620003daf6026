// common.h -- internal declarations of libjfem_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/jfem_b200.h"

#define JFEM_NSTATE 13  // eps_p(6) alpha(6) kappa  (PlasticityState, src/materials/perfect_plasticity.jl:107-127)

void jfem_set_error(const char *fmt, ...);

#define JFEM_CUDA(call)                                                                                   \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            jfem_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__,       \
                           cudaGetErrorString(e__));                                                      \
            return JFEM_ECUDA;                                                                            \
        }                                                                                                 \
    } while (0)

#define JFEM_TRY(call)                 \
    do {                               \
        int rc__ = (call);             \
        if (rc__ != JFEM_OK) return rc__; \
    } while (0)

// ---- patch node word: [31] interface  [30:28] fixed-dof mask (x,y,z)  [27] coords needed  [26:0] node id
#define PN_ID_MASK 0x07FFFFFFu
#define PN_NEEDX (1u << 27)
#define PN_FIXSHIFT 28
#define PN_IFACE (1u << 31)
#define JFEM_MAX_NODES (1 << 27)

enum { CLASS_GENERAL = 0, CLASS_AFFINE = 1, N_CLASSES = 2 };

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return JFEM_OK;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            jfem_set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
            p = nullptr; n = 0;
            return JFEM_ECUDA;
        }
        return JFEM_OK;
    }
    int upload(const std::vector<T> &v) {
        JFEM_TRY(alloc(v.size()));
        if (!v.empty()) JFEM_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        return JFEM_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr; n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// Host-side description of one homogeneous set of patches (all elements of one class).
struct PatchSetHost {
    int cls = CLASS_GENERAL;
    int nnpe = 0, EP = 0;              // nodes per element, elements per patch (= threads per block)
    int64_t n_elems = 0;               // elements in this set
    int n_patches = 0, max_nodes = 0;
    std::vector<int64_t> elem_perm;    // internal order -> caller element index (0-based)
    std::vector<int32_t> pnode_ptr;    // n_patches+1
    std::vector<uint32_t> pnodes;      // patch node words; interface nodes first
    std::vector<int32_t> n_iface;      // per patch
    std::vector<int32_t> ipart_base;   // per patch: first interface-partial slot (in nodes)
    std::vector<uint16_t> lconn;       // [(p*nnpe+k)*EP + t] local node index (0xFFFF = no element)
    std::vector<uint16_t> goff;        // per patch Np+1 offsets, at pnode_ptr[p]+p
    std::vector<uint16_t> gslots;      // per patch [k*EP + t]: position of element t's node k in the node-major staging tile
    std::vector<uint16_t> xslot;       // per patch node: slot in the compact coordinate tile (0xFFFF: coordinates not needed)
    int max_nx = 0;                    // max #nodes per patch whose coordinates are needed
    // everything above packed per patch into one 16-byte aligned blob (what the kernel streams in by TMA bulk copy):
    //   [0,16) header {np, n_iface | nx<<16, ipart_base, n_elems} | pnodes u32[max_nodes] | xlist u32[max_nx] (ids of the nodes
    //   whose coordinates are needed) | xslot u16[max_nodes] | goff u16[max_nodes+1] | rank u8[nnpe*EP] | lconn u16[nnpe*EP]
    std::vector<uint8_t> blob;
    int off_pn = 0, off_xl = 0, off_xs = 0, off_go = 0, off_gs = 0, off_lc = 0, stride = 0;
};

struct PatchSetDev {
    int cls = 0, nnpe = 0, EP = 0, n_patches = 0, max_nodes = 0;
    int64_t n_elems = 0, elem_offset = 0;  // offset of this set in the internal element order
    int max_nx = 0, off_pn = 0, off_xl = 0, off_xs = 0, off_go = 0, off_gs = 0, off_lc = 0, stride = 0;
    DevBuf<uint8_t> blob;
    size_t bytes() const { return blob.bytes(); }
    void release() { blob.release(); }
};

struct InterfaceHost {
    std::vector<uint32_t> inodes;   // node words (id + fixed mask) of interface nodes, ascending id
    std::vector<int32_t> iptr;      // n_inodes+1
    std::vector<int32_t> islots;    // partial slots (node units) in ascending (set, patch) order
    int64_t n_partials = 0;         // total interface incidences
};

// Builds patches for the element subset `elems` (caller indices) of class `cls`.
// touch[] (size n_nodes) must hold, for every node, the number of patches (over all sets) touching it;
// it is produced by build_patch_sets().
struct MeshHost {
    int nnpe = 0;
    int64_t n_nodes = 0, n_elems = 0;
    std::vector<double> coords;     // 3*n_nodes
    std::vector<int32_t> conn;      // nnpe*n_elems, 0-based
    std::vector<uint8_t> fixed;     // n_dofs, 1 = Dirichlet
    std::vector<uint8_t> cls;       // per element class
};

int build_patch_sets(const MeshHost &m, int EP, bool use_affine, PatchSetHost sets[N_CLASSES], InterfaceHost &iface);
void classify_elements(MeshHost &m, bool use_affine);
