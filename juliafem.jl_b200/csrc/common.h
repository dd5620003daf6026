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

// ---- patch node word: [31] interface  [30:28] fixed-dof mask (x,y,z)  [26:0] node id (interior) / partial slot (interface)
#define PN_ID_MASK 0x07FFFFFFu
#define PN_FIXSHIFT 28
#define PN_IFACE (1u << 31)
#define JFEM_MAX_NODES (1 << 27)

enum { CLASS_GENERAL = 0, CLASS_AFFINE = 1, N_CLASSES = 2 };
// modes of the element operator
enum { OP_LINEAR = 0, OP_RESIDUAL = 1, OP_TANGENT = 2 };

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

// Per-patch metadata blob (what the kernels stream into shared memory by TMA bulk copies).  Three parts with different
// lifetimes inside the kernel, each starting with the same 16-byte header {np, ncx | ncX << 16, ne, nrows | ghost flag << 16}:
//   part A  gather lists      cx u32[max_ncx]    16-byte CHUNKS of the dof vector the patch reads, ascending: chunk c = doubles
//                                                2c, 2c+1 of x (node n occupies doubles 3n..3n+2, i.e. chunks floor(3n/2) and
//                                                floor(3n/2)+1; consecutive node ids share chunks).  One 128-bit asynchronous
//                                                copy (LDGSTS.128) per chunk lands them back to back in the x tile.  (Runs of
//                                                consecutive chunks as TMA bulk copies were measured slower: a structured
//                                                mesh has ~100 runs of ~270 bytes per patch, and the bulk copies issue at
//                                                ~40 cycles each.)
//                             cX u32[max_ncX]    same for the coordinate array, only the nodes whose coordinates are needed
//                                                (affine Tet10: the element vertices); empty when every node needs them:
//                                                cx is used and the coordinate tile mirrors the x tile
//   part B  element table     et u32[(nnpe + nxr) * EP]   row k < nnpe, lane t:  offset (in doubles) of node k in the x tile |
//                                                         staging entry << 16;  rows nnpe.. (affine elements only): offsets of
//                                                         the 4 geometry nodes in the coordinate tile, two u16 per word
//   part C  reduce tables     qn u32[max_nodes]  node words in REDUCE order q (descending contribution count): interface
//                                                flag, fixed-dof mask and the node id (interior node: the sum goes
//                                                to y) or the partial slot of this (patch, node) (interface node)
//                             ql u8 [max_nodes]  contribution count of the node
//                             jo u16[max_rows+1] row offsets of the jagged staging tile
// Staging tile (shared memory, AoS, 3 doubles per entry): row r holds the r-th contribution of every node that has more
// than r contributions, nodes in q order -> entry(r, q) = jo[r] + q.  The reduction reads it with consecutive lanes on
// consecutive words; the element threads scatter into it through the precomputed entry index.
struct PatchLayout {
    int offA = 0, off_cx = 0, off_cX = 0;
    int offB = 0, off_et = 0;
    int offC = 0, off_qn = 0, off_ql = 0, off_jo = 0;
    int stride = 0;
    int bytesA() const { return offB - offA; }
    int bytesB() const { return offC - offB; }
    int bytesC() const { return stride - offC; }
};

// Host-side description of one homogeneous set of patches (all elements of one class).
struct PatchSetHost {
    int cls = CLASS_GENERAL;
    int nnpe = 0, EP = 0;              // nodes per element, elements per patch (= element threads per block)
    int nxr = 0;                       // extra element-table rows holding coordinate slots (2 for affine Tet10, else 0)
    int64_t n_elems = 0;               // elements in this set
    int n_patches = 0, max_nodes = 0, max_ncx = 0, max_ncX = 0, max_rows = 0, max_entries = 0;   // max_nc*: chunks per patch
    std::vector<int64_t> elem_perm;    // internal order -> caller element index (0-based); lane t of patch p = elem_perm[p*EP+t]
    std::vector<int32_t> pnode_ptr;    // n_patches+1
    std::vector<uint32_t> qnodes;      // node words in reduce order (kept to re-embed the Dirichlet mask)
    std::vector<int32_t> qids;         // node id of every entry of qnodes
    std::vector<uint8_t> ghosty;       // per patch: reads ghost values (nodes >= n_owned of a partitioned mesh)
    PatchLayout L;
    std::vector<uint8_t> blob;
    // layout statistics (modelled shared-memory wavefronts of the element threads per patch, before / after lane assignment)
    double wf_before = 0, wf_after = 0, wf_ideal = 0;
};

struct PatchSetDev {
    int cls = 0, nnpe = 0, EP = 0, nxr = 0, n_patches = 0, max_nodes = 0, max_ncx = 0, max_ncX = 0, max_rows = 0, max_entries = 0;
    int64_t n_elems = 0, elem_offset = 0;  // offset of this set in the internal element order
    PatchLayout L;
    DevBuf<uint8_t> blob;
    size_t bytes() const { return blob.bytes(); }
    void release() { blob.release(); }
};

struct InterfaceHost {
    std::vector<uint32_t> inodes;   // ids of interface nodes (touched by more than one patch), ascending
    std::vector<int32_t> ibase;     // first partial slot of each interface node (its slots are contiguous, ascending (set, patch))
    std::vector<uint32_t> orphans;  // nodes no element touches (y = 0)
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

int build_patch_sets(const MeshHost &m, int EP, bool use_affine, int lane_window, int64_t n_owned, PatchSetHost sets[N_CLASSES], InterfaceHost &iface);
// Node adjacency of the mesh = block pattern of the reference's sparse(I, J, V) (src/sparse/sparse.jl:121-132): adj[ap[a] .. ap[a+1])
// are the nodes sharing an element with node a, ascending; eblk[(e*nnpe+k)*nnpe+l] = position of node l of element e in the
// row of its node k.  Returns JFEM_EINVAL if a node has more than 65535 neighbours.
int build_node_adjacency(const MeshHost &m, std::vector<int64_t> &ap, std::vector<int32_t> &adj, std::vector<uint16_t> &eblk);
// Greedy element colouring in ascending element id (src/preprocess.jl:331-398): elements of one colour share no node.
// colour_ptr has n_colours+1 entries, celems lists the elements colour by colour (ascending id inside a colour).
void greedy_colouring(const MeshHost &m, std::vector<int64_t> &colour_ptr, std::vector<int32_t> &celems);
void classify_elements(MeshHost &m, bool use_affine);
