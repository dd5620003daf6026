// Phase-1 occupancy micro-benchmark: the element contraction of the Tet10 patch kernel (tet10_affine_linear through the same
// shared-memory accessors as patch_elem.cuh) run ALONE -- no gather, no reduction -- with 8 / 12 / 16 element warps per SM,
// to see how far more resident element warps raise the fp64 pipe utilisation (floor: 513 fp64 instr x 2 cycles per warp-instr
// per SMSP).  Synthetic but representative tables: 600-node x tile, random node picks, conflict-free staging rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../juliafem.jl_b200/csrc -o phase1_occupancy phase1_occupancy.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "patch_elem.cuh"

using namespace jf;

template <int T, int STAGGER>
__global__ void __launch_bounds__(T, 1) k(const uint32_t *et_g, const double *x_g, const double *X_g, int n_nodes, int n_x, int iters, long long *cyc,
                                           double *sink) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint32_t *et = reinterpret_cast<uint32_t *>(sm);                 // 12 rows x T
    double *xs = reinterpret_cast<double *>(sm + 12 * T * 4);
    double *Xs = xs + 3 * n_nodes;
    double *stage = Xs + 3 * n_x;
    for (int i = threadIdx.x; i < 12 * T; i += T) et[i] = et_g[i];
    for (int i = threadIdx.x; i < 3 * n_nodes; i += T) xs[i] = x_g[i];
    for (int i = threadIdx.x; i < 3 * n_x; i += T) Xs[i] = X_g[i];
    __syncthreads();
    PtLinear pt;
    pt.la = 1.2e11; pt.mu = 8e10; pt.pe = nullptr; pt.pe_n = 0;
    if (STAGGER) {   // desynchronise the warps of a scheduler: warp w starts w/4 * STAGGER cycles late
        long long t = clock64() + (long long)(threadIdx.x >> 7) * STAGGER;
        while (clock64() < t) { }
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        element_phase<10, CLASS_AFFINE, OP_LINEAR, PtLinear, T>(pt, 0, et, threadIdx.x, xs, Xs, nullptr, stage);
        if (!STAGGER) __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * T + threadIdx.x] = stage[3 * threadIdx.x];
}

template <int T, int STAGGER>
void run() {
    const int n_nodes = 600 * T / 256, n_x = 140 * T / 256, iters = 64;
    std::vector<uint32_t> et(12 * T);
    srand(1);
    for (int t = 0; t < T; t++) {
        // nodes of neighbouring elements overlap (as in a real patch): pick from a window that moves with the lane
        int base = (int)((long long)t * (n_nodes - 40) / T);
        for (int k = 0; k < 10; k++) et[k * T + t] = (uint32_t)(base + rand() % 40) | ((uint32_t)(k * T + t) << 16);
        int xb = (int)((long long)t * (n_x - 12) / T);
        int v[4];
        for (int q = 0; q < 4; q++) v[q] = xb + 3 * q + rand() % 3;
        et[10 * T + t] = v[0] | (v[1] << 16);
        et[11 * T + t] = v[2] | (v[3] << 16);
    }
    std::vector<double> x(3 * n_nodes), X(3 * n_x);
    for (auto &v : x) v = 1e-3 * (rand() / (double)RAND_MAX - 0.5);
    for (int i = 0; i < n_x; i++) { X[3 * i] = 0.01 * i + 0.003 * (rand() / (double)RAND_MAX); X[3 * i + 1] = 0.02 * (i % 7) + 0.004 * (rand() / (double)RAND_MAX); X[3 * i + 2] = 0.015 * (i % 5) + 0.005 * (rand() / (double)RAND_MAX) + 0.001 * i * i; }
    uint32_t *d_et; double *d_x, *d_X, *sink; long long *cyc;
    cudaMalloc(&d_et, et.size() * 4); cudaMalloc(&d_x, x.size() * 8); cudaMalloc(&d_X, X.size() * 8);
    cudaMalloc(&sink, 148 * T * 8); cudaMalloc(&cyc, 148 * 8);
    cudaMemcpy(d_et, et.data(), et.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_x, x.data(), x.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_X, X.data(), X.size() * 8, cudaMemcpyHostToDevice);
    size_t smem = 12 * T * 4 + 8 * 3 * (size_t)(n_nodes + n_x) + 8 * 30 * (size_t)T;
    auto kern = k<T, STAGGER>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    for (int rep = 0; rep < 2; rep++) kern<<<148, T, smem>>>(d_et, d_x, d_X, n_nodes, n_x, iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = (double)h[0] / iters;
    // floor: T/32 warps x 513 fp64 instr x 2 cycles / 4 SMSPs
    double floor_c = (T / 32) * 513.0 * 2.0 / 4.0;
    printf("T=%3d (%2d warps, %d/SMSP) regs=%3d spill=%zuB stagger=%4d : %7.0f cycles per %d elements = %5.2f cyc/element, fp64 pipe %.0f %%  [%s]\n", T,
           T / 32, T / 128, fa.numRegs, (size_t)fa.localSizeBytes, STAGGER, c, T, c / T, 100.0 * floor_c / c, cudaGetErrorString(e));
    cudaFree(d_et); cudaFree(d_x); cudaFree(d_X); cudaFree(sink); cudaFree(cyc);
}

int main() {
    run<128, 0>(); run<256, 0>(); run<384, 0>(); run<512, 0>();
    run<256, 600>(); run<384, 400>(); run<512, 300>();
    return 0;
}
