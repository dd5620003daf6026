// fp64 pipe micro-benchmark: cycles per DFMA per SMSP as a function of warps per SMSP and ILP per thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu && ./fp64_pipe
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, long long *cyc, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps_per_block) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<ILP><<<148, 32 * warps_per_block>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    k<ILP><<<148, 32 * warps_per_block>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_dfma_warp = (double)h / (iters * ILP);                       // cycles per DFMA seen by one warp
    double per_smsp = (double)h / (iters * ILP * (warps_per_block / 4.0));  // cycles per warp-DFMA per SMSP
    printf("warps/SM %2d (per SMSP %.1f) ILP %d : %6.2f cyc per DFMA per warp, %5.2f cyc per warp-DFMA per SMSP\n", warps_per_block,
           warps_per_block / 4.0, ILP, per_dfma_warp, warps_per_block >= 4 ? per_smsp : 0.0);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
