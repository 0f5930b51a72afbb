// membench.cu -- what HBM bandwidth do the access mixes of the PCG kernels get on this GPU?
// (diagnostic: a: copy 1:1, b: 2 fp64 reads + 1 byte read + 1 fp64 write, grid-stride, 128-bit accesses)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void k_copy(const double2* __restrict__ a, double2* __restrict__ b, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void k_mix(const double2* __restrict__ a, const double2* __restrict__ b, const uchar2* __restrict__ c, double2* __restrict__ o, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double2 x = a[i], y = b[i]; uchar2 m = c[i];
        o[i] = make_double2(x.x + y.x * m.x, x.y + y.y * m.y);
    }
}
__global__ void k_mix_inplace(const double2* __restrict__ a, double2* __restrict__ b, const uchar2* __restrict__ c, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double2 x = a[i], y = b[i]; uchar2 m = c[i];
        b[i] = make_double2(x.x + y.x * m.x, x.y + y.y * m.y);
    }
}
int main()
{
    const size_t n = (size_t)4096 * 4096 / 2;   // double2 elements of one L=4096 vector
    double2 *a, *b, *o; uchar2* c;
    cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMalloc(&o, n * 16); cudaMalloc(&c, n * 2);
    cudaMemset(a, 0, n * 16); cudaMemset(b, 0, n * 16); cudaMemset(c, 0, n * 2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) for (int bs : {256, 512}) {
        float ms;
        for (int v = 0; v < 3; ++v) {
            for (int w = 0; w < 3; ++w) { if (v == 0) k_copy<<<grid, bs>>>(a, o, n); else if (v == 1) k_mix<<<grid, bs>>>(a, b, c, o, n); else k_mix_inplace<<<grid, bs>>>(a, b, c, n); }
            cudaEventRecord(e0);
            const int reps = 20;
            for (int w = 0; w < reps; ++w) { if (v == 0) k_copy<<<grid, bs>>>(a, o, n); else if (v == 1) k_mix<<<grid, bs>>>(a, b, c, o, n); else k_mix_inplace<<<grid, bs>>>(a, b, c, n); }
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
            double bytes = v == 0 ? n * 32.0 : n * 50.0;
            printf("%s grid %5d bs %3d: %.4f ms  %.0f GB/s\n", v == 0 ? "copy 1:1      " : v == 1 ? "2r+byte -> 1w " : "2r+byte inplace", grid, bs, ms, bytes / ms / 1e6);
        }
    }
    return 0;
}
