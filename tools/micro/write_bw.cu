// What HBM bandwidth can a WRITE-dominated stream reach on this box?  The fused kernel's unique bytes are ~90 % stores
// (C3 shard: 674 MB written, 78 MB read per launch), the measured peak it is scored against is a COPY (half reads, half
// writes).  Three streaming kernels over buffers far larger than the L2: copy (1 : 1), write only, and 1 read : 9 writes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_variants/write_bw tools/micro/write_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_copy(const uint4 *__restrict__ a, uint4 *__restrict__ b, size_t n)
{ for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i]; }
__global__ void k_write(uint4 *__restrict__ b, size_t n)
{ for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = make_uint4((unsigned)i, 1u, 2u, 3u); }
// one 16-byte load feeds nine 16-byte stores
__global__ void k_mix(const uint4 *__restrict__ a, uint4 *__restrict__ b, size_t n_in)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_in; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = a[i];
        #pragma unroll
        for (int k = 0; k < 9; k++) { b[(size_t)k * n_in + i] = v; v.x += 1u; }
    }
}
template <typename F> static double time_ms(F f, cudaStream_t s)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaStreamSynchronize(s);
    float best = 1e30f;
    for (int r = 0; r < 7; r++) { cudaEventRecord(e0, s); f(); cudaEventRecord(e1, s); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}
int main()
{
    cudaStream_t s; cudaStreamCreate(&s);
    const size_t bytes = (size_t)2 << 30, n = bytes / 16;                 // 2 GiB per buffer
    uint4 *a, *b; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes * 9 / 8 + 4096); cudaMemset(a, 1, bytes);
    for (int cta : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        double c = time_ms([&] { k_copy<<<cta, 512, 0, s>>>(a, b, n); }, s);
        double w = time_ms([&] { k_write<<<cta, 512, 0, s>>>(b, n); }, s);
        const size_t n_in = n / 8;                                       // 256 MiB read, 2.25 GiB written
        double m = time_ms([&] { k_mix<<<cta, 512, 0, s>>>(a, b, n_in); }, s);
        printf("grid %5d x 512: copy %.0f GB/s (read + write)   write only %.0f GB/s   1 read : 9 writes %.0f GB/s\n", cta,
               2.0 * bytes / c / 1e6, (double)bytes / w / 1e6, 10.0 * n_in * 16 / m / 1e6);
    }
    printf("cudaMemsetAsync 2 GiB: %.0f GB/s\n", bytes / time_ms([&] { cudaMemsetAsync(b, 0, bytes, s); }, s) / 1e6);
    printf("cudaMemcpyAsync D2D 2 GiB: %.0f GB/s (read + write)\n", 2.0 * bytes / time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, s); }, s) / 1e6);
    return 0;
}
