// What does a tiny kernel cost on a B200 when it sits in a chain of dependent stream operations?  Each variant is
// launched 200 times back to back on one stream (CUDA events around the lot -> us per launch), so that the per-launch
// figure is the dependency-to-dependency cost, not the launch latency of one kernel alone.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_variants/small_kernels tools/micro/small_kernels.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_empty() {}
__global__ void k_store(unsigned *p) { p[blockIdx.x * blockDim.x + threadIdx.x] = threadIdx.x; }
__global__ void k_atomic(unsigned *p) { atomicAdd(&p[(blockIdx.x * blockDim.x + threadIdx.x) & 1023], 1u); }
__global__ void k_atomic64(unsigned long long *p) { atomicAdd(&p[threadIdx.x & 1], 1ull); }
__global__ void k_host(unsigned *h) { h[blockIdx.x * blockDim.x + threadIdx.x] = threadIdx.x; }
__global__ void k_chain(const unsigned *idx, unsigned *out)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned a = idx[i & 4095]; a = idx[a & 4095]; a = idx[a & 4095]; a = idx[a & 4095];
    out[i] = a;
}
template <typename F> static float run(const char *name, F launch, cudaStream_t s)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 20; i++) launch();
    cudaStreamSynchronize(s);
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, s);
        for (int i = 0; i < 200; i++) launch();
        cudaEventRecord(e1, s); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-46s %6.2f us per launch\n", name, best * 1e3f / 200);
    return best;
}
int main()
{
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    unsigned *d, *h, *idx; cudaMalloc(&d, 1 << 22); cudaMemset(d, 0, 1 << 22); cudaMallocHost(&h, 1 << 20);
    cudaMalloc(&idx, 4096 * 4); cudaMemset(idx, 0, 4096 * 4);
    for (int g : {1, 16, 64, 148}) {
        char n[96];
        snprintf(n, 96, "empty            grid %3d x 256", g);  run(n, [&] { k_empty<<<g, 256, 0, s>>>(); }, s);
        snprintf(n, 96, "store (device)   grid %3d x 256", g);  run(n, [&] { k_store<<<g, 256, 0, s>>>(d); }, s);
        snprintf(n, 96, "atomicAdd u32    grid %3d x 256", g);  run(n, [&] { k_atomic<<<g, 256, 0, s>>>(d); }, s);
        snprintf(n, 96, "atomicAdd u64 x2 grid %3d x 256", g);  run(n, [&] { k_atomic64<<<g, 256, 0, s>>>((unsigned long long *)d); }, s);
        snprintf(n, 96, "store (host)     grid %3d x 256", g);  run(n, [&] { k_host<<<g, 256, 0, s>>>(h); }, s);
        snprintf(n, 96, "4 dependent loads grid %3d x 256", g); run(n, [&] { k_chain<<<g, 256, 0, s>>>(idx, d); }, s);
    }
    // a copy node in the chain, for comparison
    run("cudaMemcpyAsync D2H 16 KB", [&] { cudaMemcpyAsync(h, d, 16384, cudaMemcpyDeviceToHost, s); }, s);
    run("cudaMemcpyAsync H2D 165 KB", [&] { cudaMemcpyAsync(d, h, 165 * 1024, cudaMemcpyHostToDevice, s); }, s);
    run("cudaMemsetAsync 8 KB", [&] { cudaMemsetAsync(d, 0, 8192, s); }, s);
    run("H2D 165 KB + empty kernel", [&] { cudaMemcpyAsync(d, h, 165 * 1024, cudaMemcpyHostToDevice, s); k_empty<<<16, 256, 0, s>>>(); }, s);
    run("empty kernel + D2H 16 KB", [&] { k_empty<<<16, 256, 0, s>>>(); cudaMemcpyAsync(h, d, 16384, cudaMemcpyDeviceToHost, s); }, s);
    run("store(device) kernel + host-store kernel", [&] { k_store<<<16, 256, 0, s>>>(d); k_host<<<16, 256, 0, s>>>(h); }, s);
    return 0;
}
