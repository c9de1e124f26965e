// Gather microbenchmark: what this B200 sustains for independent 32-byte sector
// reads at random addresses -- from a 64 MiB buffer (L2-resident, like the 256^3 cube) and from
// a 2 GiB buffer (HBM-resident) -- and, for scale, a plain streaming read.  Prints one JSON line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_roofline tools/gather_roofline.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                    \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ unsigned mix(unsigned h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

// every thread reads `iters` random 32-byte sectors (two 16-byte loads of one sector)
__global__ void __launch_bounds__(256) gather32(const float4* __restrict__ buf, unsigned long long nsectors,
                                                int iters, float* sink) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    unsigned h = mix(tid * 2654435761u + 12345u);
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        h = mix(h + i);
        const unsigned long long s = ((unsigned long long)h * nsectors) >> 32;
        const float4 a = __ldg(buf + 2 * s);
        const float4 b = __ldg(buf + 2 * s + 1);
        acc += a.x + a.w + b.y + b.z;
    }
    if (acc == 123.456f) *sink = acc;
}

__global__ void __launch_bounds__(256) stream_read(const float4* __restrict__ buf, unsigned long long n, float* sink) {
    float acc = 0.f;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(buf + i);
        acc += a.x + a.y + a.z + a.w;
    }
    if (acc == 123.456f) *sink = acc;
}

static double time_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

int main() {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float* sink;
    CK(cudaMalloc(&sink, 4));
    const size_t sizes[2] = {size_t(64) << 20, size_t(2) << 30};
    double gather_gbs[2], stream_gbs[2];
    for (int k = 0; k < 2; ++k) {
        float4* buf;
        CK(cudaMalloc(&buf, sizes[k]));
        CK(cudaMemset(buf, 0, sizes[k]));
        const unsigned long long nsectors = sizes[k] / 32;
        const int blocks = 148 * 8 * 4, iters = 256;
        for (int rep = 0; rep < 3; ++rep) gather32<<<blocks, 256>>>(buf, nsectors, iters, sink);
        CK(cudaDeviceSynchronize());
        double best = 1e30;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            gather32<<<blocks, 256>>>(buf, nsectors, iters, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            const double ms = time_ms(e0, e1);
            if (ms < best) best = ms;
        }
        gather_gbs[k] = (double)blocks * 256 * iters * 32 / (best * 1e-3) / 1e9;
        best = 1e30;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            stream_read<<<148 * 16, 256>>>(buf, sizes[k] / 16, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            const double ms = time_ms(e0, e1);
            if (ms < best) best = ms;
        }
        stream_gbs[k] = (double)sizes[k] / (best * 1e-3) / 1e9;
        CK(cudaFree(buf));
    }
    printf("{\"gather32_l2_64MiB_GBs\": %.1f, \"gather32_hbm_2GiB_GBs\": %.1f, \"stream_read_64MiB_GBs\": %.1f, "
           "\"stream_read_2GiB_GBs\": %.1f, \"how\": \"random independent 32 B sector reads, 1.2M threads x 256 reads, "
           "best of 5, CUDA events\"}\n",
           gather_gbs[0], gather_gbs[1], stream_gbs[0], stream_gbs[1]);
    return 0;
}
