// Microbenchmark: how many L1 data-pipe wavefronts does one warp-wide 32-byte-per-lane load (LDG.E.256) cost, as a
// function of which 128-byte lines its lanes touch and how the lanes are grouped?  One kernel per pattern, run under
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__lsu_writeback_active.sum,gpu__time_duration.sum
// The march kernels are bound by exactly this quantity (profiles/r2b_march_orbit_fast.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ldg256_wavefronts tools/ldg256_wavefronts.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float load32(const float* p) {
    float a0, a1, a2, a3, b0, b1, b2, b3;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3)
                 : "l"(p));
    return a0 + a1 + a2 + a3 + b0 + b1 + b2 + b3;
}
__device__ __forceinline__ float load16(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v.x + v.y + v.z + v.w;
}

// record index (32-byte units) of lane `l` for pattern P
template <int P>
__device__ __forceinline__ int pattern(int l) {
    switch (P) {
        case 0: return 0;                          // all lanes one record
        case 1: return l;                          // 32 consecutive records: 8 lines, 4 consecutive lanes per line
        case 2: return 4 * l;                      // every lane its own line (32 lines)
        case 3: return (l & 3) + 4 * (l >> 3);     // 4 lines, 8 consecutive lanes per line
        case 4: return (l & 3) + 4 * (l >> 4);     // 2 lines, 16 consecutive lanes per line
        case 5: return l >> 1;                     // lane pairs share a record: 16 records = 4 lines
        case 6: return 4 * (l & 1) + ((l >> 1) & 3);  // 2 lines, lanes alternate between them
        case 7: return 4 * (l & 3) + ((l >> 2) & 3);  // 4 lines, lanes interleaved 0123 0123 ...
        case 8: return 4 * (l & 7) + ((l >> 3) & 3);  // 8 lines, lanes interleaved
        case 9: return (l & 1) + 4 * (l >> 1);        // 16 lines, 2 consecutive lanes per line
        case 10: return l & 3;                        // one line, 4 records, lanes cycle through them
        case 11: return (l >> 3);                     // 4 records of one line, 8 consecutive lanes per record
        case 12: return 4 * (l >> 2);                 // 8 lines, the 4 lanes of a group read the SAME record
        case 13: return 4 * (l >> 2) + (l & 1);       // 8 lines, 2 records per group
        default: return l;
    }
}

template <int P, bool WIDE>
__global__ void probe(const float* __restrict__ base, float* out, int iters, int span_records) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int rec = pattern<P>(lane) + warp * 128;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        const float* p = base + (size_t)(rec % span_records) * 8;
        acc += WIDE ? load32(p) : (load16(p) + load16(p + 4));
        rec += 1024;
    }
    if (acc == 12345.678f) out[0] = acc;
}

template <int P>
void run(const float* d, float* out, int span) {
    probe<P, true><<<148 * 8, 256>>>(d, out, 64, span);
    probe<P, false><<<148 * 8, 256>>>(d, out, 64, span);
}

int main() {
    const int span = 1 << 18;  // 8 MiB of records: L2-resident, L1 misses
    float *d, *out;
    cudaMalloc(&d, (size_t)span * 32);
    cudaMemset(d, 0, (size_t)span * 32);
    cudaMalloc(&out, 16);
    run<0>(d, out, span); run<1>(d, out, span); run<2>(d, out, span); run<3>(d, out, span); run<4>(d, out, span);
    run<5>(d, out, span); run<6>(d, out, span); run<7>(d, out, span); run<8>(d, out, span); run<9>(d, out, span);
    run<10>(d, out, span); run<11>(d, out, span); run<12>(d, out, span); run<13>(d, out, span);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
