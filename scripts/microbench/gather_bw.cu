// gather_bw.cu — what can one SM pull through sector-sparse 64-bit gathers?  (evidence for DESIGN.md §4.5)
// Every lane loads an aligned 8-byte word at byte stride S from its neighbour (the warp kernel's access shape: 9 B per lane
// at 3x decimation, 23 B at 7.7x), U independent loads in flight per thread, rows 5760 B apart (1080p pitch).
// Prints sectors/us/SM and GB/s of 32-byte sectors for L2-resident and DRAM-resident arrays.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int U>
__global__ void gather(const unsigned char *__restrict__ base, unsigned row_mask, int stride, int iters, unsigned long long *sink) {
    const int lane = threadIdx.x & 31;
    const size_t warp_global = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    unsigned long long acc = 0;
    size_t row = warp_global;
    for (int it = 0; it < iters; ++it) {
        uint2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            // a fresh 1080p-like row per load: rows are 5760 B apart, the warp's span starts at a pseudo-random column
            // (32-bit hash + masks: the address arithmetic must stay far below the load cost being measured)
            const unsigned r = ((unsigned)row + (unsigned)u * (unsigned)nwarps) * 2654435761u;
            const size_t off = (size_t)((r >> 8) & row_mask) * 5760 + ((r >> 3) & 2047u) + (unsigned)lane * (unsigned)stride;
            v[u] = __ldg(reinterpret_cast<const uint2 *>(base + (off & ~(size_t)7)));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y;
        row += (size_t)U * nwarps;
    }
    if (acc == 0x1234567ull) *sink = acc;
}

// the same gather through the texture path: tex1Dfetch of 8-byte texels on a linear texture over the array
template <int U>
__global__ void gather_tex(cudaTextureObject_t tex, unsigned row_mask, int stride, int iters, unsigned long long *sink) {
    const int lane = threadIdx.x & 31;
    const size_t warp_global = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    unsigned long long acc = 0;
    size_t row = warp_global;
    for (int it = 0; it < iters; ++it) {
        uint2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned r = ((unsigned)row + (unsigned)u * (unsigned)nwarps) * 2654435761u;
            const size_t off = (size_t)((r >> 8) & row_mask) * 5760 + ((r >> 3) & 2047u) + (unsigned)lane * (unsigned)stride;
            v[u] = tex1Dfetch<uint2>(tex, (int)(off >> 3));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y;
        row += (size_t)U * nwarps;
    }
    if (acc == 0x1234567ull) *sink = acc;
}

// half of the gathers through the LSU, half through the texture path: do the two front ends add up?
template <int U>
__global__ void gather_mix(const unsigned char *__restrict__ base, cudaTextureObject_t tex, unsigned row_mask, int stride, int iters,
                           unsigned long long *sink) {
    const int lane = threadIdx.x & 31;
    const size_t warp_global = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    unsigned long long acc = 0;
    size_t row = warp_global;
    for (int it = 0; it < iters; ++it) {
        uint2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned r = ((unsigned)row + (unsigned)u * (unsigned)nwarps) * 2654435761u;
            const size_t off = (size_t)((r >> 8) & row_mask) * 5760 + ((r >> 3) & 2047u) + (unsigned)lane * (unsigned)stride;
            if (u & 1) v[u] = tex1Dfetch<uint2>(tex, (int)(off >> 3));
            else v[u] = __ldg(reinterpret_cast<const uint2 *>(base + (off & ~(size_t)7)));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y;
        row += (size_t)U * nwarps;
    }
    if (acc == 0x1234567ull) *sink = acc;
}
template <int U>
static void run_mix(const unsigned char *d, cudaTextureObject_t tex, size_t bytes, int stride, int ctas_per_sm, int threads, int sms,
                    unsigned long long *sink) {
    const int iters = 2000 / U;
    const int grid = sms * ctas_per_sm;
    unsigned rows = 1;
    while ((size_t)(rows * 2) * 5760 <= bytes) rows *= 2;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    gather_mix<U><<<grid, threads>>>(d, tex, rows - 1, stride, iters, sink);
    cudaEventRecord(a);
    gather_mix<U><<<grid, threads>>>(d, tex, rows - 1, stride, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double warp_loads = (double)grid * (threads / 32) * iters * U;
    printf("  MIX bytes=%4zu MB stride=%2d U=%2d ctas/SM=%d thr=%4d : %7.1f us  %6.0f warp-requests/us/SM\n", bytes >> 20, stride, U, ctas_per_sm,
           threads, ms * 1e3, warp_loads / (ms * 1e3) / sms);
}

template <int U>
static void run_tex(cudaTextureObject_t tex, size_t bytes, int stride, int ctas_per_sm, int threads, int sms, unsigned long long *sink) {
    const int iters = 2000 / U;
    const int grid = sms * ctas_per_sm;
    unsigned rows = 1;
    while ((size_t)(rows * 2) * 5760 <= bytes) rows *= 2;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    gather_tex<U><<<grid, threads>>>(tex, rows - 1, stride, iters, sink);
    cudaEventRecord(a);
    gather_tex<U><<<grid, threads>>>(tex, rows - 1, stride, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double warp_loads = (double)grid * (threads / 32) * iters * U;
    printf("  TEX bytes=%4zu MB stride=%2d U=%2d ctas/SM=%d thr=%4d : %7.1f us  %6.0f warp-fetches/us/SM\n", bytes >> 20, stride, U, ctas_per_sm,
           threads, ms * 1e3, warp_loads / (ms * 1e3) / sms);
}

template <int U>
static void run(const unsigned char *d, size_t bytes, int stride, int ctas_per_sm, int threads, int sms, unsigned long long *sink) {
    const int iters = 2000 / U;
    const int grid = sms * ctas_per_sm;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    unsigned rows = 1;
    while ((size_t)(rows * 2) * 5760 <= bytes) rows *= 2;
    gather<U><<<grid, threads>>>(d, rows - 1, stride, iters, sink);
    cudaEventRecord(a);
    gather<U><<<grid, threads>>>(d, rows - 1, stride, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double warp_loads = (double)grid * (threads / 32) * iters * U;
    const double sect_per_load = (32.0 * stride + 8) / 32.0 + 0.5;   // sectors a warp's span touches (approx.)
    const double sectors = warp_loads * sect_per_load;
    printf("  bytes=%4zu MB stride=%2d U=%2d ctas/SM=%d thr=%4d : %7.1f us  %6.0f warp-loads/us/SM  ~%5.0f sectors/us/SM  ~%5.2f TB/s of sectors\n",
           bytes >> 20, stride, U, ctas_per_sm, threads, ms * 1e3, warp_loads / (ms * 1e3) / sms, sectors / (ms * 1e3) / sms,
           sectors * 32 / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    unsigned long long *sink;
    cudaMalloc(&sink, 8);
    for (size_t mb : {90, 720}) {
        const size_t bytes = mb << 20;
        unsigned char *d;
        cudaMalloc(&d, bytes + 65536);
        cudaMemset(d, 1, bytes + 65536);
        printf("%s array (%zu MB)\n", mb <= 100 ? "L2-resident" : "DRAM-resident", mb);
        for (int stride : {9, 23}) {
            run<4>(d, bytes, stride, 4, 224, sms, sink);
            run<8>(d, bytes, stride, 4, 224, sms, sink);
            run<16>(d, bytes, stride, 4, 224, sms, sink);
            run<16>(d, bytes, stride, 8, 256, sms, sink);
        }
        if (mb <= 1000) {
            cudaResourceDesc rd = {};
            rd.resType = cudaResourceTypeLinear;
            rd.res.linear.devPtr = d;
            rd.res.linear.desc = cudaCreateChannelDesc<uint2>();
            rd.res.linear.sizeInBytes = bytes;
            cudaTextureDesc td = {};
            td.readMode = cudaReadModeElementType;
            cudaTextureObject_t tex = 0;
            cudaError_t e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
            if (e != cudaSuccess) printf("  texture: %s\n", cudaGetErrorString(e));
            else
                for (int stride : {9, 23}) {
                    run_tex<4>(tex, bytes, stride, 4, 224, sms, sink);
                    run_tex<8>(tex, bytes, stride, 4, 224, sms, sink);
                    run_tex<8>(tex, bytes, stride, 8, 256, sms, sink);
                    run_mix<4>(d, tex, bytes, stride, 4, 224, sms, sink);
                    run_mix<8>(d, tex, bytes, stride, 4, 224, sms, sink);
                }
        }
        cudaFree(d);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
