// gather_px.cu — how should one output pixel of the fixed-point warp fetch its 2 x 6 tap bytes?  (evidence for DESIGN.md §4.5)
// A "pixel" = a 6-byte tap pair at an arbitrary byte offset on two rows 5760 B apart; 32 lanes = 32 consecutive output
// pixels whose taps are `stride` bytes apart (3 B x decimation).  Variants of the fetch:
//   V0  4 x LDG.64   (aligned word + next word, two rows)          — warp_fixed_kernel as of round 1
//   V1  2 x LDG.128  (aligned 16-byte block, two rows)             — lower bound: pretends every phase fits one block
//   V2  2 x LDG.128 + 2 x LDG.64 predicated on phase > 10          — second block only for the lanes that straddle
//   V3  2 x LDG.64   (half of V0: does the time follow the request count?)
//   V4  2 x LDG.128 + neighbour shuffle for straddling lanes (stride <= 16), predicated LDG.64 for the warp's last lanes
// Prints pixels/us/SM for DRAM-resident frames (720 MB) and L2-resident ones (90 MB).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int V, int U>
__global__ void __launch_bounds__(224) gather_px(const unsigned char *__restrict__ base, unsigned row_mask, int stride, int iters,
                                                 unsigned long long *sink) {
    const int lane = threadIdx.x & 31;
    const size_t warp_global = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    unsigned acc = 0;
    size_t row = warp_global;
    for (int it = 0; it < iters; ++it) {
        uint4 a[U], b[U];
        uint2 c[U], d[U];
        unsigned ph[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned r = ((unsigned)row + (unsigned)u * (unsigned)nwarps) * 2654435761u;
            const size_t off = (size_t)((r >> 8) & row_mask) * 5760 + ((r >> 3) & 2047u) + (unsigned)lane * (unsigned)stride;
            const unsigned char *p0 = base + off, *p1 = p0 + 5760;
            a[u] = b[u] = make_uint4(0, 0, 0, 0);
            c[u] = d[u] = make_uint2(0, 0);
            if (V == 0 || V == 3) {
                const uint2 *q0 = reinterpret_cast<const uint2 *>(reinterpret_cast<size_t>(p0) & ~(size_t)7);
                const uint2 *q1 = reinterpret_cast<const uint2 *>(reinterpret_cast<size_t>(p1) & ~(size_t)7);
                const uint2 x0 = __ldg(q0), y0 = __ldg(q1);
                a[u].x = x0.x; a[u].y = x0.y; b[u].x = y0.x; b[u].y = y0.y;
                if (V == 0) {
                    c[u] = __ldg(q0 + 1);
                    d[u] = __ldg(q1 + 1);
                }
            } else {
                const uint4 *q0 = reinterpret_cast<const uint4 *>(reinterpret_cast<size_t>(p0) & ~(size_t)15);
                const uint4 *q1 = reinterpret_cast<const uint4 *>(reinterpret_cast<size_t>(p1) & ~(size_t)15);
                a[u] = __ldg(q0);
                b[u] = __ldg(q1);
                ph[u] = (unsigned)(reinterpret_cast<size_t>(p0) & 15);
                if (V == 2 && ph[u] > 10) {
                    c[u] = __ldg(reinterpret_cast<const uint2 *>(q0 + 1));
                    d[u] = __ldg(reinterpret_cast<const uint2 *>(q1 + 1));
                }
                if (V == 4) {
                    // the block after mine is held by lane + j (j = 1 or 2 for stride >= 8) when stride <= 16
                    const size_t mine = reinterpret_cast<size_t>(q0);
                    const size_t n1 = __shfl_down_sync(0xffffffffu, mine, 1), n2 = __shfl_down_sync(0xffffffffu, mine, 2);
                    const int j = (n1 == mine + 16) ? 1 : ((n2 == mine + 16) ? 2 : 0);
                    const bool have = j != 0 && lane + j < 32;
                    const unsigned ax = __shfl_sync(0xffffffffu, a[u].x, lane + j), ay = __shfl_sync(0xffffffffu, a[u].y, lane + j);
                    const unsigned bx = __shfl_sync(0xffffffffu, b[u].x, lane + j), by = __shfl_sync(0xffffffffu, b[u].y, lane + j);
                    if (ph[u] > 10) {
                        if (have) { c[u] = make_uint2(ax, ay); d[u] = make_uint2(bx, by); }
                        else {
                            c[u] = __ldg(reinterpret_cast<const uint2 *>(q0 + 1));
                            d[u] = __ldg(reinterpret_cast<const uint2 *>(q1 + 1));
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            acc += a[u].x ^ a[u].y ^ a[u].z ^ a[u].w ^ b[u].x ^ b[u].y ^ b[u].z ^ b[u].w ^ c[u].x ^ c[u].y ^ d[u].x ^ d[u].y;
        row += (size_t)U * nwarps;
    }
    if (acc == 0x12345u) *sink = acc;
}

template <int V, int U>
static void run(const unsigned char *d, size_t bytes, int stride, int sms, unsigned long long *sink) {
    const int iters = 1600 / U, threads = 224, grid = sms * 4;
    unsigned rows = 1;
    while ((size_t)(rows * 2) * 5760 <= bytes - 8192) rows *= 2;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    gather_px<V, U><<<grid, threads>>>(d, rows - 1, stride, iters, sink);
    cudaEventRecord(a);
    gather_px<V, U><<<grid, threads>>>(d, rows - 1, stride, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double px = (double)grid * threads * iters * U;
    printf("  V%d stride=%2d U=%d : %8.1f us  %7.0f pixels/us/SM  (%5.1f warp-pixel-rows/us/SM)\n", V, stride, U, ms * 1e3, px / (ms * 1e3) / sms,
           px / 32 / (ms * 1e3) / sms);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    unsigned long long *sink;
    cudaMalloc(&sink, 8);
    for (size_t mb : {720, 90}) {
        const size_t bytes = mb << 20;
        unsigned char *d;
        cudaMalloc(&d, bytes + 65536);
        cudaMemset(d, 1, bytes + 65536);
        printf("%s frames (%zu MB)\n", mb <= 100 ? "L2-resident" : "DRAM-resident", mb);
        for (int stride : {6, 9, 14, 23}) {
            run<0, 4>(d, bytes, stride, sms, sink);
            run<3, 4>(d, bytes, stride, sms, sink);
            run<1, 4>(d, bytes, stride, sms, sink);
            run<2, 4>(d, bytes, stride, sms, sink);
            run<4, 4>(d, bytes, stride, sms, sink);
            run<1, 2>(d, bytes, stride, sms, sink);
            run<2, 2>(d, bytes, stride, sms, sink);
        }
        cudaFree(d);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
