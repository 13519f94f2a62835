// Microbenchmark (development aid): shared-memory / constant-bank load throughput per SM on sm_100a for the access shapes the
// tile sweep kernel uses.  Prints bytes/clk/SM and issue cycles per warp instruction.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_paths smem_paths.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int NT = 512;
constexpr int ITER = 2048;

struct Params { double c[2048]; };

template <int MODE>
__global__ void __launch_bounds__(NT, 1) k(const __grid_constant__ Params prm, double* out, long long* clk, int stride_sel, int nwarps_active) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* tile = reinterpret_cast<double2*>(smem);
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 8192; i += NT) tile[i] = make_double2(i * 0.5, i * 0.25);
    __syncthreads();
    double acc0 = 0, acc1 = 0;
    long long t0 = 0, t1 = 0;
    if ((tid >> 5) < nwarps_active) {
        t0 = clock64();
        if (MODE == 0) {  // conflict-free LDS.128, 16 per iteration, addresses permuted by an XOR (like partner rows)
            uint32_t w = (tid << 4) ^ ((tid) & 7);
            for (int it = 0; it < ITER; ++it) {
                const uint32_t pw = w ^ ((it * 37) & 0x1ff0);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const double2 p = tile[pw ^ r];
                    acc0 += p.x;
                    acc1 += p.y;
                }
            }
        } else if (MODE == 1) {  // broadcast LDS.128 (all lanes same address; address in a vector register)
            for (int it = 0; it < ITER; ++it) {
                const uint32_t b = ((it * 37 + stride_sel * lane) & 0x1ff0);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const double2 p = tile[b + r];
                    acc0 += p.x;
                    acc1 += p.y;
                }
            }
        } else if (MODE == 2) {  // broadcast LDS.64
            const double* t64 = reinterpret_cast<const double*>(smem);
            for (int it = 0; it < ITER; ++it) {
                const uint32_t b = ((it * 37 + stride_sel * lane) & 0x1ff0);
#pragma unroll
                for (int r = 0; r < 16; ++r) acc0 += t64[b + r];
            }
        } else if (MODE == 3) {  // broadcast LDS.32
            const float* t32 = reinterpret_cast<const float*>(smem);
            for (int it = 0; it < ITER; ++it) {
                const uint32_t b = ((it * 37 + stride_sel * lane) & 0x1ff0);
#pragma unroll
                for (int r = 0; r < 16; ++r) acc0 += t32[b + r];
            }
        } else if (MODE == 4) {  // constant bank (kernel parameter), dynamic warp-uniform index
            for (int it = 0; it < ITER; ++it) {
                const uint32_t b = ((it * 37 + stride_sel * lane) & 0x7f0);
#pragma unroll
                for (int r = 0; r < 16; ++r) acc0 += prm.c[b + r];
            }
        } else if (MODE == 5) {  // LDS.128 with half of every quarter-warp pattern: lanes with bit 3 set idle
            uint32_t w = (tid << 4) ^ ((tid) & 7);
            if (!((lane >> 3) & 1)) {
                for (int it = 0; it < ITER; ++it) {
                    const uint32_t pw = w ^ ((it * 37) & 0x1ff0);
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const double2 p = tile[pw ^ r];
                        acc0 += p.x;
                        acc1 += p.y;
                    }
                }
            }
        } else if (MODE == 6) {  // LDS.128, lanes with bit 0 set idle (no whole quarter warp idles)
            uint32_t w = (tid << 4) ^ ((tid) & 7);
            if (!(lane & 1)) {
                for (int it = 0; it < ITER; ++it) {
                    const uint32_t pw = w ^ ((it * 37) & 0x1ff0);
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const double2 p = tile[pw ^ r];
                        acc0 += p.x;
                        acc1 += p.y;
                    }
                }
            }
        } else if (MODE == 8) {  // partner-row LDS.128 x16 interleaved with 8 indexed LDC.64 (constant path next to the shared path)
            uint32_t w = (tid << 4) ^ ((tid) & 7);
            for (int it = 0; it < ITER; ++it) {
                const uint32_t pw = w ^ ((it * 37) & 0x1ff0);
                const uint32_t b = ((it * 37 + stride_sel * lane) & 0x7f0);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const double2 p = tile[pw ^ r];
                    acc0 += p.x;
                    acc1 += p.y;
                    if (r & 1) acc1 += prm.c[b + (r >> 1)];
                }
            }
        } else if (MODE == 9) {  // partner-row LDS.128 x16 interleaved with 4 broadcast LDS.128 (today's term records)
            uint32_t w = (tid << 4) ^ ((tid) & 7);
            for (int it = 0; it < ITER; ++it) {
                const uint32_t pw = w ^ ((it * 37) & 0x1ff0);
                const uint32_t b = ((it * 37 + stride_sel * lane) & 0x1ff0);
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const double2 p = tile[pw ^ r];
                    acc0 += p.x;
                    acc1 += p.y;
                    if ((r & 3) == 3) {
                        const double2 q = tile[b + (r >> 2)];
                        acc0 += q.x;
                        acc1 += q.y;
                    }
                }
            }
        } else if (MODE == 7) {  // conflict-free LDS.64 (two per amplitude)
            const double* t64 = reinterpret_cast<const double*>(smem);
            for (int it = 0; it < ITER; ++it) {
                const uint32_t b = ((tid * 32) ^ ((it * 37) & 0x3fe0)) & 0x3fff;
#pragma unroll
                for (int r = 0; r < 16; ++r) acc0 += t64[(b + 2 * r) & 0x3fff];
            }
        }
        t1 = clock64();
    }
    out[blockIdx.x * NT + tid] = acc0 + acc1;
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double bytes_per_load, int stride_sel, int nwarps) {
    Params* hp = new Params();
    for (int i = 0; i < 2048; ++i) hp->c[i] = i;
    double* out;
    long long* clk;
    cudaMalloc(&out, 148 * NT * 8);
    cudaMalloc(&clk, 148 * 8);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    k<MODE><<<148, NT, 131072>>>(*hp, out, clk, stride_sel, nwarps);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, NT, 131072>>>(*hp, out, clk, stride_sel, nwarps);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    const double loads = (double)nwarps * ITER * 16;  // warp-level load instructions per SM
    printf("%-58s warps=%2d  %.1f us  clk(warp0)=%lld  clk per warp-load=%.2f  bytes/clk/SM=%.1f  err=%s\n", name, nwarps, ms * 1e3, h[0], h[0] / loads,
           loads * bytes_per_load / h[0], cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(clk);
    delete hp;
}

int main() {
    for (int nw : {16, 8}) {
        run<0>("LDS.128 conflict-free (partner rows)", 512, 0, nw);
        run<1>("LDS.128 broadcast (term record)", 16, 0, nw);
        run<2>("LDS.64 broadcast", 8, 0, nw);
        run<3>("LDS.32 broadcast", 4, 0, nw);
        run<4>("LDC.64 param bank, uniform dynamic index", 8, 0, nw);
        run<5>("LDS.128 half the quarter-warps idle (lane bit 3)", 256, 0, nw);
        run<6>("LDS.128 odd lanes idle (lane bit 0)", 256, 0, nw);
        run<8>("16 LDS.128 partner + 8 LDC.64 indexed per iteration", 512, 0, nw);
        run<9>("16 LDS.128 partner + 4 LDS.128 broadcast per iteration", 512, 0, nw);
    }
    return 0;
}
