// ubench_pipes.cu -- per-SM throughput of the warp-level primitives a radix ranking can be built
// from (sm_100a).  Prints cycles per warp-instruction per SM at full occupancy (64 warps / SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 2048
#define UNROLL 8

__device__ __forceinline__ uint32_t lcg(uint32_t& s) {
    s = s * 1664525u + 1013904223u;
    return s >> 10;
}

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t* out, long long* clk) {
    __shared__ uint32_t sm[32 * 256];
    // 32 warps per CTA, 256 words per warp (1 KB): random within the warp's region
    uint32_t* reg = sm + (threadIdx.x >> 5) * 256;
    const int lane = threadIdx.x & 31;
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 97u + 1u;
    uint32_t acc = 0;
    for (int i = lane; i < 256; i += 32) reg[i] = 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t r = lcg(s);
            if (OP == 0) {          // baseline: the lcg only
                acc += r;
            } else if (OP == 1) {   // vote.ballot
                acc += __ballot_sync(0xffffffffu, r & 1);
            } else if (OP == 2) {   // shfl idx
                acc += __shfl_sync(0xffffffffu, r, (r >> 3) & 31);
            } else if (OP == 3) {   // match.any on 10 bits
                acc += __match_any_sync(0xffffffffu, r & 1023);
            } else if (OP == 4) {   // ATOMS.OR no return, random word of 256
                atomicOr(&reg[r & 255], 1u << lane);
            } else if (OP == 5) {   // ATOMS.ADD no return
                atomicAdd(&reg[r & 255], 1u);
            } else if (OP == 6) {   // ATOMS.ADD with return
                acc += atomicAdd(&reg[r & 255], 1u);
            } else if (OP == 7) {   // LDS random
                acc += reg[r & 255];
            } else if (OP == 8) {   // STS random
                reg[r & 255] = r;
            } else if (OP == 9) {   // LDS conflict-free
                acc += reg[(lane + r * 32) & 255];
            } else if (OP == 10) {  // redux
                acc += __reduce_add_sync(0xffffffffu, r & 15);
            } else if (OP == 11) {  // LDS.U16 random
                acc += reinterpret_cast<uint16_t*>(reg)[r & 511];
            } else if (OP == 12) {  // shfl up
                acc += __shfl_up_sync(0xffffffffu, r, 1);
            } else if (OP == 13) {  // popc + lop
                acc += __popc(r & acc);
            } else if (OP == 14) {  // STS same-digit-heavy (zipf-ish: 50% of lanes hit word 0)
                reg[(r & 1) ? 0 : (r & 255)] = r;
            } else if (OP == 15) {  // ATOMS.OR zipf-ish
                atomicOr(&reg[(r & 1) ? 0 : (r & 255)], 1u << lane);
            }
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) out[0] = acc + reg[lane];
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const char* name, uint32_t* out, long long* clk, int nblk, double base) {
    k<OP><<<nblk, 1024>>>(out, clk);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    k<OP><<<nblk, 1024>>>(out, clk);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    long long* h = new long long[nblk];
    cudaMemcpy(h, clk, sizeof(long long) * nblk, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < nblk; ++i) mx = h[i] > mx ? h[i] : mx;
    delete[] h;
    // 2 CTAs x 32 warps per SM
    const double per = (double)mx / ((double)ITERS * UNROLL * 64.0);
    printf("%-28s %8.3f ms  %10lld cyc  %7.3f cyc/warp-instr/SM  (minus baseline %7.3f)\n", name,
           ms, mx, per, per - base);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int nblk = p.multiProcessorCount * 2;
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    uint32_t* out;
    long long* clk;
    cudaMalloc(&out, 4096);
    cudaMalloc(&clk, sizeof(long long) * nblk);
    // baseline first
    k<0><<<nblk, 1024>>>(out, clk);
    cudaDeviceSynchronize();
    long long* h = new long long[nblk];
    cudaMemcpy(h, clk, sizeof(long long) * nblk, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < nblk; ++i) mx = h[i] > mx ? h[i] : mx;
    const double base = (double)mx / ((double)ITERS * UNROLL * 64.0);
    run<0>("baseline (lcg)", out, clk, nblk, base);
    run<1>("vote.ballot", out, clk, nblk, base);
    run<2>("shfl.idx", out, clk, nblk, base);
    run<12>("shfl.up", out, clk, nblk, base);
    run<3>("match.any", out, clk, nblk, base);
    run<10>("redux.add", out, clk, nblk, base);
    run<13>("popc+lop", out, clk, nblk, base);
    run<4>("atoms.or noret random", out, clk, nblk, base);
    run<5>("atoms.add noret random", out, clk, nblk, base);
    run<6>("atoms.add ret random", out, clk, nblk, base);
    run<15>("atoms.or noret half-same", out, clk, nblk, base);
    run<7>("lds random", out, clk, nblk, base);
    run<11>("lds.u16 random", out, clk, nblk, base);
    run<9>("lds conflict-free", out, clk, nblk, base);
    run<8>("sts random", out, clk, nblk, base);
    run<14>("sts half-same", out, clk, nblk, base);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
