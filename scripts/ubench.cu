// ubench.cu -- shared-memory primitive throughput on B200, to size the AUCell kernel design.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ubench scripts/ubench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 2048;

__device__ __forceinline__ uint32_t xs(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

// mode 0: LDS u16 random over n; 1: atomicAdd u32 random; 2: atomicAdd u64 random; 3: atomicOr u32 random;
// 4: atomicAdd double random; 5: match_any on random value in [0,n); 6: atomicAdd u32 same address per warp;
// 7: LDS u32 random
template <int MODE>
__global__ void k(int n, int smem_words, long long* cyc, uint32_t* sink) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint16_t* t16 = (uint16_t*)sm;
    uint32_t* t32 = (uint32_t*)sm;
    unsigned long long* t64 = (unsigned long long*)sm;
    double* tf = (double*)sm;
    for (int i = threadIdx.x; i < smem_words; i += blockDim.x) t32[i] = i & 1023;
    __syncthreads();
    // pre-generated random indices per thread (registers)
    uint32_t s = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x * 7919u;
    uint32_t ids[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) ids[j] = xs(s) % (uint32_t)n;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS / 16; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t a = ids[j];
            if (MODE == 0) acc += t16[a];
            if (MODE == 7) acc += t32[a];
            if (MODE == 1) atomicAdd(&t32[a], 1u);
            if (MODE == 2) atomicAdd(&t64[a], 1ull);
            if (MODE == 3) atomicOr(&t32[a], 1u << (a & 31));
            if (MODE == 4) atomicAdd(&tf[a], 1.0);
            if (MODE == 5) acc += __popc(__match_any_sync(0xffffffffu, a));
            if (MODE == 6) atomicAdd(&t32[(a & ~31u) % n], 1u);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0xdeadbeef) sink[0] = acc + t32[1];
}

template <int MODE>
int run(const char* name, int n, int threads, int blocks_per_sm, int smem) {
    long long* d_c; uint32_t* d_s;
    int grid = 148 * blocks_per_sm;
    CK(cudaMalloc(&d_c, grid * 8)); CK(cudaMalloc(&d_s, 64));
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k<MODE><<<grid, threads, smem>>>(n, smem / 4, d_c, d_s);
    CK(cudaDeviceSynchronize());
    k<MODE><<<grid, threads, smem>>>(n, smem / 4, d_c, d_s);
    CK(cudaDeviceSynchronize());
    std::vector<long long> c(grid);
    CK(cudaMemcpy(c.data(), d_c, grid * 8, cudaMemcpyDeviceToHost));
    double mean = 0; for (auto v : c) mean += v; mean /= grid;
    double ops_per_sm = (double)ITERS * threads * blocks_per_sm;
    printf("%-34s n=%6d thr=%4d x%d  cycles=%9.0f  lane-ops/clk/SM=%7.3f  (warp-instr/clk/SM=%6.3f)\n", name, n, threads,
           blocks_per_sm, mean, ops_per_sm / mean, ops_per_sm / mean / 32.0);
    cudaFree(d_c); cudaFree(d_s);
    return 0;
}

int main() {
    for (int thr : {256, 1024}) {
        int bps = thr == 256 ? 4 : 1;
        run<0>("LDS.u16 random", 28000, thr, bps, 55 * 1024);
        run<7>("LDS.u32 random", 14000, thr, bps, 55 * 1024);
        run<1>("ATOMS.ADD.u32 random", 512, thr, bps, 8 * 1024);
        run<1>("ATOMS.ADD.u32 random", 4096, thr, bps, 40 * 1024);
        run<2>("ATOMS.ADD.u64 random", 512, thr, bps, 16 * 1024);
        run<2>("ATOMS.ADD.u64 random", 4096, thr, bps, 80 * 1024);
        run<3>("ATOMS.OR.u32 random", 875, thr, bps, 8 * 1024);
        run<4>("atomicAdd(double) smem random", 512, thr, bps, 16 * 1024);
        run<5>("match_any random 64 values", 64, thr, bps, 1024);
        run<5>("match_any random 4096 values", 4096, thr, bps, 40 * 1024);
        run<6>("ATOMS.ADD.u32 same addr/warp", 4096, thr, bps, 40 * 1024);
    }
    return 0;
}
