// Microbenchmark: atomic throughput on B200 for the table-update patterns the terms/histogram kernels use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/atom_bench tools/atom_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix64(uint64_t z){z^=z>>30;z*=0xBF58476D1CE4E5B9ull;z^=z>>27;z*=0x94D049BB133111EBull;z^=z>>31;return z;}
enum { G_RED64, G_RED32, G_ADDF64, G_MINCHK64, G_MINCHK64_CG, G_LD_CG, G_LD_CA, S_ADD32, S_ADD64, S_MIN64, S_ADDF64, S_MINCHK64, S_PLAIN32 };
template<int MODE> __global__ void k(uint64_t* g, uint32_t nkeys, uint64_t per_thread) {
    extern __shared__ uint64_t s[];
    uint32_t* s32 = (uint32_t*)s;
    if (MODE >= S_ADD32) { for (uint32_t i = threadIdx.x; i < nkeys; i += blockDim.x) s[i] = MODE == S_MIN64 || MODE == S_MINCHK64 ? ~0ull : 0; __syncthreads(); }
    uint64_t x = mix64(blockIdx.x * 1315423911ull + threadIdx.x);
    for (uint64_t i = 0; i < per_thread; i++) {
        x = mix64(x + i);
        uint32_t key = (uint32_t)(x % nkeys);
        uint64_t v = x >> 20;
        if (MODE == G_RED64) atomicAdd((unsigned long long*)g + key, 1ull);
        if (MODE == G_RED32) atomicAdd((unsigned int*)g + key, 1u);
        if (MODE == G_ADDF64) atomicAdd((double*)g + key, (double)v);
        if (MODE == G_MINCHK64) { if (v < g[key]) atomicMin((unsigned long long*)g + key, (unsigned long long)v); }
        if (MODE == G_MINCHK64_CG) { if (v < __ldcg(g + key)) atomicMin((unsigned long long*)g + key, (unsigned long long)v); }
        if (MODE == G_LD_CG) { x ^= __ldcg(g + key); }
        if (MODE == G_LD_CA) { x ^= g[key]; }
        if (MODE == S_ADD32) atomicAdd(s32 + key, 1u);
        if (MODE == S_ADD64) atomicAdd((unsigned long long*)s + key, 1ull);
        if (MODE == S_MIN64) atomicMin((unsigned long long*)s + key, (unsigned long long)v);
        if (MODE == S_ADDF64) atomicAdd((double*)s + key, (double)v);
        if (MODE == S_MINCHK64) { if (v < s[key]) atomicMin((unsigned long long*)s + key, (unsigned long long)v); }
        if (MODE == S_PLAIN32) s32[key] += 1;
    }
    if (MODE >= S_ADD32) { __syncthreads(); if (threadIdx.x == 0) g[blockIdx.x] = s[threadIdx.x % nkeys]; }
}
template<int MODE> void run(const char* name, uint32_t nkeys, int threads, int ctas_per_sm) {
    uint64_t* g; cudaMalloc(&g, (size_t)(nkeys > 4096 ? nkeys : 4096) * 8 + 1024 * 1024); cudaMemset(g, 0xff, (size_t)nkeys * 8);
    if (MODE != G_MINCHK64 && MODE != G_MINCHK64_CG) cudaMemset(g, 0, (size_t)nkeys * 8);
    size_t smem = MODE >= S_ADD32 ? (size_t)nkeys * 8 : 0;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int grid = 148 * ctas_per_sm; uint64_t per_thread = 2000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<grid, threads, smem>>>(g, nkeys, 100); cudaDeviceSynchronize();
    cudaEventRecord(a); k<MODE><<<grid, threads, smem>>>(g, nkeys, per_thread); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)grid * threads * per_thread;
    printf("%-12s keys=%-7u threads=%-4d cta/sm=%d  %.3f ms  %.1f Gops/s  (%s)\n", name, nkeys, threads, ctas_per_sm, ms, ops / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    cudaFree(g);
}
int main() {
    for (uint32_t nk : {10000u, 100000u, 1000000u, 16000000u}) {
        run<G_RED64>("g_red64", nk, 256, 8); run<G_ADDF64>("g_addf64", nk, 256, 8); run<G_RED32>("g_red32", nk, 256, 8);
        run<G_ADDF64>("g_addf64", nk, 512, 4);
    }
    for (uint32_t nk : {11u, 1000u, 10000u, 100000u, 1000000u}) {
        run<G_MINCHK64>("g_minchk64", nk, 256, 8); run<G_MINCHK64_CG>("g_minchk_cg", nk, 256, 8); run<G_LD_CG>("g_ld_cg", nk, 256, 8); run<G_LD_CA>("g_ld_ca", nk, 256, 8);
    }
    for (uint32_t nk : {10000u}) {
        int c = nk <= 1000 ? 4 : 1;
        run<S_ADD32>("s_add32", nk, 512, c); run<S_ADD64>("s_add64", nk, 512, c); run<S_MIN64>("s_min64", nk, 512, c);
        run<S_ADDF64>("s_addf64", nk, 512, c); run<S_MINCHK64>("s_minchk64", nk, 512, c); run<S_PLAIN32>("s_plain32", nk, 512, c);
    }
    run<S_ADD32>("s_add32", 10000, 1024, 1); run<S_ADD32>("s_add32", 10000, 1024, 2); run<S_ADDF64>("s_addf64", 10000, 1024, 2); run<S_MINCHK64>("s_minchk64", 10000, 1024, 2);
    return 0;
}
