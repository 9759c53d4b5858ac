// Microbenchmark of tcgen05.mma issue cost on sm_100a (kind::tf32, K = 8 per instruction), the design input of the
// conv engine: cycles per instruction as a function of N, operand majorness, A-from-TMEM, cta_group and concurrent
// generic-proxy shared-memory stores (the hi/lo producers).  Operand contents are whatever is in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_bench tools/mma_bench.cu && gpurun_out/mma_bench
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "../bts_b200/csrc/tc_common.cuh"

using namespace tc;

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// mode 0: SS K-major   1: SS MN-major   2: TS (A in TMEM), B K-major   3: cta_group::2 SS K-major (M = 256 over the pair)
// bg: number of background warps streaming st.shared.v4 into a separate 64 KB region while the MMAs run
// distinct: 1 -> the 4 k-steps x `stages` stage slots are walked (distinct smem addresses), 0 -> one address re-read
template <int mode>
__global__ void __launch_bounds__(320, 1) mma_bench(int N, int iters, int bg, int pattern, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *sm = smem_raw + (base - smem_u32(smem_raw));
    // [A 4 x 16 KB][B 4 x 32 KB][bg 32 KB][bars]
    const uint32_t a_off = 0, b_off = 64 * 1024, bg_off = 192 * 1024, bar_off = 224 * 1024;
    const uint32_t bar = base + bar_off;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + bar_off + 16);
    volatile int *stop = reinterpret_cast<volatile int *>(sm + bar_off + 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool two = mode == 3;
    uint32_t rank = 0;
    if constexpr (two) rank = cluster_rank();
    for (uint32_t i = threadIdx.x; i < 224 * 1024 / 16; i += blockDim.x) st_shared_v4(base + i * 16, 1.f, 0.5f, 0.25f, 2.f);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        *stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (two) tmem_alloc2(smem_u32(tmem_slot), 512);
        else tmem_alloc(smem_u32(tmem_slot), 512);
    }
    fence_proxy_async();
    tc_fence_before();
    if constexpr (two) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    long long t0 = 0, t1 = 0;
    if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const int M = two ? 256 : 128;
            const uint32_t idesc = (mode == 1) ? make_idesc(M, N, 1, 1) : make_idesc(M, N, 0, 0);
            t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                const int s = pattern ? (it & 3) : 0;
                const uint32_t a = base + a_off + s * 16384, b = base + b_off + s * 32768;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if constexpr (mode == 0) umma_tf32(tmem_base, make_desc(a) + 2 * k, make_desc(b) + 2 * k, idesc, 1);
                    else if constexpr (mode == 1) umma_tf32(tmem_base, make_desc_mn(a + k * 1024, 4096), make_desc_mn(b + k * 1024, 4096), idesc, 1);
                    else if constexpr (mode == 2) umma_tf32_ts(tmem_base, tmem_base + 256 + (s * 4 + k) * 8, make_desc(b) + 2 * k, idesc, 1);
                    else umma_tf32_2cta(tmem_base, make_desc(a) + 2 * k, make_desc(b) + 2 * k, idesc, 1);
                }
            }
            if constexpr (two) umma_commit_2cta(bar, 3); else umma_commit(bar);
            mbar_wait(bar, 0);
            t1 = clock64();
            *stop = 1;
        } else if (lane == 0) {
            mbar_wait(bar, 0);         // non-leader CTA of the pair: the multicast commit arrives here too
            *stop = 1;
        }
    } else if (warp >= 2 && warp < 2 + bg) {
        uint32_t o = (uint32_t)(warp - 2) * 4096u + (uint32_t)lane * 16u;
        int n = 0;
        while (!*stop) {
#pragma unroll
            for (int j = 0; j < 8; ++j) st_shared_v4(base + bg_off + ((o + j * 512u) & 32767u), 1.f, 2.f, 3.f, 4.f);
            o += 4096u * 8u;
            if (++n > (1 << 26)) break;
        }
    }
    tc_fence_before();
    if constexpr (two) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        if constexpr (two) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
    if (threadIdx.x == 32 && rank == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
    int dev = 0;
    cudaSetDevice(dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    const int sms = prop.multiProcessorCount;
    long long *d_out;
    cudaMalloc(&d_out, sizeof(long long) * sms);
    const int smem = 226 * 1024;
    cudaFuncSetAttribute(mma_bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mma_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mma_bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mma_bench<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2048;
    printf("mode,N,bg_warps,pattern,cycles_per_mma_median,cycles_min,cycles_max\n");
    const int Ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
    for (int mode = 0; mode < 4; ++mode)
        for (int bg = 0; bg <= 8; bg += 8)
            for (int pattern = 0; pattern < 2; ++pattern)
                for (int N : Ns) {
                    if (mode == 3 && N < 32) continue;
                    cudaMemset(d_out, 0, sizeof(long long) * sms);
                    cudaError_t e = cudaSuccess;
                    const int grid = mode == 3 ? (sms / 2) * 2 : sms;
                    if (mode == 0) mma_bench<0><<<grid, 320, smem>>>(N, iters, bg, pattern, d_out);
                    else if (mode == 1) mma_bench<1><<<grid, 320, smem>>>(N, iters, bg, pattern, d_out);
                    else if (mode == 2) mma_bench<2><<<grid, 320, smem>>>(N, iters, bg, pattern, d_out);
                    else {
                        cudaLaunchConfig_t cfg{};
                        cfg.gridDim = dim3(grid);
                        cfg.blockDim = dim3(320);
                        cfg.dynamicSmemBytes = smem;
                        cudaLaunchAttribute at[1];
                        at[0].id = cudaLaunchAttributeClusterDimension;
                        at[0].val.clusterDim.x = 2;
                        at[0].val.clusterDim.y = 1;
                        at[0].val.clusterDim.z = 1;
                        cfg.attrs = at;
                        cfg.numAttrs = 1;
                        e = cudaLaunchKernelEx(&cfg, mma_bench<3>, N, iters, bg, pattern, d_out);
                    }
                    if (e == cudaSuccess) e = cudaGetLastError();
                    if (e == cudaSuccess) e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) {
                        printf("%d,%d,%d,%d,ERROR %s\n", mode, N, bg, pattern, cudaGetErrorString(e));
                        if (mode == 3) break;          // keep the other modes' results
                        return 1;
                    }
                    std::vector<long long> h(sms);
                    cudaMemcpy(h.data(), d_out, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
                    std::vector<double> v;
                    for (int i = 0; i < grid; ++i)
                        if (h[i] > 0) v.push_back((double)h[i] / (iters * 4.0));
                    std::sort(v.begin(), v.end());
                    printf("%d,%d,%d,%d,%.1f,%.1f,%.1f\n", mode, N, bg, pattern, v[v.size() / 2], v.front(), v.back());
                    fflush(stdout);
                }
    return 0;
}
