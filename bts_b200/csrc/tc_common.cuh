// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers shared by the tensor-core kernels (sm_100a).
#pragma once
#include "common.cuh"

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > 200000000u) __trap();   // watchdog: a protocol bug must fail loudly, not hang the box
    }
}
// same protocol, for waits that are expected to be long (an epilogue warp waiting for a whole tile of MMAs, the
// weight loader waiting for a free stage): back off between polls so the spinning lanes do not take issue slots
// from the producer warps (ncu round 1: 30 % of all issued instructions were try_wait / branch / yield).
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(32);
        if (++spins > 50000000u) __trap();
    }
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// TMA im2col load (cp.async.bulk.tensor ... .im2col -> SASS UTMALDG): `pixelsPerColumn` consecutive base pixels starting at
// tensor coordinate (w, h, n), each shifted by the filter offset (off_w, off_h), `channelsPerPixel` channels from c; pixels
// that fall outside the tensor (the convolution's zero padding) arrive as zeros.  Completes `bytes` on the mbarrier.
__device__ __forceinline__ void tma_im2col_4d(uint32_t dst, const void *tmap, int c, int w, int h, int n, uint32_t bar,
                                              uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6], "
        "{%7, %8};" ::"r"(dst),
        "l"(tmap), "r"(c), "r"(w), "r"(h), "r"(n), "r"(bar), "h"(off_w), "h"(off_h)
        : "memory");
}
// plain (tiled) 2-D tensor load: coordinates {c0 = innermost, c1}; out-of-range elements are zero-filled
__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const void *tmap, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// one lane of a converged warp (elect.sync): the lane that issues the single-thread tcgen05 instructions
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 |
// SBO = 1024 B (8 rows x 128 B) | version 1 (sm_100) | layout SWIZZLE_128B (=2)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major TF32 operands: the only layout the hardware accepts is SWIZZLE_128B_BASE32B (layout type 1,
// cute Swizzle<2,5,2>): rows of 128 B (32 fp32 along MN), atoms of 4 K-rows (512 B), the 32-byte chunk index
// (address bits [5,7)) XOR-ed with the K-row index mod 4 (address bits [7,9)).
// The tile is [MN chunk of 32 fp32][K rows][128 B]; LBO = byte stride between MN chunks, SBO = 512 (4 K rows).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
// byte offset of 16-byte unit `unit` (0..7) of K-row `row` inside one MN chunk under that swizzle
__device__ __forceinline__ uint32_t mn_swizzle_off(int row, int unit) {
    return (uint32_t)row * 128u + (uint32_t)(((((unit >> 1) ^ (row & 3)) << 1) | (unit & 1)) << 4);
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

struct __align__(16) F4 { float v[4]; };

// Division of n < 2^31 by a runtime constant d >= 1 as multiply-high + shift (the divisor's magic numbers are computed
// once on the host): with s = ceil(log2 d) and mul = ceil(2^(31+s) / d) the quotient is umulhi(n, mul) >> (s-1),
// exact for every n < 2^31.  Replaces the ~25-instruction I2F/MUFU.RCP/F2I sequences of `/` and `%`.
struct FastDiv {
    uint32_t mul, shr, d;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    f.mul = 0;
    f.shr = 0;
    if (d > 1) {
        uint32_t s = 0;
        while ((1ull << s) < d) ++s;                        // s = ceil(log2 d) >= 1
        f.mul = (uint32_t)(((1ull << (31 + s)) + d - 1) / d);
        f.shr = s - 1;
    }
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv &f) {
    return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr);
}

// Ampere-style async copies (LDGSTS): 16-byte / 4-byte global -> shared with zero-fill when src_bytes == 0
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int pending) {   // wait until <= pending groups are in flight
    switch (pending) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}


}  // namespace tc
