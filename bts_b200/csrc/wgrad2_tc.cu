// wgrad for narrow-output 3x3 convolutions (Cout <= 64: the DenseNet 3x3 dense-layer convs 192->48 and the full- /
// half-resolution decoder convs conv1/upconv1/conv2/upconv2) on tcgen05 + TMEM, 3xTF32 -- "shifted dY" formulation.
//
//   dW[co, ci, tap] = sum_q  x~[q, ci] * dY[q - off(tap), co]            q = INPUT pixel, off(tap) = tap*dil - pad
//
// wgrad_tc.cu puts the tap in the grid, so the 128-channel activation tile -- the expensive operand: loads, BN/ReLU
// pre-op, hi/lo split, 32 KB of shared-memory writes per k-block -- is produced 9 times, and with a narrow N the MMA
// is too short to hide it (dense 3x3: 38 TF/s, conv1: 12 TF/s).  Here ONE CTA owns all taps of a (128 ci, <=48 co)
// block: per k-block of 16 input pixels the activation tile is produced once and multiplied against the 9 shifted dY
// tiles (cheap: <= 48 channels each), accumulating into 9 column blocks of TMEM (9 x 48 = 432 of 512 columns).
// Same MN-major SWIZZLE_128B_BASE32B operand tiles and deterministic split-K partial layout as wgrad_tc.cu.
//
// Operand staging (round 2): the 9 shifted dY windows of a k-block overlap -- they are KH row segments of 16 + (KW-1)*dil
// consecutive output pixels -- and ncu showed the producers 58% stalled on the scoreboard of their global loads (one k-block
// of register prefetch, L1 hit rate 13%).  With TMA = true a loader lane copies the raw fp32 segments (and the raw x tile
// unless it is read through the nearest-neighbour up-sample) into a small landing ring, several k-blocks ahead
// (cp.async.bulk.tensor.2d, no swizzle, zero fill past the tensor ends); the producers then read shared memory, apply the
// border masks / pre-op, split hi/lo and write the swizzled operand tiles as before.
#include <cuda.h>

#include <cstring>

#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int BLOCK_CI = 128;
constexpr int KP = 16;                       // input pixels per k-block (2 k-groups of 8)
constexpr int CHUNK = KP * 128;              // one 32-channel chunk of a k-block: 2 KB
constexpr int A_BYTES = 4 * CHUNK;           // 8 KB (hi or lo)
constexpr int MAX_TAPS = 9;
constexpr int MAX_STAGES = 4;
constexpr int NUM_THREADS = 576;               // 2 control warps + 16 producer warps (4 quarters of 128 threads)
constexpr int PRODUCERS = 512;
constexpr int SMEM_BUDGET = 224 * 1024;
constexpr int MAX_RING = 4;                  // landing-ring slots (TMA staging)
constexpr int X_RAW_BYTES = KP * BLOCK_CI * 4;   // raw x tile of a k-block: 8 KB

struct W2Params {
    const float *x; long long xs;
    int B, Hs, Ws, up, Cin;
    int KH, KW, pad, dil;
    const float *pre_scale, *pre_shift;
    const float *dy; long long dys;
    int Cout, Hout, Wout, Hin, Win;
    int cg, nb;              // output channels per CTA (multiple of 16, <= 48) and their 32-channel chunks
    int nchunks;             // 32-slot chunks of the tap-packed dY tile: ceil(taps*cg/32)
    float *part;             // [splitK][taps][Cin][Cout]
    int splitK, kb_per_split, KBq;
    int Mq;                  // B*Hin*Win input pixels
    int stages, stage_bytes, precision;
    int legacy;              // 1: round-1 single-lane MMA issue loop (fallback switch)
    // TMA landing ring (TMA = true): slot = [raw x tile 8 KB (unless up) | KH segments of segw pixels x cg channels]
    int ring, slot_bytes, seg_bytes, segw, tox_max, slot_tx;
};

template <int PRE, bool UP, bool VEC, bool TMA>
__global__ void __launch_bounds__(NUM_THREADS, 1) wgrad2_tc_kernel(const W2Params p, const __grid_constant__ CUtensorMap tmx,
                                                                   const __grid_constant__ CUtensorMap tmd) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *sm = smem_raw + (base - smem_u32(smem_raw));
    const int taps = p.KH * p.KW;
    const int S = p.stages;
    const uint32_t ring_off = (uint32_t)S * (uint32_t)p.stage_bytes;
    const uint32_t pre_off = ring_off + (TMA ? (uint32_t)p.ring * (uint32_t)p.slot_bytes : 0u);
    float *s_scale = reinterpret_cast<float *>(sm + pre_off);
    float *s_shift = s_scale + BLOCK_CI;
    const uint32_t bar0 = base + pre_off + 2 * BLOCK_CI * 4;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + pre_off + 2 * BLOCK_CI * 4);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
    const uint32_t accum_full = bar0 + 8u * (2 * MAX_STAGES);
    auto rfull = [&](int j) { return bar0 + 8u * (2 * MAX_STAGES + 1 + j); };
    auto rempty = [&](int j) { return bar0 + 8u * (2 * MAX_STAGES + 1 + MAX_RING + j); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 1 + 2 * MAX_RING);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci_tile = blockIdx.x, cgi = blockIdx.y, split = blockIdx.z;
    const int co0 = cgi * p.cg;
    const int ncol = min(p.cg, ((p.Cout - co0 + 15) >> 4) << 4);     // live columns of this CTA (multiple of 16)
    const int nb = p.nb;
    // dY operand: the taps are packed DENSELY along N -- tap t owns the N slots [t*ncol, (t+1)*ncol) of one long MN-major
    // tile of ceil(taps*ncol/32) chunks -- so that a single tcgen05.mma (N up to 256) multiplies the activation tile
    // against many taps at once (see the MMA issuer).  hi and lo tiles follow each other.
    const int n_total = taps * ncol;
    const uint32_t b_half = (uint32_t)p.nchunks * CHUNK;
    const int kb0 = split * p.kb_per_split;
    int kb1 = kb0 + p.kb_per_split;
    if (kb1 > p.KBq) kb1 = p.KBq;
    const int nkb = kb1 > kb0 ? kb1 - kb0 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            mbar_init(full(s), PRODUCERS);
            mbar_init(empty(s), 1);
        }
        mbar_init(accum_full, 1);
        for (int j = 0; j < MAX_RING; ++j) {
            mbar_init(rfull(j), 1);
            mbar_init(rempty(j), PRODUCERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    if (PRE >= 2) {
        for (int c = threadIdx.x; c < BLOCK_CI; c += NUM_THREADS) {
            const int ch = ci_tile * BLOCK_CI + c;
            s_scale[c] = ch < p.Cin ? p.pre_scale[ch] : 0.f;
            s_shift[c] = ch < p.Cin ? p.pre_shift[ch] : 0.f;
        }
    }
    // zero every operand stage once: dead channel chunks / dead units are never written afterwards
    for (int i = threadIdx.x; i < S * p.stage_bytes / 16; i += NUM_THREADS) st_shared_v4(base + i * 16, 0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (TMA && warp == 0) {
        // ---- loader: one lane keeps the landing ring `ring` k-blocks ahead of the producers
        if (lane == 0) {
            if (!UP) tma_prefetch_desc(&tmx);
            tma_prefetch_desc(&tmd);
            int j = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nkb; ++it) {
                mbar_wait(rempty(j), ph ^ 1);
                const uint32_t dst = base + ring_off + (uint32_t)j * (uint32_t)p.slot_bytes;
                const int q0 = (kb0 + it) * KP;
                mbar_arrive_expect_tx(rfull(j), (uint32_t)p.slot_tx);
                uint32_t seg = dst;
                if (!UP) {
                    tma_tile_2d(dst, &tmx, ci_tile * BLOCK_CI, q0, rfull(j));
                    seg += X_RAW_BYTES;
                }
                for (int ky = 0; ky < p.KH; ++ky)     // output pixels q - (ky*dil - pad)*W - tox, tox <= tox_max
                    tma_tile_2d(seg + (uint32_t)ky * (uint32_t)p.seg_bytes, &tmd, co0,
                                q0 - (ky * p.dil - p.pad) * p.Wout - p.tox_max, rfull(j));
                if (++j == p.ring) { j = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: whole warp in the loop, one elected lane issues (see conv_tc.cu: the round-1 single-lane loops paid
        // ~4 cycles x hundreds of issue-warp instructions per k-block).  With the taps packed along N the three products
        // of a k-group are 3 x ceil(taps*ncol/256) instructions instead of 3 x taps.
        if (p.legacy) {
        if (lane == 0) {
            // A tcgen05.mma with M=128, K=8 costs ~115-130 cycles whatever N is (round-1 measurement: 54 per-tap MMAs
            // per 16-pixel k-block = 5.5k cycles, exactly the observed k-block time).  With the taps packed along N the
            // three products of a k-group are 3 x ceil(taps*ncol/256) instructions instead of 3 x taps.
            const int n0 = n_total < 256 ? n_total : 256, n1 = n_total - n0;        // multiples of 16
            const uint32_t idesc0 = make_idesc(BLOCK_CI, n0, 1, 1);
            const uint32_t idesc1 = make_idesc(BLOCK_CI, n1 > 0 ? n1 : 16, 1, 1);
            const uint64_t dah0 = make_desc_mn(base, CHUNK), dal0 = make_desc_mn(base + A_BYTES, CHUNK);
            const uint64_t dbh0 = make_desc_mn(base + 2 * A_BYTES, CHUNK), dbl0 = make_desc_mn(base + 2 * A_BYTES + b_half, CHUNK);
            const uint64_t piece1 = (uint64_t)((8 * CHUNK) >> 4), stage_step = (uint64_t)(p.stage_bytes >> 4);
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nkb; ++it) {
                mbar_wait(full(s), ph);
                tc_fence_after();
                const uint64_t so = (uint64_t)s * stage_step;
#pragma unroll
                for (int kg = 0; kg < KP / 8; ++kg) {
                    const uint64_t o = so + (uint64_t)(kg * (1024 >> 4));   // start-address field: + 1024 B per k-group
                    const uint32_t accumulate = (it | kg) != 0;
                    if (p.precision == 0) {
                        umma_tf32(tmem_base, dal0 + o, dbh0 + o, idesc0, accumulate);
                        umma_tf32(tmem_base, dah0 + o, dbl0 + o, idesc0, 1);
                        umma_tf32(tmem_base, dah0 + o, dbh0 + o, idesc0, 1);
                        if (n1 > 0) {
                            umma_tf32(tmem_base + 256, dal0 + o, dbh0 + o + piece1, idesc1, accumulate);
                            umma_tf32(tmem_base + 256, dah0 + o, dbl0 + o + piece1, idesc1, 1);
                            umma_tf32(tmem_base + 256, dah0 + o, dbh0 + o + piece1, idesc1, 1);
                        }
                    } else {
                        umma_tf32(tmem_base, dah0 + o, dbh0 + o, idesc0, accumulate);
                        if (n1 > 0) umma_tf32(tmem_base + 256, dah0 + o, dbh0 + o + piece1, idesc1, accumulate);
                    }
                }
                umma_commit(empty(s));
                if (++s == S) { s = 0; ph ^= 1; }
            }
            umma_commit(accum_full);
        }
        } else {
        const int n0 = n_total < 256 ? n_total : 256, n1 = n_total - n0;        // multiples of 16
        const uint32_t idesc0 = make_idesc(BLOCK_CI, n0, 1, 1);
        const uint32_t idesc1 = make_idesc(BLOCK_CI, n1 > 0 ? n1 : 16, 1, 1);
        const uint64_t dah0 = make_desc_mn(base, CHUNK), dal0 = make_desc_mn(base + A_BYTES, CHUNK);
        const uint64_t dbh0 = make_desc_mn(base + 2 * A_BYTES, CHUNK), dbl0 = make_desc_mn(base + 2 * A_BYTES + b_half, CHUNK);
        const uint64_t piece1 = (uint64_t)((8 * CHUNK) >> 4), stage_step = (uint64_t)(p.stage_bytes >> 4);
        const bool leader = elect_one();
        const bool two = n1 > 0, single = p.precision != 0;
        int s = 0;
        uint32_t ph = 0;
        uint64_t so = 0;
        uint32_t bfull = full(0), bempty = empty(0);
        for (int it = 0; it < nkb; ++it) {
            mbar_wait(bfull, ph);
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int kg = 0; kg < KP / 8; ++kg) {
                    const uint64_t o = so + (uint64_t)(kg * (1024 >> 4));   // start-address field: + 1024 B per k-group
                    const uint32_t accumulate = kg != 0 ? 1u : (uint32_t)(it != 0);
                    if (!single) {
                        umma_tf32(tmem_base, dal0 + o, dbh0 + o, idesc0, accumulate);
                        umma_tf32(tmem_base, dah0 + o, dbl0 + o, idesc0, 1);
                        umma_tf32(tmem_base, dah0 + o, dbh0 + o, idesc0, 1);
                        if (two) {
                            umma_tf32(tmem_base + 256, dal0 + o, dbh0 + o + piece1, idesc1, accumulate);
                            umma_tf32(tmem_base + 256, dah0 + o, dbl0 + o + piece1, idesc1, 1);
                            umma_tf32(tmem_base + 256, dah0 + o, dbh0 + o + piece1, idesc1, 1);
                        }
                    } else {
                        umma_tf32(tmem_base, dah0 + o, dbh0 + o, idesc0, accumulate);
                        if (two) umma_tf32(tmem_base + 256, dah0 + o, dbh0 + o + piece1, idesc1, accumulate);
                    }
                }
                umma_commit(bempty);
                if (it == nkb - 1) umma_commit(accum_full);
            }
            __syncwarp();
            so += stage_step; bfull += 8u; bempty += 8u;
            if (++s == S) { s = 0; ph ^= 1; so = 0; bfull = full(0); bempty = empty(0); }
        }
        if (nkb == 0 && leader) umma_commit(accum_full);
        }   // lean issue loop
    } else if (warp >= 2) {
        // 16 producer warps (round 1 ran 8: ncu showed ~9 cycles between a warp's instructions and 2.5 warps per
        // scheduler -- the per-thread instruction stream, not the tensor pipe, set the k-block time).  Quarter q of the
        // producers owns the x chunk q and the dY taps t == q (mod 4): at most 1 + 6 sixteen-byte units per thread.
        const int pt = threadIdx.x - 64;           // 0..511
        const int unit = pt & 7;                   // 16-byte unit of the 128-byte row
        const int row = (pt >> 3) & 15;            // pixel row of the 16-pixel k-block
        const int half = pt >> 7;                  // quarter 0..3: A chunk `half`;  dY: taps t == half (mod 4)
        constexpr bool AFF = PRE >= 2;
        constexpr bool RELU = (PRE & 1) != 0;
        const uint32_t roff = mn_swizzle_off(row, unit);
        const int xs = (int)p.xs, dys = (int)p.dys;
        const int cbx = ci_tile * BLOCK_CI + half * 32 + unit * 4;    // first x channel of this thread (chunk `half`)
        const int cbd = co0 + unit * 4;                               // first dY channel (chunk 0)
        const float *__restrict__ xg = p.x;
        const float *__restrict__ dg = p.dy;
        float sc[1][4], sh[1][4];
        if (AFF) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sc[0][e] = s_scale[half * 32 + unit * 4 + e];
                sh[0][e] = s_shift[half * 32 + unit * 4 + e];
            }
        }
        // input-pixel coordinates of this thread's row, advanced by 16 pixels per k-block (no divisions in the loop)
        int qx, qy, qb;
        {
            const int q = kb0 * KP + row;
            qx = q % p.Win;
            const int r = q / p.Win;
            qy = r % p.Hin;
            qb = r / p.Hin;
        }
        constexpr int NBU = 6;                     // dY units per thread per k-block: <= 3 taps x 2 chunks
        constexpr int NAU = 1;                     // x units per thread per k-block
        // Unit slots (j, ch), j = 0..2, ch = 0..1.  Multi-tap layers: tap t = half + 4j, 32-channel chunk ch of <= 2.
        // Single-tap (1x1) layers: tap 0, chunk half + 4*(2j + ch) of <= 8 (an output tile up to 256 channels wide).
        const bool single = taps == 1;
        auto unit_tap = [&](int j) { return single ? 0 : half + 4 * j; };
        auto unit_chunk = [&](int j, int ch) { return single ? half + 4 * (2 * j + ch) : ch; };
        // tap offsets of this thread's taps, computed once: no divisions in the k-loop
        int toy[NBU / 2], tox[NBU / 2], tky[NBU / 2];
#pragma unroll
        for (int j = 0; j < NBU / 2; ++j) {
            const int t = unit_tap(j);
            const int ky = t / p.KW, kx = t - ky * p.KW;
            toy[j] = ky * p.dil - p.pad;
            tox[j] = kx * p.dil - p.pad;
            tky[j] = ky < p.KH ? ky : 0;
        }
        // shared-memory offsets of this thread's dY units in the densely packed tile: N slot = t*ncol + channel
        uint32_t boff[NBU];
        bool blive[NBU];
#pragma unroll
        for (int j = 0; j < NBU / 2; ++j)
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                const int t = unit_tap(j), cc = unit_chunk(j, ch);
                const int c = unit * 4 + cc * 32;                         // channel within this CTA's group
                const int n = t * ncol + c;
                blive[j * 2 + ch] = t < taps && cc < nb && c < ncol;
                boff[j * 2 + ch] = (uint32_t)(n >> 5) * CHUNK + mn_swizzle_off(row, (n & 31) >> 2);
            }
        struct Cursor { int qx, qy, qb; };
        Cursor cur = {qx, qy, qb};
        auto advance = [&](Cursor &c) {
            c.qx += KP;
            while (c.qx >= p.Win) {
                c.qx -= p.Win;
                if (++c.qy == p.Hin) { c.qy = 0; ++c.qb; }
            }
        };
        // ---- x~ tile through the load path: this thread's pixel, channels of its chunk
        auto load_x = [&](int it, const Cursor &c_, F4(&va)[NAU], bool &okx) {
            const int q = (kb0 + it) * KP + row;
            okx = q < p.Mq;
            const int sy = UP ? (c_.qy >> 1) : c_.qy, sx = UP ? (c_.qx >> 1) : c_.qx;
            const int xoff = ((c_.qb * p.Hs + sy) * p.Ws + sx) * xs;
#pragma unroll
            for (int ch = 0; ch < NAU; ++ch) {
                const int c = cbx + ch * 32;
                const bool live = okx && c < p.Cin;
                if (VEC) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live) v = __ldg(reinterpret_cast<const float4 *>(xg + xoff + c));
                    if (c + 3 >= p.Cin) {
                        if (c + 1 >= p.Cin) v.y = 0.f;
                        if (c + 2 >= p.Cin) v.z = 0.f;
                        v.w = 0.f;
                    }
                    va[ch].v[0] = v.x; va[ch].v[1] = v.y; va[ch].v[2] = v.z; va[ch].v[3] = v.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float v = 0.f;
                        if (live && c + e < p.Cin) v = __ldg(xg + xoff + c + e);
                        va[ch].v[e] = v;
                    }
                }
            }
        };
        // ---- shifted dY tiles through the load path: taps t = half, half+4, ...; output pixel (qy - dy_t, qx - dx_t)
        auto load_d = [&](const Cursor &c_, bool okx, F4(&vb)[NBU]) {
#pragma unroll
            for (int j = 0; j < NBU / 2; ++j) {
                const int t = unit_tap(j);
                const int py = c_.qy - toy[j], px = c_.qx - tox[j];
                const bool okd = okx && t < taps && (unsigned)py < (unsigned)p.Hout && (unsigned)px < (unsigned)p.Wout;
                const int doff = ((c_.qb * p.Hout + py) * p.Wout + px) * dys;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const int c = cbd + unit_chunk(j, ch) * 32;
                    const bool live = okd && blive[j * 2 + ch] && c < p.Cout;
                    F4 &dst = vb[j * 2 + ch];
                    if (VEC) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (live) v = __ldg(reinterpret_cast<const float4 *>(dg + doff + c));
                        if (c + 3 >= p.Cout) {
                            if (c + 1 >= p.Cout) v.y = 0.f;
                            if (c + 2 >= p.Cout) v.z = 0.f;
                            v.w = 0.f;
                        }
                        dst.v[0] = v.x; dst.v[1] = v.y; dst.v[2] = v.z; dst.v[3] = v.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float v = 0.f;
                            if (live && c + e < p.Cout) v = __ldg(dg + doff + c + e);
                            dst.v[e] = v;
                        }
                    }
                }
            }
        };
        auto load = [&](int it, F4(&va)[NAU], F4(&vb)[NBU], bool &okx) {
            load_x(it, cur, va, okx);
            load_d(cur, okx, vb);
            advance(cur);
        };
        // ---- the same operands out of a landing-ring slot (TMA): raw x tile [16 px][128 ch], then KH segments [segw px][cg ch]
        auto ring_read = [&](int it, uint32_t slot, F4(&va)[NAU], F4(&vb)[NBU], bool &okx) {
            const int q = (kb0 + it) * KP + row;
            okx = q < p.Mq;
            uint32_t segb = slot;
            if (!UP) {
                const float4 v = ld_shared_v4(slot + (uint32_t)((row * BLOCK_CI + half * 32 + unit * 4) * 4));
                va[0].v[0] = v.x; va[0].v[1] = v.y; va[0].v[2] = v.z; va[0].v[3] = v.w;    // channels >= Cin, pixels >= Mq: zero fill
                segb += X_RAW_BYTES;
            }
#pragma unroll
            for (int j = 0; j < NBU / 2; ++j) {
                const int t = unit_tap(j);
                const int py = cur.qy - toy[j], px = cur.qx - tox[j];
                const bool okd = okx && t < taps && (unsigned)py < (unsigned)p.Hout && (unsigned)px < (unsigned)p.Wout;
                const uint32_t rowb = segb + (uint32_t)tky[j] * (uint32_t)p.seg_bytes +
                                      (uint32_t)((row + p.tox_max - tox[j]) * p.cg + unit * 4) * 4u;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    F4 &dst = vb[j * 2 + ch];
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (okd && blive[j * 2 + ch]) v = ld_shared_v4(rowb + (uint32_t)unit_chunk(j, ch) * 128u);
                    dst.v[0] = v.x; dst.v[1] = v.y; dst.v[2] = v.z; dst.v[3] = v.w;
                }
            }
            advance(cur);
        };
        auto split_store = [&](uint32_t hi_addr, uint32_t lo_addr, const F4 &v) {
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a = v.v[e];
                const float hh = __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
                hi[e] = hh;
                lo[e] = a - hh;
            }
            st_shared_v4(hi_addr, hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(lo_addr, lo[0], lo[1], lo[2], lo[3]);
        };
        int st_s = 0;
        uint32_t st_ph = 0;
        auto store = [&](int it, F4(&va)[NAU], F4(&vb)[NBU], bool okx) {
            const int s = st_s;
            const uint32_t ph = st_ph;
            if (++st_s == S) { st_s = 0; st_ph ^= 1; }
            if (PRE != 0) {
#pragma unroll
                for (int ch = 0; ch < NAU; ++ch)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float a = va[ch].v[e];
                        if (AFF) {
                            a = fmaf(a, sc[ch][e], sh[ch][e]);
                            if (RELU) a = fmaxf(a, 0.f);
                            a = okx ? a : 0.f;
                        } else {
                            a = fmaxf(a, 0.f);
                        }
                        va[ch].v[e] = a;
                    }
            }
            mbar_wait(empty(s), ph ^ 1);
            const uint32_t a_hi = base + (uint32_t)s * (uint32_t)p.stage_bytes, a_lo = a_hi + A_BYTES;
            const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + b_half;
#pragma unroll
            for (int ch = 0; ch < NAU; ++ch) {
                const uint32_t o = (uint32_t)(half + ch) * CHUNK + roff;
                split_store(a_hi + o, a_lo + o, va[ch]);
            }
#pragma unroll
            for (int j = 0; j < NBU; ++j)
                if (blive[j]) split_store(b_hi + boff[j], b_lo + boff[j], vb[j]);
            fence_proxy_async();
            mbar_arrive(full(s));
        };
        if constexpr (!TMA) {
            F4 a0[NAU], a1[NAU], b0[NBU], b1[NBU];
            bool k0 = false, k1 = false;
            int it = 0;
            if (it < nkb) load(it, a0, b0, k0);
            for (; it < nkb; it += 2) {
                const bool more = it + 1 < nkb;
                if (more) load(it + 1, a1, b1, k1);
                store(it, a0, b0, k0);
                if (more) {
                    if (it + 2 < nkb) load(it + 2, a0, b0, k0);
                    store(it + 1, a1, b1, k1);
                }
            }
        } else {
            // consume the landing ring; an up-sampled x still comes through the load path, one k-block ahead (own cursor)
            F4 xc[NAU], xn[NAU], va[NAU], vb[NBU];
            bool kc = false, kn = false, okx = false;
            Cursor pre = cur;
            if (UP && nkb > 0) { load_x(0, pre, xc, kc); advance(pre); }
            int rj = 0;
            uint32_t rph = 0;
            for (int it = 0; it < nkb; ++it) {
                if (UP && it + 1 < nkb) { load_x(it + 1, pre, xn, kn); advance(pre); }
                mbar_wait(rfull(rj), rph);
                ring_read(it, base + ring_off + (uint32_t)rj * (uint32_t)p.slot_bytes, va, vb, okx);
                mbar_arrive(rempty(rj));               // the slot's values are in registers: the loader may refill it
                if (++rj == p.ring) { rj = 0; rph ^= 1; }
                if (UP) {
                    store(it, xc, vb, okx);
#pragma unroll
                    for (int ch = 0; ch < NAU; ++ch) xc[ch] = xn[ch];
                    kc = kn;
                } else {
                    store(it, va, vb, okx);
                }
            }
        }

        // ---- epilogue: TMEM lane = input channel; column block t*ncol.. = tap t; taps split between the two halves
        mbar_wait(accum_full, 0);
        tc_fence_after();
        const int q4 = warp & 3;
        const int ci = ci_tile * BLOCK_CI + q4 * 32 + lane;
        const bool ovec = (p.Cout & 3) == 0 && ((((uintptr_t)p.part) & 15) == 0) && ((co0 & 3) == 0);
        // multi-tap: the taps are split between the four quarters; single-tap: the 8-column blocks of the one tap are
        for (int t = single ? 0 : half; t < taps; t += single ? 1 : 4) {
            float *prow = p.part + (((long long)split * taps + t) * p.Cin + (ci < p.Cin ? ci : 0)) * p.Cout + co0;
            for (int cc = single ? half * 8 : 0; cc < ncol; cc += single ? 32 : 8) {
                uint32_t r[8];
                tmem_ld8(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(t * ncol + cc), r);
                tmem_ld_wait();
                if (ci < p.Cin) {
#pragma unroll
                    for (int e4 = 0; e4 < 8; e4 += 4) {
                        const int c = co0 + cc + e4;
                        if (ovec && c + 3 < p.Cout) {
                            *reinterpret_cast<float4 *>(prow + cc + e4) =
                                make_float4(nkb ? __uint_as_float(r[e4]) : 0.f, nkb ? __uint_as_float(r[e4 + 1]) : 0.f,
                                            nkb ? __uint_as_float(r[e4 + 2]) : 0.f, nkb ? __uint_as_float(r[e4 + 3]) : 0.f);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (c + e < p.Cout) prow[cc + e4 + e] = nkb ? __uint_as_float(r[e4 + e]) : 0.f;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int bts_issue_legacy();

// co-group width for the shifted-dY kernel: multiple of 16, taps*cg <= 512 TMEM columns, groups as even as possible
int bts_wgrad2_cg(int Cout, int taps) {
    int cap = (512 / taps) / 16 * 16;
    if (taps == 1) {
        if (cap > 256) cap = 256;              // 1x1: one tile up to 256 output channels wide (one tcgen05.mma)
    } else if (cap > 48) {
        cap = 48;
    }
    if (cap < 16) return 0;
    const int groups = (Cout + cap - 1) / cap;
    int cg = ((Cout + groups - 1) / groups + 15) / 16 * 16;
    if (cg > cap) cg = cap;
    return cg;
}

// measured (B200, K16 shapes, profiles/r02_w2_sweep.txt): with the TMA landing ring it beats the tap-in-grid kernel on
// every narrow-output 3x3 layer down to the 22x44 maps of block 3 (40 vs 71 us with >= 8 k-blocks per split-K CTA and one
// wave of CTAs); on the 11x22 maps of block 4 (3.9k pixels) the two tie and the tap-in-grid kernel stays
static long long g_w2_min_pixels = 12000;
static int g_w2_pointwise = 1;               // 1x1 layers (64 < Cout <= 256) on this kernel too
extern "C" int bts_wgrad2_set_pointwise(int on) { g_w2_pointwise = on ? 1 : 0; return 0; }
static int g_w2_min_kb = 8;                  // fewest 16-pixel k-blocks a split-K CTA gets
extern "C" int bts_wgrad2_set_min_pixels(long long n) { g_w2_min_pixels = n < 0 ? 12000 : n; return 0; }
extern "C" int bts_wgrad2_set_min_kblocks(int n) { g_w2_min_kb = n < 1 ? 8 : n; return 0; }

bool bts_wgrad2_eligible(int Cout, int KH, int KW, int stride, long long Mq) {
    const int taps = KH * KW;
    if (stride != 1 || Mq < g_w2_min_pixels || taps > MAX_TAPS || bts_wgrad2_cg(Cout, taps) <= 0) return false;
    if (taps == 1) return g_w2_pointwise && Cout > 64 && Cout <= 256;   // 1x1 layers with one <= 256-wide output tile
    return Cout <= 64;
}

void bts_wgrad2_plan(int B, int Hin, int Win, int Cin, int Cout, int KH, int KW, int *splitK) {
    const int taps = KH * KW;
    const int cg = bts_wgrad2_cg(Cout, taps);
    const long long Mq = (long long)B * Hin * Win;
    const long long KBq = (Mq + KP - 1) / KP;
    const long long tiles = (long long)((Cin + BLOCK_CI - 1) / BLOCK_CI) * ((Cout + cg - 1) / cg);
    const int sms = bts_num_sms();
    long long max_split = (KBq + g_w2_min_kb - 1) / g_w2_min_kb;
    if (max_split < 1) max_split = 1;
    if (max_split > 512) max_split = 512;
    long long split = 1;
    double best = -1.0;
    for (long long sp = 1; sp <= max_split; ++sp) {
        const long long ctas = tiles * sp;
        const long long waves = (ctas + sms - 1) / sms;
        if (waves > 2 && sp > 1) break;
        const double eff = (double)ctas / (double)(waves * sms);
        if (eff > best + 1e-9) { best = eff; split = sp; }      // ties -> fewer splits: less partial traffic for the reduce
    }
    *splitK = (int)split;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// [pixels][channels] fp32 view with a pixel stride (channel slices of slabs are fine); box = box_c channels x box_p pixels
static bool make_rows_map(CUtensorMap *map, const float *base, long long pixel_stride, int channels, long long pixels, int box_c,
                          int box_p) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || box_c > 256 || box_p > 256) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)channels, (cuuint64_t)pixels};
    const cuuint64_t gstr[1] = {(cuuint64_t)pixel_stride * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_p};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstr, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int g_w2_tma = 1;       // 1 (default): landing ring where eligible; 0: producers load from global memory
extern "C" int bts_wgrad2_set_tma(int on) { g_w2_tma = on ? 1 : 0; return 0; }

int bts_wgrad2_launch(const float *x, long long xs, int B, int Hs, int Ws, int up, int Cin, int KH, int KW, int pad,
                      int dil, const float *pre_scale, const float *pre_shift, int pre_relu, const float *dy,
                      long long dys, int Cout, int Hout, int Wout, float *workspace, int splitK, int precision,
                      cudaStream_t st) {
    W2Params p;
    const int taps = KH * KW;
    p.x = x; p.xs = xs; p.B = B; p.Hs = Hs; p.Ws = Ws; p.up = up; p.Cin = Cin;
    p.KH = KH; p.KW = KW; p.pad = pad; p.dil = dil;
    p.pre_scale = pre_scale; p.pre_shift = pre_shift;
    p.dy = dy; p.dys = dys; p.Cout = Cout; p.Hout = Hout; p.Wout = Wout;
    p.Hin = up ? 2 * Hs : Hs; p.Win = up ? 2 * Ws : Ws;
    p.cg = bts_wgrad2_cg(Cout, taps);
    p.nb = (p.cg + 31) / 32;
    p.part = workspace; p.splitK = splitK;
    const long long Mq = (long long)B * p.Hin * p.Win;
    if (Mq > 0x7ffffff0LL) return BTS_EINVAL;
    p.Mq = (int)Mq;
    p.KBq = (int)((Mq + KP - 1) / KP);
    p.kb_per_split = (p.KBq + splitK - 1) / splitK;
    p.nchunks = (taps * p.cg + 31) / 32;
    p.stage_bytes = 2 * A_BYTES + 2 * p.nchunks * CHUNK;
    p.stages = SMEM_BUDGET / p.stage_bytes;
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    if (p.stages < 2) return BTS_EINVAL;
    p.precision = precision;
    p.legacy = bts_issue_legacy();
    const int pre = (pre_scale ? 2 : 0) | (pre_relu ? 1 : 0);
    const bool vec = bts_aligned16(x) && (xs % 4 == 0) && bts_aligned16(dy) && (dys % 4 == 0);
    // ---- landing ring (TMA): 'same' convolutions with <= 3 kernel rows; 3 operand stages if >= 3 ring slots still fit, else 2
    CUtensorMap tmx, tmd;
    memset(&tmx, 0, sizeof(tmx));
    memset(&tmd, 0, sizeof(tmd));
    bool tma = false;
    p.ring = p.slot_bytes = p.seg_bytes = p.segw = p.tox_max = p.slot_tx = 0;
    if (g_w2_tma && vec && KH <= 3 && Hout == p.Hin && Wout == p.Win) {
        p.segw = KP + (KW - 1) * dil;
        p.tox_max = (KW - 1) * dil - pad;
        p.seg_bytes = (p.segw * p.cg * 4 + 127) / 128 * 128;
        p.slot_bytes = (up ? 0 : X_RAW_BYTES) + KH * p.seg_bytes;
        p.slot_tx = (up ? 0 : X_RAW_BYTES) + KH * p.segw * p.cg * 4;
        int st_try = p.stages > 3 ? 3 : p.stages;
        for (; st_try >= 2 && !tma; --st_try) {
            const int ring = (SMEM_BUDGET - st_try * p.stage_bytes) / p.slot_bytes;
            if (ring >= 3 || (st_try == 2 && ring >= 2)) {
                p.stages = st_try;
                p.ring = ring > MAX_RING ? MAX_RING : ring;
                tma = true;
            }
        }
        if (tma && p.segw <= 256)
            tma = make_rows_map(&tmd, dy, dys, Cout, (long long)B * Hout * Wout, p.cg, p.segw) &&
                  (up || make_rows_map(&tmx, x, xs, Cin, Mq, BLOCK_CI, KP));
        else
            tma = false;
        if (!tma) {
            p.ring = 0;
            p.stages = SMEM_BUDGET / p.stage_bytes > MAX_STAGES ? MAX_STAGES : SMEM_BUDGET / p.stage_bytes;
        }
    }
    const int smem = p.stages * p.stage_bytes + p.ring * p.slot_bytes + 2 * BLOCK_CI * 4 + 256 + 1024;
    dim3 grid((Cin + BLOCK_CI - 1) / BLOCK_CI, (Cout + p.cg - 1) / p.cg, splitK);
    cudaError_t err = cudaSuccess;
#define BTS_LAUNCH(PRE, UP, VEC, TMA)                                                                                 \
    do {                                                                                                              \
        static int attr_smem_[BTS_MAX_DEVICES] = {}; int &attr_smem = attr_smem_[bts_cur_device()];                   \
        if (attr_smem < smem) {                                                                                       \
            err = cudaFuncSetAttribute(wgrad2_tc_kernel<PRE, UP, VEC, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       SMEM_BUDGET + 2 * BLOCK_CI * 4 + 256 + 1024);                                  \
            if (err != cudaSuccess) return (int)err;                                                                  \
            attr_smem = SMEM_BUDGET + 2 * BLOCK_CI * 4 + 256 + 1024;                                                  \
        }                                                                                                             \
        wgrad2_tc_kernel<PRE, UP, VEC, TMA><<<grid, NUM_THREADS, smem, st>>>(p, tmx, tmd);                            \
    } while (0)
#define BTS_DISPATCH_UV(PRE)                                                                     \
    do {                                                                                         \
        if (tma) { if (p.up) BTS_LAUNCH(PRE, true, true, true); else BTS_LAUNCH(PRE, false, true, true); } \
        else if (p.up) { if (vec) BTS_LAUNCH(PRE, true, true, false); else BTS_LAUNCH(PRE, true, false, false); }   \
        else { if (vec) BTS_LAUNCH(PRE, false, true, false); else BTS_LAUNCH(PRE, false, false, false); }      \
    } while (0)
    switch (pre) {
        case 0: BTS_DISPATCH_UV(0); break;
        case 1: BTS_DISPATCH_UV(1); break;
        case 2: BTS_DISPATCH_UV(2); break;
        default: BTS_DISPATCH_UV(3); break;
    }
#undef BTS_DISPATCH_UV
#undef BTS_LAUNCH
    BTS_LAUNCH_CHECK();
    return 0;
}
