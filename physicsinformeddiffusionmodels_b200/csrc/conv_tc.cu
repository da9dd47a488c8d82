// Convolutions of the U-Net on NHWC bf16 activations as implicit GEMMs on the 5th-generation tensor cores (tcgen05),
// operands staged by TMA, fp32 accumulators in tensor memory (TMEM).
//
//   y[b,h,w,n] = sum_{r,q,c} x[b, s*h+r-pad, s*w+q-pad, c] * Wp[n][(r*KW+q)*Cin + c] (+ bias[n]) (+ residual[b,h,w,n])
//
// One kernel covers every layer of the network and its dgrad: 3x3 / 1x1 / 7x7 stride-1 layers (ResnetBlock convs,
// res_conv, to_qkv, to_out, stem: reference unet_model.py:227,253,275,279,453), the 4x4/stride-2 down-sampling conv
// (:197, TMA elementStrides) and the 4x4/stride-2 transposed conv (:163, four output-parity classes).
//
// Tiling.  GEMM M = 128 output pixels, N = BN output channels, K = taps * Cin walked in K-steps.
//   plain mode     : the pixel tile is a TN x TH x TW box (TW = W); one K-step = one tap x BK channels; the A operand
//                    is ONE 4-D TMA box {BK, TW, TH, TN} at spatial offset (r-pad, q-pad) -- the im2col gather, the zero
//                    padding (TMA out-of-bounds fill) and the 128B/64B swizzle all happen in the copy engine.
//   row-group mode : (stride-1 KxK layers whose image splits into 16x8 tiles) one K-step = one kernel COLUMN q x BK
//                    channels; its A box holds (16 + KH - 1) x 8 pixels and the KH kernel rows are row-shifted views of
//                    that one shared-memory tile (a shift of 8 pixels = one swizzle period), so L2 -> SM operand
//                    traffic drops from KH*KW to KW*(16+KH-1)/16 tiles per output tile.
//   The B operand (K-major packed weights) is a 2-D TMA box per tap, or resident in shared memory for the whole kernel
//   when the layer has a single n-tile.  Both land in the canonical K-major swizzled layout that the UMMA descriptors
//   expect: no thread ever touches the operands.
//
// Persistent, warp-specialised (256 threads, one CTA per SM, tiles walked with stride gridDim.x):
//   warps 0,2,3  TMA producers (one elected lane each, K-step i belongs to producer i % 3), <= 12-stage mbarrier ring
//                that runs across tile boundaries
//   warp 1       MMA issuer: ONE thread, tcgen05.mma.cta_group::1.kind::f16 M128 x BN x K16, descriptors advanced by
//                integer adds; tcgen05.commit releases ring stages / publishes the accumulator
//   warp 2       also allocates TMEM (2 x BN columns: the epilogue of tile i overlaps the mainloop of tile i+1)
//   warps 4-7    epilogue: tcgen05.ld -> +bias +residual -> GroupNorm sum / sum-of-squares (optional) -> bf16 ->
//                XOR-swizzled smem transpose -> 64/128-byte coalesced row-segment stores
// What bounds it (clock64 traces, scripts/trace_conv.py; profiles/r02_conv_splitk_trace.txt): the K-step cadence of the
// main loop is set by the copy engine, not by the tensor pipe -- ~427 cycles per (128-row im2col box + weight box) at the
// 8x8 level whatever the n-tile width, ~770 cycles per 144-row halo box of 64-byte rows at Cin = 32 (5.3 cycles per row) --
// while the four to six MMAs of a K-step issue in 190-290 cycles; a launch costs >= 4.3 us (1x1 conv at 8x8).
// Two alternatives were built and measured in round 2 and are NOT in this file (git history: c60b6e5, c86e22a): split-K
// over a thread-block cluster with a DSMEM exchange of the partial accumulators (main loop 3.3x shorter, but the cluster
// barrier absorbs ~6.5 k cycles of CTA start skew and the exchange 5-12 k: 11.9 -> 12.8 us per 8x8x256 layer) and
// cp.async producer warps for the A operand instead of TMA (2100-2400 cycles per K-step: 12.8 -> 44-49 us).  Carrying
// them as opt-in paths cost the default path 3 % (register pressure in the issue loops), so they were removed.
#include "common.cuh"
#include "pidm.h"
#include <cuda.h>
#include <stdlib.h>

namespace pidm {

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 256;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(tc_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            tc_smem_u32(dst)),
        "l"(map), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            tc_smem_u32(dst)),
        "l"(map), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// K-major, swizzled UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major, 1), [32,46) SBO>>4 = 8 rows * swizzle span,
//   [46,48) version = 1, [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
template <int SWIZZLE_BYTES>
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    constexpr uint64_t layout = SWIZZLE_BYTES == 128 ? 2 : (SWIZZLE_BYTES == 64 ? 4 : 6);
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * SWIZZLE_BYTES) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

struct TcClass {             // one output-parity class of a transposed (stride-2) gather: <= 16 taps
    int n_taps, off_h, off_w;
    signed char dh[16], dw[16];
    short ktap[16];          // tap index into the packed weights (k offset = ktap * Cin)
};

struct TcParams {
    int B, GH, GW;           // pixel grid that forms the GEMM M dimension (TW == GW)
    int Cin, Cout;
    int mode;                // 0: regular conv, taps by formula (KH, KW, pad), input sampled with in_stride
                             // 1: transposed gather, per-class tap tables, output scattered with out_scale / offsets
    int KH, KW, pad;
    int in_stride, out_scale;
    int Ho, Wo;              // spatial size of the output tensor
    int TW, TH, TN;          // pixel box; TW*TH*TN == 128
    int tiles_h, tiles_w;    // GH / TH, GW / TW
    // operand staging plan (see tc_plan):
    //   rg = 1  "row-group" mode for stride-1 KxK convs on 16x8 (rows x cols) pixel tiles: one A box of
    //           (TH + KH - 1) x TW pixels per kernel COLUMN q serves the KH vertical taps as row-shifted views of the
    //           same shared-memory tile (a shift of r rows = r * TW * row_bytes, a whole swizzle period), so the
    //           L2 -> SM operand traffic drops from KH*KW to KW * (TH + KH - 1) / TH tiles per output tile
    //   nb      B (weight) tiles consumed per K-step (KH in row-group mode, else 1)
    //   resident = 1: all weights of the CTA's (single) n-tile are loaded once into shared memory
    int rg, nb, a_bytes, stage_bytes, stages, resident, res_bytes;
    int operand_bytes;       // resident weights + ring, rounded up to 1 KB: the barriers / epilogue staging follow it
    int m_tiles, n_tiles, n_classes;   // persistent tile walk: tile = (cls * n_tiles + nt) * m_tiles + mt
    const float* bias;
    const __nv_bfloat16* residual;
    __nv_bfloat16* y;
    long long* trace;        // optional debug timeline of CTA 0 (pidm_debug_set_trace): [role][event] = clock64
    float* gn_sums;          // optional [B, G, 2]: GroupNorm sum / sum-of-squares of the output, fused into the epilogue
    int gn_cpg, gn_groups;   // channels per group, groups
    TcClass cls[4];
};

// GroupNorm statistics of one 32-channel chunk held by a warp (one pixel row per lane).  Per group the channels
// are summed in registers; the 2*NG partial sums of the 32 lanes are then reduced with a butterfly that halves the
// number of live values at every step (NV + log2-many shuffles instead of 5 per value), leaving value i on lanes
// {i*32/NV ...}; those lanes issue one atomic each.
template <int NV>
__device__ __forceinline__ float butterfly_reduce(float (&v)[NV], int lane) {
    // after the step with offset `off`, a lane keeps the half of its values selected by bit `off` of its lane id
    int n = NV;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (n > 1) {
            const int half = n >> 1;
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) {
                if (i < half) {
                    const float keep = upper ? v[i + half] : v[i];
                    const float send = upper ? v[i] : v[i + half];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            n = half;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
        }
    }
    return v[0];
}

template <int CPG>
__device__ __forceinline__ void gn_stats_chunk(const float* f, float* sums_b, int first_group, int lane) {
    constexpr int NG = (CPG >= 32) ? 1 : 32 / CPG;
    constexpr int W = (CPG >= 32) ? 32 : CPG;
    constexpr int NV = 2 * NG;                      // (sum, sumsq) per group: 2, 4, 8 or 16 values
    float v[NV];
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int j = 0; j < W; ++j) { float x = f[gi * W + j]; s += x; ss += x * x; }
        v[2 * gi] = s; v[2 * gi + 1] = ss;
    }
    const float total = butterfly_reduce<NV>(v, lane);
    // value index held by this lane: built from the lane bits consumed while n > 1 (bit 16 first = most significant)
    int idx = 0, n = NV;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (n > 1) { n >>= 1; if (lane & off) idx += n; }
    }
    // lanes that differ only in the bits consumed after n reached 1 hold the same total: the lowest one publishes
    constexpr int DUP = 32 / NV;
    if ((lane & (DUP - 1)) == 0) atomicAdd(sums_b + first_group * 2 + idx, total);
}

template <int BN, int BK>
struct TcCfg {
    static constexpr int SW = BK * 2;                                     // swizzle span in bytes (128 or 64)
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int ACC_STAGES = 2;                                  // TMEM accumulators: epilogue(i) || mainloop(i+1)
    static constexpr int TMEM_COLS = ACC_STAGES * BN;                     // 64 .. 512, power of two
};
// persistent kernel, one CTA per SM: the operand ring takes (almost) all shared memory so that the TMA producers
// run many K-steps (and tiles) ahead of the tensor pipe; latency is hidden by the ring, not by co-resident CTAs
constexpr int TC_MAX_STAGES = 12;
constexpr int TC_RING_STAGES = 12;      // default ring depth (PIDM_TC_STAGES overrides)
constexpr int TC_OPERAND_BYTES = 200 * 1024;                              // resident weights + ring
constexpr int TC_SMEM_BYTES = TC_OPERAND_BYTES + 1024 /*align slack*/ + 512 /*barriers*/ + 4 * 4096 /*epilogue staging*/;

// Persistent, warp-specialised implicit-GEMM convolution.  Tiles (m_tile, n_tile, class) are walked with a static
// stride of gridDim.x by all three roles in lock step:
//   warp 0   TMA producer : keeps the smem ring full across tile boundaries
//   warp 1   MMA issuer   : tcgen05.mma into TMEM accumulator (tile & 1); commits free ring slots / publish the tile
//   warp 2   TMEM allocator
//   warps 4-7 epilogue    : drain accumulator (tile & 1) while the tensor pipe already works on the next tile
template <int BN, int BK>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                const __grid_constant__ CUtensorMap map_w, TcParams p) {
    using Cfg = TcCfg<BN, BK>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned operand ring (required by the 128B swizzle atoms)
    const uint32_t raw_addr = tc_smem_u32(smem_raw);
    const uint32_t pad_bytes = (1024 - (raw_addr & 1023)) & 1023;
    unsigned char* wres = smem_raw + pad_bytes;              // resident weights (res_bytes, may be 0)
    unsigned char* ring = wres + p.res_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + pad_bytes + p.operand_bytes);
    uint64_t* full = bars;                                   // [TC_MAX_STAGES]
    uint64_t* empty = bars + TC_MAX_STAGES;                  // [TC_MAX_STAGES]
    uint64_t* acc_full = bars + 2 * TC_MAX_STAGES;           // [ACC_STAGES]
    uint64_t* acc_empty = acc_full + Cfg::ACC_STAGES;        // [ACC_STAGES]
    uint64_t* wfull = acc_empty + Cfg::ACC_STAGES;           // resident weights have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
    unsigned char* stage_base = smem_raw + pad_bytes + p.operand_bytes + 512;   // 4 warps x 4 KB epilogue staging
    const int n_stages = p.stages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kc_per_tap = p.Cin / BK;
    const int total_tiles = p.m_tiles * p.n_tiles * p.n_classes;
    pdl_trigger();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < n_stages; ++s) { tc_mbar_init(&full[s], 1); tc_mbar_init(&empty[s], 1); }
        for (int s = 0; s < Cfg::ACC_STAGES; ++s) { tc_mbar_init(&acc_full[s], 1); tc_mbar_init(&acc_empty[s], 128); }
        tc_mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // TMEM allocation (power of two >= 32 columns), whole warp executes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)),
                     "r"(Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                 // prologue done; everything below reads what the previous kernel wrote

    if (warp == 0 || warp == 2 || warp == 3) {
        // ===== TMA producers: three warps, one elected lane each, K-step `git` belongs to producer git % 3 ===========
        // (a single thread can only issue a K-step every ~600 cycles -- wait + expect_tx + 2 TMA -- which starved the
        //  tensor pipe on the small-channel layers; the three issue streams are independent)
        const uint32_t pidx = (warp == 0) ? 0u : (uint32_t)(warp - 1);
        if (elect_one()) {
            if (p.resident && pidx == 0) {
                // all K tiles of the (single) n-tile: [tap][kc] boxes of BN x BK
                const int n_k = p.KH * p.KW * kc_per_tap;
                tc_mbar_expect_tx(wfull, (uint32_t)(n_k * Cfg::B_BYTES));
                for (int i = 0; i < n_k; ++i) tma_load_2d(wres + (size_t)i * Cfg::B_BYTES, &map_w, wfull, i * BK, 0);
            }
            // ring position (stage, phase, whose turn) is carried incrementally: no integer divisions in the loop
            uint32_t st = 0, ph = 0, turn = 0, git = 0;
            unsigned char* a_dst = ring;
            const bool tracing = p.trace != nullptr && blockIdx.x == 0;
            const int groups_m0 = p.rg ? p.KW : p.KH * p.KW;
            int tile_m = blockIdx.x % p.m_tiles, rest = blockIdx.x / p.m_tiles;      // tile = rest * m_tiles + tile_m
            const int step_m = gridDim.x % p.m_tiles, step_r = gridDim.x / p.m_tiles;
            int lt = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
                const int n0 = (rest % p.n_tiles) * BN;
                const TcClass& cl = p.cls[rest / p.n_tiles];
                const int tw_idx = tile_m % p.tiles_w;
                const int t2 = tile_m / p.tiles_w;
                const int tb = t2 / p.tiles_h, th_idx = t2 - tb * p.tiles_h;
                const int b0 = tb * p.TN, hh0 = p.in_stride * th_idx * p.TH, ww0 = p.in_stride * tw_idx * p.TW;
                const int n_groups = (p.mode == 0) ? groups_m0 : cl.n_taps;
                if (tracing && pidx == 0 && lt < 500) p.trace[lt * 2] = clock64();
                int r = 0, q = 0;                              // mode 0, plain: tap (r, q) walked incrementally
                for (int g = 0; g < n_groups; ++g) {
                    int dh, dw, ktap;
                    if (p.mode == 0) {
                        if (p.rg) { dh = -p.pad; dw = g - p.pad; }
                        else { dh = r - p.pad; dw = q - p.pad; if (++q == p.KW) { q = 0; ++r; } }
                        ktap = g;
                    } else {
                        dh = cl.dh[g]; dw = cl.dw[g]; ktap = cl.ktap[g];
                    }
                    int kcol = ktap * p.Cin;
                    for (int kc = 0; kc < kc_per_tap; ++kc, kcol += BK) {
                        if (turn == pidx) {
                            tc_mbar_wait(&empty[st], ph ^ 1);
                            if (tracing && git < 1000) p.trace[6144 + git] = clock64();
                            tc_mbar_expect_tx(&full[st], (uint32_t)p.stage_bytes);
                            tma_load_4d(a_dst, &map_x, &full[st], kc * BK, ww0 + dw, hh0 + dh, b0);
                            if (!p.resident) {
                                unsigned char* b_dst = a_dst + p.a_bytes;
                                int kj = kcol;
                                for (int j = 0; j < p.nb; ++j, kj += p.KW * p.Cin, b_dst += Cfg::B_BYTES)
                                    tma_load_2d(b_dst, &map_w, &full[st], kj, n0);   // row-group mode: tap (r = j, q = g)
                            }
                        }
                        ++git;
                        if (++turn == 3u) turn = 0;
                        if (++st == (uint32_t)n_stages) { st = 0; ph ^= 1; a_dst = ring; } else a_dst += p.stage_bytes;
                    }
                }
                tile_m += step_m; rest += step_r;
                if (tile_m >= p.m_tiles) { tile_m -= p.m_tiles; ++rest; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
        // A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(TC_BM >> 4) << 24);
        // One elected thread runs the whole loop.  The loop body is kept free of integer divisions and descriptor
        // re-encoding: a clock64 trace showed ~500 cycles of scalar overhead per K-step in the naive form, more than
        // the MMAs themselves on the narrow (N = 32) tiles.
        if (elect_one()) {
            const bool tracing = p.trace != nullptr && blockIdx.x == 0;
            // descriptor = hi(constant: SBO, version, swizzle) : lo(start >> 4 | LBO = 1 << 16)
            const uint32_t desc_hi = (uint32_t)(umma_desc<Cfg::SW>(0) >> 32);
            const uint32_t ring_lo = ((tc_smem_u32(ring) & 0x3FFFF) >> 4) | (1u << 16);
            const uint32_t wres_lo = ((tc_smem_u32(wres) & 0x3FFFF) >> 4) | (1u << 16);
            const uint32_t stage_lo = (uint32_t)p.stage_bytes >> 4;
            const uint32_t a_shift_lo = (uint32_t)(p.TW * BK * 2) >> 4;      // one tile row of pixels
            const uint32_t b_off_lo = (uint32_t)p.a_bytes >> 4;
            constexpr uint32_t b_tile_lo = (uint32_t)Cfg::B_BYTES >> 4;
            const uint32_t res_j_lo = (uint32_t)(p.KW * kc_per_tap) * b_tile_lo;   // resident: next kernel row
            const int groups_m0 = p.rg ? p.KW : p.KH * p.KW;
            const int nb = p.nb;
            const bool resident = p.resident != 0;
            uint32_t st = 0, ph = 0, a_lo = ring_lo, git = 0;
            int rest = blockIdx.x / p.m_tiles, tile_m = blockIdx.x % p.m_tiles;
            const int step_m = gridDim.x % p.m_tiles, step_r = gridDim.x / p.m_tiles;
            int lt = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
                const int n_iters = ((p.mode == 0) ? groups_m0 : p.cls[rest / p.n_tiles].n_taps) * kc_per_tap;
                const int as = lt & 1;
                if (tracing && lt < 64) p.trace[1024 + lt * 4] = clock64();
                if (lt == 0 && resident) tc_mbar_wait(wfull, 0);
                tc_mbar_wait(&acc_empty[as], ((lt >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tracing && lt < 64) p.trace[1024 + lt * 4 + 1] = clock64();
                const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
                uint32_t accum = 0;
                uint32_t res_lo = wres_lo;                           // resident weights: tile (it) of kernel row 0
                for (int it = 0; it < n_iters; ++it, res_lo += b_tile_lo) {
                    tc_mbar_wait(&full[st], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (tracing && git < 1000) p.trace[4096 + git * 2] = clock64();
                    uint32_t aj = a_lo;
                    uint32_t bj = resident ? res_lo : a_lo + b_off_lo;
                    const uint32_t bj_step = resident ? res_j_lo : b_tile_lo;
                    for (int j = 0; j < nb; ++j, aj += a_shift_lo, bj += bj_step) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(aj + 2 * k);
                            const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(bj + 2 * k);
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_acc),
                                "l"(da), "l"(db), "r"(idesc), "r"(accum)
                                : "memory");
                            accum = 1;
                        }
                    }
                    // release the smem stage once the MMAs that read it have completed
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     tc_smem_u32(&empty[st]))
                                 : "memory");
                    if (tracing && git < 1000) p.trace[4096 + git * 2 + 1] = clock64();
                    ++git;
                    if (++st == (uint32_t)n_stages) { st = 0; ph ^= 1; a_lo = ring_lo; } else a_lo += stage_lo;
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 tc_smem_u32(&acc_full[as]))
                             : "memory");
                if (tracing && lt < 64) p.trace[1024 + lt * 4 + 3] = clock64();
                tile_m += step_m; rest += step_r;
                if (tile_m >= p.m_tiles) { tile_m -= p.m_tiles; ++rest; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers (+bias, +residual, GroupNorm statistics) -> bf16 -> smem transpose -> global
        // A lane owns one accumulator row (pixel).  Writing its row straight to global memory would make every store
        // instruction touch 32 different lines (16 bytes each): the LSU, not HBM, bounds the wide-N layers that way.
        // Instead each warp stages its 32 rows x CH columns in a private, XOR-swizzled shared-memory tile and writes it
        // back with LPR lanes per row, i.e. whole 64/128-byte row segments per quarter-warp.  The residual is read the
        // same way (coalesced -> staged -> own row) so that it is still added in fp32 before the single rounding.
        constexpr int CH = BN >= 64 ? 64 : 32;        // columns per pass
        constexpr int LPR = CH / 8;                   // 16-byte units per staged row = lanes per row when storing
        constexpr int RPI = 32 / LPR;                 // rows per store instruction
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        uint4* stage = reinterpret_cast<uint4*>(stage_base) + quarter * (32 * 8);
        const int m = quarter * 32 + lane;            // accumulator row = pixel within the tile
        const int tn = m / (p.TH * p.TW);
        const int rem = m - tn * p.TH * p.TW;
        const int th = rem / p.TW, tw = rem - th * p.TW;
        const int my_sw = (LPR == 8) ? (lane & 7) : ((lane >> 1) & 3);
        const int sr = lane / LPR, su = lane % LPR;   // store phase: row within the group of RPI rows, 16-byte unit
        const bool tracing = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 128;
        int tile_m = blockIdx.x % p.m_tiles, rest = blockIdx.x / p.m_tiles;
        const int step_m = gridDim.x % p.m_tiles, step_r = gridDim.x / p.m_tiles;
        int lt = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
            const int n0 = (rest % p.n_tiles) * BN;
            const TcClass& cl = p.cls[rest / p.n_tiles];
            const int tw_idx = tile_m % p.tiles_w;
            const int t2 = tile_m / p.tiles_w;
            const int tb = t2 / p.tiles_h, th_idx = t2 - tb * p.tiles_h;
            const int b = tb * p.TN + tn, h = th_idx * p.TH + th, w = tw_idx * p.TW + tw;
            const bool row_ok = (b < p.B) && (h < p.GH);
            const int oh = p.out_scale * h + cl.off_h, ow = p.out_scale * w + cl.off_w;
            const long long row_off = row_ok ? (long long)((((size_t)b * p.Ho + oh) * p.Wo + ow) * p.Cout + n0) : -1ll;
            // offsets of the rows this lane stores (row it * RPI + sr), -1 = masked
            long long st_off[LPR];
#pragma unroll
            for (int it = 0; it < LPR; ++it) st_off[it] = __shfl_sync(0xffffffffu, row_off, it * RPI + sr);
            const int as = lt & 1;
            if (tracing && lt < 64) p.trace[2048 + lt * 4] = clock64();
            tc_mbar_wait(&acc_full[as], (lt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tracing && lt < 64) p.trace[2048 + lt * 4 + 1] = clock64();
#pragma unroll 1
            for (int c = 0; c < BN; c += CH) {
                float f[CH];
#pragma unroll
                for (int hh = 0; hh < CH / 32; ++hh) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + c + hh * 32);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                          "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                          "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                          "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[hh * 32 + j] = row_ok ? __uint_as_float(v[j]) : 0.f;
                }
                if (tracing && lt < 64 && c == 0) p.trace[3072 + lt * 4] = clock64();
                if (c + CH >= BN) {
                    // the last columns of this accumulator are in registers: hand the TMEM stage back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(&acc_empty[as])) : "memory");
                }
                if (p.bias && row_ok) {
#pragma unroll
                    for (int j = 0; j < CH; j += 4) {
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + j));
                        f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
                    }
                }
                if (p.residual) {
#pragma unroll
                    for (int it = 0; it < LPR; ++it) {
                        const int r = it * RPI + sr;
                        const int sw = (LPR == 8) ? (r & 7) : ((r >> 1) & 3);
                        uint4 rv = make_uint4(0u, 0u, 0u, 0u);
                        if (st_off[it] >= 0) rv = *reinterpret_cast<const uint4*>(p.residual + st_off[it] + c + su * 8);
                        stage[r * LPR + (su ^ sw)] = rv;
                    }
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < LPR; ++u) {
                        const uint4 rv = stage[lane * LPR + (u ^ my_sw)];
                        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            f[u * 8 + 2 * k] += __low2float(hp[k]);
                            f[u * 8 + 2 * k + 1] += __high2float(hp[k]);
                        }
                    }
                    __syncwarp();
                }
                // bf16 rows -> swizzled staging tile
#pragma unroll
                for (int u = 0; u < LPR; ++u) {
                    uint4 pk;
                    __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                    for (int k = 0; k < 4; ++k) hp[k] = __floats2bfloat162_rn(f[u * 8 + 2 * k], f[u * 8 + 2 * k + 1]);
                    stage[lane * LPR + (u ^ my_sw)] = pk;
                }
                __syncwarp();
#pragma unroll
                for (int it = 0; it < LPR; ++it) {
                    const int r = it * RPI + sr;
                    const int sw = (LPR == 8) ? (r & 7) : ((r >> 1) & 3);
                    if (st_off[it] >= 0)
                        *reinterpret_cast<uint4*>(p.y + st_off[it] + c + su * 8) = stage[r * LPR + (su ^ sw)];
                }
                if (tracing && lt < 64 && c == 0) p.trace[3072 + lt * 4 + 1] = clock64();
                if (p.gn_sums != nullptr && b < p.B) {        // warp-uniform: the 32 rows of a warp lie in one sample
                    float* sums_b = p.gn_sums + (size_t)b * p.gn_groups * 2;
#pragma unroll
                    for (int hh = 0; hh < CH / 32; ++hh) {
                        const int fg = (n0 + c + hh * 32) / p.gn_cpg;
                        const float* fh = f + hh * 32;
                        if (p.gn_cpg == 4) gn_stats_chunk<4>(fh, sums_b, fg, lane);
                        else if (p.gn_cpg == 8) gn_stats_chunk<8>(fh, sums_b, fg, lane);
                        else if (p.gn_cpg == 16) gn_stats_chunk<16>(fh, sums_b, fg, lane);
                        else gn_stats_chunk<32>(fh, sums_b, fg, lane);     // cpg >= 32 (multiple of 32): one group
                    }
                }
                __syncwarp();                                   // staging tile is reused by the next pass
                if (tracing && lt < 64 && c == 0) p.trace[3072 + lt * 4 + 2] = clock64();
            }
            if (tracing && lt < 64) p.trace[2048 + lt * 4 + 2] = clock64();
            tile_m += step_m; rest += step_r;
            if (tile_m >= p.m_tiles) { tile_m -= p.m_tiles; ++rest; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
    }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

struct TcPlan {
    int TW, TH, TN, BN, BK;
    int rg, nb, a_bytes, stage_bytes, stages, resident, res_bytes, operand_bytes;
    int cps;                 // CTAs per SM the launch is sized for (grid = SMs * cps persistent CTAs)
};

// GH x GW = pixel grid of the GEMM (output grid for regular convs, input grid for the transposed gather)
static bool tc_plan(int B, int GH, int GW, int Cin, int Cout, int KH, int KW, int mode, int in_stride, int classes,
                    TcPlan& pl) {
    if (Cin % 32 != 0 || Cout % 32 != 0) return false;
    pl.BK = (Cin % 64 == 0) ? 64 : 32;
    // row-group mode: stride-1 KxK (K > 1) convs whose image splits into 16-row x 8-column tiles
    pl.rg = (mode == 0 && in_stride == 1 && KH > 1 && KH <= 7 && GW % 8 == 0 && GH % 16 == 0) ? 1 : 0;
    if (pl.rg) {
        pl.TW = 8; pl.TH = 16; pl.TN = 1;
    } else {
        if (GW > 128 || GW < 1 || (128 % GW) != 0) return false;
        pl.TW = GW;
        int th = 128 / GW;
        if (th > GH) th = GH;
        if (GH % th != 0) return false;
        pl.TH = th;
        pl.TN = 128 / (pl.TW * pl.TH);
    }
    if (pl.TW * pl.TH * pl.TN != 128) return false;
    if (pl.TW * in_stride > 256 || pl.TH * in_stride > 256) return false;      // TMA box limit
    pl.nb = pl.rg ? KH : 1;
    pl.a_bytes = (pl.TH + (pl.rg ? KH - 1 : 0)) * pl.TW * pl.TN * pl.BK * 2;
    const long long K = (long long)KH * KW * Cin;
    const long long m_tiles = (long long)((B + pl.TN - 1) / pl.TN) * (GH / pl.TH) * (GW / pl.TW) * classes;
    const int cands[4] = {256, 128, 64, 32};
    static int force_bn = -1;                       // debugging aid: PIDM_TC_BN pins the n-tile width
    if (force_bn < 0) { const char* ev = getenv("PIDM_TC_BN"); force_bn = ev ? atoi(ev) : 0; }
    pl.BN = 0;
    for (int i = 0; i < 4; ++i) {
        const int bn = cands[i];
        if (Cout % bn != 0) continue;
        if (force_bn > 0 && bn != force_bn && Cout % force_bn == 0) continue;
        // weights resident in shared memory when the CTA only ever sees one n-tile and they leave room for the ring
        const long long wbytes = (long long)bn * K * 2;
        int resident = (mode == 0 && Cout == bn && wbytes <= 100 * 1024) ? 1 : 0;
        int stage = pl.a_bytes + (resident ? 0 : pl.nb * bn * pl.BK * 2);
        int stages = (int)((TC_OPERAND_BYTES - (resident ? wbytes : 0)) / stage);
        if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
        if (stages < 3) continue;
        static int stage_cap = -1;                  // tuning aid: PIDM_TC_STAGES caps the ring depth
        if (stage_cap < 0) { const char* ev = getenv("PIDM_TC_STAGES"); stage_cap = ev ? atoi(ev) : TC_RING_STAGES; }
        if (stage_cap >= 3 && stages > stage_cap) stages = stage_cap;
        // Two CTAs per SM for the narrow layers with resident weights (BN = 32: 126 registers, 64 TMEM columns): these
        // layers are bound by per-CTA issue chains as much as by the SM's copy engine -- measured (B200, batch 32) 64x64x32
        // 3x3 10.5 -> 8.0 us, 4x4/s2 32->32 10.7 -> 7.7 us, 1x1 256->32 13.8 -> 11.1 us with two half-size rings, while
        // layers that stream their weights (several n-tiles) or already run wide tiles do not gain.  PIDM_TC_CTAS=1 disables.
        static int max_cps = -1;
        if (max_cps < 0) { const char* ev = getenv("PIDM_TC_CTAS"); max_cps = ev ? atoi(ev) : 2; if (max_cps < 1) max_cps = 1; }
        int cps = 1;
        if (max_cps >= 2 && bn == 32 && resident) {
            const long long fixed = (long long)wbytes + 1024 + 512 + 4 * 4096 + 1024;       // + 1 KB reserved per CTA
            const int s2 = (int)(((227 * 1024) / 2 - fixed) / stage);
            if (s2 >= 3) { cps = 2; if (stages > s2) stages = s2; }
        }
        pl.cps = cps;
        pl.BN = bn; pl.resident = resident; pl.res_bytes = resident ? (int)wbytes : 0;
        pl.stage_bytes = stage; pl.stages = stages;
        pl.operand_bytes = (int)(((resident ? wbytes : 0) + (long long)stages * stage + 1023) / 1024 * 1024);
        // widest tile that still (nearly) fills the machine: 128 tiles of N = 64 beat 256 tiles of N = 32 = 1.7 waves
        // (16x16x128 3x3: 8.5 -> 6.2 us), while 64 tiles of N = 64 lose to 128 of N = 32 (8x8x256: 11.5 vs 11.6 us)
        if (m_tiles * (Cout / bn) >= 128) break;
    }
    return pl.BN != 0;
}

template <int BN, int BK>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& p, dim3 grid, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        PIDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       TC_SMEM_BYTES));
        attr = true;
    }
    const size_t smem = (size_t)p.operand_bytes + 1024 /*align slack*/ + 512 /*barriers*/ + 4 * 4096 /*epilogue staging*/;
    PIDM_CUDA(launch_pdl(conv_tc_kernel<BN, BK>, grid, dim3(TC_THREADS), smem, st, mx, mw, p));
    PIDM_LAUNCH_CHECK("conv2d_tc");
    return 0;
}

// geometry of one call -> (pixel grid, classes).  Returns false when the tensor-core kernel does not cover it.
static long long* g_tc_trace = nullptr;

static bool tc_geometry(int B, int H, int W, int Cin, int Ho, int Wo, int Cout, int KH, int KW, int stride, int pad,
                        int transposed, TcParams& p, TcPlan& pl, int& classes) {
    if (KH != KW) return false;
    p.B = B; p.Cin = Cin; p.Cout = Cout; p.KH = KH; p.KW = KW; p.pad = pad; p.Ho = Ho; p.Wo = Wo;
    for (int c = 0; c < 4; ++c) { p.cls[c].n_taps = 0; p.cls[c].off_h = 0; p.cls[c].off_w = 0; }
    if (!transposed) {
        if (stride != 1 && stride != 2) return false;
        if ((H + 2 * pad - KH) / stride + 1 != Ho || (W + 2 * pad - KW) / stride + 1 != Wo) return false;
        p.mode = 0; p.GH = Ho; p.GW = Wo; p.in_stride = stride; p.out_scale = 1; classes = 1;
    } else {
        if (stride != 2 || (KH & 1) || Ho != 2 * H || Wo != 2 * W) return false;     // 4x4/s2/p1 style up-sampling
        p.mode = 1; p.GH = H; p.GW = W; p.in_stride = 1; p.out_scale = 2; classes = 4;
        for (int pa = 0; pa < 2; ++pa)
            for (int pb = 0; pb < 2; ++pb) {
                TcClass& c = p.cls[pa * 2 + pb];
                c.off_h = pa; c.off_w = pb;
                for (int r = 0; r < KH; ++r) {
                    if ((pa + pad - r) & 1) continue;
                    for (int q = 0; q < KW; ++q) {
                        if ((pb + pad - q) & 1) continue;
                        if (c.n_taps >= 16) return false;
                        c.dh[c.n_taps] = (signed char)((pa + pad - r) / 2);
                        c.dw[c.n_taps] = (signed char)((pb + pad - q) / 2);
                        c.ktap[c.n_taps] = (short)(r * KW + q);
                        ++c.n_taps;
                    }
                }
            }
    }
    if (!tc_plan(B, p.GH, p.GW, Cin, Cout, KH, KW, p.mode, p.in_stride, classes, pl)) return false;
    p.TW = pl.TW; p.TH = pl.TH; p.TN = pl.TN; p.tiles_h = p.GH / pl.TH; p.tiles_w = p.GW / pl.TW;
    p.rg = pl.rg; p.nb = pl.nb; p.a_bytes = pl.a_bytes; p.stage_bytes = pl.stage_bytes; p.stages = pl.stages;
    p.resident = pl.resident; p.res_bytes = pl.res_bytes; p.operand_bytes = pl.operand_bytes;
    return true;
}

static int tc_run(const void* x, const void* w_packed, const float* bias, const void* residual, void* y, int B, int H,
                  int W, int Cin, int Ho, int Wo, int Cout, int KH, int KW, int stride, int pad, int transposed,
                  float* gn_sums, int gn_groups, int gn_sums_zeroed, cudaStream_t st) {
    TcParams p;
    TcPlan pl;
    int classes = 1;
    PIDM_REQUIRE(tc_geometry(B, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad, transposed, p, pl, classes),
                 "conv2d_tc: unsupported geometry");
    // cuTensorMapEncodeTiled is a driver-API call: the calling thread (e.g. the autograd worker) may not have the
    // primary context bound yet if no runtime call has run in it
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        PIDM_CUDA(cudaFree(0));
        ctx_bound = true;
    }
    EncodeTiledFn enc = get_encode();
    PIDM_REQUIRE(enc != nullptr, "conv2d_tc: cuTensorMapEncodeTiled is not available from the driver");
    PIDM_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w_packed & 15) == 0, "conv2d_tc: operands must be 16-byte aligned");
    CUtensorMap mx, mw;
    const CUtensorMapSwizzle sw = pl.BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    {
        const int s = p.in_stride;
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        // with elementStrides = s the box spans boxDim global elements and loads boxDim / s of them
        const int box_h = pl.rg ? pl.TH + KH - 1 : pl.TH * s;
        cuuint32_t box[4] = {(cuuint32_t)pl.BK, (cuuint32_t)(pl.TW * s), (cuuint32_t)box_h, (cuuint32_t)pl.TN};
        cuuint32_t es[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
        CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PIDM_REQUIRE(r == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
    }
    {
        const int K = KH * KW * Cin;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {(cuuint32_t)pl.BK, (cuuint32_t)pl.BN};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        PIDM_REQUIRE(r == CUDA_SUCCESS, "conv2d_tc: cuTensorMapEncodeTiled(w) failed with %d", (int)r);
    }
    p.bias = bias; p.residual = (const __nv_bfloat16*)residual; p.y = (__nv_bfloat16*)y;
    p.trace = g_tc_trace;
    p.gn_sums = gn_sums; p.gn_groups = gn_groups; p.gn_cpg = gn_groups > 0 ? Cout / gn_groups : 0;
    if (gn_sums) {
        PIDM_REQUIRE(gn_groups > 0 && Cout % gn_groups == 0, "conv2d_tc: bad GroupNorm group count %d", gn_groups);
        const int cpg = p.gn_cpg;
        PIDM_REQUIRE(cpg == 4 || cpg == 8 || cpg == 16 || cpg % 32 == 0, "conv2d_tc: fused GroupNorm statistics need "
                     "4, 8, 16 or a multiple of 32 channels per group (got %d)", cpg);
        if (!gn_sums_zeroed) PIDM_CUDA(cudaMemsetAsync(gn_sums, 0, (size_t)B * gn_groups * 2 * sizeof(float), st));
    }
    p.m_tiles = ((B + pl.TN - 1) / pl.TN) * p.tiles_h * p.tiles_w;
    p.n_tiles = Cout / pl.BN;
    p.n_classes = classes;
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        PIDM_CUDA(cudaGetDevice(&dev));
        PIDM_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const int total_tiles = p.m_tiles * p.n_tiles * p.n_classes;
    const int slots = sm_count * pl.cps;
    dim3 grid(total_tiles < slots ? total_tiles : slots);
#define TC_CASE(bn, bk) if (pl.BN == bn && pl.BK == bk) return launch_tc<bn, bk>(mx, mw, p, grid, st)
    TC_CASE(256, 64); TC_CASE(128, 64); TC_CASE(64, 64); TC_CASE(32, 64);
    TC_CASE(256, 32); TC_CASE(128, 32); TC_CASE(64, 32); TC_CASE(32, 32);
#undef TC_CASE
    return set_error(2, "conv2d_tc: no kernel for BN=%d BK=%d", pl.BN, pl.BK);
}

}  // namespace pidm
using namespace pidm;

// debugging aid: device buffer of >= 8192 int64 that receives a clock64 timeline of CTA 0 of every conv_tc launch
extern "C" int pidm_debug_set_trace(void* buf) {
    g_tc_trace = (long long*)buf;
    return 0;
}

extern "C" int pidm_conv2d_tc_general_supported(int B, int H, int W, int Cin, int Ho, int Wo, int Cout, int KH, int KW,
                                                int stride, int pad, int transposed) {
    TcParams p; TcPlan pl; int classes;
    return tc_geometry(B, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad, transposed, p, pl, classes) ? 1 : 0;
}

extern "C" int pidm_conv2d_tc_general(const void* x, const void* w_packed, const float* bias, const void* residual,
                                      void* y, int B, int H, int W, int Cin, int Ho, int Wo, int Cout, int KH, int KW,
                                      int stride, int pad, int transposed, float* gn_sums, int gn_groups,
                                      int gn_sums_zeroed, void* stream) {
    return tc_run(x, w_packed, bias, residual, y, B, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad, transposed, gn_sums,
                  gn_groups, gn_sums_zeroed, (cudaStream_t)stream);
}
