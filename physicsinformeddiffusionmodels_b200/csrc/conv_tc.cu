// Stride-1 KxK convolution ("same" padding) on NHWC bf16 activations as an implicit GEMM on the 5th-generation
// tensor cores (tcgen05), operands staged by TMA, fp32 accumulators in tensor memory (TMEM).
//
//   y[b,h,w,n] = sum_{r,q,c} x[b, h+r-pad, w+q-pad, c] * Wp[n][(r*KW+q)*Cin + c] (+ bias[n]) (+ residual[b,h,w,n])
//
// This is the kernel behind every 3x3 / 1x1 layer of the U-Net (ResnetBlock convs, res_conv, to_qkv, to_out,
// reference unet_model.py:227,253,275,279) and behind their dgrad (same kernel, 180-degree-rotated packed weights).
//
// Tiling.  GEMM M = 128 output pixels = a TN x TH x TW box of the image (TW = W), N = BN output channels,
// K = taps * Cin walked in (tap, BK-channel) steps.  For one K step the A operand is ONE 4-D TMA box
// {BK channels, TW, TH, TN} of the NHWC tensor at spatial offset (r-pad, q-pad): the im2col gather, the zero
// padding (TMA out-of-bounds fill) and the 128B/64B shared-memory swizzle all happen in the copy engine.
// The B operand is a 2-D TMA box {BK, BN} of the K-major packed weights.  Both land in the canonical K-major
// swizzled layout that the UMMA shared-memory descriptors expect, so no thread ever touches the operands.
//
// Warp roles (256 threads): warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one elected lane,
// tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16), warp 2 = TMEM allocator, warps 4-7 = epilogue
// (tcgen05.ld 32x32b.x32 -> +bias +residual -> bf16 -> 128-bit global stores).  smem full/empty mbarrier ring
// between producer and MMA, tcgen05.commit releases stages and publishes the accumulator.
#include "common.cuh"
#include "pidm.h"
#include <cuda.h>

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
    int tiles_h;             // GH / TH
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
__device__ __forceinline__ void gn_stats_chunk(const float f[32], float* sums_b, int first_group, int lane) {
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
    static constexpr int A_BYTES = TC_BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // persistent kernel, one CTA per SM: the operand ring takes (almost) all shared memory so that the TMA producer
    // runs many K-steps (and tiles) ahead of the tensor pipe; latency is hidden by the ring, not by co-resident CTAs
    static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 12 ? 12 : STAGES_RAW;
    static constexpr int ACC_STAGES = 2;                                  // TMEM accumulators: epilogue(i) || mainloop(i+1)
    static constexpr int TMEM_COLS = ACC_STAGES * BN;                     // 64 .. 512, power of two
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
};

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
    unsigned char* ring = smem_raw + pad_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;                                   // [STAGES]
    uint64_t* empty = bars + Cfg::STAGES;                    // [STAGES]
    uint64_t* acc_full = bars + 2 * Cfg::STAGES;             // [ACC_STAGES]
    uint64_t* acc_empty = acc_full + Cfg::ACC_STAGES;        // [ACC_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + Cfg::ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kc_per_tap = p.Cin / BK;
    const int total_tiles = p.m_tiles * p.n_tiles * p.n_classes;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { tc_mbar_init(&full[s], 1); tc_mbar_init(&empty[s], 1); }
        for (int s = 0; s < Cfg::ACC_STAGES; ++s) { tc_mbar_init(&acc_full[s], 1); tc_mbar_init(&acc_empty[s], 128); }
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

    if (warp == 0 || warp == 2 || warp == 3) {
        // ===== TMA producers: three warps, one elected lane each, K-step `git` belongs to producer git % 3 ===========
        // (a single thread can only issue a K-step every ~600 cycles -- wait + expect_tx + 2 TMA -- which starved the
        //  tensor pipe on the small-channel layers; the three issue streams are independent)
        const uint32_t pidx = (warp == 0) ? 0u : (uint32_t)(warp - 1);
        if (elect_one()) {
            uint32_t git = 0;                                  // global K-step counter (ring position)
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int tile_m = tile % p.m_tiles;
                const int rest = tile / p.m_tiles;
                const int n0 = (rest % p.n_tiles) * BN;
                const TcClass& cl = p.cls[rest / p.n_tiles];
                const int tb = tile_m / p.tiles_h, th_idx = tile_m - tb * p.tiles_h;
                const int b0 = tb * p.TN, h0 = th_idx * p.TH;
                const int n_taps = (p.mode == 0) ? p.KH * p.KW : cl.n_taps;
                if (p.trace && blockIdx.x == 0 && pidx == 0 && git < 2000) p.trace[git / kc_per_tap / (n_taps > 0 ? n_taps : 1) * 2] = clock64();
                for (int tap = 0; tap < n_taps; ++tap) {
                    int dh, dw, ktap;
                    if (p.mode == 0) {
                        const int r = tap / p.KW, q = tap - r * p.KW;
                        dh = r - p.pad; dw = q - p.pad; ktap = tap;
                    } else {
                        dh = cl.dh[tap]; dw = cl.dw[tap]; ktap = cl.ktap[tap];
                    }
                    for (int kc = 0; kc < kc_per_tap; ++kc, ++git) {
                        if (git % 3u != pidx) continue;
                        const int s = git % Cfg::STAGES;
                        const uint32_t ph = (git / Cfg::STAGES) & 1;
                        tc_mbar_wait(&empty[s], ph ^ 1);
                        unsigned char* a_dst = ring + s * Cfg::STAGE_BYTES;
                        unsigned char* b_dst = a_dst + Cfg::A_BYTES;
                        tc_mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
                        tma_load_4d(a_dst, &map_x, &full[s], kc * BK, dw, p.in_stride * h0 + dh, b0);
                        tma_load_2d(b_dst, &map_w, &full[s], ktap * p.Cin + kc * BK, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
        // A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(TC_BM >> 4) << 24);
        uint32_t git = 0;
        int lt = 0;                                            // local tile counter
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
            const int rest = tile / p.m_tiles;
            const TcClass& cl = p.cls[rest / p.n_tiles];
            const int n_iters = ((p.mode == 0) ? p.KH * p.KW : cl.n_taps) * kc_per_tap;
            const int as = lt & 1;
            if (p.trace && blockIdx.x == 0 && lane == 0 && lt < 64) p.trace[1024 + lt * 4] = clock64();
            tc_mbar_wait(&acc_empty[as], ((lt >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (p.trace && blockIdx.x == 0 && lane == 0 && lt < 64) p.trace[1024 + lt * 4 + 1] = clock64();
            const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
            for (int it = 0; it < n_iters; ++it, ++git) {
                const int s = git % Cfg::STAGES;
                const uint32_t ph = (git / Cfg::STAGES) & 1;
                tc_mbar_wait(&full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t a_addr = tc_smem_u32(ring + s * Cfg::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = umma_desc<Cfg::SW>(a_addr + k * 32);
                        const uint64_t db = umma_desc<Cfg::SW>(b_addr + k * 32);
                        const uint32_t accum = (it > 0 || k > 0) ? 1u : 0u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "setp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_acc),
                            "l"(da), "l"(db), "r"(idesc), "r"(accum)
                            : "memory");
                    }
                    // release the smem stage once the MMAs that read it have completed
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     tc_smem_u32(&empty[s]))
                                 : "memory");
                    if (it == n_iters - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                         tc_smem_u32(&acc_full[as]))
                                     : "memory");
                }
                __syncwarp();
                if (p.trace && blockIdx.x == 0 && lane == 0 && lt < 64 && it == 0) p.trace[1024 + lt * 4 + 2] = clock64();
            }
            if (p.trace && blockIdx.x == 0 && lane == 0 && lt < 64) p.trace[1024 + lt * 4 + 3] = clock64();
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> (+bias, +residual) -> bf16 -> global =====
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int m = quarter * 32 + lane;            // accumulator row = pixel within the tile
        const int tn = m / (p.TH * p.TW);
        const int rem = m - tn * p.TH * p.TW;
        const int th = rem / p.TW, tw = rem - th * p.TW;
        int lt = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
            const int tile_m = tile % p.m_tiles;
            const int rest = tile / p.m_tiles;
            const int n0 = (rest % p.n_tiles) * BN;
            const TcClass& cl = p.cls[rest / p.n_tiles];
            const int tb = tile_m / p.tiles_h, th_idx = tile_m - tb * p.tiles_h;
            const int b = tb * p.TN + tn, h = th_idx * p.TH + th;
            const bool row_ok = (b < p.B) && (h < p.GH);
            const int oh = p.out_scale * h + cl.off_h, ow = p.out_scale * tw + cl.off_w;
            const size_t row_off = (((size_t)b * p.Ho + oh) * p.Wo + ow) * p.Cout + n0;
            const int as = lt & 1;
            if (p.trace && blockIdx.x == 0 && threadIdx.x == 128 && lt < 64) p.trace[2048 + lt * 4] = clock64();
            tc_mbar_wait(&acc_full[as], (lt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (p.trace && blockIdx.x == 0 && threadIdx.x == 128 && lt < 64) p.trace[2048 + lt * 4 + 1] = clock64();
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + c);
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
                if (p.trace && blockIdx.x == 0 && threadIdx.x == 128 && lt < 64 && c == 0) p.trace[3072 + lt * 4] = clock64();
                if (c + 32 >= BN) {
                    // the last chunk of this accumulator is in registers: hand the TMEM stage back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(&acc_empty[as])) : "memory");
                }
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = row_ok ? __uint_as_float(v[j]) : 0.f;
                if (row_ok) {
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 bv = *reinterpret_cast<const float4*>(p.bias + n0 + c + j);
                            f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
                        }
                    }
                    if (p.residual) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            float rv[8];
                            ld8(p.residual + row_off + c + j, rv);
#pragma unroll
                            for (int k = 0; k < 8; ++k) f[j + k] += rv[k];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 8) st8(p.y + row_off + c + j, f + j);
                }
                if (p.trace && blockIdx.x == 0 && threadIdx.x == 128 && lt < 64 && c == 0) p.trace[3072 + lt * 4 + 1] = clock64();
                if (p.gn_sums != nullptr && b < p.B) {        // warp-uniform: the 32 rows of a warp lie in one sample
                    float* sums_b = p.gn_sums + (size_t)b * p.gn_groups * 2;
                    const int fg = (n0 + c) / p.gn_cpg;
                    if (p.gn_cpg == 4) gn_stats_chunk<4>(f, sums_b, fg, lane);
                    else if (p.gn_cpg == 8) gn_stats_chunk<8>(f, sums_b, fg, lane);
                    else if (p.gn_cpg == 16) gn_stats_chunk<16>(f, sums_b, fg, lane);
                    else gn_stats_chunk<32>(f, sums_b, fg, lane);     // cpg >= 32 (multiple of 32): chunk inside one group
                }
                if (p.trace && blockIdx.x == 0 && threadIdx.x == 128 && lt < 64 && c == 0) p.trace[3072 + lt * 4 + 2] = clock64();
            }
            if (p.trace && blockIdx.x == 0 && threadIdx.x == 128 && lt < 64) p.trace[2048 + lt * 4 + 2] = clock64();
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
};

// GH x GW = pixel grid of the GEMM (output grid for regular convs, input grid for the transposed gather)
static bool tc_plan(int B, int GH, int GW, int Cin, int Cout, int in_stride, int classes, TcPlan& pl) {
    if (GW > 128 || GW < 1 || (128 % GW) != 0) return false;
    if (Cin % 32 != 0 || Cout % 32 != 0) return false;
    pl.TW = GW;
    int th = 128 / GW;
    if (th > GH) th = GH;
    if (GH % th != 0) return false;
    pl.TH = th;
    pl.TN = 128 / (pl.TW * pl.TH);
    if (pl.TW * pl.TH * pl.TN != 128) return false;
    if (pl.TW * in_stride > 256 || pl.TH * in_stride > 256) return false;      // TMA box limit
    pl.BK = (Cin % 64 == 0) ? 64 : 32;
    const long long m_tiles = (long long)((B + pl.TN - 1) / pl.TN) * (GH / pl.TH) * classes;
    const int cands[4] = {256, 128, 64, 32};
    pl.BN = 0;
    for (int i = 0; i < 4; ++i) {
        int bn = cands[i];
        if (Cout % bn != 0) continue;
        pl.BN = bn;
        if (m_tiles * (Cout / bn) >= 148) break;    // widest tile that still fills the machine
    }
    return pl.BN != 0;
}

template <int BN, int BK>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mw, const TcParams& p, dim3 grid, cudaStream_t st) {
    using Cfg = TcCfg<BN, BK>;
    static bool attr = false;
    if (!attr) {
        PIDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
        attr = true;
    }
    conv_tc_kernel<BN, BK><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(mx, mw, p);
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
    if (!tc_plan(B, p.GH, p.GW, Cin, Cout, p.in_stride, classes, pl)) return false;
    p.TW = pl.TW; p.TH = pl.TH; p.TN = pl.TN; p.tiles_h = p.GH / pl.TH;
    return true;
}

static int tc_run(const void* x, const void* w_packed, const float* bias, const void* residual, void* y, int B, int H,
                  int W, int Cin, int Ho, int Wo, int Cout, int KH, int KW, int stride, int pad, int transposed,
                  float* gn_sums, int gn_groups, cudaStream_t st) {
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
        cuuint32_t box[4] = {(cuuint32_t)pl.BK, (cuuint32_t)(pl.TW * s), (cuuint32_t)(pl.TH * s), (cuuint32_t)pl.TN};
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
        PIDM_CUDA(cudaMemsetAsync(gn_sums, 0, (size_t)B * gn_groups * 2 * sizeof(float), st));
    }
    p.m_tiles = ((B + pl.TN - 1) / pl.TN) * p.tiles_h;
    p.n_tiles = Cout / pl.BN;
    p.n_classes = classes;
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        PIDM_CUDA(cudaGetDevice(&dev));
        PIDM_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const int total_tiles = p.m_tiles * p.n_tiles * p.n_classes;
    dim3 grid(total_tiles < sm_count ? total_tiles : sm_count);
#define TC_CASE(bn, bk) if (pl.BN == bn && pl.BK == bk) return launch_tc<bn, bk>(mx, mw, p, grid, st)
    TC_CASE(256, 64); TC_CASE(128, 64); TC_CASE(64, 64); TC_CASE(32, 64);
    TC_CASE(256, 32); TC_CASE(128, 32); TC_CASE(64, 32); TC_CASE(32, 32);
#undef TC_CASE
    return set_error(2, "conv2d_tc: no kernel for BN=%d BK=%d", pl.BN, pl.BK);
}

}  // namespace pidm
using namespace pidm;

// debugging aid: device buffer of >= 4096 int64 that receives a clock64 timeline of CTA 0 of every conv_tc launch
extern "C" int pidm_debug_set_trace(void* buf) {
    g_tc_trace = (long long*)buf;
    return 0;
}

extern "C" int pidm_conv2d_tc_supported(int B, int H, int W, int Cin, int Cout, int KH, int KW, int pad) {
    TcParams p; TcPlan pl; int classes;
    return (2 * pad == KH - 1 && tc_geometry(B, H, W, Cin, H, W, Cout, KH, KW, 1, pad, 0, p, pl, classes)) ? 1 : 0;
}

extern "C" int pidm_conv2d_tc(const void* x, const void* w_packed, const float* bias, const void* residual, void* y,
                              int B, int H, int W, int Cin, int Cout, int KH, int KW, int pad, void* stream) {
    return tc_run(x, w_packed, bias, residual, y, B, H, W, Cin, H, W, Cout, KH, KW, 1, pad, 0, nullptr, 0,
                  (cudaStream_t)stream);
}

extern "C" int pidm_conv2d_tc_general_supported(int B, int H, int W, int Cin, int Ho, int Wo, int Cout, int KH, int KW,
                                                int stride, int pad, int transposed) {
    TcParams p; TcPlan pl; int classes;
    return tc_geometry(B, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad, transposed, p, pl, classes) ? 1 : 0;
}

extern "C" int pidm_conv2d_tc_general(const void* x, const void* w_packed, const float* bias, const void* residual,
                                      void* y, int B, int H, int W, int Cin, int Ho, int Wo, int Cout, int KH, int KW,
                                      int stride, int pad, int transposed, float* gn_sums, int gn_groups, void* stream) {
    return tc_run(x, w_packed, bias, residual, y, B, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad, transposed, gn_sums,
                  gn_groups, (cudaStream_t)stream);
}
