// Weight gradient of a stride-1 3x3 convolution on tcgen05, "tap-complete" tiling.
//
//   dW[co][ci][tap] += sum_{pixels m} dy[m][co] * x[m shifted by tap][ci]
//
// wgrad_tc.cu gives every CTA 128 rows of the (tap, ci) dimension; the rows of one CTA then belong to a few taps of
// MANY input channels, their addresses in the framework layout [co][ci][tap] are 9 floats apart, and the split-K
// epilogue degenerates into scattered 4-byte red.global.add (16 K transactions per CTA; it dominated every layer
// with few pixels).  Here a CTA owns ALL nine taps of a 32-channel chunk of ci and an NP-wide tile of co:
//   * three TMEM accumulators (M = 128 rows each = 4 atoms of 32 channels),
//   * for a fixed co its 9 x 32 results are 288 CONTIGUOUS floats of dW: the epilogue transposes through shared
//     memory and issues fully coalesced 128-byte reductions,
//   * the dy tile (B operand) is fetched once per K' step for all nine taps (it was fetched by three CTAs before).
// Two operand-staging modes:
//   RG = true  (image splits into 16x8 pixel tiles): per kernel COLUMN q one halo box of (16 + 3) x 8 pixels; the
//              three kernel rows r are row-shifted views of it -- atom j of accumulator q starts j * 8 pixel rows
//              further down, which the MN-major descriptor expresses as LBO = 8 rows (the 4th atom is discarded).
//   RG = false (small images, e.g. 8x8): nine separate boxes of 64 pixels, 12 atom slots (3 unused).
#include "common.cuh"
#include "pidm.h"
#include <cuda.h>

namespace pidm {

constexpr int W3_THREADS = 256;

__device__ __forceinline__ uint32_t w3_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void w3_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(w3_smem(bar)), "r"(count));
}
__device__ __forceinline__ void w3_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(w3_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void w3_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(w3_smem(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void w3_tma_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            w3_smem(dst)),
        "l"(map), "r"(w3_smem(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ bool w3_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// MN-major swizzled UMMA descriptor: LBO = byte stride between swizzle atoms along M/N, SBO = between 8-row K groups
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t w3_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    constexpr uint64_t layout = ROW_BYTES == 128 ? 2 : (ROW_BYTES == 64 ? 4 : 6);
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

struct W3Params {
    int B, pad;
    int TW, TH, TN, tiles_h, tiles_w;    // pixel tile of one K' step and the tile grid per TN samples
    int n_pix_tiles, tiles_per_split;
    float* dw;
    long long s_col;                     // dw index = cB * s_col + cA * 9 + tap
};

template <int NP, int AB, bool RG>
struct W3Cfg {
    static constexpr int PX = RG ? 128 : 64;                      // pixels per K' step
    static constexpr int A_ATOM = PX * 64;                        // [PX][32 ch] bf16 (RG = false)
    static constexpr int A_BOX_RG = 19 * 8 * 64;                  // (16 + 3) rows x 8 pixels x 32 ch
    static constexpr int A_RG_ALLOC = 10240;
    static constexpr int A_BYTES = RG ? 3 * A_RG_ALLOC : 12 * A_ATOM;
    static constexpr int NBOX = NP / AB;
    static constexpr int B_TILE = PX * AB * 2;
    static constexpr int B_BYTES = NBOX * B_TILE;
    static constexpr int STAGE_BYTES = B_BYTES + A_BYTES;         // B tiles first (they need the stricter alignment)
    static constexpr int TX_BYTES = B_BYTES + (RG ? 3 * A_BOX_RG : 9 * A_ATOM);
    static constexpr int STAGES_RAW = (184 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 4 ? 4 : STAGES_RAW;
    static constexpr int TMEM_COLS = 3 * NP <= 128 ? 128 : (3 * NP <= 256 ? 256 : 512);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    static_assert(STAGES * STAGE_BYTES >= 32 * 288 * 4, "epilogue staging must fit in the ring");
};

template <int NP, int AB, bool RG>
__global__ void __launch_bounds__(W3_THREADS, 1) wgrad3_kernel(const __grid_constant__ CUtensorMap map_x,
                                                               const __grid_constant__ CUtensorMap map_dy, W3Params p) {
    using Cfg = W3Cfg<NP, AB, RG>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw_addr = w3_smem(smem_raw);
    unsigned char* ring = smem_raw + ((1024 - (raw_addr & 1023)) & 1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + Cfg::STAGES;
    uint64_t* acc_full = bars + 2 * Cfg::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    const int c0 = blockIdx.x * 32, n0 = blockIdx.y * NP;
    const int pt_begin = blockIdx.z * p.tiles_per_split;
    int pt_end = pt_begin + p.tiles_per_split;
    if (pt_end > p.n_pix_tiles) pt_end = p.n_pix_tiles;
    const int n_iters = pt_end - pt_begin;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { w3_mbar_init(&full[s], 1); w3_mbar_init(&empty[s], 1); }
        w3_mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(w3_smem(tmem_slot)),
                     "r"(Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (n_iters > 0) {
        if (warp == 0) {
            if (w3_elect_one()) {
                // pixel tile index -> (sample group, tile row, tile column), walked incrementally
                int tw_idx = pt_begin % p.tiles_w;
                int t2 = pt_begin / p.tiles_w;
                int th_idx = t2 % p.tiles_h, tb = t2 / p.tiles_h;
                uint32_t st = 0, ph = 0;
                unsigned char* stage = ring;
                for (int it = 0; it < n_iters; ++it) {
                    w3_mbar_wait(&empty[st], ph ^ 1);
                    const int b0 = tb * p.TN, h0 = th_idx * p.TH, w0 = tw_idx * p.TW;
                    unsigned char* a_dst = stage + Cfg::B_BYTES;
                    w3_mbar_expect_tx(&full[st], Cfg::TX_BYTES);
#pragma unroll
                    for (int j = 0; j < Cfg::NBOX; ++j)
                        w3_tma_4d(stage + j * Cfg::B_TILE, &map_dy, &full[st], n0 + j * AB, w0, h0, b0);
                    if (RG) {
#pragma unroll
                        for (int q = 0; q < 3; ++q)
                            w3_tma_4d(a_dst + q * Cfg::A_RG_ALLOC, &map_x, &full[st], c0, w0 + q - p.pad, h0 - p.pad, b0);
                    } else {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap)
                            w3_tma_4d(a_dst + tap * Cfg::A_ATOM, &map_x, &full[st], c0, w0 + tap % 3 - p.pad,
                                      h0 + tap / 3 - p.pad, b0);
                    }
                    if (++tw_idx == p.tiles_w) { tw_idx = 0; if (++th_idx == p.tiles_h) { th_idx = 0; ++tb; } }
                    if (++st == (uint32_t)Cfg::STAGES) { st = 0; ph ^= 1; stage = ring; } else stage += Cfg::STAGE_BYTES;
                }
            }
        } else if (warp == 1) {
            // D = f32, A = B = bf16, both MN-major (bits 15, 16), N>>3 at [17,23), M>>4 at [24,29)
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                       ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            if (w3_elect_one()) {
                const uint32_t ring_addr = w3_smem(ring);
                // accumulator t: RG -> halo box of kernel column t, atoms (kernel rows) 8 pixel rows = 512 B apart;
                //                else -> atom slots 4t .. 4t+3, one atom apart
                constexpr uint32_t a_lbo = RG ? 8 * 64 : Cfg::A_ATOM;
                constexpr uint32_t a_acc_stride = RG ? Cfg::A_RG_ALLOC : 4 * Cfg::A_ATOM;
                const uint64_t da0 = w3_desc_mn<64>(ring_addr + Cfg::B_BYTES, a_lbo);
                const uint64_t db0 = w3_desc_mn<AB * 2>(ring_addr, Cfg::B_TILE);
                constexpr uint32_t stage_lo = Cfg::STAGE_BYTES >> 4;
                constexpr uint32_t ka_lo = (16 * 64) >> 4, kb_lo = (16 * AB * 2) >> 4, acc_lo = a_acc_stride >> 4;
                uint32_t st = 0, ph = 0, off_lo = 0, accum = 0;
                for (int it = 0; it < n_iters; ++it) {
                    w3_mbar_wait(&full[st], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
#pragma unroll
                        for (int k = 0; k < Cfg::PX / 16; ++k) {
                            const uint64_t da = da0 + (uint64_t)(off_lo + t * acc_lo + k * ka_lo);
                            const uint64_t db = db0 + (uint64_t)(off_lo + k * kb_lo);
                            const uint32_t acc_k = (k == 0) ? accum : 1u;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base + (uint32_t)(t * NP)),
                                "l"(da), "l"(db), "r"(idesc), "r"(acc_k)
                                : "memory");
                        }
                    }
                    accum = 1;
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     w3_smem(&empty[st]))
                                 : "memory");
                    if (++st == (uint32_t)Cfg::STAGES) { st = 0; ph ^= 1; off_lo = 0; } else off_lo += stage_lo;
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 w3_smem(acc_full))
                             : "memory");
            }
            __syncwarp();
        } else if (warp >= 4) {
            // ===== epilogue: warp q reads atom slot q of every accumulator (TMEM lane quarter q), lane = channel ci
            const int quarter = warp & 3;
            float* S = reinterpret_cast<float*>(ring);             // [32 co][288] staging, the ring is idle by now
            w3_mbar_wait(acc_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < NP; c += 32) {
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const int tap = RG ? quarter * 3 + t : t * 4 + quarter;        // RG: (r = quarter, q = t)
                    const bool valid = RG ? quarter < 3 : tap < 9;                 // warp-uniform
                    if (!valid) continue;
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * NP + c);
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
                    // S[co][ci * 9 + tap]: lanes are 9 floats apart -> conflict-free
#pragma unroll
                    for (int j = 0; j < 32; ++j) S[j * 288 + lane * 9 + tap] = __uint_as_float(v[j]);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                // 288 contiguous floats of dW per co: coalesced reductions, 8 output channels per warp
#pragma unroll 1
                for (int j = quarter; j < 32; j += 4) {
                    float* dst = p.dw + (long long)(n0 + c + j) * p.s_col + (long long)c0 * 9;
#pragma unroll
                    for (int i = 0; i < 9; ++i) atomicAdd(dst + i * 32 + lane, S[j * 288 + i * 32 + lane]);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
    }
}

typedef CUresult (*W3EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static W3EncodeFn w3_get_encode() {
    static W3EncodeFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (W3EncodeFn)ptr;
    }
    return fn;
}

static int w3_encode(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int atom, int bw, int bh, int bn) {
    W3EncodeFn enc = w3_get_encode();
    PIDM_REQUIRE(enc != nullptr, "wgrad3: cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)atom, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, atom == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PIDM_REQUIRE(r == CUDA_SUCCESS, "wgrad3: cuTensorMapEncodeTiled failed with %d", (int)r);
    return 0;
}

template <int NP, int AB, bool RG>
static int w3_launch(const CUtensorMap& mx, const CUtensorMap& my, const W3Params& p, dim3 grid, cudaStream_t st) {
    using Cfg = W3Cfg<NP, AB, RG>;
    static bool attr = false;
    if (!attr) {
        PIDM_CUDA(cudaFuncSetAttribute(wgrad3_kernel<NP, AB, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::SMEM_BYTES));
        attr = true;
    }
    PIDM_CUDA(launch_pdl(wgrad3_kernel<NP, AB, RG>, grid, dim3(W3_THREADS), Cfg::SMEM_BYTES, st, mx, my, p));
    PIDM_LAUNCH_CHECK("conv2d_wgrad_tc(3x3)");
    return 0;
}

// Is this call covered?  a = x [B,HA,WA,CA], b = dy [B,GH,GW,CB], stride 1, 3x3, "same" padding, framework layout
// dw[cB][cA][tap] (s_row == 9), no channel padding.
bool wgrad3_supported(int B, int HA, int WA, int CA, int CA_real, int GH, int GW, int CB, int KH, int KW, int a_stride,
                      int pad, long long s_row) {
    if (KH != 3 || KW != 3 || a_stride != 1 || pad != 1 || s_row != 9) return false;
    if (CA % 32 != 0 || CA_real != CA || CB % 32 != 0 || HA != GH || WA != GW) return false;
    if (GW % 8 == 0 && GH % 16 == 0) return true;                       // RG
    if (GW > 64 || 64 % GW != 0) return false;
    const int th = 64 / GW;
    if (th <= GH) return GH % th == 0;
    const int tn = 64 / (GW * GH);
    return GW * GH * tn == 64 && B % tn == 0;
}

int wgrad3_run(const void* a, const void* b, float* dw, int B, int HA, int WA, int CA, int GH, int GW, int CB,
               long long s_col, cudaStream_t st) {
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        PIDM_CUDA(cudaFree(0));
        ctx_bound = true;
    }
    const bool rg = (GW % 8 == 0 && GH % 16 == 0);
    W3Params p;
    p.B = B; p.pad = 1; p.dw = dw; p.s_col = s_col;
    if (rg) { p.TW = 8; p.TH = 16; p.TN = 1; }
    else {
        p.TW = GW;
        int th = 64 / GW;
        if (th > GH) th = GH;
        p.TH = th;
        p.TN = 64 / (p.TW * p.TH);
    }
    p.tiles_h = GH / p.TH; p.tiles_w = GW / p.TW;
    p.n_pix_tiles = (B / p.TN) * p.tiles_h * p.tiles_w;
    const int NP = (CB % 128 == 0) ? 128 : ((CB % 64 == 0) ? 64 : 32);
    const int AB = (CB % 64 == 0) ? 64 : 32;
    CUtensorMap mx, my;
    if (int e = w3_encode(&mx, a, B, HA, WA, CA, 32, p.TW, rg ? p.TH + 3 : p.TH, p.TN)) return e;
    if (int e = w3_encode(&my, b, B, GH, GW, CB, AB, p.TW, p.TH, p.TN)) return e;
    const int chunks = CA / 32, n_tiles = CB / NP;
    int splits = 148 / (chunks * n_tiles);
    if (splits > p.n_pix_tiles) splits = p.n_pix_tiles;
    if (splits < 1) splits = 1;
    p.tiles_per_split = (p.n_pix_tiles + splits - 1) / splits;
    splits = (p.n_pix_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
    dim3 grid(chunks, n_tiles, splits);
#define W3_CASE(np, ab) \
    if (NP == np && AB == ab) return rg ? w3_launch<np, ab, true>(mx, my, p, grid, st) : w3_launch<np, ab, false>(mx, my, p, grid, st)
    W3_CASE(128, 64); W3_CASE(64, 64); W3_CASE(32, 32);
#undef W3_CASE
    return set_error(2, "wgrad3: no kernel for NP=%d AB=%d", NP, AB);
}

}  // namespace pidm
