// Linear attention FUSED with its to_qkv 1x1 projection, for the full-resolution level of the U-Net (C = 32 input
// channels, 8 heads x 32; reference unet_model.py:275-297).
//
// At 64x64 the qkv tensor is 24x larger than the tensor it is projected from (768 vs 32 channels): writing it and
// streaming it back through the statistics / context / output / backward kernels was ~1.2 GB of HBM traffic per layer.
// Here every warp (one head) recomputes its q / k / v slices on the tensor cores from the 64-byte pixel rows of the
// normalised input xn -- a 32x32x32 mma.sync product per 32 pixels -- so that forward reads xn (8 MB) and writes
// `out` only, and backward reads xn + dout and writes dqkv once (consumed by the unchanged dgrad / wgrad of to_qkv).
// All intermediate tiles stay in registers: accumulator fragments are converted to A fragments directly and to
// transposed (K-major) fragments with movmatrix; only xn / dout tiles and the output staging touch shared memory.
// The two paths differ only by the bf16 rounding of the (here never materialised) q, k, v.
#define PIDM_PDL_GROUP 1
#include "common.cuh"
#include "mma_util.cuh"
#include "pidm.h"

namespace pidm {

constexpr int LF_C = 32;                       // channels of xn

__device__ __forceinline__ uint32_t movm_t(uint32_t x) {
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

// B fragments of a 32-row block of the K-major projection weights W[n][c] (bf16): B[k = c][n] = W[n][c]
__device__ __forceinline__ void load_w_frags(uint32_t (&w)[2][4][2], const __nv_bfloat16* __restrict__ Wrows, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const uint32_t* row = reinterpret_cast<const uint32_t*>(Wrows + (size_t)(nt * 8 + g) * LF_C + ks * 16 + 2 * t);
            w[ks][nt][0] = __ldg(row);
            w[ks][nt][1] = __ldg(row + 4);
        }
}
// A fragments of MT 16-pixel row blocks of an xn tile [rows][LW_PITCH]
template <int MT>
__device__ __forceinline__ void load_x_frags(uint32_t (&a)[MT][2][4], const __nv_bfloat16* Xs, int lane) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) frag_a_rowmajor(a[mt][ks], Xs, LW_PITCH, mt * 16, ks * 16, lane);
}
// c[mt][nt] = xn_tile(mt) * W^T in fp32.  (The unfused path rounds q, k, v to bf16 when it materialises them; the
// fused path keeps the fp32 products -- closer to the fp32 reference, and conversions share the XU pipe with exp.)
template <int MT>
__device__ __forceinline__ void project(float (&c)[MT][4][4], const uint32_t (&a)[MT][2][4], const uint32_t (&w)[2][4][2]) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) c[mt][nt][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) mma_bf16(c[mt][nt], a[mt][ks], w[ks][nt][0], w[ks][nt][1]);
        }
}
// softmax over the 32 columns of the two rows (g, g + 8) a thread shares with its quad, times mul
__device__ __forceinline__ void frag_softmax(float (&c)[4][4], float mul) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mx = fmaxf(mx, fmaxf(c[nt][half * 2], c[nt][half * 2 + 1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float s = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            c[nt][half * 2] = __expf(c[nt][half * 2] - mx);
            c[nt][half * 2 + 1] = __expf(c[nt][half * 2 + 1] - mx);
            s += c[nt][half * 2] + c[nt][half * 2 + 1];
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const float inv = mul / s;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { c[nt][half * 2] *= inv; c[nt][half * 2 + 1] *= inv; }
    }
}
// accumulator fragment [16 px][32] -> the two A fragments (k16 steps over the 32 columns) of the same matrix
__device__ __forceinline__ void c_to_a(uint32_t (&a)[2][4], const float (&c)[4][4]) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        a[ks][0] = pack_bf16(c[2 * ks][0], c[2 * ks][1]);
        a[ks][1] = pack_bf16(c[2 * ks][2], c[2 * ks][3]);
        a[ks][2] = pack_bf16(c[2 * ks + 1][0], c[2 * ks + 1][1]);
        a[ks][3] = pack_bf16(c[2 * ks + 1][2], c[2 * ks + 1][3]);
    }
}
// per-thread column constants: column (nt, j) = nt*8 + 2*(lane&3) + j
__device__ __forceinline__ void load_cols(float (&v)[8], const float* __restrict__ src, int lane) {
    const int t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { v[nt * 2] = src[nt * 8 + 2 * t]; v[nt * 2 + 1] = src[nt * 8 + 2 * t + 1]; }
}

// ---- pass 1: per-chunk column maxima of k = xn Wk^T (the exp-sums are accumulated by the context kernel) -------------
constexpr int LFS_STAGES = 2;      // a tile is consumed into registers at once: one tile of look-ahead is enough,
                                   // and 41 KB per CTA lets five CTAs share an SM
__global__ void __launch_bounds__(256) laf_kmax_kernel(const __nv_bfloat16* __restrict__ xn,
                                                       const __nv_bfloat16* __restrict__ W, float* __restrict__ part,
                                                       int N, int rows_per_chunk) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char raw[];
    const int b = blockIdx.y, chunk = blockIdx.x, lane = threadIdx.x & 31, h = threadIdx.x >> 5;
    __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(raw) + (size_t)h * (LFS_STAGES * LW_TILE);
    const int n_begin = chunk * rows_per_chunk, n_end = min(N, n_begin + rows_per_chunk);
    const int n_tiles = (n_end - n_begin) / 32;
    const __nv_bfloat16* xsrc = xn + ((size_t)b * N + n_begin) * LF_C;
    auto issue = [&](int it) {
        if (it < n_tiles) lw_issue<32>(ring + (size_t)(it % LFS_STAGES) * LW_TILE, xsrc + (size_t)it * 32 * LF_C, LF_C, lane);
        cp_commit();
    };
#pragma unroll
    for (int s = 0; s < LFS_STAGES; ++s) issue(s);
    uint32_t wk[2][4][2];
    load_w_frags(wk, W + (size_t)(LM_HID + h * LM_D) * LF_C, lane);
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
    for (int it = 0; it < n_tiles; ++it) {
        cp_wait<LFS_STAGES - 1>();
        __syncwarp();
        uint32_t ax[2][2][4];
        load_x_frags<2>(ax, ring + (size_t)(it % LFS_STAGES) * LW_TILE, lane);
        __syncwarp();
        issue(it + LFS_STAGES);
        float ck[2][4][4];
        project<2>(ck, ax, wk);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                m[nt * 2 + j] = fmaxf(fmaxf(m[nt * 2 + j], fmaxf(ck[0][nt][j], ck[0][nt][2 + j])),
                                      fmaxf(ck[1][nt][j], ck[1][nt][2 + j]));
    }
#pragma unroll
    for (int off = 4; off < 32; off <<= 1)
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], __shfl_xor_sync(0xffffffffu, m[i], off));
    if (lane < 4) {
        float* o = part + ((size_t)b * gridDim.x + chunk) * LM_HID + h * LM_D;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) o[nt * 8 + 2 * lane + j] = m[nt * 2 + j];
    }
}

// ctx[b][h][d][:] *= 1 / Z[b][h][d]  and  kzinv = 1 / Z, after the context kernel has accumulated both
__global__ void laf_finalize_kernel(float* __restrict__ ctx, float* __restrict__ kzinv, int n_rows) {
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const float zi = 1.f / kzinv[row];
    ctx[(size_t)row * LM_D + lane] *= zi;
    __syncwarp();
    if (lane == 0) kzinv[row] = zi;
}

// ---- context (MODE 0) / dcontext (MODE 1) --------------------------------------------------------------------------
//   MODE 0: ctx[h][d][e]  += sum_n exp(k[n,d] - M_d) v[n,e]          k, v projected from xn
//   MODE 1: dctx[h][d][e] += sum_n softmax_d(q[n,:])[d] s dout[n,e]  q projected from xn, dout streamed
template <int MODE>
struct LfcCfg {
    static constexpr int STAGES = MODE == 0 ? 4 : 2;                       // tiles are consumed into registers at once
    static constexpr int STAGE_ELEMS = MODE == 0 ? LW_TILE : 2 * LW_TILE;  // xn tile (| dout tile)
    static constexpr size_t SMEM = (size_t)LM_HEADS * STAGES * STAGE_ELEMS * 2 + (size_t)LM_HEADS * 2 * LM_D * 4;
};
template <int MODE>
__global__ void __launch_bounds__(256, 2) laf_ctx_kernel(const __nv_bfloat16* __restrict__ xn, const __nv_bfloat16* __restrict__ W,
                                                      const __nv_bfloat16* __restrict__ dout, const float* __restrict__ part,
                                                      int n_stat_chunks, float* __restrict__ kmax, float* __restrict__ kzinv,
                                                      float* __restrict__ ctx, int N, int chunk_px, float scale) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char raw[];
    constexpr int LFC_STAGES = LfcCfg<MODE>::STAGES, LFC_STAGE_ELEMS = LfcCfg<MODE>::STAGE_ELEMS;
    const int b = blockIdx.y, chunk = blockIdx.x, lane = threadIdx.x & 31, h = threadIdx.x >> 5;
    __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(raw) + (size_t)h * (LFC_STAGES * LFC_STAGE_ELEMS);
    float* sM = reinterpret_cast<float*>(raw + (size_t)LM_HEADS * LFC_STAGES * LFC_STAGE_ELEMS * 2) + h * 2 * LM_D;
    float* sZi = sM + LM_D;
    const int n_begin = chunk * chunk_px, n_end = min(N, n_begin + chunk_px);
    const int n_tiles = (n_end - n_begin) / 32;
    const size_t pix0 = (size_t)b * N + n_begin;
    const __nv_bfloat16* xsrc = xn + pix0 * LF_C;
    const __nv_bfloat16* gsrc = (MODE == 1) ? dout + pix0 * LM_HID + h * LM_D : nullptr;
    auto issue = [&](int it) {
        if (it < n_tiles) {
            __nv_bfloat16* buf = ring + (size_t)(it % LFC_STAGES) * LFC_STAGE_ELEMS;
            lw_issue<32>(buf, xsrc + (size_t)it * 32 * LF_C, LF_C, lane);
            if (MODE == 1) lw_issue<32>(buf + LW_TILE, gsrc + (size_t)it * 32 * LM_HID, LM_HID, lane);
        }
        cp_commit();
    };
#pragma unroll
    for (int s = 0; s < LFC_STAGES; ++s) issue(s);
    uint32_t w0[2][4][2], w1[2][4][2];               // MODE 0: Wk, Wv   MODE 1: Wq, (unused)
    load_w_frags(w0, W + (size_t)((MODE == 0 ? LM_HID : 0) + h * LM_D) * LF_C, lane);
    if (MODE == 0) load_w_frags(w1, W + (size_t)(2 * LM_HID + h * LM_D) * LF_C, lane);
    float Mc[8];
    float zacc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) zacc[i] = 0.f;
    if (MODE == 0) {                               // lane = channel d of this head: combine the per-chunk maxima
        const int c = h * LM_D + lane;
        float M = -INFINITY;
        for (int i = 0; i < n_stat_chunks; ++i) M = fmaxf(M, part[((size_t)b * n_stat_chunks + i) * LM_HID + c]);
        sM[lane] = M;
        if (chunk == 0) kmax[(size_t)b * LM_HID + c] = M;
        __syncwarp();
        load_cols(Mc, sM, lane);
    }
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
    for (int it = 0; it < n_tiles; ++it) {
        cp_wait<LFC_STAGES - 1>();
        __syncwarp();
        const __nv_bfloat16* buf = ring + (size_t)(it % LFC_STAGES) * LFC_STAGE_ELEMS;
        uint32_t ax[2][2][4];
        load_x_frags<2>(ax, buf, lane);
        uint32_t bv[2][4][2];                         // B fragments [k = px][n = e] per 16-pixel step
        if (MODE == 1) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t b01[4], b23[4];
                frag_b_krows(b01, buf + LW_TILE, LW_PITCH, ks * 16, 0, lane);
                frag_b_krows(b23, buf + LW_TILE, LW_PITCH, ks * 16, 16, lane);
                bv[ks][0][0] = b01[0]; bv[ks][0][1] = b01[1]; bv[ks][1][0] = b01[2]; bv[ks][1][1] = b01[3];
                bv[ks][2][0] = b23[0]; bv[ks][2][1] = b23[1]; bv[ks][3][0] = b23[2]; bv[ks][3][1] = b23[3];
            }
        }
        __syncwarp();                              // the tile is in registers
        issue(it + LFC_STAGES);
        float cw[2][4][4];
        project<2>(cw, ax, w0);
        if (MODE == 0) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        cw[mt][nt][i] = __expf(cw[mt][nt][i] - Mc[nt * 2 + (i & 1)]);
                        zacc[nt * 2 + (i & 1)] += cw[mt][nt][i];
                    }
            float cv[2][4][4];
            project<2>(cv, ax, w1);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    bv[mt][nt][0] = movm_t(pack_bf16(cv[mt][nt][0], cv[mt][nt][1]));
                    bv[mt][nt][1] = movm_t(pack_bf16(cv[mt][nt][2], cv[mt][nt][3]));
                }
        } else {
            frag_softmax(cw[0], scale);
            frag_softmax(cw[1], scale);
        }
        // A[d][px] = w~[px][d]^T: transposed 8x8 blocks of the accumulator fragments
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {           // 16-pixel k step
#pragma unroll
            for (int md = 0; md < 2; ++md) {       // 16-channel m tile
                uint32_t a[4];
                a[0] = movm_t(pack_bf16(cw[mt][2 * md][0], cw[mt][2 * md][1]));
                a[1] = movm_t(pack_bf16(cw[mt][2 * md + 1][0], cw[mt][2 * md + 1][1]));
                a[2] = movm_t(pack_bf16(cw[mt][2 * md][2], cw[mt][2 * md][3]));
                a[3] = movm_t(pack_bf16(cw[mt][2 * md + 1][2], cw[mt][2 * md + 1][3]));
#pragma unroll
                for (int ne = 0; ne < 4; ++ne) mma_bf16(acc[md][ne], a, bv[mt][ne][0], bv[mt][ne][1]);
            }
        }
    }
    const int g = lane >> 2, t = lane & 3;
    if (MODE == 0) {                               // Z_d = sum_n exp(k[n,d] - M_d): column sums over the 8 row-lanes
#pragma unroll
        for (int off = 4; off < 32; off <<= 1)
#pragma unroll
            for (int i = 0; i < 8; ++i) zacc[i] += __shfl_xor_sync(0xffffffffu, zacc[i], off);
        if (lane < 4) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    atomicAdd(kzinv + (size_t)b * LM_HID + h * LM_D + nt * 8 + 2 * lane + j, zacc[nt * 2 + j]);
        }
    }
    float* cb = ctx + ((size_t)b * LM_HEADS + h) * LM_D * LM_D;
#pragma unroll
    for (int md = 0; md < 2; ++md)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int d = md * 16 + g + half * 8;
            const float f = (MODE == 0) ? 1.f / (float)N : 1.f;      // MODE 0: 1 / Z_d is applied by laf_finalize_kernel
#pragma unroll
            for (int ne = 0; ne < 4; ++ne) {
                const int e = ne * 8 + 2 * t;
                atomicAdd(cb + d * LM_D + e, acc[md][ne][half * 2] * f);
                atomicAdd(cb + d * LM_D + e + 1, acc[md][ne][half * 2 + 1] * f);
            }
        }
}

// ---- out[n,h,e] = sum_d softmax_d(q[n,:])[d] * s * ctx[h][d][e],  q projected from xn --------------------------------
constexpr int LFO_STAGES = 2;      // 61 KB per CTA: three CTAs (24 warps) per SM -- ncu (r02): the kernel is bound by
                                   // fixed-latency dependencies (stall_wait 2.0 per issue at 16 warps per SM)
__global__ void __launch_bounds__(256) laf_out_kernel(const __nv_bfloat16* __restrict__ xn, const __nv_bfloat16* __restrict__ W,
                                                      const float* __restrict__ ctx, __nv_bfloat16* __restrict__ out, int N,
                                                      int chunk_px, float scale) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char raw[];
    const int b = blockIdx.y, chunk = blockIdx.x, lane = threadIdx.x & 31, h = threadIdx.x >> 5;
    __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(raw) + (size_t)h * ((LFO_STAGES + 1) * LW_TILE);
    __nv_bfloat16* Os = ring + LFO_STAGES * LW_TILE;
    const int n_begin = chunk * chunk_px, n_end = min(N, n_begin + chunk_px);
    const int n_tiles = (n_end - n_begin) / 32;
    const size_t pix0 = (size_t)b * N + n_begin;
    const __nv_bfloat16* xsrc = xn + pix0 * LF_C;
    __nv_bfloat16* odst = out + pix0 * LM_HID + h * LM_D;
    auto issue = [&](int it) {
        if (it < n_tiles) lw_issue<32>(ring + (size_t)(it % LFO_STAGES) * LW_TILE, xsrc + (size_t)it * 32 * LF_C, LF_C, lane);
        cp_commit();
    };
#pragma unroll
    for (int s = 0; s < LFO_STAGES; ++s) issue(s);
    uint32_t wq[2][4][2];
    load_w_frags(wq, W + (size_t)(h * LM_D) * LF_C, lane);
    const float* ch = ctx + ((size_t)b * LM_HEADS + h) * LM_D * LM_D;
    uint32_t bf[2][4][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) frag_b_global<true>(bf[ks][nt], ch, ks * 16, nt * 8, lane);
    const int g = lane >> 2, t = lane & 3;
    for (int it = 0; it < n_tiles; ++it) {
        cp_wait<LFO_STAGES - 1>();
        __syncwarp();
        uint32_t ax[2][2][4];
        load_x_frags<2>(ax, ring + (size_t)(it % LFO_STAGES) * LW_TILE, lane);
        __syncwarp();
        issue(it + LFO_STAGES);
        float cq[2][4][4];
        project<2>(cq, ax, wq);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            frag_softmax(cq[mt], scale);
            uint32_t a[2][4];
            c_to_a(a, cq[mt]);
            float c[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) c[i][k] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_bf16(c[nt], a[ks], bf[ks][nt][0], bf[ks][nt][1]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                __nv_bfloat16* o = Os + (size_t)(mt * 16 + g) * LW_PITCH + nt * 8 + 2 * t;
                *reinterpret_cast<uint32_t*>(o) = pack_bf16(c[nt][0], c[nt][1]);
                *reinterpret_cast<uint32_t*>(o + 8 * LW_PITCH) = pack_bf16(c[nt][2], c[nt][3]);
            }
        }
        __syncwarp();
        lw_store<32>(odst + (size_t)it * 32 * LM_HID, LM_HID, Os, lane);
        __syncwarp();
    }
}

// ---- backward per pixel: dq | dk | dv from xn, dout, ctx, dctx and the saved column statistics -------------------------
// 217 registers (the nine loop-invariant 32x32 B-operand fragment sets live in registers): one CTA = 8 warps per SM.
// Capping the kernel at 128 registers for two CTAs per SM was measured (round 2): ptxas spills ~70 fragment registers to
// local memory and the 64x64 layer goes from 104 to 154 us -- occupancy does not pay for re-reading the operands.
constexpr int LFB_ROWS = 16;
constexpr int LFB_STAGES = 4;
constexpr int LFB_STAGE_ELEMS = 2 * LFB_ROWS * LW_PITCH;          // xn tile | dout tile
constexpr int LFB_OUT_ELEMS = 3 * LFB_ROWS * LW_PITCH;            // dq | dk | dv staging
__global__ void __launch_bounds__(256) laf_bwd_kernel(const __nv_bfloat16* __restrict__ xn, const __nv_bfloat16* __restrict__ W,
                                                      const __nv_bfloat16* __restrict__ dout, const float* __restrict__ ctx,
                                                      const float* __restrict__ dctx, const float* __restrict__ kmax,
                                                      const float* __restrict__ kzinv, __nv_bfloat16* __restrict__ dqkv,
                                                      int N, int chunk_px, float scale) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char raw[];
    const int b = blockIdx.y, chunk = blockIdx.x, lane = threadIdx.x & 31, h = threadIdx.x >> 5;
    __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(raw) + (size_t)h * (LFB_STAGES * LFB_STAGE_ELEMS + LFB_OUT_ELEMS);
    __nv_bfloat16* Os = ring + LFB_STAGES * LFB_STAGE_ELEMS;
    float* scd = reinterpret_cast<float*>(raw + (size_t)LM_HEADS * (LFB_STAGES * LFB_STAGE_ELEMS + LFB_OUT_ELEMS) * 2) + h * LM_D;
    const int n_begin = chunk * chunk_px, n_end = min(N, n_begin + chunk_px);
    const int n_tiles = (n_end - n_begin) / LFB_ROWS;
    const size_t pix0 = (size_t)b * N + n_begin;
    const __nv_bfloat16* xsrc = xn + pix0 * LF_C;
    const __nv_bfloat16* gsrc = dout + pix0 * LM_HID + h * LM_D;
    __nv_bfloat16* ddst = dqkv + pix0 * 3 * LM_HID + h * LM_D;
    auto issue = [&](int it) {
        if (it < n_tiles) {
            __nv_bfloat16* buf = ring + (size_t)(it % LFB_STAGES) * LFB_STAGE_ELEMS;
            lw_issue<LFB_ROWS>(buf, xsrc + (size_t)it * LFB_ROWS * LF_C, LF_C, lane);
            lw_issue<LFB_ROWS>(buf + LFB_ROWS * LW_PITCH, gsrc + (size_t)it * LFB_ROWS * LM_HID, LM_HID, lane);
        }
        cp_commit();
    };
#pragma unroll
    for (int s = 0; s < LFB_STAGES; ++s) issue(s);
    const float* cg = ctx + ((size_t)b * LM_HEADS + h) * LM_D * LM_D;
    const float* dg = dctx + ((size_t)b * LM_HEADS + h) * LM_D * LM_D;
    {
        float s = 0.f;                 // cd[d] = sum_e dctx[d][e] ctx[d][e]   (lane = d)
#pragma unroll
        for (int e = 0; e < LM_D; e += 4) {
            const float4 x = *reinterpret_cast<const float4*>(dg + lane * LM_D + e);
            const float4 y = *reinterpret_cast<const float4*>(cg + lane * LM_D + e);
            s += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
        }
        scd[lane] = s;
    }
    __syncwarp();
    float Mc[8], Zc[8], cdc[8];
    load_cols(Mc, kmax + (size_t)b * LM_HID + h * LM_D, lane);
    load_cols(Zc, kzinv + (size_t)b * LM_HID + h * LM_D, lane);
    load_cols(cdc, scd, lane);
    uint32_t wq[2][4][2], wk[2][4][2], wv[2][4][2];
    load_w_frags(wq, W + (size_t)(h * LM_D) * LF_C, lane);
    load_w_frags(wk, W + (size_t)(LM_HID + h * LM_D) * LF_C, lane);
    load_w_frags(wv, W + (size_t)(2 * LM_HID + h * LM_D) * LF_C, lane);
    uint32_t bc[2][4][2], bd[2][4][2], bt[2][4][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            frag_b_global<false>(bc[ks][nt], cg, ks * 16, nt * 8, lane);      // B[k=e][n=d] = ctx[d][e]
            frag_b_global<false>(bd[ks][nt], dg, ks * 16, nt * 8, lane);      // B[k=e][n=d] = dctx[d][e]
            frag_b_global<true>(bt[ks][nt], dg, ks * 16, nt * 8, lane);       // B[k=d][n=e] = dctx[d][e]
        }
    const int g = lane >> 2, t = lane & 3;
    const float invN = 1.f / (float)N;
    __nv_bfloat16* Oq = Os;
    __nv_bfloat16* Ok = Os + LFB_ROWS * LW_PITCH;
    __nv_bfloat16* Ov = Os + 2 * LFB_ROWS * LW_PITCH;
    for (int it = 0; it < n_tiles; ++it) {
        cp_wait<LFB_STAGES - 1>();
        __syncwarp();
        const __nv_bfloat16* buf = ring + (size_t)(it % LFB_STAGES) * LFB_STAGE_ELEMS;
        uint32_t ax[1][2][4], ag[2][4];
        load_x_frags<1>(ax, buf, lane);
        frag_a_rowmajor(ag[0], buf + LFB_ROWS * LW_PITCH, LW_PITCH, 0, 0, lane);      // dout [px][e]
        frag_a_rowmajor(ag[1], buf + LFB_ROWS * LW_PITCH, LW_PITCH, 0, 16, lane);
        __syncwarp();                              // the tile is in registers
        issue(it + LFB_STAGES);
        float c1[1][4][4], c2[4][4];
        // ---- dq = s * p * (dp - sum_d p dp),  dp = dout ctx^T,  p = softmax_d(q) (bf16-rounded like the unfused path)
        project<1>(c1, ax, wq);
        frag_softmax(c1[0], 1.f);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) c2[nt][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) mma_bf16(c2[nt], ag[ks], bc[ks][nt][0], bc[ks][nt][1]);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float dot = 0.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                dot += c1[0][nt][half * 2] * c2[nt][half * 2] + c1[0][nt][half * 2 + 1] * c2[nt][half * 2 + 1];
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                *reinterpret_cast<uint32_t*>(Oq + (g + half * 8) * LW_PITCH + nt * 8 + 2 * t) =
                    pack_bf16(scale * c1[0][nt][half * 2] * (c2[nt][half * 2] - dot),
                              scale * c1[0][nt][half * 2 + 1] * (c2[nt][half * 2 + 1] - dot));
        }
        // ---- dk = k~ * (dk~ - cd),  dk~ = (v / N) dctx^T,  k~ = exp(k - M) Zinv
        uint32_t av[2][4], ak[2][4];
        project<1>(c1, ax, wv);
        c_to_a(av, c1[0]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) c2[nt][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) mma_bf16(c2[nt], av[ks], bd[ks][nt][0], bd[ks][nt][1]);
        }
        project<1>(c1, ax, wk);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                c1[0][nt][i] = __expf(c1[0][nt][i] - Mc[nt * 2 + (i & 1)]) * Zc[nt * 2 + (i & 1)];
        c_to_a(ak, c1[0]);
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                *reinterpret_cast<uint32_t*>(Ok + (g + half * 8) * LW_PITCH + nt * 8 + 2 * t) =
                    pack_bf16(c1[0][nt][half * 2] * (c2[nt][half * 2] * invN - cdc[nt * 2]),
                              c1[0][nt][half * 2 + 1] * (c2[nt][half * 2 + 1] * invN - cdc[nt * 2 + 1]));
        // ---- dv = (k~ dctx) / N
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) c2[nt][i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) mma_bf16(c2[nt], ak[ks], bt[ks][nt][0], bt[ks][nt][1]);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                *reinterpret_cast<uint32_t*>(Ov + (g + half * 8) * LW_PITCH + nt * 8 + 2 * t) =
                    pack_bf16(c2[nt][half * 2] * invN, c2[nt][half * 2 + 1] * invN);
        __syncwarp();
        __nv_bfloat16* d = ddst + (size_t)it * LFB_ROWS * 3 * LM_HID;
        lw_store<LFB_ROWS>(d, 3 * LM_HID, Oq, lane);
        lw_store<LFB_ROWS>(d + LM_HID, 3 * LM_HID, Ok, lane);
        lw_store<LFB_ROWS>(d + 2 * LM_HID, 3 * LM_HID, Ov, lane);
        __syncwarp();
    }
}

constexpr size_t LAF_STATS_SMEM = (size_t)LM_HEADS * LFS_STAGES * LW_TILE * 2;
constexpr size_t LAF_OUT_SMEM = (size_t)LM_HEADS * (LFO_STAGES + 1) * LW_TILE * 2;
constexpr size_t LAF_BWD_SMEM = (size_t)LM_HEADS * (LFB_STAGES * LFB_STAGE_ELEMS + LFB_OUT_ELEMS) * 2 + (size_t)LM_HEADS * LM_D * 4;

static int laf_attrs() {
    static bool done = false;
    if (!done) {
        PIDM_CUDA(cudaFuncSetAttribute(laf_kmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAF_STATS_SMEM));
        PIDM_CUDA(cudaFuncSetAttribute(laf_ctx_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LfcCfg<0>::SMEM));
        PIDM_CUDA(cudaFuncSetAttribute(laf_ctx_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LfcCfg<1>::SMEM));
        PIDM_CUDA(cudaFuncSetAttribute(laf_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAF_OUT_SMEM));
        PIDM_CUDA(cudaFuncSetAttribute(laf_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LAF_BWD_SMEM));
        done = true;
    }
    return 0;
}

static int laf_chunk_px(int B, int N, int ctas_per_sm) {
    // chunks are per sample: k chunks per sample with B * k <= resident CTA slots (one wave), 32-pixel granularity
    int k = (148 * ctas_per_sm) / B;
    if (k < 1) k = 1;
    int px = ((N + k - 1) / k + 31) / 32 * 32;
    if (px < 64) px = 64;
    if (px > N) px = N;
    return px;
}
static int laf_stat_chunks(int N) {
    int c = N / 128;
    if (c < 1) c = 1;
    if (c > 16) c = 16;                 // B * chunks CTAs: one wave at batch 32
    return c;
}

}  // namespace pidm
using namespace pidm;

extern "C" int pidm_linattn_fused_supported(int C, int heads, int N, int dtype) {
    return (C == LF_C && heads == LM_HEADS && dtype == PIDM_BF16 && N % 128 == 0) ? 1 : 0;
}

extern "C" int pidm_linattn_fused_workspace_floats(int B, int N) { return B * laf_stat_chunks(N) * LM_HID; }

// xn [B,N,32] bf16 (the PreNorm output), w_qkv [768][32] bf16 (packed to_qkv weights, K-major), out [B,N,256] bf16.
// ctx [B,8,32,32], kmax / kzinv [B,8,32] are outputs kept for backward; workspace: pidm_linattn_fused_workspace_floats.
extern "C" int pidm_linattn_fused_fwd(const void* xn, const void* w_qkv, void* out, float* ctx, float* kmax, float* kzinv,
                                      float* workspace, int B, int N, void* stream) {
    PIDM_REQUIRE(N % 128 == 0, "linattn_fused: N must be a multiple of 128 (got %d)", N);
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = laf_attrs()) return e;
    const float scale = 0.17677669529663687f;   // 32^-0.5
    const __nv_bfloat16* x = (const __nv_bfloat16*)xn;
    const __nv_bfloat16* w = (const __nv_bfloat16*)w_qkv;
    PIDM_CUDA(cudaMemsetAsync(ctx, 0, (size_t)B * LM_HEADS * LM_D * LM_D * sizeof(float), st));
    PIDM_CUDA(cudaMemsetAsync(kzinv, 0, (size_t)B * LM_HID * sizeof(float), st));
    const int chunks = laf_stat_chunks(N);
    const int rpc = N / chunks;
    PIDM_REQUIRE(rpc % 32 == 0 && rpc * chunks == N, "linattn_fused: bad statistics chunking for N=%d", N);
    PIDM_CUDA(launch_pdl(laf_kmax_kernel, dim3(dim3(chunks, B)), dim3(256), (size_t)(LAF_STATS_SMEM), st, x, w, workspace, N, rpc));
    const int cpx = laf_chunk_px(B, N, 2);
    PIDM_CUDA(launch_pdl(laf_ctx_kernel<0>, dim3(dim3((N + cpx - 1) / cpx, B)), dim3(256), (size_t)(LfcCfg<0>::SMEM), st, x, w, nullptr, workspace, chunks, kmax, kzinv,
                                                                             ctx, N, cpx, scale));
    PIDM_CUDA(launch_pdl(laf_finalize_kernel, dim3((B * LM_HID + 7) / 8), dim3(256), (size_t)(0), st, ctx, kzinv, B * LM_HID));
    const int opx = laf_chunk_px(B, N, 3);
    PIDM_CUDA(launch_pdl(laf_out_kernel, dim3(dim3((N + opx - 1) / opx, B)), dim3(256), (size_t)(LAF_OUT_SMEM), st, x, w, ctx, (__nv_bfloat16*)out, N, opx, scale));
    PIDM_LAUNCH_CHECK("linattn_fused_fwd");
    return 0;
}

// dqkv [B,N,768] bf16 is the gradient w.r.t. the (never materialised) qkv = xn W^T; dctx [B,8,32,32] is scratch.
extern "C" int pidm_linattn_fused_bwd(const void* xn, const void* w_qkv, const void* dout, const float* ctx,
                                      const float* kmax, const float* kzinv, void* dqkv, float* dctx, int B, int N,
                                      void* stream) {
    PIDM_REQUIRE(N % 128 == 0, "linattn_fused: N must be a multiple of 128 (got %d)", N);
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = laf_attrs()) return e;
    const float scale = 0.17677669529663687f;
    const __nv_bfloat16* x = (const __nv_bfloat16*)xn;
    const __nv_bfloat16* w = (const __nv_bfloat16*)w_qkv;
    PIDM_CUDA(cudaMemsetAsync(dctx, 0, (size_t)B * LM_HEADS * LM_D * LM_D * sizeof(float), st));
    const int cpx = laf_chunk_px(B, N, 2);
    PIDM_CUDA(launch_pdl(laf_ctx_kernel<1>, dim3(dim3((N + cpx - 1) / cpx, B)), dim3(256), (size_t)(LfcCfg<1>::SMEM), st, x, w, (const __nv_bfloat16*)dout, nullptr, 0,
                                                                             nullptr, nullptr, dctx, N, cpx, scale));
    const int bpx = laf_chunk_px(B, N, 1);
    PIDM_CUDA(launch_pdl(laf_bwd_kernel, dim3(dim3((N + bpx - 1) / bpx, B)), dim3(256), (size_t)(LAF_BWD_SMEM), st, x, w, (const __nv_bfloat16*)dout, ctx, dctx, kmax,
                                                                          kzinv, (__nv_bfloat16*)dqkv, N, bpx, scale));
    PIDM_LAUNCH_CHECK("linattn_fused_bwd");
    return 0;
}
