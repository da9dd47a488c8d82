// Linear attention (reference unet_model.py:286-297) on the tensor cores for bf16 activations, heads = 8, dim_head = 32.
//
// The per-(sample, head) products are 32x32 blocks -- too small for tcgen05 (UMMA M >= 64) and HBM-bound anyway
// (the whole qkv row of a pixel, 8 heads x 3 x 32 channels = 1536 B, is streamed once), so these kernels use
// warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate) with ldmatrix-fed fragments: one warp per head, a CTA
// streams full 512-byte pixel rows (all 8 heads) so that every global access is a contiguous 16-byte vector.
//   la_ctx_mma<0>: ctx[h][d][e]  += sum_n exp(k[n,d]-M_d) v[n,e]      (scaled by 1/(Z_d N) in the epilogue)
//   la_ctx_mma<1>: dctx[h][d][e] += sum_n softmax_d(q[n,:])[d]*s * dout[n,e]
//   la_out_mma   : out[n,h,e]     = sum_d softmax_d(q[n,:])[d]*s * ctx[h][d][e]
//   la_bwd_mma   : dq, dk, dv per pixel from dout, ctx, dctx and the saved column statistics
#include "common.cuh"
#include "pidm.h"

namespace pidm {

constexpr int LM_HEADS = 8, LM_D = 32, LM_HID = 256;
constexpr int LM_PITCH = LM_HID + 8;          // bf16 elements per smem row (528 B: 16-byte aligned, conflict-free ldmatrix)
constexpr int LM_CPITCH = LM_D + 8;           // ctx rows [d][e] in bf16
constexpr int LM_CHUNK = 256;                 // pixels per CTA

__device__ __forceinline__ uint32_t lm_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st8_smem(__nv_bfloat16* p, const float v[8]) {
    uint4 t = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(p) = t;
}

// A fragment (16 rows x 16 k) from a row-major smem tile S[row][col]: rows = M index, cols = K index
__device__ __forceinline__ void frag_a_rowmajor(uint32_t (&a)[4], const __nv_bfloat16* S, int pitch, int m0, int k0, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    ldsm_x4(a, lm_smem(S + (size_t)(m0 + r + 8 * (mi & 1)) * pitch + k0 + 8 * (mi >> 1)));
}
// A fragment when smem holds the transpose: S[k][m] (rows = K index, cols = M index)
__device__ __forceinline__ void frag_a_kmajor(uint32_t (&a)[4], const __nv_bfloat16* S, int pitch, int k0, int m0, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    ldsm_x4_t(a, lm_smem(S + (size_t)(k0 + r + 8 * (mi >> 1)) * pitch + m0 + 8 * (mi & 1)));
}
// B fragments of TWO adjacent n-tiles (n0..n0+15) for one k16 step, from S[k][n] (rows = K index): b[0..1] tile 0, b[2..3] tile 1
__device__ __forceinline__ void frag_b_krows(uint32_t (&b)[4], const __nv_bfloat16* S, int pitch, int k0, int n0, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    ldsm_x4_t(b, lm_smem(S + (size_t)(k0 + r + 8 * (mi & 1)) * pitch + n0 + 8 * (mi >> 1)));
}
// same from S[n][k] (rows = N index, cols = K index)
__device__ __forceinline__ void frag_b_nrows(uint32_t (&b)[4], const __nv_bfloat16* S, int pitch, int n0, int k0, int lane) {
    const int mi = lane >> 3, r = lane & 7;
    ldsm_x4(b, lm_smem(S + (size_t)(n0 + r + 8 * (mi >> 1)) * pitch + k0 + 8 * (mi & 1)));
}

// Load `rows` pixel rows of one 256-channel third of qkv (or of a [.,256] tensor) into smem with a transform:
//   XF 0: raw    XF 1: exp(x - M[c])   XF 2: exp(x - M[c]) * Zi[c]   XF 3: softmax over the head's 32 channels (* mul)
template <int XF>
__device__ __forceinline__ void load_rows(const __nv_bfloat16* __restrict__ src, size_t row_stride, int rows_valid, int rows,
                                          __nv_bfloat16* dst, const float* sM, const float* sZi, float mul) {
    const int lane = threadIdx.x & 31;
    for (int idx = threadIdx.x; idx < rows * 32; idx += blockDim.x) {
        const int row = idx >> 5;                       // octet == lane because blockDim % 32 == 0
        float v[8];
        if (row < rows_valid) ld8(src + (size_t)row * row_stride + lane * 8, v);
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (XF == 3 || XF == 0) ? 0.f : -INFINITY;
        }
        if (XF == 1 || XF == 2) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float e = __expf(v[k] - sM[lane * 8 + k]);
                v[k] = (XF == 2) ? e * sZi[lane * 8 + k] : e;
            }
        } else if (XF == 3) {
            float mx = v[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) mx = fmaxf(mx, v[k]);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { v[k] = __expf(v[k] - mx); s += v[k]; }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            const float inv = mul / s;
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] *= inv;
        }
        st8_smem(dst + (size_t)row * LM_PITCH + lane * 8, v);
    }
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lm_smem(dst)), "l"(src) : "memory");
}
// asynchronously copy `rows` pixel rows (256 channels each) of a [., row_stride] tensor into smem [rows][LM_PITCH]
__device__ __forceinline__ void issue_rows(__nv_bfloat16* dst, const __nv_bfloat16* __restrict__ src, size_t row_stride, int rows) {
    const int lane = threadIdx.x & 31;
    for (int idx = threadIdx.x; idx < rows * 32; idx += blockDim.x) {
        const int row = idx >> 5;
        cp_async16(dst + (size_t)row * LM_PITCH + lane * 8, src + (size_t)row * row_stride + lane * 8);
    }
}
// in-place transform of landed rows: XF 1: exp(x - M[c]);  XF 2: exp(x - M[c]) * Zi[c];  XF 3: softmax_d * mul
template <int XF>
__device__ __forceinline__ void transform_rows(__nv_bfloat16* buf, int rows, const float* sM, const float* sZi, float mul) {
    const int lane = threadIdx.x & 31;
    for (int idx = threadIdx.x; idx < rows * 32; idx += blockDim.x) {
        __nv_bfloat16* p = buf + (size_t)(idx >> 5) * LM_PITCH + lane * 8;
        float v[8];
        ld8(p, v);
        if (XF == 1 || XF == 2) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float e = __expf(v[k] - sM[lane * 8 + k]);
                v[k] = (XF == 2) ? e * sZi[lane * 8 + k] : e;
            }
        } else {
            float mx = v[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) mx = fmaxf(mx, v[k]);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { v[k] = __expf(v[k] - mx); sum += v[k]; }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float inv = mul / sum;
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] *= inv;
        }
        st8_smem(p, v);
    }
}

// ---- context / dcontext: 2-deep cp.async ring of 64-pixel (W | V) tiles ------------------------------------------
constexpr int LC_ROWS = 64;
constexpr int LC_TILE_ELEMS = 2 * LC_ROWS * LM_PITCH;
template <int MODE>
__global__ void __launch_bounds__(256) la_ctx_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                         const __nv_bfloat16* __restrict__ dout,
                                                         const float* __restrict__ part, int n_stat_chunks,
                                                         float* __restrict__ kmax, float* __restrict__ kzinv,
                                                         float* __restrict__ ctx, int N, float scale) {
    extern __shared__ __align__(16) unsigned char raw[];
    __nv_bfloat16* ring = reinterpret_cast<__nv_bfloat16*>(raw);               // [2][W | V][64][LM_PITCH]
    float* sM = reinterpret_cast<float*>(ring + 2 * LC_TILE_ELEMS);
    float* sZi = sM + LM_HID;
    const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, h = tid >> 5;
    const int n_begin = chunk * LM_CHUNK, n_end = min(N, n_begin + LM_CHUNK);
    const int n_tiles = (n_end - n_begin) / LC_ROWS;
    auto issue = [&](int it) {
        __nv_bfloat16* buf = ring + (size_t)(it & 1) * LC_TILE_ELEMS;
        const size_t pix = (size_t)b * N + n_begin + (size_t)it * LC_ROWS;
        if (MODE == 0) {
            issue_rows(buf, qkv + pix * 3 * LM_HID + LM_HID, 3 * LM_HID, LC_ROWS);
            issue_rows(buf + LC_ROWS * LM_PITCH, qkv + pix * 3 * LM_HID + 2 * LM_HID, 3 * LM_HID, LC_ROWS);
        } else {
            issue_rows(buf, qkv + pix * 3 * LM_HID, 3 * LM_HID, LC_ROWS);
            issue_rows(buf + LC_ROWS * LM_PITCH, dout + pix * LM_HID, LM_HID, LC_ROWS);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (n_tiles > 0) issue(0);
    if (MODE == 0) {
        float M = -INFINITY;
        for (int i = 0; i < n_stat_chunks; ++i) M = fmaxf(M, part[(((size_t)b * n_stat_chunks + i) * LM_HID + tid) * 2]);
        float Z = 0.f;
        for (int i = 0; i < n_stat_chunks; ++i) {
            const float* p = part + (((size_t)b * n_stat_chunks + i) * LM_HID + tid) * 2;
            Z += p[1] * __expf(p[0] - M);
        }
        sM[tid] = M;
        sZi[tid] = 1.f / Z;
        if (chunk == 0) { kmax[(size_t)b * LM_HID + tid] = M; kzinv[(size_t)b * LM_HID + tid] = 1.f / Z; }
    }
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
    for (int it = 0; it < n_tiles; ++it) {
        __nv_bfloat16* Ws = ring + (size_t)(it & 1) * LC_TILE_ELEMS;
        __nv_bfloat16* Vs = Ws + LC_ROWS * LM_PITCH;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                       // tile `it` landed for everyone; the other slot is no longer being read
        if (it + 1 < n_tiles) issue(it + 1);
        if (MODE == 0) transform_rows<1>(Ws, LC_ROWS, sM, sZi, 1.f);
        else transform_rows<3>(Ws, LC_ROWS, nullptr, nullptr, scale);
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a0[4], a1[4], b01[4], b23[4];
            frag_a_kmajor(a0, Ws, LM_PITCH, ks * 16, h * LM_D, lane);
            frag_a_kmajor(a1, Ws, LM_PITCH, ks * 16, h * LM_D + 16, lane);
            frag_b_krows(b01, Vs, LM_PITCH, ks * 16, h * LM_D, lane);
            frag_b_krows(b23, Vs, LM_PITCH, ks * 16, h * LM_D + 16, lane);
            mma_bf16(acc[0][0], a0, b01[0], b01[1]); mma_bf16(acc[0][1], a0, b01[2], b01[3]);
            mma_bf16(acc[0][2], a0, b23[0], b23[1]); mma_bf16(acc[0][3], a0, b23[2], b23[3]);
            mma_bf16(acc[1][0], a1, b01[0], b01[1]); mma_bf16(acc[1][1], a1, b01[2], b01[3]);
            mma_bf16(acc[1][2], a1, b23[0], b23[1]); mma_bf16(acc[1][3], a1, b23[2], b23[3]);
        }
    }
    const int g = lane >> 2, t = lane & 3;
    float* cb = ctx + ((size_t)b * LM_HEADS + h) * LM_D * LM_D;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int d = mt * 16 + g + half * 8;
            const float f = (MODE == 0) ? sZi[h * LM_D + d] / (float)N : 1.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int e = nt * 8 + 2 * t;
                atomicAdd(cb + d * LM_D + e, acc[mt][nt][half * 2] * f);
                atomicAdd(cb + d * LM_D + e + 1, acc[mt][nt][half * 2 + 1] * f);
            }
        }
}

__device__ __forceinline__ void stage_ctx_bf16(const float* __restrict__ src, __nv_bfloat16* dst) {
    for (int i = threadIdx.x; i < LM_HEADS * LM_D * LM_D; i += blockDim.x) {
        const int e = i & 31, d = (i >> 5) & 31, h = i >> 10;
        dst[(size_t)(h * LM_D + d) * LM_CPITCH + e] = __float2bfloat16_rn(src[i]);
    }
}

// ---- out = q~ ctx: 3-deep cp.async ring of 64-pixel q tiles ----------------------------------------------------------
constexpr int LO_ROWS = 64, LO_STAGES = 3;
__global__ void __launch_bounds__(256) la_out_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                         const float* __restrict__ ctx, __nv_bfloat16* __restrict__ out,
                                                         int N, float scale) {
    extern __shared__ __align__(16) unsigned char raw[];
    __nv_bfloat16* Cs = reinterpret_cast<__nv_bfloat16*>(raw);                 // [8*32][LM_CPITCH]
    __nv_bfloat16* ring = Cs + LM_HEADS * LM_D * LM_CPITCH;                    // [3][64][LM_PITCH]
    __nv_bfloat16* Os = ring + LO_STAGES * LO_ROWS * LM_PITCH;                 // [64][LM_PITCH]
    const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, h = tid >> 5;
    const int n_begin = chunk * LM_CHUNK, n_end = min(N, n_begin + LM_CHUNK);
    const int n_tiles = (n_end - n_begin) / LO_ROWS;
    const size_t pix_base = (size_t)b * N + n_begin;
    auto issue = [&](int it) {
        issue_rows(ring + (size_t)(it % LO_STAGES) * LO_ROWS * LM_PITCH, qkv + (pix_base + (size_t)it * LO_ROWS) * 3 * LM_HID,
                   3 * LM_HID, LO_ROWS);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (n_tiles > 0) issue(0);
    if (n_tiles > 1) issue(1);
    stage_ctx_bf16(ctx + (size_t)b * LM_HEADS * LM_D * LM_D, Cs);
    __syncthreads();
    // B[k = d][n = e] = ctx[d][e]: rows of Cs are the K index
    uint32_t bf[2][2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int np = 0; np < 2; ++np) frag_b_krows(bf[ks][np], Cs + (size_t)h * LM_D * LM_CPITCH, LM_CPITCH, ks * 16, np * 16, lane);
    const int g = lane >> 2, t = lane & 3;
    for (int it = 0; it < n_tiles; ++it) {
        __nv_bfloat16* Qs = ring + (size_t)(it % LO_STAGES) * LO_ROWS * LM_PITCH;
        if (it + 1 < n_tiles) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                       // tile landed; previous Os copy-out and previous q tile reads are done
        if (it + 2 < n_tiles) issue(it + 2);
        transform_rows<3>(Qs, LO_ROWS, nullptr, nullptr, scale);
        __syncthreads();
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            float c[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) c[i][k] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t a[4];
                frag_a_rowmajor(a, Qs, LM_PITCH, mt * 16, h * LM_D + ks * 16, lane);
                mma_bf16(c[0], a, bf[ks][0][0], bf[ks][0][1]); mma_bf16(c[1], a, bf[ks][0][2], bf[ks][0][3]);
                mma_bf16(c[2], a, bf[ks][1][0], bf[ks][1][1]); mma_bf16(c[3], a, bf[ks][1][2], bf[ks][1][3]);
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                __nv_bfloat16* o = Os + (size_t)(mt * 16 + g) * LM_PITCH + h * LM_D + nt * 8 + 2 * t;
                *reinterpret_cast<uint32_t*>(o) = pack_bf16(c[nt][0], c[nt][1]);
                *reinterpret_cast<uint32_t*>(o + 8 * LM_PITCH) = pack_bf16(c[nt][2], c[nt][3]);
            }
        }
        __syncthreads();
        for (int idx = tid; idx < LO_ROWS * 32; idx += 256) {
            const int row = idx >> 5, oc = idx & 31;
            *reinterpret_cast<uint4*>(out + (pix_base + (size_t)it * LO_ROWS + row) * LM_HID + oc * 8) =
                *reinterpret_cast<const uint4*>(Os + (size_t)row * LM_PITCH + oc * 8);
        }
    }
}

// ---- backward per pixel ---------------------------------------------------------------------------------------------
// Streaming kernel: a 3-deep cp.async ring of raw 16-pixel tiles (dout | q | k | v rows, 32 KiB per tile) keeps two
// tiles in flight per SM while the current one is transformed in place (q -> softmax, k -> k~), multiplied on the
// tensor cores and written back through a staging tile with 16-byte coalesced stores.
constexpr int LB_ROWS = 16;
constexpr int LB_OPITCH = 3 * LM_HID + 8;
constexpr int LB_STAGES = 3;
constexpr int LB_TILE_ELEMS = 4 * LB_ROWS * LM_PITCH;        // dout, q, k, v

__device__ __forceinline__ void lb_issue_tile(__nv_bfloat16* buf, const __nv_bfloat16* __restrict__ qkv,
                                              const __nv_bfloat16* __restrict__ dout, size_t pix0) {
    const int lane = threadIdx.x & 31;
    for (int idx = threadIdx.x; idx < LB_ROWS * 32; idx += blockDim.x) {
        const int row = idx >> 5;
        const __nv_bfloat16* qrow = qkv + (pix0 + row) * 3 * LM_HID + lane * 8;
        __nv_bfloat16* d = buf + (size_t)row * LM_PITCH + lane * 8;
        cp_async16(d, dout + (pix0 + row) * LM_HID + lane * 8);
        cp_async16(d + LB_ROWS * LM_PITCH, qrow);
        cp_async16(d + 2 * LB_ROWS * LM_PITCH, qrow + LM_HID);
        cp_async16(d + 3 * LB_ROWS * LM_PITCH, qrow + 2 * LM_HID);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// in-place transforms of a landed tile: q rows -> softmax_d (unscaled p), k rows -> exp(k - M) * Zinv
__device__ __forceinline__ void lb_transform_tile(__nv_bfloat16* buf, const float* sM, const float* sZi) {
    const int lane = threadIdx.x & 31;
    for (int idx = threadIdx.x; idx < LB_ROWS * 32; idx += blockDim.x) {
        const int row = idx >> 5;
        __nv_bfloat16* q = buf + (size_t)(LB_ROWS + row) * LM_PITCH + lane * 8;
        __nv_bfloat16* k = buf + (size_t)(2 * LB_ROWS + row) * LM_PITCH + lane * 8;
        float v[8];
        ld8(q, v);
        float mx = v[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) mx = fmaxf(mx, v[i]);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] = __expf(v[i] - mx); sum += v[i]; }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float inv = 1.f / sum;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= inv;
        st8_smem(q, v);
        ld8(k, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __expf(v[i] - sM[lane * 8 + i]) * sZi[lane * 8 + i];
        st8_smem(k, v);
    }
}

__global__ void __launch_bounds__(256) la_bwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                         const __nv_bfloat16* __restrict__ dout,
                                                         const float* __restrict__ ctx, const float* __restrict__ dctx,
                                                         const float* __restrict__ kmax, const float* __restrict__ kzinv,
                                                         __nv_bfloat16* __restrict__ dqkv, int N, float scale) {
    extern __shared__ __align__(16) unsigned char raw[];
    __nv_bfloat16* Cs = reinterpret_cast<__nv_bfloat16*>(raw);                 // ctx  [8*32][LM_CPITCH]
    __nv_bfloat16* Ds = Cs + LM_HEADS * LM_D * LM_CPITCH;                      // dctx [8*32][LM_CPITCH]
    __nv_bfloat16* ring = Ds + LM_HEADS * LM_D * LM_CPITCH;                    // [LB_STAGES][4][16][LM_PITCH]
    __nv_bfloat16* Os = ring + LB_STAGES * LB_TILE_ELEMS;                      // [16][LB_OPITCH]: dq | dk | dv
    float* sM = reinterpret_cast<float*>(Os + LB_ROWS * LB_OPITCH);
    float* sZi = sM + LM_HID;
    float* scd = sZi + LM_HID;
    const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, h = tid >> 5;
    const int n_begin = chunk * LM_CHUNK, n_end = min(N, n_begin + LM_CHUNK);
    const int n_tiles = (n_end - n_begin) / LB_ROWS;
    const size_t pix_base = (size_t)b * N + n_begin;
    // start the pipeline before the (global-memory) prologue so that both overlap
    if (n_tiles > 0) lb_issue_tile(ring, qkv, dout, pix_base);
    if (n_tiles > 1) lb_issue_tile(ring + LB_TILE_ELEMS, qkv, dout, pix_base + LB_ROWS);
    const float* cg = ctx + (size_t)b * LM_HEADS * LM_D * LM_D;
    const float* dg = dctx + (size_t)b * LM_HEADS * LM_D * LM_D;
    stage_ctx_bf16(cg, Cs);
    stage_ctx_bf16(dg, Ds);
    sM[tid] = kmax[(size_t)b * LM_HID + tid];
    sZi[tid] = kzinv[(size_t)b * LM_HID + tid];
    {
        float s = 0.f;                 // cd[h][d] = sum_e dctx[h][d][e] ctx[h][d][e]   (tid = h*32 + d)
#pragma unroll
        for (int e = 0; e < LM_D; ++e) s += dg[(size_t)tid * LM_D + e] * cg[(size_t)tid * LM_D + e];
        scd[tid] = s;
    }
    __syncthreads();
    const __nv_bfloat16* Ch = Cs + (size_t)h * LM_D * LM_CPITCH;
    const __nv_bfloat16* Dh = Ds + (size_t)h * LM_D * LM_CPITCH;
    uint32_t bc[2][2][4], bd[2][2][4], bt[2][2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int np = 0; np < 2; ++np) {
            frag_b_nrows(bc[ks][np], Ch, LM_CPITCH, np * 16, ks * 16, lane);      // B[k=e][n=d] = ctx[d][e]
            frag_b_nrows(bd[ks][np], Dh, LM_CPITCH, np * 16, ks * 16, lane);      // B[k=e][n=d] = dctx[d][e]
            frag_b_krows(bt[ks][np], Dh, LM_CPITCH, ks * 16, np * 16, lane);      // B[k=d][n=e] = dctx[d][e]
        }
    const int g = lane >> 2, t = lane & 3;
    const float invN = 1.f / (float)N;
    for (int it = 0; it < n_tiles; ++it) {
        __nv_bfloat16* buf = ring + (size_t)(it % LB_STAGES) * LB_TILE_ELEMS;
        if (it + 1 < n_tiles) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                   // tile `it` visible to all; Os of tile it-1 fully copied out
        if (it + 2 < n_tiles)                              // ring slot (it+2)%3 held tile it-1: free since the sync above
            lb_issue_tile(ring + (size_t)((it + 2) % LB_STAGES) * LB_TILE_ELEMS, qkv, dout, pix_base + (size_t)(it + 2) * LB_ROWS);
        lb_transform_tile(buf, sM, sZi);
        __syncthreads();
        const __nv_bfloat16* T0 = buf;                         // dout
        const __nv_bfloat16* T1 = buf + LB_ROWS * LM_PITCH;    // p
        const __nv_bfloat16* T2 = buf + 2 * LB_ROWS * LM_PITCH;  // k~
        const __nv_bfloat16* T3 = buf + 3 * LB_ROWS * LM_PITCH;  // v
        float cq[4][4], ck[4][4], cv[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) { cq[i][k] = 0.f; ck[i][k] = 0.f; cv[i][k] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t ag[4], av[4], ak[4];
            frag_a_rowmajor(ag, T0, LM_PITCH, 0, h * LM_D + ks * 16, lane);      // dout [px][e]
            frag_a_rowmajor(av, T3, LM_PITCH, 0, h * LM_D + ks * 16, lane);      // v    [px][e]
            frag_a_rowmajor(ak, T2, LM_PITCH, 0, h * LM_D + ks * 16, lane);      // k~   [px][d]
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                mma_bf16(cq[np * 2], ag, bc[ks][np][0], bc[ks][np][1]); mma_bf16(cq[np * 2 + 1], ag, bc[ks][np][2], bc[ks][np][3]);
                mma_bf16(ck[np * 2], av, bd[ks][np][0], bd[ks][np][1]); mma_bf16(ck[np * 2 + 1], av, bd[ks][np][2], bd[ks][np][3]);
                mma_bf16(cv[np * 2], ak, bt[ks][np][0], bt[ks][np][1]); mma_bf16(cv[np * 2 + 1], ak, bt[ks][np][2], bt[ks][np][3]);
            }
        }
        // dq = p * (dp - sum_d p dp), dp = scale * (dout ctx^T);  dk = k~ * (dk~ - cd), dk~ = (v/N) dctx^T;  dv = (k~ dctx)/N
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = g + half * 8;
            float pv[4][2], dot = 0.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const __nv_bfloat162 p2 = *reinterpret_cast<const __nv_bfloat162*>(T1 + (size_t)row * LM_PITCH + h * LM_D + nt * 8 + 2 * t);
                pv[nt][0] = __low2float(p2); pv[nt][1] = __high2float(p2);
                dot += pv[nt][0] * cq[nt][half * 2] + pv[nt][1] * cq[nt][half * 2 + 1];
            }
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int col = h * LM_D + nt * 8 + 2 * t;
                __nv_bfloat16* o = Os + (size_t)row * LB_OPITCH + col;
                *reinterpret_cast<uint32_t*>(o) = pack_bf16(scale * pv[nt][0] * (cq[nt][half * 2] - dot),
                                                           scale * pv[nt][1] * (cq[nt][half * 2 + 1] - dot));
                const __nv_bfloat162 k2 = *reinterpret_cast<const __nv_bfloat162*>(T2 + (size_t)row * LM_PITCH + col);
                *reinterpret_cast<uint32_t*>(o + LM_HID) =
                    pack_bf16(__low2float(k2) * (ck[nt][half * 2] * invN - scd[col]),
                              __high2float(k2) * (ck[nt][half * 2 + 1] * invN - scd[col + 1]));
                *reinterpret_cast<uint32_t*>(o + 2 * LM_HID) = pack_bf16(cv[nt][half * 2] * invN, cv[nt][half * 2 + 1] * invN);
            }
        }
        __syncthreads();
        for (int idx = tid; idx < LB_ROWS * 96; idx += 256) {
            const int row = idx / 96, oc = idx - row * 96;
            *reinterpret_cast<uint4*>(dqkv + (pix_base + (size_t)it * LB_ROWS + row) * 3 * LM_HID + oc * 8) =
                *reinterpret_cast<const uint4*>(Os + (size_t)row * LB_OPITCH + oc * 8);
        }
    }
}

constexpr size_t LA_CTX_SMEM = (size_t)2 * LC_TILE_ELEMS * 2 + 2 * LM_HID * 4;
constexpr size_t LA_OUT_SMEM = (size_t)LM_HEADS * LM_D * LM_CPITCH * 2 + (size_t)(LO_STAGES + 1) * LO_ROWS * LM_PITCH * 2;
constexpr size_t LA_BWD_SMEM = (size_t)2 * LM_HEADS * LM_D * LM_CPITCH * 2 + (size_t)LB_STAGES * LB_TILE_ELEMS * 2 +
                               (size_t)LB_ROWS * LB_OPITCH * 2 + 3 * LM_HID * 4;

static int la_mma_attrs() {
    static bool done = false;
    if (!done) {
        PIDM_CUDA(cudaFuncSetAttribute(la_ctx_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_CTX_SMEM));
        PIDM_CUDA(cudaFuncSetAttribute(la_ctx_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_CTX_SMEM));
        PIDM_CUDA(cudaFuncSetAttribute(la_out_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_OUT_SMEM));
        PIDM_CUDA(cudaFuncSetAttribute(la_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_BWD_SMEM));
        done = true;
    }
    return 0;
}

// entry points used by attention.cu for the bf16 / 8-head case
int la_mma_ctx(int mode, const void* qkv, const void* dout, const float* part, int n_stat_chunks, float* kmax,
               float* kzinv, float* ctx, int B, int N, float scale, cudaStream_t st) {
    if (int e = la_mma_attrs()) return e;
    dim3 grid((N + LM_CHUNK - 1) / LM_CHUNK, B);
    if (mode == 0)
        la_ctx_mma_kernel<0><<<grid, 256, LA_CTX_SMEM, st>>>((const __nv_bfloat16*)qkv, nullptr, part, n_stat_chunks, kmax,
                                                            kzinv, ctx, N, scale);
    else
        la_ctx_mma_kernel<1><<<grid, 256, LA_CTX_SMEM, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout, nullptr,
                                                            0, nullptr, nullptr, ctx, N, scale);
    PIDM_LAUNCH_CHECK("la_ctx_mma");
    return 0;
}
int la_mma_out(const void* qkv, const float* ctx, void* out, int B, int N, float scale, cudaStream_t st) {
    if (int e = la_mma_attrs()) return e;
    dim3 grid((N + LM_CHUNK - 1) / LM_CHUNK, B);
    la_out_mma_kernel<<<grid, 256, LA_OUT_SMEM, st>>>((const __nv_bfloat16*)qkv, ctx, (__nv_bfloat16*)out, N, scale);
    PIDM_LAUNCH_CHECK("la_out_mma");
    return 0;
}
int la_mma_bwd(const void* qkv, const void* dout, const float* ctx, const float* dctx, const float* kmax,
               const float* kzinv, void* dqkv, int B, int N, float scale, cudaStream_t st) {
    if (int e = la_mma_attrs()) return e;
    dim3 grid((N + LM_CHUNK - 1) / LM_CHUNK, B);
    la_bwd_mma_kernel<<<grid, 256, LA_BWD_SMEM, st>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout, ctx, dctx, kmax,
                                                     kzinv, (__nv_bfloat16*)dqkv, N, scale);
    PIDM_LAUNCH_CHECK("la_bwd_mma");
    return 0;
}

}  // namespace pidm
