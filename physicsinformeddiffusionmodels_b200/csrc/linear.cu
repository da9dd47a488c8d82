// Time-conditioning path of the U-Net, fp32 throughout (B x 128 activations: latency-, not throughput-bound):
//   * time_embed : SinusoidalPosEmb(dim) -> Linear(dim,4dim) -> GELU(erf) -> Linear(4dim,4dim) (+ SiLU of it)
//                  reference unet_model.py:147-159, 464-469
//   * block_mlps : every ResnetBlock's Linear(4dim, 2*C_out) on SiLU(t) (unet_model.py:246-249,258-262),
//                  ALL blocks in one launch through a device-side table (they share the same input).
#include "common.cuh"
#include "pidm.h"

namespace pidm {

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

// one CTA per sample, blockDim = td
__global__ void time_embed_fwd_kernel(const long long* __restrict__ t, const float* __restrict__ W1,
                                      const float* __restrict__ b1, const float* __restrict__ W2,
                                      const float* __restrict__ b2, float* __restrict__ emb /*[B,dim]*/,
                                      float* __restrict__ h1 /*[B,td]*/, float* __restrict__ temb /*[B,td]*/,
                                      float* __restrict__ silu_t /*[B,td]*/, int dim, int td) {
    extern __shared__ float sm[];   // e[dim] | a1[td]
    float* e = sm;
    float* a1 = sm + dim;
    const int b = blockIdx.x, j = threadIdx.x;
    const int half = dim / 2;
    if (j < dim) {
        int k = j < half ? j : j - half;
        float f = expf((float)k * -(logf(10000.f) / (float)(half - 1)));
        float arg = (float)t[b] * f;
        float v = j < half ? sinf(arg) : cosf(arg);
        e[j] = v;
        emb[(size_t)b * dim + j] = v;
    }
    __syncthreads();
    float acc = b1[j];
    for (int k = 0; k < dim; ++k) acc += W1[(size_t)j * dim + k] * e[k];
    h1[(size_t)b * td + j] = acc;
    a1[j] = gelu_erf(acc);
    __syncthreads();
    float o = b2[j];
    for (int k = 0; k < td; ++k) o += W2[(size_t)j * td + k] * a1[k];
    temb[(size_t)b * td + j] = o;
    silu_t[(size_t)b * td + j] = silu_f(o);
}

// backward: given d_silu_t; gradients of W1,b1,W2,b2 are ACCUMULATED with atomics (B CTAs).
__global__ void time_embed_bwd_kernel(const float* __restrict__ d_silu, const float* __restrict__ emb,
                                      const float* __restrict__ h1, const float* __restrict__ temb,
                                      const float* __restrict__ W2, float* __restrict__ dW1, float* __restrict__ db1,
                                      float* __restrict__ dW2, float* __restrict__ db2, int dim, int td) {
    extern __shared__ float sm[];   // e[dim] | a1[td] | dt[td] | dh[td]
    float* e = sm;
    float* a1 = e + dim;
    float* dt = a1 + td;
    float* dh = dt + td;
    const int b = blockIdx.x, j = threadIdx.x;
    if (j < dim) e[j] = emb[(size_t)b * dim + j];
    a1[j] = gelu_erf(h1[(size_t)b * td + j]);
    float dtj = d_silu[(size_t)b * td + j] * silu_grad_f(temb[(size_t)b * td + j]);
    dt[j] = dtj;
    __syncthreads();
    atomicAdd(&db2[j], dtj);
    for (int k = 0; k < td; ++k) atomicAdd(&dW2[(size_t)j * td + k], dtj * a1[k]);
    // da1[j] = sum_i dt[i] W2[i][j]  (column read: coalesced across threads j)
    float da = 0.f;
    for (int i = 0; i < td; ++i) da += dt[i] * W2[(size_t)i * td + j];
    float dhj = da * gelu_erf_grad(h1[(size_t)b * td + j]);
    dh[j] = dhj;
    atomicAdd(&db1[j], dhj);
    for (int k = 0; k < dim; ++k) atomicAdd(&dW1[(size_t)j * dim + k], dhj * e[k]);
}

struct MlpEntry {
    const float* W;      // [n, td]
    const float* b;      // [n]
    float* dW;           // accumulated
    float* db;           // accumulated
    float* out;          // [B, n]
    const float* dout;   // [B, n]
    int n, pad_;
};

constexpr int MLP_BCHUNK = 32;
// grid (entries, row-chunks of 64); warp per row, lanes over k; out[b, off+j] = b[j] + W[j,:] . s[b,:]
__global__ void block_mlps_fwd_kernel(const MlpEntry* __restrict__ table, const float* __restrict__ s /*[B,td]*/,
                                      int B, int td) {
    extern __shared__ float ss[];   // [MLP_BCHUNK][td]
    const MlpEntry e = table[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int r0 = blockIdx.y * 64;
    if (r0 >= e.n) return;
    for (int b0 = 0; b0 < B; b0 += MLP_BCHUNK) {
        int nb = min(MLP_BCHUNK, B - b0);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * td; i += blockDim.x) ss[i] = s[(size_t)b0 * td + i];
        __syncthreads();
        for (int j = r0 + warp; j < min(r0 + 64, e.n); j += nw) {
            const float* wr = e.W + (size_t)j * td;
            float bj = e.b[j];
            for (int b = 0; b < nb; ++b) {
                float acc = 0.f;
                for (int k = lane; k < td; k += 32) acc += wr[k] * ss[b * td + k];
                acc = warp_sum(acc);
                if (lane == 0) e.out[(size_t)(b0 + b) * e.n + j] = acc + bj;
            }
        }
    }
}

// dW[j,k] += sum_b d[b, off+j] s[b,k] ; db[j] += sum_b d[b, off+j]    (row j owned by exactly one warp)
__global__ void block_mlps_wgrad_kernel(const MlpEntry* __restrict__ table, const float* __restrict__ s, int B,
                                        int td) {
    extern __shared__ float ss[];
    const MlpEntry e = table[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int r0 = blockIdx.y * 64;
    if (r0 >= e.n) return;
    for (int b0 = 0; b0 < B; b0 += MLP_BCHUNK) {
        int nb = min(MLP_BCHUNK, B - b0);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * td; i += blockDim.x) ss[i] = s[(size_t)b0 * td + i];
        __syncthreads();
        for (int j = r0 + warp; j < min(r0 + 64, e.n); j += nw) {
            float* wr = e.dW + (size_t)j * td;
            float dbj = 0.f;
            for (int k = lane; k < td; k += 32) {
                float acc = 0.f;
                for (int b = 0; b < nb; ++b) acc += e.dout[(size_t)(b0 + b) * e.n + j] * ss[b * td + k];
                wr[k] += acc;
            }
            if (lane == 0) {
                for (int b = 0; b < nb; ++b) dbj += e.dout[(size_t)(b0 + b) * e.n + j];
                e.db[j] += dbj;
            }
        }
    }
}

// d_s[b,k] += sum_j dout_e[b, j] W_e[j,k]     grid (B, entries), block td; ds zeroed by the caller
__global__ void block_mlps_dgrad_kernel(const MlpEntry* __restrict__ table, float* __restrict__ ds, int td) {
    extern __shared__ float sd[];   // [max_rows]
    const int b = blockIdx.x, k = threadIdx.x;
    const MlpEntry e = table[blockIdx.y];
    for (int j = threadIdx.x; j < e.n; j += blockDim.x) sd[j] = e.dout[(size_t)b * e.n + j];
    __syncthreads();
    float acc = 0.f;
#pragma unroll 4
    for (int j = 0; j < e.n; ++j) acc += sd[j] * e.W[(size_t)j * td + k];
    atomicAdd(&ds[(size_t)b * td + k], acc);
}

}  // namespace pidm
using namespace pidm;

extern "C" int pidm_time_embed_fwd(const long long* t, const float* W1, const float* b1, const float* W2,
                                   const float* b2, float* emb, float* h1, float* temb, float* silu_t, int B, int dim,
                                   int td, void* stream) {
    PIDM_REQUIRE(td <= 1024 && dim <= td && dim % 2 == 0 && dim >= 4, "time_embed: need 4<=dim<=td<=1024, dim even");
    time_embed_fwd_kernel<<<B, td, (dim + td) * sizeof(float), (cudaStream_t)stream>>>(t, W1, b1, W2, b2, emb, h1, temb,
                                                                                      silu_t, dim, td);
    PIDM_LAUNCH_CHECK("time_embed_fwd");
    return 0;
}

extern "C" int pidm_time_embed_bwd(const float* d_silu_t, const float* emb, const float* h1, const float* temb,
                                   const float* W2, float* dW1, float* db1, float* dW2, float* db2, int B, int dim,
                                   int td, void* stream) {
    PIDM_REQUIRE(td <= 1024 && dim <= td, "time_embed_bwd: need dim<=td<=1024");
    time_embed_bwd_kernel<<<B, td, (dim + 3 * td) * sizeof(float), (cudaStream_t)stream>>>(
        d_silu_t, emb, h1, temb, W2, dW1, db1, dW2, db2, dim, td);
    PIDM_LAUNCH_CHECK("time_embed_bwd");
    return 0;
}

extern "C" int pidm_mlp_entry_size(void) { return (int)sizeof(MlpEntry); }

static int mlp_smem_attr(size_t bytes) {
    static bool done = false;
    if (!done && bytes > 48 * 1024) {
        PIDM_CUDA(cudaFuncSetAttribute(block_mlps_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        PIDM_CUDA(cudaFuncSetAttribute(block_mlps_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        done = true;
    }
    return 0;
}

extern "C" int pidm_block_mlps_fwd(const void* table_dev, int n_entries, int max_rows, const float* silu_t, int B,
                                   int td, void* stream) {
    PIDM_REQUIRE(td <= 768, "block_mlps: td <= 768 required");
    size_t smem = (size_t)MLP_BCHUNK * td * sizeof(float);
    if (int e = mlp_smem_attr(smem)) return e;
    dim3 grid(n_entries, ceil_div(max_rows, 64));
    block_mlps_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const MlpEntry*)table_dev, silu_t, B, td);
    PIDM_LAUNCH_CHECK("block_mlps_fwd");
    return 0;
}

// weight/bias grads accumulate into the table's dW/db pointers; d_silu_t is overwritten.
extern "C" int pidm_block_mlps_bwd(const void* table_dev, int n_entries, int max_rows, const float* silu_t,
                                   float* d_silu_t, int B, int td, void* stream) {
    PIDM_REQUIRE(td <= 768, "block_mlps: td <= 768 required");
    cudaStream_t st = (cudaStream_t)stream;
    size_t smem = (size_t)MLP_BCHUNK * td * sizeof(float);
    if (int e = mlp_smem_attr(smem)) return e;
    dim3 grid(n_entries, ceil_div(max_rows, 64));
    block_mlps_wgrad_kernel<<<grid, 256, smem, st>>>((const MlpEntry*)table_dev, silu_t, B, td);
    PIDM_CUDA(cudaMemsetAsync(d_silu_t, 0, (size_t)B * td * sizeof(float), st));
    block_mlps_dgrad_kernel<<<dim3(B, n_entries), td, max_rows * sizeof(float), st>>>((const MlpEntry*)table_dev, d_silu_t, td);
    PIDM_LAUNCH_CHECK("block_mlps_bwd");
    return 0;
}
