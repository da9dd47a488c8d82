// Linear-elasticity (topology-optimisation) residual, MATRIX-FREE.
// Reference src/residuals_mechanics_K.py:198-274 assembles a dense B x 8450 x 8450 stiffness matrix
// (285.6 MB per sample) by index_put, replaces Dirichlet rows by identity rows and multiplies by u.
// Here K(rho) u is evaluated by gathering, for every node, the <= 4 adjacent Q4 elements:
//     (K u)_i = sum_{e ni i} rho_e sum_j KE[loc_e(i)][j] u_{dof_e(j)}
// which touches ~150 KB per sample instead of >1 GB.  Node id = row*(nel+1)+col, dof = 2*node+d, element
// (er,ec) has nodes n1=(er+1,ec), n2=(er+1,ec+1), n3=(er,ec+1), n4=(er,ec) (counter-clockwise, y up).
//   residual_i = mask_i ? u_i : (K u)_i - f_i ;   compliance = sum_i u_i * (mask_i ? u_i : (K u)_i)
// Also: bilinear resize (torchvision Resize(antialias=False) == align_corners=False), reference :10-21.
#include "common.cuh"
#include "pidm.h"

namespace pidm {

__constant__ float c_KE[64];

// (K v)_{node (r,c), both dofs} for a nodal field v [2][nn][nn] of sample b
__device__ __forceinline__ void kv_node(const float* __restrict__ v, const float* __restrict__ rho, int nel, int r, int c,
                                        float& o0, float& o1) {
    const int nn = nel + 1;
    o0 = 0.f; o1 = 0.f;
    // adjacent elements and the local index of this node inside them
    const int der[4] = {-1, -1, 0, 0}, dec[4] = {-1, 0, -1, 0}, loc[4] = {1, 0, 2, 3};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        int er = r + der[a], ec = c + dec[a];
        if (er < 0 || er >= nel || ec < 0 || ec >= nel) continue;
        float re = rho[er * nel + ec];
        const int nr[4] = {er + 1, er + 1, er, er}, nc[4] = {ec, ec + 1, ec + 1, ec};
        float ue[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ue[2 * k] = v[nr[k] * nn + nc[k]];
            ue[2 * k + 1] = v[nn * nn + nr[k] * nn + nc[k]];
        }
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { s0 += c_KE[(2 * loc[a]) * 8 + j] * ue[j]; s1 += c_KE[(2 * loc[a] + 1) * 8 + j] * ue[j]; }
        o0 += re * s0; o1 += re * s1;
    }
}

// MODE 0: forward.  MODE 1: backward stage A (z and the direct terms).  MODE 2: backward stage B (du += K z).
template <int MODE>
__global__ void mech_node_kernel(const float* __restrict__ u, const float* __restrict__ rho, const float* __restrict__ bcs,
                                 float* __restrict__ residual, float* __restrict__ compliance,
                                 const float* __restrict__ g_r, const float* __restrict__ g_c, float* __restrict__ z,
                                 float* __restrict__ du, int nel) {
    const int nn = nel + 1, b = blockIdx.y;
    const float* ub = u + (size_t)b * 2 * nn * nn;
    const float* rb = rho + (size_t)b * nel * nel;
    const float* bb = bcs + (size_t)b * 4 * nn * nn;
    float csum = 0.f;
    for (int node = blockIdx.x * blockDim.x + threadIdx.x; node < nn * nn; node += gridDim.x * blockDim.x) {
        int r = node / nn, c = node - r * nn;
        if (MODE == 2) {
            float k0, k1;
            kv_node(z + (size_t)b * 2 * nn * nn, rb, nel, r, c, k0, k1);
            du[(size_t)b * 2 * nn * nn + node] += k0;
            du[(size_t)b * 2 * nn * nn + nn * nn + node] += k1;
            continue;
        }
        float k0, k1;
        kv_node(ub, rb, nel, r, c, k0, k1);
        const bool m0 = bb[node] != 0.f, m1 = bb[nn * nn + node] != 0.f;
        const float u0 = ub[node], u1 = ub[nn * nn + node];
        const float w0 = m0 ? u0 : k0, w1 = m1 ? u1 : k1;
        if (MODE == 0) {
            const float f0 = m0 ? 0.f : bb[2 * nn * nn + node], f1 = m1 ? 0.f : bb[3 * nn * nn + node];
            residual[(size_t)b * 2 * nn * nn + 2 * node] = w0 - f0;
            residual[(size_t)b * 2 * nn * nn + 2 * node + 1] = w1 - f1;
            csum += u0 * w0 + u1 * w1;
        } else {
            const float gc = g_c ? g_c[b] : 0.f;
            const float wb0 = (g_r ? g_r[(size_t)b * 2 * nn * nn + 2 * node] : 0.f) + gc * u0;
            const float wb1 = (g_r ? g_r[(size_t)b * 2 * nn * nn + 2 * node + 1] : 0.f) + gc * u1;
            z[(size_t)b * 2 * nn * nn + node] = m0 ? 0.f : wb0;
            z[(size_t)b * 2 * nn * nn + nn * nn + node] = m1 ? 0.f : wb1;
            du[(size_t)b * 2 * nn * nn + node] = gc * w0 + (m0 ? wb0 : 0.f);
            du[(size_t)b * 2 * nn * nn + nn * nn + node] = gc * w1 + (m1 ? wb1 : 0.f);
        }
    }
    if (MODE == 0 && compliance) {
        csum = warp_sum(csum);
        if ((threadIdx.x & 31) == 0) atomicAdd(&compliance[b], csum);
    }
}

// d rho_e = z_e^T KE u_e
__global__ void mech_drho_kernel(const float* __restrict__ u, const float* __restrict__ z, float* __restrict__ drho,
                                 int nel) {
    const int nn = nel + 1, b = blockIdx.y;
    const float* ub = u + (size_t)b * 2 * nn * nn;
    const float* zb = z + (size_t)b * 2 * nn * nn;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nel * nel; e += gridDim.x * blockDim.x) {
        int er = e / nel, ec = e - er * nel;
        const int nr[4] = {er + 1, er + 1, er, er}, nc[4] = {ec, ec + 1, ec + 1, ec};
        float ue[8], ze[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ue[2 * k] = ub[nr[k] * nn + nc[k]]; ue[2 * k + 1] = ub[nn * nn + nr[k] * nn + nc[k]];
            ze[2 * k] = zb[nr[k] * nn + nc[k]]; ze[2 * k + 1] = zb[nn * nn + nr[k] * nn + nc[k]];
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) t += c_KE[i * 8 + j] * ue[j];
            s += ze[i] * t;
        }
        drho[(size_t)b * nel * nel + e] = s;
    }
}

// bilinear resize, align_corners=False, no antialias (matches F.interpolate / torchvision Resize(antialias=False))
__device__ __forceinline__ void bil_src(int o, float scale, int in, int& i0, int& i1, float& w1) {
    float s = ((float)o + 0.5f) * scale - 0.5f;
    if (s < 0.f) s = 0.f;
    i0 = (int)s;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + ((i0 < in - 1) ? 1 : 0);
    w1 = s - (float)i0;
}
__global__ void bilinear_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, int in, int out) {
    const float scale = (float)in / (float)out;
    long long total = (long long)planes * out * out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ow = (int)(i % out), oh = (int)((i / out) % out);
        long long pl = i / ((long long)out * out);
        int h0, h1, w0, w1; float lh, lw;
        bil_src(oh, scale, in, h0, h1, lh);
        bil_src(ow, scale, in, w0, w1, lw);
        const float* p = x + pl * in * in;
        y[i] = (1.f - lh) * ((1.f - lw) * p[h0 * in + w0] + lw * p[h0 * in + w1]) +
               lh * ((1.f - lw) * p[h1 * in + w0] + lw * p[h1 * in + w1]);
    }
}
__global__ void bilinear_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int planes, int in, int out) {
    const float scale = (float)in / (float)out;
    long long total = (long long)planes * out * out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ow = (int)(i % out), oh = (int)((i / out) % out);
        long long pl = i / ((long long)out * out);
        int h0, h1, w0, w1; float lh, lw;
        bil_src(oh, scale, in, h0, h1, lh);
        bil_src(ow, scale, in, w0, w1, lw);
        float g = dy[i];
        float* p = dx + pl * in * in;
        atomicAdd(&p[h0 * in + w0], g * (1.f - lh) * (1.f - lw));
        atomicAdd(&p[h0 * in + w1], g * (1.f - lh) * lw);
        atomicAdd(&p[h1 * in + w0], g * lh * (1.f - lw));
        atomicAdd(&p[h1 * in + w1], g * lh * lw);
    }
}

static int upload_ke(const float* KE_dev, cudaStream_t st) {
    PIDM_CUDA(cudaMemcpyToSymbolAsync(c_KE, KE_dev, 64 * sizeof(float), 0, cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // namespace pidm
using namespace pidm;

extern "C" int pidm_mechanics_residual_fwd(const float* u, const float* rho, const float* bcs, const float* KE,
                                           float* residual, float* compliance, int B, int nel, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = upload_ke(KE, st)) return e;
    if (compliance) PIDM_CUDA(cudaMemsetAsync(compliance, 0, B * sizeof(float), st));
    const int nn = nel + 1;
    dim3 grid(ceil_div(nn * nn, 128), B);
    mech_node_kernel<0><<<grid, 128, 0, st>>>(u, rho, bcs, residual, compliance, nullptr, nullptr, nullptr, nullptr, nel);
    PIDM_LAUNCH_CHECK("mechanics_residual_fwd");
    return 0;
}

// workspace: float[B * 2 * (nel+1)^2].  grad_residual / grad_compliance may be NULL.  grad_u, grad_rho overwritten.
extern "C" int pidm_mechanics_residual_bwd(const float* u, const float* rho, const float* bcs, const float* KE,
                                           const float* grad_residual, const float* grad_compliance, float* grad_u,
                                           float* grad_rho, float* workspace, int B, int nel, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = upload_ke(KE, st)) return e;
    const int nn = nel + 1;
    dim3 grid(ceil_div(nn * nn, 128), B);
    mech_node_kernel<1><<<grid, 128, 0, st>>>(u, rho, bcs, nullptr, nullptr, grad_residual, grad_compliance, workspace,
                                             grad_u, nel);
    mech_node_kernel<2><<<grid, 128, 0, st>>>(u, rho, bcs, nullptr, nullptr, nullptr, nullptr, workspace, grad_u, nel);
    mech_drho_kernel<<<dim3(ceil_div(nel * nel, 128), B), 128, 0, st>>>(u, workspace, grad_rho, nel);
    PIDM_LAUNCH_CHECK("mechanics_residual_bwd");
    return 0;
}

extern "C" int pidm_bilinear_resize_fwd(const float* x, float* y, int planes, int in, int out, void* stream) {
    long long total = (long long)planes * out * out;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    bilinear_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, planes, in, out);
    PIDM_LAUNCH_CHECK("bilinear_resize_fwd");
    return 0;
}

// dx [planes,in,in] is zeroed here, then accumulated.
extern "C" int pidm_bilinear_resize_bwd(const float* dy, float* dx, int planes, int in, int out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PIDM_CUDA(cudaMemsetAsync(dx, 0, (size_t)planes * in * in * sizeof(float), st));
    long long total = (long long)planes * out * out;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    bilinear_bwd_kernel<<<grid, 256, 0, st>>>(dy, dx, planes, in, out);
    PIDM_LAUNCH_CHECK("bilinear_resize_bwd");
    return 0;
}
