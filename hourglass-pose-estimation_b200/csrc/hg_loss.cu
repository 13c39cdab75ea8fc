// JointsMSE loss (+ gradient) over all stacks in one launch, with the Gaussian target either read
// from memory or regenerated on the fly from per-joint centres (src/loss/mse.py:14-44,
// src/datasets/common.py:197-248 of the reference).
#include "hg_common.cuh"
#include "../../include/hg_api.h"

namespace hg {

struct StackPtrs {
    const float* pred[HG_MAX_STACKS];
    float* grad[HG_MAX_STACKS];
};

// common.py:217-227.  `int()` truncates toward zero; the patch is [mu-r, mu+r] in both axes.
__global__ void joint_centers_kernel(const double* __restrict__ joints, const double* __restrict__ vis,
                                     int* __restrict__ mu, float* __restrict__ weight, int total, int h, int w, int in_w,
                                     int in_h, int radius) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double sx = static_cast<double>(in_w) / static_cast<double>(w);
    const double sy = static_cast<double>(in_h) / static_cast<double>(h);
    const int mx = static_cast<int>(joints[3 * i] / sx + 0.5);
    const int my = static_cast<int>(joints[3 * i + 1] / sy + 0.5);
    float wt = static_cast<float>(vis[3 * i]);
    const int ulx = mx - radius, uly = my - radius, brx = mx + radius + 1, bry = my + radius + 1;
    if (ulx >= w || uly >= h || brx < 0 || bry < 0) wt = 0.f;
    mu[2 * i] = mx;
    mu[2 * i + 1] = my;
    weight[i] = wt;
}

__device__ __forceinline__ float target_at(const int* __restrict__ mu, const float* __restrict__ patch, float wt,
                                           int map, int y, int x, int radius) {
    // common.py:243-246: the patch is written only if v > 0.5
    if (!(wt > 0.5f)) return 0.f;
    const int dx = x - mu[2 * map] + radius, dy = y - mu[2 * map + 1] + radius;
    const int size = 2 * radius + 1;
    if (dx < 0 || dx >= size || dy < 0 || dy >= size) return 0.f;
    return __ldg(patch + dy * size + dx);
}

__global__ void __launch_bounds__(256) gaussian_target_kernel(const int* __restrict__ mu, const float* __restrict__ weight,
                                                              const float* __restrict__ patch, float* __restrict__ target,
                                                              long long total, int h, int w, int radius) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % w);
        const long long t = i / w;
        const int y = static_cast<int>(t % h);
        const int map = static_cast<int>(t / h);
        target[i] = target_at(mu, patch, weight[map], map, y, x, radius);
    }
}

// One thread handles 4 consecutive elements (hw % 4 == 0 path) of every stack: target / weight are read
// once, S predictions are read and S gradients written.  Block-level fp32 reduction, one atomic per CTA.
template <bool kVec>
__global__ void __launch_bounds__(256) jmse_kernel(StackPtrs ptrs, const float* __restrict__ target,
                                                   const float* __restrict__ tw, const int* __restrict__ mu,
                                                   const float* __restrict__ patch, int radius, float* __restrict__ loss_out,
                                                   int stacks, long long total, int h, int w, float loss_scale,
                                                   float grad_scale) {
    constexpr int V = kVec ? 4 : 1;
    const int hw = h * w;
    float acc = 0.f;
    const long long nvec = total / V;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long e0 = i * V;
        const int map = static_cast<int>(e0 / hw);
        const float wt = tw ? __ldg(tw + map) : 1.f;
        float g[V];
        if (target) {
            if (kVec) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(target + e0));
                g[0] = t4.x; g[V > 1 ? 1 : 0] = t4.y; g[V > 2 ? 2 : 0] = t4.z; g[V > 3 ? 3 : 0] = t4.w;
            } else {
                g[0] = __ldg(target + e0);
            }
        } else {
            const int r = static_cast<int>(e0 - static_cast<long long>(map) * hw);
            const int y = r / w, x = r - y * w;
            // mirrors generate_target: patch only when the (possibly forced-to-zero) weight is > 0.5.
            // With use_target_weight=False (tw == NULL) callers must pass an explicit target.
#pragma unroll
            for (int v = 0; v < V; ++v) g[v] = target_at(mu, patch, wt, map, y, x + v, radius);
        }
        for (int s = 0; s < stacks; ++s) {
            float p[V];
            if (kVec) {
                const float4 p4 = __ldg(reinterpret_cast<const float4*>(ptrs.pred[s] + e0));
                p[0] = p4.x; p[V > 1 ? 1 : 0] = p4.y; p[V > 2 ? 2 : 0] = p4.z; p[V > 3 ? 3 : 0] = p4.w;
            } else {
                p[0] = __ldg(ptrs.pred[s] + e0);
            }
            float d[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                d[v] = p[v] * wt - g[v] * wt;             // mse.py:32-33: both operands scaled, then subtracted
                acc = fmaf(d[v], d[v], acc);
            }
            if (ptrs.grad[s]) {
                const float gs = wt * grad_scale;
                if (kVec) {
                    float4 o;
                    o.x = d[0] * gs; o.y = d[V > 1 ? 1 : 0] * gs; o.z = d[V > 2 ? 2 : 0] * gs; o.w = d[V > 3 ? 3 : 0] * gs;
                    *reinterpret_cast<float4*>(ptrs.grad[s] + e0) = o;
                } else {
                    ptrs.grad[s][e0] = d[0] * gs;
                }
            }
        }
    }
    // block reduction
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    __shared__ float s_part[8];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? s_part[threadIdx.x] : 0.f;
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (threadIdx.x == 0 && v != 0.f) atomicAdd(loss_out, v * loss_scale);
    }
}

}  // namespace hg

using namespace hg;

extern "C" int hg_joint_centers(const double* joints, const double* vis, int32_t* mu, float* weight, int32_t b, int32_t j,
                                int32_t h, int32_t w, int32_t in_w, int32_t in_h, int32_t radius, void* stream) {
    if (!joints || !vis || !mu || !weight || b <= 0 || j <= 0 || h <= 0 || w <= 0 || in_w <= 0 || in_h <= 0 || radius < 0) {
        set_last_error("hg_joint_centers: bad arguments");
        return HG_ERR_INVALID;
    }
    const int total = b * j;
    joint_centers_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(joints, vis, mu, weight, total, h,
                                                                                              w, in_w, in_h, radius);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}

extern "C" int hg_gaussian_target(const int32_t* mu, const float* weight, const float* patch, float* target, int32_t b,
                                  int32_t j, int32_t h, int32_t w, int32_t radius, void* stream) {
    if (!mu || !weight || !patch || !target || b <= 0 || j <= 0 || h <= 0 || w <= 0 || radius < 0) {
        set_last_error("hg_gaussian_target: bad arguments");
        return HG_ERR_INVALID;
    }
    const long long total = static_cast<long long>(b) * j * h * w;
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    gaussian_target_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(mu, weight, patch, target,
                                                                                                     total, h, w, radius);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}

extern "C" int hg_jmse_loss(const float* const* preds, float* const* grads, const float* target, const float* target_weight,
                            const int32_t* mu, const float* patch, int32_t radius, float* loss_out, int32_t stacks,
                            int32_t b, int32_t j, int32_t h, int32_t w, float grad_scale, void* stream) {
    if (!preds || !loss_out || stacks <= 0 || stacks > HG_MAX_STACKS || b <= 0 || j <= 0 || h <= 0 || w <= 0) {
        set_last_error("hg_jmse_loss: bad arguments (stacks must be 1..%d)", HG_MAX_STACKS);
        return HG_ERR_INVALID;
    }
    if (!target && (!mu || !patch || !target_weight || radius < 0)) {
        set_last_error("hg_jmse_loss: without an explicit target, mu/patch/target_weight are required");
        return HG_ERR_INVALID;
    }
    StackPtrs ptrs;
    bool vec = (w % 4 == 0) && (!target || (reinterpret_cast<uintptr_t>(target) & 15u) == 0);
    for (int s = 0; s < HG_MAX_STACKS; ++s) {
        ptrs.pred[s] = s < stacks ? preds[s] : nullptr;
        ptrs.grad[s] = (s < stacks && grads) ? grads[s] : nullptr;
        if (s < stacks) {
            if (!ptrs.pred[s]) {
                set_last_error("hg_jmse_loss: preds[%d] is null", s);
                return HG_ERR_INVALID;
            }
            if ((reinterpret_cast<uintptr_t>(ptrs.pred[s]) & 15u) || (reinterpret_cast<uintptr_t>(ptrs.grad[s]) & 15u))
                vec = false;
        }
    }
    const long long total = static_cast<long long>(b) * j * h * w;
    const float denom = static_cast<float>(static_cast<double>(j) * b * h * w);
    const float loss_scale = 0.5f / denom;
    const float gscale = grad_scale / denom;
    const long long items = vec ? total / 4 : total;
    long long blocks = (items + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec)
        jmse_kernel<true><<<static_cast<int>(blocks), 256, 0, st>>>(ptrs, target, target_weight, mu, patch, radius, loss_out,
                                                                    stacks, total, h, w, loss_scale, gscale);
    else
        jmse_kernel<false><<<static_cast<int>(blocks), 256, 0, st>>>(ptrs, target, target_weight, mu, patch, radius, loss_out,
                                                                     stacks, total, h, w, loss_scale, gscale);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}
