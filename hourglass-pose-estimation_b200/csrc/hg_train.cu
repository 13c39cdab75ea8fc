// Bandwidth-bound kernels of the training step (src/runner/trainer.py:82-99): train-mode BatchNorm
// forward/backward (batch statistics), max-pool / nearest-upsample backward, gradient accumulation,
// weight packing (fp32 master -> bf16 GEMM layouts), RMSprop, and a tiny fp32 GEMM for the parameter-
// space chain rule of the merged remap convolution.
//
// All activation tensors are NHWC bf16; one thread owns one 16-byte chunk (8 channels), so per-channel
// constants live in registers/shared memory and every global access is 128-bit and coalesced.
// Per-channel reductions: a thread keeps fp32 partial sums for ITS 8 channels over a strided set of
// pixels, the block combines them through shared memory, one atomicAdd per channel per block.
#include <cstdlib>
#include "hg_common.cuh"
#include "../../include/hg_api.h"

namespace hg {
namespace tr {

constexpr int kThreads = 256;
constexpr int kMaxC = 256;
constexpr int kUnroll = 4;          // independent 128-bit loads per thread in the streaming loops (8 measured slower:
                                    // 27.9 vs 26.5 ms/step -- registers cost occupancy)

// CTAs per SM of the two per-channel reduction kernels.  Every CTA ends with one atomicAdd per channel, so the grid size is
// also the contention on each accumulator: measured on B200 (training step, batch 32) 26.5 / 24.8 / 24.8 / 25.3 / 26.1 /
// 26.6 ms for 1 / 2 / 3 / 4 / 6 / 8 CTAs per SM.
constexpr int kReducePerSm = 2;
static inline int grid_for(long long items, int per_sm = 8) {
    const long long cap = static_cast<long long>(num_sms()) * per_sm;
    return static_cast<int>(items < cap ? (items > 0 ? items : 1) : cap);
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = bf16_lo_to_f32(v.x); f[1] = bf16_hi_to_f32(v.x);
    f[2] = bf16_lo_to_f32(v.y); f[3] = bf16_hi_to_f32(v.y);
    f[4] = bf16_lo_to_f32(v.z); f[5] = bf16_hi_to_f32(v.z);
    f[6] = bf16_lo_to_f32(v.w); f[7] = bf16_hi_to_f32(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 r;
    r.x = pack_bf16x2(f[0], f[1]);
    r.y = pack_bf16x2(f[2], f[3]);
    r.z = pack_bf16x2(f[4], f[5]);
    r.w = pack_bf16x2(f[6], f[7]);
    return r;
}

// Per-channel reductions run on a 2-D grid: blockIdx.y picks a GROUP of kGroupC8 16-byte chunks (64 channels) and
// blockIdx.x a SLAB of pixels (pixel lanes x slabs, strided).  A block therefore ends with 64 (x NV) channel sums
// instead of C, which quarters the number of accumulator updates for C = 256 at the same number of blocks.
constexpr int kGroupC8 = 8;
constexpr int kGroupC = kGroupC8 * 8;
constexpr int kScratchHeader = 16;     // floats reserved at the head of a reduction scratch buffer (arrival counters)

__device__ __forceinline__ float ld_cg_f32(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// Block-level combine of per-thread partial sums acc[NV][8] (thread = (pixel lane, 8-channel chunk of this block's
// group)) and the grid-level sum:  out_v[c] += sum.  out pointers may be null (skipped); channels >= c_valid are skipped.
//   scratch == nullptr: one atomicAdd per channel per block (order of the additions = order of arrival).
//   scratch != nullptr: DETERMINISTIC.  Every block stores its partial sums in its own slot of `scratch`, takes a ticket,
//     and the block that arrives last adds all slots in slab order -- the same fp32 additions in the same order on every
//     run, whatever the scheduling -- and resets the ticket counter (scratch must be zero before its first use only).
template <int NV>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[NV][8], float* const (&out)[NV], int c_valid, float* scratch) {
    __shared__ float sm[NV][kThreads][9];      // +1 padding: conflict-free column reads
    __shared__ int s_last;
    const int tid = threadIdx.x;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int e = 0; e < 8; ++e) sm[v][tid][e] = acc[v][e];
    __syncthreads();
    constexpr int lanes = kThreads / kGroupC8;
    const int slabs = gridDim.x;
    const int c0 = blockIdx.y * kGroupC;
    const bool direct = scratch == nullptr || slabs == 1;
    float* part = direct ? nullptr
                         : scratch + kScratchHeader + (static_cast<size_t>(blockIdx.y) * slabs + blockIdx.x) * (NV * kGroupC);
    for (int idx = tid; idx < NV * kGroupC; idx += kThreads) {
        const int v = idx / kGroupC, c = idx - v * kGroupC;
        const int chunk = c >> 3, e = c & 7;
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += sm[v][l * kGroupC8 + chunk][e];
        if (!direct) {
            __stcg(part + idx, s);
        } else if (out[v] != nullptr && c0 + c < c_valid) {
            if (scratch == nullptr) atomicAdd(out[v] + c0 + c, s);
            else out[v][c0 + c] += s;          // a single slab: this block is the only writer of these channels
        }
    }
    if (direct) return;
    __threadfence();
    __syncthreads();
    unsigned int* counter = reinterpret_cast<unsigned int*>(scratch) + blockIdx.y;
    if (tid == 0) s_last = atomicAdd(counter, 1u) == static_cast<unsigned int>(slabs - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* all = scratch + kScratchHeader + static_cast<size_t>(blockIdx.y) * slabs * (NV * kGroupC);
    // two threads per output: even / odd halves of the slab list, each in ascending order, combined in a fixed order
    const int half = tid / (NV * kGroupC), idx = tid - half * (NV * kGroupC);
    float s = 0.f;
    if (half < 2) {
        const int b0 = half == 0 ? 0 : (slabs + 1) / 2, b1 = half == 0 ? (slabs + 1) / 2 : slabs;
        int b = b0;
        for (; b + 4 <= b1; b += 4) {
            const float v0 = ld_cg_f32(all + static_cast<size_t>(b) * (NV * kGroupC) + idx);
            const float v1 = ld_cg_f32(all + static_cast<size_t>(b + 1) * (NV * kGroupC) + idx);
            const float v2 = ld_cg_f32(all + static_cast<size_t>(b + 2) * (NV * kGroupC) + idx);
            const float v3 = ld_cg_f32(all + static_cast<size_t>(b + 3) * (NV * kGroupC) + idx);
            s = (((s + v0) + v1) + v2) + v3;
        }
        for (; b < b1; ++b) s += ld_cg_f32(all + static_cast<size_t>(b) * (NV * kGroupC) + idx);
    }
    __syncthreads();
    float* comb = &sm[0][0][0];
    if (half == 1) comb[idx] = s;
    __syncthreads();
    if (half == 0) {
        s += comb[idx];
        const int v = idx / kGroupC, c = idx - v * kGroupC;
        if (out[v] != nullptr && c0 + c < c_valid) out[v][c0 + c] += s;
    }
    if (tid == 0) *counter = 0u;
}
static_assert(kThreads >= 2 * 2 * kGroupC, "final stage: two threads per output");

// ---------------------------------------------------------------- per-channel sum / sum of squares
// shift != 0: the sums are taken about k[c] = x[pixel 0][c] (sum (x-k), sum (x-k)^2): the variance the BatchNorm derives
// from them, E[(x-k)^2] - E[x-k]^2, does not cancel catastrophically when |mean| >> std.
__global__ void __launch_bounds__(kThreads) colstats_kernel(const uint4* __restrict__ x, float* sum, float* sumsq,
                                                             long long pixels, int c8, int c_valid, int shift, float* scratch) {
    pdl_launch_dependents();
    pdl_wait();
    const int chunk = blockIdx.y * kGroupC8 + threadIdx.x % kGroupC8, prow = threadIdx.x / kGroupC8;
    constexpr int lanes = kThreads / kGroupC8;
    float acc[2][8], k[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[0][e] = acc[1][e] = k[e] = 0.f;
    if (shift) unpack8(ldg_nc_v4(x + chunk), k);
    // kUnroll independent 16-byte loads in flight per thread: a single load per iteration leaves ~2.4 MB in flight on
    // the whole GPU, well short of the ~6.5 MB that bandwidth x latency asks for
    const long long stride = static_cast<long long>(gridDim.x) * lanes;
    long long pix = static_cast<long long>(blockIdx.x) * lanes + prow;
    for (; pix + (kUnroll - 1) * stride < pixels; pix += kUnroll * stride) {
        uint4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) v[u] = ldg_nc_v4(x + (pix + u * stride) * c8 + chunk);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            float f[8];
            unpack8(v[u], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = f[e] - k[e];
                acc[0][e] += d;
                acc[1][e] = fmaf(d, d, acc[1][e]);
            }
        }
    }
    for (; pix < pixels; pix += stride) {
        float f[8];
        unpack8(ldg_nc_v4(x + pix * c8 + chunk), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float d = f[e] - k[e];
            acc[0][e] += d;
            acc[1][e] = fmaf(d, d, acc[1][e]);
        }
    }
    float* const outs[2] = {sum, sumsq};
    block_channel_reduce<2>(acc, outs, c_valid, scratch);
}

// ---------------------------------------------------------------- train-mode BN + ReLU (forward)
// scale/shift from the batch sums; block 0 also records them and updates the running statistics
// (torch.nn.BatchNorm2d: biased variance to normalise, unbiased for running_var, momentum 0.1).
struct BnFwdParams {
    const uint4* x;
    uint4* out;
    const float* sums;      // [2C]: sum, sum of squares
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    long long* num_batches_tracked;
    float* saved;           // [4C]: mean, invstd, scale, shift
    int n, h, w, c8;
    int halo;               // out is halo-padded [zero row][n][h+1][w+1][c]
    int relu;
    int shifted;            // sums are taken about k[c] = x[pixel 0][c] (hg_colstats_nhwc with shift)
    float eps, momentum;
};

__global__ void __launch_bounds__(kThreads) bn_train_fwd_kernel(const BnFwdParams p) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_scale[kMaxC], s_shift[kMaxC];
    const int C = p.c8 * 8;
    const long long pixels = static_cast<long long>(p.n) * p.h * p.w;
    const float inv_n = 1.f / static_cast<float>(pixels);
    for (int c = threadIdx.x; c < C; c += kThreads) {
        const float m1 = p.sums[c] * inv_n;
        const float var = fmaxf(p.sums[C + c] * inv_n - m1 * m1, 0.f);
        const float mean = p.shifted ? m1 + __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.x)[c]) : m1;
        const float invstd = rsqrtf(var + p.eps);
        const float sc = p.gamma[c] * invstd;
        const float sh = p.beta[c] - mean * sc;
        s_scale[c] = sc;
        s_shift[c] = sh;
        if (blockIdx.x == 0) {
            p.saved[c] = mean;
            p.saved[C + c] = invstd;
            p.saved[2 * C + c] = sc;
            p.saved[3 * C + c] = sh;
            if (p.running_mean != nullptr) {
                const float unbiased = pixels > 1 ? var * static_cast<float>(pixels) / static_cast<float>(pixels - 1) : var;
                p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * mean;
                p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * unbiased;
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.num_batches_tracked != nullptr) *p.num_batches_tracked += 1;
    __syncthreads();
    const long long total = pixels * p.c8;
    const int P = p.w + 1;
    auto apply = [&](long long i, const uint4& v) {
        const int chunk = static_cast<int>(i % p.c8);
        const long long pix = i / p.c8;
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            f[e] = fmaf(f[e], s_scale[chunk * 8 + e], s_shift[chunk * 8 + e]);
            if (p.relu) f[e] = fmaxf(f[e], 0.f);
        }
        long long o = i;
        if (p.halo) {
            const int xw = static_cast<int>(pix % p.w);
            const long long t = pix / p.w;
            const int y = static_cast<int>(t % p.h);
            const long long b = t / p.h;
            o = (P + (b * (p.h + 1) + y) * P + xw) * p.c8 + chunk;
        }
        p.out[o] = pack8(f);
    };
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    for (; i + (kUnroll - 1) * stride < total; i += kUnroll * stride) {     // kUnroll loads in flight per thread
        uint4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) v[u] = ldg_nc_v4(p.x + i + u * stride);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) apply(i + u * stride, v[u]);
    }
    for (; i < total; i += stride) apply(i, ldg_nc_v4(p.x + i));
}

// ---------------------------------------------------------------- BN + ReLU backward, pass 1: per-channel sums
// dY = dz * [x*scale + shift > 0];  s1 = sum dY;  s2 = sum dY * xhat,  xhat = (x - mean) * invstd
__global__ void __launch_bounds__(kThreads) bn_bwd_reduce_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ x,
                                                                  const float* __restrict__ saved, float* sums,
                                                                  long long pixels, int c8, int relu, float* scratch) {
    pdl_launch_dependents();
    pdl_wait();
    const int C = c8 * 8;
    const int chunk = blockIdx.y * kGroupC8 + threadIdx.x % kGroupC8, prow = threadIdx.x / kGroupC8;
    constexpr int lanes = kThreads / kGroupC8;
    float mean[8], invstd[8], sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = chunk * 8 + e;
        mean[e] = saved[c];
        invstd[e] = saved[C + c];
        sc[e] = saved[2 * C + c];
        sh[e] = saved[3 * C + c];
    }
    float acc[2][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[0][e] = acc[1][e] = 0.f;
    const long long stride = static_cast<long long>(gridDim.x) * lanes;
    long long pix = static_cast<long long>(blockIdx.x) * lanes + prow;
    constexpr int kU = 2;                  // 2 x (dz, x) = four 128-bit loads in flight
    for (; pix + (kU - 1) * stride < pixels; pix += kU * stride) {
        uint4 vg[kU], vx[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            vg[u] = ldg_nc_v4(dz + (pix + u * stride) * c8 + chunk);
            vx[u] = ldg_nc_v4(x + (pix + u * stride) * c8 + chunk);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            float g[8], xv[8];
            unpack8(vg[u], g);
            unpack8(vx[u], xv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float dy = (!relu || fmaf(xv[e], sc[e], sh[e]) > 0.f) ? g[e] : 0.f;
                acc[0][e] += dy;
                acc[1][e] = fmaf(dy, (xv[e] - mean[e]) * invstd[e], acc[1][e]);
            }
        }
    }
    for (; pix < pixels; pix += stride) {
        float g[8], xv[8];
        unpack8(ldg_nc_v4(dz + pix * c8 + chunk), g);
        unpack8(ldg_nc_v4(x + pix * c8 + chunk), xv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float dy = (!relu || fmaf(xv[e], sc[e], sh[e]) > 0.f) ? g[e] : 0.f;
            acc[0][e] += dy;
            acc[1][e] = fmaf(dy, (xv[e] - mean[e]) * invstd[e], acc[1][e]);
        }
    }
    float* const outs[2] = {sums, sums + C};
    block_channel_reduce<2>(acc, outs, C, scratch);
}

// ---------------------------------------------------------------- BN + ReLU backward, pass 2
// dx = scale * (dY - s1/N - xhat * s2/N) (+ add1) (+ add2);  block 0 writes dgamma = s2, dbeta = s1.
struct BnBwdParams {
    const uint4* dz;
    const uint4* x;
    const float* saved;
    const float* sums;      // [2C]: s1, s2
    const uint4* add1;
    const uint4* add2;
    uint4* out;
    float* dgamma;
    float* dbeta;
    int n, h, w, c8;
    int halo;
    int relu;
};

__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(const BnBwdParams p) {
    pdl_launch_dependents();
    pdl_wait();
    // dx = a*dY + b*x + c  with  a = scale, b = -scale*invstd*k2, c = scale*(mean*invstd*k2 - k1)
    __shared__ float s_a[kMaxC], s_b[kMaxC], s_c[kMaxC], s_sc[kMaxC], s_sh[kMaxC];
    const int C = p.c8 * 8;
    const long long pixels = static_cast<long long>(p.n) * p.h * p.w;
    const float inv_n = 1.f / static_cast<float>(pixels);
    for (int c = threadIdx.x; c < C; c += kThreads) {
        const float mean = p.saved[c], invstd = p.saved[C + c], sc = p.saved[2 * C + c];
        const float s1 = p.sums[c], s2 = p.sums[C + c];
        const float k1 = s1 * inv_n, k2 = s2 * inv_n;
        s_a[c] = sc;
        s_b[c] = -sc * invstd * k2;
        s_c[c] = sc * (mean * invstd * k2 - k1);
        s_sc[c] = sc;
        s_sh[c] = p.saved[3 * C + c];
        if (blockIdx.x == 0 && p.dgamma != nullptr) {
            p.dgamma[c] = s2;
            p.dbeta[c] = s1;
        }
    }
    __syncthreads();
    const long long total = pixels * p.c8;
    const int P = p.w + 1;
    auto apply = [&](long long i, const uint4& vg, const uint4& vx) {
        const int chunk = static_cast<int>(i % p.c8);
        float g[8], xv[8], r[8];
        unpack8(vg, g);
        unpack8(vx, xv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = chunk * 8 + e;
            const float dy = (!p.relu || fmaf(xv[e], s_sc[c], s_sh[c]) > 0.f) ? g[e] : 0.f;
            r[e] = fmaf(s_a[c], dy, fmaf(s_b[c], xv[e], s_c[c]));
        }
        if (p.add1 != nullptr) {
            float a[8];
            unpack8(ldg_v4(p.add1 + i), a);
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] += a[e];
        }
        if (p.add2 != nullptr) {
            float a[8];
            unpack8(ldg_v4(p.add2 + i), a);
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] += a[e];
        }
        long long o = i;
        if (p.halo) {
            const long long pix = i / p.c8;
            const int xw = static_cast<int>(pix % p.w);
            const long long t = pix / p.w;
            const int y = static_cast<int>(t % p.h);
            const long long b = t / p.h;
            o = (P + (b * (p.h + 1) + y) * P + xw) * p.c8 + chunk;
        }
        p.out[o] = pack8(r);
    };
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    constexpr int kU = 2;                 // 2 x (dz, x) loads in flight before the dependent work (out may alias add1/add2:
    for (; i + (kU - 1) * stride < total; i += kU * stride) {      //  those are read inside apply(), element-wise before the store)
        uint4 vg[kU], vx[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            vg[u] = ldg_nc_v4(p.dz + i + u * stride);
            vx[u] = ldg_nc_v4(p.x + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) apply(i + u * stride, vg[u], vx[u]);
    }
    for (; i < total; i += stride) apply(i, ldg_nc_v4(p.dz + i), ldg_nc_v4(p.x + i));
}

// ---------------------------------------------------------------- max-pool 2x2 backward
// dx[2y+i][2x+j] (+)= dpool[y][x] at the FIRST maximum of the window in scan order (torch's rule).
__global__ void __launch_bounds__(kThreads) maxpool_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dpool,
                                                                uint4* dx, int n, int h, int w, int c8, int accumulate) {
    pdl_launch_dependents();
    pdl_wait();
    const int oh = h >> 1, ow = w >> 1;
    const long long total = static_cast<long long>(n) * oh * ow * c8;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(i % c8);
        long long pix = i / c8;
        const int ox = static_cast<int>(pix % ow);
        pix /= ow;
        const int oy = static_cast<int>(pix % oh);
        const long long b = pix / oh;
        const long long base = ((b * h + 2 * oy) * w + 2 * ox) * c8 + ch;
        const long long offs[4] = {base, base + c8, base + static_cast<long long>(w) * c8,
                                   base + static_cast<long long>(w) * c8 + c8};
        float v[4][8], g[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) unpack8(ldg_nc_v4(x + offs[k]), v[k]);
        unpack8(ldg_nc_v4(dpool + i), g);
        float r[4][8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int best = 0;
            float m = v[0][e];
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (v[k][e] > m) {
                    m = v[k][e];
                    best = k;
                }
#pragma unroll
            for (int k = 0; k < 4; ++k) r[k][e] = (k == best) ? g[e] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (accumulate) {
                float old[8];
                unpack8(ldg_v4(dx + offs[k]), old);
#pragma unroll
                for (int e = 0; e < 8; ++e) r[k][e] += old[e];
            }
            dx[offs[k]] = pack8(r[k]);
        }
    }
}

// ---------------------------------------------------------------- nearest-upsample x2 backward = 2x2 sum
__global__ void __launch_bounds__(kThreads) sumpool_kernel(const uint4* __restrict__ dy, uint4* dlow, int n, int h, int w,
                                                            int c8, int accumulate) {
    pdl_launch_dependents();
    pdl_wait();
    const int oh = h >> 1, ow = w >> 1;
    const long long total = static_cast<long long>(n) * oh * ow * c8;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(i % c8);
        long long pix = i / c8;
        const int ox = static_cast<int>(pix % ow);
        pix /= ow;
        const int oy = static_cast<int>(pix % oh);
        const long long b = pix / oh;
        const long long base = ((b * h + 2 * oy) * w + 2 * ox) * c8 + ch;
        float a[8], t[8];
        unpack8(ldg_nc_v4(dy + base), a);
        unpack8(ldg_nc_v4(dy + base + c8), t);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += t[e];
        unpack8(ldg_nc_v4(dy + base + static_cast<long long>(w) * c8), t);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += t[e];
        unpack8(ldg_nc_v4(dy + base + static_cast<long long>(w) * c8 + c8), t);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += t[e];
        if (accumulate) {
            unpack8(ldg_v4(dlow + i), t);
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] += t[e];
        }
        dlow[i] = pack8(a);
    }
}

// ---------------------------------------------------------------- dst += src
__global__ void __launch_bounds__(kThreads) add_inplace_kernel(uint4* dst, const uint4* __restrict__ src, long long count) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float a[8], b[8];
        unpack8(ldg_v4(dst + i), a);
        unpack8(ldg_nc_v4(src + i), b);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += b[e];
        dst[i] = pack8(a);
    }
}

// ---------------------------------------------------------------- fp32 NCHW -> bf16 NHWC with channel padding
// (heat-map gradients [n][j][h][w] -> the score convolution's dgrad / wgrad operand [n][h][w][c_pad]).  A block moves a
// tile of 64 pixels through shared memory: reads run along the pixels of one channel plane (coalesced), writes along the
// channels of one pixel (coalesced 16-byte chunks) -- the naive one-thread-per-element form read 4 bytes every 16 KiB.
constexpr int kTpPix = 64;
__global__ void __launch_bounds__(kThreads) nchw_to_nhwc_pad_kernel(const float* __restrict__ in, __nv_bfloat16* out, int n,
                                                                     int c, int c_pad, int hw) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(16) __nv_bfloat16 tile[kTpPix][64 + 8];        // c_pad <= 64; +8: row pitch 144 B spreads the banks
    const int tiles_per_img = (hw + kTpPix - 1) / kTpPix;
    const long long tiles = static_cast<long long>(n) * tiles_per_img;
    const int px = threadIdx.x % kTpPix, cq = threadIdx.x / kTpPix;     // 4 channel lanes x 64 pixels
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long long b = t / tiles_per_img;
        const int p0 = static_cast<int>(t - b * tiles_per_img) * kTpPix;
        const bool in_img = p0 + px < hw;
        for (int ch = cq; ch < c_pad; ch += kThreads / kTpPix)
            tile[px][ch] = __float2bfloat16_rn((ch < c && in_img) ? in[(b * c + ch) * hw + p0 + px] : 0.f);
        __syncthreads();
        const int chunks = c_pad / 8;                                   // 16-byte chunks per pixel
        for (int i = threadIdx.x; i < kTpPix * chunks; i += kThreads) {
            const int p = i / chunks, k = i - p * chunks;
            if (p0 + p < hw)
                *reinterpret_cast<uint4*>(out + (b * hw + p0 + p) * c_pad + k * 8) = *reinterpret_cast<const uint4*>(&tile[p][k * 8]);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- weight packing (table-driven, one launch)
constexpr int kPackBlocksPerEntry = 32;      // 289 -> 259 us for the 8-stack network (the transposed dgrad writes dominate)
__global__ void __launch_bounds__(kThreads) pack_weights_kernel(const hg_pack_entry* __restrict__ table, int n_entries,
                                                                 int which) {
    // NO early pdl_launch_dependents() here: the GEMM kernels fetch their weights BEFORE griddepcontrol.wait
    // (constants on the inference path), and a chain of small grids can run that preamble several launches ahead.
    // Without the early trigger nothing launched after this kernel starts until every packed weight is visible.
    pdl_wait();
    const int ei = blockIdx.x / kPackBlocksPerEntry;
    if (ei >= n_entries) return;
    const hg_pack_entry e = table[ei];
    const int sub = blockIdx.x - ei * kPackBlocksPerEntry;
    const int row_len = e.taps * e.ci;
    const long long total = static_cast<long long>(e.co) * row_len;
    // which: bit 0 = the forward layouts (dst_fwd, dst_f32), bit 1 = the transposed dgrad layout -- two launches let the
    // slow transposed writes run beside the forward pass, which only waits for the first
    __nv_bfloat16* fwd = (which & 1) ? static_cast<__nv_bfloat16*>(e.dst_fwd) : nullptr;
    float* f32 = (which & 1) ? e.dst_f32 : nullptr;
    __nv_bfloat16* dg = (which & 2) ? static_cast<__nv_bfloat16*>(e.dst_dgrad) : nullptr;
    if (fwd == nullptr && f32 == nullptr && dg == nullptr) return;
    for (long long idx = static_cast<long long>(sub) * kThreads + threadIdx.x; idx < total;
         idx += static_cast<long long>(kPackBlocksPerEntry) * kThreads) {
        const int o = static_cast<int>(idx / row_len);
        const int r = static_cast<int>(idx - static_cast<long long>(o) * row_len);
        float v = e.src[idx];
        if (e.src2 != nullptr) v += e.src2[idx];
        if (f32 != nullptr) f32[idx] = v;
        if (fwd != nullptr) fwd[static_cast<long long>(o) * e.fwd_ld + e.fwd_col0 + r] = __float2bfloat16_rn(v);
        if (dg != nullptr) {
            const int tap = r / e.ci, i = r - tap * e.ci;
            dg[static_cast<long long>(i) * e.dgrad_ld + (e.taps - 1 - tap) * e.co + o] = __float2bfloat16_rn(v);
        }
    }
}

// ---------------------------------------------------------------- RMSprop (torch defaults: momentum 0, not centered)
__global__ void __launch_bounds__(kThreads) rmsprop_kernel(float4* p, const float4* __restrict__ g, float4* v, long long n4,
                                                            float lr, float alpha, float eps, float grad_scale) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float4 pp = p[i], vv = v[i];
        float4 gg = g[i];
        gg.x *= grad_scale; gg.y *= grad_scale; gg.z *= grad_scale; gg.w *= grad_scale;
        vv.x = alpha * vv.x + (1.f - alpha) * gg.x * gg.x;
        vv.y = alpha * vv.y + (1.f - alpha) * gg.y * gg.y;
        vv.z = alpha * vv.z + (1.f - alpha) * gg.z * gg.z;
        vv.w = alpha * vv.w + (1.f - alpha) * gg.w * gg.w;
        pp.x -= lr * gg.x / (sqrtf(vv.x) + eps);
        pp.y -= lr * gg.y / (sqrtf(vv.y) + eps);
        pp.z -= lr * gg.z / (sqrtf(vv.z) + eps);
        pp.w -= lr * gg.w / (sqrtf(vv.w) + eps);
        p[i] = pp;
        v[i] = vv;
    }
}

// ---------------------------------------------------------------- tiny strided fp32 GEMM (parameter space only)
// C[i][j] = beta*C[i][j] + (D ? D[i][j] : 0) + sum_k A[i*sai + k*sak] * B[k*sbk + j*sbj];  one warp per output,
// the k range split over the lanes and combined by shuffles.
__global__ void __launch_bounds__(kThreads) small_gemm_kernel(float* C, const float* __restrict__ A, const float* __restrict__ B,
                                                               const float* __restrict__ D, int m, int n, int k, int sai,
                                                               int sak, int sbk, int sbj, int sci, int scj, float beta) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = static_cast<long long>(m) * n;
    const int lane = threadIdx.x & 31;
    const long long warps = static_cast<long long>(gridDim.x) * (kThreads / 32);
    for (long long idx = static_cast<long long>(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5); idx < total; idx += warps) {
        const int i = static_cast<int>(idx / n), j = static_cast<int>(idx - static_cast<long long>(i) * n);
        float s = 0.f;
        for (int kk = lane; kk < k; kk += 32)
            s = fmaf(A[static_cast<long long>(i) * sai + static_cast<long long>(kk) * sak],
                     B[static_cast<long long>(kk) * sbk + static_cast<long long>(j) * sbj], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            float* c = C + static_cast<long long>(i) * sci + static_cast<long long>(j) * scj;
            float r = s;
            if (beta != 0.f) r += beta * *c;
            if (D != nullptr) r += D[static_cast<long long>(i) * sci + static_cast<long long>(j) * scj];
            *c = r;
        }
    }
}

}  // namespace tr
}  // namespace hg

using namespace hg;
using namespace hg::tr;

static bool bn_channels_ok(int c) { return c == 64 || c == 128 || c == 256; }

// slabs of a per-channel reduction over `pixels` pixels of a c-channel tensor (grid.x; grid.y = c / 64 groups)
static int reduce_slabs(long long pixels, int c) {
    constexpr int lanes = kThreads / kGroupC8;
    const int groups = c / kGroupC;
    long long cap = static_cast<long long>(num_sms()) * kReducePerSm / groups;
    if (cap > 160) cap = 160;            // the deterministic final stage reads one slot per slab
    if (cap < 1) cap = 1;
    const long long need = (pixels + lanes - 1) / lanes;
    return static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
}

extern "C" int64_t hg_colreduce_scratch_bytes(int64_t pixels, int32_t c) {
    if (pixels <= 0 || c <= 0 || c % kGroupC != 0) return 0;
    return static_cast<int64_t>(sizeof(float)) * (kScratchHeader + static_cast<int64_t>(c / kGroupC) * reduce_slabs(pixels, c) * 2 * kGroupC);
}

extern "C" int hg_colstats_nhwc(const void* x, float* sum, float* sumsq, int64_t pixels, int32_t c, int32_t c_valid,
                                int32_t shift, float* scratch, void* stream) {
    if (!x || pixels <= 0 || !bn_channels_ok(c) || c_valid <= 0 || c_valid > c || !aligned16(x)) {
        set_last_error("hg_colstats_nhwc: need c in {64,128,256}, 16-byte aligned input");
        return HG_ERR_INVALID;
    }
    HG_CUDA_OK(launch_kernel(colstats_kernel, dim3(reduce_slabs(pixels, c), c / kGroupC), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), static_cast<const uint4*>(x), sum, sumsq,
                             static_cast<long long>(pixels), c / 8, c_valid, shift, scratch));
    return HG_OK;
}

extern "C" int hg_bn_train_fwd(const void* x, const float* sums, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, int64_t* num_batches_tracked, float* saved, void* out, int32_t n,
                               int32_t h, int32_t w, int32_t c, int32_t out_halo, int32_t relu, int32_t shifted, float eps,
                               float momentum, void* stream) {
    if (!x || !sums || !gamma || !beta || !saved || !out || n <= 0 || h <= 0 || w <= 0 || !bn_channels_ok(c) ||
        !aligned16(x) || !aligned16(out)) {
        set_last_error("hg_bn_train_fwd: need c in {64,128,256}, 16-byte aligned tensors");
        return HG_ERR_INVALID;
    }
    BnFwdParams p;
    p.x = static_cast<const uint4*>(x);
    p.out = static_cast<uint4*>(out);
    p.sums = sums;
    p.gamma = gamma;
    p.beta = beta;
    p.running_mean = running_mean;
    p.running_var = running_var;
    p.num_batches_tracked = reinterpret_cast<long long*>(num_batches_tracked);
    p.saved = saved;
    p.n = n; p.h = h; p.w = w; p.c8 = c / 8;
    p.halo = out_halo;
    p.relu = relu;
    p.shifted = shifted;
    p.eps = eps;
    p.momentum = momentum;
    const long long items = static_cast<long long>(n) * h * w * (c / 8);
    HG_CUDA_OK(launch_kernel(bn_train_fwd_kernel, dim3(grid_for((items + kThreads - 1) / kThreads)), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), p));
    return HG_OK;
}

extern "C" int hg_bn_bwd_reduce(const void* dz, const void* x, const float* saved, float* sums, int64_t pixels, int32_t c,
                                int32_t relu, float* scratch, void* stream) {
    if (!dz || !x || !saved || !sums || pixels <= 0 || !bn_channels_ok(c) || !aligned16(dz) || !aligned16(x)) {
        set_last_error("hg_bn_bwd_reduce: need c in {64,128,256}, 16-byte aligned tensors");
        return HG_ERR_INVALID;
    }
    HG_CUDA_OK(launch_kernel(bn_bwd_reduce_kernel, dim3(reduce_slabs(pixels, c), c / kGroupC), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), static_cast<const uint4*>(dz), static_cast<const uint4*>(x),
                             saved, sums, static_cast<long long>(pixels), c / 8, relu, scratch));
    return HG_OK;
}

extern "C" int hg_bn_bwd_apply(const void* dz, const void* x, const float* saved, const float* sums, const void* add1,
                               const void* add2, void* out, float* dgamma, float* dbeta, int32_t n, int32_t h, int32_t w,
                               int32_t c, int32_t out_halo, int32_t relu, void* stream) {
    if (!dz || !x || !saved || !sums || !out || n <= 0 || h <= 0 || w <= 0 || !bn_channels_ok(c) || !aligned16(dz) ||
        !aligned16(x) || !aligned16(out) || !aligned16(add1) || !aligned16(add2) || (dgamma == nullptr) != (dbeta == nullptr)) {
        set_last_error("hg_bn_bwd_apply: need c in {64,128,256}, 16-byte aligned tensors");
        return HG_ERR_INVALID;
    }
    BnBwdParams p;
    p.dz = static_cast<const uint4*>(dz);
    p.x = static_cast<const uint4*>(x);
    p.saved = saved;
    p.sums = sums;
    p.add1 = static_cast<const uint4*>(add1);
    p.add2 = static_cast<const uint4*>(add2);
    p.out = static_cast<uint4*>(out);
    p.dgamma = dgamma;
    p.dbeta = dbeta;
    p.n = n; p.h = h; p.w = w; p.c8 = c / 8;
    p.halo = out_halo;
    p.relu = relu;
    const long long items = static_cast<long long>(n) * h * w * (c / 8);
    HG_CUDA_OK(launch_kernel(bn_bwd_apply_kernel, dim3(grid_for((items + kThreads - 1) / kThreads)), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), p));
    return HG_OK;
}

extern "C" int hg_maxpool2x2_bwd_nhwc(const void* x, const void* dpool, void* dx, int32_t n, int32_t h, int32_t w, int32_t c,
                                      int32_t accumulate, void* stream) {
    if (!x || !dpool || !dx || n <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || c <= 0 || (c & 7) || !aligned16(x) ||
        !aligned16(dpool) || !aligned16(dx)) {
        set_last_error("hg_maxpool2x2_bwd_nhwc: need even h,w, c %% 8 == 0, 16-byte aligned pointers");
        return HG_ERR_INVALID;
    }
    const long long items = static_cast<long long>(n) * (h / 2) * (w / 2) * (c / 8);
    HG_CUDA_OK(launch_kernel(maxpool_bwd_kernel, dim3(grid_for((items + kThreads - 1) / kThreads)), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), static_cast<const uint4*>(x), static_cast<const uint4*>(dpool),
                             static_cast<uint4*>(dx), n, h, w, c / 8, accumulate));
    return HG_OK;
}

extern "C" int hg_sumpool2x2_nhwc(const void* dy, void* dlow, int32_t n, int32_t h, int32_t w, int32_t c, int32_t accumulate,
                                  void* stream) {
    if (!dy || !dlow || n <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || c <= 0 || (c & 7) || !aligned16(dy) ||
        !aligned16(dlow)) {
        set_last_error("hg_sumpool2x2_nhwc: need even h,w, c %% 8 == 0, 16-byte aligned pointers");
        return HG_ERR_INVALID;
    }
    const long long items = static_cast<long long>(n) * (h / 2) * (w / 2) * (c / 8);
    HG_CUDA_OK(launch_kernel(sumpool_kernel, dim3(grid_for((items + kThreads - 1) / kThreads)), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), static_cast<const uint4*>(dy), static_cast<uint4*>(dlow), n, h,
                             w, c / 8, accumulate));
    return HG_OK;
}

extern "C" int hg_add_inplace_bf16(void* dst, const void* src, int64_t count, void* stream) {
    if (!dst || !src || count <= 0 || (count & 7) || !aligned16(dst) || !aligned16(src)) {
        set_last_error("hg_add_inplace_bf16: need count %% 8 == 0 and 16-byte aligned pointers");
        return HG_ERR_INVALID;
    }
    const long long items = count / 8;
    HG_CUDA_OK(launch_kernel(add_inplace_kernel, dim3(grid_for((items + kThreads - 1) / kThreads)), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), static_cast<uint4*>(dst), static_cast<const uint4*>(src),
                             items));
    return HG_OK;
}

extern "C" int hg_nchw_f32_to_nhwc_bf16_pad(const float* in, void* out, int32_t n, int32_t c, int32_t c_pad, int32_t h,
                                            int32_t w, void* stream) {
    if (!in || !out || n <= 0 || c <= 0 || c_pad < c || c_pad > 64 || c_pad % 8 != 0 || h <= 0 || w <= 0 || !aligned16(out)) {
        set_last_error("hg_nchw_f32_to_nhwc_bf16_pad: need c <= c_pad <= 64, c_pad a multiple of 8, 16-byte aligned output");
        return HG_ERR_INVALID;
    }
    const long long tiles = static_cast<long long>(n) * ((static_cast<long long>(h) * w + kTpPix - 1) / kTpPix);
    HG_CUDA_OK(launch_kernel(nchw_to_nhwc_pad_kernel, dim3(grid_for(tiles)), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), in, static_cast<__nv_bfloat16*>(out), n, c, c_pad, h * w));
    return HG_OK;
}

extern "C" int hg_pack_weights(const hg_pack_entry* table_dev, int32_t n_entries, int32_t which, void* stream) {
    if (!table_dev || n_entries <= 0 || which < 1 || which > 3) {
        set_last_error("hg_pack_weights: empty table");
        return HG_ERR_INVALID;
    }
    HG_CUDA_OK(launch_kernel(pack_weights_kernel, dim3(n_entries * kPackBlocksPerEntry), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), table_dev, n_entries, which));
    return HG_OK;
}

extern "C" int hg_rmsprop_step(float* params, const float* grads, float* square_avg, int64_t count, float lr, float alpha,
                               float eps, float grad_scale, void* stream) {
    if (!params || !grads || !square_avg || count <= 0 || (count & 3) || !aligned16(params) || !aligned16(grads) ||
        !aligned16(square_avg)) {
        set_last_error("hg_rmsprop_step: need count %% 4 == 0 and 16-byte aligned buffers");
        return HG_ERR_INVALID;
    }
    const long long n4 = count / 4;
    HG_CUDA_OK(launch_kernel(rmsprop_kernel, dim3(grid_for((n4 + kThreads - 1) / kThreads)), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), reinterpret_cast<float4*>(params),
                             reinterpret_cast<const float4*>(grads), reinterpret_cast<float4*>(square_avg), n4, lr, alpha, eps,
                             grad_scale));
    return HG_OK;
}

extern "C" int hg_small_gemm_f32(float* c, const float* a, const float* b, const float* d, int32_t m, int32_t n, int32_t k,
                                 int32_t sai, int32_t sak, int32_t sbk, int32_t sbj, int32_t sci, int32_t scj, float beta,
                                 void* stream) {
    if (!c || !a || !b || m <= 0 || n <= 0 || k <= 0) {
        set_last_error("hg_small_gemm_f32: bad arguments");
        return HG_ERR_INVALID;
    }
    const long long items = static_cast<long long>(m) * n;      // one warp per output
    HG_CUDA_OK(launch_kernel(small_gemm_kernel, dim3(grid_for((items + kThreads / 32 - 1) / (kThreads / 32))), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), c, a, b, d, m, n, k, sai, sak, sbk, sbj, sci, scj, beta));
    return HG_OK;
}
