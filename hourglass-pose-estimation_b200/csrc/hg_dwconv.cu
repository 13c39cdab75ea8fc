// Depthwise 3x3 convolution (mobile=True bottlenecks: nn.Conv2d(planes, planes, 3, padding=1, groups=planes),
// reference src/models/modules.py:15-17) on NHWC bf16 -- 9 MACs per element, so this is a CUDA-core stencil
// bound by HBM bandwidth, not a tensor-core GEMM.
//
//   forward / dgrad:  one thread = one 8-channel chunk x one column x a strip of kRows output rows; the strip's
//                     kRows+2 input rows are each loaded once (three 128-bit loads per row: left, centre, right
//                     neighbours hit L1/L2), the 72 weights of the chunk stay in registers across the grid-stride
//                     loop.  dgrad is the same stencil with the taps reversed (flip_taps).
//   wgrad:            dw[c][tap] += sum_pixels dout[p][c] * z[p + tap][c]: per-thread fp32 partial sums for its
//                     chunk, warp shuffles across the lanes sharing a chunk, shared-memory atomics per block,
//                     one red.global per (channel, tap) per block.
//
// Weights are fp32 [c][9] (channel-major: exactly torch's [c,1,3,3] memory), bias fp32 [c] or null.
#include "hg_common.cuh"
#include "../../include/hg_api.h"

namespace hg {
namespace dw {

constexpr int kThreads = 256;
constexpr int kRows = 4;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = bf16_lo_to_f32(v.x); f[1] = bf16_hi_to_f32(v.x);
    f[2] = bf16_lo_to_f32(v.y); f[3] = bf16_hi_to_f32(v.y);
    f[4] = bf16_lo_to_f32(v.z); f[5] = bf16_hi_to_f32(v.z);
    f[6] = bf16_lo_to_f32(v.w); f[7] = bf16_hi_to_f32(v.w);
}

struct Params {
    const uint4* x;
    const float* w;        // [c][9]
    const float* bias;     // [c] or null
    uint4* out;
    int n, h, w_, c8;
    int relu, flip;
};

__global__ void __launch_bounds__(kThreads) dwconv3x3_kernel(const Params p) {
    pdl_launch_dependents();
    const int c8 = p.c8;
    const int chunk = threadIdx.x & (c8 - 1);          // c8 is a power of two; the grid stride is a multiple of it
    float wr[9][8], br[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) wr[t][e] = p.w[(chunk * 8 + e) * 9 + (p.flip ? 8 - t : t)];
#pragma unroll
    for (int e = 0; e < 8; ++e) br[e] = p.bias != nullptr ? p.bias[chunk * 8 + e] : 0.f;
    pdl_wait();                                        // weights are constants of the launch; activations are not
    const int strips = (p.h + kRows - 1) / kRows;
    const long long total = static_cast<long long>(p.n) * strips * p.w_ * c8;
    const long long row_pitch = static_cast<long long>(p.w_) * c8;
    for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        long long t = i / c8;
        const int x = static_cast<int>(t % p.w_);
        t /= p.w_;
        const int strip = static_cast<int>(t % strips);
        const long long b = t / strips;
        const int y0 = strip * kRows;
        const uint4* img = p.x + b * p.h * row_pitch + chunk;
        float acc[kRows][8];
#pragma unroll
        for (int j = 0; j < kRows; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[j][e] = br[e];
#pragma unroll
        for (int r = -1; r <= kRows; ++r) {
            const int y = y0 + r;
            if (y < 0 || y >= p.h) continue;
            const uint4* row = img + y * row_pitch;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = x + kx - 1;
                if (xx < 0 || xx >= p.w_) continue;
                float f[8];
                unpack8(ldg_v4(row + static_cast<long long>(xx) * c8), f);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int j = r - (ky - 1);           // output row (within the strip) this input row feeds via tap ky
                    if (j < 0 || j >= kRows) continue;
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[j][e] = fmaf(f[e], wr[ky * 3 + kx][e], acc[j][e]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kRows; ++j) {
            const int y = y0 + j;
            if (y >= p.h) break;
            uint4 o;
            if (p.relu) {
                o.x = pack_bf16x2_relu(acc[j][0], acc[j][1]);
                o.y = pack_bf16x2_relu(acc[j][2], acc[j][3]);
                o.z = pack_bf16x2_relu(acc[j][4], acc[j][5]);
                o.w = pack_bf16x2_relu(acc[j][6], acc[j][7]);
            } else {
                o.x = pack_bf16x2(acc[j][0], acc[j][1]);
                o.y = pack_bf16x2(acc[j][2], acc[j][3]);
                o.z = pack_bf16x2(acc[j][4], acc[j][5]);
                o.w = pack_bf16x2(acc[j][6], acc[j][7]);
            }
            p.out[(b * p.h + y) * row_pitch + static_cast<long long>(x) * c8 + chunk] = o;
        }
    }
}

// dw[c][tap] += sum over pixels of dout[p][c] * z[p + tap][c]
__global__ void __launch_bounds__(kThreads) dwconv3x3_wgrad_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ z,
                                                                    float* dw, int n, int h, int w_, int c8) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_acc[256 * 9];                   // [c][9], c <= 256
    const int C = c8 * 8;
    for (int i = threadIdx.x; i < C * 9; i += kThreads) s_acc[i] = 0.f;
    __syncthreads();
    const int chunk = threadIdx.x & (c8 - 1);
    float acc[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[t][e] = 0.f;
    const long long total = static_cast<long long>(n) * h * w_ * c8;
    const long long row_pitch = static_cast<long long>(w_) * c8;
    for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        long long t = i / c8;
        const int x = static_cast<int>(t % w_);
        t /= w_;
        const int y = static_cast<int>(t % h);
        const long long b = t / h;
        float g[8];
        unpack8(ldg_nc_v4(dout + i), g);
        const uint4* img = z + b * h * row_pitch + chunk;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            if (yy < 0 || yy >= h) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = x + kx - 1;
                if (xx < 0 || xx >= w_) continue;
                float f[8];
                unpack8(ldg_v4(img + yy * row_pitch + static_cast<long long>(xx) * c8), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[ky * 3 + kx][e] = fmaf(g[e], f[e], acc[ky * 3 + kx][e]);
            }
        }
    }
    // lanes of a warp that share a chunk: lane, lane + c8, ... (c8 in {1,2,4,8,16,32})
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float v = acc[t][e];
            for (int o = 16; o >= c8; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[t][e] = v;
        }
    if ((threadIdx.x & 31) < c8 || c8 >= 32) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&s_acc[(chunk * 8 + e) * 9 + t], acc[t][e]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 9; i += kThreads) atomicAdd(dw + i, s_acc[i]);
}

static bool ok_channels(int c) {
    const int c8 = c / 8;
    return c % 8 == 0 && c8 >= 1 && c8 <= 32 && (c8 & (c8 - 1)) == 0;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace dw
}  // namespace hg

using namespace hg;
using namespace hg::dw;

extern "C" int hg_dwconv3x3_nhwc(const void* x, const float* w, const float* bias, void* out, int32_t n, int32_t h, int32_t w_,
                                 int32_t c, int32_t relu, int32_t flip_taps, void* stream) {
    if (!x || !w || !out || n <= 0 || h <= 0 || w_ <= 0 || !ok_channels(c) || !aligned16(x) || !aligned16(out) || x == out) {
        set_last_error("hg_dwconv3x3_nhwc: need c/8 a power of two <= 32, 16-byte aligned distinct tensors");
        return HG_ERR_INVALID;
    }
    Params p;
    p.x = static_cast<const uint4*>(x);
    p.w = w;
    p.bias = bias;
    p.out = static_cast<uint4*>(out);
    p.n = n; p.h = h; p.w_ = w_; p.c8 = c / 8;
    p.relu = relu;
    p.flip = flip_taps;
    const long long items = static_cast<long long>(n) * ((h + kRows - 1) / kRows) * w_ * (c / 8);
    long long blocks = (items + kThreads - 1) / kThreads;
    const long long cap = static_cast<long long>(num_sms()) * 4;
    if (blocks > cap) blocks = cap;
    HG_CUDA_OK(launch_kernel(dwconv3x3_kernel, dim3(static_cast<unsigned>(blocks)), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), p));
    return HG_OK;
}

extern "C" int hg_dwconv3x3_wgrad(const void* dout, const void* z, float* dw, int32_t n, int32_t h, int32_t w_, int32_t c,
                                  void* stream) {
    if (!dout || !z || !dw || n <= 0 || h <= 0 || w_ <= 0 || !ok_channels(c) || !aligned16(dout) || !aligned16(z)) {
        set_last_error("hg_dwconv3x3_wgrad: need c/8 a power of two <= 32, 16-byte aligned tensors");
        return HG_ERR_INVALID;
    }
    const long long items = static_cast<long long>(n) * h * w_ * (c / 8);
    long long blocks = (items + kThreads - 1) / kThreads;
    const long long cap = static_cast<long long>(num_sms()) * 2;
    if (blocks > cap) blocks = cap;
    HG_CUDA_OK(launch_kernel(dwconv3x3_wgrad_kernel, dim3(static_cast<unsigned>(blocks)), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), static_cast<const uint4*>(dout), static_cast<const uint4*>(z),
                             dw, n, h, w_, c / 8));
    return HG_OK;
}
