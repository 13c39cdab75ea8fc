// Bandwidth-bound NHWC bf16 kernels: 128-bit vectorised, one 8-channel chunk per thread, grid sized
// in multiples of the SM count with a grid-stride loop.
#include "hg_common.cuh"
#include "../../include/hg_api.h"

namespace hg {

constexpr int kEwThreads = 256;

static inline int ew_grid(long long work_items) {
    const long long blocks = (work_items + kEwThreads - 1) / kEwThreads;
    const long long cap = static_cast<long long>(num_sms()) * 8;     // 8 resident CTAs of 256 threads per SM
    return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a);
    __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b);
    __nv_bfloat162 r = __hmax2(x, y);
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t bf16x2_add_f32(uint32_t a, uint32_t b) {   // fp32 add, one rounding
    return pack_bf16x2(bf16_lo_to_f32(a) + bf16_lo_to_f32(b), bf16_hi_to_f32(a) + bf16_hi_to_f32(b));
}

// ---------------------------------------------------------------- max pool 2x2 stride 2
__global__ void __launch_bounds__(kEwThreads) maxpool2x2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                                 int n, int h, int w, int c8) {
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
        const uint4 v00 = ldg_nc_v4(in + base);
        const uint4 v01 = ldg_nc_v4(in + base + c8);
        const uint4 v10 = ldg_nc_v4(in + base + static_cast<long long>(w) * c8);
        const uint4 v11 = ldg_nc_v4(in + base + static_cast<long long>(w) * c8 + c8);
        uint4 r;
        r.x = bf16x2_max(bf16x2_max(v00.x, v01.x), bf16x2_max(v10.x, v11.x));
        r.y = bf16x2_max(bf16x2_max(v00.y, v01.y), bf16x2_max(v10.y, v11.y));
        r.z = bf16x2_max(bf16x2_max(v00.z, v01.z), bf16x2_max(v10.z, v11.z));
        r.w = bf16x2_max(bf16x2_max(v00.w, v01.w), bf16x2_max(v10.w, v11.w));
        out[i] = r;
    }
}

// ---------------------------------------------------------------- out = a + upsample2x(low)
__global__ void __launch_bounds__(kEwThreads) upsample_add_kernel(const uint4* __restrict__ a,
                                                                   const uint4* __restrict__ low,
                                                                   uint4* __restrict__ out, int n, int h, int w, int c8) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = static_cast<long long>(n) * h * w * c8;
    const int lh = h >> 1, lw = w >> 1;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(i % c8);
        long long pix = i / c8;
        const int x = static_cast<int>(pix % w);
        pix /= w;
        const int y = static_cast<int>(pix % h);
        const long long b = pix / h;
        const uint4 va = ldg_nc_v4(a + i);
        const uint4 vl = ldg_v4(low + ((b * lh + (y >> 1)) * lw + (x >> 1)) * c8 + ch);
        uint4 r;
        r.x = bf16x2_add_f32(va.x, vl.x);
        r.y = bf16x2_add_f32(va.y, vl.y);
        r.z = bf16x2_add_f32(va.z, vl.z);
        r.w = bf16x2_add_f32(va.w, vl.w);
        out[i] = r;
    }
}

// ---------------------------------------------------------------- out = relu(x*scale + shift)
__global__ void __launch_bounds__(kEwThreads) bn_relu_kernel(const uint4* __restrict__ in,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift, uint4* __restrict__ out,
                                                              long long pixels, int c8) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = pixels * c8;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c0 = static_cast<int>(i % c8) * 8;
        const uint4 v = ldg_nc_v4(in + i);
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c0));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + c0 + 4));
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c0));
        const float4 h1 = __ldg(reinterpret_cast<const float4*>(shift + c0 + 4));
        uint4 r;
        r.x = pack_bf16x2(fmaxf(fmaf(bf16_lo_to_f32(v.x), s0.x, h0.x), 0.f), fmaxf(fmaf(bf16_hi_to_f32(v.x), s0.y, h0.y), 0.f));
        r.y = pack_bf16x2(fmaxf(fmaf(bf16_lo_to_f32(v.y), s0.z, h0.z), 0.f), fmaxf(fmaf(bf16_hi_to_f32(v.y), s0.w, h0.w), 0.f));
        r.z = pack_bf16x2(fmaxf(fmaf(bf16_lo_to_f32(v.z), s1.x, h1.x), 0.f), fmaxf(fmaf(bf16_hi_to_f32(v.z), s1.y, h1.y), 0.f));
        r.w = pack_bf16x2(fmaxf(fmaf(bf16_lo_to_f32(v.w), s1.z, h1.z), 0.f), fmaxf(fmaf(bf16_hi_to_f32(v.w), s1.w, h1.w), 0.f));
        out[i] = r;
    }
}

// ---------------------------------------------------------------- stem im2col: NCHW fp32 -> [n*oh*ow][192] bf16
// k = (ky*7 + kx)*3 + c  for k < 147, zero for 147 <= k < 192.  One thread = one 16-byte chunk (8 k values).
constexpr int kStemK = 192;
__global__ void __launch_bounds__(kEwThreads) stem_im2col_kernel(const float* __restrict__ in, uint4* __restrict__ out,
                                                                  int n, int h, int w, int flip_w) {
    pdl_launch_dependents();
    pdl_wait();
    const int oh = h >> 1, ow = w >> 1;
    const long long total = static_cast<long long>(n) * oh * ow * (kStemK / 8);
    const long long plane = static_cast<long long>(h) * w;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int chunk = static_cast<int>(i % (kStemK / 8));
        long long pix = i / (kStemK / 8);
        const int ox = static_cast<int>(pix % ow);
        pix /= ow;
        const int oy = static_cast<int>(pix % oh);
        const long long b = pix / oh;
        const float* img = in + b * 3 * plane;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = chunk * 8 + e;
            float v = 0.f;
            if (k < 147) {
                const int tap = k / 3, c = k - tap * 3;
                const int ky = tap / 7, kx = tap - ky * 7;
                const int iy = 2 * oy - 3 + ky;
                int ix = 2 * ox - 3 + kx;
                if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
                    if (flip_w) ix = w - 1 - ix;
                    v = __ldg(img + c * plane + static_cast<long long>(iy) * w + ix);
                }
            }
            f[e] = v;
        }
        uint4 r;
        r.x = pack_bf16x2(f[0], f[1]);
        r.y = pack_bf16x2(f[2], f[3]);
        r.z = pack_bf16x2(f[4], f[5]);
        r.w = pack_bf16x2(f[6], f[7]);
        out[i] = r;
    }
}

// ---------------------------------------------------------------- layout converters (API edge / tests)
__global__ void __launch_bounds__(kEwThreads) nchw_to_nhwc_kernel(const float* __restrict__ in,
                                                                   __nv_bfloat16* __restrict__ out, int n, int c, int hw) {
    const long long total = static_cast<long long>(n) * hw * c;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(i % c);
        const long long t = i / c;
        const int p = static_cast<int>(t % hw);
        const long long b = t / hw;
        out[i] = __float2bfloat16_rn(in[(b * c + ch) * hw + p]);
    }
}
__global__ void __launch_bounds__(kEwThreads) nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ in,
                                                                   float* __restrict__ out, int n, int c, int hw) {
    const long long total = static_cast<long long>(n) * hw * c;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(i % hw);
        const long long t = i / hw;
        const int ch = static_cast<int>(t % c);
        const long long b = t / c;
        out[i] = __bfloat162float(in[(b * hw + p) * c + ch]);
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace hg

using namespace hg;

extern "C" int hg_maxpool2x2_nhwc(const void* in, void* out, int32_t n, int32_t h, int32_t w, int32_t c, void* stream) {
    if (!in || !out || n <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || c <= 0 || (c & 7) || !aligned16(in) ||
        !aligned16(out)) {
        set_last_error("hg_maxpool2x2_nhwc: need even h,w, c %% 8 == 0, 16-byte aligned pointers");
        return HG_ERR_INVALID;
    }
    const long long items = static_cast<long long>(n) * (h / 2) * (w / 2) * (c / 8);
    HG_CUDA_OK(launch_kernel(maxpool2x2_kernel, dim3(ew_grid(items)), dim3(kEwThreads), 0, static_cast<cudaStream_t>(stream),
                             static_cast<const uint4*>(in), static_cast<uint4*>(out), n, h, w, c / 8));
    return HG_OK;
}

extern "C" int hg_upsample2x_add_nhwc(const void* a, const void* low, void* out, int32_t n, int32_t h, int32_t w,
                                      int32_t c, void* stream) {
    if (!a || !low || !out || n <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || c <= 0 || (c & 7) || !aligned16(a) ||
        !aligned16(low) || !aligned16(out)) {
        set_last_error("hg_upsample2x_add_nhwc: need even h,w, c %% 8 == 0, 16-byte aligned pointers");
        return HG_ERR_INVALID;
    }
    const long long items = static_cast<long long>(n) * h * w * (c / 8);
    upsample_add_kernel<<<ew_grid(items), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(a), static_cast<const uint4*>(low), static_cast<uint4*>(out), n, h, w, c / 8);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}

extern "C" int hg_bn_relu_nhwc(const void* in, const float* scale, const float* shift, void* out, int64_t pixels,
                               int32_t c, void* stream) {
    if (!in || !scale || !shift || !out || pixels <= 0 || c <= 0 || (c & 7) || !aligned16(in) || !aligned16(out) ||
        !aligned16(scale) || !aligned16(shift)) {
        set_last_error("hg_bn_relu_nhwc: need c %% 8 == 0 and 16-byte aligned pointers");
        return HG_ERR_INVALID;
    }
    bn_relu_kernel<<<ew_grid(pixels * (c / 8)), kEwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(in), scale, shift, static_cast<uint4*>(out), pixels, c / 8);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}

extern "C" int hg_stem_im2col(const float* in_nchw, void* out_rows, int32_t n, int32_t h, int32_t w, int32_t flip_w,
                              void* stream) {
    if (!in_nchw || !out_rows || n <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || !aligned16(out_rows)) {
        set_last_error("hg_stem_im2col: need even h,w and a 16-byte aligned output");
        return HG_ERR_INVALID;
    }
    const long long items = static_cast<long long>(n) * (h / 2) * (w / 2) * (kStemK / 8);
    HG_CUDA_OK(launch_kernel(stem_im2col_kernel, dim3(ew_grid(items)), dim3(kEwThreads), 0, static_cast<cudaStream_t>(stream),
                             in_nchw, static_cast<uint4*>(out_rows), n, h, w, flip_w));
    return HG_OK;
}

extern "C" int hg_nchw_f32_to_nhwc_bf16(const float* in, void* out, int32_t n, int32_t c, int32_t h, int32_t w,
                                        void* stream) {
    if (!in || !out || n <= 0 || c <= 0 || h <= 0 || w <= 0) {
        set_last_error("hg_nchw_f32_to_nhwc_bf16: bad arguments");
        return HG_ERR_INVALID;
    }
    nchw_to_nhwc_kernel<<<ew_grid(static_cast<long long>(n) * c * h * w), kEwThreads, 0,
                          static_cast<cudaStream_t>(stream)>>>(in, static_cast<__nv_bfloat16*>(out), n, c, h * w);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}

extern "C" int hg_nhwc_bf16_to_nchw_f32(const void* in, float* out, int32_t n, int32_t c, int32_t h, int32_t w,
                                        void* stream) {
    if (!in || !out || n <= 0 || c <= 0 || h <= 0 || w <= 0) {
        set_last_error("hg_nhwc_bf16_to_nchw_f32: bad arguments");
        return HG_ERR_INVALID;
    }
    nhwc_to_nchw_kernel<<<ew_grid(static_cast<long long>(n) * c * h * w), kEwThreads, 0,
                          static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(in), out, n, c, h * w);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}
