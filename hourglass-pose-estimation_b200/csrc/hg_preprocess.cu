// Input preprocessing on the device: the step immediately before the hot path.
//
//   hg_normalize_u8_nhwc    : transforms.ToTensor() + Normalize(mean, std) (src/datasets/common.py:57-64) on uint8
//                             HWC crops -- float32 x/255 then (x-mean)/std, IEEE division, so the fp32 NCHW result
//                             is bit-identical to torch's; optionally also written straight into the stem's packed
//                             NHWC4 bf16 staging image (hg_stem_pack's layout), which removes the fp32 round trip.
//   hg_preprocess_frames_u8 : Estimator.preprocess_bbox (src/runner/estimator.py:39-54) for a batch of equally
//                             sized uint8 frames: /255, per-dataset mean/std in float64, bilinear resize with
//                             OpenCV's INTER_LINEAR tap positions (float32 weights from (d+0.5)*scale-0.5, horizontal
//                             pass then vertical pass, float64, no fused multiply-add), cast to float32 NCHW.
//
// Both are HBM/PCIe-side byte work: one thread per output pixel, coalesced planar stores.  uint8 input is 4x fewer
// host->device bytes than the fp32 NCHW tensor the reference ships.
#include "hg_common.cuh"
#include "../../include/hg_api.h"

namespace hg {
namespace pre {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) normalize_u8_kernel(const uint8_t* __restrict__ in, float* out_nchw, uint2* packed,
                                                                 float m0, float m1, float m2, float s0, float s1, float s2,
                                                                 int n, int h, int w, int flip_w) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = static_cast<long long>(n) * h * w;
    const long long plane = static_cast<long long>(h) * w;
    for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        const uint8_t* px = in + i * 3;
        const float r = (static_cast<float>(px[0]) / 255.0f - m0) / s0;
        const float g = (static_cast<float>(px[1]) / 255.0f - m1) / s1;
        const float b = (static_cast<float>(px[2]) / 255.0f - m2) / s2;
        const long long row = i / w;                 // image*h + y
        const int x = static_cast<int>(i - row * w);
        const long long img = row / h;
        if (out_nchw != nullptr) {
            float* o = out_nchw + img * 3 * plane + (row - img * h) * w + x;
            o[0] = r;
            o[plane] = g;
            o[2 * plane] = b;
        }
        if (packed != nullptr) {
            uint2 o;
            o.x = pack_bf16x2(r, g);
            o.y = pack_bf16x2(b, 0.f);
            packed[row * (w + 8) + 4 + (flip_w ? (w - 1 - x) : x)] = o;
        }
    }
}

struct ResizeParams {
    const uint8_t* frames;
    float* out;
    double mean[3], stdv[3];
    int normalize;
    int n, fh, fw, h, w;
    double scale_x, scale_y;
};

// OpenCV's linear tap for destination index d: source index (may be -1) and float32 weight of the second sample
__device__ __forceinline__ void linear_tap(int d, double scale, int& s, float& f) {
    const float fx = static_cast<float>(__dadd_rn(__dmul_rn(static_cast<double>(d) + 0.5, scale), -0.5));
    const float fl = floorf(fx);
    s = static_cast<int>(fl);
    f = __fsub_rn(fx, fl);
}

__device__ __forceinline__ double norm_px(const ResizeParams& p, const uint8_t* px, int c) {
    double v = __ddiv_rn(static_cast<double>(px[c]), 255.0);
    if (p.normalize) v = __ddiv_rn(__dsub_rn(v, p.mean[c]), p.stdv[c]);
    return v;
}

__global__ void __launch_bounds__(kThreads) preprocess_frames_kernel(const ResizeParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = static_cast<long long>(p.n) * p.h * p.w;
    const long long plane = static_cast<long long>(p.h) * p.w;
    const bool same = (p.fh == p.h && p.fw == p.w);
    for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        const int dx = static_cast<int>(i % p.w);
        const long long t = i / p.w;
        const int dy = static_cast<int>(t % p.h);
        const long long img = t / p.h;
        const uint8_t* frame = p.frames + img * p.fh * p.fw * 3;
        float* o = p.out + img * 3 * plane + static_cast<long long>(dy) * p.w + dx;
        if (same) {
            const uint8_t* px = frame + (static_cast<long long>(dy) * p.fw + dx) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c * plane] = static_cast<float>(norm_px(p, px, c));
            continue;
        }
        int sx, sy;
        float fx, fy;
        linear_tap(dx, p.scale_x, sx, fx);
        linear_tap(dy, p.scale_y, sy, fy);
        // horizontal taps: left of the image -> sample 0 with weight 0 on its neighbour; at or right of the last column ->
        // the last sample alone
        bool single = false;
        if (sx < 0) { sx = 0; fx = 0.f; }
        if (sx >= p.fw - 1) { sx = p.fw - 1; single = true; }
        const double a0 = static_cast<double>(__fsub_rn(1.0f, fx)), a1 = static_cast<double>(fx);
        // vertical taps: both rows clamped, weights untouched
        const int y0 = min(max(sy, 0), p.fh - 1), y1 = min(max(sy + 1, 0), p.fh - 1);
        const double b0 = static_cast<double>(__fsub_rn(1.0f, fy)), b1 = static_cast<double>(fy);
        const uint8_t* r0 = frame + (static_cast<long long>(y0) * p.fw + sx) * 3;
        const uint8_t* r1 = frame + (static_cast<long long>(y1) * p.fw + sx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double h0, h1;
            if (single) {
                h0 = norm_px(p, r0, c);
                h1 = norm_px(p, r1, c);
            } else {
                h0 = __dadd_rn(__dmul_rn(norm_px(p, r0, c), a0), __dmul_rn(norm_px(p, r0 + 3, c), a1));
                h1 = __dadd_rn(__dmul_rn(norm_px(p, r1, c), a0), __dmul_rn(norm_px(p, r1 + 3, c), a1));
            }
            o[c * plane] = static_cast<float>(__dadd_rn(__dmul_rn(h0, b0), __dmul_rn(h1, b1)));
        }
    }
}

static inline int grid_of(long long items) {
    const long long blocks = (items + kThreads - 1) / kThreads;
    const long long cap = static_cast<long long>(num_sms()) * 8;
    return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace pre
}  // namespace hg

using namespace hg;
using namespace hg::pre;

extern "C" int hg_normalize_u8_nhwc(const void* in_u8, const float* mean3, const float* std3, float* out_nchw, void* packed,
                                    int32_t n, int32_t h, int32_t w, int32_t flip_w, void* stream) {
    if (!in_u8 || !mean3 || !std3 || (!out_nchw && !packed) || n <= 0 || h <= 0 || w <= 0 ||
        (packed && (reinterpret_cast<uintptr_t>(packed) & 7u))) {
        set_last_error("hg_normalize_u8_nhwc: bad arguments (mean3/std3 are HOST float[3]; one of out_nchw / packed required)");
        return HG_ERR_INVALID;
    }
    for (int c = 0; c < 3; ++c)
        if (!(std3[c] != 0.f)) {
            set_last_error("hg_normalize_u8_nhwc: std[%d] is zero", c);
            return HG_ERR_INVALID;
        }
    HG_CUDA_OK(launch_kernel(normalize_u8_kernel, dim3(grid_of(static_cast<long long>(n) * h * w)), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), static_cast<const uint8_t*>(in_u8), out_nchw,
                             static_cast<uint2*>(packed), mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], n, h, w,
                             flip_w));
    return HG_OK;
}

extern "C" int hg_preprocess_frames_u8(const void* frames_u8, const double* mean3, const double* std3, float* out_nchw,
                                       int32_t n, int32_t fh, int32_t fw, int32_t h, int32_t w, void* stream) {
    if (!frames_u8 || !out_nchw || n <= 0 || fh <= 0 || fw <= 0 || h <= 0 || w <= 0 || ((mean3 == nullptr) != (std3 == nullptr))) {
        set_last_error("hg_preprocess_frames_u8: bad arguments (mean3/std3 are HOST double[3] or both NULL)");
        return HG_ERR_INVALID;
    }
    ResizeParams p;
    p.frames = static_cast<const uint8_t*>(frames_u8);
    p.out = out_nchw;
    p.normalize = mean3 != nullptr;
    for (int c = 0; c < 3; ++c) {
        p.mean[c] = mean3 ? mean3[c] : 0.0;
        p.stdv[c] = std3 ? std3[c] : 1.0;
        if (p.stdv[c] == 0.0) {
            set_last_error("hg_preprocess_frames_u8: std[%d] is zero", c);
            return HG_ERR_INVALID;
        }
    }
    p.n = n; p.fh = fh; p.fw = fw; p.h = h; p.w = w;
    // OpenCV: inv_scale = dsize/ssize in double, scale = 1/inv_scale
    p.scale_x = 1.0 / (static_cast<double>(w) / static_cast<double>(fw));
    p.scale_y = 1.0 / (static_cast<double>(h) / static_cast<double>(fh));
    HG_CUDA_OK(launch_kernel(preprocess_frames_kernel, dim3(grid_of(static_cast<long long>(n) * h * w)), dim3(kThreads), 0,
                             static_cast<cudaStream_t>(stream), p));
    return HG_OK;
}
