// Heat-map decode kernels: one warp per (batch, joint) map, 128-bit loads, warp-shuffle arg-max with
// first-index tie-break (torch.max semantics, src/utils/evaluation.py:14-15 of the reference).
#include "hg_common.cuh"
#include "../../include/hg_api.h"

namespace hg {

constexpr int kDecThreads = 256;            // 8 warps = 8 maps per CTA
constexpr int kMaxHW = 1 << 22;             // flat indices stay exact in fp32 far beyond this

struct ArgMax {
    float v;
    int i;
};

// Flat arg-max over hw floats by one warp.  Ties -> smallest index.  NaN-free input assumed
// (torch.max would propagate NaN; documented in DESIGN.md).
__device__ __forceinline__ ArgMax warp_argmax(const float* __restrict__ m, int hw, int lane) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    if ((hw & 3) == 0 && (reinterpret_cast<uintptr_t>(m) & 15u) == 0) {
        const float4* m4 = reinterpret_cast<const float4*>(m);
        const int n4 = hw >> 2;
        for (int k = lane; k < n4; k += 32) {
            const float4 x = __ldg(m4 + k);
            const int base = k << 2;
            if (x.x > best) { best = x.x; bi = base; }
            if (x.y > best) { best = x.y; bi = base + 1; }
            if (x.z > best) { best = x.z; bi = base + 2; }
            if (x.w > best) { best = x.w; bi = base + 3; }
        }
    } else {
        for (int k = lane; k < hw; k += 32) {
            const float x = __ldg(m + k);
            if (x > best) { best = x; bi = k; }
        }
    }
    // a lane that saw nothing better than -inf keeps bi = INT_MAX; an all -inf map resolves to index 0 below
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (bi == 0x7fffffff) bi = 0;
    return {best, bi};
}

// The reference's 1-based-index quirk (evaluation.py:22-23): x = (idx-1) % W + 1, y = floor((idx-1)/W) + 1
// evaluated with Python modulo semantics; exact in integers for idx < 2^24.
__device__ __forceinline__ void quirk_coords(int idx, int w, float& x, float& y) {
    if (idx == 0) {
        x = static_cast<float>(w);
        y = 0.f;
    } else {
        x = static_cast<float>((idx - 1) % w + 1);
        y = static_cast<float>((idx - 1) / w + 1);
    }
}

__global__ void __launch_bounds__(kDecThreads) decode_argmax_kernel(const float* __restrict__ hm, float* __restrict__ preds,
                                                                     float* __restrict__ maxval, int* __restrict__ argidx,
                                                                     int maps, int h, int w) {
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = gridDim.x * (kDecThreads / 32);
    for (int m = blockIdx.x * (kDecThreads / 32) + (threadIdx.x >> 5); m < maps; m += warps_per_grid) {
        const ArgMax a = warp_argmax(hm + static_cast<long long>(m) * h * w, h * w, lane);
        if (lane == 0) {
            float x, y;
            quirk_coords(a.i, w, x, y);
            const float mask = a.v > 0.f ? 1.f : 0.f;          // evaluation.py:25-26
            preds[2 * m] = x * mask;
            preds[2 * m + 1] = y * mask;
            if (maxval) maxval[m] = a.v;
            if (argidx) argidx[m] = a.i;
        }
    }
}

// Inverse affine of get_affine_transform(center, scale, 0, output_size, inv=1) (transforms.py:40-73):
// the reference builds three float32 point pairs and lets cv2 solve dst -> src in double.
struct Affine {
    double m[6];
};
__device__ Affine inverse_affine(double cx, double cy, double scale0, int out_w, int out_h) {
    const double src_w = scale0 * 200.0;
    const double dst_w = out_w, dst_h = out_h;
    float sx[3], sy[3], dx[3], dy[3];
    sx[0] = static_cast<float>(cx);
    sy[0] = static_cast<float>(cy);
    sx[1] = static_cast<float>(cx + 0.0);
    sy[1] = static_cast<float>(cy + src_w * -0.5);
    dx[0] = static_cast<float>(dst_w * 0.5);
    dy[0] = static_cast<float>(dst_h * 0.5);
    dx[1] = static_cast<float>(dst_w * 0.5 + 0.0);
    dy[1] = static_cast<float>(dst_h * 0.5 + static_cast<double>(static_cast<float>(dst_w * -0.5)));
    // get_3rd_point(a, b) = b + (-(a-b).y, (a-b).x), float32 arithmetic (transforms.py:82-84)
    sx[2] = sx[1] + (-(sy[0] - sy[1]));
    sy[2] = sy[1] + (sx[0] - sx[1]);
    dx[2] = dx[1] + (-(dy[0] - dy[1]));
    dy[2] = dy[1] + (dx[0] - dx[1]);
    // solve [dx dy 1] * (a b c)^T = s for both rows in double, on differences to avoid cancellation
    const double u1 = static_cast<double>(dx[1]) - dx[0], v1 = static_cast<double>(dy[1]) - dy[0];
    const double u2 = static_cast<double>(dx[2]) - dx[0], v2 = static_cast<double>(dy[2]) - dy[0];
    const double det = u1 * v2 - u2 * v1;
    Affine A;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const double s0 = r == 0 ? sx[0] : sy[0], s1 = r == 0 ? sx[1] : sy[1], s2 = r == 0 ? sx[2] : sy[2];
        const double t1 = s1 - s0, t2 = s2 - s0;
        const double a = (t1 * v2 - t2 * v1) / det;
        const double b = (u1 * t2 - u2 * t1) / det;
        A.m[3 * r + 0] = a;
        A.m[3 * r + 1] = b;
        A.m[3 * r + 2] = s0 - a * dx[0] - b * dy[0];
    }
    return A;
}

__global__ void __launch_bounds__(kDecThreads) decode_final_kernel(const float* __restrict__ hm,
                                                                    const double* __restrict__ center,
                                                                    const double* __restrict__ scale,
                                                                    double* __restrict__ out, int nb, int nj, int h, int w,
                                                                    int out_w, int out_h) {
    const int lane = threadIdx.x & 31;
    const int maps = nb * nj;
    const int warps_per_grid = gridDim.x * (kDecThreads / 32);
    for (int m = blockIdx.x * (kDecThreads / 32) + (threadIdx.x >> 5); m < maps; m += warps_per_grid) {
        const float* map = hm + static_cast<long long>(m) * h * w;
        const ArgMax a = warp_argmax(map, h * w, lane);
        if (lane == 0) {
            float x, y;
            quirk_coords(a.i, w, x, y);
            if (!(a.v > 0.f)) { x = 0.f; y = 0.f; }
            // inference.py:54-61: px = floor(x + .5), guard 1 < px < W-1, 1 < py < H-1, quarter-pixel sign shift
            const int px = static_cast<int>(floorf(x + 0.5f));
            const int py = static_cast<int>(floorf(y + 0.5f));
            if (1 < px && px < w - 1 && 1 < py && py < h - 1) {
                const float ddx = map[(py - 1) * w + px] - map[(py - 1) * w + px - 2];
                const float ddy = map[py * w + px - 1] - map[(py - 2) * w + px - 1];
                x += (ddx > 0.f ? 0.25f : (ddx < 0.f ? -0.25f : 0.f));
                y += (ddy > 0.f ? 0.25f : (ddy < 0.f ? -0.25f : 0.f));
            }
            const int b = m / nj;
            const Affine A = inverse_affine(center[2 * b], center[2 * b + 1], scale[2 * b], out_w, out_h);
            const double xd = x, yd = y;
            out[2 * m] = A.m[0] * xd + A.m[1] * yd + A.m[2];
            out[2 * m + 1] = A.m[3] * xd + A.m[4] * yd + A.m[5];
        }
    }
}

// out[b][k][y][x] = 0.5 * (hm[b][k][y][x] + hm_flip[b][perm[k]][y][w-1-x])
__global__ void __launch_bounds__(256) flip_average_kernel(const float* __restrict__ hm, const float* __restrict__ hm_flip,
                                                            const int* __restrict__ perm, float* __restrict__ out,
                                                            int nb, int nj, int h, int w) {
    const long long total = static_cast<long long>(nb) * nj * h * w;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % w);
        long long t = i / w;
        const int y = static_cast<int>(t % h);
        t /= h;
        const int k = static_cast<int>(t % nj);
        const long long b = t / nj;
        const long long src = ((b * nj + __ldg(perm + k)) * h + y) * w + (w - 1 - x);
        out[i] = 0.5f * (__ldg(hm + i) + __ldg(hm_flip + src));
    }
}

// accuracy() device part: normalised distance between arg-max of prediction and of target, or -1
__global__ void __launch_bounds__(kDecThreads) pck_dists_kernel(const float* __restrict__ out_hm,
                                                                 const float* __restrict__ tgt_hm,
                                                                 float* __restrict__ dists, int nb, int nj, int h, int w) {
    const int lane = threadIdx.x & 31;
    const int maps = nb * nj;
    const int warps_per_grid = gridDim.x * (kDecThreads / 32);
    const float norm = static_cast<float>(w) / 10.0f;                   // evaluation.py:61
    for (int m = blockIdx.x * (kDecThreads / 32) + (threadIdx.x >> 5); m < maps; m += warps_per_grid) {
        const ArgMax p = warp_argmax(out_hm + static_cast<long long>(m) * h * w, h * w, lane);
        const ArgMax g = warp_argmax(tgt_hm + static_cast<long long>(m) * h * w, h * w, lane);
        if (lane == 0) {
            float px, py, gx, gy;
            quirk_coords(p.i, w, px, py);
            quirk_coords(g.i, w, gx, gy);
            if (!(p.v > 0.f)) { px = 0.f; py = 0.f; }
            if (!(g.v > 0.f)) { gx = 0.f; gy = 0.f; }
            float d = -1.f;
            if (gx > 1.f && gy > 1.f) {                                 // evaluation.py:36
                const float ex = px - gx, ey = py - gy;
                d = sqrtf(ex * ex + ey * ey) / norm;
            }
            const int b = m / nj, j = m - b * nj;
            dists[static_cast<long long>(j) * nb + b] = d;              // [j][b] like calc_dists
        }
    }
}

// ---------------------------------------------------------------------------------------------
// DARK-style decode: get_final_preds_v2 (src/utils/inference.py:9-45,70-87), one CTA per (image, joint) map.
//   1. flat arg-max (first maximum) and the map's maximum;
//   2. 11x11 separable Gaussian blur (sigma 2, OpenCV's getGaussianKernel(11, 0) weights) of the zero-padded map in
//      float64 -- rows (taps left to right) then columns (centre, then symmetric pairs), no fused multiply-add --
//      each value stored as float32, as the reference's float32 array does (inference.py:43);
//   3. the blurred map rescaled so that its maximum is the original maximum (float32), log(max(., 1e-10));
//   4. second-order Taylor step from the 13-point stencil around the arg-max for the first `refine_joints` joints
//      (the reference's loop runs over coords.shape[1] == 2, i.e. joints 0 and 1 only), float32, 2x2 inverse by
//      LU with partial pivoting as LAPACK's gesv does;
//   5. the inverse affine of get_final_preds_v1, float64.
// ---------------------------------------------------------------------------------------------
__constant__ double c_gauss11[11] = {
    0x1.20c2564ee6772p-7, 0x1.bcb86a082c301p-6, 0x1.0ab50979aaf94p-4, 0x1.f2464c62edaf4p-4, 0x1.6a7e1d504a91dp-3,
    0x1.9ac20a36ea596p-3, 0x1.6a7e1d504a91dp-3, 0x1.f2464c62edaf4p-4, 0x1.0ab50979aaf94p-4, 0x1.bcb86a082c301p-6,
    0x1.20c2564ee6772p-7};

__global__ void __launch_bounds__(kDecThreads) decode_dark_kernel(const float* __restrict__ hm, const double* __restrict__ center,
                                                                   const double* __restrict__ scale, double* __restrict__ out,
                                                                   int nj, int h, int w, int out_w, int out_h,
                                                                   int refine_joints) {
    extern __shared__ __align__(16) unsigned char dark_smem[];
    double* rows = reinterpret_cast<double*>(dark_smem);                 // [h][w] row-pass result
    float* blur = reinterpret_cast<float*>(rows + h * w);                // [h][w] blurred map, float32
    __shared__ float s_red[kDecThreads / 32];
    __shared__ ArgMax s_arg;
    const int m = blockIdx.x;
    const int hw = h * w;
    const float* map = hm + static_cast<long long>(m) * hw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int joint = m % nj;
    const bool refine = joint < refine_joints;
    if (warp == 0) {
        const ArgMax a = warp_argmax(map, hw, lane);
        if (lane == 0) s_arg = a;
    }
    float bmax = -INFINITY;
    if (refine) {
        for (int i = threadIdx.x; i < hw; i += kDecThreads) {
            const int y = i / w, x = i - y * w;
            const float* r = map + y * w;
            double s = __dmul_rn(c_gauss11[0], (x - 5 >= 0) ? static_cast<double>(__ldg(r + x - 5)) : 0.0);
#pragma unroll
            for (int t = 1; t < 11; ++t) {
                const int xx = x + t - 5;
                const double v = (xx >= 0 && xx < w) ? static_cast<double>(__ldg(r + xx)) : 0.0;
                s = __dadd_rn(s, __dmul_rn(c_gauss11[t], v));
            }
            rows[i] = s;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < hw; i += kDecThreads) {
            const int y = i / w, x = i - y * w;
            double s = __dmul_rn(c_gauss11[5], rows[i]);
#pragma unroll
            for (int t = 1; t <= 5; ++t) {
                const double lo = (y - t >= 0) ? rows[(y - t) * w + x] : 0.0;
                const double hi = (y + t < h) ? rows[(y + t) * w + x] : 0.0;
                s = __dadd_rn(s, __dmul_rn(c_gauss11[5 + t], __dadd_rn(hi, lo)));
            }
            const float f = static_cast<float>(s);
            blur[i] = f;
            bmax = fmaxf(bmax, f);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, off));
        if (lane == 0) s_red[warp] = bmax;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const ArgMax a = s_arg;
    float x, y;
    quirk_coords(a.i, w, x, y);
    if (!(a.v > 0.f)) { x = 0.f; y = 0.f; }
    if (refine) {
        float mb = s_red[0];
        for (int i = 1; i < kDecThreads / 32; ++i) mb = fmaxf(mb, s_red[i]);
        const float ratio = __fdiv_rn(a.v, mb);                          // origin_max / max(blurred), float32
        const int px = static_cast<int>(x), py = static_cast<int>(y);    // int(): truncation
        if (1 < px && px < w - 2 && 1 < py && py < h - 2) {
            auto L = [&](int yy, int xx) -> float {
                const float v = fmaxf(__fmul_rn(blur[yy * w + xx], ratio), 1e-10f);
                return static_cast<float>(log(static_cast<double>(v)));
            };
            const float c0 = L(py, px);
            const float dx = __fmul_rn(0.5f, __fsub_rn(L(py, px + 1), L(py, px - 1)));
            const float dy = __fmul_rn(0.5f, __fsub_rn(L(py + 1, px), L(py - 1, px)));
            const float dxx = __fmul_rn(0.25f, __fadd_rn(__fsub_rn(L(py, px + 2), __fmul_rn(2.f, c0)), L(py, px - 2)));
            const float dxy = __fmul_rn(0.25f, __fadd_rn(__fsub_rn(__fsub_rn(L(py + 1, px + 1), L(py - 1, px + 1)),
                                                                   L(py + 1, px - 1)), L(py - 1, px - 1)));
            const float dyy = __fmul_rn(0.25f, __fadd_rn(__fsub_rn(L(py + 2, px), __fmul_rn(2.f, c0)), L(py - 2, px)));
            if (__fsub_rn(__fmul_rn(dxx, dyy), __fmul_rn(dxy, dxy)) != 0.f) {
                // inverse of [[dxx, dxy], [dxy, dyy]]: LU with partial pivoting, then the two unit right-hand sides
                float a00 = dxx, a01 = dxy, a10 = dxy, a11 = dyy;
                const bool swap = fabsf(a10) > fabsf(a00);
                if (swap) { float t0 = a00; a00 = a10; a10 = t0; t0 = a01; a01 = a11; a11 = t0; }
                const float l = __fmul_rn(a10, __fdiv_rn(1.f, a00));
                const float u11 = __fsub_rn(a11, __fmul_rn(l, a01));
                // columns of the inverse: solve for P*e0 and P*e1
                float inv[2][2];
#pragma unroll
                for (int col = 0; col < 2; ++col) {
                    float b0 = col == 0 ? 1.f : 0.f, b1 = col == 0 ? 0.f : 1.f;
                    if (swap) { const float t0 = b0; b0 = b1; b1 = t0; }
                    const float y1 = __fsub_rn(b1, __fmul_rn(l, b0));
                    const float x1 = __fdiv_rn(y1, u11);
                    const float x0 = __fdiv_rn(__fsub_rn(b0, __fmul_rn(a01, x1)), a00);
                    inv[0][col] = x0;
                    inv[1][col] = x1;
                }
                const float ox = -__fadd_rn(__fmul_rn(inv[0][0], dx), __fmul_rn(inv[0][1], dy));
                const float oy = -__fadd_rn(__fmul_rn(inv[1][0], dx), __fmul_rn(inv[1][1], dy));
                x = __fadd_rn(x, ox);
                y = __fadd_rn(y, oy);
            }
        }
    }
    const int b = m / nj;
    const Affine A = inverse_affine(center[2 * b], center[2 * b + 1], scale[2 * b], out_w, out_h);
    const double xd = x, yd = y;
    out[2 * m] = A.m[0] * xd + A.m[1] * yd + A.m[2];
    out[2 * m + 1] = A.m[3] * xd + A.m[4] * yd + A.m[5];
}

static int dec_grid(int maps) {
    const int blocks = (maps + kDecThreads / 32 - 1) / (kDecThreads / 32);
    const int cap = num_sms() * 8;
    return blocks < cap ? (blocks > 0 ? blocks : 1) : cap;
}

}  // namespace hg

using namespace hg;

static int check_maps(const char* fn, int32_t b, int32_t j, int32_t h, int32_t w) {
    if (b <= 0 || j <= 0 || h <= 0 || w <= 0 || static_cast<long long>(h) * w > kMaxHW ||
        static_cast<long long>(b) * j > 0x7fffffffLL / 4) {
        set_last_error("%s: bad shape b=%d j=%d h=%d w=%d", fn, b, j, h, w);
        return HG_ERR_INVALID;
    }
    return HG_OK;
}

extern "C" int hg_decode_argmax(const float* hm, float* preds, float* maxval, int32_t* argidx, int32_t b, int32_t j,
                                int32_t h, int32_t w, void* stream) {
    int rc = check_maps("hg_decode_argmax", b, j, h, w);
    if (rc) return rc;
    if (!hm || !preds) {
        set_last_error("hg_decode_argmax: null pointer");
        return HG_ERR_INVALID;
    }
    decode_argmax_kernel<<<dec_grid(b * j), kDecThreads, 0, static_cast<cudaStream_t>(stream)>>>(hm, preds, maxval, argidx,
                                                                                                  b * j, h, w);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}

extern "C" int hg_decode_final_preds(const float* hm, const double* center, const double* scale, double* out, int32_t b,
                                     int32_t j, int32_t h, int32_t w, int32_t out_w, int32_t out_h, void* stream) {
    int rc = check_maps("hg_decode_final_preds", b, j, h, w);
    if (rc) return rc;
    if (!hm || !center || !scale || !out || out_w <= 0 || out_h <= 0) {
        set_last_error("hg_decode_final_preds: null pointer or bad output size");
        return HG_ERR_INVALID;
    }
    decode_final_kernel<<<dec_grid(b * j), kDecThreads, 0, static_cast<cudaStream_t>(stream)>>>(hm, center, scale, out, b, j,
                                                                                                 h, w, out_w, out_h);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}

extern "C" int hg_flip_average(const float* hm, const float* hm_flip, const int32_t* perm, float* out, int32_t b, int32_t j,
                               int32_t h, int32_t w, void* stream) {
    int rc = check_maps("hg_flip_average", b, j, h, w);
    if (rc) return rc;
    if (!hm || !hm_flip || !perm || !out) {
        set_last_error("hg_flip_average: null pointer");
        return HG_ERR_INVALID;
    }
    const long long total = static_cast<long long>(b) * j * h * w;
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    flip_average_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(hm, hm_flip, perm, out, b, j,
                                                                                                  h, w);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}

extern "C" int hg_pck_dists(const float* out_hm, const float* tgt_hm, float* dists, int32_t b, int32_t j, int32_t h,
                            int32_t w, void* stream) {
    int rc = check_maps("hg_pck_dists", b, j, h, w);
    if (rc) return rc;
    if (!out_hm || !tgt_hm || !dists) {
        set_last_error("hg_pck_dists: null pointer");
        return HG_ERR_INVALID;
    }
    pck_dists_kernel<<<dec_grid(b * j), kDecThreads, 0, static_cast<cudaStream_t>(stream)>>>(out_hm, tgt_hm, dists, b, j, h,
                                                                                              w);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}

extern "C" int hg_decode_final_preds_v2(const float* hm, const double* center, const double* scale, double* out, int32_t b,
                                        int32_t j, int32_t h, int32_t w, int32_t out_w, int32_t out_h, int32_t refine_joints,
                                        void* stream) {
    int rc = check_maps("hg_decode_final_preds_v2", b, j, h, w);
    if (rc) return rc;
    const size_t smem = static_cast<size_t>(h) * w * (sizeof(double) + sizeof(float));
    if (!hm || !center || !scale || !out || out_w <= 0 || out_h <= 0 || refine_joints < 0 || smem > 200 * 1024) {
        set_last_error("hg_decode_final_preds_v2: null pointer, bad output size, or map larger than 128x128");
        return HG_ERR_INVALID;
    }
    HG_CUDA_OK(cudaFuncSetAttribute(decode_dark_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    decode_dark_kernel<<<b * j, kDecThreads, smem, static_cast<cudaStream_t>(stream)>>>(hm, center, scale, out, j, h, w, out_w,
                                                                                         out_h, refine_joints);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}
