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
