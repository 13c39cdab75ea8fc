// Stem of the hourglass: conv 7x7 stride 2 pad 3 (3 -> 64) + folded BatchNorm + ReLU
// (reference: src/models/hourglass.py:17-20,71-73), without ever materialising an im2col matrix.
//
//   1. hg_stem_pack: NCHW fp32 image -> NHWC4 bf16 with 4 zero pixels of padding left and right
//      ([n][h][w+8][4], 8 bytes per pixel).  Optional left-right mirroring (flip test).
//   2. hg_stem_conv: implicit GEMM on tcgen05.  For output pixel (oy, ox) and filter row ky, the seven
//      taps x three channels it needs are the 8-pixel window [2ox-4, 2ox+4) of input row 2oy-3+ky:
//      32 contiguous bf16 = 64 bytes, and consecutive ox are 2 pixels = 16 bytes apart.  So the A operand of a
//      filter row -- [128 output pixels x 32] -- IS the raw row segment itself (16*128 + 48 bytes), read through an
//      UMMA descriptor WITHOUT swizzle whose core matrices OVERLAP: in the canonical K-major layout the eight rows
//      of a core matrix lie 16 bytes apart, the next core matrix along K starts `LBO` bytes later and the next
//      eight rows `SBO` bytes later; with LBO = 16 and SBO = 128 element (m, k) is read from byte 16 m + 2 k of
//      the segment, which is exactly window m.  One 2 KB bulk copy per (tile, filter row) replaces a 128-row x
//      64-byte TMA box (8 KB of shared-memory writes, 4x the L2 reads, and the TMA unit's per-row work -- the
//      first version of this kernel was bound by exactly that: 353 us against an HBM bound of 112 us, epilogue
//      warps waiting 46 % of the time, profiles/r2_graph_trace.txt).  Filter rows outside the image are skipped
//      (zero contribution).  K = 7 filter rows x 32 = 224 (weights zero at the unused window slots); B (28 KiB,
//      64-byte swizzle) stays resident in shared memory.  Two CTAs per SM: eight epilogue warps.
#include "hg_common.cuh"
#include "../../include/hg_api.h"

#include <cudaTypedefs.h>
#include <cstring>
#include <mutex>

namespace hg {
namespace stem {

constexpr int kTileM = 128;
constexpr int kCout = 64;
constexpr int kTaps = 7;                         // filter rows = k-blocks
constexpr int kKb = 32;                          // bf16 per k-block row (8 pixels x 4 channels)
constexpr int kAStage = 2176;                    // one raw row segment: 16 * 128 + 48 = 2096 bytes, rounded up to 128
constexpr int kBBlock = kCout * kKb * 2;         // 4 KiB
constexpr int kStages = 16;                        // 16 x 2176 B = 34 KiB: keeps the staging slabs behind it 1024-byte aligned
constexpr int kStaging = kTileM * 128;           // [128 x 64 ch] bf16, 128-byte swizzle
constexpr int kSmem = 1024 + kTaps * kBBlock + kStages * kAStage + 2 * kStaging + kCout * 4 + 512;

struct Params {
    const uint8_t* packed; // [n][h][w+8][4] bf16: the packed image (hg_stem_pack)
    long long row_pitch;   // bytes per packed image row
    int h;                 // input rows
    CUtensorMap map_b;     // (224, 64) weights
    CUtensorMap map_out;   // (64, n*oh*ow) output, NHWC bf16
    const float* bias;
    unsigned int* err_word;
    int oh, ow, box_w, tiles_x, num_tiles;
};

__global__ void __launch_bounds__(256) stem_pack_kernel(const float* __restrict__ in, uint2* __restrict__ out, int n,
                                                         int h, int w, int flip_w) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = static_cast<long long>(n) * h * w;
    const long long plane = static_cast<long long>(h) * w;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % w);
        const long long row = i / w;                     // n*h + y
        const long long b = row / h;
        const long long src = b * 3 * plane + (row - b * h) * w + (flip_w == 1 ? (w - 1 - x) : x);
        const float r = __ldg(in + src), g = __ldg(in + src + plane), bl = __ldg(in + src + 2 * plane);
        uint2 o;
        o.x = pack_bf16x2(r, g);
        o.y = pack_bf16x2(bl, 0.f);
        out[row * (w + 8) + 4 + x] = o;
        // flip_w == 2: both orientations from one read -- the mirrored copy goes to the second half of a [2n] buffer
        if (flip_w == 2) out[(static_cast<long long>(n) * h + row) * (w + 8) + 4 + (w - 1 - x)] = o;
    }
}

// K-major operand WITHOUT swizzle: 8-row core matrices of 16-byte rows; `lbo` = bytes to the next core matrix along K,
// `sbo` = bytes to the next 8 rows.
__device__ __forceinline__ uint64_t umma_desc_noswizzle(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) |
           (static_cast<uint64_t>(sbo >> 4) << 32) | (static_cast<uint64_t>(1) << 46);
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(256, 2) stem_conv_kernel(const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_b = smem;                                  // 7 x 4 KiB
    uint8_t* smem_a = smem_b + kTaps * kBBlock;              // 28 KiB offset: 1024-aligned
    uint8_t* smem_out = smem_a + kStages * kAStage;
    float* s_bias = reinterpret_cast<float*>(smem_out + 2 * kStaging);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + kCout);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tmem_full_bar = bars + 2 * kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint64_t* w_bar = tmem_empty_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_bar + 1);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kCout; i += blockDim.x) s_bias[i] = p.bias[i];
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&p.map_b);
        tma_prefetch_desc(&p.map_out);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 4);
        }
        mbar_init(w_bar, 1);
        fence_mbar_init();
    }
    if (warp_idx == 1) tmem_alloc(tmem_ptr_smem, 2 * kCout);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);   // broadcast: provably warp-uniform for the issue loop
    pdl_launch_dependents();
    const uint32_t a_tx = static_cast<uint32_t>(16 * p.box_w + 48);     // bytes of one raw row segment

    if (warp_idx == 0) {
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(w_bar, kTaps * kBBlock);
            for (int kb = 0; kb < kTaps; ++kb) tma_load_2d(smem_b + kb * kBBlock, &p.map_b, w_bar, kb * kKb, 0);
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x) {
                const int tx = tile % p.tiles_x;
                const int row = tile / p.tiles_x;            // n*oh + oy
                const int n = row / p.oh, oy = row - n * p.oh;
                for (int ky = 0; ky < kTaps; ++ky) {
                    const int iy = 2 * oy - 3 + ky;
                    if (iy < 0 || iy >= p.h) continue;           // a filter row above / below the image contributes nothing
                    ok = mbar_wait(&empty_bar[stage], phase ^ 1u, p.err_word, 0x2101);
                    if (!ok) break;
                    mbar_arrive_expect_tx(&full_bar[stage], a_tx);
                    bulk_load(smem_a + stage * kAStage,
                              p.packed + (static_cast<long long>(n) * p.h + iy) * p.row_pitch + static_cast<long long>(tx) * p.box_w * 16,
                              a_tx, &full_bar[stage]);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_bf16(kTileM, kCout);
            bool ok = mbar_wait(w_bar, 0, p.err_word, 0x2203);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                ok = mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1u) ^ 1u, p.err_word, 0x2201);
                if (!ok) break;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kCout);
                const int oy = (tile / p.tiles_x) % p.oh;
                uint32_t have = 0;                               // 0 until the first filter row of this tile has been issued
                for (int ky = 0; ky < kTaps; ++ky) {
                    const int iy = 2 * oy - 3 + ky;
                    if (iy < 0 || iy >= p.h) continue;
                    ok = mbar_wait(&full_bar[stage], phase, p.err_word, 0x2202);
                    if (!ok) break;
                    tc_fence_after();
                    const uint64_t a_desc = umma_desc_noswizzle(smem_u32(smem_a + stage * kAStage), 16, 128);
                    const uint64_t b_desc = umma_desc_sw64(smem_u32(smem_b + ky * kBBlock));
#pragma unroll
                    for (int k = 0; k < kKb / 16; ++k)
                        tc_mma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (have | k) != 0 ? 1u : 0u);
                    have = 1;
                    tc_commit(&empty_bar[stage]);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (ok) tc_commit(&tmem_full_bar[acc]);
            }
        }
    } else if (warp_idx >= 4) {
        pdl_wait();
        const int q = warp_idx & 3;
        const int row = q * 32 + lane;
        int it = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const int tx = tile % p.tiles_x;
            const int orow = tile / p.tiles_x;
            ok = mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1u, p.err_word, 0x2401);
            if (!ok) break;
            tc_fence_after();
            uint8_t* stg = smem_out + (it & 1) * kStaging;
            if (warp_idx == 4 && elect_one_sync()) tma_store_wait_read<1>();
            named_bar_sync(1, 128);
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * kCout);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + half * 32, v);
                tmem_ld_wait();
                if (half == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = fmaxf(__uint_as_float(v[i * 8 + j]) + s_bias[half * 32 + i * 8 + j], 0.f);
                    uint4 o;
                    o.x = pack_bf16x2(f[0], f[1]);
                    o.y = pack_bf16x2(f[2], f[3]);
                    o.z = pack_bf16x2(f[4], f[5]);
                    o.w = pack_bf16x2(f[6], f[7]);
                    *reinterpret_cast<uint4*>(stg + row * 128 + (((half * 4 + i) ^ (row & 7)) << 4)) = o;
                }
            }
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (warp_idx == 4 && elect_one_sync()) {
                tma_store_2d(&p.map_out, stg, 0, orow * p.ow + tx * p.box_w);
                tma_store_commit();
            }
        }
        if (warp_idx == 4 && elect_one_sync()) tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * kCout);
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static std::mutex mu;
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
            set_last_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
            return nullptr;
        }
        fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

}  // namespace stem
}  // namespace hg

using namespace hg;

extern "C" int hg_stem_pack(const float* in_nchw, void* packed, int32_t n, int32_t h, int32_t w, int32_t flip_w,
                            void* stream) {
    if (!in_nchw || !packed || n <= 0 || h <= 0 || w <= 0 || (reinterpret_cast<uintptr_t>(packed) & 15u)) {
        set_last_error("hg_stem_pack: bad arguments");
        return HG_ERR_INVALID;
    }
    const long long total = static_cast<long long>(n) * h * w;
    long long blocks = (total + 255) / 256;
    const long long cap = static_cast<long long>(num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    HG_CUDA_OK(launch_kernel(stem::stem_pack_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0,
                             static_cast<cudaStream_t>(stream), in_nchw, static_cast<uint2*>(packed), n, h, w, flip_w));
    return HG_OK;
}

extern "C" int hg_stem_conv(const void* packed, const void* weight, const float* bias, void* out, unsigned int* err_word,
                            int32_t n, int32_t h, int32_t w, void* stream) {
    using namespace hg::stem;
    if (!packed || !weight || !bias || !out || n <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1)) {
        set_last_error("hg_stem_conv: bad arguments (need even h, w)");
        return HG_ERR_INVALID;
    }
    if (w / 2 > kTileM && (w / 2) % kTileM != 0) {
        // a partial last tile per output row would read past the packed row and store past the output row (the output map
        // is the flat [pixels x 64] matrix): widths above 256 must be multiples of 256
        set_last_error("hg_stem_conv: input widths above 256 must be multiples of 256 (got %d)", w);
        return HG_ERR_INVALID;
    }
    auto enc = encode_fn();
    if (!enc) return HG_ERR_CUDA;
    Params kp;
    memset(&kp, 0, sizeof(kp));
    kp.bias = bias;
    kp.err_word = err_word;
    kp.oh = h / 2;
    kp.ow = w / 2;
    kp.box_w = kp.ow < kTileM ? kp.ow : kTileM;
    kp.tiles_x = (kp.ow + kp.box_w - 1) / kp.box_w;
    kp.num_tiles = n * kp.oh * kp.tiles_x;
    kp.packed = static_cast<const uint8_t*>(packed);
    kp.row_pitch = static_cast<long long>(w + 8) * 8;                  // bytes per packed image row
    kp.h = h;
    if (reinterpret_cast<uintptr_t>(packed) & 15u) {
        set_last_error("hg_stem_conv: the packed image must be 16-byte aligned");
        return HG_ERR_INVALID;
    }
    {
        cuuint64_t gdim[2] = {kTaps * kKb, kCout};
        cuuint64_t gstr[1] = {kTaps * kKb * 2};
        cuuint32_t box[2] = {kKb, kCout};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&kp.map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(weight), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_last_error("hg_stem_conv: weight tensor map failed: CUresult %d", (int)r);
            return HG_ERR_CUDA;
        }
    }
    {
        const uint64_t m = static_cast<uint64_t>(n) * kp.oh * kp.ow;
        cuuint64_t gdim[2] = {kCout, m};
        cuuint64_t gstr[1] = {kCout * 2};
        cuuint32_t box[2] = {64, static_cast<cuuint32_t>(kp.box_w)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&kp.map_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_last_error("hg_stem_conv: output tensor map failed: CUresult %d", (int)r);
            return HG_ERR_CUDA;
        }
    }
    static std::mutex mu;
    static unsigned long long done_mask = 0;
    int dev = 0;
    HG_CUDA_OK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 64 || !(done_mask >> dev & 1ull)) {
            HG_CUDA_OK(cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
            if (dev < 64) done_mask |= 1ull << dev;
        }
    }
    const int grid = kp.num_tiles < 2 * num_sms() ? kp.num_tiles : 2 * num_sms();     // two CTAs per SM
    HG_CUDA_OK(launch_kernel(stem_conv_kernel, dim3(grid), dim3(256), kSmem, static_cast<cudaStream_t>(stream), kp));
    return HG_OK;
}
