// 1x1 convolution (pixel-major GEMM) for NHWC bf16 activations -- the HBM-bound half of the hourglass
// (conv1 / conv3 of every bottleneck, fc, merged remap).  Second-generation kernel, specialised for
// streaming: every byte of A, residual and output crosses HBM exactly once and nothing else does.
//
//   * weights [BLOCK_N x K] are loaded ONCE per persistent CTA and stay resident in shared memory
//     (<= 128 KiB); the main-loop ring carries only 16 KiB A slabs, so it is 4-8 stages deep.
//   * the residual tile (and the quarter-resolution tensor of the fused nearest-upsample add) are
//     prefetched by TMA into the SAME swizzled staging slabs the epilogue later overwrites with the
//     result and TMA-stores: no per-thread global loads anywhere in the epilogue.
//   * pre-activation bn1+ReLU prologue with a conflict-free mapping (a quarter-warp owns one
//     128-byte row; per-thread channel chunk is fixed so scale/shift live in registers per k-block).
//
//   * optional second output: the 2x2 max-pool of the result (the hourglass pools every level's input,
//     src/models/modules.py:82): a 128-pixel tile holds whole pooling windows, so the epilogue takes the maximum of
//     four rows of the bf16 slab it has just staged and TMA-stores a 32-row slab of the pooled tensor -- the pool
//     kernel's full re-read of the 256-channel tensor disappears.
//
// Warp roles: 0 = A producer (TMA), 1 = MMA issuer + TMEM owner, 2 = staging-ring producer (TMA),
//             3 = idle, 4..7 = epilogue, 8..11 = prologue (optional).
#include "hg_common.cuh"
#include "../../include/hg_api.h"

#include <cudaTypedefs.h>
#include <cstring>
#include <mutex>

namespace hg {

namespace c1 {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;
constexpr int kSlabBytes = kTileM * kBlockK * 2;     // 16 KiB: one [128 x 64] bf16 slab (A stage or staging slab)
constexpr int kUpRows = 32;                          // quarter-resolution pixels under one 128-pixel tile
constexpr int kUpBytes = kUpRows * kBlockK * 2;      // 4 KiB
constexpr int kMaxAStages = 8;
constexpr int kMaxRing = 6;
constexpr int kMaxK = 512;
constexpr int kSmemLimit = 232448;
#ifndef HG_C1_FFMA2
#define HG_C1_FFMA2 1
#endif

struct Params {
    CUtensorMap map_a;      // (c, m) over `in`
    CUtensorMap map_a2;     // (c, m) over `in2`
    CUtensorMap map_b;      // (k, cout) over weights
    CUtensorMap map_res;    // (c, m) over residual
    CUtensorMap map_up;     // (c, m/4) over the quarter-resolution tensor
    CUtensorMap map_out;    // (c, m) over out
    CUtensorMap map_pool;   // (c, m/4) over the max-pooled output (kPool)
    const float* bias;
    const float* in_scale;
    const float* in_shift;
    unsigned int* err_word;
    float* stats;           // optional fp32 [2*cout]: per-channel sum | sum of squares of the epilogue's fp32 results, ADDED
    int m_total;            // valid rows (the last tile may be ragged)
    int num_tiles;
    int kb1, kb2;           // k-blocks from `in` / from `in2`
    int cin;
    int a_stages, ring;     // pipeline depths chosen by the host from the smem budget
    int has_res, has_up, relu;
    int has_pool;           // image width when the 2x2 max-pool of the result is written too (kPool), else 0
    int out_halo, img_h, img_w;     // out_halo: map_out is the 4-D strided view (c, x, y, n) of a halo-padded buffer
    __nv_bfloat16* out_raw;         // kRagged: the halo-padded buffer itself (rows are stored one by one, no tensor map)
    // Row tiling (tiles_per_img != 0): a tile is `tile_pixels` = R whole image rows (R*w <= 128) of ONE image instead of 128
    // consecutive pixels, so that the geometric epilogues (halo / 4-D store, upsample operand) work for any width <= 128;
    // the rest of the 128-row MMA tile computes on stale shared memory and is clipped by the 4-D TMA store.
    int tile_pixels, tiles_per_img, img_hw;
    // kPoolIn: the prologue warps also write the 2x2 max-pool of the RAW input tile (hg_conv_desc.pool_in)
    __nv_bfloat16* pool_in;
    int pool_in_w, pool_in_shift;   // image width; log2 of the pooled width
};

__device__ __forceinline__ int tile_m0(const Params& p, int tile) {
    if (p.tiles_per_img == 0) return tile * kTileM;
    const int img = tile / p.tiles_per_img;
    return img * p.img_hw + (tile - img * p.tiles_per_img) * p.tile_pixels;
}

enum : uint32_t { kErrProducer = 0x1100, kErrMma = 0x1200, kErrRing = 0x1300, kErrEpilogue = 0x1400, kErrPrologue = 0x1500 };

template <int BLOCK_N>
__host__ __device__ constexpr int w_bytes(int num_kb) { return num_kb * BLOCK_N * kBlockK * 2; }

// kStats: the epilogue also adds the per-channel sum / sum of squares of its results into p.stats.  A separate
// instantiation so that the plain kernel (inference, dgrad) keeps its register budget and schedule.
__device__ __forceinline__ uint32_t bf16x2_max_u32(uint32_t a, uint32_t b) {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}

// kPool: the epilogue also writes the 2x2 max-pool of its result through p.map_pool.
// kRagged: halo-padded output for image sizes whose 128-pixel tiles are not whole image rows (e.g. the 64x48 heat-map
// grid of 256x192 inputs): the 4-D strided TMA store needs a rectangular box, so each epilogue thread stores its own
// pixel's channels at the pixel's halo position instead (like the 3x3 kernel's epilogue); pads are never written.
// kPoolIn: see Params::pool_in (prologue kernels only).
template <int BLOCK_N, bool kPrologue, bool kStats, bool kPool, bool kRagged = false, bool kPoolIn = false>
__global__ void __launch_bounds__(kPrologue ? 384 : 256, 1) conv1x1_kernel(const __grid_constant__ Params p) {
    constexpr int kBStage = BLOCK_N * kBlockK * 2;
    constexpr int kSlabs = BLOCK_N / 64;
    constexpr int kTmemCols = 2 * BLOCK_N;           // 128 / 256 / 512

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int num_kb = p.kb1 + p.kb2;
    uint8_t* smem_w = smem;                                        // num_kb x kBStage
    uint8_t* smem_a = smem_w + num_kb * kBStage;                   // a_stages x 16 KiB
    uint8_t* smem_ring = smem_a + p.a_stages * kSlabBytes;         // ring x 16 KiB
    uint8_t* smem_up = smem_ring + p.ring * kSlabBytes;            // ring x 4 KiB (only if has_up)
    uint8_t* smem_pool = smem_up + (p.has_up ? p.ring * kUpBytes : 0);     // 2 x 4 KiB pooled staging slabs (kPool)
    float* s_bias = reinterpret_cast<float*>(smem_pool + (kPool ? 2 * kUpBytes : 0));
    float* s_stats = s_bias + BLOCK_N;                             // [2*BLOCK_N] column sums of this CTA's tiles
    float* s_scale = s_stats + (kStats ? 2 * BLOCK_N : 0);
    float* s_shift = s_scale + (kPrologue ? kMaxK : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + (kPrologue ? kMaxK : 0));
    uint64_t* full_bar = bars;                                     // [kMaxAStages]
    uint64_t* empty_bar = full_bar + kMaxAStages;
    uint64_t* ready_bar = empty_bar + kMaxAStages;
    uint64_t* res_full_bar = ready_bar + kMaxAStages;              // [kMaxRing]
    uint64_t* ring_empty_bar = res_full_bar + kMaxRing;
    uint64_t* tmem_full_bar = ring_empty_bar + kMaxRing;           // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint64_t* w_bar = tmem_empty_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_bar + 1);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < BLOCK_N; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    if (kStats)
        for (int i = threadIdx.x; i < 2 * BLOCK_N; i += blockDim.x) s_stats[i] = 0.f;
    if (kPrologue) {
        for (int i = threadIdx.x; i < p.cin; i += blockDim.x) {
            s_scale[i] = p.in_scale[i];
            s_shift[i] = p.in_shift[i];
        }
    }
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&p.map_a);
        tma_prefetch_desc(&p.map_b);
        if (!kRagged) tma_prefetch_desc(&p.map_out);
        if (p.kb2) tma_prefetch_desc(&p.map_a2);
        if (p.has_res) tma_prefetch_desc(&p.map_res);
        if (p.has_up) tma_prefetch_desc(&p.map_up);
        if (kPool) tma_prefetch_desc(&p.map_pool);
        for (int s = 0; s < kMaxAStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&ready_bar[s], 4);
        }
        for (int s = 0; s < kMaxRing; ++s) {
            mbar_init(&res_full_bar[s], 1);
            mbar_init(&ring_empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 4);
        }
        mbar_init(w_bar, 1);
        fence_mbar_init();
    }
    if (warp_idx == 1) tmem_alloc(tmem_ptr_smem, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);   // broadcast: provably warp-uniform for the issue loop
    pdl_launch_dependents();     // the next kernel in the stream may begin its own setup

    if (warp_idx == 0) {
        // ===================== A producer: resident weights once, then the A slab stream =====================
        if (elect_one_sync()) {
            // weights / bias / bn vectors are constants: fetched BEFORE waiting on the predecessor kernel
            mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(num_kb * kBStage));
            for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(smem_w + kb * kBStage, &p.map_b, w_bar, kb * kBlockK, 0);
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x) {
                const int m0 = tile_m0(p, tile);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ok = mbar_wait(&empty_bar[stage], phase ^ 1u, p.err_word, kErrProducer | 1);
                    if (!ok) break;
                    mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.tile_pixels * 128));
                    if (kb < p.kb1)
                        tma_load_2d(smem_a + stage * kSlabBytes, &p.map_a, &full_bar[stage], kb * kBlockK, m0);
                    else
                        tma_load_2d(smem_a + stage * kSlabBytes, &p.map_a2, &full_bar[stage], (kb - p.kb1) * kBlockK, m0);
                    if (++stage == p.a_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== MMA issuer =====================
        if (elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BLOCK_N);
            bool ok = mbar_wait(w_bar, 0, p.err_word, kErrMma | 3);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1u;
                ok = mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u, p.err_word, kErrMma | 1);
                if (!ok) break;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
                for (int kb = 0; kb < num_kb; ++kb) {
                    ok = mbar_wait(kPrologue ? &ready_bar[stage] : &full_bar[stage], phase, p.err_word, kErrMma | 2);
                    if (!ok) break;
                    tc_fence_after();
                    const uint64_t a_desc = umma_desc_sw128(smem_u32(smem_a + stage * kSlabBytes));
                    const uint64_t b_desc = umma_desc_sw128(smem_u32(smem_w + kb * kBStage));
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)
                        tc_mma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(&empty_bar[stage]);
                    if (kb == num_kb - 1) tc_commit(&tmem_full_bar[acc]);
                    if (++stage == p.a_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp_idx == 2) {
        // ===================== staging-ring producer: residual / upsample operands by TMA =====================
        if (elect_one_sync()) {
            pdl_wait();
            int buf = 0;
            uint32_t phase = 0;
            bool ok = true;
            const uint32_t tx = static_cast<uint32_t>((p.has_res ? p.tile_pixels * 128 : 0) + (p.has_up ? p.tile_pixels * 32 : 0));
            for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x) {
                const int m0 = tile_m0(p, tile);
                for (int slab = 0; slab < kSlabs; ++slab) {
                    ok = mbar_wait(&ring_empty_bar[buf], phase ^ 1u, p.err_word, kErrRing | 1);
                    if (!ok) break;
                    if (tx != 0) {
                        mbar_arrive_expect_tx(&res_full_bar[buf], tx);
                        if (p.has_res)
                            tma_load_2d(smem_ring + buf * kSlabBytes, &p.map_res, &res_full_bar[buf], slab * 64, m0);
                        if (p.has_up)
                            tma_load_2d(smem_up + buf * kUpBytes, &p.map_up, &res_full_bar[buf], slab * 64, m0 >> 2);
                    } else {
                        mbar_arrive(&res_full_bar[buf]);       // nothing to fetch: the slab is simply free
                    }
                    if (++buf == p.ring) {
                        buf = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp_idx >= 4 && warp_idx < 8) {
        // ===================== epilogue =====================
        pdl_wait();
        const int q = warp_idx & 3;
        const int row = q * 32 + lane;
        // the slab stores are issued by ONE thread of warp 4, always the same one (elect.sync is deterministic for a given
        // mask), because bulk async-groups are tracked per thread
        // quarter-resolution row under this pixel: tiles are 128-aligned runs of whole 2x2 blocks, so the
        // 32 low-res pixels of a tile are contiguous; row r=(y,x) within the tile maps to (y/2, x/2).
        // With W = 2^k <= 64:  r = yy*W + x  ->  low = (yy/2)*(W/2) + x/2
        int low_row = 0;
        if (p.has_up) {
            const int W = p.has_up;            // has_up carries the image width
            const int yy = row / W, x = row - yy * W;
            low_row = (yy >> 1) * (W >> 1) + (x >> 1);
        }
        int it = 0;
        int buf = 0;
        uint32_t ring_phase = 0;
        int prev_buf = -1;
        int pool_it = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1u;
            const int m0 = tile_m0(p, tile);
            ok = mbar_wait(&tmem_full_bar[acc], acc_phase, p.err_word, kErrEpilogue | 1);
            if (!ok) break;
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
            __nv_bfloat16* ragged_row = nullptr;       // kRagged: where this thread's pixel lives in the halo-padded buffer
            if (kRagged) {
                const int m = m0 + row;
                if (m < p.m_total) {
                    const int hw = p.img_h * p.img_w;
                    const int n_img = m / hw, r = m - n_img * hw;
                    const int y = r / p.img_w, x = r - y * p.img_w;
                    const long long pos = static_cast<long long>(p.img_w + 1) * (1 + static_cast<long long>(n_img) * (p.img_h + 1) + y) + x;
                    ragged_row = p.out_raw + pos * BLOCK_N;
                }
            }
            for (int slab = 0; slab < kSlabs; ++slab) {
                ok = mbar_wait(&res_full_bar[buf], ring_phase, p.err_word, kErrEpilogue | 2);
                if (!ok) break;
                const uint32_t stg = smem_u32(smem_ring + buf * kSlabBytes + row * 128);
                const uint32_t upb = smem_u32(smem_up + buf * kUpBytes + low_row * 128);
                const uint32_t bias_s = smem_u32(s_bias);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int col0 = slab * 64 + half * 32;
                    uint32_t v[32];
                    tmem_ld_32x32(t_row + col0, v);
                    uint4 r_res[4], r_up[4];
                    if (p.has_res) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) r_res[i] = lds128(stg + (((half * 4 + i) ^ (row & 7)) << 4));
                    }
                    if (p.has_up) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) r_up[i] = lds128(upb + (((half * 4 + i) ^ (low_row & 7)) << 4));
                    }
                    float f[32];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {       // warp-uniform address: one broadcast wavefront per 4 channels
                        const float4 b4 = lds128f(bias_s + (col0 + i * 4) * 4);
                        f[i * 4 + 0] = b4.x;
                        f[i * 4 + 1] = b4.y;
                        f[i * 4 + 2] = b4.z;
                        f[i * 4 + 3] = b4.w;
                    }
                    tmem_ld_wait();
                    if (slab == kSlabs - 1 && half == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] += __uint_as_float(v[i]);
                    if (p.has_res) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t w4[4] = {r_res[i].x, r_res[i].y, r_res[i].z, r_res[i].w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                f[i * 8 + 2 * j] += bf16_lo_to_f32(w4[j]);
                                f[i * 8 + 2 * j + 1] += bf16_hi_to_f32(w4[j]);
                            }
                        }
                    }
                    if (p.has_up) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t w4[4] = {r_up[i].x, r_up[i].y, r_up[i].z, r_up[i].w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                f[i * 8 + 2 * j] += bf16_lo_to_f32(w4[j]);
                                f[i * 8 + 2 * j + 1] += bf16_hi_to_f32(w4[j]);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 o;
                        if (p.relu) {
                            o.x = pack_bf16x2_relu(f[i * 8 + 0], f[i * 8 + 1]);
                            o.y = pack_bf16x2_relu(f[i * 8 + 2], f[i * 8 + 3]);
                            o.z = pack_bf16x2_relu(f[i * 8 + 4], f[i * 8 + 5]);
                            o.w = pack_bf16x2_relu(f[i * 8 + 6], f[i * 8 + 7]);
                        } else {
                            o.x = pack_bf16x2(f[i * 8 + 0], f[i * 8 + 1]);
                            o.y = pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]);
                            o.z = pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]);
                            o.w = pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]);
                        }
                        if (kRagged) {
                            if (ragged_row != nullptr) *reinterpret_cast<uint4*>(ragged_row + col0 + i * 8) = o;
                        } else {
                            sts128(stg + (((half * 4 + i) ^ (row & 7)) << 4), o);
                        }
                    }
                    if (kStats) {
                        // per-channel sum / sum of squares of what this tile produced (train-mode BatchNorm statistics
                        // of the NEXT layer, fused here so that no separate pass re-reads the tensor)
                        const bool row_ok = m0 + row < p.m_total;
                        float sq[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float t = p.relu ? fmaxf(f[i], 0.f) : f[i];
                            t = row_ok ? t : 0.f;
                            f[i] = t;
                            sq[i] = t * t;
                        }
                        const float cs = warp_column_sum(f, lane);
                        const float cq = warp_column_sum(sq, lane);
                        atomicAdd(&s_stats[col0 + lane], cs);
                        atomicAdd(&s_stats[BLOCK_N + col0 + lane], cq);
                    }
                }
                fence_proxy_async_smem();
                named_bar_sync(1, 128);
                if (kPool) {
                    // 2x2 max-pool of the staged slab: 32 pooled pixels x 8 sixteen-byte chunks = 2 items per thread
                    const uint32_t slab_s = smem_u32(smem_ring + buf * kSlabBytes);
                    const uint32_t pool_s = smem_u32(smem_pool + (pool_it & 1) * kUpBytes);
                    const int W = p.has_pool, Wh = W >> 1;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int item = static_cast<int>(threadIdx.x) - 128 + 128 * k;
                        const int prow = item >> 3, chunk = item & 7;
                        const int py = prow / Wh, px = prow - py * Wh;
                        const int r0 = 2 * py * W + 2 * px;
                        const int rr[4] = {r0, r0 + 1, r0 + W, r0 + W + 1};
                        uint4 m = lds128(slab_s + rr[0] * 128 + ((chunk ^ (rr[0] & 7)) << 4));
#pragma unroll
                        for (int j = 1; j < 4; ++j) {
                            const uint4 v4 = lds128(slab_s + rr[j] * 128 + ((chunk ^ (rr[j] & 7)) << 4));
                            m.x = bf16x2_max_u32(m.x, v4.x);
                            m.y = bf16x2_max_u32(m.y, v4.y);
                            m.z = bf16x2_max_u32(m.z, v4.z);
                            m.w = bf16x2_max_u32(m.w, v4.w);
                        }
                        sts128(pool_s + prow * 128 + ((chunk ^ (prow & 7)) << 4), m);
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1, 128);
                }
                if (kRagged) {
                    // nothing staged: the slab (at most a residual operand, read above) is free as soon as all four warps are here
                    if (warp_idx == 4 && elect_one_sync()) mbar_arrive(&ring_empty_bar[buf]);
                } else if (warp_idx == 4 && elect_one_sync()) {
                    if (kPool) {
                        // same bulk group as the slab's own store: the read-completion wait below covers both
                        tma_store_2d(&p.map_pool, smem_pool + (pool_it & 1) * kUpBytes, slab * 64, m0 >> 2);
                    }
                    if (p.out_halo) {
                        const int hw = p.img_h * p.img_w;
                        const int n_img = m0 / hw;
                        const int y0 = (m0 - n_img * hw) / p.img_w;
                        tma_store_4d(&p.map_out, smem_ring + buf * kSlabBytes, slab * 64, 0, y0, n_img);
                    } else {
                        tma_store_2d(&p.map_out, smem_ring + buf * kSlabBytes, slab * 64, m0);
                    }
                    tma_store_commit();
                    if (prev_buf >= 0) {
                        tma_store_wait_read<1>();               // the previous slab's store has drained its smem
                        mbar_arrive(&ring_empty_bar[prev_buf]);
                    }
                }
                if (!kRagged) prev_buf = buf;
                if (kPool) ++pool_it;
                if (++buf == p.ring) {
                    buf = 0;
                    ring_phase ^= 1u;
                }
            }
        }
        if (kStats) {
            named_bar_sync(1, 128);
            for (int i = threadIdx.x - 128; i < 2 * BLOCK_N; i += 128) atomicAdd(p.stats + i, s_stats[i]);
        }
        if (warp_idx == 4 && elect_one_sync()) tma_store_wait<0>();
    } else if (kPrologue && warp_idx >= 8) {
        // ===================== prologue: a = relu(a*scale + shift), in place =====================
        const int w = warp_idx - 8;
        const int c = lane & 7;                 // logical 16-byte chunk = 8 channels, fixed per thread
        const int rsub = lane >> 3;
        int stage = 0;
        uint32_t phase = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x) {
            for (int kb = 0; kb < num_kb; ++kb) {
                ok = mbar_wait(&full_bar[stage], phase, p.err_word, kErrPrologue | 1);
                if (!ok) break;
                if (kb < p.kb1) {
                    const uint32_t sc = smem_u32(s_scale + kb * kBlockK + c * 8);
                    const uint32_t sh = smem_u32(s_shift + kb * kBlockK + c * 8);
                    const float4 s0 = lds128f(sc), s1 = lds128f(sc + 16);
                    const float4 h0 = lds128f(sh), h1 = lds128f(sh + 16);
#if HG_C1_FFMA2
                    const unsigned long long sc01 = f2_pack(s0.x, s0.y), sc23 = f2_pack(s0.z, s0.w), sc45 = f2_pack(s1.x, s1.y),
                                             sc67 = f2_pack(s1.z, s1.w);
                    const unsigned long long sh01 = f2_pack(h0.x, h0.y), sh23 = f2_pack(h0.z, h0.w), sh45 = f2_pack(h1.x, h1.y),
                                             sh67 = f2_pack(h1.z, h1.w);
#endif
                    const uint32_t a_base = smem_u32(smem_a + stage * kSlabBytes);
                    uint32_t addr[8];
                    uint4 d[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {                        // all eight loads in flight before any use
                        int row = i * 16 + w * 4 + rsub;                 // a quarter-warp owns one 128-byte row
                        if (kPoolIn) {
                            // the eight rows of a thread are two whole 2x2 pooling windows: pooled pixel q of the tile's 32,
                            // window element e = (dy, dx)
                            const int q = (w * 4 + rsub) * 2 + (i >> 2), e = i & 3;
                            const int qy = q >> p.pool_in_shift, qx = q - (qy << p.pool_in_shift);
                            row = (2 * qy + (e >> 1)) * p.pool_in_w + 2 * qx + (e & 1);
                        }
                        addr[i] = a_base + row * 128 + ((c ^ (row & 7)) << 4);
                        d[i] = lds128(addr[i]);
                    }
                    if (kPoolIn) {
                        // max over each window of the RAW values (what F.max_pool2d sees), 8 channels per thread: a
                        // quarter-warp writes 128 contiguous bytes of the pooled pixel's row
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            uint4 m;
                            m.x = bf16x2_max_u32(bf16x2_max_u32(d[h2 * 4].x, d[h2 * 4 + 1].x), bf16x2_max_u32(d[h2 * 4 + 2].x, d[h2 * 4 + 3].x));
                            m.y = bf16x2_max_u32(bf16x2_max_u32(d[h2 * 4].y, d[h2 * 4 + 1].y), bf16x2_max_u32(d[h2 * 4 + 2].y, d[h2 * 4 + 3].y));
                            m.z = bf16x2_max_u32(bf16x2_max_u32(d[h2 * 4].z, d[h2 * 4 + 1].z), bf16x2_max_u32(d[h2 * 4 + 2].z, d[h2 * 4 + 3].z));
                            m.w = bf16x2_max_u32(bf16x2_max_u32(d[h2 * 4].w, d[h2 * 4 + 1].w), bf16x2_max_u32(d[h2 * 4 + 2].w, d[h2 * 4 + 3].w));
                            const long long q = static_cast<long long>(tile) * 32 + (w * 4 + rsub) * 2 + h2;
                            stg_v4(p.pool_in + q * p.cin + kb * kBlockK + c * 8, m);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        uint4 o;
#if HG_C1_FFMA2
                        // packed fp32 pairs (FFMA2): bit-identical to two scalar fmaf, half the arithmetic instructions
                        const unsigned long long r0 = f2_fma(f2_pack(bf16_lo_to_f32(d[i].x), bf16_hi_to_f32(d[i].x)), sc01, sh01);
                        const unsigned long long r1 = f2_fma(f2_pack(bf16_lo_to_f32(d[i].y), bf16_hi_to_f32(d[i].y)), sc23, sh23);
                        const unsigned long long r2 = f2_fma(f2_pack(bf16_lo_to_f32(d[i].z), bf16_hi_to_f32(d[i].z)), sc45, sh45);
                        const unsigned long long r3 = f2_fma(f2_pack(bf16_lo_to_f32(d[i].w), bf16_hi_to_f32(d[i].w)), sc67, sh67);
                        o.x = pack_bf16x2_relu(f2_lo(r0), f2_hi(r0));
                        o.y = pack_bf16x2_relu(f2_lo(r1), f2_hi(r1));
                        o.z = pack_bf16x2_relu(f2_lo(r2), f2_hi(r2));
                        o.w = pack_bf16x2_relu(f2_lo(r3), f2_hi(r3));
#else
                        o.x = pack_bf16x2_relu(fmaf(bf16_lo_to_f32(d[i].x), s0.x, h0.x), fmaf(bf16_hi_to_f32(d[i].x), s0.y, h0.y));
                        o.y = pack_bf16x2_relu(fmaf(bf16_lo_to_f32(d[i].y), s0.z, h0.z), fmaf(bf16_hi_to_f32(d[i].y), s0.w, h0.w));
                        o.z = pack_bf16x2_relu(fmaf(bf16_lo_to_f32(d[i].z), s1.x, h1.x), fmaf(bf16_hi_to_f32(d[i].z), s1.y, h1.y));
                        o.w = pack_bf16x2_relu(fmaf(bf16_lo_to_f32(d[i].w), s1.z, h1.z), fmaf(bf16_hi_to_f32(d[i].w), s1.w, h1.w));
#endif
                        sts128(addr[i], o);
                    }
                    fence_proxy_async_smem();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready_bar[stage]);
                if (++stage == p.a_stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
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

// row-major bf16 matrix [rows][cols] viewed as (cols, rows); box (64, box_rows), 128-byte swizzle
static int make_map(CUtensorMap* map, const void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows) {
    auto enc = encode_fn();
    if (!enc) return HG_ERR_CUDA;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (cols %llu rows %llu box_rows %u)", (int)r,
                       (unsigned long long)cols, (unsigned long long)rows, box_rows);
        return HG_ERR_CUDA;
    }
    return HG_OK;
}

template <int BLOCK_N, bool kPrologue, bool kStats, bool kPool = false, bool kRagged = false, bool kPoolIn = false>
static int launch_variant(const Params& kp, int smem_bytes, cudaStream_t stream) {
    auto kern = conv1x1_kernel<BLOCK_N, kPrologue, kStats, kPool, kRagged, kPoolIn>;
    static std::mutex mu;
    static unsigned long long done_mask = 0;
    int dev = 0;
    HG_CUDA_OK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 64 || !(done_mask >> dev & 1ull)) {
            HG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
            if (dev < 64) done_mask |= 1ull << dev;
        }
    }
    const int grid = kp.num_tiles < num_sms() ? kp.num_tiles : num_sms();
    HG_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(kPrologue ? 384 : 256), smem_bytes, stream, kp));
    return HG_OK;
}

template <int BLOCK_N, bool kPrologue>
static int launch(const Params& kp, int smem_bytes, cudaStream_t stream) {
    return kp.stats != nullptr ? launch_variant<BLOCK_N, kPrologue, true>(kp, smem_bytes, stream)
                               : launch_variant<BLOCK_N, kPrologue, false>(kp, smem_bytes, stream);
}

}  // namespace c1

static bool halo_rectangular(const hg_conv_desc* d) {
    return d->w <= 128 && 128 % d->w == 0 && (static_cast<long long>(d->h) * d->w) % 128 == 0;
}

// the TMA-fed upsample operand of a flat tile (128 consecutive pixels) needs whole 2x2 blocks inside every tile
static bool flat_upsample_ok(const hg_conv_desc* d) {
    const int w = d->w, h = d->h;
    const bool pow2 = (w & (w - 1)) == 0;
    return pow2 && w <= 64 && w >= 2 && !(h & 1) && 128 % (2 * w) == 0;
}

// Image rows per tile when tiles are whole rows of one image (Params::tiles_per_img), 0 if the size does not allow it.
static int row_tile_rows(const hg_conv_desc* d) {
    static const bool off = getenv("HG_CONV1X1_NO_ROWTILE") != nullptr;
    if (off || d->w > 128 || d->w < 1) return 0;
    int r = 128 / d->w;
    if (r > d->h) r = d->h;
    if (d->up_low != nullptr) {
        if ((d->h | d->w) & 1) return 0;
        r &= ~1;
    }
    return r;
}

enum { kGeomFlat = 0, kGeomRows = 1, kGeomRagged = 2, kGeomNone = -1 };

// How the tile -> pixel mapping is chosen for the epilogues that depend on image geometry (halo output, upsample operand).
static int geometry_mode(const hg_conv_desc* d) {
    const bool flat_ok = (!d->out_halo || halo_rectangular(d)) && (d->up_low == nullptr || flat_upsample_ok(d));
    if (flat_ok) return kGeomFlat;
    if (row_tile_rows(d) > 0 && d->stats == nullptr && d->pool_out == nullptr) return kGeomRows;
    // per-pixel stores of the halo layout: the 64- and 128-channel producers of the 3x3 kernel's input (w <= 253 is that
    // kernel's own limit)
    if (d->out_halo && d->up_low == nullptr && d->cout != 256 && d->stats == nullptr && d->pool_out == nullptr && d->w <= 253)
        return kGeomRagged;
    return kGeomNone;
}

// Returns 1 if this kernel can run the descriptor, 0 if the generic kernel must be used.
int conv1x1_supported(const hg_conv_desc* d) {
    if (d->ksize != 1 || d->out_nchw_f32 != nullptr || d->out == nullptr) return 0;
    if (d->cout != 64 && d->cout != 128 && d->cout != 256) return 0;
    const int k = d->cin + d->cin2;
    if (k > c1::kMaxK || static_cast<long long>(k) * d->cout * 2 > 128 * 1024) return 0;
    if (d->in_scale != nullptr && d->cout == 256) return 0;
    if (geometry_mode(d) == kGeomNone) return 0;
    if (d->pool_in != nullptr) {
        // the pooled INPUT as a second output of the prologue warps: flat 128-pixel tiles made of whole 2x2 windows
        const int w = d->w, h = d->h;
        const bool pow2 = (w & (w - 1)) == 0;
        if (d->in_scale == nullptr || d->cout != 128 || d->cin2 != 0 || d->stats != nullptr || d->pool_out != nullptr) return 0;
        if (geometry_mode(d) != kGeomFlat) return 0;
        if (!pow2 || w > 64 || w < 2 || (h & 1) || 128 % (2 * w) != 0 || (static_cast<long long>(d->n) * h * w) % 128 != 0) return 0;
    }
    if (d->pool_out != nullptr) {
        // the fused max-pool output: 256-channel results without prologue / statistics, whole 2x2 windows per tile
        const int w = d->w, h = d->h;
        const bool pow2 = (w & (w - 1)) == 0;
        if (d->cout != 256 || d->in_scale != nullptr || d->stats != nullptr || d->out_halo) return 0;
        // K <= 128 only: with K = 256 the resident weights (128 KiB) leave no room for the two pooled staging slabs without
        // shrinking the staging ring, and the kernel loses more than the pool pass costs (measured: 481 vs 284 + 110 us)
        if (d->cin + d->cin2 > 128) return 0;
        if (!pow2 || w > 64 || w < 2 || (h & 1) || 128 % (2 * w) != 0) return 0;
    }
    return 1;
}

int conv1x1_launch(const hg_conv_desc* d, cudaStream_t stream) {
    using namespace c1;
    Params kp;
    memset(&kp, 0, sizeof(kp));
    const long long m = static_cast<long long>(d->n) * d->h * d->w;
    if (m > 0x7fffffffLL) {
        set_last_error("hg_conv_nhwc_bf16: too many pixels");
        return HG_ERR_INVALID;
    }
    const bool prologue = d->in_scale != nullptr;
    kp.bias = d->bias;
    kp.in_scale = d->in_scale;
    kp.in_shift = d->in_shift;
    kp.err_word = d->err_word;
    kp.stats = d->stats;
    kp.m_total = static_cast<int>(m);
    kp.num_tiles = static_cast<int>((m + kTileM - 1) / kTileM);
    const int geom = geometry_mode(d);
    const int tile_rows = geom == kGeomRows ? row_tile_rows(d) : 0;
    kp.tile_pixels = kTileM;
    kp.img_hw = d->h * d->w;
    if (geom == kGeomRows) {
        kp.tile_pixels = tile_rows * d->w;
        kp.tiles_per_img = (d->h + tile_rows - 1) / tile_rows;
        kp.num_tiles = d->n * kp.tiles_per_img;
    }
    const uint32_t box_m = static_cast<uint32_t>(kp.tile_pixels);
    kp.kb1 = d->cin / 64;
    kp.kb2 = d->cin2 / 64;
    kp.cin = d->cin;
    kp.has_res = d->residual != nullptr;
    kp.has_up = d->up_low != nullptr ? d->w : 0;
    kp.has_pool = d->pool_out != nullptr ? d->w : 0;
    kp.relu = d->relu;

    // shared-memory budget -> pipeline depths
    const int num_kb = kp.kb1 + kp.kb2;
    const int wbytes = num_kb * d->cout * kBlockK * 2;
    const int misc = (d->stats ? 3 : 1) * d->cout * 4 + (prologue ? 2 * kMaxK * 4 : 0) + 512 +   // bias, column sums, scale/shift, barriers
                     (d->pool_out ? 2 * kUpBytes : 0);                                          // pooled staging slabs
    int avail = kSmemLimit - 1024 - wbytes - misc;
    const int ring_unit = kSlabBytes + (kp.has_up ? kUpBytes : 0);
    const int slabs = d->cout / 64;
    // the staging ring wants >= one tile of slabs when it also prefetches a residual; the A ring gets the rest
    int ring = kp.has_res ? (slabs < 2 ? 2 : (slabs < kMaxRing ? slabs : kMaxRing)) : 2;
    int a_stages = (avail - ring * ring_unit) / kSlabBytes;
    while (a_stages < 3 && ring > 2) {
        --ring;
        a_stages = (avail - ring * ring_unit) / kSlabBytes;
    }
    if (a_stages > kMaxAStages) a_stages = kMaxAStages;
    if (a_stages < 2) {
        set_last_error("conv1x1: shared-memory budget exhausted (K=%d N=%d)", d->cin + d->cin2, d->cout);
        return HG_ERR_INVALID;
    }
    kp.a_stages = a_stages;
    kp.ring = ring;
    const int smem_bytes = 1024 + wbytes + a_stages * kSlabBytes + ring * ring_unit + misc;

    int rc;
    if ((rc = make_map(&kp.map_a, d->in, d->cin, m, box_m)) != HG_OK) return rc;
    if (d->in2 && (rc = make_map(&kp.map_a2, d->in2, d->cin2, m, box_m)) != HG_OK) return rc;
    if ((rc = make_map(&kp.map_b, d->weight, d->cin + d->cin2, d->cout, d->cout)) != HG_OK) return rc;
    kp.out_halo = d->out_halo;
    kp.img_h = d->h;
    kp.img_w = d->w;
    const bool ragged = geom == kGeomRagged;
    kp.out_raw = static_cast<__nv_bfloat16*>(d->out);
    if (ragged) {
        // no output tensor map: see kRagged
    } else if (d->out_halo || geom == kGeomRows) {
        // 4-D view (c, x, y, n): the interior of the halo-padded [zero row][n][h+1][w+1][c], or (row tiling) the dense NHWC
        // tensor -- there the 4-D box is what clips the last, partial tile of every image
        auto enc = encode_fn();
        if (!enc) return HG_ERR_CUDA;
        const uint64_t C = d->cout, P = d->out_halo ? d->w + 1 : d->w, HP = d->out_halo ? d->h + 1 : d->h;
        const uint32_t box_rows = geom == kGeomRows ? static_cast<uint32_t>(tile_rows) : static_cast<uint32_t>(kTileM / d->w);
        cuuint64_t gdim[4] = {C, static_cast<cuuint64_t>(d->w), static_cast<cuuint64_t>(d->h), static_cast<cuuint64_t>(d->n)};
        cuuint64_t gstr[3] = {C * 2, P * C * 2, HP * P * C * 2};
        cuuint32_t box[4] = {64, static_cast<cuuint32_t>(d->w), box_rows, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        void* base = static_cast<char*>(d->out) + (d->out_halo ? P * C * 2 : 0);        // skip the leading zero row
        kp.out_halo = 1;                                                               // = "store through the 4-D map"
        CUresult r = enc(&kp.map_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_last_error("conv1x1: halo-padded output tensor map failed: CUresult %d", (int)r);
            return HG_ERR_CUDA;
        }
    } else if ((rc = make_map(&kp.map_out, d->out, d->cout, m, kTileM)) != HG_OK) {
        return rc;
    }
    if (d->residual && (rc = make_map(&kp.map_res, d->residual, d->cout, m, box_m)) != HG_OK) return rc;
    if (d->up_low && (rc = make_map(&kp.map_up, d->up_low, d->cout, m / 4, box_m / 4)) != HG_OK) return rc;
    if (d->pool_out) {
        if ((rc = make_map(&kp.map_pool, d->pool_out, d->cout, m / 4, kUpRows)) != HG_OK) return rc;
        return launch_variant<256, false, false, true>(kp, smem_bytes, stream);
    }

    if (d->pool_in) {
        kp.pool_in = static_cast<__nv_bfloat16*>(d->pool_in);
        kp.pool_in_w = d->w;
        int sh = 0;
        while ((2 << sh) < d->w) ++sh;                     // log2(w / 2)
        kp.pool_in_shift = sh;
        return launch_variant<128, true, false, false, false, true>(kp, smem_bytes, stream);
    }

    if (ragged) {
        if (d->cout == 64)
            return prologue ? launch_variant<64, true, false, false, true>(kp, smem_bytes, stream)
                            : launch_variant<64, false, false, false, true>(kp, smem_bytes, stream);
        return prologue ? launch_variant<128, true, false, false, true>(kp, smem_bytes, stream)
                        : launch_variant<128, false, false, false, true>(kp, smem_bytes, stream);
    }
    switch (d->cout) {
        case 64: return prologue ? launch<64, true>(kp, smem_bytes, stream) : launch<64, false>(kp, smem_bytes, stream);
        case 128: return prologue ? launch<128, true>(kp, smem_bytes, stream) : launch<128, false>(kp, smem_bytes, stream);
        case 256: return launch<256, false>(kp, smem_bytes, stream);
    }
    set_last_error("conv1x1: unsupported cout %d", d->cout);
    return HG_ERR_INVALID;
}

}  // namespace hg
