// Implicit-GEMM convolution for NHWC bf16 activations on tcgen05 tensor cores (sm_100a).
//
//   D[128 pixels x BLOCK_N couts] (fp32, TMEM) = sum over k-blocks of A[128 x 64] . B[BLOCK_N x 64]^T
//
// * A k-block = 64 input channels of one filter tap, fetched by ONE 4-D TMA box load
//   (c, x, y, n) whose x/y start is shifted by the tap offset; out-of-bounds elements are
//   zero-filled by TMA, which IS the convolution's zero padding.  1x1 convs use the same path
//   with the tensor flattened to (c, n*h*w).
// * B k-block = the matching 64 K-columns of the K-major weight matrix [cout_pad][ktot].
// * Both land in 128-byte-swizzled shared memory and are consumed by tcgen05.mma (kind::f16,
//   M=128, N=BLOCK_N, K=16) issued by one thread; accumulators are double-buffered in TMEM so
//   the epilogue of tile i overlaps the main loop of tile i+1 (persistent CTAs, 1 per SM).
// * optional PROLOGUE (pre-activation bn1 + ReLU of the reference's HGBottleneck,
//   src/models/modules.py:30-32): four extra warps rewrite the landed A tile in place,
//   a = relu(a*scale[c] + shift[c]), before the MMA warp may read it.
// * EPILOGUE: + bias (+ residual) (+ nearest-upsampled low-res tensor) (ReLU) -> bf16 -> swizzled
//   staging tile -> TMA store (clips partial tiles); or fp32 NCHW heat maps for the score heads.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue, 6..9 = prologue.
#include "hg_common.cuh"
#include "../../include/hg_api.h"

#include <cudaTypedefs.h>
#include <cstring>
#include <mutex>

#include <cstdlib>

namespace hg {

// second-generation flat 1x1 kernel (hg_conv1x1.cu)
int conv1x1_supported(const hg_conv_desc* d);
int conv1x1_launch(const hg_conv_desc* d, cudaStream_t stream);

constexpr int kTileM = 128;          // pixels per tile (UMMA M)
constexpr int kBlockK = 64;          // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kAStageBytes = kTileM * kBlockK * 2;   // 16 KiB
constexpr int kNumEpiThreads = 128;
constexpr int kMaxCin = 512;         // prologue scale/shift staging capacity

struct ConvKernelParams {
    CUtensorMap map_a;     // (c, x, y, n) over `in`
    CUtensorMap map_a2;    // (c, x, y, n) over `in2`
    CUtensorMap map_b;     // (k, cout_pad) over weights
    CUtensorMap map_out;   // (c, x, y, n) over `out`
    const float* bias;
    const float* in_scale;
    const float* in_shift;
    const __nv_bfloat16* residual;
    const __nv_bfloat16* up_low;
    float* out_nchw_f32;
    unsigned int* err_word;
    // tile geometry in the (x, y, n) space of the tensor maps
    int W, H, NB;                 // extents (flat 1x1 mode: W = n*h*w, H = NB = 1)
    int box_w, box_h, box_n;      // rows used per tile = box_w*box_h*box_n <= 128
    int tiles_x, tiles_y, num_tiles;
    int img_h, img_w;             // true image geometry (for up_low / NCHW indexing)
    int taps;                     // 1 or 9
    int cin_blocks, cin2_blocks;  // k-blocks per tap of `in`; k-blocks of `in2`
    int cin;                      // channels of `in` (prologue vector length)
    int cout;                     // real output channels
    int relu;
};

template <int BLOCK_N, bool kPrologue>
struct ConvCfg {
    static constexpr int kBStageBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kAStageBytes + kBStageBytes;
    static constexpr int kStagingBytes = (BLOCK_N >= 64) ? 2 * kAStageBytes : 0;
    static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 5 : 6);
    static constexpr int kTmemCols = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;   // power of two for all BLOCK_N used
    static constexpr int kMiscBytes = BLOCK_N * 4 + (kPrologue ? 2 * kMaxCin * 4 : 0) + 256;
    static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kStagingBytes + kMiscBytes;
    static constexpr int kThreads = kPrologue ? 320 : 192;
    static_assert(kSmemBytes <= 232448, "exceeds 227 KiB of shared memory");
    static_assert((kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols <= 512, "TMEM columns must be a power of two");
};

// error codes written to err_word (role << 8 | barrier class)
enum : uint32_t { kErrProducer = 0x100, kErrMma = 0x200, kErrEpilogue = 0x300, kErrPrologue = 0x400 };

template <int BLOCK_N, bool kPrologue>
__global__ void __launch_bounds__(ConvCfg<BLOCK_N, kPrologue>::kThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvKernelParams p) {
    using Cfg = ConvCfg<BLOCK_N, kPrologue>;
    constexpr int kStages = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem_a + kStages * kAStageBytes;
    uint8_t* smem_stage_out = smem_b + kStages * Cfg::kBStageBytes;
    float* s_bias = reinterpret_cast<float*>(smem_stage_out + Cfg::kStagingBytes);
    float* s_scale = s_bias + BLOCK_N;
    float* s_shift = s_scale + (kPrologue ? kMaxCin : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + (kPrologue ? kMaxCin : 0));
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* ready_bar = bars + 2 * kStages;
    uint64_t* tmem_full_bar = bars + 3 * kStages;
    uint64_t* tmem_empty_bar = bars + 3 * kStages + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 4);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int kb1 = p.taps * p.cin_blocks;
    const int num_kb = kb1 + p.cin2_blocks;
    const int rows_used = p.box_w * p.box_h * p.box_n;
    const uint32_t stage_tx_bytes = static_cast<uint32_t>(rows_used * kBlockK * 2 + Cfg::kBStageBytes);

    // ------------------------------------------------------------------ one-time setup
    for (int i = threadIdx.x; i < BLOCK_N; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    if (kPrologue) {
        for (int i = threadIdx.x; i < p.cin; i += blockDim.x) {
            s_scale[i] = p.in_scale[i];
            s_shift[i] = p.in_shift[i];
        }
    }
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&p.map_a);
        tma_prefetch_desc(&p.map_b);
        if (p.cin2_blocks) tma_prefetch_desc(&p.map_a2);
        if (BLOCK_N >= 64) tma_prefetch_desc(&p.map_out);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&ready_bar[s], 4);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 4);
        }
        fence_mbar_init();
    }
    if (warp_idx == 1) tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);   // broadcast: provably warp-uniform for the issue loop
    pdl_launch_dependents();
    pdl_wait();      // everything below reads or overwrites tensors the predecessor kernel may still touch

    // ------------------------------------------------------------------ roles
    if (warp_idx == 0) {
        // ===================== TMA producer (one thread) =====================
        if (elect_one_sync()) {
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x) {
                const int tx = tile % p.tiles_x;
                const int ty = (tile / p.tiles_x) % p.tiles_y;
                const int tn = tile / (p.tiles_x * p.tiles_y);
                const int x0 = tx * p.box_w, y0 = ty * p.box_h, n0 = tn * p.box_n;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ok = mbar_wait(&empty_bar[stage], phase ^ 1u, p.err_word, kErrProducer | 1);
                    if (!ok) break;
                    mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
                    void* a_dst = smem_a + stage * kAStageBytes;
                    if (kb < kb1) {
                        const int tap = kb / p.cin_blocks;
                        const int cb = kb - tap * p.cin_blocks;
                        int dx = 0, dy = 0;
                        if (p.taps == 9) {
                            dy = tap / 3 - 1;
                            dx = tap - (tap / 3) * 3 - 1;
                        }
                        tma_load_4d(a_dst, &p.map_a, &full_bar[stage], cb * kBlockK, x0 + dx, y0 + dy, n0);
                    } else {
                        tma_load_4d(a_dst, &p.map_a2, &full_bar[stage], (kb - kb1) * kBlockK, x0, y0, n0);
                    }
                    tma_load_2d(smem_b + stage * Cfg::kBStageBytes, &p.map_b, &full_bar[stage], kb * kBlockK, 0);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            bool ok = true;
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
                    const uint64_t a_desc = umma_desc_sw128(smem_u32(smem_a + stage * kAStageBytes));
                    const uint64_t b_desc = umma_desc_sw128(smem_u32(smem_b + stage * Cfg::kBStageBytes));
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // +32 bytes per K=16 step inside the 128-byte swizzle row (encoded >> 4)
                        tc_mma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(&empty_bar[stage]);                      // frees this smem stage when the MMAs retire
                    if (kb == num_kb - 1) tc_commit(&tmem_full_bar[acc]);   // accumulator complete
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp_idx < 6) {
        // ===================== epilogue (4 warps, one TMEM lane quarter each) =====================
        const int q = warp_idx & 3;
        const int row = q * 32 + lane;
        const int rows_per_img = p.box_w * p.box_h;
        int it = 0;
        int slab_counter = 0;     // staging buffer ring position (continues across tiles)
        bool ok = true;
        for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1u;
            const int tx = tile % p.tiles_x;
            const int ty = (tile / p.tiles_x) % p.tiles_y;
            const int tn = tile / (p.tiles_x * p.tiles_y);
            const int x0 = tx * p.box_w, y0 = ty * p.box_h, n0 = tn * p.box_n;
            const int nl = row / rows_per_img;
            const int rem = row - nl * rows_per_img;
            const int hl = rem / p.box_w;
            const int wl = rem - hl * p.box_w;
            const bool valid = (row < rows_used) && (n0 + nl < p.NB) && (y0 + hl < p.H) && (x0 + wl < p.W);
            const long long pix = (static_cast<long long>(n0 + nl) * p.H + (y0 + hl)) * p.W + (x0 + wl);
            long long low_pix = 0;
            long long img_n = 0;
            int img_rem = 0;
            if (p.up_low != nullptr || p.out_nchw_f32 != nullptr) {
                const int hw = p.img_h * p.img_w;
                img_n = pix / hw;
                img_rem = static_cast<int>(pix - img_n * hw);
                const int y = img_rem / p.img_w, x = img_rem - y * p.img_w;
                low_pix = (img_n * (p.img_h >> 1) + (y >> 1)) * (p.img_w >> 1) + (x >> 1);
            }

            ok = mbar_wait(&tmem_full_bar[acc], acc_phase, p.err_word, kErrEpilogue | 1);
            if (!ok) break;
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);

            if constexpr (BLOCK_N < 64) {
                // ---- score heads: fp32 NCHW heat maps, lanes = consecutive pixels -> coalesced
                uint32_t v[32];
                if constexpr (BLOCK_N == 16) {
                    uint32_t v16[16];
                    tmem_ld_32x16(t_row, v16);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = v16[i];
                } else {
                    tmem_ld_32x32(t_row, v);
                    tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                if (valid) {
                    const int hw = p.img_h * p.img_w;
                    float* o = p.out_nchw_f32 + img_n * p.cout * hw + img_rem;
#pragma unroll
                    for (int c = 0; c < BLOCK_N; ++c) {
                        if (c < p.cout) {
                            float f = __uint_as_float(v[c]) + s_bias[c];
                            if (p.relu) f = fmaxf(f, 0.f);
                            o[static_cast<long long>(c) * hw] = f;
                        }
                    }
                }
            } else {
                // ---- bf16 NHWC output through a swizzled staging tile + TMA store
                constexpr int kSlabs = BLOCK_N / 64;
                for (int slab = 0; slab < kSlabs; ++slab, ++slab_counter) {
                    uint8_t* stg = smem_stage_out + (slab_counter & 1) * kAStageBytes;
                    // the TMA store that last read this buffer (two slabs ago) must have drained
                    if (warp_idx == 2 && elect_one_sync()) tma_store_wait_read<1>();
                    named_bar_sync(1, kNumEpiThreads);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int col0 = slab * 64 + half * 32;
                        uint32_t v[32];
                        tmem_ld_32x32(t_row + col0, v);
                        uint4 r_res[4], r_up[4];
                        if (p.residual != nullptr && valid) {
                            const uint4* src = reinterpret_cast<const uint4*>(p.residual + pix * p.cout + col0);
#pragma unroll
                            for (int i = 0; i < 4; ++i) r_res[i] = ldg_v4(src + i);
                        }
                        if (p.up_low != nullptr && valid) {
                            const uint4* src = reinterpret_cast<const uint4*>(p.up_low + low_pix * p.cout + col0);
#pragma unroll
                            for (int i = 0; i < 4; ++i) r_up[i] = ldg_v4(src + i);
                        }
                        tmem_ld_wait();
                        if (slab == kSlabs - 1 && half == 1) {
                            // every TMEM read of this accumulator is done: hand it back to the MMA warp
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                        }
                        float f[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) + s_bias[col0 + i];
                        if (p.residual != nullptr && valid) {
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
                        if (p.up_low != nullptr && valid) {
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
                        if (p.relu) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
                        }
                        // 4 x 16-byte chunks of this row, 128-byte swizzle: chunk' = chunk ^ (row & 7)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4 o;
                            o.x = pack_bf16x2(f[i * 8 + 0], f[i * 8 + 1]);
                            o.y = pack_bf16x2(f[i * 8 + 2], f[i * 8 + 3]);
                            o.z = pack_bf16x2(f[i * 8 + 4], f[i * 8 + 5]);
                            o.w = pack_bf16x2(f[i * 8 + 6], f[i * 8 + 7]);
                            const int chunk = (half * 4 + i) ^ (row & 7);
                            *reinterpret_cast<uint4*>(stg + row * 128 + chunk * 16) = o;
                        }
                    }
                    fence_proxy_async_smem();          // generic-proxy writes -> visible to the TMA engine
                    named_bar_sync(1, kNumEpiThreads);
                    if (warp_idx == 2 && elect_one_sync()) {
                        tma_store_4d(&p.map_out, stg, slab * 64, x0, y0, n0);
                        tma_store_commit();
                    }
                }
            }
        }
        if (BLOCK_N >= 64 && warp_idx == 2 && elect_one_sync()) tma_store_wait<0>();
    } else if (kPrologue) {
        // ===================== prologue: a = relu(a*scale + shift) in place (4 warps) =====================
        const int t = threadIdx.x - 192;
        const int g = t >> 3, l = t & 7;      // 16 groups of 8 rows; l = row within the swizzle atom
        const int row = g * 8 + l;
        int stage = 0;
        uint32_t phase = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x) {
            for (int kb = 0; kb < num_kb; ++kb) {
                ok = mbar_wait(&full_bar[stage], phase, p.err_word, kErrPrologue | 1);
                if (!ok) break;
                if (kb < kb1) {
                    const int cbase = (kb % p.cin_blocks) * kBlockK;
                    uint8_t* a_row = smem_a + stage * kAStageBytes + row * 128;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int pc = (i + l) & 7;           // physical 16-byte chunk (bank-conflict free)
                        const int c0 = cbase + ((pc ^ l) << 3);   // logical channel of its first element
                        uint4 d = *reinterpret_cast<uint4*>(a_row + pc * 16);
                        const float4 s0 = *reinterpret_cast<const float4*>(s_scale + c0);
                        const float4 s1 = *reinterpret_cast<const float4*>(s_scale + c0 + 4);
                        const float4 h0 = *reinterpret_cast<const float4*>(s_shift + c0);
                        const float4 h1 = *reinterpret_cast<const float4*>(s_shift + c0 + 4);
                        d.x = pack_bf16x2(fmaxf(fmaf(bf16_lo_to_f32(d.x), s0.x, h0.x), 0.f),
                                          fmaxf(fmaf(bf16_hi_to_f32(d.x), s0.y, h0.y), 0.f));
                        d.y = pack_bf16x2(fmaxf(fmaf(bf16_lo_to_f32(d.y), s0.z, h0.z), 0.f),
                                          fmaxf(fmaf(bf16_hi_to_f32(d.y), s0.w, h0.w), 0.f));
                        d.z = pack_bf16x2(fmaxf(fmaf(bf16_lo_to_f32(d.z), s1.x, h1.x), 0.f),
                                          fmaxf(fmaf(bf16_hi_to_f32(d.z), s1.y, h1.y), 0.f));
                        d.w = pack_bf16x2(fmaxf(fmaf(bf16_lo_to_f32(d.w), s1.z, h1.z), 0.f),
                                          fmaxf(fmaf(bf16_hi_to_f32(d.w), s1.w, h1.w), 0.f));
                        *reinterpret_cast<uint4*>(a_row + pc * 16) = d;
                    }
                    fence_proxy_async_smem();   // generic-proxy writes -> visible to tcgen05.mma
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready_bar[stage]);
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    }

    // ------------------------------------------------------------------ teardown
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// =================================================================================================
// host side
// =================================================================================================
static PFN_cuTensorMapEncodeTiled_v12000 get_tensormap_encode() {
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

// bf16 tensor viewed as (c, x, y, n) with c contiguous; box = (box_c, bw, bh, bn)
static int make_map_4d(CUtensorMap* map, const void* ptr, uint64_t C, uint64_t X, uint64_t Y, uint64_t N,
                       uint32_t box_c, uint32_t bw, uint32_t bh, uint32_t bn, CUtensorMapSwizzle swz) {
    auto enc = get_tensormap_encode();
    if (!enc) return HG_ERR_CUDA;
    cuuint64_t gdim[4] = {C, X, Y, N};
    cuuint64_t gstr[3] = {C * 2, X * C * 2, Y * X * C * 2};
    cuuint32_t box[4] = {box_c, bw, bh, bn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled(4d) failed: CUresult %d (dims %llu %llu %llu %llu box %u %u %u %u)",
                       (int)r, (unsigned long long)C, (unsigned long long)X, (unsigned long long)Y,
                       (unsigned long long)N, box_c, bw, bh, bn);
        return HG_ERR_CUDA;
    }
    return HG_OK;
}

static int make_map_2d(CUtensorMap* map, const void* ptr, uint64_t K, uint64_t R, uint32_t box_k, uint32_t box_r) {
    auto enc = get_tensormap_encode();
    if (!enc) return HG_ERR_CUDA;
    cuuint64_t gdim[2] = {K, R};
    cuuint64_t gstr[1] = {K * 2};
    cuuint32_t box[2] = {box_k, box_r};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled(2d) failed: CUresult %d (K %llu R %llu box %u %u)", (int)r,
                       (unsigned long long)K, (unsigned long long)R, box_k, box_r);
        return HG_ERR_CUDA;
    }
    return HG_OK;
}

template <int BLOCK_N, bool kPrologue>
static int launch_conv(const ConvKernelParams& kp, cudaStream_t stream) {
    using Cfg = ConvCfg<BLOCK_N, kPrologue>;
    auto kern = conv_gemm_kernel<BLOCK_N, kPrologue>;
    // opt in to >48 KiB dynamic shared memory once per (instantiation, device)
    static std::mutex mu;
    static unsigned long long done_mask = 0;
    int dev = 0;
    HG_CUDA_OK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 64 || !(done_mask >> dev & 1ull)) {
            HG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
            if (dev < 64) done_mask |= 1ull << dev;
        }
    }
    const int grid = kp.num_tiles < num_sms() ? kp.num_tiles : num_sms();
    HG_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, kp));
    return HG_OK;
}

}  // namespace hg

extern "C" int hg_conv_nhwc_bf16(const hg_conv_desc* d, void* stream_v) {
    using namespace hg;
    if (d == nullptr) {
        set_last_error("hg_conv_nhwc_bf16: null descriptor");
        return HG_ERR_INVALID;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const bool heads = d->out_nchw_f32 != nullptr;
    const int cout_pad = (d->cout + 15) / 16 * 16;
    if (d->in == nullptr || d->weight == nullptr || (d->out == nullptr && !heads) || d->n <= 0 || d->h <= 0 ||
        d->w <= 0) {
        set_last_error("hg_conv_nhwc_bf16: null pointer or empty shape");
        return HG_ERR_INVALID;
    }
    if (d->ksize != 1 && d->ksize != 3) {
        set_last_error("hg_conv_nhwc_bf16: ksize %d unsupported (1 or 3)", d->ksize);
        return HG_ERR_INVALID;
    }
    if (d->cin % 64 != 0 || d->cin <= 0 || d->cin2 % 64 != 0 || (d->in2 == nullptr) != (d->cin2 == 0)) {
        set_last_error("hg_conv_nhwc_bf16: cin=%d cin2=%d must be multiples of 64 (cin2>0 iff in2 given)", d->cin,
                       d->cin2);
        return HG_ERR_INVALID;
    }
    if (d->stats != nullptr && heads) {
        set_last_error("hg_conv_nhwc_bf16: stats need a bf16 NHWC output");
        return HG_ERR_INVALID;
    }
    if (heads) {
        if (cout_pad > 32 || d->residual || d->up_low) {
            set_last_error("hg_conv_nhwc_bf16: fp32 NCHW output supports cout<=32 without residual terms");
            return HG_ERR_INVALID;
        }
    } else if (d->cout != 64 && d->cout != 128 && d->cout != 256) {
        set_last_error("hg_conv_nhwc_bf16: bf16 NHWC output needs cout in {64,128,256}, got %d", d->cout);
        return HG_ERR_INVALID;
    }
    const bool prologue = d->in_scale != nullptr;
    if (prologue && (d->ksize != 1 || d->in_shift == nullptr || d->cin > kMaxCin || (cout_pad != 64 && cout_pad != 128))) {
        set_last_error("hg_conv_nhwc_bf16: prologue needs ksize 1, cin<=%d, cout in {64,128}", kMaxCin);
        return HG_ERR_INVALID;
    }
    if (d->up_low && ((d->h & 1) || (d->w & 1))) {
        set_last_error("hg_conv_nhwc_bf16: up_low needs even h,w");
        return HG_ERR_INVALID;
    }

    {
        // HG_CONV1X1_GENERIC=1 forces the generic kernel for 1x1 convs (A/B comparisons only)
        static const bool force_generic = getenv("HG_CONV1X1_GENERIC") != nullptr;
        if (!force_generic && conv1x1_supported(d)) return conv1x1_launch(d, stream);
    }

    if (d->pool_in) {
        set_last_error("hg_conv_nhwc_bf16: pool_in needs a 1x1 conv with prologue, cout 128, cin2 0, no stats, w a power of two "
                       "<= 64 with 128 %% (2*w) == 0 and even h");
        return HG_ERR_INVALID;
    }
    if (d->pool_out) {
        set_last_error("hg_conv_nhwc_bf16: pool_out needs a 1x1 conv with cout 256, no prologue / stats / out_halo, w a power of two "
                       "<= 64 with 128 %% (2*w) == 0 and even h");
        return HG_ERR_INVALID;
    }
    if (d->out_halo) {
        set_last_error("hg_conv_nhwc_bf16: out_halo needs a 1x1 conv with 128 %% w == 0, (h*w) %% 128 == 0, cout in {64,128,256}");
        return HG_ERR_INVALID;
    }

    ConvKernelParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.bias = d->bias;
    kp.in_scale = d->in_scale;
    kp.in_shift = d->in_shift;
    kp.residual = static_cast<const __nv_bfloat16*>(d->residual);
    kp.up_low = static_cast<const __nv_bfloat16*>(d->up_low);
    kp.out_nchw_f32 = d->out_nchw_f32;
    kp.err_word = d->err_word;
    kp.img_h = d->h;
    kp.img_w = d->w;
    kp.taps = d->ksize * d->ksize;
    kp.cin_blocks = d->cin / 64;
    kp.cin2_blocks = d->cin2 / 64;
    kp.cin = d->cin;
    kp.cout = d->cout;
    kp.relu = d->relu;

    if (d->ksize == 1) {     // flat: (c, n*h*w)
        const long long m = static_cast<long long>(d->n) * d->h * d->w;
        if (m > 0x7fffffffLL) {
            set_last_error("hg_conv_nhwc_bf16: too many pixels");
            return HG_ERR_INVALID;
        }
        kp.W = static_cast<int>(m);
        kp.H = 1;
        kp.NB = 1;
        kp.box_w = kTileM;
        kp.box_h = 1;
        kp.box_n = 1;
    } else {
        kp.W = d->w;
        kp.H = d->h;
        kp.NB = d->n;
        kp.box_w = d->w < kTileM ? d->w : kTileM;
        if (d->h * d->w <= kTileM) {
            kp.box_h = d->h;
            kp.box_n = kTileM / (d->h * d->w);
            if (kp.box_n > d->n) kp.box_n = d->n;
        } else {
            kp.box_n = 1;
            kp.box_h = kTileM / kp.box_w;
            if (kp.box_h > d->h) kp.box_h = d->h;
        }
    }
    kp.tiles_x = (kp.W + kp.box_w - 1) / kp.box_w;
    kp.tiles_y = (kp.H + kp.box_h - 1) / kp.box_h;
    const int tiles_n = (kp.NB + kp.box_n - 1) / kp.box_n;
    kp.num_tiles = kp.tiles_x * kp.tiles_y * tiles_n;

    const int ktot = kp.taps * d->cin + d->cin2;
    int rc;
    if ((rc = make_map_4d(&kp.map_a, d->in, d->cin, kp.W, kp.H, kp.NB, 64, kp.box_w, kp.box_h, kp.box_n,
                          CU_TENSOR_MAP_SWIZZLE_128B)) != HG_OK)
        return rc;
    if (d->in2 != nullptr) {
        if ((rc = make_map_4d(&kp.map_a2, d->in2, d->cin2, kp.W, kp.H, kp.NB, 64, kp.box_w, kp.box_h, kp.box_n,
                              CU_TENSOR_MAP_SWIZZLE_128B)) != HG_OK)
            return rc;
    }
    if ((rc = make_map_2d(&kp.map_b, d->weight, ktot, cout_pad, 64, cout_pad)) != HG_OK) return rc;
    if (!heads) {
        if ((rc = make_map_4d(&kp.map_out, d->out, d->cout, kp.W, kp.H, kp.NB, 64, kp.box_w, kp.box_h, kp.box_n,
                              CU_TENSOR_MAP_SWIZZLE_128B)) != HG_OK)
            return rc;
    }

    switch (cout_pad) {
        case 16: rc = launch_conv<16, false>(kp, stream); break;
        case 32: rc = launch_conv<32, false>(kp, stream); break;
        case 64: rc = prologue ? launch_conv<64, true>(kp, stream) : launch_conv<64, false>(kp, stream); break;
        case 128: rc = prologue ? launch_conv<128, true>(kp, stream) : launch_conv<128, false>(kp, stream); break;
        case 256: rc = launch_conv<256, false>(kp, stream); break;
        default:
            set_last_error("hg_conv_nhwc_bf16: unsupported cout_pad %d", cout_pad);
            return HG_ERR_INVALID;
    }
    if (rc == HG_OK && d->stats != nullptr) {
        // this general-shape kernel has no fused statistics: one pass over the (L2-resident, just written) result
        rc = hg_colstats_nhwc(d->out, d->stats, d->stats + d->cout, static_cast<int64_t>(d->n) * d->h * d->w, d->cout, d->cout,
                              0, nullptr, stream_v);
    }
    return rc;
}
