// 3x3 convolution (stride 1, pad 1) as implicit GEMM on tcgen05, second generation.
//
// The first-generation kernel (hg_conv_gemm.cu) fetches every filter tap as its own TMA box: nine
// L2 -> shared-memory reads of the same pixels per tile, which leaves it L2-bandwidth bound at ~37 % of
// the tensor pipe.  This kernel reads each input pixel ONCE per tile:
//
//   * the input lives in a HALO-PADDED layout in HBM: [1 zero row][n][h+1][w+1][c] -- one zero column
//     right of every image row and one zero row under every image (written by the producing 1x1
//     kernel through a strided TMA store; the pads are zeroed once and never touched).  In the flat
//     "position" index f every 3x3 tap is then a CONSTANT offset dy*(w+1)+dx, and the pads are the
//     convolution's zero padding on all four sides.
//   * a tile is 256 consecutive positions.  Its halo'd run (256 + 2(w+1) + 2 positions x 64 channels) is
//     loaded once into 128-byte-swizzled shared memory; the A operand of tap (dy,dx) is the same
//     buffer with the UMMA descriptor's start address shifted by whole 128-byte rows (the hardware
//     swizzles on absolute address bits, so row-shifted descriptors are exact -- probed on B200).
//   * two 128-row accumulators share every B k-block, halving weight traffic again.
//
// Per 256 outputs: A traffic 2 x 50 KiB (was 2 x 288 KiB), B traffic 288 KiB (was 576 KiB).
// Warp roles: 0 = A producer, 1 = MMA issuer + TMEM owner, 2 = B producer, 4..7 = epilogue.
//
// Three kernels live here: conv3x3_kernel (one CTA per tile; training and the small levels), conv3x3_pair_kernel (the same
// GEMM on CTA pairs, cta_group::2: half the shared memory and TMEM per CTA) and conv3x3_k3_pair_kernel, which uses that
// room to run the bottleneck's closing 1x1 (+ residual / upsample-add) on the 3x3's result without leaving the SM.
#include "hg_common.cuh"
#include "../../include/hg_api.h"

#include <cudaTypedefs.h>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace hg {
namespace c3 {

constexpr int kBM = 256;            // positions per tile (two UMMA M=128 sub-tiles)
constexpr int kBlockK = 64;
constexpr int kMaxBStages = 8;
constexpr int kSmemLimit = 232448;

struct Params {
    CUtensorMap map_a;      // (c, positions) over the halo-padded input
    CUtensorMap map_b;      // (k, cout) weights, k = tap*cin + c
    const float* bias;
    __nv_bfloat16* out;     // dense NHWC [n][h][w][cout]
    unsigned int* err_word;
    float* stats;           // optional fp32 [2*cout]: per-channel sum | sum of squares of the fp32 results, ADDED
    int H, W, P, NB;
    int cin, cout, slabs;
    int num_tiles;
    int n_split;            // output channels are split over n_split CTAs per tile (cout = n_split * BLOCK_N): small grids
    int num_work;           // num_tiles * n_split work items (tile-major)
    long long total_pos;    // NB*(H+1)*P  (positions after the leading zero row)
    int box_rows, num_boxes, region_bytes;   // halo'd run = num_boxes TMA boxes of box_rows rows
    int b_stages;
    int relu;
    // K2 + K3 fusion (conv3x3_k3_pair_kernel): the bottleneck's closing 1x1 (128 -> 256) + residual (+ upsample-add)
    CUtensorMap map_w3;             // (k = 128, cout3 = 256) weights of the 1x1
    const float* bias3;
    const __nv_bfloat16* res;       // dense NHWC [n][h][w][256] residual, or null
    const __nv_bfloat16* up;        // dense NHWC [n][h/2][w/2][256] nearest-upsampled and added, or null
    __nv_bfloat16* out3;            // dense NHWC [n][h][w][256]
};

// kStats: the epilogue also adds the per-channel sum / sum of squares of its results into p.stats (a separate
// instantiation: the plain kernel keeps its register budget and schedule).
template <int BLOCK_N, bool kStats>
__global__ void __launch_bounds__(256, 1) conv3x3_kernel(const __grid_constant__ Params p) {
    constexpr int kBStage = BLOCK_N * kBlockK * 2;
    constexpr int kTmemCols = 4 * BLOCK_N;            // 2 accumulator stages x 2 sub-tiles

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                    // 2 regions
    uint8_t* smem_b = smem_a + 2 * p.region_bytes;             // b_stages x kBStage
    float* s_bias = reinterpret_cast<float*>(smem_b + p.b_stages * kBStage);
    float* s_stats = s_bias + p.cout;                           // [2*cout]; s_bias holds all cout = n_split * BLOCK_N channels
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stats + (kStats ? 2 * p.cout : 0));
    uint64_t* a_full = bars;                  // [2]
    uint64_t* a_empty = bars + 2;             // [2]
    uint64_t* b_full = bars + 4;              // [kMaxBStages]
    uint64_t* b_empty = b_full + kMaxBStages;
    uint64_t* tmem_full_bar = b_empty + kMaxBStages;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < p.cout; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    if (kStats)
        for (int i = threadIdx.x; i < 2 * p.cout; i += blockDim.x) s_stats[i] = 0.f;
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&p.map_a);
        tma_prefetch_desc(&p.map_b);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 4);
        }
        for (int s = 0; s < kMaxBStages; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        fence_mbar_init();
    }
    if (warp_idx == 1) tmem_alloc(tmem_ptr_smem, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);   // broadcast: provably warp-uniform, so the MMA issue loop keeps it in a uniform register (no per-MMA ELECT/R2UR waterfall)
    pdl_launch_dependents();

    const int halo = p.P + 1;                 // positions before the tile start that the taps reach

    if (warp_idx == 0) {
        // ===================== A producer: one halo'd run per (tile, 64-channel slab) =====================
        if (elect_one_sync()) {
            pdl_wait();
            int it = 0;
            bool ok = true;
            for (int wi = blockIdx.x; wi < p.num_work && ok; wi += gridDim.x) {
                const int tile = wi / p.n_split;
                // position 0 of the tensor map is the leading zero row; tiles start after it
                const long long f0 = static_cast<long long>(p.P) + static_cast<long long>(tile) * kBM;
                const int row0 = static_cast<int>(f0 - halo);
                for (int slab = 0; slab < p.slabs; ++slab, ++it) {
                    const int buf = it & 1;
                    ok = mbar_wait(&a_empty[buf], ((it >> 1) & 1u) ^ 1u, p.err_word, 0x3101);
                    if (!ok) break;
                    mbar_arrive_expect_tx(&a_full[buf], static_cast<uint32_t>(p.num_boxes * p.box_rows * 128));
                    for (int b = 0; b < p.num_boxes; ++b)
                        tma_load_2d(smem_a + buf * p.region_bytes + b * p.box_rows * 128, &p.map_a, &a_full[buf],
                                    slab * kBlockK, row0 + b * p.box_rows);
                }
            }
        }
    } else if (warp_idx == 2) {
        // ===================== B producer: weight k-blocks (tap, slab), streamed =====================
        if (elect_one_sync()) {
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (int wi = blockIdx.x; wi < p.num_work && ok; wi += gridDim.x) {
                const int n0 = (wi % p.n_split) * BLOCK_N;          // this work item's output-channel window
                for (int slab = 0; slab < p.slabs && ok; ++slab) {
                    for (int tap = 0; tap < 9; ++tap) {
                        ok = mbar_wait(&b_empty[stage], phase ^ 1u, p.err_word, 0x3301);
                        if (!ok) break;
                        mbar_arrive_expect_tx(&b_full[stage], kBStage);
                        tma_load_2d(smem_b + stage * kBStage, &p.map_b, &b_full[stage], tap * p.cin + slab * kBlockK, n0);
                        if (++stage == p.b_stages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== MMA issuer =====================
        if (elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0, a_it = 0;
            bool ok = true;
            for (int wi = blockIdx.x; wi < p.num_work && ok; wi += gridDim.x, ++it) {
                const int acc = it & 1;
                ok = mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1u) ^ 1u, p.err_word, 0x3201);
                if (!ok) break;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 2 * BLOCK_N);
                for (int slab = 0; slab < p.slabs && ok; ++slab, ++a_it) {
                    const int buf = a_it & 1;
                    ok = mbar_wait(&a_full[buf], (a_it >> 1) & 1u, p.err_word, 0x3202);
                    if (!ok) break;
                    const uint32_t a_base = smem_u32(smem_a + buf * p.region_bytes);
                    for (int tap = 0; tap < 9; ++tap) {
                        ok = mbar_wait(&b_full[stage], phase, p.err_word, 0x3203);
                        if (!ok) break;
                        tc_fence_after();
                        const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                        const uint32_t row_off = static_cast<uint32_t>(halo + dy * p.P + dx);   // >= 0 by construction
                        const uint64_t b_desc = umma_desc_sw128(smem_u32(smem_b + stage * kBStage));
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const uint64_t a_desc = umma_desc_sw128(a_base + (row_off + half * 128) * 128);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_mma_bf16(d_tmem + half * BLOCK_N, a_desc + 2u * k, b_desc + 2u * k, idesc,
                                            (slab | tap | k) != 0 ? 1u : 0u);
                        }
                        tc_commit(&b_empty[stage]);
                        if (++stage == p.b_stages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    if (ok) tc_commit(&a_empty[buf]);
                }
                if (ok) tc_commit(&tmem_full_bar[acc]);
            }
        }
    } else if (warp_idx >= 4) {
        // ===================== epilogue: bias + ReLU -> bf16, dense NHWC stores (pads skipped) =====================
        pdl_wait();
        const int q = warp_idx & 3;
        const int row = q * 32 + lane;
        const long long img_pos = static_cast<long long>(p.H + 1) * p.P;
        int it = 0;
        bool ok = true;
        for (int wi = blockIdx.x; wi < p.num_work && ok; wi += gridDim.x, ++it) {
            const int tile = wi / p.n_split;
            const int n0 = (wi - tile * p.n_split) * BLOCK_N;
            const float* bias_w = s_bias + n0;
            const int acc = it & 1;
            ok = mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1u, p.err_word, 0x3401);
            if (!ok) break;
            tc_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const long long f = static_cast<long long>(tile) * kBM + half * 128 + row;     // position after the zero row
                const long long n = f / img_pos;
                const int r = static_cast<int>(f - n * img_pos);
                const int y = r / p.P, x = r - y * p.P;
                const bool valid = (f < p.total_pos) && (x < p.W) && (y < p.H);
                __nv_bfloat16* o = p.out + ((n * p.H + y) * p.W + x) * p.cout + n0;
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                       static_cast<uint32_t>(acc * 2 * BLOCK_N + half * BLOCK_N);
#pragma unroll
                for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_row + c0, v);
                    tmem_ld_wait();
                    if (half == 1 && c0 == BLOCK_N - 32) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                    }
                    if (!kStats) {
                        if (valid) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float f8[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    f8[j] = __uint_as_float(v[i * 8 + j]) + bias_w[c0 + i * 8 + j];
                                    if (p.relu) f8[j] = fmaxf(f8[j], 0.f);
                                }
                                uint4 w4;
                                w4.x = pack_bf16x2(f8[0], f8[1]);
                                w4.y = pack_bf16x2(f8[2], f8[3]);
                                w4.z = pack_bf16x2(f8[4], f8[5]);
                                w4.w = pack_bf16x2(f8[6], f8[7]);
                                *reinterpret_cast<uint4*>(o + c0 + i * 8) = w4;
                            }
                        }
                    } else {
                        float fv[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            fv[i] = __uint_as_float(v[i]) + bias_w[c0 + i];
                            if (p.relu) fv[i] = fmaxf(fv[i], 0.f);
                        }
                        if (valid) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                uint4 w4;
                                w4.x = pack_bf16x2(fv[i * 8 + 0], fv[i * 8 + 1]);
                                w4.y = pack_bf16x2(fv[i * 8 + 2], fv[i * 8 + 3]);
                                w4.z = pack_bf16x2(fv[i * 8 + 4], fv[i * 8 + 5]);
                                w4.w = pack_bf16x2(fv[i * 8 + 6], fv[i * 8 + 7]);
                                *reinterpret_cast<uint4*>(o + c0 + i * 8) = w4;
                            }
                        }
                        // train-mode BatchNorm statistics of the next layer, fused: column sums over the valid rows
                        float sq[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            fv[i] = valid ? fv[i] : 0.f;
                            sq[i] = fv[i] * fv[i];
                        }
                        const float cs = warp_column_sum(fv, lane);
                        const float cq = warp_column_sum(sq, lane);
                        atomicAdd(&s_stats[n0 + c0 + lane], cs);
                        atomicAdd(&s_stats[p.cout + n0 + c0 + lane], cq);
                    }
                }
            }
        }
        if (kStats) {
            named_bar_sync(1, 128);
            for (int i = threadIdx.x - 128; i < 2 * p.cout; i += 128) atomicAdd(p.stats + i, s_stats[i]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ================================================================================================================
// Paired-CTA variant (cta_group::2): the two CTAs of a cluster share ONE 256-position tile.  Each CTA stages the
// halo'd run of its own 128 positions and HALF of every weight k-block (64 of the 128 output channels' rows); the
// leader CTA issues M=256 MMAs that read A and B from both CTAs' shared memory and write each CTA's 128 accumulator
// rows into that CTA's TMEM.  Per CTA this halves the shared memory of the single-CTA kernel (2 x 34 KB of A regions,
// 8 KB weight stages) and its TMEM (2 x 128 columns) at the same weight traffic per position -- the room the K2+K3
// fusion needs.  Selected with HG_CONV3X3_PAIR=1 (cout = 128, no statistics).
// ================================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-SM TMA load: lands in THIS CTA's shared memory, completes its bytes on the LEADER CTA's mbarrier (same offset, peer bit
// cleared).
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// one arrival on the barrier at this offset in BOTH CTAs once every MMA issued so far has retired
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_on_leader(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}
// Same, without release semantics: for arrivals that only hand TMEM columns back (ordered by tcgen05.fence::before_thread_sync).
// A cluster-scope RELEASE waits until the thread's earlier global stores are performed -- an epilogue that has just stored its
// previous sub-slab would hold the accumulator for the store round trip.
__device__ __forceinline__ void mbar_arrive_on_leader_relaxed(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}

constexpr int kPairBStages = 16;
constexpr int kPairBStage = 64 * kBlockK * 2;          // this CTA's half of a [128 x 64] weight k-block

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) conv3x3_pair_kernel(const __grid_constant__ Params p) {
    constexpr int BLOCK_N = 128;
    constexpr int kTmemCols = 2 * BLOCK_N;              // 2 accumulator stages of this CTA's 128 rows

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                    // 2 regions
    uint8_t* smem_b = smem_a + 2 * p.region_bytes;             // b_stages x 8 KiB
    float* s_bias = reinterpret_cast<float*>(smem_b + p.b_stages * kPairBStage);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + BLOCK_N);
    uint64_t* a_full = bars;                  // [2]   (used in the leader)
    uint64_t* a_empty = bars + 2;             // [2]
    uint64_t* b_full = bars + 4;              // [kPairBStages] (used in the leader)
    uint64_t* b_empty = b_full + kPairBStages;
    uint64_t* tmem_full_bar = b_empty + kPairBStages;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2]   (used in the leader: 4 epilogue warps of each CTA)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    for (int i = threadIdx.x; i < BLOCK_N; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&p.map_a);
        tma_prefetch_desc(&p.map_b);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 8);
        }
        for (int s = 0; s < kPairBStages; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        fence_mbar_init();
    }
    if (warp_idx == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                        // both CTAs' barriers exist before either one's TMA / commit can touch them
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
    pdl_launch_dependents();

    const int halo = p.P + 1;

    if (warp_idx == 0) {
        // ===================== A producer: this CTA's 128 positions (+ halo) per (tile, slab) =====================
        if (elect_one_sync()) {
            pdl_wait();
            int it = 0;
            bool ok = true;
            for (int tile = pair; tile < p.num_tiles && ok; tile += num_pairs) {
                const long long f0 = static_cast<long long>(p.P) + static_cast<long long>(tile) * kBM + rank * 128;
                const int row0 = static_cast<int>(f0 - halo);
                for (int slab = 0; slab < p.slabs; ++slab, ++it) {
                    const int buf = it & 1;
                    ok = mbar_wait(&a_empty[buf], ((it >> 1) & 1u) ^ 1u, p.err_word, 0x3501);
                    if (!ok) break;
                    if (leader) mbar_arrive_expect_tx(&a_full[buf], static_cast<uint32_t>(2 * p.num_boxes * p.box_rows * 128));
                    for (int b = 0; b < p.num_boxes; ++b)
                        tma_load_2d_2sm(smem_a + buf * p.region_bytes + b * p.box_rows * 128, &p.map_a, &a_full[buf],
                                        slab * kBlockK, row0 + b * p.box_rows);
                }
            }
        }
    } else if (warp_idx == 2) {
        // ===================== B producer: this CTA's 64 output-channel rows of every k-block =====================
        if (elect_one_sync()) {
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (int tile = pair; tile < p.num_tiles && ok; tile += num_pairs) {
                for (int slab = 0; slab < p.slabs && ok; ++slab) {
                    for (int tap = 0; tap < 9; ++tap) {
                        ok = mbar_wait(&b_empty[stage], phase ^ 1u, p.err_word, 0x3701);
                        if (!ok) break;
                        if (leader) mbar_arrive_expect_tx(&b_full[stage], 2 * kPairBStage);
                        tma_load_2d_2sm(smem_b + stage * kPairBStage, &p.map_b, &b_full[stage], tap * p.cin + slab * kBlockK,
                                        static_cast<int>(rank) * 64);
                        if (++stage == p.b_stages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_bf16(256, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0, a_it = 0;
            bool ok = true;
            for (int tile = pair; tile < p.num_tiles && ok; tile += num_pairs, ++it) {
                const int acc = it & 1;
                ok = mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1u) ^ 1u, p.err_word, 0x3601);
                if (!ok) break;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
                for (int slab = 0; slab < p.slabs && ok; ++slab, ++a_it) {
                    const int buf = a_it & 1;
                    ok = mbar_wait(&a_full[buf], (a_it >> 1) & 1u, p.err_word, 0x3602);
                    if (!ok) break;
                    const uint32_t a_base = smem_u32(smem_a + buf * p.region_bytes);
                    for (int tap = 0; tap < 9; ++tap) {
                        ok = mbar_wait(&b_full[stage], phase, p.err_word, 0x3603);
                        if (!ok) break;
                        tc_fence_after();
                        const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                        const uint32_t row_off = static_cast<uint32_t>(halo + dy * p.P + dx);
                        const uint64_t b_desc = umma_desc_sw128(smem_u32(smem_b + stage * kPairBStage));
                        const uint64_t a_desc = umma_desc_sw128(a_base + row_off * 128);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_bf16_2sm(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (slab | tap | k) != 0 ? 1u : 0u);
                        tc_commit_2sm(&b_empty[stage]);
                        if (++stage == p.b_stages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    if (ok) tc_commit_2sm(&a_empty[buf]);
                }
                if (ok) tc_commit_2sm(&tmem_full_bar[acc]);
            }
        }
    } else if (warp_idx >= 4) {
        // ===================== epilogue: this CTA's 128 rows =====================
        pdl_wait();
        const int q = warp_idx & 3;
        const int row = q * 32 + lane;
        const long long img_pos = static_cast<long long>(p.H + 1) * p.P;
        int it = 0;
        bool ok = true;
        for (int tile = pair; tile < p.num_tiles && ok; tile += num_pairs, ++it) {
            const int acc = it & 1;
            ok = mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1u, p.err_word, 0x3801);
            if (!ok) break;
            tc_fence_after();
            const long long f = static_cast<long long>(tile) * kBM + rank * 128 + row;
            const long long n = f / img_pos;
            const int r = static_cast<int>(f - n * img_pos);
            const int y = r / p.P, x = r - y * p.P;
            const bool valid = (f < p.total_pos) && (x < p.W) && (y < p.H);
            __nv_bfloat16* o = p.out + ((n * p.H + y) * p.W + x) * p.cout;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
#pragma unroll
            for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + c0, v);
                tmem_ld_wait();
                if (c0 == BLOCK_N - 32) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_on_leader(&tmem_empty_bar[acc]);
                }
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float f8[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            f8[j] = __uint_as_float(v[i * 8 + j]) + s_bias[c0 + i * 8 + j];
                            if (p.relu) f8[j] = fmaxf(f8[j], 0.f);
                        }
                        uint4 w4;
                        w4.x = pack_bf16x2(f8[0], f8[1]);
                        w4.y = pack_bf16x2(f8[2], f8[3]);
                        w4.z = pack_bf16x2(f8[4], f8[5]);
                        w4.w = pack_bf16x2(f8[6], f8[7]);
                        *reinterpret_cast<uint4*>(o + c0 + i * 8) = w4;
                    }
                }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();                        // neither CTA may free TMEM / exit while the pair's MMAs or barriers are live
    if (warp_idx == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ================================================================================================================
// K2 + K3 in one kernel (paired CTAs): the 3x3's bf16 result never leaves the SM.  Per CTA: the 3x3 accumulators
// (2 stages x 128 TMEM columns) are drained by the epilogue warps into a [128 x 128] bf16 K-major operand in shared
// memory (bias + ReLU = the folded bn3), which the leader multiplies (M = 256 over the pair, N = 256, K = 128) with the
// resident 1x1 weights (each CTA holds 128 of the 256 output-channel rows) into a third accumulator (256 columns); the
// same warps then add bias, the residual and optionally the nearest-upsampled low-resolution tensor (both read from
// global memory at the pixel's dense position) and store the 256-channel result.  Order on the tensor pipe:
// K2(t), K3(t-1), K2(t+1), ... so the epilogue of tile t overlaps K2(t+1).
// Warps: 0 A producer, 1 MMA issuer (+TMEM), 2 B producer, 3 W3 loader, 4..11 K3 drain (two per TMEM lane quarter: the
// lower four take the low half of the columns, the upper four the high half), 12..15 K2 drain (one per lane quarter).
// The two drains run on separate warps: the K3 drain is the long pole of a tile (global loads and stores), and with the K2
// drain ahead of it in the same warps it also paid that drain, the wait for the single operand buffer and a cluster-scope
// release behind its own output stores -- 37 % of the epilogue's time (profiles/r2_ncu_k3_fused_epilogue.txt).
// ================================================================================================================
constexpr int kSlab16 = 128 * kBlockK * 2;              // [128 rows x 64 k] bf16, 128-byte swizzled

// kRes / kUp: residual / upsample operand present (separate instantiations: no runtime predicates in the issue-bound epilogue)
template <bool kRes, bool kUp>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1) conv3x3_k3_pair_kernel(const __grid_constant__ Params p) {
    constexpr int BLOCK_N = 128;
    constexpr int kTmemCols = 512;                      // K2: 2 x 128 at [0,256); K3: 256 at [256,512)
    constexpr uint32_t kAcc3 = 256;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                    // 2 regions
    uint8_t* smem_b = smem_a + 2 * p.region_bytes;             // b_stages x 8 KiB
    uint8_t* smem_a2 = smem_b + p.b_stages * kPairBStage;      // K3's A operand: two K slabs
    uint8_t* smem_w3 = smem_a2 + 2 * kSlab16;                  // this CTA's 128 rows of W3: two K slabs
    uint8_t* smem_stage = smem_w3 + 2 * kSlab16;               // 8 warps x 2 private staging buffers of 2 KiB
    float* s_bias2 = reinterpret_cast<float*>(smem_stage + 2 * kSlab16);
    float* s_bias3 = s_bias2 + BLOCK_N;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias3 + 256);
    uint64_t* a_full = bars;                  // [2]   (leader)
    uint64_t* a_empty = bars + 2;             // [2]
    uint64_t* b_full = bars + 4;              // [kPairBStages] (leader)
    uint64_t* b_empty = b_full + kPairBStages;
    uint64_t* tmem_full_bar = b_empty + kPairBStages;   // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2]   (leader; 4 K2 drain warps of each CTA)
    uint64_t* a2_full = tmem_empty_bar + 2;             // (leader; 8 warps)
    uint64_t* a2_empty = a2_full + 1;
    uint64_t* acc3_full = a2_empty + 1;
    uint64_t* acc3_empty = acc3_full + 1;               // (leader; 16 warps)
    uint64_t* w3_bar = acc3_empty + 1;                  // (leader)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w3_bar + 1);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    for (int i = threadIdx.x; i < BLOCK_N; i += blockDim.x) s_bias2[i] = p.bias ? p.bias[i] : 0.f;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_bias3[i] = p.bias3 ? p.bias3[i] : 0.f;
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&p.map_a);
        tma_prefetch_desc(&p.map_b);
        tma_prefetch_desc(&p.map_w3);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 8);
        }
        for (int s = 0; s < kPairBStages; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        mbar_init(a2_full, 8);
        mbar_init(a2_empty, 1);
        mbar_init(acc3_full, 1);
        mbar_init(acc3_empty, 16);
        mbar_init(w3_bar, 1);
        fence_mbar_init();
    }
    if (warp_idx == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
    pdl_launch_dependents();

    const int halo = p.P + 1;

    if (warp_idx == 0) {
        // ===================== A producer =====================
        if (elect_one_sync()) {
            pdl_wait();
            int it = 0;
            bool ok = true;
            for (int tile = pair; tile < p.num_tiles && ok; tile += num_pairs) {
                const long long f0 = static_cast<long long>(p.P) + static_cast<long long>(tile) * kBM + rank * 128;
                const int row0 = static_cast<int>(f0 - halo);
                for (int slab = 0; slab < p.slabs; ++slab, ++it) {
                    const int buf = it & 1;
                    ok = mbar_wait(&a_empty[buf], ((it >> 1) & 1u) ^ 1u, p.err_word, 0x3901);
                    if (!ok) break;
                    if (leader) mbar_arrive_expect_tx(&a_full[buf], static_cast<uint32_t>(2 * p.num_boxes * p.box_rows * 128));
                    for (int b = 0; b < p.num_boxes; ++b)
                        tma_load_2d_2sm(smem_a + buf * p.region_bytes + b * p.box_rows * 128, &p.map_a, &a_full[buf],
                                        slab * kBlockK, row0 + b * p.box_rows);
                }
            }
        }
    } else if (warp_idx == 2) {
        // ===================== B producer =====================
        if (elect_one_sync()) {
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (int tile = pair; tile < p.num_tiles && ok; tile += num_pairs) {
                for (int slab = 0; slab < p.slabs && ok; ++slab) {
                    for (int tap = 0; tap < 9; ++tap) {
                        ok = mbar_wait(&b_empty[stage], phase ^ 1u, p.err_word, 0x3b01);
                        if (!ok) break;
                        if (leader) mbar_arrive_expect_tx(&b_full[stage], 2 * kPairBStage);
                        tma_load_2d_2sm(smem_b + stage * kPairBStage, &p.map_b, &b_full[stage], tap * p.cin + slab * kBlockK,
                                        static_cast<int>(rank) * 64);
                        if (++stage == p.b_stages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp_idx == 3) {
        // ===================== W3: this CTA's 128 output-channel rows, resident =====================
        if (elect_one_sync()) {
            if (leader) mbar_arrive_expect_tx(w3_bar, 4 * kSlab16);
            for (int slab = 0; slab < 2; ++slab)
                tma_load_2d_2sm(smem_w3 + slab * kSlab16, &p.map_w3, w3_bar, slab * kBlockK, static_cast<int>(rank) * 128);
        }
    } else if (warp_idx == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && elect_one_sync()) {
            constexpr uint32_t idesc2 = umma_idesc_bf16(256, BLOCK_N);
            constexpr uint32_t idesc3 = umma_idesc_bf16(256, 256);
            const uint32_t a2_base = smem_u32(smem_a2), w3_base = smem_u32(smem_w3);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0, a_it = 0;
            bool ok = mbar_wait(w3_bar, 0, p.err_word, 0x3a05);
            auto issue_k3 = [&](int j) -> bool {
                if (!mbar_wait(a2_full, static_cast<uint32_t>(j & 1), p.err_word, 0x3a06)) return false;
                if (!mbar_wait(acc3_empty, static_cast<uint32_t>(j & 1) ^ 1u, p.err_word, 0x3a07)) return false;
                tc_fence_after();
#pragma unroll
                for (int slab = 0; slab < 2; ++slab) {
                    const uint64_t a_desc = umma_desc_sw128(a2_base + slab * kSlab16);
                    const uint64_t b_desc = umma_desc_sw128(w3_base + slab * kSlab16);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc_mma_bf16_2sm(tmem_base + kAcc3, a_desc + 2u * k, b_desc + 2u * k, idesc3, (slab | k) != 0 ? 1u : 0u);
                }
                tc_commit_2sm(a2_empty);
                tc_commit_2sm(acc3_full);
                return true;
            };
            for (int tile = pair; tile < p.num_tiles && ok; tile += num_pairs, ++it) {
                const int acc = it & 1;
                ok = mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1u) ^ 1u, p.err_word, 0x3a01);
                if (!ok) break;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
                for (int slab = 0; slab < p.slabs && ok; ++slab, ++a_it) {
                    const int buf = a_it & 1;
                    ok = mbar_wait(&a_full[buf], (a_it >> 1) & 1u, p.err_word, 0x3a02);
                    if (!ok) break;
                    const uint32_t a_base = smem_u32(smem_a + buf * p.region_bytes);
                    for (int tap = 0; tap < 9; ++tap) {
                        ok = mbar_wait(&b_full[stage], phase, p.err_word, 0x3a03);
                        if (!ok) break;
                        tc_fence_after();
                        const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                        const uint32_t row_off = static_cast<uint32_t>(halo + dy * p.P + dx);
                        const uint64_t b_desc = umma_desc_sw128(smem_u32(smem_b + stage * kPairBStage));
                        const uint64_t a_desc = umma_desc_sw128(a_base + row_off * 128);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            tc_mma_bf16_2sm(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc2, (slab | tap | k) != 0 ? 1u : 0u);
                        tc_commit_2sm(&b_empty[stage]);
                        if (++stage == p.b_stages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    if (ok) tc_commit_2sm(&a_empty[buf]);
                }
                if (!ok) break;
                tc_commit_2sm(&tmem_full_bar[acc]);
                if (it > 0) ok = issue_k3(it - 1);
            }
            if (ok && it > 0) issue_k3(it - 1);
        }
    } else if (warp_idx >= 12) {
        // ===================== K2 drain warps: 3x3 accumulators -> bias + ReLU (folded bn3) -> bf16 K3 operand =====================
        // One warp per TMEM lane quarter, all 128 columns.  These warps never touch global memory, so the cluster-scope RELEASE
        // that hands the operand to the leader's MMA thread has nothing to wait for (in the K3 drain warps the same arrive sat
        // behind the previous tile's output stores: MEMBAR/ERRBAR, a fifth of the epilogue's time in the ncu source page).
        const int q = warp_idx & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const uint32_t a2_row = smem_u32(smem_a2) + row * 128;
        const uint32_t bias2_s = smem_u32(s_bias2);
        int it = 0;
        bool ok = true;
        for (int tile = pair; tile < p.num_tiles && ok; tile += num_pairs, ++it) {
            const int acc = it & 1;
            ok = mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1u, p.err_word, 0x3c01);
            if (!ok) break;
            ok = mbar_wait(a2_empty, static_cast<uint32_t>(it & 1) ^ 1u, p.err_word, 0x3c02);   // K3(it-1) has read the operand buffer
            if (!ok) break;
            tc_fence_after();
#pragma unroll
            for (int slab = 0; slab < 2; ++slab) {
                uint32_t v0[32], v1[32];
                const uint32_t t2 = lane_base + static_cast<uint32_t>(acc * BLOCK_N + slab * 64);
                tmem_ld_32x32(t2, v0);
                tmem_ld_32x32(t2 + 32, v1);
                tmem_ld_wait();
                if (slab == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_on_leader_relaxed(&tmem_empty_bar[acc]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t* vv = j < 4 ? &v0[j * 8] : &v1[(j - 4) * 8];
                    const uint32_t bs = bias2_s + static_cast<uint32_t>((slab * 64 + j * 8) * 4);
                    const float4 ba = lds128f(bs), bb = lds128f(bs + 16);
                    const unsigned long long a0 = f2_add(f2_pack(__uint_as_float(vv[0]), __uint_as_float(vv[1])), f2_pack(ba.x, ba.y));
                    const unsigned long long a1 = f2_add(f2_pack(__uint_as_float(vv[2]), __uint_as_float(vv[3])), f2_pack(ba.z, ba.w));
                    const unsigned long long a2 = f2_add(f2_pack(__uint_as_float(vv[4]), __uint_as_float(vv[5])), f2_pack(bb.x, bb.y));
                    const unsigned long long a3 = f2_add(f2_pack(__uint_as_float(vv[6]), __uint_as_float(vv[7])), f2_pack(bb.z, bb.w));
                    uint4 w4;
                    w4.x = pack_bf16x2_relu(f2_lo(a0), f2_hi(a0));
                    w4.y = pack_bf16x2_relu(f2_lo(a1), f2_hi(a1));
                    w4.z = pack_bf16x2_relu(f2_lo(a2), f2_hi(a2));
                    w4.w = pack_bf16x2_relu(f2_lo(a3), f2_hi(a3));
                    sts128(a2_row + slab * kSlab16 + ((j ^ (row & 7)) << 4), w4);
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_on_leader(a2_full);
        }
    } else if (warp_idx >= 4) {
        // ===================== K3 drain warps: 1x1 accumulators + bias + residual (+ upsampled) -> output =====================
        pdl_wait();
        const int q = warp_idx & 3;
        const int ch = (warp_idx - 4) >> 2;                 // column half this warp owns
        const int row = q * 32 + lane;
        const long long img_pos = static_cast<long long>(p.H + 1) * p.P;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        long long cur_low = 0;
        int cur_px = -1;                                     // dense pixel index of this lane's row in the tile being drained
        // warp-private staging: 2 buffers of [32 rows x 32 channels] bf16 (64-byte rows, 16-byte chunks swizzled so that both
        // the row-per-lane and the four-lanes-per-row access patterns are bank-conflict free); no CTA-level barriers
        const uint32_t wstage = smem_u32(smem_stage) + static_cast<uint32_t>(warp_idx - 4) * 4096u;
        const uint32_t bias3_s = smem_u32(s_bias3);
        auto sw = [](int r, int chunk) -> uint32_t { return static_cast<uint32_t>(r * 64 + ((chunk ^ ((r >> 1) & 3)) << 4)); };
        int it = 0;
        bool ok = true;

        // residual sub-slab `sub` (this warp's 32 rows x 32 channels) -> buffer sub & 1: cp.async, four lanes per row
        // rows this lane moves in the four-lanes-per-row passes (k-th pass: row k*8 + lane/4, chunk lane%4): element offset of
        // that row's pixel in the 256-channel tensors (+ this warp's channel half + chunk), or -1 for pads; set once per tile
        long long mv_off[4] = {-1, -1, -1, -1};
        auto fetch_res = [&](int sub) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = k * 32 + lane;
                const int r = idx >> 2, chunk = idx & 3;
                if (kRes && mv_off[k] >= 0) {
                    const __nv_bfloat16* src = p.res + mv_off[k] + sub * 32;
                    const uint32_t dst = wstage + (sub & 1) * 2048u + sw(r, chunk);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // where the rows of `tile` live in the dense tensors
        auto locate = [&](int tile) {
            const long long f = static_cast<long long>(tile) * kBM + rank * 128 + row;
            const long long n = f / img_pos;
            const int r = static_cast<int>(f - n * img_pos);
            const int y = r / p.P, x = r - y * p.P;
            const bool valid = (f < p.total_pos) && (x < p.W) && (y < p.H);
            cur_px = valid ? static_cast<int>((n * p.H + y) * p.W + x) : -1;
            cur_low = (n * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int px = __shfl_sync(0xffffffffu, cur_px, k * 8 + (lane >> 2));
                mv_off[k] = px >= 0 ? static_cast<long long>(px) * 256 + ch * 128 + (lane & 3) * 8 : -1;
            }
        };

        auto drain_k3 = [&](int j) -> bool {
            // tile j's 1x1 result, 32 channels at a time: the residual waits in the warp's staging buffer, each lane adds its
            // row in place, and the sub-slab leaves with 64 contiguous bytes per row (eight rows per store instruction)
            const bool valid = cur_px >= 0;
            const __nv_bfloat16* u = kUp ? p.up + cur_low * 256 + ch * 128 : nullptr;
            uint32_t ur2[2][16];                             // upsample operand, one phase ahead
            if (kUp && valid) {
                ldg_nc_v8(u, ur2[0]);
                ldg_nc_v8(u + 16, ur2[0] + 8);
            }
            if (!mbar_wait(acc3_full, static_cast<uint32_t>(j & 1), p.err_word, 0x3c03)) return false;
            tc_fence_after();
#pragma unroll
            for (int ph = 0; ph < 4; ++ph) {
                const uint32_t buf = wstage + (ph & 1) * 2048u;
                uint32_t v[32];
                const uint32_t* ur = ur2[ph & 1];
                if (kUp && valid && ph < 3) {
                    ldg_nc_v8(u + (ph + 1) * 32, ur2[(ph + 1) & 1]);
                    ldg_nc_v8(u + (ph + 1) * 32 + 16, ur2[(ph + 1) & 1] + 8);
                }
                tmem_ld_32x32(lane_base + kAcc3 + ch * 128 + ph * 32, v);
                if (ph < 3) asm volatile("cp.async.wait_group 1;" ::: "memory");
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();                                // every lane's share of this residual sub-slab has landed
                tmem_ld_wait();
                if (ph == 3) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_on_leader_relaxed(acc3_empty);
                }
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t addr = buf + sw(lane, i);
                        uint4 r4 = make_uint4(0u, 0u, 0u, 0u);
                        if (kRes) r4 = lds128(addr);
                        const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w};
                        const uint32_t bs = bias3_s + static_cast<uint32_t>((ch * 128 + ph * 32 + i * 8) * 4);
                        const float4 ba = lds128f(bs), bb = lds128f(bs + 16);
                        const float bl[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                        uint32_t o4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            // (acc + bias) + residual (+ upsampled), two channels per FADD2 -- the order of the two-kernel path
                            unsigned long long a = f2_add(f2_pack(__uint_as_float(v[i * 8 + 2 * e]), __uint_as_float(v[i * 8 + 2 * e + 1])),
                                                          f2_pack(bl[2 * e], bl[2 * e + 1]));
                            if (kRes) a = f2_add(a, f2_pack(bf16_lo_to_f32(rr[e]), bf16_hi_to_f32(rr[e])));
                            if (kUp) a = f2_add(a, f2_pack(bf16_lo_to_f32(ur[i * 4 + e]), bf16_hi_to_f32(ur[i * 4 + e])));
                            o4[e] = pack_bf16x2(f2_lo(a), f2_hi(a));
                        }
                        sts128(addr, make_uint4(o4[0], o4[1], o4[2], o4[3]));
                    }
                }
                __syncwarp();                                // the sub-slab is complete
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int idx = k * 32 + lane;
                    const int r = idx >> 2, chunk = idx & 3;
                    if (mv_off[k] >= 0) stg_v4(p.out3 + mv_off[k] + ph * 32, lds128(buf + sw(r, chunk)));
                }
                if (ph < 2) {
                    __syncwarp();                            // the buffer may be refilled
                    fetch_res(ph + 2);
                }
            }
            return true;
        };

        if (pair < p.num_tiles) {
            locate(pair);
            fetch_res(0);
            fetch_res(1);
        }
        for (int tile = pair; tile < p.num_tiles && ok; tile += num_pairs, ++it) {
            ok = drain_k3(it);
            if (!ok) break;
            // the first two residual sub-slabs of the NEXT tile fly while its 3x3 and 1x1 run
            if (tile + num_pairs < p.num_tiles) {
                locate(tile + num_pairs);
                __syncwarp();                                // every lane has read the staging buffers of the tile just drained
                fetch_res(0);
                fetch_res(1);
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp_idx == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
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
        set_last_error("conv3x3: cuTensorMapEncodeTiled failed: CUresult %d (cols %llu rows %llu box_rows %u)", (int)r,
                       (unsigned long long)cols, (unsigned long long)rows, box_rows);
        return HG_ERR_CUDA;
    }
    return HG_OK;
}

template <int BLOCK_N, bool kStats>
static int launch_variant(const Params& kp, int smem_bytes, cudaStream_t stream) {
    auto kern = conv3x3_kernel<BLOCK_N, kStats>;
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
    const int grid = kp.num_work < num_sms() ? kp.num_work : num_sms();
    HG_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(256), smem_bytes, stream, kp));
    return HG_OK;
}

template <int BLOCK_N>
static int launch(const Params& kp, int smem_bytes, cudaStream_t stream) {
    return kp.stats != nullptr ? launch_variant<BLOCK_N, true>(kp, smem_bytes, stream)
                               : launch_variant<BLOCK_N, false>(kp, smem_bytes, stream);
}

}  // namespace c3
}  // namespace hg

// Elements a halo-padded activation buffer needs: one leading zero row + n*(h+1)*(w+1) positions.
extern "C" int64_t hg_halo_padded_elems(int32_t n, int32_t h, int32_t w, int32_t c) {
    return (static_cast<int64_t>(n) * (h + 1) * (w + 1) + (w + 1)) * c;
}

extern "C" int hg_conv3x3_halo_bf16(const void* in_padded, const void* weight, const float* bias, void* out,
                                    unsigned int* err_word, float* stats, int32_t n, int32_t h, int32_t w, int32_t cin, int32_t cout,
                                    int32_t relu, void* stream) {
    using namespace hg;
    using namespace hg::c3;
    if (!in_padded || !weight || !out || n <= 0 || h <= 0 || w <= 0 || cin % 64 != 0 || cin <= 0 ||
        (cout != 64 && cout != 128) || w > 253) {
        set_last_error("hg_conv3x3_halo_bf16: bad arguments (cin %% 64 == 0, cout in {64,128}, w <= 253)");
        return HG_ERR_INVALID;
    }
    Params kp;
    memset(&kp, 0, sizeof(kp));
    kp.bias = bias;
    kp.out = static_cast<__nv_bfloat16*>(out);
    kp.err_word = err_word;
    kp.stats = stats;
    kp.H = h;
    kp.W = w;
    kp.P = w + 1;
    kp.NB = n;
    kp.cin = cin;
    kp.cout = cout;
    kp.slabs = cin / 64;
    kp.relu = relu;
    kp.total_pos = static_cast<long long>(n) * (h + 1) * kp.P;
    kp.num_tiles = static_cast<int>((kp.total_pos + kBM - 1) / kBM);
    const int region_rows = kBM + 2 * kp.P + 2;
    kp.num_boxes = (region_rows + 255) / 256;
    kp.box_rows = ((region_rows + kp.num_boxes - 1) / kp.num_boxes + 7) / 8 * 8;
    kp.region_bytes = kp.num_boxes * kp.box_rows * 128;
    // Small grids (the 16x16 .. 4x4 levels of the hourglass at training batch sizes are 37 / 11 / 4 tiles): split the
    // output channels over up to four CTAs per tile so that more SMs share the MMA work (a 256-position tile x 128
    // channels x K=1152 is ~5 us of one SM's tensor pipe) and each CTA streams a quarter of the weights.
    int block_n = cout;
    if (cout == 128) {
        const int sms = num_sms();
        if (kp.num_tiles * 4 <= sms) block_n = 32;
        else if (kp.num_tiles * 2 <= sms) block_n = 64;
    }
    static const bool no_split = getenv("HG_CONV3X3_NO_NSPLIT") != nullptr;
    if (no_split) block_n = cout;
    const bool use_pair = getenv("HG_CONV3X3_PAIR") != nullptr;      // read per call: the tests switch it
    if (use_pair && cout == 128 && stats == nullptr && kp.num_tiles >= num_sms() / 2) {
        // paired CTAs: each CTA stages its own 128 positions (+ halo) and half of every weight k-block
        const int rows = 128 + 2 * kp.P + 2;
        const int p_boxes = (rows + 255) / 256;
        const int p_box_rows = ((rows + p_boxes - 1) / p_boxes + 7) / 8 * 8;
        const int p_region = p_boxes * p_box_rows * 128;
        const int misc_p = cout * 4 + 1024;
        int stages = (kSmemLimit - 1024 - 2 * p_region - misc_p) / kPairBStage;
        if (stages > kPairBStages) stages = kPairBStages;
        if (stages >= 4) {
            kp.num_boxes = p_boxes;
            kp.box_rows = p_box_rows;
            kp.region_bytes = p_region;
            kp.b_stages = stages;
            kp.n_split = 1;
            kp.num_work = kp.num_tiles;
            const int smem_pair = 1024 + 2 * kp.region_bytes + stages * kPairBStage + misc_p;
            const uint64_t rows_total = static_cast<uint64_t>(kp.total_pos) + kp.P;
            int rc2;
            if ((rc2 = make_map(&kp.map_a, in_padded, cin, rows_total, kp.box_rows)) != HG_OK) return rc2;
            if ((rc2 = make_map(&kp.map_b, weight, 9ull * cin, cout, 64)) != HG_OK) return rc2;
            static std::mutex mu;
            static unsigned long long done_mask = 0;
            int dev = 0;
            HG_CUDA_OK(cudaGetDevice(&dev));
            {
                std::lock_guard<std::mutex> lock(mu);
                if (dev >= 64 || !(done_mask >> dev & 1ull)) {
                    HG_CUDA_OK(cudaFuncSetAttribute(conv3x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
                    if (dev < 64) done_mask |= 1ull << dev;
                }
            }
            const int pairs = kp.num_tiles < num_sms() / 2 ? kp.num_tiles : num_sms() / 2;
            HG_CUDA_OK(launch_kernel(conv3x3_pair_kernel, dim3(2 * pairs), dim3(256), smem_pair, static_cast<cudaStream_t>(stream), kp));
            return HG_OK;
        }
    }
    kp.n_split = cout / block_n;
    kp.num_work = kp.num_tiles * kp.n_split;
    const int b_stage = block_n * 128;
    const int misc = (stats ? 3 : 1) * cout * 4 + 512;
    int b_stages = (kSmemLimit - 1024 - 2 * kp.region_bytes - misc) / b_stage;
    if (b_stages > kMaxBStages) b_stages = kMaxBStages;
    if (b_stages < 2) {
        set_last_error("hg_conv3x3_halo_bf16: image too wide for the shared-memory budget (w=%d)", w);
        return HG_ERR_INVALID;
    }
    kp.b_stages = b_stages;
    const int smem_bytes = 1024 + 2 * kp.region_bytes + b_stages * b_stage + misc;
    const uint64_t rows = static_cast<uint64_t>(kp.total_pos) + kp.P;       // incl. the leading zero row
    int rc;
    if ((rc = make_map(&kp.map_a, in_padded, cin, rows, kp.box_rows)) != HG_OK) return rc;
    if ((rc = make_map(&kp.map_b, weight, 9ull * cin, cout, block_n)) != HG_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (block_n) {
        case 32: return launch<32>(kp, smem_bytes, st);
        case 64: return launch<64>(kp, smem_bytes, st);
        default: return launch<128>(kp, smem_bytes, st);
    }
}

// ---- K2 + K3 fused (paired CTAs) ----
static int k3_pair_geometry(int32_t n, int32_t h, int32_t w, int* boxes, int* box_rows, int* region, int* stages, int* tiles) {
    using namespace hg;
    using namespace hg::c3;
    if (n <= 0 || h <= 0 || w <= 0 || w > 253) return 0;
    const int P = w + 1;
    const long long total_pos = static_cast<long long>(n) * (h + 1) * P;
    const long long num_tiles = (total_pos + kBM - 1) / kBM;
    if (num_tiles > 0x7fffffffLL / kBM || static_cast<long long>(n) * h * w > 0x7fffffffLL) return 0;     // 32-bit pixel table
    const int rows = 128 + 2 * P + 2;
    const int nb = (rows + 255) / 256;
    const int br = ((rows + nb - 1) / nb + 7) / 8 * 8;
    const int reg = nb * br * 128;
    const int misc = (128 + 256 + 256) * 4 + 1024;
    int st = (kSmemLimit - 1024 - 2 * reg - 6 * kSlab16 - misc) / kPairBStage;
    if (st > kPairBStages) st = kPairBStages;
    if (st < 4) return 0;
    *boxes = nb; *box_rows = br; *region = reg; *stages = st; *tiles = static_cast<int>(num_tiles);
    return 1;
}

extern "C" int hg_conv3x3_k3_fusable(int32_t n, int32_t h, int32_t w) {
    int a, b, c, d, tiles;
    if (!k3_pair_geometry(n, h, w, &a, &b, &c, &d, &tiles)) return 0;
    return tiles >= hg::num_sms() / 2 ? 1 : 0;       // at least one tile per CTA pair
}

extern "C" int hg_conv3x3_k3_fused_bf16(const void* in_padded, const void* w2, const float* b2, const void* w3, const float* b3,
                                        const void* residual, const void* up_low, void* out, unsigned int* err_word, int32_t n,
                                        int32_t h, int32_t w, void* stream) {
    using namespace hg;
    using namespace hg::c3;
    Params kp;
    memset(&kp, 0, sizeof(kp));
    int tiles = 0;
    if (!in_padded || !w2 || !w3 || !out || out == residual ||
        !k3_pair_geometry(n, h, w, &kp.num_boxes, &kp.box_rows, &kp.region_bytes, &kp.b_stages, &tiles) ||
        (up_low != nullptr && ((h | w) & 1))) {
        set_last_error("hg_conv3x3_k3_fused_bf16: bad arguments (w <= 253, even h and w with up_low, out != residual)");
        return HG_ERR_INVALID;
    }
    kp.bias = b2;
    kp.bias3 = b3;
    kp.res = static_cast<const __nv_bfloat16*>(residual);
    kp.up = static_cast<const __nv_bfloat16*>(up_low);
    kp.out3 = static_cast<__nv_bfloat16*>(out);
    kp.err_word = err_word;
    kp.H = h;
    kp.W = w;
    kp.P = w + 1;
    kp.NB = n;
    kp.cin = 128;
    kp.cout = 128;
    kp.slabs = 2;
    kp.relu = 1;
    kp.total_pos = static_cast<long long>(n) * (h + 1) * kp.P;
    kp.num_tiles = tiles;
    kp.n_split = 1;
    kp.num_work = tiles;
    const uint64_t rows_total = static_cast<uint64_t>(kp.total_pos) + kp.P;
    int rc;
    if ((rc = make_map(&kp.map_a, in_padded, 128, rows_total, kp.box_rows)) != HG_OK) return rc;
    if ((rc = make_map(&kp.map_b, w2, 9ull * 128, 128, 64)) != HG_OK) return rc;
    if ((rc = make_map(&kp.map_w3, w3, 128, 256, 128)) != HG_OK) return rc;
    static std::mutex mu;
    static unsigned long long done_mask = 0;
    int dev = 0;
    HG_CUDA_OK(cudaGetDevice(&dev));
    void (*kern)(Params) = residual ? (up_low ? conv3x3_k3_pair_kernel<true, true> : conv3x3_k3_pair_kernel<true, false>)
                                    : (up_low ? conv3x3_k3_pair_kernel<false, true> : conv3x3_k3_pair_kernel<false, false>);
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 64 || !(done_mask >> dev & 1ull)) {
            HG_CUDA_OK(cudaFuncSetAttribute(conv3x3_k3_pair_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
            HG_CUDA_OK(cudaFuncSetAttribute(conv3x3_k3_pair_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
            HG_CUDA_OK(cudaFuncSetAttribute(conv3x3_k3_pair_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
            HG_CUDA_OK(cudaFuncSetAttribute(conv3x3_k3_pair_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
            if (dev < 64) done_mask |= 1ull << dev;
        }
    }
    const int smem_bytes = 1024 + 2 * kp.region_bytes + kp.b_stages * kPairBStage + 6 * kSlab16 + (128 + 256 + 256) * 4 + 1024;
    const int pairs = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
    HG_CUDA_OK(launch_kernel(kern, dim3(2 * pairs), dim3(512), smem_bytes, static_cast<cudaStream_t>(stream), kp));
    return HG_OK;
}
