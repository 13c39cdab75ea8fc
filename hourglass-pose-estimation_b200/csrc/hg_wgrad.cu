// Weight gradient of a convolution (the `conv.weight.grad` autograd produces in the reference's
// loss.backward(), src/runner/trainer.py:97-98) as a split-K tcgen05 GEMM over NHWC bf16 tensors:
//
//   dW[co][tap][ci] += sum over positions f of  dOut[f][co] * Z[f + off(tap)][ci]
//
// The contraction runs over PIXELS, which is the slow dimension of both NHWC operands, so neither is
// K-major.  Instead of transposing, both operands are described to the tensor core as **MN-major**
// (instruction-descriptor bits 15/16): a TMA box of [64 channels x R rows] with the 128-byte swizzle
// IS the canonical MN-major SW128 layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -- 64
// contiguous M (or N) elements per 128-byte row, one row per k, 8-row groups 1024 B apart (SBO),
// 64-channel blocks `LBO` bytes apart.  A K=16 step of the MMA is 16 rows = +2048 B on the start
// address.
//
// 3x3 mode (taps = 9): both tensors are in the halo-padded layout of hg_conv3x3.cu with the same
// geometry, so tap (dy,dx) is the constant row offset dy*P+dx of the Z operand; the zero pads of dOut
// kill the contributions of pad positions and the zero pads of Z are the convolution's padding.  One
// CTA owns the three taps of one filter row: it loads Z rows [r + dy*P - 1, r + dy*P + 71) once per
// k-block and the three taps are row-shifted descriptors over that tile.
//
// Each CTA accumulates its K-chunk in TMEM ([128 x N] fp32 per 128 output channels per tap, <= 512
// columns) and adds it into dW with vectorised red.global.add -- dW is the step's zero-initialised flat
// gradient buffer, kept in the GEMM-natural [co][tap][ci] order (= torch channels_last), so every lane
// writes whole 128-byte lines.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 4..7 = epilogue.
#include "hg_common.cuh"
#include "../../include/hg_api.h"

#include <cudaTypedefs.h>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace hg {
namespace wg {

constexpr int kBlockK = 64;                   // rows (positions) per k-block
constexpr int kABoxBytes = kBlockK * 128;     // [64 rows x 64 channels] bf16
constexpr int kMaxStages = 8;
constexpr int kSmemLimit = 232448;

struct Params {
    CUtensorMap map_a;      // (c, rows) over dOut, box (64, 64)
    CUtensorMap map_b;      // (c, rows) over Z, box (64, b_rows)
    float* dw;
    unsigned int* err_word;
    long long rows_total;   // rows of A to contract over (starting at row 0)
    int rows_per_chunk;     // multiple of 64
    int num_chunks;
    int tap_groups;         // 1 (1x1) or 3 (3x3: one filter row per CTA)
    int taps_per_cta;       // 1 or 3
    int P;                  // halo row pitch in positions (w+1); 0 in 1x1 mode
    int m_halves;           // 128-channel halves of dOut (1 or 2)
    int n_blocks;           // 64-channel blocks of Z (1..4)
    int b_rows;             // rows per Z box: 64 (1x1) or 72 (3x3: 1 + 64 + 1, rounded to 8)
    int stages;
    int ld;                 // floats between consecutive co rows of dW
    int tap_stride;         // floats between taps inside a dW row
    int co_first, co_valid, ci_valid;      // rows [co_first, co_valid) of dout^T.z are written, to dW rows 0..
    int vec4;               // dW rows / taps are 16-byte aligned: use red.v4
    int tmem_cols;
    uint32_t lbo_a, sbo_a, lbo_b, sbo_b;   // MN-major descriptor byte offsets
};

// MN-major operand, 128-byte swizzle
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32) | (static_cast<uint64_t>(1) << 46) |
           (static_cast<uint64_t>(2) << 61);
}
// kind::f16: D=f32, A=B=bf16, BOTH MN-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float* p, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

__global__ void __launch_bounds__(256, 1) wgrad_kernel(const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = 2 * p.m_halves * kABoxBytes;
    const int b_box_bytes = p.b_rows * 128;
    const int stage_bytes = a_bytes + p.n_blocks * b_box_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kMaxStages;
    uint64_t* tmem_full_bar = bars + 2 * kMaxStages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&p.map_a);
        tma_prefetch_desc(&p.map_b);
        for (int s = 0; s < kMaxStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        fence_mbar_init();
    }
    if (warp_idx == 1) tmem_alloc(tmem_ptr_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);   // broadcast: provably warp-uniform for the issue loop
    pdl_launch_dependents();

    const int group = blockIdx.x % p.tap_groups;           // filter row (3x3) or 0
    const int chunk = blockIdx.x / p.tap_groups;
    const long long r0 = static_cast<long long>(chunk) * p.rows_per_chunk;
    long long r1 = r0 + p.rows_per_chunk;
    if (r1 > p.rows_total) r1 = p.rows_total;
    const int num_kb = static_cast<int>((r1 - r0 + kBlockK - 1) / kBlockK);
    const int N = p.n_blocks * 64;

    if (warp_idx == 0) {
        if (elect_one_sync()) {
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            const long long b_shift = p.tap_groups == 3 ? static_cast<long long>(group - 1) * p.P - 1 : 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                if (!mbar_wait(&empty_bar[stage], phase ^ 1u, p.err_word, 0x5101)) break;
                mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes));
                uint8_t* sa = smem + stage * stage_bytes;
                uint8_t* sb = sa + a_bytes;
                const int row = static_cast<int>(r0 + static_cast<long long>(kb) * kBlockK);
                for (int m = 0; m < 2 * p.m_halves; ++m)
                    tma_load_2d(sa + m * kABoxBytes, &p.map_a, &full_bar[stage], m * 64, row);
                for (int n = 0; n < p.n_blocks; ++n)
                    tma_load_2d(sb + n * b_box_bytes, &p.map_b, &full_bar[stage], n * 64, row + static_cast<int>(b_shift));
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp_idx == 1) {
        if (elect_one_sync()) {
            const uint32_t idesc = umma_idesc_bf16_mn(128, N);
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (int kb = 0; kb < num_kb && ok; ++kb) {
                ok = mbar_wait(&full_bar[stage], phase, p.err_word, 0x5201);
                if (!ok) break;
                tc_fence_after();
                const uint32_t a_base = smem_u32(smem + stage * stage_bytes);
                const uint32_t b_base = a_base + a_bytes;
                for (int t = 0; t < p.taps_per_cta; ++t) {
                    const uint32_t shift = p.tap_groups == 3 ? static_cast<uint32_t>(t) : 0u;
                    for (int h = 0; h < p.m_halves; ++h) {
                        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>((h * p.taps_per_cta + t) * N);
#pragma unroll
                        for (int ks = 0; ks < kBlockK / 16; ++ks) {
                            const uint64_t a_desc =
                                umma_desc_mn_sw128(a_base + h * 2 * kABoxBytes + ks * 16 * 128, p.lbo_a, p.sbo_a);
                            const uint64_t b_desc = umma_desc_mn_sw128(b_base + (shift + ks * 16) * 128, p.lbo_b, p.sbo_b);
                            tc_mma_bf16(d_tmem, a_desc, b_desc, idesc, (kb | ks) != 0 ? 1u : 0u);
                        }
                    }
                }
                tc_commit(&empty_bar[stage]);
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (ok) tc_commit(tmem_full_bar);
        }
    } else if (warp_idx >= 4) {
        const int q = warp_idx & 3;
        if (num_kb > 0 && mbar_wait(tmem_full_bar, 0, p.err_word, 0x5401)) {
            tc_fence_after();
            for (int h = 0; h < p.m_halves; ++h) {
                const int co = h * 128 + q * 32 + lane;
                for (int t = 0; t < p.taps_per_cta; ++t) {
                    const int tap = group * p.taps_per_cta + t;
                    float* dst = p.dw + static_cast<long long>(co - p.co_first) * p.ld + static_cast<long long>(tap) * p.tap_stride;
                    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                           static_cast<uint32_t>((h * p.taps_per_cta + t) * N);
                    for (int c0 = 0; c0 < N; c0 += 32) {
                        if (c0 >= p.ci_valid) break;          // warp-uniform
                        uint32_t v[32];
                        tmem_ld_32x32(t_row + c0, v);
                        tmem_ld_wait();
                        if (co >= p.co_first && co < p.co_valid) {
                            if (p.vec4) {
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    if (c0 + i * 4 < p.ci_valid)
                                        red_add_v4(dst + c0 + i * 4, __uint_as_float(v[i * 4]), __uint_as_float(v[i * 4 + 1]),
                                                   __uint_as_float(v[i * 4 + 2]), __uint_as_float(v[i * 4 + 3]));
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i)
                                    if (c0 + i < p.ci_valid) red_add(dst + c0 + i, __uint_as_float(v[i]));
                            }
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
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
        set_last_error("wgrad: cuTensorMapEncodeTiled failed: CUresult %d (cols %llu rows %llu box_rows %u)", (int)r,
                       (unsigned long long)cols, (unsigned long long)rows, box_rows);
        return HG_ERR_CUDA;
    }
    return HG_OK;
}

}  // namespace wg
}  // namespace hg

extern "C" int hg_wgrad_bf16(const void* dout, const void* z, float* dw, unsigned int* err_word, int64_t rows, int32_t co,
                             int32_t co_first, int32_t co_valid, int32_t ci, int32_t ci_valid, int32_t taps, int32_t halo_pitch, int32_t ld,
                             int32_t tap_stride, int32_t max_ctas_arg, void* stream) {
    using namespace hg;
    using namespace hg::wg;
    if (!dout || !z || !dw || rows <= 0 || rows > 0x7fffff00LL || co <= 0 || co > 256 || ci <= 0 || ci > 256 ||
        co % 8 != 0 || ci % 64 != 0 || (taps != 1 && taps != 9) || (taps == 9 && halo_pitch < 2) || ci_valid <= 0 ||
        ci_valid > ci || co_valid <= 0 || co_valid > co || co_first < 0 || co_first >= co_valid) {
        set_last_error("hg_wgrad_bf16: bad arguments (co<=256 multiple of 8, ci<=256 multiple of 64, taps 1|9)");
        return HG_ERR_INVALID;
    }
    Params kp;
    memset(&kp, 0, sizeof(kp));
    kp.dw = dw;
    kp.err_word = err_word;
    kp.rows_total = rows;
    kp.tap_groups = taps == 9 ? 3 : 1;
    kp.taps_per_cta = taps == 9 ? 3 : 1;
    kp.P = taps == 9 ? halo_pitch : 0;
    kp.m_halves = co > 128 ? 2 : 1;
    kp.n_blocks = ci / 64;
    kp.b_rows = taps == 9 ? 72 : 64;
    kp.ld = ld;
    kp.tap_stride = tap_stride;
    kp.co_first = co_first;
    kp.co_valid = co_valid;
    kp.ci_valid = ci_valid;
    kp.vec4 = (ld % 4 == 0 && tap_stride % 4 == 0 && ci_valid % 4 == 0 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0) ? 1 : 0;
    const int cols = kp.m_halves * kp.taps_per_cta * ci;
    if (cols > 512) {
        set_last_error("hg_wgrad_bf16: %d TMEM columns needed (co=%d ci=%d taps=%d) > 512", cols, co, ci, taps);
        return HG_ERR_INVALID;
    }
    kp.tmem_cols = 32;
    while (kp.tmem_cols < cols) kp.tmem_cols <<= 1;
    const int stage_bytes = 2 * kp.m_halves * kABoxBytes + kp.n_blocks * kp.b_rows * 128;
    int stages = (kSmemLimit - 1024 - 512) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) {
        set_last_error("hg_wgrad_bf16: shared-memory budget exhausted");
        return HG_ERR_INVALID;
    }
    kp.stages = stages;
    const int smem_bytes = 1024 + stages * stage_bytes + 512;
    // split K: about one CTA per SM, but never less than 8 k-blocks per CTA (bounds the red.add traffic)
    const long long total_kb = (rows + kBlockK - 1) / kBlockK;
    // Weight gradients are leaves of the backward DAG: they run beside the critical chain, so they are confined to about
    // a quarter of the SMs (36 CTAs; HG_WGRAD_CTAS overrides) -- which also quarters the fp32 red.add traffic, every CTA
    // adding a full [cout x cin] tile at its end.  Measured on B200 (training step, batch 32, 8 streams): 24.9 / 24.5 /
    // 24.1 / 24.0 / 23.8 / 24.1 / 25.0 ms for 148 / 96 / 72 / 48 / 36 / 24 / 16 CTAs.
    static const int env_ctas = getenv("HG_WGRAD_CTAS") ? atoi(getenv("HG_WGRAD_CTAS")) : 36;
    // max_ctas_arg > 0 overrides; 1 = no split along the pixels at all: every dW element is then produced by ONE CTA in
    // one fixed accumulation order (deterministic, for validation runs; slow)
    const int max_ctas = max_ctas_arg > 0 ? max_ctas_arg : env_ctas;
    constexpr int min_kb = 8;
    long long chunks = (max_ctas > 0 ? (max_ctas < num_sms() ? max_ctas : num_sms()) : num_sms()) / kp.tap_groups;
    if (chunks > total_kb / min_kb) chunks = total_kb / min_kb;
    if (chunks < 1) chunks = 1;
    const long long kb_per_chunk = (total_kb + chunks - 1) / chunks;
    kp.rows_per_chunk = static_cast<int>(kb_per_chunk * kBlockK);
    kp.num_chunks = static_cast<int>((total_kb + kb_per_chunk - 1) / kb_per_chunk);
    // canonical MN-major SW128: LBO = distance between 64-channel blocks, SBO = distance between 8-row groups
    kp.lbo_a = kABoxBytes;
    kp.sbo_a = 1024;
    kp.lbo_b = kp.b_rows * 128;
    kp.sbo_b = 1024;
    int rc;
    if ((rc = make_map(&kp.map_a, dout, co, static_cast<uint64_t>(rows), 64)) != HG_OK) return rc;
    if ((rc = make_map(&kp.map_b, z, ci, static_cast<uint64_t>(rows), kp.b_rows)) != HG_OK) return rc;

    static std::mutex mu;
    static unsigned long long done_mask = 0;
    int dev = 0;
    HG_CUDA_OK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 64 || !(done_mask >> dev & 1ull)) {
            HG_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
            if (dev < 64) done_mask |= 1ull << dev;
        }
    }
    HG_CUDA_OK(launch_kernel(wgrad_kernel, dim3(kp.num_chunks * kp.tap_groups), dim3(256), smem_bytes,
                             static_cast<cudaStream_t>(stream), kp));
    return HG_OK;
}
