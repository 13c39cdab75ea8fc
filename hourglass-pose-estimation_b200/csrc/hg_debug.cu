// Hardware probes used while developing the kernels (exported for tests/gpu_probe.py; not on the product path).
//
// hg_debug_shifted_desc: does a 128-byte-swizzled K-major UMMA operand work when the descriptor's start
// address is shifted by whole 128-byte rows (not 1024-byte aligned)?  This is what an implicit-GEMM 3x3
// convolution needs to read all nine taps out of ONE halo'd shared-memory tile.
#include "hg_common.cuh"
#include "../../include/hg_api.h"

#include <cudaTypedefs.h>
#include <cstring>

namespace hg {

constexpr int kProbeRows = 192;     // rows of A resident in smem
constexpr int kProbeN = 64;

struct ProbeParams {
    CUtensorMap map_a;   // (64, kProbeRows)
    CUtensorMap map_b;   // (64, kProbeN)
    float* out;          // [num_shifts][2 variants][128][kProbeN]
    int shifts[16];
    int num_shifts;
};

__global__ void __launch_bounds__(128, 1) probe_shifted_desc_kernel(const __grid_constant__ ProbeParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                 // kProbeRows x 128 B
    uint8_t* smem_b = smem + kProbeRows * 128;              // 64 x 128 B
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_ptr, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_load, kProbeRows * 128 + kProbeN * 128);
        tma_load_2d(smem_a, &p.map_a, &bar_load, 0, 0);
        tma_load_2d(smem_b, &p.map_b, &bar_load, 0, 0);
    }
    mbar_wait(&bar_load, 0, nullptr, 0);
    uint32_t phase = 0;
    for (int si = 0; si < p.num_shifts; ++si) {
        for (int variant = 0; variant < 2; ++variant) {
            if (threadIdx.x == 0) {
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem_a) + p.shifts[si] * 128;
                uint64_t a_desc = umma_desc_sw128(a_addr);
                if (variant == 1) a_desc |= static_cast<uint64_t>((a_addr >> 7) & 7u) << 49;   // base_offset field
                const uint64_t b_desc = umma_desc_sw128(smem_u32(smem_b));
                constexpr uint32_t idesc = umma_idesc_bf16(128, kProbeN);
                for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base, a_desc + 2u * k, b_desc + 2u * k, idesc, k ? 1u : 0u);
                tc_commit(&bar_mma);
            }
            mbar_wait(&bar_mma, phase, nullptr, 0);
            phase ^= 1u;
            tc_fence_after();
            uint32_t v[32];
            float* o = p.out + (static_cast<size_t>(si * 2 + variant) * 128 + warp * 32 + lane) * kProbeN;
            for (int h = 0; h < 2; ++h) {
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + h * 32, v);
                tmem_ld_wait();
                for (int i = 0; i < 32; ++i) o[h * 32 + i] = __uint_as_float(v[i]);
            }
            tc_fence_before();
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

}  // namespace hg

// a: bf16 [192][64], b: bf16 [64][64], out: fp32 [num_shifts][2][128][64]
extern "C" int hg_debug_shifted_desc(const void* a, const void* b, float* out, const int32_t* shifts, int32_t num_shifts,
                                     void* stream) {
    using namespace hg;
    if (num_shifts <= 0 || num_shifts > 16) return HG_ERR_INVALID;
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    HG_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qres));
    auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fnp);
    ProbeParams p;
    memset(&p, 0, sizeof(p));
    auto mk = [&](CUtensorMap* m, const void* ptr, uint64_t rows) {
        cuuint64_t gdim[2] = {64, rows};
        cuuint64_t gstr[1] = {128};
        cuuint32_t box[2] = {64, static_cast<cuuint32_t>(rows)};
        cuuint32_t estr[2] = {1, 1};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    if (mk(&p.map_a, a, kProbeRows) != CUDA_SUCCESS || mk(&p.map_b, b, kProbeN) != CUDA_SUCCESS) {
        set_last_error("hg_debug_shifted_desc: tensor map encode failed");
        return HG_ERR_CUDA;
    }
    p.out = out;
    p.num_shifts = num_shifts;
    for (int i = 0; i < num_shifts; ++i) {
        if (shifts[i] < 0 || shifts[i] + 128 > kProbeRows) return HG_ERR_INVALID;
        p.shifts[i] = shifts[i];
    }
    const int smem = 1024 + kProbeRows * 128 + kProbeN * 128;
    probe_shifted_desc_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
    HG_CUDA_OK(cudaGetLastError());
    return HG_OK;
}
