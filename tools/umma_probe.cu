// tools/umma_probe.cu — bring-up probe for the tcgen05 path (development tool, not product code).
// Answers, on a real B200, the questions the conv kernel design depends on:
//   1. are my instruction / shared-memory descriptors right (M=128, N=32, K-major bf16, SWIZZLE_128B)?
//   2. can the A operand start at an arbitrary ROW of a swizzled tile (tap shift = +k rows), and does the
//      descriptor's base_offset field have to carry (row & 7)?
//   3. same for the un-swizzled "interleaved" layout (row stride 16 B inside a K-chunk plane);
//   4. does a 3-D TMA box with negative / past-the-end row coordinates and a partial channel box zero-fill
//      the way Conv1d zero padding needs, and does the resulting tile feed the MMA directly?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int layout_type, int base_offset) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                      // version = 1 (Blackwell)
    d |= (uint64_t)(base_offset & 7) << 49;
    d |= (uint64_t)(layout_type & 7) << 61;
    return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int N_ = 32;
constexpr int ROWS = 160;          // A tile rows available (128 + max shift + slack)

// layout: 0 = SW128 (row-major 128B rows, 16B chunk index XOR (row&7)); 1 = interleaved (no swizzle): [kchunk][row][8]
__device__ __forceinline__ uint32_t a_off(int layout, int row, int col, int rows_total) {
    if (layout == 0) return row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1));
    return (col >> 3) * (rows_total * 16) + row * 16 + ((col & 7) << 1);
}

extern "C" __global__ void __launch_bounds__(128) probe_manual(const __nv_bfloat16 *A, const __nv_bfloat16 *Bm, float *D,
                                                               int shift, int layout, int bo_mode)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                       // ROWS * 128 B
    uint8_t *sB = smem + ROWS * 128;          // N_ * 128 B   (ROWS*128 = 20480, multiple of 1024)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < ROWS * 64; i += 128) {
        int r = i / 64, c = i % 64;
        *(__nv_bfloat16 *)(sA + a_off(layout, r, c, ROWS)) = A[i];
    }
    for (int i = tid; i < N_ * 64; i += 128) {
        int r = i / 64, c = i % 64;
        *(__nv_bfloat16 *)(sB + a_off(layout, r, c, N_)) = Bm[i];
    }
    if (tid == 0) mbar_init(&bar, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    // idesc: c=f32 (1<<4), a=bf16 (1<<7), b=bf16 (1<<10), K-major both, N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N_ >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (tid == 0) {
        for (int k = 0; k < 4; ++k) {
            uint64_t da, db;
            if (layout == 0) {
                const uint32_t a_addr = smem_u32(sA) + shift * 128 + k * 32;
                const int bo = (bo_mode == 0) ? 0 : (shift & 7);
                da = make_desc(a_addr, 16, 1024, 2, bo);
                db = make_desc(smem_u32(sB) + k * 32, 16, 1024, 2, 0);
            } else {
                da = make_desc(smem_u32(sA) + shift * 16 + k * 2 * (ROWS * 16), ROWS * 16, 128, 0, 0);
                db = make_desc(smem_u32(sB) + k * 2 * (N_ * 16), N_ * 16, 128, 0, 0);
            }
            umma_bf16(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[tid * N_ + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32));
}

// ---- TMA-fed conv tile: X [B][L][C] bf16 (NLC), 15 taps, Cin=C (<64, zero-filled by TMA), Cout = 32 --------------
constexpr int KS = 15, PADL = 7, TROWS = 128 + KS - 1;    // 142
extern "C" __global__ void __launch_bounds__(128) probe_tma(const __grid_constant__ CUtensorMap tmap, const __nv_bfloat16 *W /*[KS][32][64]*/,
                                                            float *D, int l0, int b, int bo_mode)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                          // 142 rows * 128 B = 18176 -> pad to 18432
    uint8_t *sB = smem + 18432;                  // KS * 32 * 128 B = 61440
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < KS * N_ * 64; i += 128) {
        int t = i / (N_ * 64), r = (i / 64) % N_, c = i % 64;
        *(__nv_bfloat16 *)(sB + t * (N_ * 128) + a_off(0, r, c, N_)) = W[i];
    }
    if (tid == 0) { mbar_init(&bar_tma, 1); mbar_init(&bar_mma, 1); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N_ >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (tid == 0) {
        mbar_expect_tx(&bar_tma, TROWS * 128);
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(sA)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&bar_tma)), "r"(0), "r"(l0 - PADL), "r"(b) : "memory");
        mbar_wait(&bar_tma, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int t = 0; t < KS; ++t) {
            for (int k = 0; k < 2; ++k) {        // only the first 32 channels can be non-zero (C=24): 2 K-steps
                const uint32_t a_addr = smem_u32(sA) + t * 128 + k * 32;
                const int bo = (bo_mode == 0) ? 0 : (t & 7);
                uint64_t da = make_desc(a_addr, 16, 1024, 2, bo);
                uint64_t db = make_desc(smem_u32(sB) + t * (N_ * 128) + k * 32, 16, 1024, 2, 0);
                umma_bf16(tmem_base, da, db, idesc, (t | k) ? 1u : 0u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    mbar_wait(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[tid * N_ + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32));
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main()
{
    srand(1);
    std::vector<float> Af(ROWS * 64), Bf(N_ * 64);
    std::vector<__nv_bfloat16> Ah(ROWS * 64), Bh(N_ * 64);
    for (size_t i = 0; i < Af.size(); ++i) { Af[i] = bf((rand() % 2001 - 1000) / 1000.f); Ah[i] = __float2bfloat16(Af[i]); }
    for (size_t i = 0; i < Bf.size(); ++i) { Bf[i] = bf((rand() % 2001 - 1000) / 1000.f); Bh[i] = __float2bfloat16(Bf[i]); }
    __nv_bfloat16 *dA, *dB; float *dD;
    CK(cudaMalloc(&dA, Ah.size() * 2)); CK(cudaMalloc(&dB, Bh.size() * 2)); CK(cudaMalloc(&dD, 128 * N_ * 4));
    CK(cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice));
    const int smem_bytes = ROWS * 128 + N_ * 128 + 2048;
    CK(cudaFuncSetAttribute(probe_manual, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    std::vector<float> Dh(128 * N_);
    const int shifts[] = {0, 1, 2, 3, 7, 8, 9, 14};
    for (int layout = 0; layout < 2; ++layout)
        for (int bo_mode = 0; bo_mode < (layout == 0 ? 2 : 1); ++bo_mode)
            for (int s : shifts) {
                CK(cudaMemset(dD, 0, 128 * N_ * 4));
                probe_manual<<<1, 128, smem_bytes>>>(dA, dB, dD, s, layout, bo_mode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("manual layout=%d bo_mode=%d shift=%d: CUDA error %s\n", layout, bo_mode, s, cudaGetErrorString(e)); return 2; }
                CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
                double maxerr = 0;
                for (int i = 0; i < 128; ++i)
                    for (int j = 0; j < N_; ++j) {
                        double ref = 0;
                        for (int k = 0; k < 64; ++k) ref += (double)Af[(i + s) * 64 + k] * Bf[j * 64 + k];
                        maxerr = fmax(maxerr, fabs(ref - Dh[i * N_ + j]));
                    }
                printf("manual layout=%s bo_mode=%s shift=%2d  max_err=%.3e  %s\n", layout == 0 ? "SW128" : "INTERLEAVE",
                       bo_mode == 0 ? "0" : "row&7", s, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
            }

    // ---- TMA conv probe -------------------------------------------------------------------------------------
    const int Bn = 2, L = 256, C = 24;
    std::vector<float> Xf((size_t)Bn * L * C), Wf((size_t)KS * N_ * 64, 0.f);
    std::vector<__nv_bfloat16> Xh(Xf.size()), Wh(Wf.size());
    for (size_t i = 0; i < Xf.size(); ++i) { Xf[i] = bf((rand() % 2001 - 1000) / 1000.f); Xh[i] = __float2bfloat16(Xf[i]); }
    for (int t = 0; t < KS; ++t) for (int co = 0; co < N_; ++co) for (int c = 0; c < 64; ++c) {
        float v = (c < C) ? bf((rand() % 2001 - 1000) / 4000.f) : bf(0.37f);   // weights beyond C are NON-zero: TMA must zero-fill X
        Wf[((size_t)t * N_ + co) * 64 + c] = v; Wh[((size_t)t * N_ + co) * 64 + c] = __float2bfloat16(v);
    }
    __nv_bfloat16 *dX, *dW;
    CK(cudaMalloc(&dX, Xh.size() * 2)); CK(cudaMalloc(&dW, Wh.size() * 2));
    CK(cudaMemcpy(dX, Xh.data(), Xh.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, Wh.data(), Wh.size() * 2, cudaMemcpyHostToDevice));
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 3; }
    CUtensorMap tmap;
    cuuint64_t gdim[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)Bn};
    cuuint64_t gstr[2] = {(cuuint64_t)C * 2, (cuuint64_t)L * C * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)TROWS, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dX, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 4; }
    const int smem2 = 18432 + KS * N_ * 128 + 2048;
    CK(cudaFuncSetAttribute(probe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    for (int bo_mode = 0; bo_mode < 2; ++bo_mode)
        for (int l0 : {0, 128}) for (int b : {0, 1}) {
            CK(cudaMemset(dD, 0, 128 * N_ * 4));
            probe_tma<<<1, 128, smem2>>>(tmap, dW, dD, l0, b, bo_mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("tma bo_mode=%d l0=%d b=%d: CUDA error %s\n", bo_mode, l0, b, cudaGetErrorString(e)); return 5; }
            CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
            double maxerr = 0;
            for (int i = 0; i < 128; ++i) for (int co = 0; co < N_; ++co) {
                double ref = 0;
                for (int t = 0; t < KS; ++t) {
                    int l = l0 + i + t - PADL;
                    if (l < 0 || l >= L) continue;
                    for (int c = 0; c < C; ++c) ref += (double)Xf[((size_t)b * L + l) * C + c] * Wf[((size_t)t * N_ + co) * 64 + c];
                }
                maxerr = fmax(maxerr, fabs(ref - Dh[i * N_ + co]));
            }
            printf("tma conv bo_mode=%s l0=%3d b=%d  max_err=%.3e  %s\n", bo_mode == 0 ? "0" : "row&7", l0, b, maxerr, maxerr < 2e-3 ? "OK" : "MISMATCH");
        }
    return 0;
}
