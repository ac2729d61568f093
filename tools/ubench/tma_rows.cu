// tma_rows.cu -- how fast does one SM's TMA unit move boxes made of short rows?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_rows tools/ubench/tma_rows.cu && build/tma_rows
//
// One block per SM, W warps per block; lane 0 of every warp keeps DEPTH loads (and optionally as many
// stores) of [ROWS x ROW_BYTES] boxes in flight out of an L2-resident [R, pitch] fp32 matrix (row pitch
// 576 B like the observation rows) and counts completed boxes for a fixed number of iterations.
// Prints boxes / rows / bytes per cycle per SM.  Answers: is a 32 x 64 B box limited by rows (requests)
// or by bytes?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void mbar_init(unsigned m) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(m) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned m, unsigned b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_try(unsigned m, unsigned p) {
    unsigned d;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(d) : "r"(m), "r"(p) : "memory");
    return d;
}

template <int DEPTH>
__global__ void __launch_bounds__(1024, 1) k(const __grid_constant__ CUtensorMap tm, int box_bytes, int iters, int n_row_tiles,
                                               int do_store, long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long bars[32 * DEPTH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const unsigned sb = (unsigned)__cvta_generic_to_shared(smem) + warp * DEPTH * box_bytes;
    const unsigned mb = (unsigned)__cvta_generic_to_shared(bars) + warp * DEPTH * 8;
    if (lane == 0) for (int d = 0; d < DEPTH; ++d) mbar_init(mb + 8 * d);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    long long t0 = clock64();
    if (lane == 0) {
        int tile = (blockIdx.x * W + warp) % n_row_tiles;
        for (int d = 0; d < DEPTH; ++d) {
            mbar_expect(mb + 8 * d, box_bytes);
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(sb + d * box_bytes), "l"(&tm), "r"(0), "r"(tile * 32), "r"(mb + 8 * d) : "memory");
            tile = (tile + gridDim.x * W) % n_row_tiles;
        }
        for (int i = 0; i < iters; ++i) {
            const int d = i % DEPTH;
            while (!mbar_try(mb + 8 * d, (i / DEPTH) & 1)) {}
            if (do_store) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(&tm), "r"(0), "r"(tile * 32), "r"(sb + d * box_bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            if (i + DEPTH < iters) {
                mbar_expect(mb + 8 * d, box_bytes);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(sb + d * box_bytes), "l"(&tm), "r"(0), "r"(tile * 32), "r"(mb + 8 * d) : "memory");
                tile = (tile + gridDim.x * W) % n_row_tiles;
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
    EncodeTiledFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &q));
    int sms;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int R = 32768, pitch = 144;                       // 32768 rows x 576 B = 18.9 MB: L2 resident
    float *buf;
    CK(cudaMalloc(&buf, (size_t)R * pitch * 4));
    CK(cudaMemset(buf, 0, (size_t)R * pitch * 4));
    long long *cyc;
    CK(cudaMalloc(&cyc, sms * 8));
    CK(cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    printf("| box | warps | depth | store | boxes/kcycle/SM | rows/cycle/SM | B/cycle/SM |\n|---|---|---|---|---|---|---|\n");
    const int row_floats[] = {8, 16, 32, 64};
    for (int rf : row_floats)
        for (int W : {1, 4, 14, 28})
            for (int depth : {2, 4})
                for (int st : {0, 1}) {
                    const int box_bytes = 32 * rf * 4;
                    if ((size_t)W * depth * box_bytes > 200 * 1024) continue;
                    CUtensorMap tm;
                    const cuuint64_t gdim[2] = {128, (cuuint64_t)R};
                    const cuuint64_t gstr[1] = {pitch * 4};
                    const cuuint32_t box[2] = {(cuuint32_t)rf, 32}, es[2] = {1, 1};
                    const CUtensorMapSwizzle sw = rf * 4 >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                  : rf * 4 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
                    if (rf * 4 > 128) continue;
                    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf + 16, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                        printf("encode failed\n");
                        return 1;
                    }
                    const int iters = 400;
                    for (int rep = 0; rep < 2; ++rep) {
                        if (depth == 2) k<2><<<sms, W * 32, (size_t)W * depth * box_bytes>>>(tm, box_bytes, iters, R / 32, st, cyc);
                        else k<4><<<sms, W * 32, (size_t)W * depth * box_bytes>>>(tm, box_bytes, iters, R / 32, st, cyc);
                        CK(cudaDeviceSynchronize());
                    }
                    std::vector<long long> h(sms);
                    CK(cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost));
                    double mean = 0;
                    for (auto c : h) mean += c;
                    mean /= sms;
                    const double boxes = (double)W * iters * (st ? 2 : 1);
                    printf("| 32 x %d B | %d | %d | %d | %.1f | %.3f | %.1f |\n", rf * 4, W, depth, st, 1e3 * boxes / mean, 32 * boxes / mean,
                           boxes * box_bytes / mean);
                }
    return 0;
}
