// Probe: which TMA tile configurations work on this box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <class R, int BW, int BH>
__global__ void probe(const __grid_constant__ CUtensorMap map, R* out, int x, int y) {
    extern __shared__ __align__(128) unsigned char smem[];
    R* tile = reinterpret_cast<R*>(smem);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + ((BW * BH * sizeof(R) + 127) / 128 * 128));
    unsigned b = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((unsigned)(BW * BH * sizeof(R))) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(tile)), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(x), "r"(y), "r"(b) : "memory");
    }
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(b), "r"(0) : "memory");
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <class R, int BW, int BH> void run(EncodeFn enc, const char* name, CUtensorMapDataType dt, int cols, int rows, int pitch) {
    std::vector<R> h((size_t)rows * pitch);
    for (int y = 0; y < rows; ++y) for (int x = 0; x < pitch; ++x) h[(size_t)y * pitch + x] = (R)(y * 1000 + x);
    R *d, *o;
    cudaMalloc(&d, h.size() * sizeof(R)); cudaMalloc(&o, BW * BH * sizeof(R));
    cudaMemcpy(d, h.data(), h.size() * sizeof(R), cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t gstr[1] = {(cuuint64_t)pitch * sizeof(R)};
    cuuint32_t box[2] = {BW, BH}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, dt, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int smem = (BW * BH * sizeof(R) + 127) / 128 * 128 + 64;
    cudaFuncSetAttribute(probe<R, BW, BH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int starts[][2] = {{0, 0}, {4, 3}, {0, -1}, {-4, -1}, {252, 250}, {2, 0}, {1, 0}, {-1, 0}};
    for (auto& st : starts) {
        const int xs = st[0], ys = st[1];
        probe<R, BW, BH><<<1, 128, smem>>>(tm, o, xs, ys);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<R> got(BW * BH);
        cudaMemcpy(got.data(), o, got.size() * sizeof(R), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int j = 0; j < BH; ++j) for (int i = 0; i < BW; ++i) {
            int gx = xs + i, gy = ys + j;
            R want = (gx < 0 || gy < 0 || gx >= cols || gy >= rows) ? (R)0 : (R)(gy * 1000 + gx);
            if (got[j * BW + i] != want) ++bad;
        }
        printf("%-28s encode=%d start=(%3d,%3d) -> %s, mismatches=%d\n", name, (int)r, xs, ys, cudaGetErrorString(e), bad);
        if (e != cudaSuccess) { return; }
    }
}

#include <cstdlib>
int main(int argc, char** argv) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 0) run<float, 68, 10>(enc, "f32 68x10", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 256, 256, 256);
    if (which == 1) run<double, 66, 10>(enc, "f64 66x10", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 256, 256, 256);
    if (which == 2) run<double, 66, 10>(enc, "f64 66x10 cols96", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 96, 96, 96);
    if (which == 3) run<double, 68, 10>(enc, "f64 68x10", CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 256, 256, 256);
    return 0;
}
