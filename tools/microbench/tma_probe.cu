// Probe of cp.async.bulk.tensor.2d variants on B200 (one variant per process: a fault kills the context).
// usage: tma_probe <desc: 0 param | 1 global> <c0> <boxw> <r0>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int GLOBAL>
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int c0, int r0, int boxw, int R, float* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* buf = (float*)sm;
    unsigned long long* bar = (unsigned long long*)(sm + 32768);
    const CUtensorMap* m = GLOBAL ? gmap : &pmap;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (GLOBAL) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(m) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(boxw * R * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(buf)), "l"(m), "r"(c0), "r"(r0), "r"(smem_u32(bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(bar)) : "memory");
    for (int i = threadIdx.x; i < boxw * R; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char** argv) {
    int glob = atoi(argv[1]), c0 = atoi(argv[2]), boxw = atoi(argv[3]), r0 = atoi(argv[4]);
    const int nx = 500, ny = 500, ld = 512, R = 4;
    std::vector<float> h((size_t)ld * ny + 64);
    for (int j = 0; j < ny; ++j) for (int i = 0; i < ld; ++i) h[(size_t)j * ld + i] = (i < nx) ? 1000.f * j + i : -7.f;
    float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 65536);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)nx, (cuuint64_t)ny}, strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)boxw, (cuuint32_t)R}, es[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(map)); cudaMemcpy(gm, &map, sizeof(map), cudaMemcpyHostToDevice);
    if (glob) probe<1><<<1, 128, 32768 + 64>>>(map, gm, c0, r0, boxw, R, out);
    else probe<0><<<1, 128, 32768 + 64>>>(map, gm, c0, r0, boxw, R, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("glob=%d c0=%d boxw=%d r0=%d -> %s\n", glob, c0, boxw, r0, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> o(boxw * R); cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < R; ++rr) for (int i = 0; i < boxw; ++i) {
            int c = c0 + i, j = r0 + rr; float want = (c >= 0 && c < nx && j >= 0 && j < ny) ? 1000.f * j + c : 0.f;
            if (o[rr * boxw + i] != want) ++bad;
        }
        printf("  mismatches: %d (first row: %g %g %g %g)\n", bad, o[0], o[1], o[2], o[3]);
    }
    return 0;
}
