// Throw-away probe: which way of passing a TMA descriptor / which rank works on this box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
struct alignas(64) Maps { CUtensorMap m[8]; };
__device__ __forceinline__ uint32_t s32(const void* p){ return (uint32_t)__cvta_generic_to_shared(p); }
template<int RANK>
__device__ void run(const CUtensorMap* map, int x, int y, int z, unsigned* out, int bytes, int flags) {
    __shared__ alignas(128) unsigned char win[8192];
    __shared__ unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bar)) : "memory");
    }
    if (!(flags & 2)) {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(&bar)), "r"((flags & 1) ? 0 : bytes) : "memory");
        if (flags & 1) {}
        else if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                :: "r"(s32(win)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(s32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                :: "r"(s32(win)), "l"(map), "r"(x), "r"(y), "r"(s32(&bar)) : "memory");
    }
    unsigned spins = 0, ok = 0;
    while (!ok && spins < (1u<<20)) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(s32(&bar)), "r"(0) : "memory");
        ++spins;
    }
    __syncthreads();
    if (threadIdx.x == 0) { out[0] = ok; out[1] = spins; unsigned s = 0; for (int i = 0; i < bytes; ++i) s += win[i]; out[2] = s; out[3] = win[0]; }
}
__global__ void k_param3(const __grid_constant__ Maps maps, int lvl, int x, int y, int z, unsigned* out, int bytes, int flags) { run<3>(&maps.m[lvl], x, y, z, out, bytes, flags); }
__global__ void k_param2(const __grid_constant__ Maps maps, int lvl, int x, int y, int z, unsigned* out, int bytes, int flags) { run<2>(&maps.m[lvl], x, y, z, out, bytes, flags); }
__global__ void k_glob3(const CUtensorMap* maps, int lvl, int x, int y, int z, unsigned* out, int bytes, int flags) { run<3>(maps + lvl, x, y, z, out, bytes, flags); }
__global__ void k_glob2(const CUtensorMap* maps, int lvl, int x, int y, int z, unsigned* out, int bytes, int flags) { run<2>(maps + lvl, x, y, z, out, bytes, flags); }
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    int x = argc > 2 ? atoi(argv[2]) : 100; unsigned bw = argc > 3 ? atoi(argv[3]) : 32, bh = argc > 4 ? atoi(argv[4]) : 24; int flags = argc > 5 ? atoi(argv[5]) : 0;
    int cols = 620, rows = 188, pitch = 624, B = 2;
    size_t slot = (size_t)rows * pitch;
    unsigned char* d; cudaMalloc(&d, B * slot);
    std::vector<unsigned char> h(B * slot);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 7 + (i >> 8));
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    Enc enc = (Enc)fp;
    Maps* maps = new Maps(); memset(maps, 0, sizeof(Maps));
    int rank = (variant & 1) ? 2 : 3;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)(rank == 3 ? rows : rows * B), (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)slot};
    cuuint32_t box[3] = {bw, bh, 1}, es[3] = {1, 1, 1};
    CUresult rc = enc(&maps->m[1], CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, (flags & 4) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d rank %d encode rc=%d\n", variant, rank, (int)rc);
    unsigned* out; cudaMalloc(&out, 16); cudaMemset(out, 0, 16);
    CUtensorMap* dmaps; cudaMalloc(&dmaps, sizeof(Maps)); cudaMemcpy(dmaps, maps, sizeof(Maps), cudaMemcpyHostToDevice);
    int y = 50, z = 1; int bytes = bw * bh;
    if (variant == 0) k_param3<<<1, 32>>>(*maps, 1, x, y, z, out, bytes, flags);
    if (variant == 1) k_param2<<<1, 32>>>(*maps, 1, x, y + rows, z, out, bytes, flags);
    if (variant == 2) k_glob3<<<1, 32>>>(dmaps, 1, x, y, z, out, bytes, flags);
    if (variant == 3) k_glob2<<<1, 32>>>(dmaps, 1, x, y + rows, z, out, bytes, flags);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned r[4]; cudaMemcpy(r, out, 16, cudaMemcpyDeviceToHost);
    unsigned expect = 0; for (int j = 0; j < (int)bh; ++j) for (int i = 0; i < (int)bw; ++i) expect += h[slot + (size_t)(y + j) * pitch + x + i];
    printf("variant %d x=%d box=%ux%u flags=%d: sync=%s ok=%u spins=%u sum=%u expect=%u first=%u expect_first=%u\n", variant, x, bw, bh, flags, cudaGetErrorString(e), r[0], r[1], r[2], expect, r[3], (unsigned)h[slot + (size_t)y * pitch + x]);
    return 0;
}
