// Development probe: one 3-D TMA tile load of doubles into shared memory, copied back out for checking.
// usage: tma_probe boxw boxrows boxz x0 y0 z0 [px py nz]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe_k(const __grid_constant__ CUtensorMap map, int x0, int y0, int z0, int nbytes, double* out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + 65536);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(nbytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(sm)), "l"((unsigned long long)&map), "r"(x0), "r"(y0), "r"(z0), "r"(smem_u32(bar)) : "memory");
    }
    unsigned done = 0;
    for (unsigned spin = 0; !done && spin < (1u << 20); ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
    if (threadIdx.x == 0) out[0] = done ? 1.0 : -1.0;
    const double* s = reinterpret_cast<const double*>(sm);
    for (int i = threadIdx.x; i < nbytes / 8; i += blockDim.x) out[1 + i] = s[i];
}
int main(int argc, char** argv)
{
    int a[9] = {34, 3, 4, 3, 1, 0, 72, 10, 32};
    for (int i = 1; i < argc && i <= 9; ++i) a[i - 1] = atoi(argv[i]);
    const int bw = a[0], br = a[1], bz = a[2], x0 = a[3], y0 = a[4], z0 = a[5], px = a[6], py = a[7], nz = a[8];
    std::vector<double> h((size_t)px * py * nz);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i;
    double *d, *out;
    cudaMalloc(&d, h.size() * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    const int nbytes = bw * br * bz * 8;
    cudaMalloc(&out, nbytes + 8);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)px, (cuuint64_t)py, (cuuint64_t)nz}, strides[2] = {(cuuint64_t)px * 8, (cuuint64_t)px * py * 8};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)br, (cuuint32_t)bz}, es[3] = {1, 1, 1};
    CUresult r = ((Enc)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d  box %dx%dx%d at (%d,%d,%d) in %dx%dx%d: ", (int)r, bw, br, bz, x0, y0, z0, px, py, nz);
    cudaFuncSetAttribute(probe_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
    probe_k<<<1, 128, 65536 + 64>>>(map, x0, y0, z0, nbytes, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("KERNEL ERROR: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<double> o(nbytes / 8 + 1);
    cudaMemcpy(o.data(), out, nbytes + 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int z = 0; z < bz; ++z)
        for (int y = 0; y < br; ++y)
            for (int x = 0; x < bw; ++x) {
                const int gx = x0 + x, gy = y0 + y, gz = z0 + z;
                const double want = (gx < 0 || gx >= px || gy < 0 || gy >= py || gz < 0 || gz >= nz) ? 0.0 : (double)(gx + (size_t)px * (gy + (size_t)py * gz));
                if (o[1 + x + bw * (y + br * z)] != want) ++bad;
            }
    printf("done=%g mismatches=%d\n", o[0], bad);
    return 0;
}
