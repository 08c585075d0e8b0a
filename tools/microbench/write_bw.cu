// Write-only bandwidth probes for the render roofline discussion (DESIGN.md "Render roofline").
//   a) cudaMemsetAsync           b) st.global.v4 grid-stride kernel
//   c) TMA bulk store (cp.async.bulk.global.shared::cta) of a resident smem tile, persistent CTAs
//   d) copy of an L2-resident 1.26 MB source replicated to a large destination with LDG.128/STG.128
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_bw write_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void st_kernel(uint4 *dst, size_t n16) {
    uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

__global__ void copy_rep_kernel(uint4 *dst, const uint4 *__restrict__ src, size_t src16, size_t reps) {
    // each CTA copies whole replicas: dst[rep][j] = src[j]
    for (size_t rep = blockIdx.x; rep < reps; rep += gridDim.x) {
        uint4 *d = dst + rep * src16;
        for (size_t j = threadIdx.x; j < src16; j += blockDim.x) d[j] = __ldg(src + j);
    }
}

template <int CHUNK>
__global__ void tma_store_kernel(uint8_t *dst, size_t nchunks) {
    extern __shared__ __align__(128) uint8_t buf[];
    for (int i = threadIdx.x; i < CHUNK / 4; i += blockDim.x) ((uint32_t *)buf)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a = (uint32_t)__cvta_generic_to_shared(buf);
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * CHUNK), "r"(a), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main() {
    const size_t bytes = (size_t)8 << 30;
    uint8_t *dst, *src;
    CK(cudaMalloc(&dst, bytes));
    const size_t frame = 1257984;
    CK(cudaMalloc(&src, frame));
    CK(cudaMemset(src, 7, frame));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a); CK(cudaMemsetAsync(dst, 1, bytes)); cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b); printf("memset            %7.1f GB/s\n", bytes / ms / 1e6);
        for (int mult : {2, 4, 8, 16}) {
            cudaEventRecord(a); st_kernel<<<sms * mult, 512>>>((uint4 *)dst, bytes / 16); cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b); printf("st.v4 grid %2dxSM  %7.1f GB/s\n", mult, bytes / ms / 1e6);
        }
        size_t reps = bytes / frame;
        for (int mult : {2, 4, 8}) {
            cudaEventRecord(a); copy_rep_kernel<<<sms * mult, 512>>>((uint4 *)dst, (const uint4 *)src, frame / 16, reps); cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b); printf("ldg/stg replicate %dxSM %7.1f GB/s\n", mult, reps * frame / ms / 1e6);
        }
        {
            constexpr int CH = 32768;
            cudaFuncSetAttribute(tma_store_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH);
            for (int mult : {1, 2, 4}) {
                cudaEventRecord(a); tma_store_kernel<CH><<<sms * mult, 128, CH>>>(dst, bytes / CH); cudaEventRecord(b); cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms, a, b); printf("tma store 32K %dxSM %7.1f GB/s\n", mult, bytes / ms / 1e6);
            }
            constexpr int CH2 = 98304;
            cudaFuncSetAttribute(tma_store_kernel<CH2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH2);
            for (int mult : {1, 2}) {
                cudaEventRecord(a); tma_store_kernel<CH2><<<sms * mult, 128, CH2>>>(dst, bytes / CH2); cudaEventRecord(b); cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms, a, b); printf("tma store 96K %dxSM %7.1f GB/s\n", mult, (bytes / CH2) * CH2 / ms / 1e6);
            }
        }
        CK(cudaGetLastError());
    }
    return 0;
}
