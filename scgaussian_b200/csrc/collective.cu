// collective.cu -- the path's single collective as a hand-written NVSwitch (NVLS) kernel (sm_100a).
//
// the gradients of the N views rendered by the N ranks are summed with ONE
// all-reduce over one flat fp32 buffer per step.  When that buffer lives in symmetric memory with a
// multicast mapping (torch.distributed._symmetric_memory; scgaussian_b200/parallel.py), the sum is
// done here in two shots through the switch instead of by NCCL:
//   shot 1  rank r pulls ITS 1/N of the buffer with multimem.ld_reduce: the switch reads the N
//           replicas and returns their sum (in-network reduction, fp32 accumulate);
//   shot 2  rank r pushes the reduced values with multimem.st: the switch writes them into all N
//           replicas.
// Every GPU sends and receives the buffer once; every replica ends up with bit-identical sums
// (one reduction per element, broadcast).  The caller brackets the kernel with cross-rank barriers
// (all replicas written before shot 1, all shards broadcast before anyone reads).
// HBM / NVLink-bound streaming of fp32: no tensor cores.
#include <cstdlib>

#include "common.cuh"

namespace scgr {

namespace {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, const float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

template <int NVLS_THREADS, int NVLS_UNROLL>
__global__ void __launch_bounds__(NVLS_THREADS)
nvls_allreduce_kernel(float4* __restrict__ mc, const size_t first, const size_t count) {
    const size_t stride = (size_t)gridDim.x * NVLS_THREADS;
    size_t i = (size_t)blockIdx.x * NVLS_THREADS + threadIdx.x;
    // NVLS_UNROLL independent reductions in flight per thread
    for (; i + (NVLS_UNROLL - 1) * stride < count; i += NVLS_UNROLL * stride) {
        float4 v[NVLS_UNROLL];
#pragma unroll
        for (int u = 0; u < NVLS_UNROLL; u++) v[u] = multimem_ld_reduce_add(mc + first + i + u * stride);
#pragma unroll
        for (int u = 0; u < NVLS_UNROLL; u++) multimem_st(mc + first + i + u * stride, v[u]);
    }
    for (; i < count; i += stride) multimem_st(mc + first + i, multimem_ld_reduce_add(mc + first + i));
}

}  // namespace

void launch_nvls_allreduce(void* multicast_ptr, size_t n_floats, int rank, int world, const Launch& L) {
    const size_t n4 = n_floats / 4;                       // caller guarantees n_floats % (4 * world) == 0
    const size_t per = n4 / (size_t)world;
    if (per == 0) return;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const int ctas_per_sm = getenv("SCGR_NVLS_CTAS") ? atoi(getenv("SCGR_NVLS_CTAS")) : 2;
    static const int variant = getenv("SCGR_NVLS_VARIANT") ? atoi(getenv("SCGR_NVLS_VARIANT")) : 0;
    float4* const mc = reinterpret_cast<float4*>(multicast_ptr);
    const size_t first = per * (size_t)rank;
    auto grid_for = [&](int threads, int unroll) {
        const size_t want = (per + (size_t)threads * unroll - 1) / ((size_t)threads * unroll);
        const size_t cap = (size_t)ctas_per_sm * sms;
        return (unsigned)(want < cap ? want : cap);
    };
    begin_kernel("nvls_allreduce", L);
    switch (variant) {
        case 1: nvls_allreduce_kernel<512, 8><<<grid_for(512, 8), 512, 0, L.stream>>>(mc, first, per); break;
        case 2: nvls_allreduce_kernel<1024, 4><<<grid_for(1024, 4), 1024, 0, L.stream>>>(mc, first, per); break;
        case 3: nvls_allreduce_kernel<256, 2><<<grid_for(256, 2), 256, 0, L.stream>>>(mc, first, per); break;
        default: nvls_allreduce_kernel<512, 4><<<grid_for(512, 4), 512, 0, L.stream>>>(mc, first, per); break;
    }
    check_launch("nvls_allreduce", L);
}

}  // namespace scgr
