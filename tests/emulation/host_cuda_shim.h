// host_cuda_shim.h -- TEST INFRASTRUCTURE (tests/ only; the product never builds or loads this).
//
// Just enough of the CUDA programming model to execute the kernels of scgaussian_b200/csrc/{model,preprocess,binning,render,loss,knn}.cu on the HOST,
// thread for thread: every block of a launch is run by `blockDim.x` real threads, `__syncthreads()` is a real barrier,
// `__shared__` arrays are per-process statics (blocks run one after the other).  What this checks before a GPU is
// available: indexing, bounds, the segment tables, the shared-memory staging and its barrier placement -- against
// the same oracle and golden vectors as the GPU tests.  What it cannot check: warp intrinsics (none are used by
// these kernels beyond full-mask votes / shuffles / match / reduce), memory-model subtleties, performance.
#pragma once
#define SCGR_HOST_EMULATION 1      // common.cuh: chain() launches through emu_launch, griddepcontrol is a no-op
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
#define __align__(x) __attribute__((aligned(x)))
#define __shared__ static
#define __constant__ static

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline float3 make_float3(float a, float b, float c) { return {a, b, c}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return {a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static thread_local dim3 blockIdx(0, 0, 0), threadIdx(0, 0, 0), blockDim, gridDim;
typedef void* cudaStream_t;
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { return a + b; }      // compiled with -ffp-contract=off
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline void __trap() { std::abort(); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline int cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicMin(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline float atomicAdd(float* p, float v) {
    unsigned old = __atomic_load_n(reinterpret_cast<unsigned*>(p), __ATOMIC_SEQ_CST), want;
    float f;
    do {
        std::memcpy(&f, &old, 4);
        f += v;
        std::memcpy(&want, &f, 4);
    } while (!__atomic_compare_exchange_n(reinterpret_cast<unsigned*>(p), &old, want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    std::memcpy(&f, &old, 4);
    return f;
}
static inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }
// ---- the slice of the CUDA runtime the host side of the library uses.  Everything is synchronous here: a kernel has
// finished when its launch returns, so streams and events have nothing to order and a "mapped" pinned pointer is the
// host pointer itself.
typedef int cudaError_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorNotReady = 600, cudaErrorNoDevice = 100, cudaDevAttrMultiProcessorCount = 16,
       cudaMemcpyDeviceToHost = 2, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
// device queries: device 0 exists as far as stream bookkeeping goes; attribute queries fail, so callers fall back to
// their defaults (e.g. 148 SMs)
static inline cudaError_t cudaGetDevice(int* d) { if (d) *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int*, int, int) { return cudaErrorNoDevice; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "host emulation: no CUDA error"; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = reinterpret_cast<cudaStream_t>(1); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(1); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = reinterpret_cast<cudaEvent_t>(1); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, int, cudaStream_t) { std::memcpy(dst, src, n); return cudaSuccess; }
static inline cudaError_t cudaHostGetDevicePointer(void** dev, void* host, unsigned) { *dev = host; return cudaSuccess; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
// dynamic shared memory: `extern __shared__ T name[];` is rewritten to a pointer into this buffer
alignas(16) static unsigned char g_dyn_smem[256 * 1024];
#define cudaFuncAttributeMaxDynamicSharedMemorySize 0
template <class F> static inline int cudaFuncSetAttribute(F, int, int) { return 0; }
using std::max;
using std::min;

static std::unique_ptr<std::barrier<>> g_block_barrier;
static inline void __syncthreads() { g_block_barrier->arrive_and_wait(); }

// warp votes: the 32 threads of a warp meet at a per-warp barrier, publish their predicate, and read all 32
// (full-mask votes only: every lane of the warp must take part, which is what the kernels do)
static std::vector<std::unique_ptr<std::barrier<>>> g_warp_barrier;
static unsigned g_votes[64][32];
static unsigned g_warp_lanes[64];
static inline unsigned __ballot_sync(unsigned, int pred) {
    const unsigned w = threadIdx.x / 32, lane = threadIdx.x % 32;
    g_votes[w][lane] = pred ? 1u : 0u;
    g_warp_barrier[w]->arrive_and_wait();
    unsigned r = 0;
    for (unsigned l = 0; l < g_warp_lanes[w]; l++) r |= g_votes[w][l] << l;
    g_warp_barrier[w]->arrive_and_wait();
    return r;
}
// warp shuffles: every lane publishes its value, the warp meets, every lane reads the lane it names (its own value
// when that lane does not exist, as the hardware does)
static unsigned long long g_shfl[64][32];
template <class T>
static inline T emu_shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 64 bits");
    const unsigned w = threadIdx.x / 32, lane = threadIdx.x % 32;
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    g_shfl[w][lane] = bits;
    g_warp_barrier[w]->arrive_and_wait();
    T out = v;
    if (src >= 0 && (unsigned)src < g_warp_lanes[w]) std::memcpy(&out, &g_shfl[w][src], sizeof(T));
    g_warp_barrier[w]->arrive_and_wait();
    return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu_shfl(v, (int)((threadIdx.x % 32) ^ (unsigned)lane_mask)); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl(v, src & 31); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    const int lane = (int)(threadIdx.x % 32);
    return emu_shfl(v, lane >= (int)delta ? lane - (int)delta : -1);
}
static inline void __syncwarp(unsigned = 0xffffffffu) { g_warp_barrier[threadIdx.x / 32]->arrive_and_wait(); }
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    unsigned m = v;
    for (int o = 16; o > 0; o >>= 1) m = std::max(m, __shfl_xor_sync(0xffffffffu, m, o));
    return m;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    unsigned m = v;
    for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    return m;
}
static inline int __any_sync(unsigned, int pred) { return __ballot_sync(0xffffffffu, pred) != 0u; }
static inline unsigned __match_any_sync(unsigned, unsigned v) {
    unsigned peers = 0;
    for (int l = 0; l < 32; l++) {
        const unsigned other = __shfl_sync(0xffffffffu, v, l);
        if (other == v && (unsigned)l < g_warp_lanes[threadIdx.x / 32]) peers |= 1u << l;
    }
    return peers;
}

// kernel<<<grid, threads, 0, stream>>>(args) is rewritten (tests/emulation/build.py) into
// emu_launch(grid, threads, [=] { kernel(args); }).  `threads` workers walk the blocks in order; a barrier closes
// every block, so a kernel must use __syncthreads() uniformly within a block (as CUDA requires anyway).
template <class F>
static void emu_launch(dim3 grid, unsigned threads, F f) {
    g_block_barrier.reset(new std::barrier<>(threads));
    g_warp_barrier.clear();
    for (unsigned w = 0; w * 32 < threads; w++) {
        g_warp_lanes[w] = std::min(32u, threads - w * 32);
        g_warp_barrier.emplace_back(new std::barrier<>(g_warp_lanes[w]));
    }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; t++)
        pool.emplace_back([=] {
            threadIdx = dim3(t, 0, 0);
            blockDim = dim3(threads, 1, 1);
            gridDim = grid;
            for (unsigned bz = 0; bz < grid.z; bz++)
                for (unsigned by = 0; by < grid.y; by++)
                    for (unsigned bx = 0; bx < grid.x; bx++) {
                        blockIdx = dim3(bx, by, bz);
                        f();
                        g_block_barrier->arrive_and_wait();
                    }
        });
    for (auto& th : pool) th.join();
}

// ---- mbarrier + bulk-async copy (common.cuh's TMA plumbing), emulated with a real phase protocol: the 64-bit barrier word
// holds {phase:1 | init count:15 | pending arrivals:16 | transaction bytes:32 (signed, may go negative transiently)}; a
// phase completes when no arrival is pending and the byte count is zero, exactly as the hardware object does.  The copy
// itself is a synchronous memcpy followed by its complete_tx. ----
static inline void emu_mbar_update(unsigned long long* bar, int d_pending, long long d_tx) {
    unsigned long long old = __atomic_load_n(bar, __ATOMIC_SEQ_CST), neu;
    do {
        unsigned phase = (unsigned)(old >> 63), count = (unsigned)((old >> 48) & 0x7fffu);
        int pending = (int)((old >> 32) & 0xffffu) + d_pending;
        long long tx = (long long)(int)(unsigned)(old & 0xffffffffu) + d_tx;
        if (pending == 0 && tx == 0) { phase ^= 1u; pending = (int)count; }
        neu = ((unsigned long long)phase << 63) | ((unsigned long long)count << 48) | ((unsigned long long)(pending & 0xffff) << 32) |
              (unsigned long long)(unsigned)(int)tx;
    } while (!__atomic_compare_exchange_n(bar, &old, neu, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
}
static inline void emu_mbar_init(unsigned long long* bar, unsigned count) {
    __atomic_store_n(bar, ((unsigned long long)count << 48) | ((unsigned long long)count << 32), __ATOMIC_SEQ_CST);
}
static inline void emu_mbar_wait(unsigned long long* bar, unsigned parity) {
    while ((unsigned)(__atomic_load_n(bar, __ATOMIC_SEQ_CST) >> 63) == (parity & 1u)) std::this_thread::yield();
}
