// host_cuda_shim.h -- TEST INFRASTRUCTURE (tests/ only; the product never builds or loads this).
//
// Just enough of the CUDA programming model to execute the kernels of scgaussian_b200/csrc/model.cu on the HOST,
// thread for thread: every block of a launch is run by `blockDim.x` real threads, `__syncthreads()` is a real barrier,
// `__shared__` arrays are per-process statics (blocks run one after the other).  What this checks before a GPU is
// available: indexing, bounds, the segment tables, the shared-memory staging and its barrier placement -- against
// the same oracle and golden vectors as the GPU tests.  What it cannot check: warp intrinsics (none are used by
// these kernels), memory-model subtleties, performance.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(x)
#define __grid_constant__
#define __restrict__
#define __align__(x) __attribute__((aligned(x)))
#define __shared__ static

struct float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
struct EmuDim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local EmuDim3 blockIdx, threadIdx;
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { return a + b; }      // compiled with -ffp-contract=off
static inline float __fmul_rn(float a, float b) { return a * b; }
using std::max;
using std::min;
typedef void* cudaStream_t;

static std::unique_ptr<std::barrier<>> g_block_barrier;
static inline void __syncthreads() { g_block_barrier->arrive_and_wait(); }

// kernel<<<grid, threads, 0, stream>>>(args) is rewritten (tests/emulation/build.py) into
// emu_launch(grid, threads, [=] { kernel(args); }).  `threads` workers walk the blocks in order; a barrier closes
// every block, so a kernel must use __syncthreads() uniformly within a block (as CUDA requires anyway).
template <class F>
static void emu_launch(unsigned grid, unsigned threads, F f) {
    g_block_barrier.reset(new std::barrier<>(threads));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; t++)
        pool.emplace_back([=] {
            for (unsigned b = 0; b < grid; b++) {
                blockIdx.x = b;
                threadIdx.x = t;
                f();
                g_block_barrier->arrive_and_wait();
            }
        });
    for (auto& th : pool) th.join();
}
