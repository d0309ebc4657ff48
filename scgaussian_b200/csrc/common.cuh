// common.cuh -- shared definitions of the sm_100a rasterizer kernels (internal; the public
// surface is include/scgr.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/scgr.h"

namespace scgr {

constexpr int TILE = SCGR_TILE;             // 16x16 pixel tiles
constexpr int TILE_PIX = TILE * TILE;
constexpr float NEAR_Z = 0.2f;              // A.1
constexpr float ALPHA_MIN = 1.0f / 255.0f;  // A.8
constexpr float ALPHA_MAX = 0.99f;          // A.8
constexpr float T_EPS = 1e-4f;              // A.8
constexpr float DILATION = 0.3f;            // A.4
constexpr uint32_t CULLED_KEY = 0xFFFFFFFFu;

// ---- per-Gaussian record written by preprocess, gathered by render (48 B, 3 x 16-B loads) ----
//   q0 = {mean_x, mean_y, cA, cB}      cA = -0.5 log2e A, cB = -log2e B, cC = -0.5 log2e C  (A,B,C = conic)
//   q1 = {cC, opacity, depth, pmin2}   so that G = exp2(cA dx^2 + cB dx dy + cC dy^2); pmin2 = -log2(255 opacity)
//   q2 = {r, g, b, bits}               bits (uint32) = radius | flags << 28; flags bit0..2 = SH colour clamped
struct __align__(16) Record {
    float4 q0, q1, q2;
};
static_assert(sizeof(Record) == 48, "Record must be 48 bytes");

// ---- per-Gaussian screen-space gradient accumulator written by render-backward (48 B) ----
// raw sums over (pixel, Gaussian) pairs; preprocess-backward applies the constant factors:
// with u = opacity * G * dL/dalpha per pair:
//   a0 = {Sx, Sy, SA, SB} = sum {u dx, u dy, u dx^2, u dx dy}    dL/dmean (pixels) = ln2 * (2cA Sx + cB Sy, 2cC Sy + cB Sx)
//   a1 = {SC, Su, dL/ddepth, 0} = sum {u dy^2, u, ...}           dL/dconic = (-0.5 SA, -SB, -0.5 SC), dL/dopacity = Su / opacity
//   a2 = {dL/dr, dL/dg, dL/db, 0}
struct __align__(16) ScreenGrad {
    float4 a0, a1, a2;
};
static_assert(sizeof(ScreenGrad) == 48, "ScreenGrad must be 48 bytes");

// ---- exact (Gaussian, pixel-rectangle) culling, shared by preprocess (count), emission (write) and
// render (per-slot skip).  Maximum over the pixel rectangle [x0, x1] x [y0, y1] of
//   p2(d) = cA dx^2 + cB dx dy + cC dy^2,  d = mean - pixel   (negative definite, max 0 at d = 0).
// The maximiser of a concave function whose global maximum lies outside the box is on one of the
// two faces that look at the origin; each face is a 1-D concave parabola (kx = -cB/(2cA),
// ky = -cB/(2cC) give its vertex).  Written with explicit round-to-nearest intrinsics so that the
// counting pass and the emission pass -- two different kernels -- take bit-identical decisions.
#ifdef __CUDACC__
__device__ __forceinline__ float max_power_over_rect(const float cA, const float cB, const float cC,
                                                     const float kx, const float ky, const float mx,
                                                     const float my, const float x0, const float x1,
                                                     const float y0, const float y1) {
    const float dxl = __fsub_rn(mx, x1), dxh = __fsub_rn(mx, x0);
    const float dyl = __fsub_rn(my, y1), dyh = __fsub_rn(my, y0);
    const float cx = fminf(fmaxf(0.f, dxl), dxh);
    const float cy = fminf(fmaxf(0.f, dyl), dyh);
    const float dy1 = fminf(fmaxf(__fmul_rn(ky, cx), dyl), dyh);   // face dx = cx
    const float f1 = __fmaf_rn(cx, __fmaf_rn(cA, cx, __fmul_rn(cB, dy1)), __fmul_rn(__fmul_rn(cC, dy1), dy1));
    const float dx2 = fminf(fmaxf(__fmul_rn(kx, cy), dxl), dxh);   // face dy = cy
    const float f2 = __fmaf_rn(dx2, __fmaf_rn(cA, dx2, __fmul_rn(cB, cy)), __fmul_rn(__fmul_rn(cC, cy), cy));
    return fmaxf(f1, f2);
}
// log2 units; keeps every rectangle bound conservative under fp32 rounding of the per-pixel test
constexpr float CULL_MARGIN = 0.02f;

struct CullParams {   // per-Gaussian constants of the test, derived from the stored record only
    float cA, cB, cC, kx, ky, mx, my, thr;
};
__device__ __forceinline__ CullParams make_cull(const float4 q0, const float4 q1) {
    CullParams c;
    c.mx = q0.x; c.my = q0.y; c.cA = q0.z; c.cB = q0.w; c.cC = q1.x;
    c.kx = __fdiv_rn(-c.cB, __fmul_rn(2.f, c.cA));
    c.ky = __fdiv_rn(-c.cB, __fmul_rn(2.f, c.cC));
    c.thr = __fsub_rn(q1.w, CULL_MARGIN);
    return c;
}
// ---- the same test for a whole ROW of tiles at once: which tiles of tile row `ty` does the ellipse
//   { d : p2(d) >= thr }   (the region in which alpha >= 1/255 is possible, with the margin)
// intersect?  Over the row band dy in [dyl, dyh] the ellipse's extreme dx are reached at the dy of its
// right-/leftmost point clamped into the band (dx_max(dy) = kx dy + sqrt(E - D dy^2) / (-2 cA) is
// concave), so one row costs two square roots instead of one rectangle test per tile.
// D = 4 cA cC - cB^2 suffers cancellation for needle-shaped Gaussians: `robust` is false then and the
// callers fall back to the per-tile test.  Explicit rounding + one hardware sqrt: the counting pass
// (preprocess) and the emission pass take bit-identical decisions.
__device__ __forceinline__ float sqrt_hw(const float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float SPAN_SLACK = 0.01f;      // pixels; on top of CULL_MARGIN
struct SpanParams {
    float mx, my, kx, hA, D, E, dyT, dyR;
    bool any;       // false: no pixel anywhere can reach the threshold
    bool robust;    // false: use tile_may_contribute() per tile
};
__device__ __forceinline__ SpanParams make_span(const CullParams& c) {
    SpanParams s;
    s.mx = c.mx; s.my = c.my; s.kx = c.kx;
    const float ac4 = __fmul_rn(__fmul_rn(4.f, c.cA), c.cC);
    s.D = __fsub_rn(ac4, __fmul_rn(c.cB, c.cB));
    s.E = __fmul_rn(__fmul_rn(4.f, c.cA), c.thr);
    s.hA = __fdiv_rn(-0.5f, c.cA);
    s.any = c.thr < 0.f;
    s.robust = s.D > __fmul_rn(1e-3f, ac4) && s.D > 0.f;
    const float invD = __fdiv_rn(1.f, s.D);
    s.dyT = __fadd_rn(sqrt_hw(fmaxf(0.f, __fmul_rn(s.E, invD))), SPAN_SLACK);
    const float dxR = sqrt_hw(fmaxf(0.f, __fmul_rn(__fmul_rn(__fmul_rn(4.f, c.cC), c.thr), invD)));
    s.dyR = __fmul_rn(c.ky, dxR);
    return s;
}
// tiles [*first, *first + return value) of tile row ty, clipped to the rect columns [x0r, x1r)
__device__ __forceinline__ int row_span(const SpanParams& s, const int ty, const int x0r, const int x1r, int* first) {
    const float y0 = (float)(ty * TILE);
    const float dyl = __fsub_rn(s.my, y0 + (float)(TILE - 1)), dyh = __fsub_rn(s.my, y0);
    *first = x0r;
    if (!s.any || dyl > s.dyT || dyh < -s.dyT) return 0;
    const float d1 = fminf(fmaxf(s.dyR, dyl), dyh);
    const float d2 = fminf(fmaxf(-s.dyR, dyl), dyh);
    const float r1 = sqrt_hw(fmaxf(0.f, __fmaf_rn(-s.D, __fmul_rn(d1, d1), s.E)));
    const float r2 = sqrt_hw(fmaxf(0.f, __fmaf_rn(-s.D, __fmul_rn(d2, d2), s.E)));
    const float dxmax = __fmaf_rn(s.kx, d1, __fmul_rn(r1, s.hA));
    const float dxmin = __fmaf_rn(s.kx, d2, -__fmul_rn(r2, s.hA));
    const float xa = __fsub_rn(__fsub_rn(s.mx, dxmax), SPAN_SLACK);      // leftmost / rightmost pixel abscissa reached
    const float xb = __fadd_rn(__fsub_rn(s.mx, dxmin), SPAN_SLACK);
    // tile tx covers pixels [16 tx, 16 tx + 15]
    const float fa = ceilf(__fmul_rn(__fsub_rn(xa, (float)(TILE - 1)), 1.f / TILE));
    const float fb = floorf(__fmul_rn(xb, 1.f / TILE));
    const int a = max(x0r, (int)fmaxf(fa, -1.f));
    const int b = min(x1r - 1, (int)fminf(fb, 70000.f));
    *first = a;
    return max(0, b - a + 1);
}

// may any pixel of tile (tx, ty) see alpha >= 1/255 from this Gaussian?
__device__ __forceinline__ bool tile_may_contribute(const CullParams& c, const int tx, const int ty) {
    const float x0 = (float)(tx * TILE), y0 = (float)(ty * TILE);
    return max_power_over_rect(c.cA, c.cB, c.cC, c.kx, c.ky, c.mx, c.my, x0, x0 + (float)(TILE - 1), y0,
                               y0 + (float)(TILE - 1)) >= c.thr;
}
#endif

// ---- scratch carving (256-B aligned sub-allocations) ----
__host__ __device__ inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

constexpr int RADIX_THREADS = 256;
constexpr int RADIX_ITEMS = 16;                                // items per thread
constexpr int RADIX_TILE = RADIX_THREADS * RADIX_ITEMS;        // 4096 items per block
constexpr int RADIX_BINS = 256;
constexpr int SCAN_BLOCK = 1024;

constexpr int MAX_TILE_PASSES = 4;   // tile ids are < 2^32

inline uint32_t radix_blocks(int64_t n) { return (uint32_t)((n + RADIX_TILE - 1) / RADIX_TILE); }

// Per-pass "onesweep" state, all zero before the pass starts:
//   [0, 256)                       global digit histogram of the pass
//   [256]                          dynamic tile ticket;  [257..259] padding
//   [260, 260 + 256 * nblocks)     decoupled look-back words: flag (2 msb) | count (30 lsb), [block][digit]
inline size_t sweep_pass_words(int64_t n) { return 260 + (size_t)RADIX_BINS * radix_blocks(n > 0 ? n : 1); }
inline size_t sweep_words(int64_t n, int passes) { return sweep_pass_words(n) * (size_t)passes; }
inline size_t scan_state_bytes(int64_t n) { return ((size_t)(n > 0 ? n : 1) / 4096 + 3) * 8; }

struct GeometryLayout {
    Record* rec;              // [P]
    uint32_t* depth_key;      // == sort_keys[0]: float bits of view depth, CULLED_KEY behind the near plane
    uint32_t* tiles_touched;  // [P]
    uint2* rect;              // [P] {minx | miny << 16, maxx | maxy << 16}
    unsigned long long* tile_mask;  // [P] bit (ty * rect_w + tx) = tile of the rect survives culling (rects <= 64 tiles)
    uint32_t* sort_keys[2];   // [P] ping-pong
    uint32_t* sort_vals[2];   // [P] ping-pong (Gaussian ids); final depth order in sort_vals[0]
    uint32_t* offsets;        // [P] inclusive scan of tiles_touched in depth order
    uint32_t* sweep;          // onesweep state of the 4 depth passes, zeroed once per forward (sweep_words(P, 4))
    unsigned long long* scan_state;  // chained-scan ticket + look-back words; directly after `sweep`, zeroed with it
    ScreenGrad* screen_grad;  // [P] (used by backward only; lives here so backward allocates nothing)
    int64_t* status;          // [2] {R, overflow}
    size_t bytes;
};

inline GeometryLayout carve_geometry(void* base, int32_t P) {
    GeometryLayout L;
    size_t o = 0;
    char* b = (char*)base;
    size_t Pa = P > 0 ? (size_t)P : 1;
    auto take = [&](size_t n) { char* p = b + o; o += align_up(n); return (void*)p; };
    L.status = (int64_t*)take(2 * sizeof(int64_t));
    L.rec = (Record*)take(Pa * sizeof(Record));
    L.tiles_touched = (uint32_t*)take(Pa * 4);
    L.rect = (uint2*)take(Pa * 8);
    L.tile_mask = (unsigned long long*)take(Pa * 8);
    for (int i = 0; i < 2; i++) L.sort_keys[i] = (uint32_t*)take(Pa * 4);
    for (int i = 0; i < 2; i++) L.sort_vals[i] = (uint32_t*)take(Pa * 4);
    L.depth_key = L.sort_keys[0];   // the depth-key kernel writes the sort input in place
    L.offsets = (uint32_t*)take(Pa * 4);
    L.sweep = (uint32_t*)take(align_up(sweep_words(Pa, 4) * 4));
    L.scan_state = (unsigned long long*)take(scan_state_bytes(Pa));   // contiguous with sweep (both 256-B multiples)
    L.screen_grad = (ScreenGrad*)take(Pa * sizeof(ScreenGrad));
    L.bytes = o;
    return L;
}

struct BinningLayout {
    uint32_t* keys[2];        // [capacity] tile ids, ping-pong
    uint32_t* vals[2];        // [capacity] Gaussian ids, ping-pong
    uint32_t* sweep;          // onesweep state of the (<= 2) tile passes (sweep_words(capacity, 2))
    uint2* ranges;            // [tiles]
    size_t bytes;
};

inline BinningLayout carve_binning(void* base, int32_t W, int32_t H, int64_t capacity) {
    BinningLayout L;
    size_t o = 0;
    char* b = (char*)base;
    size_t C = capacity > 0 ? (size_t)capacity : 1;
    size_t tiles = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    if (tiles == 0) tiles = 1;
    auto take = [&](size_t n) { char* p = b + o; o += align_up(n); return (void*)p; };
    L.ranges = (uint2*)take(tiles * sizeof(uint2));
    for (int i = 0; i < 2; i++) L.keys[i] = (uint32_t*)take(C * 4);
    for (int i = 0; i < 2; i++) L.vals[i] = (uint32_t*)take(C * 4);
    L.sweep = (uint32_t*)take(sweep_words(C, MAX_TILE_PASSES) * 4);
    L.bytes = o;
    return L;
}

struct ImageLayout {
    uint32_t* n_contrib;      // [H*W]
    float* final_T;           // [H*W]
    uint4* tile_todo;         // [tiles] max n_contrib of the (up to 4) forward work items of the tile
    uint32_t* tile_order;     // [tiles] tiles by descending work, built by the backward (longest first)
    size_t bytes;
};

inline ImageLayout carve_image(void* base, int32_t W, int32_t H) {
    ImageLayout L;
    size_t o = 0;
    char* b = (char*)base;
    size_t N = (size_t)W * H;
    if (N == 0) N = 1;
    auto take = [&](size_t n) { char* p = b + o; o += align_up(n); return (void*)p; };
    L.n_contrib = (uint32_t*)take(N * 4);
    L.final_T = (float*)take(N * 4);
    size_t tiles = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    if (tiles == 0) tiles = 1;
    L.tile_todo = (uint4*)take(tiles * sizeof(uint4));
    L.tile_order = (uint32_t*)take(tiles * 4);
    L.bytes = o;
    return L;
}

// ---- kernel launchers (defined in the .cu files) ----
struct Launch {
    cudaStream_t stream;
    bool debug;
};

// ---- programmatic dependent launch (SCGR_PDL=1) ----
// The kernels of a stage form a chain on one stream (scan -> emission -> partition passes -> render; prologue -> render
// backward -> preprocess backward; depth keys -> radix passes).  Launched through chain(...) with the programmatic
// stream-serialization attribute, a kernel's CTAs become resident while the tail of its predecessor is still running and
// block in griddepcontrol.wait -- the FIRST statement of every chained kernel -- until the predecessor has completed and
// its writes are visible: the grid launch latency and the kernel's own prologue leave the critical path, nothing else
// changes (no chained kernel touches memory before the wait).  Every chained kernel also releases its own successor
// at once (pdl_trigger): the successor can only take SM resources this grid no longer needs, since the trigger fires
// when ALL of this grid's CTAs have started.  Without the attribute (switch off, a stream being captured into a graph, a
// predecessor that is not a kernel) both instructions do nothing and the launch is an ordinary one.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifdef SCGR_HOST_EMULATION
template <class... KA>
struct ChainLaunch {
    void (*kernel)(KA...);
    dim3 grid;
    unsigned threads;
    template <class... A>
    void operator()(A... args) const {
        void (*k)(KA...) = kernel;
        emu_launch(grid, threads, [=] { k(args...); });
    }
};
template <class... KA>
inline ChainLaunch<KA...> chain(void (*kernel)(KA...), dim3 grid, dim3 block, size_t, const Launch&) {
    return ChainLaunch<KA...>{kernel, grid, block.x};
}
#else
bool pdl_allowed(cudaStream_t stream);      // capi.cu: the switch, and the stream is not being captured
template <class... KA>
struct ChainLaunch {
    void (*kernel)(KA...);
    dim3 grid, block;
    size_t smem;
    cudaStream_t stream;
    template <class... A>
    void operator()(A&&... args) const {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr = {};
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = pdl_allowed(stream) ? 1u : 0u;
        (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<A&&>(args)...);      // errors surface in check_launch
    }
};
template <class... KA>
inline ChainLaunch<KA...> chain(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, const Launch& L) {
    return ChainLaunch<KA...>{kernel, grid, block, smem, L.stream};
}
#endif

void launch_preprocess_forward(const ScgrView& v, const ScgrGaussians& g, const GeometryLayout& G,
                               int32_t* radii, const Launch& L);
void launch_preprocess_backward(const ScgrView& v, const ScgrGaussians& g, const GeometryLayout& G,
                                const ScgrGrads& out, const Launch& L);
void launch_mark_visible(const float* means3D, int32_t P, const float* viewmatrix, uint8_t* present,
                         const Launch& L);

void launch_depth_keys(const ScgrView& v, const ScgrGaussians& g, const GeometryLayout& G, const Launch& L);
void launch_depth_sort(const ScgrView& v, const ScgrGaussians& g, const GeometryLayout& G, const Launch& L);
void launch_scan_offsets(const GeometryLayout& G, int32_t P, int64_t* status_mapped, const Launch& L);
void launch_binning_prologue(const ScgrView& v, const BinningLayout& B, int32_t P, int64_t capacity, const Launch& L);
void launch_emit_and_partition(const ScgrView& v, const GeometryLayout& G, const BinningLayout& B,
                               int32_t P, int64_t capacity, int* final_buffer, const Launch& L);

void launch_render_forward(const ScgrView& v, const GeometryLayout& G, const BinningLayout& B,
                           const uint32_t* point_list, int64_t capacity, const ImageLayout& I,
                           float* out_color, float* out_depth, float* out_alpha, const Launch& L);
void launch_render_backward(const ScgrView& v, const GeometryLayout& G, const BinningLayout& B,
                            const uint32_t* point_list, int64_t capacity, const ImageLayout& I,
                            const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                            int32_t P, const Launch& L);
int tile_partition_final_buffer(uint32_t n_tiles);

// the gradient all-reduce through NVSwitch multicast (collective.cu)
void launch_nvls_allreduce(void* multicast_ptr, size_t n_floats, int rank, int world, const Launch& L);
void launch_nvls_allreduce_rows(void* multicast_rows, const float* live_count, long long n_rows, int row_floats, int rank,
                                int world, const Launch& L);
void launch_nvls_allreduce_fused(const ScgrNvlsFused& f, const Launch& L);

// mean squared distance to the 3 nearest neighbours (knn.cu)
void launch_knn3(const float* points, int32_t n, float* out, const Launch& L);

// activations + hybrid assembly, Adam (model.cu)
void launch_assemble_forward(const ScgrModel& m, const ScgrActivated& out, const Launch& L);
void launch_assemble_backward(const ScgrModel& m, const ScgrActivatedGrads& g, const ScgrModelGrads& out,
                              const Launch& L);
void launch_adam(const ScgrAdamGroup* groups, int32_t n_groups, double beta1, double beta2, double eps,
                 const Launch& L);
void launch_gather_rows(const ScgrRowGather* arrays, int32_t n_arrays, const int64_t* index, int64_t n_out,
                        const Launch& L);
void launch_copy_segments(const ScgrSegmentCopy* segs, int32_t n_segs, const Launch& L);
void launch_densification_stats(const float* dL_dmeans2D, const uint8_t* update_filter, const int32_t* radii, int32_t P,
                                float* xyz_gradient_accum, float* denom, float* max_radii2D, const Launch& L);

// fused photometric loss (loss.cu)
size_t photometric_scratch_bytes(int C, int H, int W);
void launch_photometric_forward(const float* img, const float* gt, int C, int H, int W, float lambda, void* scratch,
                                bool want_grad, float* out3, const Launch& L);
void launch_photometric_backward(const float* img, const float* gt, int C, int H, int W, float lambda,
                                 const void* scratch, const float* upstream, float* dL_dimg, const Launch& L);

// match-prior loss, DTU background term (prior.cu)
void launch_match_loss_forward(const float* depth, int H, int W, float width, float height, const ScgrMatchPair* pairs,
                               int n_pairs, float* scratch, float* out, const Launch& L);
void launch_match_loss_backward(const float* depth, int H, int W, float width, float height, const ScgrMatchPair* pairs,
                                int n_pairs, const float* scratch, const float* upstream, float* dL_ddepth, const Launch& L);
void launch_bg_mask(float* gt, int C, int H, int W, float threshold, int window, uint8_t* mask, float* count, const Launch& L);
size_t masked_mean_scratch_bytes(size_t n);
void launch_masked_mean_forward(const float* values, const uint8_t* mask, size_t n, void* scratch, float* out2, const Launch& L);
void launch_masked_mean_backward(const uint8_t* mask, size_t n, const float* out2, const float* upstream, float* dL_dvalues,
                                 const Launch& L);

// Every kernel launch is bracketed:  begin_kernel(name, L); kernel<<<...>>>(...) or chain(kernel, ...)(...); check_launch(name, L);
// begin_kernel records a start event when profiling is on (scgr_profile_enable); check_launch
// counts the launch, records the stop event, and turns CUDA errors into std::runtime_error
// (synchronising first when the view's `debug` flag is set, like the reference's CHECK_CUDA).
void begin_kernel(const char* what, const Launch& L);
void check_launch(const char* what, const Launch& L);
void check_stage(const char* what, const Launch& L);

// ---- TMA (cp.async.bulk) + mbarrier plumbing, shared by render.cu (every lane pulls the 48-byte record of its Gaussian
// into the warp's staging buffer) and preprocess.cu (every thread pulls the 192-byte SH row of its Gaussian): one
// bulk-async copy per lane, completion is a transaction count on an mbarrier, no register holds data in flight. ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Per-thread asynchronous copies (LDGSTS): unlike a bulk copy, which takes its addresses from uniform registers and is
// therefore issued one lane at a time when every lane has its own source (8 instructions x 32 rounds per warp), one
// cp.async serves the 32 lanes of a warp at once.  Completion is per thread (wait_group), then a warp barrier.
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// generic-proxy accesses of a staging buffer (reads, in-place writes) ordered before the async-proxy writes that refill it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// arrival without bytes: a thread that takes part in the phase but issues no copy
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SCGR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SCGR_DONE;\n"
        "bra SCGR_WAIT;\n"
        "SCGR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}


}  // namespace scgr
