// capi.cu -- the C ABI of libscgr.so (include/scgr.h): argument checking, scratch carving,
// stage orchestration, error translation.  Replaces the external extension's
// RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible + Rasterizer::forward /
// backward without torch types, allocations or host syncs.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace scgr {

static thread_local std::string g_last_error;

// ---- launch accounting + optional per-kernel event timing (scgr_profile_*) ----
struct ProfEntry {
    const char* name;
    cudaEvent_t start, stop;
    bool stop_pending;
};
static std::mutex g_prof_mutex;
static std::vector<ProfEntry> g_prof;
static std::atomic<bool> g_prof_on{false};
static std::atomic<long long> g_kernel_launches{0};

// SCGR_PDL (default 1): programmatic dependent launch between the kernels of a stage (common.cuh); B200, config 3:
// 1.035 -> 1.026 ms per step.  Never inside a stream capture: graph kernel nodes are already launched back to back by
// the device.
#ifndef SCGR_HOST_EMULATION
bool pdl_allowed(cudaStream_t stream) {
    static const bool on = getenv("SCGR_PDL") ? atoi(getenv("SCGR_PDL")) != 0 : true;
    if (!on) return false;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return st == cudaStreamCaptureStatusNone;
}
#endif

void begin_kernel(const char* what, const Launch& L) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfEntry e{what, nullptr, nullptr, true};
    cudaEventCreate(&e.start);
    cudaEventCreate(&e.stop);
    cudaEventRecord(e.start, L.stream);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof.push_back(e);
}

void check_stage(const char* what, const Launch& L) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && L.debug) e = cudaStreamSynchronize(L.stream);   // reference `debug` semantics
    if (e != cudaSuccess)
        throw std::runtime_error(std::string("scgr: ") + what + ": " + cudaGetErrorString(e));
}

void check_launch(const char* what, const Launch& L) {
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    if (g_prof_on.load(std::memory_order_relaxed)) {
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        // the matching begin_kernel pushed the last entry with this name
        for (size_t i = g_prof.size(); i-- > 0;)
            if (std::strcmp(g_prof[i].name, what) == 0 && g_prof[i].stop_pending) { g_prof[i].stop_pending = false; cudaEventRecord(g_prof[i].stop, L.stream); break; }
    }
    check_stage(what, L);
}

static void require(bool cond, const char* msg) {
    if (!cond) throw std::invalid_argument(std::string("scgr: ") + msg);
}

static void validate(const ScgrView* v, const ScgrGaussians* g) {
    require(v && g, "null view / gaussians");
    require(v->image_width >= 0 && v->image_height >= 0, "negative image size");
    require((v->image_width + TILE - 1) / TILE < 65536 && (v->image_height + TILE - 1) / TILE < 65536,
            "image too large for 16-bit tile coordinates");
    require(g->P >= 0, "negative P");
    if (g->P == 0) return;
    require(v->bg && v->viewmatrix && v->projmatrix && v->campos, "null camera tensors");
    require(g->means3D && g->opacities, "null means3D / opacities");
    const bool split = g->sh_dc[0] != nullptr || g->sh_dc[1] != nullptr;
    require((int)(g->shs != nullptr) + (int)(g->colors_precomp != nullptr) + (int)split == 1,
            "provide exactly one of shs / colors_precomp");
    if (split) {
        require(g->sh_coeffs == 16, "split SH arrays (sh_dc / sh_rest) need 16 coefficients per Gaussian");
        require(g->sh_n0 >= 0 && g->sh_n0 <= g->P, "split SH arrays: sh_n0 out of range");
        require(g->sh_n0 == 0 || (g->sh_dc[0] && g->sh_rest[0]), "split SH arrays: set 0 is missing");
        require(g->sh_n0 == g->P || (g->sh_dc[1] && g->sh_rest[1]), "split SH arrays: set 1 is missing");
    }
    const bool sr = g->scales != nullptr && g->rotations != nullptr;
    require(sr != (g->cov3D_precomp != nullptr) && (sr || (!g->scales && !g->rotations)),
            "provide exactly one of (scales, rotations) / cov3D_precomp");
    if (g->shs || split) {
        require(v->sh_degree >= 0 && v->sh_degree <= 3, "sh_degree must be 0..3");
        require(g->sh_coeffs >= (v->sh_degree + 1) * (v->sh_degree + 1), "shs has too few coefficients for sh_degree");
    }
    if (g->rotations) require((reinterpret_cast<uintptr_t>(g->rotations) & 15) == 0, "rotations must be 16-byte aligned");
}

template <typename F>
static int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return 1;
    } catch (...) {
        g_last_error = "scgr: unknown error";
        return 2;
    }
}

// ---- auxiliary stream: stage 1 forks -- depth keys + depth sort run beside the heavy preprocess
// kernel and join it before the prefix sum.  One stream and one fork/join event pair per host
// thread and device (thread-local: concurrent callers never share events), created on first use
// and kept for the life of the thread. ----
struct AuxStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static AuxStream& aux_stream() {
    constexpr int kMaxDevices = 64;
    static thread_local AuxStream aux[kMaxDevices];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) throw std::runtime_error("scgr: cannot query the current device");
    AuxStream& a = aux[dev];
    if (!a.stream) {
        // highest priority: the sort's few, short CTAs must not queue behind the thousands of preprocess CTAs
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { (void)cudaGetLastError(); greatest = 0; }
        if (cudaStreamCreateWithPriority(&a.stream, cudaStreamNonBlocking, greatest) != cudaSuccess ||
            cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming) != cudaSuccess) {
            a = AuxStream{};
            throw std::runtime_error(std::string("scgr: cannot create the auxiliary stream: ") + cudaGetErrorString(cudaGetLastError()));
        }
    }
    return a;
}

// stage 1 proper (shared by scgr_forward_geometry and scgr_forward): P > 0
// `early`: the R-independent part of stage 2 (empty ranges, zeroed onesweep words of the tile partition) when the caller
// already holds a binning buffer: enqueued behind the preprocess, where the main stream otherwise idles until the depth
// sort on the auxiliary stream is done (0.10 ms against 0.08 ms) -- not between the scan and the emission.
static void enqueue_geometry_stage(const ScgrView* view, const ScgrGaussians* g, const GeometryLayout& G, int32_t* radii,
                                   int64_t* status_mapped, const Launch& L, const BinningLayout* early = nullptr,
                                   int64_t early_capacity = 0) {
    static const bool serial = getenv("SCGR_SERIAL_STAGE1") != nullptr;    // A/B switch: everything on the caller's stream
    if (serial) {
        launch_depth_sort(*view, *g, G, L);
        launch_preprocess_forward(*view, *g, G, radii, L);
    } else {
        AuxStream& a = aux_stream();
        const Launch La{a.stream, L.debug};
        cudaEventRecord(a.fork, L.stream);                 // inputs are ready at this point of the caller's stream
        cudaStreamWaitEvent(a.stream, a.fork, 0);
        launch_depth_sort(*view, *g, G, La);               // aux:  sweep memset, depth keys + histograms, 4 radix passes
        cudaEventRecord(a.join, a.stream);
        launch_preprocess_forward(*view, *g, G, radii, L); // main: project, covariance, SH, tile counts
        if (early) launch_binning_prologue(*view, *early, g->P, early_capacity, L);
        cudaStreamWaitEvent(L.stream, a.join, 0);
    }
    if (early && serial) launch_binning_prologue(*view, *early, g->P, early_capacity, L);
    launch_scan_offsets(G, g->P, status_mapped, L);
}

static void copy_status(const GeometryLayout& G, int64_t* status_host, cudaStream_t s) {
    if (status_host) cudaMemcpyAsync(status_host, G.status, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, s);
}

}  // namespace scgr

using namespace scgr;

extern "C" {

int scgr_version(void) { return SCGR_VERSION; }
const char* scgr_last_error(void) { return g_last_error.c_str(); }

size_t scgr_geometry_bytes(int32_t P) { return carve_geometry(nullptr, P).bytes; }
size_t scgr_binning_bytes(int32_t P, int32_t W, int32_t H, int64_t capacity) {
    (void)P;
    return carve_binning(nullptr, W, H, capacity).bytes;
}
size_t scgr_image_bytes(int32_t W, int32_t H) { return carve_image(nullptr, W, H).bytes; }

int scgr_forward_geometry(const ScgrView* view, const ScgrGaussians* g, void* geometry_scratch,
                          int32_t* radii, int64_t* status_host, scgr_stream_t stream) {
    return guarded([&] {
        validate(view, g);
        require(geometry_scratch != nullptr, "null geometry scratch");
        require((reinterpret_cast<uintptr_t>(geometry_scratch) & 255) == 0, "geometry scratch must be 256-byte aligned");
        const Launch L{(cudaStream_t)stream, view->debug != 0};
        const GeometryLayout G = carve_geometry(geometry_scratch, g->P);
        if (g->P == 0) {
            cudaMemsetAsync(G.status, 0, 2 * sizeof(int64_t), L.stream);
        } else {
            require(radii != nullptr, "null radii");
            enqueue_geometry_stage(view, g, G, radii, nullptr, L);
        }
        copy_status(G, status_host, L.stream);
        check_stage("forward_geometry", L);
    });
}

// stage 2 proper (shared by scgr_forward_render and scgr_forward)
static void enqueue_render_stage(const ScgrView* view, const ScgrGaussians* g, void* geometry_scratch,
                                 void* binning_scratch, int64_t capacity, void* image_scratch, float* out_color,
                                 float* out_depth, float* out_alpha, bool prologue_done, const Launch& L) {
    const int W = view->image_width, H = view->image_height;
    const GeometryLayout G = carve_geometry(geometry_scratch, g->P);
    const BinningLayout B = carve_binning(binning_scratch, W, H, capacity);
    const ImageLayout I = carve_image(image_scratch, W, H);
    if (!prologue_done) launch_binning_prologue(*view, B, g->P, capacity, L);
    if (g->P == 0) {
        // section 8b: P = 0 returns all-zero images (not background-filled)
        const size_t N = (size_t)W * H;
        cudaMemsetAsync(out_color, 0, 3 * N * sizeof(float), L.stream);
        cudaMemsetAsync(out_depth, 0, N * sizeof(float), L.stream);
        cudaMemsetAsync(out_alpha, 0, N * sizeof(float), L.stream);
        cudaMemsetAsync(I.n_contrib, 0, N * sizeof(uint32_t), L.stream);
        cudaMemsetAsync(I.final_T, 0, N * sizeof(float), L.stream);
    } else {
        int fin = 0;
        launch_emit_and_partition(*view, G, B, g->P, capacity, &fin, L);
        launch_render_forward(*view, G, B, B.vals[fin], capacity, I, out_color, out_depth, out_alpha, L);
    }
}

int scgr_forward_render(const ScgrView* view, const ScgrGaussians* g, void* geometry_scratch,
                        void* binning_scratch, int64_t capacity, void* image_scratch,
                        float* out_color, float* out_depth, float* out_alpha, int64_t* status_host,
                        scgr_stream_t stream) {
    return guarded([&] {
        validate(view, g);
        require(geometry_scratch && binning_scratch && image_scratch, "null scratch");
        require(out_color && out_depth && out_alpha, "null outputs");
        require(capacity >= 0 && capacity < (int64_t)0xFFFFFFFFll, "capacity out of range");
        const Launch L{(cudaStream_t)stream, view->debug != 0};
        enqueue_render_stage(view, g, geometry_scratch, binning_scratch, capacity, image_scratch, out_color,
                             out_depth, out_alpha, false, L);
        copy_status(carve_geometry(geometry_scratch, g->P), status_host, L.stream);
        check_stage("forward_render", L);
    });
}

// Host wait for R between the two stages of scgr_forward().  The scan kernel stores R straight into
// the caller's pinned word (zero-copy), so the host sees it a PCIe write after the kernel retires
// and stage 2 is enqueued a few microseconds later -- no D2H memcpy, no stream synchronisation, no
// return to the caller's language in between.  A failed stream is detected by polling it.
static int64_t wait_for_count(volatile int64_t* word, int64_t sentinel, cudaStream_t stream) {
    for (uint64_t spins = 1;; spins++) {
        const int64_t v = *word;
        if (v != sentinel) return v;
        if ((spins & 0x3ff) == 0) {
            const cudaError_t q = cudaStreamQuery(stream);
            if (q == cudaErrorNotReady) continue;
            if (q != cudaSuccess) throw std::runtime_error(std::string("scgr: forward stage 1: ") + cudaGetErrorString(q));
            const int64_t w = *word;   // stream drained: the store must be visible by now
            if (w != sentinel) return w;
            throw std::runtime_error("scgr: forward stage 1 finished without publishing num_rendered");
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
}

int scgr_forward(const ScgrView* view, const ScgrGaussians* g, void* geometry_scratch, int32_t* radii,
                 void* binning_scratch, int64_t capacity, void* image_scratch, float* out_color,
                 float* out_depth, float* out_alpha, int64_t* status_host, scgr_stream_t stream) {
    bool need_capacity = false;
    const int rc = guarded([&] {
        validate(view, g);
        require(geometry_scratch && image_scratch, "null scratch");
        require((reinterpret_cast<uintptr_t>(geometry_scratch) & 255) == 0, "geometry scratch must be 256-byte aligned");
        require(out_color && out_depth && out_alpha, "null outputs");
        require(status_host != nullptr, "scgr_forward needs a pinned host status word");
        require(capacity >= 0 && capacity < (int64_t)0xFFFFFFFFll, "capacity out of range");
        const Launch L{(cudaStream_t)stream, view->debug != 0};
        const GeometryLayout G = carve_geometry(geometry_scratch, g->P);
        int64_t R = 0;
        bool prologue_done = false, stage2_enqueued = false;
        if (g->P == 0) {
            cudaMemsetAsync(G.status, 0, 2 * sizeof(int64_t), L.stream);
            status_host[0] = 0;
            status_host[1] = 0;
        } else {
            require(radii != nullptr, "null radii");
            constexpr int64_t kSentinel = INT64_MIN;
            void* mapped = nullptr;
            const bool zero_copy = cudaHostGetDevicePointer(&mapped, status_host, 0) == cudaSuccess && mapped != nullptr;
            if (!zero_copy) (void)cudaGetLastError();
            status_host[1] = 0;
            *(volatile int64_t*)status_host = kSentinel;
            if (binning_scratch) {   // R-independent part of stage 2: enqueued inside stage 1, behind the preprocess
                const BinningLayout Bearly = carve_binning(binning_scratch, view->image_width, view->image_height, capacity);
                enqueue_geometry_stage(view, g, G, radii, zero_copy ? (int64_t*)mapped : nullptr, L, &Bearly, capacity);
                prologue_done = true;
            } else {
                enqueue_geometry_stage(view, g, G, radii, zero_copy ? (int64_t*)mapped : nullptr, L);
            }
            if (zero_copy && binning_scratch) {
                // Stage 2 is enqueued BEFORE R is known: its kernels read R from device memory and do nothing
                // when it exceeds `capacity`, so the GPU never idles between the stages.  The host then only
                // waits (briefly: R lands while stage 2 is still running) to tell the caller which case it was.
                enqueue_render_stage(view, g, geometry_scratch, binning_scratch, capacity, image_scratch, out_color,
                                     out_depth, out_alpha, prologue_done, L);
                stage2_enqueued = true;
            }
            if (zero_copy) {
                R = wait_for_count((volatile int64_t*)status_host, kSentinel, L.stream);
            } else {   // status_host is not device-mapped: the reference's protocol (copy + synchronise)
                copy_status(G, status_host, L.stream);
                const cudaError_t e = cudaStreamSynchronize(L.stream);
                if (e != cudaSuccess) throw std::runtime_error(std::string("scgr: forward stage 1: ") + cudaGetErrorString(e));
                R = status_host[0];
            }
        }
        if (g->P != 0 && (binning_scratch == nullptr || R > capacity)) {
            need_capacity = true;     // stage 1 is complete and stays valid: size the buffer, call scgr_forward_render
            return;                   // (an eagerly enqueued stage 2 has refused to run on the device)
        }
        require(binning_scratch != nullptr, "null binning scratch");
        if (!stage2_enqueued)
            enqueue_render_stage(view, g, geometry_scratch, binning_scratch, capacity, image_scratch, out_color, out_depth,
                                 out_alpha, prologue_done, L);
        check_stage("forward", L);
    });
    if (rc == 0 && need_capacity) {
        g_last_error = "scgr: binning capacity too small (not an error: call scgr_forward_render with capacity >= status_host[0])";
        return SCGR_NEED_CAPACITY;
    }
    return rc;
}

int scgr_backward(const ScgrView* view, const ScgrGaussians* g, const void* geometry_scratch,
                  const void* binning_scratch, int64_t capacity, const void* image_scratch,
                  const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                  const ScgrGrads* grads, scgr_stream_t stream) {
    return guarded([&] {
        validate(view, g);
        require(grads != nullptr, "null grads");
        if (g->P == 0) return;
        require(geometry_scratch && binning_scratch && image_scratch, "null scratch");
        require(dL_dcolor && dL_ddepth && dL_dalpha, "null upstream gradients");
        require(grads->dL_dmeans3D && grads->dL_dmeans2D && grads->dL_dopacities, "null gradient outputs");
        require((g->shs != nullptr) == (grads->dL_dshs != nullptr), "dL_dshs must match shs");
        for (int k = 0; k < 2; k++)
            require((g->sh_dc[k] != nullptr) == (grads->dL_dsh_dc[k] != nullptr) &&
                    (g->sh_rest[k] != nullptr) == (grads->dL_dsh_rest[k] != nullptr), "dL_dsh_dc / dL_dsh_rest must match sh_dc / sh_rest");
        require((g->colors_precomp != nullptr) == (grads->dL_dcolors_precomp != nullptr), "dL_dcolors_precomp must match colors_precomp");
        require((g->scales != nullptr) == (grads->dL_dscales != nullptr) &&
                (g->rotations != nullptr) == (grads->dL_drotations != nullptr), "dL_dscales / dL_drotations must match inputs");
        require((g->cov3D_precomp != nullptr) == (grads->dL_dcov3D_precomp != nullptr), "dL_dcov3D_precomp must match cov3D_precomp");
        if (grads->dL_drotations) require((reinterpret_cast<uintptr_t>(grads->dL_drotations) & 15) == 0, "dL_drotations must be 16-byte aligned");
        if (grads->densification_stats) {
            require(grads->radii != nullptr, "densification_stats needs radii");
            require((reinterpret_cast<uintptr_t>(grads->densification_stats) & 7) == 0, "densification_stats must be 8-byte aligned");
        }
        const Launch L{(cudaStream_t)stream, view->debug != 0};
        const int W = view->image_width, H = view->image_height;
        const GeometryLayout G = carve_geometry(const_cast<void*>(geometry_scratch), g->P);
        const BinningLayout B = carve_binning(const_cast<void*>(binning_scratch), W, H, capacity);
        const ImageLayout I = carve_image(const_cast<void*>(image_scratch), W, H);
        const uint32_t n_tiles = (uint32_t)((W + TILE - 1) / TILE) * (uint32_t)((H + TILE - 1) / TILE);
        const int fin = tile_partition_final_buffer(n_tiles);
        launch_render_backward(*view, G, B, B.vals[fin], capacity, I, dL_dcolor, dL_ddepth, dL_dalpha, g->P, L);
        launch_preprocess_backward(*view, *g, G, *grads, L);
        check_stage("backward", L);
    });
}

int scgr_mark_visible(const float* means3D, int32_t P, const float* viewmatrix, uint8_t* present,
                      scgr_stream_t stream) {
    return guarded([&] {
        require(P >= 0, "negative P");
        if (P == 0) return;
        require(means3D && viewmatrix && present, "null argument");
        const Launch L{(cudaStream_t)stream, false};
        launch_mark_visible(means3D, P, viewmatrix, present, L);
    });
}

size_t scgr_photometric_scratch_bytes(int32_t C, int32_t H, int32_t W) {
    if (C <= 0 || H <= 0 || W <= 0) return 256;
    return photometric_scratch_bytes(C, H, W);
}

int scgr_photometric_forward(const float* image, const float* gt, int32_t C, int32_t H, int32_t W, float lambda_dssim,
                             void* scratch, int32_t want_grad, float* out3, scgr_stream_t stream) {
    return guarded([&] {
        require(C > 0 && H > 0 && W > 0, "photometric loss: empty image");
        require((int64_t)C * (((int64_t)H + 15) / 16) < 65536 * 1024ll, "photometric loss: too many planes");
        require(image && gt && scratch && out3, "photometric loss: null argument");
        require((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "photometric loss: scratch must be 256-byte aligned");
        const Launch L{(cudaStream_t)stream, false};
        launch_photometric_forward(image, gt, C, H, W, lambda_dssim, scratch, want_grad != 0, out3, L);
    });
}

int scgr_photometric_backward(const float* image, const float* gt, int32_t C, int32_t H, int32_t W, float lambda_dssim,
                              const void* scratch, const float* upstream, float* dL_dimage, scgr_stream_t stream) {
    return guarded([&] {
        require(C > 0 && H > 0 && W > 0, "photometric loss: empty image");
        require(image && gt && scratch && dL_dimage, "photometric loss: null argument");
        const Launch L{(cudaStream_t)stream, false};
        launch_photometric_backward(image, gt, C, H, W, lambda_dssim, scratch, upstream, dL_dimage, L);
    });
}

int scgr_match_loss_forward(const float* depth, int32_t H, int32_t W, float width, float height, const ScgrMatchPair* pairs,
                            int32_t n_pairs, float* scratch, float* out, scgr_stream_t stream) {
    return guarded([&] {
        require(H > 0 && W > 0 && width > 0.f && height > 0.f, "match loss: empty image");
        require(n_pairs >= 0 && n_pairs <= SCGR_MATCH_MAX_PAIRS, "match loss: 0..SCGR_MATCH_MAX_PAIRS pairs per call");
        require(depth && scratch && out && (pairs || n_pairs == 0), "match loss: null argument");
        for (int i = 0; i < n_pairs; i++) {
            require(pairs[i].n >= 0, "match loss: negative match count");
            if (pairs[i].n > 0)
                require(pairs[i].uv0 && pairs[i].rays_o && pairs[i].rays_d && pairs[i].cam_rays_d && pairs[i].uv1,
                        "match loss: null array in a pair");
        }
        const Launch L{(cudaStream_t)stream, false};
        launch_match_loss_forward(depth, H, W, width, height, pairs, n_pairs, scratch, out, L);
    });
}

int scgr_match_loss_backward(const float* depth, int32_t H, int32_t W, float width, float height, const ScgrMatchPair* pairs,
                             int32_t n_pairs, const float* scratch, const float* upstream, float* dL_ddepth,
                             scgr_stream_t stream) {
    return guarded([&] {
        require(H > 0 && W > 0 && width > 0.f && height > 0.f, "match loss: empty image");
        require(n_pairs >= 0 && n_pairs <= SCGR_MATCH_MAX_PAIRS, "match loss: 0..SCGR_MATCH_MAX_PAIRS pairs per call");
        require(depth && scratch && dL_ddepth && (pairs || n_pairs == 0), "match loss: null argument");
        const Launch L{(cudaStream_t)stream, false};
        launch_match_loss_backward(depth, H, W, width, height, pairs, n_pairs, scratch, upstream, dL_ddepth, L);
    });
}

int scgr_bg_mask(float* gt, int32_t C, int32_t H, int32_t W, float threshold, int32_t window, uint8_t* mask, float* count,
                 scgr_stream_t stream) {
    return guarded([&] {
        require(C > 0 && H > 0 && W > 0, "bg mask: empty image");
        require(window >= 1, "bg mask: window must be >= 1");
        require(gt && mask && count, "bg mask: null argument");
        const Launch L{(cudaStream_t)stream, false};
        launch_bg_mask(gt, C, H, W, threshold, window, mask, count, L);
    });
}

size_t scgr_masked_mean_scratch_bytes(int64_t n) { return masked_mean_scratch_bytes(n > 0 ? (size_t)n : 1); }

int scgr_masked_mean_forward(const float* values, const uint8_t* mask, int64_t n, void* scratch, float* out2,
                             scgr_stream_t stream) {
    return guarded([&] {
        require(n > 0, "masked mean: empty input");
        require(values && mask && scratch && out2, "masked mean: null argument");
        require((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, "masked mean: scratch must be 16-byte aligned");
        const Launch L{(cudaStream_t)stream, false};
        launch_masked_mean_forward(values, mask, (size_t)n, scratch, out2, L);
    });
}

int scgr_masked_mean_backward(const uint8_t* mask, int64_t n, const float* out2, const float* upstream, float* dL_dvalues,
                              scgr_stream_t stream) {
    return guarded([&] {
        require(n > 0, "masked mean: empty input");
        require(mask && out2 && dL_dvalues, "masked mean: null argument");
        const Launch L{(cudaStream_t)stream, false};
        launch_masked_mean_backward(mask, (size_t)n, out2, upstream, dL_dvalues, L);
    });
}

int scgr_nvls_allreduce(void* multicast_ptr, size_t n_floats, int32_t rank, int32_t world, scgr_stream_t stream) {
    return guarded([&] {
        require(multicast_ptr != nullptr, "nvls all-reduce: null multicast pointer");
        require(world >= 1 && rank >= 0 && rank < world, "nvls all-reduce: bad rank / world size");
        require((reinterpret_cast<uintptr_t>(multicast_ptr) & 15) == 0, "nvls all-reduce: buffer must be 16-byte aligned");
        require(n_floats % (4 * (size_t)world) == 0, "nvls all-reduce: element count must be a multiple of 4 * world");
        const Launch L{(cudaStream_t)stream, false};
        launch_nvls_allreduce(multicast_ptr, n_floats, rank, world, L);
    });
}

int scgr_nvls_allreduce_rows(void* multicast_rows, const float* live_count, int64_t n_rows, int32_t row_floats,
                             int32_t rank, int32_t world, scgr_stream_t stream) {
    return guarded([&] {
        require(world >= 1 && rank >= 0 && rank < world, "nvls all-reduce: bad rank / world size");
        require(n_rows >= 0, "nvls all-reduce: negative row count");
        if (n_rows == 0) return;
        require(multicast_rows != nullptr && live_count != nullptr, "nvls all-reduce: null pointer");
        require((reinterpret_cast<uintptr_t>(multicast_rows) & 15) == 0, "nvls all-reduce: rows must be 16-byte aligned");
        require(row_floats > 0 && row_floats % 4 == 0, "nvls all-reduce: row_floats must be a positive multiple of 4");
        const Launch L{(cudaStream_t)stream, false};
        launch_nvls_allreduce_rows(multicast_rows, live_count, n_rows, row_floats, rank, world, L);
    });
}

int scgr_nvls_allreduce_fused(const ScgrNvlsFused* f, scgr_stream_t stream) {
    return guarded([&] {
        require(f != nullptr, "nvls all-reduce: null arguments");
        require(f->world >= 1 && f->world <= SCGR_NVLS_MAX_WORLD && f->rank >= 0 && f->rank < f->world,
                "nvls all-reduce: bad rank / world size (at most 8 ranks: one NVSwitch box)");
        require(f->multicast_ptr != nullptr && (reinterpret_cast<uintptr_t>(f->multicast_ptr) & 15) == 0,
                "nvls all-reduce: buffer must be a 16-byte aligned multicast address");
        require(f->dense_floats % (4 * (size_t)f->world) == 0, "nvls all-reduce: element count must be a multiple of 4 * world");
        if (f->multicast_rows) {
            require(f->n_rows >= 0 && f->live_count != nullptr, "nvls all-reduce: rows need their live counts");
            require((reinterpret_cast<uintptr_t>(f->multicast_rows) & 15) == 0, "nvls all-reduce: rows must be 16-byte aligned");
            require(f->row_floats > 0 && f->row_floats % 4 == 0, "nvls all-reduce: row_floats must be a positive multiple of 4");
        }
        require(f->sync_local != nullptr, "nvls all-reduce: null sync words");
        for (int q = 0; q < f->world; q++) require(f->flags[q] != nullptr, "nvls all-reduce: null flag array");
        const Launch L{(cudaStream_t)stream, false};
        launch_nvls_allreduce_fused(*f, L);
    });
}

int scgr_knn3_mean_dist2(const float* points, int32_t n, float* out, scgr_stream_t stream) {
    return guarded([&] {
        require(n >= 0, "knn3: negative point count");
        if (n == 0) return;
        require(points && out, "knn3: null argument");
        require(n <= (1 << 20), "knn3: exact all-pairs scan is meant for initial clouds of up to 2^20 points");
        const Launch L{(cudaStream_t)stream, false};
        launch_knn3(points, n, out, L);
    });
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static void validate_model(const ScgrModel* m) {
    require(m != nullptr, "assemble: null model");
    require(m->sh_rest >= 0 && m->sh_rest <= 4095, "assemble: sh_rest out of range");
    int64_t P = 0;
    for (int k = 0; k < 2; k++) {
        const ScgrModelSet& s = m->set[k];
        require(s.n >= 0, "assemble: negative set size");
        P += s.n;
        if (s.n == 0) continue;
        require((s.rayo != nullptr) != (s.xyz != nullptr), "assemble: a set has either xyz or (rayo, rayd, zval)");
        if (s.rayo) require(s.rayd && s.zval, "assemble: ray-based set needs rayo, rayd and zval");
        require(s.scaling && s.rotation && s.opacity && s.features_dc, "assemble: null parameter array");
        require(m->sh_rest == 0 || s.features_rest, "assemble: null features_rest");
        require(aligned16(s.rotation), "assemble: rotation must be 16-byte aligned");
    }
    require(P * 3 * (int64_t)(m->sh_rest + 1) < (int64_t(1) << 31), "assemble: model too large for 32-bit SH indexing");
}

int scgr_assemble_forward(const ScgrModel* model, const ScgrActivated* out, scgr_stream_t stream) {
    return guarded([&] {
        validate_model(model);
        if (model->set[0].n + model->set[1].n == 0) return;
        require(out && out->means3D && out->scales && out->rotations && out->opacities, "assemble: null output array");
        require(aligned16(out->rotations) && aligned16(out->shs), "assemble: outputs must be 16-byte aligned");      // (shs may be NULL: split SH layout)
        const Launch L{(cudaStream_t)stream, false};
        launch_assemble_forward(*model, *out, L);
    });
}

int scgr_assemble_backward(const ScgrModel* model, const ScgrActivatedGrads* grads, const ScgrModelGrads* out,
                           scgr_stream_t stream) {
    return guarded([&] {
        validate_model(model);
        if (model->set[0].n + model->set[1].n == 0) return;
        require(grads && grads->dL_dmeans3D && grads->dL_dscales && grads->dL_drotations && grads->dL_dopacities,
                "assemble: null incoming gradient");      // (dL_dshs may be NULL: split SH layout)
        require(aligned16(grads->dL_drotations) && aligned16(grads->dL_dshs),
                "assemble: incoming gradients must be 16-byte aligned");
        require(out != nullptr, "assemble: null gradient outputs");
        for (int k = 0; k < 2; k++) {
            const ScgrModelSet& s = model->set[k];
            const ScgrModelSetGrads& d = out->set[k];
            if (s.n == 0) continue;
            require(s.rayo ? d.dL_dzval != nullptr : d.dL_dxyz != nullptr, "assemble: null position gradient output");
            require(d.dL_dscaling && d.dL_drotation && d.dL_dopacity, "assemble: null gradient output");
            if (grads->dL_dshs) {
                require(d.dL_dfeatures_dc != nullptr, "assemble: null gradient output");
                require(model->sh_rest == 0 || d.dL_dfeatures_rest, "assemble: null dL_dfeatures_rest");
            }
            require(aligned16(d.dL_drotation), "assemble: dL_drotation must be 16-byte aligned");
        }
        const Launch L{(cudaStream_t)stream, false};
        launch_assemble_backward(*model, *grads, *out, L);
    });
}

int scgr_densification_stats(const float* dL_dmeans2D, const uint8_t* update_filter, const int32_t* radii, int32_t P,
                             float* xyz_gradient_accum, float* denom, float* max_radii2D, scgr_stream_t stream) {
    return guarded([&] {
        require(P >= 0, "densification_stats: negative P");
        if (P == 0) return;
        require(update_filter || radii, "densification_stats: need update_filter or radii");
        require((xyz_gradient_accum != nullptr) == (denom != nullptr),
                "densification_stats: xyz_gradient_accum and denom go together");
        require(!xyz_gradient_accum || dL_dmeans2D, "densification_stats: null dL_dmeans2D");
        require(xyz_gradient_accum || (max_radii2D && radii), "densification_stats: nothing to update");
        const Launch L{(cudaStream_t)stream, false};
        launch_densification_stats(dL_dmeans2D, update_filter, radii, P, xyz_gradient_accum, denom, max_radii2D, L);
    });
}

int scgr_gather_rows(const ScgrRowGather* arrays, int32_t n_arrays, const int64_t* index, int64_t n_out,
                     scgr_stream_t stream) {
    return guarded([&] {
        require(n_arrays >= 0 && n_arrays <= SCGR_GATHER_MAX_ARRAYS, "gather_rows: 0..SCGR_GATHER_MAX_ARRAYS arrays per call");
        require(n_out >= 0 && n_out < (int64_t(1) << 31), "gather_rows: n_out must be in [0, 2^31)");
        if (n_arrays == 0 || n_out == 0) return;
        require(arrays && index, "gather_rows: null argument");
        for (int a = 0; a < n_arrays; a++) {
            require(arrays[a].row_floats >= 0 && arrays[a].row_floats <= 65536, "gather_rows: row_floats out of range");
            if (arrays[a].row_floats == 0) continue;
            require(arrays[a].src && arrays[a].dst, "gather_rows: null array");
            require(n_out * (int64_t)arrays[a].row_floats < (int64_t(1) << 40), "gather_rows: array too large");
        }
        const Launch L{(cudaStream_t)stream, false};
        launch_gather_rows(arrays, n_arrays, index, n_out, L);
    });
}

int scgr_copy_segments(const ScgrSegmentCopy* segments, int32_t n_segments, scgr_stream_t stream) {
    return guarded([&] {
        require(n_segments >= 0 && n_segments <= SCGR_COPY_MAX_SEGMENTS, "copy_segments: 0..SCGR_COPY_MAX_SEGMENTS segments per call");
        if (n_segments == 0) return;
        require(segments != nullptr, "copy_segments: null segment table");
        for (int i = 0; i < n_segments; i++) {
            require(segments[i].n_floats >= 0 && segments[i].n_floats < (int64_t(1) << 40), "copy_segments: segment size out of range");
            require(segments[i].n_floats == 0 || segments[i].dst != nullptr, "copy_segments: null destination");
        }
        const Launch L{(cudaStream_t)stream, false};
        launch_copy_segments(segments, n_segments, L);
    });
}

int scgr_adam_step(const ScgrAdamGroup* groups, int32_t n_groups, double beta1, double beta2, double eps,
                   scgr_stream_t stream) {
    return guarded([&] {
        require(n_groups >= 0 && n_groups <= SCGR_ADAM_MAX_GROUPS, "adam: 0..SCGR_ADAM_MAX_GROUPS groups per call");
        if (n_groups == 0) return;
        require(groups != nullptr, "adam: null group table");
        require(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "adam: bad betas / eps");
        for (int i = 0; i < n_groups; i++) {
            const ScgrAdamGroup& G = groups[i];
            require(G.n >= 0 && G.n < (int64_t(1) << 31), "adam: group size must be in [0, 2^31)");
            if (G.n == 0) continue;
            require(G.param && G.grad && G.exp_avg && G.exp_avg_sq, "adam: null array in a group");
            require(G.step >= 1, "adam: step counts from 1 (torch increments before the update)");
        }
        const Launch L{(cudaStream_t)stream, false};
        launch_adam(groups, n_groups, beta1, beta2, eps, L);
    });
}

long long scgr_kernel_launch_count(void) { return g_kernel_launches.load(); }

int scgr_profile_enable(int on) {
    return guarded([&] {
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        for (auto& e : g_prof) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
        g_prof.clear();
        g_prof_on.store(on != 0);
    });
}

int scgr_profile_fetch(const char** names, float* ms, int max_entries) {
    int n = 0;
    int rc = guarded([&] {
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        for (auto& e : g_prof) {
            float t = 0.f;
            cudaError_t err = cudaEventSynchronize(e.stop);
            if (err == cudaSuccess) err = cudaEventElapsedTime(&t, e.start, e.stop);
            if (err != cudaSuccess) { (void)cudaGetLastError(); t = -1.f; }
            if (n < max_entries) { names[n] = e.name; ms[n] = t; n++; }
            cudaEventDestroy(e.start);
            cudaEventDestroy(e.stop);
        }
        g_prof.clear();
    });
    return rc == 0 ? n : -1;
}

int scgr_debug_views(int32_t P, int32_t W, int32_t H, int64_t capacity, const void* geometry_scratch,
                     const void* binning_scratch, const void* image_scratch, ScgrDebugViews* out) {
    return guarded([&] {
        require(out != nullptr, "null out");
        std::memset(out, 0, sizeof(*out));
        if (geometry_scratch) {
            const GeometryLayout G = carve_geometry(const_cast<void*>(geometry_scratch), P);
            out->record = reinterpret_cast<const float*>(G.rec);
            out->tiles_touched = G.tiles_touched;
            out->depth_order = G.sort_vals[0];
            out->num_rendered = G.status;
        }
        if (binning_scratch) {
            const BinningLayout B = carve_binning(const_cast<void*>(binning_scratch), W, H, capacity);
            const uint32_t n_tiles = (uint32_t)((W + TILE - 1) / TILE) * (uint32_t)((H + TILE - 1) / TILE);
            out->point_list = B.vals[tile_partition_final_buffer(n_tiles)];
            out->ranges = reinterpret_cast<const uint32_t*>(B.ranges);
        }
        if (image_scratch) {
            const ImageLayout I = carve_image(const_cast<void*>(image_scratch), W, H);
            out->n_contrib = I.n_contrib;
            out->final_T = I.final_T;
        }
    });
}

}  // extern "C"
