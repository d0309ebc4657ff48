// emu_preprocess.cpp -- TEST INFRASTRUCTURE: scgaussian_b200/csrc/preprocess.cu, binning.cu and render.cu (+ the device
// helpers of common.cuh) compiled for the host (see host_cuda_shim.h).  `_build/common_host.cuh` and `_build/preprocess_body.inc` are the two
// sources with the include of <cuda_runtime.h> dropped, the one inline-PTX statement (sqrt.approx) replaced by sqrtf
// and the launches rewritten by tests/emulation/build.py; nothing else is changed.
#define __CUDACC__ 1
#include "host_cuda_shim.h"

#include <stdexcept>
#include <string>

#include "_build/common_host.cuh"

namespace scgr {
void begin_kernel(const char*, const Launch&) {}
void check_launch(const char*, const Launch&) {}
void check_stage(const char*, const Launch&) {}
}  // namespace scgr

#include "_build/preprocess_body.inc"
#include "_build/binning_body.inc"
#include "_build/render_body.inc"

static const scgr::Launch kHost{nullptr, false};

extern "C" {
size_t emu_geometry_bytes(int32_t P) { return scgr::carve_geometry(nullptr, P).bytes; }
// byte offsets of the arrays the tests read: rec, tiles_touched, rect, tile_mask, depth_key, sort_vals[0], screen_grad, sweep
void emu_geometry_offsets(int32_t P, size_t* off) {
    const scgr::GeometryLayout G = scgr::carve_geometry(nullptr, P);
    off[0] = (size_t)G.rec; off[1] = (size_t)G.tiles_touched; off[2] = (size_t)G.rect; off[3] = (size_t)G.tile_mask;
    off[4] = (size_t)G.depth_key; off[5] = (size_t)G.sort_vals[0]; off[6] = (size_t)G.screen_grad; off[7] = (size_t)G.sweep;
}
int emu_preprocess_forward(const ScgrView* v, const ScgrGaussians* g, void* geometry, int32_t* radii) {
    scgr::launch_preprocess_forward(*v, *g, scgr::carve_geometry(geometry, g->P), radii, kHost);
    return 0;
}
int emu_depth_keys(const ScgrView* v, const ScgrGaussians* g, void* geometry) {
    scgr::launch_depth_keys(*v, *g, scgr::carve_geometry(geometry, g->P), kHost);
    return 0;
}
int emu_preprocess_backward(const ScgrView* v, const ScgrGaussians* g, void* geometry, const ScgrGrads* out) {
    scgr::launch_preprocess_backward(*v, *g, scgr::carve_geometry(geometry, g->P), *out, kHost);
    return 0;
}
size_t emu_binning_bytes(int32_t W, int32_t H, int64_t capacity) { return scgr::carve_binning(nullptr, W, H, capacity).bytes; }
// byte offsets into the binning scratch: ranges, the final (tile id, Gaussian id) buffers
void emu_binning_offsets(int32_t W, int32_t H, int64_t capacity, size_t* off) {
    const scgr::BinningLayout B = scgr::carve_binning(nullptr, W, H, capacity);
    const uint32_t n_tiles = (uint32_t)((W + 15) / 16) * (uint32_t)((H + 15) / 16);
    const int fin = scgr::tile_partition_final_buffer(n_tiles);
    off[0] = (size_t)B.ranges; off[1] = (size_t)B.vals[fin]; off[2] = (size_t)B.keys[0]; off[3] = (size_t)B.vals[0];
}
size_t emu_offsets_offset(int32_t P) { return (size_t)scgr::carve_geometry(nullptr, P).offsets; }
size_t emu_status_offset(int32_t P) { return (size_t)scgr::carve_geometry(nullptr, P).status; }
// depth sort (4 radix passes) + scan: everything of stage 1 that follows the preprocess kernel
int emu_depth_sort_and_scan(const ScgrView* v, const ScgrGaussians* g, void* geometry) {
    const scgr::GeometryLayout G = scgr::carve_geometry(geometry, g->P);
    scgr::launch_depth_sort(*v, *g, G, kHost);
    scgr::launch_scan_offsets(G, g->P, nullptr, kHost);
    return 0;
}
// stage 2 up to the per-tile lists: empty ranges, emission, tile partition (+ ranges in its last pass)
int emu_emit_and_partition(const ScgrView* v, const ScgrGaussians* g, void* geometry, void* binning, int64_t capacity) {
    const scgr::GeometryLayout G = scgr::carve_geometry(geometry, g->P);
    const scgr::BinningLayout B = scgr::carve_binning(binning, v->image_width, v->image_height, capacity);
    scgr::launch_binning_prologue(*v, B, g->P, capacity, kHost);
    int fin = -1;
    scgr::launch_emit_and_partition(*v, G, B, g->P, capacity, &fin, kHost);
    return fin;
}
size_t emu_image_bytes(int32_t W, int32_t H) { return scgr::carve_image(nullptr, W, H).bytes; }
void emu_image_offsets(int32_t W, int32_t H, size_t* off) {
    const scgr::ImageLayout I = scgr::carve_image(nullptr, W, H);
    off[0] = (size_t)I.n_contrib; off[1] = (size_t)I.final_T;
}
static const uint32_t* final_list(const ScgrView* v, const scgr::BinningLayout& B) {
    const uint32_t n_tiles = (uint32_t)((v->image_width + 15) / 16) * (uint32_t)((v->image_height + 15) / 16);
    return B.vals[scgr::tile_partition_final_buffer(n_tiles)];
}
int emu_render_forward(const ScgrView* v, const ScgrGaussians* g, void* geometry, void* binning, int64_t capacity, void* image,
                       float* color, float* depth, float* alpha) {
    const scgr::GeometryLayout G = scgr::carve_geometry(geometry, g->P);
    const scgr::BinningLayout B = scgr::carve_binning(binning, v->image_width, v->image_height, capacity);
    const scgr::ImageLayout I = scgr::carve_image(image, v->image_width, v->image_height);
    scgr::launch_render_forward(*v, G, B, final_list(v, B), capacity, I, color, depth, alpha, kHost);
    return 0;
}
// backward prologue (tile order + zeroed accumulators) + render backward: fills the per-Gaussian ScreenGrad sums
int emu_render_backward(const ScgrView* v, const ScgrGaussians* g, void* geometry, void* binning, int64_t capacity, void* image,
                        const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha) {
    const scgr::GeometryLayout G = scgr::carve_geometry(geometry, g->P);
    const scgr::BinningLayout B = scgr::carve_binning(binning, v->image_width, v->image_height, capacity);
    const scgr::ImageLayout I = scgr::carve_image(image, v->image_width, v->image_height);
    scgr::launch_render_backward(*v, G, B, final_list(v, B), capacity, I, dL_dcolor, dL_ddepth, dL_dalpha, g->P, kHost);
    return 0;
}
int emu_mark_visible(const float* means3D, int32_t P, const float* viewmatrix, uint8_t* present) {
    scgr::launch_mark_visible(means3D, P, viewmatrix, present, kHost);
    return 0;
}
}
