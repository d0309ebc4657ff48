// emu_preprocess.cpp -- TEST INFRASTRUCTURE: scgaussian_b200/csrc/preprocess.cu (+ the device helpers of common.cuh)
// compiled for the host (see host_cuda_shim.h).  `_build/common_host.cuh` and `_build/preprocess_body.inc` are the two
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
int emu_mark_visible(const float* means3D, int32_t P, const float* viewmatrix, uint8_t* present) {
    scgr::launch_mark_visible(means3D, P, viewmatrix, present, kHost);
    return 0;
}
}
