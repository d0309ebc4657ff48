// emu_loss_knn.cpp -- TEST INFRASTRUCTURE: scgaussian_b200/csrc/loss.cu and knn.cu compiled for the host (see
// host_cuda_shim.h); sources transformed by tests/emulation/build.py as for emu_preprocess.cpp.
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

#include "_build/loss_body.inc"
#include "_build/knn_body.inc"

static const scgr::Launch kHost{nullptr, false};

extern "C" {
size_t emu_photometric_scratch_bytes(int C, int H, int W) { return scgr::photometric_scratch_bytes(C, H, W); }
int emu_photometric_forward(const float* img, const float* gt, int C, int H, int W, float lambda, void* scratch, int want_grad,
                            float* out3) {
    scgr::launch_photometric_forward(img, gt, C, H, W, lambda, scratch, want_grad != 0, out3, kHost);
    return 0;
}
int emu_photometric_backward(const float* img, const float* gt, int C, int H, int W, float lambda, const void* scratch,
                             const float* upstream, float* dL_dimg) {
    scgr::launch_photometric_backward(img, gt, C, H, W, lambda, scratch, upstream, dL_dimg, kHost);
    return 0;
}
int emu_knn3(const float* points, int32_t n, float* out) { scgr::launch_knn3(points, n, out, kHost); return 0; }
}
