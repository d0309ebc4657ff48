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
#include "_build/prior_body.inc"

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
int emu_match_loss_forward(const float* depth, int H, int W, float width, float height, const ScgrMatchPair* pairs, int n_pairs,
                           float* scratch, float* out) {
    scgr::launch_match_loss_forward(depth, H, W, width, height, pairs, n_pairs, scratch, out, kHost);
    return 0;
}
int emu_match_loss_backward(const float* depth, int H, int W, float width, float height, const ScgrMatchPair* pairs, int n_pairs,
                            const float* scratch, const float* upstream, float* dL_ddepth) {
    scgr::launch_match_loss_backward(depth, H, W, width, height, pairs, n_pairs, scratch, upstream, dL_ddepth, kHost);
    return 0;
}
int emu_bg_mask(float* gt, int C, int H, int W, float threshold, int window, uint8_t* mask, float* count) {
    scgr::launch_bg_mask(gt, C, H, W, threshold, window, mask, count, kHost);
    return 0;
}
size_t emu_masked_mean_scratch_bytes(long long n) { return scgr::masked_mean_scratch_bytes((size_t)n); }
int emu_masked_mean_forward(const float* values, const uint8_t* mask, long long n, void* scratch, float* out2) {
    scgr::launch_masked_mean_forward(values, mask, (size_t)n, scratch, out2, kHost);
    return 0;
}
int emu_masked_mean_backward(const uint8_t* mask, long long n, const float* out2, const float* upstream, float* dL_dvalues) {
    scgr::launch_masked_mean_backward(mask, (size_t)n, out2, upstream, dL_dvalues, kHost);
    return 0;
}
}
