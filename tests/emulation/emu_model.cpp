// emu_model.cpp -- TEST INFRASTRUCTURE: scgaussian_b200/csrc/model.cu compiled for the host (see host_cuda_shim.h).
// `model_body.inc` is model.cu with its launches rewritten by tests/emulation/build.py; nothing else is changed.
#include "host_cuda_shim.h"

#include "../../include/scgr.h"

namespace scgr {
struct Launch {
    cudaStream_t stream;
    bool debug;
};
static void begin_kernel(const char*, const Launch&) {}
static void check_launch(const char*, const Launch&) {}
}  // namespace scgr

#include "_build/model_body.inc"

static const scgr::Launch kHost{nullptr, false};

extern "C" {
int emu_assemble_forward(const ScgrModel* m, const ScgrActivated* o) { scgr::launch_assemble_forward(*m, *o, kHost); return 0; }
int emu_assemble_backward(const ScgrModel* m, const ScgrActivatedGrads* g, const ScgrModelGrads* o) {
    scgr::launch_assemble_backward(*m, *g, *o, kHost);
    return 0;
}
int emu_adam_step(const ScgrAdamGroup* g, int n, double b1, double b2, double e) { scgr::launch_adam(g, n, b1, b2, e, kHost); return 0; }
int emu_densification_stats(const float* g, const uint8_t* f, const int32_t* r, int P, float* a, float* d, float* m) {
    scgr::launch_densification_stats(g, f, r, P, a, d, m, kHost);
    return 0;
}
int emu_gather_rows(const ScgrRowGather* a, int n, const int64_t* idx, int64_t n_out) { scgr::launch_gather_rows(a, n, idx, n_out, kHost); return 0; }
int emu_copy_segments(const ScgrSegmentCopy* s, int n) { scgr::launch_copy_segments(s, n, kHost); return 0; }
}
