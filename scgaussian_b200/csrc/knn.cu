// knn.cu -- mean squared distance to the 3 nearest neighbours (sm_100a).  SURVEY.md section 8(f), row f4.
//
// Replaces `simple_knn._C.distCUDA2` (external package), which the reference
// calls once, at initialisation, on the <= ~10^4-10^5 points of the initial cloud
// (reference scene/gaussian_model.py:20 import, :444 `dist2 = clamp_min(distCUDA2(points), 1e-7)`)
// to size the initial Gaussians: out[i] = mean over the 3 nearest OTHER points j of |p_i - p_j|^2
// (coincident points count with distance 0).  The external kernel prunes with Morton-ordered boxes; at these sizes an
// exact tiled all-pairs scan is simpler and takes well under a millisecond (10^4 points = 10^8 pair
// evaluations): one thread per query, the candidate points staged 256 at a time in shared memory.
// Not on the per-step hot path.
#include <cfloat>

#include "common.cuh"

namespace scgr {

namespace {

constexpr int KNN_THREADS = 256;

__global__ void __launch_bounds__(KNN_THREADS)
knn3_kernel(const float* __restrict__ pts, const int n, float* __restrict__ out) {
    __shared__ float4 s_p[KNN_THREADS];
    const int i = blockIdx.x * KNN_THREADS + threadIdx.x;
    float3 q = make_float3(0.f, 0.f, 0.f);
    if (i < n) q = make_float3(__ldg(pts + 3 * (size_t)i), __ldg(pts + 3 * (size_t)i + 1), __ldg(pts + 3 * (size_t)i + 2));
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;      // ascending: the three smallest squared distances so far
    for (int base = 0; base < n; base += KNN_THREADS) {
        const int j = base + threadIdx.x;
        __syncthreads();
        if (j < n)
            s_p[threadIdx.x] = make_float4(__ldg(pts + 3 * (size_t)j), __ldg(pts + 3 * (size_t)j + 1), __ldg(pts + 3 * (size_t)j + 2), 0.f);
        __syncthreads();
        const int cnt = min(KNN_THREADS, n - base);
        for (int k = 0; k < cnt; k++) {
            const float4 p = s_p[k];
            const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
            const float d = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            if (base + k != i && d < b2) {
                if (d < b1) {
                    b2 = b1;
                    if (d < b0) { b1 = b0; b0 = d; } else { b1 = d; }
                } else {
                    b2 = d;
                }
            }
        }
    }
    if (i < n) {
        // the external kernel starts its three best distances at FLT_MAX and averages all three; with fewer than
        // 4 points that would be ~1e38 -- here the neighbours that do not exist are simply left out
        float sum = 0.f;
        int m = 0;
        if (b0 < FLT_MAX) { sum += b0; m++; }
        if (b1 < FLT_MAX) { sum += b1; m++; }
        if (b2 < FLT_MAX) { sum += b2; m++; }
        out[i] = m == 3 ? sum / 3.0f : (m > 0 ? sum / (float)m : 0.f);
    }
}

}  // namespace

void launch_knn3(const float* points, int32_t n, float* out, const Launch& L) {
    if (n <= 0) return;
    begin_kernel("knn3_mean_dist2", L);
    knn3_kernel<<<(n + KNN_THREADS - 1) / KNN_THREADS, KNN_THREADS, 0, L.stream>>>(points, n, out);
    check_launch("knn3_mean_dist2", L);
}

}  // namespace scgr
