// Event-driven frame representation ("simple_rep") of lens/tools/dvstools.py:173-361 on the GPU.
//
// The reference walks a "t x y p" text file event by event: a frame stays open while
// |t - current_time| <= 1/fps; the first later event that is not a hot pixel closes it, is itself
// dropped, and its timestamp starts the next frame (dvstools.py:297-349).  Inside a frame every event
// whose pixel lies in one of the 3x3 patches around `pixels` random centroids adds accum_factor to that
// centroid's slot of a uint8 vector (dvstools.py:310-322).
//
// GPU formulation: the timestamps are sorted, so a frame is a contiguous index range.  The ranges form
// a serial chain (each start depends on the previous close), found by ONE warp with a gallop + 32-ary
// search per frame (a handful of dependent loads instead of one per event); the float64 comparison is
// the reference's own expression, so the boundaries agree bit for bit.  After that the time array is
// never read again: the slot histogram kernel streams x / y once (4 B/event) through a pixel -> slot
// lookup table, private shared-memory histograms, and one global atomic per touched slot.
#include "common.cuh"

namespace lens {

constexpr int kLutHot = -2;    // hot pixel: the event does not exist (dvstools.py:293-294)

__device__ __forceinline__ bool in_frame(double t, double c, double interval)
{
    return fabs(__dsub_rn(t, c)) <= interval;                  // dvstools.py:297
}

__device__ __forceinline__ int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

// first i in [lo, hi) for which pred(i) is false (pred is monotone true...false); whole-warp call
template <class Pred>
__device__ int64_t warp_partition_point(int64_t lo, int64_t hi, int lane, Pred pred)
{
    if (lo >= hi) return hi;
    // gallop from lo: lane l probes lo + 32 * 2^l - 1 (clipped to the last element)
    for (;;) {
        const int64_t want = lo + ((int64_t)32 << lane) - 1;
        const bool clipped = want >= hi;
        const bool ok = pred(clipped ? hi - 1 : want);
        const unsigned bad = __ballot_sync(0xffffffffu, !ok);
        const unsigned any_clipped = __ballot_sync(0xffffffffu, clipped);
        if (bad) {
            const int f = __ffs(bad) - 1;                        // probe f is false, probe f - 1 is true
            const int64_t nlo = f == 0 ? lo : lo + ((int64_t)32 << (f - 1));
            hi = min64(hi, lo + ((int64_t)32 << f));             // element hi - 1 is false from here on
            lo = nlo;
            break;
        }
        if (any_clipped) return hi;                              // the last element is true: all are
        lo += (int64_t)32 << 31;
    }
    // 32-ary refinement of [lo, hi); its last element is false
    while (hi - lo > 32) {
        const int64_t stride = (hi - lo + 31) / 32;
        const int64_t probe = min64(lo + (int64_t)(lane + 1) * stride - 1, hi - 1);
        const bool ok = pred(probe);
        const unsigned bad = __ballot_sync(0xffffffffu, !ok);   // never 0: lane 31 probes hi - 1
        const int f = __ffs(bad) - 1;
        const int64_t nlo = lo + (int64_t)f * stride;
        hi = min64(hi, lo + (int64_t)(f + 1) * stride);
        lo = nlo;
    }
    const int64_t probe = lo + lane;
    const bool ok = probe < hi ? pred(probe) : false;
    const unsigned bad = __ballot_sync(0xffffffffu, !ok);
    return lo + (__ffs(bad) - 1);
}

// One warp walks the chain of frames.  flags[0] = 1 if the timestamps are not ascending (set by
// sorted_check_kernel beforehand).  start_mode 0: start = current = t[0] (the reference's offset == 0 case,
// dvstools.py:286-290, where the very first event defines the offset even if it is a hot pixel).
__global__ void event_windows_kernel(const double *__restrict__ t, const uint16_t *__restrict__ x,
                                     const uint16_t *__restrict__ y, int64_t n, const int16_t *__restrict__ lut, int W,
                                     int H, int use_first, double start, double interval, int64_t max_windows,
                                     int64_t *__restrict__ win_begin, int64_t *__restrict__ win_end,
                                     double *__restrict__ win_t0, int64_t *__restrict__ n_windows)
{
    const int lane = threadIdx.x;
    int64_t k = 0;
    if (n > 0) {
        double c = use_first ? t[0] : start;
        const double s0 = c;
        // events before the start timestamp are skipped (dvstools.py:293)
        int64_t b = warp_partition_point(0, n, lane, [&](int64_t i) { return t[i] < s0; });
        while (k < max_windows) {
            const double cc = c;
            const int64_t e = warp_partition_point(b, n, lane, [&](int64_t i) { return in_frame(t[i], cc, interval); });
            // the closing event is the first one at or after e that is not a hot pixel
            int64_t z = e;
            while (z < n) {
                const int64_t i = z + lane;
                bool hot = false;
                if (i < n) {
                    const int xe = x[i], ye = y[i];
                    hot = xe < W && ye < H && lut[ye * W + xe] == kLutHot;
                }
                const unsigned live = __ballot_sync(0xffffffffu, i < n && !hot);
                if (live) { z += __ffs(live) - 1; break; }
                z += 32;
            }
            if (z >= n) break;                                   // the last frame is never closed, hence never saved
            if (lane == 0) { win_begin[k] = b; win_end[k] = e; win_t0[k] = cc; }
            ++k;
            c = t[z];
            b = z + 1;                                           // the closing event itself is dropped
        }
    }
    if (lane == 0) *n_windows = k;
}

__global__ void sorted_check_kernel(const double *__restrict__ t, int64_t n, int *__restrict__ unsorted)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i + 1 < n && t[i + 1] < t[i]) *unsorted = 1;
}

// grid = (n_windows, n_segments): a CTA histograms its share of one frame's events into shared memory
// and adds the touched slots to counts[window][slot].
__global__ void __launch_bounds__(256) bin_lut_kernel(const uint16_t *__restrict__ x, const uint16_t *__restrict__ y,
                                                      const int16_t *__restrict__ lut, int W, int H,
                                                      const int64_t *__restrict__ win_begin,
                                                      const int64_t *__restrict__ win_end, int n_slots,
                                                      uint32_t *__restrict__ counts)
{
    extern __shared__ uint32_t hist[];
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const int64_t w = blockIdx.x;
    const int64_t b = win_begin[w], e = win_end[w];
    const int64_t len = e - b;
    const int64_t s0 = b + len * blockIdx.y / gridDim.y, s1 = b + len * (blockIdx.y + 1) / gridDim.y;
    auto one = [&](int xe, int ye) {
        if (xe >= W || ye >= H) return;
        const int slot = lut[ye * W + xe];
        if (slot >= 0) atomicAdd(&hist[slot], 1u);
    };
    int64_t a0 = (s0 + 7) & ~(int64_t)7;
    if (a0 > s1) a0 = s1;
    const int64_t a1 = a0 + ((s1 - a0) & ~(int64_t)7);
    for (int64_t i = s0 + threadIdx.x; i < a0; i += blockDim.x) one(x[i], y[i]);
    const uint4 *x8 = reinterpret_cast<const uint4 *>(x + a0), *y8 = reinterpret_cast<const uint4 *>(y + a0);
    const int64_t n8 = (a1 - a0) >> 3;
    for (int64_t j = threadIdx.x; j < n8; j += blockDim.x) {
        const uint4 xv = __ldg(x8 + j), yv = __ldg(y8 + j);
        const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            one(xs[h] & 0xffff, ys[h] & 0xffff);
            one(xs[h] >> 16, ys[h] >> 16);
        }
    }
    for (int64_t i = a1 + threadIdx.x; i < s1; i += blockDim.x) one(x[i], y[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) {
        const uint32_t v = hist[i];
        if (v) atomicAdd(&counts[w * n_slots + i], v);
    }
}

// frame[slot] = uint8(count * weight): `frame_data[index] += accum_factor` on a uint8 vector
// (dvstools.py:322) applied `count` times wraps mod 256.
__global__ void slots_to_u8_kernel(const uint32_t *__restrict__ counts, int64_t n, uint32_t weight,
                                   uint8_t *__restrict__ frames)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) frames[i] = (uint8_t)((counts[i] * weight) & 255u);
}

}  // namespace lens

using namespace lens;

extern "C" int lens_event_windows(const double *t, const uint16_t *x, const uint16_t *y, int64_t n_events,
                                  const int16_t *lut, int sensor_w, int sensor_h, int use_first_event, double start,
                                  double interval, int64_t max_windows, int64_t *win_begin, int64_t *win_end,
                                  double *win_t0, int64_t *n_windows, int *unsorted, void *stream)
{
    LENS_CHECK_ARG(n_events >= 0 && max_windows >= 0, "lens_event_windows: negative size");
    LENS_CHECK_ARG(n_events == 0 || (t && x && y), "lens_event_windows: NULL event array");
    LENS_CHECK_ARG(lut && sensor_w > 0 && sensor_h > 0, "lens_event_windows: bad lookup table");
    LENS_CHECK_ARG(win_begin && win_end && win_t0 && n_windows, "lens_event_windows: NULL output");
    LENS_CHECK_ARG(interval >= 0.0, "lens_event_windows: interval must be >= 0");
    cudaStream_t st = as_stream(stream);
    if (unsorted) {
        LENS_CUDA(cudaMemsetAsync(unsorted, 0, sizeof(int), st));
        if (n_events > 1) {
            sorted_check_kernel<<<(unsigned)ceil_div64(n_events - 1, 256), 256, 0, st>>>(t, n_events, unsorted);
            LENS_LAUNCH_CHECK();
        }
    }
    event_windows_kernel<<<1, 32, 0, st>>>(t, x, y, n_events, lut, sensor_w, sensor_h, use_first_event, start, interval,
                                           max_windows, win_begin, win_end, win_t0, n_windows);
    LENS_LAUNCH_CHECK();
    return 0;
}

extern "C" int lens_bin_events_lut(const uint16_t *x, const uint16_t *y, const int16_t *lut, int sensor_w, int sensor_h,
                                   const int64_t *win_begin, const int64_t *win_end, int64_t n_windows, int n_slots,
                                   int weight, int64_t events_per_window_hint, uint32_t *counts, uint8_t *frames,
                                   void *stream)
{
    LENS_CHECK_ARG(n_windows >= 0 && n_windows <= 2147483647LL, "lens_bin_events_lut: bad window count");
    LENS_CHECK_ARG(n_slots > 0 && n_slots <= 8192, "lens_bin_events_lut: n_slots=%d outside [1, 8192]", n_slots);
    LENS_CHECK_ARG(weight >= 0 && weight <= 255, "lens_bin_events_lut: weight outside [0, 255]");
    LENS_CHECK_ARG(lut && sensor_w > 0 && sensor_h > 0, "lens_bin_events_lut: bad lookup table");
    if (n_windows == 0) return 0;
    LENS_CHECK_ARG(x && y && win_begin && win_end && counts && frames, "lens_bin_events_lut: NULL buffer");
    LENS_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "lens_bin_events_lut: x / y must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int64_t total = n_windows * n_slots;
    LENS_CUDA(cudaMemsetAsync(counts, 0, (size_t)total * sizeof(uint32_t), st));
    // enough CTAs per frame that each streams about 32 K events, and at least ~4 CTAs per SM in total
    int64_t segs = std::max<int64_t>(1, events_per_window_hint / 32768);
    const int64_t want = (int64_t)std::max(sm_count(), 1) * 4;
    if (n_windows * segs < want) segs = std::min<int64_t>(ceil_div64(want, n_windows), std::max<int64_t>(1, events_per_window_hint / 2048));
    segs = std::max<int64_t>(1, std::min<int64_t>(segs, 65535));
    dim3 grid((unsigned)n_windows, (unsigned)segs);
    bin_lut_kernel<<<grid, 256, (size_t)n_slots * 4, st>>>(x, y, lut, sensor_w, sensor_h, win_begin, win_end, n_slots,
                                                           counts);
    LENS_LAUNCH_CHECK();
    slots_to_u8_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(counts, total, (uint32_t)weight, frames);
    LENS_LAUNCH_CHECK();
    return 0;
}
