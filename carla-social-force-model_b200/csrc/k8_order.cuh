// K8 -- staged slot order: which slot of the rank's staged block (sfm_common.cuh) a pedestrian row occupies.
//
// The pair kernel reads a 256-slot tile through the origins of its four 64-slot runs whenever those runs are spatially
// compact (k1_sym.cuh, "local" path: one subtraction per coordinate and pair instead of three).  Row order belongs to the
// caller -- PedState.state is in spawn order (pedestrian_state.py:26-43) and every result is per row -- so compactness is
// arranged one level below: rows are STAGED in the order of a Hilbert curve over their xy positions.  K3 runs in slot
// order (thread s stages row row_of_slot[s]), k1_sym_finish / k1_sym_repair read a row's accumulators at
// slot_of_row[row]; nothing above the staged planes ever sees the permutation.  The order is rebuilt from the current
// positions every `reorder_interval` ticks (pedestrians walk ~0.07 m per tick: an order stays good for tens of ticks) --
// three small kernels and a stable radix sort (k2_cells.cuh), deterministic: ties keep row order.
//
// A Hilbert curve rather than a Z-order one: consecutive cells of a Hilbert curve are neighbours in the plane, while a
// Z-order run that crosses a quadrant boundary spans half the domain.
#pragma once

#include <limits.h>

#include "sfm_common.cuh"

namespace sfm {

constexpr int ORDER_BITS = 15;          // Hilbert cells per axis = 2^15 over the crowd's bounding square (30-bit keys)

// float -> int, monotone (so atomicMin / atomicMax on ints order floats)
__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k8_bbox_init(int* box) {
    if (threadIdx.x < 2) box[threadIdx.x] = INT_MAX;
    else if (threadIdx.x < 4) box[threadIdx.x] = INT_MIN;
}

// box = (min x, min y, max x, max y) of the float32-rounded xy positions, in the ordered-int encoding
__global__ void __launch_bounds__(256) k8_bbox(const double4* __restrict__ locr, int n, int* box) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int lo_x = INT_MAX, lo_y = INT_MAX, hi_x = INT_MIN, hi_y = INT_MIN;
    if (i < n) {
        const double4 L = locr[i];
        const float x = (float)L.x, y = (float)L.y;
        if (x == x && y == y) {                     // NaN rows do not shape the box
            lo_x = hi_x = float_to_ordered(x);
            lo_y = hi_y = float_to_ordered(y);
        }
    }
    lo_x = __reduce_min_sync(0xffffffffu, lo_x); lo_y = __reduce_min_sync(0xffffffffu, lo_y);
    hi_x = __reduce_max_sync(0xffffffffu, hi_x); hi_y = __reduce_max_sync(0xffffffffu, hi_y);
    if ((threadIdx.x & 31) == 0) {
        if (lo_x != INT_MAX) { atomicMin(box + 0, lo_x); atomicMin(box + 1, lo_y); }
        if (hi_x != INT_MIN) { atomicMax(box + 2, hi_x); atomicMax(box + 3, hi_y); }
    }
}

// Hilbert index of cell (x, y) on a 2^bits x 2^bits grid
__device__ __forceinline__ unsigned hilbert_index(unsigned x, unsigned y, int bits) {
    const unsigned n = 1u << bits;
    unsigned d = 0;
    for (unsigned s = n >> 1; s > 0; s >>= 1) {
        const unsigned rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (!ry) {
            if (rx) { x = n - 1 - x; y = n - 1 - y; }
            const unsigned t = x; x = y; y = t;
        }
    }
    return d;
}

__global__ void __launch_bounds__(256) k8_keys(const double4* __restrict__ locr, int n, const int* __restrict__ box,
                                               unsigned* __restrict__ key, int* __restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x0 = ordered_to_float(box[0]), y0 = ordered_to_float(box[1]);
    const float span = fmaxf(ordered_to_float(box[2]) - x0, ordered_to_float(box[3]) - y0);
    const float cells = (float)(1 << ORDER_BITS);
    const float scale = (span > 0.0f && span < 3.0e38f) ? cells / span : 0.0f;
    const double4 L = locr[i];
    const float fx = ((float)L.x - x0) * scale, fy = ((float)L.y - y0) * scale;
    const unsigned qx = (unsigned)fminf(fmaxf(fx, 0.0f), cells - 1.0f);        // NaN -> 0
    const unsigned qy = (unsigned)fminf(fmaxf(fy, 0.0f), cells - 1.0f);
    key[i] = hilbert_index(qx, qy, ORDER_BITS);
    val[i] = i;
}

// sorted rows -> the two maps; slots beyond the live rows are pad slots (row -1)
__global__ void __launch_bounds__(256) k8_fill_order(const int* __restrict__ sorted_row, int n, int rows_pad,
                                                     int* __restrict__ slot_of_row, int* __restrict__ row_of_slot) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= rows_pad) return;
    if (s < n) {
        const int r = sorted_row[s];
        row_of_slot[s] = r;
        slot_of_row[r] = s;
    } else {
        row_of_slot[s] = -1;
    }
}

// a caller-supplied order (sfm_set_slot_order; validated on the host): slot_of_row -> row_of_slot
__global__ void __launch_bounds__(256) k8_invert_order(const int* __restrict__ slot_of_row, int n, int rows_pad,
                                                       int* __restrict__ row_of_slot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_pad) return;
    if (i < n) row_of_slot[slot_of_row[i]] = i;
    else row_of_slot[i] = -1;          // slots n .. rows_pad - 1 are never the image of a row (the order permutes [0, n))
}

}  // namespace sfm
