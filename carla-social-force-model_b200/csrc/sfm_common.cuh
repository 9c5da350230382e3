// Shared definitions of the sfm_b200 CUDA translation unit (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sfm_b200.h"

namespace sfm {

// ---- staged float32 planes read by the all-pairs kernel -------------------------------------------------------------
// Gather buffer layout: [world][NPLANES][rows_pad] float32.  One rank block is what a rank contributes to the per-step
// all-gather (44 B per pedestrian row).
//
// Positions are staged as float32 PAIRS (double-single): x - origin = hi + lo with hi on the 2^-6 m lattice (exact in
// float32 while |x - origin| < 2^18 m) and |lo| <= 2^-7 m.  The pair kernel forms d = (hi_j - hi_i) + (lo_j - lo_i): the
// first difference is exact (both operands are lattice points), the second is exact to 2^-31 m, so d carries one float32
// rounding RELATIVE TO |d| -- independent of how far the crowd is from the origin.  (A single float32 per coordinate
// rounds to 1.5e-5 m at |x| >= 256 m, which alone breaks the 1e-4 / 1e-5 force tolerance on integrated states.)
constexpr int NPLANES = 11;
enum Plane { PX = 0, PY = 1, PZ = 2, PXL = 3, PYL = 4, PZL = 5, PR = 6, PVX = 7, PVY = 8, PVZ = 9, PFLAG = 10 };
constexpr int ROW_ALIGN = 256;              // rows_pad granularity == j-tile length of the pair kernel
constexpr float PAD_POS = 1.0e15f;          // padded rows sit this far away: exp(-dist/B) underflows to exactly 0
constexpr double POS_LATTICE = 64.0;        // hi parts are multiples of 1 / POS_LATTICE metres

// Parameters of the float32 Moussaid evaluation, pre-folded on the host (see k1_ped_pairs.cuh for the algebra).
struct PairParams {
    float lambda;          // lambda_weight (applied when staging velocities)
    float eps_gamma;       // epsilon * gamma
    float neg_l2e_over_gamma;   // -log2(e) / gamma
    float c_nprime;        // (n' * gamma)^2 * log2(e)
    float c_n;             // (n  * gamma)^2 * log2(e)
    float log2A;           // log2(A)
};

// Parameters of the float64 Moussaid evaluation used by the obstacle kernels (reference operation order).
struct MoussaidD {
    double lambda, A, gamma, n, n_prime, epsilon, threshold;
};

// Uniform grid over the centres of a point set (sections or obstacles); cell edge >= the largest cutoff of the set.
struct CellGrid {
    double x0, y0, cell, inv_cell;
    int nx, ny;
};

// A CSR point set on the device.
struct SegmentSet {
    int64_t count = 0;          // sections / obstacles
    int64_t n_points = 0;
    double2* center = nullptr;  // [count]
    double* cutoff = nullptr;   // [count]  per-section length, or the perception threshold replicated
    double2* velocity = nullptr;// [count]  obstacle velocity (zeros for borders / static)
    int* offset = nullptr;      // [count + 1]
    double2* point = nullptr;   // [n_points]
    float* tol = nullptr;       // [count]  float32 bracket width of the two-stage nearest-point search (k2_cells.cuh)
    int* chunk_first = nullptr; // [count]  first pruning chunk of the item in `chunk`, -1 = none (null: no tables)
    float4* chunk = nullptr;    // two float4 per chunk of 16 points: chord + deviation (k2_cells.cuh)
    float4* chord0 = nullptr;   // two float4 per item: whole-item chord + uniform-spacing deviation (k2_cells.cuh, direct path)
    CellGrid grid{};
    int* cell_start = nullptr;  // [nx * ny + 1]
    int* cell_item = nullptr;   // [count]  set items ordered by (cell, index)
    int64_t cap_count = 0, cap_points = 0, cap_cells = 0;
};

}  // namespace sfm
