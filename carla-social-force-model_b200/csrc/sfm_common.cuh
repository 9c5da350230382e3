// Shared definitions of the sfm_b200 CUDA translation unit (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sfm_b200.h"

namespace sfm {

// ---- staged float32 planes read by the all-pairs kernel -------------------------------------------------------------
// Gather buffer layout: [world][NPLANES][rows_pad] float32.  One rank block is what a rank contributes to the per-step
// all-gather (60 B per pedestrian row).
//
// Positions are staged as float32 PAIRS (double-single): x - origin = hi + lo with hi on the 2^-6 m lattice (exact in
// float32 while |x - origin| < 2^18 m) and |lo| <= 2^-7 m.  The pair kernel forms d = (hi_j - hi_i) + (lo_j - lo_i): the
// first difference is exact (both operands are lattice points), the second is exact to 2^-31 m, so d carries one float32
// rounding RELATIVE TO |d| -- independent of how far the crowd is from the origin.  (A single float32 per coordinate
// rounds to 1.5e-5 m at |x| >= 256 m, which alone breaks the 1e-4 / 1e-5 force tolerance on integrated states.)
//
// Run-local copies (the pair kernel's "local" path).  Every run of SUB_ROWS = 64 consecutive slots (a quarter of a
// 256-slot tile) also carries its positions as ONE float32 per coordinate relative to the run's own origin c -- the centre
// of its bounding box, on the same lattice: xr = float32(x - origin - c).  The kernel can then form d = xr_j - m_i with
// m_i = (hi_i - c_J) + lo_i computed once per (row, run): ONE subtraction per coordinate and pair instead of three.
// Rounding: xr_j to 2^-25 |xr_j|, m_i to 2^-25 |x_i - c_J| (hi_i - c_J is exact: lattice points), the difference to
// 2^-25 |d| -- an ABSOLUTE error of about 2^-24 ext for a run of half-extent ext (5e-7 m at 8 m), where the double-single
// form is exact relative to |d|.  That is harmless in the far field and not good enough for the pairs that carry the force
// (the model amplifies a relative error of d by 2 (n' B)^2 theta ~ 20-50), so the choice is made per TILE pair, by
// geometry alone: only when the xy bounding boxes of the two 256-slot tiles are at least
//     max(LOCAL_SEP, LOCAL_SEP_FACTOR * ext)      (ext = largest half-extent of the partner tile's four runs)
// apart.  Every pair closer than that is evaluated in double-single form (or by the guarded diagonal code); on the local
// path |d| >= ext / 4, so the relative error of d stays below 2^-22 = 2.4e-7 whatever the run looks like -- at most a tenth
// of the force tolerance after the model's amplification, and the far field carries little force to begin with
// (profiles/local_origin_emulation.py: staging error alone, both paths, against the force tolerance).  What the path
// saves is the work on the 93-99 % of the tile pairs that are not neighbours -- provided consecutive slots are close in
// space, which is what the staged slot order (k8_order.cuh) arranges; rows staged in arbitrary order have runs as wide as
// the crowd and simply never qualify: the order is a matter of speed, never of accuracy.
// The plane PMETA carries, in the first 16 slots of every tile, (c_x, c_y, c_z, ext) of its four runs (ext = +inf for a
// run that does not qualify: wider than LOCAL_LIMIT, or too far from the staging origin for hi - c to be exact), and in
// slots META_BOX .. META_BOX + 3 the tile's xy bounding box (min x, min y, max x, max y, origin-relative).
constexpr int NPLANES = 15;
enum Plane { PX = 0, PY = 1, PZ = 2, PXL = 3, PYL = 4, PZL = 5, PR = 6, PVX = 7, PVY = 8, PVZ = 9, PFLAG = 10,
             PXR = 11, PYR = 12, PZR = 13, PMETA = 14 };
constexpr int ROW_ALIGN = 256;              // rows_pad granularity == j-tile length of the pair kernel
constexpr float PAD_POS = 1.0e15f;          // padded rows sit this far away: exp(-dist/B) underflows to exactly 0
constexpr double POS_LATTICE = 64.0;        // hi parts are multiples of 1 / POS_LATTICE metres
constexpr int SUB_ROWS = 64;                // rows per sub-tile run (two warps of the staging CTA)
constexpr int SUBS_PER_TILE = ROW_ALIGN / SUB_ROWS;
constexpr double LOCAL_LIMIT = 64.0;        // runs wider than this (half-extent, metres) never take the local path
constexpr float LOCAL_SEP = 1.0f;           // tile pairs whose bounding boxes are closer than this stay double-single ...
constexpr float LOCAL_SEP_FACTOR = 0.25f;   // ... or closer than this times the partner tile's largest run half-extent
constexpr int META_BOX = 16;                // first slot of the tile's bounding box in its PMETA plane
constexpr double LOCAL_RANGE = 131072.0;    // run origins beyond 2^17 m from the staging origin: hi - c no longer exact

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
