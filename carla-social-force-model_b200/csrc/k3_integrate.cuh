// K3 -- fused acceleration force + force sum + Euler step + speed clamp + restaging, float64, HBM-bound.
//
// Replaces AccelerationForce._get_force (reference forces.py:46-53 with stateutils.desired_directions :7-15), the
// dict-order force sum of PedestrianSimulation.tick (pedestrian_simulation.py:81), calculate_new_velocities (:117-124)
// with stateutils.cap_velocity (stateutils.py:18-23) and PedState.max_speed (pedestrian_state.py:72-73), and -- when
// integrate_positions is set -- the position update the reference leaves to the CARLA server (run_simulation.py:77-87),
// defined here as semi-implicit Euler x+ = x + dt v+.
//
// The master state stays float64 (the reference dtype, pedestrian_state.py:17): products and sums use the unfused
// __dmul_rn/__dadd_rn forms in numpy's operation order, so with identical force inputs the new velocities are
// bit-identical to numpy's.  The same pass re-emits the float32 staging planes the pair kernel reads next step.
#pragma once

#include <math_constants.h>

#include "k4_lifecycle.cuh"
#include "sfm_common.cuh"

namespace sfm {

struct StepArgs {
    // master state of the local rows
    double4* locr;               // (x, y, z, radius)
    double4* vels;               // (vx, vy, vz, target_speed)
    const double2* wp;           // next waypoint xy
    const uint8_t* mode;
    int64_t n;                   // live local rows
    int64_t rows_pad;            // staged rows of one rank block
    const int* row_of_slot;      // staged slot -> row (-1: pad slot), k8_order.cuh; nullptr: slot s holds row s
    // force inputs
    const double* ped_force;     // [n][3] reduced pair force (k1_reduce_fixup)
    const double2* f_border;     // [n] or nullptr
    const double2* f_static;
    const double2* f_dynamic;
    // outputs
    double* f_total;             // [n][3]
    double* f_accel;             // [n][3] or nullptr (kept only when class forces are requested)
    float* planes_own;           // this rank's block of the gather buffer: [NPLANES][rows_pad]
    float* planes_peer[7];       // the same block inside the other ranks' gather buffers (peer memory, K7) ...
    int n_peer;                  // ... when the all-gather is fused into this kernel; 0 otherwise
    // parameters
    double dt, tau, max_speed_factor, lambda_ped;
    double ox, oy, oz;
    int enable_accel, enable_ped;
    int update_velocity;         // 0: forces only (Force.get_force path)
    int integrate_positions;
    // fused arrival test + waypoint hand-over (K4b), evaluated at the position the forces were computed at -- the
    // reference checks arrivals right after the tick, before the simulator moves anybody (run_simulation.py:114-118)
    int advance_routes;
    Routes routes;
    ModeMachines mm;
    double2* wp_rw;
    uint8_t* mode_rw;
    double sim_time;
};

__device__ __forceinline__ void store_row(float* planes, int64_t rows_pad, int64_t i, const float (&v)[NPLANES]) {
#pragma unroll
    for (int p = 0; p < NPLANES; ++p) planes[p * rows_pad + i] = v[p];
}

// one staged row into this rank's block of the gather buffer and -- fused all-gather -- into every peer's copy of it
__device__ __forceinline__ void store_row_everywhere(const StepArgs& a, int64_t i, const float (&v)[NPLANES]) {
    store_row(a.planes_own, a.rows_pad, i, v);
    for (int r = 0; r < a.n_peer; ++r) store_row(a.planes_peer[r], a.rows_pad, i, v);
}

// x (float64, origin-relative) -> lattice point hi (multiple of 2^-6 m) + remainder lo, both float32.  lo is taken against
// the float32 value actually stored, so hi + lo represents x to 2^-31 m even where hi itself had to round (|x| >= 2^18 m).
__device__ __forceinline__ void split_hi_lo(double x, float& hi, float& lo) {
    hi = (float)(rint(x * POS_LATTICE) * (1.0 / POS_LATTICE));
    lo = (float)(x - (double)hi);
}

// Bounds of one run of SUB_ROWS slots, from the per-warp bounding boxes `bb[warp][min xyz, max xyz]`: the run's origin c
// (centre of the box on the position lattice; 0 for a run without live rows) and its half-extent -- the largest
// |x - origin - c| over its live rows and coordinates -- or +inf when the run cannot take the local path: wider than
// LOCAL_LIMIT, or so far from the staging origin that hi - c is no longer exact (sfm_common.cuh).
__device__ __forceinline__ float run_origin(const double (*bb)[6], int run, double (&c)[3]) {
    double ext = 0.0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double mn = fmin(bb[2 * run][k], bb[2 * run + 1][k]);
        const double mx = fmax(bb[2 * run][3 + k], bb[2 * run + 1][3 + k]);
        if (!(mn <= mx)) {                      // no live row in this run
            c[k] = 0.0;
            continue;
        }
        c[k] = rint(0.5 * (mn + mx) * POS_LATTICE) * (1.0 / POS_LATTICE);
        ext = fmax(ext, fmax(mx - c[k], c[k] - mn));
        ok = ok && (fabs(c[k]) <= LOCAL_RANGE);
    }
    return (ok && ext <= LOCAL_LIMIT) ? (float)ext : CUDART_INF_F;      // (a NaN extent fails the comparison: +inf)
}

// Stages one 256-slot tile: EVERY thread of the CTA (thread t <-> slot 256 * blockIdx.x + t) calls this, live or not.
// Besides the (hi, lo) parts each row gets its position relative to the origin of its 64-row run (the pair kernel's
// "local" path, sfm_common.cuh); the run origins and half-extents go to the first 16 slots of the tile's PMETA plane,
// the tile's xy bounding box to the next four.
__device__ __forceinline__ void stage_tile(const StepArgs& a, int64_t i, bool live, double x, double y, double z, double r,
                                           double vx, double vy, double vz) {
    __shared__ double bb[256 / 32][6];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double rel[3] = {x - a.ox, y - a.oy, z - a.oz};
    double lo[3], hi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = live ? rel[k] : CUDART_INF;
        hi[k] = live ? rel[k] : -CUDART_INF;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            bb[wid][k] = lo[k];
            bb[wid][3 + k] = hi[k];
        }
    }
    __syncthreads();
    float v[NPLANES];
#pragma unroll
    for (int p = 0; p < NPLANES; ++p) v[p] = 0.0f;
    if (live) {
        double c[3];
        run_origin(bb, wid >> 1, c);
        split_hi_lo(rel[0], v[PX], v[PXL]);
        split_hi_lo(rel[1], v[PY], v[PYL]);
        split_hi_lo(rel[2], v[PZ], v[PZL]);
        v[PXR] = (float)(rel[0] - c[0]);
        v[PYR] = (float)(rel[1] - c[1]);
        v[PZR] = (float)(rel[2] - c[2]);
        v[PR] = (float)r;
        v[PVX] = (float)(a.lambda_ped * vx);
        v[PVY] = (float)(a.lambda_ped * vy);
        v[PVZ] = (float)(a.lambda_ped * vz);
        // non-planar flag read by the symmetric pair kernel (its z-free fast path needs z == origin and v_z == 0)
        v[PFLAG] = (v[PZ] != 0.0f || v[PZL] != 0.0f || v[PVZ] != 0.0f) ? 1.0f : 0.0f;
    } else {
        v[PX] = v[PY] = v[PXR] = v[PYR] = PAD_POS;
    }
    if (threadIdx.x < 4 * SUBS_PER_TILE) {
        double c[3];
        const float ext = run_origin(bb, threadIdx.x >> 2, c);
        const int k = threadIdx.x & 3;
        v[PMETA] = (k == 0) ? (float)c[0] : (k == 1) ? (float)c[1] : (k == 2) ? (float)c[2] : ext;
    } else if (threadIdx.x < META_BOX + 4) {
        // the tile's xy bounding box (min x, min y, max x, max y; +inf / -inf for a tile without live rows)
        const int k = threadIdx.x - META_BOX, col = (k & 1) + ((k >> 1) ? 3 : 0);
        double e = bb[0][col];
#pragma unroll
        for (int wv = 1; wv < 256 / 32; ++wv) e = (k >> 1) ? fmax(e, bb[wv][col]) : fmin(e, bb[wv][col]);
        v[PMETA] = (float)e;
    }
    store_row_everywhere(a, i, v);
}

// The CTAs of K3 walk the staged block in SLOT order: thread s handles the row staged at slot s (-1: a pad slot).
__device__ __forceinline__ int64_t row_of(const StepArgs& a, int64_t s) {
    if (a.row_of_slot) return a.row_of_slot[s];
    return s < a.n ? s : -1;
}

// Staging only: master state -> float32 planes (after an upload or a kinematics refresh).
__global__ void __launch_bounds__(256) k3_stage(StepArgs a) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // rows_pad is a multiple of the CTA size
    const int64_t i = row_of(a, s);
    const bool live = i >= 0;
    const double4 L = live ? a.locr[i] : make_double4(0.0, 0.0, 0.0, 0.0);
    const double4 V = live ? a.vels[i] : make_double4(0.0, 0.0, 0.0, 0.0);
    stage_tile(a, s, live, L.x, L.y, L.z, L.w, V.x, V.y, V.z);
}

__global__ void __launch_bounds__(256) k3_integrate(StepArgs a) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = row_of(a, s);
    const bool live = i >= 0;                  // pad slots of the last tile only take part in the restaging (stage_tile)
    const double4 L = live ? a.locr[i] : make_double4(0.0, 0.0, 0.0, 0.0);
    const double4 V = live ? a.vels[i] : make_double4(0.0, 0.0, 0.0, 0.0);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    if (live && a.enable_accel) {
        const double2 w = a.wp[i];
        const double ex = __dsub_rn(w.x, L.x), ey = __dsub_rn(w.y, L.y);
        const double nrm = __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
        const double dv = (nrm == 0.0) ? 1.0 : nrm;                        // stateutils.py:88-90
        const double dirx = __ddiv_rn(ex, dv), diry = __ddiv_rn(ey, dv);
        const double inv_tau = __ddiv_rn(1.0, a.tau);                      // forces.py:51  1.0 / tau * (...)
        const double ax = __dmul_rn(inv_tau, __dsub_rn(__dmul_rn(V.w, dirx), V.x));
        const double ay = __dmul_rn(inv_tau, __dsub_rn(__dmul_rn(V.w, diry), V.y));
        const double az = __dmul_rn(inv_tau, __dsub_rn(__dmul_rn(V.w, 0.0), V.z));
        if (a.f_accel) {
            a.f_accel[3 * i + 0] = ax;
            a.f_accel[3 * i + 1] = ay;
            a.f_accel[3 * i + 2] = az;
        }
        fx = __dadd_rn(fx, ax);
        fy = __dadd_rn(fy, ay);
        fz = __dadd_rn(fz, az);
    }
    if (live && a.enable_ped) {
        fx = __dadd_rn(fx, a.ped_force[3 * i + 0]);
        fy = __dadd_rn(fy, a.ped_force[3 * i + 1]);
        fz = __dadd_rn(fz, a.ped_force[3 * i + 2]);
    }
    if (live && a.f_border) {
        const double2 f = a.f_border[i];
        fx = __dadd_rn(fx, f.x);
        fy = __dadd_rn(fy, f.y);
    }
    if (live && a.f_static) {
        const double2 f = a.f_static[i];
        fx = __dadd_rn(fx, f.x);
        fy = __dadd_rn(fy, f.y);
    }
    if (live && a.f_dynamic) {
        const double2 f = a.f_dynamic[i];
        fx = __dadd_rn(fx, f.x);
        fy = __dadd_rn(fy, f.y);
    }
    if (live) {
        a.f_total[3 * i + 0] = fx;
        a.f_total[3 * i + 1] = fy;
        a.f_total[3 * i + 2] = fz;
    }
    if (!a.update_velocity) return;            // uniform over the grid

    // pedestrian_simulation.py:120-121, stateutils.py:18-23
    double vx = __dadd_rn(V.x, __dmul_rn(a.dt, fx));
    double vy = __dadd_rn(V.y, __dmul_rn(a.dt, fy));
    double vz = __dadd_rn(V.z, __dmul_rn(a.dt, fz));
    double sp = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
    if (sp == 0.0) sp = 1.0;
    const double vmax = __dmul_rn(V.w, a.max_speed_factor);               // pedestrian_state.py:72-73
    const double factor = fmin(1.0, __ddiv_rn(vmax, sp));
    vx = __dmul_rn(vx, factor);
    vy = __dmul_rn(vy, factor);
    vz = __dmul_rn(vz, factor);
    if (live) a.vels[i] = make_double4(vx, vy, vz, V.w);
    double x = L.x, y = L.y, z = L.z;
    if (a.integrate_positions) {                                          // CARLA stub: x += v+ * dt
        x = __dadd_rn(x, __dmul_rn(vx, a.dt));
        y = __dadd_rn(y, __dmul_rn(vy, a.dt));
        z = __dadd_rn(z, __dmul_rn(vz, a.dt));
        if (live) a.locr[i] = make_double4(x, y, z, L.w);
    }
    stage_tile(a, s, live, x, y, z, L.w, vx, vy, vz);
    if (live && a.advance_routes) advance_waypoint(a.routes, a.mm, i, L.x, L.y, a.wp_rw, a.mode_rw, a.sim_time);
}

// calculate_new_velocities for a force array the caller composed itself (pedestrian_simulation.py:117-124 with
// stateutils.cap_velocity, stateutils.py:18-23): v' = v + dt F, clamped to target_speed * max_speed_factor -- numpy's
// operation order, unfused, so the result equals the reference's bit for bit.  Writes the new velocity into `vels`.
__global__ void __launch_bounds__(256) k3_apply_force(int64_t n, double4* vels, const double* __restrict__ force, double dt,
                                                      double max_speed_factor) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 V = vels[i];
    double vx = __dadd_rn(V.x, __dmul_rn(dt, force[3 * i + 0]));
    double vy = __dadd_rn(V.y, __dmul_rn(dt, force[3 * i + 1]));
    double vz = __dadd_rn(V.z, __dmul_rn(dt, force[3 * i + 2]));
    double sp = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
    if (sp == 0.0) sp = 1.0;
    const double factor = fmin(1.0, __ddiv_rn(__dmul_rn(V.w, max_speed_factor), sp));
    vels[i] = make_double4(__dmul_rn(vx, factor), __dmul_rn(vy, factor), __dmul_rn(vz, factor), V.w);
}

}  // namespace sfm
