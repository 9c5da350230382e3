// K4 / K5 / K6 -- the per-tick bookkeeping around the force kernels, moved onto the device.
//
//  K4a  k4_tick_modes +      PedState.apply_current_mode (reference pedestrian_state.py:94-95), PedModeManager.tick
//       k4_gap_acceptance    (ped_mode_manager.py:30-35) and the gap-acceptance loop of PedestrianSimulation.tick
//                            (pedestrian_simulation.py:67-73) with check_traffic (check_traffic.py:7-61) as a
//                            closed-form segment test -- waiting pedestrians compacted first, one thread each, vehicles
//                            staged in shared memory.
//  K4b  advance_waypoint()   arrival test (pedestrian_simulation.py:88-97) + waypoint hand-over
//                            (run_simulation.py:118-132, pedestrian_state.py:83-92) + PedModeManager.set_mode with its
//                            detours (ped_mode_manager.py:37-47); called from K3 (fused) or from k4_advance_waypoints.
//  K5   k5_vehicle_rings     ellipse rings around vehicles (obstacles.py:269-281 as called from :297-329): one thread
//                            per ring point; k5_advance_vehicles moves the centres ballistically (headless stub of the
//                            CARLA vehicle actors, like the position integration in K3).
//  K6   k6_record_frame      one snapshot of (x, y, v_x, v_y, mode) per pedestrian into a device-resident frame buffer
//                            -- the columns pedestrian.csv needs (output_generator.py:35) -- instead of the O(N)
//                            structured-array copy per tick (pedestrian_state.py:100-104).
//
// All arithmetic is float64 in the reference's operation order (unfused), so mode decisions, waypoint hand-overs and
// ring points are reproducible against the host classes.
#pragma once

#include "sfm_common.cuh"

namespace sfm {

// Per-pedestrian mode machine (the fields of PedModeManager that outlive a call; current_mode is sfm_ctx::mode).
struct ModeMachines {
    double* mode_speed = nullptr;        // mode.target_speed          (ped_mode_manager.py:25, 49-69)
    const double* initial_speed = nullptr;   // mode.initial_target_speed  (:21)
    const double* crossing_speed = nullptr;  // mode.crossing_speed        (:22)
    const double* safety_margin = nullptr;   // mode.crossing_safety_margin (:23)
    double* next_mode_time = nullptr;    // mode.next_mode_time        (:27, :54)
    double waiting_time = 5.0;           // mode.waiting_time          (:28)
};

// PedModeManager._activate_mode (ped_mode_manager.py:49-69).  `sim_time` is the machine's last tick time.
__device__ __forceinline__ void activate_mode(const ModeMachines& mm, int64_t i, int mode, double sim_time, uint8_t& cur) {
    if (mode == SFM_IDLE) {
        mm.mode_speed[i] = 0.0;
        mm.next_mode_time[i] = __dadd_rn(sim_time, mm.waiting_time);
    } else if (mode == SFM_WALKING_SIDEWALK) {
        mm.mode_speed[i] = mm.initial_speed[i];
    } else if (mode == SFM_CROSSING_ROAD) {
        mm.mode_speed[i] = mm.crossing_speed[i];
    } else if (mode == SFM_CHECKING_TRAFFIC) {
        mm.mode_speed[i] = 0.0;
    } else if (mode != SFM_ROAD_TO_SIDEWALK) {
        return;                              // unknown mode: ignored like the reference's if/elif chain
    }
    cur = (uint8_t)mode;
}

// PedModeManager.set_mode (ped_mode_manager.py:37-47): the two detours through intermediate modes.
__device__ __forceinline__ void request_mode(const ModeMachines& mm, int64_t i, int wanted, double sim_time, uint8_t& cur) {
    if (cur == SFM_WALKING_SIDEWALK && wanted == SFM_CROSSING_ROAD) activate_mode(mm, i, SFM_CHECKING_TRAFFIC, sim_time, cur);
    else if (cur == SFM_CROSSING_ROAD && wanted == SFM_WALKING_SIDEWALK) activate_mode(mm, i, SFM_ROAD_TO_SIDEWALK, sim_time, cur);
    else activate_mode(mm, i, wanted, sim_time, cur);
}

// ---- gap acceptance ---------------------------------------------------------------------------------------------
struct Traffic {
    int count = 0;
    const double2* center = nullptr;     // vehicle centres               (check_traffic.py:33)
    const double2* velocity = nullptr;   // vehicle velocities            (:34)
    double ext0_x = 0.0, ext0_y = 0.0;   // vehicle_extents[:][0]: the FIRST vehicle's extent, sic (:35-36)
};

__device__ __forceinline__ double norm2d(double x, double y) {
    return __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
}

// Intersection of segment p0-p1 with q0-q1 as shapely's LineString.intersection gives it (host mirror:
// check_traffic._segment_intersection).  Returns false when empty; otherwise the hit as the segment h0-h1 -- a point has
// h0 == h1, collinear overlapping segments yield both ends of the overlap, and a zero-length p (the pedestrian stands on
// its waypoint) hits iff that point lies on q.
__device__ __forceinline__ bool segment_hit(double p0x, double p0y, double p1x, double p1y, double q0x, double q0y,
                                            double q1x, double q1y, double& h0x, double& h0y, double& h1x, double& h1y) {
    const double rx = __dsub_rn(p1x, p0x), ry = __dsub_rn(p1y, p0y);
    const double sx = __dsub_rn(q1x, q0x), sy = __dsub_rn(q1y, q0y);
    const double denom = __dsub_rn(__dmul_rn(rx, sy), __dmul_rn(ry, sx));
    const double qpx = __dsub_rn(q0x, p0x), qpy = __dsub_rn(q0y, p0y);
    if (denom != 0.0) {
        const double t = __ddiv_rn(__dsub_rn(__dmul_rn(qpx, sy), __dmul_rn(qpy, sx)), denom);
        const double u = __ddiv_rn(__dsub_rn(__dmul_rn(qpx, ry), __dmul_rn(qpy, rx)), denom);
        if (!(t >= 0.0 && t <= 1.0 && u >= 0.0 && u <= 1.0)) return false;
        h0x = h1x = __dadd_rn(p0x, __dmul_rn(t, rx));
        h0y = h1y = __dadd_rn(p0y, __dmul_rn(t, ry));
        return true;
    }
    if (__dsub_rn(__dmul_rn(qpx, ry), __dmul_rn(qpy, rx)) != 0.0) return false;      // parallel, not collinear
    const double rr = __dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry));
    if (rr == 0.0) {
        const double ss = __dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy));
        h0x = h1x = p0x;
        h0y = h1y = p0y;
        if (ss == 0.0) return qpx == 0.0 && qpy == 0.0;
        if (__dsub_rn(__dmul_rn(qpx, sy), __dmul_rn(qpy, sx)) != 0.0) return false;
        const double t = __ddiv_rn(-__dadd_rn(__dmul_rn(qpx, sx), __dmul_rn(qpy, sy)), ss);
        return t >= 0.0 && t <= 1.0;
    }
    const double a = __ddiv_rn(__dadd_rn(__dmul_rn(qpx, rx), __dmul_rn(qpy, ry)), rr);
    const double b = __ddiv_rn(__dadd_rn(__dmul_rn(__dsub_rn(q1x, p0x), rx), __dmul_rn(__dsub_rn(q1y, p0y), ry)), rr);
    const double lo = fmax(fmin(a, b), 0.0), hi = fmin(fmax(a, b), 1.0);
    if (!(lo <= hi)) return false;
    h0x = __dadd_rn(p0x, __dmul_rn(lo, rx));
    h0y = __dadd_rn(p0y, __dmul_rn(lo, ry));
    h1x = __dadd_rn(p0x, __dmul_rn(hi, rx));
    h1y = __dadd_rn(p0y, __dmul_rn(hi, ry));
    return true;
}

// intersection.distance(Point(x)) (check_traffic.py:52-54): to the point, or to the nearest point of the overlap segment
__device__ __forceinline__ double hit_distance(double h0x, double h0y, double h1x, double h1y, double x, double y) {
    const double vx = __dsub_rn(h1x, h0x), vy = __dsub_rn(h1y, h0y);
    const double vv = __dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy));
    if (vv == 0.0) return norm2d(__dsub_rn(h0x, x), __dsub_rn(h0y, y));
    const double t = fmin(1.0, fmax(0.0, __ddiv_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, h0x), vx), __dmul_rn(__dsub_rn(y, h0y), vy)), vv)));
    return norm2d(__dsub_rn(__dadd_rn(h0x, __dmul_rn(t, vx)), x), __dsub_rn(__dadd_rn(h0y, __dmul_rn(t, vy)), y));
}

struct ModeTickArgs {
    int64_t n;
    const double4* locr;
    double4* vels;                       // .w = state['target_speed'] (what the clamp reads)
    const double2* wp;
    uint8_t* mode;
    ModeMachines mm;
    Traffic tr;
    double sim_time;
    unsigned long long* counters;        // [0] pedestrians that entered CROSSING_ROAD this tick, [1] idle wake-ups
    int* check_list;                     // [n] rows in CHECKING_TRAFFIC this tick (compacted)
    int* check_count;                    // [1]
    uint8_t* blocked;                    // [n] per list entry: some vehicle blocks the crossing
};

constexpr int K4_THREADS = 128;
constexpr int K4_VEH_TILE = 256;

// Phase 1, one thread per pedestrian: apply_current_mode, the machines' tick, and the list of pedestrians that stand at the
// kerb (CHECKING_TRAFFIC) -- compacted, so that phase 2 runs on dense warps.  Without vehicles everybody crosses at once
// (pedestrian_simulation.py:68-73).
__global__ void __launch_bounds__(K4_THREADS) k4_tick_modes(const ModeTickArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    uint8_t cur = a.mode[i];
    // pedestrian_state.py:94-95 -- runs BEFORE the machines tick: this tick clamps against the speed the mode had at the
    // end of the previous tick
    double4 V = a.vels[i];
    V.w = a.mm.mode_speed[i];
    a.vels[i] = V;
    // ped_mode_manager.py:30-35
    if (cur == SFM_IDLE && a.mm.next_mode_time[i] <= a.sim_time) {
        activate_mode(a.mm, i, SFM_WALKING_SIDEWALK, a.sim_time, cur);
        atomicAdd(a.counters + 1, 1ull);
    }
    if (cur == SFM_CHECKING_TRAFFIC) {
        if (a.tr.count > 0) {
            a.check_list[atomicAdd(a.check_count, 1)] = (int)i;           // order is irrelevant: decisions are per pedestrian
        } else {
            request_mode(a.mm, i, SFM_CROSSING_ROAD, a.sim_time, cur);
            atomicAdd(a.counters + 0, 1ull);
        }
    }
    a.mode[i] = cur;
}

// Phase 2, one thread per (waiting pedestrian, tile of 256 vehicles): check_traffic.py:7-61.  grid.y walks the vehicle
// tiles, so a crowd with few waiting pedestrians still fills the machine; a tile that finds a blocking vehicle raises the
// pedestrian's flag.  Per tile the vehicles are staged in shared memory together with what depends on the vehicle alone
// (half-length vector heading * extent[0], speed).  A vehicle whose swept segment back -> goal cannot touch the
// pedestrian's path (bounding boxes apart by more than a rounding margin) is dismissed after a handful of operations --
// exactly the cases in which the segment test says "no".
__global__ void __launch_bounds__(K4_THREADS) k4_gap_acceptance(const ModeTickArgs a) {
    __shared__ double2 s_center[K4_VEH_TILE], s_vel[K4_VEH_TILE], s_half[K4_VEH_TILE];
    __shared__ double s_speed[K4_VEH_TILE];
    __shared__ float4 s_fb[K4_VEH_TILE];          // float32 copies for the dismissal test: (back.x, back.y, front.x, front.y)
    __shared__ float2 s_uf[K4_VEH_TILE];          //                                         velocity
    const int count = *a.check_count;
    if ((int)(blockIdx.x * K4_THREADS) >= count) return;
    const int v0 = blockIdx.y * K4_VEH_TILE;
    const int m = min(K4_VEH_TILE, a.tr.count - v0);
    for (int v = threadIdx.x; v < m; v += K4_THREADS) {
        const double2 c = a.tr.center[v0 + v], u = a.tr.velocity[v0 + v];
        const double vs = norm2d(u.x, u.y);
        const double dn = (vs == 0.0) ? 1.0 : vs;                                              // stateutils.py:88-90
        s_center[v] = c;
        s_vel[v] = u;
        const double2 h = make_double2(__dmul_rn(__ddiv_rn(u.x, dn), a.tr.ext0_x), __dmul_rn(__ddiv_rn(u.y, dn), a.tr.ext0_y));
        s_half[v] = h;
        s_speed[v] = vs;
        s_fb[v] = make_float4((float)(c.x - h.x), (float)(c.y - h.y), (float)(c.x + h.x), (float)(c.y + h.y));
        s_uf[v] = make_float2((float)u.x, (float)u.y);
    }
    __syncthreads();
    for (int base = blockIdx.x * K4_THREADS; base < count; base += gridDim.x * K4_THREADS) {
        const int k = base + threadIdx.x;
        if (k >= count) continue;
        const int64_t i = a.check_list[k];
        const double margin = a.mm.safety_margin[i];
        if (margin < 0.0) continue;                                        // crosses without looking (check_traffic.py:24)
        const double4 L = a.locr[i];
        const double2 w = a.wp[i];
        const double px = L.x, py = L.y, gx = w.x, gy = w.y;
        const double speed = a.mm.crossing_speed[i];
        const double time_ped = __ddiv_rn(norm2d(__dsub_rn(gx, px), __dsub_rn(gy, py)), speed);      // :27-28
        const double horizon = __dadd_rn(time_ped, margin);
        const double eps = 1.0e-9 * (1.0 + fabs(px) + fabs(py) + fabs(gx) + fabs(gy));
        const double x0 = fmin(px, gx) - eps, x1 = fmax(px, gx) + eps, y0 = fmin(py, gy) - eps, y1 = fmax(py, gy) + eps;
        // the same box in float32, widened by far more than any float32 rounding of the quantities compared with it
        // (1e-5 relative to the coordinates involved, 160 float32 ulps): a first, cheap dismissal on the FP32 pipe
        const float hz = (float)horizon;
        const float wide = 1.0e-5f * (1.0f + (float)(fabs(px) + fabs(py) + fabs(gx) + fabs(gy)) + fabsf(hz));
        const float fx0 = (float)x0 - wide, fx1 = (float)x1 + wide, fy0 = (float)y0 - wide, fy1 = (float)y1 + wide;
        for (int v = 0; v < m; ++v) {
            {
                const float4 fb = s_fb[v];
                const float2 uf = s_uf[v];
                const float txf = fmaf(uf.x, hz, fb.z), tyf = fmaf(uf.y, hz, fb.w);
                const float grow = 1.0e-5f * (fabsf(fb.x) + fabsf(fb.y) + fabsf(txf) + fabsf(tyf) + (fabsf(uf.x) + fabsf(uf.y)) * fabsf(hz));
                if ((fmaxf(fb.x, txf) + grow < fx0) | (fminf(fb.x, txf) - grow > fx1) | (fmaxf(fb.y, tyf) + grow < fy0) |
                    (fminf(fb.y, tyf) - grow > fy1))
                    continue;
            }
            const double2 c = s_center[v], u = s_vel[v], h = s_half[v];
            const double fx = __dadd_rn(c.x, h.x), fy = __dadd_rn(c.y, h.y);                 // front (:35)
            const double bx = __dsub_rn(c.x, h.x), by = __dsub_rn(c.y, h.y);                 // back  (:36)
            const double tx = __dadd_rn(fx, __dmul_rn(u.x, horizon)), ty = __dadd_rn(fy, __dmul_rn(u.y, horizon));
            if (fmax(bx, tx) < x0 || fmin(bx, tx) > x1 || fmax(by, ty) < y0 || fmin(by, ty) > y1) continue;
            double ix, iy, jx, jy;
            if (!segment_hit(px, py, gx, gy, bx, by, tx, ty, ix, iy, jx, jy)) continue;
            const double vs = s_speed[v];
            if (vs == 0.0) continue;                                                          // :48-49
            const double tti_ped = __ddiv_rn(hit_distance(ix, iy, jx, jy, px, py), speed);
            const double tti_front = __ddiv_rn(hit_distance(ix, iy, jx, jy, fx, fy), vs);
            const double tti_back = __ddiv_rn(hit_distance(ix, iy, jx, jy, bx, by), vs);
            if (__dsub_rn(tti_front, margin) < tti_ped && tti_ped < __dadd_rn(tti_back, margin)) {   // :57
                a.blocked[k] = 1;
                break;
            }
        }
    }
}

// Phase 3: every waiting pedestrian no vehicle blocks requests CROSSING_ROAD (pedestrian_simulation.py:72-73).
__global__ void __launch_bounds__(K4_THREADS) k4_gap_commit(const ModeTickArgs a) {
    const int count = *a.check_count;
    for (int k = blockIdx.x * K4_THREADS + threadIdx.x; k < count; k += gridDim.x * K4_THREADS) {
        if (a.blocked[k]) continue;
        const int64_t i = a.check_list[k];
        uint8_t cur = SFM_CHECKING_TRAFFIC;
        request_mode(a.mm, i, SFM_CROSSING_ROAD, a.sim_time, cur);
        a.mode[i] = cur;
        atomicAdd(a.counters + 0, 1ull);
    }
}

// ---- routes ----------------------------------------------------------------------------------------------------
struct Routes {
    const int* end = nullptr;            // [n] one past the pedestrian's last waypoint in `waypoint`
    int* cursor = nullptr;               // [n] next waypoint to hand out (run_simulation.py:123 pop(0))
    const double* waypoint = nullptr;    // [W][3]
    const uint8_t* crossing = nullptr;   // [W] the tuple's crossing_road flag (pedestrian_state.py:84-90)
    double* next_wp3 = nullptr;          // [n][3] full next_waypoint column (z is carried, never used by the forces)
    uint8_t* finished = nullptr;         // [n] arrived with no waypoint left (despawn candidates, run_simulation.py:127)
    double threshold = 2.0;              // walker_config['waypoint_threshold'] (run_simulation.py:39)
    unsigned long long* counters = nullptr;   // [0] hand-overs, [1] newly finished
};

// One pedestrian's arrival test at position (px, py) against its current waypoint w.
__device__ __forceinline__ void advance_waypoint(const Routes& r, const ModeMachines& mm, int64_t i, double px, double py,
                                                 double2* wp, uint8_t* mode, double sim_time) {
    const double2 w = wp[i];
    const double dist = norm2d(__dsub_rn(w.x, px), __dsub_rn(w.y, py));                     // pedestrian_simulation.py:92-93
    if (!(dist < r.threshold)) return;
    const int c = r.cursor[i];
    if (c < r.end[i]) {
        const double nx = r.waypoint[3 * (size_t)c], ny = r.waypoint[3 * (size_t)c + 1], nz = r.waypoint[3 * (size_t)c + 2];
        wp[i] = make_double2(nx, ny);
        r.next_wp3[3 * i] = nx; r.next_wp3[3 * i + 1] = ny; r.next_wp3[3 * i + 2] = nz;
        r.cursor[i] = c + 1;
        uint8_t cur = mode[i];
        request_mode(mm, i, r.crossing[c] ? SFM_CROSSING_ROAD : SFM_WALKING_SIDEWALK, sim_time, cur);
        mode[i] = cur;
        atomicAdd(r.counters + 0, 1ull);
    } else if (!r.finished[i]) {
        r.finished[i] = 1;
        atomicAdd(r.counters + 1, 1ull);
    }
}

__global__ void __launch_bounds__(256) k4_advance_waypoints(int64_t n, const double4* locr, double2* wp, uint8_t* mode,
                                                            Routes r, ModeMachines mm, double sim_time) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 L = locr[i];
    advance_waypoint(r, mm, i, L.x, L.y, wp, mode, sim_time);
}

// ---- despawn (run_simulation.py:127-132: pedestrians that arrived with no waypoint left are destroyed) ---------------
// keep[i] = !finished[i]; an exclusive scan of keep gives every surviving row its new index (order preserved, like the
// boolean-mask copy of PedState.remove_pedestrian, pedestrian_state.py:42-43); one gather per column.
__global__ void k4_keep_flags(int64_t n, const uint8_t* finished, int* keep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep[i] = finished[i] ? 0 : 1;
}

template <typename T>
__global__ void k4_compact(int64_t n, const uint8_t* finished, const int* new_index, const T* src, T* dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !finished[i]) dst[new_index[i]] = src[i];
}

struct double3s { double x, y, z; };      // one next_waypoint row

// ---- vehicle rings ---------------------------------------------------------------------------------------------
struct VehicleArgs {
    int count;
    double2* center;                     // [V]
    const double* yaw_deg;               // [V]  transform.rotation.yaw (degrees, obstacles.py:316)
    const double2* velocity;             // [V]
    const double2* extent;               // [V]  bounding_box.extent.x / .y
    const int* offset;                   // [V + 1]
    double2* point;                      // [offset[V]]
    double size_factor;                  // sqrt(2), obstacles.py:269
    double dt;
};

// centres += velocity * dt  (headless stand-in for the CARLA vehicle actors)
__global__ void k5_advance_vehicles(VehicleArgs a) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.count) return;
    const double2 c = a.center[v], u = a.velocity[v];
    a.center[v] = make_double2(__dadd_rn(c.x, __dmul_rn(u.x, a.dt)), __dadd_rn(c.y, __dmul_rn(u.y, a.dt)));
}

// obstacles.py:269-281: samples = offset[v+1] - offset[v] points at theta = 2 pi i / samples on the ellipse with
// semi-axes extent * size_factor, rotated by the yaw and translated to the centre (carla.Transform with pitch = roll = 0).
__global__ void __launch_bounds__(256) k5_vehicle_rings(VehicleArgs a, int n_points) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_points) return;
    int lo = 0, hi = a.count;                        // largest v with offset[v] <= q
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.offset[mid] <= q) lo = mid; else hi = mid;
    }
    const int v = lo;
    const int samples = a.offset[v + 1] - a.offset[v], k = q - a.offset[v];
    const double theta = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, 3.141592653589793), (double)k), (double)samples);
    const double2 e = a.extent[v], c = a.center[v];
    const double lx = __dmul_rn(__dmul_rn(e.x, cos(theta)), a.size_factor);
    const double ly = __dmul_rn(__dmul_rn(e.y, sin(theta)), a.size_factor);
    const double yaw = __dmul_rn(a.yaw_deg[v], 3.141592653589793 / 180.0);
    const double cy = cos(yaw), sy = sin(yaw);
    a.point[q] = make_double2(__dadd_rn(c.x, __dsub_rn(__dmul_rn(cy, lx), __dmul_rn(sy, ly))),
                              __dadd_rn(c.y, __dadd_rn(__dmul_rn(sy, lx), __dmul_rn(cy, ly))));
}

// ---- recording -------------------------------------------------------------------------------------------------
// frame layout: [n] double4 (x, y, v_x, v_y) followed by [n] uint8 mode -- the columns of pedestrian.csv
__global__ void __launch_bounds__(256) k6_record_frame(int64_t n, const double4* locr, const double4* vels,
                                                       const uint8_t* mode, double4* frame_xyv, uint8_t* frame_mode) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 L = locr[i], V = vels[i];
    frame_xyv[i] = make_double4(L.x, L.y, V.x, V.y);
    frame_mode[i] = mode[i];
}

}  // namespace sfm
