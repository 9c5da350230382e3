// K2 -- uniform-grid cell lists and the cutoff-limited pedestrian x point-set forces, float64, sm_100a.
//
// Replaces BorderForce._get_force (reference forces.py:138-179) and ObstacleForce._get_force (forces.py:208-283), whose
// per-pedestrian linear filters (forces.py:149-151, :222-225) become a cell-list lookup.
//
//  K2a  binning.  Set items (border sections / obstacles) are keyed by the grid cell of their *centre* -- the point the
//       reference's cutoff is tested against -- and ordered by (cell, index) with a stable LSD radix sort, so every
//       pedestrian visits its candidate items in one fixed order.  The cell edge is >= the largest cutoff of the set,
//       hence every item within the cutoff of a pedestrian lies in the 3x3 cells around it.  Pedestrians are binned
//       too (Morton-ordered counting sort) purely for locality: a CTA then owns pedestrians that share candidates.
//  K2b/c  one CTA = 32 spatially adjacent pedestrians (one per lane) x 4 warps that split the candidate items.  The
//       warps walk the cells overlapping the pedestrians' bounding box (+1 ring), reject items whose cutoff disc misses
//       the box, stage an item's points in shared memory and let every lane that passes the reference's exact filter
//       scan them.
//
// Bit-exact enumeration: the filter sqrt(dx*dx + dy*dy) < cutoff is evaluated in float64 with numpy's operation order
// (np.linalg.norm = sqrt(add.reduce(x*x)), unfused).  The nearest-point argmin is found in two stages: a float32
// scan over centre-relative coordinates brackets every point that could be the float64 minimum -- squared distances within
// the item's rigorous rounding bound `tol` of the float32 minimum, a contiguous index window that is almost always 1-3
// points wide -- and that window is then scanned with numpy's exact float64 arithmetic, np.argmin's first-index tie rule
// included.  So the (pedestrian, item, nearest point) triplets equal the reference's exactly, and the forces (float64,
// reference operation order) agree to a few ulp.
//
// Direct path (items with a chord record, i.e. every host-uploaded item): the item's points are modelled as UNIFORMLY
// spaced on the chord from its first to its last point, g_k = a + k/(P-1) u, with E >= max_k |q_k - g_k| measured at
// upload (float64, against the float32-rounded a and u, plus the pedestrian's own float32 rounding).  With kappa the
// unclamped projection index of the pedestrian on the chord line, k* = clamp(round(kappa)), D* = |p - g_k*| and spacing s,
//     dist(p, q_k) <= min_k dist(p, q_k)   ==>   |p - g_k| <= D* + 2E   ==>   (k - kappa)^2 <= (k* - kappa)^2 + 4E(D* + E)/s^2,
// so the float64 argmin -- exact ties included -- lies in an index window around kappa that is 1-2 points wide for a
// straight section (E ~ rounding) and grows with curvature; windows wider than 8 points (any lane of the warp) fall back
// to the chunk-pruned float32 scan below.  Inside the window the distances are evaluated with numpy's float64 arithmetic
// directly -- no float32 stage, no staging in shared memory: ~100 instructions per (pedestrian, item) instead of ~1,300.
//
// Why the bracket is a superset: coordinates are taken relative to the item's centre c in float64 and rounded once to
// float32, |x~ - (x - c)| <= 2^-24 |x - c|; with M >= every |component| involved (M = max(cutoff, ring extent), per item,
// computed at upload) a difference carries <= 3.1 * 2^-24 M and the squared distance <= 42 * 2^-24 M^2 =: Delta of
// absolute error, all float32 roundings included.  If q* is the float64 argmin and q~ the float32 one,
// d~(q*) <= D(q*) + Delta <= D(q~) + Delta <= d~(q~) + 2 Delta.  tol = 128 * 2^-24 M^2 >= 2 Delta with slack for the
// float32 evaluation of the threshold itself.
#pragma once

#include "sfm_common.cuh"

namespace sfm {

constexpr int K2_WARPS = 4;             // warps sharing one group of 32 pedestrians
constexpr int K2_THREADS = 32 * K2_WARPS;
constexpr int K2_CHUNK = 256;           // points (float2, centre-relative) one warp stages per pass
constexpr int SORT_THREADS = 512;
constexpr int SORT_RADIX_BITS = 4;

__device__ __forceinline__ double norm2_np(double dx, double dy) {   // np.linalg.norm of a 2-vector
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

__device__ __forceinline__ int cell_coord(double v, double v0, double inv_cell, int n) {
    int c = (int)floor((v - v0) * inv_cell);
    return min(max(c, 0), n - 1);
}

// ---- K2a: set items -> (cell, index) order ------------------------------------------------------------------------
__global__ void k2_item_keys(const double2* __restrict__ center, int count, CellGrid g, unsigned* __restrict__ key,
                             int* __restrict__ val) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    const double2 c = center[s];
    key[s] = (unsigned)(cell_coord(c.y, g.y0, g.inv_cell, g.ny) * g.nx + cell_coord(c.x, g.x0, g.inv_cell, g.nx));
    val[s] = s;
}

// Stable LSD radix sort of (key, val) by one CTA (sets of a few thousand items: one launch, lowest latency); each thread
// owns a contiguous run of the input, so equal keys keep their input (= index) order without any atomics.  Larger sets
// take the many-CTA sort below (sfm_api.cu: sort_items).
__global__ void __launch_bounds__(SORT_THREADS) k2_radix_sort(unsigned* key_a, int* val_a, unsigned* key_b, int* val_b,
                                                              int count, int key_bits) {
    constexpr int R = 1 << SORT_RADIX_BITS;
    __shared__ int hist[R][SORT_THREADS];
    __shared__ int digit_base[R];
    const int tid = threadIdx.x;
    const int per = (count + SORT_THREADS - 1) / SORT_THREADS;
    const int lo = min(tid * per, count), hi = min(lo + per, count);
    unsigned* kin = key_a;
    int* vin = val_a;
    unsigned* kout = key_b;
    int* vout = val_b;
    for (int shift = 0; shift < key_bits; shift += SORT_RADIX_BITS) {
        int cnt[R];
#pragma unroll
        for (int d = 0; d < R; ++d) cnt[d] = 0;
        for (int i = lo; i < hi; ++i) {
            const int d = (kin[i] >> shift) & (R - 1);
#pragma unroll
            for (int e = 0; e < R; ++e) cnt[e] += (e == d);
        }
#pragma unroll
        for (int d = 0; d < R; ++d) hist[d][tid] = cnt[d];
        __syncthreads();
        // exclusive scan over (digit-major, thread-minor): per-digit scan across threads, then digit bases
        if (tid < R) {
            int run = 0;
            for (int t = 0; t < SORT_THREADS; ++t) {
                const int c = hist[tid][t];
                hist[tid][t] = run;
                run += c;
            }
            digit_base[tid] = run;
        }
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int d = 0; d < R; ++d) {
                const int c = digit_base[d];
                digit_base[d] = run;
                run += c;
            }
        }
        __syncthreads();
        int pos[R];
#pragma unroll
        for (int d = 0; d < R; ++d) pos[d] = digit_base[d] + hist[d][tid];
        for (int i = lo; i < hi; ++i) {
            const unsigned k = kin[i];
            const int d = (k >> shift) & (R - 1);
            int p = 0;
#pragma unroll
            for (int e = 0; e < R; ++e) {
                if (e == d) {
                    p = pos[e];
                    pos[e] = p + 1;
                }
            }
            kout[p] = k;
            vout[p] = vin[i];
        }
        __syncthreads();
        unsigned* tk = kin; kin = kout; kout = tk;
        int* tv = vin; vin = vout; vout = tv;
    }
    // result is in (kin, vin); make sure it ends in buffer a
    if (kin != key_a) {
        for (int i = lo; i < hi; ++i) {
            key_a[i] = kin[i];
            val_a[i] = vin[i];
        }
    }
}

// ---- the same sort for large sets: stable LSD radix sort over many CTAs, 8-bit digits -----------------------------
// One CTA owns a contiguous run of MSORT_TILE items.  Per pass: (1) per-CTA digit histograms, written digit-major so that
// (2) one exclusive scan over [digit][CTA] yields every CTA's first output slot per digit, (3) a stable scatter -- the CTA
// walks its run in sub-tiles of 256 items; inside a sub-tile an item's rank among equal digits is its rank inside the warp
// (match.any + popc of the lower lanes) plus the counts of the lower warps, so equal keys keep their input order and the
// result is deterministic: no atomics on global memory anywhere.
constexpr int MSORT_THREADS = 256;
constexpr int MSORT_ITEMS = 16;
constexpr int MSORT_TILE = MSORT_THREADS * MSORT_ITEMS;
constexpr int MSORT_BITS = 8;
constexpr int MSORT_R = 1 << MSORT_BITS;

__global__ void __launch_bounds__(MSORT_THREADS) k2_msort_hist(const unsigned* __restrict__ key, int count, int shift,
                                                              int* __restrict__ hist, int nblk) {
    __shared__ int h[MSORT_R];
    const int tid = threadIdx.x;
    h[tid] = 0;
    __syncthreads();
    const int base = blockIdx.x * MSORT_TILE;
    for (int t = 0; t < MSORT_ITEMS; ++t) {
        const int i = base + t * MSORT_THREADS + tid;
        if (i < count) atomicAdd(&h[(key[i] >> shift) & (MSORT_R - 1)], 1);      // shared-memory counter: order-free
    }
    __syncthreads();
    hist[tid * nblk + blockIdx.x] = h[tid];
}

__global__ void __launch_bounds__(MSORT_THREADS) k2_msort_scatter(const unsigned* __restrict__ kin, const int* __restrict__ vin,
                                                                 unsigned* __restrict__ kout, int* __restrict__ vout,
                                                                 int count, int shift, const int* __restrict__ first_slot,
                                                                 int nblk) {
    constexpr int W = MSORT_THREADS / 32;
    __shared__ int slot[MSORT_R];                 // next output slot of this CTA per digit
    __shared__ int cnt[W][MSORT_R];               // per warp: items of each digit in the current sub-tile
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    slot[tid] = first_slot[tid * nblk + blockIdx.x];
    const int base = blockIdx.x * MSORT_TILE;
    for (int t = 0; t < MSORT_ITEMS; ++t) {
#pragma unroll
        for (int w = 0; w < W; ++w) cnt[w][tid] = 0;
        __syncthreads();
        const int i = base + t * MSORT_THREADS + tid;
        const bool live = i < count;
        const unsigned k = live ? kin[i] : 0u;
        const int d = live ? (int)((k >> shift) & (MSORT_R - 1)) : MSORT_R + lane;        // dead lanes match nobody
        const unsigned same = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(same & ((1u << lane) - 1u));
        if (live && rank == 0) cnt[wid][d] = __popc(same);
        __syncthreads();
        if (live) {
            int before = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) before += (w < wid) ? cnt[w][d] : 0;
            const int p = slot[d] + before + rank;
            kout[p] = k;
            vout[p] = vin[i];
        }
        __syncthreads();
        int total = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) total += cnt[w][tid];
        slot[tid] += total;
        __syncthreads();
    }
}

// ---- exclusive scan over many CTAs: tiles of SCAN_TILE entries scanned in place, the tile totals scanned by one CTA
//      (k2_exclusive_scan below), then added back ----------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_PER_THREAD = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

__global__ void __launch_bounds__(SCAN_THREADS) k2_scan_tiles(int* data, int n, int* __restrict__ tile_sum) {
    __shared__ int warp_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int base = blockIdx.x * SCAN_TILE + tid * SCAN_PER_THREAD;
    int v[SCAN_PER_THREAD], mine = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0;
        mine += v[k];
    }
    int x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int w = warp_sum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        warp_sum[lane] = w;
    }
    __syncthreads();
    int run = (wid ? warp_sum[wid - 1] : 0) + x - mine;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
    if (tid == SCAN_THREADS - 1) tile_sum[blockIdx.x] = run;
}

__global__ void k2_scan_add(int* data, int n, const int* __restrict__ tile_prefix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] += tile_prefix[i / SCAN_TILE];
}

// cell_start[c] = first position in the sorted key array whose key is >= c  (c in [0, ncell])
__global__ void k2_cell_bounds(const unsigned* __restrict__ sorted_key, int count, int ncell,
                               int* __restrict__ cell_start) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ncell) return;
    int lo = 0, hi = count;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sorted_key[mid] < (unsigned)c) lo = mid + 1; else hi = mid;
    }
    cell_start[c] = lo;
}

// ---- K2a: pedestrians -> Morton-ordered permutation (locality only; order inside a cell is irrelevant) ------------
__device__ __forceinline__ unsigned morton2(unsigned x, unsigned y) {
    auto spread = [](unsigned v) {
        v &= 0xffffu;
        v = (v | (v << 8)) & 0x00ff00ffu;
        v = (v | (v << 4)) & 0x0f0f0f0fu;
        v = (v | (v << 2)) & 0x33333333u;
        v = (v | (v << 1)) & 0x55555555u;
        return v;
    };
    return spread(x) | (spread(y) << 1);
}

__global__ void k2_ped_count(const double4* __restrict__ locr, int n, CellGrid g, unsigned* __restrict__ ped_cell,
                             int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 L = locr[i];
    const unsigned c = morton2((unsigned)cell_coord(L.x, g.x0, g.inv_cell, g.nx),
                               (unsigned)cell_coord(L.y, g.y0, g.inv_cell, g.ny));
    ped_cell[i] = c;
    atomicAdd(&count[c], 1);
}

// in-place exclusive scan of data[0..n) by one CTA (n is a few 10^5 at most here)
__global__ void __launch_bounds__(1024) k2_exclusive_scan(int* data, int n) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = (i < n) ? data[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int w = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sum[lane] = w;
        }
        __syncthreads();
        const int prefix = carry + (wid ? warp_sum[wid - 1] : 0) + x - v;
        if (i < n) data[i] = prefix;
        __syncthreads();
        if (tid == 1023) carry = prefix + v;
        __syncthreads();
    }
}

__global__ void k2_ped_fill(const unsigned* __restrict__ ped_cell, int n, const int* __restrict__ cell_start,
                            int* __restrict__ cursor, int* __restrict__ perm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned c = ped_cell[i];
    perm[cell_start[c] + atomicAdd(&cursor[c], 1)] = i;
}

// ---- K2b / K2c ---------------------------------------------------------------------------------------------------
struct SegArgs {
    const double4* locr;
    const double4* vels;
    const uint8_t* mode;
    const int* perm;
    int n;
    // the set
    const double2* center;
    const double* cutoff;
    const double2* velocity;
    const int* offset;
    const double2* point;
    const float* tol;               // [count] float32 bracket width of the two-stage nearest-point search
    const int* chunk_first;         // [count] first entry of the item's pruning chunks in `chunk`, or -1 (may be null)
    const float4* chunk;            // two float4 per chunk: (a_x, a_y, u_x, u_y), (1/|u|^2, deviation, -, -), centre-relative
    const float4* chord0;           // [count] two float4 per item: (a_x, a_y, u_x, u_y), (1/|u|^2, E, P-1, 1/(P-1)); may be null
    CellGrid grid;
    const int* cell_start;
    const int* cell_item;
    // parameters
    MoussaidD mp;
    double border_a, border_b, neg_inv_border_b;   // -1 / b, so that exp(-dist / b) costs no division
    int use_radius;
    double2* f_out;                 // [n]
    int n_groups;                   // ceil(n / 32) pedestrian groups
    int* work_counter;              // persistent mode: next group to hand out (zeroed before the launch); null: one CTA per group
    unsigned long long* eval_count; // optional [2]: (pedestrian, item) pairs inside the cutoff, sum of their point counts
    long long* emit;                // optional [capacity][3]
    unsigned long long* emit_count;
    long long emit_capacity;
};

// `nrm` = np.linalg.norm(loc - point) as the nearest-point search already evaluated it (numpy's arithmetic), or < 0 to have
// it computed here.  Unit vectors are formed with one reciprocal and two products instead of two divisions (a few ulp from
// numpy's quotient; the enumeration, which is what must be bit-exact, does not depend on them).
template <int KIND>   // 0: border (exponential repulsion), 1: obstacle (Moussaid)
__device__ __forceinline__ double2 segment_force(const SegArgs& a, double px, double py, double vx, double vy,
                                                 double radius, double2 P, double2 ovel, double nrm) {
    if (KIND == 0) {
        // forces.py:158-165: direction from the border point to the pedestrian
        const double dx = __dsub_rn(px, P.x), dy = __dsub_rn(py, P.y);
        if (nrm < 0.0) nrm = norm2_np(dx, dy);
        const double rinv = __ddiv_rn(1.0, (nrm == 0.0) ? 1.0 : nrm);
        const double ex = __dmul_rn(dx, rinv), ey = __dmul_rn(dy, rinv);
        double dist = nrm;
        if (a.use_radius) dist = __dsub_rn(dist, radius);
        const double mag = exp(__dmul_rn(dist, a.neg_inv_border_b));
        return make_double2(__dmul_rn(__dmul_rn(ex, a.border_a), mag), __dmul_rn(__dmul_rn(ey, a.border_a), mag));
    } else {
        // forces.py:233-270
        const double dx = __dsub_rn(P.x, px), dy = __dsub_rn(P.y, py);
        if (nrm < 0.0) nrm = norm2_np(dx, dy);
        const double rinv = __ddiv_rn(1.0, (nrm == 0.0) ? 1.0 : nrm);
        const double ex = __dmul_rn(dx, rinv), ey = __dmul_rn(dy, rinv);
        double dl = nrm;
        if (a.use_radius) dl = __dsub_rn(dl, radius);
        const double wx = __dsub_rn(vx, ovel.x), wy = __dsub_rn(vy, ovel.y);
        const double Dx = __dadd_rn(__dmul_rn(a.mp.lambda, wx), ex), Dy = __dadd_rn(__dmul_rn(a.mp.lambda, wy), ey);
        const double Dn = norm2_np(Dx, Dy);
        const double Dinv = __ddiv_rn(1.0, (Dn == 0.0) ? 1.0 : Dn);
        const double tx = __dmul_rn(Dx, Dinv), ty = __dmul_rn(Dy, Dinv);
        const double nx = __dmul_rn(ty, -1.0), ny = tx;
        // stateutils.py:104-112: angle(e) - angle(t) wrapped once into [-pi, pi] == atan2(t x e, t . e): one atan2 instead
        // of two (equal to the reference's difference to a few ulp; exactly 0 when e == t, e.g. a standing pedestrian
        // before a static obstacle, so sign(theta') keeps the reference's value there)
        double th = atan2(__dsub_rn(__dmul_rn(tx, ey), __dmul_rn(ty, ex)), __dadd_rn(__dmul_rn(tx, ex), __dmul_rn(ty, ey)));
        if (Dn == 0.0 || nrm == 0.0) th = __dsub_rn(atan2(ey, ex), atan2(ty, tx));   // zero vectors: numpy's atan2(0, 0) = 0 terms
        const double B = __dmul_rn(a.mp.gamma, Dn);
        th = __dadd_rn(th, __dmul_rn(B, -a.mp.epsilon));
        const double base = __ddiv_rn(__dmul_rn(-1.0, dl), B);
        const double qv = __dmul_rn(__dmul_rn(a.mp.n_prime, B), th);
        const double qt = __dmul_rn(__dmul_rn(a.mp.n, B), th);
        const double fv = __dmul_rn(__dmul_rn(-1.0, a.mp.A), exp(__dsub_rn(base, __dmul_rn(qv, qv))));
        const double sgn = (th > 0.0) ? 1.0 : ((th < 0.0) ? -1.0 : th);      // np.sign (nan stays nan)
        const double ft = __dmul_rn(__dmul_rn(__dmul_rn(-1.0, a.mp.A), sgn), exp(__dsub_rn(base, __dmul_rn(qt, qt))));
        return make_double2(__dadd_rn(__dmul_rn(fv, tx), __dmul_rn(ft, nx)),
                            __dadd_rn(__dmul_rn(fv, ty), __dmul_rn(ft, ny)));
    }
}

// np.argmin(np.linalg.norm(loc - points, axis=-1)) with numpy's arithmetic over the bracket [o0, o1); `best` receives the
// winning distance (what the force needs next) or -1 when no point compared finite.
__device__ __forceinline__ int exact_argmin(const double2* __restrict__ point, int o0, int o1, double px, double py,
                                            double& best_out) {
    double best = __longlong_as_double(0x7ff0000000000000LL);
    int best_q = o0;
    for (int q = o0; q < o1; ++q) {
        const double2 P = point[q];
        const double d = norm2_np(__dsub_rn(px, P.x), __dsub_rn(py, P.y));
        if (d < best) {
            best = d;
            best_q = q;
        }
    }
    best_out = (best < __longlong_as_double(0x7ff0000000000000LL)) ? best : -1.0;
    return best_q;
}


// ---- stage 1 helpers: float32 scans over staged, centre-relative points ---------------------------------------------
constexpr int K2_PRUNE_CHUNK = 16;      // points per pruning chunk (chord + deviation, built at upload)
constexpr int K2_PRUNE_MAX = K2_CHUNK / K2_PRUNE_CHUNK;      // chunk tables cover items that fit one staging pass
constexpr float K2_FAR = 1.0e18f;       // pad value: d2 = 2e36, finite, never a candidate

// One pass over sp[q0 .. q1) (q0, q1 multiples of 4; base0 = global index of sp[0]): keeps the running minimum m1 of the
// squared distances and an index window [lo, hi] that contains EVERY point with d2 <= (final m1) + tol.  Invariant: a
// point enters the window when it is within tol of the running minimum (which only decreases, so nothing that ends within
// tol of the final minimum is ever skipped); the window restarts only when a group's minimum undercuts the running one
// by more than tol -- everything seen before is then > final minimum + tol.
__device__ __forceinline__ void scan_pass(const float2* __restrict__ sp, int q0, int q1, float pxf, float pyf, float tol,
                                          int base0, float& m1, int& lo, int& hi) {
#pragma unroll 2
    for (int q = q0; q < q1; q += 4) {
        const float4 A = *reinterpret_cast<const float4*>(&sp[q]);
        const float4 B = *reinterpret_cast<const float4*>(&sp[q + 2]);
        const float ax = pxf - A.x, ay = pyf - A.y, bx = pxf - A.z, by = pyf - A.w;
        const float cx = pxf - B.x, cy = pyf - B.y, ex = pxf - B.z, ey = pyf - B.w;
        const float d0 = fmaf(ax, ax, ay * ay), d1 = fmaf(bx, bx, by * by);
        const float d2 = fmaf(cx, cx, cy * cy), d3 = fmaf(ex, ex, ey * ey);
        const float m4 = fminf(fminf(d0, d1), fminf(d2, d3));
        if (m4 <= m1 + tol) {
            if (m4 < m1 - tol) { lo = 0x7fffffff; hi = -1; }
            m1 = fminf(m1, m4);
            const float thr = m1 + tol;
            const int base = base0 + q;
            if (d0 <= thr) { lo = min(lo, base); hi = max(hi, base); }
            if (d1 <= thr) { lo = min(lo, base + 1); hi = max(hi, base + 1); }
            if (d2 <= thr) { lo = min(lo, base + 2); hi = max(hi, base + 2); }
            if (d3 <= thr) { lo = min(lo, base + 3); hi = max(hi, base + 3); }
        }
    }
}

// centre-relative float32 copy of points [c0, c0 + m) into sp[0 .. round_up(m, pad)), padded with K2_FAR
__device__ __forceinline__ void stage_points(float2* __restrict__ sp, const double2* __restrict__ point, int c0, int m, int pad,
                                             double cxs, double cys, int lane) {
    const int mp = (m + pad - 1) & ~(pad - 1);
    __syncwarp();
    for (int q = lane; q < mp; q += 32) {
        float2 v = make_float2(K2_FAR, K2_FAR);
        if (q < m) {
            const double2 P = point[c0 + q];
            v = make_float2((float)(P.x - cxs), (float)(P.y - cys));
        }
        sp[q] = v;
    }
    __syncwarp();
}

// One CTA = 32 spatially adjacent pedestrians (lane = pedestrian) x K2_WARPS warps.  All warps walk the same candidate
// list -- metadata of 32 items is fetched lane-parallel and broadcast by shuffles -- and share it by item index: warp k
// takes the items with index % K2_WARPS == k that survive the bounding-box test.  Each warp stages its item's points in its own
// shared-memory slice (warp-level barriers only), scans them, and keeps a partial force per pedestrian; the partials
// are combined in warp order at the end, so the summation order is fixed.
#ifndef SFM_K2_MINB
#define SFM_K2_MINB 6                  // min CTAs per SM (register cap: 80) of the segment kernels, profiles/k2_minb_sweep_r1.log
#endif
template <int KIND>
__global__ void __launch_bounds__(K2_THREADS, SFM_K2_MINB) k2_segments(const SegArgs a) {
    __shared__ __align__(16) float2 sp[K2_WARPS][K2_CHUNK];
    __shared__ __align__(16) float4 sc[K2_WARPS][2 * K2_PRUNE_MAX];
    __shared__ double2 part[K2_WARPS][32];
    __shared__ int s_group;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // Persistent mode (work_counter != null): the grid is a few CTAs per SM that pull pedestrian groups from a counter, so
    // the kernel occupies a fixed share of every SM and the pair kernel runs beside it for its whole duration.
    for (bool first = true;; first = false) {
    if (a.work_counter) {
        if (tid == 0) s_group = atomicAdd(a.work_counter, 1);
        __syncthreads();
    } else if (!first) {
        break;
    }
    const int group = a.work_counter ? s_group : (int)blockIdx.x;
    if (group >= a.n_groups) break;
    const int slot = group * 32 + lane;
    const bool active = slot < a.n;
    const int i = active ? a.perm[slot] : -1;
    double px = 0.0, py = 0.0, radius = 0.0, vx = 0.0, vy = 0.0;
    if (active) {
        const double4 L = a.locr[i];
        const double4 V = a.vels[i];
        px = L.x; py = L.y; radius = L.w; vx = V.x; vy = V.y;
    }
    // bounding box of the warp's live pedestrians
    const double BIG = 1.0e300;
    double x0 = active ? px : BIG, x1 = active ? px : -BIG, y0 = active ? py : BIG, y1 = active ? py : -BIG;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x0 = fmin(x0, __shfl_xor_sync(0xffffffffu, x0, o));
        x1 = fmax(x1, __shfl_xor_sync(0xffffffffu, x1, o));
        y0 = fmin(y0, __shfl_xor_sync(0xffffffffu, y0, o));
        y1 = fmax(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    const CellGrid g = a.grid;
    const int cx0 = max(cell_coord(x0, g.x0, g.inv_cell, g.nx) - 1, 0);
    const int cx1 = min(cell_coord(x1, g.x0, g.inv_cell, g.nx) + 1, g.nx - 1);
    const int cy0 = max(cell_coord(y0, g.y0, g.inv_cell, g.ny) - 1, 0);
    const int cy1 = min(cell_coord(y1, g.y0, g.inv_cell, g.ny) + 1, g.ny - 1);
    const double INF = __longlong_as_double(0x7ff0000000000000LL);

    double fx = 0.0, fy = 0.0;
    for (int cy = cy0; cy <= cy1; ++cy) {
        // cells of one grid row are contiguous in the (cell, index) order: one item range per row
        const int k_begin = a.cell_start[cy * g.nx + cx0], k_end = a.cell_start[cy * g.nx + cx1 + 1];
        for (int k0 = k_begin; k0 < k_end; k0 += 32) {
            const int kk = k0 + lane;
            int s_l = -1, o0_l = 0, o1_l = 0;
            double2 cen_l = make_double2(0.0, 0.0);
            double cut_l = 0.0;
            float tol_l = 0.0f;
            int cf_l = -1;
            bool accept_l = false;
            if (kk < k_end) {
                s_l = a.cell_item[kk];
                cen_l = a.center[s_l];
                cut_l = a.cutoff[s_l];
                tol_l = a.tol[s_l];
                if (a.chunk_first) cf_l = a.chunk_first[s_l];
                o0_l = a.offset[s_l];
                o1_l = a.offset[s_l + 1];
                // conservative reject: the cutoff disc misses the warp's bounding box (slack covers rounding)
                const double gx = fmax(fmax(x0 - cen_l.x, cen_l.x - x1), 0.0);
                const double gy = fmax(fmax(y0 - cen_l.y, cen_l.y - y1), 0.0);
                accept_l = !(gx * gx + gy * gy > cut_l * cut_l * 1.000000001);
            }
            // item -> warp by the item's own index, so a pedestrian's summation order does not depend on which other
            // pedestrians share its group (results are reproducible under any pedestrian ordering)
            unsigned todo = __ballot_sync(0xffffffffu, accept_l && (s_l % K2_WARPS) == wid);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int s = __shfl_sync(0xffffffffu, s_l, src);
                const double cxs = __shfl_sync(0xffffffffu, cen_l.x, src), cys = __shfl_sync(0xffffffffu, cen_l.y, src);
                const double cut = __shfl_sync(0xffffffffu, cut_l, src);
                const int o0 = __shfl_sync(0xffffffffu, o0_l, src), o1 = __shfl_sync(0xffffffffu, o1_l, src);
                // the reference's filter, bit for bit: norm(loc - centre) < cutoff  (forces.py:149-150, :222-223) -- decided in
                // float32 on centre-relative coordinates unless the squared distance sits within 1e-5 (relative) of cutoff^2,
                // 40x the float32 evaluation error; only that margin takes the float64 test
                const float pxf = (float)(px - cxs), pyf = (float)(py - cys);
                bool pass;
                {
                    const float cutf = (float)cut, c2 = cutf * cutf, d2f = fmaf(pxf, pxf, pyf * pyf);
                    if (cutf > 0.0f && d2f < c2 * 0.99999f) pass = true;
                    else if (cutf > 0.0f && d2f > c2 * 1.00001f) pass = false;
                    else pass = norm2_np(__dsub_rn(px, cxs), __dsub_rn(py, cys)) < cut;
                    pass = pass && active;
                }
                if (!__any_sync(0xffffffffu, pass)) continue;
                const int np = o1 - o0;
                if (a.eval_count) {
                    const int hits = __popc(__ballot_sync(0xffffffffu, pass));
                    if (lane == 0) {
                        atomicAdd(a.eval_count, (unsigned long long)hits);
                        atomicAdd(a.eval_count + 1, (unsigned long long)hits * (unsigned long long)np);
                    }
                }
                int best_q = o0;
                // ---- direct path: index window around the projection on the item's chord (header), float64 inside it
                bool direct = false;
                int k_lo = 0, k_cnt = 0;
                double best = -1.0;                 // distance to the nearest point where the search already has it
                float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
                if (a.chord0) {
                    c0 = __ldg(&a.chord0[2 * s]);
                    c1 = __ldg(&a.chord0[2 * s + 1]);
                }
                if (c1.x > 0.0f || (a.chord0 && c1.z == 0.0f)) {        // (uniform: the item has a usable chord record)
                    const float wx = pxf - c0.x, wy = pyf - c0.y, nm1 = c1.z;
                    const float kap = fmaf(wx, c0.z, wy * c0.w) * c1.x * nm1;
                    const float ks = fminf(fmaxf(rintf(kap), 0.0f), nm1);
                    const float fr = ks * c1.w;
                    const float gx = fmaf(-fr, c0.z, wx), gy = fmaf(-fr, c0.w, wy);
                    const float Ds = sqrtf(fmaf(gx, gx, gy * gy));
                    const float del = kap - ks;
                    const float Q = 4.0f * c1.y * (Ds + c1.y) * (nm1 * nm1) * c1.x;
                    const float rho = fmaf(sqrtf(fmaf(del, del, Q)), 1.001f, 0.004f);
                    const float lo_f = fminf(fmaxf(ceilf(kap - rho), 0.0f), ks), hi_f = fmaxf(fminf(floorf(kap + rho), nm1), ks);
                    const bool ok = (hi_f - lo_f) < 8.0f;                                          // NaN / inf -> false
                    direct = __all_sync(0xffffffffu, !pass || ok);
                    if (direct && pass) {
                        k_lo = (int)lo_f;
                        k_cnt = (int)hi_f - k_lo + 1;
                    }
                }
                if (direct) {
                    best = __longlong_as_double(0x7ff0000000000000LL);
                    for (int j = 0; __any_sync(0xffffffffu, j < k_cnt); ++j) {
                        if (j < k_cnt) {
                            const int q = o0 + k_lo + j;
                            const double2 P = a.point[q];
                            const double d = norm2_np(__dsub_rn(px, P.x), __dsub_rn(py, P.y));
                            if (d < best) {                 // strict: first index on exact ties (forces.py:154, :228)
                                best = d;
                                best_q = q;
                            }
                        }
                    }
                    if (pass && !(best < __longlong_as_double(0x7ff0000000000000LL))) {
                        best_q = exact_argmin(a.point, o0, o1, px, py, best);      // non-finite coordinates: full scan
                    }
                } else {
                // Nearest point, stage 1 (float32, all lanes in lock step over the staged points): the smallest squared
                // distance m1 and the index window [lo, hi] of every point within tol of it, in one pass.
                const float tol = __shfl_sync(0xffffffffu, tol_l, src);
                float m1 = 3.0e38f;
                int lo = 0x7fffffff, hi = -1;
                const int cf = __shfl_sync(0xffffffffu, cf_l, src);
                if (cf >= 0) {
                    // Pruned search (items with a chunk table: 48 .. 256 points).  Every run of 16 points is covered by
                    // its chord a + t u (t in [0, 1]) and the largest distance `dev` of its points from that chord, so no
                    // point of the chunk is closer than dist(p, chord) - dev.  A chunk is scanned only if that bound does
                    // not exceed the distance to the nearest chunk START (an actual point, hence an upper bound on the
                    // minimum) -- for at least one pedestrian of the warp that passed the filter.
                    const int nch = (np + K2_PRUNE_CHUNK - 1) / K2_PRUNE_CHUNK;
                    __syncwarp();
                    if (lane < 2 * nch) sc[wid][lane] = a.chunk[2 * cf + lane];
                    __syncwarp();
                    float ds[K2_PRUNE_MAX];
                    float ub = 3.0e38f;
#pragma unroll
                    for (int c = 0; c < K2_PRUNE_MAX; ++c) {
                        ds[c] = 3.0e38f;
                        if (c < nch) {
                            const float4 g = sc[wid][2 * c], h = sc[wid][2 * c + 1];
                            const float wx = pxf - g.x, wy = pyf - g.y;
                            ub = fminf(ub, fmaf(wx, wx, wy * wy));
                            const float t = __saturatef(fmaf(wx, g.z, wy * g.w) * h.x);
                            const float rx = fmaf(-t, g.z, wx), ry = fmaf(-t, g.w, wy);
                            ds[c] = fmaf(rx, rx, ry * ry);
                        }
                    }
                    const float b = ub + tol;
                    const float rb2 = 2.000002f * sqrtf(b);
                    const float slack = b + 4.0f * tol;          // float32 error of the chord distance: <= 2 tol, doubled
                    unsigned need = 0;
#pragma unroll
                    for (int c = 0; c < K2_PRUNE_MAX; ++c) {
                        if (c < nch) {
                            const float dev = sc[wid][2 * c + 1].y;
                            const bool wanted = pass && !(ds[c] > fmaf(dev, rb2 + dev, slack));      // NaN-safe
                            if (__any_sync(0xffffffffu, wanted)) need |= 1u << c;
                        }
                    }
                    // only the chunks somebody needs are fetched: two per pass (16 lanes each), at their own positions
                    for (unsigned mm = need; mm;) {
                        const int ca = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const int cb = mm ? __ffs(mm) - 1 : ca;
                        mm &= mm - 1;
                        const int q = ((lane < 16) ? ca : cb) * K2_PRUNE_CHUNK + (lane & 15);
                        float2 v = make_float2(K2_FAR, K2_FAR);
                        if (q < np) {
                            const double2 P = a.point[o0 + q];
                            v = make_float2((float)(P.x - cxs), (float)(P.y - cys));
                        }
                        sp[wid][q] = v;
                    }
                    __syncwarp();
                    for (unsigned mm = need; mm; mm &= mm - 1) {
                        const int q0 = (__ffs(mm) - 1) * K2_PRUNE_CHUNK;
                        scan_pass(sp[wid], q0, q0 + K2_PRUNE_CHUNK, pxf, pyf, tol, o0, m1, lo, hi);
                    }
                } else {
                    for (int c0 = o0; c0 < o1; c0 += K2_CHUNK) {
                        const int m = min(K2_CHUNK, o1 - c0);
                        stage_points(sp[wid], a.point, c0, m, 4, cxs, cys, lane);
                        scan_pass(sp[wid], 0, (m + 3) & ~3, pxf, pyf, tol, c0, m1, lo, hi);
                    }
                }
                if (pass)
                    // stage 2: numpy's exact arithmetic over the bracket (first index on exact ties, forces.py:154, :228)
                    best_q = (hi < lo) ? exact_argmin(a.point, o0, o1, px, py, best)      // non-finite coordinates: full scan
                                       : exact_argmin(a.point, lo, min(hi + 1, o1), px, py, best);
                }
                if (pass) {
                    const double2 f = segment_force<KIND>(a, px, py, vx, vy, radius, a.point[best_q],
                                                          KIND ? a.velocity[s] : make_double2(0.0, 0.0), best);
                    fx = __dadd_rn(fx, f.x);
                    fy = __dadd_rn(fy, f.y);
                    if (a.emit) {
                        const unsigned long long e = atomicAdd(a.emit_count, 1ull);
                        if ((long long)e < a.emit_capacity) {
                            a.emit[3 * e + 0] = i;
                            a.emit[3 * e + 1] = s;
                            a.emit[3 * e + 2] = best_q - o0;
                        }
                    }
                }
            }
        }
    }
    part[wid][lane] = make_double2(fx, fy);
    __syncthreads();
    if (wid == 0 && active) {
        double sx = part[0][lane].x, sy = part[0][lane].y;
#pragma unroll
        for (int w = 1; w < K2_WARPS; ++w) {
            sx = __dadd_rn(sx, part[w][lane].x);
            sy = __dadd_rn(sy, part[w][lane].y);
        }
        if (KIND == 0) {
            const uint8_t md = a.mode[i];                   // forces.py:176-177
            if (md == SFM_CROSSING_ROAD || md == SFM_ROAD_TO_SIDEWALK) { sx = __dmul_rn(sx, 0.0); sy = __dmul_rn(sy, 0.0); }
        }
        a.f_out[i] = make_double2(sx, sy);
    }
    __syncthreads();                     // part[] and s_group are reused by the next group
    }
}

__global__ void k2_zero(double2* f, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f[i] = make_double2(0.0, 0.0);
}

}  // namespace sfm
