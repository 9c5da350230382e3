// C ABI of sfm_b200 (include/sfm_b200.h): context, device memory, launch sequencing.  sm_100a only, no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "k1_ped_pairs.cuh"
#include "k1_sym.cuh"
#include "k2_cells.cuh"
#include "k3_integrate.cuh"
#include "k4_lifecycle.cuh"
#include "k7_peer.cuh"
#include "k8_order.cuh"

using namespace sfm;

namespace {

thread_local std::string g_error;
unsigned long long g_alloc_epoch = 0;      // bumped by every device (re)allocation: captured graphs bake pointers in

int fail(const std::string& msg) {
    g_error = msg;
    return 1;
}

#define SFM_CUDA(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t err__ = (expr);                                                                         \
        if (err__ != cudaSuccess)                                                                           \
            return fail(std::string(#expr) + ": " + cudaGetErrorString(err__) + " (" __FILE__ ":" +         \
                        std::to_string(__LINE__) + ")");                                                    \
    } while (0)

#define SFM_TRY(expr)            \
    do {                         \
        int rc__ = (expr);       \
        if (rc__ != 0) return rc__; \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (p) SFM_CUDA(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 256;
        SFM_CUDA(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
        g_alloc_epoch += 1;
        return 0;
    }
    // like ensure(), but the first `keep` elements survive a reallocation (device-to-device copy on `stream`)
    int grow_keep(size_t n, size_t keep, cudaStream_t stream) {
        if (n <= cap) return 0;
        T* old = p;
        size_t want = n + n / 4 + 256;
        T* fresh = nullptr;
        SFM_CUDA(cudaMalloc(&fresh, want * sizeof(T)));
        if (old && keep) SFM_CUDA(cudaMemcpyAsync(fresh, old, keep * sizeof(T), cudaMemcpyDeviceToDevice, stream));
        SFM_CUDA(cudaStreamSynchronize(stream));
        if (old) SFM_CUDA(cudaFree(old));
        p = fresh;
        cap = want;
        g_alloc_epoch += 1;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

enum StatClass { ST_PAIRS = 0, ST_CELLS = 1, ST_SEGMENTS = 2, ST_INTEGRATE = 3, ST_LIFECYCLE = 4, ST_COUNT = 5 };

struct TimedSpan {
    cudaEvent_t start, stop;
    int cls;
};

struct SetStorage {
    SegmentSet s;
    DevBuf<double2> center, velocity, point;
    DevBuf<double> cutoff;
    DevBuf<int> offset, cell_start, cell_item, val_tmp;
    DevBuf<unsigned> key, key_tmp;
    DevBuf<int> sort_hist, sort_scan;      // many-CTA sort: [256][CTAs] digit histogram, tile totals of its scan
    DevBuf<float> tol;
    DevBuf<int> chunk_first;
    DevBuf<float4> chunk, chord0;
    double threshold = 0.0;        // perception_threshold the cutoffs / bracket widths / grid of this set were built from
    bool stale = false;            // sfm_set_params changed that threshold afterwards: the set must be uploaded again
    void release() {
        tol.release(); chunk_first.release(); chunk.release(); chord0.release(); sort_hist.release(); sort_scan.release();
        center.release(); velocity.release(); point.release(); cutoff.release(); offset.release();
        cell_start.release(); cell_item.release(); val_tmp.release(); key.release(); key_tmp.release();
    }
};

}  // namespace

struct sfm_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t aux_stream = nullptr;       // cell-list kernels run here, concurrently with the pair kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap = true, join_pending = false;
    sfm_params params{};
    bool have_params = false;
    double ox = 0.0, oy = 0.0, oz = 0.0;
    bool origin_set = false;        // false: oz follows the first pedestrian's z so flat crowds take the planar path
    int world = 1, rank = 0;
    int64_t rows_pad = 0;
    bool partition_fixed = false;
    int64_t n = 0;
    bool staged = false;            // planes_own reflects the master state

    DevBuf<double4> locr, vels;
    DevBuf<double2> wp;
    DevBuf<uint8_t> mode;
    DevBuf<float> planes;           // [world][NPLANES][rows_pad]
    int nsplit = 1;
    DevBuf<double2> f_border, f_static, f_dynamic;
    DevBuf<double> f_total, f_accel, f_ped;
    DevBuf<double> raw_a, raw_b, raw_c, raw_d, raw_e;   // upload / download scratch
    DevBuf<uint8_t> rec_bytes;      // sfm_tick_records: the caller's AoS table as uploaded
    DevBuf<double4> rec_out;        //                   (new velocity, target speed) per pedestrian
    DevBuf<unsigned long long> rec_ident;   //               the table's identity column as adopted (8 bytes per record)
    DevBuf<int> rec_ident_flag;     //                   [1] set by records_identity when a record's entry differs
    int64_t rec_ident_n = -1;       //                   rows rec_ident describes (-1: nothing adopted)
    double4* rec_pinned = nullptr;  //                   pinned landing zone of rec_out
    size_t rec_pinned_cap = 0;
    DevBuf<uint8_t> raw_mode;
    // pedestrian binning
    CellGrid ped_grid{};
    int ped_cells = 0;
    DevBuf<int> perm, ped_start, ped_cursor, ped_scan_tmp;
    int sort_single_max = 8192;     // SFM_SORT_SINGLE_MAX: sets up to this many items are sorted by one CTA (one launch)
    DevBuf<unsigned> ped_cell;
    bool perm_valid = false;
    // point sets
    SetStorage borders, stat, dyn;
    // enumeration scratch
    DevBuf<long long> emit;
    DevBuf<unsigned long long> emit_count;
    DevBuf<unsigned long long> fixup_rows;
    bool fixup_zeroed = false;
    bool k1_local = true;           // pair kernel: tiles whose 64-row runs are compact are read through the runs' origins
    // staged slot order (k8_order.cuh): rows are staged along a Hilbert curve so that those runs ARE compact
    DevBuf<int> slot_of_row, row_of_slot, order_box;
    SetStorage order_sort;          // key / value / scratch buffers of the radix sort
    int64_t order_n = -1, order_rows_pad = -1;     // the (n, rows_pad) the maps were built for; anything else = identity
    int reorder_interval = 0;       // sfm_set_reorder_interval: rebuild the order every this many ticks (0: never)
    int64_t ticks_total = 0;        // ticks stepped by this context (never reset)
    int64_t order_tick = 0;         // ticks_total when the order was last rebuilt
    bool order_due = false;         // new positions were uploaded / the interval was set: rebuild at the next (re)staging
    int64_t reorders = 0;
    // accounting
    bool profiling = false;
    int64_t launches = 0, steps = 0, pair_launches = 0, pair_evals = 0;
    double ms[ST_COUNT] = {0, 0, 0, 0, 0};
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> event_pool;
    int k1_target_ctas = 148 * 4 * 32;     // grid depth of the pair kernel: CTAs of 1-2 partner tiles at cfg3 (profiles/k1_target_ctas_r2_v5.log:
                                           // 3.49 -> 3.43 ms alone against 148 * 4 * 16; flat from 24 to 64 waves)
    bool k1_first = false;
    int k2_persist = 2;             // SFM_K2_PERSIST: CTAs per SM of the persistent cell-list kernels beside the pair kernel (0: one CTA
                                    // per group); profiles/k2_persist_sweep_r1.log: 4.54 -> 4.48 ms per tick, 4.71 -> 4.56 ms through host buffers
    int sm_count = 148;
    bool k2_persist_multi = false;  // SFM_K2_PERSIST_MULTI=1: persistent mode on multi-rank contexts too
    DevBuf<int> k2_counter;
    bool k2_prune = true;           // SFM_K2_PRUNE=0: cell-list kernels scan every point of an item (no chunk bounds)
    bool k2_direct = true;          // SFM_K2_DIRECT=0: no chord-projection windows (every item takes the float32 scan)
    int k1_smem_pad = 0;            // dynamic shared memory added to the pair kernel: caps its CTAs per SM so that the
                                    // cell-list kernels stay co-resident on every SM (see step_begin)
    DevBuf<long long> facc;         // [world * rows_pad][4] fixed-point force accumulators + poison counter
    // ---- lifecycle: mode machines, traffic, routes, device-generated vehicle rings, recorder
    DevBuf<double> mm_speed, mm_initial, mm_crossing, mm_margin, mm_next_time;
    bool have_mm = false;
    double waiting_time = 5.0, sim_time = 0.0;
    DevBuf<double2> tr_center, tr_vel;
    int tr_count = 0;
    double tr_ext0[2] = {0.0, 0.0};
    DevBuf<int> rt_end, rt_cursor;
    DevBuf<double> rt_wp, next_wp3;
    DevBuf<uint8_t> rt_cross, finished;
    bool have_routes = false, routes_fused = false;
    double route_threshold = 2.0;
    DevBuf<unsigned long long> life_counters;      // [0] crossings started [1] idle wake-ups [2] hand-overs [3] finished
    DevBuf<double2> veh_extent;
    DevBuf<double> veh_yaw;
    bool have_vehicles = false;
    double veh_size_factor = 1.4142135623730951;
    DevBuf<double4> rec_xyv;
    DevBuf<uint8_t> rec_mode;
    int64_t rec_capacity = 0, rec_rows = 0, rec_count = 0;
    std::vector<double> rec_times;
    DevBuf<double4> cmp4; DevBuf<double2> cmp2; DevBuf<double3s> cmp3; DevBuf<double> cmp1; DevBuf<int> cmpi;
    DevBuf<uint8_t> cmpb;                          // scratch columns of sfm_despawn_finished
    DevBuf<int> bad_rows;                          // [0] count, [1..] rows the pair-force repair kernel recomputes
    DevBuf<int> check_list;                        // [0] count, [1..] rows waiting at the kerb this tick
    DevBuf<uint8_t> check_blocked;
    int64_t rt_total = 0;                          // waypoints stored in rt_wp / rt_cross
    std::vector<int> rt_begin;                     // host copy of the routes' first entries (cursor downloads are relative)
    // ---- peer-memory exchange (K7): mapped buffers of the other ranks, flag barrier, double-buffered gather buffer
    bool p2p = false;
    int parity = 0;                                  // which half of `planes` the pair kernel reads this tick
    void* peer_planes[MAX_PEERS] = {};               // [rank] base of that rank's `planes` (own entry = own pointer)
    void* peer_facc[MAX_PEERS] = {};
    void* peer_flags[MAX_PEERS] = {};
    DevBuf<unsigned> flags;                          // [MAX_PEERS] epochs signalled by the peers + [MAX_PEERS] error word
    unsigned epoch = 0;
    int64_t barriers = 0;
    long long barrier_timeout_cycles = 20000000000LL;   // ~10 s at 1.9 GHz; SFM_BARRIER_TIMEOUT_MS overrides
    // ---- CUDA graph of one single-rank tick (sfm_step), opt-in (SFM_GRAPH=1): the ~12 launches of a tick replayed as one
    //      graph launch; re-captured whenever a pointer, a size or a parameter a kernel argument was built from may have
    //      changed.  Measured (profiles/small_n_steps_r1.log): 58 -> 55 us per tick at N = 64, 158 -> 145 us at N = 4,096,
    //      738 -> 785 us at N = 16,384 -- the small-crowd tick is bound by the latency of its chain of dependent kernels on
    //      the GPU, not by launch overhead on the host, so replay is not the default.
    bool use_graph = false;
    cudaGraphExec_t graph_exec = nullptr;
    unsigned long long graph_alloc_epoch = 0, config_epoch = 0, graph_config_epoch = 0;
    int graph_integrate = -1, graph_warm = 0;
    cudaStream_t graph_stream = nullptr;
    int64_t graph_launches = 0, graph_pair_launches = 0, graph_pair_evals = 0, graph_replays = 0;
    bool pairs_pending = false;
    bool step_open = false;         // sfm_step_begin done, sfm_step_end outstanding     // symmetric accumulation launched, finish kernel not yet run
};

namespace {

struct SpanGuard {
    sfm_ctx* c;
    int idx = -1;
    cudaStream_t st;
    SpanGuard(sfm_ctx* ctx, int cls, cudaStream_t stream = nullptr) : c(ctx), st(stream ? stream : ctx->stream) {
        // NVTX range per tick phase (host side of the launches; a no-op unless a tool is attached)
        static const char* const names[ST_COUNT] = {"sfm:pairs (K1)", "sfm:cells (K2a)", "sfm:segments (K2b/c)",
                                                    "sfm:integrate (K3)", "sfm:lifecycle (K4-K6)"};
        nvtxRangePushA(names[cls]);
        if (!c->profiling) return;
        TimedSpan sp;
        auto get = [&]() {
            cudaEvent_t e;
            if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        sp.start = get();
        sp.stop = get();
        sp.cls = cls;
        cudaEventRecord(sp.start, st);
        c->spans.push_back(sp);
        idx = (int)c->spans.size() - 1;
    }
    ~SpanGuard() {
        if (idx >= 0) cudaEventRecord(c->spans[idx].stop, st);
        nvtxRangePop();
    }
};

int drain_spans(sfm_ctx* c) {
    if (c->spans.empty()) return 0;
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    for (auto& sp : c->spans) {
        float t = 0.f;
        SFM_CUDA(cudaEventElapsedTime(&t, sp.start, sp.stop));
        c->ms[sp.cls] += t;
        c->event_pool.push_back(sp.start);
        c->event_pool.push_back(sp.stop);
    }
    c->spans.clear();
    return 0;
}

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- small packing kernels (host layout [n][3] float64 <-> device double4) ----------------------------------------
__global__ void pack_state(int64_t n, const double* loc, const double* vel, const double* wp3, const double* radius,
                           const double* speed, const uint8_t* mode_in, double4* locr, double4* vels, double2* wp,
                           uint8_t* mode) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (loc) {
        const double r = radius ? radius[i] : locr[i].w;
        locr[i] = make_double4(loc[3 * i], loc[3 * i + 1], loc[3 * i + 2], r);
    }
    if (vel) {
        const double s = speed ? speed[i] : vels[i].w;
        vels[i] = make_double4(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2], s);
    } else if (speed) {
        vels[i].w = speed[i];
    }
    if (wp3) wp[i] = make_double2(wp3[3 * i], wp3[3 * i + 1]);
    if (mode_in) mode[i] = mode_in[i];
}

__global__ void unpack_state(int64_t n, const double4* locr, const double4* vels, double* loc, double* vel) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (loc) {
        const double4 L = locr[i];
        loc[3 * i] = L.x; loc[3 * i + 1] = L.y; loc[3 * i + 2] = L.z;
    }
    if (vel) {
        const double4 V = vels[i];
        vel[3 * i] = V.x; vel[3 * i + 1] = V.y; vel[3 * i + 2] = V.z;
    }
}

// ---- the reference's AoS pedestrian table (pedestrian_state.py:17-19, 132-byte records) on the device ----------------
// Fields sit at 4-byte-aligned offsets of a 4-byte-aligned stride: float64 values are read as two 32-bit halves.
__device__ __forceinline__ double load_f64_u32(const uint8_t* p) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
    return __hiloint2double((int)q[1], (int)q[0]);
}

struct RecordLayout {
    int64_t stride, off_loc, off_vel, off_wp, off_radius, off_speed;
};

// records -> master state.  take_speed: the target speed comes from the record too (host-managed modes); otherwise the
// device keeps its own (K4a applies the mode machines' speed).
__global__ void unpack_records(int64_t n, const uint8_t* rec, RecordLayout L, int take_speed, double4* locr, double4* vels,
                               double2* wp, double* next_wp3) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* r = rec + i * L.stride;
    locr[i] = make_double4(load_f64_u32(r + L.off_loc), load_f64_u32(r + L.off_loc + 8), load_f64_u32(r + L.off_loc + 16),
                           load_f64_u32(r + L.off_radius));
    const double speed = take_speed ? load_f64_u32(r + L.off_speed) : vels[i].w;
    vels[i] = make_double4(load_f64_u32(r + L.off_vel), load_f64_u32(r + L.off_vel + 8), load_f64_u32(r + L.off_vel + 16), speed);
    const double wx = load_f64_u32(r + L.off_wp), wy = load_f64_u32(r + L.off_wp + 8), wz = load_f64_u32(r + L.off_wp + 16);
    wp[i] = make_double2(wx, wy);
    next_wp3[3 * i] = wx; next_wp3[3 * i + 1] = wy; next_wp3[3 * i + 2] = wz;
}

// The 8 bytes at `offset` of every record (the drop-in's `mode` column: object pointers) adopted as the table's identity,
// or compared with the adopted column: any difference raises the flag.
__global__ void records_identity(int64_t n, const uint8_t* rec, int64_t stride, int64_t offset, int adopt,
                                 unsigned long long* ident, int* flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(rec + i * stride + offset);
    const unsigned long long v = (unsigned long long)q[0] | ((unsigned long long)q[1] << 32);
    if (adopt) ident[i] = v;
    else if (ident[i] != v) *flag = 1;
}

// (new velocity, target speed the clamp used) per pedestrian, one 32-byte item each: the D2H payload of the record tick
__global__ void pack_velocities(int64_t n, const double4* vels, double4* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = vels[i];
}

__global__ void expand_xy(int64_t n, const double2* f, double* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 v = f[i];
    out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = 0.0;
}

int check_ctx(sfm_ctx* c) {
    if (!c) return fail("null context");
    SFM_CUDA(cudaSetDevice(c->device));
    return 0;
}

PairParams make_pair_params(const sfm_moussaid_params& m) {
    const double l2e = 1.4426950408889634;
    PairParams p;
    p.lambda = (float)m.lambda_weight;
    p.eps_gamma = (float)(m.epsilon * m.gamma);
    p.neg_l2e_over_gamma = (float)(-l2e / m.gamma);
    p.c_nprime = (float)(m.n_prime * m.gamma * m.n_prime * m.gamma * l2e);
    p.c_n = (float)(m.n * m.gamma * m.n * m.gamma * l2e);
    p.log2A = (float)std::log2(m.A);
    return p;
}

MoussaidD make_moussaid_d(const sfm_moussaid_params& m) {
    MoussaidD d;
    d.lambda = m.lambda_weight; d.A = m.A; d.gamma = m.gamma; d.n = m.n; d.n_prime = m.n_prime;
    d.epsilon = m.epsilon; d.threshold = m.perception_threshold;
    return d;
}

int ensure_layout(sfm_ctx* c, int64_t n) {
    if (!c->partition_fixed) {
        c->world = 1;
        c->rank = 0;
        c->rows_pad = std::max<int64_t>(ROW_ALIGN, (n + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN);
    } else if (n > c->rows_pad) {
        return fail("row count exceeds the rows_pad given to sfm_set_partition");
    }
    if (c->rows_pad * (int64_t)c->world > (int64_t)1 << 30) return fail("too many staged rows");
    if (c->p2p && (size_t)2 * c->world * NPLANES * c->rows_pad > c->planes.cap)
        return fail("the row layout changed after the peer-memory exchange was set up");
    SFM_TRY(c->planes.ensure((size_t)(c->world > 1 ? 2 : 1) * c->world * NPLANES * c->rows_pad));
    return 0;
}

bool order_valid(const sfm_ctx* c) { return c->n > 0 && c->order_n == c->n && c->order_rows_pad == c->rows_pad; }

// The gather buffer the pair kernel reads this tick, and the one K3 stages the next tick's rows into (the same buffer
// unless the peer-memory exchange alternates between the two halves).
float* planes_cur(sfm_ctx* c) { return c->planes.p + (size_t)c->parity * c->world * NPLANES * c->rows_pad; }
int next_parity(sfm_ctx* c) { return c->p2p ? c->parity ^ 1 : c->parity; }

// `target_parity`: the half of the gather buffer the staged rows go to (K3: the next tick's; k3_stage: either)
StepArgs step_args(sfm_ctx* c, int target_parity = -1) {
    if (target_parity < 0) target_parity = next_parity(c);
    StepArgs a{};
    a.locr = c->locr.p; a.vels = c->vels.p; a.wp = c->wp.p; a.mode = c->mode.p;
    a.n = c->n; a.rows_pad = c->rows_pad;
    a.row_of_slot = order_valid(c) ? c->row_of_slot.p : nullptr;
    a.ped_force = c->f_ped.p;
    a.f_total = c->f_total.p;
    const size_t own_off = (size_t)c->rank * NPLANES * c->rows_pad;
    const size_t half = (size_t)target_parity * c->world * NPLANES * c->rows_pad;
    a.planes_own = c->planes.p + half + own_off;
    if (c->p2p) {                                   // fused all-gather: the same block inside every peer's buffer
        for (int r = 0; r < c->world; ++r)
            if (r != c->rank) a.planes_peer[a.n_peer++] = reinterpret_cast<float*>(c->peer_planes[r]) + half + own_off;
    }
    a.dt = c->params.step_length; a.tau = c->params.tau; a.max_speed_factor = c->params.max_speed_factor;
    a.lambda_ped = c->params.ped.lambda_weight;
    a.ox = c->ox; a.oy = c->oy; a.oz = c->oz;
    return a;
}

ModeMachines mode_machines(sfm_ctx* c) {
    ModeMachines mm;
    mm.mode_speed = c->mm_speed.p; mm.initial_speed = c->mm_initial.p; mm.crossing_speed = c->mm_crossing.p;
    mm.safety_margin = c->mm_margin.p; mm.next_mode_time = c->mm_next_time.p; mm.waiting_time = c->waiting_time;
    return mm;
}

Routes routes_of(sfm_ctx* c) {
    Routes r;
    r.end = c->rt_end.p; r.cursor = c->rt_cursor.p; r.waypoint = c->rt_wp.p; r.crossing = c->rt_cross.p;
    r.next_wp3 = c->next_wp3.p; r.finished = c->finished.p; r.threshold = c->route_threshold;
    r.counters = c->life_counters.p + 2;
    return r;
}

int ensure_life_counters(sfm_ctx* c) {
    if (c->life_counters.p) return 0;
    SFM_TRY(c->life_counters.ensure(4));
    SFM_CUDA(cudaMemsetAsync(c->life_counters.p, 0, 4 * sizeof(unsigned long long), c->stream));
    return 0;
}

int launch_barrier(sfm_ctx* c) {
    PeerPtrs fl{};
    for (int r = 0; r < c->world; ++r) fl.p[r] = c->peer_flags[r];
    c->epoch += 1;
    k7_barrier<<<1, 32, 0, c->stream>>>(fl, c->flags.p, c->world, c->rank, c->epoch, c->flags.p + MAX_PEERS,
                                        c->barrier_timeout_cycles);
    c->launches += 1;
    c->barriers += 1;
    SFM_CUDA(cudaGetLastError());
    return 0;
}

int maybe_reorder(sfm_ctx* c);

// master state -> staged rows of the current tick.  With the peer-memory exchange the rows (pad rows included) go to
// both halves of every rank's gather buffer, and a barrier makes sure everybody's rows have arrived: a collective call.
int launch_stage(sfm_ctx* c) {
    SFM_TRY(maybe_reorder(c));
    {
        SpanGuard g(c, ST_INTEGRATE);
        for (int half = 0; half < (c->p2p ? 2 : 1); ++half) {
            StepArgs a = step_args(c, c->p2p ? half : c->parity);
            k3_stage<<<cdiv(c->rows_pad, 256), 256, 0, c->stream>>>(a);
            c->launches += 1;
        }
        SFM_CUDA(cudaGetLastError());
    }
    c->staged = true;
    if (c->p2p) SFM_TRY(launch_barrier(c));
    return 0;
}

// Device counters of the pair kernels: [0] rows the repair path recomputed, [1] tile pairs that took the local path.
int ensure_pair_counters(sfm_ctx* c) {
    SFM_TRY(c->fixup_rows.ensure(2));
    if (!c->fixup_zeroed) {
        SFM_CUDA(cudaMemsetAsync(c->fixup_rows.p, 0, 2 * sizeof(unsigned long long), c->stream));
        c->fixup_zeroed = true;
    }
    return 0;
}

// Symmetric pair kernel, phase 1: zero the fixed-point accumulators and add every tile pair this rank owns.
int launch_pairs_accumulate(sfm_ctx* c) {
    if (!c->staged && c->p2p) return fail("peer-memory contexts restage collectively: call sfm_stage on every rank first");
    if (!c->staged) SFM_TRY(launch_stage(c));
    const int own_tiles = (int)(c->rows_pad / K1_TJ);
    const int total_tiles = own_tiles * c->world;
    const size_t slots = (size_t)c->world * c->rows_pad;
    SFM_TRY(c->facc.ensure(slots * 4));
    const int half = std::max(1, total_tiles / 2);
    int nsplit = std::max(1, cdiv(c->k1_target_ctas, own_tiles));
    nsplit = std::min(nsplit, half);
    c->nsplit = nsplit;
    SymArgs a{};
    a.planes = planes_cur(c); a.rows_pad = (int)c->rows_pad; a.total_tiles = total_tiles;
    a.own_first_tile = c->rank * own_tiles; a.facc = c->facc.p; a.pp = make_pair_params(c->params.ped);
    SFM_TRY(ensure_pair_counters(c));
    a.use_local = c->k1_local ? 1 : 0; a.local_pairs = c->fixup_rows.p + 1;
    SpanGuard g(c, ST_PAIRS);
    SFM_CUDA(cudaMemsetAsync(c->facc.p, 0, slots * 4 * sizeof(long long), c->stream));
    dim3 grid(own_tiles, nsplit);
    // experiment knob: extra dynamic shared memory caps the pair kernel's CTAs per SM while cell-list kernels are in flight
    const int pad = c->join_pending ? c->k1_smem_pad : 0;
    // SIGN0 variant: np.sign(0) = 0 reproduced in the fast path, needed only when epsilon == 0 (k1_sym.cuh)
    const bool rad = c->params.use_ped_radius != 0, sign0 = a.pp.eps_gamma == 0.0f;
    if (rad && sign0) k1_sym_pairs<true, true><<<grid, KS_THREADS, pad, c->stream>>>(a);
    else if (rad) k1_sym_pairs<true, false><<<grid, KS_THREADS, pad, c->stream>>>(a);
    else if (sign0) k1_sym_pairs<false, true><<<grid, KS_THREADS, pad, c->stream>>>(a);
    else k1_sym_pairs<false, false><<<grid, KS_THREADS, pad, c->stream>>>(a);
    c->launches += 1;
    c->pair_launches += 1;
    for (int t = 0; t < own_tiles; ++t) {        // pair terms this launch evaluates: half shell + the diagonal tile
        const int I = a.own_first_tile + t;
        const int H = (total_tiles & 1) ? (total_tiles - 1) / 2 : ((I < total_tiles / 2) ? total_tiles / 2 : total_tiles / 2 - 1);
        c->pair_evals += (int64_t)(H + 1) * K1_TJ * K1_TJ;
    }
    SFM_CUDA(cudaGetLastError());
    c->pairs_pending = true;
    return 0;
}

// Symmetric pair kernel, phase 2 (after the cross-rank reduce-scatter on multi-GPU runs): fixed point -> float64 force.
int launch_pairs_finish(sfm_ctx* c) {
    if (!c->pairs_pending) return 0;
    SFM_TRY(c->f_ped.ensure((size_t)3 * std::max<int64_t>(c->n, 1)));
    SFM_TRY(ensure_pair_counters(c));
    FinishArgs f{};
    f.planes = planes_cur(c); f.rows_pad = (int)c->rows_pad; f.world = c->world; f.own_block = c->rank;
    f.n_local = (int)c->n; f.facc_own = c->facc.p + (size_t)c->rank * c->rows_pad * 4; f.f_ped = c->f_ped.p;
    f.fixup_rows = c->fixup_rows.p; f.pp = make_pair_params(c->params.ped);
    f.slot_of_row = order_valid(c) ? c->slot_of_row.p : nullptr;
    if (c->p2p)                                     // fused reduce-scatter: this rank's rows inside every peer's accumulator
        for (int r = 0; r < c->world; ++r)
            if (r != c->rank)
                f.facc_peer[f.n_peer++] = reinterpret_cast<const long long*>(c->peer_facc[r]) + (size_t)c->rank * c->rows_pad * 4;
    SFM_TRY(c->bad_rows.ensure(c->n + 1));
    f.bad_count = c->bad_rows.p; f.bad_list = c->bad_rows.p + 1;
    SpanGuard g(c, ST_PAIRS);
    SFM_CUDA(cudaMemsetAsync(f.bad_count, 0, sizeof(int), c->stream));
    k1_sym_finish<<<cdiv(c->n, 256), 256, 0, c->stream>>>(f);
    const int repair_grid = (int)std::min<int64_t>(c->n, (int64_t)c->sm_count * 8);
    if (c->params.use_ped_radius) k1_sym_repair<true><<<repair_grid, KS_REPAIR_THREADS, 0, c->stream>>>(f);
    else k1_sym_repair<false><<<repair_grid, KS_REPAIR_THREADS, 0, c->stream>>>(f);
    c->launches += 2;
    SFM_CUDA(cudaGetLastError());
    c->pairs_pending = false;
    return 0;
}

int launch_pairs(sfm_ctx* c) {
    SFM_TRY(launch_pairs_accumulate(c));
    return launch_pairs_finish(c);
}

// In-place exclusive scan: one CTA for short arrays, tiles + scanned tile totals + add-back beyond that.
int launch_exclusive_scan(sfm_ctx* c, int* data, int n, DevBuf<int>& tmp, cudaStream_t st) {
    if (n <= 4 * SCAN_TILE) {
        k2_exclusive_scan<<<1, 1024, 0, st>>>(data, n);
        c->launches += 1;
    } else {
        const int tiles = cdiv(n, SCAN_TILE);
        SFM_TRY(tmp.ensure(tiles));
        k2_scan_tiles<<<tiles, SCAN_THREADS, 0, st>>>(data, n, tmp.p);
        k2_exclusive_scan<<<1, 1024, 0, st>>>(tmp.p, tiles);
        k2_scan_add<<<cdiv(n, 256), 256, 0, st>>>(data, n, tmp.p);
        c->launches += 3;
    }
    SFM_CUDA(cudaGetLastError());
    return 0;
}

// (key, cell_item) of a set into (cell, index) order: one CTA up to sort_single_max items, the many-CTA sort beyond.
int sort_items(sfm_ctx* c, SetStorage& st, int count, int key_bits) {
    if (count <= c->sort_single_max) {
        k2_radix_sort<<<1, SORT_THREADS, 0, c->stream>>>(st.key.p, st.cell_item.p, st.key_tmp.p, st.val_tmp.p, count, key_bits);
        c->launches += 1;
        SFM_CUDA(cudaGetLastError());
        return 0;
    }
    const int nblk = cdiv(count, MSORT_TILE);
    SFM_TRY(st.sort_hist.ensure((size_t)MSORT_R * nblk));
    unsigned* kin = st.key.p;
    int* vin = st.cell_item.p;
    unsigned* kout = st.key_tmp.p;
    int* vout = st.val_tmp.p;
    for (int shift = 0; shift < key_bits; shift += MSORT_BITS) {
        k2_msort_hist<<<nblk, MSORT_THREADS, 0, c->stream>>>(kin, count, shift, st.sort_hist.p, nblk);
        c->launches += 1;
        SFM_TRY(launch_exclusive_scan(c, st.sort_hist.p, MSORT_R * nblk, st.sort_scan, c->stream));
        k2_msort_scatter<<<nblk, MSORT_THREADS, 0, c->stream>>>(kin, vin, kout, vout, count, shift, st.sort_hist.p, nblk);
        c->launches += 1;
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    SFM_CUDA(cudaGetLastError());
    if (kin != st.key.p) {
        SFM_CUDA(cudaMemcpyAsync(st.key.p, kin, sizeof(unsigned) * count, cudaMemcpyDeviceToDevice, c->stream));
        SFM_CUDA(cudaMemcpyAsync(st.cell_item.p, vin, sizeof(int) * count, cudaMemcpyDeviceToDevice, c->stream));
    }
    return 0;
}

int rebin_peds(sfm_ctx* c, cudaStream_t st = nullptr) {
    if (!st) st = c->stream;
    const int n = (int)c->n;
    SFM_TRY(c->perm.ensure(n));
    SFM_TRY(c->ped_cell.ensure(n));
    SFM_TRY(c->ped_start.ensure(c->ped_cells + 1));
    SFM_TRY(c->ped_cursor.ensure(c->ped_cells + 1));
    SpanGuard g(c, ST_CELLS, st);
    SFM_CUDA(cudaMemsetAsync(c->ped_start.p, 0, sizeof(int) * (c->ped_cells + 1), st));
    SFM_CUDA(cudaMemsetAsync(c->ped_cursor.p, 0, sizeof(int) * (c->ped_cells + 1), st));
    k2_ped_count<<<cdiv(n, 256), 256, 0, st>>>(c->locr.p, n, c->ped_grid, c->ped_cell.p, c->ped_start.p);
    SFM_TRY(launch_exclusive_scan(c, c->ped_start.p, c->ped_cells + 1, c->ped_scan_tmp, st));
    k2_ped_fill<<<cdiv(n, 256), 256, 0, st>>>(c->ped_cell.p, n, c->ped_start.p, c->ped_cursor.p, c->perm.p);
    c->launches += 2;
    SFM_CUDA(cudaGetLastError());
    c->perm_valid = true;
    return 0;
}

// Staged slot order from the current positions (k8_order.cuh): bounding box, Hilbert keys, stable radix sort, the two maps.
// Stream-ordered on the main stream; the caller restages (K3 or k3_stage) afterwards.
int rebuild_order(sfm_ctx* c) {
    const int n = (int)c->n;
    if (n <= 0) return 0;
    SetStorage& st = c->order_sort;
    SFM_TRY(st.key.ensure(n)); SFM_TRY(st.key_tmp.ensure(n)); SFM_TRY(st.cell_item.ensure(n)); SFM_TRY(st.val_tmp.ensure(n));
    SFM_TRY(c->slot_of_row.ensure(n)); SFM_TRY(c->row_of_slot.ensure(c->rows_pad)); SFM_TRY(c->order_box.ensure(4));
    SpanGuard g(c, ST_CELLS);
    k8_bbox_init<<<1, 32, 0, c->stream>>>(c->order_box.p);
    k8_bbox<<<cdiv(n, 256), 256, 0, c->stream>>>(c->locr.p, n, c->order_box.p);
    k8_keys<<<cdiv(n, 256), 256, 0, c->stream>>>(c->locr.p, n, c->order_box.p, st.key.p, st.cell_item.p);
    c->launches += 3;
    SFM_CUDA(cudaGetLastError());
    SFM_TRY(sort_items(c, st, n, 2 * ORDER_BITS));
    k8_fill_order<<<cdiv(c->rows_pad, 256), 256, 0, c->stream>>>(st.cell_item.p, n, (int)c->rows_pad, c->slot_of_row.p,
                                                                 c->row_of_slot.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    c->order_n = c->n; c->order_rows_pad = c->rows_pad;
    c->order_tick = c->ticks_total;
    c->order_due = false;
    c->reorders += 1;
    c->config_epoch += 1;                       // a captured tick graph holds the old maps' pointers (or none)
    return 0;
}

// Called wherever rows are about to be (re)staged: rebuilds the order when it is due.  Crowds of fewer than 8 tiles stay
// in row order -- nearly all of their tile pairs are neighbours, which take the double-single path anyway.
int maybe_reorder(sfm_ctx* c) {
    if (c->reorder_interval <= 0 || c->n <= 0 || c->rows_pad * (int64_t)c->world < 8 * ROW_ALIGN) return 0;
    if (!order_valid(c) || c->order_due || c->ticks_total - c->order_tick >= c->reorder_interval) return rebuild_order(c);
    return 0;
}

// Pedestrian grid from the host-side bounding box at upload time; later positions are clamped into it (the
// permutation only provides locality, never correctness).
void plan_ped_grid(sfm_ctx* c, int64_t n, const double* loc) {
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int64_t i = 0; i < n; ++i) {
        const double x = loc[3 * i], y = loc[3 * i + 1];
        if (x < x0) x0 = x;
        if (x > x1) x1 = x;
        if (y < y0) y0 = y;
        if (y > y1) y1 = y;
    }
    if (n == 0 || !(x1 >= x0) || !(y1 >= y0)) { x0 = y0 = 0.0; x1 = y1 = 1.0; }
    const double w = std::max(x1 - x0, 1.0), h = std::max(y1 - y0, 1.0);
    x0 -= 0.05 * w; y0 -= 0.05 * h;
    const double W = 1.1 * w, H = 1.1 * h;
    double cell = std::max(1.0, std::sqrt(16.0 * W * H / (double)std::max<int64_t>(n, 1)));
    int bits = 1;
    for (;;) {
        const int nx = (int)std::ceil(W / cell), ny = (int)std::ceil(H / cell);
        bits = 1;
        while ((1 << bits) < std::max(nx, ny)) ++bits;
        if (bits <= 9) break;          // at most 512 x 512 cells
        cell *= 2.0;
    }
    c->ped_grid.x0 = x0; c->ped_grid.y0 = y0; c->ped_grid.cell = cell; c->ped_grid.inv_cell = 1.0 / cell;
    c->ped_grid.nx = 1 << bits; c->ped_grid.ny = 1 << bits;
    c->ped_cells = (1 << bits) * (1 << bits);
}

int build_set_grid(sfm_ctx* c, SetStorage& st, int64_t count, const double* centers, double max_cut) {
    SegmentSet& s = st.s;
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int64_t i = 0; i < count; ++i) {
        x0 = std::min(x0, centers[2 * i]); x1 = std::max(x1, centers[2 * i]);
        y0 = std::min(y0, centers[2 * i + 1]); y1 = std::max(y1, centers[2 * i + 1]);
    }
    double cell = std::max(max_cut, 1e-3);
    if (!std::isfinite(cell)) cell = std::max(std::max(x1 - x0, y1 - y0) * 2.0, 1.0);      // everything in one cell
    for (;;) {
        const double nx = std::floor((x1 - x0) / cell) + 1.0, ny = std::floor((y1 - y0) / cell) + 1.0;
        if (nx * ny <= 1048576.0 && nx <= 4096.0 && ny <= 4096.0) break;
        cell *= 2.0;
    }
    s.grid.x0 = x0; s.grid.y0 = y0; s.grid.cell = cell; s.grid.inv_cell = 1.0 / cell;
    s.grid.nx = (int)(std::floor((x1 - x0) / cell) + 1.0);
    s.grid.ny = (int)(std::floor((y1 - y0) / cell) + 1.0);
    const int ncell = s.grid.nx * s.grid.ny;
    SFM_TRY(st.cell_start.ensure(ncell + 2));
    SFM_TRY(st.cell_item.ensure(count));
    SFM_TRY(st.val_tmp.ensure(count));
    SFM_TRY(st.key.ensure(count));
    SFM_TRY(st.key_tmp.ensure(count));
    int key_bits = 1;
    while ((1 << key_bits) < ncell) ++key_bits;
    SpanGuard g(c, ST_CELLS);
    k2_item_keys<<<cdiv(count, 256), 256, 0, c->stream>>>(st.center.p, (int)count, s.grid, st.key.p, st.cell_item.p);
    c->launches += 1;
    SFM_TRY(sort_items(c, st, (int)count, key_bits));
    k2_cell_bounds<<<cdiv(ncell + 1, 256), 256, 0, c->stream>>>(st.key.p, (int)count, ncell, st.cell_start.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    s.cell_start = st.cell_start.p;
    s.cell_item = st.cell_item.p;
    return 0;
}

// Bracket width of the float32 nearest-point stage for an item whose centre-relative coordinates (pedestrians inside
// the cutoff and the item's own points) are bounded by M in every component: 128 * 2^-24 * M^2 (see k2_cells.cuh).
float bracket_tolerance(double M) {
    if (!std::isfinite(M)) return INFINITY;
    const double t = 128.0 * std::ldexp(1.0, -24) * M * M * 1.0001;
    return t < 3.0e38 ? (float)t : INFINITY;
}

int upload_set(sfm_ctx* c, SetStorage& st, int64_t count, const double* centers, const double* cutoffs,
               double uniform_cut, const double* velocities, const int64_t* offsets, const double* points) {
    SegmentSet& s = st.s;
    s.count = 0;
    if (count <= 0) return 0;
    if (!centers || !offsets || !points) return fail("null point-set array");
    const int64_t np = offsets[count];
    if (np > (int64_t)INT32_MAX || count > (int64_t)INT32_MAX) return fail("point set too large");
    for (int64_t i = 0; i < count; ++i)
        if (offsets[i + 1] <= offsets[i]) return fail("every section / obstacle needs at least one point");
    SFM_TRY(st.center.ensure(count));
    SFM_TRY(st.cutoff.ensure(count));
    SFM_TRY(st.velocity.ensure(count));
    SFM_TRY(st.offset.ensure(count + 1));
    SFM_TRY(st.point.ensure(np));
    std::vector<double> cut(count);
    std::vector<int> off(count + 1);
    double max_cut = 0.0;
    for (int64_t i = 0; i < count; ++i) {
        cut[i] = cutoffs ? cutoffs[i] : uniform_cut;
        // an infinite cutoff reaches every pedestrian (forces.py:149-150 accepts it everywhere): one grid cell
        if (!std::isnan(cut[i])) max_cut = std::max(max_cut, cut[i]);
    }
    for (int64_t i = 0; i <= count; ++i) off[i] = (int)offsets[i];
    std::vector<float> tol(count);
    for (int64_t i = 0; i < count; ++i) {
        double M = std::isfinite(cut[i]) ? std::fabs(cut[i]) : INFINITY;
        for (int64_t q = offsets[i]; q < offsets[i + 1]; ++q) {
            M = std::max(M, std::fabs(points[2 * q] - centers[2 * i]));
            M = std::max(M, std::fabs(points[2 * q + 1] - centers[2 * i + 1]));
        }
        tol[i] = bracket_tolerance(M);
    }
    SFM_TRY(st.tol.ensure(count));
    // Pruning chunks (k2_cells.cuh): every run of 16 points of a 48..256-point item as chord a + t u and the largest
    // distance of its points from that chord, centre-relative float32, the deviation rounded up.
    std::vector<int> chunk_first(count, -1);
    std::vector<float4> chunk;
    for (int64_t i = 0; i < count; ++i) {
        const int64_t b = offsets[i], e = offsets[i + 1];
        if (e - b < 48 || e - b > K2_CHUNK || !std::isfinite((double)tol[i])) continue;
        chunk_first[i] = (int)(chunk.size() / 2);
        const double cx = centers[2 * i], cy = centers[2 * i + 1];
        double M = 0.0;
        for (int64_t q = b; q < e; ++q) M = std::max(M, std::max(std::fabs(points[2 * q] - cx), std::fabs(points[2 * q + 1] - cy)));
        for (int64_t first = b; first < e; first += K2_PRUNE_CHUNK) {
            const int64_t last = std::min<int64_t>(first + K2_PRUNE_CHUNK, e) - 1;
            const float ax = (float)(points[2 * first] - cx), ay = (float)(points[2 * first + 1] - cy);
            const float ux = (float)(points[2 * last] - points[2 * first]), uy = (float)(points[2 * last + 1] - points[2 * first + 1]);
            const double uu = (double)ux * ux + (double)uy * uy;
            const float inv = uu > 0.0 ? (float)(1.0 / uu) : 0.0f;
            double dev = 0.0;
            for (int64_t q = first; q <= last; ++q) {           // distance of each point from the (float32-rounded) chord
                const double wx = points[2 * q] - cx - ax, wy = points[2 * q + 1] - cy - ay;
                double t = (wx * ux + wy * uy) * inv;
                t = std::min(1.0, std::max(0.0, t));
                const double rx = wx - t * ux, ry = wy - t * uy;
                dev = std::max(dev, std::sqrt(rx * rx + ry * ry));
            }
            const float devf = std::nextafter((float)(dev * 1.0001 + 8.0 * std::ldexp(1.0, -24) * M), INFINITY);
            chunk.push_back(make_float4(ax, ay, ux, uy));
            chunk.push_back(make_float4(inv, devf, 0.0f, 0.0f));
        }
    }
    // Whole-item chord records of the direct path (k2_cells.cuh): first point a, chord u to the last point, 1/|u|^2, and E >=
    // the largest distance of any point from its uniformly spaced model position a + k/(P-1) u -- measured in float64
    // against the float32-rounded a and u the kernel will use, plus the pedestrian's own float32 rounding (4 * 2^-24 * M).
    std::vector<float4> chord0(2 * (size_t)count);
    for (int64_t i = 0; i < count; ++i) {
        const int64_t b = offsets[i], e = offsets[i + 1], P = e - b;
        const double cx = centers[2 * i], cy = centers[2 * i + 1];
        const float ax = (float)(points[2 * b] - cx), ay = (float)(points[2 * b + 1] - cy);
        const float ux = (float)(points[2 * (e - 1)] - points[2 * b]), uy = (float)(points[2 * (e - 1) + 1] - points[2 * b + 1]);
        const double uu = (double)ux * ux + (double)uy * uy;
        double M = std::isfinite(cut[i]) ? std::fabs(cut[i]) : 0.0, E = 0.0;
        for (int64_t q = b; q < e; ++q) {
            const double rx = points[2 * q] - cx, ry = points[2 * q + 1] - cy;
            M = std::max(M, std::max(std::fabs(rx), std::fabs(ry)));
            const double f = P > 1 ? (double)(q - b) / (double)(P - 1) : 0.0;
            E = std::max(E, std::hypot(rx - ((double)ax + f * ux), ry - ((double)ay + f * uy)));
        }
        const double Em = E * 1.0001 + 8.0 * std::ldexp(1.0, -24) * M;
        float inv = (uu > 0.0 && P > 1) ? (float)(1.0 / uu) : 0.0f;
        float Ef = std::nextafter((float)Em, INFINITY);
        if (!std::isfinite(Em) || !std::isfinite((double)inv) || !std::isfinite(cut[i]) || P > (1 << 22)) { inv = 0.0f; Ef = 0.0f; }
        // the window half-width is at least 2 E (P-1) / |u| index units: beyond 4 the 8-point limit can never be met
        // (closed rings, strongly curved sections) and the kernel does not even try
        if (inv > 0.0f && 2.0 * Em * (double)(P - 1) * std::sqrt((double)inv) > 4.0) inv = 0.0f;
        const float nm1 = (inv > 0.0f) ? (float)(P - 1) : (P == 1 && std::isfinite(Em) ? 0.0f : 1.0f);   // 1 with inv = 0: direct path off
        chord0[2 * i] = make_float4(ax, ay, ux, uy);
        chord0[2 * i + 1] = make_float4(inv, Ef, nm1, nm1 > 0.0f ? 1.0f / nm1 : 0.0f);
    }
    SFM_TRY(st.chord0.ensure(chord0.size()));
    SFM_TRY(st.chunk_first.ensure(count));
    SFM_TRY(st.chunk.ensure(std::max<size_t>(chunk.size(), 1)));
    // synchronous copies from pageable host memory: the inputs may be temporaries of the caller
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    SFM_CUDA(cudaMemcpy(st.center.p, centers, sizeof(double2) * count, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(st.cutoff.p, cut.data(), sizeof(double) * count, cudaMemcpyHostToDevice));
    if (velocities) SFM_CUDA(cudaMemcpy(st.velocity.p, velocities, sizeof(double2) * count, cudaMemcpyHostToDevice));
    else SFM_CUDA(cudaMemset(st.velocity.p, 0, sizeof(double2) * count));
    SFM_CUDA(cudaMemcpy(st.offset.p, off.data(), sizeof(int) * (count + 1), cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(st.point.p, points, sizeof(double2) * np, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(st.tol.p, tol.data(), sizeof(float) * count, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(st.chunk_first.p, chunk_first.data(), sizeof(int) * count, cudaMemcpyHostToDevice));
    if (!chunk.empty())
        SFM_CUDA(cudaMemcpy(st.chunk.p, chunk.data(), sizeof(float4) * chunk.size(), cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(st.chord0.p, chord0.data(), sizeof(float4) * chord0.size(), cudaMemcpyHostToDevice));
    s.chord0 = st.chord0.p;
    s.tol = st.tol.p;
    s.chunk_first = st.chunk_first.p;
    s.chunk = st.chunk.p;
    s.center = st.center.p; s.cutoff = st.cutoff.p; s.velocity = st.velocity.p; s.offset = st.offset.p;
    s.point = st.point.p; s.n_points = np;
    SFM_TRY(build_set_grid(c, st, count, centers, max_cut));
    s.count = count;
    return 0;
}

int launch_segments(sfm_ctx* c, int cls, bool emit, int64_t emit_capacity, cudaStream_t strm = nullptr,
                    bool persistent = false, unsigned long long* eval_count = nullptr) {
    if (!strm) strm = c->stream;
    SetStorage& st = (cls == SFM_FORCE_BORDER) ? c->borders : (cls == SFM_FORCE_STATIC_OBSTACLE ? c->stat : c->dyn);
    DevBuf<double2>& out = (cls == SFM_FORCE_BORDER) ? c->f_border
                                                     : (cls == SFM_FORCE_STATIC_OBSTACLE ? c->f_static : c->f_dynamic);
    const int n = (int)c->n;
    SFM_TRY(out.ensure(n));
    if (st.s.count && st.stale)
        return fail("perception_threshold changed after this obstacle set was uploaded (its cutoffs and cell grid are built "
                    "from it): call sfm_set_obstacles / sfm_set_vehicles again");
    if (st.s.count == 0) {       // forces.py:140-141, :209-210: zeros
        SpanGuard g(c, ST_SEGMENTS, strm);
        k2_zero<<<cdiv(n, 256), 256, 0, strm>>>(out.p, n);
        c->launches += 1;
        SFM_CUDA(cudaGetLastError());
        return 0;
    }
    if (!c->perm_valid) SFM_TRY(rebin_peds(c, strm));
    SegArgs a{};
    a.locr = c->locr.p; a.vels = c->vels.p; a.mode = c->mode.p; a.perm = c->perm.p; a.n = n;
    a.center = st.s.center; a.cutoff = st.s.cutoff; a.velocity = st.s.velocity; a.offset = st.s.offset;
    a.point = st.s.point; a.tol = st.s.tol; a.chunk_first = c->k2_prune ? st.s.chunk_first : nullptr; a.chunk = st.s.chunk;
    a.chord0 = c->k2_direct ? st.s.chord0 : nullptr;
    a.grid = st.s.grid; a.cell_start = st.s.cell_start; a.cell_item = st.s.cell_item;
    a.mp = make_moussaid_d(cls == SFM_FORCE_DYNAMIC_OBSTACLE ? c->params.dynamic_obs : c->params.static_obs);
    a.border_a = c->params.border_a; a.border_b = c->params.border_b; a.neg_inv_border_b = -1.0 / c->params.border_b;
    a.use_radius = c->params.use_ped_radius;
    a.f_out = out.p;
    a.eval_count = eval_count;
    if (emit) {
        a.emit = c->emit.p; a.emit_count = c->emit_count.p; a.emit_capacity = emit_capacity;
    }
    a.n_groups = cdiv(n, 32);
    int grid = a.n_groups;
    SpanGuard g(c, ST_SEGMENTS, strm);
    // (single-rank ticks only: on 2 GPUs the persistent mode shortens the device-timed tick 4.43 -> 4.31 ms but lengthens
    //  the host-buffer tick 4.77 -> 5.50 ms -- measured, not yet understood -- so multi-rank contexts keep one CTA per group)
    if (persistent && c->k2_persist > 0 && (c->world == 1 || c->k2_persist_multi) && a.n_groups > c->k2_persist * c->sm_count) {
        // beside the pair kernel: k2_persist CTAs per SM pulling groups from a counter (see k2_cells.cuh)
        SFM_TRY(c->k2_counter.ensure(4));
        a.work_counter = c->k2_counter.p + (cls - SFM_FORCE_BORDER);
        SFM_CUDA(cudaMemsetAsync(a.work_counter, 0, sizeof(int), strm));
        grid = c->k2_persist * c->sm_count;
    }
    if (cls == SFM_FORCE_BORDER) k2_segments<0><<<grid, K2_THREADS, 0, strm>>>(a);
    else k2_segments<1><<<grid, K2_THREADS, 0, strm>>>(a);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    return 0;
}

int ensure_force_buffers(sfm_ctx* c) {
    SFM_TRY(c->f_total.ensure((size_t)3 * c->n));
    SFM_TRY(c->f_accel.ensure((size_t)3 * c->n));
    SFM_TRY(c->f_ped.ensure((size_t)3 * c->n));
    return 0;
}

// First half of a tick: everything that does not need other ranks' force contributions -- the pair accumulation and
// the three cell-list forces.
int step_begin(sfm_ctx* c) {
    const sfm_params& P = c->params;
    SFM_TRY(ensure_force_buffers(c));
    const bool any_set = (P.enable[SFM_FORCE_BORDER] && c->borders.s.count) ||
                         (P.enable[SFM_FORCE_STATIC_OBSTACLE] && c->stat.s.count) ||
                         (P.enable[SFM_FORCE_DYNAMIC_OBSTACLE] && c->dyn.s.count);
    // The cell-list kernels run on the auxiliary (high-priority) stream, forked from the main stream here and joined
    // before K3; the pair kernel fills the machine behind them.  SFM_K1_FIRST / SFM_AUX_PRIORITY / SFM_K1_SMEM_PAD are
    // the knobs of the co-residency experiments (profiles/overlap_sweep.sh).
    const bool forked = any_set && c->overlap && P.enable[SFM_FORCE_PEDESTRIAN];
    cudaStream_t st2 = forked ? c->aux_stream : c->stream;
    if (forked) {
        SFM_CUDA(cudaEventRecord(c->ev_fork, c->stream));
        SFM_CUDA(cudaStreamWaitEvent(st2, c->ev_fork, 0));
        c->join_pending = true;
    }
    auto pairs = [&]() -> int {
        if (!P.enable[SFM_FORCE_PEDESTRIAN]) return 0;
        return launch_pairs_accumulate(c);
    };
    if (forked && c->k1_first) SFM_TRY(pairs());
    if (any_set && !c->perm_valid) SFM_TRY(rebin_peds(c, st2));
    if (P.enable[SFM_FORCE_BORDER] && c->borders.s.count)
        SFM_TRY(launch_segments(c, SFM_FORCE_BORDER, false, 0, st2, forked));
    if (P.enable[SFM_FORCE_STATIC_OBSTACLE] && c->stat.s.count)
        SFM_TRY(launch_segments(c, SFM_FORCE_STATIC_OBSTACLE, false, 0, st2, forked));
    if (P.enable[SFM_FORCE_DYNAMIC_OBSTACLE] && c->dyn.s.count)
        SFM_TRY(launch_segments(c, SFM_FORCE_DYNAMIC_OBSTACLE, false, 0, st2, forked));
    if (forked) SFM_CUDA(cudaEventRecord(c->ev_join, st2));
    if (!(forked && c->k1_first)) SFM_TRY(pairs());
    c->step_open = true;
    return 0;
}

// Second half: finish the pair force (after the reduce-scatter on multi-GPU runs), then K3.
int step_end(sfm_ctx* c, bool update_velocity, bool integrate_positions, bool keep_class_forces) {
    const sfm_params& P = c->params;
    if (P.enable[SFM_FORCE_PEDESTRIAN]) SFM_TRY(launch_pairs_finish(c));
    if (c->join_pending) {
        SFM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        c->join_pending = false;
    }
    // the order may be rebuilt here: the pair force of this tick has been read out under the old one (finish / repair
    // above), K3 below stages the next tick's rows under the new one.  (Not while a tick is being captured as a graph.)
    if (update_velocity && !c->use_graph) SFM_TRY(maybe_reorder(c));
    StepArgs a = step_args(c);
    if (P.enable[SFM_FORCE_BORDER] && c->borders.s.count) a.f_border = c->f_border.p;
    if (P.enable[SFM_FORCE_STATIC_OBSTACLE] && c->stat.s.count) a.f_static = c->f_static.p;
    if (P.enable[SFM_FORCE_DYNAMIC_OBSTACLE] && c->dyn.s.count) a.f_dynamic = c->f_dynamic.p;
    a.enable_accel = P.enable[SFM_FORCE_ACCELERATION];
    a.enable_ped = P.enable[SFM_FORCE_PEDESTRIAN];
    a.update_velocity = update_velocity;
    a.integrate_positions = integrate_positions;
    if (keep_class_forces) a.f_accel = c->f_accel.p;
    if (update_velocity && c->have_routes && c->routes_fused && c->have_mm) {
        a.advance_routes = 1;
        a.routes = routes_of(c);
        a.mm = mode_machines(c);
        a.wp_rw = c->wp.p;
        a.mode_rw = c->mode.p;
        a.sim_time = c->sim_time;
    }
    {
        SpanGuard g(c, ST_INTEGRATE);
        k3_integrate<<<cdiv(std::max<int64_t>(c->n, 1), 256), 256, 0, c->stream>>>(a);
        c->launches += 1;
        SFM_CUDA(cudaGetLastError());
    }
    if (update_velocity) {
        c->steps += 1;
        c->ticks_total += 1;
        c->perm_valid = false;      // positions moved; rebin before the next segment pass
    }
    c->step_open = false;
    return 0;
}

int step_once(sfm_ctx* c, bool update_velocity, bool integrate_positions, bool keep_class_forces) {
    SFM_TRY(step_begin(c));
    return step_end(c, update_velocity, integrate_positions, keep_class_forces);
}

int download3(sfm_ctx* c, const double* dev, int64_t n, double* out) {
    SFM_CUDA(cudaMemcpyAsync(out, dev, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // namespace

namespace {

template <typename T>
int compact_column(sfm_ctx* c, DevBuf<T>& col, DevBuf<T>& scratch, int64_t n, const uint8_t* finished, const int* new_index) {
    if (!col.p) return 0;
    SFM_TRY(scratch.ensure(n));
    k4_compact<T><<<cdiv(n, 256), 256, 0, c->stream>>>(n, finished, new_index, col.p, scratch.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    std::swap(col.p, scratch.p);
    std::swap(col.cap, scratch.cap);
    return 0;
}

}  // namespace

namespace {

int tick_modes_impl(sfm_ctx* c, double sim_time) {
    if (!c->have_mm) return fail("sfm_set_mode_machines must be called first");
    c->sim_time = sim_time;
    if (c->n == 0) return 0;
    SFM_TRY(ensure_life_counters(c));
    ModeTickArgs a{};
    a.n = c->n; a.locr = c->locr.p; a.vels = c->vels.p; a.wp = c->wp.p; a.mode = c->mode.p;
    a.mm = mode_machines(c);
    if (c->have_vehicles) {                      // device-resident vehicle set (sfm_set_vehicles)
        a.tr.count = (int)c->dyn.s.count; a.tr.center = c->dyn.s.center; a.tr.velocity = c->dyn.s.velocity;
    } else {
        a.tr.count = c->tr_count; a.tr.center = c->tr_center.p; a.tr.velocity = c->tr_vel.p;
    }
    a.tr.ext0_x = c->tr_ext0[0]; a.tr.ext0_y = c->tr_ext0[1];
    a.sim_time = sim_time; a.counters = c->life_counters.p;
    SFM_TRY(c->check_list.ensure(c->n + 1));
    SFM_TRY(c->check_blocked.ensure(c->n));
    a.check_list = c->check_list.p + 1; a.check_count = c->check_list.p; a.blocked = c->check_blocked.p;
    SpanGuard g(c, ST_LIFECYCLE);
    SFM_CUDA(cudaMemsetAsync(a.check_count, 0, sizeof(int), c->stream));
    SFM_CUDA(cudaMemsetAsync(a.blocked, 0, c->n, c->stream));
    k4_tick_modes<<<cdiv(c->n, K4_THREADS), K4_THREADS, 0, c->stream>>>(a);
    c->launches += 1;
    if (a.tr.count > 0) {
        const int gx = (int)std::min<int64_t>(cdiv(c->n, K4_THREADS), 148 * 4);
        k4_gap_acceptance<<<dim3(gx, cdiv(a.tr.count, K4_VEH_TILE)), K4_THREADS, 0, c->stream>>>(a);
        k4_gap_commit<<<gx, K4_THREADS, 0, c->stream>>>(a);
        c->launches += 2;
    }
    SFM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" {

int sfm_abi_version(void) { return SFM_ABI_VERSION; }

const char* sfm_last_error(void) { return g_error.c_str(); }

int sfm_device_count(int* count) {
    if (!count) return fail("null pointer");
    SFM_CUDA(cudaGetDeviceCount(count));
    return 0;
}

int sfm_create(int device, sfm_ctx** out) {
    if (!out) return fail("null pointer");
    *out = nullptr;
    int count = 0;
    SFM_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail("no such CUDA device");
    cudaDeviceProp prop;
    SFM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(std::string("sfm_b200 needs an sm_100 (B200) device, found ") + prop.name + " (sm_" +
                    std::to_string(prop.major) + std::to_string(prop.minor) + "); there is no fallback path");
    SFM_CUDA(cudaSetDevice(device));
    sfm_ctx* c = new sfm_ctx();
    c->device = device;
    int prio_lo = 0, prio_hi = 0;
    SFM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    // Measured on cfg3 (profiles/overlap_sweep_r1*.log): both kernels are issue-slot bound, so co-residency buys nothing
    // -- the best schedule lets the cell-list kernels claim the machine first (high priority, enqueued first, 4.93 ms
    // per tick) and the pair kernel fill in behind them; pair-kernel-first / equal priorities / capped CTAs: 5.24-5.73.
    int own_prio = prio_lo, aux_prio = prio_hi;
    if (const char* env = std::getenv("SFM_AUX_PRIORITY")) {
        if (std::strcmp(env, "low") == 0) { own_prio = prio_hi; aux_prio = prio_lo; }
        if (std::strcmp(env, "equal") == 0) { own_prio = prio_lo; aux_prio = prio_lo; }
    }
    SFM_CUDA(cudaStreamCreateWithPriority(&c->own_stream, cudaStreamNonBlocking, own_prio));
    c->stream = c->own_stream;
    SFM_CUDA(cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, aux_prio));
    c->k1_smem_pad = 0;
    if (const char* env = std::getenv("SFM_K1_SMEM_PAD")) c->k1_smem_pad = std::max(0, std::atoi(env));
    if (const char* env = std::getenv("SFM_K1_FIRST")) c->k1_first = std::atoi(env) != 0;
    if (const char* env = std::getenv("SFM_K2_PRUNE")) c->k2_prune = std::atoi(env) != 0;
    if (const char* env = std::getenv("SFM_K2_DIRECT")) c->k2_direct = std::atoi(env) != 0;
    if (const char* env = std::getenv("SFM_SORT_SINGLE_MAX")) c->sort_single_max = std::max(0, std::atoi(env));
    if (const char* env = std::getenv("SFM_BARRIER_TIMEOUT_MS"))
        c->barrier_timeout_cycles = std::max(1LL, (long long)(std::atof(env) * 1.0e-3 * (double)prop.clockRate * 1.0e3));
    if (const char* env = std::getenv("SFM_GRAPH")) c->use_graph = std::atoi(env) != 0;
    if (const char* env = std::getenv("SFM_K2_PERSIST")) c->k2_persist = std::max(0, std::atoi(env));
    if (const char* env = std::getenv("SFM_K2_PERSIST_MULTI")) c->k2_persist_multi = std::atoi(env) != 0;
    c->sm_count = prop.multiProcessorCount;
    SFM_CUDA(cudaFuncSetAttribute(k1_sym_pairs<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    SFM_CUDA(cudaFuncSetAttribute(k1_sym_pairs<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    SFM_CUDA(cudaFuncSetAttribute(k1_sym_pairs<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    SFM_CUDA(cudaFuncSetAttribute(k1_sym_pairs<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    SFM_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    SFM_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    if (const char* env = std::getenv("SFM_OVERLAP")) c->overlap = std::atoi(env) != 0;
    if (const char* env = std::getenv("SFM_K1_LOCAL")) c->k1_local = std::atoi(env) != 0;
    if (const char* env = std::getenv("SFM_K1_TARGET_CTAS")) c->k1_target_ctas = std::max(1, std::atoi(env));
    else c->k1_target_ctas = prop.multiProcessorCount * 4 * 32;
    *out = c;
    return 0;
}

int sfm_destroy(sfm_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& sp : c->spans) { cudaEventDestroy(sp.start); cudaEventDestroy(sp.stop); }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    c->locr.release(); c->vels.release(); c->wp.release(); c->mode.release(); c->planes.release();
    c->f_border.release(); c->f_static.release(); c->f_dynamic.release();
    c->f_total.release(); c->f_accel.release(); c->f_ped.release();
    c->raw_a.release(); c->raw_b.release(); c->raw_c.release(); c->raw_d.release(); c->raw_e.release();
    c->rec_bytes.release(); c->rec_out.release();
    if (c->rec_pinned) cudaFreeHost(c->rec_pinned);
    c->ped_scan_tmp.release();
    c->raw_mode.release(); c->perm.release(); c->ped_start.release(); c->ped_cursor.release(); c->ped_cell.release();
    c->order_sort.release(); c->slot_of_row.release(); c->row_of_slot.release(); c->order_box.release();
    c->borders.release(); c->stat.release(); c->dyn.release(); c->emit.release(); c->emit_count.release(); c->fixup_rows.release(); c->facc.release();
    if (c->p2p)
        for (int r = 0; r < c->world; ++r)
            if (r != c->rank) {
                cudaIpcCloseMemHandle(c->peer_planes[r]); cudaIpcCloseMemHandle(c->peer_facc[r]);
                cudaIpcCloseMemHandle(c->peer_flags[r]);
            }
    c->flags.release(); c->bad_rows.release(); c->check_list.release(); c->check_blocked.release(); c->k2_counter.release();
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    c->cmp4.release(); c->cmp2.release(); c->cmp3.release(); c->cmp1.release(); c->cmpi.release(); c->cmpb.release();
    c->mm_speed.release(); c->mm_initial.release(); c->mm_crossing.release(); c->mm_margin.release();
    c->mm_next_time.release(); c->tr_center.release(); c->tr_vel.release(); c->rt_end.release(); c->rt_cursor.release();
    c->rt_wp.release(); c->next_wp3.release(); c->rt_cross.release(); c->finished.release(); c->life_counters.release();
    c->veh_extent.release(); c->veh_yaw.release(); c->rec_xyv.release(); c->rec_mode.release();
    cudaStreamDestroy(c->own_stream);
    cudaStreamDestroy(c->aux_stream);
    cudaEventDestroy(c->ev_fork);
    cudaEventDestroy(c->ev_join);
    delete c;
    return 0;
}

int sfm_set_stream(sfm_ctx* c, void* cuda_stream) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return 0;
}

int sfm_synchronize(sfm_ctx* c) {
    SFM_TRY(check_ctx(c));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int sfm_set_params(sfm_ctx* c, const sfm_params* p) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    if (!p) return fail("null params");
    if (!(p->step_length > 0.0)) return fail("step_length must be positive");
    if (!(p->tau > 0.0)) return fail("tau must be positive");
    const sfm_moussaid_params* sets[3] = {&p->ped, &p->static_obs, &p->dynamic_obs};
    for (auto* m : sets)
        if (!(m->gamma > 0.0) || !(m->A > 0.0)) return fail("Moussaid parameters gamma and A must be positive");
    if (!(p->border_b != 0.0)) return fail("border_force.b must be non-zero");
    const bool lambda_changed = !c->have_params || c->params.ped.lambda_weight != p->ped.lambda_weight;
    // the obstacle sets bake the perception threshold into their cutoffs, bracket widths and cell grid (upload_set)
    if (c->stat.s.count && p->static_obs.perception_threshold != c->stat.threshold) c->stat.stale = true;
    if (c->dyn.s.count && p->dynamic_obs.perception_threshold != c->dyn.threshold) c->dyn.stale = true;
    if (c->stat.s.count && p->static_obs.perception_threshold == c->stat.threshold) c->stat.stale = false;
    if (c->dyn.s.count && p->dynamic_obs.perception_threshold == c->dyn.threshold) c->dyn.stale = false;
    c->params = *p;
    c->have_params = true;
    if (lambda_changed) c->staged = false;
    return 0;
}

int sfm_set_origin(sfm_ctx* c, double ox, double oy, double oz) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    c->ox = ox; c->oy = oy; c->oz = oz;
    c->origin_set = true;
    c->staged = false;
    return 0;
}

int sfm_set_partition(sfm_ctx* c, int world, int rank, int64_t rows_pad) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    if (world < 1 || rank < 0 || rank >= world) return fail("bad world / rank");
    if (rows_pad <= 0 || rows_pad % ROW_ALIGN != 0) return fail("rows_pad must be a positive multiple of 256");
    c->world = world; c->rank = rank; c->rows_pad = rows_pad;
    c->partition_fixed = true;
    c->staged = false;
    return 0;
}

int sfm_upload_state(sfm_ctx* c, int64_t n, const double* loc, const double* vel, const double* wp3,
                     const double* radius, const double* speed, const uint8_t* mode) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    if (!c->have_params) return fail("sfm_set_params must be called first");
    if (n < 0 || n > (int64_t)1 << 28) return fail("bad row count");
    if (n > 0 && (!loc || !vel || !wp3 || !radius || !speed || !mode)) return fail("null state array");
    SFM_TRY(ensure_layout(c, n));
    c->n = n;
    c->rec_ident_n = -1;           // a new table: sfm_tick_records must adopt its identity column again
    c->staged = false;
    c->perm_valid = false;
    c->order_due = true;                       // new positions: the staged order is due
    if (n == 0) return 0;
    SFM_TRY(c->locr.ensure(n)); SFM_TRY(c->vels.ensure(n)); SFM_TRY(c->wp.ensure(n)); SFM_TRY(c->mode.ensure(n));
    SFM_TRY(c->raw_a.ensure(3 * n)); SFM_TRY(c->raw_b.ensure(3 * n)); SFM_TRY(c->raw_c.ensure(3 * n));
    SFM_TRY(c->raw_d.ensure(n)); SFM_TRY(c->raw_e.ensure(n)); SFM_TRY(c->raw_mode.ensure(n));
    SFM_CUDA(cudaMemcpyAsync(c->raw_a.p, loc, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    SFM_CUDA(cudaMemcpyAsync(c->raw_b.p, vel, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    SFM_CUDA(cudaMemcpyAsync(c->raw_c.p, wp3, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    SFM_CUDA(cudaMemcpyAsync(c->raw_d.p, radius, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    SFM_CUDA(cudaMemcpyAsync(c->raw_e.p, speed, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    SFM_CUDA(cudaMemcpyAsync(c->raw_mode.p, mode, n, cudaMemcpyHostToDevice, c->stream));
    pack_state<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->raw_a.p, c->raw_b.p, c->raw_c.p, c->raw_d.p, c->raw_e.p,
                                                    c->raw_mode.p, c->locr.p, c->vels.p, c->wp.p, c->mode.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    SFM_TRY(c->next_wp3.ensure(3 * n));
    SFM_CUDA(cudaMemcpyAsync(c->next_wp3.p, c->raw_c.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice, c->stream));
    c->have_mm = c->have_routes = false;             // per-row tables belong to the previous row set
    c->rec_capacity = c->rec_count = 0;
    plan_ped_grid(c, n, loc);
    if (!c->origin_set && c->world == 1) c->oz = loc[2];
    SFM_CUDA(cudaStreamSynchronize(c->stream));      // host arrays may be released by the caller on return
    return 0;
}

int sfm_update_kinematics(sfm_ctx* c, int64_t n, const double* loc, const double* vel) {
    SFM_TRY(check_ctx(c));
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (n == 0) return 0;
    if (!loc || !vel) return fail("null state array");
    SFM_CUDA(cudaMemcpyAsync(c->raw_a.p, loc, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    SFM_CUDA(cudaMemcpyAsync(c->raw_b.p, vel, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    pack_state<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->raw_a.p, c->raw_b.p, nullptr, nullptr, nullptr, nullptr,
                                                    c->locr.p, c->vels.p, c->wp.p, c->mode.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    c->staged = false;
    c->perm_valid = false;
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int sfm_update_targets(sfm_ctx* c, int64_t n, const double* wp3, const double* speed, const uint8_t* mode) {
    SFM_TRY(check_ctx(c));
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (n == 0) return 0;
    if (wp3) SFM_CUDA(cudaMemcpyAsync(c->raw_c.p, wp3, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    if (speed) SFM_CUDA(cudaMemcpyAsync(c->raw_e.p, speed, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    if (mode) SFM_CUDA(cudaMemcpyAsync(c->raw_mode.p, mode, n, cudaMemcpyHostToDevice, c->stream));
    pack_state<<<cdiv(n, 256), 256, 0, c->stream>>>(n, nullptr, nullptr, wp3 ? c->raw_c.p : nullptr, nullptr,
                                                    speed ? c->raw_e.p : nullptr, mode ? c->raw_mode.p : nullptr,
                                                    c->locr.p, c->vels.p, c->wp.p, c->mode.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    if (wp3 && c->next_wp3.p)
        SFM_CUDA(cudaMemcpyAsync(c->next_wp3.p, c->raw_c.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int sfm_download_state(sfm_ctx* c, int64_t n, double* loc, double* vel) {
    SFM_TRY(check_ctx(c));
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (n == 0) return 0;
    unpack_state<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->locr.p, c->vels.p, loc ? c->raw_a.p : nullptr,
                                                      vel ? c->raw_b.p : nullptr);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    if (loc) SFM_CUDA(cudaMemcpyAsync(loc, c->raw_a.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    if (vel) SFM_CUDA(cudaMemcpyAsync(vel, c->raw_b.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int sfm_set_borders(sfm_ctx* c, int64_t n_sections, const double* center, const double* length, const int64_t* offsets,
                    const double* points) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    if (n_sections > 0 && !length) return fail("null section_length");
    return upload_set(c, c->borders, n_sections, center, length, 0.0, nullptr, offsets, points);
}

int sfm_set_obstacles(sfm_ctx* c, int which, int64_t n_obstacles, const double* centers, const double* velocities,
                      const int64_t* offsets, const double* points) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    if (!c->have_params) return fail("sfm_set_params must be called first");
    if (which != SFM_FORCE_STATIC_OBSTACLE && which != SFM_FORCE_DYNAMIC_OBSTACLE) return fail("bad obstacle class");
    const bool dynamic = which == SFM_FORCE_DYNAMIC_OBSTACLE;
    if (dynamic) c->have_vehicles = false;       // host-provided rings replace a device-generated vehicle set
    const double thr = dynamic ? c->params.dynamic_obs.perception_threshold : c->params.static_obs.perception_threshold;
    SetStorage& st = dynamic ? c->dyn : c->stat;
    st.threshold = thr;
    st.stale = false;
    return upload_set(c, st, n_obstacles, centers, nullptr, thr, velocities, offsets, points);
}

int sfm_force(sfm_ctx* c, int cls, int64_t n, double* out) {
    SFM_TRY(check_ctx(c));
    if (!c->have_params) return fail("sfm_set_params must be called first");
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (cls < 0 || cls >= SFM_FORCE_COUNT) return fail("bad force class");
    if (n == 0) return 0;
    if (!out) return fail("null output");
    SFM_TRY(ensure_force_buffers(c));
    if (cls == SFM_FORCE_ACCELERATION || cls == SFM_FORCE_PEDESTRIAN) {
        if (cls == SFM_FORCE_PEDESTRIAN && c->world > 1)
            return fail("the per-class pedestrian force of a multi-rank context needs the reduce-scatter; use the step API");
        if (cls == SFM_FORCE_PEDESTRIAN) SFM_TRY(launch_pairs(c));
        StepArgs a = step_args(c);
        a.enable_accel = cls == SFM_FORCE_ACCELERATION;
        a.enable_ped = cls == SFM_FORCE_PEDESTRIAN;
        a.update_velocity = 0;
        a.f_accel = c->f_accel.p;
        {
            SpanGuard g(c, ST_INTEGRATE);
            k3_integrate<<<cdiv(n, 256), 256, 0, c->stream>>>(a);
            c->launches += 1;
            SFM_CUDA(cudaGetLastError());
        }
        return download3(c, cls == SFM_FORCE_ACCELERATION ? c->f_accel.p : c->f_ped.p, n, out);
    }
    SFM_TRY(launch_segments(c, cls, false, 0));
    DevBuf<double2>& f = (cls == SFM_FORCE_BORDER) ? c->f_border
                                                   : (cls == SFM_FORCE_STATIC_OBSTACLE ? c->f_static : c->f_dynamic);
    SFM_TRY(c->raw_a.ensure(3 * n));
    expand_xy<<<cdiv(n, 256), 256, 0, c->stream>>>(n, f.p, c->raw_a.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    return download3(c, c->raw_a.p, n, out);
}

int sfm_enumerate_pairs(sfm_ctx* c, int cls, int64_t capacity, int64_t* triplets, int64_t* count) {
    SFM_TRY(check_ctx(c));
    if (cls != SFM_FORCE_BORDER && cls != SFM_FORCE_STATIC_OBSTACLE && cls != SFM_FORCE_DYNAMIC_OBSTACLE)
        return fail("only the cutoff-limited classes enumerate pairs");
    if (!count || capacity < 0 || (capacity > 0 && !triplets)) return fail("bad enumeration buffer");
    *count = 0;
    if (c->n == 0) return 0;
    SFM_TRY(c->emit.ensure((size_t)3 * std::max<int64_t>(capacity, 1)));
    SFM_TRY(c->emit_count.ensure(2));
    SFM_CUDA(cudaMemsetAsync(c->emit_count.p, 0, sizeof(unsigned long long), c->stream));
    SFM_TRY(launch_segments(c, cls, true, capacity));
    unsigned long long total = 0;
    SFM_CUDA(cudaMemcpyAsync(&total, c->emit_count.p, sizeof(total), cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    *count = (int64_t)total;
    const int64_t got = std::min<int64_t>((int64_t)total, capacity);
    if (got > 0) SFM_CUDA(cudaMemcpy(triplets, c->emit.p, sizeof(long long) * 3 * got, cudaMemcpyDeviceToHost));
    return 0;
}

int sfm_count_point_evaluations(sfm_ctx* c, int cls, int64_t* pairs, int64_t* point_evaluations) {
    SFM_TRY(check_ctx(c));
    if (cls != SFM_FORCE_BORDER && cls != SFM_FORCE_STATIC_OBSTACLE && cls != SFM_FORCE_DYNAMIC_OBSTACLE)
        return fail("only the cutoff-limited classes evaluate point sets");
    if (!pairs || !point_evaluations) return fail("null pointer");
    *pairs = *point_evaluations = 0;
    if (c->n == 0) return 0;
    SFM_TRY(c->emit_count.ensure(2));
    SFM_CUDA(cudaMemsetAsync(c->emit_count.p, 0, 2 * sizeof(unsigned long long), c->stream));
    SFM_TRY(launch_segments(c, cls, false, 0, nullptr, false, c->emit_count.p));
    unsigned long long v[2] = {0, 0};
    SFM_CUDA(cudaMemcpyAsync(v, c->emit_count.p, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    *pairs = (int64_t)v[0];
    *point_evaluations = (int64_t)v[1];
    return 0;
}

int sfm_step(sfm_ctx* c, int n_steps, int integrate_positions) {
    SFM_TRY(check_ctx(c));
    if (!c->have_params) return fail("sfm_set_params must be called first");
    if (n_steps < 0) return fail("negative step count");
    if (c->n == 0) return 0;                       // pedestrian_simulation.py:60 early-out
    if (c->world > 1)
        return fail("multi-rank contexts step with sfm_step_begin / reduce-scatter / sfm_step_end / all-gather");
    const bool integrate = integrate_positions != 0;
    for (int s = 0; s < n_steps; ++s) {
        const bool graphable = c->use_graph && !c->profiling && c->staged && !(c->have_routes && c->routes_fused);
        if (!graphable) {
            SFM_TRY(step_once(c, true, integrate, false));
            c->graph_warm = 0;
            continue;
        }
        const bool fresh = c->graph_exec && c->graph_alloc_epoch == g_alloc_epoch && c->graph_config_epoch == c->config_epoch &&
                           c->graph_integrate == (int)integrate && c->graph_stream == c->stream;
        if (fresh) {
            SFM_CUDA(cudaGraphLaunch(c->graph_exec, c->stream));
            c->launches += c->graph_launches; c->pair_launches += c->graph_pair_launches;
            c->pair_evals += c->graph_pair_evals;
            c->steps += 1; c->ticks_total += 1; c->graph_replays += 1;
            c->perm_valid = false;
            continue;
        }
        if (c->graph_warm < 2) {                    // two eager ticks first: every buffer a tick touches exists afterwards
            const unsigned long long before = g_alloc_epoch;
            SFM_TRY(step_once(c, true, integrate, false));
            c->graph_warm = (before == g_alloc_epoch) ? c->graph_warm + 1 : 0;
            continue;
        }
        // capture one tick (the auxiliary stream joins the capture through the fork / join events)
        if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
        const int64_t l0 = c->launches, p0 = c->pair_launches, e0 = c->pair_evals, st0 = c->steps;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            const int rc = step_once(c, true, integrate, false);
            const cudaError_t end = cudaStreamEndCapture(c->stream, &graph);
            ok = rc == 0 && end == cudaSuccess && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&c->graph_exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        c->graph_launches = c->launches - l0; c->graph_pair_launches = c->pair_launches - p0;
        c->graph_pair_evals = c->pair_evals - e0;
        c->launches = l0; c->pair_launches = p0; c->pair_evals = e0; c->steps = st0;   // nothing ran during the capture
        if (!ok) {                                   // capture not possible here: stay eager for good
            cudaGetLastError();
            c->use_graph = false;
            c->graph_exec = nullptr;
            c->join_pending = false;
            c->step_open = false;
            SFM_TRY(step_once(c, true, integrate, false));
            continue;
        }
        c->graph_alloc_epoch = g_alloc_epoch; c->graph_config_epoch = c->config_epoch;
        c->graph_integrate = (int)integrate; c->graph_stream = c->stream;
        --s;                                         // the tick itself: replay the fresh graph
    }
    return 0;
}

int sfm_tick_host(sfm_ctx* c, int64_t n, const double* loc, const double* vel, double* new_vel, double* new_loc) {
    SFM_TRY(check_ctx(c));
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (n == 0) return 0;
    if (!loc || !vel || !new_vel) return fail("null array");
    SFM_CUDA(cudaMemcpyAsync(c->raw_a.p, loc, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    SFM_CUDA(cudaMemcpyAsync(c->raw_b.p, vel, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    pack_state<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->raw_a.p, c->raw_b.p, nullptr, nullptr, nullptr, nullptr,
                                                    c->locr.p, c->vels.p, c->wp.p, c->mode.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    c->staged = false;
    c->perm_valid = false;
    SFM_TRY(step_once(c, true, new_loc != nullptr, false));
    unpack_state<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->locr.p, c->vels.p, new_loc ? c->raw_a.p : nullptr,
                                                      c->raw_b.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    if (new_loc)
        SFM_CUDA(cudaMemcpyAsync(new_loc, c->raw_a.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaMemcpyAsync(new_vel, c->raw_b.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int sfm_tick_records(sfm_ctx* c, int64_t n, void* records, int64_t stride, const int64_t* offsets5, double sim_time,
                     int tick_modes, int64_t* counters4, int64_t identity_offset, int identity_mode,
                     int* identity_changed) {
    SFM_TRY(check_ctx(c));
    if (!c->have_params) return fail("sfm_set_params must be called first");
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (c->world > 1) return fail("the record tick serves the single-process drop-in (one rank)");
    if (n == 0) return 0;
    if (!records || !offsets5) return fail("null pointer");
    RecordLayout L{stride, offsets5[0], offsets5[1], offsets5[2], offsets5[3], offsets5[4]};
    const int64_t offs[5] = {L.off_loc, L.off_vel, L.off_wp, L.off_radius, L.off_speed};
    const int64_t widths[5] = {24, 24, 24, 8, 8};
    if (stride <= 0 || stride % 4 != 0 || (reinterpret_cast<uintptr_t>(records) & 3u) != 0)
        return fail("records must be 4-byte aligned with a positive stride that is a multiple of 4");
    for (int k = 0; k < 5; ++k)
        if (offs[k] < 0 || offs[k] % 4 != 0 || offs[k] + widths[k] > stride) return fail("bad field offset");
    if (tick_modes && !c->have_mm) return fail("sfm_set_mode_machines must be called first");
    if (identity_changed) *identity_changed = 0;
    if (identity_mode < 0 || identity_mode > 2) return fail("identity_mode must be 0 (off), 1 (adopt) or 2 (check)");
    if (identity_mode != 0 && (identity_offset < 0 || identity_offset % 4 != 0 || identity_offset + 8 > stride))
        return fail("bad identity offset");
    if (identity_mode == 2 && !identity_changed) return fail("identity check needs identity_changed");
    const size_t span = (size_t)(n - 1) * stride + (size_t)std::max<int64_t>(std::max<int64_t>(offs[0] + 24, std::max<int64_t>(offs[1] + 24,
                            std::max<int64_t>(offs[2] + 24, std::max<int64_t>(offs[3] + 8, offs[4] + 8)))),
                            identity_mode != 0 ? identity_offset + 8 : 0);
    SFM_TRY(c->rec_bytes.ensure(span + 8));
    SFM_TRY(c->rec_out.ensure(n));
    // large tables take their results back by 2-D DMA straight into the records (profiles/microbench/copy2d.cu: 0.26 ms for
    // both columns at N = 65,536 against 0.45 ms for a packed copy + a host scatter loop); small ones keep the packed copy
    const bool dma2d = n >= 4096;
    if (!dma2d && c->rec_pinned_cap < (size_t)n) {
        if (c->rec_pinned) SFM_CUDA(cudaFreeHost(c->rec_pinned));
        c->rec_pinned = nullptr;
        c->rec_pinned_cap = 0;
        SFM_CUDA(cudaMallocHost(&c->rec_pinned, sizeof(double4) * ((size_t)n + n / 8 + 256)));
        c->rec_pinned_cap = (size_t)n + n / 8 + 256;
    }
    // 1. the table as it lies in host memory (one copy; the unused columns ride along), unpacked on the device
    SFM_CUDA(cudaMemcpyAsync(c->rec_bytes.p, records, span, cudaMemcpyHostToDevice, c->stream));
    // 1b. identity column (the drop-in's object pointers): adopted, or compared with the adopted one.  A table that holds
    //     other objects than the resident mode machines describe is handed back untouched: the only mid-tick
    //     synchronisation, ~0.05 ms, where the host-side strided comparison it replaces took 0.25 ms at N = 65,536
    if (identity_mode != 0) {
        SFM_TRY(c->rec_ident.ensure(n));
        SFM_TRY(c->rec_ident_flag.ensure(1));
        if (identity_mode == 2 && c->rec_ident_n != n) {
            *identity_changed = 1;
            SFM_CUDA(cudaStreamSynchronize(c->stream));
            return 0;
        }
        if (identity_mode == 2) SFM_CUDA(cudaMemsetAsync(c->rec_ident_flag.p, 0, sizeof(int), c->stream));
        records_identity<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->rec_bytes.p, stride, identity_offset, identity_mode == 1,
                                                           c->rec_ident.p, c->rec_ident_flag.p);
        c->launches += 1;
        SFM_CUDA(cudaGetLastError());
        if (identity_mode == 1) {
            c->rec_ident_n = n;
        } else {
            int changed = 0;
            SFM_CUDA(cudaMemcpyAsync(&changed, c->rec_ident_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            SFM_CUDA(cudaStreamSynchronize(c->stream));
            if (changed) {
                *identity_changed = 1;
                c->rec_ident_n = -1;
                return 0;
            }
        }
    }
    unpack_records<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->rec_bytes.p, L, tick_modes ? 0 : 1, c->locr.p, c->vels.p,
                                                         c->wp.p, c->next_wp3.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    c->staged = false;
    c->perm_valid = false;
    // 2. pedestrian_simulation.py:63-73 on the device: apply_current_mode, the machines' tick, gap acceptance
    if (tick_modes) SFM_TRY(tick_modes_impl(c, sim_time));
    // 3. pedestrian_simulation.py:81-83: forces, sum, new velocities (positions are the simulator's business)
    SFM_TRY(step_once(c, true, false, false));
    // 4. new velocities (+ the target speed of this tick) back into the table: state[['id','vel']] is a view of it
    pack_velocities<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->vels.p, c->rec_out.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    uint8_t* base = static_cast<uint8_t*>(records);
    if (dma2d) {
        SFM_CUDA(cudaMemcpy2DAsync(base + L.off_vel, (size_t)stride, c->rec_out.p, sizeof(double4), 24, (size_t)n,
                                   cudaMemcpyDeviceToHost, c->stream));
        if (tick_modes)                                                  // apply_current_mode (pedestrian_state.py:94-95)
            SFM_CUDA(cudaMemcpy2DAsync(base + L.off_speed, (size_t)stride, reinterpret_cast<uint8_t*>(c->rec_out.p) + 24,
                                       sizeof(double4), 8, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    } else {
        SFM_CUDA(cudaMemcpyAsync(c->rec_pinned, c->rec_out.p, sizeof(double4) * n, cudaMemcpyDeviceToHost, c->stream));
    }
    unsigned long long cnt[4] = {0, 0, 0, 0};
    if (counters4 && c->life_counters.p)
        SFM_CUDA(cudaMemcpyAsync(cnt, c->life_counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    if (!dma2d) {
        for (int64_t i = 0; i < n; ++i) {
            const double4 v = c->rec_pinned[i];
            uint8_t* r = base + i * stride;
            std::memcpy(r + L.off_vel, &v.x, 24);
            if (tick_modes) std::memcpy(r + L.off_speed, &v.w, 8);
        }
    }
    if (counters4) for (int k = 0; k < 4; ++k) counters4[k] = (int64_t)cnt[k];
    return 0;
}

int sfm_host_register(void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return fail("bad host range");
    SFM_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return 0;
}

int sfm_host_unregister(void* ptr) {
    if (!ptr) return fail("null pointer");
    SFM_CUDA(cudaHostUnregister(ptr));
    return 0;
}

int sfm_host_column_gather(const void* records, int64_t stride, int64_t offset, int64_t width, int64_t n, void* packed) {
    if (n < 0 || width <= 0 || (n > 0 && (!records || !packed))) return fail("bad column description");
    const uint8_t* base = static_cast<const uint8_t*>(records) + offset;
    uint8_t* out = static_cast<uint8_t*>(packed);
    for (int64_t i = 0; i < n; ++i) std::memcpy(out + i * width, base + i * stride, (size_t)width);
    return 0;
}

int sfm_host_column_equal(const void* records, int64_t stride, int64_t offset, int64_t width, int64_t n,
                          const void* packed, int* equal) {
    if (!equal || n < 0 || width <= 0 || (n > 0 && (!records || !packed))) return fail("bad column description");
    const uint8_t* base = static_cast<const uint8_t*>(records) + offset;
    const uint8_t* want = static_cast<const uint8_t*>(packed);
    *equal = 1;
    for (int64_t i = 0; i < n; ++i)
        if (std::memcmp(base + i * stride, want + i * width, (size_t)width) != 0) { *equal = 0; break; }
    return 0;
}

int sfm_apply_force(sfm_ctx* c, int64_t n, const double* force, double* new_vel) {
    SFM_TRY(check_ctx(c));
    if (!c->have_params) return fail("sfm_set_params must be called first");
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (n == 0) return 0;
    if (!force || !new_vel) return fail("null array");
    SFM_CUDA(cudaMemcpyAsync(c->raw_a.p, force, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    {
        SpanGuard g(c, ST_INTEGRATE);
        k3_apply_force<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->vels.p, c->raw_a.p, c->params.step_length,
                                                             c->params.max_speed_factor);
        unpack_state<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->locr.p, c->vels.p, nullptr, c->raw_b.p);
        c->launches += 2;
        SFM_CUDA(cudaGetLastError());
    }
    c->staged = false;
    SFM_CUDA(cudaMemcpyAsync(new_vel, c->raw_b.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int sfm_download_force(sfm_ctx* c, int64_t n, double* out) {
    SFM_TRY(check_ctx(c));
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (n == 0) return 0;
    if (!out || !c->f_total.p) return fail("no force available");
    return download3(c, c->f_total.p, n, out);
}

int sfm_download_class_force(sfm_ctx* c, int cls, int64_t n, double* out) {
    // Per-class arrays are produced on demand from the current state (Force.get_force semantics).
    return sfm_force(c, cls, n, out);
}

int sfm_gather_buffer(sfm_ctx* c, void** device_ptr, size_t* bytes_per_rank) {
    SFM_TRY(check_ctx(c));
    if (!device_ptr || !bytes_per_rank) return fail("null pointer");
    if (!c->planes.p) return fail("no state uploaded yet");
    *device_ptr = c->planes.p;
    *bytes_per_rank = sizeof(float) * NPLANES * (size_t)c->rows_pad;
    return 0;
}

int sfm_step_begin(sfm_ctx* c) {
    SFM_TRY(check_ctx(c));
    if (!c->have_params) return fail("sfm_set_params must be called first");
    if (c->step_open) return fail("sfm_step_begin called twice without sfm_step_end");
    if (c->n == 0 && c->world == 1) return 0;
    return step_begin(c);
}

int sfm_step_end(sfm_ctx* c, int integrate_positions) {
    SFM_TRY(check_ctx(c));
    if (c->n == 0 && c->world == 1) return 0;
    if (!c->step_open) return fail("sfm_step_end without sfm_step_begin");
    return step_end(c, true, integrate_positions != 0, false);
}

int sfm_force_accumulator(sfm_ctx* c, void** device_ptr, size_t* bytes_per_rank) {
    SFM_TRY(check_ctx(c));
    if (!device_ptr || !bytes_per_rank) return fail("null pointer");
    if (!c->planes.p) return fail("no state uploaded yet");
    SFM_TRY(c->facc.ensure((size_t)c->world * c->rows_pad * 4));
    *device_ptr = c->facc.p;
    *bytes_per_rank = sizeof(long long) * 4 * (size_t)c->rows_pad;
    return 0;
}

int sfm_stage(sfm_ctx* c) {
    SFM_TRY(check_ctx(c));
    if (!c->have_params) return fail("sfm_set_params must be called first");
    if (!c->planes.p) return fail("no state uploaded yet");
    if (!c->staged) SFM_TRY(launch_stage(c));
    return 0;
}

/* ================================================================================================================
 * lifecycle: mode machines, gap acceptance, routes, device-generated vehicle rings, recorder
 * ================================================================================================================ */

int sfm_set_reorder_interval(sfm_ctx* c, int ticks) {
    SFM_TRY(check_ctx(c));
    if (ticks < 0) return fail("negative interval");
    c->reorder_interval = ticks;
    c->order_due = true;                        // due at the next (re)staging
    return 0;
}

int sfm_reorder_slots(sfm_ctx* c) {
    SFM_TRY(check_ctx(c));
    if (c->n <= 0) return fail("no state uploaded yet");
    if (c->step_open) return fail("a step is open: call sfm_step_end first");
    SFM_TRY(rebuild_order(c));
    c->staged = false;
    return 0;
}

int sfm_get_slot_order(sfm_ctx* c, int64_t n, int32_t* slot_of_row) {
    SFM_TRY(check_ctx(c));
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (!slot_of_row) return fail("null array");
    if (!order_valid(c)) {
        for (int64_t i = 0; i < n; ++i) slot_of_row[i] = (int32_t)i;
        return 0;
    }
    SFM_CUDA(cudaMemcpyAsync(slot_of_row, c->slot_of_row.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int sfm_set_slot_order(sfm_ctx* c, int64_t n, const int32_t* slot_of_row) {
    SFM_TRY(check_ctx(c));
    if (n != c->n || n <= 0) return fail("row count differs from the uploaded state");
    if (!slot_of_row) return fail("null array");
    if (c->step_open) return fail("a step is open: call sfm_step_end first");
    std::vector<uint8_t> seen((size_t)n, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t s = slot_of_row[i];
        if (s < 0 || s >= n || seen[(size_t)s]) return fail("slot_of_row is not a permutation of [0, n)");
        seen[(size_t)s] = 1;
    }
    SFM_TRY(c->slot_of_row.ensure(n)); SFM_TRY(c->row_of_slot.ensure(c->rows_pad));
    SFM_CUDA(cudaMemcpyAsync(c->slot_of_row.p, slot_of_row, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    k8_invert_order<<<cdiv(c->rows_pad, 256), 256, 0, c->stream>>>(c->slot_of_row.p, (int)n, (int)c->rows_pad, c->row_of_slot.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    SFM_CUDA(cudaStreamSynchronize(c->stream));          // the host array may go away
    c->order_n = c->n; c->order_rows_pad = c->rows_pad;
    c->order_tick = c->ticks_total;
    c->order_due = false;
    c->staged = false;
    c->config_epoch += 1;
    return 0;
}

int sfm_set_mode_machines(sfm_ctx* c, int64_t n, const double* initial_speed, const double* crossing_speed,
                          const double* safety_margin, const double* mode_speed, const double* next_mode_time,
                          double waiting_time) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (n > 0 && (!initial_speed || !crossing_speed || !safety_margin || !mode_speed || !next_mode_time))
        return fail("null mode-machine array");
    c->waiting_time = waiting_time;
    SFM_TRY(ensure_life_counters(c));
    if (n == 0) { c->have_mm = true; return 0; }
    SFM_TRY(c->mm_speed.ensure(n)); SFM_TRY(c->mm_initial.ensure(n)); SFM_TRY(c->mm_crossing.ensure(n));
    SFM_TRY(c->mm_margin.ensure(n)); SFM_TRY(c->mm_next_time.ensure(n));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    const size_t b = sizeof(double) * n;
    SFM_CUDA(cudaMemcpy(c->mm_initial.p, initial_speed, b, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(c->mm_crossing.p, crossing_speed, b, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(c->mm_margin.p, safety_margin, b, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(c->mm_speed.p, mode_speed, b, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(c->mm_next_time.p, next_mode_time, b, cudaMemcpyHostToDevice));
    c->have_mm = true;
    return 0;
}

int sfm_set_traffic(sfm_ctx* c, int64_t n_vehicles, const double* centers, const double* velocities,
                    const double* extents) {
    SFM_TRY(check_ctx(c));
    if (n_vehicles < 0 || n_vehicles > (int64_t)1 << 24) return fail("bad vehicle count");
    c->tr_count = 0;
    if (n_vehicles == 0) return 0;
    if (!centers || !velocities || !extents) return fail("null traffic array");
    SFM_TRY(c->tr_center.ensure(n_vehicles)); SFM_TRY(c->tr_vel.ensure(n_vehicles));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    SFM_CUDA(cudaMemcpy(c->tr_center.p, centers, sizeof(double2) * n_vehicles, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(c->tr_vel.p, velocities, sizeof(double2) * n_vehicles, cudaMemcpyHostToDevice));
    c->tr_ext0[0] = extents[0]; c->tr_ext0[1] = extents[1];          // vehicle_extents[:][0], check_traffic.py:35-36
    c->tr_count = (int)n_vehicles;
    return 0;
}

int sfm_tick_modes(sfm_ctx* c, double sim_time) {
    SFM_TRY(check_ctx(c));
    return tick_modes_impl(c, sim_time);
}

int sfm_download_modes(sfm_ctx* c, int64_t n, uint8_t* mode, double* mode_speed, double* next_mode_time,
                       double* target_speed) {
    SFM_TRY(check_ctx(c));
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (n == 0) return 0;
    if ((mode_speed || next_mode_time) && !c->have_mm) return fail("no mode machines on the device");
    if (mode) SFM_CUDA(cudaMemcpyAsync(mode, c->mode.p, n, cudaMemcpyDeviceToHost, c->stream));
    if (mode_speed) SFM_CUDA(cudaMemcpyAsync(mode_speed, c->mm_speed.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (next_mode_time)
        SFM_CUDA(cudaMemcpyAsync(next_mode_time, c->mm_next_time.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (target_speed)
        SFM_CUDA(cudaMemcpy2DAsync(target_speed, sizeof(double), &c->vels.p[0].w, sizeof(double4), sizeof(double), n,
                                   cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int sfm_set_routes(sfm_ctx* c, int64_t n, const int64_t* offsets, const double* waypoints, const uint8_t* crossing,
                   double distance_threshold, int fused) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (!c->have_mm) return fail("sfm_set_mode_machines must be called first (a hand-over requests a mode)");
    if (n > 0 && !offsets) return fail("null route offsets");
    c->route_threshold = distance_threshold;
    c->routes_fused = fused != 0;
    SFM_TRY(ensure_life_counters(c));
    if (n == 0) { c->have_routes = true; return 0; }
    const int64_t total = offsets[n];
    if (total < 0 || total > (int64_t)INT32_MAX) return fail("bad route table");
    if (total > 0 && (!waypoints || !crossing)) return fail("null waypoint array");
    std::vector<int> cursor(n), end(n);
    for (int64_t i = 0; i < n; ++i) {
        if (offsets[i + 1] < offsets[i]) return fail("route offsets must be non-decreasing");
        cursor[i] = (int)offsets[i];
        end[i] = (int)offsets[i + 1];
    }
    SFM_TRY(c->rt_cursor.ensure(n)); SFM_TRY(c->rt_end.ensure(n)); SFM_TRY(c->finished.ensure(n));
    SFM_TRY(c->rt_wp.ensure(3 * std::max<int64_t>(total, 1))); SFM_TRY(c->rt_cross.ensure(std::max<int64_t>(total, 1)));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    SFM_CUDA(cudaMemcpy(c->rt_cursor.p, cursor.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    c->rt_begin = cursor;
    c->rt_total = total;
    SFM_CUDA(cudaMemcpy(c->rt_end.p, end.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemset(c->finished.p, 0, n));
    if (total > 0) {
        SFM_CUDA(cudaMemcpy(c->rt_wp.p, waypoints, sizeof(double) * 3 * total, cudaMemcpyHostToDevice));
        SFM_CUDA(cudaMemcpy(c->rt_cross.p, crossing, total, cudaMemcpyHostToDevice));
    }
    c->have_routes = true;
    return 0;
}

int sfm_advance_waypoints(sfm_ctx* c) {
    SFM_TRY(check_ctx(c));
    if (!c->have_routes || !c->have_mm) return fail("sfm_set_routes must be called first");
    if (c->n == 0) return 0;
    SpanGuard g(c, ST_LIFECYCLE);
    k4_advance_waypoints<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->n, c->locr.p, c->wp.p, c->mode.p, routes_of(c),
                                                                   mode_machines(c), c->sim_time);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    return 0;
}

int sfm_download_routes(sfm_ctx* c, int64_t n, int64_t* cursor, uint8_t* finished, double* next_waypoint) {
    SFM_TRY(check_ctx(c));
    if (n != c->n) return fail("row count differs from the uploaded state");
    if (n == 0) return 0;
    if ((cursor || finished) && !c->have_routes) return fail("no routes on the device");
    std::vector<int> cur;
    if (cursor) {
        cur.resize(n);
        SFM_CUDA(cudaMemcpyAsync(cur.data(), c->rt_cursor.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    }
    if (finished) SFM_CUDA(cudaMemcpyAsync(finished, c->finished.p, n, cudaMemcpyDeviceToHost, c->stream));
    if (next_waypoint)
        SFM_CUDA(cudaMemcpyAsync(next_waypoint, c->next_wp3.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    if (cursor) for (int64_t i = 0; i < n; ++i) cursor[i] = cur[i] - c->rt_begin[i];
    return 0;
}


int sfm_append_pedestrians(sfm_ctx* c, int64_t m, const double* loc, const double* vel, const double* wp3,
                           const double* radius, const double* speed, const uint8_t* mode, const double* initial_speed,
                           const double* crossing_speed, const double* safety_margin, const double* mode_speed,
                           const double* next_mode_time, const int64_t* route_offsets, const double* waypoints,
                           const uint8_t* crossing) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;
    if (!c->have_params) return fail("sfm_set_params must be called first");
    if (c->world > 1) return fail("spawning changes the row partition: single-rank contexts only");
    if (c->step_open) return fail("spawn between sfm_step_begin and sfm_step_end");
    if (m < 0) return fail("negative row count");
    if (m == 0) return 0;
    if (!loc || !vel || !wp3 || !radius || !speed || !mode) return fail("null state array");
    if (c->have_mm && (!initial_speed || !crossing_speed || !safety_margin || !mode_speed || !next_mode_time))
        return fail("this context carries mode machines: the new pedestrians need theirs");
    if (c->have_routes && !route_offsets) return fail("this context carries routes: the new pedestrians need theirs");
    const int64_t n0 = c->n, n1 = c->n + m;
    if (n1 > (int64_t)1 << 28) return fail("bad row count");
    cudaStream_t st = c->stream;
    // per-row tables grow in place (contents of the first n0 rows kept)
    SFM_TRY(c->locr.grow_keep(n1, n0, st)); SFM_TRY(c->vels.grow_keep(n1, n0, st)); SFM_TRY(c->wp.grow_keep(n1, n0, st));
    SFM_TRY(c->mode.grow_keep(n1, n0, st)); SFM_TRY(c->next_wp3.grow_keep(3 * n1, 3 * n0, st));
    SFM_TRY(c->raw_a.ensure(3 * n1)); SFM_TRY(c->raw_b.ensure(3 * n1)); SFM_TRY(c->raw_c.ensure(3 * n1));
    SFM_TRY(c->raw_d.ensure(n1)); SFM_TRY(c->raw_e.ensure(n1)); SFM_TRY(c->raw_mode.ensure(n1));
    SFM_CUDA(cudaMemcpyAsync(c->raw_a.p, loc, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, st));
    SFM_CUDA(cudaMemcpyAsync(c->raw_b.p, vel, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, st));
    SFM_CUDA(cudaMemcpyAsync(c->raw_c.p, wp3, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, st));
    SFM_CUDA(cudaMemcpyAsync(c->raw_d.p, radius, sizeof(double) * m, cudaMemcpyHostToDevice, st));
    SFM_CUDA(cudaMemcpyAsync(c->raw_e.p, speed, sizeof(double) * m, cudaMemcpyHostToDevice, st));
    SFM_CUDA(cudaMemcpyAsync(c->raw_mode.p, mode, m, cudaMemcpyHostToDevice, st));
    pack_state<<<cdiv(m, 256), 256, 0, st>>>(m, c->raw_a.p, c->raw_b.p, c->raw_c.p, c->raw_d.p, c->raw_e.p, c->raw_mode.p,
                                              c->locr.p + n0, c->vels.p + n0, c->wp.p + n0, c->mode.p + n0);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    SFM_CUDA(cudaMemcpyAsync(c->next_wp3.p + 3 * n0, c->raw_c.p, sizeof(double) * 3 * m, cudaMemcpyDeviceToDevice, st));
    if (c->have_mm) {
        DevBuf<double>* cols[5] = {&c->mm_initial, &c->mm_crossing, &c->mm_margin, &c->mm_speed, &c->mm_next_time};
        const double* src[5] = {initial_speed, crossing_speed, safety_margin, mode_speed, next_mode_time};
        for (int k = 0; k < 5; ++k) {
            SFM_TRY(cols[k]->grow_keep(n1, n0, st));
            SFM_CUDA(cudaMemcpyAsync(cols[k]->p + n0, src[k], sizeof(double) * m, cudaMemcpyHostToDevice, st));
        }
    }
    if (c->have_routes) {
        const int64_t add = route_offsets[m] - route_offsets[0];
        if (add < 0 || c->rt_total + add > (int64_t)INT32_MAX) return fail("bad route table");
        if (add > 0 && (!waypoints || !crossing)) return fail("null waypoint array");
        std::vector<int> cursor(m), end(m);
        for (int64_t i = 0; i < m; ++i) {
            if (route_offsets[i + 1] < route_offsets[i]) return fail("route offsets must be non-decreasing");
            cursor[i] = (int)(c->rt_total + route_offsets[i] - route_offsets[0]);
            end[i] = (int)(c->rt_total + route_offsets[i + 1] - route_offsets[0]);
        }
        SFM_TRY(c->rt_cursor.grow_keep(n1, n0, st)); SFM_TRY(c->rt_end.grow_keep(n1, n0, st));
        SFM_TRY(c->finished.grow_keep(n1, n0, st));
        SFM_TRY(c->rt_wp.grow_keep(3 * std::max<int64_t>(c->rt_total + add, 1), 3 * c->rt_total, st));
        SFM_TRY(c->rt_cross.grow_keep(std::max<int64_t>(c->rt_total + add, 1), c->rt_total, st));
        SFM_CUDA(cudaStreamSynchronize(st));
        SFM_CUDA(cudaMemcpy(c->rt_cursor.p + n0, cursor.data(), sizeof(int) * m, cudaMemcpyHostToDevice));
        SFM_CUDA(cudaMemcpy(c->rt_end.p + n0, end.data(), sizeof(int) * m, cudaMemcpyHostToDevice));
        SFM_CUDA(cudaMemset(c->finished.p + n0, 0, m));
        if (add > 0) {
            SFM_CUDA(cudaMemcpy(c->rt_wp.p + 3 * c->rt_total, waypoints + 3 * route_offsets[0], sizeof(double) * 3 * add,
                                cudaMemcpyHostToDevice));
            SFM_CUDA(cudaMemcpy(c->rt_cross.p + c->rt_total, crossing + route_offsets[0], add, cudaMemcpyHostToDevice));
        }
        c->rt_begin.insert(c->rt_begin.end(), cursor.begin(), cursor.end());
        c->rt_total += add;
    }
    c->n = n1;
    c->rec_ident_n = -1;
    SFM_TRY(ensure_layout(c, n1));               // the staging planes may have to grow: every slot is restaged
    c->staged = false;
    c->perm_valid = false;
    c->rec_capacity = 0;
    SFM_CUDA(cudaStreamSynchronize(st));         // host arrays may be released by the caller on return
    return 0;
}

int sfm_despawn_finished(sfm_ctx* c, int64_t* n_after, int64_t* n_removed) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    if (!c->have_routes || !c->have_mm) return fail("sfm_set_routes must be called first");
    if (c->world > 1) return fail("despawning changes the row partition: single-rank contexts only");
    if (c->step_open) return fail("despawn between sfm_step_begin and sfm_step_end");
    const int64_t n = c->n;
    if (n_after) *n_after = n;
    if (n_removed) *n_removed = 0;
    if (n == 0) return 0;
    DevBuf<int>& idx = c->ped_cursor;                      // scratch of the pedestrian binning, rebuilt every tick anyway
    SFM_TRY(idx.ensure(n + 1));
    {
        SpanGuard g(c, ST_LIFECYCLE);
        k4_keep_flags<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->finished.p, idx.p);
        c->launches += 1;
        SFM_TRY(launch_exclusive_scan(c, idx.p, (int)n + 1, c->ped_scan_tmp, c->stream));      // idx[n] = survivors
        SFM_CUDA(cudaGetLastError());
    }
    int survivors = 0;
    SFM_CUDA(cudaMemcpyAsync(&survivors, idx.p + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    if (survivors == n) return 0;
    {
        SpanGuard g(c, ST_LIFECYCLE);
        DevBuf<double4>& s4 = c->cmp4;
        DevBuf<double2>& s2 = c->cmp2;
        DevBuf<double3s>& s3 = c->cmp3;
        DevBuf<double>& s1 = c->cmp1;
        DevBuf<int>& si = c->cmpi;
        DevBuf<uint8_t>& sb = c->cmpb;
        const uint8_t* fin = c->finished.p;
        SFM_TRY(compact_column(c, c->locr, s4, n, fin, idx.p));
        SFM_TRY(compact_column(c, c->vels, s4, n, fin, idx.p));
        SFM_TRY(compact_column(c, c->wp, s2, n, fin, idx.p));
        SFM_TRY(compact_column(c, c->mode, sb, n, fin, idx.p));
        {   // next_wp3 is a double[n][3] buffer
            DevBuf<double3s> view; view.p = reinterpret_cast<double3s*>(c->next_wp3.p); view.cap = c->next_wp3.cap / 3;
            SFM_TRY(compact_column(c, view, s3, n, fin, idx.p));
            c->next_wp3.p = reinterpret_cast<double*>(view.p); c->next_wp3.cap = view.cap * 3;
        }
        SFM_TRY(compact_column(c, c->mm_speed, s1, n, fin, idx.p));
        SFM_TRY(compact_column(c, c->mm_initial, s1, n, fin, idx.p));
        SFM_TRY(compact_column(c, c->mm_crossing, s1, n, fin, idx.p));
        SFM_TRY(compact_column(c, c->mm_margin, s1, n, fin, idx.p));
        SFM_TRY(compact_column(c, c->mm_next_time, s1, n, fin, idx.p));
        SFM_TRY(compact_column(c, c->rt_cursor, si, n, fin, idx.p));
        SFM_TRY(compact_column(c, c->rt_end, si, n, fin, idx.p));
        // the route table's host copy of the first entries (relative cursors) follows the same mask
        std::vector<uint8_t> fin_host(n);
        SFM_CUDA(cudaMemcpyAsync(fin_host.data(), fin, n, cudaMemcpyDeviceToHost, c->stream));
        SFM_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<int> begin;
        begin.reserve(survivors);
        for (int64_t i = 0; i < n; ++i) if (!fin_host[i]) begin.push_back(c->rt_begin[i]);
        c->rt_begin.swap(begin);
        SFM_CUDA(cudaMemsetAsync(c->finished.p, 0, n, c->stream));
    }
    c->n = survivors;
    c->rec_ident_n = -1;
    c->staged = false;                   // every staged slot moved
    c->perm_valid = false;
    c->rec_capacity = 0;                 // recorded frames have a fixed row count: the recorder must be re-armed
    if (n_after) *n_after = survivors;
    if (n_removed) *n_removed = n - survivors;
    return 0;
}

int sfm_lifecycle_counters(sfm_ctx* c, int64_t* out4) {
    SFM_TRY(check_ctx(c));
    if (!out4) return fail("null pointer");
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    if (!c->life_counters.p) return 0;
    unsigned long long v[4];
    SFM_CUDA(cudaMemcpyAsync(v, c->life_counters.p, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 4; ++k) out4[k] = (int64_t)v[k];
    return 0;
}

namespace {

VehicleArgs vehicle_args(sfm_ctx* c, double dt) {
    VehicleArgs a{};
    a.count = (int)c->dyn.s.count; a.center = c->dyn.s.center; a.yaw_deg = c->veh_yaw.p; a.velocity = c->dyn.s.velocity;
    a.extent = c->veh_extent.p; a.offset = c->dyn.s.offset; a.point = c->dyn.s.point;
    a.size_factor = c->veh_size_factor; a.dt = dt;
    return a;
}

// (cell, index) order of the dynamic set on its FIXED grid (centres outside the grid are clamped into the border cells,
// which never widens the index distance between a pedestrian's cell and an obstacle within the cutoff)
int rebin_dynamic(sfm_ctx* c) {
    SetStorage& st = c->dyn;
    const int count = (int)st.s.count, ncell = st.s.grid.nx * st.s.grid.ny;
    int key_bits = 1;
    while ((1 << key_bits) < ncell) ++key_bits;
    SpanGuard g(c, ST_CELLS);
    k2_item_keys<<<cdiv(count, 256), 256, 0, c->stream>>>(st.center.p, count, st.s.grid, st.key.p, st.cell_item.p);
    c->launches += 1;
    SFM_TRY(sort_items(c, st, count, key_bits));
    k2_cell_bounds<<<cdiv(ncell + 1, 256), 256, 0, c->stream>>>(st.key.p, count, ncell, st.cell_start.p);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    return 0;
}

int launch_vehicle_rings(sfm_ctx* c) {
    SpanGuard g(c, ST_LIFECYCLE);
    const int np = (int)c->dyn.s.n_points;
    k5_vehicle_rings<<<cdiv(np, 256), 256, 0, c->stream>>>(vehicle_args(c, 0.0), np);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int sfm_set_vehicles(sfm_ctx* c, int64_t n_vehicles, const double* centers, const double* yaw_deg,
                     const double* velocities, const double* extents, double resolution, double size_factor) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    if (!c->have_params) return fail("sfm_set_params must be called first");
    SetStorage& st = c->dyn;
    st.s.count = 0;
    c->have_vehicles = false;
    if (n_vehicles <= 0) return 0;
    if (!centers || !yaw_deg || !velocities || !extents) return fail("null vehicle array");
    if (!(resolution > 0.0)) return fail("resolution must be positive");
    if (n_vehicles > (int64_t)1 << 24) return fail("too many vehicles");
    std::vector<int> off(n_vehicles + 1, 0);
    for (int64_t v = 0; v < n_vehicles; ++v) {           // obstacles.py:273-274
        const double circumference = 2.0 * extents[2 * v] + 2.0 * extents[2 * v + 1];
        const long long samples = std::max<long long>(6, (long long)(circumference / resolution));
        if (off[v] + samples > (long long)INT32_MAX) return fail("vehicle rings too large");
        off[v + 1] = off[v] + (int)samples;
    }
    const int64_t np = off[n_vehicles];
    SFM_TRY(st.center.ensure(n_vehicles)); SFM_TRY(st.cutoff.ensure(n_vehicles)); SFM_TRY(st.velocity.ensure(n_vehicles));
    SFM_TRY(st.offset.ensure(n_vehicles + 1)); SFM_TRY(st.point.ensure(np));
    SFM_TRY(c->veh_extent.ensure(n_vehicles)); SFM_TRY(c->veh_yaw.ensure(n_vehicles));
    const double thr = c->params.dynamic_obs.perception_threshold;
    st.threshold = thr;
    st.stale = false;
    std::vector<double> cut(n_vehicles, thr);
    std::vector<float> tol(n_vehicles);
    for (int64_t v = 0; v < n_vehicles; ++v)          // ring points lie within size_factor * max(extent) of the centre
        tol[v] = bracket_tolerance(std::max(std::fabs(thr), 1.001 * std::fabs(size_factor) *
                                                                std::max(std::fabs(extents[2 * v]), std::fabs(extents[2 * v + 1]))));
    SFM_TRY(st.tol.ensure(n_vehicles));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    SFM_CUDA(cudaMemcpy(st.center.p, centers, sizeof(double2) * n_vehicles, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(st.velocity.p, velocities, sizeof(double2) * n_vehicles, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(st.cutoff.p, cut.data(), sizeof(double) * n_vehicles, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(st.offset.p, off.data(), sizeof(int) * (n_vehicles + 1), cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(c->veh_extent.p, extents, sizeof(double2) * n_vehicles, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(c->veh_yaw.p, yaw_deg, sizeof(double) * n_vehicles, cudaMemcpyHostToDevice));
    SFM_CUDA(cudaMemcpy(st.tol.p, tol.data(), sizeof(float) * n_vehicles, cudaMemcpyHostToDevice));
    SegmentSet& s = st.s;
    s.tol = st.tol.p;
    s.chunk_first = nullptr;                       // rings are regenerated every tick: no chunk / chord tables
    s.chord0 = nullptr;
    s.center = st.center.p; s.cutoff = st.cutoff.p; s.velocity = st.velocity.p; s.offset = st.offset.p;
    s.point = st.point.p; s.n_points = np;
    c->veh_size_factor = size_factor;
    c->tr_ext0[0] = extents[0]; c->tr_ext0[1] = extents[1];
    SFM_TRY(build_set_grid(c, st, n_vehicles, centers, std::isnan(thr) ? 0.0 : thr));
    s.count = n_vehicles;
    c->have_vehicles = true;
    return launch_vehicle_rings(c);
}

int sfm_advance_vehicles(sfm_ctx* c, double dt) {
    SFM_TRY(check_ctx(c));
    if (!c->have_vehicles) return fail("sfm_set_vehicles must be called first");
    {
        SpanGuard g(c, ST_LIFECYCLE);
        k5_advance_vehicles<<<cdiv(c->dyn.s.count, 256), 256, 0, c->stream>>>(vehicle_args(c, dt));
        c->launches += 1;
        SFM_CUDA(cudaGetLastError());
    }
    SFM_TRY(launch_vehicle_rings(c));
    return rebin_dynamic(c);
}

int sfm_download_vehicles(sfm_ctx* c, int64_t n_vehicles, double* centers, int64_t* offsets, int64_t point_capacity,
                          double* points) {
    SFM_TRY(check_ctx(c));
    if (!c->have_vehicles) return fail("no device-resident vehicle set");
    if (n_vehicles != c->dyn.s.count) return fail("vehicle count differs");
    std::vector<int> off(n_vehicles + 1);
    SFM_CUDA(cudaMemcpyAsync(off.data(), c->dyn.s.offset, sizeof(int) * (n_vehicles + 1), cudaMemcpyDeviceToHost, c->stream));
    if (centers)
        SFM_CUDA(cudaMemcpyAsync(centers, c->dyn.s.center, sizeof(double2) * n_vehicles, cudaMemcpyDeviceToHost, c->stream));
    if (points) {
        if (point_capacity < c->dyn.s.n_points) return fail("point buffer too small");
        SFM_CUDA(cudaMemcpyAsync(points, c->dyn.s.point, sizeof(double2) * c->dyn.s.n_points, cudaMemcpyDeviceToHost, c->stream));
    }
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    if (offsets) for (int64_t v = 0; v <= n_vehicles; ++v) offsets[v] = off[v];
    return 0;
}

int sfm_record_begin(sfm_ctx* c, int64_t capacity_frames) {
    SFM_TRY(check_ctx(c));
    if (capacity_frames < 0) return fail("negative capacity");
    c->rec_rows = c->n;
    c->rec_count = 0;
    c->rec_times.clear();
    c->rec_capacity = 0;
    if (capacity_frames == 0 || c->n == 0) return 0;
    SFM_TRY(c->rec_xyv.ensure((size_t)capacity_frames * c->n));
    SFM_TRY(c->rec_mode.ensure((size_t)capacity_frames * c->n));
    c->rec_capacity = capacity_frames;
    return 0;
}

int sfm_record_frame(sfm_ctx* c, double sim_time) {
    SFM_TRY(check_ctx(c));
    if (c->rec_capacity == 0) return fail("sfm_record_begin must be called first");
    if (c->rec_rows != c->n) return fail("row count changed since sfm_record_begin");
    if (c->rec_count >= c->rec_capacity) return fail("frame buffer full: download and call sfm_record_begin again");
    SpanGuard g(c, ST_LIFECYCLE);
    const size_t at = (size_t)c->rec_count * c->n;
    k6_record_frame<<<cdiv(c->n, 256), 256, 0, c->stream>>>(c->n, c->locr.p, c->vels.p, c->mode.p, c->rec_xyv.p + at,
                                                              c->rec_mode.p + at);
    c->launches += 1;
    SFM_CUDA(cudaGetLastError());
    c->rec_times.push_back(sim_time);
    c->rec_count += 1;
    return 0;
}

int sfm_download_frames(sfm_ctx* c, int64_t first, int64_t count, double* xyv, uint8_t* mode, double* times,
                        int64_t* frames_recorded) {
    SFM_TRY(check_ctx(c));
    if (frames_recorded) *frames_recorded = c->rec_count;
    if (count == 0) return 0;
    if (first < 0 || count < 0 || first + count > c->rec_count) return fail("frame range out of bounds");
    const size_t at = (size_t)first * c->rec_rows, items = (size_t)count * c->rec_rows;
    if (xyv) SFM_CUDA(cudaMemcpyAsync(xyv, c->rec_xyv.p + at, sizeof(double4) * items, cudaMemcpyDeviceToHost, c->stream));
    if (mode) SFM_CUDA(cudaMemcpyAsync(mode, c->rec_mode.p + at, items, cudaMemcpyDeviceToHost, c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    if (times) for (int64_t k = 0; k < count; ++k) times[k] = c->rec_times[first + k];
    return 0;
}

/* ---- peer-memory exchange (K7) ------------------------------------------------------------------------------------- */
int sfm_peer_export(sfm_ctx* c, void* handles) {
    SFM_TRY(check_ctx(c));
    if (!handles) return fail("null pointer");
    if (!c->partition_fixed || c->world < 2) return fail("sfm_set_partition (world >= 2) must be called first");
    if (c->world > MAX_PEERS) return fail("the peer-memory exchange supports up to 8 ranks (one box)");
    if (!c->planes.p) return fail("no state uploaded yet");
    SFM_TRY(c->facc.ensure((size_t)c->world * c->rows_pad * 4));
    SFM_TRY(c->flags.ensure(2 * MAX_PEERS));
    SFM_CUDA(cudaMemsetAsync(c->flags.p, 0, 2 * MAX_PEERS * sizeof(unsigned), c->stream));
    SFM_CUDA(cudaStreamSynchronize(c->stream));
    cudaIpcMemHandle_t* h = reinterpret_cast<cudaIpcMemHandle_t*>(handles);
    SFM_CUDA(cudaIpcGetMemHandle(&h[0], c->planes.p));
    SFM_CUDA(cudaIpcGetMemHandle(&h[1], c->facc.p));
    SFM_CUDA(cudaIpcGetMemHandle(&h[2], c->flags.p));
    return 0;
}

int sfm_peer_import(sfm_ctx* c, const void* all_handles) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    if (!all_handles) return fail("null pointer");
    if (!c->flags.p) return fail("sfm_peer_export must be called first");
    const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(all_handles);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) {
            c->peer_planes[r] = c->planes.p; c->peer_facc[r] = c->facc.p; c->peer_flags[r] = c->flags.p;
            continue;
        }
        SFM_CUDA(cudaIpcOpenMemHandle(&c->peer_planes[r], h[3 * r + 0], cudaIpcMemLazyEnablePeerAccess));
        SFM_CUDA(cudaIpcOpenMemHandle(&c->peer_facc[r], h[3 * r + 1], cudaIpcMemLazyEnablePeerAccess));
        SFM_CUDA(cudaIpcOpenMemHandle(&c->peer_flags[r], h[3 * r + 2], cudaIpcMemLazyEnablePeerAccess));
    }
    c->p2p = true;
    c->parity = 0;
    c->staged = false;
    return 0;
}

int sfm_peer_barrier(sfm_ctx* c) {
    SFM_TRY(check_ctx(c));
    if (!c->p2p) return fail("the peer-memory exchange is not set up");
    return launch_barrier(c);
}

int sfm_step_peer(sfm_ctx* c, int n_steps, int integrate_positions) {
    SFM_TRY(check_ctx(c));
    if (!c->p2p) return fail("the peer-memory exchange is not set up (sfm_peer_export / sfm_peer_import)");
    if (!c->have_params) return fail("sfm_set_params must be called first");
    for (int s = 0; s < n_steps; ++s) {
        SFM_TRY(step_begin(c));                          // pair accumulation into this rank's accumulator + cell-list forces
        SFM_TRY(launch_barrier(c));                      // every rank's accumulator is complete
        SFM_TRY(step_end(c, true, integrate_positions != 0, false));   // pull-reduce + finish, K3 with the fused push
        SFM_TRY(launch_barrier(c));                      // everybody has read my accumulator and received my rows
        c->parity ^= 1;
    }
    return 0;
}

int sfm_peer_status(sfm_ctx* c, int64_t* barriers, int* timed_out) {
    SFM_TRY(check_ctx(c));
    if (barriers) *barriers = c->barriers;
    if (timed_out) {
        *timed_out = 0;
        if (c->flags.p) {
            unsigned e = 0;                  // bit r: a barrier of this rank gave up waiting for rank r
            SFM_CUDA(cudaMemcpyAsync(&e, c->flags.p + MAX_PEERS, sizeof(e), cudaMemcpyDeviceToHost, c->stream));
            SFM_CUDA(cudaStreamSynchronize(c->stream));
            *timed_out = (int)e;
        }
    }
    return 0;
}

int sfm_set_profiling(sfm_ctx* c, int enabled) {
    SFM_TRY(check_ctx(c));
    c->config_epoch += 1;                      // a captured tick graph (sfm_step) must be rebuilt
    SFM_TRY(drain_spans(c));
    c->profiling = enabled != 0;
    return 0;
}

int sfm_reset_stats(sfm_ctx* c) {
    SFM_TRY(check_ctx(c));
    SFM_TRY(drain_spans(c));
    c->launches = c->steps = c->pair_launches = c->pair_evals = 0;
    for (double& m : c->ms) m = 0.0;
    c->graph_replays = 0;
    if (c->fixup_rows.p) SFM_CUDA(cudaMemsetAsync(c->fixup_rows.p, 0, 2 * sizeof(unsigned long long), c->stream));
    c->fixup_zeroed = c->fixup_rows.p != nullptr;
    return 0;
}

int sfm_get_stats(sfm_ctx* c, sfm_stats* out) {
    SFM_TRY(check_ctx(c));
    if (!out) return fail("null pointer");
    SFM_TRY(drain_spans(c));
    out->launches = c->launches; out->steps = c->steps; out->pair_launches = c->pair_launches;
    out->ms_pairs = c->ms[ST_PAIRS]; out->ms_cells = c->ms[ST_CELLS]; out->ms_segments = c->ms[ST_SEGMENTS];
    out->ms_integrate = c->ms[ST_INTEGRATE];
    out->ms_lifecycle = c->ms[ST_LIFECYCLE];
    out->graph_replays = c->graph_replays;
    out->fixup_rows = 0;
    out->local_tile_pairs = 0;
    out->pair_evaluations = c->pair_evals;
    if (c->fixup_rows.p && c->fixup_zeroed) {
        unsigned long long v[2] = {0, 0};
        SFM_CUDA(cudaMemcpyAsync(v, c->fixup_rows.p, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
        SFM_CUDA(cudaStreamSynchronize(c->stream));
        out->fixup_rows = (int64_t)v[0];
        out->local_tile_pairs = (int64_t)v[1];
    }
    return 0;
}

}  // extern "C"
