// K1 -- all-pairs pedestrian interaction force (Moussaid et al. 2009), float32, sm_100a.
//
// Replaces PedestrianForce._get_force (reference forces.py:74-117) together with stateutils.all_diffs / all_sums /
// normalize / angle_diff_2d (stateutils.py:32-128): for every ordered pair i != j
//     d = p_j - p_i (3-D), dist = |d|, e = d / dist, dl = dist - (r_i + r_j) [use_ped_radius]
//     D = lambda (v_i - v_j) + e, t = D / |D|, n = (-t_y, t_x, 0), B = gamma |D|
//     theta = angle_xy(e) - angle_xy(t) wrapped to (-pi, pi], theta' = theta - epsilon B
//     F_i += -A exp(-dl/B - (n' B theta')^2) t  -  A sign(theta') exp(-dl/B - (n B theta')^2) n
//
// Algebra used here (all exact identities, checked against the oracle in tests/):
//   * velocities are staged pre-multiplied by lambda: w = lambda v_i - lambda v_j, D = d * (1/dist) + w  (3 FFMA, e is
//     never materialised);
//   * theta = atan2(D_xy x d_xy, D_xy . d_xy), and D_xy x d_xy == w_xy x d_xy because e is parallel to d -- the
//     cancellation-free form;  atan2 is an octant-reduced degree-15 odd minimax polynomial (2.2e-7 relative);
//   * the force is accumulated against the *unnormalised* D with both coefficients scaled by 1/|D|;
//   * exp(x) = ex2(x log2 e), with -log2(e)/gamma, (n gamma)^2 log2 e and log2 A folded into constants, so each
//     exponent is one FFMA on top of the shared term.
// Degenerate pairs follow the reference's zero-safe normalisation (stateutils.py:88-90): |d| == 0 gives e = 0 and
// angle_xy(e) = atan2(0, 0) = 0; |D| == 0 gives t = 0, B = 0 and a vanishing contribution whenever dl > 0.
//
// Parallelisation: one CTA = 128 threads x IR rows, looping over a range of j-tiles (256 rows of the staged SoA planes)
// that TMA bulk copies (cp.async.bulk + mbarrier, two stages) bring into shared memory; every lane reads the same j
// (LDS.128 broadcast of 4 consecutive j per plane) and keeps its rows' partial force in registers.  grid.y splits the
// j range so the grid is many waves deep on 148 SMs; the per-split partial sums are written once (no atomics) and
// reduced in a fixed order by k1_reduce_fixup, which makes the result deterministic for a given launch geometry.
#pragma once

#include "sfm_common.cuh"

namespace sfm {

constexpr int K1_THREADS = 128;
constexpr int K1_TJ = ROW_ALIGN;      // j rows per staged tile
constexpr int K1_STAGES = 2;
constexpr int K1_PLANES = 7;          // PX..PVZ (PSPARE is not staged)
constexpr float K1_TINY = 1.0e-30f;

__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// atan2(y, x) in (-pi, pi]; atan2(0, 0) = 0 (numpy's convention for +0 arguments).  Max relative error 2.2e-7.
__device__ __forceinline__ float atan2_poly(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float q = mn * rcp_approx(fmaxf(mx, K1_TINY));
    const float s = q * q;
    float p = -4.693274875e-03f;
    p = fmaf(p, s, 2.425239913e-02f);
    p = fmaf(p, s, -5.948638773e-02f);
    p = fmaf(p, s, 9.914292465e-02f);
    p = fmaf(p, s, -1.401948078e-01f);
    p = fmaf(p, s, 1.996972388e-01f);
    p = fmaf(p, s, -3.333199075e-01f);
    p = fmaf(p, s, 9.999999010e-01f);
    p = p * q;
    if (ay > ax) p = 1.57079632679489662f - p;
    if (x < 0.0f) p = 3.14159265358979324f - p;
    return copysignf(p, y);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct PairAcc {
    float gx, gy, gz;     // accumulates -F
};

// One ordered pair (i <- j).  `self` pairs (same staged slot) are removed like the reference removes the diagonal
// (stateutils.py:44): their coefficients are replaced by zero.
template <bool RADIUS, bool DIAG>
__device__ __forceinline__ void pair_force(const float xi, const float yi, const float zi, const float ri,
                                           const float vxi, const float vyi, const float vzi, const float xj,
                                           const float yj, const float zj, const float rj, const float vxj,
                                           const float vyj, const float vzj, const bool self, const PairParams& pp,
                                           PairAcc& acc) {
    const float dx = xj - xi, dy = yj - yi, dz = zj - zi;
    const float dxy2 = fmaf(dy, dy, dx * dx);
    const float d2 = fmaf(dz, dz, dxy2);
    const float rinv = rsqrt_approx(fmaxf(d2, K1_TINY));
    const float dist = d2 * rinv;
    const float wx = vxi - vxj, wy = vyi - vyj, wz = vzi - vzj;
    const float Dx = fmaf(dx, rinv, wx), Dy = fmaf(dy, rinv, wy), Dz = fmaf(dz, rinv, wz);
    const float D2 = fmaf(Dz, Dz, fmaf(Dy, Dy, Dx * Dx));
    const float Dinv = rsqrt_approx(fmaxf(D2, K1_TINY));
    const float Dn = D2 * Dinv;
    float cross = fmaf(wx, dy, -(wy * dx));
    float dot = fmaf(Dx, dx, Dy * dy);
    if (dxy2 == 0.0f) {       // angle_xy(e) = atan2(0, 0) = 0: measure the angle of t against the x axis
        cross = -Dy;
        dot = Dx;
    }
    const float theta = atan2_poly(cross, dot);
    const float thp = fmaf(-pp.eps_gamma, Dn, theta);
    const float u = Dn * thp;
    const float u2 = u * u;
    float dl = dist;
    if (RADIUS) dl = dist - (ri + rj);
    const float y = fmaf(dl * Dinv, pp.neg_l2e_over_gamma, pp.log2A);
    const float e1 = ex2_approx(fmaf(-pp.c_nprime, u2, y));
    const float e2 = ex2_approx(fmaf(-pp.c_n, u2, y));
    float a = e1 * Dinv;
    float b = copysignf(e2 * Dinv, thp);
    if (thp == 0.0f) b = 0.0f;                       // np.sign(0) == 0 (forces.py:108)
    if (DIAG && self) {
        a = 0.0f;
        b = 0.0f;
    }
    acc.gx = fmaf(a, Dx, acc.gx);
    acc.gx = fmaf(-b, Dy, acc.gx);
    acc.gy = fmaf(a, Dy, acc.gy);
    acc.gy = fmaf(b, Dx, acc.gy);
    acc.gz = fmaf(a, Dz, acc.gz);
}

template <int IR, bool RADIUS, bool DIAG>
__device__ __forceinline__ void tile_pairs(const float (*__restrict__ tl)[K1_TJ], const float (&xi)[IR],
                                           const float (&yi)[IR], const float (&zi)[IR], const float (&ri)[IR],
                                           const float (&vxi)[IR], const float (&vyi)[IR], const float (&vzi)[IR],
                                           const int (&self_j)[IR], const PairParams& pp, PairAcc (&acc)[IR]) {
#pragma unroll 1
    for (int j = 0; j < K1_TJ; j += 4) {
        const float4 X = *reinterpret_cast<const float4*>(&tl[PX][j]);
        const float4 Y = *reinterpret_cast<const float4*>(&tl[PY][j]);
        const float4 Z = *reinterpret_cast<const float4*>(&tl[PZ][j]);
        float4 R = make_float4(0.f, 0.f, 0.f, 0.f);
        if (RADIUS) R = *reinterpret_cast<const float4*>(&tl[PR][j]);
        const float4 VX = *reinterpret_cast<const float4*>(&tl[PVX][j]);
        const float4 VY = *reinterpret_cast<const float4*>(&tl[PVY][j]);
        const float4 VZ = *reinterpret_cast<const float4*>(&tl[PVZ][j]);
#pragma unroll
        for (int r = 0; r < IR; ++r) {
            pair_force<RADIUS, DIAG>(xi[r], yi[r], zi[r], ri[r], vxi[r], vyi[r], vzi[r], X.x, Y.x, Z.x, R.x, VX.x, VY.x,
                                     VZ.x, DIAG && (self_j[r] == j + 0), pp, acc[r]);
            pair_force<RADIUS, DIAG>(xi[r], yi[r], zi[r], ri[r], vxi[r], vyi[r], vzi[r], X.y, Y.y, Z.y, R.y, VX.y, VY.y,
                                     VZ.y, DIAG && (self_j[r] == j + 1), pp, acc[r]);
            pair_force<RADIUS, DIAG>(xi[r], yi[r], zi[r], ri[r], vxi[r], vyi[r], vzi[r], X.z, Y.z, Z.z, R.z, VX.z, VY.z,
                                     VZ.z, DIAG && (self_j[r] == j + 2), pp, acc[r]);
            pair_force<RADIUS, DIAG>(xi[r], yi[r], zi[r], ri[r], vxi[r], vyi[r], vzi[r], X.w, Y.w, Z.w, R.w, VX.w, VY.w,
                                     VZ.w, DIAG && (self_j[r] == j + 3), pp, acc[r]);
        }
    }
}

// ---- packed float32x2 fast path (Blackwell FFMA2 / FMUL2 / FADD2) --------------------------------------------------
// The kernel is bound by instruction issue, and sm_100 can retire two FP32 operations per lane per issued instruction
// when they are packed in a 64-bit register pair.  One packed value holds the same quantity for two consecutive j
// (an LDS.128 of a staged plane yields two such pairs for free); MUFU and the few compare/select steps stay scalar on
// the halves.  This path carries NO zero guards: a degenerate pair (|d| = 0, |D| = 0, d_xy = 0) produces a NaN that
// poisons the row's partial sum, and k1_reduce_fixup recomputes exactly those rows with the guarded scalar code above.
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 neg2(f32x2 a) { return a ^ 0x8000000080000000ULL; }
__device__ __forceinline__ f32x2 rsqrt2(f32x2 a) {
    float lo, hi;
    unpack2(a, lo, hi);
    return pack2(rsqrt_approx(lo), rsqrt_approx(hi));
}
__device__ __forceinline__ f32x2 ex2_2(f32x2 a) {
    float lo, hi;
    unpack2(a, lo, hi);
    return pack2(ex2_approx(lo), ex2_approx(hi));
}
__device__ __forceinline__ f32x2 splat2(float v) { return pack2(v, v); }

// octant bookkeeping of atan2 on one half: p = atan(min/max) in [0, pi/4] -> angle of (x, y) in (-pi, pi]
__device__ __forceinline__ float octant_fix(float p, float ax, float ay, float x, float y) {
    if (ay > ax) p = 1.57079632679489662f - p;
    if (x < 0.0f) p = 3.14159265358979324f - p;
    return copysignf(p, y);
}

struct PackedConst {
    f32x2 eps_gamma_neg, k_exp, log2A, c_nprime_neg, c_n_neg;
    f32x2 a7, a6, a5, a4, a3, a2, a1, a0;
};

__device__ __forceinline__ PackedConst make_packed_const(const PairParams& pp) {
    PackedConst c;
    c.eps_gamma_neg = splat2(-pp.eps_gamma);
    c.k_exp = splat2(pp.neg_l2e_over_gamma);
    c.log2A = splat2(pp.log2A);
    c.c_nprime_neg = splat2(-pp.c_nprime);
    c.c_n_neg = splat2(-pp.c_n);
    c.a7 = splat2(-4.693274875e-03f);
    c.a6 = splat2(2.425239913e-02f);
    c.a5 = splat2(-5.948638773e-02f);
    c.a4 = splat2(9.914292465e-02f);
    c.a3 = splat2(-1.401948078e-01f);
    c.a2 = splat2(1.996972388e-01f);
    c.a1 = splat2(-3.333199075e-01f);
    c.a0 = splat2(9.999999010e-01f);
    return c;
}

struct PackedAcc {
    f32x2 ax, bx, ay, az;     // sum a*Dx, sum b*Dy, sum (a*Dy + b*Dx), sum a*Dz   (-F = (ax - bx, ay, az))
};

// Two ordered pairs (i <- j0) and (i <- j1) at once.
template <bool RADIUS>
__device__ __forceinline__ void pair_force2(const f32x2 xi, const f32x2 yi, const f32x2 zi, const f32x2 ri,
                                            const f32x2 vxi, const f32x2 vyi, const f32x2 vzi, const f32x2 xj,
                                            const f32x2 yj, const f32x2 zj, const f32x2 rj, const f32x2 vxj,
                                            const f32x2 vyj, const f32x2 vzj, const PackedConst& c, PackedAcc& acc) {
    const f32x2 dx = sub2(xj, xi), dy = sub2(yj, yi), dz = sub2(zj, zi);
    const f32x2 d2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
    const f32x2 rinv = rsqrt2(d2);
    const f32x2 dist = mul2(d2, rinv);
    const f32x2 wx = sub2(vxi, vxj), wy = sub2(vyi, vyj), wz = sub2(vzi, vzj);
    const f32x2 Dx = fma2(dx, rinv, wx), Dy = fma2(dy, rinv, wy), Dz = fma2(dz, rinv, wz);
    const f32x2 D2 = fma2(Dz, Dz, fma2(Dy, Dy, mul2(Dx, Dx)));
    const f32x2 Dinv = rsqrt2(D2);
    const f32x2 Dn = mul2(D2, Dinv);
    const f32x2 cross = fma2(wx, dy, neg2(mul2(wy, dx)));
    const f32x2 dot = fma2(Dx, dx, mul2(Dy, dy));
    // atan2(cross, dot): scalar octant reduction on the halves, packed polynomial
    float cl, ch, tl, th;
    unpack2(cross, cl, ch);
    unpack2(dot, tl, th);
    const float axl = fabsf(tl), ayl = fabsf(cl), axh = fabsf(th), ayh = fabsf(ch);
    const f32x2 mn = pack2(fminf(axl, ayl), fminf(axh, ayh));
    const f32x2 mxr = pack2(rcp_approx(fmaxf(axl, ayl)), rcp_approx(fmaxf(axh, ayh)));
    const f32x2 q = mul2(mn, mxr);
    const f32x2 s = mul2(q, q);
    f32x2 p = fma2(c.a7, s, c.a6);
    p = fma2(p, s, c.a5);
    p = fma2(p, s, c.a4);
    p = fma2(p, s, c.a3);
    p = fma2(p, s, c.a2);
    p = fma2(p, s, c.a1);
    p = fma2(p, s, c.a0);
    p = mul2(p, q);
    float pl, ph;
    unpack2(p, pl, ph);
    const f32x2 theta = pack2(octant_fix(pl, axl, ayl, tl, cl), octant_fix(ph, axh, ayh, th, ch));
    const f32x2 thp = fma2(c.eps_gamma_neg, Dn, theta);
    const f32x2 u = mul2(Dn, thp);
    const f32x2 u2 = mul2(u, u);
    f32x2 dl = dist;
    if (RADIUS) dl = sub2(sub2(dist, ri), rj);
    const f32x2 y = fma2(mul2(dl, Dinv), c.k_exp, c.log2A);
    const f32x2 e1 = ex2_2(fma2(c.c_nprime_neg, u2, y));
    const f32x2 e2 = ex2_2(fma2(c.c_n_neg, u2, y));
    const f32x2 a = mul2(e1, Dinv);
    // b carries sign(theta'): e2 * Dinv >= 0, so OR-ing theta's sign bits in is copysign (NaNs stay NaNs)
    const f32x2 b = mul2(e2, Dinv) | (thp & 0x8000000080000000ULL);
    acc.ax = fma2(a, Dx, acc.ax);
    acc.bx = fma2(b, Dy, acc.bx);
    acc.ay = fma2(a, Dy, acc.ay);
    acc.ay = fma2(b, Dx, acc.ay);
    acc.az = fma2(a, Dz, acc.az);
}

template <int IR, bool RADIUS>
__device__ __forceinline__ void tile_pairs_packed(const float (*__restrict__ tl)[K1_TJ], const f32x2 (&xi)[IR],
                                                  const f32x2 (&yi)[IR], const f32x2 (&zi)[IR], const f32x2 (&ri)[IR],
                                                  const f32x2 (&vxi)[IR], const f32x2 (&vyi)[IR],
                                                  const f32x2 (&vzi)[IR], const PackedConst& c, PackedAcc (&acc)[IR]) {
#pragma unroll 1
    for (int j = 0; j < K1_TJ; j += 4) {
        const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(&tl[PX][j]);
        const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(&tl[PY][j]);
        const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(&tl[PZ][j]);
        ulonglong2 R = make_ulonglong2(0ull, 0ull);
        if (RADIUS) R = *reinterpret_cast<const ulonglong2*>(&tl[PR][j]);
        const ulonglong2 VX = *reinterpret_cast<const ulonglong2*>(&tl[PVX][j]);
        const ulonglong2 VY = *reinterpret_cast<const ulonglong2*>(&tl[PVY][j]);
        const ulonglong2 VZ = *reinterpret_cast<const ulonglong2*>(&tl[PVZ][j]);
#pragma unroll
        for (int r = 0; r < IR; ++r) {
            pair_force2<RADIUS>(xi[r], yi[r], zi[r], ri[r], vxi[r], vyi[r], vzi[r], X.x, Y.x, Z.x, R.x, VX.x, VY.x, VZ.x,
                                c, acc[r]);
            pair_force2<RADIUS>(xi[r], yi[r], zi[r], ri[r], vxi[r], vyi[r], vzi[r], X.y, Y.y, Z.y, R.y, VX.y, VY.y, VZ.y,
                                c, acc[r]);
        }
    }
}

// planes:      [world][NPLANES][rows_pad] staged rows of all ranks (after the all-gather)
// own_block:   index of the rank block whose rows this launch computes forces for
// partial:     [gridDim.y][partial_stride] float4, (-> K3 sums over the splits)
template <int IR, bool RADIUS, int MINB>
__global__ void __launch_bounds__(K1_THREADS, MINB) k1_ped_pairs(const float* __restrict__ planes, const int rows_pad,
                                                           const int total_tiles, const int own_block,
                                                           float4* __restrict__ partial, const int partial_stride,
                                                           const PairParams pp) {
    __shared__ __align__(128) float tile[K1_STAGES][K1_PLANES][K1_TJ];
    __shared__ __align__(8) uint64_t bar[K1_STAGES];

    const int tid = threadIdx.x;
    const int rows_per_cta = K1_THREADS * IR;
    const int i_base = blockIdx.x * rows_per_cta;
    const int tiles_per_rank = rows_pad / K1_TJ;
    const int nsplit = gridDim.y, split = blockIdx.y;
    const int t_begin = (int)(((long long)total_tiles * split) / nsplit);
    const int t_end = (int)(((long long)total_tiles * (split + 1)) / nsplit);

    if (tid == 0) {
        for (int s = 0; s < K1_STAGES; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int t, int stage) {
        const int q = t / tiles_per_rank;
        const int off = (t - q * tiles_per_rank) * K1_TJ;
        const float* src = planes + ((size_t)q * NPLANES) * rows_pad + off;
        mbar_expect_tx(&bar[stage], K1_PLANES * K1_TJ * sizeof(float));
#pragma unroll
        for (int p = 0; p < K1_PLANES; ++p)
            bulk_copy_g2s(&tile[stage][p][0], src + (size_t)p * rows_pad, K1_TJ * sizeof(float), &bar[stage]);
    };
    if (tid == 0 && t_begin < t_end) issue(t_begin, 0);

    // this thread's rows
    const float* own = planes + ((size_t)own_block * NPLANES) * rows_pad;
    float xi[IR], yi[IR], zi[IR], ri[IR], vxi[IR], vyi[IR], vzi[IR];
    int self_j[IR];
    PairAcc acc[IR];
#pragma unroll
    for (int r = 0; r < IR; ++r) {
        const int row = i_base + r * K1_THREADS + tid;
        xi[r] = own[(size_t)PX * rows_pad + row];
        yi[r] = own[(size_t)PY * rows_pad + row];
        zi[r] = own[(size_t)PZ * rows_pad + row];
        ri[r] = own[(size_t)PR * rows_pad + row];
        vxi[r] = own[(size_t)PVX * rows_pad + row];
        vyi[r] = own[(size_t)PVY * rows_pad + row];
        vzi[r] = own[(size_t)PVZ * rows_pad + row];
        acc[r].gx = acc[r].gy = acc[r].gz = 0.0f;
    }
    // packed duplicates of the row data and the packed accumulators of the fast path
    const PackedConst pc = make_packed_const(pp);
    f32x2 xi2[IR], yi2[IR], zi2[IR], ri2[IR], vxi2[IR], vyi2[IR], vzi2[IR];
    PackedAcc acc2[IR];
#pragma unroll
    for (int r = 0; r < IR; ++r) {
        xi2[r] = splat2(xi[r]); yi2[r] = splat2(yi[r]); zi2[r] = splat2(zi[r]); ri2[r] = splat2(ri[r]);
        vxi2[r] = splat2(vxi[r]); vyi2[r] = splat2(vyi[r]); vzi2[r] = splat2(vzi[r]);
        acc2[r].ax = acc2[r].bx = acc2[r].ay = acc2[r].az = 0ull;
    }
    // the CTA's rows are K1_TJ-aligned blocks, so they meet the j range in rows_per_cta / K1_TJ diagonal tiles
    const int diag_first = (own_block * rows_pad + i_base) / K1_TJ;
    const int diag_last = (own_block * rows_pad + i_base + rows_per_cta - 1) / K1_TJ;

    for (int t = t_begin; t < t_end; ++t) {
        const int k = t - t_begin;
        const int stage = k & 1;
        if (tid == 0 && t + 1 < t_end) issue(t + 1, stage ^ 1);      // stage^1 was drained before the last barrier
        while (!mbar_try_wait(&bar[stage], (k >> 1) & 1)) {}
        if (t >= diag_first && t <= diag_last) {
#pragma unroll
            for (int r = 0; r < IR; ++r)
                self_j[r] = own_block * rows_pad + i_base + r * K1_THREADS + tid - t * K1_TJ;   // slot inside this tile
            tile_pairs<IR, RADIUS, true>(tile[stage], xi, yi, zi, ri, vxi, vyi, vzi, self_j, pp, acc);
        } else {
            tile_pairs_packed<IR, RADIUS>(tile[stage], xi2, yi2, zi2, ri2, vxi2, vyi2, vzi2, pc, acc2);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < IR; ++r) {
        const int row = i_base + r * K1_THREADS + tid;
        float axl, axh, bxl, bxh, ayl, ayh, azl, azh;
        unpack2(acc2[r].ax, axl, axh);
        unpack2(acc2[r].bx, bxl, bxh);
        unpack2(acc2[r].ay, ayl, ayh);
        unpack2(acc2[r].az, azl, azh);
        const float gx = acc[r].gx + ((axl + axh) - (bxl + bxh));
        const float gy = acc[r].gy + (ayl + ayh);
        const float gz = acc[r].gz + (azl + azh);
        partial[(size_t)split * partial_stride + row] = make_float4(-gx, -gy, -gz, 0.0f);
    }
}

// K1r -- reduce the per-split partial sums (fixed order, float64) and repair poisoned rows.
//
// A row whose sum is not finite met a degenerate pair in the unguarded packed path (or genuinely overflows); its warp
// recomputes it cooperatively with the guarded scalar pair_force, lanes striding over every staged slot.  Healthy crowds
// never take that branch, so the kernel is a plain N x nsplit x 16 B read.
struct ReduceArgs {
    const float* planes;
    int rows_pad, world, own_block, n_local;
    const float4* partial;
    int nsplit;
    double* f_ped;                  // [n_local][3]
    unsigned long long* fixup_rows; // running count of repaired rows (statistics)
    PairParams pp;
};

template <bool RADIUS>
__global__ void __launch_bounds__(256) k1_reduce_fixup(const ReduceArgs a) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = row < a.n_local;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    if (live) {
        for (int s = 0; s < a.nsplit; ++s) {
            const float4 p = a.partial[(size_t)s * a.rows_pad + row];
            sx += (double)p.x;
            sy += (double)p.y;
            sz += (double)p.z;
        }
    }
    const bool bad = live && !(isfinite(sx) && isfinite(sy) && isfinite(sz));
    unsigned mask = __ballot_sync(0xffffffffu, bad);
    while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const int r = __shfl_sync(0xffffffffu, row, src);
        const float* own = a.planes + ((size_t)a.own_block * NPLANES) * a.rows_pad;
        const float xi = own[(size_t)PX * a.rows_pad + r], yi = own[(size_t)PY * a.rows_pad + r];
        const float zi = own[(size_t)PZ * a.rows_pad + r], ri = own[(size_t)PR * a.rows_pad + r];
        const float vxi = own[(size_t)PVX * a.rows_pad + r], vyi = own[(size_t)PVY * a.rows_pad + r];
        const float vzi = own[(size_t)PVZ * a.rows_pad + r];
        const int islot = a.own_block * a.rows_pad + r;
        const int total = a.world * a.rows_pad;
        double gx = 0.0, gy = 0.0, gz = 0.0;
        for (int j = lane; j < total; j += 32) {
            const int q = j / a.rows_pad;
            const float* b = a.planes + ((size_t)q * NPLANES) * a.rows_pad + (j - q * a.rows_pad);
            PairAcc acc = {0.0f, 0.0f, 0.0f};
            pair_force<RADIUS, true>(xi, yi, zi, ri, vxi, vyi, vzi, b[(size_t)PX * a.rows_pad], b[(size_t)PY * a.rows_pad],
                                     b[(size_t)PZ * a.rows_pad], b[(size_t)PR * a.rows_pad], b[(size_t)PVX * a.rows_pad],
                                     b[(size_t)PVY * a.rows_pad], b[(size_t)PVZ * a.rows_pad], j == islot, a.pp, acc);
            gx += (double)acc.gx;
            gy += (double)acc.gy;
            gz += (double)acc.gz;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            gx += __shfl_xor_sync(0xffffffffu, gx, o);
            gy += __shfl_xor_sync(0xffffffffu, gy, o);
            gz += __shfl_xor_sync(0xffffffffu, gz, o);
        }
        if (lane == src) {
            sx = -gx;
            sy = -gy;
            sz = -gz;
            atomicAdd(a.fixup_rows, 1ull);
        }
    }
    if (live) {
        a.f_ped[3 * (size_t)row + 0] = sx;
        a.f_ped[3 * (size_t)row + 1] = sy;
        a.f_ped[3 * (size_t)row + 2] = sz;
    }
}

}  // namespace sfm
