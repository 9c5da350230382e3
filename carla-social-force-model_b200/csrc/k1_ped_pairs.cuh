// K1 -- all-pairs pedestrian interaction force (Moussaid et al. 2009), float32, sm_100a: the pair arithmetic shared by
// the symmetric kernel (k1_sym.cuh), its guarded diagonal / repair paths, and the PTX primitives they use.
//
// Replaces PedestrianForce._get_force (reference forces.py:74-117) together with stateutils.all_diffs / all_sums /
// normalize / angle_diff_2d (stateutils.py:32-128): for every ordered pair i != j
//     d = p_j - p_i (3-D), dist = |d|, e = d / dist, dl = dist - (r_i + r_j) [use_ped_radius]
//     D = lambda (v_i - v_j) + e, t = D / |D|, n = (-t_y, t_x, 0), B = gamma |D|
//     theta = angle_xy(e) - angle_xy(t) wrapped to (-pi, pi], theta' = theta - epsilon B
//     F_i += -A exp(-dl/B - (n' B theta')^2) t  -  A sign(theta') exp(-dl/B - (n B theta')^2) n
//
// Algebra used here (all exact identities, checked against the oracle in tests/):
//   * positions are staged as (hi, lo) float32 pairs, hi on a 2^-6 m lattice (sfm_common.cuh), and
//     d = (hi_j - hi_i) + (lo_j - lo_i): exact lattice difference + tiny remainder, one rounding relative to |d|;
//   * velocities are staged pre-multiplied by lambda: w = lambda v_i - lambda v_j, D = d * (1/dist) + w  (3 FFMA, e is
//     never materialised);
//   * theta = atan2(D_xy x d_xy, D_xy . d_xy), and D_xy x d_xy == w_xy x d_xy because e is parallel to d -- the
//     cancellation-free form;  atan2 is an octant-reduced degree-15 odd minimax polynomial (2.2e-7 relative);
//   * the force is accumulated against the *unnormalised* D with both coefficients scaled by 1/|D|;
//   * exp(x) = ex2(x log2 e), with -log2(e)/gamma, (n gamma)^2 log2 e and log2 A folded into constants, so each
//     exponent is one FFMA on top of the shared term.
// Degenerate pairs follow the reference's zero-safe normalisation (stateutils.py:88-90): |d| == 0 gives e = 0 and
// angle_xy(e) = atan2(0, 0) = 0; |D| == 0 gives t = 0, B = 0 and a vanishing contribution whenever dl > 0.
#pragma once

#include "sfm_common.cuh"

namespace sfm {

constexpr int K1_TJ = ROW_ALIGN;      // j rows per staged tile
constexpr int K1_STAGES = 2;
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

// One staged row (scalar copy): position hi / lo parts, radius, lambda * velocity.
struct RowF {
    float x, y, z, xl, yl, zl, r, vx, vy, vz;
};

__device__ __forceinline__ RowF load_row(const float* __restrict__ block, const size_t stride, const size_t row) {
    RowF o;
    o.x = block[(size_t)PX * stride + row]; o.y = block[(size_t)PY * stride + row]; o.z = block[(size_t)PZ * stride + row];
    o.xl = block[(size_t)PXL * stride + row]; o.yl = block[(size_t)PYL * stride + row];
    o.zl = block[(size_t)PZL * stride + row];
    o.r = block[(size_t)PR * stride + row];
    o.vx = block[(size_t)PVX * stride + row]; o.vy = block[(size_t)PVY * stride + row];
    o.vz = block[(size_t)PVZ * stride + row];
    return o;
}

// One ordered pair (i <- j), guarded: numpy's zero-safe semantics for |d| = 0, |D| = 0, theta' = 0.  `self` pairs (same
// staged slot) are removed like the reference removes the diagonal (stateutils.py:44): their coefficients are zeroed.
template <bool RADIUS, bool DIAG>
__device__ __forceinline__ void pair_force(const RowF& I, const RowF& J, const bool self, const PairParams& pp,
                                           PairAcc& acc) {
    const float dx = (J.x - I.x) + (J.xl - I.xl), dy = (J.y - I.y) + (J.yl - I.yl), dz = (J.z - I.z) + (J.zl - I.zl);
    const float dxy2 = fmaf(dy, dy, dx * dx);
    const float d2 = fmaf(dz, dz, dxy2);
    const float rinv = rsqrt_approx(fmaxf(d2, K1_TINY));
    const float dist = d2 * rinv;
    const float wx = I.vx - J.vx, wy = I.vy - J.vy, wz = I.vz - J.vz;
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
    if (RADIUS) dl = dist - (I.r + J.r);
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

// The guarded code over one staged tile (the diagonal tile of the symmetric kernel: self pairs removed).
template <int IR, bool RADIUS, bool DIAG>
__device__ __forceinline__ void tile_pairs(const float (*__restrict__ tl)[K1_TJ], const RowF (&I)[IR],
                                           const int (&self_j)[IR], const PairParams& pp, PairAcc (&acc)[IR]) {
#pragma unroll 1
    for (int j = 0; j < K1_TJ; j += 4) {
        float q[10][4];
        constexpr int plane[10] = {PX, PY, PZ, PXL, PYL, PZL, PR, PVX, PVY, PVZ};
#pragma unroll
        for (int p = 0; p < 10; ++p) {
            const float4 v = *reinterpret_cast<const float4*>(&tl[plane[p]][j]);
            q[p][0] = v.x; q[p][1] = v.y; q[p][2] = v.z; q[p][3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const RowF J = {q[0][k], q[1][k], q[2][k], q[3][k], q[4][k], q[5][k], q[6][k], q[7][k], q[8][k], q[9][k]};
#pragma unroll
            for (int r = 0; r < IR; ++r) pair_force<RADIUS, DIAG>(I[r], J, DIAG && (self_j[r] == j + k), pp, acc[r]);
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

}  // namespace sfm
