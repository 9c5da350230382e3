// K1s -- all-pairs pedestrian force evaluated once per UNORDERED pair (Newton's third law), float32 packed, sm_100a.
//
// The reference's pair force (forces.py:74-117) is exactly antisymmetric: swapping i and j negates d = p_j - p_i and
// w = lambda (v_i - v_j), hence D = d/|d| + w -> -D, while |d|, |D|, B = gamma |D|, the radius term r_i + r_j and the
// interaction angle theta = atan2(D_xy x d_xy, D_xy . d_xy) are unchanged; so f_ji = -f_ij.  Each tile pair {I, J} of
// 256-row tiles is therefore visited once ("half shell": tile I takes the partners I+1 .. I+T/2 cyclically) and the
// result is added to the rows of I and subtracted from the rows of J.
//
// * Thread t of the CTA owns 4 rows of tile I (registers).  Lanes are STAGGERED over the partner tile (sym_tile: four
//   phases, the lower half-warp on the 16 quads of run ph, the upper on run ph + 2, lane l on quad (k + l) mod 16), so
//   the J-side accumulation (a per-warp shared-memory array, read-modify-write of 2-3 float4 per step) never sees two
//   lanes of a warp on the same j; no shared-memory atomics.
// * Two forms of d = p_j - p_i.  Double-single, d = (hi_j - hi_i) + (lo_j - lo_i): one float32 rounding relative to |d|
//   wherever the crowd sits -- 3 packed operations per coordinate.  Run-local, d = xr_j - m_i with xr_j the partner's
//   position relative to the origin of its 64-row run and m_i = (hi_i - c_run) + lo_i kept in registers for the 16 steps
//   a lane spends on that run -- 1 operation per coordinate, float32 rounding at the scale of the run's extent (~5e-7 m).
//   The choice is per tile pair and purely geometric (sfm_common.cuh): run-local only for partner tiles whose bounding
//   box is >= max(1 m, a quarter of their widest run's half-extent) away from tile I's -- never for a pair close enough
//   to carry much force, and always with |d| large against the rounding; the staged slot order (k8_order.cuh) makes that
//   the 93-99 % of the tile pairs that are not neighbours.
// * Tile partials (256 terms per row, float32) are converted to 64-bit fixed point (2^-32 m/s^2) and accumulated with
//   integer atomics: integer addition is associative, so the result does not depend on CTA scheduling, on the launch
//   geometry or on how many GPUs share the crowd -- and the multi-GPU exchange is an integer reduce-scatter.
// * The diagonal tile {I, I} runs the guarded asymmetric code (self pairs removed).  A non-finite or out-of-range tile
//   partial (degenerate pair) bumps the row's poison counter instead; k1_sym_finish recomputes
//   poisoned rows with the guarded scalar path.
#pragma once

#include "k1_ped_pairs.cuh"

namespace sfm {

#ifndef SFM_KS_IR
#define SFM_KS_IR 4
#endif
#ifndef SFM_KS_ASIN
#define SFM_KS_ASIN 1           // planar tiles: interaction angle from sin/cos (no MUFU.RCP) with an asin polynomial
#endif
#ifndef SFM_KS_ASIN_TERMS
#define SFM_KS_ASIN_TERMS 7
#endif
#ifndef SFM_KS_ANGLE
#define SFM_KS_ANGLE 2          // planar tiles: 0 = octant reduction on min(|sin|, |cos|); 1, 2 = first-quadrant angle from
#endif                          //   (|sin| - |cos|) / sqrt 2 (no min / compare / select), signs by 1: one product + two
                                //   transfers, 2: two transfers (profiles/tune_k1s_r2*.log: 3.95 -> 3.83 -> 3.75 ms)
#ifndef SFM_KS_NEGSUB
#define SFM_KS_NEGSUB 1         // a*b - c*d written as sub2(mul2, mul2): no 64-bit XOR negations (2 LOP3 each)
#endif
#ifndef SFM_KS_UNROLL
#define SFM_KS_UNROLL 1
#endif
#ifndef SFM_KS_MINB
#define SFM_KS_MINB 4           // min CTAs per SM handed to __launch_bounds__ (register cap 128: no spills; 5 -> 96 registers
                                // spills the fixed-point accumulators and is 1.5-4 % slower, profiles/tune_k1s_r2*.log)
#endif
#ifndef SFM_KS_POLY
#define SFM_KS_POLY 0           // asin polynomial: 0 Horner, 1 Estrin, 2 even / odd split
#endif
#ifndef SFM_KS_SIGNS
#define SFM_KS_SIGNS 0          // 1: theta' = |theta| - copysign(eps gamma, cross) |D| (transfer off the polynomial's chain),
#endif                          //    sign(theta') = sign(cross) sign(that) by one product
#ifndef SFM_KS_ABLATE
#define SFM_KS_ABLATE 0         // timing-only experiments (WRONG results): bit 0 no sign transfers, 1 no ex2, 2 short polynomial,
#endif                          //   3 no lo parts, 4 no J-side read-modify-write, 5 no __syncwarp, 6 no rsqrt
constexpr int KS_UNROLL = SFM_KS_UNROLL;                 // unroll depth of the j-quad loop
constexpr int KS_IR = SFM_KS_IR;                         // rows per thread
constexpr int KS_THREADS = K1_TJ / KS_IR;                // one 256-row tile per CTA
constexpr int KS_WARPS = KS_THREADS / 32;
constexpr int KS_QUADS = K1_TJ / 4;
constexpr int KS_PLANES = NPLANES;                       // every staged plane, the per-row non-planar flag included
constexpr float KS_FIXED_SCALE = 4294967296.0f;          // 2^32 counts per m/s^2
constexpr float KS_FIXED_LIMIT = 1.0e9f;                 // |partial| beyond this goes to the repair path

struct SymArgs {
    const float* planes;       // [world][NPLANES][rows_pad]
    int rows_pad;
    int total_tiles;           // world * rows_pad / 256
    int own_first_tile;        // first tile of this rank's rows
    long long* facc;           // [total_slots][4]: fixed-point force x, y, z and the poison counter
    PairParams pp;
    int use_local;             // 0: every tile takes the double-single path (SFM_K1_LOCAL=0)
    unsigned long long* local_pairs;   // [1] tile pairs that took the local path (sfm_stats.local_tile_pairs)
};

// asin(x) = x P(x^2) on [0, sin(pi/4)]: minimax fits, max abs error 6.2e-7 (6 terms) / 1.0e-7 (7 terms)
struct AsinConst {
    f32x2 s0, s1, s2, s3, s4, s5, s6;
};
template <bool DIFF_FORM>
__device__ __forceinline__ AsinConst make_asin_const() {
    AsinConst c;
#if SFM_KS_ASIN_TERMS == 6
    c.s0 = splat2(9.999993443e-01f); c.s1 = splat2(1.667610258e-01f); c.s2 = splat2(7.292483002e-02f);
    c.s3 = splat2(6.089710817e-02f); c.s4 = splat2(-2.424139343e-02f); c.s5 = splat2(9.637616575e-02f);
    c.s6 = 0ull;
#else
    c.s0 = splat2(1.000000119e+00f); c.s1 = splat2(1.666489840e-01f); c.s2 = splat2(7.553865016e-02f);
    c.s3 = splat2(3.859527037e-02f); c.s4 = splat2(6.176946312e-02f); c.s5 = splat2(-5.651333556e-02f);
    c.s6 = splat2(1.019138768e-01f);
#endif
    if (DIFF_FORM) {
        // asin(t / sqrt 2) = t Q(t^2): coefficient k scaled by 2^-(k + 1/2)
        f32x2* cs[7] = {&c.s0, &c.s1, &c.s2, &c.s3, &c.s4, &c.s5, &c.s6};
        float scale = 0.70710678118654752f;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            float lo, hi;
            unpack2(*cs[k], lo, hi);
            *cs[k] = splat2(lo * scale);
            scale *= 0.5f;
        }
    }
    return c;
}

// -f_ij for two consecutive j at once: g = (a Dx - b Dy, a Dy + b Dx, a Dz); F_i -= g, F_j += g.
// One row of tile I held by a thread: every value splat over both halves of a packed register.
struct RowP {
    f32x2 x, y, z, xl, yl, zl, r, vx, vy, vz;
};

__device__ __forceinline__ RowP splat_row(const RowF& f) {
    RowP o;
    o.x = splat2(f.x); o.y = splat2(f.y); o.z = splat2(f.z); o.xl = splat2(f.xl); o.yl = splat2(f.yl); o.zl = splat2(f.zl);
    o.r = splat2(f.r); o.vx = splat2(f.vx); o.vy = splat2(f.vy); o.vz = splat2(f.vz);
    return o;
}

// SIGN0: reproduce np.sign(0) = 0 (forces.py:108) in the fast path -- only instantiated for epsilon == 0, where theta' = 0
// is reached by every pair with w parallel to d (e.g. a standing crowd); with epsilon != 0 the event has measure zero.
// LOCAL: the partner tile is read through the origins of its 64-row runs (sfm_common.cuh) -- xj, yj, zj are the XR / YR /
// ZR planes and I.x, I.y, I.z hold m = (hi_i - c) + lo_i for the run the lane is working on: d = xr_j - m_i, one
// subtraction per coordinate where the double-single form takes three.
template <bool RADIUS, bool PLANAR, bool SIGN0, bool LOCAL>
__device__ __forceinline__ void pair_terms2(const RowP& I, const f32x2 xj, const f32x2 yj, const f32x2 zj,
                                            const f32x2 xlj, const f32x2 ylj, const f32x2 zlj, const f32x2 rj,
                                            const f32x2 vxj, const f32x2 vyj, const f32x2 vzj, const PackedConst& c,
                                            const AsinConst& sc, f32x2& gx, f32x2& gy, f32x2& gz) {
    // d = (hi_j - hi_i) + (lo_j - lo_i): exact lattice difference + remainder (sfm_common.cuh)
    // PLANAR: every pedestrian of both tiles has the same z and no vertical velocity, so d_z = w_z = D_z = 0 exactly
#if SFM_KS_ABLATE & 8
    const f32x2 dx = sub2(xj, I.x), dy = sub2(yj, I.y);
#else
    const f32x2 dx = LOCAL ? sub2(xj, I.x) : add2(sub2(xj, I.x), sub2(xlj, I.xl));
    const f32x2 dy = LOCAL ? sub2(yj, I.y) : add2(sub2(yj, I.y), sub2(ylj, I.yl));
#endif
    f32x2 dz = 0ull, wz = 0ull, Dz = 0ull;
    f32x2 d2 = fma2(dy, dy, mul2(dx, dx));
    const f32x2 dxy2 = d2;
    if (!PLANAR) {
        dz = LOCAL ? sub2(zj, I.z) : add2(sub2(zj, I.z), sub2(zlj, I.zl));
        d2 = fma2(dz, dz, d2);
    }
#if SFM_KS_ABLATE & 64
    const f32x2 rinv = fma2(d2, splat2(1e-3f), splat2(0.5f));
#else
    const f32x2 rinv = rsqrt2(d2);
#endif
    const f32x2 wx = sub2(I.vx, vxj), wy = sub2(I.vy, vyj);
    const f32x2 Dx = fma2(dx, rinv, wx), Dy = fma2(dy, rinv, wy);
    f32x2 D2 = fma2(Dy, Dy, mul2(Dx, Dx));
    const f32x2 Dxy2 = D2;
    if (!PLANAR) {
        wz = sub2(I.vz, vzj);
        Dz = fma2(dz, rinv, wz);
        D2 = fma2(Dz, Dz, D2);
    }
#if SFM_KS_ABLATE & 64
    const f32x2 Dinv = fma2(D2, splat2(1e-3f), splat2(0.5f));
#else
    const f32x2 Dinv = rsqrt2(D2);
#endif
    const f32x2 Dn = mul2(D2, Dinv);
#if SFM_KS_NEGSUB
    const f32x2 cross = sub2(mul2(wx, dy), mul2(wy, dx));
#else
    const f32x2 cross = fma2(wx, dy, neg2(mul2(wy, dx)));
#endif
    const f32x2 dot = fma2(Dx, dx, mul2(Dy, dy));
    float cl, ch, tl, th;
    unpack2(cross, cl, ch);
    unpack2(dot, tl, th);
    constexpr bool ASIN = PLANAR && (SFM_KS_ASIN != 0);
    constexpr bool DIFF = (SFM_KS_ASIN != 0) && (SFM_KS_ANGLE != 0) && !SIGN0;
    f32x2 R = 0ull, theta;
    // (SIGN0 keeps the octant form: it returns theta = +-0 exactly for w parallel to d, which the difference form cannot)
    if (DIFF) {
        // first-quadrant angle phi = atan2(|sin|, |cos|) from sqrt2 sin(phi - pi/4) = |sin| - |cos|; then
        // theta = sign(cross) (pi/2 + sign(dot) (phi - pi/2)) = copysign(pi/2, cross) + sign(cross dot) (phi - pi/2).
        // sin, cos = cross, dot / (|d_xy| |D_xy|): planar tiles have |d_xy| = |d|, |D_xy| = |D| (no extra MUFU); general
        // tiles take one rsqrt of the product of the xy parts -- where the reciprocal of the octant form used to be
        R = mul2(rinv, Dinv);
        const f32x2 Rxy = PLANAR ? R : rsqrt2(mul2(dxy2, Dxy2));
        const f32x2 q = mul2(pack2(fabsf(cl) - fabsf(tl), fabsf(ch) - fabsf(th)), Rxy);
        const f32x2 s = mul2(q, q);
        f32x2 p;
#if SFM_KS_POLY == 1
        // Estrin: 8 operations, depth 3 (Horner: 6 operations, depth 6)
        {
            const f32x2 s2 = mul2(s, s);
            const f32x2 pa = fma2(sc.s1, s, sc.s0), pb = fma2(sc.s3, s, sc.s2), pcc = fma2(sc.s5, s, sc.s4);
            const f32x2 s4 = mul2(s2, s2);
            const f32x2 pe = fma2(sc.s6, s2, pcc), pf = fma2(pb, s2, pa);
            p = fma2(pe, s4, pf);
        }
#elif SFM_KS_POLY == 2
        // even / odd split: 7 operations, depth 5
        {
            const f32x2 s2 = mul2(s, s);
            f32x2 pe = fma2(sc.s6, s2, sc.s4);
            f32x2 po = fma2(sc.s5, s2, sc.s3);
            pe = fma2(pe, s2, sc.s2);
            po = fma2(po, s2, sc.s1);
            pe = fma2(pe, s2, sc.s0);
            p = fma2(po, s, pe);
        }
#else
#if SFM_KS_ASIN_TERMS == 6
        p = fma2(sc.s5, s, sc.s4);
#else
        p = fma2(sc.s6, s, sc.s5);
        p = fma2(p, s, sc.s4);
#endif
#if !(SFM_KS_ABLATE & 4)
        p = fma2(p, s, sc.s3);
        p = fma2(p, s, sc.s2);
        p = fma2(p, s, sc.s1);
#endif
        p = fma2(p, s, sc.s0);
#endif
        const f32x2 g = fma2(p, q, splat2(-0.78539816339744831f));                 // phi - pi/2 in [-pi/2, 0]
#if SFM_KS_ANGLE == 1
        const f32x2 mg = g ^ (mul2(cross, dot) & 0x8000000080000000ULL);
        const f32x2 h = (cross & 0x8000000080000000ULL) | splat2(1.57079632679489662f);
        theta = add2(h, mg);
#elif SFM_KS_ABLATE & 1
        theta = add2(g, splat2(1.57079632679489662f));
#elif SFM_KS_SIGNS == 1
        theta = add2(g ^ (dot & 0x8000000080000000ULL), splat2(1.57079632679489662f));   // |theta| in [0, pi]
#else
        const f32x2 tabs = add2(g ^ (dot & 0x8000000080000000ULL), splat2(1.57079632679489662f));   // |theta| in [0, pi]
        theta = tabs ^ (cross & 0x8000000080000000ULL);
#endif
    } else {
        const float axl = fabsf(tl), ayl = fabsf(cl), axh = fabsf(th), ayh = fabsf(ch);
        f32x2 q, p;
        if (ASIN) {
            // planar: |D_xy| = |D| and |d_xy| = |d|, so sin/cos(theta) = cross/dot * (1/|d|)(1/|D|) -- no reciprocal
            R = mul2(rinv, Dinv);
            q = mul2(pack2(fminf(axl, ayl), fminf(axh, ayh)), R);
            const f32x2 s = mul2(q, q);
#if SFM_KS_ASIN_TERMS == 6
            p = fma2(sc.s5, s, sc.s4);
#else
            p = fma2(sc.s6, s, sc.s5);
            p = fma2(p, s, sc.s4);
#endif
            p = fma2(p, s, sc.s3);
            p = fma2(p, s, sc.s2);
            p = fma2(p, s, sc.s1);
            p = fma2(p, s, sc.s0);
        } else {
            const f32x2 mxr = pack2(rcp_approx(fmaxf(axl, ayl)), rcp_approx(fmaxf(axh, ayh)));
            q = mul2(pack2(fminf(axl, ayl), fminf(axh, ayh)), mxr);
            const f32x2 s = mul2(q, q);
            p = fma2(c.a7, s, c.a6);
            p = fma2(p, s, c.a5);
            p = fma2(p, s, c.a4);
            p = fma2(p, s, c.a3);
            p = fma2(p, s, c.a2);
            p = fma2(p, s, c.a1);
            p = fma2(p, s, c.a0);
        }
        p = mul2(p, q);
        float pl, ph;
        unpack2(p, pl, ph);
        theta = pack2(octant_fix(pl, axl, ayl, tl, cl), octant_fix(ph, axh, ayh, th, ch));
    }
#if SFM_KS_SIGNS == 1
    // theta = sigma |theta| (sigma = sign(cross)): theta' = sigma (|theta| - sigma eps gamma |D|); its square is sign-free
    const f32x2 thq = (DIFF) ? fma2(c.eps_gamma_neg ^ (cross & 0x8000000080000000ULL), Dn, theta) : 0ull;
    const f32x2 thp = (DIFF) ? mul2(thq, cross) : fma2(c.eps_gamma_neg, Dn, theta);      // DIFF: only its sign is used below
    const f32x2 u = mul2(Dn, (DIFF) ? thq : thp);
#else
    const f32x2 thp = fma2(c.eps_gamma_neg, Dn, theta);
    const f32x2 u = mul2(Dn, thp);
#endif
    const f32x2 u2 = mul2(u, u);
    f32x2 y;
    if (RADIUS) {
        const f32x2 dl = sub2(sub2(mul2(d2, rinv), I.r), rj);
        y = fma2(mul2(dl, Dinv), c.k_exp, c.log2A);
    } else if (ASIN || DIFF) {
        y = fma2(mul2(d2, R), c.k_exp, c.log2A);                          // |d| / |D| = d2 (1/|d|)(1/|D|)
    } else {
        y = fma2(mul2(mul2(d2, rinv), Dinv), c.k_exp, c.log2A);
    }
#if SFM_KS_ABLATE & 2
    const f32x2 e1 = fma2(c.c_nprime_neg, u2, y);
    const f32x2 e2 = fma2(c.c_n_neg, u2, y);
#else
    const f32x2 e1 = ex2_2(fma2(c.c_nprime_neg, u2, y));
    const f32x2 e2 = ex2_2(fma2(c.c_n_neg, u2, y));
#endif
    const f32x2 a = mul2(e1, Dinv);
#if SFM_KS_ABLATE & 1
    f32x2 b = mul2(e2, Dinv);
#else
    f32x2 b = mul2(e2, Dinv) | (thp & 0x8000000080000000ULL);            // copysign(e2 / |D|, theta')
#endif
    if (SIGN0) {
        float bl, bh, hl, hh;
        unpack2(b, bl, bh);
        unpack2(thp, hl, hh);
        b = pack2(hl == 0.0f ? 0.0f : bl, hh == 0.0f ? 0.0f : bh);
    }
#if SFM_KS_NEGSUB
    gx = sub2(mul2(a, Dx), mul2(b, Dy));        // ptxas folds the subtraction into FFMA2 with a negated addend
#else
    gx = fma2(a, Dx, mul2(neg2(b), Dy));
#endif
    gy = fma2(a, Dy, mul2(b, Dx));
    gz = PLANAR ? 0ull : mul2(a, Dz);
}

__device__ __forceinline__ bool fixed_ok(float v) { return fabsf(v) < KS_FIXED_LIMIT; }     // false for NaN / inf too
__device__ __forceinline__ long long to_fixed(float v) { return __float2ll_rn(v * KS_FIXED_SCALE); }

// One partner tile against this thread's rows: accumulates -F_i partials in registers (returned in gi) and +g into the
// warp's private J-side slice.  Lanes are staggered over the j-quads so no two lanes of a warp touch the same j: the tile
// is walked in four phases, and in a phase the lower half-warp works on the 16 quads of run `ph` of the tile while the
// upper half-warp works on run `ph + 2` (lane l on quad (k + l) mod 16 of its run at step k).  A lane therefore stays on
// one 64-row run for 16 steps, which is what lets the LOCAL path keep m_i = (hi_i - c_run) + lo_i in registers.
template <bool RADIUS, bool PLANAR, bool SIGN0, bool LOCAL>
__device__ __forceinline__ void sym_tile(const float (*__restrict__ tl)[K1_TJ], float (*__restrict__ accw)[K1_TJ],
                                         const int lane, const RowF (&If)[KS_IR], const RowP (&I)[KS_IR],
                                         const PackedConst& pc, const AsinConst& sc, float (&gi)[KS_IR][3]) {
    f32x2 Gx[KS_IR], Gy[KS_IR], Gz[KS_IR];
#pragma unroll
    for (int r = 0; r < KS_IR; ++r) Gx[r] = Gy[r] = Gz[r] = 0ull;
    const ulonglong2 zero = make_ulonglong2(0ull, 0ull);
    constexpr int RUN_QUADS = SUB_ROWS / 4;                          // 16 quads per run
    const int half = lane >> 4, l16 = lane & (RUN_QUADS - 1);
#pragma unroll 1
    for (int ph = 0; ph < SUBS_PER_TILE; ++ph) {
        const int run = (ph + 2 * half) & (SUBS_PER_TILE - 1);
        RowP Iw[KS_IR];
#pragma unroll
        for (int r = 0; r < KS_IR; ++r) Iw[r] = I[r];
        if (LOCAL) {
            const float4 meta = *reinterpret_cast<const float4*>(&tl[PMETA][4 * run]);     // (c_x, c_y, c_z, half-extent)
#pragma unroll
            for (int r = 0; r < KS_IR; ++r) {
                Iw[r].x = splat2((If[r].x - meta.x) + If[r].xl);     // hi_i - c: exact (lattice points)
                Iw[r].y = splat2((If[r].y - meta.y) + If[r].yl);
                if (!PLANAR) Iw[r].z = splat2((If[r].z - meta.z) + If[r].zl);
            }
        }
#pragma unroll KS_UNROLL
        for (int step = 0; step < RUN_QUADS; ++step) {
            const int j = (run * RUN_QUADS + ((step + l16) & (RUN_QUADS - 1))) * 4;   // staggered: distinct j per lane
            auto ld = [&](int plane) { return *reinterpret_cast<const ulonglong2*>(&tl[plane][j]); };
            const ulonglong2 X = ld(LOCAL ? PXR : PX), Y = ld(LOCAL ? PYR : PY);
            const ulonglong2 XL = LOCAL ? zero : ld(PXL), YL = LOCAL ? zero : ld(PYL);
            const ulonglong2 Z = PLANAR ? zero : ld(LOCAL ? PZR : PZ), ZL = (PLANAR || LOCAL) ? zero : ld(PZL);
            const ulonglong2 VZ = PLANAR ? zero : ld(PVZ);
            const ulonglong2 R = RADIUS ? ld(PR) : zero;
            const ulonglong2 VX = ld(PVX), VY = ld(PVY);
            f32x2 jx0 = 0ull, jy0 = 0ull, jz0 = 0ull, jx1 = 0ull, jy1 = 0ull, jz1 = 0ull;   // sum over my rows
#pragma unroll
            for (int r = 0; r < KS_IR; ++r) {
                f32x2 gx, gy, gz;
                pair_terms2<RADIUS, PLANAR, SIGN0, LOCAL>(Iw[r], X.x, Y.x, Z.x, XL.x, YL.x, ZL.x, R.x, VX.x, VY.x, VZ.x, pc,
                                                          sc, gx, gy, gz);
                Gx[r] = add2(Gx[r], gx); Gy[r] = add2(Gy[r], gy);
                jx0 = r ? add2(jx0, gx) : gx; jy0 = r ? add2(jy0, gy) : gy;
                if (!PLANAR) { Gz[r] = add2(Gz[r], gz); jz0 = r ? add2(jz0, gz) : gz; }
                pair_terms2<RADIUS, PLANAR, SIGN0, LOCAL>(Iw[r], X.y, Y.y, Z.y, XL.y, YL.y, ZL.y, R.y, VX.y, VY.y, VZ.y, pc,
                                                          sc, gx, gy, gz);
                Gx[r] = add2(Gx[r], gx); Gy[r] = add2(Gy[r], gy);
                jx1 = r ? add2(jx1, gx) : gx; jy1 = r ? add2(jy1, gy) : gy;
                if (!PLANAR) { Gz[r] = add2(Gz[r], gz); jz1 = r ? add2(jz1, gz) : gz; }
            }
            // J side: F_j += g, into this warp's private slice (lanes hold distinct j, so plain read-modify-write)
#if SFM_KS_ABLATE & 16
            Gx[0] = add2(Gx[0], add2(jx0, jx1)); Gy[0] = add2(Gy[0], add2(jy0, jy1));
#else
            ulonglong2* ax = reinterpret_cast<ulonglong2*>(&accw[0][j]);
            ulonglong2* ay = reinterpret_cast<ulonglong2*>(&accw[1][j]);
            ulonglong2 vx = *ax, vy = *ay;
            vx.x = add2(vx.x, jx0); vx.y = add2(vx.y, jx1);
            vy.x = add2(vy.x, jy0); vy.y = add2(vy.y, jy1);
            *ax = vx; *ay = vy;
#endif
            if (!PLANAR) {
                ulonglong2* az = reinterpret_cast<ulonglong2*>(&accw[2][j]);
                ulonglong2 vz = *az;
                vz.x = add2(vz.x, jz0); vz.y = add2(vz.y, jz1);
                *az = vz;
            }
#if !(SFM_KS_ABLATE & 32)
            __syncwarp();                               // next step another lane owns this quad
#endif
        }
    }
#pragma unroll
    for (int r = 0; r < KS_IR; ++r) {
        float lo, hi;
        unpack2(Gx[r], lo, hi); gi[r][0] = lo + hi;
        unpack2(Gy[r], lo, hi); gi[r][1] = lo + hi;
        unpack2(Gz[r], lo, hi); gi[r][2] = lo + hi;
    }
}

#ifdef SFM_KS_MAXNREG
#define SFM_KS_BOUNDS __maxnreg__(SFM_KS_MAXNREG)
#else
#define SFM_KS_BOUNDS __launch_bounds__(KS_THREADS, SFM_KS_MINB)
#endif
template <bool RADIUS, bool SIGN0>
__global__ void SFM_KS_BOUNDS k1_sym_pairs(const SymArgs a) {
    __shared__ __align__(128) float tile[K1_STAGES][KS_PLANES][K1_TJ];
    __shared__ __align__(16) float accj[KS_WARPS][3][K1_TJ];
    __shared__ __align__(8) uint64_t bar[K1_STAGES];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int T = a.total_tiles;
    const int I = a.own_first_tile + blockIdx.x;
    const int tiles_per_rank = a.rows_pad / K1_TJ;
    // half shell: partners I+1 .. I+H (cyclic); for even T the opposite tile belongs to the lower half only
    const int H = (T & 1) ? (T - 1) / 2 : ((I < T / 2) ? T / 2 : T / 2 - 1);
    const int nsplit = gridDim.y, split = blockIdx.y;
    const int m_begin = (int)(((long long)H * split) / nsplit), m_end = (int)(((long long)H * (split + 1)) / nsplit);
    const int has_diag = (split == 0) ? 1 : 0;
    const int n_items = has_diag + (m_end - m_begin);
    if (n_items == 0) return;
    auto item_tile = [&](int k) {
        if (has_diag && k == 0) return I;
        int j = I + 1 + m_begin + (k - has_diag);
        return (j >= T) ? j - T : j;
    };

    if (tid == 0) {
        for (int s = 0; s < K1_STAGES; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    for (int e = tid; e < KS_WARPS * 3 * K1_TJ; e += KS_THREADS) (&accj[0][0][0])[e] = 0.0f;
    __syncthreads();

    auto issue = [&](int t, int stage) {
        const int q = t / tiles_per_rank;
        const int off = (t - q * tiles_per_rank) * K1_TJ;
        const float* src = a.planes + ((size_t)q * NPLANES) * a.rows_pad + off;
        mbar_expect_tx(&bar[stage], KS_PLANES * K1_TJ * sizeof(float));
#pragma unroll
        for (int p = 0; p < KS_PLANES; ++p)
            bulk_copy_g2s(&tile[stage][p][0], src + (size_t)p * a.rows_pad, K1_TJ * sizeof(float), &bar[stage]);
    };
    if (tid == 0) issue(item_tile(0), 0);

    // this thread's rows of tile I
    const int qi = I / tiles_per_rank;
    const float* own = a.planes + ((size_t)qi * NPLANES) * a.rows_pad + (size_t)(I - qi * tiles_per_rank) * K1_TJ;
    RowF rowf[KS_IR];
    RowP rowp[KS_IR];
    long long fix[KS_IR][3];
    int bad[KS_IR];
#pragma unroll
    for (int r = 0; r < KS_IR; ++r) {
        rowf[r] = load_row(own, (size_t)a.rows_pad, (size_t)(r * KS_THREADS + tid));
        rowp[r] = splat_row(rowf[r]);
        fix[r][0] = fix[r][1] = fix[r][2] = 0;
        bad[r] = 0;
    }
    const PackedConst pc = make_packed_const(a.pp);
    const AsinConst sc = make_asin_const<(SFM_KS_ANGLE != 0) && !SIGN0>();
    // planar fast path: K3 flags every staged row whose z (relative to the origin) or vertical velocity is non-zero
    int own_flag = 0;
#pragma unroll
    for (int r = 0; r < KS_IR; ++r) own_flag |= (own[(size_t)PFLAG * a.rows_pad + r * KS_THREADS + tid] != 0.0f);
    const bool planar_own = __syncthreads_or(own_flag) == 0;

    float own_box[4];                            // xy bounding box of tile I (PMETA, written by K3)
#pragma unroll
    for (int k = 0; k < 4; ++k) own_box[k] = own[(size_t)PMETA * a.rows_pad + META_BOX + k];
    int n_local = 0;
    for (int k = 0; k < n_items; ++k) {
        const int stage = k & 1;
        const int J = item_tile(k);
        if (tid == 0 && k + 1 < n_items) issue(item_tile(k + 1), stage ^ 1);   // stage^1 was released by the barriers below
        while (!mbar_try_wait(&bar[stage], (k >> 1) & 1)) {}
        float gi[KS_IR][3];                                   // this tile's -F_i partial per row
        int flag_j = 0;
#pragma unroll
        for (int r = 0; r < KS_IR; ++r) flag_j |= (tile[stage][PFLAG][tid + r * KS_THREADS] != 0.0f);
        const bool tile_nonplanar = __syncthreads_or(flag_j) != 0;
        // local path: the two tiles' bounding boxes are at least max(LOCAL_SEP, LOCAL_SEP_FACTOR * ext) apart, ext the
        // largest half-extent of the partner tile's four runs (+inf for a run that does not qualify; K3 decides, per
        // tick) -- every pair closer than that stays double-single (sfm_common.cuh)
        const float* meta = tile[stage][PMETA];
        const float sep = fmaxf(fmaxf(meta[META_BOX] - own_box[2], own_box[0] - meta[META_BOX + 2]),
                                fmaxf(meta[META_BOX + 1] - own_box[3], own_box[1] - meta[META_BOX + 3]));
        const float ext = fmaxf(fmaxf(meta[3], meta[7]), fmaxf(meta[11], meta[15]));
        const bool tile_local = a.use_local && sep >= fmaxf(LOCAL_SEP, LOCAL_SEP_FACTOR * ext);
        if (J == I) {
            // diagonal tile: guarded asymmetric evaluation with self pairs removed, rows of I only
            PairAcc acc[KS_IR];
            int self_j[KS_IR];
#pragma unroll
            for (int r = 0; r < KS_IR; ++r) {
                acc[r].gx = acc[r].gy = acc[r].gz = 0.0f;
                self_j[r] = r * KS_THREADS + tid;
            }
            tile_pairs<KS_IR, RADIUS, true>(tile[stage], rowf, self_j, a.pp, acc);
#pragma unroll
            for (int r = 0; r < KS_IR; ++r) { gi[r][0] = acc[r].gx; gi[r][1] = acc[r].gy; gi[r][2] = acc[r].gz; }
            __syncthreads();
        } else {
            const bool planar = planar_own && !tile_nonplanar;
            if (planar && tile_local) sym_tile<RADIUS, true, SIGN0, true>(tile[stage], accj[wid], lane, rowf, rowp, pc, sc, gi);
            else if (planar) sym_tile<RADIUS, true, SIGN0, false>(tile[stage], accj[wid], lane, rowf, rowp, pc, sc, gi);
            else if (tile_local) sym_tile<RADIUS, false, SIGN0, true>(tile[stage], accj[wid], lane, rowf, rowp, pc, sc, gi);
            else sym_tile<RADIUS, false, SIGN0, false>(tile[stage], accj[wid], lane, rowf, rowp, pc, sc, gi);
            n_local += tile_local ? 1 : 0;
            __syncthreads();                                    // every warp's J-side slice is complete
            // flush the J side: sum the warps' slices in fixed order, fixed-point atomics into the global accumulator
            for (int e = tid; e < K1_TJ; e += KS_THREADS) {
                float v[3];
                bool ok = true;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    v[c] = accj[0][c][e];
                    accj[0][c][e] = 0.0f;
#pragma unroll
                    for (int wv = 1; wv < KS_WARPS; ++wv) {           // fixed order: deterministic
                        v[c] += accj[wv][c][e];
                        accj[wv][c][e] = 0.0f;
                    }
                    ok = ok && fixed_ok(v[c]);
                }
                unsigned long long* dst = reinterpret_cast<unsigned long long*>(a.facc + ((size_t)J * K1_TJ + e) * 4);
                if (ok) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        if (v[c] != 0.0f) atomicAdd(dst + c, (unsigned long long)to_fixed(v[c]));
                } else {
                    atomicAdd(dst + 3, 1ull);
                }
            }
            __syncthreads();                                    // slices are zero again; the tile stage is free
        }
        // I side: this tile's partial into the thread's fixed-point registers (F_i -= g)
#pragma unroll
        for (int r = 0; r < KS_IR; ++r) {
            if (fixed_ok(gi[r][0]) && fixed_ok(gi[r][1]) && fixed_ok(gi[r][2])) {
                fix[r][0] -= to_fixed(gi[r][0]);
                fix[r][1] -= to_fixed(gi[r][1]);
                fix[r][2] -= to_fixed(gi[r][2]);
            } else {
                bad[r] += 1;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < KS_IR; ++r) {
        unsigned long long* dst =
            reinterpret_cast<unsigned long long*>(a.facc + ((size_t)I * K1_TJ + r * KS_THREADS + tid) * 4);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (fix[r][c] != 0) atomicAdd(dst + c, (unsigned long long)fix[r][c]);
        if (bad[r]) atomicAdd(dst + 3, (unsigned long long)bad[r]);
    }
    if (tid == 0 && n_local) atomicAdd(a.local_pairs, (unsigned long long)n_local);
}

// Fixed-point accumulators -> float64 pair force of the local rows.  Rows whose poison counter is set (a degenerate or
// overflowing pair met the unguarded fast path) are appended to a list; k1_sym_repair recomputes exactly those rows with
// the guarded scalar code (numpy's zero-safe semantics), one CTA per row striding over every staged slot.
struct FinishArgs {
    const float* planes;
    int rows_pad, world, own_block, n_local;
    const long long* facc_own;      // [rows_pad][4] this rank's block (after the reduce-scatter) ...
    const long long* facc_peer[7];  // ... or, with peer memory (K7), the same block inside every other rank's accumulator:
    int n_peer;                     //     the reduce-scatter is the sum below
    double* f_ped;                  // [n_local][3]
    unsigned long long* fixup_rows;
    int* bad_list;                  // [n_local] rows to repair
    int* bad_count;                 // [1], zeroed before the launch
    PairParams pp;
    const int* slot_of_row;         // row -> staged slot (k8_order.cuh); nullptr: row r is staged at slot r
};

__device__ __forceinline__ int slot_of(const FinishArgs& a, int row) { return a.slot_of_row ? a.slot_of_row[row] : row; }

__global__ void __launch_bounds__(256) k1_sym_finish(const FinishArgs a) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.n_local) return;
    const size_t slot = (size_t)slot_of(a, row);
    const longlong4 own = *reinterpret_cast<const longlong4*>(a.facc_own + slot * 4);
    long long fx = own.x, fy = own.y, fz = own.z, poison = own.w;
    for (int r = 0; r < a.n_peer; ++r) {            // integer sums: associative, so any order gives the same bits
        const longlong4 o = *reinterpret_cast<const longlong4*>(a.facc_peer[r] + slot * 4);
        fx += o.x; fy += o.y; fz += o.z; poison += o.w;
    }
    const double inv = 1.0 / 4294967296.0;
    a.f_ped[3 * (size_t)row + 0] = (double)fx * inv;
    a.f_ped[3 * (size_t)row + 1] = (double)fy * inv;
    a.f_ped[3 * (size_t)row + 2] = (double)fz * inv;
    if (poison != 0) a.bad_list[atomicAdd(a.bad_count, 1)] = row;       // list order is irrelevant: rows are independent
}

constexpr int KS_REPAIR_THREADS = 256;

template <bool RADIUS>
__global__ void __launch_bounds__(KS_REPAIR_THREADS) k1_sym_repair(const FinishArgs a) {
    __shared__ double red[3][KS_REPAIR_THREADS];
    const int count = *a.bad_count;
    const int tid = threadIdx.x;
    const int total = a.world * a.rows_pad;
    for (int b = blockIdx.x; b < count; b += gridDim.x) {
        const int r = a.bad_list[b];
        const int slot = slot_of(a, r);
        const RowF I = load_row(a.planes + ((size_t)a.own_block * NPLANES) * a.rows_pad, (size_t)a.rows_pad, (size_t)slot);
        const int islot = a.own_block * a.rows_pad + slot;
        double gx = 0.0, gy = 0.0, gz = 0.0;
        for (int j = tid; j < total; j += KS_REPAIR_THREADS) {
            const int q = j / a.rows_pad;
            const RowF J = load_row(a.planes + ((size_t)q * NPLANES) * a.rows_pad, (size_t)a.rows_pad,
                                    (size_t)(j - q * a.rows_pad));
            PairAcc acc = {0.0f, 0.0f, 0.0f};
            pair_force<RADIUS, true>(I, J, j == islot, a.pp, acc);
            gx += (double)acc.gx;
            gy += (double)acc.gy;
            gz += (double)acc.gz;
        }
        red[0][tid] = gx; red[1][tid] = gy; red[2][tid] = gz;
        __syncthreads();
        for (int o = KS_REPAIR_THREADS / 2; o > 0; o >>= 1) {        // fixed tree: deterministic
            if (tid < o) {
                red[0][tid] += red[0][tid + o];
                red[1][tid] += red[1][tid + o];
                red[2][tid] += red[2][tid + o];
            }
            __syncthreads();
        }
        if (tid == 0) {
            a.f_ped[3 * (size_t)r + 0] = -red[0][0];
            a.f_ped[3 * (size_t)r + 1] = -red[1][0];
            a.f_ped[3 * (size_t)r + 2] = -red[2][0];
            atomicAdd(a.fixup_rows, 1ull);
        }
        __syncthreads();
    }
}

}  // namespace sfm
