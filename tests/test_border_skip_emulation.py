"""Soundness of the border kernel's negligible-term rule (csrc/k2_cells.cuh, `skip_on`), restated in numpy on the CPU.

The kernel drops a (pedestrian, section) term a exp(-(dist - r) / b) (forces.py:158-165) when the pedestrian is provably
farther than R = b ln(a / 1e-17) from every point of the section: dist >= Ds - E, with Ds the distance to the nearest
point of the section's uniform chord model and E the largest deviation of a real point from its model position (the
chord record built at upload, sfm_api.cu).  Here: the same float32 arithmetic on cfg2's sections, checked against the
float64 nearest distances -- every dropped term really is below 1e-17, and four candidate terms in five are dropped.
"""
import numpy as np

from sfm_b200 import synth

f32 = np.float32


def chord_record(points, centre):
    """(a, u, 1/|u|^2, E, P - 1) as sfm_api.cu builds it (centre-relative float32 a, u; E in float64 + rounding slack)."""
    P = len(points)
    a = (points[0] - centre).astype(f32)
    u = (points[-1] - points[0]).astype(f32)
    rel = points - centre
    frac = (np.arange(P) / (P - 1))[:, None] if P > 1 else np.zeros((1, 1))
    model = a.astype(np.float64)[None, :] + frac * u.astype(np.float64)[None, :]
    E = np.hypot(*(rel - model).T).max()
    M = max(np.abs(rel).max(), 20.0)
    Em = f32(np.nextafter(f32(E * 1.0001 + 8.0 * 2.0 ** -24 * M), f32(np.inf)))
    uu = float(u[0]) ** 2 + float(u[1]) ** 2
    return a, u, f32(1.0 / uu), Em, f32(P - 1)


def test_dropped_border_terms_are_below_1e_17():
    w = synth.make_config(2)
    a_par, b_par = 3.0, 0.1                                     # forces.py:135-136 defaults (sfm_config.toml has the same)
    R = f32(b_par * np.log(a_par / 1e-17))
    rng = np.random.default_rng(5)
    peds = rng.choice(w.n, 256, replace=False)
    dropped = kept = 0
    for s, pts in enumerate(w.borders):
        centre, cut = w.section_center[s], w.section_length[s]
        rel_p = w.loc[peds, :2] - centre
        inside = np.hypot(rel_p[:, 0], rel_p[:, 1]) < cut         # forces.py:149-150
        if not inside.any():
            continue
        a, u, inv, E, nm1 = chord_record(pts, centre)
        p32 = rel_p[inside].astype(f32)
        wv = p32 - a[None, :]
        kap = (wv[:, 0] * u[0] + wv[:, 1] * u[1]) * inv * nm1
        ks = np.clip(np.rint(kap), f32(0), nm1)
        fr = ks * (f32(1.0) / nm1)
        g = wv - fr[:, None] * u[None, :]
        Ds = np.sqrt(g[:, 0] * g[:, 0] + g[:, 1] * g[:, 1]).astype(f32)
        radf = (w.radius[peds][inside]).astype(f32) * f32(1.000001) + f32(1e-6)
        for radius_on in (False, True):
            r32 = radf if radius_on else f32(0)
            far = Ds * f32(0.9999) - E - f32(1e-3) - r32 > R
            true_min = np.hypot(w.loc[peds, None, 0][inside] - pts[None, :, 0], w.loc[peds, None, 1][inside] - pts[None, :, 1]).min(1)
            dist = true_min - (w.radius[peds][inside] if radius_on else 0.0)
            assert (a_par * np.exp(-dist[far] / b_par) < 1e-17).all()
        dropped += int(far.sum())
        kept += int((~far).sum())
    assert dropped > 3 * kept > 0, (dropped, kept)               # four candidate terms in five are numerically nothing
