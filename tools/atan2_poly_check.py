"""Accuracy of fpv_math.cuh:atan2_poly, emulated in numpy float32 (FMA through float64): max / mean ulp error against float64
arctan2 over random arguments of every octant, next to numpy's own float32 arctan2.  usage: python tools/atan2_poly_check.py"""
import numpy as np
f = np.float32
C = [0.00282363896258175373077393, -0.0159569028764963150024414, 0.0425049886107444763183594, -0.0748900920152664184570312,
     0.106347933411598205566406, -0.142027363181114196777344, 0.199926957488059997558594, -0.333331018686294555664062]


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f)


def atan2_poly(y, x):
    ax, ay = np.abs(x), np.abs(y)
    mx = np.maximum(np.maximum(ax, ay), f(1e-30)); mn = np.minimum(ax, ay)
    rc = (1.0 / mx.astype(np.float64)).astype(f)
    t = (mn * rc).astype(f); t = fma(fma(-mx, t, mn), rc, t)
    s = (t * t).astype(f)
    p = np.full_like(s, f(C[0]))
    for k in C[1:]:
        p = fma(p, s, np.full_like(s, f(k)))
    r = fma((p * s).astype(f), t, t)
    r = np.where(ay > ax, (f(np.pi / 2) - r).astype(f), r)
    r = np.where(x < 0, (f(np.pi) - r).astype(f), r)
    return np.copysign(r, y).astype(f)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    ang = rng.uniform(-np.pi, np.pi, 4_000_000); rad = np.exp(rng.uniform(-3, 1, ang.size))
    y = (rad * np.sin(ang)).astype(f); x = (rad * np.cos(ang)).astype(f)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    ulp = np.spacing(np.abs(ref).astype(f)).astype(np.float64)
    for name, got in (("atan2_poly", atan2_poly(y, x)), ("numpy float32 arctan2", np.arctan2(y, x))):
        e = np.abs(got.astype(np.float64) - ref) / ulp
        print(f"{name}: max {e.max():.2f} ulp, mean {e.mean():.3f} ulp, max abs {np.abs(got - ref).max():.2e}")
    print("atan2_poly(0, 0) =", atan2_poly(np.zeros(1, f), np.zeros(1, f))[0])
