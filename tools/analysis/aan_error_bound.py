#!/usr/bin/env python
"""Rigorous worst-case error bound of the FP32 AAN filter transform used by k_bgr_to_coef's fast path.

Model: every FP32 operation z = RN(op) has |error| <= u * |z_exact_of_inputs| with u = 2^-24; nodes that are sums
of integers below 2^24 are exact.  Each node is tracked as (linear form over the 64 inputs, absolute error bound).
Inputs are integers in [-128, 127].  The bound for output (u,v) is in units of the final quantised value
v~ = out * K[u,v]  with  K = true_scale/(4*q).
"""
import numpy as np
U = 2.0 ** -24
f32 = lambda x: float(np.float32(x))
C707, C382, C541, C1306 = f32(0.70710678118654752), f32(0.38268343236508977), f32(0.54119610014619698), f32(1.30656296487637652)

class Node:
    __slots__ = ("lin", "err", "exact_int")
    def __init__(s, lin, err=0.0, exact_int=True):
        s.lin, s.err, s.exact_int = lin, err, exact_int
    def mag(s):
        return 128.0 * np.abs(s.lin).sum() + s.err
def add(a, b, sign=1.0):
    lin = a.lin + sign * b.lin
    z = Node(lin, a.err + b.err, a.exact_int and b.exact_int)
    if not (z.exact_int and z.mag() < 2 ** 24):
        z.exact_int = False
        z.err += U * z.mag()
    return z
def fma(a, c, b):      # a*c + b, one rounding
    z = Node(a.lin * c + b.lin, abs(c) * a.err + b.err, False)
    z.err += U * z.mag()
    return z
def mul(a, c):
    z = Node(a.lin * c, abs(c) * a.err, False)
    z.err += U * z.mag()
    return z

def aan(d):
    t0, t7 = add(d[0], d[7]), add(d[0], d[7], -1)
    t1, t6 = add(d[1], d[6]), add(d[1], d[6], -1)
    t2, t5 = add(d[2], d[5]), add(d[2], d[5], -1)
    t3, t4 = add(d[3], d[4]), add(d[3], d[4], -1)
    t10, t13 = add(t0, t3), add(t0, t3, -1)
    t11, t12 = add(t1, t2), add(t1, t2, -1)
    o = [None] * 8
    o[0], o[4] = add(t10, t11), add(t10, t11, -1)
    s = add(t12, t13)
    o[2], o[6] = fma(s, C707, t13), fma(s, -C707, t13)
    a10, a11, a12 = add(t4, t5), add(t5, t6), add(t6, t7)
    z5 = mul(add(a10, a12, -1), C382)
    z2, z4 = fma(a10, C541, z5), fma(a12, C1306, z5)
    z11, z13 = fma(a11, C707, t7), fma(a11, -C707, t7)
    o[5], o[3] = add(z13, z2), add(z13, z2, -1)
    o[1], o[7] = add(z11, z4), add(z11, z4, -1)
    return o

QUANT = {
    "luma": [16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
             18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99],
    "chroma": [17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99] + [99] * 32}


def position_bounds(verbose=False):
    """{comp: 8x8 array B[v][u]}: worst-case |d*K - v_true| of every coefficient, in units of the quantised value."""
    # 2-D: rows first (index [y][x] -> pass over x), then columns
    X = [[Node(np.eye(64)[y * 8 + x]) for x in range(8)] for y in range(8)]
    R = [aan(X[y]) for y in range(8)]                       # R[y][u]
    Cc = [aan([R[y][u] for y in range(8)]) for u in range(8)]  # Cc[u][v]
    # true unnormalised 2-D DCT linear forms (reference convention): F[v][u] = sum_y sum_x p[y][x] cos((2y+1)v pi/16) cos((2x+1)u pi/16)
    n = np.arange(8)
    Ct = np.cos((2 * n[:, None] + 1) * n[None, :] * np.pi / 16)   # Ct[t][f]
    out = {}
    for comp, q in QUANT.items():
        B = np.zeros((8, 8))
        for v in range(8):
            for u in range(8):
                node = Cc[u][v]
                true = np.outer(Ct[:, v], Ct[:, u]).reshape(64)        # [y*8+x]
                r = float(true @ true) / float(node.lin @ true)        # AAN output scale: node.lin ~= true / r
                cu = (0.5 ** 0.5 if u == 0 else 1.0) * (0.5 ** 0.5 if v == 0 else 1.0)
                K = f32(r * cu / (4.0 * q[v * 8 + u]))                  # the FP32 multiplier the kernel brackets around
                sys_err = 128.0 * np.abs(node.lin * K - true * cu / (4.0 * q[v * 8 + u])).sum()   # constants' rounding + K rounding
                rnd_err = node.err * K
                B[v, u] = sys_err + rnd_err
                if verbose and (u, v) in ((0, 0), (1, 0), (0, 1), (1, 1), (7, 7), (4, 4)):
                    print(f"{comp} (u={u},v={v}) q={q[v*8+u]:3d} r={r:.6f} max|v~|={node.mag()*K:8.2f} sys={sys_err:.2e} rnd={rnd_err:.2e}")
        out[comp] = B
    return out


def main():
    B = position_bounds(verbose=True)
    np.set_printoptions(linewidth=200, precision=2)
    for comp, b in B.items():
        print(comp, "worst-case |v~ - v_true| =", b.max(), "; per position (x 1e-5, rows = vertical frequency):")
        print(b * 1e5)
    print("bound:", max(b.max() for b in B.values()))

if __name__ == "__main__":
    main()
