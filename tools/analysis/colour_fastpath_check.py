#!/usr/bin/env python
"""Exhaustive check of the FP32/integer colour fast path of k_bgr_to_coef (fast kernel).

For every reachable numerator n:
  floor: RZ_f32(n * INV + 2^23) - 2^23 == n // D        (FMA: exact product, one rounding toward zero)
  ties : (n * M mod 2^32) < 2^19  whenever n % D == 0    (superset test; false-positive rate reported)
Y : n = 299R+587G+114B, D = 1000 ; Cb/Cr : n = 4e6 + ... , D = 31250 (coefficients divided by 32).
"""
import numpy as np

def ru_f32(x):            # smallest float32 >= x
    f = np.float32(x)
    if float(f) < x:
        f = np.nextafter(f, np.float32(np.inf))
    return f

def check(D, nmax, name):
    inv = ru_f32(1.0 / D)
    M = -(-2 ** 32 // D)
    n = np.arange(0, nmax + 1, dtype=np.int64)
    prod = n.astype(np.float64) * float(inv)                 # exact: 23-bit n times 24-bit inv fits in 53 bits
    fl = np.floor(prod).astype(np.int64)                     # RZ at integer grid == floor for non-negative values (2^23 + prod < 2^24)
    assert (prod < 2 ** 23).all()
    bad = np.nonzero(fl != n // D)[0]
    lo = (n * M) & 0xFFFFFFFF
    ties = (n % D) == 0
    miss = np.nonzero(ties & ~(lo < 2 ** 19))[0]
    fp = np.count_nonzero(~ties & (lo < 2 ** 19)) / n.size
    print(f"{name}: D={D} inv={float(inv)!r} (bits {inv.view(np.uint32):#x}) M={M} n<= {nmax}: floor mismatches {bad.size}, missed ties {miss.size}, false-positive rate {fp:.2e}")
    assert bad.size == 0 and miss.size == 0

def check_remainder(D, nmax, name):
    """k_pixels_to_tokens' exact screen: r' = n_bits - q_bits * D (mod 2^32) on the float bit patterns 0x4B000000 + n and
    0x4B000000 + n // D.  It must equal K_D = 0x4B000000 * (1 - D) mod 2^32 exactly when D divides n, and stay in
    [K_D, K_D + D) (no wrap-around), so that the running unsigned minimum detects a tie."""
    n = np.arange(0, nmax + 1, dtype=np.uint64)
    nb, qb = (0x4B000000 + n) & 0xFFFFFFFF, (0x4B000000 + n // D) & 0xFFFFFFFF
    r = (nb - qb * D) & 0xFFFFFFFF
    K = (0x4B000000 * (1 - D)) % (1 << 32)
    assert K + D < (1 << 32), "wrap-around"
    assert ((r >= K) & (r < K + D)).all() and ((r == K) == (n % D == 0)).all()
    print(f"{name}: exact remainder screen K_D = {K:#010x}, range [K_D, K_D + {D}) without wrap-around: ok")

check(1000, 255 * 1000, "Y ")
check_remainder(1000, 255 * 1000, "Y ")
check_remainder(31250, 4_000_000 + 255 * 15625, "Cb/Cr")
check(31250, 4_000_000 + 255 * 15625, "Cb/Cr")
# reachable numerator ranges
R, G, B = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
cb = 4_000_000 - 5273 * R - 10352 * G + 15625 * B
cr = 4_000_000 + 15625 * R - 13084 * G - 2541 * B
print("Cb numerator range", cb.min(), cb.max(), " Cr", cr.min(), cr.max(), " (must be in [0, 2^23) =", 2 ** 23, ")")
# cross-check the /32 reduction against the reference's 1e6-denominator form
cb6 = 128_000_000 - 168736 * R - 331264 * G + 500000 * B
cr6 = 128_000_000 + 500000 * R - 418688 * G - 81312 * B
assert (cb6 == 32 * cb).all() and (cr6 == 32 * cr).all()
print("ok")


# ---- round 2: the FADD-free form (ycc_row8n in dct_core.cuh) ---------------------------------------------------------
# x = T + n is the float whose bits are 0x4B000000 + (T - 2^23) + n; q_bits = RZ_f32(x * inv + C) with C = 2^23 + off - T / D.
def mant_exp(bits):
    return (bits & 0x7FFFFF) | 0x800000, 150 - ((bits >> 23) & 0xFF)      # value = M * 2^-e exactly

def check_nofadd(D, T, off, inv_bits, lo, hi, name, tie_delta):
    M, e = mant_exp(inv_bits)
    n = np.arange(lo, hi + 1, dtype=np.int64)                              # numerator without any bias
    assert T % D == 0 and T + lo >= 2 ** 23 and T + hi < 2 ** 24
    C = 2 ** 23 + off - T // D
    q = (((T + n) * M + (C << e)) >> e) - 2 ** 23 - off                    # exact product + C, rounded toward zero (positive)
    want = n // D
    tie = (n % D) == 0
    assert (q[~tie] == want[~tie]).all(), name
    assert (q[tie] == want[tie] + tie_delta).all(), name
    # remainder screen on the bit patterns
    xb = (0x4B000000 + T - 2 ** 23 + n) & 0xFFFFFFFF
    qb = (0x4B000000 + off + q) & 0xFFFFFFFF
    r = (xb - qb * D) & 0xFFFFFFFF
    K = (0x4B000000 * (1 - D) + T - 2 ** 23 - off * D) % (1 << 32)
    assert K + D < (1 << 32)
    flag = K if tie_delta == 0 else K + D
    assert ((r == flag) == tie).all() and (r >= K).all() and (r <= K + D).all()
    print(f"{name}: FADD-free floor exact off ties, ties read {tie_delta:+d}; screen constant {flag:#010x} ({'min' if tie_delta == 0 else 'max'})")

check_nofadd(1000, 8_389_000, 0, 0x3a83126f, 0, 255_000, "Y  (no FADD)", 0)
check_nofadd(31250, 12_500_000, 128, 0x380637bd, -255 * 15625, 255 * 15625, "C  (no FADD)", -1)
