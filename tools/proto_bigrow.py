#!/usr/bin/env python
"""Index-algebra prototype of the whole-row-in-shared-memory waterfall kernel (fft_bigrow_kernel in
csrc/fft_bigrow.cuh): in-place DIF radix-16 x3 + radix-R3 on a padded row, two adjacent butterflies
per thread through 16-byte accesses. Checks (a) the transform against numpy, (b) that every 16-byte
shared-memory access pattern is conflict-free per quarter warp, (c) the output permutation.
Run on CPU; no GPU needed.  usage: proto_bigrow.py [13|14]"""
import sys
import numpy as np

LOGL = int(sys.argv[1]) if len(sys.argv) > 1 else 14
L = 1 << LOGL
R3 = L // 4096
B1, B2 = L // 16, L // 256
S2 = B2           # no padding inside a chunk: each chunk of B1 elements is ONE contiguous bulk copy
S1 = B1 + 2       # chunk stride in 16-byte units is odd: lanes that differ in d0 never share a bank group
NT = L // 32
SIGN = +1.0  # backward transform (waterfall)


def phys(p):
    d0, r = divmod(p, B1)
    d1, r = divmod(r, B2)
    return d0 * S1 + d1 * S2 + r


def check16(addrs, what):
    """addrs: per-thread element offsets (even) of one 16-byte access instruction"""
    a = np.asarray(addrs)
    assert np.all(a % 2 == 0), what
    for q in range(0, NT, 8):
        ch = (a[q:q + 8] // 2) % 8
        if len(set(ch.tolist())) != 8:
            return 8 // len(set(ch.tolist())) if len(set(ch.tolist())) in (1, 2, 4) else False
    return True


def w(n, e):
    return np.exp(SIGN * 2j * np.pi * (e % n) / n)


rng = np.random.default_rng(1)
x = rng.standard_normal(L) + 1j * rng.standard_normal(L)
buf = np.zeros(16 * S1, complex)
for p in range(L):
    buf[phys(p)] = x[p]

conf = {}
t = np.arange(NT)
# stage 0: pairs j = 2t, 2t+1 in [0, B1); elements at stride S1
j0 = 2 * t
off0 = (j0 // B2) * S2 + (j0 % B2)
conf["s0"] = all(check16(off0 + i * S1, "s0") for i in range(16))
for th in range(NT):
    for bfly in range(2):
        j = 2 * th + bfly
        a = np.array([buf[off0[th] + bfly + i * S1] for i in range(16)])
        y = np.array([sum(a[m] * w(16, i * m) for m in range(16)) for i in range(16)])
        y *= np.array([w(L, i * j) for i in range(16)])
        for i in range(16):
            buf[off0[th] + bfly + i * S1] = y[i]
# stage 1: t -> d0 = t // (B2/2), jp = t % (B2/2); stride S2; twiddle W_B1^(i j)
d0 = t // (B2 // 2)
jp = t % (B2 // 2)
off1 = d0 * S1 + 2 * jp
conf["s1"] = all(check16(off1 + i * S2, "s1") for i in range(16))
for th in range(NT):
    for bfly in range(2):
        j = 2 * jp[th] + bfly
        a = np.array([buf[off1[th] + bfly + i * S2] for i in range(16)])
        y = np.array([sum(a[m] * w(16, i * m) for m in range(16)) for i in range(16)])
        y *= np.array([w(B1, i * j) for i in range(16)])
        for i in range(16):
            buf[off1[th] + bfly + i * S2] = y[i]
# stage 2: t -> blk = t // (R3/2) = (d0, d1), jp = t % (R3/2); stride R3; twiddle W_B2^(i j)
sd0 = t & 15
rest = t >> 4
jp2 = rest % (R3 // 2)
sd1 = rest // (R3 // 2)
off2 = sd0 * S1 + sd1 * S2 + 2 * jp2
conf["s2"] = all(check16(off2 + i * R3, "s2") for i in range(16))
for th in range(NT):
    for bfly in range(2):
        j = 2 * jp2[th] + bfly
        a = np.array([buf[off2[th] + bfly + i * R3] for i in range(16)])
        y = np.array([sum(a[m] * w(16, i * m) for m in range(16)) for i in range(16)])
        y *= np.array([w(B2, i * j) for i in range(16)])
        for i in range(16):
            buf[off2[th] + bfly + i * R3] = y[i]
# last stage, pass 1: radix R3 on contiguous groups (in place); thread takes 32 / R3 groups
ng = 32 // R3
ok = True
for g in range(ng):
    # lanes vary d0 (chunk stride odd in 16-byte units): conflict free for both radices
    gd0 = t & 15
    rest = (t >> 4) + (NT // 16) * g
    gd1, gd2 = rest >> 4, rest & 15
    base = gd0 * S1 + gd1 * S2 + gd2 * R3
    for c in range(R3 // 2):
        ok &= check16(base + 2 * c, "last")
    for th in range(NT):
        a = np.array([buf[base[th] + m] for m in range(R3)])
        y = np.array([sum(a[m] * w(R3, i * m) for m in range(R3)) for i in range(R3)])
        for m in range(R3):
            buf[base[th] + m] = y[m]
conf["last_pass1"] = ok
# pass 2: lane -> (d0 = lane & 15, d1lo = lane >> 4); warp -> d1hi = w & 7, sel = w >> 3 (R3/2 values);
# iteration it -> d2; 16-byte read of (d3 = 2 sel, 2 sel + 1) -> k, k + L/R3
out = np.zeros(L, complex)
lane, wid = t & 31, t >> 5
pd0, pd1 = lane & 15, (wid & 7) * 2 + (lane >> 4)
sel = wid >> 3
ok = True
for it in range(16):
    addr = pd0 * S1 + pd1 * S2 + it * R3 + 2 * sel
    ok &= check16(addr, "pass2")
    k = pd0 + 16 * pd1 + 256 * it + 4096 * (2 * sel)
    for th in range(NT):
        out[k[th]] = buf[addr[th]]
        out[k[th] + 4096] = buf[addr[th] + 1]
    # coalescing: lanes of a warp write 32 consecutive k
    for wv in range(NT // 32):
        kk = k[wv * 32:(wv + 1) * 32]
        assert np.array_equal(kk, kk[0] + np.arange(32))
conf["pass2"] = ok
ref = np.fft.ifft(x) * L if SIGN > 0 else np.fft.fft(x)
err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
print(f"L=2^{LOGL} R3={R3} S2={S2} S1={S1} buf_elems={16 * S1} ({16 * S1 * 8} B) NT={NT}")
print("conflict-free:", conf)
print("rel-L2 vs numpy:", err)
assert err < 1e-12
