"""numpy check of the register-resident N = 512 real FFT schedule of csrc/fft16.cuh: 16 lanes x 16 registers per transform
(256-point packed-complex FFT as 16 x 16, one shared-memory transpose), the lane-pair real post-processing and its inverse."""
import numpy as np

rng = np.random.default_rng(1)
N, NC, L = 512, 256, 16
x = rng.standard_normal(N)
z = x[0::2] + 1j * x[1::2]
W = lambda n, e: np.exp(-2j * np.pi * e / n)

# forward: lane b holds v[a] = z[16 a + b]
v = np.array([[z[16 * a + b] for a in range(16)] for b in range(L)])
Y = np.array([[sum(v[b][a] * W(16, a * c) for a in range(16)) for c in range(16)] for b in range(L)])
Y *= np.array([[W(NC, b * c) for c in range(16)] for b in range(L)])
u = Y.T.copy()                                     # lane c holds u[b] = Y'_b[c]
Z = np.array([[sum(u[c][b] * W(16, b * d) for b in range(16)) for d in range(16)] for c in range(L)])   # Z[c][d] = FFT(z)[c + 16 d]
ref = np.fft.fft(z)
assert np.allclose([[ref[c + 16 * d] for d in range(16)] for c in range(16)], Z)

# real post-processing: lane c gets r[i] = Z[(16 - c) & 15][i]; partner of (c, d) is r[15 - d] (c >= 1) or own Z[0][(16 - d) & 15]
X = np.zeros((16, 16), complex)
for c in range(16):
    r = Z[(16 - c) & 15]
    for d in range(16):
        pz = Z[0][(16 - d) & 15] if c == 0 else r[15 - d]
        zk = Z[c][d]
        e = 0.5 * (zk + np.conj(pz)); o = -0.5j * (zk - np.conj(pz))
        X[c][d] = e + W(N, c) * W(32, d) * o
Xref = np.fft.rfft(x)
assert np.allclose([[Xref[c + 16 * d] for d in range(16)] for c in range(16)], X)
nyq = Z[0][0].real - Z[0][0].imag
assert np.allclose(nyq, Xref[256].real)

# inverse: lane c holds X[c + 16 d]; lane 0 also X[256]
Xh = np.array([[Xref[c + 16 * d] for d in range(16)] for c in range(16)])
Zi = np.zeros((16, 16), complex)
for c in range(16):
    r = Xh[(16 - c) & 15]
    for d in range(16):
        if c == 0:
            px = Xref[256] if d == 0 else Xh[0][16 - d]
        else:
            px = r[15 - d]
        xk = Xh[c][d]
        if c == 0 and d == 0:
            xk = xk.real + 0j; px = px.real + 0j
        e = 0.5 * (xk + np.conj(px)); dd = 0.5 * (xk - np.conj(px))
        Zi[c][d] = e + 1j * np.conj(W(N, c) * W(32, d)) * dd
assert np.allclose(Zi, Z)
U = np.array([[sum(Zi[c][d] * np.conj(W(16, b * d)) for d in range(16)) for b in range(16)] for c in range(16)])
U *= np.array([[np.conj(W(NC, b * c)) for b in range(16)] for c in range(16)])
w = U.T.copy()                                     # lane b holds w[c]
zz = np.array([[sum(w[b][c] * np.conj(W(16, a * c)) for c in range(16)) for a in range(16)] for b in range(16)]) / NC
assert np.allclose([[z[16 * a + b] for a in range(16)] for b in range(16)], zz)

# swizzled transpose buffer: element (row, col) at row * 16 + (col ^ row), 8-byte elements, 16 lanes per 64-bit wavefront
addr = lambda row, col: row * 16 + (col ^ row)
for fixed in range(16):
    assert len({addr(fixed, b) % 16 for b in range(16)}) == 16      # writer: fixed row, lanes = col
    assert len({addr(c, fixed) % 16 for c in range(16)}) == 16      # reader: fixed col, lanes = row
print("ok")
