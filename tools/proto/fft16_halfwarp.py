"""numpy check of the register-resident real-FFT schedule of csrc/fft16.cuh for N = 512 (R = 16 points per lane) and N = 1024 (R = 32):
16 lanes per transform (16 R-point packed-complex FFT as R x 16, one shared-memory transpose), the lane-pair real post-processing, its
inverse, and the bank-conflict freedom of the transpose buffer."""
import numpy as np

rng = np.random.default_rng(1)
W = lambda n, e: np.exp(-2j * np.pi * e / n)


def check(R):
    NC, N = 16 * R, 32 * R
    x = rng.standard_normal(N)
    z = x[0::2] + 1j * x[1::2]
    # forward: lane b holds v[a] = z[16 a + b], a < R
    v = np.array([[z[16 * a + b] for a in range(R)] for b in range(16)])
    Y = np.array([[sum(v[b][a] * W(R, a * c) for a in range(R)) for c in range(R)] for b in range(16)])          # dft_R over a
    Y *= np.array([[W(NC, b * c) for c in range(R)] for b in range(16)])
    Z = np.zeros((16, R), complex)                                                                              # Z[lane][e] = FFT(z)[lane + 16 e]
    for lane in range(16):
        for h in range(R // 16):                                                                                # rows lane + 16 h of the transpose
            u = np.array([Y[b][lane + 16 * h] for b in range(16)])
            out = np.array([sum(u[b] * W(16, b * d) for b in range(16)) for d in range(16)])                    # k = lane + 16 h + R d
            for d in range(16):
                Z[lane][(R // 16) * d + h] = out[d]
    ref = np.fft.fft(z)
    assert np.allclose([[ref[c + 16 * e] for e in range(R)] for c in range(16)], Z)

    # real post-processing: lane c gets r[i] = Z[(16 - c) & 15][i]; the partner of (c, e) is r[R - 1 - e] (c >= 1) or own Z[0][(R - e) % R]
    X = np.zeros((16, R), complex)
    for c in range(16):
        r = Z[(16 - c) & 15]
        for e in range(R):
            pz = Z[0][(R - e) % R] if c == 0 else r[R - 1 - e]
            zk = Z[c][e]
            X[c][e] = 0.5 * (zk + np.conj(pz)) + W(N, c) * W(2 * R, e) * (-0.5j) * (zk - np.conj(pz))
    Xref = np.fft.rfft(x)
    assert np.allclose([[Xref[c + 16 * e] for e in range(R)] for c in range(16)], X)
    assert np.allclose(Z[0][0].real - Z[0][0].imag, Xref[NC].real)                                              # the Nyquist bin

    # inverse pre-processing (lane 0 also holds X[N/2]) and the inverse schedule
    Xh = np.array([[Xref[c + 16 * e] for e in range(R)] for c in range(16)])
    Zi = np.zeros((16, R), complex)
    for c in range(16):
        r = Xh[(16 - c) & 15]
        for e in range(R):
            px = (Xref[NC] if e == 0 else Xh[0][R - e]) if c == 0 else r[R - 1 - e]
            xk = Xh[c][e]
            if c == 0 and e == 0:
                xk = xk.real + 0j; px = px.real + 0j
            Zi[c][e] = 0.5 * (xk + np.conj(px)) + 1j * np.conj(W(N, c) * W(2 * R, e)) * 0.5 * (xk - np.conj(px))
    assert np.allclose(Zi, Z)

    # swizzled transpose buffer: element (row c, col b) at 16 c + (b ^ (c & 15)), 8-byte elements, 16 lanes per 64-bit wavefront
    addr = lambda row, col: row * 16 + (col ^ (row & 15))
    for row in range(R):
        assert len({addr(row, b) % 16 for b in range(16)}) == 16                      # writer: fixed row, lanes = col
    for col in range(16):
        for h in range(R // 16):
            assert len({addr(lane + 16 * h, col) % 16 for lane in range(16)}) == 16   # reader: fixed col, lanes = row - 16 h


check(16)
check(32)
print("ok")
