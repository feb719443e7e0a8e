#!/usr/bin/env python
"""Prototype (numpy, float64) of the decimated inverse transform proposed in DESIGN.md §9 for the GCC-PHAT lag kernel.

The kernel needs r[l] = sum_{k=0}^{N-1} G[k] exp(+2 pi i k l / N) only for |l| <= L (L = 28, N = 1024), G Hermitian.
With k = B q + s (B = 32 sub-sequences of A = N / B = 32 bins):

    r[l] = sum_{s=0}^{B-1} exp(2 pi i s l / N) h_s[l mod A],      h_s[m] = sum_{q=0}^{A-1} G[B q + s] exp(2 pi i q m / A)

and Hermitian symmetry pairs the sub-sequences: h_{B-s}[m] = exp(-2 pi i m / A) conj(h_s[m]), so the terms s and B - s are complex
conjugates of each other and only s = 0 .. B/2 are transformed (17 transforms of 32 points instead of one of 1024):

    r[l] = Re h_0[m] + Re(exp(2 pi i (B/2) l / N) h_{B/2}[m]) + 2 sum_{s=1}^{B/2-1} Re(exp(2 pi i s l / N) h_s[m]),   m = l mod A.

This script checks the identity against numpy's inverse FFT on random Hermitian spectra and counts the arithmetic.
usage: python tools/proto/decimated_gcc.py
"""
import numpy as np


def lags_full(G_half, N, L):
    """reference: full one-sided -> real inverse transform, window |l| <= L (unnormalised sum)"""
    r = np.fft.irfft(G_half, n=N) * N
    return np.concatenate([r[N - L:], r[:L + 1]])


def lags_decimated(G_half, N, L, B=32):
    A = N // B
    k = np.arange(N)
    G = np.where(k <= N // 2, G_half[np.minimum(k, N // 2)], np.conj(G_half[np.minimum(N - k, N // 2)]))   # full Hermitian spectrum
    ls = np.arange(-L, L + 1)
    m = ls % A
    out = np.zeros(len(ls))
    for s in range(B // 2 + 1):
        h = np.fft.ifft(G[s::B]) * A                       # h_s[m], one A-point transform
        term = np.real(np.exp(2j * np.pi * s * ls / N) * h[m])
        out += term if s in (0, B // 2) else 2.0 * term
    return out


def main():
    rng = np.random.default_rng(0)
    N, L = 1024, 28
    worst = 0.0
    for _ in range(200):
        ph = rng.uniform(-np.pi, np.pi, N // 2 + 1)
        G = np.exp(1j * ph)                                 # PHAT: unit-modulus cross-spectrum
        G[0] = np.sign(np.cos(ph[0])) or 1.0
        G[-1] = np.sign(np.cos(ph[-1])) or 1.0              # DC / Nyquist real
        a, b = lags_full(G, N, L), lags_decimated(G, N, L)
        worst = max(worst, np.max(np.abs(a - b)) / np.max(np.abs(a)))
        assert np.argmax(a) == np.argmax(b)
    print(f"decimated == full inverse on the lag window: worst relative difference {worst:.2e} over 200 random spectra")
    A = B = 32
    full = 2.5 * (N // 2) * np.log2(N // 2) + 24 * (N // 4)          # packed N/2-point transform + even/odd build (FMA slots)
    dec = (B // 2 + 1) * 2.5 * A * np.log2(A) + (2 * L + 1) * (B // 2 + 1) * 2 + 4 * (B // 2 + 1) * A
    print(f"FMA slots per pair: packed full transform ~{full:.0f} (pruned last pass: ~{full - 1900:.0f}), decimated ~{dec:.0f}")


if __name__ == "__main__":
    main()
