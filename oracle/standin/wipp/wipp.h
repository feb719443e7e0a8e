// TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
//
// Stand-in for the WIPP vector library (jordi-adell/wipp, un-vendored, un-pinned:
// /root/reference/CMakeLists.txt:31, /root/reference/cmake/FindWIPP.cmake:9-10).
// WIPP's source is absent, so every primitive below restates the IPP-style
// semantics that the reference's own call sites imply.
// PARITY UNPINNED at this boundary: each assumption is listed in
// oracle/CONVENTIONS.md.
#ifndef ORACLE_STANDIN_WIPP_H
#define ORACLE_STANDIN_WIPP_H

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

namespace wipp {

typedef struct { double re; double im; } wipp_complex_t;

// ---- init / copy -------------------------------------------------------------------------
template <class T> inline void setZeros(T *x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] = T(); }
template <class T> inline void set(T v, T *x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] = v; }
inline void set(double v, double *x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] = v; }
template <class T> inline void copyBuffer(const T *src, T *dst, size_t n) { if (n) std::memmove(dst, src, n * sizeof(T)); }

// ---- arithmetic; IPP "in-place" convention: the LAST buffer is src-and-dst ------------------
// add(a, b, n): b += a          (SteeringBeamforming.cpp:140)
inline void add(const double *a, double *b, size_t n) { for (size_t i = 0; i < n; ++i) b[i] += a[i]; }
inline void add(const double *a, const double *b, double *c, size_t n) { for (size_t i = 0; i < n; ++i) c[i] = a[i] + b[i]; }
inline void add(const wipp_complex_t *a, wipp_complex_t *b, size_t n) { for (size_t i = 0; i < n; ++i) { b[i].re += a[i].re; b[i].im += a[i].im; } }
inline void add(const int16_t *a, int16_t *b, size_t n) { for (size_t i = 0; i < n; ++i) b[i] = int16_t(b[i] + a[i]); }
// sub(a, b, c, n): c = b - a    (comment "y[n] = x[n] - x[n-1]" at SteeringBeamforming.cpp:158-159)
inline void sub(const double *a, const double *b, double *c, size_t n) { for (size_t i = 0; i < n; ++i) c[i] = b[i] - a[i]; }
// mult(a, b, n): b *= a         (SteeringBeamforming.cpp:173)
inline void mult(const double *a, double *b, size_t n) { for (size_t i = 0; i < n; ++i) b[i] *= a[i]; }
inline void mult(const double *a, const double *b, double *c, size_t n) { for (size_t i = 0; i < n; ++i) c[i] = a[i] * b[i]; }
inline void mult(const wipp_complex_t *a, wipp_complex_t *b, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    double re = a[i].re * b[i].re - a[i].im * b[i].im, im = a[i].re * b[i].im + a[i].im * b[i].re;
    b[i].re = re; b[i].im = im;
  }
}
inline void mult(const wipp_complex_t *a, const wipp_complex_t *b, wipp_complex_t *c, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    double re = a[i].re * b[i].re - a[i].im * b[i].im, im = a[i].re * b[i].im + a[i].im * b[i].re;
    c[i].re = re; c[i].im = im;
  }
}
inline void multC(double C, double *x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] *= C; }
inline void multC(double C, const double *src, double *dst, size_t n) { for (size_t i = 0; i < n; ++i) dst[i] = src[i] * C; }
inline void divC(double C, double *x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] /= C; }
inline void divC(int C, int16_t *x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] = int16_t(x[i] / C); }
inline void addC(double C, double *x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] += C; }
inline void subC(double C, double *x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] -= C; }
inline void sqr(double *x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] *= x[i]; }
inline void sqr(const double *s, double *d, size_t n) { for (size_t i = 0; i < n; ++i) d[i] = s[i] * s[i]; }
inline void sqrt(double *x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] = std::sqrt(x[i]); }
inline void sqrt(const double *s, double *d, size_t n) { for (size_t i = 0; i < n; ++i) d[i] = std::sqrt(s[i]); }
inline void abs(const double *s, double *d, size_t n) { for (size_t i = 0; i < n; ++i) d[i] = std::fabs(s[i]); }
inline void abs(double *x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] = std::fabs(x[i]); }

// ---- complex helpers ---------------------------------------------------------------------
inline void magnitude(const wipp_complex_t *c, double *m, size_t n) { for (size_t i = 0; i < n; ++i) m[i] = std::sqrt(c[i].re * c[i].re + c[i].im * c[i].im); }
inline void real(const wipp_complex_t *c, double *re, size_t n) { for (size_t i = 0; i < n; ++i) re[i] = c[i].re; }
inline void imag(const wipp_complex_t *c, double *im, size_t n) { for (size_t i = 0; i < n; ++i) im[i] = c[i].im; }
inline void conj(const wipp_complex_t *s, wipp_complex_t *d, size_t n) { for (size_t i = 0; i < n; ++i) { d[i].re = s[i].re; d[i].im = -s[i].im; } }
inline void real2complex(const double *re, const double *im, wipp_complex_t *c, size_t n) {
  for (size_t i = 0; i < n; ++i) { c[i].re = re[i]; c[i].im = im ? im[i] : 0.0; }
}
inline void polar2cart(const double *mag, const double *phase, wipp_complex_t *c, size_t n) {
  for (size_t i = 0; i < n; ++i) { c[i].re = mag[i] * std::cos(phase[i]); c[i].im = mag[i] * std::sin(phase[i]); }
}

// ---- generators --------------------------------------------------------------------------
// ramp(x, n, offset, slope): x[i] = offset + slope*i   (Beamformer.cpp:58-59)
template <class T> inline void ramp(T *x, size_t n, double offset, double slope) { for (size_t i = 0; i < n; ++i) x[i] = T(offset + slope * double(i)); }
// tone: magn*cos(2*pi*f*i + phase), IPP ippsTone_Direct  (test_mcarray.cpp:916-917)
inline void tone(int16_t *x, size_t n, double magn, double freq, double phase) {
  for (size_t i = 0; i < n; ++i) x[i] = int16_t(std::lrint(magn * std::cos(2 * M_PI * freq * double(i) + phase)));
}
inline void tone(double *x, size_t n, double magn, double freq, double phase) {
  for (size_t i = 0; i < n; ++i) x[i] = magn * std::cos(2 * M_PI * freq * double(i) + phase);
}
// triangle: only used for an unused member (BinauralLocalisation.cpp:347); symmetric unit triangle.
inline void triangle(double *x, size_t n, double period, double phase) {
  for (size_t i = 0; i < n; ++i) {
    double t = std::fmod(double(i) / period + phase / (2 * M_PI), 1.0);
    x[i] = 1.0 - 4.0 * std::fabs(t - 0.5);
  }
}

inline void triangle(double *x, size_t n, double freq, double phase, double, double) { triangle(x, n, freq > 0 ? 1.0 / freq : double(n), phase); }

// ---- thresholds / filters ----------------------------------------------------------------
// strict comparisons as in ippsThreshold_LTValGTVal (SteeringBeamforming.cpp:160-161)
inline void threshold_lt_gt(double *x, size_t n, double tl, double vl, double tg, double vg) {
  for (size_t i = 0; i < n; ++i) { if (x[i] < tl) x[i] = vl; else if (x[i] > tg) x[i] = vg; }
}
// odd-length sliding median, centred, borders replicate the edge sample (legacy ippsFilterMedian).
inline void median_filter(const double *in, double *out, size_t n, size_t mask) {
  const long h = long(mask / 2);
  std::vector<double> w(mask);
  for (long i = 0; i < long(n); ++i) {
    for (long j = -h; j <= h; ++j) { long q = std::min(std::max(i + j, 0L), long(n) - 1); w[size_t(j + h)] = in[q]; }
    std::nth_element(w.begin(), w.begin() + h, w.end());
    out[i] = w[size_t(h)];
  }
}

// ---- statistics --------------------------------------------------------------------------
inline void sum(const double *x, size_t n, double *s) { double a = 0; for (size_t i = 0; i < n; ++i) a += x[i]; *s = a; }
inline void mean(const double *x, size_t n, double *m) { double a = 0; for (size_t i = 0; i < n; ++i) a += x[i]; *m = n ? a / double(n) : 0.0; }
inline void mean(const wipp_complex_t *x, size_t n, wipp_complex_t *m) {
  double a = 0, b = 0; for (size_t i = 0; i < n; ++i) { a += x[i].re; b += x[i].im; }
  m->re = n ? a / double(n) : 0.0; m->im = n ? b / double(n) : 0.0;
}
inline void min(const double *x, size_t n, double *m) { double a = x[0]; for (size_t i = 1; i < n; ++i) a = std::min(a, x[i]); *m = a; }
inline void max(const double *x, size_t n, double *m) { double a = x[0]; for (size_t i = 1; i < n; ++i) a = std::max(a, x[i]); *m = a; }
// first maximum wins (ippsMaxIndx; SteeringBeamforming.cpp:187-188)
inline void maxidx(const double *x, size_t n, double *mx, size_t *idx) {
  double a = x[0]; size_t k = 0; for (size_t i = 1; i < n; ++i) if (x[i] > a) { a = x[i]; k = i; }
  *mx = a; *idx = k;
}
inline void stddev(const double *x, size_t n, double *s) {
  double m; mean(x, n, &m); double a = 0; for (size_t i = 0; i < n; ++i) a += (x[i] - m) * (x[i] - m);
  *s = n > 1 ? std::sqrt(a / double(n - 1)) : 0.0;
}
// cross_corr: off the hot path (TemporalGCC / FastBinauralMasking::localise, both unused on the path)
inline void cross_corr(const double *a, size_t na, const double *b, size_t nb, double *out, size_t nout, int lowlag) {
  for (size_t i = 0; i < nout; ++i) {
    long lag = long(i) - lowlag; double s = 0;
    for (size_t n = 0; n < na; ++n) { long m = long(n) + lag; if (m >= 0 && m < long(nb)) s += a[n] * b[size_t(m)]; }
    out[i] = s;
  }
}

}  // namespace wipp

#endif
