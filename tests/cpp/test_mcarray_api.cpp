// C++ API tests in the shape of the reference's test/test_mcarray.cpp (gtest is not in this image: plain asserts).
//   test_mcarray_api array                    -> testArrayDescription (test_mcarray.cpp:518-580), host only
//   test_mcarray_api ssl <in.f64> <M> <n> <fs> <x0,x1,..> <chunk> <out_prefix>
//        runs mca::SourceSeparationAndLocalisation through process(std::vector<double*>&, ...) in `chunk`-sample calls the
//        way mcabeamf.cpp:101-119 does, and the frame-level mca::BeamformingSeparationAndLocalisation on the same spectra is
//        left to the Python parity tests; writes <out_prefix>.doa (text, one callback per line) and <out_prefix>.out (f64).
//   test_mcarray_api facade <in.s16> <n> <fs> <dist> <chunk> <out_prefix>
//        the ArrayModules facades (include/mcarray/ArrayModules.h:41-106): mca::SoundLocalisation (2 microphones ->
//        FreqGCCBinauralLocalisation, usePowerFloor = true, deterministic DOA tracker) and mca::BinauralMasking (FastBinauralMasking,
//        500-5000 Hz, RELATIVE / BOTH) fed int16 PCM in `chunk`-sample calls; writes <out_prefix>.doa and <out_prefix>.out (s16, planar).
//   test_mcarray_api frame <fs> <N>           -> SteeringBeamforming / Beamformer / BSAL frame-level classes on a synthetic
//        plane wave: the selected DOA must be the source cell and the beamformer steered there must return the source.
#include <mcarray/micarray.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#define EXPECT(cond)                                                                  \
  do {                                                                                \
    if (!(cond)) { std::cerr << __FILE__ << ":" << __LINE__ << ": EXPECT failed: " #cond << std::endl; ++g_failures; } \
  } while (0)
#define EXPECT_DOUBLE_EQ(a, b) EXPECT(std::fabs((a) - (b)) <= 4 * 2.220446049250313e-16 * std::max(std::fabs(a), std::fabs(b)))
static int g_failures = 0;

using namespace mca;

static void testArrayDescription() {
  ArrayDescription description;
  ArrayDescription::ElementId l, cl, cr, r;
  EXPECT(description.size() == 0);
  l = description.pushPosition(0, 2, 3, "left");
  cl = description.pushPosition(0.035 * 2, 2, 3, "center-left");
  cr = description.pushPosition(0.035 * 5, 2, 3, "center-right");
  r = description.pushPosition(0.035 * 6, 2, 3, "right");
  EXPECT(description.size() == 4);
  EXPECT(description.getName(l) == "left" && description.getName(cl) == "center-left" && description.getName(cr) == "center-right" && description.getName(r) == "right");
  EXPECT_DOUBLE_EQ(description.getX(l), 0.000); EXPECT_DOUBLE_EQ(description.getX(cl), 0.070);
  EXPECT_DOUBLE_EQ(description.getX(cr), 0.175); EXPECT_DOUBLE_EQ(description.getX(r), 0.210);
  for (int i = 0; i < 4; ++i) { EXPECT_DOUBLE_EQ(description.getY(i), 2.0); EXPECT_DOUBLE_EQ(description.getZ(i), 3.0); }
  const double want[4][4] = {{0, .070, .175, .210}, {.070, 0, .105, .140}, {.175, .105, 0, .035}, {.210, .140, .035, 0}};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) EXPECT_DOUBLE_EQ(description.distance(i, j), want[i][j]);
  EXPECT_DOUBLE_EQ(description.maxDistance(), 0.210);
  EXPECT_DOUBLE_EQ(description.distance("left", "center-left"), 0.070);
  EXPECT_DOUBLE_EQ(description.distance("left", "center-right"), 0.175);
  EXPECT_DOUBLE_EQ(description.distance("center-right", "center-left"), 0.105);
  EXPECT_DOUBLE_EQ(description.distance("right", "center-left"), 0.140);
  bool thrown = false;
  try { description.pushPosition(1, 1, 1, "left"); } catch (const MCArrayException &) { thrown = true; }
  EXPECT(thrown);                                                     // ArrayDescription.cpp:114-115
  ArrayDescription lin = ArrayDescription::make_linear_array_description(std::vector<double>{-2.25, -1.25, 1.25, 2.25});
  EXPECT(lin.size() == 4 && lin.getName(2) == "2");
  EXPECT_DOUBLE_EQ(lin.getBandwidth(), 346.1 / (2 * 4.5));
  EXPECT(ShortTimeProcessor::calculateOrderFromSampleRate(16000, 0.025f) == 9 && ShortTimeProcessor::calculateOrderFromSampleRate(48000, 0.025f) == 11);
}

class RecordingCallback : public LocalisationCallback {
 public:
  explicit RecordingCallback(std::ostream &os) : os_(os) {}
  virtual void setDOA(SignalPtr doa, SignalPtr prob, double power, int n) {
    char buf[128];
    for (int i = 0; i < n; ++i) { std::snprintf(buf, sizeof buf, "%.17g %.9g %.9g ", doa[i], prob[i], power); os_ << buf; }
    os_ << "\n";
    ++count;
  }
  int count = 0;
 private:
  std::ostream &os_;
};

static std::vector<double> parse_list(const char *s) {
  std::vector<double> v; std::stringstream ss(s); std::string tok;
  while (std::getline(ss, tok, ',')) v.push_back(std::atof(tok.c_str()));
  return v;
}

static int runSsl(int argc, char **argv) {
  if (argc < 9) return 64;
  const int M = std::atoi(argv[3]), n = std::atoi(argv[4]), fs = std::atoi(argv[5]), chunk = std::atoi(argv[7]);
  const std::vector<double> xs = parse_list(argv[6]);
  std::vector<double> x(size_t(M) * n);
  std::ifstream f(argv[2], std::ios::binary);
  f.read(reinterpret_cast<char *>(x.data()), std::streamsize(x.size() * 8));
  if (!f) { std::cerr << "short read" << std::endl; return 65; }
  ArrayDescription array = ArrayDescription::make_linear_array_description(xs);
  SourceSeparationAndLocalisation sss(fs, array, 1, false);
  std::ofstream doa(std::string(argv[8]) + ".doa");
  RecordingCallback cb(doa);
  sss.setCallback(cb);
  const int cap = chunk + sss.getMaxLatency();
  SignalVector out;
  std::vector<double *> rin(M), rout(M);
  for (int c = 0; c < M; ++c) { out.push_back(SignalPtr(new double[cap])); rout[c] = out[c].get(); }
  std::ofstream audio(std::string(argv[8]) + ".out", std::ios::binary);
  for (int pos = 0; pos < n; pos += chunk) {
    const int m = std::min(chunk, n - pos);
    for (int c = 0; c < M; ++c) rin[c] = x.data() + size_t(c) * n + pos;
    const int got = sss.process(rin, m, rout, cap);
    audio.write(reinterpret_cast<const char *>(out[0].get()), std::streamsize(got) * 8);
    for (int c = 1; c < M; ++c)
      for (int i = 0; i < got; ++i) EXPECT(out[c][i] == 0.0);          // channels >= numOfSources are zeroed (BSAL.cpp:117-118)
  }
  std::cout << "frames with callbacks: " << cb.count << std::endl;
  return g_failures ? 1 : 0;
}

// MultibandBinarualLocalisation (N2) on a planar stereo signal: one line per callback
static int runMultiband(int argc, char **argv) {
  if (argc < 8) return 64;
  const int n = std::atoi(argv[3]), fs = std::atoi(argv[4]), chunk = std::atoi(argv[6]);
  const double dist = std::atof(argv[5]);
  std::vector<double> x(size_t(2) * n);
  std::ifstream f(argv[2], std::ios::binary);
  f.read(reinterpret_cast<char *>(x.data()), std::streamsize(x.size() * 8));
  if (!f) { std::cerr << "short read" << std::endl; return 65; }
  ArrayDescription array = ArrayDescription::make_linear_array_description(std::vector<double>{0.0, dist});
  MultibandBinarualLocalisation mbl(fs, array, 15, false);
  EXPECT(mbl.getNumberOfBins() == 15 && mbl.getWindowSize() == 512);
  std::ofstream doa(std::string(argv[7]) + ".doa");
  RecordingCallback cb(doa);
  mbl.setCallback(cb);
  std::vector<double *> rin(2);
  for (int pos = 0; pos < n; pos += chunk) {
    for (int c = 0; c < 2; ++c) rin[c] = x.data() + size_t(c) * n + pos;
    mbl.process(rin, std::min(chunk, n - pos));
  }
  std::cout << "frames with callbacks: " << cb.count << std::endl;
  return g_failures ? 1 : 0;
}

// the two facade classes of ArrayModules.h on a planar int16 stereo signal
static int runFacade(int argc, char **argv) {
  if (argc < 8) return 64;
  const int n = std::atoi(argv[3]), fs = std::atoi(argv[4]), chunk = std::atoi(argv[6]);
  const double dist = std::atof(argv[5]);
  std::vector<int16_t> x(size_t(2) * n);
  std::ifstream f(argv[2], std::ios::binary);
  f.read(reinterpret_cast<char *>(x.data()), std::streamsize(x.size() * 2));
  if (!f) { std::cerr << "short read" << std::endl; return 65; }
  ArrayDescription array = ArrayDescription::make_linear_array_description(std::vector<double>{0.0, dist});
  std::ofstream doa(std::string(argv[7]) + ".doa");
  RecordingCallback cb(doa);
  SoundLocalisation loc(fs, array, &cb);
  BinauralMasking mask(fs, array, 500, 5000, BinauralMasking::RELATIVE, BinauralMasking::BOTH);
  EXPECT(loc.getFrameSize() > 0 && mask.getFrameSize() > 0);
  const int cap = chunk + mask.getMaxLatency();
  std::vector<int16_t> o0(cap), o1(cap);
  std::vector<int16_t *> rin(2), rout{o0.data(), o1.data()};
  std::vector<int16_t> y0, y1;
  for (int pos = 0; pos < n; pos += chunk) {
    const int m = std::min(chunk, n - pos);
    for (int c = 0; c < 2; ++c) rin[c] = x.data() + size_t(c) * n + pos;
    loc.process(rin, m);
    const int got = mask.process(rin, m, rout, cap);
    y0.insert(y0.end(), o0.begin(), o0.begin() + got); y1.insert(y1.end(), o1.begin(), o1.begin() + got);
  }
  std::ofstream audio(std::string(argv[7]) + ".out", std::ios::binary);
  audio.write(reinterpret_cast<const char *>(y0.data()), std::streamsize(y0.size() * 2));
  audio.write(reinterpret_cast<const char *>(y1.data()), std::streamsize(y1.size() * 2));
  std::cout << "frames with callbacks: " << cb.count << ", samples out: " << y0.size() << std::endl;
  return g_failures ? 1 : 0;
}

// plane wave from grid cell `cell` on a linear array: X_c[k] = S[k] exp(+j 2 pi k fs/N x_c sin(theta)/c)
static int runFrame(int argc, char **argv) {
  const int fs = argc > 2 ? std::atoi(argv[2]) : 16000, N = argc > 3 ? std::atoi(argv[3]) : 512, ccs = N + 2, K = N / 2 + 1;
  const std::vector<double> xs{0, 0.07, 0.175, 0.21};                  // Reem-C array, test_mcarray.cpp:397
  ArrayDescription array = ArrayDescription::make_linear_array_description(xs);
  const int M = int(xs.size()), cell = 24;                             // 24 * 5 - 90 = 30 degrees
  const double theta = double(float(float(cell) * float(5 * M_PI / 180) - M_PI / 2));
  SignalVector frames, outs, wiener;
  std::vector<double> S(size_t(2) * K);
  unsigned lcg = 12345;
  for (int k = 0; k < K; ++k) { lcg = lcg * 1664525u + 1013904223u; const double ph = 2 * M_PI * (lcg >> 8) / 16777216.0; S[2 * k] = 1000 * std::cos(ph); S[2 * k + 1] = (k == 0 || k == K - 1) ? 0 : 1000 * std::sin(ph); }
  for (int c = 0; c < M; ++c) {
    frames.push_back(SignalPtr(new double[ccs])); outs.push_back(SignalPtr(new double[ccs]));
    for (int k = 0; k < K; ++k) {
      const double a = 2 * M_PI * k * double(fs) / N * xs[c] * std::sin(theta) / 346.1, cr = std::cos(a), ci = std::sin(a);
      frames[c][2 * k] = S[2 * k] * cr - S[2 * k + 1] * ci;
      frames[c][2 * k + 1] = S[2 * k] * ci + S[2 * k + 1] * cr;
    }
  }
  BeamformingSeparationAndLocalisation bsal(fs, ccs, array, 1, false);
  std::ostringstream rec;
  RecordingCallback cb(rec);
  bsal.setCallback(cb);
  for (int t = 0; t < 4; ++t) bsal.processFrameLocalisation(frames, wiener);
  EXPECT(cb.count == 4);
  std::istringstream last(rec.str().substr(rec.str().rfind('\n', rec.str().size() - 2) + 1));
  double deg = 0; last >> deg;
  std::cout << "frame-level DOA " << deg << " degrees (source at " << theta * 180 / M_PI << ")" << std::endl;
  EXPECT(std::fabs(deg - theta * 180 / M_PI) < 1e-9);
  bsal.processFrameSeparation(frames, outs);
  double err = 0, ref = 0;
  for (int i = 0; i < ccs; ++i) { err = std::max(err, std::fabs(outs[0][i] - S[i])); ref = std::max(ref, std::fabs(S[i])); }
  std::cout << "beamformer steered at the source: max error " << err << " of " << ref << std::endl;
  EXPECT(err <= 1e-6 + 2e-4 * ref);
  for (int c = 1; c < M; ++c) for (int i = 0; i < ccs; ++i) EXPECT(outs[c][i] == 0.0);
  // configuration errors throw MCArrayException (FastBinauralMasking.cpp:88-91 / BinauralLocalisation)
  bool thrown = false;
  try { FreqGCCBinauralLocalisation bad(16000, array); } catch (const MCArrayException &) { thrown = true; }
  EXPECT(thrown);
  return g_failures ? 1 : 0;
}

int main(int argc, char **argv) {
  const std::string mode = argc > 1 ? argv[1] : "array";
  try {
    if (mode == "array") { testArrayDescription(); std::cout << (g_failures ? "FAILED" : "testArrayDescription ok") << std::endl; return g_failures ? 1 : 0; }
    if (mode == "ssl") return runSsl(argc, argv);
    if (mode == "frame") return runFrame(argc, argv);
    if (mode == "multiband") return runMultiband(argc, argv);
    if (mode == "facade") return runFacade(argc, argv);
  } catch (const std::exception &e) {
    std::cerr << "exception: " << e.what() << std::endl;
    return 3;
  }
  return 64;
}
