// mcbeam (B200 build) — command-line twin of the reference's src/programs/mcabeamf.cpp:
//   mcbeam -i in.wav -o out.wav [-d doafile | -d -] [-a positions] [-b frames]
// reads a multichannel WAV, runs mca::SourceSeparationAndLocalisation(sampleRate, array, numOfSources = 1,
// usePowerFloor = false) on it (mcabeamf.cpp:182-194) and writes channel 0 of the separated signal as a mono WAV
// (:114-119); with -d every localised frame prints "[DOA: x, p=y, P=z] " (:53-73).  The array defaults to the reference's
// hard-coded 4-microphone line {-2.25, -1.25, 1.25, 2.25} (:182); -a, parsed but unused there, takes a comma-separated
// list of x positions here.  -b sets how many frames are batched per device call (the reference feeds 1024 samples at a
// time, :80; the GPU wants far more per launch, and the result does not depend on it).
#include <mcarray/SourceSeparationAndLocalisation.h>

#include <unistd.h>

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>

#include "wav.h"

namespace {

void usage(const char *argv0) {
  std::cout << "Use: " << argv0 << " [options]" << std::endl;
  std::cout << "  options:" << std::endl;
  std::cout << "      -i file    Input file" << std::endl;
  std::cout << "      -o file    Output file" << std::endl;
  std::cout << "      -a list    Array description: comma separated x positions in metres" << std::endl;
  std::cout << "      -d file    Print DOA in file [use - for stdout]" << std::endl;
  std::cout << "      -b frames  Frames per device call (default 256)" << std::endl;
  std::cout << "      -h         This help message" << std::endl;
  std::cout << std::endl;
}

class McBeamCallback : public mca::LocalisationCallback {
 public:
  explicit McBeamCallback(std::ostream &os) : os_(os) {}
  virtual void setDOA(mca::SignalPtr doas, mca::SignalPtr probs, double power, int numOfSources) {
    for (int i = 0; i < numOfSources; ++i) os_ << "[DOA: " << doas[i] << ", p=" << probs[i] << ", P=" << power << "] ";
    os_ << std::endl;
  }
 private:
  std::ostream &os_;
};

}  // namespace

int main(int argc, char *argv[]) {
  std::string input_file, output_file, array_description, doa_file;
  int batch_frames = 256, c;
  while ((c = getopt(argc, argv, "i:o:a:d:b:gh")) != -1) {
    switch (c) {
      case 'i': input_file = optarg; break;
      case 'o': output_file = optarg; break;
      case 'a': array_description = optarg; break;
      case 'd': doa_file = optarg; break;
      case 'b': batch_frames = std::atoi(optarg); break;
      case 'g': break;   // the reference's optional DSPONE GUI: not part of this build
      default: usage(argv[0]); return 1;
    }
  }
  if (input_file.empty() || output_file.empty() || batch_frames < 2) { usage(argv[0]); return 1; }

  std::ostream *os = NULL;
  std::unique_ptr<std::ofstream> fos;
  if (!doa_file.empty()) {
    if (doa_file == "-") os = &std::cout;
    else { fos.reset(new std::ofstream(doa_file.c_str())); os = fos.get(); }
  }
  std::vector<double> mics_positions;
  if (array_description.empty()) {
    const double def[] = {-2.25, -1.25, 1.25, 2.25};
    mics_positions.assign(def, def + 4);
  } else {
    std::stringstream ss(array_description);
    std::string tok;
    while (std::getline(ss, tok, ',')) mics_positions.push_back(std::atof(tok.c_str()));
  }
  try {
    mca::ArrayDescription array = mca::ArrayDescription::make_linear_array_description(mics_positions);
    wav::Reader in(input_file);
    const int nchannels = in.info().channels;
    if (nchannels != int(array.size())) throw std::runtime_error("the input has " + std::to_string(nchannels) + " channels, the array " + std::to_string(array.size()));
    wav::Info out_info = in.info();
    out_info.channels = 1;
    wav::Writer out(output_file, out_info);

    mca::SourceSeparationAndLocalisation sss(in.info().sample_rate, array, 1, false, 1, batch_frames);
    std::unique_ptr<McBeamCallback> callback;
    if (os) { callback.reset(new McBeamCallback(*os)); sss.setCallback(*callback); }

    const int nframes = (batch_frames - 1) * sss.getFrameSize();           // samples per process() call
    const int out_len = nframes + sss.getMaxLatency();
    std::vector<double> buffer(size_t(nframes) * nchannels);
    mca::SignalVector audio_channels, output;
    std::vector<double *> raw_audio_channels, raw_output;
    for (int ch = 0; ch < nchannels; ++ch) {
      audio_channels.push_back(mca::SignalPtr(new double[nframes]));
      raw_audio_channels.push_back(audio_channels.back().get());
      output.push_back(mca::SignalPtr(new double[out_len]));
      raw_output.push_back(output.back().get());
    }
    int read_frames;
    while ((read_frames = in.readf_double(buffer.data(), nframes)) > 0) {
      for (int f = 0; f < read_frames; ++f)
        for (int ch = 0; ch < nchannels; ++ch) audio_channels[ch][f] = buffer[size_t(f) * nchannels + ch];
      const int processed = sss.process(raw_audio_channels, read_frames, raw_output, out_len);
      out.writef_double(output[0].get(), processed);                        // mono: channel 0 only (mcabeamf.cpp:114-119)
    }
  } catch (const std::exception &e) {
    std::cout << "error" << std::endl;
    std::cerr << "mcbeam: " << e.what() << std::endl;
    return 2;
  }
  return 0;
}
