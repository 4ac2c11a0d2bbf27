// Probe compiled against the REFERENCE (oracle/_ref/libkaldi_ref.so) by oracle/build_ref.py.
// TEST INFRASTRUCTURE ONLY.  Dumps the reference's intermediate values that no Kaldi binary exposes:
// for each 400-sample frame read from stdin the windowed frame (feature-window.cc ExtractWindow),
// its packed split-radix spectrum (srfft.cc) and the MFCC (feature-mfcc.cc), as raw float32.
// Used to pin the bit-exactness of the FFT restatement (tests/golden/make_golden.py).
#include "matrix/srfft.h"
#include "feat/feature-mfcc.h"
#include "feat/feature-window.h"
#include <cstdio>
using namespace kaldi;
int main(int argc, char **argv) {
  // stdin: int32 nframes, then nframes*400 float samples; stdout: windows[512], fft[512], mel-log[40], mfcc[40] per frame
  int32 n; fread(&n, 4, 1, stdin);
  MfccOptions opts; opts.mel_opts.num_bins = 40; opts.num_ceps = 40; opts.use_energy = false; opts.frame_opts.dither = 0.0;
  opts.mel_opts.low_freq = 20; opts.mel_opts.high_freq = -400;
  FeatureWindowFunction wf(opts.frame_opts);
  MfccComputer comp(opts);
  SplitRadixRealFft<BaseFloat> fft(512);
  for (int i = 0; i < n; i++) {
    Vector<BaseFloat> wave(400); fread(wave.Data(), 4, 400, stdin);
    Vector<BaseFloat> window;
    ExtractWindow(0, wave, 0, opts.frame_opts, wf, &window, NULL);
    fwrite(window.Data(), 4, 512, stdout);
    Vector<BaseFloat> f(window);
    fft.Compute(f.Data(), true);
    fwrite(f.Data(), 4, 512, stdout);
    Vector<BaseFloat> feat(40);
    comp.Compute(0.0, 1.0, &window, &feat);
    fwrite(feat.Data(), 4, 40, stdout);
  }
}
