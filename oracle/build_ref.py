#!/usr/bin/env python3
"""Build the REAL reference (vendored Kaldi 5.5 CPU path) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path.

The reference sources are compiled *where they lie* under /root/reference (nothing is
copied into this repository); only objects, one shared library and a handful of
binaries are written to oracle/_ref/ (git-ignored, but shipped to the GPU box).

What is built:
  libkaldi_ref.so : kaldi/src/{base,matrix,util,feat,tree,gmm,transform,hmm,lat,decoder,
                    cudamatrix(CPU mode),chain,nnet3,ivector,online2}/*.cc + kaldi/openfst/src/lib/*.cc
  binaries        : the two decoders rhasspy shells out to
                    (online2-wav-nnet3-latgen-faster, online2-cli-nnet3-decode-faster),
                    the lattice post-processing pair (lattice-to-nbest, nbest-to-linear),
                    and per-stage probes (compute-mfcc-feats, ivector-extract-online2,
                    nnet3-compute, latgen-faster-mapped, lattice-best-path, copy tools).

The reference's own build system (cmake) is NOT run.  BLAS/LAPACK come from the
OpenBLAS 0.3.15 that ships inside the python venv (same image on the GPU box);
headers are the CLAPACK ones vendored in kaldi/tools/CLAPACK.  The one generated
file Kaldi needs (base/version.h, written by base/get_version.sh) is supplied as a
two-line stub in oracle/_ref/include/base/.  Flags follow the reference build
(script/build_kaldi.sh: CMake Release, C++14, HAVE_CLAPACK, no -march => no FMA
contraction, which matters for bit-parity of the restatements).
"""
import glob
import os
import subprocess
import sys

REF = os.environ.get("RS_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
K = os.path.join(REF, "kaldi", "src")
OFST = os.path.join(REF, "kaldi", "openfst", "src")

LIB_DIRS = ["base", "matrix", "util", "feat", "tree", "gmm", "transform", "hmm", "lat",
            "decoder", "cudamatrix", "chain", "nnet3", "ivector", "online2", "fstext"]

# files that pull in model families outside the hot path (nnet2)
EXCLUDE = {"online-nnet2-decoding-threaded.cc", "online-nnet2-decoding.cc"}

BINARIES = {
    "online2-wav-nnet3-latgen-faster": "online2bin",
    "online2-cli-nnet3-decode-faster": "online2bin",
    "ivector-extract-online2": "online2bin",
    "lattice-to-nbest": "latbin",
    "nbest-to-linear": "latbin",
    "lattice-best-path": "latbin",
    "lattice-copy": "latbin",
    "compute-mfcc-feats": "featbin",
    "copy-feats": "featbin",
    "nnet3-compute": "nnet3bin",
    "nnet3-am-copy": "nnet3bin",
    "nnet3-am-info": "nnet3bin",
    "latgen-faster-mapped": "bin",
    "copy-matrix": "bin",
    "copy-vector": "bin",
    "copy-transition-model": "bin",
    "gmm-global-copy": "gmmbin",
    "ivector-extractor-copy": "ivectorbin",
}


def find_openblas():
    import site
    cands = []
    for sp in site.getsitepackages() + [os.path.dirname(os.path.dirname(os.__file__))]:
        cands += glob.glob(os.path.join(sp, "opencv_python_headless.libs", "libopenblas*.so*"))
        cands += glob.glob(os.path.join(sp, "**", "libopenblasp-r0-*.3.15.so"), recursive=False)
    cands += glob.glob("/usr/lib/x86_64-linux-gnu/libopenblas.so*")
    if not cands:
        raise SystemExit("no OpenBLAS found for the reference build")
    return cands[0]


def lib_sources():
    srcs = []
    for d in LIB_DIRS:
        for f in sorted(glob.glob(os.path.join(K, d, "*.cc"))):
            b = os.path.basename(f)
            if b.endswith("-test.cc") or b.endswith("_test.cc"):
                continue
            if b in EXCLUDE:
                continue
            with open(f, "r", errors="replace") as fh:
                txt = fh.read()
            if "int main(" in txt or "int main (" in txt:
                continue
            srcs.append(f)
    for f in sorted(glob.glob(os.path.join(OFST, "lib", "*.cc"))):
        srcs.append(f)
    return srcs


def main():
    if not os.path.isdir(K):
        print("reference tree not present (%s): using prebuilt oracle/_ref if any" % K)
        return 0
    os.makedirs(os.path.join(OUT, "include", "base"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "bin"), exist_ok=True)
    vh = os.path.join(OUT, "include", "base", "version.h")
    if not os.path.exists(vh):
        with open(vh, "w") as f:
            f.write('// stub for the file base/get_version.sh would generate\n#define KALDI_VERSION "5.5-ref"\n')
    blas = find_openblas()
    blas_dir = os.path.dirname(blas)
    cxxflags = ("-std=c++14 -O2 -fPIC -w -pthread -DKALDI_DOUBLEPRECISION=0 -DHAVE_EXECINFO_H=1 "
                "-DHAVE_CXXABI_H -DHAVE_CLAPACK=1 -DKALDI_NO_PORTAUDIO=1 -DNDEBUG "
                "-I%s -I%s -I%s -I%s" % (os.path.join(OUT, "include"), K,
                                         os.path.join(OFST, "include"),
                                         os.path.join(REF, "kaldi", "tools", "CLAPACK")))
    ldflags = "-pthread -Wl,--disable-new-dtags -Wl,-rpath,'$$ORIGIN/..' -Wl,-rpath,%s -L%s -L%s -lkaldi_ref -l:%s -ldl -lm" % (
        blas_dir, OUT, blas_dir, os.path.basename(blas))
    n = ["cxx = g++", "cxxflags = " + cxxflags, "",
         "rule cc", "  command = $cxx $cxxflags -MMD -MF $out.d -c $in -o $out",
         "  depfile = $out.d", "  deps = gcc", "  description = CC $out", "",
         "rule solib",
         "  command = $cxx -shared -o $out @$out.rsp -Wl,--disable-new-dtags -Wl,-rpath,%s -l:%s -L%s -lpthread -ldl -lm" % (
             blas_dir, os.path.basename(blas), blas_dir),
         "  rspfile = $out.rsp", "  rspfile_content = $in", "  description = SOLIB $out", "",
         "rule link", "  command = $cxx -o $out $in %s" % ldflags, "  description = LINK $out", ""]
    objs = []
    for s in lib_sources():
        rel = os.path.relpath(s, os.path.join(REF, "kaldi")).replace("/", "__")
        o = os.path.join(OUT, "obj", rel[:-3] + ".o")
        objs.append(o)
        n.append("build %s: cc %s" % (o, s))
    lib = os.path.join(OUT, "libkaldi_ref.so")
    n.append("build %s: solib %s" % (lib, " ".join(objs)))
    targets = []
    for b, d in BINARIES.items():
        s = os.path.join(K, d, b + ".cc")
        o = os.path.join(OUT, "obj", "bin__" + b + ".o")
        e = os.path.join(OUT, "bin", b)
        n.append("build %s: cc %s" % (o, s))
        n.append("build %s: link %s | %s" % (e, o, lib))
        targets.append(e)
    # intermediate-value probe (our own source, linked against the reference library)
    probe_o = os.path.join(OUT, "obj", "bin__ref-probe.o")
    probe_e = os.path.join(OUT, "bin", "ref-probe")
    n.append("build %s: cc %s" % (probe_o, os.path.join(HERE, "ref_probe.cc")))
    n.append("build %s: link %s | %s" % (probe_e, probe_o, lib))
    targets.append(probe_e)
    # OpenFst side of the fuzzy matcher (fstcompile / fstcompose / fstshortestpath / ... through the library)
    fz_o = os.path.join(OUT, "obj", "bin__fuzzy-probe.o")
    fz_e = os.path.join(OUT, "bin", "fuzzy-probe")
    n.append("build %s: cc %s" % (fz_o, os.path.join(HERE, "fuzzy_probe.cc")))
    n.append("build %s: link %s | %s" % (fz_e, fz_o, lib))
    targets.append(fz_e)
    n.append("default " + " ".join(targets))
    with open(os.path.join(OUT, "build.ninja"), "w") as f:
        f.write("\n".join(n) + "\n")
    jobs = os.environ.get("RS_REF_JOBS", str(os.cpu_count() or 4))
    r = subprocess.run(["ninja", "-C", OUT, "-j", jobs] + sys.argv[1:])
    return r.returncode


if __name__ == "__main__":
    sys.exit(main())
