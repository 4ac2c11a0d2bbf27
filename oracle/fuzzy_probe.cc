// Probe compiled against the REFERENCE's OpenFst (headers under /root/reference/kaldi/openfst/src/include, library
// code inside oracle/_ref/libkaldi_ref.so) by oracle/build_ref.py.  TEST INFRASTRUCTURE ONLY.
//
// Runs, with the reference's own library classes, the two OpenFst command lines of rhasspy-speech's fuzzy matcher:
//   compile <text.fst> <words.txt> <out.fst>
//       fstcompile --isymbols=words.txt --osymbols=words.txt --keep_isymbols=true --keep_osymbols=true
//       (rhasspy_speech/kaldi.py:390-407: how G.fuzzy.fst is produced)
//   fuzzy <G.fuzzy.fst> <words.txt>      (text FST of the n-best hypotheses on stdin)
//       fstcompile | fstcompose - G.fuzzy.fst | fstshortestpath | fstrmepsilon | fsttopsort |
//       fstproject --project_type=output | fstprint --osymbols=words.txt
//       (rhasspy_speech/transcribe_util.py:46-60); prints what fstprint prints.
#include <fst/fstlib.h>
#include <fst/script/compile-impl.h>
#include <fst/script/print-impl.h>

#include <fstream>
#include <iostream>
#include <memory>

using fst::StdArc;
using fst::StdVectorFst;
using fst::SymbolTable;

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  const std::string mode = argv[1];
  if (mode == "compile" && argc == 5) {
    std::ifstream in(argv[2]);
    std::unique_ptr<SymbolTable> syms(SymbolTable::ReadText(argv[3]));
    if (!in || !syms) return 3;
    fst::FstCompiler<StdArc> comp(in, argv[2], syms.get(), syms.get(), nullptr, false, true, true, false, false);
    StdVectorFst out(comp.Fst());
    return out.Write(argv[4]) ? 0 : 4;
  }
  if (mode == "fuzzy" && argc == 4) {
    fst::FstCompiler<StdArc> comp(std::cin, "stdin", nullptr, nullptr, nullptr, false, false, false, false, false);
    StdVectorFst input(comp.Fst());
    std::unique_ptr<StdVectorFst> fuzzy(StdVectorFst::Read(argv[2]));
    std::unique_ptr<SymbolTable> syms(SymbolTable::ReadText(argv[3]));
    if (!fuzzy || !syms) return 3;
    StdVectorFst composed, best;
    fst::Compose(input, *fuzzy, &composed);
    fst::ShortestPath(composed, &best);
    fst::RmEpsilon(&best);
    fst::TopSort(&best);
    fst::Project(&best, fst::PROJECT_OUTPUT);
    fst::FstPrinter<StdArc> printer(best, nullptr, syms.get(), nullptr, false, false, "\t");
    printer.Print(&std::cout, "stdout");
    return 0;
  }
  return 2;
}
