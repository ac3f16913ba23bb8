// TEST INFRASTRUCTURE ONLY: compares the parallel reader of the `siga` front-end (siga_b200/csrc/host/seqio.cpp) with
// its serial reader on the files given on the command line.  Prints one line per file: "<path> same <n reads>" or
// "<path> DIFF ...".
#include <cstdio>
#include <string>

#include "../../siga_b200/csrc/host/seqio.h"

int main(int argc, char** argv) {
  int bad = 0;
  for (int i = 1; i < argc; ++i) {
    sgxhost::ReadSet a, b;
    a.keep_quals = b.keep_quals = true;  // FASTQ: the quality strings must agree too (what `siga correct` reads)
    std::string ea, eb;
    const bool ra = sgxhost::read_sequences_serial(argv[i], &a, &ea);
    const bool rb = sgxhost::read_sequences(argv[i], &b, &eb);
    const bool same = ra == rb && a.names == b.names && a.comments == b.comments && a.bases == b.bases && a.off == b.off &&
                      a.quals == b.quals && (a.quals.empty() || a.quals.size() == a.bases.size());
    printf("%s %s %zu%s\n", argv[i], same ? "same" : "DIFF", b.size(), b.quals.empty() ? "" : " q");
    bad += same ? 0 : 1;
  }
  return bad;
}
