// Test infrastructure only: dumps what the UNMODIFIED reference reader (sequence_io.cpp, linked from
// oracle/_ref/obj) returns for a file or file pair, one line per query:
//   index \t header \t seq1 \t seq2
// usage: mc_ref_reader <file1> [file2]      (file2 == file1: pairs of consecutive sequences)
#include "sequence_io.hpp"
#include <cstdio>
#include <string>

int main (int argc, char** argv) {
    if (argc < 2) return 2;
    try {
        mc::sequence_pair_reader reader{argv[1], argc > 2 ? argv[2] : ""};
        while (reader.has_next()) {
            auto q = reader.next();
            std::string h = q.first.header;
            std::string a(q.first.data.begin(), q.first.data.end());
            std::string b(q.second.data.begin(), q.second.data.end());
            printf("%llu\t%s\t%s\t%s\n", (unsigned long long)q.first.index, h.c_str(), a.c_str(), b.c_str());
        }
    } catch (std::exception& e) {
        printf("ERROR\t%s\n", e.what());
        return 1;
    }
    return 0;
}
