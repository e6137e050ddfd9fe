/******************************************************************************
 * shim_reader_dump - prints what mcb200::sequence_pair_reader (the shim's mirror
 * of the reference class, sequence_io.hpp:123-190) returns, one line per query:
 *   index TAB header TAB seq1 TAB seq2
 * usage: shim_reader_dump <file1> [file2] [skip]     (needs no GPU)
 ******************************************************************************/
#include "mcb200_shim.hpp"

#include <cstdio>
#include <iostream>

int main (int argc, char** argv)
{
    if (argc < 2) { std::cerr << "usage: shim_reader_dump <file1> [file2] [skip]\n"; return 2; }
    try {
        mcb200::sequence_pair_reader reader{argv[1], argc > 2 ? argv[2] : ""};
        if (argc > 3) reader.skip(std::stoull(argv[3]));
        while (reader.has_next()) {
            const auto q = reader.next();
            std::printf("%llu\t%s\t%s\t%s\n", (unsigned long long)q.first.index, q.first.header.c_str(),
                        q.first.data.c_str(), q.second.data.c_str());
        }
    } catch (const std::exception& e) {
        std::printf("ERROR\t%s\n", e.what());
        return 1;
    }
    return 0;
}
