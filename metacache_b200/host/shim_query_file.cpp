/******************************************************************************
 * shim_query_file - `metacache query <db> <reads.fa|fq[.gz]> [mates]` below the
 * printing layer: the reference's query_batched (database_query.hpp:170-303) with
 * the shims' reader + worker threads (mcb200::query_files).
 *
 *   usage: shim_query_file <db.cacheN> <file1> [file2|-] [threads] [maxcand]
 *   output: one line per query in file order: header TAB tgt:hits:beg:end,...
 ******************************************************************************/
#include "mcb200_shim.hpp"

#include <cstdio>
#include <iostream>
#include <map>
#include <string>

using namespace mcb200;

int main (int argc, char** argv)
{
    if (argc < 3) { std::cerr << "usage: shim_query_file <db.cacheN> <file1> [file2|-] [threads] [maxcand]\n"; return 2; }
    const std::string file1 = argv[2];
    const std::string file2 = (argc > 3 && std::string(argv[3]) != "-") ? argv[3] : "";
    const unsigned threads = argc > 4 ? unsigned(std::stoul(argv[4])) : 4;
    const std::uint32_t maxcand = argc > 5 ? std::uint32_t(std::stoul(argv[5])) : 2;
    try {
        gpu_hashmap<feature, location> store;
        store.prepare_query_tables(1, 1);
        std::ifstream is(argv[1], std::ios::binary);
        if (!is) { std::cerr << "cannot open " << argv[1] << "\n"; return 1; }
        read_binary(is, store, 0);
        const sketching_opt sk;
        std::map<std::pair<unsigned, std::uint64_t>, std::string> out;      // (thread, batch) = file order
        query_files(store, file1, file2, sk, maxcand, 0, threads, 4096, true,
            [&] (unsigned t, std::uint64_t batchNo, const std::vector<std::string>& headers,
                 query_batch<location>::query_host_data& hd) {
                std::string text;
                for (std::uint32_t s = 0; s < hd.num_queries(); ++s) {
                    text += headers[s]; text += '\t';
                    bool any = false;
                    for (const auto& c : hd.top_candidates(s)) {
                        if (!c.hits) break;
                        char buf[96];
                        std::snprintf(buf, sizeof buf, "%s%u:%u:%u:%u", any ? "," : "", c.tgt, c.hits, c.pos.beg, c.pos.end);
                        text += buf; any = true;
                    }
                    text += '\n';
                }
                out[{t, batchNo}] = std::move(text);
            });
        for (const auto& kv : out) std::fputs(kv.second.c_str(), stdout);
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
