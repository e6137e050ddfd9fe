/******************************************************************************
 * shim_query - the reference's `query_gpu` loop (database_query.hpp:87-124)
 * written against mcb200_shim.hpp, to show (and test) that the shims carry the
 * reference's call sequence unchanged:
 *
 *   usage: shim_query <db.cacheN> <reads.txt> [maxcand] [copyAllHits]
 *   reads.txt: one query per line, "SEQ1" or "SEQ1 SEQ2", "-" = empty
 *   output   : one line per query: id TAB tgt:hits:beg:end,... TAB nallhits
 ******************************************************************************/
#include "mcb200_shim.hpp"

#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

using namespace mcb200;

int main (int argc, char** argv)
{
    if (argc < 3) { std::cerr << "usage: shim_query <db.cacheN> <reads.txt> [maxcand] [allhits]\n"; return 2; }
    const std::size_t maxcand = argc > 3 ? std::stoul(argv[3]) : 2;
    const bool allhits = argc > 4 && std::stoi(argv[4]) != 0;
    try {
        gpu_hashmap<feature, location> store;                       // database::featureStore_
        store.prepare_query_tables(1, 1);
        std::ifstream is(argv[1], std::ios::binary);
        if (!is) { std::cerr << "cannot open " << argv[1] << "\n"; return 1; }
        read_binary(is, store, 0);                                  // database::read_cache

        const sketching_opt sk;                                     // db.target_sketching()
        const std::size_t batchSize = 8192;                         // options.hpp:228
        query_batch<location> batch(batchSize, batchSize * sk.winlen, sk.sketchlen, sk.sketchlen * 254,
                                    maxcand, allhits, 1, store.table_count(), 0);
        std::ifstream rs(argv[2]);
        std::string line;
        std::vector<std::pair<std::string, std::string>> reads;
        while (std::getline(rs, line)) {
            const auto sp = line.find(' ');
            std::string a = line.substr(0, sp), b = sp == std::string::npos ? "" : line.substr(sp + 1);
            if (a == "-") a.clear();
            if (b == "-") b.clear();
            reads.emplace_back(a, b);
        }
        std::size_t qid = 0, first = 0;
        auto flush = [&] {
            auto& hd = batch.host_data(0);
            if (hd.num_queries() == 0) return;
            store.query_async(batch, 0, sk, 0);                     // db.query_gpu_async (under scheduleMtx)
            hd.wait_for_results();
            for (std::size_t s = 0; s < hd.num_queries(); ++s) {
                std::printf("%zu\t", first + s);
                bool any = false;
                for (const auto& c : hd.top_candidates(s)) {
                    if (!c.hits) break;
                    std::printf("%s%u:%u:%u:%u", any ? "," : "", c.tgt, c.hits, c.pos.beg, c.pos.end);
                    any = true;
                }
                std::printf("\t%zu\n", hd.allhits(s).size());
            }
            first += hd.num_queries();
            hd.clear();
        };
        for (; qid < reads.size(); ++qid) {
            candidate_generation_rules rules;                       // make_candidate_generation_rules
            rules.maxWindowsInRange = window_id(2 + (reads[qid].first.size() + reads[qid].second.size()) / sk.winstride);
            rules.maxCandidates = maxcand;
            if (!batch.add_paired_read(0, reads[qid].first, reads[qid].second, sk, rules)) {
                flush();
                if (!batch.add_paired_read(0, reads[qid].first, reads[qid].second, sk, rules))
                    std::cerr << "query batch is too small for a single read!\n";
            }
        }
        flush();
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
