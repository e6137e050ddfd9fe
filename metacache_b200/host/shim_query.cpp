/******************************************************************************
 * shim_query - the reference's `query_gpu` loop (database_query.hpp:87-124)
 * written against mcb200_shim.hpp, to show (and test) that the shims carry the
 * reference's call sequence unchanged:
 *
 *   usage: shim_query <db.cacheN | db> <reads.txt> [maxcand] [copyAllHits]
 *          (<db> = base name of a multi-part database: db.cache0, db.cache1, ...; with several GPUs
 *           prepare_query_tables puts one part on each, in this one process)
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
        std::vector<std::string> files;
        if (std::ifstream(argv[1], std::ios::binary)) files.push_back(argv[1]);
        else for (int p = 0; std::ifstream(std::string(argv[1]) + ".cache" + std::to_string(p), std::ios::binary); ++p)
            files.push_back(std::string(argv[1]) + ".cache" + std::to_string(p));
        if (files.empty()) { std::cerr << "cannot open " << argv[1] << "\n"; return 1; }
        store.prepare_query_tables(part_id(files.size()), 1);       // database::read (database.cpp:183-242)
        for (std::size_t p = 0; p < files.size(); ++p) {
            std::ifstream is(files[p], std::ios::binary);
            read_binary(is, store, part_id(p));                     // database::read_cache
        }
        std::printf("# parts=%u devices=%u\n", unsigned(store.table_count()), unsigned(store.gpu_count()));

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
