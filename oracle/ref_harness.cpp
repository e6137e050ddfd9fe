/******************************************************************************
 * TEST INFRASTRUCTURE ONLY - never linked or executed by the product path.
 *
 * Reference harness: links the UNMODIFIED reference objects (compiled from
 * /root/reference/src by oracle/Makefile) and drives the reference's own
 * query hot path for a list of reads:
 *
 *   mc::sketcher::for_each_sketch          (hash_dna.hpp:207-255)
 *   mc::database::query_host               (database.hpp:399-407)
 *     -> host_hashmap::query_host_hashmap  (host_hashmap.hpp:695-723)
 *   mc::make_candidate_generation_rules    (candidate_structs.hpp:134-151)
 *
 * and dumps, per read, the window sketches, the sorted all-hits list and the
 * top candidates as little-endian binary, so that oracle/mc_oracle.c and the
 * CUDA path can be checked stage by stage against the reference itself.
 * It also serves as the CPU baseline for bench.py (hot path only, T threads).
 *
 * usage: mc_ref_harness <db> <reads.txt> <out.bin|-> [key=value ...]
 *   reads.txt : one query per line, "SEQ1" or "SEQ1 SEQ2" (paired); the token
 *               "-" stands for an empty sequence
 *   keys      : maxcand=2 insert=0 part=-1 threads=1 repeat=1 sketches=1
 *               allhits=1 tops=<file> first=0 count=all
 *   tops=<file>: fixed-size dump of the top candidates only (for parity checks at benchmark scale):
 *               u32 LE [nreads][maxcand][4] = {tgt,hits,beg,end}, unused entries {~0,0,0,0} -
 *               the layout of the library's mcb200_candidate rows.  first/count select a slice of
 *               the reads file.
 * output (all u32 LE):
 *   magic 0x4d435246 ("MCRF"), nreads
 *   per read: nsk, {len, feat[len]} x nsk, nall, {win,tgt} x nall,
 *             ntop, {tgt,hits,beg,end} x ntop
 * stdout: "pass=<i> seconds=<s>" per pass, then one line
 *         "reads=<n> threads=<t> seconds=<best s> reads_per_s=<r> load_seconds=<l>"
 *****************************************************************************/
#include "database.hpp"
#include "query_handler.hpp"
#include "candidate_generation.hpp"
#include "options.hpp"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

using namespace mc;

namespace {

struct read_pair { sequence s1, s2; };

struct seq_query {            // what make_candidate_generation_rules needs
    const sequence& seq1;
    const sequence& seq2;
};

sequence to_seq (const std::string& s, size_t b, size_t e) {
    sequence q;
    if (e - b == 1 && s[b] == '-') return q;
    q.resize(e - b);
    if (e > b) memcpy(q.data(), s.data() + b, e - b);
    return q;
}

struct result_blob { std::vector<uint32_t> w; };

const char* str_arg_of (int argc, char** argv, const char* key) {
    const size_t n = strlen(key);
    for (int i = 4; i < argc; ++i)
        if (!strncmp(argv[i], key, n) && argv[i][n] == '=') return argv[i] + n + 1;
    return nullptr;
}

long arg_of (int argc, char** argv, const char* key, long dflt) {
    const size_t n = strlen(key);
    for (int i = 4; i < argc; ++i)
        if (!strncmp(argv[i], key, n) && argv[i][n] == '=') return atol(argv[i] + n + 1);
    return dflt;
}

} // namespace


int main (int argc, char** argv)
{
    if (argc < 4) {
        std::cerr << "usage: mc_ref_harness <db> <reads.txt> <out.bin|-> [key=value ...]\n";
        return 2;
    }
    const std::string dbname = argv[1], readsFile = argv[2], outFile = argv[3];
    const long maxcand = arg_of(argc, argv, "maxcand", 2);
    const long insert  = arg_of(argc, argv, "insert", 0);
    const long part    = arg_of(argc, argv, "part", -1);
    const long threads = std::max(1L, arg_of(argc, argv, "threads", 1));
    const long repeat  = std::max(1L, arg_of(argc, argv, "repeat", 1));
    const bool wantSk  = arg_of(argc, argv, "sketches", 1) != 0;
    const bool wantAll = arg_of(argc, argv, "allhits", 1) != 0;
    const bool dump    = outFile != "-";
    const char* topsFile = str_arg_of(argc, argv, "tops");
    const long first   = std::max(0L, arg_of(argc, argv, "first", 0));
    const long count   = arg_of(argc, argv, "count", -1);

    const auto tload = std::chrono::steady_clock::now();
    database db = make_database(dbname, int(part), database::scope::everything,
                                info_level::silent);
    const double loadSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tload).count();

    // query sketching = target sketching (querying.cpp:225-266 default)
    const sketching_opt skopt = db.target_sketching();

    classification_options copt;
    copt.insertSizeMax = size_t(insert);
    copt.maxNumCandidatesPerQuery = size_t(maxcand);

    std::vector<read_pair> reads;
    {
        std::ifstream is(readsFile);
        std::string line;
        long lineNo = -1;
        while (std::getline(is, line)) {
            ++lineNo;
            if (lineNo < first) continue;
            if (count >= 0 && lineNo >= first + count) break;
            while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
            const auto sp = line.find(' ');
            read_pair r;
            if (sp == std::string::npos) r.s1 = to_seq(line, 0, line.size());
            else { r.s1 = to_seq(line, 0, sp); r.s2 = to_seq(line, sp + 1, line.size()); }
            reads.push_back(std::move(r));
        }
    }
    const size_t n = reads.size();
    std::vector<result_blob> out(dump ? n : 0);
    std::vector<uint32_t> tops(topsFile ? n * size_t(maxcand) * 4 : 0);

    auto work = [&] (size_t tid, bool record) {
        query_handler<location> handler;
        sketcher sk;
        for (size_t i = tid; i < n; i += size_t(threads)) {
            const auto& r = reads[i];
            seq_query q{r.s1, r.s2};
            auto rules = make_candidate_generation_rules(q, copt, skopt.winstride);
            db.query_host(r.s1, r.s2, handler, skopt, rules);
            if (topsFile) {
                uint32_t* t = tops.data() + i * size_t(maxcand) * 4;
                size_t c = 0;
                for (const auto& cand : handler.tophits()) {
                    if (c >= size_t(maxcand)) break;
                    t[4 * c] = cand.tgt; t[4 * c + 1] = cand.hits; t[4 * c + 2] = cand.pos.beg; t[4 * c + 3] = cand.pos.end;
                    ++c;
                }
                for (; c < size_t(maxcand); ++c) { t[4 * c] = 0xFFFFFFFFu; t[4 * c + 1] = t[4 * c + 2] = t[4 * c + 3] = 0; }
            }
            if (!record) continue;
            auto& w = out[i].w;
            if (wantSk) {
                const size_t pos = w.size();
                w.push_back(0);
                uint32_t nsk = 0;
                auto grab = [&] (const auto& s) {
                    ++nsk; w.push_back(uint32_t(s.size()));
                    for (auto f : s) w.push_back(uint32_t(f));
                };
                sk.for_each_sketch(r.s1.begin(), r.s1.end(), skopt, grab);
                sk.for_each_sketch(r.s2.begin(), r.s2.end(), skopt, grab);
                w[pos] = nsk;
            } else w.push_back(0);
            if (wantAll) {
                auto all = handler.allhits();
                w.push_back(uint32_t(all.size()));
                for (const auto& l : all) { w.push_back(l.win); w.push_back(l.tgt); }
            } else w.push_back(0);
            auto top = handler.tophits();
            w.push_back(uint32_t(top.size()));
            for (const auto& c : top) {
                w.push_back(c.tgt); w.push_back(c.hits);
                w.push_back(c.pos.beg); w.push_back(c.pos.end);
            }
        }
    };

    double best = 1e300;
    for (long rep = 0; rep < repeat; ++rep) {
        const bool record = dump && rep == 0;
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        for (long t = 1; t < threads; ++t) pool.emplace_back(work, size_t(t), record);
        work(0, record);
        for (auto& t : pool) t.join();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("pass=%ld seconds=%.6f\n", rep, s);
        fflush(stdout);
        // a recording pass also pays for the dump; prefer non-recording passes
        if (!record || repeat == 1) best = std::min(best, s);
    }

    if (dump) {
        FILE* f = fopen(outFile.c_str(), "wb");
        if (!f) { std::cerr << "cannot write " << outFile << "\n"; return 1; }
        const uint32_t hdr[2] = {0x4d435246u, uint32_t(n)};
        fwrite(hdr, 4, 2, f);
        for (const auto& b : out) fwrite(b.w.data(), 4, b.w.size(), f);
        fclose(f);
    }
    if (topsFile) {
        FILE* f = fopen(topsFile, "wb");
        if (!f) { std::cerr << "cannot write " << topsFile << "\n"; return 1; }
        fwrite(tops.data(), 4, tops.size(), f);
        fclose(f);
    }
    printf("reads=%zu threads=%ld seconds=%.6f reads_per_s=%.1f load_seconds=%.3f\n",
           n, threads, best, double(n) / best, loadSeconds);
    return 0;
}
