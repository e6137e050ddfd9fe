/******************************************************************************
 * mc_oracle.c - TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement, in plain C, of MetaCache's query hot path as implemented by
 * the reference CPU code.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product path
 * (metacache_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file
 * against (1) the known-answer vectors generated from the reference headers
 * (SURVEY.md 4.3), (2) fixtures produced by the reference itself
 * (oracle/ref_harness.cpp linking the unmodified reference objects; committed
 * under tests/golden/ with the generating script oracle/make_golden.py) and
 * (3) the reference's own golden file test/data/classified.expected (per-read
 * all_hits/top_hits columns) when oracle/_ref/c1 is present.
 *
 * Every function names the reference file:line whose behaviour it restates.
 * Nothing here is copied from the reference; it is re-derived from behaviour.
 *****************************************************************************/
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t k;          /* kmerlen   (1..16)                */
    uint32_t s;          /* sketchlen (>=1)                  */
    uint32_t w;          /* winlen                           */
    uint32_t stride;     /* winstride                        */
} mco_sketch_opt;

/* ---------------------------------------------------------------------------
 * hash_int.hpp:41-48  thomas_mueller_hash, the feature hash for 32-bit k-mers
 * (same_size_hash<uint32_t>, hash_int.hpp:171-177).                          */
uint32_t mco_hash32 (uint32_t x)
{
    x = ((x >> 16) ^ x) * 0x45d9f3bu;
    x = ((x >> 16) ^ x) * 0x45d9f3bu;
    return (x >> 16) ^ x;
}

/* dna_encoding.hpp:168-177  reverse complement of a 2-bit encoded k-mer:
 * reverse the order of the 2-bit groups of the 32-bit word, complement all
 * bits, keep the top k groups shifted down.                                   */
uint32_t mco_revcomp32 (uint32_t x, uint32_t k)
{
    uint32_t r = 0;
    for (int i = 0; i < 16; ++i) {          /* group i -> group 15-i */
        r |= ((x >> (2 * i)) & 3u) << (2 * (15 - i));
    }
    r = ~r;
    return r >> (32 - 2 * k);
}

/* dna_encoding.hpp:215-226  canonical = min(kmer, revcomp) */
uint32_t mco_canonical32 (uint32_t x, uint32_t k)
{
    const uint32_t rc = mco_revcomp32(x, k);
    return x < rc ? x : rc;
}

/* hash_dna.hpp:54-75  number of windows for_each_window visits.
 * len <= w: one window.  Otherwise every full window at i*stride, plus one
 * trailing partial window if the next start is still inside the sequence.   */
uint64_t mco_num_windows (uint64_t len, uint32_t w, uint32_t stride)
{
    if (len <= w) return 1;
    uint64_t full = (len - w) / stride + 1;
    return full + ((full * stride < len) ? 1 : 0);
}

/* window i covers [beg, end) */
static void window_bounds (uint64_t len, uint32_t w, uint32_t stride, uint64_t i,
                           uint64_t* beg, uint64_t* end)
{
    if (len <= w) { *beg = 0; *end = len; return; }
    *beg = i * stride;
    *end = (*beg + w <= len) ? *beg + w : len;
}

/* dna_encoding.hpp:270-316 (rolling 2-bit encoder with ambiguity shift register)
 * + dna_encoding.hpp:433-444 (skip ambiguous, canonicalise)
 * + hash_dna.hpp:219-252 (s smallest UNIQUE hashes, ascending, ~0 sentinels
 *   stripped; s = min(sketchlen, n-k+1); windows with n < k are not sketched).
 * Returns number of features written to out (<= opt->s), or -1 if the window
 * is skipped by the reference (n < k) - such a window produces no sketch.   */
int mco_sketch_window (const char* seq, uint64_t n, const mco_sketch_opt* opt, uint32_t* out)
{
    const uint32_t k = opt->k;
    if (n < k) return -1;
    uint64_t s64 = n - k + 1;
    uint32_t s = (s64 < opt->s) ? (uint32_t)s64 : opt->s;
    if (s < 1) return -1;

    uint32_t* sk = (uint32_t*)malloc(sizeof(uint32_t) * s);
    for (uint32_t i = 0; i < s; ++i) sk[i] = 0xFFFFFFFFu;

    const uint32_t kmask = (k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    const uint32_t amask = (k >= 16) ? 0xFFFFu : ((1u << k) - 1u);
    uint32_t kmer = 0, ambig = 0;
    for (uint64_t i = 0; i < n; ++i) {
        kmer <<= 2; ambig <<= 1;
        switch (seq[i]) {
            case 'A': case 'a': break;
            case 'C': case 'c': kmer |= 1; break;
            case 'G': case 'g': kmer |= 2; break;
            case 'T': case 't': case 'U': case 'u': kmer |= 3; break;
            default: ambig |= 1; break;
        }
        if (i + 1 < k) continue;
        kmer &= kmask; ambig &= amask;
        if (ambig) continue;
        const uint32_t h = mco_hash32(mco_canonical32(kmer, k));
        if (h < sk[s - 1]) {
            /* lower_bound */
            uint32_t lo = 0, hi = s;
            while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (sk[mid] < h) lo = mid + 1; else hi = mid; }
            if (lo < s && sk[lo] != h) {
                memmove(sk + lo + 1, sk + lo, sizeof(uint32_t) * (s - 1 - lo));
                sk[lo] = h;
            }
        }
    }
    uint32_t m = 0;
    while (m < s && sk[m] != 0xFFFFFFFFu) { out[m] = sk[m]; ++m; }
    free(sk);
    return (int)m;
}

/* hash_dna.hpp:207-255  all window sketches of one sequence.
 * feats: [nwin][opt->s] (row-major, unused tail undefined), counts[nwin] gets
 * the feature count or -1 for windows the reference does not sketch.
 * Returns the number of windows.                                             */
uint64_t mco_sketch_sequence (const char* seq, uint64_t len, const mco_sketch_opt* opt,
                              uint32_t* feats, int32_t* counts, uint64_t max_windows)
{
    const uint64_t nw = mco_num_windows(len, opt->w, opt->stride);
    for (uint64_t i = 0; i < nw && i < max_windows; ++i) {
        uint64_t b, e;
        window_bounds(len, opt->w, opt->stride, i, &b, &e);
        counts[i] = mco_sketch_window(seq + b, e - b, opt, feats + i * opt->s);
    }
    return nw;
}

/* ---------------------------------------------------------------------------
 * feature -> bucket lookup.  hash_multimap.hpp:1086-1098 (find_occupied_slot)
 * returns the bucket whose key equals the feature or "not found"; the probing
 * order (quadratic on key % nbuckets, hash_multimap.hpp:135-175) does not
 * influence WHAT is found, so any exact-match dictionary restates it.  We use
 * linear probing on a power-of-two table.                                    */
typedef struct {
    uint64_t  nslots;        /* power of two */
    uint32_t* slot_key;
    uint64_t* slot_idx;      /* bucket index + 1, 0 = empty */
    uint64_t  nkeys;
    uint64_t* offsets;       /* [nkeys+1] */
    const uint64_t* values;  /* (tgt << 32) | win, owned copy */
    uint64_t* values_own;
} mco_table;

static uint64_t mix64 (uint64_t x) { x *= 0x9E3779B97F4A7C15ull; return x ^ (x >> 29); }

mco_table* mco_table_build (const uint32_t* keys, const uint8_t* sizes, uint64_t nkeys,
                            const uint64_t* values, uint64_t nvalues)
{
    mco_table* t = (mco_table*)calloc(1, sizeof(mco_table));
    uint64_t ns = 16; while (ns < nkeys * 2) ns <<= 1;
    t->nslots = ns; t->nkeys = nkeys;
    t->slot_key = (uint32_t*)calloc(ns, sizeof(uint32_t));
    t->slot_idx = (uint64_t*)calloc(ns, sizeof(uint64_t));
    t->offsets  = (uint64_t*)malloc(sizeof(uint64_t) * (nkeys + 1));
    t->values_own = (uint64_t*)malloc(sizeof(uint64_t) * (nvalues ? nvalues : 1));
    memcpy(t->values_own, values, sizeof(uint64_t) * nvalues);
    t->values = t->values_own;
    uint64_t off = 0;
    for (uint64_t i = 0; i < nkeys; ++i) {
        t->offsets[i] = off; off += sizes[i];
        uint64_t p = mix64(keys[i]) & (ns - 1);
        while (t->slot_idx[p]) p = (p + 1) & (ns - 1);
        t->slot_key[p] = keys[i]; t->slot_idx[p] = i + 1;
    }
    t->offsets[nkeys] = off;
    return t;
}

void mco_table_free (mco_table* t)
{
    if (!t) return;
    free(t->slot_key); free(t->slot_idx); free(t->offsets); free(t->values_own); free(t);
}

/* returns bucket size (0 = not found) and sets *first to its first value */
uint32_t mco_table_find (const mco_table* t, uint32_t key, const uint64_t** first)
{
    uint64_t p = mix64(key) & (t->nslots - 1);
    while (t->slot_idx[p]) {
        if (t->slot_key[p] == key) {
            const uint64_t i = t->slot_idx[p] - 1;
            *first = t->values + t->offsets[i];
            return (uint32_t)(t->offsets[i + 1] - t->offsets[i]);
        }
        p = (p + 1) & (t->nslots - 1);
    }
    return 0;
}

/* ---------------------------------------------------------------------------
 * growable u64 list + run boundaries (query_handler.hpp:37-72 matches_sorter) */
typedef struct { uint64_t* v; uint64_t n, cap; } vec64;
static void vpush (vec64* a, const uint64_t* src, uint64_t n)
{
    if (a->n + n > a->cap) {
        while (a->n + n > a->cap) a->cap = a->cap ? a->cap * 2 : 1024;
        a->v = (uint64_t*)realloc(a->v, a->cap * sizeof(uint64_t));
    }
    memcpy(a->v + a->n, src, n * sizeof(uint64_t)); a->n += n;
}

/* query_handler.hpp:75-101  bottom-up pairwise merge of the appended bucket
 * runs.  The order relation is location::operator< = (tgt, win)
 * (database.hpp:151-156); with locations packed as (tgt<<32)|win that is plain
 * u64 '<'.  std::merge is stable (left run first on ties).  The reference's
 * buffer ping-pong is an implementation detail; for a single part its result
 * is the stable merge of all runs, which is what is restated here.  (For more
 * than one part the reference's ping-pong corrupts earlier parts - SURVEY F5;
 * the oracle implements the evidently intended per-part sort.)              */
static void merge_runs (uint64_t* a, uint64_t* tmp, const uint64_t* off, uint64_t nruns)
{
    if (nruns < 2) return;
    uint64_t* src = a; uint64_t* dst = tmp;
    const uint64_t base = off[0], total = off[nruns];
    for (uint64_t s = 1; s < nruns; s *= 2) {
        for (uint64_t i = 0; i < nruns; i += 2 * s) {
            const uint64_t b = off[i];
            const uint64_t m = (i + s <= nruns) ? off[i + s] : off[nruns];
            const uint64_t e = (i + 2 * s <= nruns) ? off[i + 2 * s] : off[nruns];
            uint64_t x = b, y = m, o = b;
            while (x < m && y < e) dst[o++] = (src[y] < src[x]) ? src[y++] : src[x++];
            while (x < m) dst[o++] = src[x++];
            while (y < e) dst[o++] = src[y++];
        }
        uint64_t* t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a + base, src + base, (total - base) * sizeof(uint64_t));
}

typedef struct { uint32_t tgt, hits, beg, end; } mco_candidate;

/* candidate_generation.hpp:172-231  best_distinct_matches_...::insert.
 * tax_of_tgt == NULL  <=> mergeBelow == Sequence (every target is its own
 * taxon, cached_taxon_of_target never null for valid targets).
 * Otherwise tax_of_tgt[tgt] is an opaque non-zero taxon key at the merge rank
 * (0 = no ancestor at/above that rank => candidate dropped, :191).           */
static void top_insert (mco_candidate* top, uint64_t* toptax, uint32_t* ntop, uint32_t maxc,
                        mco_candidate c, const uint64_t* tax_of_tgt)
{
    if (maxc == 0) return;
    if (*ntop == maxc && top[*ntop - 1].hits >= c.hits) return;      /* early exit :180 */
    uint64_t tax = tax_of_tgt ? tax_of_tgt[c.tgt] : (uint64_t)c.tgt + 1;
    if (!tax) return;
    if (tax_of_tgt) {
        uint32_t i = 0;
        while (i < *ntop && toptax[i] != tax) ++i;
        if (i < *ntop) {                                             /* :209-214 */
            if (c.hits > top[i].hits) {
                top[i] = c;
                /* std::sort(begin, i+1, greater): only element i is out of
                 * place (it grew); sort is not stable, but equal-hits
                 * neighbours are distinguishable - restate as insertion from
                 * the right, moving past strictly smaller hits only.         */
                while (i > 0 && top[i - 1].hits < top[i].hits) {
                    mco_candidate t = top[i - 1]; top[i - 1] = top[i]; top[i] = t;
                    uint64_t x = toptax[i - 1]; toptax[i - 1] = toptax[i]; toptax[i] = x;
                    --i;
                }
            }
            return;
        }
    }
    /* upper_bound with greater: first position whose hits < c.hits */
    uint32_t pos = 0;
    while (pos < *ntop && top[pos].hits >= c.hits) ++pos;
    if (pos < *ntop || *ntop < maxc) {
        uint32_t last = (*ntop < maxc) ? *ntop : maxc - 1;   /* element that falls off if full */
        for (uint32_t j = last; j > pos; --j) { top[j] = top[j - 1]; toptax[j] = toptax[j - 1]; }
        top[pos] = c; toptax[pos] = tax;
        if (*ntop < maxc) ++*ntop;
    }
}

/* candidate_generation.hpp:47-108  for_all_contiguous_window_ranges:
 * per target a two-pointer scan; hits counts list ENTRIES (multiplicity) whose
 * window lies within maxWin of the left end; the first strictly best range
 * wins; one candidate per target, emitted in list order.                    */
uint32_t mco_candidates (const uint64_t* locs, uint64_t n, uint32_t max_win,
                         uint32_t maxc, const uint64_t* tax_of_tgt, mco_candidate* top)
{
    uint32_t ntop = 0;
    if (n == 0 || maxc == 0) return 0;
    uint64_t* toptax = (uint64_t*)calloc(maxc, sizeof(uint64_t));
    uint64_t fst = 0;
    uint32_t hits = 1;
    mco_candidate cur = { (uint32_t)(locs[0] >> 32), 1, (uint32_t)locs[0], (uint32_t)locs[0] };
    for (uint64_t lst = 1; lst < n; ++lst) {
        const uint32_t tgt = (uint32_t)(locs[lst] >> 32), win = (uint32_t)locs[lst];
        if (tgt == cur.tgt) {
            ++hits;
            while (fst != lst && (uint32_t)(win - (uint32_t)locs[fst]) >= max_win) { --hits; ++fst; }
            if (hits > cur.hits) { cur.hits = hits; cur.beg = (uint32_t)locs[fst]; cur.end = win; }
        } else {
            top_insert(top, toptax, &ntop, maxc, cur, tax_of_tgt);
            fst = lst; hits = 1;
            cur.tgt = tgt; cur.hits = 1; cur.beg = win; cur.end = win;
        }
    }
    top_insert(top, toptax, &ntop, maxc, cur, tax_of_tgt);
    free(toptax);
    return ntop;
}

/* candidate_structs.hpp:134-151 */
uint32_t mco_max_windows_in_range (uint64_t len1, uint64_t len2, uint64_t insert_size_max,
                                   uint32_t winstride)
{
    uint64_t m = len1 + len2; if (insert_size_max > m) m = insert_size_max;
    return (uint32_t)(2 + m / winstride);
}

/* host_hashmap.hpp:629-723  query_host_hashmap for nparts tables:
 * part 0: sketch both mates window by window, look every feature up, append
 * the bucket as one run, merge-sort; parts 1..: look up the concatenated
 * sketch, append, sort that part's region; then candidates over the whole
 * list (part-major).  allhits (may be NULL) receives up to allhits_cap sorted
 * locations; *n_all the true count.  Returns number of top candidates.       */
uint32_t mco_query (const mco_table* const* tables, uint32_t nparts,
                    const char* s1, uint64_t len1, const char* s2, uint64_t len2,
                    const mco_sketch_opt* opt, uint32_t max_win, uint32_t maxc,
                    const uint64_t* tax_of_tgt,
                    uint64_t* allhits, uint64_t allhits_cap, uint64_t* n_all,
                    mco_candidate* top)
{
    vec64 locs = {0, 0, 0}, runs = {0, 0, 0};
    /* all features of all windows of both mates, in order */
    uint32_t* feats = 0; uint64_t nfeat = 0;
    for (int mate = 0; mate < 2; ++mate) {
        const char* s = mate ? s2 : s1; const uint64_t len = mate ? len2 : len1;
        /* for_each_window on an empty sequence: len(0) <= w -> one window with
         * n = 0 < k -> not sketched.                                         */
        const uint64_t nw = mco_num_windows(len, opt->w, opt->stride);
        uint32_t* f = (uint32_t*)malloc(sizeof(uint32_t) * (nw * opt->s + 1));
        int32_t* c = (int32_t*)malloc(sizeof(int32_t) * (nw + 1));
        mco_sketch_sequence(s, len, opt, f, c, nw);
        feats = (uint32_t*)realloc(feats, sizeof(uint32_t) * (nfeat + nw * opt->s + 1));
        for (uint64_t i = 0; i < nw; ++i)
            for (int32_t j = 0; j < c[i]; ++j) feats[nfeat++] = f[i * opt->s + j];
        free(f); free(c);
    }
    for (uint32_t p = 0; p < nparts; ++p) {
        runs.n = 0;
        uint64_t start = locs.n; vpush(&runs, &start, 1);          /* sorter.next() */
        for (uint64_t i = 0; i < nfeat; ++i) {
            const uint64_t* first; const uint32_t sz = mco_table_find(tables[p], feats[i], &first);
            if (sz) { vpush(&locs, first, sz); uint64_t e = locs.n; vpush(&runs, &e, 1); }
        }
        if (runs.n >= 3) {
            uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * locs.n);
            merge_runs(locs.v, tmp, runs.v, runs.n - 1);
            free(tmp);
        }
    }
    free(feats);
    if (n_all) *n_all = locs.n;
    if (allhits) memcpy(allhits, locs.v, sizeof(uint64_t) * (locs.n < allhits_cap ? locs.n : allhits_cap));
    const uint32_t ntop = mco_candidates(locs.v, locs.n, max_win, maxc, tax_of_tgt, top);
    free(locs.v); free(runs.v);
    return ntop;
}

/* Stable merge of per-part top lists in part order: what the reference's
 * documented partitioned flow computes (docs/partitioning.md:116-142 +
 * mode_merge.cpp:158-240 re-inserting candidates through the same insert()).
 * parts_top: [nparts][maxc], parts_n[nparts].                                */
uint32_t mco_merge_tops (const mco_candidate* parts_top, const uint32_t* parts_n,
                         uint32_t nparts, uint32_t maxc, mco_candidate* top)
{
    uint32_t ntop = 0;
    uint64_t* toptax = (uint64_t*)calloc(maxc ? maxc : 1, sizeof(uint64_t));
    for (uint32_t p = 0; p < nparts; ++p)
        for (uint32_t i = 0; i < parts_n[p]; ++i)
            top_insert(top, toptax, &ntop, maxc, parts_top[(uint64_t)p * maxc + i], 0);
    free(toptax);
    return ntop;
}

/* ---------------------------------------------------------------------------
 * classification.cpp:146-189  classify(): LCA over the ranked lineages of the
 * candidates whose hits exceed (hits0 - hitsMin) * hitsDiffFraction (float).
 * taxonomy.hpp:1291-1301 ranked_lca: first rank >= lowest where both lineages
 * hold the same non-null taxon.  lineages: [n_targets][21], taxon ordinal + 1,
 * 0 = none (taxonomy.hpp:576-597 make_ranks; index = rank).  `lowest` is the
 * rank candidates were generated at (cand.tax = lowest_ranked_ancestor).
 * Returns taxon ordinal + 1 (0 = unclassified); *rank_out = its rank.          */
uint32_t mco_classify (const mco_candidate* top, uint32_t ntop, const uint32_t* lineages,
                       uint32_t n_targets, uint32_t hits_min, float hits_diff_fraction,
                       uint32_t lowest, uint32_t highest, uint32_t* rank_out)
{
    if (rank_out) *rank_out = 21;
    if (ntop == 0 || top[0].tgt >= n_targets) return 0;
    const uint32_t* lin0 = lineages + (uint64_t)top[0].tgt * 21;
    uint32_t r = lowest;
    while (r < 21 && !lin0[r]) ++r;                     /* cand[0].tax */
    if (r >= 21) return 0;
    if (top[0].hits < hits_min) return 0;
    uint32_t lca = lin0[r];
    const float threshold = top[0].hits > hits_min ? (float)(top[0].hits - hits_min) * hits_diff_fraction : 0.0f;
    for (uint32_t i = 1; i < ntop; ++i) {
        if (!((float)top[i].hits > threshold)) break;
        if (top[i].tgt >= n_targets) return 0;
        const uint32_t* lin = lineages + (uint64_t)top[i].tgt * 21;
        uint32_t x = r;
        while (x <= 20 && !(lin0[x] && lin0[x] == lin[x])) ++x;
        if (x > 20) return 0;
        r = x; lca = lin0[x];
        if (r > highest) return 0;
    }
    if (r > highest) return 0;
    if (rank_out) *rank_out = r;
    return lca;
}

#ifdef __cplusplus
}
#endif
