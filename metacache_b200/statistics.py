"""Per-rank classification statistics of a query run (row N3 of SURVEY.md 8f): the counters of
`classification_statistics::assign` (classification_statistics.hpp:41-200) and the summary block of
`show_taxon_statistics` (printing.cpp:502-540), fed with the (taxon, rank) pairs the device classifier
returns (`QueryHostData.classifications()`), a whole batch per call; and the per-taxon read counts with
the abundance estimation of `-abundances` / `-abundance-per <rank>` (`estimate_abundance`,
classification.cpp:304-377; tables of printing.cpp:424-497)."""
from __future__ import annotations

import numpy as np

from .formatting import RANK_NAMES

RANK_NONE = 21
# the ranks the reference prints (printing.cpp:506-513)
_SHOWN = [0, 3, 4, 6, 10, 12, 14, 16, 18, 19, 20]


class ClassificationStatistics:
    def __init__(self):
        self.assigned_ = np.zeros(RANK_NONE + 1, np.int64)

    def assign_batch(self, classifications: np.ndarray):
        """classifications: [n, 2] (taxon ordinal + 1 or 0 = unclassified, rank); one assign() per row"""
        c = np.asarray(classifications).reshape(-1, 2)
        rank = np.where(c[:, 0] == 0, RANK_NONE, c[:, 1]).astype(np.int64)
        self.assigned_ += np.bincount(rank, minlength=RANK_NONE + 1)[:RANK_NONE + 1]

    def assign(self, rank: int):
        self.assigned_[rank] += 1

    def assigned(self, rank: int = 20) -> int:
        """number of assignments on a rank and below it (more specific), assigned(rank) of the reference"""
        return int(self.assigned_[:rank + 1].sum())

    def unassigned(self) -> int:
        return int(self.assigned_[RANK_NONE])

    def total(self) -> int:
        return self.assigned() + self.unassigned()

    def classification_rate(self, rank: int) -> float:
        return self.assigned(rank) / self.total() if self.total() else 0.0

    def unclassified_rate(self) -> float:
        return self.unassigned() / self.total() if self.total() else 0.0

    def summary_lines(self, prefix: str = "# "):
        """show_taxon_statistics without ground truth (numbers print like a default C++ ostream: %g)"""
        if self.assigned() < 1:
            return ["None of the input sequences could be classified."]
        out = []
        if self.unassigned() > 0:
            out.append(f"{prefix}unclassified: {100 * self.unclassified_rate():g}% ({self.unassigned()})")
        out.append(f"{prefix}classified:")
        for r in _SHOWN:
            if self.assigned(r) > 0:
                out.append(f"{prefix}  {RANK_NAMES[r]:<11s}{100 * self.classification_rate(r):g}% ({self.assigned(r)})")
        return out


# --------------------------------------------------------------------------------------
# per-taxon read counts and abundance estimation (-abundances / -abundance-per <rank>)
# --------------------------------------------------------------------------------------
NUM_RANKS = RANK_NONE            # taxonomy.hpp:103: a ranked lineage has one entry per rank below `none`
RANK_SEQUENCE = 0


class TaxonCounts:
    """`taxon_count_map` (classification.hpp:48-56): taxon -> number of reads (double), iterated in
    `rank_higher` order = rank descending (root first), then taxon id ascending.  Taxa are named by their
    ordinal + 1 into `DbMeta.taxa` (what the device classifier returns); `count_batch` is
    `++taxCounts[cls.best]` (classification.cpp:552-554) for a whole batch."""

    def __init__(self, taxa):
        self.taxa = taxa
        self.counts = {}                                            # ordinal + 1 -> float
        self._by_id = None
        self._lineages = {}

    def count_batch(self, classifications: np.ndarray):
        c = np.asarray(classifications).reshape(-1, 2)
        hit = c[:, 0][c[:, 0] != 0].astype(np.int64)
        for o, n in zip(*np.unique(hit, return_counts=True)):
            self.counts[int(o)] = self.counts.get(int(o), 0.0) + float(n)

    def merge(self, other: "TaxonCounts"):
        """publish_results (classification.cpp:569-578): a worker's counts into the global map"""
        for o, n in other.counts.items():
            self.counts[o] = self.counts.get(o, 0.0) + n

    # -- order and lineages -------------------------------------------------------------
    def _key(self, o):
        t = self.taxa[o - 1]
        return (-t.rank, t.id)

    def ordered(self):
        return sorted(self.counts, key=self._key)

    def ranked_lineage(self, o):
        """`taxonomy::make_ranks` (taxonomy.hpp:576-597): the taxon itself and its ancestors by rank"""
        lin = self._lineages.get(o)
        if lin is not None:
            return lin
        if self._by_id is None:
            self._by_id = {t.id: i + 1 for i, t in enumerate(self.taxa)}
        lin = [0] * NUM_RANKS
        t = self.taxa[o - 1]
        if t.rank != RANK_NONE:
            lin[t.rank] = o
        pid = t.parent
        while pid != 0:
            a = self._by_id.get(pid)
            if a is None:
                break
            ta = self.taxa[a - 1]
            if ta.rank != RANK_NONE:
                lin[ta.rank] = a
            if ta.parent == pid:                                    # cycles end the walk
                break
            pid = ta.parent
        self._lineages[o] = lin
        return lin

    # -- estimate_abundance (classification.cpp:304-377) ----------------------------------
    def estimate_abundance(self, rank: int):
        """Counts of taxa below `rank` move up to their ancestor on (or above) that rank; counts of taxa
        above it are distributed over their counted descendants in proportion to the descendants' weights;
        what remains are the leaves.  Arithmetic as in the reference: counts are doubles, weights integers."""
        counts = self.counts
        if rank != RANK_SEQUENCE:
            # everything from lower_bound(taxon{id 0, rank - 1}) on: lower ranks, and rank - 1 with id >= 0
            below = [o for o in self.ordered()
                     if self.taxa[o - 1].rank < rank - 1 or (self.taxa[o - 1].rank == rank - 1 and self.taxa[o - 1].id >= 0)]
            for o in below:
                lin = self.ranked_lineage(o)
                anc = next((lin[i] for i in range(rank, NUM_RANKS) if lin[i]), 0)
                if anc:
                    counts[anc] = counts.get(anc, 0.0) + counts[o]
                    del counts[o]
        order = self.ordered()
        weights = {o: 0 for o in order}
        children = {}
        for o in reversed(order):                                   # leaves to root
            lin = self.ranked_lineage(o)
            r = self.taxa[o - 1].rank
            for i in range(min(r + 1, 256), NUM_RANKS):
                p = lin[i]
                if p and p in weights:
                    weights[p] = int(weights[p] + (weights[o] + counts[o]))
                    children.setdefault(p, []).append(o)
                    break
        for o in order:                                             # root to leaves
            ch = children.get(o)
            if ch:
                total = weights[o]
                for c in ch:
                    counts[c] = counts[c] + counts[o] * (counts[c] + weights[c]) / total
                del counts[o]

    # -- show_abundance_table (printing.cpp:424-468) ---------------------------------------
    def table_lines(self, statistics: ClassificationStatistics, title: str, prefix: str = "# ", column: str = "\t|\t"):
        out = [prefix + title, f"{prefix}rank:name{column}taxid{column}number of reads{column}abundance"]
        total = statistics.total()
        for o in self.ordered():
            t = self.taxa[o - 1]
            n = self.counts[o]
            taxid = t.parent if t.rank == RANK_SEQUENCE else t.id
            # whole numbers print through the default stream format, fractions with 15 significant digits
            num = f"{n:g}" if float(n).is_integer() else f"{n:.15g}"
            out.append(f"{RANK_NAMES[t.rank]}:{t.name}{column}{taxid}{column}{num}{column}{n / float(total) * 100:g}%")
        out.append(f"unclassified{column}--{column}0{column}{statistics.unassigned()}{column}"
                   f"{statistics.unclassified_rate() * 100:g}%")
        return out

    def abundance_lines(self, statistics, **kw):
        """show_abundances (printing.cpp:473-481)"""
        return self.table_lines(statistics, "query summary: number of queries mapped per taxon", **kw)

    def estimate_lines(self, statistics, rank: int, **kw):
        """show_abundance_estimates (printing.cpp:486-497), after estimate_abundance(rank)"""
        return self.table_lines(statistics, f"estimated abundance (number of queries) per {RANK_NAMES[rank]}", **kw)
