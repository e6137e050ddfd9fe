"""Per-rank classification statistics of a query run (row N3 of SURVEY.md 8f): the counters of
`classification_statistics::assign` (classification_statistics.hpp:41-200) and the summary block of
`show_taxon_statistics` (printing.cpp:502-540), fed with the (taxon, rank) pairs the device classifier
returns (`QueryHostData.classifications()`), a whole batch per call."""
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
