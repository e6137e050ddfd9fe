"""Per-read output columns as the reference prints them (used to diff against the
reference's own golden file test/data/classified.expected).

  show_matches     printing.cpp:315-340   all_hits  = "name/win:count," per distinct location
  show_candidates  printing.cpp:283-295   top_hits  = "name:hits" comma separated
Only the `-lowest sequence` (default) variants are mirrored.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def format_all_hits(allhits: np.ndarray, target_names: Sequence[str]) -> str:
    if len(allhits) == 0:
        return ""
    uniq, counts = np.unique(allhits, return_counts=True)   # sorted u64 -> (tgt, win) order kept
    out = []
    for key, c in zip(uniq.tolist(), counts.tolist()):
        out.append(f"{target_names[key >> 32]}/{key & 0xFFFFFFFF}:{c},")
    return "".join(out)


def format_top_hits(top, target_names: Sequence[str]) -> str:
    """top: iterable of (tgt, hits, beg, end); stops at the first entry with hits == 0."""
    out = []
    for tgt, hits, _beg, _end in top:
        if hits == 0:
            break
        out.append(f"{target_names[tgt]}:{hits}")
    return ",".join(out)


RANK_NAMES = ["sequence", "form", "variety", "subspecies", "species", "subgenus", "genus", "subtribe", "tribe",
              "subfamily", "family", "suborder", "order", "subclass", "class", "subphylum", "phylum",
              "subkingdom", "kingdom", "domain", "root", "none"]          # taxonomy.hpp:226-252


def format_classification(taxon_ordinal: int, rank: int, taxa) -> str:
    """show_taxon (printing.cpp:250-280) in its default style: `rank:name`, `--` if unclassified.
    taxon_ordinal = index + 1 into `taxa` (DbMeta.taxa), 0 = unclassified."""
    if taxon_ordinal == 0:
        return "--"
    return f"{RANK_NAMES[rank]}:{taxa[taxon_ordinal - 1].name}"
