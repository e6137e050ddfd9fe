"""The synthetic workload generators (BASELINE.md section 3 recipes): the torch generator that fills
HBM for the benchmark and its numpy twin used by the tests give the same bytes."""
import numpy as np
import torch

from metacache_b200 import synth

NT, TL = 20, 30000


def test_targets_and_r150_torch_equals_numpy():
    tb, off = synth.make_targets(NT, TL, 10, synth.SEED_DB)
    tt, toff = synth.make_targets(NT, TL, 10, synth.SEED_DB, device=torch.device("cpu"))
    assert np.array_equal(tb, tt.numpy()) and np.array_equal(off.astype(np.int64), toff.numpy())
    a = synth.make_reads_150(500, tb, NT, TL)
    b = synth.make_reads_150(500, tt, NT, TL, device=torch.device("cpu"))
    assert a.shape == (500, 150) and np.array_equal(a, b.numpy())
    assert set(np.unique(a)) <= set(b"ACGTN")
    # family members differ by about 1 %
    d = (tb[:TL] != tb[TL:2 * TL]).mean()
    assert 0.005 < d < 0.04


def test_long_reads_torch_equals_numpy_and_follow_the_length_recipe():
    tb, _ = synth.make_targets(NT, TL, 10, synth.SEED_DB)
    a, ao = synth.make_long_reads(300, tb, NT, TL)
    b, bo = synth.make_long_reads(300, torch.from_numpy(tb), NT, TL, device=torch.device("cpu"))
    assert np.array_equal(a, b.numpy()) and np.array_equal(ao, bo.numpy())
    lens = np.diff(ao)
    assert np.array_equal(lens, np.minimum(synth.long_read_lengths(300), 19000))
    assert lens.min() >= 200 and lens.max() <= 19000 and 350 < np.median(lens) < 650
    # a read is a mutated substring of a target or of its reverse complement (~5 % substitutions)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    r = bytes(a[ao[0]:ao[1]])
    tgt = tb.tobytes()

    def best(seq):
        # anchor on the first exactly matching 24-mer
        for p in range(0, len(seq) - 24):
            i = tgt.find(seq[p:p + 24])
            if i >= 0:
                s = i - p
                if s < 0 or s + len(seq) > len(tgt):
                    continue
                ref = tgt[s:s + len(seq)]
                return sum(x != y for x, y in zip(seq, ref)) / len(seq)
        return 1.0
    assert min(best(r), best(r.translate(comp)[::-1])) < 0.12
