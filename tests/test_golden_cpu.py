"""Golden fixtures on REAL data (tests/golden/, made by tests/golden/make_golden.py from the
reference's own sample FITS: HST WFPC2 M16 at 502 / 656 / 673 nm, the BASELINE config-1 fixtures).
CPU side: the oracle must keep reproducing the committed vectors bit for bit (guards the checker
against drift) and the raw big-endian data units must decode to the same pixels numpy reads."""
import os
import sys

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)


def test_fixture_files_and_decode(orc):
    import make_golden as mg

    crops = mg.load_crops()
    for name, c in zip(mg.CHANNELS, crops):
        raw = open(os.path.join(GOLD, f"m16_{name}_crop{mg.N}.bef32"), "rb").read()
        assert len(raw) == mg.N * mg.N * 4
        assert np.array_equal(c, np.frombuffer(raw, ">f4").astype(np.float32).reshape(mg.N, mg.N))
        assert (c == 0).sum() > 1000 and c.max() > 100.0      # mosaic padding and real signal are in the crop


def test_oracle_reproduces_golden_vectors(orc):
    import make_golden as mg

    gold = np.load(os.path.join(GOLD, "m16_golden.npz"))
    got = mg.compute(mg.load_crops())
    assert sorted(got) == sorted(gold.files)
    for k in gold.files:
        assert np.array_equal(got[k], gold[k]), k
    assert gold["stars_656nmos"].shape[0] >= 10 and gold["rgb8"].shape == (mg.N, mg.N, 3)
