"""The committed golden vectors ARE what the unmodified reference produces: regenerates every file of ``tests/golden``
by running ``oracle/make_golden.py`` (which executes ``/root/reference/src/scripts`` through ``oracle/ref_harness.py``)
into a scratch directory and compares it with the committed copy.  Dev container only: skipped where
``/root/reference`` does not exist (the GPU box); test infrastructure, nothing here is imported by the product."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

REF = "/root/reference/src/scripts"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not present on this machine")


@pytest.fixture(scope="module")
def regenerated(tmp_path_factory):
    import make_golden as mg
    out = tmp_path_factory.mktemp("golden")
    old = mg.OUT
    mg.OUT = str(out)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mg.main()
    finally:
        mg.OUT = old
    return str(out)


def _same(x, y):
    if x.shape != y.shape:
        return False
    if x.dtype.kind in "iubUSO":
        return np.array_equal(x, y)
    return np.allclose(x, y, rtol=1e-6, atol=1e-9, equal_nan=True)


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLD) if f.endswith(".npz")))
def test_committed_golden_file_is_reproduced_by_the_reference(regenerated, name):
    new = np.load(os.path.join(regenerated, name), allow_pickle=True)
    old = np.load(os.path.join(GOLD, name), allow_pickle=True)
    assert sorted(new.files) == sorted(old.files)
    for key in new.files:
        if key == "cases_json":
            continue
        assert _same(new[key], old[key]), f"{name}:{key}"


def test_hybrid_cases_match_up_to_exact_ties(regenerated):
    """The hybrid cases store the reference's final ids and scores: the scores reproduce exactly; two ids may swap where
    the reference's blended scores tie exactly (its sort is stable on a dict whose order follows BLAS rounding)."""
    new = json.loads(str(np.load(os.path.join(regenerated, "hybrid_golden.npz"), allow_pickle=True)["cases_json"]))
    old = json.loads(str(np.load(os.path.join(GOLD, "hybrid_golden.npz"), allow_pickle=True)["cases_json"]))
    assert len(new) == len(old)
    for a, b in zip(new, old):
        assert a["n_valid"] == b["n_valid"] and a["liked"] == b["liked"] and a["kwargs"] == b["kwargs"]
        assert [x is None for x in a["scores"]] == [x is None for x in b["scores"]]      # (short lists are padded)
        np.testing.assert_allclose([x for x in a["scores"] if x is not None], [x for x in b["scores"] if x is not None],
                                   rtol=1e-6, atol=1e-9)
        assert sorted(a["ids"]) == sorted(b["ids"])
        for p, (x, y) in enumerate(zip(a["ids"], b["ids"])):
            if x != y:                                   # a swap is legal only inside a run of (near-)equal scores
                q = b["ids"].index(x)
                assert abs(b["scores"][p] - b["scores"][q]) <= 1e-6 * max(1.0, abs(b["scores"][p]))
