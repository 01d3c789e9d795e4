"""CPU: the float64 model of the block-end corrections (tools/edge_model.py) against SciPy's own edge handling, and the
generated weight tables against the model that wrote them."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import edge_model as em  # noqa: E402


@pytest.fixture(scope="module")
def tables():
    return em.build_tables()


@pytest.mark.parametrize("n,fo", [(16384, 0.0), (20003, 1234.5), (20007, -12500.0), (131072, 0.0), (24000, 7000.0)])
def test_model_reproduces_scipy_block_ends(tables, n, fo):
    """reference(x) - cascade(x zero-extended), both in float64 with the exact IIRs, equals the staged correction at
    both ends (every residue of n mod 10 changes the decimation phase of the right end)."""
    err_l, err_r, interior, size_l, size_r = em.check(n, fo, seed=n, tab=tables)
    assert err_l < 2e-9 and err_r < 2e-9, (err_l, err_r)
    assert interior < 1e-9                      # away from the ends the chain is shift-invariant
    assert size_l > 0.01 and size_r > 0.01      # the corrections are not small: the ends matter


def test_generated_header_matches_model(tables):
    path = os.path.join(ROOT, "tetraear_b200", "csrc", "edge_tables_generated.h")
    txt = open(path).read()
    defs = dict(re.findall(r"#define (ET_\w+) (\d+)", txt))
    assert int(defs["ET_G"]) == em.G_HALF and int(defs["ET_NC"]) == em.NC and int(defs["ET_NAC"]) == em.NAC
    assert int(defs["ET_TD"]) == em.TD and int(defs["ET_NRING"]) == em.NRING and int(defs["ET_T2"]) == em.T2
    for name, key in (("ET_G1", "g1"), ("ET_WC", "wc"), ("ET_WAC", "wac"), ("ET_RINGC", "ringc"), ("ET_RING", "ring"), ("ET_U", "u")):
        m = re.search(r"static const double %s\[(\d+)\] = \{(.*?)\};" % name, txt, re.S)
        vals = np.array([float(v) for v in m.group(2).replace("\n", " ").split(",") if v.strip()])
        ref = np.asarray(tables[key]).reshape(-1)
        assert int(m.group(1)) == len(ref) == len(vals)
        assert np.abs(vals - ref).max() <= 1e-15 * max(1.0, np.abs(ref).max())
