"""GPU: the block-end correction kernel (k_edge_correct, through tetra_edge_corrections) against its float64 model, which
tests/test_edge_model.py pins to SciPy's sosfiltfilt / filtfilt edge handling."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import edge_model as em  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tables():
    return em.build_tables()


@pytest.mark.parametrize("n", [16384, 20003, 131072, 100007])
def test_edge_corrections_match_model(gpu_processor, tables, n):
    sp = gpu_processor
    rng = np.random.default_rng(n)
    fos = np.array([0.0, 1234.5, -12500.0, 6000.25])
    x = (rng.standard_normal((4, n)) + 1j * rng.standard_normal((4, n))).astype(np.complex64)
    x[:, :60] += 2.0                                   # block ends far from zero: large corrections
    x[:, -60:] -= 1.5j
    out = np.zeros((4, 2, 168), dtype=np.complex64)
    sp._check(sp._lib.tetra_edge_corrections(sp._ctx, x.ctypes.data, 4, n, n, fos.ctypes.data, out.ctypes.data), "edge_corrections")
    for c in range(4):
        dl, dr = em.edge_corrections(x[c].astype(np.complex128), float(fos[c]), tables)
        scale = max(np.abs(dl).max(), np.abs(dr).max())
        assert scale > 0.05
        assert np.abs(out[c, 0] - dl).max() <= 3e-7 * scale, (c, np.abs(out[c, 0] - dl).max() / scale)     # float32 output
        assert np.abs(out[c, 1] - dr).max() <= 3e-7 * scale, (c, np.abs(out[c, 1] - dr).max() / scale)


def test_block_ends_of_process_follow_the_reference(gpu_processor):
    """The first and last symbols of a block -- the ones the corrections touch -- against the oracle, on an input whose
    ends are far from zero (a strong unmodulated carrier: the odd extension and zi * x0 matter at full scale)."""
    from oracle import ref_dsp
    from tetraear_b200 import synth
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    n = 1 << 17
    x = synth.carrier_iq(n, 321, snr_db=25.0, alphabet="centred").astype(np.complex64) + np.complex64(0.8 - 0.3j)
    for fo in (0.0, 2500.0):
        res = sp.process_batch(x[None, :], [fo], want_symbols=True)
        r = ref_dsp.process(x.astype(np.complex128), fo, 2.4e6)
        nd = int(res["n_dibits"][0])
        assert nd == len(r["dibits"]) and int(res["best_phase"][0]) == r["best_phase"]
        assert np.array_equal(res["dibits"][0, :nd], r["dibits"])
        s = res["symbols"][0, : nd + 1]
        scale = np.abs(r["symbols"]).max()
        assert np.abs(s[:16] - r["symbols"][:16]).max() / scale <= 1e-5
        assert np.abs(s[-16:] - r["symbols"][-16:]).max() / scale <= 1e-5
        assert np.abs(s - r["symbols"]).max() / scale <= 1e-5
