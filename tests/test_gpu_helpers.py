"""GPU: the public helper methods of SignalProcessor (processor.py:35-219) through the C ABI against goldens written by
the reference itself (tests/golden/helpers.npz, oracle/make_golden.py:110-126), plus the assertions of the reference's
own unit tests (tests/unit/test_signal_processor.py:14-116) restated against the drop-in class."""
import numpy as np
import pytest

from conftest import load_golden
from oracle.make_golden import helper_signal, input_digest

pytestmark = pytest.mark.gpu
TOL = 1e-9           # complex128 in, complex128 out: relative to the largest output magnitude


@pytest.fixture(scope="module")
def helpers():
    g = load_golden("helpers")
    x = helper_signal()
    assert input_digest(x) == str(g["input_sha256"]), "helper_signal drifted from the golden fixture"
    return g, x


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max()


@pytest.mark.parametrize("key,kw,tol", [("filter_25k", dict(bandwidth=25000), 1e-8), ("filter_50k", dict(bandwidth=50000), TOL),
                                        ("filter_240k", dict(bandwidth=25000, sample_rate=240000.0), TOL)])
def test_filter_signal(gpu_processor, helpers, key, kw, tol):
    """The literal float64 recursion (SciPy's lfilter order of operations). At fs = 2.4 MHz and 25 kHz bandwidth the
    Butterworth poles sit within 2e-2 of the unit circle (wn = 0.0104, numerator ~1e-8): the direct form amplifies the
    last-bit differences between the two coefficient designs to 1.3e-9 (measured on B200), hence 1e-8 for that case."""
    g, x = helpers
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    y = sp.filter_signal(x, **kw)
    assert y.dtype == np.complex128 and len(y) == len(x)
    assert _rel(y, g[key]) <= tol


@pytest.mark.parametrize("key,fo,kw", [("shift_1k", 1000, {}), ("shift_m7k_240k", -7777.7, dict(sample_rate=240000.0))])
def test_frequency_shift(gpu_processor, helpers, key, fo, kw):
    g, x = helpers
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    y = sp.frequency_shift(x, fo, **kw)
    assert y.dtype == np.complex128 and _rel(y, g[key]) <= TOL


@pytest.mark.parametrize("key,kw", [("extract_2p4M", {}), ("extract_1M", dict(sample_rate=1.0e6)),
                                    ("extract_240k", dict(sample_rate=240000.0))])
def test_extract_symbols(gpu_processor, helpers, key, kw):
    """sps = 133 (9 phases, step 16), 55 and 13: the gather is a copy, so the result is bit-identical."""
    g, x = helpers
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    y = sp.extract_symbols(x, **kw)
    assert np.iscomplexobj(y) and len(y) == len(g[key])
    assert np.array_equal(y, g[key])


def test_demodulate_dqpsk_on_raw_iq(gpu_processor, helpers):
    g, x = helpers
    d = gpu_processor.demodulate_dqpsk(x)
    assert d.dtype == np.uint8 and np.array_equal(d, g["demod"])


@pytest.mark.parametrize("key,sl,rate", [("resample_half", slice(None), 1.2e6), ("resample_up", slice(0, 3000), 3.6e6),
                                         ("resample_odd", slice(0, 3001), 1.0e6)])
def test_resample(gpu_processor, helpers, key, sl, rate):
    g, x = helpers
    sp = gpu_processor
    sp.sample_rate = 2.4e6
    y = sp.resample(x[sl], rate)
    assert y.dtype == np.complex128 and len(y) == len(g[key])
    assert _rel(y, g[key]) <= TOL


class TestReferenceUnitTests:
    """tests/unit/test_signal_processor.py of the reference, assertion by assertion, on the drop-in class."""

    @pytest.fixture
    def iq(self):
        # tests/conftest.py:53-67 (`sample_iq_samples`), seeded
        rng = np.random.default_rng(99)
        t = np.arange(0, 0.01, 1 / 2.4e6)
        return np.exp(1j * 2 * np.pi * 0 * t) + (rng.standard_normal(len(t)) + 1j * rng.standard_normal(len(t))) * 0.1

    def test_processor_initialization(self):
        from tetraear_b200.processor import SignalProcessor
        p = SignalProcessor()
        assert p.sample_rate == 2.4e6 and p.symbol_rate == 18000 and p.samples_per_symbol > 0
        p.close()

    def test_processor_custom_sample_rate(self):
        from tetraear_b200.processor import SignalProcessor
        p = SignalProcessor(sample_rate=1.0e6)
        assert p.sample_rate == 1.0e6 and p.symbol_rate == 18000
        p.close()

    def test_resample(self, gpu_processor, iq):
        gpu_processor.sample_rate = 2.4e6
        r = gpu_processor.resample(iq, 1.2e6)
        assert len(r) > 0 and isinstance(r, np.ndarray) and np.iscomplexobj(r)

    def test_filter_signal_empty(self, gpu_processor):
        assert len(gpu_processor.filter_signal(np.array([]))) == 0

    def test_filter_signal(self, gpu_processor, iq):
        gpu_processor.sample_rate = 2.4e6
        r = gpu_processor.filter_signal(iq, bandwidth=25000)
        assert len(r) == len(iq) and isinstance(r, np.ndarray)

    def test_filter_signal_custom_bandwidth(self, gpu_processor, iq):
        assert len(gpu_processor.filter_signal(iq, bandwidth=50000)) == len(iq)

    def test_frequency_shift(self, gpu_processor, iq):
        r = gpu_processor.frequency_shift(iq, 1000)
        assert len(r) == len(iq) and isinstance(r, np.ndarray) and np.iscomplexobj(r)

    def test_frequency_shift_zero(self, gpu_processor, iq):
        r = gpu_processor.frequency_shift(iq, 0)
        assert len(r) == len(iq) and np.allclose(r, iq)

    def test_demodulate_dqpsk_empty(self, gpu_processor):
        r = gpu_processor.demodulate_dqpsk(np.array([]))
        assert len(r) == 0 and isinstance(r, np.ndarray)

    def test_demodulate_dqpsk_single_sample(self, gpu_processor):
        assert len(gpu_processor.demodulate_dqpsk(np.array([1.0 + 1.0j]))) == 0

    def test_demodulate_dqpsk(self, gpu_processor, iq):
        r = gpu_processor.demodulate_dqpsk(iq)
        assert len(r) > 0 and isinstance(r, np.ndarray) and r.dtype == np.uint8 and all(0 <= s <= 3 for s in r)

    def test_extract_symbols_empty(self, gpu_processor):
        assert len(gpu_processor.extract_symbols(np.array([]))) == 0

    def test_extract_symbols(self, gpu_processor, iq):
        gpu_processor.sample_rate = 2.4e6
        r = gpu_processor.extract_symbols(iq)
        assert len(r) > 0 and isinstance(r, np.ndarray) and np.iscomplexobj(r)

    def test_extract_symbols_custom_rate(self, gpu_processor, iq):
        assert len(gpu_processor.extract_symbols(iq, sample_rate=1.0e6)) > 0
