"""Debug: where do the config-3 / u8+offset parity failures sit? (run on the GPU box)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tetraear_b200 import synth
from tetraear_b200.processor import SignalProcessor
from oracle import ref_dsp

sp = SignalProcessor(2.4e6)
n = 1 << 17
x, active, freqs = synth.wideband_capture(n, seed=3)
pick = [0, 1, 17, 47, 48, 49, 80, 95]
res = sp.process_wideband(x, freqs[pick], want_symbols=True, want_match=True)
x128 = x.astype(np.complex128)
for row, k in enumerate(pick):
    r = ref_dsp.process(ref_dsp.nco(x128, freqs[k], 2.4e6), 0.0, 2.4e6)
    nd = int(res["n_dibits"][row])
    m = min(nd + 1, len(r["symbols"]))
    e = np.abs(res["symbols"][row, :m] - r["symbols"][:m]) / np.abs(r["symbols"]).max()
    bad = np.flatnonzero(e > 1e-5)
    print("c3 k=%d nd=%d/%d best=%d/%d maxerr=%.3g nbad=%d first/last bad=%s" % (k, nd, len(r["dibits"]), int(res["best_phase"][row]), r["best_phase"],
          e.max(), len(bad), (bad[:3].tolist(), bad[-3:].tolist()) if len(bad) else None))
sys.stdout.flush()
n, n_car = 16384 + 640, 300
base = []
for k in range(4):
    xx = synth.carrier_iq(n, 540 + k, snr_db=24.0, alphabet="centred" if k & 1 else "pi4")
    z = xx / np.abs(xx).max() * 0.9
    base.append(np.stack([np.clip(np.round((z.real + 1.0) * 127.5), 0, 255), np.clip(np.round((z.imag + 1.0) * 127.5), 0, 255)], axis=-1).astype(np.uint8))
raw = np.stack([base[c % 4] for c in range(n_car)])
fos = np.random.default_rng(8).uniform(-12000.0, 12000.0, size=n_car)
fos[::7] = 0.0
res = sp.process_batch_u8(raw, fos, want_symbols=True)
nbadc = 0
for c in range(n_car):
    b = base[c % 4]
    x128 = (b[:, 0].astype(np.float64) / 127.5 - 1.0) + 1j * (b[:, 1].astype(np.float64) / 127.5 - 1.0)
    r = ref_dsp.process(x128, float(fos[c]), 2.4e6)
    nd = int(res["n_dibits"][c])
    m = min(nd + 1, len(r["symbols"]))
    e = np.abs(res["symbols"][c, :m] - r["symbols"][:m]) / np.abs(r["symbols"]).max()
    if nd != len(r["dibits"]) or e.max() > 1e-5:
        nbadc += 1
        if nbadc < 12:
            bad = np.flatnonzero(e > 1e-5)
            print("u8fo c=%d fo=%.1f nd=%d/%d best=%d/%d maxerr=%.3g nbad=%d/%d" % (c, fos[c], nd, len(r["dibits"]), int(res["best_phase"][c]), r["best_phase"], e.max(), len(bad), m))
print("u8fo bad carriers:", nbadc, "of", n_car)
