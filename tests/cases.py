"""The parity cases: must stay in step with oracle/make_golden.py:CASES."""
CASES = [
    ("cfg1_pi4_2p20", dict(seed=0, alphabet="pi4", snr_db=30.0), 1 << 20, 2.4e6, 0.0),
    ("cfg1_pi4_2p20_fo", dict(seed=0, alphabet="pi4", snr_db=30.0), 1 << 20, 2.4e6, 1234.5),
    ("centred_2p18", dict(seed=1, alphabet="centred", snr_db=30.0), 1 << 18, 2.4e6, 0.0),
    ("short_24000", dict(seed=2, alphabet="centred", snr_db=25.0), 24000, 2.4e6, 0.0),
    ("gui_131072", dict(seed=3, alphabet="centred", snr_db=15.0), 131072, 2.4e6, 0.0),
    ("gui_131072_fo", dict(seed=3, alphabet="pi4", snr_db=20.0), 131072, 2.4e6, -5000.0),
    ("odd_100003", dict(seed=4, alphabet="pi4", snr_db=20.0), 100003, 2.4e6, 0.0),
    ("min_fast_16384", dict(seed=5, alphabet="centred", snr_db=30.0), 16384, 2.4e6, 0.0),
    ("rate_1p8M", dict(seed=6, alphabet="centred", snr_db=30.0, sps=98), 65536, 1.8e6, 0.0),
    ("rate_2p048M", dict(seed=7, alphabet="centred", snr_db=30.0, sps=112), 65536, 2.048e6, 250.0),
    ("rate_1M", dict(seed=8, alphabet="centred", snr_db=30.0, sps=52), 50000, 1.0e6, 0.0),
    ("rate_240k", dict(seed=9, alphabet="centred", snr_db=30.0, sps=13), 20000, 240e3, 0.0),
    ("tiny_100", dict(seed=10, alphabet="centred", snr_db=30.0), 100, 2.4e6, 0.0),
    ("tiny_20", dict(seed=11, alphabet="centred", snr_db=30.0), 20, 2.4e6, 0.0),
    ("tiny_300", dict(seed=12, alphabet="centred", snr_db=30.0), 300, 2.4e6, 0.0),
]
SYNC_THRESHOLDS = (0.90, 0.85, 0.80, 0.78)
SMALL = [c for c in CASES if c[2] <= 131072]
