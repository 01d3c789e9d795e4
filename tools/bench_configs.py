#!/usr/bin/env python
"""Timings of the BASELINE configs that are not the bench line (configs[1], [2], [4]): single-carrier latency,
96-channel wideband capture, waterfall STFT. CUDA-event timed through the C ABI with device-resident buffers
where the entry point accepts them. Writes one JSON object (stdout) for profiles/.

    python tools/bench_configs.py > gpurun_out/configs.json
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(only=None, device=0):
    """Measure the selected configs ("2", "3", "u8", "5"; None = TETRA_CONFIGS or all) on `device`; returns the dict."""
    import torch
    from tetraear_b200 import synth
    from tetraear_b200.processor import SignalProcessor
    out = {}
    dev = torch.device("cuda", device)
    sp = SignalProcessor(2.4e6, device=device)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    def timed(fn, reps=20, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps

    if only is None:
        only = [t for t in os.environ.get("TETRA_CONFIGS", "").split(",") if t]      # e.g. TETRA_CONFIGS=3 under ncu

    def want(tag):
        return not only or tag in only

    # ---- config 2: one carrier, 2^20 samples, device-resident (latency-bound) ----
    n = 1 << 20
    cap = sp.dibit_capacity(n)
    x = torch.view_as_real(torch.from_numpy(synth.carrier_iq(n, 0, snr_db=30.0)).to(dev)).contiguous()
    dib = torch.zeros((1, cap), dtype=torch.uint8, device=dev)
    nd = torch.zeros(1, dtype=torch.int32, device=dev)
    sym = torch.zeros((1, cap + 1, 2), dtype=torch.float32, device=dev)
    mt = torch.zeros((1, 2 * cap, 2), dtype=torch.uint8, device=dev)
    ph = torch.zeros(1, dtype=torch.int32, device=dev)
    for name, fo in (("config2_one_carrier_fo0", None), ("config2_one_carrier_fo1234.5", [1234.5])) if want("2") else ():
        ms, wall = timed(lambda: sp.process_batch_device(x.data_ptr(), 1, n, n, dib.data_ptr(), cap, nd.data_ptr(), sym.data_ptr(),
                                                         ph.data_ptr(), mt.data_ptr(), stream=0, freq_offsets=fo))
        out[name] = {"ms_per_block": ms, "host_wall_ms": wall, "MS_per_s": n / ms / 1e3, "x_real_time": (n / 2.4e6) / (ms * 1e-3)}
    if want("2"):
        # the class-level call a TetraEar user makes (host numpy in, host numpy out, H2D + D2H + Python inside), at the block
        # of config 2 and at the GUI's chunk (ui/modern.py:1912 reads 128*1024 samples), for the three input forms
        def host_call_ms(fn, reps=20):
            fn(); fn()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            return (time.perf_counter() - t0) * 1e3 / reps

        calls = {}
        for label, ns in (("2^20", n), ("gui_chunk_131072", 128 * 1024)):
            hx = synth.carrier_iq(ns, 0, snr_db=30.0)
            hx128 = hx.astype(np.complex128)
            z = hx / np.abs(hx).max() * 0.9
            raw = np.stack([np.clip(np.round((z.real + 1.0) * 127.5), 0, 255), np.clip(np.round((z.imag + 1.0) * 127.5), 0, 255)],
                           axis=-1).astype(np.uint8)[None]
            hp = torch.from_numpy(hx).pin_memory().numpy()
            calls[label] = {
                "process_complex64_ms": host_call_ms(lambda: sp.process(hx)),
                "process_complex64_afc_ms": host_call_ms(lambda: sp.process(hx, freq_offset=1234.5)),
                "process_complex64_pinned_ms": host_call_ms(lambda: sp.process(hp)),
                "process_complex128_ms": host_call_ms(lambda: sp.process(hx128)),
                "process_batch_u8_afc_ms": host_call_ms(lambda: sp.process_batch_u8(raw, [1234.5])),
                "block_duration_ms": ns / 2.4e3,
            }
        out["config2_host_calls"] = calls
        # the other RTL-SDR rates (signal/capture.py:83-87) take the exact-recursion kernel (KX), one thread per carrier
        other = {}
        for rate in (2.048e6, 1.8e6):
            sp.sample_rate = rate
            for label, ns in (("2^20", n), ("gui_chunk_131072", 128 * 1024)):
                hx = synth.carrier_iq(ns, 0, snr_db=30.0)
                other["%.3f MS/s %s" % (rate / 1e6, label)] = {"process_complex64_ms": host_call_ms(lambda: sp.process(hx), reps=3),
                                                               "block_duration_ms": ns / rate * 1e3}
        sp.sample_rate = 2.4e6
        out["config2_other_rates_exact_path"] = other
        out["config2_process_host_call_ms"] = calls["2^20"]["process_complex64_ms"]

    if want("3"):
        # ---- config 3: 96 channels of one 2^20-sample wideband capture ----
        xw, active, freqs = synth.wideband_capture(n, seed=3)
        t0 = time.perf_counter()
        reps = 5
        sp.process_wideband(xw, freqs, want_symbols=True, want_match=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            sp.process_wideband(xw, freqs, want_symbols=True, want_match=True)
        wall = (time.perf_counter() - t0) / reps
        out["config3_wideband_96ch"] = {"host_call_ms": wall * 1e3, "wideband_MS_per_s": n / wall / 1e6, "channel_MS_per_s": 96 * n / wall / 1e6,
                                        "x_real_time": (n / 2.4e6) / wall, "note": "host capture in, host dibits/symbols/match out (PCIe + 96 result rows inside)"}

        # device-resident variant: capture and every output already in HBM
        xw_d = torch.view_as_real(torch.from_numpy(xw).to(dev)).contiguous()
        fr = np.ascontiguousarray(freqs, dtype=np.float64)
        dib3 = torch.zeros((96, cap), dtype=torch.uint8, device=dev)
        nd3 = torch.zeros(96, dtype=torch.int32, device=dev)
        sym3 = torch.zeros((96, cap + 1, 2), dtype=torch.float32, device=dev)
        ph3 = torch.zeros(96, dtype=torch.int32, device=dev)
        mt3 = torch.zeros((96, 2 * cap, 2), dtype=torch.uint8, device=dev)
        sp._lib.tetra_set_stream(sp._ctx, 1)
        ms, wall = timed(lambda: sp._lib.tetra_process_wideband(sp._ctx, xw_d.data_ptr(), n, fr.ctypes.data, 96, dib3.data_ptr(), cap,
                                                                nd3.data_ptr(), sym3.data_ptr(), ph3.data_ptr(), mt3.data_ptr()), reps=10)
        sp._lib.tetra_set_stream(sp._ctx, None)
        # fp32 work of the per-channel stages (FFMA2 per input sample: proto 4.1 [x2 with modulated complex taps],
        # half-band 0.65, fir120 6.4, interpolation 0.8) against the measured FFMA2 issue peak (tools/microbench/pipes.cu:
        # 0.5 warp instructions per clock and SM sub-partition)
        pfb_on = os.environ.get("TETRA_PFB", "1") != "0"
        ffma2_per_sample = (0.65 + 6.4 + 0.8) if pfb_on else (2 * 4.1 + 0.65 + 6.4 + 0.8)
        ffma2_peak = 148 * 4 * 0.5 * 32 * 1.965e9
        out["config3_wideband_96ch_device_resident"] = {"ms_per_capture": ms, "wideband_MS_per_s": n / ms / 1e3,
                                                        "channel_MS_per_s": 96 * n / ms / 1e3, "x_real_time": (n / 2.4e6) / (ms * 1e-3),
                                                        "front_end": "k_pfb96 polyphase DFT + MODE 5" if pfb_on else "per-channel modulated proto (MODE 2)",
                                                        "ffma2_lane_ops_per_capture": 96 * n * ffma2_per_sample,
                                                        "fp32_pipe_utilisation": 96 * n * ffma2_per_sample / (ms * 1e-3) / ffma2_peak,
                                                        }

    if want("u8"):
        # ---- RTL-SDR bytes, device-resident, the configs[3] batch (4096 carriers x 2^20 samples at 2 B/sample) ----
        cu = int(os.environ.get("TETRA_U8_CARRIERS", "4096"))
        g = torch.Generator(device=dev); g.manual_seed(7)
        raw = torch.randint(0, 256, (cu, n, 2), dtype=torch.uint8, device=dev, generator=g)     # timing only: any bytes will do
        dib8 = torch.zeros((cu, cap), dtype=torch.uint8, device=dev)
        nd8 = torch.zeros(cu, dtype=torch.int32, device=dev)
        sym8 = torch.zeros((cu, cap + 1, 2), dtype=torch.float32, device=dev)
        ph8 = torch.zeros(cu, dtype=torch.int32, device=dev)
        mt8 = torch.zeros((cu, 2 * cap, 2), dtype=torch.uint8, device=dev)
        sp._lib.tetra_set_stream(sp._ctx, 1)
        sp._lib.tetra_enable_kernel_timing(sp._ctx, 1)
        ms, wall = timed(lambda: sp._lib.tetra_process_batch_u8(sp._ctx, raw.data_ptr(), cu, n, n, None, dib8.data_ptr(), cap, nd8.data_ptr(),
                                                                sym8.data_ptr(), ph8.data_ptr(), mt8.data_ptr(), None, 0, None), reps=5)
        k1_ms = sp.kernel_time_ms()
        sp._lib.tetra_enable_kernel_timing(sp._ctx, 0)
        sp._lib.tetra_set_stream(sp._ctx, None)
        out["u8_ingest_device_resident"] = {"carriers": cu, "ms_per_batch": ms, "MS_per_s": cu * n / ms / 1e3, "fused_kernel_ms": k1_ms[0] / max(k1_ms[1], 1),
                                            "GB_per_s_at_2.1B_per_sample": 2.1 * cu * n / (ms * 1e-3) / 1e9,
                                            "note": "2 B/sample read + the outputs' 0.1 B/sample; stage A of the fused kernel filters the bytes as they are"}
        del raw, dib8, sym8, mt8

    if want("5"):
        # ---- config 5: waterfall STFT 4096 / hop 1024 on 1 s of IQ, device-resident ----
        ns = 2_400_000
        xs = torch.view_as_real(torch.from_numpy(synth.stft_test_signal(ns, 5)).to(dev)).contiguous()
        rows = (ns - 4096) // 1024 + 1
        o = torch.zeros((rows, 4096), dtype=torch.float32, device=dev)
        import ctypes as C
        r64 = C.c_int64(0)
        ms, wall = timed(lambda: sp._lib.tetra_stft_db(sp._ctx, xs.data_ptr(), ns, 4096, 1024, o.data_ptr(), C.byref(r64)), reps=50)
        by = 24.0 * ns
        out["config5_stft_4096_hop1024"] = {"ms_per_second_of_iq": ms, "rows_per_s": rows / (ms * 1e-3), "MS_per_s": ns / ms / 1e3,
                                            "x_real_time": 1e3 / ms, "GB_per_s_at_24B_per_sample": by / (ms * 1e-3) / 1e9,
                                            "frac_of_measured_hbm_peak": by / (ms * 1e-3) / 1e9 / peak, "frames_at_60fps_rows": rows / 60.0}
        # the same on 30 s of IQ in one call: the per-call launch + synchronise cost (tens of microseconds) no longer shows
        reps30 = 30
        xl = xs.repeat(reps30, 1)
        nl = ns * reps30
        rows_l = (nl - 4096) // 1024 + 1
        ol = torch.zeros((rows_l, 4096), dtype=torch.float32, device=dev)
        ms, wall = timed(lambda: sp._lib.tetra_stft_db(sp._ctx, xl.data_ptr(), nl, 4096, 1024, ol.data_ptr(), C.byref(r64)), reps=10)
        by = 24.0 * nl
        out["config5_stft_4096_hop1024_30s"] = {"ms_per_call": ms, "rows_per_s": rows_l / (ms * 1e-3), "MS_per_s": nl / ms / 1e3,
                                                "x_real_time": 30e3 / ms, "GB_per_s_at_24B_per_sample": by / (ms * 1e-3) / 1e9,
                                                "frac_of_measured_hbm_peak": by / (ms * 1e-3) / 1e9 / peak}
    sp.close()
    return out


if __name__ == "__main__":
    print(json.dumps(run(), indent=1))
