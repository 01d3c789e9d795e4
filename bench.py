#!/usr/bin/env python
"""Benchmark of the IQ -> dibit hot path (BASELINE.json metric: IQ MS/s demodulated).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                      # the reference algorithm on the host cores (oracle port)

Workload (config.workload): BASELINE.json configs[3] -- 4096 independent 25 kHz carriers, 2^20 complex64
samples each at 2.4 MS/s (32 GiB, far larger than L2, resident in HBM), sharded over the ranks (strong
scaling: total fixed). A step = one pass of process() over every local carrier (fused channelize+demod
kernel, exact edge windows, timing pick + slicer, TS1/TS2 correlator) and, for N > 1, one NCCL all-gather
of the dibit streams. One JSON line on stdout from rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = 1 << 20
TOTAL_CARRIERS = 4096
BYTES_PER_SAMPLE = 8.10      # 8 B read + (1 B dibit + 8 B soft symbol + ~4 B match) per 130 samples
# dram__bytes_read.sum + dram__bytes_write.sum per input sample from the committed ncu pass AT THIS CONFIGURATION
# (profiles/r02_traffic.json <- profiles/r02_launches_4096carriers_ncu.csv): the fused kernel alone and all kernels of a step
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_traffic.json")


def ncu_traffic():
    try:
        t = json.load(open(TRAFFIC_FILE))
        return float(t["k1_dram_bytes_per_sample"]), float(t["step_dram_bytes_per_sample"]), "ncu at %d carriers x 2^20, profiles/r02_traffic.json" % t["carriers"]
    except Exception:
        return None, None, "no committed ncu capture found"


METRIC = "IQ MS/s demodulated"
WORKLOAD = "configs[3]: %d carriers x 2^20 complex64 samples @2.4 MS/s, sharded %d per GPU"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        # under load = upper half of the samples (idle samples before/after the region drag the median down)
        return {"sm_mhz": float(np.median(sorted(sm)[len(sm) // 2:])), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port: same SciPy calls as the reference) on host cores
# ---------------------------------------------------------------------------------------------
_WORKER_X = None


def _cpu_init(barrier):
    """Pool initializer: every worker process makes its own carrier stream ONCE (so input generation never sits inside a
    timed region), runs the path once (imports, caches, page faults) and waits for the other workers."""
    global _WORKER_X
    os.environ["OMP_NUM_THREADS"] = "1"
    from tetraear_b200 import synth
    from oracle import ref_dsp
    _WORKER_X = synth.carrier_iq(N_SAMPLES, os.getpid() % 4096, snr_db=25.0).astype(np.complex128)
    _cpu_worker(1)
    barrier.wait(timeout=600)


def _cpu_worker(n_rep):
    from oracle import ref_dsp
    x = _WORKER_X
    t0 = time.perf_counter()
    for _ in range(n_rep):
        r = ref_dsp.process(x, 0.0, 2.4e6)
        bits = ref_dsp.symbols_to_bits(r["dibits"])
        ref_dsp.match_counts(bits)
    return time.perf_counter() - t0


def cpu_pool(cores=None):
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    ctx = mp.get_context("spawn")
    pool = ctx.Pool(cores, initializer=_cpu_init, initargs=(ctx.Barrier(cores),))
    pool.map(_cpu_worker, [0] * cores, chunksize=1)                       # returns once the workers have left the barrier
    return pool, cores


def cpu_rate(reps, cores, pool):
    """MS/s of the oracle port with one process per host core, each repeating the path on its own resident carrier stream."""
    t0 = time.perf_counter()
    pool.map(_cpu_worker, [reps] * cores, chunksize=1)
    dt = time.perf_counter() - t0
    return cores * reps * N_SAMPLES / dt / 1e6, cores, dt


def cpu_baseline_sample(target_s=20.0):
    """One bounded sample (about target_s seconds of wall time on all host cores) of the same workload."""
    pool, cores = cpu_pool()
    try:
        _, _, dt4 = cpu_rate(4, cores, pool)                                # calibration pass
        reps = int(max(8, min(512, round(4 * target_s / max(dt4, 1e-3)))))
        v, _, dt = cpu_rate(reps, cores, pool)
    finally:
        pool.close(); pool.join()
    return v, cores, dt, reps


def run_reference(a):
    """--impl reference: the reference's CPU algorithm (oracle port: the same SciPy calls the reference makes,
    plus symbols_to_bits and the TS1/TS2 match counts) on every host core; a step = a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pool, cores = cpu_pool()
    budget_s = 200.0
    try:
        _, _, dt4 = cpu_rate(4, cores, pool)
        steps = max(1, a.steps)
        warm = max(0, a.warmup)
        # every worker's input is resident before any timed region; a step repeats the path `reps` times per core
        reps = int(max(8, min(64, (budget_s / (steps + warm)) / max(dt4 / 4, 1e-3))))
        for _ in range(warm):
            cpu_rate(reps, cores, pool)
        t_all = time.perf_counter()
        for _ in range(steps):
            cpu_rate(reps, cores, pool)
        wall = time.perf_counter() - t_all
    finally:
        pool.close(); pool.join()
    v = cores * reps * steps * N_SAMPLES / wall / 1e6
    sample = (f"{steps} steps x {cores} processes x {reps} carrier-blocks x 2^20 samples "
              f"(process + symbols_to_bits + TS1/TS2 match counts), {wall:.1f} s")
    a.out.write(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "MS/s", "n_gpus": a.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % (a.carriers, a.carriers // max(1, a.gpus)) + " (bounded sample of it on the host cores)",
                   "n_samples": N_SAMPLES, "carriers": a.carriers},
        "cpu_baseline": {"value": v, "unit": "MS/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }) + "\n")
    a.out.flush()


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def carrier_snr_db():
    return np.random.default_rng(4).uniform(15.0, 35.0, size=TOTAL_CARRIERS)


def fill_carrier(torch, dst, c, base_d, snr, gen):
    """Carrier c of the workload into dst [N, 2] float32 (on the device): base stream c % n_base + seeded white noise."""
    gen.manual_seed(1000 + c)
    sigma = float(np.sqrt(10.0 ** (-snr[c % len(snr)] / 10.0) / 2.0))
    dst.normal_(0.0, sigma, generator=gen)
    dst += base_d[c % base_d.shape[0]]


def make_inputs(torch, dev, n_local, first_carrier, n_base=16):
    """Carrier c = base[c % n_base] * gain_c + white noise (SNR 15..35 dB, seeded): distinct streams, generated on
    the device so the 32 GiB never cross PCIe. The base streams come from the seeded host generator."""
    from tetraear_b200 import synth
    base = np.stack([synth.carrier_iq(N_SAMPLES, s, snr_db=40.0, alphabet="centred" if s % 2 else "pi4")
                     for s in range(n_base)])
    base_d = torch.view_as_real(torch.from_numpy(base).to(dev))           # [n_base, N, 2] float32
    x = torch.empty((n_local, N_SAMPLES, 2), dtype=torch.float32, device=dev)
    snr = carrier_snr_db()
    gen = torch.Generator(device=dev)
    for i in range(n_local):
        fill_carrier(torch, x[i], first_carrier + i, base_d, snr, gen)
    return x, base_d


def run_ours(a):
    import torch
    import torch.distributed as dist
    from tetraear_b200.processor import SignalProcessor

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from tetraear_b200 import shard
    total = a.carriers
    first_carrier, n_local = shard.partition(total, world, rank)
    sp = SignalProcessor(2.4e6, device=local)
    cap = sp.dibit_capacity(N_SAMPLES)

    x, base_d = make_inputs(torch, dev, n_local, first_carrier)
    # the exchange: the library's own two kernels over NVLink peer memory (default), or pack -> ncclAllGather -> unpack
    transport = "nccl" if a.gather == "nccl" else "p2p"
    packed = None
    if total % world == 0:
        try:
            packed = shard.PackedStreams(total, cap, device=dev, packer=sp, transport=transport)
        except Exception as e:                                # no CUDA IPC on this box: say so and use NCCL
            if transport != "p2p":
                raise
            sys.stderr.write("bench.py: peer-memory exchange unavailable (%s); using NCCL\n" % e)
            transport = "nccl"
            packed = shard.PackedStreams(total, cap, device=dev, packer=sp, transport="nccl")
    if packed is not None:                                    # dibits + lengths in one buffer: ONE all-gather per step
        dib, nd, cap_row = packed.dibits, packed.n_dibits, packed.cap
    else:
        dib = torch.zeros((n_local, cap), dtype=torch.uint8, device=dev)
        nd = torch.zeros(n_local, dtype=torch.int32, device=dev)
        cap_row = cap
    cap = cap_row                                             # row stride of every per-carrier output below
    sym = torch.zeros((n_local, cap + 1, 2), dtype=torch.float32, device=dev)
    ph = torch.zeros(n_local, dtype=torch.int32, device=dev)
    mt = torch.zeros((n_local, 2 * cap, 2), dtype=torch.uint8, device=dev)
    if world > 1 and packed is None:
        all_dib = torch.zeros((total, cap), dtype=torch.uint8, device=dev)
        all_nd = torch.zeros(total, dtype=torch.int32, device=dev)
    # every kernel of the step, the NCCL gather and the timing events share ONE explicit stream
    work = torch.cuda.Stream(device=dev)
    sp.enable_kernel_timing(True)
    fos = None
    if a.fo_max > 0:          # not the BASELINE workload: per-carrier AFC-style offsets exercise the freq_offset != 0 kernel
        fos = np.random.default_rng(6).uniform(-a.fo_max, a.fo_max, size=total)[first_carrier:first_carrier + n_local]

    fused = world > 1 and packed is not None and transport == "p2p" and a.gather == "p2p-fused"

    def step():
        if fused:                                                     # slicer + push in one kernel, then wait + unpack
            sp.process_batch_allgather_device(x.data_ptr(), n_local, N_SAMPLES, N_SAMPLES, dib.data_ptr(), cap, nd.data_ptr(),
                                              sym.data_ptr(), ph.data_ptr(), mt.data_ptr(), packed.out.data_ptr(),
                                              packed.out_n.data_ptr(), stream=work.cuda_stream, freq_offsets=fos)
            return
        sp.process_batch_device(x.data_ptr(), n_local, N_SAMPLES, N_SAMPLES, dib.data_ptr(), cap, nd.data_ptr(),
                                sym.data_ptr(), ph.data_ptr(), mt.data_ptr(), stream=work.cuda_stream, freq_offsets=fos)
        if world > 1:
            if packed is not None:
                packed.gather()                                       # 2-bit pack + exchange + unpack
            else:
                shard.gather_dibits(dib, nd, total, all_dib, all_nd)

    torch.cuda.synchronize()
    with torch.cuda.stream(work):
        for _ in range(a.warmup):
            step()
        torch.cuda.synchronize()
        sp.kernel_time_ms()                                  # drop the warm-up launches from the record
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.5)
        if world > 1:
            dist.barrier()                                   # every rank enters the timed region together
        l0 = sp.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record(work)
        for _ in range(a.steps):
            step()
        ev1.record(work)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = sp.launch_count() - l0
    k_total_ms, k_n = sp.kernel_time_ms()                    # the fused kernel's launches INSIDE the timed region
    phases = sp.last_phase_ms()                              # timeline of the last timed step
    if rank == 0:
        time.sleep(0.3)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = shard.max_over_ranks(ms_total, dev)
    ms_step = ms_total / a.steps
    value = total * N_SAMPLES / (ms_step * 1e-3) / 1e6

    # ---- parity spot check of what was just computed (outside the timed region) ----
    parity = None
    if rank == 0:
        from oracle import ref_dsp
        ok = True
        for i in (0, 1, n_local - 1):
            xi = torch.view_as_complex(x[i]).cpu().numpy()
            r = ref_dsp.process(xi.astype(np.complex128), float(fos[i]) if fos is not None else 0.0, 2.4e6)
            n_i = int(nd[i].item())
            ok &= n_i == len(r["dibits"]) and bool(np.array_equal(dib[i, :n_i].cpu().numpy(), r["dibits"]))
            s = torch.view_as_complex(sym[i, : n_i + 1]).cpu().numpy()
            ok &= bool(np.abs(s - r["symbols"]).max() / np.abs(r["symbols"]).max() < 1e-5)
            bits = ref_dsp.symbols_to_bits(r["dibits"])
            ok &= bool(np.array_equal(mt[i, : 2 * n_i - 21].cpu().numpy(), ref_dsp.match_counts(bits)))
        parity = bool(ok)

    # ---- the gathered streams: one carrier of every other rank, as rank 0 received it, against the oracle on that
    #      carrier's regenerated input (same seeded device generator) ----
    gather_parity = None
    if world > 1 and packed is not None and rank == 0 and fos is None:
        from oracle import ref_dsp
        # what the last timed step left on rank 0
        g_dib = packed.out
        g_nd = packed.out_n if transport == "p2p" else packed.all.view(world, packed.block)[:, packed.packed_bytes:].view(torch.int32)
        tmp = torch.empty((N_SAMPLES, 2), dtype=torch.float32, device=dev)
        gen = torch.Generator(device=dev)
        snr = carrier_snr_db()
        ok, checked = True, []
        for r in range(1, world):
            f_r, n_r = shard.partition(total, world, r)
            i_r = (7 * r) % n_r
            fill_carrier(torch, tmp, f_r + i_r, base_d, snr, gen)
            ref = ref_dsp.process(torch.view_as_complex(tmp).cpu().numpy().astype(np.complex128), 0.0, 2.4e6)
            n_i = int(g_nd[r, i_r].item())
            ok &= n_i == len(ref["dibits"]) and bool(np.array_equal(g_dib[r, i_r, :n_i].cpu().numpy(), ref["dibits"]))
            checked.append(f_r + i_r)
        gather_parity = {"ok": bool(ok), "carriers_checked": checked,
                         "transport": a.gather if transport == "p2p" else transport, "p2p_status": sp.p2p_status() if transport == "p2p" else None,
                         "what": "dibits of one carrier per remote rank, read from rank 0's all-gather output, vs the oracle"}

    # ---- e2e: host buffers through the public C-ABI call, H2D + D2H inside the timed region ----
    ce = min(a.e2e_carriers, n_local)
    hx = torch.view_as_complex(x[:ce]).cpu().pin_memory()
    hx_np = hx.numpy()
    sp._lib.tetra_set_stream(sp._ctx, None)
    sp.process_batch(hx_np, None, want_symbols=True, want_match=False)           # warm-up (allocations)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        r0 = sp.process_batch(hx_np, None, want_symbols=True, want_match=False)
    dt = (time.perf_counter() - t0) / reps
    dt = shard.max_over_ranks(dt, dev)
    e2e = {"value": world * ce * N_SAMPLES / dt / 1e6, "unit": "MS/s",
           "h2d_bytes_per_step": int(ce * N_SAMPLES * 8), "d2h_bytes_per_step": int(ce * (cap + 8 + 8 * (cap + 1))),
           "carriers_per_rank": ce, "timed": "host wall clock around SignalProcessor.process_batch, max over ranks",
           "note": "pinned host IQ -> tetra_process_batch (32 MiB chunks of carriers, the H2D copy of the chunks ahead beside the kernels "
                   "and result copies of the current one) -> host dibits + soft symbols + timing phase; PCIe-bound"}

    # ---- the same through the RTL-SDR byte format: 2 bytes per sample cross PCIe ----
    scale = float(x[:ce].abs().max().item())
    raw = torch.clamp(torch.round((x[:ce] / scale * 0.9 + 1.0) * 127.5), 0, 255).to(torch.uint8).cpu().pin_memory()
    raw_np = raw.numpy()
    sp.process_batch_u8(raw_np, None, want_symbols=True, want_match=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        sp.process_batch_u8(raw_np, None, want_symbols=True, want_match=False)
    dt8 = shard.max_over_ranks((time.perf_counter() - t0) / reps, dev)
    e2e["u8_ingest"] = {"value": world * ce * N_SAMPLES / dt8 / 1e6, "unit": "MS/s", "h2d_bytes_per_step": int(ce * N_SAMPLES * 2),
                        "note": "pinned host uint8 I/Q (RTL-SDR native) -> tetra_process_batch_u8 -> host dibits"}

    if rank == 0:
        peak, peak_src = measured_peak()
        k1_bps, step_bps, traffic_src = ncu_traffic()
        # a large batch is cut into a few launches of the fused kernel (whole slots each, tetra_b200.cu): bytes and time are
        # summed over the launches of a step, i.e. achieved = sum of algorithmic bytes / sum of launch durations
        k_avg = (k_total_ms / a.steps) if k_n else None
        k_per_step = (k_n / a.steps) if k_n else None
        achieved = (BYTES_PER_SAMPLE * n_local * N_SAMPLES / (k_avg * 1e-3) / 1e9) if k_avg else None
        cpu = None
        if world == 1 and not a.no_cpu:
            cpu_v, cpu_cores, cpu_dt, cpu_reps = cpu_baseline_sample()
            cpu = {"value": cpu_v, "unit": "MS/s", "cores": cpu_cores, "kind": "port",
                   "sample": "%d processes x %d carrier-blocks x 2^20 samples (process + symbols_to_bits + TS match counts), %.1f s"
                             % (cpu_cores, cpu_reps, cpu_dt)}
        out = {
            "metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "carriers_per_s": value / 2.4,
            "config": {"workload": WORKLOAD % (total, n_local),
                       "n_samples": N_SAMPLES, "carriers": total,
                       "freq_offset": 0 if fos is None else "uniform +-%g Hz per carrier" % a.fo_max,
                       "l2": "inputs (%.1f GiB per GPU) far larger than L2; no flush needed" % (n_local * N_SAMPLES * 8 / 2**30),
                       "outputs": "dibits + soft symbols + best phase + TS1/TS2 match counts"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": a.traffic if a.traffic is not None else (k1_bps * n_local * N_SAMPLES if k1_bps else None),
                         "traffic_step_all_kernels": step_bps * n_local * N_SAMPLES if step_bps else None,
                         "traffic_source": traffic_src + " (per-sample figure x samples per launch)",
                         "kernel": "k1_channelize_demod<%d>" % (1 if fos is not None else 0), "kernel_ms": k_avg, "kernel_launches_timed": k_n,
                         "kernel_launches_per_step": k_per_step,
                         "kernel_share_of_step": (k_total_ms / ms_total) if k_n else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_sample": BYTES_PER_SAMPLE,
                         "algorithmic_bytes_per_launch": BYTES_PER_SAMPLE * n_local * N_SAMPLES / (k_per_step or 1),
                         "algorithmic_bytes_per_step": BYTES_PER_SAMPLE * n_local * N_SAMPLES,
                         "last_step_timeline_ms": {"fused_kernel_and_edge_join": phases[0], "edge_kernel_span": phases[1],
                                                   "finalize_and_sync": phases[2]}},
            "cpu_baseline": cpu,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "parity_spot_check": parity,
            "gather_parity": gather_parity,
        }
        if world == 1 and not a.no_extra:
            # the other BASELINE configs on the same box, same process (not the bench metric; tools/bench_configs.py):
            # configs[2] (96-channel wideband capture) and configs[4] (waterfall STFT), device-resident, CUDA-event timed
            del x, sym, mt
            torch.cuda.empty_cache()
            try:
                import importlib.util
                spec = importlib.util.spec_from_file_location("bench_configs", os.path.join(ROOT, "tools", "bench_configs.py"))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                oc = mod.run(only=["2", "3", "5"], device=local)
                out["other_configs"] = {k: oc[k] for k in ("config2_one_carrier_fo0", "config2_host_calls", "config2_other_rates_exact_path", "config3_wideband_96ch_device_resident",
                                                           "config5_stft_4096_hop1024", "config5_stft_4096_hop1024_30s") if k in oc}
            except Exception as e:                            # never lose the bench line over the extras
                out["other_configs"] = {"error": repr(e)}
        a.out.write(json.dumps(out) + "\n")
        a.out.flush()
    if world > 1:
        dist.destroy_process_group()
    sp.close()


def _claim_stdout():
    """Libraries (NCCL's version banner, ...) write to fd 1; the contract is ONE JSON line on stdout. Point fd 1
    at stderr for the run and return a writer bound to the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--carriers", type=int, default=TOTAL_CARRIERS)
    ap.add_argument("--e2e-carriers", type=int, default=256)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs appended to the N = 1 line")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "p2p-fused", "nccl"],
                    help="N > 1: exchange of the dibit streams over NVLink peer memory -- a pack + push kernel behind finalize (p2p, default: "
                         "the fastest measured), the push inside the finalize kernel (p2p-fused) -- or pack -> NCCL all-gather -> unpack (nccl)")
    ap.add_argument("--fo-max", type=float, default=0.0, help="per-carrier freq_offset drawn from +-this (Hz); 0 = BASELINE workload")
    ap.add_argument("--traffic", type=float, default=None,
                    help="dram bytes per launch of the fused kernel from an ncu --set full capture (profiles/), if known")
    a = ap.parse_args()
    a.out = _claim_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
