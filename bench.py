#!/usr/bin/env python
"""bench.py — VM cycles/sec (witness rows) of the batched EraVM witness generator on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--vms V] [--transfers T]

Workload (BASELINE.json configs[1]): V = 65 536 parallel VMs per GPU, synthetic ERC-20 transfer bytecode
(T = 8 transfers per VM), full witness trace on.  One "step" = one pass of the hot path over one batch:
every VM runs from its bootloader entry to the end of execution, all witness streams written to HBM.

value     device throughput: inputs resident in HBM, each step = device-side restore of the initial batch state
          (zkb_restore, D2D) + ONE launch of the persistent interpreter kernel; CUDA events, max over ranks.
e2e       the same metric through the public host API with HOST buffers: reset + populate (H2D) + run + packed
          fetch of all six witness streams into pinned host memory (D2H), every step.
roofline  algorithmic bytes (exact byte length of the emitted streams) / average kernel duration, against the
          measured HBM copy bandwidth in MEASURED_PEAKS.json.
--impl reference: the CPU restatement of the reference path (oracle/, "port": the Rust crate cannot be built in
          this image) on all host threads, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "VM cycles/sec (witness rows)"
UNIT = "cycles/s"


def load_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_ncu_traffic(workload):
    """dram bytes per launch from the committed ncu --set full summary, if one exists for this workload."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


class ClockSampler:
    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._idx = device_index
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self._idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=3)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def cpu_reference_run(workload, vm_ids, threads=0, repeats=1):
    """times the oracle (C++ restatement of the reference path) on `vm_ids`; returns (cycles/s, cycles, threads, seconds)."""
    import oracle
    best = None
    for _ in range(repeats):
        cfg = workload.config(len(vm_ids))
        b = oracle.OracleBatch(cfg)
        workload.setup(b, vm_ids)
        t0 = time.perf_counter()
        b.run_threads(0, threads)
        dt = time.perf_counter() - t0
        cycles, _ = b.totals()
        st = b.vm_status()
        assert (st[:, 0] == 1).all(), "oracle: not all VMs ended"
        b.close()
        if best is None or dt < best[1]:
            best = (cycles, dt)
    n_thr = threads or (os.cpu_count() or 1)
    return best[0] / best[1], best[0], min(n_thr, len(vm_ids)), best[1]


def sized_cpu_sample(workload, target_seconds=6.0, max_vms=16384):
    """bounded sample: calibrate on 512 VMs, then size the sample for ~target_seconds of wall time on all threads."""
    rate, cycles, thr, dt = cpu_reference_run(workload, np.arange(512))
    per_vm = cycles / 512
    n = int(min(max_vms, max(1024, rate * target_seconds / per_vm)))
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--vms", type=int, default=65536, help="VMs per GPU")
    ap.add_argument("--transfers", type=int, default=8)
    ap.add_argument("--workload", default="erc20", choices=["erc20", "alu_loop", "keccak", "storage"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from era_zk_evm_b200 import workloads, records

    if args.workload == "erc20":
        w = workloads.Erc20(n_transfers=args.transfers)
        wl_name = f"erc20_transfer x{args.transfers}, {args.vms} VMs/GPU, full witness_trace"
    else:
        w = workloads.WORKLOADS[args.workload]()
        wl_name = f"{args.workload}, {args.vms} VMs/GPU, full witness_trace"
    config = {"workload": wl_name, "vms_per_gpu": args.vms, "seed": hex(w.seed),
              "l2_policy": "inputs+outputs (>2 GB state, >10 GB streams per step) far larger than the 126 MB L2"}

    # ------------------------------------------------------------------ reference arm (CPU) --------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        n = sized_cpu_sample(w)
        ids = np.arange(n)
        for _ in range(max(1, min(args.warmup, 1))):
            cpu_reference_run(w, ids)
        times, cyc = [], 0
        for _ in range(args.steps):
            rate, cyc, thr, dt = cpu_reference_run(w, ids)
            times.append(dt)
        dt = float(np.mean(times))
        value = cyc / dt
        sample = f"{n} VMs of the same workload per step ({cyc} cycles), witness recording on"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u256 (8 x u32 limbs)",
                "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample,
                                 "note": "C++ restatement of the reference path (oracle/), not the Rust crate: no rustc/cargo in this image"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm --------------------------
    import torch
    import torch.distributed as dist
    from era_zk_evm_b200 import GpuVmBatch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    vm_ids = np.arange(args.vms, dtype=np.uint64) + np.uint64(rank * args.vms)   # static VM-range partition
    cfg = w.config(args.vms, device=local_rank)
    batch = GpuVmBatch(cfg)
    w.setup(batch, vm_ids)
    batch.snapshot()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        batch.restore()
        batch.run(sync=False)

    for _ in range(max(args.warmup, 0)):
        step()
        batch.sync()
    cycles, sbytes = batch.totals()
    st = batch.vm_status()
    if not (st[:, 0] == 1).all():
        raise SystemExit(f"bench.py: {int((st[:, 0] != 1).sum())} VMs did not end: {st[st[:, 0] != 1][:4]}")
    alg_bytes = int(sum(sbytes))

    kernel_ms = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            step()
            batch.sync()
            kernel_ms.append(batch.last_run_ms()[0])
        ev1.record()
        barrier()
        total_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    c = torch.tensor([float(cycles)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    total_ms, total_cycles = float(t.item()), float(c.item())
    ms_per_step = total_ms / args.steps
    value = total_cycles / (ms_per_step * 1e-3)

    peak, peak_src = load_peak()
    k_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "zkb_run_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": load_ncu_traffic(args.workload), "peak_source": peak_src, "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_cycle": alg_bytes / cycles,
                "note": "interpreter is integer-issue/latency bound; see profiles/ for pipe utilisation"}

    # ---- e2e through the public host API with host buffers (H2D inputs + D2H witness inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        caps = [int(cycles_k) for cycles_k in sbytes]
        pinned = [torch.empty(max(nb, 8), dtype=torch.uint8, pin_memory=True) for nb in caps]
        e2e_steps = max(1, min(args.steps, 2))

        def e2e_step():
            batch.reset()
            w.setup(batch, vm_ids)
            batch.run()
            for k in range(records.N_STREAMS):
                batch.fetch_stream_packed(k, pinned[k].data_ptr(), pinned[k].numel())

        e2e_step()   # warm-up (also sizes the pack buffer)
        batch.transfer_stats(reset=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        h2d, d2h = batch.transfer_stats()
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        e2e = {"value": total_cycles / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
               "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "path": "GpuVmBatch.reset + Workload.setup (populate_* from host arrays) + run + fetch_stream_packed x6 into pinned host"}
        del pinned

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n = sized_cpu_sample(w)
        rate, cyc, thr, dt = cpu_reference_run(w, np.arange(n))
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": thr, "kind": "port",
                        "sample": f"{n} VMs of the same workload ({cyc} cycles, {dt:.2f} s wall), witness recording on",
                        "note": "C++ restatement of the reference path (oracle/), not the Rust crate"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u256 (8 x u32 limbs)", "data": "synthetic", "config": config, "clocks": clocks.summary(),
                "e2e": e2e, "gpu_launches": args.steps, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "cycles_per_step": total_cycles, "stream_bytes_per_step_per_gpu": dict(zip(records.STREAM_NAMES, sbytes))}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    batch.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
