#!/usr/bin/env python
"""bench.py -- stage-1 structural indexing throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one stage-1 pass (structural indexing + UTF-8 validation) over one synthetic
document per GPU.  Workload at N=1: configs[1] of BASELINE.json, "synthetic 64 MiB
random-structure JSON, stage1 index on 1xB200" (seeded generator simdjson_b200/corpus.py,
SURVEY.md section 8(d) item 2).  At N>1 every rank holds one 64 MiB shard of an N x 64 MiB stream
(byte-range sharding): each rank scans its shard with a speculated incoming
scanner state, the ranks all-gather {6-bit carry transducer, count} over NCCL (the path's one
real exchange step), fold their true incoming state / index base, and re-scan only if the
speculation was wrong -> weak scaling.

The JSON line carries:
  value     input GB/s with the input resident in HBM (device-timed, max over ranks)
  e2e       the same metric through the host-pointer C-ABI call (pinned host input, H2D copy and
            D2H of the n indexes inside the timed region)
  roofline  algorithmic bytes (1 B read per input byte + 4 B written per structural + 12 B of
            sentinels) / the scan kernel's mean duration, measured with CUDA events
            recorded around the kernel on its launch stream, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own CPU stage 1 (oracle/_ref, compiled from the unmodified
            reference) timed on this box's host cores on the same document (bounded sample)
`--impl reference` times that CPU implementation as its own arm (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DOC_BYTES = 64 << 20
ROTATE = 4  # distinct 64 MiB inputs used round-robin: 256 MiB > 126 MB of L2, so no step finds its input in L2
METRIC = "stage1 GB/s (bytes in / s) vs HBM roofline at 1/2/4/8 B200; CPU ref GB/s"
UNIT = "GB/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """samples SM clocks / throttle reasons through NVML while the timed region runs (every ~1 ms; the timed region
    of a 64 MiB stage-1 pass is milliseconds long, far shorter than one `nvidia-smi` invocation)"""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.stop, self.th, self.h = index, [], threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv, self.h, self.max_sm = None, None, None

    def _run(self):
        while not self.stop.is_set():
            try:
                sm = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                try:
                    rs = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    rs = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs))
            except Exception:  # noqa: BLE001
                pass
            self.stop.wait(0.001)

    def __enter__(self):
        if self.h is not None:
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.th:
            self.th.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["nvml unavailable"], "samples": 0}
        sm = sorted(r[0] for r in self.rows)
        reasons = set()
        for _, rs in self.rows:
            for bit, nm in self.REASONS.items():
                if rs & bit:
                    reasons.add(nm)
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(self.max_sm), "reasons": sorted(reasons), "samples": len(self.rows)}


def make_doc(seed_offset):
    from simdjson_b200 import corpus
    return corpus.random_json(DOC_BYTES, seed=corpus.SEED + 7919 * seed_offset)


# =============================================================================== reference arm
def run_reference(args, rank):
    """the reference's own CPU stage 1 (oracle/_ref = unmodified reference, compiled in the build
    container) on this box's host cores; else the oracle port.  Rank 0 only."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    doc = make_doc(0)
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, int(os.environ.get("SJB200_REF_THREADS", cores))))
    if O.have_ref():
        ref = O.Ref("")
        kind, name = "reference", ref.name
        # bounded sample: every thread runs one full 64 MiB stage-1 call per round (independent parsers, the only
        # way the reference can use more than one core for this path); `steps` rounds after `warmup` rounds
        for _ in range(max(0, args.warmup - 1)):
            ref.time(0, doc, 0, threads, 1)
        t0 = time.perf_counter()
        best, err = ref.time(0, doc, 0, threads, max(1, args.steps))
        wall = time.perf_counter() - t0
        gbs = threads * len(doc) / best / 1e9
        one, _ = ref.time(0, doc, 0, 1, 2)
        sample = f"{threads} threads x one {DOC_BYTES >> 20} MiB stage1 call per round, best of {args.steps} rounds ({wall:.1f} s wall); 1 thread: {len(doc)/one/1e9:.2f} GB/s"
    else:
        port = O.Port()
        kind, name, threads = "port", "oracle port (scalar C)", 1
        t0 = time.perf_counter()
        r = port.stage1(doc[: 16 << 20], 0)
        best = time.perf_counter() - t0
        gbs = (16 << 20) / best / 1e9
        err = r.err
        sample = "1 thread x 16 MiB prefix of the document, one pass"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(best * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "synthetic 64 MiB random-structure JSON, stage1 (CPU reference, %s kernel)" % name, "bytes_per_call": DOC_BYTES, "threads": threads},
        "cpu_baseline": {"value": round(gbs, 3), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample, "error_code": err},
        "e2e": {"value": round(gbs, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =============================================================================== our arm
def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist

    import simdjson_b200 as sj
    from simdjson_b200 import capi

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    impl = sj.get_active_implementation(local)
    L = sj.lib()

    # ---- workload: ROTATE distinct documents per rank, resident in HBM before the timed region
    docs = [make_doc(rank * ROTATE + k) for k in range(ROTATE)]
    # (N>1: rank r's document is shard r of the stream doc_0 doc_1 ... doc_{N-1}; every shard is a complete document)
    d_docs = [torch.from_numpy(d.copy()).to(dev) for d in docs]
    pinned = [torch.from_numpy(d.copy()).pin_memory() for d in docs]
    rc, parser = impl.create_dom_parser_implementation(DOC_BYTES)
    if rc != sj.SUCCESS:
        raise RuntimeError("create_dom_parser_implementation failed: " + capi.ERROR_NAMES.get(rc, str(rc)))
    parser.set_option("time_kernel", 1)
    parsers = [parser]
    stream = torch.cuda.Stream(device=dev)  # the stream every scan of the timed region is launched on
    torch.cuda.set_stream(stream)
    words = L.sjb200_index_words(DOC_BYTES)
    d_idxs = [torch.empty(words, dtype=torch.int32, device=dev) for _ in range(2)]
    shard_res = [torch.zeros(3, dtype=torch.int64, device=dev) for _ in range(128)]
    shard_gather = [torch.zeros(3 * world, dtype=torch.int64, device=dev) for _ in range(128)] if world > 1 else []
    gather_in = torch.zeros(4, dtype=torch.int64, device=dev)
    gather_out = torch.zeros(4 * world, dtype=torch.int64, device=dev) if world > 1 else None

    kernel_ms, n_struct = [], []

    def sharded_step(i):
        """one sharded pass (simdjson_b200/sharding.py): scan with speculated state 0, all-gather {transducer,count,flags},
        fold, re-scan if the speculation was wrong"""
        from simdjson_b200 import sharding

        def scan(state_in):
            rc, res = parser.stage1_shard_device(d_docs[i % ROTATE], state_in, rank == world - 1, d_idx=d_idxs[i % 2], stream=stream)
            if rc != 0:
                raise RuntimeError("shard scan failed: " + parser.last_cuda_error())
            return int(res.ttable), int(res.count), int(res.flags)

        def all_gather(v):
            gather_in.copy_(torch.from_numpy(v))
            dist.all_gather_into_tensor(gather_out, gather_in)
            return gather_out.view(world, 4).cpu().numpy()

        r = sharding.exchange(scan, rank, world, all_gather)
        return r["count"], r["base"]

    def run_steps(k, record):
        t_ev0, t_ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        parser.get_stat("kernel_ms_mean")  # reset the per-launch event log
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_ev0.record(stream)
        if world == 1:
            # K steps = K documents through the batch entry point: every scan is queued back to back on `stream`
            res = parser.stage1_device_batch([d_docs[i % ROTATE] for i in range(k)], [d_idxs[i % 2] for i in range(k)], sj.REGULAR, stream=stream)
            for err, n in res:
                if err != sj.SUCCESS:
                    raise RuntimeError("stage1 failed: " + capi.ERROR_NAMES.get(err, str(err)) + " " + parser.last_cuda_error())
                if record:
                    n_struct.append(n)
        else:
            # pipelined speculation: scan_k -> all_gather_k are queued for every step without a host round trip; the
            # host verifies all K speculations afterwards and falls back to the synchronous protocol for a step whose
            # speculation was wrong (never the case for shards that start at a document boundary)
            from simdjson_b200 import sharding
            for i in range(k):
                rc = parser.stage1_shard_device_enqueue(d_docs[i % ROTATE], shard_res[i % len(shard_res)], d_idx=d_idxs[i % 2], stream=stream)
                if rc != 0:
                    raise RuntimeError("shard scan failed: " + parser.last_cuda_error())
                dist.all_gather_into_tensor(shard_gather[i % len(shard_res)], shard_res[i % len(shard_res)])
            t_ev1.record(stream)
            torch.cuda.synchronize()
            for i in range(min(k, len(shard_res))):
                okk, state_in, _base = sharding.verify_speculation(shard_gather[i].view(world, 3).cpu().numpy(), rank)
                if not okk:
                    sharded_step(i)
                if record:
                    n_struct.append(sharding.unpack_result(shard_res[i].cpu().numpy())[1])
        if world == 1:
            t_ev1.record(stream)
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if record:
            kernel_ms.append(parser.get_stat("kernel_ms_mean"))
        return t_ev0.elapsed_time(t_ev1)

    launches0 = sum(p.get_stat("launches") for p in parsers)
    run_steps(max(3, args.warmup), False)
    launches1 = sum(p.get_stat("launches") for p in parsers)
    with ClockSampler(local) as clocks:
        total_ms = run_steps(args.steps, True)
    launches2 = sum(p.get_stat("launches") for p in parsers)
    _ = launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * DOC_BYTES / (ms_per_step * 1e-3) / 1e9

    # ---- e2e through the host-pointer C-ABI call (pinned host input; H2D + D2H inside the timed region)
    p = parsers[0]
    hosts = [x.numpy() for x in pinned]
    for i in range(3):
        p.stage1(hosts[i % ROTATE], sj.REGULAR)
    e2e_steps = max(3, min(args.steps, 10))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_n = 0
    for i in range(e2e_steps):
        rc = p.stage1(hosts[i % ROTATE], sj.REGULAR)
        assert rc == 0, rc
        e2e_n = p.n_structural_indexes
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * DOC_BYTES * e2e_steps / float(te.item()) / 1e9

    if rank == 0:
        peak, peak_src = peaks()
        kms = float(np.mean(kernel_ms))
        nmean = float(np.mean(n_struct))
        algo_bytes = DOC_BYTES + 4.0 * nmean + 12.0
        achieved = algo_bytes / (kms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "synthetic 64 MiB random-structure JSON, stage1 index on 1xB200 (BASELINE.json configs[1])" if world == 1 else
                       f"{world} x 64 MiB shards of one random-structure JSON stream, stage1 sharded by byte range + NCCL carry/offset all-gather",
                       "bytes_per_gpu_per_step": DOC_BYTES, "mode": "regular" if world == 1 else "shard", "structurals_per_step": int(nmean),
                       "l2": f"{ROTATE} distinct inputs used round-robin ({ROTATE * DOC_BYTES >> 20} MiB > 126 MB L2)",
                       "api": "sjb200_stage1_dev_batch: the K documents of the timed region are queued back to back on one stream" if world == 1 else "sjb200_stage1_shard_dev_enqueue + NCCL all_gather_into_tensor per step, verified after the timed region"},
            "clocks": clocks.summary(),
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": DOC_BYTES, "d2h_bytes_per_step": int(4 * e2e_n + 24), "steps": e2e_steps},
            "gpu_launches": int(launches2 - launches1),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": ncu_traffic()[0],
                         "traffic_source": ncu_traffic()[1], "peak_source": peak_src, "kernel": "sjb200::scan4_deferred_kernel / scan4_kernel (sjb200_scan4.cuh)", "kernel_ms": round(kms, 5), "algorithmic_bytes": int(algo_bytes),
                         "input_gbs_kernel_only": round(DOC_BYTES / (kms * 1e-3) / 1e9, 1)},
        }
        line["cpu_baseline"] = cpu_baseline(docs[0])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic():
    """DRAM bytes of one launch of the stage-1 kernel on the bench document, from the committed `ncu --set full` capture
    (profiles/, produced by tools/run_gpu_round.sh + tools/ncu_summary.py); None when there is no capture"""
    path = os.path.join(ROOT, "profiles", "r1b_scan4_kernel_ncu_full.json")
    try:
        d = json.load(open(path))
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = d[k]
            tot += float(v) * unit[u]
        return int(tot), os.path.relpath(path, ROOT)
    except Exception:  # noqa: BLE001
        return None, None


def cpu_baseline(doc):
    """rank 0, N=1 style bounded sample of the reference's CPU stage 1 on this host"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    cores = os.cpu_count() or 1
    if O.have_ref():
        ref = O.Ref("")
        one, _ = ref.time(0, doc, 0, 1, 3)
        threads = max(1, min(cores, int(os.environ.get("SJB200_REF_THREADS", cores))))
        many, _ = ref.time(0, doc, 0, threads, 3)
        return {"value": round(threads * len(doc) / many / 1e9, 3), "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": f"{ref.name} kernel; {threads} threads x one 64 MiB stage1 call, best of 3 rounds; single thread: {len(doc)/one/1e9:.2f} GB/s"}
    port = O.Port()
    t0 = time.perf_counter()
    port.stage1(doc[: 16 << 20], 0)
    dt = time.perf_counter() - t0
    return {"value": round((16 << 20) / dt / 1e9, 3), "unit": UNIT, "cores": 1, "kind": "port", "sample": "scalar C port, 16 MiB prefix, one pass"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and args.gpus > 1:
        sys.stderr.write(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})\n")
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
