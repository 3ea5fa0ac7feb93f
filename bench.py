#!/usr/bin/env python3
"""bench.py — Msamples/s through RRC -> GFSK demod -> DMR decoder for N x 4096 synthetic 48 kHz channels.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
  * a step = one pass of the hot path over one batch: 4096 channels x 48000 samples (1 s of signal) per GPU
    (BASELINE.json configs[1]); the per-channel streams continue across steps (state is carried);
  * `value`   : whole-job Msamples/s with the batch resident in HBM, timed with CUDA events on the launching
                stream over exactly K steps, max over ranks;
  * `e2e`     : the same metric through the C ABI with HOST buffers: pinned host -> device copy of every batch,
                the three kernels, device -> host read of the decoded frames + metadata events and their replay;
  * `roofline`: the dominant kernel (K1, the RRC FIR): algorithmic bytes (8 B/sample) / its live CUDA-event time;
  * `cpu_baseline` (rank 0, N = 1): the reference's own CPU modules (oracle/_ref, compiled from the unmodified
                reference sources) on a bounded sample of the same batch, all host cores.
`--impl reference` times only that CPU arm, same metric/config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CHANNELS_PER_GPU = 4096
SAMPLES_PER_STEP = 48000
METRIC = "Msamples/s RRC->GFSK->DMR pipe"
ALGO_BYTES_PER_SAMPLE_K1 = 8.0       # K1 stand-alone: 4 B in + 4 B out
ALGO_BYTES_PER_SAMPLE_PIPE = 4.119   # fused-pipe figure, reported for context

# The bench line is the DMR pipe (BASELINE configs[1]).  `--workload` runs the other pipes through the same
# harness (BASELINE configs[2] / [4] and the SURVEY 8f decoders): name, default channels per GPU, samples per
# symbol, dominant stage (0 = K1 RRC, 1 = K2 demodulator) and its algorithmic bytes per sample.
WORKLOADS = {
    "dmr": dict(channels=4096, sps=10, dominant=0, algo=8.0, pipe_bytes=4.119,
                desc="WideRrcFilter -> GfskDemodulator(10) -> Dmr::Decoder (BASELINE configs[1])"),
    "ysf": dict(channels=8192, sps=10, dominant=0, algo=8.0, pipe_bytes=4.12,
                desc="WideRrcFilter -> GfskDemodulator(10) -> Ysf::Decoder (BASELINE configs[2])"),
    "nxdn": dict(channels=4096, sps=20, dominant=0, algo=8.0, pipe_bytes=4.06,
                 desc="NarrowRrcFilter -> GfskDemodulator(20) -> Nxdn::Decoder (SURVEY 8f rank 1)"),
    "dstar": dict(channels=8192, sps=10, dominant=1, algo=4.1, pipe_bytes=4.11,
                  desc="FskDemodulator(10) -> DStar::Decoder (SURVEY 8f rank 3)"),
    "pocsag": dict(channels=32768, sps=40, dominant=1, algo=4.025, pipe_bytes=4.03,
                   desc="FskDemodulator(40, invert) -> Pocsag::Decoder (BASELINE configs[4])"),
}


def workload_signal(name, channels, n, seed, device):
    """Seeded synthetic input [channels, pitch] float32 of one workload (device or cpu)."""
    import numpy as np
    from digiham_b200 import synth
    if name == "dmr":
        return synth.dmr_channel_bank(channels, n, seed=seed, device=device)[0]
    w = WORKLOADS[name]
    gen = {"ysf": lambda k: synth.ysf_symbols(12, seed=seed * 131 + k, mode="mix", lead_in=0),
           "nxdn": lambda k: synth.nxdn_symbols(16, seed=seed * 131 + k, lead_in=0),
           "dstar": lambda k: np.concatenate([np.tile(np.array([1, 0], dtype=np.uint8), 100),
                                              synth.dstar_symbols(60, seed=seed * 131 + k, lead_in=0)]),
           "pocsag": lambda k: synth.pocsag_bits([(1000 + k, 3, "B200 BENCH %d" % k), (77 + k, 3, "73")],
                                                 seed=seed * 131 + k, lead_in=0, preamble=200)}[name]
    levels = {"ysf": synth.LEVELS4, "nxdn": synth.LEVELS4, "dstar": synth.LEVELS2, "pocsag": synth.LEVELS2[::-1].copy()}[name]
    nsym = n // w["sps"] + 8
    pool = np.stack([np.resize(gen(k), nsym) for k in range(24)])
    rng = np.random.default_rng(seed)
    sym = np.empty((channels, nsym), dtype=np.uint8)
    shifts = rng.integers(0, nsym, size=channels)
    for c in range(channels):
        sym[c] = np.roll(pool[c % 24], int(shifts[c]))
    return synth.modulate_batch(sym, n, sps=w["sps"], levels=levels, amplitude=rng.choice([0.25, 0.5, 0.8], size=channels),
                                ppm=rng.choice([0.0, 20.0, -20.0, 50.0, -50.0], size=channels),
                                phase=rng.integers(0, 4 * w["sps"], size=channels).astype(np.float64),
                                snr_db=rng.choice([np.inf, 20.0, 12.0, 8.0], size=channels), seed=seed + 1, device=device)


def peaks():
    """HBM peak for the roofline: the driver-written MEASURED_PEAKS.json when present (any numeric entry whose key
    names HBM bandwidth; a kernel timed inside a long step uses the sustained figure when both exist), else the
    fallback of B200_PROFILING.md."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                d = json.load(f)
            flat = {}

            def walk(prefix, node):
                if isinstance(node, dict):
                    for k, v in node.items():
                        walk(prefix + "." + str(k) if prefix else str(k), v)
                elif isinstance(node, (int, float)) and not isinstance(node, bool):
                    flat[prefix.lower()] = float(node)

            walk("", d)
            cands = {k: v for k, v in flat.items() if "hbm" in k and v > 100.0}
            for pref in ("sustained", "hbm_gbs", ""):
                for k in sorted(cands):
                    if pref in k:
                        v = cands[k]
                        return (v * 1000.0 if v < 50.0 else v), "measured (MEASURED_PEAKS.json:%s)" % k
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed regions (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
                for k, nme in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "sm_mhz_min": min(sm) if sm else None, "power_w_median": statistics.median(power) if power else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


def proto_ids(name):
    import oracle_lib
    import digiham_b200 as dh
    return {"dmr": (dh.PROTO_DMR, oracle_lib.PROTO_DMR), "ysf": (dh.PROTO_YSF, oracle_lib.PROTO_YSF),
            "nxdn": (dh.PROTO_NXDN, oracle_lib.PROTO_NXDN), "dstar": (dh.PROTO_DSTAR, oracle_lib.PROTO_DSTAR),
            "pocsag": (dh.PROTO_POCSAG, oracle_lib.PROTO_POCSAG)}[name]


def reference_arm(args, x_host_rows, cores):
    """Times the reference CPU modules (or the port when the compiled reference is absent) on host cores.
    x_host_rows: float32 numpy [channels, n]: one step's batch.  Runs args.warmup + args.steps steps unless the time
    budget (--ref-budget seconds) ends the run earlier.  Returns (Msamples/s, kind, sample description, seconds per
    step, steps timed, warm-up steps done)."""
    import oracle_lib
    orc = oracle_lib.best()
    nch, n = x_host_rows.shape

    def one():
        t0 = time.perf_counter()
        orc.pipe_batch(args.orc_proto, x_host_rows, threads=cores, chunk=4096)
        return time.perf_counter() - t0

    dt0 = one()                                    # first warm-up step, also sizes the run
    warm, steps = max(1, args.warmup), args.steps
    if dt0 * (warm + steps) > args.ref_budget:     # a CPU loop needs no long warm-up; keep the timed steps
        warm = 1
        steps = max(1, min(steps, int(args.ref_budget / dt0) - 1))
    for _ in range(warm - 1):
        one()
    times = [one() for _ in range(steps)]
    sec = sum(times) / len(times)
    return (nch * n / sec / 1e6, orc.kind, "%d channels x %d samples per step (the full per-GPU batch), %d step(s) timed"
            % (nch, n, len(times)), sec, len(times), warm)


def s16_conversion_seconds(x_host_rows, cores):
    """`csdr convert -i s16 -o float` of one step's batch on all host cores (numpy releases the GIL): the extra CPU
    work of the reference pipe when the ingest is int16 (reference examples/dmr-decoder.sh:13-15)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    s = np.clip(np.rint(x_host_rows * 20000.0), -32768, 32767).astype(np.int16)
    out = np.empty_like(x_host_rows)
    parts = np.array_split(np.arange(s.shape[0]), max(1, cores))

    def conv(idx):
        if len(idx):
            np.divide(s[idx[0]:idx[-1] + 1].astype(np.float32), np.float32(32767), out=out[idx[0]:idx[-1] + 1])

    best = None
    with ThreadPoolExecutor(max_workers=max(1, cores)) as ex:
        for _ in range(3):
            t0 = time.perf_counter()
            list(ex.map(conv, parts))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return best


def bind_to_gpu_numa_node(local_rank):
    """N > 1: keep this rank (and the pinned host blocks it allocates) on the NUMA node of its GPU, so that the
    per-GPU PCIe uploads of the e2e arm do not cross the socket interconnect.  Best effort; returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = "/sys/bus/pci/devices/%s/local_cpulist" % bus.lower()[-12:]
        if not os.path.exists(path):
            path = "/sys/bus/pci/devices/%s/local_cpulist" % bus.lower()
        cpus = set()
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "%d cpus local to %s" % (len(allowed), bus)
        return "no local cpu in the allowed set"
    except Exception as ex:   # affinity is an optimisation, never a requirement
        return "unavailable (%s)" % type(ex).__name__


def build_host_sample(workload, channels, n, seed):
    """CPU-side generation of a bounded sample of the workload (used by --impl reference on a box whose GPU arm
    is not running): same generator, same seed, first `channels` channels."""
    x = workload_signal(workload, channels, n, seed, "cpu")
    return x[:, :n].contiguous().numpy()


def sharded_arm(args, dh, dist, dev, rank, world, stream, barrier, max_over_ranks, orc_proto):
    """BASELINE configs[3]: `--shard-channels` (8192) channels per GPU, all blocks on rank 0, scatter -> compute ->
    gather pipelined through dh_shard_*.  Returns the `nccl` object of the bench line (all ranks call this)."""
    import numpy as np
    import torch
    from digiham_b200 import shard, synth
    Cs, L = args.shard_channels, args.samples
    total = Cs * world
    out = {"channels_per_gpu": Cs, "channels_total": total, "samples_per_channel_per_step": L}
    check_per_rank = 64
    # rank 0 holds two different steps' blocks of ALL channels (per-rank seeds differ, so a misplaced shard would
    # show); the int16 blocks are the float32 ones quantised like an rtl_fm discriminator output
    pitch32 = (L + 3) & ~3
    pitch16 = (L + 7) & ~7
    blocks32 = blocks16 = ref32 = ref16 = None
    if rank == 0:
        blocks32 = [torch.zeros((total, pitch32), dtype=torch.float32, device=dev) for _ in range(2)]
        blocks16 = [torch.zeros((total, pitch16), dtype=torch.int16, device=dev) for _ in range(2)]
        ref32, ref16 = [], []
        for r in range(world):
            xr = synth.dmr_channel_bank(Cs, 2 * L, seed=777 + r, device=dev)[0][:, :2 * L]
            xs = torch.clamp(torch.round(xr * 20000.0), -32768, 32767).to(torch.int16)
            for k in range(2):
                blocks32[k][r * Cs:(r + 1) * Cs, :L] = xr[:, k * L:(k + 1) * L]
                blocks16[k][r * Cs:(r + 1) * Cs, :L] = xs[:, k * L:(k + 1) * L]
            ref32.append(xr[:check_per_rank].cpu().numpy())
            ref16.append(xs[:check_per_rank].cpu().numpy().astype(np.float32) / np.float32(32767))
            del xr, xs
    for fmt_name, fmt, dtype in (("f32", dh.FMT_F32, torch.float32), ("s16", dh.FMT_S16, torch.int16)):
        sp = shard.ShardedPipe(total, dh.PROTO_DMR, max_chunk=L, device=dev, fmt=fmt)
        pitch = sp.pitch
        blocks = blocks16 if fmt == dh.FMT_S16 else blocks32
        ref_rows = ref16 if fmt == dh.FMT_S16 else ref32
        assert pitch == (pitch16 if fmt == dh.FMT_S16 else pitch32)
        # parity first, on the fresh streams: two steps through scatter / kernels / gather / read-back, then the
        # gathered frames + metadata of the first `check_per_rank` channels of EVERY rank against the reference
        sp.submit(blocks[0] if rank == 0 else None, L, scatter=True)
        sp.submit(blocks[1] if rank == 0 else None, L, scatter=True)
        sp.collect_step()
        sp.collect_step()
        parity = None
        if rank == 0:
            try:
                import oracle_lib
                orc = oracle_lib.best()
                _, outs, metas = orc.pipe_batch(orc_proto, np.concatenate(ref_rows, axis=0), threads=os.cpu_count() or 1,
                                                chunk=4096, meta_cap=1 << 15)
                bad, nbytes = [], 0
                for r in range(world):
                    for c in range(check_per_rank):
                        g = r * Cs + c
                        i = r * check_per_rank + c
                        nbytes += len(outs[i])
                        if sp.output(g) != outs[i].tobytes() or sp.meta(g) != metas[i]:
                            bad.append(g)
                parity = {"checked_channels": world * check_per_rank, "per_rank": check_per_rank, "steps": 2,
                          "reference_bytes": nbytes, "equal": not bad, "oracle": orc.kind}
                if bad:
                    parity["first_mismatches"] = bad[:8]
            except Exception as ex:
                parity = {"equal": None, "error": str(ex)}
            sp.clear()
        # timed loop: device-resident blocks on the ingest rank, results left in the gather buffers (discard)
        def run(k):
            for i in range(k):
                sp.submit(blocks[i & 1] if rank == 0 else None, L, scatter=True)
                sp.discard_step()
            sp.sync()
        run(max(3, args.warmup))
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(args.steps)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        barrier()
        # the same loop with the read-back + metadata replay of ALL channels on the ingest rank (wall clock)
        k2 = max(2, min(args.steps, 10))
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        sp.submit(blocks[0] if rank == 0 else None, L, scatter=True)
        for i in range(1, k2):
            sp.submit(blocks[i & 1] if rank == 0 else None, L, scatter=True)
            sp.collect_step()
            if rank == 0:
                sp.clear()
        sp.collect_step()
        sp.sync()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        launches, wire_bytes, _ = sp.stats()
        out["gather_path" + ("" if fmt_name == "f32" else "_s16")] = {1: "nccl send/recv", 2: "pack kernel stores -> CUDA IPC root buffer"}.get(sp.gather_path, "none")
        out["scatter_path" + ("" if fmt_name == "f32" else "_s16")] = {1: "nccl send/recv", 2: "copy engines -> CUDA IPC peer slots"}.get(sp.scatter_path, "none")
        esz = 2 if fmt == dh.FMT_S16 else 4
        scatter_bytes = (world - 1) * Cs * pitch * esz
        key = "" if fmt_name == "f32" else "_s16"
        out["pipelined_value" + key] = total * L / (ms * 1e-3) / 1e6
        out["pipelined_ms_per_step" + key] = ms
        out["pipelined_collect_value" + key] = total * L * k2 / dt / 1e6
        out["scatter_bytes_per_step" + key] = scatter_bytes
        out["scatter_GBps_egress_if_bound" + key] = scatter_bytes / (ms * 1e-3) / 1e9
        out["parity" + key] = parity
        out["gather_wire_bytes_per_rank_per_step"] = wire_bytes
        sp.close()
    del blocks32, blocks16
    torch.cuda.empty_cache()
    out["unit"] = "Msamples/s"
    out["note"] = ("dh_shard_*: rank 0 holds every channel's block in HBM; per step NCCL scatter (grouped send/recv) -> "
                   "K1/K2/K3 on each rank -> pack -> NCCL gather to rank 0, consecutive steps overlapped on three streams "
                   "and two communicators; both collectives inside the timed loop.  The ingest rank's NVLink egress "
                   "bounds the step (scatter bytes / step time = scatter_GBps_egress_if_bound); `value` above is the "
                   "same kernels with every rank's block already local")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dmr", choices=sorted(WORKLOADS),
                    help="dmr = the bench line (BASELINE configs[1]); the others run the same harness on another pipe")
    ap.add_argument("--channels", type=int, default=None, help="channels per GPU (default: the workload's)")
    ap.add_argument("--samples", type=int, default=SAMPLES_PER_STEP, help="samples per channel per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline sample")
    ap.add_argument("--ref-budget", type=float, default=150.0,
                    help="--impl reference: wall-clock budget (s); fewer steps are timed when it would be exceeded")
    ap.add_argument("--no-shard", action="store_true", help="N > 1: skip the sharded (scatter/gather) pipeline arm")
    ap.add_argument("--shard-channels", type=int, default=8192,
                    help="N > 1: channels per GPU of the sharded pipeline arm (BASELINE configs[3]: 65536 / 8)")
    ap.add_argument("--wc", action="store_true", help="e2e arms: write-combined pinned host blocks")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-async", action="store_true",
                    help="device arm: three kernels back to back on one stream instead of cross-step pipelining")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    wl = WORKLOADS[args.workload]
    if args.channels is None:
        args.channels = wl["channels"]
    metric = METRIC if args.workload == "dmr" else "Msamples/s %s pipe" % args.workload.upper()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    config = {"workload": "%d ch/GPU x %d samples/step synthetic 48 kHz %s: %s" % (
                  args.channels, args.samples, args.workload.upper(), wl["desc"]),
              "channels_per_gpu": args.channels, "samples_per_channel_per_step": args.samples,
              "l2": "inputs larger than L2 (%.0f MB/step/GPU)" % (args.channels * args.samples * 4 / 1e6),
              "sharding": "contiguous channel ranges per rank; `value`/`e2e`: every rank's block local, no data-path "
                          "collective; N > 1 adds `nccl`: one ingest rank, NCCL scatter/gather pipelined (configs[3])"}

    if args.impl == "reference":
        if rank != 0:
            return
        # the full per-GPU batch of the configuration (4096 channels x 48000 samples for the bench line), every step
        x = build_host_sample(args.workload, args.channels, args.samples, seed=1234)
        args.orc_proto = proto_ids(args.workload)[1]
        val, kind, sample, sec, steps_done, warm_done = reference_arm(args, x, cores)
        line = {"metric": metric, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": steps_done,
                "warmup": warm_done, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        if wl["dominant"] == 0:
            conv = s16_conversion_seconds(x, cores)
            line["e2e_s16"] = {"value": args.channels * args.samples / (sec + conv) / 1e6, "unit": "Msamples/s",
                               "convert_ms_per_step": conv * 1e3,
                               "note": "int16 ingest: csdr convert -i s16 -o float (numpy float32 division on all cores) "
                                       "in front of the same reference modules"}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import digiham_b200 as dh
    from digiham_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA path is the only path (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    arm = {}     # how this arm ran (kept out of `config`, which both arms print identically)
    if world > 1:
        arm["numa"] = bind_to_gpu_numa_node(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    C, L = args.channels, args.samples
    dh_proto, orc_proto = proto_ids(args.workload)
    x = workload_signal(args.workload, C, L, 1234 + rank, dev)             # [C, pitch] float32 in HBM
    pipe = dh.Pipe(C, dh_proto, max_chunk=L, device=dev)
    stream = torch.cuda.current_stream()

    sampler = ClockSampler(local_rank)
    # ---- kernel-level arm: batch resident in HBM ---------------------------------------------------------------
    # cross-step software pipelining (dh_pipe_set_async): K1 of step i+1 overlaps K2 + decoder of step i on two
    # internal streams; every kernel of all K steps still runs inside the timed region (joined by pipe.sync).
    # Per-stage timing events are OFF in the timed region; the stage times come from a second pass right after it.
    pipe.set_async(not args.no_async)
    arm["pipelining"] = ("none" if args.no_async or wl["dominant"] != 0 else
                         "K1(i+1) overlaps K2+K3(i) on two streams (dh_pipe_set_async)")
    for _ in range(args.warmup):
        pipe.process(x, n=L)
        pipe.discard()
    pipe.sync()
    torch.cuda.synchronize()
    arm["demod_schedule"] = ("split: search chain -> per-symbol window sums -> per-block slicing (3 kernels)"
                             if pipe.demod_kernels_per_call == 3 else "one kernel")
    launches0 = pipe.launch_count
    barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        pipe.process(x, n=L)
        pipe.discard()               # results stay in HBM; only the device-side counters are reset
    pipe.sync()                      # the timing stream waits for the last decoder kernel
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    launches = pipe.launch_count - launches0
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = world * C * L / (ms_step * 1e-3) / 1e6
    # second pass, same schedule, with CUDA events around every kernel on its launching stream: K1's in-region time
    prof_steps = min(args.steps, 50)
    pipe.set_profiling(True)
    for _ in range(prof_steps):
        pipe.process(x, n=L)
        pipe.discard()
    pipe.sync()
    torch.cuda.synchronize()
    stage_ms, calls = pipe.stage_times()
    pipe.set_async(False)
    # the same kernels once more WITHOUT cross-step overlap: K1's launch time with the SMs to itself
    iso_steps = 0 if args.no_async else min(10, args.steps)
    for _ in range(iso_steps):
        pipe.process(x, n=L)
        pipe.discard()
    iso_ms, iso_calls = pipe.stage_times() if iso_steps else ([0.0, 0.0, 0.0], 0)
    pipe.set_profiling(False)

    # ---- end-to-end arms: host buffers through the C ABI -------------------------------------------------------
    def e2e_arm(dtype):
        """dh_pipe_submit_host[_s16] / dh_pipe_collect_step with two pinned blocks alternating: the upload of step
        k+1 overlaps kernels, read-back and metadata replay of step k; every byte of every step crosses PCIe inside
        the timed region.  Returns the e2e object."""
        s16 = dtype == torch.int16
        pitch = pipe.host_pitch_s16 if s16 else pipe.host_pitch
        esz = 2 if s16 else 4
        src = x[:, :L]
        if s16:
            src = torch.clamp(torch.round(src * 20000.0), -32768, 32767).to(torch.int16)
        blocks = [dh.PinnedBlock(C, pitch, dtype=dtype, write_combined=args.wc) for _ in range(2)]
        for b in blocks:
            b.tensor.zero_()
            b.tensor[:, :L].copy_(src)
        bufs = [b.tensor for b in blocks]

        def steps(k):
            pipe.submit(bufs[0], n=L)
            for i in range(1, k):
                pipe.submit(bufs[i & 1], n=L)   # H2D + K1 + K2 + decoder kernel, asynchronous
                pipe.collect_step()             # D2H of frames/events + host metadata replay of the previous step
                pipe.decoder.clear()
            pipe.collect_step()
            pipe.decoder.clear()

        steps(min(args.warmup, 3))
        # the link itself: the same pinned block copied host -> device with nothing else going on on this GPU, all
        # ranks at the same time (barrier first), so the N > 1 figure is the CONCURRENT link ceiling of the box
        link = torch.empty((C, pitch), dtype=dtype, device=dev)
        for _ in range(2):
            link.copy_(bufs[0], non_blocking=True)
        torch.cuda.synchronize()
        barrier()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record(stream)
        for _ in range(5):
            link.copy_(bufs[0], non_blocking=True)
        l1.record(stream)
        torch.cuda.synchronize()
        link_ms = max_over_ranks(l0.elapsed_time(l1) / 5)
        del link
        _, d2h0 = pipe.decoder.stats()
        k_e2e = args.steps
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        steps(k_e2e)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        _, d2h1 = pipe.decoder.stats()
        for b in blocks:
            b.close()
        val = world * C * L * k_e2e / dt / 1e6
        link_rate = world * C * L / link_ms / 1e3
        return {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": C * pitch * esz,
                "d2h_bytes_per_step": (d2h1 - d2h0) // k_e2e, "steps": k_e2e, "ms_per_step": dt / k_e2e * 1e3,
                "host_sample_format": "int16 (csdr convert fused into K1)" if s16 else "float32",
                "pinned": "write-combined" if args.wc else "default",
                "h2d_copy_only": {"ms_per_step_block": link_ms, "GBps_per_gpu": C * pitch * esz / link_ms / 1e6,
                                  "GBps_all_gpus": world * C * pitch * esz / link_ms / 1e6,
                                  "Msamples_per_s": link_rate, "concurrent_ranks": world,
                                  "note": "pinned host -> device copy of one step's block alone, all ranks at once "
                                          "(max over ranks): the PCIe ceiling of e2e on this box"},
                "frac_of_link": val / link_rate,
                "path": "dh_pipe_submit_host%s (pinned H2D + 3 kernels) / dh_pipe_collect_step (D2H + metadata replay) "
                        "per step, two steps in flight" % ("_s16" if s16 else "")}

    e2e = e2e_s16 = None
    if not args.no_e2e:
        e2e = e2e_arm(torch.float32)
        if wl["dominant"] == 0:          # pipes with an RRC stage take int16 samples
            e2e_s16 = e2e_arm(torch.int16)
    clocks = sampler.stop() if rank == 0 else None

    # ---- N > 1: the sharded pipeline of BASELINE configs[3] ------------------------------------------------------
    # One ingest rank holds the blocks of ALL channels; per step: NCCL scatter -> K1/K2/K3 on every rank -> pack ->
    # NCCL gather of frames + metadata events to the ingest rank, the three phases of consecutive steps overlapping
    # (dh_shard_*, digiham_b200/csrc/shard.cu).  Both collectives are INSIDE the timed loop.
    nccl = None
    if world > 1 and not args.no_shard and args.workload == "dmr":
        nccl = sharded_arm(args, dh, dist, dev, rank, world, stream, barrier, max_over_ranks, orc_proto)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1) ------------------------------------------------------------------
    peak, peak_src = peaks()
    # stage sums cover every launch of the profiled pass; per step = / prof_steps
    dom = wl["dominant"]
    algo_bytes = wl["algo"]
    k1_ms = stage_ms[dom] / prof_steps           # launch time of the dominant kernel (K1, or K2 for the FSK pipes)
    achieved = C * L * algo_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("channels") == C and tj.get("samples") == L and args.workload == "dmr":
            traffic = tj.get("dram_bytes_per_launch")
    iso = None
    if iso_calls:
        iso_k1 = iso_ms[dom] / iso_calls
        iso_ach = C * L * algo_bytes / (iso_k1 * 1e-3) / 1e9
        iso = {"ms_per_launch": iso_k1, "achieved": iso_ach, "frac": iso_ach / peak, "launches": iso_calls,
               "stage_ms_per_step": {"k1_rrc": iso_ms[0] / iso_calls, "k2_demod": iso_ms[1] / iso_calls,
                                     "k3_k4_dmr": iso_ms[2] / iso_calls},
               "note": "same kernels, same inputs, three kernels back to back on one stream right after the timed "
                       "region: K1 with the SMs to itself.  In the timed region K1 of step i+1 shares the SMs with K2/K3 "
                       "of step i, which stretches its launch duration while shortening the step"}
    kernel_name = {"dmr": "rrc_fir_kernel<80> (K1)", "ysf": "rrc_fir_kernel<80> (K1)", "nxdn": "rrc_fir_kernel<160> (K1)",
                   "dstar": "demod_kernel<10,10> (K2)", "pocsag": "demod_kernel<10,40> (K2)"}[args.workload]
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "not_overlapped": iso,
                "algorithmic_bytes_per_sample": algo_bytes,
                "ms_per_launch": k1_ms * prof_steps / max(1, calls), "k1_ms_per_step": k1_ms,
                "stage_ms_per_step": {"k1_rrc": stage_ms[0] / prof_steps, "k2_demod": stage_ms[1] / prof_steps,
                                      "k3_k4_dmr": stage_ms[2] / prof_steps,
                                      "launches_per_stage_per_step": calls / prof_steps, "profiled_steps": prof_steps,
                                      "note": "per-kernel CUDA events on the launching streams, taken in a second pass "
                                              "of the same pipelined schedule right after the timed region (the "
                                              "timed region itself records no per-stage events); K1 of step i+1 runs "
                                              "beside K2/K3 of step i, so the sum exceeds ms_per_step"},
                "fp32_issue_bound": {"note": "K1 executes 162 separately rounded fp32 ops/sample (no FMA, bit-exact); "
                                             "the binding ceiling is FP32 issue, not HBM (SURVEY.md D9)",
                                     "fp32_ops_per_s": C * L * (322 if args.workload == "nxdn" else 162) / (k1_ms * 1e-3)
                                     if k1_ms > 0 and dom == 0 else None,
                                     "peak_fp32_lane_ops_per_s_at_max_clock": 148 * 128 * 1.965e9,
                                     "frac_in_region": (C * L * (322 if args.workload == "nxdn" else 162) / (k1_ms * 1e-3)
                                                        / (148 * 128 * 1.965e9)) if k1_ms > 0 and dom == 0 else None,
                                     "frac_not_overlapped": (C * L * (322 if args.workload == "nxdn" else 162)
                                                             / (iso["ms_per_launch"] * 1e-3) / (148 * 128 * 1.965e9))
                                     if iso and dom == 0 else None},
                "pipe_bytes_per_sample": wl["pipe_bytes"],
                "pipe_frac_of_hbm": value * 1e6 / world * wl["pipe_bytes"] / 1e9 / peak}

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            import oracle_lib
            orc = oracle_lib.best()
            probe = x[:cores, :L].cpu().numpy()
            t0 = time.perf_counter()
            orc.pipe_batch(orc_proto, probe, threads=cores, chunk=4096)
            per_ch = (time.perf_counter() - t0) / 1.0            # seconds for `cores` channels in parallel
            nch = int(max(cores, min(C, cores * max(1.0, args.cpu_seconds / max(per_ch, 1e-3)))))
            xs = x[:nch, :L].cpu().numpy()
            t0 = time.perf_counter()
            _, outs, metas = orc.pipe_batch(orc_proto, xs, threads=cores, chunk=4096, meta_cap=1 << 15)
            dt = time.perf_counter() - t0
            cpu = {"value": nch * L / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": orc.kind,
                   "sample": "first %d of %d channels x %d samples (one step's batch), %.1f s" % (nch, C, L, dt)}
            # the same sample doubles as an in-run parity check of the GPU path
            chk = dh.Pipe(nch, dh_proto, max_chunk=L, device=dev)
            chk.process(x[:nch], n=L)
            chk.collect()
            ok = all(chk.output(c) == outs[c].tobytes() and chk.meta(c) == metas[c] for c in range(nch))
            cpu["gpu_matches_cpu_on_sample"] = bool(ok)
            chk.close()
        except Exception as ex:   # the checker is optional for the measurement itself
            cpu = {"value": None, "unit": "Msamples/s", "cores": cores, "kind": "unavailable", "sample": str(ex)}

    line = {"metric": metric, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "arm": arm}
    if e2e_s16:
        line["e2e_s16"] = e2e_s16
    if nccl:
        line["nccl"] = nccl
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
