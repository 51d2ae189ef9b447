#!/usr/bin/env python
"""bench.py -- throughput of the encrypted-imputation cloud evaluation (cloud_compute_score) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload at N = 1 (BASELINE.json configs[1]): iDASH-scale synthetic, 1004 samples x 16184 tag SNPs x 80882 target
SNPs, neighbors = 5: 48 552 input TRLWE ciphertexts -> 242 646 output ciphertexts per batch of 1004
samples. One step = one pass of cloud_compute_score over the batch(es). With N > 1 GPUs the job is BASELINE
configs[4] as written: N batches of 1004 samples (8032 samples at N = 8), neighbors = 20, sharded by contiguous
target-SNP range: rank r evaluates targets [G r/N, G (r+1)/N) of every batch from the tag-ciphertext slab its
band touches; no collective is on the data path (weak scaling: per-GPU work is constant).

PARITY GATE: before a value is printed, the reference's own cloud_compute_score (oracle/_ref)
is run on the SAME ciphertext words the GPU was given and every output word and variance is compared; the line
carries "parity": {"checked_words": N, "equal": true} and the run fails without printing a value otherwise.

metric = samples x targets x 3 per second (packed imputed slots/s); `out_ct_per_s` is the same in output
ciphertexts. `value` is timed with the inputs resident in HBM; `e2e` goes through the host-buffer C-ABI
call (H2D + kernels + D2H inside the timed region). The reference arm (--impl reference) and the
`cpu_baseline` object time the reference's own cloud_compute_score (oracle/_ref, compiled from the
unmodified reference) on this box's host cores.

Further records of the line: `sustained` (seconds of back-to-back steps with their own clock / power samples),
`decrypt` (the decrypt stage on the step's own outputs: `roofline` against HBM plus `roofline.tensor` against the
int8 tensor rate, `e2e` through the host-buffer call, `parity` against the exact integer phase, the reference's
decrypt_predictions as `cpu_baseline`; at N > 1 also `sharded`: every rank decrypts its own rows), `nvlink` (N > 1:
the scatter of the input slabs from rank 0 and the gather of the rows back, and sharded == unsharded on the GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "imputed target-SNP ciphertext slots/sec (samples x targets x 3)"
UNIT = "slots/s"
S, T, G, NEIGHBORS = 1004, 16184, 80882, 5
SEED = 1234
CT_BYTES = 8192


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--neighbors", type=int, default=0, help="default: 5 at --gpus 1 (configs[1]), 20 at --gpus > 1 (configs[4])")
    ap.add_argument("--targets", type=int, default=G, help="(debug) fewer target SNPs")
    ap.add_argument("--tags", type=int, default=T)
    ap.add_argument("--samples", type=int, default=S)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the extra timing runs of the CPU baseline (the parity run still happens)")
    ap.add_argument("--no-parity", action="store_true", help="(debug / profiler runs) skip the parity gate; the line says so")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back steps for the `sustained` record (0 = skip)")
    ap.add_argument("--no-decrypt", action="store_true", help="skip the secondary decrypt-kernel timing")
    ap.add_argument("--batches", type=int, default=0, help="(debug) batches per step; default = number of GPUs")
    ap.add_argument("--no-batched", action="store_true", help="N > 1: one launch per batch instead of one batched launch per step")
    ap.add_argument("--kernel", default="auto", choices=["auto", "imad", "tensor", "tile", "ring"],
                    help="cloud kernel: auto (tensor-core when the model is eligible), imad, tensor (ring if eligible, "
                         "else tile), tile (one CTA per tile), ring (persistent)")
    a = ap.parse_args()
    if a.neighbors <= 0:
        a.neighbors = NEIGHBORS if a.gpus == 1 else 20
    return a


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.dev = device_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.dev)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            r = [x.strip() for x in r]
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for nme, v in zip(names, r[5:9]):
                if v.lower() == "active":
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(local_rank: int):
    """One process per GPU: run on the CPU cores NVML reports as local to that GPU, so that the pinned staging buffers of the
    end-to-end leg are first touched on the GPU's own NUMA node (a buffer on the other socket makes every PCIe copy cross
    the inter-socket link). Best effort: returns the core count bound to, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cores = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cores &= os.sched_getaffinity(0)
        if cores:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception:
        pass
    return None


def measured_peak_gbs():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def measured_peaks() -> dict:
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        return {}


def ncu_traffic(kernel: str, args, world: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    of this same workload (profiles/traffic.json, written from the .ncu-rep by tools/ncu_summary.py), or None."""
    if world != 1 or (args.samples, args.tags, args.targets) != (S, T, G):
        return None
    try:
        t = json.loads((ROOT / "profiles" / "traffic.json").read_text())
        return t[kernel][str(args.neighbors)]["dram_bytes_per_launch"]
    except Exception:
        return None


def workload_name(args, world=1, n_batches=1):
    base = f"iDASH-scale synthetic {args.samples} samples x {args.tags} tag x {args.targets} target SNPs, neighbors={args.neighbors}"
    if world > 1:
        return (f"{base} x {n_batches} sample batches = {n_batches * args.samples} samples (BASELINE configs[4]), sharded by "
                f"target range over {world} GPUs")
    return base + " (BASELINE configs[1])"


def reference_full(args, sub, cts, var, NR, RS, threads, want_output):
    """The reference's cloud_compute_score (oracle/_ref; the C restatement where it is absent) on a whole (sub-)model.
    -> (out_ct, out_var, seconds, kind)"""
    from oracle import pyoracle as po
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    idx = np.arange(len(cts), dtype=np.uint32)
    if po.have_ref():
        o, v, t = po.cloud_ref(args.samples, NR, RS, idx, cts, var, sub.out_bidx, sub.row_ptr, sub.col, sub.coef, want_output=want_output)
        return o, v, t, "reference"
    t0 = time.perf_counter()
    o, v = po.cloud_port(args.samples, NR, RS, idx, cts, var, sub.row_ptr, sub.col, sub.coef, threads=threads)
    return o, v, time.perf_counter() - t0, "port"


def parity_gate(args, api, ctx, m, sub, x0, NR, RS, world):
    """Every output word + variance + index of one evaluation on this rank's inputs x0 (torch CUDA tensor) against the reference on
    the same words. -> (parity dict, reference seconds, kind, cores)"""
    import torch
    from oracle import pyoracle as po
    n_rows, slab = sub.n_out, x0.shape[0]
    out = torch.zeros((n_rows, 2048), dtype=torch.int32, device="cuda")
    ovar = torch.zeros(n_rows, dtype=torch.float64, device="cuda")
    oidx = torch.zeros(n_rows, dtype=torch.int32, device="cuda")
    xv = torch.full((slab,), 2.0 ** -50, dtype=torch.float64, device="cuda")
    api.cloud_compute_score_device(ctx, m, x0, out, in_var=xv, out_index=oidx, out_var=ovar)
    torch.cuda.synchronize()
    ctx.check_device_status()
    cts = x0.cpu().numpy().view(np.uint32)
    threads = max(1, po.host_threads() // world)
    ref_out, ref_var, secs, kind = reference_full(args, sub, cts, np.full(slab, 2.0 ** -50), NR, RS, threads, True)
    bad = 0
    for lo in range(0, n_rows, 16384):
        r = torch.from_numpy(ref_out[lo:lo + 16384].view(np.int32)).cuda()
        bad += int((out[lo:lo + 16384] != r).any(dim=1).sum())
    var_ok = bool(np.array_equal(ovar.cpu().numpy(), ref_var))
    idx_ok = bool(np.array_equal(oidx.cpu().numpy().view(np.uint32), sub.out_bidx))
    par = {"checked_words": int(n_rows) * 2048, "checked_variances": int(n_rows), "differing_ciphertexts": bad,
           "equal": bool(bad == 0 and var_ok and idx_ok), "against": kind,
           "how": "reference cloud_compute_score on the same input ciphertext words, all output words + variances + indices compared"}
    del out, ref_out
    return par, secs, kind, threads


def build_workload(args):
    from idash2019_2_b200 import synth
    tag, tgt = synth.make_positions(args.tags, args.targets, SEED)
    model = synth.make_model(tag, tgt, args.neighbors, SEED)
    return model


# ---------------------------------------------------------------------------------------------------
def reference_sample_runner(args, model):
    """Returns (fn(sample_targets) -> seconds of the reference cloud_compute_score, kind, cores)."""
    from idash2019_2_b200 import synth
    from oracle import pyoracle as po
    cores = po.host_threads()
    os.environ["OMP_NUM_THREADS"] = str(cores)
    kind = "reference" if po.have_ref() else "port"
    NRr, RSr = 1024 // args.samples, 1024 // (1024 // args.samples)
    cache = {}

    def run(sample_targets: int) -> float:
        if sample_targets not in cache:
            sub = model.rows(0, 3 * sample_targets)
            real = sub.col != 0xFFFFFFFF
            n_in = int(sub.col[real].max()) // NRr + 1
            cache.clear()
            cache[sample_targets] = (sub, synth.random_ciphertexts(n_in, SEED), np.full(n_in, 2.0 ** -50),
                                     np.arange(n_in, dtype=np.uint32))
        sub, cts, var, idx = cache[sample_targets]
        if kind == "reference":
            return po.cloud_ref(args.samples, NRr, RSr, idx, cts, var, sub.out_bidx, sub.row_ptr, sub.col, sub.coef,
                                want_output=False)[2]
        t0 = time.perf_counter()
        po.cloud_port(args.samples, NRr, RSr, idx, cts, var, sub.row_ptr, sub.col, sub.coef, threads=cores)
        return time.perf_counter() - t0

    return run, kind, cores


def pick_sample(run, n_targets, budget_s=2.5):
    """Bounded sample of the workload: as many leading target SNPs as the reference evaluates in ~budget_s."""
    probe = min(n_targets, 1500)
    t = run(probe)
    if probe == n_targets:
        return probe
    est = int(probe * budget_s / max(t, 1e-3))
    return max(probe, min(n_targets, est))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = build_workload(args)
    run, kind, cores = reference_sample_runner(args, model)
    sample = pick_sample(run, args.targets)
    for _ in range(args.warmup):
        run(sample)
    times = [run(sample) for _ in range(args.steps)]
    total = sum(times)
    value = args.samples * sample * 3 * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "out_ct_per_s": 3 * sample * args.steps / total,
        "config": {"workload": workload_name(args, args.gpus, args.batches or args.gpus), "neighbors": args.neighbors, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"first {sample} of {args.targets} target SNPs ({3 * sample} output ciphertexts) of one 1004-sample batch "
                                   f"per step, reference cloud_compute_score 'fhe wall time' with {cores} OpenMP threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def decrypt_sharded_ms(args, api, ctx, outs, n_rows):
    """N > 1: every rank decrypts its own rows of every batch (the stage shards by ciphertext, nothing is exchanged). Milliseconds
    for all batches of this rank, CUDA events around the back-to-back launches, mean of 5."""
    import torch
    key = np.random.default_rng(5).integers(0, 2, 1024).astype(np.int32)
    scores = torch.empty((n_rows, args.samples), dtype=torch.float32, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for o in outs:
        api.decrypt_predictions_device(ctx, key, args.samples, o, scores)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        for o in outs:
            api.decrypt_predictions_device(ctx, key, args.samples, o, scores)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5


def bench_decrypt(args, api, ctx, out_ct, n_rows, peak):
    """The decrypt stage (decrypt_predictions, eval/idash.cpp:681-761) on the step's own output ciphertexts: kernel roofline,
    end to end through the host-buffer entry point, and the reference's decrypt_predictions on a bounded sample."""
    import torch
    key = np.random.default_rng(5).integers(0, 2, 1024).astype(np.int32)
    scores = torch.empty((n_rows, args.samples), dtype=torch.float32, device="cuda")
    for _ in range(2):
        api.decrypt_predictions_device(ctx, key, args.samples, out_ct, scores)
    torch.cuda.synchronize()
    ctx.timing_enable(10)
    for _ in range(10):
        api.decrypt_predictions_device(ctx, key, args.samples, out_ct, scores)
    torch.cuda.synchronize()
    d_ms = statistics.mean(ctx.timing_read(10))
    ctx.timing_enable(0)
    d_bytes = n_rows * (CT_BYTES + 4 * args.samples)
    kname = {api.DECRYPT_TENSOR_PAIR: "decrypt_pair_kernel", api.DECRYPT_TENSOR: "decrypt_tc_kernel"}.get(ctx.last_decrypt_kernel(), "decrypt_kernel")
    rec = {"metric": "decrypted output ciphertexts/sec", "value": n_rows / (d_ms * 1e-3), "unit": "ct/s", "ciphertexts": n_rows,
           "roofline": {"bound": "hbm", "achieved": d_bytes / (d_ms * 1e-3) * 1e-9, "peak": peak, "unit": "GB/s",
                        "frac": d_bytes / (d_ms * 1e-3) * 1e-9 / peak, "traffic": (json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(kname, {}).get(str(n_rows), {}).get("dram_bytes_per_launch") if (ROOT / "profiles" / "traffic.json").exists() else None), "kernel": kname, "kernel_ms": d_ms,
                        "algorithmic_bytes": d_bytes, "int8_mac_per_s": n_rows * 4 * 1024 * 1024 / (d_ms * 1e-3)}}
    # the other roofline of this kernel: the phase is a GEMM (1024 x 1024 Toeplitz matrix x 4 byte planes per ciphertext), and on CTA
    # pairs it runs the int8 tensor pipe at 97 % active cycles (profiles/r02_ncu_decrypt_pair.txt). Peak: int8 dense = 2 x the bf16
    # figure MEASURED_PEAKS.json holds (the profiling recipe's table: fp8 / int8 dense = 2 x bf16 dense)
    bf16 = measured_peaks().get("bf16_tflops")
    if bf16 and kname != "decrypt_kernel":
        tops = 2 * n_rows * 4 * 1024 * 1024 / (d_ms * 1e-3) * 1e-12
        rec["roofline"]["tensor"] = {"achieved": tops, "peak": 2 * bf16, "unit": "TOP/s (int8)", "frac": tops / (2 * bf16),
                                     "peak_source": "2 x measured bf16 TFLOP/s (MEASURED_PEAKS.json bf16_tflops)"}
    # end to end: pinned host ciphertexts in, pinned host scores out
    try:
        h_ct = torch.empty((n_rows, 2048), dtype=torch.int32, pin_memory=True)
        h_ct.copy_(out_ct)
        h_sc = torch.empty((n_rows, args.samples), dtype=torch.float32, pin_memory=True)
        np_ct, np_sc = h_ct.numpy().view(np.uint32), h_sc.numpy()
        api.decrypt_predictions(ctx, key, args.samples, np_ct, out_scores=np_sc)
        t0 = time.perf_counter()
        for _ in range(2):
            api.decrypt_predictions(ctx, key, args.samples, np_ct, out_scores=np_sc)
        t = (time.perf_counter() - t0) / 2
        rec["e2e"] = {"value": n_rows / t, "unit": "ct/s", "ms_per_step": t * 1e3, "h2d_bytes_per_step": n_rows * CT_BYTES,
                      "d2h_bytes_per_step": n_rows * 4 * args.samples,
                      "matches_device_path": bool(torch.equal(h_sc.cuda(), scores))}
        # CPU baseline + parity on a bounded sample: the reference's decrypt_predictions (FFT path: +-1 LSB of the phase) and the
        # exact integer phase of the oracle (bit-exact)
        from oracle import pyoracle as po
        n_s = min(n_rows, 3 * 8000)
        sample = np.ascontiguousarray(np_ct[:n_s])
        exact = po.decode_port(args.samples, po.phase_exact_port(key, sample))
        rec["parity"] = {"checked_scores": int(n_s) * args.samples, "equal": bool(np.array_equal(np_sc[:n_s], exact)),
                         "against": "oracle exact integer phase (TFHE's exact product), decoded as eval/idash.cpp:717-719"}
        if po.have_ref():
            ref_sc, secs = po.decrypt_ref(args.samples, key, sample)
            rec["cpu_baseline"] = {"value": n_s / secs, "unit": "ct/s", "cores": po.host_threads(), "kind": "reference",
                                   "sample": f"first {n_s} of {n_rows} output ciphertexts, reference decrypt_predictions ('decrypt wall time')",
                                   "max_abs_score_diff_vs_reference_fft": float(np.abs(ref_sc.astype(np.float64) - np_sc[:n_s]).max())}
    except Exception as e:  # secondary record: say why a part is missing
        rec["e2e_error"] = repr(e)
    return rec


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:   # convenience: relaunch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29533", __file__] + sys.argv[1:]
            os.execv(sys.executable, cmd)
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as ge
    if rank == 0:
        ge.build_native()
    if world > 1:
        dist.barrier()
    from idash2019_2_b200 import api

    model = build_workload(args)
    NR = 1024 // args.samples
    RS = 1024 // NR
    from idash2019_2_b200 import shard as shard_mod
    sh = shard_mod.make_shard(model, NR, args.targets, rank, world)     # contiguous target range + the input slab it reads
    sub, slab, t_lo, t_hi = sh.model, sh.n_ct, sh.target_lo, sh.target_hi
    n_rows = sub.n_out
    n_batches = args.batches or world
    ctx = api.Context(local_rank)
    ctx.set_kernel({"auto": api.KERNEL_AUTO, "imad": api.KERNEL_IMAD, "tensor": api.KERNEL_TENSOR,
                    "tile": api.KERNEL_TENSOR_TILE, "ring": api.KERNEL_TENSOR_RING}[args.kernel])
    t0 = time.perf_counter()
    m = api.Model(ctx, args.samples, NR, RS, sub.out_bidx, sub.row_ptr, sub.col, sub.coef)
    m.free()
    t0 = time.perf_counter()
    m = api.Model(ctx, args.samples, NR, RS, sub.out_bidx, sub.row_ptr, sub.col, sub.coef)     # warm: allocator and page cache primed
    model_upload_s = time.perf_counter() - t0
    gen = torch.Generator(device="cuda").manual_seed(SEED + rank)
    ins = [torch.randint(-2 ** 31, 2 ** 31, (slab, 2048), dtype=torch.int32, device="cuda", generator=gen)
           for _ in range(n_batches)]
    outs = [torch.empty((n_rows, 2048), dtype=torch.int32, device="cuda") for _ in range(n_batches)]

    # ---- N > 1: the data plane of the sharded job. The job's ciphertexts are resident on GPU 0 -- all tag ciphertexts of
    # every batch -- and its results are wanted there: rank r RECEIVES the slab its target range reads, of every batch (scatter), and
    # SENDS BACK its rows of every batch's output (gather), as grouped NCCL send / recv over NVLink / NVSwitch. There is no collective
    # inside the evaluation itself. The timed steps below consume the scattered slabs, and the gathered result of batch 0 is compared
    # on GPU 0 with the UNSHARDED evaluation of the whole model on the same inputs (sharded == unsharded, bit for bit).
    nvlink = None
    if world > 1:
        real = model.col != np.uint32(0xFFFFFFFF)
        n_in_total = int(model.col[real].max()) // NR + 1
        metas = [shard_mod.make_shard(model, NR, args.targets, r, world) for r in range(world)]
        meta = [(s_.ct_min, s_.n_ct, s_.row_lo, s_.row_hi) for s_ in metas]
        del metas
        full_ins = full_outs = None
        if rank == 0:
            g0 = torch.Generator(device="cuda").manual_seed(SEED + 1000)
            full_ins = [torch.randint(-2 ** 31, 2 ** 31, (n_in_total, 2048), dtype=torch.int32, device="cuda", generator=g0) for _ in range(n_batches)]
            full_outs = [torch.empty((3 * args.targets, 2048), dtype=torch.int32, device="cuda") for _ in range(n_batches)]

        def scatter():
            ops = []
            if rank == 0:
                for b in range(n_batches):
                    ins[b].copy_(full_ins[b][meta[0][0]:meta[0][0] + meta[0][1]])
                    for r in range(1, world):
                        ops.append(dist.P2POp(dist.isend, full_ins[b][meta[r][0]:meta[r][0] + meta[r][1]], r))
            else:
                for b in range(n_batches):
                    ops.append(dist.P2POp(dist.irecv, ins[b], 0))
            for q in dist.batch_isend_irecv(ops):
                q.wait()

        def gather():
            ops = []
            if rank == 0:
                for b in range(n_batches):
                    full_outs[b][meta[0][2]:meta[0][3]].copy_(outs[b])
                    for r in range(1, world):
                        ops.append(dist.P2POp(dist.irecv, full_outs[b][meta[r][2]:meta[r][3]], r))
            else:
                for b in range(n_batches):
                    ops.append(dist.P2POp(dist.isend, outs[b], 0))
            for q in dist.batch_isend_irecv(ops):
                q.wait()

        def timed(fn, reps=3):
            fn()                                   # warm-up: NCCL sets up its peer channels on first use
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            tt = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt[0])

        scatter_ms = timed(scatter)
        if n_batches > 1 and not args.no_batched:
            api.cloud_compute_score_device_batched(ctx, m, ins, outs)
        else:
            for b in range(n_batches):
                api.cloud_compute_score_device(ctx, m, ins[b], outs[b])
        torch.cuda.synchronize()
        gather_ms = timed(gather)
        same = None
        if rank == 0:
            mf = api.Model(ctx, args.samples, NR, RS, model.out_bidx, model.row_ptr, model.col, model.coef)
            whole = torch.empty((3 * args.targets, 2048), dtype=torch.int32, device="cuda")
            api.cloud_compute_score_device(ctx, mf, full_ins[0], whole)
            torch.cuda.synchronize()
            same = bool(torch.equal(whole, full_outs[0]))
            mf.free()
            del whole
        sc_bytes = n_batches * sum(meta[r][1] for r in range(1, world)) * CT_BYTES
        ga_bytes = n_batches * sum(meta[r][3] - meta[r][2] for r in range(1, world)) * CT_BYTES
        nvlink = {"scatter_ms": scatter_ms, "gather_ms": gather_ms, "scatter_bytes": sc_bytes, "gather_bytes": ga_bytes,
                  "scatter_gbs": sc_bytes / (scatter_ms * 1e-3) * 1e-9, "gather_gbs": ga_bytes / (gather_ms * 1e-3) * 1e-9,
                  "sharded_equals_unsharded": same,
                  "how": "job data resident on GPU 0: every rank receives its slab of every batch and returns its rows of every batch, grouped "
                         "ncclSend / ncclRecv (torch.distributed.batch_isend_irecv), CUDA events, max over ranks; the timed steps consume the "
                         "scattered slabs; GPU 0 compares the gathered batch 0 with its own unsharded evaluation of the whole model"}
        del full_ins, full_outs
        torch.cuda.empty_cache()

    # The batches of a step are independent evaluations of the same model on the rank's target range. With more than one they
    # go through the batched entry point: ONE launch of the persistent kernel walks the tiles of every batch, so the kernel's
    # ramp-up and tail are paid once per step and not once per batch (--no-batched: one launch per batch, round-robin on two
    # side streams, the round's earlier scheme).
    batched = n_batches > 1 and not args.no_batched
    side = [torch.cuda.Stream() for _ in range(2)] if n_batches > 1 and not batched else []
    launches_per_step = 1 if (batched or n_batches == 1) else n_batches

    def step():
        if batched:
            api.cloud_compute_score_device_batched(ctx, m, ins, outs)
            return
        if not side:
            api.cloud_compute_score_device(ctx, m, ins[0], outs[0])
            return
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        for s_ in side:
            s_.wait_event(fork)
        for b in range(n_batches):
            api.cloud_compute_score_device(ctx, m, ins[b], outs[b], stream=side[b % 2].cuda_stream)
        for s_ in side:
            join = torch.cuda.Event()
            join.record(s_)
            main.wait_event(join)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    ctx.check_device_status()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.kernel_launches()
    ctx.timing_enable(args.steps * launches_per_step)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.start()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = ctx.timing_read(args.steps * launches_per_step)
    ctx.timing_enable(0)
    launches = ctx.kernel_launches() - launches0
    kernel_used = ctx.last_kernel()

    # ---- sustained: seconds of back-to-back steps with their own clock samples (does the burst figure survive thermally?)
    sustained = None
    if args.sustain > 0:
        n_sus = max(args.steps, int(args.sustain * 1e3 / max(ms_total / args.steps, 1e-3)))
        sampler2 = ClockSampler(local_rank) if rank == 0 else None
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if sampler2:
            sampler2.start()
        es0.record()
        for _ in range(n_sus):
            step()
        es1.record()
        barrier()
        sus_ms = es0.elapsed_time(es1)
        sustained = {"steps": n_sus, "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / n_sus,
                     "clocks": sampler2.stop() if sampler2 else None}

    # ---- end to end through the host-buffer entry point (pinned host memory, H2D + D2H per step)
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    h_in = torch.empty((slab, 2048), dtype=torch.int32, pin_memory=True)
    h_in.copy_(ins[0])
    h_out = torch.empty((n_rows, 2048), dtype=torch.int32, pin_memory=True)
    np_in, np_out = h_in.numpy().view(np.uint32), h_out.numpy().view(np.uint32)

    def e2e_step():
        for _b in range(n_batches):
            api.cloud_compute_score(ctx, m, np_in, out_ct=np_out)

    e2e_step()                      # warm-up: sizes the library's device staging buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    # what the host link alone allows: the same device->host bytes as a bare pinned copy, all ranks at the same time
    eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eb0.record()
    for _b in range(n_batches):
        h_out.copy_(outs[0], non_blocking=True)
    eb1.record()
    torch.cuda.synchronize()
    bare_d2h_ms = eb0.elapsed_time(eb1)
    barrier()
    clocks = sampler.stop() if sampler else None
    h_out_host_path = None
    if True:
        api.cloud_compute_score(ctx, m, np_in, out_ct=np_out)
        h_out_host_path = torch.equal(h_out.cuda(), outs[0])   # host path and device path agree on the same input

    # ---- parity gate: the reference on the same input words (every rank checks its own shard, batch 0)
    parity, ref_secs, ref_kind, ref_cores = None, None, None, None
    if not args.no_parity:
        parity, ref_secs, ref_kind, ref_cores = parity_gate(args, api, ctx, m, sub, ins[0], NR, RS, world)

    # ---- the decrypt stage on the step's own output ciphertexts (secondary record beside the headline)
    decrypt, dec_ms = None, 0.0
    if world > 1 and not args.no_decrypt:
        dec_ms = decrypt_sharded_ms(args, api, ctx, outs, n_rows)
    if rank == 0 and not args.no_decrypt:
        decrypt = bench_decrypt(args, api, ctx, outs[0], n_rows, measured_peak_gbs()[0])

    par_ok = 1.0 if (parity is None or parity["equal"]) else 0.0
    par_words = float(parity["checked_words"]) if parity else 0.0
    t = torch.tensor([ms_total, t_e2e * 1e3, bare_d2h_ms, sustained["ms_per_step"] if sustained else 0.0, dec_ms], dtype=torch.float64, device="cuda")
    pt = torch.tensor([par_ok, 1.0 if h_out_host_path else 0.0], dtype=torch.float64, device="cuda")
    ps = torch.tensor([par_words], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(pt, op=dist.ReduceOp.MIN)
        dist.all_reduce(ps, op=dist.ReduceOp.SUM)
    ms_total, ms_e2e, bare_d2h_ms, sus_ms_step, dec_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3]), float(t[4])
    all_equal, host_same = bool(pt[0] > 0.5), bool(pt[1] > 0.5)

    rc = 0
    if rank == 0:
        slots_per_step = args.samples * args.targets * 3 * n_batches
        ct_per_step = args.targets * 3 * n_batches
        value = slots_per_step * args.steps / (ms_total * 1e-3)
        k_ms = statistics.mean(kernel_ms) if kernel_ms else float("nan")
        alg_bytes = CT_BYTES * (slab + n_rows) * (n_batches // launches_per_step)   # per launch: every input ct read once, every output written once
        peak, peak_src = measured_peak_gbs()
        achieved = alg_bytes / (k_ms * 1e-3) * 1e-9
        how = "algorithmic bytes of one launch / its mean duration (CUDA events on the launch stream)"
        if side:
            # the launches of a step overlap on two streams, so a launch's own duration includes the time it shared the SMs:
            # the GPU's HBM rate is the bytes of all launches of the step over the step time
            achieved = launches_per_step * alg_bytes / (ms_total / args.steps * 1e-3) * 1e-9
            how = f"{n_batches} launches per step overlap on 2 streams: algorithmic bytes of the step's launches / step time"
        h2d = n_batches * (slab * CT_BYTES)
        d2h = n_batches * (n_rows * (CT_BYTES + 4 + 8))
        kernel_name = {api.KERNEL_IMAD: "cloud_eval_kernel", api.KERNEL_TENSOR_TILE: "cloud_tc_kernel",
                       api.KERNEL_TENSOR_RING: "cloud_ring_kernel"}.get(kernel_used, "?")
        e2e_value = slots_per_step * e2e_steps / (ms_e2e * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "out_ct_per_s": ct_per_step * args.steps / (ms_total * 1e-3),
            "config": {"workload": workload_name(args, world, n_batches),
                       "neighbors": args.neighbors, "batches": n_batches, "targets_per_gpu": t_hi - t_lo,
                       "in_ct_per_gpu_batch": slab, "out_ct_per_gpu_batch": n_rows,
                       "l2": "inputs+outputs per step (2.4 GB) exceed the 126 MB L2; no explicit flush",
                       "streams": ("batches of a step round-robin on 2 side streams" if side else
                                   f"single stream, one batched launch of {n_batches} batches per step" if batched else "single stream"),
                       "numa_local_cores": numa},
            "parity": (dict(parity, checked_words=int(ps[0]), equal=all_equal, ranks=world) if parity else
                       {"skipped": "--no-parity (debug / profiler run): this line is not a benchmark result"}),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(kernel_name, args, world),
                         "kernel": kernel_name, "kernel_ms": k_ms, "algorithmic_bytes": alg_bytes,
                         "peak_source": peak_src, "kernel_share_of_step": k_ms * launches_per_step / (ms_total / args.steps), "how": how},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "matches_device_path": host_same,
                    "excludes": "one-time costs outside the per-step call: model compile + upload (model_upload_ms below, once per model), "
                                "CUDA context creation, page-locking of the host buffers",
                    "model_upload_ms": model_upload_s * 1e3,
                    "value_with_model_upload": slots_per_step / (ms_e2e / e2e_steps * 1e-3 + model_upload_s),
                    "d2h_gbs_per_gpu": d2h / (ms_e2e / e2e_steps * 1e-3) * 1e-9,
                    "bare_d2h_ms_per_step": bare_d2h_ms,
                    "bare_d2h_gbs_per_gpu": d2h / (bare_d2h_ms * 1e-3) * 1e-9,
                    "limiter": "device->host copy of the output ciphertexts over the GPU's host link: a bare pinned copy of the same bytes "
                               "(all ranks at once) takes bare_d2h_ms_per_step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if sustained:
            sustained["ms_per_step"] = sus_ms_step
            sustained["value"] = slots_per_step / (sus_ms_step * 1e-3)
            sustained["frac"] = (CT_BYTES * (slab + n_rows) * n_batches / (sus_ms_step * 1e-3) * 1e-9) / peak
            sustained["vs_burst"] = (ms_total / args.steps) / sus_ms_step
            line["sustained"] = sustained
        if decrypt:
            if world > 1 and dec_ms > 0:
                # the whole job's decrypt: every rank its own rows of every batch, max over ranks
                n_all = n_rows * n_batches * world
                decrypt["sharded"] = {"n_gpus": world, "ciphertexts": n_all, "ms": dec_ms, "value": n_all / (dec_ms * 1e-3), "unit": "ct/s",
                                      "how": "every rank decrypts its rows of every batch (no exchange), CUDA events, max over ranks",
                                      "hbm_frac_per_gpu": n_rows * n_batches * (CT_BYTES + 4 * args.samples) / (dec_ms * 1e-3) * 1e-9 / measured_peak_gbs()[0]}
            line["decrypt"] = decrypt
        if nvlink:
            nvlink["job_ms_gpu0_resident"] = nvlink["scatter_ms"] + ms_total / args.steps + nvlink["gather_ms"]
            line["nvlink"] = nvlink
            if nvlink["sharded_equals_unsharded"] is False:
                all_equal = False
        if world == 1 and parity is not None:
            try:
                best = ref_secs
                if not args.no_cpu_baseline:
                    cts0 = ins[0].cpu().numpy().view(np.uint32)
                    v0 = np.full(slab, 2.0 ** -50)
                    for _ in range(2):
                        best = min(best, reference_full(args, sub, cts0, v0, NR, RS, ref_cores, False)[2])
                line["cpu_baseline"] = {"value": args.samples * args.targets * 3 / best, "unit": UNIT, "cores": ref_cores, "kind": ref_kind,
                                        "sample": f"the whole workload: all {args.targets} target SNPs ({n_rows} output ciphertexts) of one "
                                                  f"{args.samples}-sample batch, best of {1 if args.no_cpu_baseline else 3} runs, cloud_compute_score "
                                                  f"only ('fhe wall time'); the first run is the parity gate's",
                                        "seconds": best}
            except Exception as e:  # the checker is optional for the bench line; say why it is missing
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
        if (parity is not None and not all_equal) or (nvlink and nvlink["sharded_equals_unsharded"] is False):
            # parity is a gate -- no value without it
            line["value"] = None
            line["e2e"]["value"] = None
            line["error"] = "parity gate failed: the CUDA path's output differs from the reference's on the same inputs"
            rc = 1
        print(json.dumps(line), flush=True)
    m.free()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rc:
        raise SystemExit(rc)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
