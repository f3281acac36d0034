#!/usr/bin/env python
"""bench.py -- headline benchmark of the LivingScenes hot path on B200 (see BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

metric   instances/sec (N=1024 points) for encode + match + pose
step     one pass over one batch of synthetic input PER GPU: 4 scene pairs x (32 ref + 32 rescan
         instances) = 256 instances of 1024 points (BASELINE config[1] batch, C3-style pairs):
         Shape_Prior.encode of all 256 -> [N>1: one NCCL all-gather of the packed 1028-float codes] ->
         sequential_matcher per pair -> Kabsch SE(3) per matched ref instance.
value    whole-job instances/s with inputs resident in HBM (CUDA events around every step, L2 flushed
         between steps, max over ranks).
e2e      the same through the public API from PINNED HOST buffers: H2D of the clouds and D2H of
         matches / R / t inside the timed region.
roofline the fused kNN+EdgeConv launch with the largest share, timed live with CUDA events on the
         launching stream (ls_profile_* hooks of the C ABI), against MEASURED_PEAKS.json.
--impl reference  times the CPU restatement of the reference path (oracle/, kind "port"; the Python
         reference itself cannot travel to the GPU box) on the host cores, same metric / config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_POINTS = 1024
PAIRS_PER_GPU = 4
INST_PER_SET = 32
INST_PER_GPU = PAIRS_PER_GPU * 2 * INST_PER_SET  # 256
N_INPUT_SETS = 4
SHIPPED = os.path.join(ROOT, "livingscenes_b200", "_weights", "shipped_fp32.pt")
FEAT = [32, 32, 64, 64, 128, 256, 512]
DOWN = {2: 2, 4: 4, 5: 4}


def load_state_dict():
    from livingscenes_b200 import synthetic as S

    if os.path.exists(SHIPPED):
        return torch.load(SHIPPED, map_location="cpu", weights_only=True), "shipped checkpoint weights"
    return S.random_state_dict(0), "seeded random weights (shipped checkpoint not present)"


def make_scene_batch(n_pairs: int, seed: int, n_set: int = INST_PER_SET):
    """[n_pairs*2*n_set, 3, 1024]: per pair n_set ref instances (asymmetric ``synth_parts`` objects) then n_set rescan
    instances = permuted, rotated, translated exact rigid copies of the ref ones (SURVEY.md 8d, config C3 at
    N=1024).  Also returns the planted truth per pair: the expected matches0 (inverse permutation) and the SE(3)
    of every ref instance, which the step must recover (``check_planted``)."""
    from livingscenes_b200 import synthetic as R

    g = torch.Generator().manual_seed(seed)
    out, truth = [], []
    for p in range(n_pairs):
        ref = R.synth_parts(n_set, N_POINTS, seed * 1000 + p)
        perm = torch.randperm(n_set, generator=g)
        Rg = R.random_rotations(n_set, seed * 1000 + 500 + p)
        tg = torch.randn(n_set, 3, 1, generator=g)
        res = Rg @ ref[perm] + tg
        out += [ref, res]
        inv = torch.argsort(perm)
        truth.append({"matches0": inv, "R": Rg[inv], "t": tg[inv]})
    return torch.cat(out, 0).contiguous(), truth


def check_planted(m0, R, t, truth):
    """matches0 [P*n], R [P*n,3,3], t [P*n,3,1] of one step against the planted permutation / poses."""
    import math

    gm = torch.cat([tr["matches0"] for tr in truth])
    gR = torch.cat([tr["R"] for tr in truth])
    gt = torch.cat([tr["t"] for tr in truth])
    m0, R, t = m0.cpu(), R.cpu().float(), t.cpu().float()
    cos = ((torch.einsum("bij,bij->b", R, gR) - 1) / 2).clamp(-1, 1)
    rre = torch.acos(cos) * 180 / math.pi
    rte = (t - gt).norm(dim=1).reshape(-1)
    return {"match_recall": float((m0 == gm).float().mean()), "rre_deg_median": float(rre.median()),
            "rre_deg_max": float(rre.max()), "rte_median": float(rte.median())}


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- roofline
def layer_geometry(N: int):
    geo, n = [], N
    for i, co in enumerate(FEAT):
        ns = n
        n //= DOWN.get(i, 1)
        geo.append(dict(layer=i, c_in=1 if i == 0 else FEAT[i - 1], c_out=co, n_src=ns, n_dst=n))
    return geo


def knn_edge_algorithmic(g: dict):
    """Per instance, SURVEY.md 8d: bytes = 4*C_in*3*N_src + 4*C_out*3*N_dst + 8*N_dst*K;
    FLOPs = 3*N_dst*N_src*3*C_in (kNN) + the layer's reference-form VN FLOPs (edge-level V[,K] branches:
    2*(2C_in)*C_out*3*N_dst*K for lin + 2*C_out^2*3*N_dst*K for lin_dir, per branch)."""
    ci, co, ns, nd = g["c_in"], g["c_out"], g["n_src"], g["n_dst"]
    byt = 4 * ci * 3 * ns + 4 * co * 3 * nd + 8 * nd * 16
    knn = 3 * nd * ns * 3 * ci
    cg = 3 if g["layer"] == 0 else 2 * ci
    branches = 1 if g["layer"] < 2 else 2
    vn = branches * (2 * cg * co + 2 * co * co) * 3 * nd * 16
    return byt, knn + vn


def measured_peaks():
    """(HBM GB/s, source, SM MHz, TF32 dense TFLOP/s proxy = measured bf16 / 2, source)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return (float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0)),
                float(d["bf16_tflops"]) / 2, "measured bf16_tflops / 2 (TF32 runs at half the bf16 rate)")
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0, 1590.0 / 2, "fallback bf16 1590 / 2"


# --------------------------------------------------------------------------------------- CPU arm
def cpu_path_step(sd, x_pair, use_reference_modules):
    """encode + sequential match + Kabsch of one small scene pair on the host cores."""
    from oracle import restatement as R

    n = x_pair.shape[0] // 2
    with torch.no_grad():
        if use_reference_modules is not None:
            sp, mods = use_reference_modules
            ca, cb = sp.encode(x_pair[:n]), sp.encode(x_pair[n:])
            m = mods.matcher_new.sequential_matcher(ca["z_inv"], cb["z_inv"])["matches0"]
            mods.pose_estimation.kabsch_transformation_estimation(ca["z_so3"] + ca["t"], (cb["z_so3"] + cb["t"])[m])
        else:
            ca, cb = R.encode(sd, x_pair[:n]), R.encode(sd, x_pair[n:])
            m = R.sequential_match(ca["z_inv"], cb["z_inv"])["matches0"]
            R.kabsch(ca["z_so3"] + ca["t"], (cb["z_so3"] + cb["t"])[m])


def cpu_setup(sample_inst: int):
    from livingscenes_b200 import synthetic as S
    from oracle import p3d_shim, ref_loader  # the CPU-baseline leg is the one place bench.py executes oracle/

    p3d_shim.EXACT = False  # timing leg: cheapest honest fp32 kNN instead of the fp64 parity kNN
    torch.set_num_threads(os.cpu_count() or 1)
    sd, wdesc = load_state_dict()
    half = sample_inst // 2
    ref = S.synth_instances(half, N_POINTS, 777)
    res = S.random_rotations(half, 778) @ ref.flip(0) + 0.1
    x = torch.cat([ref, res], 0)
    mods, kind = None, "port"
    if ref_loader.available() and ref_loader.checkpoint_available():
        try:
            mods = (ref_loader.shape_prior(sd), ref_loader.load())
            kind = "reference"
        except Exception:
            mods, kind = None, "port"
    return sd, x, mods, kind, wdesc


WORKLOAD = ("BASELINE config[1]: 256 synthetic instances x 1024 points per GPU = 4 scene pairs x "
            "(32 ref + 32 rescan); Shape_Prior.encode + sequential_matcher + Kabsch per matched pair")


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (its own modules where /root/reference is
    mounted, else the oracle port) on all host threads.  One step = one bounded sample of the workload (a scene
    pair of 2+2 instances, ~3.6 s on 16 cores; 1+1 when K > 40) so that K steps end within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 60))
    sample = 4 if steps <= 40 else 2
    sd, x, mods, kind, wdesc = cpu_setup(sample)
    warm = max(1, min(args.warmup, 2))
    for _ in range(warm):
        cpu_path_step(sd, x, mods)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_path_step(sd, x, mods)
    dt = (time.perf_counter() - t0) / steps
    val = sample / dt
    cores = torch.get_num_threads()
    sample_desc = (f"{sample} instances of {N_POINTS} points per step (1 scene pair of {sample // 2}+{sample // 2}), "
                   f"{steps} timed steps after {warm} warm-up; the full 256-instance step is {256 // sample} such samples "
                   f"(CPU cost per instance is batch-independent: extrapolated {256 * dt / sample:.0f} s per full step)")
    print(json.dumps({
        "impl": "reference", "metric": "instances/sec (N=1024 pts) encode+match+pose", "value": val,
        "unit": "instances/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": f"synthetic ({wdesc})",
        "config": {"workload": WORKLOAD, "instances_per_gpu": INST_PER_GPU, "n_points": N_POINTS,
                   "pairs_per_gpu": PAIRS_PER_GPU, "sample": sample_desc,
                   "parallelism": f"host CPU, {cores} threads (rank 0 only)"},
        "cpu_baseline": {"value": val, "unit": "instances/s", "cores": cores, "kind": kind, "sample": sample_desc},
        "e2e": {"value": val, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------- B200 arm
def capture_graph(fn, barrier):
    """Warm ``fn`` on a side stream, then capture it into one CUDA graph; returns (graph, static outputs)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = fn()
    barrier()
    return graph, out


def gpu_eager_leg(sd, x_dev, sizes, chunk=16, timed_passes=2):
    """The reference's eager PyTorch form of the SAME 256-instance step on the same B200: oracle/restatement.py (the
    reference's edge-level fp32 op sequence) with device='cuda'.  pytorch3d is not installable
    here, so kNN = torch fp32 matmul form + topk on the device and FPS = this repo's CUDA FPS op standing in for
    pytorch3d's CUDA kernel (both favour the baseline).  The edge tensors of 256 instances (~0.76 GB each) do not fit:
    the batch is encoded in chunks of ``chunk`` instances, as the reference would have to."""
    from livingscenes_b200.ops import sample_farthest_points
    from oracle import p3d_shim
    from oracle import restatement as R

    dev = x_dev.device
    sdd = {k: v.to(dev) for k, v in sd.items() if k.startswith("encoder.")}
    old = (p3d_shim.ON_DEVICE, p3d_shim.FPS_IMPL, p3d_shim.EXACT)
    p3d_shim.ON_DEVICE, p3d_shim.FPS_IMPL = True, (lambda pts, K: sample_farthest_points(pts, K))
    n_set = sizes[0]

    def one_pass():
        with torch.no_grad():
            codes = [R.encode(sdd, x_dev[i:i + chunk]) for i in range(0, x_dev.shape[0], chunk)]
            z_inv = torch.cat([c["z_inv"] for c in codes])
            z_so3 = torch.cat([c["z_so3"] for c in codes])
            t = torch.cat([c["t"] for c in codes])
            for p in range(len(sizes)):
                a, b = 2 * p * n_set, (2 * p + 1) * n_set
                m = R.sequential_match(z_inv[a:a + n_set], z_inv[b:b + n_set])["matches0"]
                R.kabsch(z_so3[a:a + n_set] + t[a:a + n_set], (z_so3[b:b + n_set] + t[b:b + n_set])[m])

    try:
        with torch.no_grad():
            R.encode(sdd, x_dev[:chunk])  # warm-up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(timed_passes):
            one_pass()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / timed_passes
        return {"value": x_dev.shape[0] / (ms * 1e-3), "unit": "instances/s", "ms_per_step": ms, "kind": "port",
                "device": torch.cuda.get_device_name(dev), "same_config": True,
                "what": f"eager PyTorch restatement of the reference path on the GPU, the same {x_dev.shape[0]}-instance "
                        f"step in chunks of {chunk}; torch matmul-form kNN + this repo's FPS op in place of pytorch3d; "
                        f"{timed_passes} timed passes after a warm-up chunk"}
    except Exception as e:  # the baseline leg must never take the bench line down
        return {"error": f"{type(e).__name__}: {e}"[:300]}
    finally:
        p3d_shim.ON_DEVICE, p3d_shim.FPS_IMPL, p3d_shim.EXACT = old
        torch.cuda.empty_cache()


def oracle_sample_check(ls, model, sd, x_batch, n_set, n_take=4):
    """<= 8 instances of the timed workload (n_take ref + their planted rescans are not contiguous, so simply the first
    n_take ref and first n_take rescan instances of pair 0) through the CUDA path and through oracle.restatement on
    the CPU: same assignments, codes within 1e-4 except where an fp32 near-tie flipped a neighbour (counted)."""
    from oracle import restatement as R

    xa, xb = x_batch[:n_take].cpu(), x_batch[n_set:n_set + n_take].cpu()
    dev = x_batch.device
    out = ls.More_Solver(model).solve_scene_pair(xa.to(dev), xb.to(dev))
    with torch.no_grad():
        ca, cb = R.encode(sd, xa), R.encode(sd, xb)
        m = R.sequential_match(ca["z_inv"], cb["z_inv"])["matches0"]
        Rk, tk, _ = R.kabsch(ca["z_so3"] + ca["t"], (cb["z_so3"] + cb["t"])[m])
    rel = lambda u, v: ((u.cpu() - v).reshape(u.shape[0], -1).abs().amax(1) / v.reshape(v.shape[0], -1).abs().amax(1))
    err = torch.cat([torch.stack([rel(out[side][k], c[k]) for k in ("z_so3", "z_inv", "s", "t")]).amax(0)
                     for side, c in (("ref_codes", ca), ("rescan_codes", cb))])   # per instance, 2 * n_take
    return {"instances": 2 * n_take, "matches_equal": bool(torch.equal(out["matches"]["matches0"].cpu(), m)),
            "codes_within_1e-4": int((err < 1e-4).sum()), "codes_max_rel_err": float(err.max()),
            "R_max_abs_err": float((out["R"].cpu() - Rk).abs().max())}


def sdf_section(model, sd, dev, with_cpu=True, n_inst=64, M=100_000, timed=3):
    """BASELINE config[4] / SURVEY C5: SDF decoder reconstruction query, 64 instances x 100 000 query points on one GPU.
    Codes come from encoding 64 synthetic instances; queries are uniform in the 1.1-padded unit cube of every
    instance's canonical frame (mesh_extractor2.py:100).  Device-resident timing (CUDA events) and an end-to-end leg
    from pinned host memory (H2D of the queries, D2H of the SDF values inside the timed region)."""
    from livingscenes_b200 import synthetic as S

    x = S.synth_parts(n_inst, N_POINTS, 555).to(dev)
    codes = model.encode(x)
    q_host = S.sdf_queries(codes["s"].cpu(), codes["t"].cpu(), M, 1239).pin_memory()
    q = q_host.to(dev)
    fn = lambda qq: model.decoder(qq, None, codes, return_sdf=True)
    for _ in range(2):
        out = fn(q)
    torch.cuda.synchronize()
    ms = []
    for _ in range(timed):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn(q)
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    h_out = torch.empty(n_inst, M).pin_memory()
    e2e = []
    for _ in range(timed):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        h_out.copy_(fn(q_host.to(dev, non_blocking=True)), non_blocking=True)
        b.record()
        b.synchronize()
        e2e.append(a.elapsed_time(b))
    t_ms, e_ms = statistics.median(ms), statistics.median(e2e)
    _, _, _, tf32_peak, tf32_src = measured_peaks()
    macs = 257 * 768 + 2 * 768 * 768 + 768 * 255 + 512 * 768 + 3 * 768 * 768 + 768   # per point after the layer-0/4 collapse
    issued = 3 * 2 * macs * n_inst * M / (t_ms * 1e-3) / 1e12
    # self-check on a sample against the CPU oracle
    chk = None
    cpu = None
    if with_cpu:
        from oracle import restatement as R

        cc = {k: v[:2].cpu() for k, v in codes.items()}
        qs = q_host[:2, :20000]
        t0 = time.perf_counter()
        with torch.no_grad():
            ref = R.sdf_decode(sd, qs, cc)
        dt = time.perf_counter() - t0
        chk = float((out[:2, :20000].cpu() - ref).abs().max())
        assert chk < 1e-4, f"SDF values differ from the oracle: {chk}"
        cpu = {"value": 40000 / dt, "unit": "query points/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"2 instances x 20 000 query points through oracle.restatement.sdf_decode ({dt:.1f} s)"}
    return {"workload": f"BASELINE config[4]: SDF decoder query, {n_inst} instances x {M} points, 1 GPU", "metric": "query points/sec",
            "value": n_inst * M / (t_ms * 1e-3), "unit": "query points/s", "ms": t_ms,
            "e2e": {"value": n_inst * M / (e_ms * 1e-3), "unit": "query points/s", "ms": e_ms,
                    "h2d_bytes": q_host.numel() * 4, "d2h_bytes": h_out.numel() * 4},
            "roofline": {"bound": "tensor", "kernel": "k_gemm_tc3 (tcgen05 3xTF32, activations in tensor memory, 128 x 256 x 16 MMA blocks): the 8 hidden layers of the DeepSDF MLP",
                         "achieved": issued, "peak": tf32_peak, "unit": "TFLOP/s", "frac": issued / tf32_peak,
                         "peak_source": tf32_src, "flops": "executed: 3 TF32 MMA passes x 2 x 3.35 M MAC per point (SURVEY.md 7.1 fact 4)",
                         "fp32_equivalent_TFLOPs": issued / 3},
            "max_abs_err_vs_oracle_sample": chk, "cpu_baseline": cpu}


def run_sdf(args):
    import livingscenes_b200 as ls

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sd, wdesc = load_state_dict()
    model = ls.Shape_Prior.from_state_dict(sd).to(dev).eval()
    sampler = ClockSampler(dev.index)
    sampler.start()
    sec = sdf_section(model, sd, dev, with_cpu=not args.no_cpu_baseline, timed=max(3, min(args.steps, 10)))
    sec.update({"n_gpus": 1, "higher_is_better": True, "dtype": "f32", "data": f"synthetic ({wdesc})", "clocks": sampler.stop(),
                "config": {"workload": sec.pop("workload")}, "vs_baseline": None, "scaling": "weak"})
    print(json.dumps(sec))


def run_b200(args):
    import torch.distributed as dist

    import livingscenes_b200 as ls
    from livingscenes_b200 import _lib
    from livingscenes_b200.dist import all_gather_codes, encode_sharded, shard_range, unpack_codes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner to STDOUT (fd 1) when the first communicator is created: point fd 1 at
        # stderr until the JSON line is due, so that stdout carries exactly one line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"

    sd, wdesc = load_state_dict()
    model = ls.Shape_Prior.from_state_dict(sd).to(dev).eval()
    # Rank r holds, per scene pair, ITS ref scan and the rescan of rank r-1's ref scan: every rank then matches its ref
    # scans against rescan codes that were encoded on rank r+1 and arrive only through the all-gather, so a wrong or
    # stale gathered table breaks the planted-permutation check below.  (world == 1: both scans are local.)
    host_sets, dev_sets, truths = [], [], []
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    for k in range(N_INPUT_SETS):
        mine, truth = make_scene_batch(PAIRS_PER_GPU, 100 + 17 * k + 1000 * rank)
        x = mine
        if world > 1:
            theirs, _ = make_scene_batch(PAIRS_PER_GPU, 100 + 17 * k + 1000 * prv)
            x = mine.clone().view(PAIRS_PER_GPU, 2, INST_PER_SET, 3, N_POINTS)
            x[:, 1] = theirs.view(PAIRS_PER_GPU, 2, INST_PER_SET, 3, N_POINTS)[:, 1]
            x = x.view(-1, 3, N_POINTS).contiguous()
        host_sets.append(x.pin_memory())
        dev_sets.append(x.to(dev))
        truths.append(truth)
    sizes = [INST_PER_SET] * PAIRS_PER_GPU
    n_total = world * INST_PER_GPU
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def step(x):
        rec = model.encode_packed(x)["packed"]                      # [256,1028]
        other = rec
        if world > 1:
            full = all_gather_codes(rec, n_total)                    # one NCCL all-gather
            other = full[nxt * INST_PER_GPU:(nxt + 1) * INST_PER_GPU]  # rescans of MY ref scans live on rank r+1
        ref = unpack_codes(rec.view(PAIRS_PER_GPU, 2, INST_PER_SET, -1)[:, 0].reshape(-1, rec.shape[1]))
        res = unpack_codes(other.view(PAIRS_PER_GPU, 2, INST_PER_SET, -1)[:, 1].reshape(-1, rec.shape[1]))
        m = ls.sequential_matcher_batched(ref["z_inv"].contiguous(), res["z_inv"].contiguous(), sizes, sizes)
        # pair-local rescan index -> row in `res`
        base = torch.arange(PAIRS_PER_GPU, device=dev).repeat_interleave(INST_PER_SET) * INST_PER_SET
        gm = torch.where(m["matches0"] >= 0, m["matches0"] + base, m["matches0"])
        R, t, _ = ls.kabsch_from_codes(ref, res, gm)
        return m["matches0"], R, t, rec

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(max(args.warmup, 3)):
        step(dev_sets[w % N_INPUT_SETS])
    barrier()

    # ---- capture the whole step (encode -> [all-gather] -> match -> pose) in ONE CUDA graph.
    # Every C-ABI call launches on torch's current stream, takes caller-owned workspaces and never
    # synchronises, so the step is capturable as is; the input lives in a static buffer that is refilled
    # (device copy for `value`, H2D copy for `e2e`) before each replay.  NCCL all-gather is captured too.
    use_graph = not args.no_graph
    x_static = dev_sets[0].clone()
    graph, static_out = None, None
    if use_graph:
        graph, static_out = capture_graph(lambda: step(x_static), barrier)

    def run_step(x_src, non_blocking=False):
        if graph is None:
            return step(x_src if x_src.is_cuda else x_src.to(dev, non_blocking=non_blocking))
        x_static.copy_(x_src, non_blocking=non_blocking)
        graph.replay()
        return static_out

    # ---------------- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    total_ms = 0.0
    for k in range(args.steps):
        flush.zero_()                                   # evict L2 between timed steps (not timed)
        ev[k][0].record()
        out = run_step(dev_sets[k % N_INPUT_SETS])
        ev[k][1].record()
        ev[k][1].synchronize()
        total_ms += ev[k][0].elapsed_time(ev[k][1])
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3

    # ---------------- the same K steps launched eagerly with per-stage CUDA events (ls_profile_*): kernel
    # durations for the roofline, and the launch count of one step (a graph replay issues the same kernels)
    _lib.profile_enable(True)
    stage_ms = {}
    launches0 = _lib.kernel_launches()
    eager_ms = 0.0
    for k in range(args.steps):
        flush.zero_()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        step(dev_sets[k % N_INPUT_SETS])
        b_.record()
        b_.synchronize()
        eager_ms += a_.elapsed_time(b_)
        per_step = {}
        for name, layer, ms in _lib.profile_read():   # a stage may be issued in several waves: sum them per step
            per_step[(name, layer)] = per_step.get((name, layer), 0.0) + ms
        for key, ms in per_step.items():
            stage_ms.setdefault(key, []).append(ms)
    launches = (_lib.kernel_launches() - launches0) // max(args.steps, 1)
    _lib.profile_enable(False)
    barrier()

    # ---------------- timed region 2: end to end from pinned host memory
    h_m = torch.empty(PAIRS_PER_GPU * INST_PER_SET, dtype=torch.int64).pin_memory()
    h_R = torch.empty(PAIRS_PER_GPU * INST_PER_SET, 3, 3).pin_memory()
    h_t = torch.empty(PAIRS_PER_GPU * INST_PER_SET, 3, 1).pin_memory()
    e2e_ms = 0.0
    barrier()
    for k in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m0, R, t, _ = run_step(host_sets[k % N_INPUT_SETS], non_blocking=True)
        h_m.copy_(m0, non_blocking=True)
        h_R.copy_(R, non_blocking=True)
        h_t.copy_(t, non_blocking=True)
        b.record()
        b.synchronize()
        e2e_ms += a.elapsed_time(b)
    barrier()
    clocks = sampler.stop() if rank == 0 else None  # sampled over all three timed regions (value, per-stage, e2e)
    h2d = host_sets[0].numel() * 4
    d2h = h_m.numel() * 8 + h_R.numel() * 4 + h_t.numel() * 4

    # ---------------- self-check of the timed step (not timed): (1) the graph replay equals a plain eager run bit for
    # bit, (2) every input set's matches / poses equal the planted permutation / SE(3) -- with N > 1 the rescan codes
    # come from ANOTHER rank through the all-gather, (3) rank 0, N = 1: a <= 8-instance sample against oracle/.
    checked = {"graph_equals_eager": True, "planted": None, "oracle_sample": None}
    worst = {"match_recall": 1.0, "rre_deg_median": 0.0, "rre_deg_max": 0.0, "rte_median": 0.0}
    for k in range(N_INPUT_SETS):
        g_out = [o.clone() for o in run_step(dev_sets[k])[:3]]
        e_out = step(dev_sets[k])[:3]
        torch.cuda.synchronize()
        checked["graph_equals_eager"] &= all(torch.equal(u, v) for u, v in zip(g_out, e_out))
        c = check_planted(g_out[0], g_out[1], g_out[2], truths[k])
        worst = {"match_recall": min(worst["match_recall"], c["match_recall"]),
                 **{q: max(worst[q], c[q]) for q in ("rre_deg_median", "rre_deg_max", "rte_median")}}
    checked["planted"] = worst
    ok_local = checked["graph_equals_eager"] and worst["match_recall"] == 1.0 and worst["rre_deg_median"] < 1.0
    okt = torch.tensor([1.0 if ok_local else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    assert float(okt) == 1.0, f"self-check of the timed step failed on some rank (rank {rank}: {checked})"

    # ---------------- strong scaling, BASELINE config[3] / SURVEY C4: ONE scene of 128 instances (64 ref + 64 rescan)
    # block-partitioned over the ranks (dist.shard_range), encode_sharded = local encode + the all-gather, then the
    # match + poses run replicated on the GATHERED table on every rank.
    c4 = None
    if not args.no_c4:
        n_scene = 64
        xs, truth4 = make_scene_batch(1, 4242, n_set=n_scene)
        xs = xs.to(dev)

        def c4_step():
            code = encode_sharded(model, xs)                         # [128] codes on every rank
            ref = {k: v[:n_scene].contiguous() for k, v in code.items()}
            res = {k: v[n_scene:].contiguous() for k, v in code.items()}
            m = ls.sequential_matcher(ref["z_inv"], res["z_inv"])
            R, t, _ = ls.kabsch_from_codes(ref, res, m["matches0"])
            return m["matches0"], R, t, code

        for _ in range(3):
            c4_step()
        barrier()
        g4, out4 = (capture_graph(c4_step, barrier) if use_graph else (None, None))
        c4_ms = 0.0
        n4 = max(args.steps, 10)
        for k in range(n4):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if g4 is not None:
                g4.replay()
            else:
                out4 = c4_step()
            b.record()
            b.synchronize()
            c4_ms += a.elapsed_time(b)
        barrier()
        # the gathered table must equal a 1-GPU encode of the full list bit for bit, on every rank
        full_local = model.encode(xs)
        same = all(torch.equal(out4[3][k].reshape(-1), full_local[k].reshape(-1)) for k in ("z_so3", "z_inv", "s", "t"))
        pl = check_planted(out4[0], out4[1], out4[2], truth4)
        # fixed cost at this batch size: per-stage events of one local shard encode
        _lib.profile_enable(True)
        lo, hi = shard_range(2 * n_scene, rank, world)
        model.encode_packed(xs[lo:hi])
        torch.cuda.synchronize()
        st4 = {}
        for name, layer, ms in _lib.profile_read():
            st4[name] = st4.get(name, 0.0) + ms
        _lib.profile_enable(False)
        t4 = torch.tensor([c4_ms, 0.0 if (same and pl["match_recall"] == 1.0) else 1.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t4, op=dist.ReduceOp.MAX)
        c4 = {"workload": "BASELINE config[3]: one scene of 128 instances x 1024 points (64 ref + 64 rescan), block-partitioned "
                          f"over {world} GPU(s) ({(2 * n_scene + world - 1) // world}/GPU), one all-gather, replicated match + Kabsch",
              "scaling": "strong", "instances": 2 * n_scene, "steps": n4, "ms_per_step": float(t4[0]) / n4,
              "instances_per_s": 2 * n_scene / (float(t4[0]) / n4 * 1e-3),
              "gathered_table_bit_identical_to_1gpu_encode": bool(float(t4[1]) == 0.0 and same),
              "planted": pl, "shard_stage_ms": {k: round(v, 4) for k, v in sorted(st4.items())}}
        assert float(t4[1]) == 0.0, f"C4: gathered table / matches differ from the 1-GPU run (rank {rank}: same={same}, {pl})"
        del g4

    # max over ranks
    tt = torch.tensor([total_ms, e2e_ms, wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, wall_ms = (float(v) for v in tt.tolist())

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = n_total / (ms_per_step * 1e-3)
        e2e_val = n_total / (e2e_ms / args.steps * 1e-3)
        # ---- roofline of the dominant kNN+EdgeConv layer.  The fused path of layer l is four launches on the
        # same stream: k_knn_pack (features -> UMMA images), k_knn_tc (tcgen05 candidate filter), k_knn_rerank
        # (exact fp32 re-rank) and k_knn_edge (EdgeConv + pooling); its duration is the sum of their event times.
        peak, peak_src, sm_max, tf32_peak, tf32_src = measured_peaks()
        geo = layer_geometry(N_POINTS)
        mean = lambda name: {l: statistics.mean(v) for (n, l), v in stage_ms.items() if n == name}
        edge, filt, rer = mean("knn_edgeconv"), mean("knn_filter"), mean("knn_rerank")
        tabs, gconv = mean("gemm_tables"), mean("global_conv")
        knn = {l: edge[l] + filt.get(l, 0.0) + rer.get(l, 0.0) for l in edge}
        dom = max(knn, key=knn.get)
        byt, flops = knn_edge_algorithmic(geo[dom])
        dur_s = knn[dom] * 1e-3
        achieved = byt * INST_PER_GPU / dur_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "knn_edgeconv_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(f"layer{dom}_dram_bytes_per_launch")
        stages = {}
        for (n, l), v in sorted(stage_ms.items()):
            stages[f"{n}" + (f"[{l}]" if l >= 0 else "")] = round(statistics.mean(v), 4)
        enc_ms = sum(stages.values())
        all_layers = []
        pad = lambda n, m: (n + m - 1) // m * m
        for l in sorted(knn):
            b_, f_ = knn_edge_algorithmic(geo[l])
            g_ = geo[l]
            row = {"layer": l, "ms": round(knn[l], 4), "filter_ms": round(filt.get(l, 0.0), 4),
                   "rerank_ms": round(rer.get(l, 0.0), 4), "edgeconv_ms": round(edge[l], 4),
                   "algorithmic_GBps": round(b_ * INST_PER_GPU / (knn[l] * 1e-3) / 1e9, 2)}
            if l in filt:
                # tensor work of the filter: 3 TF32 MMAs (hi*hi + hi*lo + lo*hi) over padded 128-tiles
                tf = 3 * 2 * pad(g_["n_dst"], 128) * pad(g_["n_src"], 128) * pad(3 * g_["c_in"], 8) * INST_PER_GPU
                row["filter_tensor_TFLOPs_incl_pack"] = round(tf / (filt[l] * 1e-3) / 1e12, 1)
            if l >= 1:
                # executed work of the EdgeConv kernel: gathered table rows (L2) -- 16 neighbour rows + 1 dst row per point
                nb = 2 if l >= 2 else 1
                row_s, row_d = 2 * nb * g_["c_out"] * 3 * 4, (2 * nb + (2 if l >= 2 else 0)) * g_["c_out"] * 3 * 4
                gath = (16 * row_s + row_d) * g_["n_dst"] * INST_PER_GPU
                row["edgeconv_gather_GBps_from_L2"] = round(gath / (edge[l] * 1e-3) / 1e9, 1)
            if l in tabs:
                # executed tensor work of the two table GEMMs (3 TF32 passes, padded to the 128 x 128 x 16 tiles)
                nb = 2 if l >= 2 else 1
                r_s, r_d = 2 * nb * g_["c_out"], (2 * nb + (2 if l >= 2 else 0)) * g_["c_out"]
                tf = 3 * 2 * pad(g_["c_in"], 16) * 3 * (pad(r_s, 128) * g_["n_src"] + pad(r_d, 128) * g_["n_dst"]) * INST_PER_GPU
                row["table_gemm_ms"] = round(tabs[l], 4)
                row["table_gemm_tensor_TFLOPs"] = round(tf / (tabs[l] * 1e-3) / 1e12, 1)
                row["table_gemm_write_GBps"] = round((r_s * g_["n_src"] + r_d * g_["n_dst"]) * 12 * INST_PER_GPU / (tabs[l] * 1e-3) / 1e9, 1)
            all_layers.append(row)
        gemm_ms = sum(tabs.values()) + sum(gconv.values())
        # What the EdgeConv kernel is actually bound by: the gathered table rows all come through the L2 slices.  Peak =
        # the LTS cap of B300_MICROARCH.md (~6 300 B/cycle, path independent) at the SM clock sampled under load.
        l2_peak = 6300.0 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e9
        l2_best = max((r for r in all_layers if "edgeconv_gather_GBps_from_L2" in r), key=lambda r: r["edgeconv_gather_GBps_from_L2"])
        l2_rec = {"bound": "l2", "kernel": f"k_knn_edge (EdgeConv + pooling) of layer {l2_best['layer']}: 16 gathered neighbour rows + 1 dst row "
                                           "of the point-level tables per dst point",
                  "achieved": l2_best["edgeconv_gather_GBps_from_L2"], "peak": round(l2_peak, 1), "unit": "GB/s",
                  "frac": round(l2_best["edgeconv_gather_GBps_from_L2"] / l2_peak, 3),
                  "peak_source": "B300_MICROARCH.md LTS throughput cap 6300 B/cycle x sampled SM clock (no measured L2 figure in MEASURED_PEAKS.json)",
                  "note": "useful gathered bytes only (ncu lts__t_bytes of the same launch is ~1.5x: dst rows, outputs, DRAM fills); "
                          "measured-not-adopted attempts to go past it: profiles/r02/experiments.md sections 1 and 6"}
        # ---- CPU baseline (bounded sample, N=1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sample = 16  # ~15 s of host work (the spec asks for a bounded 10-30 s sample)
            sd_c, x_c, mods, kind, _ = cpu_setup(sample)
            cpu_path_step(sd_c, torch.cat([x_c[:1], x_c[sample // 2:sample // 2 + 1]]), mods)  # warm-up on a 1+1 pair
            t0 = time.perf_counter()
            cpu_path_step(sd_c, x_c, mods)
            dt = time.perf_counter() - t0
            cpu = {"value": sample / dt, "unit": "instances/s", "cores": torch.get_num_threads(), "kind": kind,
                   "sample": f"{sample} instances of {N_POINTS} points (1 scene pair of {sample // 2}+{sample // 2}): encode + sequential "
                             f"match + Kabsch, 1 timed pass after 1 warm-up ({dt:.1f} s); the full 256-instance step is "
                             f"16 such samples (per-instance cost is batch-independent on the CPU: extrapolated {256 / (sample / dt):.0f} s/step)"}
            from oracle import p3d_shim
            p3d_shim.EXACT = True
            checked["oracle_sample"] = oracle_sample_check(ls, model, sd, dev_sets[0], INST_PER_SET)
            assert checked["oracle_sample"]["matches_equal"], checked
        gpu_eager = None
        if world == 1 and not args.no_gpu_eager:
            gpu_eager = gpu_eager_leg(sd, dev_sets[0], sizes)
        sdf_c5 = None
        if world == 1 and not args.no_sdf:
            torch.cuda.empty_cache()
            sdf_c5 = sdf_section(model, sd, dev, with_cpu=not args.no_cpu_baseline)
        line = {
            "metric": "instances/sec (N=1024 pts) encode+match+pose", "value": value, "unit": "instances/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": f"synthetic ({wdesc})",
            "config": {"workload": WORKLOAD,
                       "instances_per_gpu": INST_PER_GPU, "n_points": N_POINTS, "pairs_per_gpu": PAIRS_PER_GPU,
                       "parallelism": f"instance-sharded x{world}" + (" + 1 NCCL all-gather of packed codes; every rank matches its ref "
                                                                      "scans against rescan codes encoded on rank r+1 (taken from the gathered table)" if world > 1 else ""),
                       "l2": "256 MiB buffer written between timed steps (L2 flush); per-step working set ~2 GB >> 126 MB L2",
                       "timing": "CUDA events per step on the launching stream, summed over K steps, max over ranks",
                       "launch": "one CUDA graph replay per step" if graph is not None else "eager launches"},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "instances/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "checked": bool(ok_local), "checks": checked,
            "roofline": {"bound": "hbm",
                         "kernel": f"fused kNN+EdgeConv path of layer {dom}: k_knn_pack + k_knn_tc (tcgen05 filter) + "
                                   "k_knn_rerank + k_knn_edge (EdgeConv + attention pool)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "launch_ms": knn[dom], "algorithmic_bytes_per_launch": byt * INST_PER_GPU,
                         "share_of_encoder": knn[dom] / enc_ms,
                         "traffic_source": "ncu dram__bytes of the same four launches in scripts/one_forward.py (same shapes, "
                                           "not the bench process): profiles/knn_edgeconv_traffic.json",
                         "limiter": "NOT the HBM pipe: BASELINE.json's metric asks for HBM GB/s over the algorithmic bytes "
                                    "(SURVEY.md 8d: layer input + output + int64 graph), but the path is bound by L2 row gathers "
                                    "(edgeconv_gather_GBps_from_L2 vs ~12 TB/s of L2) and the tensor pipe of the filter; per layer in all_layers",
                         "all_layers": all_layers},
            "roofline_l2": l2_rec,
            "roofline_tensor": {"bound": "tensor", "kernel": "k_gemm_tc2 (K < 128) / k_gemm_tc3 (K >= 128), 3xTF32 tcgen05: point-level table GEMMs of layers 1-6",
                                "unit": "TFLOP/s", "peak": tf32_peak, "peak_source": tf32_src, "gemm_ms_per_step": round(gemm_ms, 4),
                                "achieved": sum(r.get("table_gemm_tensor_TFLOPs", 0.0) * r.get("table_gemm_ms", 0.0) for r in all_layers)
                                / max(sum(r.get("table_gemm_ms", 0.0) for r in all_layers), 1e-9),
                                "note": "executed FLOPs (3 MMA passes over padded tiles) of the table GEMMs / their event time"},
            "cpu_baseline": cpu,
            "gpu_eager_baseline": gpu_eager,
            "strong_c4": c4,
            "sdf_c5": sdf_c5,
            "stages_ms": stages,
            "eager_ms_per_step": eager_ms / args.steps,
            "wall_ms_per_step_incl_flush": wall_ms / args.steps,
        }
        line["roofline_tensor"]["frac"] = line["roofline_tensor"]["achieved"] / tf32_peak
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line))
    if world > 1:
        # The captured step holds NCCL work: drop the graph before tearing the communicator down, and leave with a
        # hard exit once every rank is past the barrier -- destroy_process_group() next to a live captured
        # collective was seen to hang the job after the JSON line had been printed.
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        del graph, static_out
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the eager-PyTorch-on-GPU baseline leg")
    ap.add_argument("--no-c4", action="store_true", help="skip the strong-scaling config[3] section")
    ap.add_argument("--no-sdf", action="store_true", help="skip the SDF decoder config[4] section")
    ap.add_argument("--workload", default="encode", choices=["encode", "sdf"],
                    help="encode: the headline encode+match+pose step; sdf: BASELINE config[4] alone (query points/s)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "sdf":
        run_sdf(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
