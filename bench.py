#!/usr/bin/env python
"""bench.py -- headline benchmark of the GKR prover hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one GKR proof (gkr_prove through the C ABI) of the synthetic layered add/mul circuit of
BASELINE.json config 3: 2^20 gates/layer x 16 layers, BN254 Fr, MiMC7 transcript on the host.
  value : ms per proof with circuit + witness already resident in HBM (device time, CUDA events on the
          prover's stream, max over ranks).  N > 1: every rank proves one independent proof of the same
          shape (batch distribution, no data-path collective) -> weak scaling, value = ms per proof
          amortised over the job (max-rank time / N); `latency_ms_per_proof` and `proofs_per_s` say the same two ways.
  parity: every timed path is checked inside this script -- gkr_verify on the proofs of all seeds, on the 2^24-gate layer
          and on a sample of the batch; the verifier's chain / transcript / final-evaluation checks on the 2^24 and
          2^28 sumchecks; at N > 1 the sharded sumcheck and the table-sharded proof bit-exact against rank 0's
          single-GPU results.
  e2e   : the same metric through the public host API with HOST buffers: pinned input layer -> H2D ->
          on-device circuit evaluation -> proof -> Proof arrays back on the host.
Extra objects: roofline (dominant kernel class, per-launch CUDA events inside the library),
cpu_baseline (dense CPU oracle port on a bounded sample), sumcheck (standalone 3-table product
sumcheck, Melem/s + HBM roofline fraction), clocks.
--impl reference times the CPU oracle port of the reference algorithm (the Rust reference cannot be built
here: no toolchain) on the same config, each step a bounded sample.
Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's own banner / debug lines (printed to stdout when the environment sets
# NCCL_DEBUG) go to stderr instead
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402

METRIC = "gkr_prove_ms"
UNIT = "ms"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _workload_name(k, layers):
    return f"synthetic layered add/mul circuit, 2^{k} gates/layer x {layers} layers, BN254 Fr, MiMC7-91 transcript"


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background during the timed region)
# --------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

        def away_from_the_prover():
            # the proof is a serial chain of host hashes on one core: keep the sampler (NVML start-up and one query
            # every 250 ms) on the last CPU this process may use, where the scheduler will not put that thread
            try:
                cpus = sorted(os.sched_getaffinity(0))
                if len(cpus) > 1:
                    os.sched_setaffinity(0, {cpus[-1]})
            except OSError:
                pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "250", "-i", str(index)], stdout=self.tmp, stderr=subprocess.DEVNULL,
                                         preexec_fn=away_from_the_prover)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "no samples"}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the dense oracle port on a bounded sample
# --------------------------------------------------------------------------------------------------
def _all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every core of the box"""
    from oracle import oracle as orc
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    orc.set_num_threads(n)
    return n


def cpu_sample_ms(k: int, layers: int, sample_layers: int, seed: int = 1):
    """ms for a full `layers`-layer proof, extrapolated from proving the first `sample_layers` layers
    (all layers have the same shape and cost) with the dense CPU oracle on all host threads."""
    from gkr_b200 import synthetic as syn
    from oracle import oracle as orc
    _all_host_threads()
    sl = min(sample_layers, layers)
    circ = syn.layered_circuit(seed, k, sl)
    inputs = syn.input_values(seed, k)
    ol = [orc.DenseLayer(L.k_out, L.k_in, L.gtype, L.left, L.right) for L in circ]
    vals = orc.evaluate_circuit(ol, inputs.view(np.uint8).reshape(-1, 32))
    t0 = time.perf_counter()
    orc.gkr_prove(ol, vals)
    dt = time.perf_counter() - t0
    return dt * 1e3 * (layers / sl), orc.num_threads(), f"{sl} of {layers} layers of the same circuit, scaled x{layers / sl:g}"


def literal_reference_seconds(ks=(5, 6, 7, 8)):
    """seconds for ONE layer of 2^k gates with the LITERAL restatement of the reference's term-list algorithm
    (oracle/l0_reference.py, pure Python, 1 thread): documents its O(4^k) growth (x3.3-4 per extra variable: k = 20
    would take on the order of 10^7 s per layer) -- it cannot reach k = 16..20."""
    import random

    from oracle import l0_reference as l0
    out = {}
    rng = random.Random(1)
    l0.mimc7_constants()
    for k in ks:
        gates = [(rng.randrange(2), rng.randrange(1 << k), rng.randrange(1 << k)) for _ in range(1 << k)]
        circ = l0.build_reference_circuit([(k, gates)], k)
        inp, _ = l0.calculate_input([(k, gates)], [rng.randrange(l0.P) for _ in range(1 << k)])
        t0 = time.perf_counter()
        l0.prove(circ, inp)
        out[f"k={k}"] = round(time.perf_counter() - t0, 4)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k, layers = args.k, args.layers
    # a quarter of the layers per step (the per-proof fixed work -- two Moebius transforms, the witness -- is then
    # over-counted x4 instead of x16; bench.py's own cpu_baseline object times the whole proof once)
    sample_layers = max(1, layers // 4) if k >= 18 else layers
    _all_host_threads()
    for _ in range(args.warmup):
        cpu_sample_ms(k, layers, sample_layers)
    vals = []
    cores, sample = 0, ""
    for _ in range(args.steps):
        ms, cores, sample = cpu_sample_ms(k, layers, sample_layers)
        vals.append(ms)
    ms = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 (BN254 Fr, 254-bit integers)", "data": "synthetic",
        "config": {"workload": _workload_name(k, layers), "k": k, "layers": layers},
        "cpu_baseline": {"value": ms, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "dense CPU oracle (oracle/gkr_dense.c, OpenMP); the Rust reference cannot be built here "
                                 "and its term-list algorithm is O(4^k) per round (infeasible at this size)",
                         "literal_reference_algorithm_seconds_per_layer": literal_reference_seconds()},
        "e2e": {"value": ms, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import gkr_b200
    from gkr_b200 import synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; gkr_b200 has no CPU fallback")
    placement = None
    if world > 1 and os.environ.get("GKR_BENCH_PIN", "1") != "0":
        from gkr_b200 import dist as gd0
        placement = gd0.pin_rank_near_gpu(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    hbm_peak, peak_src = _peaks()
    k, layers = args.k, args.layers
    pv = gkr_b200.Prover(local)
    ext = torch.cuda.ExternalStream(pv.stream, device=torch.device("cuda", local))
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.zero_()
        torch.cuda.synchronize()

    seed = 1 + rank
    circ_layers = syn.layered_circuit(seed, k, layers)
    inputs_np = syn.input_values(seed, k)
    pinned = torch.empty(inputs_np.shape, dtype=torch.int32, pin_memory=True)
    pinned_np = pinned.numpy().view(np.uint32)
    pinned_np[...] = inputs_np
    circuit = pv.circuit(circ_layers)
    witness = pv.witness_eval(circuit, pinned_np)

    def step_resident():
        ptr = pv.prove_raw(circuit, witness)
        pv.free_raw(ptr)

    def step_e2e():
        w = pv.witness_eval(circuit, pinned_np)
        ptr = pv.prove_raw(circuit, w)
        pv.free_raw(ptr)
        w.close()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        total_ms = 0.0
        for _ in range(steps):
            flush_l2()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            fn()
            e1.record(ext)
            e1.synchronize()
            barrier()
            total_ms += e0.elapsed_time(e1)
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # the proof is a serial chain of ~640 host hashes: bring the host core to its working clock before timing, and
    # start the clock sampler first -- nvidia-smi's start-up (NVML init, ~0.5 s of host and driver time) otherwise
    # lands in the first timed steps and inflates them by several ms
    clocks = Clocks(local) if rank == 0 else None
    # (with several ranks on one host NVML enumerates every GPU and the cores share their boost budget: spin longer)
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < (1.5 if world == 1 else 4.0):
        step_resident()
    pv.stats(reset=True)
    total_ms = timed(step_resident, args.steps, args.warmup)
    st = pv.stats(reset=True)
    # host transcript time of every rank (contention between the ranks' host threads shows up here first)
    host_per_rank = None
    if world > 1:
        tt = torch.tensor([1e3 * st["transcript_seconds"] / (args.steps + args.warmup)], dtype=torch.float64, device="cuda")
        gathered = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(gathered, tt)
        host_per_rank = [round(float(g.item()), 3) for g in gathered]
    launches_timed = st["kernel_launches"] * args.steps // (args.steps + args.warmup)
    e2e_ms = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    st2 = pv.stats(reset=True)
    n_e2e = args.steps + max(1, args.warmup // 2)
    clock_info = clocks.stop() if clocks else None

    ms_per_step = total_ms / args.steps
    value = ms_per_step / world                      # amortised ms per proof over the whole job
    e2e_value = e2e_ms / args.steps / world
    parity = {}

    # ---- seeds 1..3 for the timing of this config; every seed's proof goes through gkr_verify (the
    # complete verifier: claim chain, transcript hashes, add_i/mult_i on the device, q, z, the input layer) ----------
    seeds = None
    if world == 1 and args.seeds > 1:
        seeds = {str(seed): {"ms": round(ms_per_step, 4)}}
        for sd in range(2, args.seeds + 1):
            cl = syn.layered_circuit(sd, k, layers)
            iv = syn.input_values(sd, k)
            cc = pv.circuit(cl)
            ww = pv.witness_eval(cc, iv)

            def step_sd(cc=cc, ww=ww):
                pv.free_raw(pv.prove_raw(cc, ww))
            n_sd = max(2, args.steps // 2)
            seeds[str(sd)] = {"ms": round(timed(step_sd, n_sd, 1) / n_sd, 4)}
            seeds[str(sd)]["verified"] = bool(pv.verify(cc, pv.prove(cc, ww), iv)[0])
            ww.close()
            cc.close()
        seeds[str(seed)]["verified"] = bool(pv.verify(circuit, pv.prove(circuit, witness), pinned_np)[0])
        parity["c3_all_seeds_verified"] = all(v["verified"] for v in seeds.values())
    # ---- BASELINE.json config 2 (2^16 gates x 8 layers, single GPU): timed and verified beside the headline ----------
    config2 = None
    if world == 1 and k == 20 and args.seeds > 1:
        cl = syn.layered_circuit(1, 16, 8)
        iv = syn.input_values(1, 16)
        cc = pv.circuit(cl)
        ww = pv.witness_eval(cc, iv)

        def step_c2(cc=cc, ww=ww):
            pv.free_raw(pv.prove_raw(cc, ww))
        n_c2 = max(4, args.steps)
        config2 = {"workload": "synthetic layered add/mul circuit, 2^16 gates/layer x 8 layers",
                   "ms": round(timed(step_c2, n_c2, 3) / n_c2, 4),
                   "verified": bool(pv.verify(cc, pv.prove(cc, ww), iv)[0])}
        parity["c2_2p16x8_verified"] = config2["verified"]
        ww.close()
        cc.close()
    elif rank == 0:
        parity["c3_proof_verified"] = bool(pv.verify(circuit, pv.prove(circuit, witness), pinned_np)[0])

    # ---- proofs made by the reference's own Python prover (committed fixture): the CUDA path must reproduce them -------
    refpy = os.path.join(ROOT, "tests", "golden", "refpy_vectors.json")
    if rank == 0 and os.path.exists(refpy):
        from gkr_b200 import verify as gv_ref
        parity["reference_python_prover_vectors"] = gv_ref.check_reference_python_vectors(pv, refpy)

    # ---- per-kernel-class device timing (library-side CUDA events around every launch) --------------
    pv.profile(1)
    step_resident()
    prof = pv.profile(0)
    classes = {n: d for n, d in prof.items() if d["launches"]}
    # the class that takes the most device time, over ALL classes (the ~500 latency-bound launches on small tables are
    # the class gkr_round_tail; `line` and `mobius` run on the low-priority stream, overlapped with the rounds)
    dom = max(classes, key=lambda n: prof[n]["ms"])
    d = prof[dom]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except Exception:
            traffic = None
    achieved = d["algo_bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
    proof_bytes = sum(x["algo_bytes"] for x in prof.values())
    dev_ms_all = sum(x["ms"] for x in prof.values())
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "launches": d["launches"], "avg_launch_us": 1e3 * d["ms"] / max(1, d["launches"]),
                "algo_bytes_per_launch": d["algo_bytes"] / max(1, d["launches"]),
                "share_of_device_time": d["ms"] / max(1e-9, dev_ms_all),
                "whole_proof": {"algo_bytes": proof_bytes, "gbs_over_wall_time": proof_bytes / (ms_per_step * 1e-3) / 1e9,
                                "frac_of_hbm_peak_over_wall_time": proof_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                                "gbs_over_device_busy_time": proof_bytes / (dev_ms_all * 1e-3) / 1e9 if dev_ms_all else None,
                                "note": "wall time of a proof is bounded by its 640 serial host hashes, see `host`"},
                "note": "tables of this workload (3 x 32 MiB per phase) mostly fit the 126 MB L2; the HBM-bound "
                        "measurements are the `sumcheck` objects (tables larger than L2)"}

    # ---- standalone product sumcheck (BASELINE.json config 4; tables larger than L2) --------------------------
    # N = 1: whole tables on one GPU.  N > 1: tables sharded on the low log2(N) index bits, per-round NCCL
    # all-gather of the partial sums (strong scaling); rank 0 also times the unsharded run for the speed-up.
    sumcheck = None
    if args.sumcheck_vars:
        from gkr_b200 import dist as gd
        from gkr_b200._lib import GkrError
        v = args.sumcheck_vars
        N = 1 << v
        sc_seed = 1

        def make_tables(n_loc, first, stride):
            return [pv.dev_table_synth(sc_seed, syn.TABLE_STREAM + t, n_loc, first=first, stride=stride) for t in range(3)]

        def sc_profile(fn):
            pv.profile(1)
            fn()
            sp = pv.profile(0)
            kern_ms = sp["prod3_round"]["ms"] + sp["prod3_round_fused"]["ms"]
            kern_bytes = sp["prod3_round"]["algo_bytes"] + sp["prod3_round_fused"]["algo_bytes"]
            return {"ms": kern_ms, "gbs": kern_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms else 0.0,
                    "frac": kern_bytes / (kern_ms * 1e-3) / 1e9 / hbm_peak if kern_ms else 0.0,
                    "first_round_ms": sp["prod3_round"]["ms"],
                    "first_round_gbs": sp["prod3_round"]["algo_bytes"] / (sp["prod3_round"]["ms"] * 1e-3) / 1e9
                    if sp["prod3_round"]["ms"] else 0.0}
        algo = 32.0 * 3 * (4 * N - 6)
        from gkr_b200 import verify as gv
        from gkr_b200.field import fr_to_ints

        def unpack(raw):
            msgs, mlen, chal, fin = raw
            return [fr_to_ints(msgs[j, :mlen[j]]) for j in range(v)], fr_to_ints(chal), fr_to_ints(fin)
        try:
            if world == 1:
                tabs = make_tables(N, 0, 1)
                sc_ms = timed(lambda: pv.sumcheck_prod_raw(tabs, v), args.steps, args.warmup) / args.steps
                rk = sc_profile(lambda: pv.sumcheck_prod_raw(tabs, v))
                # the verifier's checks on the timed computation: chain, transcript, final product, and the final
                # evaluations T_i(r) recomputed on the device with an eq table and a dot product
                m_, c_, f_ = unpack(pv.sumcheck_prod_raw(tabs, v))
                chk = gv.check_sumcheck_prod(m_, c_, f_, evals=[pv.dev_table_eval(t, c_) for t in tabs])
                parity["sumcheck_2p%d" % v] = chk
                extra = {}
            else:
                gd.init_comm(pv)
                tabs = make_tables(N // world, rank, world)
                sc_ms = timed(lambda: pv.sumcheck_prod_sharded_raw(tabs, v), args.steps, args.warmup) / args.steps
                rk = sc_profile(lambda: pv.sumcheck_prod_sharded_raw(tabs, v))
                sharded_raw = pv.sumcheck_prod_sharded_raw(tabs, v)
                for t in tabs:
                    t.close()
                tabs = []
                single_ms = None
                if rank == 0:
                    # the unsharded run on rank 0, on a context of its own: timing reference AND bit-exact parity reference
                    p1 = gkr_b200.Prover(local)
                    ext1 = torch.cuda.ExternalStream(p1.stream, device=torch.device("cuda", local))
                    full = [p1.dev_table_synth(sc_seed, syn.TABLE_STREAM + t, N) for t in range(3)]
                    for _ in range(2):
                        p1.sumcheck_prod_raw(full, v)
                    t0 = torch.cuda.Event(enable_timing=True)
                    t1 = torch.cuda.Event(enable_timing=True)
                    t0.record(ext1)
                    for _ in range(args.steps):
                        single_raw = p1.sumcheck_prod_raw(full, v)
                    t1.record(ext1)
                    t1.synchronize()
                    single_ms = t0.elapsed_time(t1) / args.steps
                    same = all(np.array_equal(a, b) for a, b in zip(single_raw, sharded_raw))
                    m_, c_, f_ = unpack(sharded_raw)
                    chk = gv.check_sumcheck_prod(m_, c_, f_, evals=[p1.dev_table_eval(t, c_) for t in full])
                    chk["sharded_equals_single_gpu"] = bool(same)
                    chk["ok"] = bool(chk["ok"] and same)
                    parity["sumcheck_2p%d_sharded_x%d" % (v, world)] = chk
                    for t in full:
                        t.close()
                    p1.close()
                barrier()
                # one proof of the headline circuit with every layer table-sharded over all ranks (strong scaling;
                # latency-bound: 640 serial rounds, the large ones sharded, the exchange hidden behind the host hash)
                try:
                    sh_layers = syn.layered_circuit(1, k, layers)
                    sh_inputs = syn.input_values(1, k)
                    sh_circ = pv.circuit(sh_layers)                     # created after init_comm => per-rank CSRs
                    sh_wit = pv.witness_eval(sh_circ, sh_inputs)

                    def step_sharded():
                        ptr = pv.prove_raw(sh_circ, sh_wit)
                        pv.free_raw(ptr)
                    gkr_sharded_ms = timed(step_sharded, max(2, args.steps // 2), 2) / max(2, args.steps // 2)
                    sh_proof = pv.prove(sh_circ, sh_wit)
                    sh_wit.close()
                    if rank == 0:
                        # bit-exact against the same proof from one GPU (rank 0's seed is 1 too), which gkr_verify accepts
                        p1 = gkr_b200.Prover(local)
                        c1 = p1.circuit(sh_layers)
                        w1 = p1.witness_eval(c1, sh_inputs)
                        one = p1.prove(c1, w1)
                        fields = ("sumcheck_proofs", "sumcheck_r", "q", "z", "r", "d_coef", "input_coef")
                        parity["gkr_table_sharded_x%d" % world] = {
                            "equals_single_gpu_proof": all(getattr(one, f) == getattr(sh_proof, f) for f in fields),
                            "verified": bool(p1.verify(c1, sh_proof, sh_inputs)[0])}
                        parity["gkr_table_sharded_x%d" % world]["ok"] = all(parity["gkr_table_sharded_x%d" % world].values())
                        w1.close()
                        c1.close()
                        p1.close()
                    barrier()
                except GkrError as e:
                    gkr_sharded_ms = f"error: {e}"
                extra = {"sharding": f"low log2({world}) index bits; per-round partial sums through a shared pinned host block, "
                                     "summed by every rank's host (no NCCL call, no extra launch per round)",
                         "gkr_table_sharded_ms_per_proof": gkr_sharded_ms,
                         "single_gpu_ms_same_run": single_ms,
                         "speedup_vs_single_gpu": (single_ms / sc_ms) if single_ms else None}
            sumcheck = {"n_vars": v, "tables": 3, "n_gpus": world, "ms": sc_ms, "melem_s": N / (sc_ms * 1e-3) / 1e6,
                        "algo_bytes": algo, "gbs_whole_sumcheck": algo / (sc_ms * 1e-3) / 1e9,
                        "frac_whole_sumcheck": algo / (sc_ms * 1e-3) / 1e9 / (hbm_peak * world),
                        "round_kernels_this_rank": rk, "l2": "inputs larger than L2 (3 x %d MiB per rank)" % (N * 32 // world >> 20),
                        **extra}
            for t in tabs:
                t.close()
            # the second size BASELINE.json names (2^24), single GPU only: same measurement, same checks
            if world == 1 and args.sumcheck_vars_small and args.sumcheck_vars_small != v:
                v2 = args.sumcheck_vars_small
                N2 = 1 << v2
                tabs2 = make_tables(N2, 0, 1)
                ms2 = timed(lambda: pv.sumcheck_prod_raw(tabs2, v2), args.steps, args.warmup) / args.steps
                msgs, mlen, chal, fin = pv.sumcheck_prod_raw(tabs2, v2)
                m_ = [fr_to_ints(msgs[j, :mlen[j]]) for j in range(v2)]
                c_, f_ = fr_to_ints(chal), fr_to_ints(fin)
                parity["sumcheck_2p%d" % v2] = gv.check_sumcheck_prod(m_, c_, f_, evals=[pv.dev_table_eval(t, c_) for t in tabs2])
                algo2 = 32.0 * 3 * (4 * N2 - 6)
                sumcheck["second_size"] = {"n_vars": v2, "ms": ms2, "melem_s": N2 / (ms2 * 1e-3) / 1e6,
                                           "gbs_whole_sumcheck": algo2 / (ms2 * 1e-3) / 1e9,
                                           "frac_whole_sumcheck": algo2 / (ms2 * 1e-3) / 1e9 / hbm_peak}
                for t in tabs2:
                    t.close()
        except GkrError as e:
            sumcheck = {"n_vars": v, "error": str(e)}

    # ---- the degree-2 GKR round kernels on HBM-sized tables (one 2^24-gate layer: 3 x 512 MiB per phase) ----------
    gkr_large = None
    if args.large_layer_k and world == 1:
        try:
            lk = args.large_layer_k
            big_layers = syn.layered_circuit(seed, lk, 1)
            big_c = pv.circuit(big_layers)
            big_w = pv.witness_eval(big_c, syn.input_values(seed, lk))
            big_inputs = syn.input_values(seed, lk)
            for _ in range(2):
                pv.free_raw(pv.prove_raw(big_c, big_w))
            pv.profile(1)
            pv.free_raw(pv.prove_raw(big_c, big_w))
            bp = pv.profile(0)
            # the streaming (lazy-accumulation) kernels of this size are exercised nowhere else: verify what they produced
            parity["gkr_layer_2p%d_verified" % lk] = bool(pv.verify(big_c, pv.prove(big_c, big_w), big_inputs)[0])
            big_w.close()
            big_c.close()

            def cls(n):
                x = bp[n]
                g = x["algo_bytes"] / (x["ms"] * 1e-3) / 1e9 if x["ms"] else 0.0
                return {"launches": x["launches"], "ms": round(x["ms"], 4), "gbs": round(g, 1), "frac_of_hbm_peak": round(g / hbm_peak, 3)}
            gkr_large = {"k": lk, "layers": 1, "tables": "3 x %d MiB per phase (larger than L2)" % ((32 << lk) >> 20),
                         "gkr_round_fused": cls("gkr_round_fused"), "lookahead_start_nofold": cls("gkr_round"),
                         "wiring_plus_first_round": cls("wiring"),
                         "note": "launches of >= 2^16 pairs; CUDA events around every launch. gkr_round_fused = direct "
                                 "fused rounds above 2^19 entries + look-ahead rounds below (fold + next message as a "
                                 "polynomial); the first round of each phase is fused into the wiring-sum kernel"}
        except Exception as e:  # noqa: BLE001 - an OOM here must not lose the headline numbers
            gkr_large = {"error": str(e)}

    # ---- several headline-size proofs in flight (throughput of config 3; `value` above stays the latency of ONE proof) --
    # One proof is a chain of 640 host hashes with the device idle most of the time; the lockstep batch prover keeps
    # `in_flight` proofs of the same circuit (different inputs) going per GPU: they share the device and, per host thread,
    # one SIMD hash call per round.  Every rank proves its own `in_flight` proofs; wall clock of the proving alone.
    throughput = None
    if args.in_flight > 1:
        try:
            from gkr_b200.batch import NativeBatch
            n_thr = max(1, min(4, (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 4)))
            lanes_tp = max(1, (args.in_flight + n_thr - 1) // n_thr)
            tp_jobs = [(circ_layers, syn.input_values(1000 + rank * args.in_flight + j, k)) for j in range(args.in_flight)]
            with NativeBatch(n_thr, lanes_tp, local) as nbt:
                # with several proofs in flight the device is the limit: look-ahead rounds (45 % more arithmetic than direct
                # rounds, there to hide the device behind the hash of ONE proof) only for tables of at most 2^16 entries
                nbt.set_option("lookahead_log2", args.in_flight_lookahead_log2)
                nbt.load(tp_jobs)
                tp_proofs = nbt.prove()                       # warm-up + the proofs the verifier sees
                barrier()
                tp_s = 1e30
                for _ in range(2):
                    nbt.prove(keep=False)
                    tp_s = min(tp_s, nbt.seconds)
            torch.cuda.synchronize()
            tt = torch.tensor([tp_s], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tp_s = float(tt.item())
            pqt = gkr_b200.Prover(local)
            ct = pqt.circuit(circ_layers)
            ok_tp = all(pqt.verify(ct, tp_proofs[j], tp_jobs[j][1])[0] for j in (0, args.in_flight - 1))
            ct.close()
            pqt.close()
            del tp_proofs, tp_jobs
            okt = torch.tensor([1 if ok_tp else 0], dtype=torch.int64, device="cuda")
            if world > 1:
                dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            parity["throughput_mode_verified"] = bool(int(okt.item()))
            throughput = {"proofs_in_flight_per_gpu": args.in_flight, "host_threads_per_gpu": n_thr,
                          "proofs_in_lockstep_per_thread": lanes_tp, "lookahead_log2": args.in_flight_lookahead_log2,
                          "ms_total": 1e3 * tp_s,
                          "ms_per_proof_amortised": 1e3 * tp_s / (args.in_flight * world),
                          "proofs_per_s": args.in_flight * world / tp_s,
                          "note": "same circuit, different inputs; gkr_batch (csrc/batch.cpp); the first and last proof of "
                                  "every rank go through gkr_verify"}
        except Exception as e:  # noqa: BLE001 - must not lose the headline numbers
            throughput = {"error": str(e)}

    # ---- batch of small independent proofs (BASELINE.json configs 1 and 5: the sub-circuits of rust/t.circom) ------
    # 364-constraint MiMC7-91 system per input -> 12 sub-circuits (k <= 7, 3-5 layers); inputs in1 = 2 + j are dealt
    # round-robin to the ranks, each rank proves its sub-circuits on a pool of host threads (one context each), as the
    # reference does under rayon (aggregator.rs:413-416).  GKR stage only: the front end runs before the timed region.
    tcircom = None
    if args.tcircom_inputs:
        try:
            from gkr_b200 import frontend as fe
            mine = [j for j in range(args.tcircom_inputs) if j % world == rank]
            jobs = []
            for j in mine:
                r1, w1 = fe.mimc7_constraint_system(2 + j)
                subs, _ = fe.convert_r1cs_wtns_gkr(r1, w1)
                jobs += [(sc.layers, sc.input_values) for sc in subs]
            n_cpus = len(os.sched_getaffinity(0)) if placement and placement.get("pinned") else (os.cpu_count() or 1) // world
            workers = int(os.environ.get("GKR_BATCH_WORKERS", 0)) or max(1, min(16, n_cpus))
            from gkr_b200.batch import NativeBatch, timed_prove_stage
            # the library's lockstep batch prover (gkr_batch: `workers` pinned threads x several proofs per thread hashed
            # in SIMD lanes); the clock is the library's own, around the proving alone, all workers starting together
            nb = NativeBatch(workers, int(os.environ.get("GKR_BATCH_LANES", 0)), local)
            nb.load(jobs)
            batch_proofs = nb.prove()                  # untimed pass: warm-up, and the proofs the verifier samples below
            barrier()
            nb.prove(keep=False)
            dt_local = nb.seconds
            torch.cuda.synchronize()
            dt = torch.tensor([dt_local], dtype=torch.float64, device="cuda")
            barrier()
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            batch_lanes, batch_simd = nb.lanes, nb.simd_hash
            nb.close()
            # one input alone (12 proofs), one proof per thread: the latency of a single aggregation step
            with NativeBatch(min(12, workers), 1, local) as nb1:
                nb1.load(jobs[:12])
                nb1.prove(keep=False)
                one = 1e30
                for _ in range(3):
                    nb1.prove(keep=False)
                    one = min(one, nb1.seconds)
            # the older pool -- one host thread and one scalar transcript per proof -- for comparison
            dt_pool = timed_prove_stage(jobs, workers, local)
            dt = float(dt.item())
            # a sample of the timed path's own proofs through the complete verifier (every rank checks some of its own)
            ok_n = 0
            step_s = max(1, len(jobs) // 24)
            sample = list(zip(jobs, batch_proofs))[::step_s][:24]
            pq = gkr_b200.Prover(local)          # a context without a communicator (pv has one at N > 1)
            for (lay, inp), prf in sample:
                cj = pq.circuit(lay)
                ok_n += 1 if pq.verify(cj, prf, inp)[0] else 0
                cj.close()
            del batch_proofs
            okt = torch.tensor([ok_n, len(sample)], dtype=torch.int64, device="cuda")
            if world > 1:
                dist.all_reduce(okt, op=dist.ReduceOp.SUM)
            parity["batch_sample_verified"] = {"verified": int(okt[0].item()), "of": int(okt[1].item()),
                                               "ok": int(okt[0].item()) == int(okt[1].item())}
            # ---- one recursive round (config 5: "plus one recursive C'_i verifier-circuit round"), rank 0: the 12 proofs
            # of input 0 are packaged, the combined circuit C'_1 = t.circom + one VerifyGKR instance per proof is built
            # as a constraint system (hand-written stand-in for what circom would emit, see
            # frontend.aggregated_constraint_system), compiled by the native front end and proved and verified
            recursive = None
            if rank == 0:
                from gkr_b200.prover import dense_to_proof
                prev = []
                for lay, inp in jobs[:12]:
                    cj = pq.circuit(lay)
                    wj = pq.witness_eval(cj, inp)
                    prev.append(dense_to_proof(pq.prove(cj, wj)))
                    wj.close()
                    cj.close()
                t_fe = time.perf_counter()
                r2, w2 = fe.aggregated_constraint_system(2 + 1, prev)
                subs2 = fe.compile_native(fe.write_r1cs(r2), fe.write_wtns(w2))
                t_fe = time.perf_counter() - t_fe
                jobs2 = [(sc.layers, sc.input_values) for sc in subs2]
                with NativeBatch(min(len(jobs2), workers), 0, local) as nb2:
                    nb2.load(jobs2)
                    proofs2 = nb2.prove()                                                     # warm-up + verifier input
                    dt2 = 1e30
                    for _ in range(3):
                        nb2.prove(keep=False)
                        dt2 = min(dt2, nb2.seconds)
                ok2 = 0
                for (lay, inp), prf in zip(jobs2, proofs2):
                    cj = pq.circuit(lay)
                    ok2 += 1 if pq.verify(cj, prf, inp)[0] else 0
                    cj.close()
                parity["recursive_round_verified"] = {"verified": ok2, "of": len(jobs2), "ok": ok2 == len(jobs2)}
                recursive = {"constraints": len(r2.constraints), "of_which_user_circuit": 364, "sub_circuits": len(jobs2),
                             "max_k": max(max(sc.k) for sc in subs2), "gkr_stage_ms": 1e3 * dt2,
                             "front_end_ms": 1e3 * t_fe,
                             "note": "synthetic stand-in for C'_1: the Horner steps VerifyGKR(meta) adds for the 12 proofs of "
                                     "input 0 (verifier.circom:39-71 with circom's linear constraints simplified away) on top "
                                     "of the t.circom constraints; circom itself is not available here"}
            pq.close()
            tcircom = {"inputs": args.tcircom_inputs, "constraints_per_input": 364, "sub_circuits_per_input": 12,
                       "proofs": 12 * args.tcircom_inputs, "host_threads_per_gpu": workers,
                       "proofs_in_lockstep_per_thread": batch_lanes, "simd_transcript_hash": batch_simd,
                       "thread_pool_proofs_per_s_this_rank": len(jobs) / dt_pool, "ms_total": 1e3 * dt,
                       "ms_per_input": 1e3 * dt / args.tcircom_inputs, "proofs_per_s": 12 * args.tcircom_inputs / dt,
                       "ms_one_input_alone": 1e3 * one, "host_cores": os.cpu_count(), "recursive_round": recursive,
                       "note": "hand-built constraint system of the shape circom emits for rust/t.circom (no circom "
                               "here): an approximation; wall clock of the GKR stage alone (circuits uploaded and "
                               "witnesses evaluated beforehand, as the reference times it, aggregator.rs:406-418), "
                               "max over ranks"}
        except Exception as e:  # noqa: BLE001 - must not lose the headline numbers
            tcircom = {"error": str(e)}

    # ---- integer-multiply ceiling of this device (register-resident Montgomery products) ------------------------
    gmul = pv.bench_field_mul(4, 4, 2000) / 1e9
    int_roof = {"gmul_per_s": gmul, "unit": "G Montgomery products/s (8x32-bit limbs, IMAD.WIDE)",
                "note": "IMAD.WIDE issues at 32 lanes/clk/SM; a fused degree-3 round needs ~1095 wide multiplies per 576 B, "
                        "so the integer pipe, not HBM, bounds the round kernels"}

    # ---- CPU baseline on rank 0, N = 1 only -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            ms1, cores, sample1 = cpu_sample_ms(k, layers, 1 if k >= 20 else 2)
            # one complete proof on all host threads (~10-20 s at the headline size): the number itself, not an extrapolation
            ms, cores, sample = cpu_sample_ms(k, layers, layers)
            cpu = {"value": ms, "unit": UNIT, "cores": cores, "kind": "port", "sample": "the whole proof, once (" + sample + ")",
                   "extrapolated_from_one_layer_ms": ms1,
                   "note": "dense CPU oracle (oracle/gkr_dense.c, OpenMP on every host thread); the Rust reference cannot be built here",
                   "literal_reference_algorithm_seconds_per_layer": literal_reference_seconds()}
        except Exception as e:  # the oracle is a checker, never a dependency of the measured path
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (BN254 Fr, 254-bit integers)", "data": "synthetic",
            "config": {"workload": _workload_name(k, layers), "k": k, "layers": layers,
                       "parallelism": "1 proof per GPU" if world > 1 else "single GPU", "placement_rank0": placement,
                       "l2": "256 MiB flush between timed iterations; witness tables total %d MiB" % ((layers + 1) * (32 << k) >> 20),
                       "timing": "CUDA events on the prover stream, per step, summed; max over ranks"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": st2["h2d_bytes"] // n_e2e,
                    "d2h_bytes_per_step": st2["d2h_bytes"] // n_e2e,
                    "note": "pinned input layer -> H2D -> device circuit evaluation -> gkr_prove -> Proof on host"},
            "gpu_launches": int(launches_timed),
            "latency_ms_per_proof": ms_per_step, "proofs_per_s": 1e3 * world / ms_per_step,
            "parity": dict(parity, all_ok=all((v.get("ok", True) if isinstance(v, dict) else bool(v)) for v in parity.values())),
            "seeds": seeds,
            "config2_2p16x8": config2,
            "roofline": roofline, "cpu_baseline": cpu, "sumcheck": sumcheck, "clocks": clock_info,
            "integer_roofline": int_roof, "gkr_rounds_large_tables": gkr_large,
            "kernel_classes": {n: {"launches": x["launches"], "ms": round(x["ms"], 4),
                                   "gbs": round(x["algo_bytes"] / (x["ms"] * 1e-3) / 1e9, 1) if x["ms"] else None}
                               for n, x in classes.items()},
            "throughput_mode": throughput,
            "t_circom_like_batch": tcircom,
            "host": {"transcript_ms_per_step_per_rank": host_per_rank,
                     "transcript_ms_per_step": 1e3 * st["transcript_seconds"] / (args.steps + args.warmup),
                     "wait_ms_per_step": 1e3 * st["wait_seconds"] / (args.steps + args.warmup)},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--k", type=int, default=20, help="log2 gates per layer (BASELINE config 3: 20)")
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--sumcheck-vars", type=int, default=28, help="standalone 3-table sumcheck size (0 = skip)")
    ap.add_argument("--sumcheck-vars-small", type=int, default=24, help="second standalone sumcheck size, N = 1 only (0 = skip)")
    ap.add_argument("--seeds", type=int, default=3, help="time and verify the headline circuit for seeds 1..SEEDS (N = 1)")
    ap.add_argument("--large-layer-k", type=int, default=24, help="also profile the GKR round kernels on one 2^k-gate layer (0 = skip)")
    ap.add_argument("--tcircom-inputs", type=int, default=64, help="batch of t.circom-like input proofs (0 = skip)")
    ap.add_argument("--in-flight", type=int, default=8, help="headline-size proofs in flight per GPU for the throughput figure (0/1 = skip)")
    ap.add_argument("--in-flight-lookahead-log2", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
