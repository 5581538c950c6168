#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric: 7-card hand evals/s).

Step = one exhaustive evaluation of all C(52,7) = 133,784,560 seven-card hands (BASELINE config 2) producing the
9-bin category histogram and the 7463-bin rank histogram.  With N GPUs the colex combination range is split
into N contiguous shards (one per rank) and the two histograms are all-reduced (NCCL) inside the step, so the
total work per step is fixed ("strong" scaling).  One JSON line is printed by rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Extra objects on the line: `roofline` (integer-issue model), `cpu_baseline` (C oracle on the host
cores), `e2e` (through the public host-buffer API), `equity_mc` (BASELINE config 3: 4096 six-max states,
sample range sharded over the ranks + one all-reduce of the counts), `eval7_hbm` (materialised 2^25-hand batch).
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

NUM_HANDS = 133784560
W_EVAL_LANE_OPS = 150          # algorithmic integer lane-ops per independent 7-card evaluation
NOMINAL_INT_PEAK = 148 * 128 * 1.965e9   # lane-ops/s, balanced alu+fma issue at max clock


def read_json(path):
    try:
        return json.load(open(path))
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_enum(threads):
    """C oracle (restated phevaluator algorithm) on the host cores: the whole C(52,7) enumeration."""
    from oracle import oracle as O
    O.lib()
    t0 = time.perf_counter()
    h9, h = O.enum7(0, NUM_HANDS, nthreads=threads)
    dt = time.perf_counter() - t0
    assert h9.tolist() == O.K1_HIST9
    return {"value": NUM_HANDS / dt, "unit": "evals/s", "cores": threads, "kind": "port",
            "sample": f"all {NUM_HANDS} hands, C port of the phevaluator algorithm (oracle/phe_oracle.c), {threads} threads, {dt:.2f} s"}


def reference_arm(args):
    """The reference's own CPU implementation of the path: pure-Python phevaluator algorithm + Python loop
    (oracle/pyport.py; the PyPI wheel and /root/reference are not on the GPU box), all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import pyport
    cores = os.cpu_count() or 1
    per_step = 250000 * cores      # bounded sample of the same workload: the first hands in colex order
    chunks = [(i * 250000, (i + 1) * 250000) for i in range(cores)]
    with mp.Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(pyport.eval_range, [(b, b + 2000) for b, _ in chunks])
        t0 = time.perf_counter()
        total = 0
        for _ in range(args.steps):
            for n, _hist in pool.map(pyport.eval_range, chunks):
                total += n
        dt = time.perf_counter() - t0
    val = total / dt
    line = {
        "impl": "reference", "metric": "hand_evals_per_s", "value": val, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "exhaustive_c52_7 (bounded sample)", "hands_per_step": per_step},
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} consecutive colex hands per step; pure-Python phevaluator algorithm + Python loop (oracle/pyport.py), "
                                   f"{cores} processes; the PyPI phevaluator wheel is not installable offline"},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mc-rollouts", type=int, default=1_000_000, help="rollouts per state of the equity_mc side measurement (config 3)")
    ap.add_argument("--skip-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import texasholdempocker_rl_b200 as phe

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: the product has no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.gpus = world
    warm = max(args.warmup, 3)

    L = phe._native.lib()
    stream = torch.cuda.current_stream()
    hist = torch.zeros(9 + 7463, dtype=torch.int64, device=dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    b, e = phe.shard_range(NUM_HANDS, rank, world)
    import ctypes as C

    # N > 1: the all-reduce of the histograms is fused into the enumeration kernel (phe_enum7_allreduce: the last CTA
    # exchanges the 60 KB through peer memory over NVLink).  If the symmetric-memory plumbing is unavailable the step
    # falls back to kernel + NCCL all-reduce; `config.collective` says which one ran.
    fused = None
    if world > 1:
        try:
            fused = phe.FusedEnum7()
            hist = fused.out
        except Exception as ex:  # noqa: BLE001
            fused = None
            if rank == 0:
                print(f"fused all-reduce unavailable ({type(ex).__name__}: {ex}); using NCCL", file=sys.stderr)
        ok = torch.tensor([1 if fused is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            fused = None
            hist = torch.zeros(9 + 7463, dtype=torch.int64, device=dev)

    def step():
        if fused is not None:
            fused.run()
            return
        hist.zero_()
        phe._native.check(L.phe_enum7_part(C.c_void_p(hist.data_ptr()), C.c_void_p(hist.data_ptr() + 72), 0, NUM_HANDS, rank, world,
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        if world > 1:
            dist.all_reduce(hist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step()
    barrier()
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "hist7463.json")))
    hh = hist.cpu().numpy()
    assert hh[:9].tolist() == gold["hist9"] and hh[9:].tolist() == gold["hist7463"], "histogram differs from the golden fixture"

    # The step (memset + enumeration kernel [+ all-reduce]) is ~0.2 ms of GPU work issued by five host calls, so it is
    # captured once in a CUDA graph and replayed: one launch per step instead of a launch-bound Python loop.
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    barrier()
    with torch.cuda.graph(graph):
        step()
    for _ in range(warm):
        graph.replay()
    barrier()
    hh = hist.cpu().numpy()
    assert hh[:9].tolist() == gold["hist9"] and hh[9:].tolist() == gold["hist7463"], "graph replay differs from the golden fixture"

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = phe._native.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s0, s1 in ev:
        flush_buf.fill_(1)                      # L2 flush between timed iterations (outside the timed events)
        s0.record()
        graph.replay()
        s1.record()
    barrier()
    # the enumeration kernel alone (this rank's shard), timed on its launching stream
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for k0, k1 in kev:
        hist.zero_()
        k0.record()
        phe._native.check(L.phe_enum7_part(C.c_void_p(hist.data_ptr()), C.c_void_p(hist.data_ptr() + 72), 0, NUM_HANDS, rank, world, C.c_void_p(stream.cuda_stream)))
        k1.record()
    barrier()
    launches = args.steps                                                      # one k_enum7 per replayed graph
    assert phe._native.launch_count() - launches0 == 10                        # the ten un-graphed kernel timings above
    step_ms = sum(a.elapsed_time(c) for a, c in ev)
    kern_ms = sorted(a.elapsed_time(c) for a, c in kev)[len(kev) // 2]         # median, counter memset + kernel
    t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = NUM_HANDS * args.steps / (total_ms * 1e-3)

    extras = {}
    if not args.skip_extras:
        # --- config 3: batched Monte-Carlo equity, sample range sharded over ranks + one all-reduce ---
        states = torch.from_numpy(phe.random_states(4096, 20261019)).to(dev)
        R = args.mc_rollouts
        rb, re_ = phe.shard_range(R, rank, world)
        wtl = torch.zeros(4096, 3, dtype=torch.int64, device=dev)

        def mc_step():
            wtl.zero_()
            phe.equity_mc(states, 5, re_ - rb, seed=0x5EED20261019, sample_offset=rb, out=wtl)
            if world > 1:
                dist.all_reduce(wtl)
        mc_step()
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        m0.record()
        for _ in range(reps):
            mc_step()
        m1.record()
        barrier()
        mt = torch.tensor([m0.elapsed_time(m1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        assert bool((wtl.sum(1) == R).all())
        nb = np.bitwise_count((phe.random_states(4096, 20261019)[:, 0]).astype(np.uint64)).astype(np.int64) - 2
        d = (5 - nb) + 10
        w_roll = float(np.mean(6 * 150 + 70 * np.ceil(d / 4) + 12 * d + 4 * 6))   # SURVEY 8d W_rollout(P=6, b)
        mc_val = 4096 * R / (float(mt.item()) * 1e-3)
        extras["equity_mc"] = {"value": mc_val, "unit": "rollouts/s", "ms_per_step": float(mt.item()),
                               "config": {"workload": "batched_mc_equity", "states": 4096, "players": 6, "rollouts_per_state": R,
                                          "sharding": f"sample index over {world} ranks + all_reduce(int64[4096,3])"},
                               "model_lane_ops_per_rollout": w_roll,
                               "roofline_frac_nominal": mc_val * w_roll / (NOMINAL_INT_PEAK * world),
                               "checksum": int(wtl[:, 0].sum().item())}
        if rank == 0 and world == 1:
            # the same batch end to end through the host-buffer API: numpy states in (H2D), numpy counts out (D2H)
            hstates = phe.random_states(4096, 20261019)
            t0 = time.perf_counter()
            hw = phe.equity_mc(hstates, 5, R, seed=0x5EED20261019)
            mdt = time.perf_counter() - t0
            assert int(hw[:, 0].sum()) == extras["equity_mc"]["checksum"]
            extras["equity_mc"]["e2e"] = {"value": 4096 * R / mdt, "unit": "rollouts/s", "h2d_bytes_per_step": int(hstates.nbytes),
                                          "d2h_bytes_per_step": int(hw.nbytes)}
        # --- materialised-hands micro-config: 2^25 card sets (268 MB > L2) -> ranks ---
        # (this and the two sections after it are single-GPU quantities: reported at N = 1 only)
        if rank == 0 and world == 1:
            g = torch.Generator(device="cpu").manual_seed(1)
            base = torch.rand(1 << 18, 52, generator=g).argsort(1)[:, :7].numpy().astype(np.uint8)
            masks = torch.from_numpy(phe.ids_to_masks(base)).to(dev).repeat(128)
            out = torch.empty(masks.shape[0], dtype=torch.int16, device=dev)
            for _ in range(3):
                phe.eval7(masks, out=out)
            torch.cuda.synchronize()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(10):
                phe.eval7(masks, out=out)
            h1.record()
            torch.cuda.synchronize()
            ms = h0.elapsed_time(h1) / 10
            n = masks.shape[0]
            peaks = read_json(os.path.join(ROOT, "MEASURED_PEAKS.json")) or {}
            hbm = peaks.get("hbm_gbs", 6650.0)
            extras["eval7_hbm"] = {"value": n / (ms * 1e-3), "unit": "evals/s", "hands": n, "ms": ms,
                                   "roofline": {"bound": "hbm", "achieved": n * 10 / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                                "frac": n * 10 / (ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes_per_eval": 10,
                                                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}
            del masks, out
            # --- the reference's equity tasks end to end (simulate_processes replacement, default plans) ---
            from oracle import oracle as O
            from oracle import pyport
            rng = np.random.default_rng(0)
            D = phe.DECK
            tasks = []
            for gid in range(128):          # a finished heads-up game emits 4 rounds x 2 players = 8 tasks (env/env.py:119-147)
                cid = rng.permutation(52)[:9].tolist()
                for p in range(2):
                    hand = [D[i] for i in cid[2 * p: 2 * p + 2]]
                    flop, turn, river = [D[i] for i in cid[4:7]], [D[cid[7]]], [D[cid[8]]]
                    for rnd in range(4):
                        tasks.append((gid, f"p{p}", rnd, flop if rnd >= 1 else None, turn if rnd >= 2 else None, river if rnd >= 3 else None, hand))
            plans = {k: v for k, v in O.DEFAULT_PLANS.items()}
            phe.solve_tasks(tasks[:16], plans)
            t0 = time.perf_counter()
            res = phe.solve_tasks(tasks, plans)
            sdt = time.perf_counter() - t0
            deals = sum(r[3] for r in res)
            t0 = time.perf_counter()
            pdeals = 0
            for t in tasks[3:64:8]:          # eight river tasks through the pure-Python port, one core
                ids_ = [phe.card_id(c) for c in t[6] + t[3] + t[4] + t[5]]
                pdeals += pyport.run_plan(plans[3][0], ids_)[1]
            pdt = time.perf_counter() - t0
            extras["equity_service"] = {"value": len(tasks) / sdt, "unit": "tasks/s", "deals_per_s": deals / sdt, "tasks": len(tasks), "ms": sdt * 1e3,
                                        "config": {"workload": "reference task tuples (env/env.py:240) -> result tuples (env/workflow.py:225), default plans, 128 games"},
                                        "cpu_python_port": {"value": pdeals / pdt, "unit": "deals/s", "cores": 1,
                                                            "sample": f"{pdeals} deals of river tasks through oracle/pyport.run_plan"}}
            # --- policy / value leaf inference: default.yml model, 1024 observations per call, bf16 ---
            pv = phe.policy_value
            net = pv.fill_parameters(pv.PolicyValueNet(**pv.DEFAULT_MODEL_PARAMS), 1)
            pred = pv.BatchedPredictor(net, dtype=torch.bfloat16, max_batch=1024, buckets=[1, 1024])
            pred.warmup()
            obs = pv.random_observations(1024, seed=5)
            dobs = torch.from_numpy(obs).to(dev)
            pred.predict_batch(dobs)
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(10):
                pred.predict_batch(dobs)
            p1.record()
            torch.cuda.synchronize()
            pms = p0.elapsed_time(p1) / 10
            t0 = time.perf_counter()
            for _ in range(10):
                probs, val = pred.predict_batch(obs)              # numpy in, numpy out: pinned H2D + graph replay + D2H
            pe2e = (time.perf_counter() - t0) / 10
            t0 = time.perf_counter()
            for _ in range(50):
                pred.predict(obs[0])
            p1lat = (time.perf_counter() - t0) / 50
            extras["policy_value"] = {"value": 1024 / (pms * 1e-3), "unit": "observations/s", "ms_per_batch": pms, "dtype": "bf16",
                                      "config": {"workload": "TransformerAlphaGoZeroModel inference, config/default.yml dims (d=640, 8 layers, 98 tokens), "
                                                             "random-init weights, synthetic observations, batch 1024, CUDA graph"},
                                      "e2e": {"value": 1024 / pe2e, "unit": "observations/s", "h2d_bytes_per_step": int(obs.nbytes),
                                              "d2h_bytes_per_step": int(probs.nbytes + val.nbytes)},
                                      "predict_latency_us": p1lat * 1e6}
            del pred, net, dobs
        barrier()

    clocks = sampler.stop() if rank == 0 else None   # sampled through the timed steps and the Monte-Carlo section
    if rank == 0:
        # --- e2e: the public host-buffer API (what a reference-side caller binds), results land in numpy ---
        t0 = time.perf_counter()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(e2e_steps):
            h9, h = phe.enumerate_c52_7(b, e)       # rank 0's shard (the whole range at N=1)
        e2e_dt = (time.perf_counter() - t0) / e2e_steps
        e2e_val = (e - b) * world / e2e_dt if world == 1 else None
        int_peaks = read_json(os.path.join(ROOT, "profiles", "int_peaks.json")) or {}
        peak = int_peaks.get("balanced_lane_ops_per_s", NOMINAL_INT_PEAK)
        achieved = (e - b) * W_EVAL_LANE_OPS / (kern_ms * 1e-3)
        ncu = (read_json(os.path.join(ROOT, "profiles", "ncu_enum7_r01.json")) or [{}])[0]

        def _num(key):
            try:
                return float(str(ncu.get(key, "")).split()[0])
            except Exception:
                return None
        inst = _num("smsp__inst_executed.sum")
        hw = {"source": "profiles/ncu_enum7_r01.json (ncu --set full of this kernel, full range, 1 GPU)",
              "warp_inst_per_hand": inst / NUM_HANDS if inst else None,
              "issue_slot_util_pct": _num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
              "alu_pipe_pct": _num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
              "fma_pipe_pct": _num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
              "active_lanes_per_inst": _num("smsp__thread_inst_executed_per_inst_executed.ratio"),
              "smem_wavefronts": _num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
              "dram_bytes": 1e3 * (_num("dram__bytes_read.sum") or 0) + (_num("dram__bytes_write.sum") or 0)}
        line = {
            "metric": "hand_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": "exhaustive_c52_7", "hands_per_step": NUM_HANDS, "card_sets": "u64 bit sets -> u16 ranks", "outputs": "int64[9] + int64[7463] histograms",
                       "sharding": f"{world} interleaved tile sets of the colex combination range (phe_enum7_part)",
                       "collective": None if world == 1 else ("fused into k_enum7: peer-memory exchange of int64[7472] by the last CTA (phe_enum7_allreduce)"
                                                              if fused is not None else "NCCL all_reduce(int64[7472])"),
                       "l2": "256 MB L2 flush between timed steps; hands are generated on chip (no HBM-resident input)",
                       "launch": "step captured in a CUDA graph and replayed"},
            "roofline": {"bound": "int_issue", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "T lane-ops/s",
                         "frac": achieved / peak, "traffic": hw["dram_bytes"] or None, "algorithmic_lane_ops_per_eval": W_EVAL_LANE_OPS,
                         "kernel_ms": kern_ms, "kernel": "k_enum7 (this rank's shard; median of 10 un-graphed launches)",
                         "peak_source": "profiles/int_peaks.json (measured LOP3+IMAD stream)" if "balanced_lane_ops_per_s" in int_peaks
                         else "nominal 148 SM x 128 lane-ops/clk x 1.965 GHz (no measured integer peak yet)",
                         "nominal_peak": NOMINAL_INT_PEAK / 1e12,
                         "note": "frac is against SURVEY 8d's 150-lane-op model of the REFERENCE algorithm; this kernel needs ~28 warp "
                                 "instructions per 32 hands (sorted-multiset index, one table read per hand), so frac > 1 is the "
                                 "algorithmic saving, not skipped work: every step's histograms are checked against the golden fixture "
                                 "before timing. Hardware-side utilisation is in `hw`.",
                         "hw": hw},
            "cpu_baseline": cpu_baseline_enum(os.cpu_count() or 1) if world == 1 else None,
            "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": (9 + 7463) * 8,
                    "ms_per_step": e2e_dt * 1e3,
                    "note": "phe.enumerate_c52_7() -> phe_enum7_host: memset + kernel + D2H of both histograms + sync; the workload has no input buffer"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        # drop the graph (it holds NCCL kernels) before the communicator goes away; never block the exit on teardown
        del graph, fused
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
