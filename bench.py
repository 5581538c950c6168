#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric: MC equity rollouts/s at 1/2/4/8 B200; 7-card evals/s).

Default step (BASELINE config 3): Monte-Carlo equity of S = 4096 random six-max game states, R = 1,000,000 rollouts per
state, through k_equity_mc_smem.  With N GPUs the global sample range [0, R) of every state is split over the ranks and
the int64[S,3] (win, tie, loss) counts are all-reduced inside the step: total work is fixed ("strong" scaling) and the
counts are bit-identical for every N (checksum on the line).  Rank 0 prints ONE JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 3|2|1|5] [--skip-extras]

  --config 3   batched MC equity, 4096 x 6-max x 1e6                   (default, the line the driver scores)
  --config 2   exhaustive C(52,7) enumeration, both histograms         (round-1 headline; also the `enum7` side object)
  --config 1   single-state equity, 10,000 rollouts, hero As Ah        (latency through the host API)
  --config 5   sharded sweep, 64 six-max states x 1e9 rollouts         (also the `config5` side object when --full-sweep)

Objects on the line: `roofline` (integer-issue roofline of the dominant kernel: kernel-native model + the SURVEY 8d
reference model + ncu hardware counters), `cpu_baseline` (C oracle on the host cores), `e2e` (public host-buffer API:
numpy states in, numpy counts out, at every N), side objects `enum7`, `eval7_hbm`, `equity_service`, `policy_value`,
`mcts_selfplay` (BASELINE config 4: the reference's own MCTS on the drop-ins + the lock-step wave driver), `config1`.

--impl reference times the reference's OWN code (unmodified checkout in baseline/_ref + the phevaluator shim) on the
host cores for the same metric and config: see baseline/ref_arm.py.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_HANDS = 133784560
W_EVAL_LANE_OPS = 150                      # integer lane-ops per independent evaluation of the REFERENCE algorithm
NOMINAL_INT_PEAK = 148 * 128 * 1.965e9     # lane-ops/s, balanced alu+fma issue at max clock
MC_STATES, MC_PLAYERS, MC_ROLLOUTS = 4096, 6, 1_000_000
MC_SEED, STATE_SEED = 0x5EED20261019, 20261019
MC_CHECKSUM_1E6 = 622098221                # sum of win counts, 4096 states x 1e6 rollouts (identical at every N, round 1 and 2)


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


# ---------------------------------------------------------------------------------------------------------------------
# roofline helpers
# ---------------------------------------------------------------------------------------------------------------------
def int_peak():
    p = read_json(os.path.join(ROOT, "profiles", "int_peaks.json")) or {}
    if "balanced_lane_ops_per_s" in p:
        return float(p["balanced_lane_ops_per_s"]), "profiles/int_peaks.json (LOP3+IMAD stream measured on this pool's B200 by csrc/phe_peaks.cu)"
    return NOMINAL_INT_PEAK, "nominal 148 SM x 128 lane-ops/clk x 1.965 GHz (no measured integer peak)"


def kernel_model(name):
    """Executed-work model of THIS kernel's algorithm (tools/kernel_model.py from the committed ncu summaries)."""
    return (read_json(os.path.join(ROOT, "profiles", "kernel_models.json")) or {}).get(name, {})


def int_roofline(kernel, units_per_s, unit_name, kernel_ms, reference_lane_ops_per_unit):
    """achieved = useful lane-ops/s of this kernel's own algorithm (thread-level instructions ncu counted per unit x units/s);
    frac <= 1 by construction.  frac_vs_reference_model is the SURVEY 8d figure (> 1 = algorithmic saving)."""
    peak, src = int_peak()
    km = kernel_model(kernel)
    per_unit = km.get("lane_ops_per_unit")
    achieved = units_per_s * per_unit if per_unit else None
    return {"bound": "int_issue", "achieved": achieved / 1e12 if achieved else None, "peak": peak / 1e12, "unit": "T lane-ops/s",
            "frac": achieved / peak if achieved else None, "traffic": km.get("dram_bytes_per_launch"),
            "kernel": kernel, "kernel_ms": kernel_ms, "peak_source": src, "nominal_peak": NOMINAL_INT_PEAK / 1e12,
            "kernel_model": {"lane_ops_per_" + unit_name: per_unit, "warp_inst_per_32_" + unit_name + "s": km.get("warp_inst_per_32_units"),
                             "smem_wavefronts_per_32_" + unit_name + "s": km.get("smem_wavefronts_per_32_units"), "source": km.get("source")},
            "frac_vs_reference_model": units_per_s * reference_lane_ops_per_unit / peak,
            "reference_model_lane_ops_per_" + unit_name: reference_lane_ops_per_unit,
            "hw": km.get("hw"),
            "note": "frac = thread-level instructions this kernel executes per " + unit_name + " (ncu, active lanes) x measured " + unit_name +
                    "s/s / measured LOP3+IMAD peak; frac_vs_reference_model uses SURVEY 8d's lane-op count of the REFERENCE algorithm and "
                    "exceeds frac by the algorithmic saving. Algorithmic HBM bytes are ~0 (inputs 16 B/state, hands generated on chip)."}


def w_rollout_reference_model(states_np, n_opp):
    """W_rollout(P, b) = P*150 + 70*ceil(d/4) + 12*d + 4*P, d = dealt cards, averaged over the states."""
    import numpy as np
    nb = np.bitwise_count(states_np[:, 0].astype(np.uint64)).astype(np.int64) - 2
    d = (5 - nb) + 2 * n_opp
    P = n_opp + 1
    return float(np.mean(P * W_EVAL_LANE_OPS + 70 * np.ceil(d / 4) + 12 * d + 4 * P))


# ---------------------------------------------------------------------------------------------------------------------
# CPU baselines (C oracle: test infrastructure, used here only as the timed CPU port)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_baseline_enum(threads):
    from oracle import oracle as O
    O.lib()
    t0 = time.perf_counter()
    h9, h = O.enum7(0, NUM_HANDS, nthreads=threads)
    dt = time.perf_counter() - t0
    assert h9.tolist() == O.K1_HIST9
    return {"value": NUM_HANDS / dt, "unit": "evals/s", "cores": threads, "kind": "port",
            "sample": f"all {NUM_HANDS} hands, C port of the phevaluator algorithm (oracle/phe_oracle.c), {threads} threads, {dt:.2f} s"}


def cpu_baseline_mc(states_np, n_opp, threads, rollouts=4096):
    """C oracle Monte-Carlo (same Philox stream, same dealing contract) over the host cores: every state, `rollouts` each."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.lib()
    S = len(states_np)

    def one(s):
        return O.equity_mc(int(states_np[s, 0]), int(states_np[s, 1]), n_opp, rollouts, seed=MC_SEED, state_idx=s)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(one, range(S)))
    dt = time.perf_counter() - t0
    return {"value": S * rollouts / dt, "unit": "rollouts/s", "cores": threads, "kind": "port",
            "sample": f"the same {S} six-max states x {rollouts} rollouts each (first {rollouts} samples of the shared Philox stream), "
                      f"C oracle oracle/phe_oracle.c:orc_equity_mc, {threads} threads, {dt:.2f} s"}, res


# ---------------------------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------------------------
def mc_config(world):
    return {"workload": "batched_mc_equity", "states": MC_STATES, "players": MC_PLAYERS, "rollouts_per_state": MC_ROLLOUTS,
            "state_seed": STATE_SEED, "sharding": f"sample index range split over {world} rank(s) + all_reduce(int64[{MC_STATES},3])",
            "l2": "inputs are 64 KB of game states; hands are generated on chip (no HBM-resident input); 256 MB L2 flush between timed steps",
            "rng": "Philox4x32-10, key = seed, counter = (sample lo, sample hi, state index, block)"}


def reference_arm(args):
    """The reference's own CPU implementation (baseline/_ref + shim), all host cores, rank 0 only.  See baseline/ref_arm.py."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np
    from baseline import ref_arm as R
    R.setup_paths()
    from texasholdempocker_rl_b200.equity import random_states          # the synthetic-state generator only (inputs, not compute)
    cores = os.cpu_count() or 1
    steps, warm = args.steps, args.warmup
    extras = {}
    if args.config == 2:
        # exhaustive enumeration has no reference-side loop; its evaluator is timed through get_winner_set_v2 (2 evaluations per call)
        pool = R.Pool(cores)
        for _ in range(warm):
            R.winner_set_rate(pool, 2000, cores)
        per = 20000
        t0 = time.perf_counter()
        for _ in range(steps):
            R.winner_set_rate(pool, per, cores)
        dt = time.perf_counter() - t0
        val = 2 * per * cores * steps / dt
        pool.close()
        metric, unit, config = "hand_evals_per_s", "evals/s", {"workload": "exhaustive_c52_7", "hands_per_step": NUM_HANDS}
        sample = f"{2 * per * cores} evaluations per step: GameEnv.get_winner_set_v2 (env/game.py:663-681) on random heads-up deals, {cores} processes"
    else:
        S = MC_STATES if args.config == 3 else (64 if args.config == 5 else 1)
        if args.config == 1:
            from texasholdempocker_rl_b200.equity import pack_states
            st = pack_states([["As", "Ah"]], [[]]).view(np.uint64)
            n_opp, per_state = 1, 10000
            config = {"workload": "single_state_equity", "hero": "As Ah", "board": "", "players": 2, "rollouts": 10000}
        else:
            st = random_states(S, STATE_SEED + (1 if args.config == 5 else 0)).view(np.uint64)
            n_opp = MC_PLAYERS - 1
            per_state = max(1, round(45000 * cores / S))                  # ~1 s of work per step on this box's cores
            config = mc_config(max(1, args.gpus)) if args.config == 3 else {"workload": "sharded_mc_equity_sweep", "states": 64, "players": 6,
                                                                           "rollouts_per_state": 10**9}
        states = [(int(a), int(b)) for a, b in st]
        pool = R.Pool(min(cores, len(states)))
        for _ in range(warm):
            pool.mc6(states, 1, n_opp, 1)
        t0 = time.perf_counter()
        tot = [0, 0, 0]
        for k in range(steps):
            w, t, l = pool.mc6(states, per_state, n_opp, 1000 * (k + 1))
            tot = [tot[0] + w, tot[1] + t, tot[2] + l]
        dt = time.perf_counter() - t0
        val = len(states) * per_state * steps / dt
        extras["counts"] = {"win": tot[0], "tie": tot[1], "loss": tot[2], "equity": (tot[0] + 0.5 * tot[1]) / max(1, sum(tot))}
        metric, unit = "mc_equity_rollouts_per_s", "rollouts/s"
        sample = (f"{len(states)} states x {per_state} rollouts per step ({n_opp + 1} players): random.sample dealing (env/workflow.py:190) + "
                  f"GameEnv.get_winner_set_v2 (env/game.py:663-681) + the +1/0/-1 mapping of env/workflow.py:142-147, {pool.procs} processes")
        if not args.skip_extras:
            # get_winner_set_v2 in a loop, simulate_processes on the config-1 and default plans
            extras["get_winner_set_v2"] = {"evals_per_s_all_cores": R.winner_set_rate(pool, 20000, pool.procs), "evals_per_s_1_core": R.winner_set_rate(pool, 20000, 1),
                                           "unit": "evals/s (2 per call)", "cores": pool.procs}
            pool.close()
            params = R.reference_params()
            plans_cfg1 = {0: ([(7, 10000, True)], 10000), 1: ([(4, 10000, True)], 10000)}
            tasks = [t for t in R.random_tasks(max(2, cores // 2), (0, 1), 5)]
            res, sdt = R.simulate_processes_rate(tasks, plans_cfg1, cores)
            extras["simulate_processes_config1"] = {"deals_per_s": sum(r[3] for r in res) / sdt, "tasks": len(tasks), "workers": cores, "seconds": sdt,
                                                    "plans": "[(7,10000,True)] / [(4,10000,True)] (config/debug.yml shape, 10,000 rollouts)"}
            dplans = params_default_plans(R)
            tasks = R.random_tasks(max(1, cores // 4), (2, 3), 6)
            res, sdt = R.simulate_processes_rate(tasks, dplans, cores)
            extras["simulate_processes_default_plans"] = {"deals_per_s": sum(r[3] for r in res) / sdt, "tasks": len(tasks), "workers": cores, "seconds": sdt,
                                                          "plans": "config/default.yml rounds 2 and 3 (45,540 / 990 deals, exhaustive)"}
            # BASELINE config 4: the reference's own SingleThreadMCTS, stock path, one process (it is single-threaded by design)
            import torch
            torch.set_num_threads(1)
            p, mp_, model, env = R.build_model_and_env()
            sims, mdt, moves = R.mcts_moves(env, lambda: R.make_search("reference", p, mp_, model, 100), 6, 3)
            extras["mcts_selfplay"] = {"simulations_per_s": sims / mdt, "moves": moves, "simulations_per_move": 100, "seconds": mdt, "cores": 1,
                                       "what": "rl.MCTS.SingleThreadMCTS.simulate + Env (config/debug.yml model dims, heads-up 25/50), CPU model, shim evaluator"}
        else:
            pool.close()
    line = {
        "impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "scaffolding+shim", "sample": sample,
                         "note": "unmodified reference code from baseline/_ref; `phevaluator` (PyPI, not installable offline) is the restated "
                                 "pure-Python algorithm in oracle/shim"},
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line.update(extras)
    print(json.dumps(line))


def params_default_plans(R):
    import yaml
    cfg = yaml.safe_load(open(os.path.join(R.setup_paths(), "config", "default.yml"), encoding="utf-8"))
    return cfg["simulation_recurrent_params"]


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 5])
    ap.add_argument("--mc-rollouts", type=int, default=MC_ROLLOUTS, help="rollouts per state of config 3 (the line is only valid at the default)")
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
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(step, steps):
        """K steps, CUDA events around each (L2 flush outside the events), barrier + synchronize on both sides; sum, max over ranks."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for s0, s1 in ev:
            flush_buf.fill_(1)
            s0.record()
            step()
            s1.record()
        barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))

    # ---- Monte-Carlo equity (configs 3, 5, 1) --------------------------------------------------------------------
    def mc_measure(S, R, state_seed, steps, n_opp=MC_PLAYERS - 1, states_np=None, kernel_reps=5):
        st_np = states_np if states_np is not None else phe.random_states(S, state_seed)
        states = torch.from_numpy(st_np).to(dev)
        rb, re_ = phe.shard_range(R, rank, world)
        wtl = torch.zeros(S, 3, dtype=torch.int64, device=dev)

        def step():
            wtl.zero_()
            phe.equity_mc(states, n_opp, re_ - rb, seed=MC_SEED, sample_offset=rb, out=wtl)
            if world > 1:
                dist.all_reduce(wtl)
        for _ in range(warm if steps > 2 else 1):
            step()
        launches0 = phe._native.launch_count()
        total_ms = timed_steps(step, steps)
        launches = phe._native.launch_count() - launches0
        counts = wtl.cpu().numpy()
        assert (counts.sum(1) == R).all(), "win + tie + loss must equal the rollouts of every state"
        # the kernel alone (this rank's shard) on its launching stream
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(kernel_reps)]
        scratch = torch.zeros_like(wtl)
        for k0, k1 in kev:
            k0.record()
            phe.equity_mc(states, n_opp, re_ - rb, seed=MC_SEED, sample_offset=rb, out=scratch)
            k1.record()
        barrier()
        kern_ms = sorted(a.elapsed_time(b) for a, b in kev)[len(kev) // 2]
        # end to end through the public host-buffer API: numpy states in (H2D), numpy counts out (D2H), every rank
        e2e_reps = 2 if R >= 10**8 else 3
        phe.sharded_equity_mc(st_np, n_opp, min(R, 100000), MC_SEED)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_reps):
            hw = phe.sharded_equity_mc(st_np, n_opp, R, MC_SEED)
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_reps)
        assert (hw == counts).all(), "host-API counts differ from the device-resident run"
        return {"states": st_np, "counts": counts, "total_ms": total_ms, "launches": launches, "kern_ms": kern_ms, "shard_rollouts": S * (re_ - rb),
                "value": S * R * steps / (total_ms * 1e-3), "e2e": {"value": S * R / e2e_s, "unit": "rollouts/s", "h2d_bytes_per_step": int(st_np.nbytes),
                                                                    "d2h_bytes_per_step": int(hw.nbytes), "ms_per_step": e2e_s * 1e3,
                                                                    "api": "phe.sharded_equity_mc(numpy states) -> numpy counts on every rank: H2D, "
                                                                           "k_equity_mc_smem on the rank's sample shard, all_reduce, D2H"}}

    # ---- exhaustive enumeration (config 2) ----------------------------------------------------------------------------
    def enum_measure(steps):
        hist = torch.zeros(9 + 7463, dtype=torch.int64, device=dev)
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
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "hist7463.json")))

        def check(what):
            hh = hist.cpu().numpy()
            assert hh[:9].tolist() == gold["hist9"] and hh[9:9 + 7463].tolist() == gold["hist7463"], f"{what}: histogram differs from the golden fixture"
        for _ in range(warm):
            step()
        barrier()
        check("eager step")
        # ~0.1 ms of GPU work behind several host calls: captured once in a CUDA graph and replayed
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
        check("graph replay")
        total_ms = timed_steps(graph.replay, steps)
        if fused is not None:
            fused.check()                      # raises if a peer missed an exchange (the histogram of that step would be partial)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        scratch = torch.zeros(9 + 7463, dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream()
        for k0, k1 in kev:
            scratch.zero_()
            k0.record()
            phe._native.check(L.phe_enum7_part(C.c_void_p(scratch.data_ptr()), C.c_void_p(scratch.data_ptr() + 72), 0, NUM_HANDS, rank, world, C.c_void_p(stream.cuda_stream)))
            k1.record()
        barrier()
        kern_ms = sorted(a.elapsed_time(b) for a, b in kev)[len(kev) // 2]
        e2e = None
        if world == 1:
            t0 = time.perf_counter()
            for _ in range(5):
                phe.enumerate_c52_7()
            e2e_dt = (time.perf_counter() - t0) / 5
            e2e = {"value": NUM_HANDS / e2e_dt, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": (9 + 7463) * 8, "ms_per_step": e2e_dt * 1e3,
                   "api": "phe.enumerate_c52_7() -> phe_enum7_host: memset + kernel + D2H of both histograms; the workload has no input buffer"}
        out = {"value": NUM_HANDS * steps / (total_ms * 1e-3), "unit": "evals/s", "ms_per_step": total_ms / steps, "kern_ms": kern_ms, "e2e": e2e,
               "collective": None if world == 1 else (("fused into k_enum7: last CTA exchanges the 30 KB u32 histogram, " + ("NVLS multimem.red.add + multimem.st flag (phe_enum7_allreduce_nvls)" if fused.nvls else "peer stores + release flags (phe_enum7_allreduce)")) if fused is not None
                                                      else "NCCL all_reduce(int64[7472])"),
               "hands_this_rank": NUM_HANDS // world}
        del graph
        return out, fused

    clocks = None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    extras = {}
    fused_keepalive = None

    if args.config in (3, 1, 5):
        if args.config == 3:
            S, R, sseed, steps, st_np, n_opp = MC_STATES, args.mc_rollouts, STATE_SEED, args.steps, None, MC_PLAYERS - 1
            config = mc_config(world)
        elif args.config == 5:
            S, R, sseed, steps, st_np, n_opp = 64, 10**9, STATE_SEED + 1, min(args.steps, 3), None, MC_PLAYERS - 1
            config = {"workload": "sharded_mc_equity_sweep", "states": 64, "players": 6, "rollouts_per_state": 10**9, "state_seed": sseed,
                      "sharding": f"sample index range split over {world} rank(s) + all_reduce(int64[64,3])", "l2": "no HBM-resident input (on-chip dealing)"}
        else:
            S, R, sseed, steps, n_opp = 1, 10000, None, args.steps, 1
            st_np = phe.pack_states([["As", "Ah"]], [[]])
            config = {"workload": "single_state_equity", "hero": "As Ah", "board": "", "players": 2, "rollouts": 10000, "seed": STATE_SEED,
                      "note": "latency-bound: one small launch (k_equity_mc_gmem); SURVEY 8d config 1"}
        m = mc_measure(S, R, sseed, steps, n_opp=n_opp, states_np=st_np)
        checksum = int(m["counts"][:, 0].sum())
        if args.config == 3 and R == MC_ROLLOUTS:
            assert checksum == MC_CHECKSUM_1E6, f"win-count checksum {checksum} != {MC_CHECKSUM_1E6}"
        per_rank_units_per_s = m["shard_rollouts"] / (m["kern_ms"] * 1e-3)
        roof = int_roofline("k_equity_mc_smem" if args.config != 1 else "k_equity_mc_gmem", per_rank_units_per_s, "rollout", m["kern_ms"],
                            w_rollout_reference_model(m["states"], n_opp))
        value, total_ms, launches, e2e = m["value"], m["total_ms"], m["launches"], m["e2e"]
        metric, unit = "mc_equity_rollouts_per_s", "rollouts/s"
        extras["checksum_wins"] = checksum
        extras["equity_state0"] = float(phe.win_probability(m["counts"][0]))
        cpu = None
        if rank == 0 and world == 1:
            cpu, cres = cpu_baseline_mc(m["states"], n_opp, os.cpu_count() or 1, rollouts=4096 if args.config != 5 else 262144)
            if args.config != 1:
                # the C oracle's counts for those first samples equal the device's (shared stream): one more parity check on the box
                k = min(S, 64)
                got = phe.equity_mc(m["states"][:k], n_opp, 4096 if args.config != 5 else 262144, MC_SEED)
                assert all((got[s] == cres[s]).all() for s in range(k)), "device counts differ from the C oracle on the shared stream"
    else:
        em, fused_keepalive = enum_measure(args.steps)
        value, total_ms, launches, e2e = em["value"], em["ms_per_step"] * args.steps, args.steps, em["e2e"] or {"value": None, "unit": "evals/s"}
        metric, unit = "hand_evals_per_s", "evals/s"
        config = {"workload": "exhaustive_c52_7", "hands_per_step": NUM_HANDS, "outputs": "int64[9] + int64[7463] histograms",
                  "sharding": f"{world} interleaved tile sets of the colex range (phe_enum7_part)", "collective": em["collective"],
                  "l2": "256 MB L2 flush between timed steps; hands are generated on chip", "launch": "step captured in a CUDA graph and replayed"}
        roof = int_roofline("k_enum7", em["hands_this_rank"] / (em["kern_ms"] * 1e-3), "hand", em["kern_ms"], W_EVAL_LANE_OPS)
        cpu = cpu_baseline_enum(os.cpu_count() or 1) if rank == 0 and world == 1 else None

    if not args.skip_extras and args.config == 3:
        # ---- config 2 as a side object at every N --------------------------------------------------------------------------
        # the side objects must never cost the headline line: a Python-level failure in one of them is reported in its place
        try:
            em, fused_keepalive = enum_measure(min(args.steps, 20))
            extras["enum7"] = {"value": em["value"], "unit": "evals/s", "ms_per_step": em["ms_per_step"], "kernel_ms": em["kern_ms"], "e2e": em["e2e"],
                               "config": {"workload": "exhaustive_c52_7 (BASELINE config 2)", "hands_per_step": NUM_HANDS, "collective": em["collective"]},
                               "roofline": int_roofline("k_enum7", em["hands_this_rank"] / (em["kern_ms"] * 1e-3), "hand", em["kern_ms"], W_EVAL_LANE_OPS)}
        except Exception as ex:  # noqa: BLE001
            extras["enum7"] = {"error": f"{type(ex).__name__}: {ex}"}
        if rank == 0 and world == 1:
            try:
                extras.update(single_gpu_extras(phe, torch, np, dev))
            except Exception as ex:  # noqa: BLE001
                extras["single_gpu_extras_error"] = f"{type(ex).__name__}: {ex}"
        barrier()
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps if args.config != 5 else min(args.steps, 3), "warmup": warm,
            "ms_per_step": total_ms / (args.steps if args.config != 5 else min(args.steps, 3)), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        # drop graphs / symmetric memory before the communicator goes away; never block the exit on teardown
        del fused_keepalive
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def single_gpu_extras(phe, torch, np, dev):
    """Single-GPU side measurements (reported at N = 1 only)."""
    out = {}
    # --- config 1: single-state equity latency through the host API ---
    st1 = phe.pack_states([["As", "Ah"], ["As", "Ah"]], [[], ["2c", "7d", "Tc"]])
    phe.equity_mc(st1[:1], 1, 10000, STATE_SEED)
    t0 = time.perf_counter()
    for _ in range(50):
        c1 = phe.equity_mc(st1[:1], 1, 10000, STATE_SEED)
    lat = (time.perf_counter() - t0) / 50
    c2 = phe.equity_mc(st1[1:], 1, 10000, STATE_SEED)
    out["config1"] = {"latency_us": lat * 1e6, "rollouts": 10000, "equity_preflop_AsAh": float(phe.win_probability(c1[0])),
                      "equity_flop_AsAh_2c7dTc": float(phe.win_probability(c2[0])), "counts_preflop": c1[0].tolist(),
                      "config": "SURVEY 8d config 1: heads-up, hero As Ah, 10,000 rollouts, numpy in / numpy out (one k_equity_mc_gmem launch + sync)"}
    # --- materialised-hands micro-config: 2^25 card sets (268 MB > L2) -> ranks ---
    g = torch.Generator(device="cpu").manual_seed(1)
    base = torch.rand(1 << 18, 52, generator=g).argsort(1)[:, :7].numpy().astype(np.uint8)
    masks = torch.from_numpy(phe.ids_to_masks(base)).to(dev).repeat(128)
    res = torch.empty(masks.shape[0], dtype=torch.int16, device=dev)
    for _ in range(3):
        phe.eval7(masks, out=res)
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(10):
        phe.eval7(masks, out=res)
    h1.record()
    torch.cuda.synchronize()
    ms = h0.elapsed_time(h1) / 10
    n = masks.shape[0]
    peaks = read_json(os.path.join(ROOT, "MEASURED_PEAKS.json")) or {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    out["eval7_hbm"] = {"value": n / (ms * 1e-3), "unit": "evals/s", "hands": n, "ms": ms,
                        "roofline": {"bound": "hbm", "achieved": n * 10 / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                     "frac": n * 10 / (ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes_per_eval": 10,
                                     "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}
    ids_dev = torch.from_numpy(base).to(dev).repeat(128, 1)
    for _ in range(2):
        phe.eval7(ids_dev, out=res)
    torch.cuda.synchronize()
    h0.record()
    for _ in range(5):
        phe.eval7(ids_dev, out=res)
    h1.record()
    torch.cuda.synchronize()
    ms_ids = h0.elapsed_time(h1) / 5
    out["eval7_hbm"]["ids_layout"] = {"value": n / (ms_ids * 1e-3), "unit": "evals/s", "ms": ms_ids, "gbs": n * 9 / (ms_ids * 1e-3) / 1e9,
                                      "algorithmic_bytes_per_eval": 9, "layout": "uint8[N,7] card ids (what evaluate_cards(*ids) callers hold)"}
    del masks, res, ids_dev
    # --- the reference's equity tasks end to end (simulate_processes replacement, default plans) ---
    from oracle import oracle as O
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
    res_t = phe.solve_tasks(tasks, plans)
    sdt = time.perf_counter() - t0
    deals = sum(r[3] for r in res_t)
    out["equity_service"] = {"value": len(tasks) / sdt, "unit": "tasks/s", "deals_per_s": deals / sdt, "tasks": len(tasks), "ms": sdt * 1e3,
                             "config": {"workload": "reference task tuples (env/env.py:240) -> result tuples (env/workflow.py:225), default plans, 128 games"}}
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
    out["policy_value"] = {"value": 1024 / (pms * 1e-3), "unit": "observations/s", "ms_per_batch": pms, "dtype": "bf16",
                           "config": {"workload": "TransformerAlphaGoZeroModel inference, config/default.yml dims (d=640, 8 layers, 98 tokens), "
                                                  "random-init weights, synthetic observations, batch 1024, CUDA graph"},
                           "e2e": {"value": 1024 / pe2e, "unit": "observations/s", "h2d_bytes_per_step": int(obs.nbytes),
                                   "d2h_bytes_per_step": int(probs.nbytes + val.nbytes)},
                           "predict_latency_us": p1lat * 1e6}
    del pred, net, dobs
    # --- BASELINE config 4: MCTS self-play on this engine (reference tree + betting engine, device leaf stages) ---
    try:
        out["mcts_selfplay"] = mcts_selfplay(phe, torch)
    except Exception as ex:  # noqa: BLE001  (the reference checkout may be absent: say so instead of failing the line)
        out["mcts_selfplay"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
    return out


def _farm_worker(rank, barrier, queue, n_simulation, moves):
    """One self-play process of the farm: its own CUDA context on cuda:0, the wave driver on the reference's Env."""
    import torch
    import texasholdempocker_rl_b200 as phe
    from baseline import ref_arm as R
    from texasholdempocker_rl_b200.mcts_wave import GpuBackends, WaveMCTS
    R.setup_paths()
    from env.constants import ChoiceMethod
    torch.set_num_threads(1)
    torch.cuda.set_device(0)
    params, mp_, model, env = R.build_model_and_env(seed=rank)
    net = phe.PolicyValueNet(**mp_)
    net.load_reference_state_dict(model.state_dict())
    predictor = phe.BatchedPredictor(net, dtype=torch.float32, max_batch=64, buckets=[1, 8, 64])
    predictor.warmup()
    backends = GpuBackends(predictor, phe.ObsEncoder(mp_["num_output_class"], mp_["historical_action_per_round"]), seed=STATE_SEED + rank)

    def make_wave():
        return WaveMCTS(mp_["num_output_class"], params["small_blind"], backends, n_simulation=n_simulation, c_puct=params["mcts_c_puct"],
                        tau=params["mcts_tau"], model_Q_epsilon=params["mcts_model_Q_epsilon"], choice_method=ChoiceMethod(params["mcts_choice_method"]), wave=64)
    R.mcts_moves(env, make_wave, 1, 100 + rank)                  # warm-up move: kernels loaded, graphs captured
    barrier.wait()
    t0 = time.perf_counter()
    sims, dt, mv = R.mcts_moves(env, make_wave, moves, 200 + rank)
    queue.put((rank, sims, t0, t0 + dt))


def mcts_farm(workers, n_simulation=800, moves=3):
    """The reference runs its game loops in many processes (main.py:140-147); so does this: `workers` self-play processes
    share ONE GPU, each with the lock-step wave driver.  Aggregate simulations/s = all simulations / (last finish - first start)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    barrier, queue = ctx.Barrier(workers), ctx.Queue()
    procs = [ctx.Process(target=_farm_worker, args=(r, barrier, queue, n_simulation, moves), daemon=True) for r in range(workers)]
    for p in procs:
        p.start()
    rows = []
    try:
        for _ in procs:
            rows.append(queue.get(timeout=300))
    finally:
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
    sims = sum(r[1] for r in rows)
    span = max(r[3] for r in rows) - min(r[2] for r in rows)
    return {"workers": workers, "simulations_per_s": sims / span, "simulations": sims, "seconds": span, "moves_per_worker": moves,
            "what": "self-play processes sharing one B200 (own CUDA context each), WaveMCTS with 64 simulations per wave, debug.yml parameters"}


def mcts_selfplay(phe, torch, n_simulation=800, moves=4):
    """config/debug.yml game parameters (800 simulations per move, heads-up, blinds 25/50, debug model dims).
    Three searches over the same seeded self-play: (a) the unpatched reference on the host (CPU model, shim evaluator),
    (b) the reference's SingleThreadMCTS with every drop-in installed (one-at-a-time leaves through the device),
    (c) the lock-step wave driver (64 simulations per wave).  Reports simulations/s, leaf showdowns/s, NN rows/s."""
    from baseline import ref_arm as R
    from texasholdempocker_rl_b200 import dropin
    from texasholdempocker_rl_b200.mcts_wave import GpuBackends, WaveMCTS
    R.setup_paths()
    import env.game as game_mod
    from env.constants import ChoiceMethod
    torch.set_num_threads(1)
    params, mp_, model, env = R.build_model_and_env()
    net = phe.PolicyValueNet(**mp_)
    net.load_reference_state_dict(model.state_dict())
    predictor = phe.BatchedPredictor(net, dtype=torch.float32, max_batch=64, buckets=[1, 8, 64])
    predictor.warmup()
    encoder = phe.ObsEncoder(mp_["num_output_class"], mp_["historical_action_per_round"])
    out = {"config": {"workload": "MCTS self-play (BASELINE config 4)", "simulations_per_move": n_simulation, "players": 2, "blinds": "25/50",
                      "model": "config/debug.yml dims (d=96, 2 layers), random init, fp32", "moves_timed": moves}}
    terminals = [0]
    stock_finish = game_mod.GameEnv.finish_game
    # (a) unpatched reference, fewer simulations per move (it is ~100x slower; the rate is what is reported)
    _, _, _, env_a = R.build_model_and_env()
    sims, dt, mv = R.mcts_moves(env_a, lambda: R.make_search("reference", params, mp_, model, max(50, n_simulation // 8)), 2, 3)
    out["reference_cpu"] = {"simulations_per_s": sims / dt, "simulations_per_move": max(50, n_simulation // 8), "moves": mv, "cores": 1,
                            "what": "unpatched rl.MCTS.SingleThreadMCTS, CPU model, shim evaluator (baseline/_ref)"}
    # (b) the same search with every drop-in installed
    dropin.install(predictor=predictor, encoder=encoder, settle=True)
    try:
        settle_finish = game_mod.GameEnv.finish_game            # the drop-in (phe_settle)

        def counting_finish(self):
            terminals[0] += 1
            return settle_finish(self)
        game_mod.GameEnv.finish_game = counting_finish
        r0 = predictor.replays
        _, _, _, env_b = R.build_model_and_env()
        sims, dt, mv = R.mcts_moves(env_b, lambda: R.make_search("reference", params, mp_, model, n_simulation // 4), 2, 3)
        out["dropin_sequential"] = {"simulations_per_s": sims / dt, "leaf_showdowns_per_s": terminals[0] / dt, "nn_calls_per_s": (predictor.replays - r0) / dt,
                                    "simulations_per_move": n_simulation // 4, "moves": mv,
                                    "what": "reference SingleThreadMCTS + Env with dropin.install(predictor, encoder, settle=True): one GPU round trip per leaf stage"}
    finally:
        game_mod.GameEnv.finish_game = settle_finish
        dropin.uninstall()
    assert game_mod.GameEnv.finish_game is stock_finish
    # (c) lock-step waves
    backends = GpuBackends(predictor, encoder, seed=STATE_SEED)
    stats = WaveMCTS(mp_["num_output_class"], params["small_blind"], backends).stats          # one counter set shared by every move's search

    def make_wave():
        return WaveMCTS(mp_["num_output_class"], params["small_blind"], backends, n_simulation=n_simulation, c_puct=params["mcts_c_puct"],
                        tau=params["mcts_tau"], model_Q_epsilon=params["mcts_model_Q_epsilon"],
                        choice_method=ChoiceMethod(params["mcts_choice_method"]), wave=64, stats=stats)
    _, _, _, env_c = R.build_model_and_env()
    sims, dt, mv = R.mcts_moves(env_c, make_wave, moves, 3)
    out["wave"] = {"simulations_per_s": sims / dt, "leaf_showdowns_per_s": stats["terminals"] / dt, "nn_rows_per_s": stats["nn_rows"] / dt,
                   "nn_calls_per_s": stats["nn_calls"] / dt, "wave": 64, "simulations_per_move": n_simulation, "moves": mv,
                   "seconds_by_stage": {k[2:]: round(v, 4) for k, v in stats.items() if k.startswith("t_")},
                   "what": "mcts_wave.WaveMCTS: reference tree nodes + GameEnv.step on the host, determinise / encode / predict_batch / settle_envs batched per wave"}
    try:
        out["farm"] = mcts_farm(min(8, max(2, (os.cpu_count() or 2) // 2)))
    except Exception as ex:  # noqa: BLE001
        out["farm"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
    out["value"] = out["wave"]["simulations_per_s"]
    out["unit"] = "simulations/s"
    out["speedup_wave_vs_reference_cpu"] = out["wave"]["simulations_per_s"] / out["reference_cpu"]["simulations_per_s"]
    out["speedup_wave_vs_dropin_sequential"] = out["wave"]["simulations_per_s"] / out["dropin_sequential"]["simulations_per_s"]
    return out


if __name__ == "__main__":
    main()
