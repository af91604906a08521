#!/usr/bin/env python
"""bench.py — self-play hot path throughput on BASELINE.json configs[1]: 9x9 Go AlphaZero, 400 simulations,
256 parallel games per GPU, 6-block x 256-channel network, random-init weights.

A "step" is one whole move search for every game of this rank: (S+1) = 401 cycles of select -> leaf transition ->
features -> network -> expand/backup for 256 games = 102 656 leaf evaluations, followed by the move itself.
Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for every field.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GAMES, SIMS, BOARD = 256, 400, 9
BLOCKS, HIDDEN, VALUE_HIDDEN, ACTIONS, IN_CH = 6, 256, 256, 82, 18
NET = os.path.join(ROOT, "oracle", "_ref", "nets", "go9_az_6bx256.pt")
FLOPS_PER_EVAL = 1.1535e9          # 2*MACs of convs + FCs of one position
FLOPS_PER_CONV_LAUNCH = 95551488.0 * GAMES  # one hidden->hidden 3x3 conv over 256 positions (2*81*256*256*9 each)
CONFIG = {"workload": "go9x9_alphazero_400sims_256games_6bx256 (BASELINE configs[1])", "games_per_gpu": GAMES, "simulations": SIMS, "board": "9x9",
          "net": "6bx256 fp16 tensor-core / fp32 accumulate", "evals_per_step_per_gpu": GAMES * (SIMS + 1),
          "l2": "per-step working set (node pools 219 MB + activations) exceeds the 126 MB L2; no explicit flush"}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one tower launch from the committed `ncu --set full` capture (profiles/)."""
    best = None
    for name in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
        if name.endswith("_ncu_full_summary.json"):
            best = os.path.join(ROOT, "profiles", name)
    try:
        with open(best) as f:
            return json.load(f)["tower"]["dram_traffic_bytes_per_launch"], os.path.basename(best)
    except Exception:
        return None, None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"], "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    except Exception:
        return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])), mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def reference_conf(threads):
    return (f"env_board_size={BOARD}:actor_num_simulation={SIMS}:zero_num_parallel_games={GAMES}:zero_num_threads={threads}:nn_type_name=alphazero:"
            f"nn_file_name={NET}:program_seed=1:program_auto_seed=false:program_quiet=true")


def run_reference(warm_cycles, cycles):
    """The UNMODIFIED reference actor path (oracle/_ref/ref_actor_group_go: reference sources compiled in place, network on
    the CPU through libtorch, tree/env on zero_num_threads host threads), timed over whole ActorGroup cycles."""
    binary = os.path.join(ROOT, "oracle", "_ref", "ref_actor_group_go")
    cores = os.cpu_count() or 1
    if not (os.path.exists(binary) and os.path.exists(NET)):
        return None
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([binary, "bench", reference_conf(cores), str(warm_cycles), str(cycles), "-1"], capture_output=True, text=True, env=env, timeout=3000)
    m = re.search(r"REFBENCH evals=(\d+) seconds=([0-9.eE+-]+) threads=(\d+)", res.stdout)
    if not m:
        raise RuntimeError("reference bench produced no REFBENCH line: " + res.stdout[-300:] + res.stderr[-300:])
    evals, secs = int(m.group(1)), float(m.group(2))
    return {"value": evals / secs, "unit": "leaf-evals/s", "cores": cores, "kind": "reference", "seconds": secs,
            "sample": f"{cycles} ActorGroup cycles x {GAMES} games = {evals} leaf evaluations of the same workload (after {warm_cycles} warm-up cycles), all-CPU"}


def port_baseline(budget_evals):
    """Fallback when oracle/_ref is absent: the C restatement for the tree/env work + the TorchScript-free fp32 C network is far
    too slow at 6bx256, so only the tree/env part is timed and the sample says so."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    lib = oracle_lib.load()
    orc = oracle_lib.OracleSearch(lib, oracle_lib.GAME_GO, BOARD, 16, SIMS)
    rng = np.random.default_rng(0)
    pol = rng.dirichlet([1.0] * ACTIONS, size=16).astype(np.float32)
    lg, val = np.log(pol), np.zeros(16, np.float32)
    t0, n = time.perf_counter(), 0
    while n < budget_evals:
        orc.select(None)
        orc.apply(pol, lg, val, None)
        n += 16
    secs = time.perf_counter() - t0
    return {"value": n / secs, "unit": "leaf-evals/s", "cores": 1, "kind": "port", "seconds": secs,
            "sample": f"{n} simulations of the oracle port's tree/env path only (network excluded: oracle/_ref not available)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL prints its version banner there) are pointed at
    # stderr, the line goes to a private copy of the original descriptor
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        if rank != 0:
            return 0
        cycles_per_step = 2  # bounded sample of a step: 2 of its 401 cycles (512 leaf evaluations)
        ref = run_reference(max(1, args.warmup) * cycles_per_step, max(1, args.steps) * cycles_per_step)
        if ref is None:
            ref = port_baseline(2000)
        line = {"impl": "reference", "metric": "selfplay_leaf_evals_per_sec", "value": ref["value"], "unit": "leaf-evals/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ref["seconds"] / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic (random-init weights, empty-board start)", "config": dict(CONFIG, reference_step=f"{cycles_per_step} cycles"),
                "games_per_sec_at_163_moves": ref["value"] / (SIMS + 1) / 163.0,
                "cpu_baseline": {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": ref["value"], "unit": "leaf-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=out, flush=True)
        return 0

    import torch
    import minizero_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    eng = minizero_b200.Engine(minizero_b200.GAME_GO, BOARD, GAMES, SIMS, device=local_rank)
    dims = dict(num_input_channels=IN_CH, input_height=BOARD, input_width=BOARD, num_hidden_channels=HIDDEN, num_blocks=BLOCKS, action_size=ACTIONS,
                num_value_hidden_channels=VALUE_HIDDEN, discrete_value_size=1)
    # model "broadcast": rank 0 reads the .pt (or draws random-init weights of the same architecture) and packs it; the packed blob
    # goes to the other ranks with one NCCL broadcast over NVLink — the only collective of the path
    weights = "reference create_network() random init (torch.manual_seed(0)) from oracle/_ref/nets" if os.path.exists(NET) else "numpy random init"
    if rank == 0:
        if os.path.exists(NET):
            eng.load_network(NET)
        else:
            import __graft_entry__ as ge
            eng.load_network((dims, ge.make_random_state(dims, np.random.default_rng(0))))
    else:
        eng.configure_network_empty(dims)
    if dist is not None:
        from minizero_b200 import dist as mzdist
        ptr, nbytes = eng.weight_blob()
        blob = torch.as_tensor(mzdist.DeviceBlob(ptr, nbytes), device=torch.device("cuda", local_rank))
        mzdist.broadcast_blob(dist, blob, src=0)
        torch.cuda.synchronize()

    rng = np.random.default_rng(1234 + rank)
    S1 = SIMS + 1
    rot = torch.empty((S1, GAMES), dtype=torch.uint8).pin_memory()
    noise = torch.empty((GAMES, ACTIONS), dtype=torch.float32).pin_memory()

    def draw_inputs():
        # training-default stochasticity: random rotation per evaluation, Dirichlet(0.03) noise at the root
        rot.numpy()[...] = rng.integers(0, 8, size=(S1, GAMES), dtype=np.uint8)
        noise.numpy()[...] = rng.dirichlet([0.03] * ACTIONS, size=GAMES).astype(np.float32)

    def e2e_step():
        """public-API step with host buffers: draw + upload the search's randomness, search, read the root tables back, choose the
        moves on the host (softmax-count, T=1), play them, restart finished games."""
        draw_inputs()
        eng.set_search_inputs(rot.numpy(), noise.numpy())
        eng.search(wait=False)
        r = eng.get_roots()
        cnt = r["count"].astype(np.float64)
        cum = np.cumsum(cnt, axis=1)
        pick = (rng.random(GAMES)[:, None] * cum[:, -1:] < cum).argmax(axis=1)
        actions = r["action"][np.arange(GAMES), pick].astype(np.int32)
        res = eng.play_all(actions)
        for g in np.nonzero(res["terminal"])[0]:
            eng.reset_game(int(g))
        return int(r["root_count"].sum())

    h2d = rot.numel() + noise.numel() * 4 + GAMES * 4
    d2h = GAMES * 16 + GAMES * ACTIONS * 4 * 7 + GAMES * 20

    # ---- warm-up (also instantiates the CUDA graph) -------------------------------------------------
    for _ in range(max(3, args.warmup)):
        e2e_step()
    draw_inputs()
    eng.set_search_inputs(rot.numpy(), noise.numpy())
    eng.sync()

    # ---- timed region 1: inputs resident in HBM, device clock ------------------------------------------
    sampler = ClockSampler(local_rank)
    launches0 = eng.launch_count()
    barrier()
    sampler.start()
    eng.timer_begin()
    for _ in range(args.steps):
        eng.search(wait=False)
        eng.play_max_count(auto_reset=True, read_back=False)
    dev_ms = eng.timer_end()
    barrier()
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0

    # ---- timed region 2: end to end through the public API with host buffers -----------------------------
    barrier()
    t0 = time.perf_counter()
    evals_e2e = 0
    for _ in range(args.steps):
        evals_e2e += e2e_step()
    eng.sync()
    barrier()
    e2e_s = time.perf_counter() - t0

    eng.search(wait=True)  # full trees (no move played) so that the tree kernel is profiled on a representative state
    prof = eng.profile_kernels(50)

    if dist is not None:
        dev = torch.device("cuda", local_rank)
        dev_ms, e2e_s = mzdist.max_over_ranks(dist, [dev_ms, e2e_s], device=dev)
        launches = int(mzdist.sum_over_ranks(dist, [float(launches)], device=dev)[0])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    pk = peaks()
    evals = world * GAMES * S1 * args.steps
    value = evals / (dev_ms * 1e-3)
    e2e_value = world * GAMES * S1 * args.steps / e2e_s
    layers = eng.conv_layers_per_launch()
    if layers > 1:  # fused tower: stem (18 real input channels) + 2 convs per block in one launch
        conv_kernel = f"conv_tower_kernel (all {layers} 3x3 conv layers of the 6bx256 tower, 256 positions, one launch)"
        flops_per_launch = GAMES * (2.0 * 81 * 9 * IN_CH * HIDDEN + (layers - 1) * 95551488.0)
    else:
        conv_kernel = "conv3x3 kernel (one hidden->hidden 3x3 conv layer, 256 positions)"
        flops_per_launch = FLOPS_PER_CONV_LAUNCH
    conv_tflops = flops_per_launch / (prof["conv_ms"] * 1e-3) / 1e12
    traffic, traffic_src = ncu_traffic() if layers > 1 else (None, None)
    line = {
        "metric": "selfplay_leaf_evals_per_sec", "value": value, "unit": "leaf-evals/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate; tree work f32/f64/int)",
        "data": "synthetic: " + weights + "; games from the empty board; Dirichlet(0.03) root noise + random rotations drawn on the host",
        "config": CONFIG,
        "games_per_sec_at_163_moves": value / S1 / 163.0,
        "frac_of_conv_flop_roofline": value * FLOPS_PER_EVAL / (world * pk["bf16_tflops_sustained"] * 1e12),
        "e2e": {"value": e2e_value, "unit": "leaf-evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": conv_kernel, "flops_per_launch": flops_per_launch, "bound": "tensor", "achieved": conv_tflops, "peak": pk["bf16_tflops"],
                     "unit": "TFLOP/s", "frac": conv_tflops / pk["bf16_tflops"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk["source"] + " burst (kernel timed alone, 50 launches)",
                     "launch_ms": prof["conv_ms"]},
        "kernels_ms": {"conv3x3": prof["conv_ms"], "tree_select_transition": prof["tree_ms"], "heads": prof["heads_ms"]},
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            cb = run_reference(2, 8) or port_baseline(2000)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the baseline is a report, never a reason to lose the measurement
            line["cpu_baseline"] = {"value": None, "unit": "leaf-evals/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: " + str(ex)[:200]}
    print(json.dumps(line), file=out, flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
