#!/usr/bin/env python
"""bench.py — self-play hot path throughput on the BASELINE.json configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

--config selects the workload (numbering of SURVEY.md §8: config k = BASELINE.json configs[k - 1]):
  2 (default; the configuration BASELINE's metric is quoted on): 9x9 Go AlphaZero, 400 simulations, 256 games per GPU, 6b x 256
  3: 8x8 Othello Gumbel MuZero, n = 16, m = 16, 512 games per GPU, 3b x 128
  4: 19x19 Go AlphaZero, 800 simulations, 128 games per GPU, 20b x 256
  5: Atari ms_pacman MuZero, 50 simulations, 256 environments per GPU, 1b x 256 (reference default size), synthetic 96 x 96 screens
A "step" is one whole move search for every game of the rank — (S + 1) cycles of select -> leaf transition / hidden-state gather ->
network -> expand / backup — followed by the move itself. Prints ONE JSON line (rank 0). See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
NETS = os.path.join(ROOT, "oracle", "_ref", "nets")
GUMBEL_CONF = "actor_use_gumbel=true:actor_use_gumbel_noise=true:actor_gumbel_sample_size=16:actor_gumbel_sigma_visit_c=50:actor_gumbel_sigma_scale_c=1:actor_use_dirichlet_noise=false:"

# FLOPs: SURVEY.md §8d (2 * MACs of convs + FCs of one position); tower_flops: the 3x3 convs of the tower launch profiled for the roofline
WORKLOADS = {
    2: dict(name="go9x9_alphazero_400sims_256games_6bx256 (BASELINE configs[1])", game="go", board=9, games=256, sims=400, net="go9_az_6bx256", muzero=0,
            dims=dict(num_input_channels=18, input_height=9, input_width=9, num_hidden_channels=256, num_blocks=6, action_size=82, num_value_hidden_channels=256, discrete_value_size=1),
            flops_per_eval=1.1535e9, flops_per_move=401 * 1.1535e9, moves_per_game=163, ref_binary="ref_actor_group_go",
            ref_conf="env_board_size=9:nn_type_name=alphazero:", tower="all 13 3x3 conv layers of the 6bx256 tower",
            tower_flops=lambda g: g * (2.0 * 81 * 9 * 18 * 256 + 12 * 95551488.0)),
    3: dict(name="othello8x8_gumbel_muzero_16sims_512games_3bx128 (BASELINE configs[2])", game="othello", board=8, games=512, sims=16, net="othello_mz_3bx128", muzero=1,
            engine=dict(use_gumbel=1, gumbel_noise=1, gumbel_sample_size=16), flops_per_eval=(0.1139e9 + 16 * 0.1324e9) / 17, flops_per_move=0.1139e9 + 16 * 0.1324e9,
            moves_per_game=60, ref_binary="ref_actor_group_othello", ref_conf="nn_type_name=muzero:" + GUMBEL_CONF,
            tower="dynamics tower: stem (128 + 1 planes) + 6 convs of 128 channels", tower_flops=lambda g: g * 2.0 * 64 * 9 * (129 * 128 + 6 * 128 * 128)),
    4: dict(name="go19x19_alphazero_800sims_128games_20bx256 (BASELINE configs[3])", game="go", board=19, games=128, sims=800, net="go19_az_20bx256", muzero=0,
            dims=dict(num_input_channels=18, input_height=19, input_width=19, num_hidden_channels=256, num_blocks=20, action_size=362, num_value_hidden_channels=256, discrete_value_size=1),
            flops_per_eval=17.07e9, flops_per_move=801 * 17.07e9, moves_per_game=400, ref_binary="ref_actor_group_go",
            ref_conf="env_board_size=19:nn_type_name=alphazero:", tower="all 41 3x3 conv layers of the 20bx256 tower",
            tower_flops=lambda g: g * 2.0 * 361 * 9 * (18 * 256 + 40 * 256 * 256)),
    5: dict(name="atari_ms_pacman_muzero_50sims_256envs_1bx256 (BASELINE configs[4]; synthetic 96x96 screens, ALE absent)", game="atari", board=6, games=256, sims=50,
            net="atari_mz_1bx256", muzero=1, engine=dict(value_rescale=1, reward_discount=0.997), flops_per_eval=(3.65e9 + 50 * 0.131e9) / 51, flops_per_move=3.65e9 + 50 * 0.131e9,
            moves_per_game=200, ref_binary="ref_actor_group_atari",
            ref_conf="env_atari_name=ms_pacman:actor_mcts_value_rescale=true:actor_mcts_reward_discount=0.997:nn_type_name=muzero:",
            tower="dynamics tower at 6x6: stem (256 + 18 planes) + 2 convs of 256 channels", tower_flops=lambda g: g * 2.0 * 36 * 9 * (274 * 256 + 2 * 256 * 256)),
}


def ncu_traffic(cfg):
    """dram__bytes_read.sum + dram__bytes_write.sum of one tower launch from the committed `ncu --set full` capture (profiles/)."""
    best = None
    for name in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
        if name.endswith("_ncu_full_summary.json"):
            best = os.path.join(ROOT, "profiles", name)
    try:
        with open(best) as f:
            j = json.load(f)
        key = "tower" if cfg == 2 else "tower_cfg%d" % cfg
        return j[key]["dram_traffic_bytes_per_launch"], os.path.basename(best)
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


def reference_conf(w, threads):
    return (w["ref_conf"] + f"actor_num_simulation={w['sims']}:zero_num_parallel_games={w['games']}:zero_num_threads={threads}:"
            f"nn_file_name={os.path.join(NETS, w['net'] + '.pt')}:program_seed=1:program_auto_seed=false:program_quiet=true")


def run_reference(w, warm_cycles, cycles, device=-1):
    """The UNMODIFIED reference actor path (oracle/_ref/ref_actor_group_*: reference sources compiled in place), tree / environment on
    zero_num_threads = nproc host threads, timed over whole ActorGroup cycles. device -1: network on the CPU through libtorch (the
    reference's CPU path); device >= 0: TorchScript on that GPU, i.e. the reference as shipped."""
    binary = os.path.join(ROOT, "oracle", "_ref", w["ref_binary"])
    net = os.path.join(NETS, w["net"] + ".pt")
    cores = os.cpu_count() or 1
    if not (os.path.exists(binary) and os.path.exists(net)):
        return None
    env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    if device < 0:
        env["CUDA_VISIBLE_DEVICES"] = ""
    res = subprocess.run([binary, "bench", reference_conf(w, cores), str(warm_cycles), str(cycles), str(device)], capture_output=True, text=True, env=env, timeout=3000)
    m = re.search(r"REFBENCH evals=(\d+) seconds=([0-9.eE+-]+) threads=(\d+)", res.stdout)
    if not m:
        raise RuntimeError("reference bench produced no REFBENCH line: " + res.stdout[-300:] + res.stderr[-300:])
    evals, secs = int(m.group(1)), float(m.group(2))
    where = "all-CPU" if device < 0 else f"tree / environment on {cores} host threads, TorchScript fp32 on GPU {device} (the reference as shipped)"
    emu = " over the synthetic frame source (stand-in ALE)" if w["game"] == "atari" else ""
    return {"value": evals / secs, "unit": "leaf-evals/s", "cores": cores, "kind": "reference", "seconds": secs,
            "sample": f"{cycles} ActorGroup cycles x {w['games']} games = {evals} leaf evaluations of the same workload (after {warm_cycles} warm-up cycles), {where}{emu}"}


def port_baseline(w, budget_evals):
    """Fallback when oracle/_ref is absent: the C restatement's tree / environment work only (its scalar fp32 network would take minutes)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    lib = oracle_lib.load()
    game = {"go": oracle_lib.GAME_GO, "othello": oracle_lib.GAME_OTHELLO, "atari": oracle_lib.GAME_ATARI}[w["game"]]
    opts = dict(w.get("engine", {}), muzero=w["muzero"])
    orc = oracle_lib.OracleSearch(lib, game, w["board"], 16, w["sims"], **opts)
    rng = np.random.default_rng(0)
    A = orc.A
    pol = rng.dirichlet([1.0] * A, size=16).astype(np.float32)
    lg, val = np.log(pol), np.zeros(16, np.float32)
    t0, n = time.perf_counter(), 0
    while n < budget_evals:
        for g in range(16):
            if orc.sims_done(g) == w["sims"] + 1:
                orc.lib.mzo_reset_search(orc.h, g)
        orc.lib.mzo_select(orc.h, None, None)
        orc.apply(pol, lg, val, None)
        n += 16
    secs = time.perf_counter() - t0
    return {"value": n / secs, "unit": "leaf-evals/s", "cores": 1, "kind": "port", "seconds": secs,
            "sample": f"{n} simulations of the oracle port's tree / environment path only (network excluded: oracle/_ref not available)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 6 (configs 2 and 4: a step is 0.15 s / 1.9 s), 300 (config 3) or 60 (config 5): about a second of timed region")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.config]
    if args.steps is None:
        args.steps = {3: 300, 5: 60}.get(args.config, 6)
    GAMES, SIMS = w["games"], w["sims"]
    S1 = SIMS + 1
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL prints its version banner there) are pointed at
    # stderr, the line goes to a private copy of the original descriptor
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    config = {"workload": w["name"], "games_per_gpu": GAMES, "simulations": SIMS, "evals_per_step_per_gpu": GAMES * S1,
              "net": w["net"] + " fp16 tensor-core / fp32 accumulate",
              "l2": "per-step working set (node pools + activations + hidden states) exceeds the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cycles_per_step = 2  # bounded sample of a step: 2 of its S + 1 cycles
        try:
            ref = run_reference(w, max(1, args.warmup) * cycles_per_step, max(1, args.steps) * cycles_per_step)
        except Exception as ex:
            print("reference binary failed: " + str(ex)[:300], file=sys.stderr)
            ref = None
        if ref is None:
            ref = port_baseline(w, 2000)
        line = {"impl": "reference", "metric": "selfplay_leaf_evals_per_sec", "value": ref["value"], "unit": "leaf-evals/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ref["seconds"] / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic (random-init weights, games from the initial position)",
                "config": dict(config, net=w["net"] + " fp32 (TorchScript on the host cores)", reference_step=f"{cycles_per_step} cycles"),
                "games_per_sec": ref["value"] / S1 / w["moves_per_game"],
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

    game_id = {"go": minizero_b200.GAME_GO, "othello": minizero_b200.GAME_OTHELLO, "atari": minizero_b200.GAME_ATARI}[w["game"]]
    eng = minizero_b200.Engine(game_id, w["board"], GAMES, SIMS, device=local_rank, muzero=w["muzero"], **w.get("engine", {}))
    A = eng.A
    atari, gumbel = (w["game"] == "atari"), bool(w.get("engine", {}).get("use_gumbel"))
    # model "broadcast": rank 0 reads the .pt (or draws random-init weights of the same architecture) and packs it; the packed blob
    # goes to the other ranks with one NCCL broadcast over NVLink — the only collective of the path
    net_path = os.path.join(NETS, w["net"] + ".pt")
    have_pt = os.path.exists(net_path)
    if not have_pt and "dims" not in w:
        raise SystemExit(f"{net_path} is missing (oracle/gen_nets.py writes it in the build container)")
    weights = "reference create_network() random init (torch.manual_seed(0)) from oracle/_ref/nets" if have_pt else "numpy random init of the same architecture"
    if dist is not None:  # every rank needs the dims; only rank 0 needs the values
        from minizero_b200 import dist as mzdist
    if rank == 0 or dist is None:
        if have_pt:
            eng.load_network(net_path)
        else:
            import __graft_entry__ as ge
            eng.load_network((w["dims"], ge.make_random_state(w["dims"], np.random.default_rng(0))))
        dims = eng.net_dims
    if dist is not None:
        dims = mzdist.broadcast_object(dist, dict(eng.net_dims) if rank == 0 else None, src=0)
        if rank != 0:
            eng.configure_network_empty(dims)
        ptr, nbytes = eng.weight_blob()
        blob = torch.as_tensor(mzdist.DeviceBlob(ptr, nbytes), device=torch.device("cuda", local_rank))
        mzdist.broadcast_blob(dist, blob, src=0)
        torch.cuda.synchronize()

    rng = np.random.default_rng(1234 + rank)
    use_rot = not w["muzero"]
    # two sets of pinned input buffers: the randomness of search k + 1 is drawn on the host while search k runs on the device
    rots = [torch.empty((S1, GAMES), dtype=torch.uint8).pin_memory() if use_rot else None for _ in range(2)]
    noises = [torch.empty((GAMES, A), dtype=torch.float32).pin_memory() for _ in range(2)]
    rot, noise = rots[0], noises[0]
    frames = torch.empty((GAMES, 3, 96, 96), dtype=torch.uint8).pin_memory() if atari else None
    frame_pool = rng.integers(0, 256, size=(4, GAMES, 3, 96, 96), dtype=np.uint8) if atari else None

    def draw_inputs(i=0):
        # training-default stochasticity: random rotation per evaluation and Dirichlet(0.03) root noise (AlphaZero),
        # Gumbel root noise (config 3), Dirichlet(0.25) over the 9 legal actions (Atari; quick-run's MuZero Atari settings)
        if use_rot:
            rots[i].numpy()[...] = rng.integers(0, 8, size=(S1, GAMES), dtype=np.uint8)
        if gumbel:
            noises[i].numpy()[...] = rng.gumbel(size=(GAMES, A)).astype(np.float32)
        else:
            noises[i].numpy()[...] = rng.dirichlet([0.25 if atari else 0.03] * A, size=GAMES).astype(np.float32)

    step_no = [0]
    cur = [0]

    def e2e_step():
        """public-API step with host buffers: draw + upload the search's randomness (Atari: and the emulator's new screens), search, read
        the root tables back, choose the moves on the host (softmax-count, T=1; Gumbel: the best candidate), play them, restart finished games."""
        i = cur[0]
        eng.set_search_inputs(rots[i].numpy() if use_rot else None, noises[i].numpy())  # host -> device, every step
        eng.search(wait=False)
        draw_inputs(1 - i)  # the next search's randomness, drawn while this one runs
        cur[0] = 1 - i
        r = eng.get_roots()
        if gumbel:
            actions = eng.gumbel_best_actions().astype(np.int32)
        else:
            cnt = r["count"].astype(np.float64)
            cum = np.cumsum(cnt, axis=1)
            pick = (rng.random(GAMES)[:, None] * cum[:, -1:] < cum).argmax(axis=1)
            actions = r["action"][np.arange(GAMES), pick].astype(np.int32)
        res = eng.play_all(actions)
        if atari:  # the host's emulator answers with a screen per environment
            frames.numpy()[...] = frame_pool[step_no[0] % 4]
            step_no[0] += 1
            eng.observe_all(actions, frames.numpy())
        for g in np.nonzero(res["terminal"])[0]:
            eng.reset_game(int(g))
        return int(r["root_count"].sum())

    h2d = (rot.numel() if use_rot else 0) + noise.numel() * 4 + GAMES * 4 + (GAMES * 3 * 96 * 96 + GAMES * 4 if atari else 0)
    d2h = GAMES * 16 + GAMES * A * 4 * 7 + GAMES * A * 4 + GAMES * 12 + GAMES * 20 + (GAMES * 4 if gumbel else 0)

    if atari:
        eng.observe_all(np.full(GAMES, -1, np.int32), frame_pool[0])
    # ---- warm-up (also instantiates the CUDA graph) -------------------------------------------------
    draw_inputs(0)
    for _ in range(max(3, args.warmup)):
        e2e_step()
    draw_inputs()
    eng.set_search_inputs(rot.numpy() if use_rot else None, noise.numpy())
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
    tower_kernel = "conv_tower_wide_kernel" if eng.tower_is_wide() == 1 else "conv_tower_kernel"
    layers = eng.conv_layers_per_launch()
    flops_per_launch = w["tower_flops"](GAMES)
    conv_tflops = flops_per_launch / (prof["conv_ms"] * 1e-3) / 1e12
    traffic, traffic_src = ncu_traffic(args.config)
    line = {
        "metric": "selfplay_leaf_evals_per_sec", "value": value, "unit": "leaf-evals/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate; tree work f32/f64/int)",
        "data": "synthetic: " + weights + "; games from the initial position; root noise" + (" + random rotations" if use_rot else "") + " drawn on the host"
                + ("; random 96x96 RGB screens stand in for the emulator" if atari else ""),
        "config": config,
        "games_per_sec": value / S1 / w["moves_per_game"], "moves_per_game_assumed": w["moves_per_game"],
        "frac_of_conv_flop_roofline": value / S1 * w["flops_per_move"] / (world * pk["bf16_tflops_sustained"] * 1e12),
        "e2e": {"value": e2e_value, "unit": "leaf-evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": f"{tower_kernel} ({w['tower']}; {GAMES} positions, one launch covering {layers} layers)", "flops_per_launch": flops_per_launch, "bound": "tensor",
                     "achieved": conv_tflops, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": conv_tflops / pk["bf16_tflops"], "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": pk["source"] + " burst (kernel timed alone, 50 launches)", "launch_ms": prof["conv_ms"]},
        "kernels_ms": {"conv_tower": prof["conv_ms"], "tree_select_transition": prof["tree_ms"], "heads": prof["heads_ms"]},
    }
    if args.config == 2:
        line["games_per_sec_at_163_moves"] = value / S1 / 163.0
    if world == 1 and not args.no_cpu_baseline:
        try:
            cb = run_reference(w, 2, 8 if args.config != 4 else 3) or port_baseline(w, 2000)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the baseline is a report, never a reason to lose the measurement
            line["cpu_baseline"] = {"value": None, "unit": "leaf-evals/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: " + str(ex)[:200]}
    if world == 1 and not args.no_gpu_reference:
        # the reference as shipped: its host-thread tree + TorchScript on THIS GPU, after our engine released it
        try:
            eng.close()
            torch.cuda.empty_cache()
            gr = run_reference(w, 4, 24 if args.config != 4 else 6, device=local_rank)
            line["gpu_reference"] = {k: gr[k] for k in ("value", "unit", "cores", "kind", "sample")} if gr else None
        except Exception as ex:
            line["gpu_reference"] = {"value": None, "unit": "leaf-evals/s", "sample": "failed: " + str(ex)[:200]}
    print(json.dumps(line), file=out, flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
