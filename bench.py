#!/usr/bin/env python
"""bench.py -- microbe-steps/s (advect + interact) of the B200 hot path, one JSON line on stdout.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...          # the reference's CPU path on the host cores

A *step* is one pass of the hot path over all microbes: RK4 advection, binning, radius pair search
fused with rock-paper-scissors resolution, pair list emitted (BASELINE.json metric:
"microbe-steps/sec (advect+interact)").

Workloads (``config.workload``):
  shard    (default) the per-GPU shard of BASELINE config 4 "100M microbes basin-scale North Pacific
           across 8xB200": 12.5M microbes per GPU, uniform-random over lon 180-240 x a 7.5-degree
           latitude strip per GPU (N = 8 gives lat 0-60 and exactly config 4), r = 0.01 deg, p = 0.55,
           time-varying synthetic random-Fourier velocity on the OSCAR 1/3-degree grid, dt = 1 h.
           Weak scaling: per-GPU work fixed as N grows.
  config2  BASELINE config 2: 490,000 microbes (700x700 lattice, 25-35N 205-215E), time-varying field,
           spun up ``--spinup`` steps so that the timed steps see the stirred state.
  config3  BASELINE config 3: 10M microbes uniform in the 10x10-degree patch (rho = 15.7).
  config5  BASELINE config 5, one point of the interaction-only sweep: ``--microbes`` (default 50 M) uniform-random at the
           config-1 areal density (4,900 per square degree), ``--radius`` (default 0.01), NO advection -- pair search + RPS on
           resident positions, the pair list emitted.  tools/sweep_interact.py runs the whole sweep (1-50 M, 0.5-5 km, p).
  config1  BASELINE config 1, the reference's own CPU-runnable case: 490,000 microbes on the 700x700 lattice, STEADY
           synthetic velocity, 24 hourly steps.  The lattice spacing exceeds r, so the configuration as specified has no
           pairs; ``--spinup`` (default 120) advection-only steps strain the lattice first so that the timed steps
           interact.  ``--impl reference`` runs the SAME 490,000 microbes through the reference's CPU path: the one
           same-size GPU/CPU comparison.
Inputs are larger than L2 for shard/config3 (state 162 MB + pair list 440 MB per step vs 126 MB L2).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "microbe-steps/sec (advect+interact)"
UNIT = "microbe-steps/s"
RADIUS = 0.01
P_RPS = (0.55, 0.55, 0.55)
DT = 3600.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------
def workload_particles(name, n, rank, world, seed=0):
    """Returns lon, lat (float64), species int8, description dict."""
    rng = np.random.default_rng(seed * 1000 + rank)
    if name == "shard":
        strip = 7.5
        lat0 = 30.0 - 0.5 * strip * world + strip * rank
        lon = 180.0 + 60.0 * rng.random(n)
        lat = lat0 + strip * rng.random(n)
        desc = dict(lon=[180.0, 240.0], lat=[30.0 - 0.5 * strip * world, 30.0 + 0.5 * strip * world])
    elif name == "config3":
        lon = 205.0 + 10.0 * rng.random(n)
        lat = 25.0 + 10.0 * rng.random(n)
        desc = dict(lon=[205.0, 215.0], lat=[25.0, 35.0])
    elif name == "config5":
        side = float(np.sqrt(n / 4900.0))
        lon = 205.0 + side * rng.random(n)
        lat = 10.0 + side * rng.random(n)
        desc = dict(lon=[205.0, 205.0 + side], lat=[10.0, 10.0 + side])
    elif name in ("config2", "config1"):
        from lagrangian_microbes_b200.particle_advecter import uniform_particle_locations
        lon, lat = uniform_particle_locations(n, 25, 35, 205, 215)
        desc = dict(lon=[205.0, 215.0], lat=[25.0, 35.0])
    else:
        raise ValueError(name)
    species = rng.integers(1, 4, n).astype(np.int8)
    return lon, lat, species, desc


def default_n(name):
    return {"shard": 12_500_000, "config3": 10_000_000, "config2": 490_000, "config1": 490_000, "config5": 50_000_000}[name]


def make_fieldset(n_modes, kind="random_fourier"):
    from lagrangian_microbes_b200 import velocity_fields
    from lagrangian_microbes_b200.particle_advecter import HostFieldSet
    velocity_fields.configure_synthetic(kind=kind, seed=0, n_modes=n_modes, rms_speed=0.2)
    return HostFieldSet(velocity_fields.oscar_dataset(2017))


def field_kind(workload):
    return "steady" if workload == "config1" else "random_fourier"


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# CPU reference arm / baseline
# ------------------------------------------------------------------------------------------------------
def cpu_reference_step_setup(workload, n_sample, n_modes):
    """The workload itself (config1, config2: the 490,000 microbes of the configuration) or a bounded sample of it:
    the same areal density, fewer microbes (shard, config3)."""
    from oracle import rk4 as ork4
    n_full = default_n(workload)
    lattice = workload in ("config1", "config2")
    lon, lat, species, desc = workload_particles(workload, n_full if lattice else n_sample, 0, 1)
    if not lattice:
        # shrink the box so that the density (hence pairs per microbe) is that of the full workload
        frac = n_sample / float(n_full)
        if workload == "shard":
            lat_lo = 30.0 - 3.75
        else:
            lat_lo = desc["lat"][0]
        s = np.sqrt(frac)
        lon = desc["lon"][0] + (lon - desc["lon"][0]) * s
        lat = lat_lo + (lat - lat_lo) * s
    hfs = make_fieldset(n_modes, field_kind(workload))
    fs = ork4.FieldSet(hfs.lon, hfs.lat, hfs.time, hfs.u, hfs.v)
    return fs, lon.astype(np.float32), lat.astype(np.float32), species


def cpu_reference_step(fs, lon, lat, params, props, t, ti, threads):
    """One step of the reference's CPU path.  Returns (ti, n_pairs, dict of phase seconds)."""
    from oracle import pairs as opairs, rk4 as ork4, rps as orps
    t0 = time.perf_counter()
    ti, _ = ork4.rk4_step_c(fs, lon, lat, t, DT, ti, threads=threads)            # parcels' JIT'd C, restated; tiles over all cores
    t1 = time.perf_counter()
    pair_set = opairs.query_pairs_reference(lon, lat, RADIUS)                    # cKDTree build + query_pairs (set): the reference's call
    t2 = time.perf_counter()
    # the reference's loop (interaction_simulator.py:104-105) over the set, calling the pair function with the reference's
    # signature and per-call work (dict look-ups, np.random.rand() when the species differ): interactions.py:13-40
    fn = orps.reference_pair_interaction
    for pair in pair_set:
        fn(params, props, pair[0], pair[1])
    t3 = time.perf_counter()
    return ti, len(pair_set), {"advect_s": t1 - t0, "tree_query_s": t2 - t1, "rps_loop_s": t3 - t2}


def run_cpu_reference(workload, n_sample, steps, warmup, n_modes, spinup=0):
    from oracle import rk4 as ork4, rps as orps
    orps.build_c()
    threads = os.cpu_count() or 1
    fs, lon, lat, species = cpu_reference_step_setup(workload, n_sample, n_modes)
    params = {"pRS": P_RPS[0], "pPR": P_RPS[1], "pSP": P_RPS[2]}
    props = {"species": species}
    np.random.seed(0)
    t, ti = 0.0, 0
    for _ in range(spinup):                      # advection only (untimed): strain the lattice until pairs exist
        ti, _ = ork4.rk4_step_c(fs, lon, lat, t, DT, ti, threads=threads)
        t += DT
    phases = {"advect_s": 0.0, "tree_query_s": 0.0, "rps_loop_s": 0.0}
    pairs = 0
    tot = 0.0
    for s in range(warmup + steps):
        ti, npairs, ph = cpu_reference_step(fs, lon, lat, params, props, t, ti, threads)
        t += DT
        if s >= warmup:
            tot += ph["advect_s"] + ph["tree_query_s"] + ph["rps_loop_s"]
            pairs += npairs
            for k2 in phases:
                phases[k2] += ph[k2]
    value = lon.size * steps / tot
    return {"value": value, "seconds": tot, "n_sample": int(lon.size), "pairs_per_step": pairs / max(steps, 1),
            "phases_s_per_step": {k2: v / max(steps, 1) for k2, v in phases.items()}, "threads": threads}


# ------------------------------------------------------------------------------------------------------
# parity twin: correctness evidence printed WITH the throughput (outside the timed region)
# ------------------------------------------------------------------------------------------------------
def parity_twin(world, rank, hfs, steps=4, per_rank=40_000, transport="peer"):
    """A reduced-size twin of the workload (same areal density, same field, same code path as the timed run) stepped
    ``steps`` times.  N = 1: every step checked against the CPU oracle -- positions vs the float64 RK4 restatement
    (1e-6 relative), the emitted pair set vs cKDTree.query_pairs (exact), the species vs the reference rule run
    sequentially in the canonical order (exact).  N > 1: the twin is sharded over the N ranks as latitude strips (NCCL
    halo exchange + migration, as in the timed run) AND stepped by one handle on rank 0; positions and species of every
    microbe must agree bit for bit.  Returns the ``parity`` block of the JSON line (identical on all ranks)."""
    import hashlib
    import torch
    from lagrangian_microbes_b200.simulation import FusedSimulation
    n = per_rank * world
    rng = np.random.default_rng(12345)
    side = float(np.sqrt(n / 27_800.0))                 # the shard's areal density: rho ~ 4.4 pairs per microbe
    lon = (200.0 + side * rng.random(n)).astype(np.float32)
    lat = (30.0 + side * rng.random(n)).astype(np.float32)
    sp = rng.integers(1, 4, n).astype(np.int8)
    out = {"twin_microbes": n, "steps": steps}

    def digest(a, b, c):
        h = hashlib.sha256()
        for x in (a, b, c):
            h.update(np.ascontiguousarray(x).tobytes())
        return h.hexdigest()[:16]

    single = None
    if rank == 0:
        single = FusedSimulation(lon, lat, sp, RADIUS, *P_RPS, hfs, dt_seconds=DT, seed=7, emit_pairs=(world == 1),
                                 pair_capacity=24 * n, regrid_every=2, grid_margin=0.5)
    if world == 1:
        from oracle import pairs as opairs, philox, rk4 as ork4, rps as orps
        fs = ork4.FieldSet(hfs.lon, hfs.lat, hfs.time, hfs.u, hfs.v)
        pairs_ok = species_ok = True
        rel, pairs_total = 0.0, 0
        l0, a0, s0 = lon, lat, sp
        for step in range(steps):
            grid = single.grid.as_dict()
            t0, ti0 = single.clock.t, single.clock.ti
            st = single.step(check=True)
            gl, ga, gs = single.download()
            a64, b64, _, _ = ork4.rk4_step_f64(fs, l0, a0, t0, DT, ti0)
            rel = max(rel, float(np.max(np.abs(gl - a64) / np.abs(a64))), float(np.max(np.abs(ga - b64) / np.abs(b64))))
            want = opairs.query_pairs_reference_array(gl, ga, RADIUS)
            got = opairs.sort_pairs(single.pairs[:st.n_pairs].cpu().numpy())
            pairs_ok = pairs_ok and st.n_pairs == want.shape[0] and bool(np.array_equal(got, want))
            order, _ = orps.canonical_order(want, gl, ga, grid, mode=single.engine.interact_mode)
            u = philox.pair_uniforms(order[:, 0], order[:, 1], step, 7)
            s_ref, _ = orps.rps_sequential_c(s0.copy(), order, u, *P_RPS)
            species_ok = species_ok and bool(np.array_equal(gs, s_ref))
            pairs_total += int(st.n_pairs)
            l0, a0, s0 = gl, ga, gs
        out.update({"against": "CPU oracle (float64 RK4 restatement, cKDTree.query_pairs, reference rule in canonical order)",
                    "positions_rel_err": rel, "pairs_exact": pairs_ok, "species_exact": species_ok, "pairs": pairs_total,
                    "checksum": digest(l0, a0, s0), "match": bool(pairs_ok and species_ok and rel < 1e-6)})
        single.engine.close()
        return out
    from lagrangian_microbes_b200.strips import DistTransport, PeerTransport, StripSet
    import torch.distributed as dist
    mine = slice(rank * per_rank, (rank + 1) * per_rank)                  # the reference's contiguous tiles
    ids = np.arange(n, dtype=np.int32)
    ss = StripSet(PeerTransport() if transport == "peer" else DistTransport(), lon[mine], lat[mine], sp[mine], ids[mine], n, RADIUS, *P_RPS, hfs, dt_seconds=DT, seed=7,
                  emit_pairs=False, slack=max(1.6, float(world)), grid_margin=0.5, regrid_every=2)
    # (slack: the twin hands every rank a contiguous TILE of ids, i.e. microbes from all over the domain, and settle() routes
    #  them one strip per pass -- the inner strips hold the traffic of both directions on the way: room for all of it)
    for _ in range(steps):
        ss.step()
        if single is not None:
            single.step()
    gl, ga, gs = ss.gather()                                               # every rank: all microbes in id order
    ok = torch.ones(1, dtype=torch.int32, device="cuda")
    if single is not None:
        wl, wa, ws = single.download()
        same = bool(np.array_equal(gl, wl) and np.array_equal(ga, wa) and np.array_equal(gs, ws))
        ok[0] = 1 if same else 0
        out["checksum_single_handle"] = digest(wl, wa, ws)
        single.engine.close()
    dist.broadcast(ok, src=0)
    moved = ss.transport.all_sum([[float(s.engine.sync_stats().n_moved_in)] for s in ss.strips])
    out.update({"against": "one handle stepping all twin microbes on rank 0 (bit for bit: positions, species)",
                "checksum_strips": digest(gl, ga, gs), "strip_edges": [int(e) for e in ss.edges],
                "migrated_last_step": int(moved[0]), "match": bool(int(ok[0]) == 1)})
    ss.close()
    return out


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default 200; config1: the configuration's 24)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="shard", choices=["shard", "config1", "config2", "config3", "config5"])
    ap.add_argument("--radius", type=float, default=0.01, help="config5: interaction radius in degrees (0.005 - 0.05 = 0.5 - 5 km)")
    ap.add_argument("--microbes", type=int, default=0, help="microbes per GPU (0 = the workload's size)")
    ap.add_argument("--spinup", type=int, default=-1, help="untimed steps before warm-up (config2 default 1500; config1 default 120, advection only)")
    ap.add_argument("--modes", type=int, default=64, help="Fourier modes of the synthetic velocity field")
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="microbes in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity twin (tuning sweeps)")
    ap.add_argument("--packed-record", action="store_true",
                    help="e2e (N = 1, experiment): positions leave as the lossless delta-packed record (DESIGN.md 4.7), "
                         "stored packed -- not decoded inside the timed region")
    ap.add_argument("--resolve-upl", type=int, default=0, help="LM_OPT_RESOLVE_UPL (tuning experiments)")
    ap.add_argument("--no-overlap", action="store_true", help="LM_OPT_OVERLAP = 0 (tuning experiments)")
    ap.add_argument("--resolve-mode", type=int, default=0, choices=[0, 1],
                    help="LM_OPT_RESOLVE_MODE: 0 nine phase launches (default), 1 tiled resolver (experimental)")
    ap.add_argument("--tile-smem", type=int, default=0, help="LM_OPT_RESOLVE_TILE_SMEM (with --resolve-mode 1)")
    ap.add_argument("--interact-mode", type=int, default=2, choices=[0, 1, 2],
                    help="LM_OPT_INTERACT_MODE: 2 hybrid (default: round-1 pipeline for the light units + rounds of matchings for "
                         "the queued heavy units), 1 fused tile kernel, 0 the round-1 pipeline alone (A/B)")
    ap.add_argument("--advect-mode", type=int, default=1, choices=[0, 1],
                    help="LM_OPT_ADVECT_MODE: 1 float32 RK4 within north_star's 1e-6 relative (default here), 0 bit-faithful "
                         "to the float32 restatement of Parcels' kernel (A/B)")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="N > 1: strip exchange through peer memory (CUDA IPC, peer stores + flags; default) or NCCL send/recv (A/B)")
    ap.add_argument("--draw-batch", type=int, default=0, help="LM_OPT_DRAW_BATCH (tuning experiments)")
    ap.add_argument("--tile-cap", type=int, default=0, help="LM_OPT_TILE_CAP (tuning experiments)")
    args = ap.parse_args()
    assert args.warmup >= 3 or args.impl == "reference", "timing rules: at least 3 warm-up steps"
    global RADIUS
    interact_only = args.workload == "config5"
    if interact_only:
        RADIUS = args.radius
        args.no_e2e = True              # no advection, no record: the device-timed interaction step is the figure
        assert args.impl == "b200" and int(os.environ.get("WORLD_SIZE", "1")) == 1, "config5: one GPU, b200 arm"
    if args.steps <= 0:
        args.steps = 24 if args.workload == "config1" else (20 if interact_only else 200)
    spinup = args.spinup if args.spinup >= 0 else {"config2": 1500, "config1": 120}.get(args.workload, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # shard: weak scaling (the per-GPU shard of config 4 on every GPU).  config3 is ONE configuration -- 10 M microbes in the
    # dense patch -- on 1 or 2 GPUs (BASELINE.json: "1xB200 vs 2xB200 latitude-strip sharding"): strong scaling.
    strong = args.workload == "config3" and world > 1
    n_per_gpu = args.microbes or (default_n(args.workload) // world if strong else default_n(args.workload))
    config = {"workload": args.workload, "microbes_per_gpu": n_per_gpu, "microbes_total": n_per_gpu * world,
              "radius_deg": RADIUS, "p": P_RPS[0], "dt_s": DT, "field": ("none (interaction only)" if args.workload == "config5" else "synthetic %s, OSCAR 1/3-degree grid "
              "(72x481x1201), %d modes" % ("steady eddy field" if args.workload == "config1" else "random-Fourier", args.modes)), "l2": "inputs larger than L2" if n_per_gpu >= 5_000_000
              else "state fits L2 (config as specified)",
              "regrid": "bounding box read back every 16 steps (one small sync), cell grid re-fitted when the cloud nears its edge"}

    if args.resolve_mode:
        config["resolver"] = "tiled (LM_OPT_RESOLVE_MODE=1, experimental)"
    if args.impl == "reference":
        if rank != 0:
            return
        same_size = args.workload in ("config1", "config2")
        steps = max(1, min(args.steps, 24 if args.workload == "config1" else 3))
        warm = 1 if args.warmup > 0 else 0
        res = run_cpu_reference(args.workload, args.cpu_sample, steps, warm, args.modes, spinup=spinup)
        if same_size:
            sample = "the configuration itself: all %d microbes, %d steps after %d advection-only spin-up steps (%.0f pairs/step)" % (
                res["n_sample"], steps, spinup, res["pairs_per_step"])
        else:
            sample = "%d steps of a %d-microbe sub-box of the workload at the same areal density (%.0f pairs/step)" % (
                steps, res["n_sample"], res["pairs_per_step"])
        # the reference line describes what the reference arm RAN: the sample, not the GPU arm's particle count
        config["microbes_per_gpu"] = config["microbes_total"] = res["n_sample"]
        config["sample_of"] = None if same_size else "%d microbes per GPU in the b200 arm" % n_per_gpu
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": 1e3 * res["seconds"] / steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 state / f64 arithmetic", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["threads"], "kind": "port",
                                 "sample": sample, "phases_s_per_step": res["phases_s_per_step"],
                                 "note": "advect: C restatement of parcels' JIT kernel, contiguous tiles over all cores "
                                         "(particle_advecter.py:143-148); pair search: the reference's own "
                                         "cKDTree.query_pairs (1 core); RPS: the reference's Python loop over the set "
                                         "calling the pair function restated with the reference's signature, dict look-ups "
                                         "and np.random.rand() (1 core; interactions.py:13-40, interaction_simulator.py:104-105)"},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version banner ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback in the product path)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from lagrangian_microbes_b200.simulation import FusedSimulation
    from lagrangian_microbes_b200.engine import Engine
    Engine.DEFAULT_INTERACT_MODE = args.interact_mode           # every handle of this run (single or strips)
    Engine.DEFAULT_ADVECT_MODE = args.advect_mode
    config["interact"] = {2: "hybrid: pair search -> hand-off -> nine phase launches for the light units, device-wide queue of "
                             "heavy units resolved in rounds of matchings; cell-round order (LM_OPT_INTERACT_MODE=2)",
                          1: "fused tile kernel, tile-round order (LM_OPT_INTERACT_MODE=1)",
                          0: "round-1 pipeline: pair search -> hand-off -> nine phase launches (LM_OPT_INTERACT_MODE=0)"}[args.interact_mode]
    config["advect"] = ("float32 RK4, positions within 1e-6 relative of the float64 RK4 (LM_OPT_ADVECT_MODE=1)"
                        if args.advect_mode == 1 else "bit-faithful to the float32 restatement of Parcels' kernel (LM_OPT_ADVECT_MODE=0)")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_setup = time.time()
    hfs = None if interact_only else make_fieldset(args.modes, field_kind(args.workload))
    lon, lat, species, desc = workload_particles(args.workload, n_per_gpu, rank, world)
    config.update(desc)
    log("[rank %d] setup: field + particles in %.1f s" % (rank, time.time() - t_setup))

    # pair-list capacity per microbe: rho grows as the flow gathers the microbes (4.4 -> 6.9 after 1,000 steps of the
    # default workload, 15.7 -> 18.9 after 400 of config 3); an overflow is an error, not a silent truncation
    ppp = 14 if args.workload != "config3" else 36
    if interact_only:
        rho_est = 0.5 * np.pi * RADIUS * RADIUS * 4900.0
        ppp = 1.3 * rho_est + 4.0 * np.sqrt(rho_est / n_per_gpu) + 0.05

    class Sharded:
        """N > 1: one latitude strip per rank (lagrangian_microbes_b200/strips.py), NCCL between neighbours."""

        def __init__(self, stream_field):
            from lagrangian_microbes_b200.strips import DistTransport, PeerTransport, StripSet
            ids = (rank * n_per_gpu + np.arange(n_per_gpu)).astype(np.int32)
            self.ss = StripSet(PeerTransport() if args.transport == "peer" else DistTransport(), lon, lat, species, ids, n_per_gpu * world, RADIUS, *P_RPS, hfs,
                               dt_seconds=DT, seed=0, emit_pairs=True, pairs_per_particle=ppp, slack=1.4,
                               grid_margin=0.5, stream_field=stream_field, regrid_every=16, rebalance_every=64)
            self.engine = self.ss.strips[0].engine
            self.regrid_every = 0
            self.k = 0

        def step(self, check=False, timing=False):
            self.ss.step(check=check, timing=timing)

        def stats(self):
            return self.ss.stats()[0]

        @property
        def h2d_bytes_last_step(self):
            return self.ss.streamer.h2d_bytes_last_step if self.ss.streamer is not None else 0

        def record_to_host(self, *_):
            self.k ^= 1
            return self.ss.record_to_host(self.k)

    def new_sim(stream_field):
        if world > 1:
            return Sharded(stream_field)
        return FusedSimulation(lon, lat, species, RADIUS, *P_RPS, hfs, dt_seconds=DT, seed=0, emit_pairs=True,
                               pair_capacity=int(max(1 << 20, ppp * n_per_gpu)), advect=not interact_only,
                               regrid_every=0 if interact_only else 16, grid_margin=0.1 if interact_only else 0.5,
                               stream_field=stream_field and not interact_only)

    parity = None
    if not args.no_parity and not interact_only:
        t_par = time.time()
        parity = parity_twin(world, rank, hfs, transport=args.transport)
        log("[rank %d] parity twin: %s in %.1f s" % (rank, "MATCH" if parity["match"] else "MISMATCH", time.time() - t_par))
        torch.cuda.empty_cache()

    sim = new_sim(False)
    if args.draw_batch or args.tile_cap:
        from lagrangian_microbes_b200._lib import LM_OPT_DRAW_BATCH, LM_OPT_TILE_CAP
        sim.engine.set_option(LM_OPT_DRAW_BATCH, args.draw_batch)
        sim.engine.set_option(LM_OPT_TILE_CAP, args.tile_cap)
    if args.resolve_upl:
        from lagrangian_microbes_b200._lib import LM_OPT_RESOLVE_UPL
        sim.engine.set_option(LM_OPT_RESOLVE_UPL, args.resolve_upl)
    if args.no_overlap:
        from lagrangian_microbes_b200._lib import LM_OPT_OVERLAP
        sim.engine.set_option(LM_OPT_OVERLAP, 0)
    if args.resolve_mode:
        from lagrangian_microbes_b200._lib import LM_OPT_RESOLVE_MODE, LM_OPT_RESOLVE_TILE_SMEM
        sim.engine.set_option(LM_OPT_RESOLVE_MODE, args.resolve_mode)
        if args.tile_smem:
            sim.engine.set_option(LM_OPT_RESOLVE_TILE_SMEM, args.tile_smem)
    def spin(sm):
        if args.workload == "config1":               # advection only: strain the lattice until pairs exist (both arms do this)
            assert world == 1
            sm.interact, emit = False, sm.emit_pairs
            sm.emit_pairs = False
            for _ in range(spinup):
                sm.step()
            sm.interact, sm.emit_pairs = True, emit
        else:
            for _ in range(spinup):
                sm.step()

    spin(sim)
    if args.workload in ("config1", "config2"):
        sim.step(check=True)
    for _ in range(args.warmup):
        sim.step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = sim.engine.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        sim.step()
    sim.engine.join()                              # the last step's RPS phases run on the library's side stream
    ev1.record()
    barrier()
    wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = sim.engine.launch_count() - launches0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    st = sim.stats()
    rho = st.n_pairs / float(n_per_gpu)
    if world > 1:
        config["parallelism"] = "%d latitude strips (one per GPU), %s between neighbours: migration + one-row halo + boundary " \
                                "species, strip edges %s" % (world, "peer-memory stores over NVLink (CUDA IPC) + flags" if args.transport == "peer"
                                                             else "NCCL send/recv", sim.ss.edges)

    # per-phase device times (3 extra steps with events between the phases)
    phase = np.zeros(5)
    for _ in range(3):
        sim.step(timing=True)
        phase += np.array(sim.engine.phase_times())
    phase /= 3.0
    barrier()

    # end to end: velocity snapshots streamed H2D from pinned host memory every step, per-step record
    # (lon, lat, species in particle-id order) read back D2H into pinned host buffers every step.
    e2e = None
    if not args.no_e2e:
        del sim
        torch.cuda.empty_cache()
        sim2 = new_sim(True)
        spin(sim2)
        rec = [(torch.empty(n_per_gpu, dtype=torch.float32).pin_memory(), torch.empty(n_per_gpu, dtype=torch.float32).pin_memory(),
                torch.empty(n_per_gpu, dtype=torch.int8).pin_memory()) if world == 1 else () for _ in range(2)]
        packer = None
        if args.packed_record:
            assert world == 1, "--packed-record: single handle only"
            from lagrangian_microbes_b200.record import DeltaRecordPacker
            packer = DeltaRecordPacker(n_per_gpu)
            lon_d = torch.empty(n_per_gpu, dtype=torch.float32, device="cuda"); lat_d = torch.empty_like(lon_d)

        host_in_record = []

        def e2e_step(k):
            if packer is not None:
                sim2.step()
                if len(packer.pending) == 2:
                    packer.pop(decode=False)       # the packed record of step k - 2 is in pinned memory
                sim2.engine.state_get(lon_d, lat_d, None)
                packer.push(lon_d, lat_d)
                sim2.engine.state_get_host(None, None, rec[k & 1][2])
            elif world == 1:
                sim2.step(record=rec[k & 1])       # record scattered + copied under the step (lm_record_next_step)
            elif os.environ.get("LM_E2E_VARIANT", "") == "after":         # (measurement only: the record issued after the step)
                sim2.step()
                th = time.perf_counter()
                sim2.record_to_host()
                host_in_record.append(time.perf_counter() - th)
            elif os.environ.get("LM_E2E_VARIANT", "") == "norecord":      # (measurement only)
                sim2.step()
            else:
                sim2.k ^= 1
                sim2.ss.step(record=sim2.k)        # the strip's record copied inside the step (lm_record_next_step_ids)

        for k in range(args.warmup):
            e2e_step(k)
        (sim2.ss if world > 1 else sim2.engine).host_copies_sync()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw0 = time.perf_counter()
        e0.record()
        h2d = 0
        packed_bytes0 = packer.bytes_d2h if packer is not None else 0
        for k in range(args.steps):
            e2e_step(k)
            h2d += sim2.h2d_bytes_last_step
        (sim2.ss if world > 1 else sim2.engine).host_copies_sync()
        if packer is not None:
            while packer.pending:
                last_packed = packer.pop(decode=False)
        e1.record()
        barrier()
        tw1 = time.perf_counter()
        ms_e2e = max(e0.elapsed_time(e1), 1e3 * (tw1 - tw0))     # device events and the host clock around the copies
        # N = 1: lon, lat, species in particle-id order (9 B); strips: ids travel with the record (13 B)
        e2e = {"ms": ms_e2e, "h2d": h2d / args.steps, "d2h": (9 if world == 1 else 13) * n_per_gpu}
        if world == 1 and packer is None:
            e2e["record"] = "lon, lat, species in particle-id order: scattered and copied inside the step (lm_record_next_step)"
        elif world > 1 and os.environ.get("LM_E2E_VARIANT", "") == "":
            e2e["d2h"] = 13 * int(sim2.ss._record_slot_count[sim2.k])        # this rank's strip at the last step
            e2e["record"] = ("ids, lon, lat, species of the rank's strip in storage order: copied inside the step "
                             "(lm_record_next_step_ids); bytes of rank 0's strip at the last step")
        if packer is not None:
            # counted from the copies issued: int16 lon + lat, the escape list and its counter (or a plain key frame
            # when the list overflowed), plus int8 species
            e2e["d2h"] = (packer.bytes_d2h - packed_bytes0) / args.steps + n_per_gpu
            e2e["record"] = "delta16 packed (lossless), stored packed: not decoded inside the timed region; last record: " + last_packed[0]
            lon_chk = sim2.download()[0]
        elif world == 1:
            lon_chk = rec[(args.steps - 1) & 1][0].numpy()
        elif os.environ.get("LM_E2E_VARIANT", "") == "":
            lon_chk = sim2.ss.record_view(sim2.k)[1]
            assert lon_chk.size == sim2.engine.state_size()
        else:
            lon_chk = None
        if lon_chk is not None:
            assert np.isfinite(lon_chk).all() and lon_chk.min() > 100.0
        if world > 1 and host_in_record:
            e2e["host_ms_inside_record_to_host"] = 1e3 * float(np.mean(host_in_record[-args.steps:]))
        sim = sim2

    # max over ranks
    vals = torch.tensor([ms, e2e["ms"] if e2e else 0.0, float(st.n_pairs), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, ms_e2e = float(mx[0]), float(mx[1])
        pairs_total, launches_total = float(sm[2]), int(sm[3])
    else:
        ms_e2e = e2e["ms"] if e2e else None
        pairs_total, launches_total = float(st.n_pairs), int(launches)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    n_total = n_per_gpu * world
    value = n_total * args.steps / (ms * 1e-3)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    field_bytes = 0 if interact_only else 2 * 2 * hfs.u.shape[1] * hfs.u.shape[2] * 4   # two snapshots of U and V
    # bytes per microbe-step; interaction only: no advection R/W (16 B)
    b_alg = (10.0 if interact_only else 26.0) + 8.0 * rho + field_bytes / float(n_per_gpu)
    # Dominant kernel: find_pairs_kernel<RPS,EMIT> -- ONE launch per step (radius search + Philox decision bits +
    # pair list + hand-off), timed live with CUDA events recorded around it on the launching stream
    # (LM_STEP_TIMING).  Algorithmic bytes of the pair search: read lon/lat 8 B per microbe, write
    # the pair list 8 B per pair = (8 + 8 rho) N per launch.  `traffic` = dram__bytes_read + dram__bytes_write of the
    # same kernel from the committed ncu --set full capture of this workload (profiles/), per launch.
    pair_bytes = (8.0 + 8.0 * rho) * n_per_gpu
    pair_gbs = pair_bytes / (phase[2] * 1e-3) / 1e9 if phase[2] > 0 else 0.0
    traffic = None
    prof_path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof_path):
        prof = json.load(open(prof_path)).get(args.workload, {}).get("find_pairs_kernel")
        if prof and prof.get("microbes_per_gpu") == n_per_gpu:
            traffic = prof["dram_bytes_read"] + prof["dram_bytes_write"]
    if args.interact_mode == 1:
        # fused tile kernel: pair search + RPS in one launch: (10 + 8 rho) N algorithmic bytes (positions 8, species R 1 + W 1, pairs 8 rho)
        kname, kbytes, kms = "interact_tile_kernel<RPS> (1 launch per step)", (10.0 + 8.0 * rho) * n_per_gpu, float(phase[2])
        knote = ("shared-memory tiles: DRAM traffic ~1.02x the algorithmic bytes, but ~135 warp instructions per microbe at 16 of 32 "
                 "lanes: issue-bound (profiles/)")
        kprof = "interact_tile_kernel"
    else:
        kname, kbytes, kms = "find_pairs_kernel<RPS,EMIT> (1 launch per step)", pair_bytes, float(phase[2])
        knote = ("issue-bound, not HBM-bound: ~2,000 warp instructions per 32 microbes (13 distance tests + 4.4 Philox2x32-10 draws "
                 "per microbe + the hand-off layout), sm__throughput ~70 % of peak in ncu")
        kprof = "find_pairs_kernel"
    kgbs = kbytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    traffic = None
    if os.path.exists(prof_path):
        prof = json.load(open(prof_path)).get(args.workload, {}).get(kprof)
        if prof and prof.get("microbes_per_gpu") == n_per_gpu:
            traffic = prof["dram_bytes_read"] + prof["dram_bytes_write"]
    roofline = {"kernel": kname, "bound": "hbm", "achieved": kgbs, "peak": peak, "unit": "GB/s", "frac": kgbs / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": kbytes, "launch_ms": kms, "note": knote}
    line = {
        "metric": METRIC if not interact_only else "microbe-steps/sec (interact only: BASELINE config 5)", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": ("f32 state; f32 RK4 arithmetic (positions within 1e-6 relative of the f64 RK4), f64-exact pair predicate"
                  if args.advect_mode == 1 else "f32 state / f64 arithmetic"), "data": "synthetic", "config": config,
        "pairs_per_step": pairs_total, "pairs_per_s": pairs_total * args.steps / (ms * 1e-3), "rho": rho,
        "gpu_launches": launches_total,
        "parity": parity,
        "clocks": clocks,
        "phases_ms": {"advect": float(phase[0]), "bin": float(phase[1]), "pair_search": float(phase[2]),
                      "rps_resolve": float(phase[3]), "stats": float(phase[4])},
        "step_roofline": {"b_alg_bytes_per_microbe_step": b_alg, "achieved_gbs_per_gpu": value * b_alg / 1e9 / world,
                          "frac": value * b_alg / 1e9 / world / peak},
        "roofline": roofline,
    }
    if e2e:
        line["e2e"] = {"value": n_total * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                       "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": ms_e2e / args.steps}
        for key in ("host_ms_inside_record_to_host",):
            if key in e2e:
                line["e2e"][key] = e2e[key]
        if "record" in e2e:
            line["e2e"]["record"] = e2e["record"]
    if world == 1 and not args.no_cpu_baseline and not interact_only:
        res = run_cpu_reference(args.workload, args.cpu_sample, 2, 1, args.modes, spinup=spinup)
        line["cpu_baseline"] = {"value": res["value"], "unit": UNIT, "cores": res["threads"], "kind": "port",
                                "sample": ("2 steps of the configuration itself, all %d microbes (%.0f pairs/step)"
                                           if args.workload in ("config1", "config2") else
                                           "2 steps of a %d-microbe sub-box at the workload's areal density (%.0f pairs/step)")
                                          % (res["n_sample"], res["pairs_per_step"]),
                                "phases_s_per_step": res["phases_s_per_step"]}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
