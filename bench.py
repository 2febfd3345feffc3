#!/usr/bin/env python
"""bench.py — element-substeps/s of the small-step XPBD tet hot path on N B200s (one process per GPU).

A "step" is one 60-fps frame of the reference's web default: `substeps_per_step` = 3000/60 = 50 calls of
Geo::Substep on the headline scene: a 55x55x55 MeshGen tet block (998 250 tets,
175 616 verts), Energy_YeohSkinFast, simultaneous solve, nu = 0.5, compliance 1, gravity (0,-0.4905),
dt = 1/3000, lock-left, ground plane just below the block.  Data is synthetic (the reference's own
generator, restated in xf_generate_tet_block).

  value      whole-job element-substeps/s with the state resident in HBM (CUDA events, max over ranks)
  e2e        the same metric through the C ABI with HOST buffers: every step uploads X,V from pinned host
             memory, runs the substeps, and downloads X,V (what a Geo user does once per frame)
  roofline   SURVEY §8d: B_HBM = 56 + 112*nV/nT bytes per element-substep against the measured HBM copy peak
             (headline, the working set exceeds half of L2); the L2 figure (184 B) is reported beside it
  cpu_baseline  the unmodified reference (oracle/_ref, -O3 -mavx2 -mfma) on one host core, bounded sample

N > 1: every rank steps its own independent scene (batch sharding, no data-path collective) => weak scaling.
`--impl reference` times the reference's own CPU code on all host cores (one independent scene per thread).
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

METRIC = "element-substeps/sec"
UNIT = "element-substeps/s"
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--cells", type=int, default=55, help="block is cells^3 hexes -> 6*cells^3 tets")
    ap.add_argument("--substeps-per-step", type=int, default=50)
    ap.add_argument("--precision", choices=["exact", "fast"], default="exact")
    ap.add_argument("--schedule", choices=["dataflow", "bricks", "persistent", "per_color"], default="dataflow")
    ap.add_argument("--grouping", choices=["auto", "elements", "chains", "clusters"], default="auto",
                    help="xf_grouping: chains = vertex records shared with the thread's next element stay in private shared memory")
    ap.add_argument("--energy", choices=["yeohskinfast", "mixedsel", "mixed", "yeohskin"], default="yeohskinfast")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hint-order", choices=["ring", "type"], default="ring",
                    help="numbering of the 24 lattice colour classes: ring = 6*class + type (xf_generate_tet_block), type = 4*type + class")
    ap.add_argument("--no-hint", action="store_true", help="use the generic colouring instead of the lattice 24-colouring")
    return ap.parse_args()


ENERGY_IDS = {"mixed": 3, "mixedsel": 4, "yeohskin": 5, "yeohskinfast": 7}


def workload_name(args):
    return "meshgen_tet_block_%dx%dx%d_%s_nu0.5_simultaneous_dt1/3000_lockleft_ground" % (args.cells, args.cells, args.cells, args.energy)


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smmax, reasons = [], [], set()
        for t, line in self.lines:
            if t < t0 - 0.05 or t > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples inside the timed region"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smmax), "reasons": sorted(reasons), "samples": len(sm)}


def make_scene(xf, args, device, stream):
    nodes, idx, hint = xf.GenerateTetBlock(args.cells, args.cells)
    if args.hint_order == "type":
        hint = (4 * (hint % 6) + hint // 6).astype(hint.dtype)
    geo = xf.GeoLinear3dCuda(nodes, idx, device=device, stream=stream,
                             precision=xf.PRECISION_EXACT if args.precision == "exact" else xf.PRECISION_FAST,
                             schedule={"dataflow": xf.SCHEDULE_DATAFLOW, "bricks": xf.SCHEDULE_BRICKS, "persistent": xf.SCHEDULE_PERSISTENT, "per_color": xf.SCHEDULE_LAUNCH_PER_COLOR}[args.schedule],
                             color_hint=None if args.no_hint else hint,
                             grouping={"auto": xf.GROUPING_AUTO, "elements": xf.GROUPING_ELEMENTS, "chains": xf.GROUPING_CHAINS,
                                       "clusters": xf.GROUPING_CLUSTERS}[args.grouping])
    y_min = float(nodes.reshape(-1, 3)[:, 1].min())
    geo.set_ground(True, y_min - 1.0e-3, 0.0)
    st = xf.make_settings(energy=ENERGY_IDS[args.energy], simultaneous=True, poisson=0.5, compliance=1.0, gravity=(0.0, -0.4905),
                          lock_left=True)
    return geo, st, nodes, idx


def cpu_baseline_leg(args, kind_pref="fast"):
    """Reference CPU path on ONE core, bounded sample of the same workload (oracle/ is the checker here, never the product)."""
    import numpy as np
    from oracle import bindings as ob
    dt = np.float32(1.0 / 3000.0)
    st = ob.make_settings(energy=ENERGY_IDS[args.energy], simultaneous=True, poisson=0.5)
    if ob.have_ref(kind_pref):
        scene = ob.RefScene.block(args.cells, args.cells, kind=kind_pref)
        kind = "reference"
    else:
        nodes, idx = ob.generate_tet_block(args.cells, args.cells)
        scene = ob.OracleScene(nodes, idx)
        kind = "port"
    scene.time_substeps(st, dt, 1)  # warm-up
    n, spent = 0, 0.0
    while spent < 8.0 and n < 400:
        spent += scene.time_substeps(st, dt, 2)
        n += 2
    value = scene.nT * n / spent
    return {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "%d substeps of the same %d-tet scene, 1 thread, g++ -O3 -mavx2 -mfma" % (n, scene.nT)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation, one independent scene per host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import bindings as ob
    dt = np.float32(1.0 / 3000.0)
    st = ob.make_settings(energy=ENERGY_IDS[args.energy], simultaneous=True, poisson=0.5)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = max(1, min(cores, 64))
    use_ref = ob.have_ref("fast")
    scenes = [None] * threads

    def build(i):
        if use_ref:
            scenes[i] = ob.RefScene.block(args.cells, args.cells, kind="fast")
        else:
            nodes, idx = ob.generate_tet_block(args.cells, args.cells)
            scenes[i] = ob.OracleScene(nodes, idx)

    ts = [threading.Thread(target=build, args=(i,)) for i in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    nT = scenes[0].nT
    sub_per_step = 1  # bounded sample: one substep per scene per step (the frame is 50; throughput is per substep)

    def step_all():
        th = [threading.Thread(target=lambda s=s: s.substep(st, dt, sub_per_step)) for s in scenes]
        [t.start() for t in th]
        [t.join() for t in th]

    for _ in range(args.warmup):
        step_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_all()
    elapsed = time.perf_counter() - t0
    value = threads * nT * sub_per_step * args.steps / elapsed
    sample = "%d independent %d-tet scenes (one per host thread), %d substep per step" % (threads, nT, sub_per_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 math / f64 state", "data": "synthetic",
        "config": {"workload": workload_name(args), "substeps_per_step": sub_per_step, "note": "CPU reference, bounded sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference" if use_ref else "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    xf = load_package()
    xf.lib()  # fail loudly if the CUDA library is missing

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    # a real (non-default) stream: torch's default stream handle is 0, which the C ABI reads as "make your own"
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    geo, st, nodes, idx = make_scene(xf, args, local_rank, stream)
    nT, nV = geo.nT, geo.nV
    dt = np.float32(1.0 / 3000.0)
    sub = args.substeps_per_step
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > L2 (126 MB)

    def one_step():
        geo.Substep(st, dt, sub)

    # ---- device-resident throughput ----
    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = geo.info()["kernelLaunches"]
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    torch.cuda.synchronize()
    t0 = time.time()
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pair)
        starts[k].record()
        one_step()
        ends[k].record()
    torch.cuda.synchronize()
    t1 = time.time()
    barrier()
    clocks = sampler.stop(t0, t1)
    launches = geo.info()["kernelLaunches"] - launches0
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    ms_t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = world * nT * sub * args.steps / (ms_max * 1e-3)
    info = geo.info()

    # ---- end to end through the C ABI with pinned host buffers ----
    hX = torch.empty((nV, 3), dtype=torch.float64).pin_memory()
    hV = torch.empty((nV, 3), dtype=torch.float64).pin_memory()
    X0, V0, _ = geo.get_state()
    hX.copy_(torch.from_numpy(X0))
    hV.copy_(torch.from_numpy(V0))

    def e2e_step():
        geo.set_state_async(hX.data_ptr(), hV.data_ptr())
        geo.Substep(st, dt, sub)
        geo.get_state_async(hX.data_ptr(), hV.data_ptr())
        geo.Sync()  # the host reads the result every frame

    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    barrier()
    e_ms = 0.0
    for k in range(args.steps):
        flush.zero_()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record()
        e2e_step()
        e_ev.record()
        torch.cuda.synchronize()
        e_ms += s_ev.elapsed_time(e_ev)
    e_t = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * nT * sub * args.steps / (float(e_t.item()) * 1e-3)
    state_bytes = 2 * nV * 3 * 8

    # sanity: the state must be finite and volume preserved, else the number is meaningless
    stats = geo.stats(st)
    if stats["nonfinite"] != 0:
        raise RuntimeError("simulation blew up during the benchmark")

    if rank == 0:
        peak, peak_src = hbm_peak()
        b_hbm = 56.0 + 112.0 * nV / nT
        b_l2 = 184.0
        kernel_ms = ms_max / args.steps  # persistent schedule: the whole step is ONE launch of the dominant kernel
        per_launch_units = nT * sub
        if args.schedule == "per_color":
            per_launch_units = None
        ach_hbm = (nT * sub * b_hbm) / (kernel_ms * 1e-3) / 1e9
        ach_l2 = (nT * sub * b_l2) / (kernel_ms * 1e-3) / 1e9
        ws = 56 * nT + 48 * nV
        kernel = {"dataflow": "k_substeps_dataflow", "bricks": "k_substeps_bricks", "persistent": "k_substeps_persistent",
                  "per_color": "k_sweep_color (x colours)"}[args.schedule]
        if args.schedule == "dataflow" and info["chainedPermille"] > 0 and info["maxColorSize"] <= info["gridBlocks"] * info["blockThreads"]:
            kernel = "k_substeps_chain"
        elif args.schedule == "dataflow" and args.grouping == "clusters":
            kernel = "k_substeps_cluster"
        traffic = None  # dram__bytes_read+write of ONE launch of the dominant kernel, from the committed ncu captures
        try:
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tj = json.load(f)
            for cap in tj.get("captures", [tj]):
                if (args.cells, args.precision, args.energy) == (55, "exact", "yeohskinfast") and cap["kernel"].split("<")[0] == kernel:
                    traffic = cap["dram_bytes_per_element_substep"] * nT * sub  # per launch of `sub` substeps
        except Exception:
            traffic = None
        roofline = {
            "bound": "hbm", "achieved": ach_hbm, "peak": peak, "unit": "GB/s", "frac": ach_hbm / peak, "traffic": traffic,
            "traffic_source": "profiles/r1_traffic.json (ncu --set full, same kernel and workload)" if traffic else None,
            "algorithmic_bytes_per_launch": nT * sub * b_hbm,
            "peak_source": peak_src, "kernel": kernel,
            "bytes_per_element_substep": b_hbm, "units_per_launch": per_launch_units, "kernel_ms": kernel_ms,
            "achieved_l2_gbs": ach_l2, "bytes_per_element_substep_l2": b_l2, "working_set_bytes": ws, "l2_bytes": info["l2Bytes"],
            "headline_rule": "SURVEY 8d names the L2 figure when the working set is <= l2/2 (here: %s); `achieved`/`frac` always use "
                             "the conservative HBM figure against the measured HBM peak, the L2 figure is reported beside it"
                             % ("yes" if ws <= info["l2Bytes"] / 2 else "no"),
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 math / f64 state (%s)" % args.precision, "data": "synthetic",
            "config": {"workload": workload_name(args), "tets": nT, "verts": nV, "colors": info["colorCount"],
                       "substeps_per_step": sub, "precision": args.precision, "schedule": args.schedule,
                       "grouping": args.grouping, "chained_permille": info["chainedPermille"], "hint_order": args.hint_order,
                       "grid": [info["gridBlocks"], info["blockThreads"]], "l2": "flushed between timed steps (512 MiB memset)",
                       "sharding": "one independent scene per GPU, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "volume_ratio": stats["volume"] / (nT and float(np.float64(geo.get_elements()["volume"].astype(np.float64).sum()))),
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_leg(args)
        elif not args.no_cpu_baseline:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "measured at N=1 only"}
        print(json.dumps(line))
    geo.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
