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
  roofline   SURVEY §8d, both figures, reproducible from the line alone: `achieved`/`frac` = B_HBM = 56 + 112*nV/nT bytes per
             element-substep against the measured HBM copy peak (the conservative figure, always the headline here);
             `achieved_l2_gbs`/`frac_l2` = B_L2 = 184 B against `l2_peak_gbs`, an L2-resident 256-bit copy measured live by
             xf_debug_l2_bandwidth (read + write bytes).  `working_set_bytes` vs `l2_bytes` says which one SURVEY's rule names.
             `frac_l2_bytes_vs_hbm_peak` = B_L2 bytes against the HBM peak, the arithmetic of SURVEY 8d's own cross-check of the
             north_star target (informational).
  cpu_baseline  the unmodified reference (oracle/_ref, -O3 -mavx2 -mfma) on one host core: median of 5 timed windows after
             a warm-up, bounded to ~20 s
  extra      driver-visible numbers of the other BASELINE configs, same process group (skipped with --no-extras):
             extra.damped       the headline scene with the web demo's default damping (ui.js:76-88)
             extra.native_rate  one frame at the reference's native 20 000 substeps/s (333 substeps): ms per frame / fps
             extra.batch        4096 x Box L and 4096 x Beam L sharded over the ranks (config 3): value, e2e, issue figure
             extra.partitioned  ONE mesh over the N ranks through k_part_dataflow (config 4: 150^3 = 20.25M tets at N >= 2,
                                the same mesh on the single-GPU kernel at N = 1), with `parity_ok` from a 384k-tet bit-exact
                                check against the oracle; the process exits non-zero on a mismatch

N > 1: `value` = every rank steps its own independent scene (batch sharding, no data-path collective) => weak scaling.
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
    ap.add_argument("--schedule", choices=["dataflow", "persistent", "per_color"], default="dataflow")
    ap.add_argument("--grouping", choices=["auto", "elements", "chains", "clusters"], default="auto",
                    help="xf_grouping: chains = vertex records shared with the thread's next element stay in private shared memory")
    ap.add_argument("--energy", choices=["yeohskinfast", "mixedsel", "mixed", "yeohskin"], default="yeohskinfast")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline leg only (profiling runs)")
    ap.add_argument("--extras", default="damped,native_rate,batch,partitioned", help="comma list of extra legs")
    ap.add_argument("--part-cells", type=int, default=150, help="extra.partitioned: cells^3 hexes (150 -> 20.25M tets)")
    ap.add_argument("--batch-scenes", type=int, default=4096)
    ap.add_argument("--damped", action="store_true", help="headline leg with the web demo's default damping (ui.js:76-88)")
    ap.add_argument("--hint-order", choices=["ring", "type"], default="ring",
                    help="numbering of the 24 lattice colour classes: ring = 6*class + type (xf_generate_tet_block), type = 4*type + class")
    ap.add_argument("--no-hint", action="store_true", help="use the generic colouring instead of the lattice 24-colouring")
    return ap.parse_args()


ENERGY_IDS = {"mixed": 3, "mixedsel": 4, "yeohskin": 5, "yeohskinfast": 7}


def workload_name(args):
    return "meshgen_tet_block_%dx%dx%d_%s_nu0.5_simultaneous_dt1/3000_lockleft_ground" % (args.cells, args.cells, args.cells, args.energy)


def shared_config(args, nT, nV):
    """The workload-defining part of `config`, identical in both arms (b200 and --impl reference)."""
    return {"workload": workload_name(args) + ("_damped_webdefault" if args.damped else ""), "tets": int(nT), "verts": int(nV),
            "substeps_per_step": args.substeps_per_step,
            "l2": "b200 arm: flushed between timed steps (512 MiB memset outside the event pairs); cpu arm: not applicable",
            "throughput_unit": "per substep (one element-substep = one SolveElement of one tet inside one Geo::Substep)"}


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
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
                             schedule={"dataflow": xf.SCHEDULE_DATAFLOW, "persistent": xf.SCHEDULE_PERSISTENT, "per_color": xf.SCHEDULE_LAUNCH_PER_COLOR}[args.schedule],
                             color_hint=None if args.no_hint else hint,
                             grouping={"auto": xf.GROUPING_AUTO, "elements": xf.GROUPING_ELEMENTS, "chains": xf.GROUPING_CHAINS,
                                       "clusters": xf.GROUPING_CLUSTERS}[args.grouping])
    y_min = float(nodes.reshape(-1, 3)[:, 1].min())
    geo.set_ground(True, y_min - 1.0e-3, 0.0)
    st = xf.make_settings(energy=ENERGY_IDS[args.energy], simultaneous=True, poisson=0.5, compliance=1.0, gravity=(0.0, -0.4905),
                          lock_left=True)
    return geo, st, nodes, idx


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_baseline_leg(args, kind_pref="fast"):
    """Reference CPU path on ONE core, bounded sample of the same workload (oracle/ is the checker here, never the product):
    warm-up, then the median of 5 timed windows, ~20 s in total at 1M tets."""
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
    # size the windows from one probe substep: ~3 s of warm-up (at most 100 substeps), five windows of ~3 s
    probe = scene.time_substeps(st, dt, 1)
    warm = int(min(100, max(1, 3.0 / max(probe, 1e-6))))
    scene.time_substeps(st, dt, warm)
    per_window = int(min(2000, max(2, 3.0 / max(probe, 1e-6))))
    windows = [scene.nT * per_window / scene.time_substeps(st, dt, per_window) for _ in range(5)]
    return {"value": statistics.median(windows), "unit": UNIT, "cores": 1, "kind": kind, "cpu": cpu_model(),
            "windows": windows,
            "sample": "median of 5 windows of %d substeps (after %d warm-up substeps) of the same %d-tet scene, 1 thread, "
                      "g++ -O3 -mavx2 -mfma; throughput is per substep" % (per_window, warm, scene.nT)}


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
    # bounded sample: ONE substep per scene per step of this arm (a frame of the b200 arm is `substeps_per_step` substeps of the
    # same scene with the same settings; the metric is throughput per substep, so the two lines are comparable as they stand)
    sub_per_step = 1

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
        "config": shared_config(args, nT, scenes[0].nV),
        "note": "CPU reference, bounded sample: %d substep per scene per step of this arm (a b200 step is %d substeps); the metric is "
                "throughput per substep, same scene and settings as the b200 arm; cpu: %s" % (sub_per_step, args.substeps_per_step, cpu_model()),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference" if use_ref else "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def source_hash():
    """sha256/16 of the sources of the dominant kernel: ties profiles/r2_traffic.json (ncu capture) to the build that is measured."""
    import hashlib
    h = hashlib.sha256()
    for name in ("xf_dataflow.cu", "xf_dataflow.cuh", "xf_dataflow_general.cu", "xf_element.cuh", "xf_element_packed.cuh", "xf_phase.cuh", "xf_scene.h",
                 "xf_dispatch.cuh"):
        with open(os.path.join(ROOT, "xpbd-fem_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def load_traffic(kernel, args):
    """dram__bytes_read + dram__bytes_write of ONE launch of the dominant kernel, from the ncu capture tools/ncu_traffic.sh commits.
    A capture taken from other kernel sources than the ones being measured is STALE: reported loudly, never used."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        with open(path) as f:
            tj = json.load(f)
    except Exception:
        return None, "no capture (run tools/ncu_traffic.sh)"
    want = (args.cells, args.precision, args.energy, args.substeps_per_step, bool(args.damped))
    for cap in tj.get("captures", []):
        have = (cap.get("cells"), cap.get("precision"), cap.get("energy"), cap.get("substeps_per_launch"), bool(cap.get("damped", False)))
        if cap["kernel"].split("<")[0] != kernel or have != want:
            continue
        if cap.get("source_hash") != source_hash():
            sys.stderr.write("bench.py: profiles/r2_traffic.json was captured from other kernel sources (%s, now %s): roofline.traffic "
                             "is null until tools/ncu_traffic.sh is re-run\n" % (cap.get("source_hash"), source_hash()))
            return None, "STALE capture (kernel sources changed since tools/ncu_traffic.sh ran)"
        return float(cap["dram_bytes_per_launch"]), "profiles/r2_traffic.json (ncu --set full of this build, %s)" % cap.get("source", "")
    return None, "no capture for this flag combination (run tools/ncu_traffic.sh with the same flags)"


def web_default_damping(xf, st):
    """The web demo's default damping (wasm/ui.js:76-88) with Sim::Update's time-corrected constants (Demo.cpp:51-63) at 3000 substeps/s."""
    st.damping = 0.005
    st.pbdDamping = 0.03
    st.drag = 0.002
    st.flags = (st.flags & ~(3 << xf.Settings_RayleighTypeBit)) | (xf.Rayleigh_PostAmortized << xf.Settings_RayleighTypeBit)
    fs = xf.new_frame_state()
    xf.frame_constants(st, fs)  # bookkeeping only (scene = NULL): fills the derived constants the way Sim::Update does
    return st


class Bench:
    """Shared plumbing of the legs: process group, stream, timing with CUDA events on the launching stream, max over ranks."""

    def __init__(self, args):
        import numpy as np
        import torch
        import torch.distributed as dist
        from __graft_entry__ import load_package
        self.np, self.torch, self.dist, self.args = np, torch, dist, args
        self.xf = load_package()
        self.xf.lib()  # fail loudly if the CUDA library is missing
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        # a real (non-default) stream: torch's default stream handle is 0, which the C ABI reads as "make your own"
        self.tstream = torch.cuda.Stream()
        torch.cuda.set_stream(self.tstream)
        self.stream = self.tstream.cuda_stream
        assert self.stream != 0
        self.flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > L2 (126 MB)
        self.dt = np.float32(1.0 / 3000.0)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step, steps, warmup, flush=True, sync_each=False):
        """warm-up, barrier + synchronize, `steps` x (L2 flush, event, step, event), synchronize + barrier; ms summed over the steps,
        max over ranks.  Returns (ms_total_max, wall t0, wall t1)."""
        torch = self.torch
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        t0 = time.time()
        for a, b in ev:
            if flush:
                self.flush.zero_()  # L2 flush between timed iterations (outside the event pair)
            a.record()
            step()
            b.record()
            if sync_each:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        t1 = time.time()
        self.barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        return self.max_over_ranks(ms), t0, t1


def leg_batch(B, shape, scenes):
    """BASELINE config 3: `scenes` independent small scenes sharded over the ranks by scene index, no data-path collective."""
    xf, np, torch, args = B.xf, B.np, B.torch, B.args
    first, count = xf.shard_scenes(scenes, B.world, B.rank)
    dims = (8, 2) if shape == "beamL" else (8, 8)
    nodes, idx, hint = xf.GenerateTetBlock(*dims)
    b = xf.GeoBatchCuda(nodes, idx, count, device=B.local_rank, precision=xf.PRECISION_EXACT, color_hint=hint, stream=B.stream)
    arr = (xf.Settings * count)()
    for k in range(count):
        s = first + k  # global scene index: per-scene gravity / compliance do not depend on the sharding
        arr[k] = xf.make_settings(energy=7, poisson=0.5, gravity=(0.0, -0.4905 * (1 + 0.1 * (s % 7))), compliance=1.0 + 0.25 * (s % 4))
    sub, steps = args.substeps_per_step, 5
    l0 = b.info()["launches"]
    ms, _, _ = B.timed(lambda: b.Substep(arr, B.dt, sub), steps, 3, flush=False)
    launches = b.info()["launches"] - l0 - 3
    value = scenes * b.nT * sub * steps / (ms * 1e-3)
    # e2e: per step the host uploads X,V of every scene of this rank (pinned), steps, and reads X,V back
    n3 = count * b.nV * 3
    hX = torch.empty(n3, dtype=torch.float64).pin_memory()
    hV = torch.empty(n3, dtype=torch.float64).pin_memory()
    X0, V0, _ = b.get_state()
    hX.copy_(torch.from_numpy(X0.reshape(-1)))
    hV.copy_(torch.from_numpy(V0.reshape(-1)))
    aX, aV = hX.numpy().reshape(count, b.nV, 3), hV.numpy().reshape(count, b.nV, 3)
    L = xf.lib()
    import ctypes as C

    def e2e_step():
        xf._check(L.xf_batch_set_state(b._h, 0, count, C.c_void_p(hX.data_ptr()), C.c_void_p(hV.data_ptr()), None))
        b.Substep(arr, B.dt, sub)
        xf._check(L.xf_batch_get_state(b._h, 0, count, C.c_void_p(hX.data_ptr()), C.c_void_p(hV.data_ptr()), None))

    e_ms, _, _ = B.timed(e2e_step, steps, 2, flush=False, sync_each=True)
    finite = bool(np.isfinite(aX).all() and np.isfinite(aV).all())
    info = b.info()
    out = {"workload": "%d x %s (%d tets, %d verts each), yeohskinfast nu=0.5 simultaneous, per-scene gravity/compliance from the global scene index"
                       % (scenes, shape, b.nT, b.nV),
           "scenes": scenes, "scenes_rank0": count, "sharding": "xf.shard_scenes: contiguous scene ranges, no collective",
           "value": value, "unit": UNIT, "ms_per_step": ms / steps, "substeps_per_step": sub, "steps": steps,
           "e2e": {"value": scenes * b.nT * sub * steps / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * n3 * 8, "d2h_bytes_per_step": 2 * n3 * 8},
           "gpu_launches": int(launches), "kernel": "k_batch_substeps", "group_threads": info["groupThreads"], "block_threads": info["blockThreads"],
           "smem_bytes": info["smemBytes"], "finite": finite}
    # issue-bound figure: warp instructions per element-substep (static property of the kernel, from the committed ncu capture)
    # x achieved element-substeps/s against the SMs' issue peak (4 schedulers x 1 warp instruction per clock)
    try:
        with open(os.path.join(ROOT, "profiles", "r2_batch_inst.json")) as f:
            bi = json.load(f)[shape]
        smc = xf.GeoLinear3dCuda  # noqa (documentation anchor)
        prop = torch.cuda.get_device_properties(B.local_rank)
        peak = prop.multi_processor_count * 4 * 1.965e9
        per_gpu = value / B.world
        out["roofline"] = {"bound": "issue", "warp_inst_per_element_substep": bi["warp_inst_per_element_substep"],
                           "achieved": per_gpu * bi["warp_inst_per_element_substep"], "peak": peak, "unit": "warp-inst/s",
                           "frac": per_gpu * bi["warp_inst_per_element_substep"] / peak, "source": bi["source"],
                           "note": "state lives in shared memory for the whole launch, element planes are L1/L2 resident: no HBM or L2 bound applies"}
    except Exception:
        out["roofline"] = None
    b.close()
    if not finite:
        raise RuntimeError("batched scenes blew up during the benchmark")
    return out


def part_connect(B, part):
    torch, dist = B.torch, B.dist
    blob = torch.from_numpy(part.ipc_export()).cuda()
    allb = [torch.empty_like(blob) for _ in range(B.world)]
    dist.all_gather(allb, blob)
    part.ipc_connect(torch.stack(allb).cpu().numpy())
    dist.barrier()


def leg_partitioned(B):
    """BASELINE config 4: ONE mesh over the N ranks (x-slabs), versioned records mirrored by in-kernel peer stores over NVLink
    (k_part_dataflow); NCCL only carries the 128-byte IPC handles and the timing reduction.  N = 1: the same mesh on the
    single-GPU barrier-free kernel (strong-scaling anchor)."""
    xf, np, torch, dist, args = B.xf, B.np, B.torch, B.dist, B.args
    out = {"substeps_per_step": args.substeps_per_step}
    st = xf.make_settings(energy=xf.Energy_MixedSel, simultaneous=True, poisson=0.5)
    # ---- parity: 40^3 = 384k tets, 8 substeps, against the unpartitioned oracle in the schedule's equivalent serial order
    pn, pi, ph = xf.GenerateTetBlock(40, 40, wonkiness=0.2)
    if B.world > 1:
        part = xf.GeoPartitionCuda(pn, pi, B.world, B.rank, device=B.local_rank, color_hint=ph, stream=B.stream)
        part_connect(B, part)
        for n in (1, 7):
            part.Substep(st, B.dt, n)
        X, V, w = part.get_state()
        l2g, order = part.local_verts(), part.get_order()
        dist.barrier()  # nobody unmaps / frees while a peer may still hold the mapping
        part.close()
        dist.barrier()
        gathered = [None] * B.world
        dist.gather_object((l2g, X, V, w), gathered if B.rank == 0 else None, dst=0)
    else:
        geo = xf.GeoLinear3dCuda(pn, pi, device=B.local_rank, stream=B.stream, color_hint=ph)
        for n in (1, 7):
            geo.Substep(st, B.dt, n)
        X, V, w = geo.get_state()
        order = geo.get_order()
        geo.close()
        gathered = [(np.arange(X.shape[0]), X, V, w)]
    ok = True
    if B.rank == 0:
        from oracle import bindings as ob  # the checker, never the thing measured
        o = ob.OracleScene(pn, pi)
        o.set_order(order)
        o.substep(ob.make_settings(energy=ob.Energy_MixedSel, simultaneous=True, poisson=0.5), B.dt, 8)
        Xo, Vo, wo = o.get_state()
        for g, Xr, Vr, wr in gathered:
            ok = ok and bool(np.array_equal(Xr, Xo[g]) and np.array_equal(Vr, Vo[g]) and np.array_equal(wr, wo[g]))
    ok = B.max_over_ranks(0.0 if ok else 1.0) == 0.0
    out["parity_ok"] = ok
    out["parity_case"] = "40^3 block (384 000 tets, wonkiness 0.2), MixedSel nu=0.5, 8 substeps in 2 calls, X,V,w bit-exact against the C oracle"
    if not ok:
        return out
    # ---- timing
    c = args.part_cells
    nodes, idx, hint = xf.GenerateTetBlock(c, c)
    sub, steps = args.substeps_per_step, 4
    if B.world > 1:
        # The barrier-free cross-GPU schedule is fail-stop: its two store-order hazards are closed by timing margins,
        # and a violation ends in a REPORTED stall, never in a wrong result.  One such stall was seen in this leg (20M tets, 4 GPUs,
        # profiles/r2_bench_n4_stall.err) among otherwise clean runs, so a stall here is recorded and the leg is measured again on the
        # flag protocol (ordered by construction) instead of failing the whole line.
        # The stalls are intermittent (3 of ~14 runs at 20M tets), so the barrier-free schedule gets ONE more attempt on a fresh partition
        # before the fallback; every stall stays in `stalls` whatever is measured in the end.
        attempts = [("dataflow", xf.SCHEDULE_AUTO, "k_part_dataflow"), ("dataflow (second attempt)", xf.SCHEDULE_AUTO, "k_part_dataflow"),
                    ("flag protocol (per-colour launches)", xf.SCHEDULE_LAUNCH_PER_COLOR, "k_part_sweep (x colours)")]
        for name, schedule, kernel in attempts:
            part = xf.GeoPartitionCuda(nodes, idx, B.world, B.rank, device=B.local_rank, color_hint=hint, stream=B.stream, schedule=schedule)
            part_connect(B, part)
            nT, nV = part.nTGlobal, part.nVGlobal
            l0 = part.info()["launches"]
            stalled, ms = 0.0, 0.0
            try:
                if os.environ.get("XF_BENCH_FAKE_STALL") and name == "dataflow":  # exercises the fallback path (tools/r2_final_fallback.sh)
                    raise xf.XfError(xf.XF_ERR_CUDA, "fake stall (XF_BENCH_FAKE_STALL)")
                ms, _, _ = B.timed(lambda: part.Substep(st, B.dt, sub), steps, 2, flush=False)
                part.Sync()
            except xf.XfError as err:
                stalled = 1.0
                out.setdefault("stalls", []).append("%s: %s" % (name, err))
                torch.cuda.synchronize()
            stalled = B.max_over_ranks(stalled)  # every rank takes the same branch
            if stalled == 0.0:
                launches = part.info()["launches"] - l0 - 2
                Xl, _, _ = part.get_state()
                finite = bool(np.isfinite(Xl).all())
                shared = int(sum(len(part.halo(cc, s, True)) for cc in range(part.nColors) for s in range(part.nPeers)))
                out.update({"kernel": kernel, "schedule": name, "local_tets_rank0": part.nT, "local_verts_rank0": part.nV, "peers_rank0": part.nPeers,
                            "shared_vertex_stores_per_substep_rank0": shared,
                            "link": "NVLink peer stores, one 32-byte record per shared vertex per writing element; nobody polls remote memory"})
            dist.barrier()
            part.close()
            dist.barrier()
            if stalled == 0.0:
                break
        else:
            out["error"] = "every schedule stalled"
            return out
    else:
        geo = xf.GeoLinear3dCuda(nodes, idx, device=B.local_rank, stream=B.stream, color_hint=hint)
        nT, nV = geo.nT, geo.nV
        l0 = geo.info()["kernelLaunches"]
        ms, _, _ = B.timed(lambda: geo.Substep(st, B.dt, sub), steps, 2, flush=False)
        launches = geo.info()["kernelLaunches"] - l0 - 2
        finite = geo.stats(st)["nonfinite"] == 0
        out.update({"kernel": "k_substeps_dataflow (one GPU: strong-scaling anchor)"})
        geo.close()
    finite = B.max_over_ranks(0.0 if finite else 1.0) == 0.0
    b_hbm = 56.0 + 112.0 * nV / nT
    value = nT * sub * steps / (ms * 1e-3)
    peak, _ = hbm_peak()
    out.update({"workload": "meshgen_tet_block_%d^3 (%d tets, %d verts), MixedSel nu=0.5 simultaneous, one mesh over %d GPU(s)" % (c, nT, nV, B.world),
                "tets": int(nT), "verts": int(nV), "value": value, "unit": UNIT, "scaling": "strong", "us_per_substep": 1e3 * ms / (steps * sub),
                "ms_per_step": ms / steps, "steps": steps, "gpu_launches": int(launches), "finite": finite,
                "roofline": {"bound": "hbm", "bytes_per_element_substep": b_hbm, "achieved": value * b_hbm / 1e9, "peak": peak * B.world,
                             "unit": "GB/s", "frac": value * b_hbm / 1e9 / (peak * B.world)}})
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    B = Bench(args)
    xf, np, torch, dist = B.xf, B.np, B.torch, B.dist
    world, rank, local_rank = B.world, B.rank, B.local_rank
    geo, st, nodes, idx = make_scene(xf, args, local_rank, B.stream)
    if args.damped:
        web_default_damping(xf, st)
    nT, nV = geo.nT, geo.nV
    dt, sub = B.dt, args.substeps_per_step
    warm = max(args.warmup, 3)

    # ---- device-resident throughput ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.2)
    launches0 = None

    def one_step():
        geo.Substep(st, dt, sub)

    for _ in range(warm):
        one_step()
    launches0 = geo.info()["kernelLaunches"]
    ms_max, t0, t1 = B.timed(one_step, args.steps, 0)
    launches = geo.info()["kernelLaunches"] - launches0
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

    e_ms, _, t2 = B.timed(e2e_step, args.steps, 3, sync_each=True)
    clocks = sampler.stop(t0, t2)
    e2e_value = world * nT * sub * args.steps / (e_ms * 1e-3)
    state_bytes = 2 * nV * 3 * 8

    # sanity: the state must be finite and volume preserved, else the number is meaningless
    stats = geo.stats(st)
    if stats["nonfinite"] != 0:
        raise RuntimeError("simulation blew up during the benchmark")
    volume_ratio = stats["volume"] / float(geo.get_elements()["volume"].astype(np.float64).sum())
    l2_peak = xf.l2_bandwidth(local_rank, 0, 16, passes=200, reps=5, blocks_per_sm=16)

    extra = {}
    failed = False
    wanted = [] if args.no_extras else [x for x in args.extras.split(",") if x]
    if "damped" in wanted and not args.damped:
        dst = web_default_damping(xf, xf.make_settings(energy=ENERGY_IDS[args.energy], simultaneous=True, poisson=0.5, compliance=1.0,
                                                       gravity=(0.0, -0.4905), lock_left=True))
        l0 = geo.info()["kernelLaunches"]
        d_ms, _, _ = B.timed(lambda: geo.Substep(dst, dt, sub), 10, 3)
        dinfo = geo.info()
        extra["damped"] = {"workload": workload_name(args) + "_damped_webdefault",
                           "settings": "Rayleigh_PostAmortized damping 0.005, pbdDamping 0.03, drag 0.002 (wasm/ui.js:76-88), time-corrected as Sim::Update does",
                           "value": world * nT * sub * 10 / (d_ms * 1e-3), "unit": UNIT, "ms_per_step": d_ms / 10, "steps": 10,
                           "vs_undamped": (ms_max / args.steps) / (d_ms / 10), "gpu_launches": int(dinfo["kernelLaunches"] - l0 - 3),
                           "last_kernel": dinfo.get("lastKernel"), "nonfinite": int(geo.stats(dst)["nonfinite"])}
    if "native_rate" in wanted:
        nst = xf.make_settings(energy=ENERGY_IDS[args.energy], simultaneous=True, poisson=0.5, compliance=1.0, gravity=(0.0, -0.4905),
                               lock_left=True, substeps_per_second=20000.0)
        ndt = np.float32(1.0 / 20000.0)
        n_ms, _, _ = B.timed(lambda: geo.Substep(nst, ndt, 333), 5, 2)
        extra["native_rate"] = {"what": "one 60-fps frame at the reference's native rate (Demo.cpp:22: 20 000 substeps/s -> 333 substeps, dt = 1/20000)",
                                "ms_per_frame": n_ms / 5, "frames_per_s": 1000.0 / (n_ms / 5), "us_per_substep": 1e3 * n_ms / (5 * 333),
                                "value": world * nT * 333 * 5 / (n_ms * 1e-3), "unit": UNIT, "target": ">= 60 frames/s"}
    geo.close()
    if "batch" in wanted:
        extra["batch"] = {shape: leg_batch(B, shape, args.batch_scenes) for shape in ("boxL", "beamL")}
    if "partitioned" in wanted:
        extra["partitioned"] = leg_partitioned(B)
        failed = failed or not extra["partitioned"]["parity_ok"]

    if rank == 0:
        peak, peak_src = hbm_peak()
        b_hbm = 56.0 + 112.0 * nV / nT
        b_l2 = 184.0
        kernel_ms = ms_max / args.steps  # barrier-free / persistent schedules: the whole step is ONE launch of the dominant kernel
        per_launch_units = nT * sub
        if args.schedule == "per_color":
            per_launch_units = None
        ach_hbm = (nT * sub * b_hbm) / (kernel_ms * 1e-3) / 1e9
        ach_l2 = (nT * sub * b_l2) / (kernel_ms * 1e-3) / 1e9
        ws = 56 * nT + 48 * nV
        kernel = info.get("lastKernel") or {"dataflow": "k_substeps_dataflow", "persistent": "k_substeps_persistent",
                                            "per_color": "k_sweep_color (x colours)"}[args.schedule]
        traffic, traffic_src = load_traffic(kernel, args)
        roofline = {
            "bound": "hbm", "achieved": ach_hbm, "peak": peak, "unit": "GB/s", "frac": ach_hbm / peak, "traffic": traffic,
            "traffic_source": traffic_src, "source_hash": source_hash(),
            "algorithmic_bytes_per_launch": nT * sub * b_hbm,
            "peak_source": peak_src, "kernel": kernel,
            "bytes_per_element_substep": b_hbm, "units_per_launch": per_launch_units, "kernel_ms": kernel_ms,
            "achieved_l2_gbs": ach_l2, "bytes_per_element_substep_l2": b_l2, "l2_peak_gbs": l2_peak, "frac_l2": ach_l2 / l2_peak,
            "l2_peak_source": "measured live: xf_debug_l2_bandwidth, L2-resident copy of 2 x 16 MiB with 256-bit accesses, read + write bytes, best of 5",
            "frac_l2_bytes_vs_hbm_peak": ach_l2 / peak,
            "frac_l2_bytes_vs_hbm_peak_note": "SURVEY 8d cross-checks north_star's '>= 50 % of the roofline' as B_L2 = 184 B x rate against the "
                                              "measured HBM copy peak (2.0e10 x 184 B = 3.7 TB/s = 56 %); this is that figure, reported beside "
                                              "the two stricter ones, never instead of them",
            "working_set_bytes": ws, "l2_bytes": info["l2Bytes"],
            "headline_rule": "SURVEY 8d names the L2 figure when the working set 56*nT + 48*nV is <= l2_bytes/2 (here: %s); `achieved`/`frac` "
                             "are ALWAYS the conservative HBM figure against the measured HBM copy peak, `achieved_l2_gbs`/`frac_l2` the L2 figure "
                             "against the measured L2 copy peak" % ("yes" if ws <= info["l2Bytes"] / 2 else "no"),
        }
        cfg = shared_config(args, nT, nV)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 math / f64 state (%s)" % args.precision, "data": "synthetic",
            "config": cfg,
            "impl_config": {"colors": info["colorCount"], "precision": args.precision, "schedule": args.schedule,
                            "grouping": args.grouping, "chained_permille": info["chainedPermille"], "hint_order": args.hint_order,
                            "grid": [info["gridBlocks"], info["blockThreads"]],
                            "sharding": "value: one independent scene per GPU, no collective; extra.batch: scenes sharded by index; "
                                        "extra.partitioned: one mesh over all ranks, NVLink peer stores"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "volume_ratio": volume_ratio,
            "extra": extra,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_leg(args)
        elif not args.no_cpu_baseline:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "measured at N=1 only"}
        print(json.dumps(line))
        if failed:
            sys.stderr.write("bench.py: PARITY FAILURE in extra.partitioned (see parity_ok)\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
