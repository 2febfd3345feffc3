"""Worker for the partitioned-mesh tests and measurements; launched once per rank by torch.distributed.run.

  --mode emulate : CPU only (gloo).  Each rank builds the host-only partition plan (device = -1) and EMULATES the
                   partitioned algorithm with the C oracle as the element solver and gloo send/recv for the halo
                   lists; rank 0 compares the gathered result with the unpartitioned oracle (bit-exact).
  --mode gpu     : one GPU per rank (nccl only moves the 128-byte IPC blobs); the library steps the mesh with
                   in-kernel peer stores; rank 0 compares with the unpartitioned oracle (bit-exact) and prints timing.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from __graft_entry__ import load_package

ap = argparse.ArgumentParser()
ap.add_argument("--mode", choices=["emulate", "gpu"], required=True)
ap.add_argument("--dims", type=int, nargs=2, default=[8, 4])
ap.add_argument("--wonk", type=float, default=0.2)
ap.add_argument("--pattern", type=int, default=0)
ap.add_argument("--energy", type=int, default=7)
ap.add_argument("--serial", action="store_true")
ap.add_argument("--poisson", type=float, default=0.5)
ap.add_argument("--substeps", type=int, default=12)
ap.add_argument("--no-hint", action="store_true")
ap.add_argument("--check", type=int, default=1)
ap.add_argument("--time-substeps", type=int, default=0)
ap.add_argument("--time-calls", type=int, default=1, help="back-to-back calls of --time-substeps substeps in the timed region")
ap.add_argument("--schedule", choices=["auto", "dataflow", "persistent", "per_color"], default="dataflow")
ap.add_argument("--partition", choices=["slabs", "graph"], default="slabs")
ap.add_argument("--damping", type=float, default=0.0, help="Rayleigh damping (with --rayleigh) + PBD damping 0.03: gpu mode only")
ap.add_argument("--rayleigh", type=int, default=3)
ap.add_argument("--volume-passes", type=int, default=0)
ap.add_argument("--mix", action="store_true", help="plain calls before and after the damped ones: the two cross-GPU protocols hand over to each other")
ap.add_argument("--mesh", choices=["block", "armadillo"], default="block", help="armadillo: the reference's asset (needs oracle/_ref)")
a = ap.parse_args()
xf = load_package()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
DT = np.float32(1.0 / 3000.0)
AUTO_RESIZE = False
if a.mesh == "armadillo":  # Demo::UpdateSettings' Armadillo path: autoResize = true, density 2 (Demo.cpp:123, 336)
    from oracle import bindings as ob_mesh
    nodes, idx = ob_mesh.RefScene.armadillo().get_mesh()
    hint, AUTO_RESIZE = None, True
else:
    nodes, idx, hint = xf.GenerateTetBlock(a.dims[0], a.dims[1], wonkiness=a.wonk, pattern=a.pattern)
    hint = None if (a.no_hint or a.pattern != 0) else hint
PARTITION = {"slabs": xf.PARTITION_SLABS, "graph": xf.PARTITION_GRAPH}[a.partition]
DENSITY = 2.0 if AUTO_RESIZE else 1.0
kw = dict(energy=a.energy, simultaneous=not a.serial, poisson=a.poisson, volume_passes=a.volume_passes)
if a.damping > 0.0:
    kw.update(damping=a.damping, rayleigh=a.rayleigh, pbd_damping=0.03, drag_tc=0.0007)
GENERAL = a.damping > 0.0 or a.volume_passes > 0


def with_constants(st):
    if a.damping > 0.0:
        st.volumeAndTimeCorrectedPbdDamping = 1e-6
        st.amortizedVolumeAndTimeCorrectedPbdDamping = 7e-6
    return st



PLAIN_KW = dict(energy=a.energy, simultaneous=not a.serial, poisson=a.poisson)
# the calls of a run: (settings keywords, substeps); --mix wraps the damped calls in plain ones
CALLS = [(kw, 1), (kw, a.substeps - 1)]
if a.mix:
    CALLS = [(PLAIN_KW, 2)] + CALLS + [(PLAIN_KW, 3), (kw, 2)]


def run_calls(scene, make_settings, step):
    tick = 0
    for k, n in CALLS:
        st = make_settings(**k)
        if k is kw:
            with_constants(st)
        st.tickId = tick
        step(scene, st, n)
        tick += n


def reference_state(order, n):
    from oracle import bindings as ob
    o = ob.OracleScene(nodes, idx, DENSITY, AUTO_RESIZE)
    o.set_order(order)
    run_calls(o, ob.make_settings, lambda sc, st, k: sc.substep(st, DT, k))
    return o.get_state()


if a.mode == "emulate":
    from oracle import bindings as ob
    dist.init_process_group("gloo")
    part = xf.GeoPartitionCuda(nodes, idx, world, rank, device=-1, color_hint=hint, partition=PARTITION, density=DENSITY, auto_resize=AUTO_RESIZE)
    l2g = part.local_verts()
    elems, cs = part.local_elements()
    peers = part.peers()
    tets = idx.reshape(-1, 5)[:, 1:]
    g2l = np.full(part.nVGlobal, -1, dtype=np.int64)
    g2l[l2g] = np.arange(part.nV)
    local_stream = np.empty((part.nT, 5), dtype=np.uint32)
    local_stream[:, 0] = 4
    local_stream[:, 1:] = g2l[tets[elems]]
    # the local sub-mesh with the FULL mesh's rest positions (after autoResize), masses and flags
    full = ob.OracleScene(nodes, idx, DENSITY, AUTO_RESIZE)
    X0full = full.get_rest()[0]
    sub = ob.OracleScene(X0full[l2g].astype(np.float32).reshape(-1), local_stream.reshape(-1))
    sub.set_order(np.arange(part.nT, dtype=np.uint32))
    w, flags = part.initial()
    sub.copy_elements_from(full, elems, rest=X0full[l2g])
    sub.set_state(w=w)
    sub.set_flags(flags)
    # global serial position of every local element (the amortised damping slices are ranges of the FULL mesh's serial order)
    full_order = part.get_order()
    serial_pos = np.empty(part.nTGlobal, dtype=np.int64)
    serial_pos[full_order] = np.arange(part.nTGlobal)
    local_pos = serial_pos[elems]
    gcs = part.global_color_start()
    tag = [0]

    def exchange(c, which):
        """after a phase of colour c: push the values of the vertices this rank's elements of that colour touch to the other copies"""
        X, V, ww = sub.get_state()
        arr = X if which == "X" else V
        reqs, bufs = [], []
        tag[0] += 1
        for slot, q in enumerate(peers):
            snd = part.halo(c, slot, True)
            rcv = part.halo(c, slot, False)
            out = torch.from_numpy(np.ascontiguousarray(arr[snd]))
            inn = torch.empty((len(rcv), 3), dtype=torch.float64)
            if len(snd):
                reqs.append(dist.isend(out, int(q), tag=tag[0]))
            if len(rcv):
                reqs.append(dist.irecv(inn, int(q), tag=tag[0]))
            bufs.append((rcv, inn, out))
        for r in reqs:
            r.wait()
        for rcv, inn, _ in bufs:
            if len(rcv):
                arr[rcv] = inn.numpy()
        if which == "X":
            sub.set_state(X=arr)
        else:
            sub.set_state(V=arr)

    def emulate(st, n):
        rayleigh = (st.flags >> ob.Settings_RayleighTypeBit) & 3
        do_damp = rayleigh >= 2 and st.damping > 0.0
        do_pbd = st.pbdDamping > 0.0
        for s_ in range(n):
            sub.phase_predict(st, DT)
            for c in range(part.nColors):
                sub.phase_elems(st, DT, 0, np.arange(cs[c], cs[c + 1]))
                exchange(c, "X")
            for _ in range(st.volumePasses):
                for c in range(part.nColors):
                    sub.phase_elems(st, DT, 1, np.arange(cs[c], cs[c + 1]))
                    exchange(c, "X")
            sub.phase_post(st, DT)
            if do_damp or do_pbd:
                lo, hi = 0, part.nTGlobal
                if rayleigh == 3:
                    k = st.tickId % 8
                    lo, hi = part.nTGlobal * k // 8, part.nTGlobal * (k + 1) // 8
                for kind, on in ((2, do_damp), (3, do_pbd)):
                    if not on:
                        continue
                    for c in range(part.nColors):
                        if not (gcs[c] < hi and gcs[c + 1] > lo):
                            continue  # the same verdict on every rank
                        mine = np.arange(cs[c], cs[c + 1])
                        mine = mine[(local_pos[mine] >= lo) & (local_pos[mine] < hi)]
                        sub.phase_elems(st, DT, kind, mine)
                        exchange(c, "V")
            st.tickId += 1

    run_calls(sub, ob.make_settings, lambda sc, st, k: emulate(st, k))
    X, V, ww = sub.get_state()
    order = part.get_order()
else:
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    part = xf.GeoPartitionCuda(nodes, idx, world, rank, device=local_rank, color_hint=hint, partition=PARTITION, density=DENSITY, auto_resize=AUTO_RESIZE,
                               schedule={"auto": xf.SCHEDULE_AUTO, "dataflow": xf.SCHEDULE_DATAFLOW, "persistent": xf.SCHEDULE_PERSISTENT,
                                         "per_color": xf.SCHEDULE_LAUNCH_PER_COLOR}[a.schedule])
    blob = torch.from_numpy(part.ipc_export()).cuda()
    allb = [torch.empty_like(blob) for _ in range(world)]
    dist.all_gather(allb, blob)
    part.ipc_connect(torch.stack(allb).cpu().numpy())
    dist.barrier()
    st = with_constants(xf.make_settings(**kw))
    run_calls(part, xf.make_settings, lambda sc, s_, k: sc.Substep(s_, DT, k))
    X, V, ww = part.get_state()
    l2g = part.local_verts()
    order = part.get_order()
    timing = None
    if a.time_substeps:
        dist.barrier()
        part.Substep(st, DT, 5)
        part.Sync()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(a.time_calls):
            part.Substep(st, DT, a.time_substeps)
        part.Sync()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        timing = float(el.item())

# gather to rank 0 and compare with the unpartitioned oracle
gathered = [None] * world
dist.gather_object((l2g, X, V, ww), gathered if rank == 0 else None, dst=0)
if rank == 0:
    nVg = nodes.size // 3
    ok = True
    msg = ""
    if a.check:
        Xo, Vo, wo = reference_state(order, a.substeps)
        for r, (g, Xr, Vr, wr) in enumerate(gathered):
            if not (np.array_equal(Xr, Xo[g]) and np.array_equal(Vr, Vo[g]) and np.array_equal(wr, wo[g])):
                ok = False
                msg = "rank %d differs: max|dX| = %.3e" % (r, np.abs(Xr - Xo[g]).max())
    copies = np.zeros(nVg, dtype=np.int64)
    for g, _, _, _ in gathered:
        copies[g] += 1
    out = {"mode": a.mode, "schedule": a.schedule, "partition": a.partition, "max_copies": int(copies.max()), "world": world, "tets": int(idx.size // 5), "verts": int(nVg), "ok": ok, "msg": msg,
           "shared_verts_rank0": int(sum(len(part.halo(c, s, True)) for c in range(part.nColors) for s in range(part.nPeers)))}
    if a.mode == "gpu" and a.time_substeps:
        out["us_per_substep"] = 1e6 * timing / (a.time_substeps * a.time_calls)
        out["element_substeps_per_s"] = (idx.size // 5) * a.time_substeps * a.time_calls / timing
    print("PART_RESULT " + json.dumps(out))
dist.barrier()
dist.destroy_process_group()
