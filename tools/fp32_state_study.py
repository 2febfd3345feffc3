#!/usr/bin/env python
"""Decides SURVEY section 7's open question by measurement: can vertex state be kept in fp32 (16-byte {x,y,z,tag} records)?

CPU only.  The C oracle (bit-identical to the reference in its normal mode) gets a study mode that rounds the positions
(mode 1) or positions, velocities and previous positions (mode 3) to fp32 after EVERY write, exactly what an fp32 record would
keep; the element arithmetic is untouched (it is fp32 after an fp64 difference in the reference anyway, Fem.cpp:453).
Free-running trajectories of the demo's Box L block (3072 tets, lock-left, gravity, nu = 0.495 and 0.5, undamped and with the
web default damping) at the web substep rate (3000/s) and the native one (20 000/s), compared with the fp64-state run:
centre of mass, volume ratio, total energy.  The reference's own FMA-vs-no-FMA spread (oracle/_ref fast vs strict build) is
the noise floor the differences are judged against.  Output: profiles/r2_fp32_state_study.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import bindings as ob


def run(kind, rate, frames, poisson, damped, mode):
    nodes, idx = ob.generate_tet_block(8, 8)
    if kind == "oracle":
        sc = ob.OracleScene(nodes, idx)
        sc.set_state_precision(mode)
    else:
        sc = ob.RefScene.mesh(nodes, idx, kind=kind)
    dt = np.float32(1.0 / rate)
    st = ob.make_settings(energy=ob.Energy_MixedSel, simultaneous=True, poisson=poisson, substeps_per_second=float(rate))
    if damped:  # web default, time-corrected as Sim::Update does (Demo.cpp:51-63)
        st.damping, st.pbdDamping, st.drag = 0.005, 0.03, 0.002
        st.flags = (st.flags & ~(3 << ob.Settings_RayleighTypeBit)) | (ob.Rayleigh_PostAmortized << ob.Settings_RayleighTypeBit)
        sdt = 1.0 / rate
        k = (20.0 / 100.0) / 31.0
        st.volumeAndTimeCorrectedPbdDamping = (1.0 - (1.0 - 0.03) ** (1000.0 * sdt)) * 6.0 * k * k
        st.amortizedVolumeAndTimeCorrectedPbdDamping = (1.0 - (1.0 - 0.03) ** (8000.0 * sdt)) * 6.0 * k * k
        st.timeCorrectedDrag = 1.0 - (1.0 - 0.002) ** (1000.0 * sdt)
    per_frame = int(round(rate / 60.0))
    v0 = sc.volume()
    out = {"com": [], "vol": [], "tip": []}
    tip = int(np.argmax(sc.get_state()[0][:, 0]))
    for f in range(frames):
        sc.substep(st, dt, per_frame)
        st.tickId += per_frame
        if f % 10 == 9 or f == frames - 1:
            X, V, w = sc.get_state()
            out["com"].append(X.mean(axis=0).tolist())
            out["tip"].append(X[tip].tolist())
            out["vol"].append(sc.volume() / v0)
    out["bbox"] = float(np.ptp(sc.get_rest()[0], axis=0).max()) if hasattr(sc, "get_rest") else None
    return out


def main():
    frames = int(os.environ.get("XF_STUDY_FRAMES", "1000"))
    cases = []
    for rate, nfr in ((3000, frames), (20000, max(100, frames // 4))):
        for poisson in (0.495, 0.5):
            for damped in (False, True):
                t0 = time.time()
                base = run("oracle", rate, nfr, poisson, damped, 0)
                runs = {"fp32_X": run("oracle", rate, nfr, poisson, damped, 1), "fp32_XVO": run("oracle", rate, nfr, poisson, damped, 3)}
                if ob.have_ref("fast"):
                    runs["reference_fma_build"] = run("fast", rate, nfr, poisson, damped, 0)
                bbox = base["bbox"]
                rec = {"substeps_per_s": rate, "frames": nfr, "poisson": poisson, "damped_webdefault": damped, "bbox": bbox,
                       "volume_ratio_fp64_final": base["vol"][-1], "tip_deflection_fp64_final": base["tip"][-1]}
                b_tip, b_com = np.array(base["tip"]), np.array(base["com"])
                for name, r in runs.items():
                    tipd = np.abs(np.array(r["tip"]) - b_tip).max(axis=1) / bbox
                    comd = np.abs(np.array(r["com"]) - b_com).max(axis=1) / bbox
                    rec[name] = {"tip_gap_over_bbox_max": float(tipd.max()), "tip_gap_over_bbox_final": float(tipd[-1]),
                                 "com_gap_over_bbox_max": float(comd.max()), "volume_ratio_final": r["vol"][-1],
                                 "volume_ratio_gap_max": float(np.abs(np.array(r["vol"]) - np.array(base["vol"])).max())}
                rec["seconds"] = time.time() - t0
                cases.append(rec)
                print(json.dumps(rec), flush=True)
    with open(os.path.join(ROOT, "profiles", "r2_fp32_state_study.json"), "w") as f:
        json.dump({"what": __doc__, "cases": cases}, f, indent=1)


if __name__ == "__main__":
    main()
