"""Throughput of the css_decode_sim drop-in (both sectors, per-shot channel update, logical checks) on a B200,
next to the per-shot reference-shaped loop driven by the CPU oracle (tests/ref_harness.py, one core)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bp_osd_b200 import codes
from bp_osd_b200.css_decode_sim import css_decode_sim

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", type=int, default=2)
ap.add_argument("--p", type=float, default=0.05)
ap.add_argument("--shots", type=int, default=400000)
ap.add_argument("--cpu-shots", type=int, default=150)
ap.add_argument("--update", default="x->z")
a = ap.parse_args()
code = codes.config_code(a.cfg)
kw = dict(bp_method="ms", ms_scaling_factor=0, max_iter=0, osd_method="osd_cs", osd_order=7)
upd = None if a.update == "none" else a.update
out = {"cfg": a.cfg, "p": a.p, "channel_update": upd, "decoder": kw}
for bs in (65536,):
    sim = css_decode_sim(code.hx, code.hz, error_rate=a.p, xyz_error_bias=[1, 1, 1], target_runs=bs, seed=11, channel_update=upd,
                         tqdm_disable=1, run_sim=0, batch_size=bs, check_code=0, **kw)
    sim.run_decode_sim()                       # warm-up: workspaces, layouts
    sim = css_decode_sim(code.hx, code.hz, error_rate=a.p, xyz_error_bias=[1, 1, 1], target_runs=a.shots, seed=12, channel_update=upd,
                         tqdm_disable=1, run_sim=0, batch_size=bs, check_code=0, error_bar_precision_cutoff=0, **kw)
    t = time.perf_counter()
    res = json.loads(json.loads(sim.run_decode_sim()))
    dt = time.perf_counter() - t
    out["gpu"] = {"shots": res["run_count"], "shots_per_s": res["run_count"] / dt, "batch_size": bs,
                  "osdw_logical_error_rate": res["osdw_logical_error_rate"], "osdw_eb": res["osdw_logical_error_rate_eb"],
                  "osd0_logical_error_rate": res["osd0_logical_error_rate"], "bp_logical_error_rate": res["bp_logical_error_rate"],
                  "bp_converge_x": res["bp_converge_count_x"], "bp_converge_z": res["bp_converge_count_z"]}
from tests import ref_harness
t = time.perf_counter()
ref = ref_harness.run(code.hx, code.hz, sim.lx, sim.lz, np.array(sim.channel_probs_x), np.array(sim.channel_probs_y),
                      np.array(sim.channel_probs_z), 12, a.cpu_shots, upd, **kw)
dt = time.perf_counter() - t
out["cpu_port_one_core"] = {"shots": a.cpu_shots, "shots_per_s": a.cpu_shots / dt,
                            "osdw_logical_error_rate": 1 - ref["osdw_success_count"] / a.cpu_shots}
print(json.dumps(out))
