"""css_decode_sim on the GPU against a per-shot reference-shaped loop driven by the CPU oracle."""
import json

import numpy as np
import pytest

from bp_osd_b200 import codes
from bp_osd_b200.hgp import hgp

pytestmark = pytest.mark.gpu

COUNTERS = ("run_count", "bp_converge_count_x", "bp_converge_count_z", "bp_success_count", "osd0_success_count",
            "osdw_success_count", "min_logical_weight")


@pytest.fixture(scope="module")
def sim_cls(cuda_lib):
    import torch
    assert torch.cuda.is_available()
    from bp_osd_b200.css_decode_sim import css_decode_sim
    return css_decode_sim


@pytest.mark.parametrize("update,bias,rotate", [("x->z", [1, 1, 1], 0), ("z->x", [1, 2, 3], 0), (None, [1, 1, 1], 0),
                                                 ("x->z", [0, 0, 1], 0), ("x->z", [3, 1, 1], 1)])
def test_counters_equal_per_shot_loop(sim_cls, oracle_mod, update, bias, rotate):
    from tests import ref_harness
    code = hgp(codes.rep_code(4))   # d=4 surface code, [[25,1,4]]
    kw = dict(bp_method="ms", ms_scaling_factor=0, max_iter=6, osd_method="osd_cs", osd_order=3)
    shots = 1500
    sim = sim_cls(code.hx, code.hz, error_rate=0.12, xyz_error_bias=bias, target_runs=shots, seed=1234,
                  channel_update=update, hadamard_rotate=rotate, hadamard_rotate_sector1_length=16,
                  tqdm_disable=1, run_sim=0, batch_size=512, **kw)
    text = sim.run_decode_sim()
    got = json.loads(json.loads(text))   # doubly encoded, as in the reference (css_decode_sim.py:555,567)
    ref = ref_harness.run(code.hx, code.hz, sim.lx, sim.lz, np.array(sim.channel_probs_x), np.array(sim.channel_probs_y),
                          np.array(sim.channel_probs_z), 1234, shots, update, **kw)
    for k in COUNTERS:
        assert got[k] == ref[k], (k, got[k], ref[k])
    n = shots
    assert got["osdw_logical_error_rate"] == pytest.approx(1 - ref["osdw_success_count"] / n)
    ler = got["osdw_logical_error_rate"]
    assert got["osdw_logical_error_rate_eb"] == pytest.approx(np.sqrt((1 - ler) * ler / n))
    assert got["K"] == 1 and got["N"] == 25
    for key in ("bp_logical_error_rate", "osd0_logical_error_rate", "osdw_word_error_rate", "runtime", "start_date",
                "seed", "error_rate", "target_runs", "osd_order"):
        assert key in got


def test_batch_size_does_not_change_the_result(sim_cls):
    code = hgp(codes.rep_code(3))
    outs = []
    for bs in (64, 1000):
        sim = sim_cls(code.hx, code.hz, error_rate=0.1, target_runs=1000, seed=77, tqdm_disable=1, run_sim=0,
                      max_iter=4, osd_order=2, batch_size=bs)
        outs.append(json.loads(json.loads(sim.run_decode_sim())))
    for k in COUNTERS:
        assert outs[0][k] == outs[1][k]


def test_resume_and_output_file(sim_cls, tmp_path):
    code = hgp(codes.rep_code(3))
    f = tmp_path / "out.json"
    sim = sim_cls(code.hx, code.hz, error_rate=0.1, target_runs=300, seed=5, tqdm_disable=1, output_file=str(f),
                  max_iter=4, osd_order=2)
    saved = json.loads(f.read_text())
    assert saved["run_count"] == 300 and "hx" not in saved and "channel_probs_x" not in saved
    saved["target_runs"] = 500
    sim2 = sim_cls(code.hx, code.hz, **saved)       # a dumped dict resumes the run (css_decode_sim.py:87-88,511)
    assert sim2.run_count == 500 and sim2.osdw_success_count >= sim.osdw_success_count
    assert sim2.seed != 5                            # re-randomised when run_count != 0 (css_decode_sim.py:135-136)


def test_invalid_code_raises(sim_cls):
    h = codes.rep_code(3)
    with pytest.raises(Exception, match="invalid CSS code"):
        sim_cls(h, h, error_rate=0.1, run_sim=0)     # README.md:125-136: hx = hz = rep code is not a CSS code


def test_reference_style_script_through_the_shim(cuda_lib, tmp_path):
    """The flow of the reference's example and README (examples/qldpc_decode_example.py, README.md:150-216) written with
    the REFERENCE's import lines -- `from bposd.hgp import hgp`, `from bposd.css_decode_sim import css_decode_sim`,
    `from bposd import bposd_decoder` -- runs on the B200 decoder through the bposd/ import shim."""
    script = '''
import numpy as np
from bposd.hgp import hgp
from bposd.css_decode_sim import css_decode_sim
from bposd import bposd_decoder

h = np.array(H_SEED)
qcode = hgp(h)
qcode.test()
osd_options = {"error_rate": 0.05, "target_runs": 600, "xyz_error_bias": [0, 0, 1], "output_file": OUT, "bp_method": "ms",
               "ms_scaling_factor": 0, "osd_method": "osd_cs", "osd_order": 42, "channel_update": None, "seed": 42,
               "max_iter": 0, "tqdm_disable": 1}
lk = css_decode_sim(hx=qcode.hx, hz=qcode.hz, **osd_options)
bpd = bposd_decoder(qcode.hz, error_rate=0.05, channel_probs=[None], max_iter=qcode.N, bp_method="ms", ms_scaling_factor=0,
                    osd_method="osd_cs", osd_order=7)
error = np.zeros(qcode.N).astype(int)
error[[5, 12]] = 1
bpd.decode(qcode.hz @ error % 2)
residual = (bpd.osdw_decoding + error) % 2
RESULT.update(N=lk.N, K=lk.K, runs=lk.run_count, ler=lk.osdw_logical_error_rate, ok=not (qcode.lz @ residual % 2).any())
'''
    out = str(tmp_path / "test.json")
    env = dict(H_SEED=np.asarray(codes.mkmn_16_4_6()).tolist(), OUT=out, RESULT={})
    exec(compile(script, "reference_style_script", "exec"), env)
    r = env["RESULT"]
    assert (r["N"], r["K"]) == (400, 16) and r["runs"] == 600 and r["ok"]
    assert 0 <= r["ler"] < 0.2
    assert json.load(open(out))["run_count"] == 600   # the output file holds the JSON text of the run, as in the reference
