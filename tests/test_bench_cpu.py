"""bench.py on the CPU: the reference arm's JSON contract (it runs the oracle port on the host cores, no GPU needed)
and the roofline bookkeeping (algorithmic FLOPs / bytes per unit)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=580)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train rays/s" and d["unit"] == "rays/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "hotdog-synthetic-train-step"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_flop_and_byte_accounting():
    sys.path.insert(0, ROOT)
    import bench

    # forward MACs per sample x 2 (K = 4, no specular): base 32->64->16, two heads 27->64->64->4
    fwd = 2 * (32 * 64 + 64 * 16) + 2 * 2 * (27 * 64 + 64 * 64 + 64 * 4)
    assert fwd == 30464
    assert bench.mlp_flops("umhs_field_mlp_tc_fwd", 1, 4, False) == (fwd, fwd)
    assert bench.mlp_flops("umhs_density_tc_fwd", 1, 4, False) == (6144, 6144)
    # backward = dgrad + wgrad = 2 x forward; the recomputing kernel executes 3 x, the saved-activation kernel 2 x
    assert bench.mlp_flops("umhs_field_mlp_tc_bwd", 10, 4, False) == (10 * 2 * fwd, 10 * 3 * fwd)
    assert bench.mlp_flops("umhs_field_mlp_tc_bwd_saved", 10, 4, False) == (10 * 2 * fwd, 10 * 2 * fwd)
    assert bench.mlp_flops("umhs_hash_encode_fwd_tiles", 1, 4, False) == (None, None)
    # hash kernels, tile form: 8 corners x 16 levels x 8 B gathered (read-modify-written) + point + bf16 row
    assert bench.algorithmic_bytes("umhs_hash_encode_fwd_tiles", 1, 50, 4) == 8 * 16 * 8 + 16 + 64
    assert bench.algorithmic_bytes("umhs_hash_encode_bwd_tiles", 1, 50, 4) == 64 + 16 + 2 * 8 * 16 * 8
    # Adam: touched entries move 32 B, untouched ones 12 B
    assert bench.algorithmic_bytes("umhs_adam_step_dev", 100, 50, 4, 0.25) == 100 * (32 * 0.25 + 12 * 0.75)
    # the committed ncu capture names the kernels the step launches (tile-form hash kernels through their aliases)
    for k in ("umhs_hash_encode_fwd_tiles", "umhs_hash_encode_bwd_tiles", "umhs_field_mlp_tc_bwd_saved",
              "umhs_field_mlp_tc_fwd", "umhs_adam_step_dev"):
        assert bench.NCU_TRAFFIC[k]["dram_bytes"] > 0, k
