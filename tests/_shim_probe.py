"""Helper of tests/test_shim_cpu.py (run as a subprocess, once per mode):

  shim  the reference's OWN rotation / GPTQ modules imported through ``compat/`` (our quant_utils,
        hadamard_utils, utils and fast_hadamard_transform underneath)
  ref   the unmodified reference with the two stubs of oracle/_ref_import.py

Both run the same seeded CPU workload and dump the results; the test compares them."""
import os
import sys
import types

import numpy as np
import torch
import transformers  # noqa: F401  (before the stubs: it probes for accelerate)

mode, out = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
for name in ("unfoldNd", "accelerate", "accelerate.utils"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["unfoldNd"].UnfoldNd = object
sys.modules["accelerate"].dispatch_model = None
sys.modules["accelerate"].infer_auto_device_map = None
sys.modules["accelerate.utils"].get_balanced_memory = None
torch.cuda.synchronize = lambda *a, **k: None  # gptq_utils.fasterquant calls it unconditionally
torch.Tensor.cuda = lambda self, *a, **k: self  # hadamard_utils.py:146 moves weights to the GPU (none here)
if mode == "shim":
    os.environ["MQUANT_REFERENCE_ROOT"] = REF
    sys.path[:0] = [os.path.join(ROOT, "compat"), ROOT]
else:
    sys.path.insert(0, ROOT)
    from oracle._ref_import import import_reference
    import_reference()

import fake_quant  # noqa: E402
from fake_quant import hadamard_utils, quant_utils, rotation_utils, utils  # noqa: E402,F401
from fake_quant.gptq import gptq_utils  # noqa: E402
import fast_hadamard_transform  # noqa: E402
for sub in ("qwenvl_rotation" if os.path.exists(os.path.join(REF, "fake_quant", "qwenvl_rotation.py")) else "rotation_utils",
            "qwen2vl_rotation", "internvl_rotation", "minicpmv_rotation"):
    __import__("fake_quant." + sub)
for sub in ("qwenvl_gptq_plus", "qwen2vl_gptq_plus", "internvl_gptq_plus", "minicpmv_gptq_plus"):
    __import__("fake_quant.gptq." + sub)

res = {"quant_utils_file": np.array(quant_utils.__file__), "fht_file": np.array(getattr(fast_hadamard_transform, "__file__", "stub"))}
torch.manual_seed(0)
# --- GPTQ (gptq_utils.py:160-303) on top of WeightQuantizer (quant_utils.py:415-524)
for mse in (False, True):
    lin = torch.nn.Linear(96, 40, bias=False)
    g = gptq_utils.GPTQ(lin)
    g.quantizer = quant_utils.WeightQuantizer()
    g.quantizer.configure(4, perchannel=True, sym=True, mse=mse)
    for _ in range(3):
        x = torch.randn(2, 17, 96)
        g.add_batch(x, None)
    g.fasterquant(blocksize=32, percdamp=0.01, groupsize=-1, actorder=False, static_groups=False)
    res[f"gptq_w_mse{int(mse)}"] = lin.weight.data.numpy().copy()
    res[f"gptq_scale_mse{int(mse)}"] = g.quantizer.scale.numpy().copy()
if mode == "shim":
    # mquant_b200.gptq.accelerate_reference(): the kernel-backed add_batch / fasterquant are installed into the
    # reference's class; a configuration the kernels do not take (here: CPU tensors) runs the reference's own loop
    from mquant_b200 import gptq as mq_gptq

    assert mq_gptq.accelerate_reference() and gptq_utils.GPTQ.add_batch is mq_gptq.add_batch
    rng_state = torch.get_rng_state()  # (the cases below must see the same random stream in both modes)
    torch.manual_seed(0)
    lin = torch.nn.Linear(96, 40, bias=False)
    g = gptq_utils.GPTQ(lin)
    g.quantizer = quant_utils.WeightQuantizer()
    g.quantizer.configure(4, perchannel=True, sym=True, mse=False)
    for _ in range(3):
        g.add_batch(torch.randn(2, 17, 96), None)
    g.fasterquant(blocksize=32, percdamp=0.01, groupsize=-1, actorder=False, static_groups=False)
    res["gptq_w_accelerated_cpu"] = lin.weight.data.numpy().copy()
    torch.set_rng_state(rng_state)
# --- RTN (rtn path of *_gptq_plus.py)
W = torch.randn(24, 64)
wq = quant_utils.WeightQuantizer()
wq.configure(4, perchannel=True, sym=True, mse=False)
wq.find_params(W)
res["rtn_w"] = wq.quantize(W).numpy()
# --- offline rotation helpers (rotation_utils.py:15-60, hadamard_utils.py:135-191)
ln = torch.nn.LayerNorm(64)
ln.weight.data = torch.rand(64) + 0.5
ln.bias.data = torch.randn(64) * 0.1
l2 = torch.nn.Linear(64, 48)
rotation_utils.fuse_ln_linear(ln, [l2])
res["fuse_w"], res["fuse_b"] = l2.weight.data.numpy().copy(), l2.bias.data.numpy().copy()
for n, had_dim, output in ((128, -1, False), (128, -1, True), (256, 64, False), (256, 64, True)):
    l3 = torch.nn.Linear(n, n)
    hadamard_utils.apply_exact_had_to_linear(l3, had_dim=had_dim, output=output)
    res[f"had_w_{n}_{had_dim}_{int(output)}"] = l3.weight.data.numpy().copy()
    res[f"had_b_{n}_{had_dim}_{int(output)}"] = l3.bias.data.numpy().copy()
res["rand_had"] = hadamard_utils.random_hadamard_matrix(64, "cpu").numpy()
x = torch.randn(5, 3, 256)
res["fht_direct"] = fast_hadamard_transform.hadamard_transform(x, scale=0.0625).numpy()
np.savez(out, **res)
print("ok", mode)
