"""Recipe for ``oracle/_ref/``: the reference's OWN Python for the quantized-linear path, staged
where the GPU box can run it.

TEST / BASELINE INFRASTRUCTURE (never imported by the product).  The reference is 100 % Python:
there is nothing to compile, so "building" it means staging the unmodified sources of the path --
``fake_quant/`` (ActQuantWrapper, quantizers, observers, hadamard_utils) and ``model/`` (QWenBlock,
model/modeling_qwen_opt.py:390-420) -- from ``/root/reference`` into ``oracle/_ref/``.  That
directory is git-ignored (no reference source enters the history) but not gpurun-ignored, so like
the built ``.so`` it travels to the GPU box, where ``bench.py --impl reference`` and the
``cpu_baseline`` leg time the REAL reference wrappers on the host cores
(``cpu_baseline.kind = "reference"``).  ``__graft_entry__.build()`` calls this when
``/root/reference`` is present; on a box without it the previously staged copy is used.

    python oracle/build_ref.py            # stage (idempotent)
"""
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
PACKAGES = ("fake_quant", "model")


def stage(src=REF_SRC, dst=REF_DST):
    """Copy the .py files of the path's packages; returns dst, or None when src is absent."""
    if not os.path.isdir(os.path.join(src, "fake_quant")):
        return dst if available(dst) else None
    for pkg in PACKAGES:
        for root, _dirs, files in os.walk(os.path.join(src, pkg)):
            rel = os.path.relpath(root, src)
            for f in files:
                if not f.endswith(".py"):
                    continue
                out = os.path.join(dst, rel, f)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                s = os.path.join(root, f)
                if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(s) or \
                        os.path.getsize(out) != os.path.getsize(s):
                    shutil.copy2(s, out)
    with open(os.path.join(dst, "STAGED_FROM"), "w") as fh:
        fh.write(src + "\n")
    return dst


def available(dst=REF_DST):
    return os.path.exists(os.path.join(dst, "fake_quant", "quant_utils.py")) and \
        os.path.exists(os.path.join(dst, "model", "modeling_qwen_opt.py"))


def import_staged(dst=REF_DST):
    """Import the staged reference on the CPU with the two stubs of SURVEY.md 8(c)
    (``fast_hadamard_transform``: fp32 butterfly, result in x.dtype; ``accelerate``: unused names).
    Returns (quant_utils, hadamard_utils, modeling_qwen_opt)."""
    import torch
    import transformers  # noqa: F401  (before the stubs: it probes for accelerate)

    if "fast_hadamard_transform" not in sys.modules:
        fht = types.ModuleType("fast_hadamard_transform")

        def hadamard_transform(x, scale=1.0):
            n = x.shape[-1]
            t = x.float().reshape(-1, n)
            h = 1
            while h < n:  # Sylvester butterflies, fp32 inside like the CUDA extension
                t = t.reshape(-1, n // (2 * h), 2, h)
                t = torch.stack((t[:, :, 0, :] + t[:, :, 1, :], t[:, :, 0, :] - t[:, :, 1, :]), dim=2).reshape(-1, n)
                h *= 2
            return (t.reshape(x.shape) * float(scale)).to(x.dtype)

        fht.hadamard_transform = hadamard_transform
        sys.modules["fast_hadamard_transform"] = fht
    if "accelerate" not in sys.modules:
        acc = types.ModuleType("accelerate")
        acc.dispatch_model = None
        acc.infer_auto_device_map = None
        accu = types.ModuleType("accelerate.utils")
        accu.get_balanced_memory = None
        acc.utils = accu
        sys.modules["accelerate"] = acc
        sys.modules["accelerate.utils"] = accu
    if dst not in sys.path:
        sys.path.insert(0, dst)
    from fake_quant import hadamard_utils, quant_utils
    from model import modeling_qwen_opt
    return quant_utils, hadamard_utils, modeling_qwen_opt


if __name__ == "__main__":
    d = stage()
    print("staged:" if d else "reference not available:", d or REF_SRC)
