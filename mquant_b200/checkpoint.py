"""Packed W4A8 checkpoint.

The reference persists a quantised model as a whole-model ``torch.save`` pickle of fp16
fake-quantised weights and loses the calibrated activation scales every run
(exam/quant_qwenvl.py:194-196).  ``save_packed`` writes, per frozen ``QuantLinear``:

    <name>.w4         uint8 [N, Kp/2]   INT4 codes, the reference's interchange nibble order
                                        (quant_utils.py:61-68: even k -> low nibble), K zero padded
                                        to Kp (a multiple of 128)
    <name>.s_w        fp32  [N]         per-output-channel weight scales (WeightQuantizer.scale)
    <name>.bias       fp16  [N]         optional
    <name>.w0         fp16  [N]         optional: the fp channel-0 column of the RMS split
    <name>.act_scale  fp32  [1] or [2]  static activation scale(s): {text} or {text, visual} (MSQ)
    <name>.meta       int64 [8]         N, K, Kp, had_k, pad_to, online_full_had, split, fp32_had

as one flat dict of tensors (``torch.load(..., weights_only=True)`` reads it; no pickled code),
plus ``__format__``.  ``load_packed`` installs the entries into a model whose Linears are already
wrapped (``add_actquant``): the GEMM-layout weights are rebuilt on the device with
``mq_repack_w4_from_ref`` and the wrapper goes straight to the integer path -- no GPTQ, no
calibration, no fp16 weights needed.
"""
from __future__ import annotations

import torch

from . import hadamard_utils, ops
from .quant_utils import ActQuantWrapper, find_qlayers, pack_i4

FORMAT = "mquant_b200.packed.v1"


@torch.no_grad()
def packed_state_dict(model):
    """name -> tensor for every frozen wrapper under `model` (CPU tensors)."""
    out = {}
    for name, w in sorted(find_qlayers(model, layers=[ActQuantWrapper]).items()):
        if not w.real_quant:
            continue
        pk = w._packed
        codes = ops.unpack_w4(pk["w4"], pk["N"], pk["Kp"])
        out[f"{name}.w4"] = pack_i4(codes).cpu()
        out[f"{name}.s_w"] = pk["s_w"].detach().reshape(-1).float().cpu()
        if pk["bias"] is not None:
            out[f"{name}.bias"] = pk["bias"].detach().cpu()
        if pk["w0"] is not None:
            out[f"{name}.w0"] = pk["w0"].detach().cpu()
        qz = w.quantizer.quantizer
        scales = [qz.scale.detach().float().reshape(-1)[:1].cpu()]
        if qz.scale_visual is not None:
            scales.append(qz.scale_visual.detach().float().reshape(-1)[:1].cpu())
        out[f"{name}.act_scale"] = torch.cat(scales)
        out[f"{name}.act_dtype"] = torch.zeros(0, dtype=qz.scale.dtype)  # statistics keep the activation dtype
        out[f"{name}.meta"] = torch.tensor([pk["N"], pk["K"], pk["Kp"], int(w.K), int(w.pad_to or 0),
                                            int(w.online_full_had), int(w.split), int(w.fp32_had)], dtype=torch.int64)
    return out


def save_packed(model, path):
    sd = packed_state_dict(model)
    if not sd:
        raise RuntimeError("save_packed: no frozen QuantLinear found (call freeze() / model_freeze first)")
    sd["__format__"] = torch.tensor(list(FORMAT.encode()), dtype=torch.uint8)
    torch.save(sd, path)
    return len([k for k in sd if k.endswith(".meta")])


@torch.no_grad()
def load_packed(model, path_or_dict, device=None, strict=True):
    """Install a packed checkpoint into the wrapped Linears of `model`.  Returns the number of
    layers switched to the integer path."""
    sd = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu", weights_only=True)
    fmt = bytes(sd["__format__"].tolist()).decode() if "__format__" in sd else None
    if fmt != FORMAT:
        raise RuntimeError(f"load_packed: unknown format {fmt!r} (expected {FORMAT!r})")
    layers = find_qlayers(model, layers=[ActQuantWrapper])
    names = sorted(k[:-5] for k in sd if k.endswith(".meta"))
    missing = [n for n in names if n not in layers]
    if missing and strict:
        raise KeyError(f"load_packed: the model has no wrapped layer named {missing[:4]} ...")
    n_loaded = 0
    for name in names:
        if name not in layers:
            continue
        w = layers[name]
        dev = torch.device(device) if device is not None else next(
            (p.device for p in w.parameters() if p.device.type == "cuda"), torch.device("cuda", torch.cuda.current_device()))
        N, K, Kp, had_k, pad_to, full_had, split, fp32_had = (int(v) for v in sd[f"{name}.meta"].tolist())
        if isinstance(w.module, torch.nn.Linear) and (w.module.out_features != N or
                                                      max(w.module.in_features, pad_to) != max(K, pad_to)):
            raise RuntimeError(f"load_packed: {name} is {w.module.in_features}->{w.module.out_features}, "
                               f"the checkpoint holds {K}->{N}")
        w.online_full_had, w.split, w.fp32_had = bool(full_had), bool(split), bool(fp32_had)
        w.pad_to = pad_to or None
        w.K = had_k
        w.had_K = hadamard_utils.get_hadK(pad_to or K)[0] if (full_had and had_k > 1) else None
        act = sd[f"{name}.act_scale"].float()
        sdt = sd[f"{name}.act_dtype"].dtype if f"{name}.act_dtype" in sd else torch.float16
        w.quantizer.configure(bits=8, sym=True, static=True, observer_type="minmax", msq=act.numel() > 1)
        w.quantizer.quantizer.set_scales(act[0].to(sdt).to(dev), act[1].to(sdt).to(dev) if act.numel() > 1 else None)
        w.quantizer.quant = True
        w._packed = {
            "w4": ops.repack_w4_from_ref(sd[f"{name}.w4"].to(dev), N, Kp),
            "s_w": sd[f"{name}.s_w"].to(dev).contiguous(),
            "w0": sd[f"{name}.w0"].to(dev).contiguous() if f"{name}.w0" in sd else None,
            "bias": sd[f"{name}.bias"].to(dev).contiguous() if f"{name}.bias" in sd else None,
            "N": N, "K": K, "Kp": Kp,
        }
        w.real_quant = True
        n_loaded += 1
    return n_loaded
