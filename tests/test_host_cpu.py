"""CPU-side tests: host logic mirrors the reference API, the C-ABI library loads and exports every
symbol the header declares (no compute without a GPU), and the product never touches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge

    ge.build()
    import mquant_b200

    return mquant_b200


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "mquant_b200.h")).read()
    declared = set(re.findall(r"\b(mq_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mq_status"}
    from mquant_b200 import _lib

    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    h = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(h, name), name
    assert built.lib().mq_version() == 100
    assert built.lib().mq_w4_packed_bytes(300, 1024) == 2 * 8 * 16384
    assert built.lib().mq_w4_packed_bytes(300, 100) == 0


def test_no_cpu_fallback(built):
    from mquant_b200 import ops

    with pytest.raises(built.MQuantLibraryError):
        ops.act_quant_i8(torch.zeros(4, 64, dtype=torch.float16), torch.ones(2))
    q = built.quant_utils.ActQuantizer()
    q.configure(bits=8, sym=True, static=True)
    q.calibrate = True
    with pytest.raises(built.MQuantLibraryError):
        q(torch.zeros(1, 4, 64, dtype=torch.float16))


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mquant_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/mq_oracle.py:qlinear_int", ""), (dirpath, f)


def test_wrapper_api_surface(built):
    qu = built.quant_utils
    lin = torch.nn.Linear(256, 64)
    w = qu.ActQuantWrapper(lin)
    for attr in ("module", "weight", "bias", "quantizer", "out_quantizer", "had_K", "K", "online_full_had",
                 "online_partial_had", "had_dim", "fp32_had", "split"):
        assert hasattr(w, attr), attr
    assert w.had_K is None and w.K == 1 and w.quantizer.bits == 16 and qu.QuantLinear is qu.ActQuantWrapper
    w.split_weights()
    assert w.L1.weight.shape == (64, 1) and w.L2.weight.shape == (64, 255)
    assert set(qu.find_qlayers(w, layers=[torch.nn.Linear])) == {"module", "L1", "L2"}
    w.register_forward_pre_hook(lambda m, i: i)  # drivers do this (exam/quant_qwenvl.py:154)
    w.quantizer.configure(bits=8, groupsize=-1, sym=True, clip_ratio=1.0, static=True, observer_type="minmax")
    assert (w.quantizer.calibrate, w.quantizer.last_calibrate, w.quantizer.quant) == (False, False, False)
    with pytest.raises(KeyError):
        qu.ActQuantizer().configure(bits=4, sym=True, static=True)  # reference: BIT_TYPE_DICT["int4"]
    with pytest.raises(NotImplementedError):
        qu.ActQuantizer().configure(bits=8, static=True, observer_type="ema")


def test_add_actquant_and_flags(built):
    qu = built.quant_utils

    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.q_proj = torch.nn.Linear(8, 8)
            self.mlp = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 8))
            self.lst = torch.nn.ModuleList([torch.nn.Linear(8, 8), torch.nn.LayerNorm(8)])

    class Sub(torch.nn.Linear):
        pass

    m = torch.nn.ModuleList([Block(), Block()])
    m[1].sub = Sub(8, 8)  # exact-type match only: subclasses are NOT wrapped (reference :638)
    qu.add_actquant(m)
    ql = qu.find_qlayers(m, layers=[qu.ActQuantWrapper])
    assert sorted(ql) == sorted(f"{i}.{n}" for i in range(2) for n in ("q_proj", "mlp.0", "mlp.2", "lst.0"))
    assert type(m[1].sub) is Sub
    qu.add_actquant(m)  # idempotent
    assert len(qu.find_qlayers(m, layers=[qu.ActQuantWrapper])) == 8
    for l in ql.values():
        l.quantizer.configure(bits=8, sym=True, static=True)
    args = type("A", (), {"skip_names": ["mlp.2"]})()
    qu.model_open_calibrate(m, args)
    qu.model_open_last_calibrate(m, args)
    assert all(l.quantizer.calibrate == ("mlp.2" not in n) for n, l in ql.items())
    assert all(l.quantizer.last_calibrate == ("mlp.2" not in n) for n, l in ql.items())
    qu.model_close_calibrate(m, args)
    qu.model_quant(m, args)
    assert all((not l.quantizer.calibrate) and l.quantizer.quant == ("mlp.2" not in n) for n, l in ql.items())
    qu.model_no_quant(m, args)
    assert not any(l.quantizer.quant for l in ql.values())


def test_hadamard_host_utils(built, golden):
    hu = built.hadamard_utils
    g = golden("hadamard")
    for n, k in zip(g["hadk_sizes"], g["hadk_K"]):
        h, kk = hu.get_hadK(int(n))
        assert kk == k
        if k > 1:
            assert torch.equal(h @ h.T, k * torch.eye(k))
            assert torch.equal(hu.get_hadK(int(n), transpose=True)[0], h.T)
    for n, p in zip(g["pad_in"], g["pad_out"]):
        assert hu.auto_pad_size(int(n)) == p
    for n in (256, 344, 688):
        y = hu.matmul_hadU(torch.from_numpy(g[f"hadU_x_{n}"]))
        np.testing.assert_allclose(y.numpy(), g[f"hadU_y_{n}"], atol=1e-12)
    torch.manual_seed(0)
    assert torch.equal(hu.random_hadamard_matrix(64, "cpu"), torch.from_numpy(g["rand_had_64"]))
    # offline weight rotation cancels the online transform: (x H)(W H)^T = x W^T
    lin = torch.nn.Linear(344, 16, bias=False).double()
    W0 = lin.weight.data.clone()
    x = torch.randn(5, 344, dtype=torch.float64)
    hu.apply_exact_had_to_linear(lin, had_dim=-1, output=False)
    np.testing.assert_allclose((hu.matmul_hadU(x) @ lin.weight.data.double().T).numpy(), (x @ W0.T).numpy(),
                               atol=1e-5)
    with pytest.raises(RuntimeError):
        hu.matmul_hadU_cuda(torch.zeros(2, 64, dtype=torch.float16), None, 1)


def test_weight_quantizer_and_pack_match_reference(built, golden):
    qu = built.quant_utils
    g = golden("weight_quant")
    for tag, dt in (("f16", torch.float16), ("f32", torch.float32)):
        W = torch.from_numpy(g[f"{tag}_W"])
        for mse in (0, 1):
            wq = qu.WeightQuantizer()
            wq.configure(4, perchannel=True, sym=True, mse=bool(mse))
            wq.find_params(W)
            assert wq.scale.dtype == torch.float32 and wq.scale.shape == (W.shape[0], 1)
            assert torch.equal(wq.scale, torch.from_numpy(g[f"{tag}_mse{mse}_scale"]))
            assert torch.equal(wq.quantize(W), torch.from_numpy(g[f"{tag}_mse{mse}_fq"]))
            assert bool(wq.ready()) and bool(wq.enabled())
    q = torch.from_numpy(g["pack_q"])
    assert torch.equal(qu.pack_i4(q), torch.from_numpy(g["pack_p"]))
    assert torch.equal(qu.unpack_i4(torch.from_numpy(g["pack_p"])), torch.from_numpy(g["pack_u"]))


def test_compat_shim_resolves_fake_quant(built):
    import importlib
    import sys

    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        fq = importlib.import_module("fake_quant")
        from fake_quant import hadamard_utils, quant_utils, utils  # noqa: F401
        from fake_quant.observer import build_observer  # noqa: F401
        from fake_quant.quantizer import build_quantizer  # noqa: F401
        from fake_quant.bit_type import BIT_TYPE_DICT  # noqa: F401

        assert quant_utils is built.quant_utils and fq is not None
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))


def test_fuse_siblings_gating(built):
    """fuse_siblings only groups wrappers for which sharing is exact (host logic; the packed
    weights are stand-in byte buffers, nothing is launched)."""
    import torch
    from mquant_b200 import quant_utils as qu

    def wrapper(n_out, scale=0.05, scale_v=0.1, k=256, had=False):
        w = qu.ActQuantWrapper(torch.nn.Linear(k, n_out, bias=False))
        w.quantizer.configure(bits=8, sym=True, static=True, observer_type="minmax", msq=scale_v is not None)
        w.quantizer.quantizer.set_scales(torch.tensor(scale, dtype=torch.float16),
                                         None if scale_v is None else torch.tensor(scale_v, dtype=torch.float16))
        w._packed = {"w4": torch.zeros(n_out * k // 2, dtype=torch.uint8), "s_w": torch.ones(n_out), "w0": None,
                     "bias": None, "N": n_out, "K": k, "Kp": k}
        w.real_quant = True
        w.online_full_had = had
        return w

    a, b, c = wrapper(512), wrapper(256), wrapper(104)
    g = qu.fuse_siblings([a, b, c])          # ragged last member is fine
    assert g is not None and g.N == 872 and g.offsets == [0, 512, 768] and a._group is g and c._group is g
    assert g.w4.numel() == 872 * 256 // 2 and g.s_w.numel() == 872 and g.bias is None
    assert qu.fuse_siblings([a, wrapper(256)]) is None                      # already grouped
    assert qu.fuse_siblings([wrapper(104), wrapper(256)]) is None           # ragged member not last
    assert qu.fuse_siblings([wrapper(256), wrapper(256, scale=0.06)]) is None   # different text scale
    assert qu.fuse_siblings([wrapper(256), wrapper(256, scale_v=0.2)]) is None  # different visual scale
    assert qu.fuse_siblings([wrapper(256), wrapper(256, scale_v=None)]) is None # MSQ vs not
    assert qu.fuse_siblings([wrapper(256), wrapper(256, had=True)]) is None     # online rotation
    assert qu.fuse_siblings([wrapper(256), wrapper(256, k=512)]) is None        # different K
    assert qu.fuse_siblings([wrapper(256)]) is None
    unfrozen = wrapper(256)
    unfrozen.real_quant = False
    assert qu.fuse_siblings([wrapper(256), unfrozen]) is None


def test_update_all_quantization_params_equals_per_wrapper_update():
    """One host round trip for all wrappers == the per-wrapper update (which mirrors the reference's
    MinmaxObserver.get_quantization_params): statistics installed as the kernels leave them, text-only,
    visual-only, both, bf16 statistics, an uncalibrated wrapper."""
    import copy

    import torch

    from mquant_b200 import quant_utils as qu

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.b, self.c = (torch.nn.Linear(16, 8) for _ in range(3))
            self.d, self.e = torch.nn.Linear(16, 8), torch.nn.Linear(16, 8)

    net = Net()
    qu.add_actquant(net)
    cfg = {"a": (True, (-1.5, 2.0), (-8.0, 6.0), [10, 4], torch.float16),
           "b": (True, (-0.3, 0.0), None, [7, 0], torch.float16),
           "c": (True, None, (-2.0, 9.0), [0, 5], torch.float16),
           "d": (False, (-100.0, 31.0), None, None, torch.bfloat16),
           "e": (False, None, None, None, torch.float16)}
    for name, (msq, text, vis, seen, dt) in cfg.items():
        q = getattr(net, name).quantizer
        q.configure(bits=8, sym=True, static=True, observer_type="minmax", msq=msq)
        for o, st in ((q.observer, text), (q.observer_visual, vis)):
            if st is not None:
                o.state = torch.tensor(st, dtype=torch.float32)
                o._dtype, o._channels = dt, 16
        if seen is not None:
            q._seen = torch.tensor(seen, dtype=torch.int32)
    ref = copy.deepcopy(net)
    for name in cfg:
        q = getattr(ref, name).quantizer
        if qu._observed(q.observer) or (q.msq and qu._observed(q.observer_visual)):
            q.update_quantization_params()
    n = qu.update_all_quantization_params(net)
    assert n == 4
    for name in cfg:
        a, b = getattr(net, name).quantizer.quantizer, getattr(ref, name).quantizer.quantizer
        for attr in ("scale", "scale_visual"):
            x, y = getattr(a, attr), getattr(b, attr)
            assert (x is None) == (y is None), (name, attr)
            if x is not None:
                assert x.dtype == y.dtype and torch.equal(x.reshape(()), y.reshape(())), (name, attr, x, y)
        if a.scale is not None:
            assert torch.equal(a.device_scales(torch.device("cpu")), b.device_scales(torch.device("cpu"))), name


def test_aifs_reorder_keeps_causal_attention_invariant():
    """visual tokens first (model/modeling_qwen_opt.py:717-727, commented sketch).
    A toy decoder layer (rotary attention + position-wise MLP) on reordered tokens with
    aifs.causal_mask and permuted rotary phases gives the reordered output of the original layer;
    the plain lower-triangular mask does not.  After the reorder the MSQ mask is two segments."""
    import torch
    from mquant_b200 import aifs

    torch.manual_seed(0)
    B, S, D, H = 2, 24, 32, 4
    x = torch.randn(B, S, D, dtype=torch.float64)
    wq, wk, wv, wo, w1 = (torch.randn(D, D, dtype=torch.float64) / D ** 0.5 for _ in range(5))
    ang = torch.arange(S, dtype=torch.float64)[:, None] * (1.0 / 10.0 ** (torch.arange(D // H // 2, dtype=torch.float64) / 4))
    cos, sin = ang.cos().repeat(1, 2), ang.sin().repeat(1, 2)  # [S, D/H]

    def rot(t, c, s):  # t [B,H,S,d]; c, s [B,S,d]
        t1, t2 = t[..., : t.shape[-1] // 2], t[..., t.shape[-1] // 2:]
        return t * c[:, None] + torch.cat((-t2, t1), -1) * s[:, None]

    def layer(h, c, s, mask):  # mask [B,S,S] bool
        def heads(t):
            return t.reshape(B, S, H, D // H).transpose(1, 2)
        q, k, v = rot(heads(h @ wq), c, s), rot(heads(h @ wk), c, s), heads(h @ wv)
        att = (q @ k.transpose(-1, -2)) / (D // H) ** 0.5
        att = att.masked_fill(~mask[:, None], float("-inf")).softmax(-1)
        o = (att @ v).transpose(1, 2).reshape(B, S, D) @ wo
        return h + o + torch.nn.functional.silu((h + o) @ w1)

    cb, sb = cos.expand(B, S, -1), sin.expand(B, S, -1)
    tril = torch.ones(S, S, dtype=torch.bool).tril().expand(B, S, S)
    ref = layer(x, cb, sb, tril)

    img_pos = [(0, 3, 12), (1, 0, 7), (1, 15, 20)]  # (sequence, image-start marker, image-end marker)
    p = aifs.plan(img_pos, S, B)
    assert p["n_visual"].tolist() == [8, 10]
    assert p["perm"][0, :8].tolist() == list(range(4, 12)) and p["perm"][0, 8:12].tolist() == [0, 1, 2, 3]
    xr = aifs.reorder(x, p)
    assert torch.equal(aifs.restore(xr, p), x)
    out = layer(xr, aifs.reorder(cb, p), aifs.reorder(sb, p), aifs.causal_mask(p))
    assert torch.allclose(aifs.restore(out, p), ref, rtol=1e-12, atol=1e-12)
    wrong = layer(xr, aifs.reorder(cb, p), aifs.reorder(sb, p), tril)
    assert not torch.allclose(aifs.restore(wrong, p), ref, rtol=1e-6, atol=1e-6)
    m = aifs.modality_mask(p, S).reshape(B, S)
    assert m[0].tolist() == [1] * 8 + [0] * 16 and m[1].tolist() == [1] * 10 + [0] * 14
    with pytest.raises(ValueError):
        aifs.plan([(0, 5, 40)], S, B)
