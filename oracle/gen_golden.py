"""Generate tests/golden/*.npz by EXECUTING the unmodified reference on CPU.

Generator-side only: run in the build container (``python oracle/gen_golden.py``) where
/root/reference is mounted.  The reference ships no tests or golden vectors (SURVEY.md
section 4), so these frozen input/output pairs are what pins ``oracle/mq_oracle.py``.
Each case names the reference entry point it drives.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    from oracle._ref_import import import_reference

    qu, hu = import_reference()
    import torch

    torch.manual_seed(42)  # drivers default --seed 42 (exam/quant_qwenvl.py:407)
    torch.set_num_threads(1)
    os.makedirs(OUT, exist_ok=True)

    def rnd(*shape, scale=3.0, dtype=torch.float16):
        x = torch.randn(*shape) * scale
        # outlier channels to exercise the clamp, as in SURVEY.md 8(d)
        if shape[-1] >= 64:
            x[..., 7] *= 40
            x[..., shape[-1] // 2 + 3] *= 25
        return x.to(dtype)

    # ---------------------------------------------------------------- 1. static act quant
    # ActQuantizer.configure/forward (quant_utils.py:116-179) -> MinmaxObserver ->
    # UniformQuantizer; three calibration batches, the last with last_calibrate.
    case = {}
    for tag, dtype, mode in (("f16", torch.float16, "layer_wise"),
                             ("f32", torch.float32, "layer_wise"),
                             ("f16cw", torch.float16, "channel_wise")):
        q = qu.ActQuantizer()
        q.configure(bits=8, sym=True, static=True, observer_type="minmax", calibration_mode=mode)
        calib = [rnd(2, 24, 192, dtype=dtype) for _ in range(3)]
        q.calibrate = True
        for i, c in enumerate(calib):
            if i == len(calib) - 1:
                q.last_calibrate = True
            out = q(c)
            assert out is c
        q.calibrate = False
        q.quant = True
        x = rnd(2, 40, 192, scale=4.0, dtype=dtype)
        x[0, 0, :8] = torch.tensor([0.0, -0.0, 65504, -65504, 1e-4, -1e-4, 0.5, -0.5]).to(dtype)
        y = q(x)
        codes = q.quantizer.quant(x.float())
        case[f"{tag}_calib"] = torch.stack(calib).numpy()
        case[f"{tag}_x"] = x.numpy()
        case[f"{tag}_scale"] = q.quantizer.scale.numpy()
        case[f"{tag}_zp"] = q.quantizer.zero_point.numpy()
        case[f"{tag}_min"] = np.asarray(q.observer.min_val.numpy())
        case[f"{tag}_max"] = np.asarray(q.observer.max_val.numpy())
        case[f"{tag}_codes"] = codes.numpy()
        case[f"{tag}_y"] = y.numpy()
    # half-way cases for round-half-even with an exactly representable scale
    q = qu.ActQuantizer()
    q.configure(bits=8, sym=True, static=True, observer_type="minmax")
    q.quantizer.scale = torch.tensor(0.25, dtype=torch.float16)
    q.quantizer.zero_point = torch.tensor(0, dtype=torch.int64)
    q.quant = True
    x = (torch.arange(-1100, 1100, dtype=torch.float32) * 0.125 / 4).to(torch.float16).reshape(1, -1, 8)
    case["tie_x"] = x.numpy()
    case["tie_scale"] = q.quantizer.scale.numpy()
    case["tie_codes"] = q.quantizer.quant(x.float()).numpy()
    case["tie_y"] = q(x).numpy()
    np.savez_compressed(os.path.join(OUT, "act_quant.npz"), **case)

    # ---------------------------------------------------------------- 2. minmax -> scale sweep
    # MinmaxObserver.get_quantization_params (minmax.py:30-52) on hand-set statistics:
    # every non-negative finite fp16 value as max (and its negation as min), plus fp32 samples.
    from fake_quant.observer import build_observer
    from fake_quant.bit_type import BIT_TYPE_DICT

    ob = build_observer("minmax", "activation", BIT_TYPE_DICT["int8"], "channel_wise")
    allpos = torch.arange(0, 0x7C00, dtype=torch.int32).to(torch.int16).view(torch.float16)
    perm = torch.randperm(allpos.numel())
    ob.max_val = allpos.clone()
    ob.min_val = -allpos[perm]
    s16, _ = ob.get_quantization_params()
    mx32 = torch.rand(4096) * torch.logspace(-9, 4, 4096)
    mn32 = -torch.rand(4096) * torch.logspace(-9, 4, 4096).flip(0)
    ob.max_val, ob.min_val = mx32.clone(), mn32.clone()
    s32, _ = ob.get_quantization_params()
    np.savez_compressed(os.path.join(OUT, "minmax_params.npz"),
                        max16=allpos.numpy(), min16=(-allpos[perm]).numpy(), scale16=s16.numpy(),
                        max32=mx32.numpy(), min32=mn32.numpy(), scale32=s32.numpy())

    # ---------------------------------------------------------------- 3. percentile observer
    # PercentileObserver.update/get_quantization_params (percentile.py:26-79), 3 updates.
    ob = build_observer("percentile", "activation", BIT_TYPE_DICT["int8"], "layer_wise")
    pc = [rnd(1, 640, 192) for _ in range(3)]
    stats = []
    for c in pc:
        ob.update(c)
        stats.append([float(ob.min_val), float(ob.max_val)])
    ps, _ = ob.get_quantization_params()
    np.savez_compressed(os.path.join(OUT, "percentile.npz"), calib=torch.stack(pc).numpy(),
                        stats=np.asarray(stats, dtype=np.float32), scale=ps.numpy())

    # ---------------------------------------------------------------- 4. weight quantizer
    # WeightQuantizer.configure/find_params/quantize (quant_utils.py:415-524), RTN and clip search
    case = {}
    lin = torch.nn.Linear(256, 96, bias=False)
    W = lin.weight.data.clone()
    W[3] *= 0  # all-zero row -> clamp(min=1e-5)
    W[5, 17] = 0.9
    for tag, dtype in (("f16", torch.float16), ("f32", torch.float32)):
        for mse in (False, True):
            wq = qu.WeightQuantizer()
            wq.configure(4, perchannel=True, sym=True, mse=mse)
            Wd = W.to(dtype)
            wq.find_params(Wd)
            fq = wq.quantize(Wd)
            codes = torch.round(fq.float() / wq.scale)
            case[f"{tag}_mse{int(mse)}_scale"] = wq.scale.numpy()
            case[f"{tag}_mse{int(mse)}_fq"] = fq.numpy()
            case[f"{tag}_mse{int(mse)}_codes"] = codes.numpy().astype(np.int8)
        case[f"{tag}_W"] = W.to(dtype).numpy()
    # ---- pack_i4 / unpack_i4 (quant_utils.py:61-94)
    qi = torch.randint(-8, 8, (32, 64), dtype=torch.int32)
    packed = qu.pack_i4(qi)
    case["pack_q"] = qi.numpy()
    case["pack_p"] = packed.numpy()
    case["pack_u"] = qu.unpack_i4(packed).numpy()
    np.savez_compressed(os.path.join(OUT, "weight_quant.npz"), **case)

    # ---------------------------------------------------------------- 5. hadamard
    case = {}
    sizes = [64, 256, 4096, 8192, 11008, 13824, 14336, 19968, 3584, 5120, 1664, 344, 96, 1280]
    ks = []
    for n in sizes:
        hk, k = hu.get_hadK(n)
        ks.append(k)
    case["hadk_sizes"] = np.asarray(sizes)
    case["hadk_K"] = np.asarray(ks)
    pads_in = [11008, 13696, 18944, 14336, 4096, 1000, 3584, 5120, 27392, 8192, 29568]
    case["pad_in"] = np.asarray(pads_in)
    case["pad_out"] = np.asarray([hu.auto_pad_size(n) for n in pads_in])
    # pure-torch matmul_hadU in fp64 (hadamard_utils.py:79-100): the transform's anchor
    for n in (256, 344, 11008 // 16):  # 688 = 172*4
        x = torch.randn(3, n, dtype=torch.float64)
        case[f"hadU_x_{n}"] = x.numpy()
        case[f"hadU_y_{n}"] = hu.matmul_hadU(x).numpy()
    # matmul_hadU_cuda with the FHT stub, fp16 in/out (hadamard_utils.py:115-128)
    for n in (512, 344, 11008):
        hk, k = hu.get_hadK(n)
        x = rnd(4, n)
        case[f"hadcuda_x_{n}"] = x.numpy()
        case[f"hadcuda_y_{n}"] = hu.matmul_hadU_cuda(x, hk, k).numpy()
        case[f"hadcuda_y32_{n}"] = hu.matmul_hadU_cuda(x.float(), hk, k).to(torch.float16).numpy()
    # random_hadamard_matrix determinism (hadamard_utils.py:107-112)
    torch.manual_seed(0)
    case["rand_had_64"] = hu.random_hadamard_matrix(64, "cpu").numpy()
    np.savez_compressed(os.path.join(OUT, "hadamard.npz"), **case)

    # ---------------------------------------------------------------- 6. wrapper forward
    # ActQuantWrapper.forward (quant_utils.py:330-391): plain, online Hadamard (K=1 and K=172),
    # split.  W4 RTN weights, static int8 minmax activations calibrated on 2 batches.
    torch.manual_seed(7)
    case = {}
    for tag, K_in, N_out, bias, had, split in (
            ("plain", 256, 128, True, False, False),
            ("had256", 256, 64, False, True, False),
            ("had344", 344, 64, False, True, False),
            ("split", 344, 64, False, True, True)):
        lin = torch.nn.Linear(K_in, N_out, bias=bias).half()
        if had:
            hu_apply = lin.weight.data.float()
            hk, k = hu.get_hadK(K_in)
            lin.weight.data = hu.matmul_hadU_cuda(hu_apply, hk, k).half()
        w = qu.ActQuantWrapper(lin)
        if had:
            w.online_full_had = True
            w.had_K, w.K = hu.get_hadK(K_in)
        if split:
            w.split = True
            w.split_weights()
        # weight RTN as qwenvl_llm_rtn does (gptq/qwenvl_gptq_plus.py:450-481)
        targets = [w.L2] if split else [w.module]
        wscales = []
        for t in targets:
            wq = qu.WeightQuantizer()
            wq.configure(4, perchannel=True, sym=True, mse=False)
            wq.find_params(t.weight.data)
            t.weight.data = wq.quantize(t.weight.data).to(torch.float16)
            wscales.append(wq.scale.numpy())
        w.quantizer.configure(bits=8, sym=True, static=True, observer_type="minmax")
        w.quantizer.calibrate = True
        for i in range(2):
            if i == 1:
                w.quantizer.last_calibrate = True
            w(rnd(1, 32, K_in))
        w.quantizer.calibrate = False
        w.quantizer.quant = True
        x = rnd(2, 16, K_in)
        y = w(x.clone())
        case[f"{tag}_x"] = x.numpy()
        case[f"{tag}_y"] = y.detach().numpy()
        case[f"{tag}_act_scale"] = w.quantizer.quantizer.scale.numpy()
        if split:
            case[f"{tag}_W"] = torch.cat([w.L1.weight.data.half(), w.L2.weight.data.half()], 1).numpy()
        else:
            case[f"{tag}_W"] = w.module.weight.data.numpy()
        case[f"{tag}_wscale"] = wscales[0]
        if bias:
            case[f"{tag}_b"] = lin.bias.data.numpy()
    np.savez_compressed(os.path.join(OUT, "wrapper.npz"), **case)

    # ---------------------------------------------------------------- 7. dynamic per-token
    q = qu.ActQuantizer()
    q.configure(bits=8, sym=True, static=False)
    x = rnd(2, 12, 192)
    x[0, 3] = 0
    q.find_params(x)
    y = q(x)
    np.savez_compressed(os.path.join(OUT, "dynamic.npz"), x=x.numpy(),
                        scale=q.scale[..., 0].reshape(-1).numpy(), y=y.numpy())
    # ---------------------------------------------------------------- 8. bf16 activations
    # (appended last so that the random stream of the cases above is unchanged)
    # InternVL2 / Qwen2-VL run in bf16: statistics, scale and output stay bf16 in the reference;
    # stored as uint16 bit patterns (numpy has no bf16)
    q = qu.ActQuantizer()
    q.configure(bits=8, sym=True, static=True, observer_type="minmax", calibration_mode="layer_wise")
    calib = [rnd(2, 24, 192, dtype=torch.bfloat16) for _ in range(3)]
    q.calibrate = True
    for i, c in enumerate(calib):
        if i == len(calib) - 1:
            q.last_calibrate = True
        q(c)
    q.calibrate = False
    q.quant = True
    x = rnd(2, 40, 192, scale=4.0, dtype=torch.bfloat16)
    x[0, 0, :8] = torch.tensor([0.0, -0.0, 60000, -60000, 1e-3, -1e-3, 0.5, -0.5]).to(torch.bfloat16)
    y = q(x)

    def bits(t):
        return t.contiguous().view(torch.int16).numpy().view(np.uint16)

    np.savez_compressed(os.path.join(OUT, "act_quant_bf16.npz"), calib=bits(torch.stack(calib)), x=bits(x),
                        scale=bits(q.quantizer.scale.reshape(1)), codes=q.quantizer.quant(x.float()).numpy(),
                        y=bits(y))
    # ---------------------------------------------------------------- 9. wrapper forward, GPU-sized
    # (appended after the cases above so that their random stream is unchanged)  The had344 / split
    # cases of section 6 have n / had_k = 2, below what the CUDA kernels take (>= 64); the same
    # wrapper configurations at K = 768 = 12 x 64 run on the GPU: online Hadamard K = 12, and the
    # RMS split on top of it.
    torch.manual_seed(9)
    case = {}
    for tag, K_in, N_out, bias, had, split in (("had768", 768, 64, False, True, False),
                                               ("split768", 768, 64, True, True, True)):
        lin = torch.nn.Linear(K_in, N_out, bias=bias).half()
        hk, k = hu.get_hadK(K_in)
        lin.weight.data = hu.matmul_hadU_cuda(lin.weight.data.float(), hk, k).half()
        w = qu.ActQuantWrapper(lin)
        w.online_full_had = True
        w.had_K, w.K = hu.get_hadK(K_in)
        if split:
            w.split = True
            w.split_weights()
        t = w.L2 if split else w.module
        wq = qu.WeightQuantizer()
        wq.configure(4, perchannel=True, sym=True, mse=False)
        wq.find_params(t.weight.data)
        t.weight.data = wq.quantize(t.weight.data).to(torch.float16)
        w.quantizer.configure(bits=8, sym=True, static=True, observer_type="minmax")
        w.quantizer.calibrate = True
        for i in range(2):
            if i == 1:
                w.quantizer.last_calibrate = True
            w(rnd(1, 32, K_in))
        w.quantizer.calibrate = False
        w.quantizer.quant = True
        x = rnd(2, 16, K_in)
        y = w(x.clone())
        case[f"{tag}_x"] = x.numpy()
        case[f"{tag}_y"] = y.detach().numpy()
        case[f"{tag}_act_scale"] = w.quantizer.quantizer.scale.numpy()
        if split:
            case[f"{tag}_W"] = torch.cat([w.L1.weight.data.half(), w.L2.weight.data.half()], 1).numpy()
        else:
            case[f"{tag}_W"] = w.module.weight.data.numpy()
        case[f"{tag}_wscale"] = wq.scale.numpy()
        if bias:
            case[f"{tag}_b"] = lin.bias.data.numpy()
    np.savez_compressed(os.path.join(OUT, "wrapper768.npz"), **case)

    # ---------------------------------------------------------------- 10. BASELINE config 1: QWenBlock
    # One reference decoder layer (model/modeling_qwen_opt.py:390-420) at hidden 4096 / ffn 11008 /
    # seq 1024 with its 7 ActQuantWrappers (RTN W4, static INT8 minmax, online Hadamard K = 172 on
    # mlp.c_proj), through oracle/ref_block.py.  The weights are regenerated from per-layer seeds by
    # the test; the fixture keeps the calibrated scales, the block output and every wrapper's
    # input / output on sampled token rows.
    from oracle import ref_block as RB
    from model import modeling_qwen_opt as mq

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    block, wrappers, wscales = RB.build_reference_block(qu, hu, mq)
    rows = np.unique(np.concatenate([np.arange(0, 4), np.arange(1020, 1024),
                                     np.random.default_rng(0).integers(4, 1020, 8)]))
    ridx = torch.from_numpy(rows)
    case = {"rows": rows, "seed0": np.asarray(RB.SEED0), "calib_seeds": np.asarray([11, 12]),
            "x_seed": np.asarray(13), "tokens": np.asarray(1024)}
    with torch.no_grad():
        RB.calibrate(block, wrappers, mq, [RB.hidden_states(11, 1024), RB.hidden_states(12, 1024)])
        grabbed = {}
        hooks = [w.register_forward_hook(lambda m, i, o, n=n: grabbed.__setitem__(n, (i[0], o)))
                 for n, w in wrappers.items()]
        y = RB.run_block(block, mq, RB.hidden_states(13, 1024))
        for h in hooks:
            h.remove()
    for n, w in wrappers.items():
        key = n.replace(".", "_")
        xi, yo = grabbed[n]
        case[f"{key}_x"] = xi.reshape(-1, xi.shape[-1])[ridx].numpy()
        case[f"{key}_y"] = yo.reshape(-1, yo.shape[-1])[ridx].numpy()
        case[f"{key}_act_scale"] = w.quantizer.quantizer.scale.numpy()
        case[f"{key}_wscale_sum"] = np.asarray(wscales[n].double().sum().item())
    case["block_y"] = y[0, ridx].numpy()
    case["block_y_rownorm"] = y[0].float().norm(dim=-1).numpy()
    np.savez_compressed(os.path.join(OUT, "qwen_block.npz"), **case)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def gen_rmsn():
    """RMSN.forward (fake_quant/module_util.py:42-61) followed by a static ActQuantizer
    (quant_utils.py:97-179), as in front of q/k/v and w1/w2 once the LayerNorm weights are fused
    into the Linears (rotation_utils.py fuse_layer_norms).  Separate entry point so that the
    other fixtures need not be regenerated: ``python oracle/gen_golden.py rmsn``."""
    from oracle._ref_import import import_reference

    qu, hu = import_reference()
    import torch
    from fake_quant.module_util import RMSN

    torch.manual_seed(4242)
    torch.set_num_threads(1)
    case = {}
    for tag, M, C in (("h4096", 20, 4096), ("h3584", 9, 3584), ("h1000", 7, 1000)):
        x = (torch.randn(M, C) * 2.5)
        x[:, 5] *= 30
        x[M // 2] *= 0.01  # one quiet token: eps matters
        x = x.to(torch.float16)
        ln = RMSN(C, eps=1e-6 if C == 3584 else 1e-5)
        y = ln(x)
        q = qu.ActQuantizer()
        q.configure(bits=8, sym=True, static=True, observer_type="minmax", calibration_mode="layer_wise")
        q.calibrate, q.quant = True, False
        q(y[: M // 2])
        q.last_calibrate = True
        q(y[M // 2:])
        q.calibrate, q.quant, q.last_calibrate = False, True, False
        fq = q(y)
        scale = q.quantizer.scale.float()
        case[f"{tag}_x"] = x.numpy()
        case[f"{tag}_eps"] = np.float32(ln.eps)
        case[f"{tag}_y"] = y.numpy()
        case[f"{tag}_scale"] = scale.numpy()
        case[f"{tag}_codes"] = torch.round(fq.float() / scale).to(torch.int8).numpy()  # exact: fq = code * scale in fp16
    np.savez_compressed(os.path.join(OUT, "rmsn.npz"), **case)
    print("rmsn.npz", os.path.getsize(os.path.join(OUT, "rmsn.npz")))


def gen_gptq():
    """GPTQ (fake_quant/gptq/gptq_utils.py:171-310) on one small Linear, on the CPU: Hessian after
    three add_batch calls, WeightQuantizer scales and the final fake-quantised weight.
    ``python oracle/gen_golden.py gptq`` (unfoldNd, used by GPTQConv only, is absent: stubbed)."""
    import types

    from oracle._ref_import import import_reference

    qu, hu = import_reference()
    import torch

    if "unfoldNd" not in sys.modules:
        m = types.ModuleType("unfoldNd")
        m.UnfoldNd = object
        sys.modules["unfoldNd"] = m
    from fake_quant.gptq.gptq_utils import GPTQ

    torch.manual_seed(7)
    torch.set_num_threads(1)
    rows, cols = 96, 256
    layer = torch.nn.Linear(cols, rows, bias=False)
    layer.weight.data = (torch.randn(rows, cols) * 0.05).half().float()
    W0 = layer.weight.data.clone()
    g = GPTQ(layer)
    g.quantizer = qu.WeightQuantizer()
    g.quantizer.configure(4, perchannel=True, sym=True, mse=False)
    basis = torch.randn(cols, cols) / cols ** 0.5  # correlated inputs: the Hessian is far from diagonal
    xs = []
    for b in range(3):
        x = ((torch.randn(2, 64, cols) @ basis) * (1 + 3 * torch.rand(cols))).half()
        xs.append(x)
        g.add_batch(x, None)
    H = g.H.clone()
    torch.cuda.synchronize = lambda *a, **k: None  # (fasterquant calls it unconditionally, :285; no device here)
    g.fasterquant(percdamp=0.01, groupsize=-1)
    case = {"W0": W0.numpy(), "H": H.numpy(), "nsamples": np.int64(g.nsamples), "scale": g.quantizer.scale.reshape(-1).numpy(),
            "Wq": layer.weight.data.numpy()}
    for b, x in enumerate(xs):
        case[f"x{b}"] = x.numpy()
    np.savez_compressed(os.path.join(OUT, "gptq.npz"), **case)
    print("gptq.npz", os.path.getsize(os.path.join(OUT, "gptq.npz")))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "rmsn":
        gen_rmsn()
    elif len(sys.argv) > 1 and sys.argv[1] == "gptq":
        gen_gptq()
    else:
        main()
