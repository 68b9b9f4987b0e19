"""Pin oracle/mq_oracle.py against outputs of the reference's own Python
(fixtures frozen by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import mq_oracle as O


def bits16(a):
    return np.asarray(a, dtype=np.float16).view(np.uint16)


@pytest.mark.parametrize("tag", ["f16", "f32", "f16cw"])
def test_static_act_quant_matches_reference(golden, tag):
    g = golden("act_quant")
    mode = "channel_wise" if tag.endswith("cw") else "layer_wise"
    ob = O.MinmaxObserver(8, mode)
    for c in g[f"{tag}_calib"]:
        ob.update(c)
    assert np.array_equal(np.asarray(ob.min_val), g[f"{tag}_min"])
    assert np.array_equal(np.asarray(ob.max_val), g[f"{tag}_max"])
    scale, zp = ob.get_quantization_params()
    assert scale.dtype == g[f"{tag}_scale"].dtype
    assert np.array_equal(scale, g[f"{tag}_scale"])
    assert np.array_equal(zp, g[f"{tag}_zp"])
    x = g[f"{tag}_x"]
    codes = O.act_quant_codes(x, scale)
    assert np.array_equal(codes, g[f"{tag}_codes"])
    assert codes.min() == -128 and codes.max() == 127  # the clamp is exercised
    y = O.act_fake_quant(x, scale)
    assert y.dtype == g[f"{tag}_y"].dtype
    assert np.array_equal(y.view(np.uint16 if y.dtype == np.float16 else np.uint32),
                          g[f"{tag}_y"].view(np.uint16 if y.dtype == np.float16 else np.uint32))


def test_round_half_even_ties(golden):
    g = golden("act_quant")
    codes = O.act_quant_codes(g["tie_x"], g["tie_scale"])
    assert np.array_equal(codes, g["tie_codes"])
    assert np.array_equal(bits16(O.act_fake_quant(g["tie_x"], g["tie_scale"])), bits16(g["tie_y"]))
    # at least some exact .5 cases exist and went to even
    r = g["tie_x"].astype(np.float32) / np.float32(g["tie_scale"])
    ties = np.abs(r - np.floor(r) - 0.5) == 0
    assert ties.sum() > 100
    inr = ties & (np.abs(r) < 120)
    assert np.all(codes[inr] % 2 == 0)


def test_minmax_params_all_fp16(golden):
    g = golden("minmax_params")
    s16, _ = O.minmax_params(g["min16"], g["max16"])
    assert s16.dtype == np.float16
    assert np.array_equal(bits16(s16), bits16(g["scale16"]))
    s32, _ = O.minmax_params(g["min32"], g["max32"])
    assert s32.dtype == np.float32
    assert np.array_equal(s32.view(np.uint32), g["scale32"].view(np.uint32))


def test_percentile_observer(golden):
    g = golden("percentile")
    ob = O.PercentileObserver()
    for c, st in zip(g["calib"], g["stats"]):
        ob.update(c)
        assert np.float32(ob.min_val) == st[0]
        assert np.float32(ob.max_val) == st[1]
    s, _ = ob.get_quantization_params()
    assert np.array_equal(np.float32(s).view(np.uint32), g["scale"].view(np.uint32))


@pytest.mark.parametrize("tag", ["f16", "f32"])
@pytest.mark.parametrize("mse", [0, 1])
def test_weight_quantizer(golden, tag, mse):
    g = golden("weight_quant")
    W = g[f"{tag}_W"]
    scale = O.weight_find_params(W, 4, mse=bool(mse))
    ref = g[f"{tag}_mse{mse}_scale"]
    if mse:
        # clip search compares fp32 error sums whose reduction order differs between torch and
        # numpy; near-ties may pick a neighbouring shrink step.  RTN (the drivers' --llm_w_rtn
        # path) is exact; the search must agree on >= 95 % of rows and be a valid grid point.
        same = (scale == ref).mean()
        assert same >= 0.95, same
        scale = ref
    else:
        assert np.array_equal(scale.view(np.uint32), ref.view(np.uint32))
    codes = O.weight_quant_codes(W, scale)
    assert np.array_equal(codes, g[f"{tag}_mse{mse}_codes"])
    fq = O.weight_fake_quant(W, scale)
    assert np.array_equal(fq, g[f"{tag}_mse{mse}_fq"])


def test_pack_unpack(golden):
    g = golden("weight_quant")
    assert np.array_equal(O.pack_i4(g["pack_q"]), g["pack_p"])
    assert np.array_equal(O.unpack_i4(g["pack_p"]), g["pack_u"])
    assert np.array_equal(O.unpack_i4(O.pack_i4(g["pack_q"])), g["pack_q"])


def test_hadk_table_and_pad(golden):
    g = golden("hadamard")
    for n, k in zip(g["hadk_sizes"], g["hadk_K"]):
        h, kk = O.get_hadK(int(n))
        assert kk == k
        if k > 1:
            assert np.array_equal(h @ h.T, k * np.eye(k, dtype=np.float32))
    for n, p in zip(g["pad_in"], g["pad_out"]):
        assert O.auto_pad_size(int(n)) == p


def test_matmul_hadU_fp64(golden):
    g = golden("hadamard")
    for n in (256, 344, 688):
        y = O.matmul_hadU(g[f"hadU_x_{n}"])
        np.testing.assert_allclose(y, g[f"hadU_y_{n}"], rtol=0, atol=1e-12)
    # involution for the symmetric K=1 case (SURVEY 8(c) self-check)
    x = g["hadU_x_256"]
    np.testing.assert_allclose(O.matmul_hadU(O.matmul_hadU(x)), x, atol=1e-12)
    # random_hadamard_matrix = matmul_hadU(diag(+-1)) (hadamard_utils.py:107-112)
    q = g["rand_had_64"]
    sgn = np.sign(q[:, 0])
    np.testing.assert_allclose(O.matmul_hadU(np.diag(sgn)), q, atol=1e-12)


def test_matmul_hadU_cuda_rounding_points(golden):
    """The FHT stub used at generation time sums in GEMM order, the oracle in butterfly order:
    equal up to one fp16 ulp on a tiny fraction of entries."""
    g = golden("hadamard")
    for n in (512, 344, 11008):
        x = g[f"hadcuda_x_{n}"]
        for key, fp32_had in ((f"hadcuda_y_{n}", False), (f"hadcuda_y32_{n}", True)):
            y = O.matmul_hadU_cuda(x, fp32_had=fp32_had)
            ref = g[key]
            d = np.abs(bits16(y).astype(np.int32) - bits16(ref).astype(np.int32))
            # ulp distance is meaningless for entries that cancel to ~0: bound those absolutely
            a = np.abs(y.astype(np.float32) - ref.astype(np.float32))
            rms = np.sqrt((ref.astype(np.float32) ** 2).mean())
            assert np.all((d <= 1) | (a <= 1e-4 * rms)), (n, key, d.max(), a.max())
            assert (d > 0).mean() < 5e-3, (n, key, (d > 0).mean())


WRAPPER_CASES = [("wrapper", "plain"), ("wrapper", "had256"), ("wrapper", "had344"), ("wrapper", "split"),
                 ("wrapper768", "had768"), ("wrapper768", "split768")]


@pytest.mark.parametrize("gname,tag", WRAPPER_CASES)
def test_wrapper_forward(golden, gname, tag):
    g = golden(gname)
    x = g[f"{tag}_x"]
    b = g[f"{tag}_b"] if f"{tag}_b" in g.files else None
    y = O.qlinear_fake(x, g[f"{tag}_W"], b, g[f"{tag}_act_scale"], online_had=tag != "plain",
                       split=tag.startswith("split"))
    ref = g[f"{tag}_y"].astype(np.float32)
    err = np.abs(y.astype(np.float32) - ref).max()
    cos = (y.astype(np.float64) * ref).sum() / np.sqrt((y.astype(np.float64) ** 2).sum() * (ref.astype(np.float64) ** 2).sum())
    assert cos > 0.99999, cos
    assert err <= 2.0 ** -9 * np.abs(ref).max() + 1e-3, err


@pytest.mark.parametrize("gname,tag", WRAPPER_CASES)
def test_int_path_matches_fake_quant_within_tolerance(golden, gname, tag):
    """the real-integer layer differs from the reference's fake-quant fp16 layer
    only by fp16 rounding of dequantised operands: cos >= 0.9999, max-abs <= 2^-9 max|y|."""
    g = golden(gname)
    x = g[f"{tag}_x"].reshape(-1, g[f"{tag}_x"].shape[-1])
    if tag != "plain":
        x = O.matmul_hadU_cuda(x)
    W = g[f"{tag}_W"]
    s_w = g[f"{tag}_wscale"].reshape(-1)
    s_a = np.float32(g[f"{tag}_act_scale"])
    b = g[f"{tag}_b"] if f"{tag}_b" in g.files else None
    if tag.startswith("split"):
        cw = np.rint(W[:, 1:].astype(np.float32) / s_w[:, None]).astype(np.int8)
        cw = np.concatenate([np.zeros((cw.shape[0], 1), np.int8), cw], 1)
        ca = O.act_quant_codes(x, s_a)
        ca[:, 0] = 0
        y = O.qlinear_int(ca, s_a, cw, s_w, b, x0=x[:, 0], w0=W[:, 0])
    else:
        cw = np.rint(W.astype(np.float32) / s_w[:, None]).astype(np.int8)
        assert np.array_equal((cw.astype(np.float32) * s_w[:, None]).astype(np.float16), W)
        y = O.qlinear_int(O.act_quant_codes(x, s_a), s_a, cw, s_w, b)
    ref = g[f"{tag}_y"].reshape(y.shape).astype(np.float64)
    cos = (y.astype(np.float64) * ref).sum() / np.sqrt((y.astype(np.float64) ** 2).sum() * (ref ** 2).sum())
    err = np.abs(y.astype(np.float64) - ref).max()
    assert cos >= 0.9999, cos
    assert err <= 2.0 ** -9 * np.abs(ref).max() * 4, (err, np.abs(ref).max())


def test_dynamic_per_token(golden):
    g = golden("dynamic")
    x = g["x"]
    s = O.dynamic_per_token_scales(x)
    assert np.array_equal(s.view(np.uint32), g["scale"].view(np.uint32))
    codes = O.dynamic_quant_codes(x, s)
    y = (s[:, None] * codes).astype(np.float16).reshape(x.shape)
    assert np.array_equal(bits16(y), bits16(g["y"]))


@pytest.mark.parametrize("tag", ["plain", "had256", "had344"])
def test_torch_port_matches_reference(golden, tag):
    """oracle/torch_port.py (what bench.py's CPU legs time) against the reference wrapper output."""
    import torch

    from oracle import torch_port as TP

    g = golden("wrapper")
    x = torch.from_numpy(g[f"{tag}_x"])
    W = torch.from_numpy(g[f"{tag}_W"])
    b = torch.from_numpy(g[f"{tag}_b"]) if f"{tag}_b" in g.files else None
    hk, K = (None, None)
    if tag != "plain":
        hk, K = TP.hadK_tensor(x.shape[-1])
    y = TP.qlinear_fake(x, W, b, torch.from_numpy(g[f"{tag}_act_scale"]), hk, K).numpy()
    ref = g[f"{tag}_y"].astype(np.float32)
    assert np.abs(y.astype(np.float32) - ref).max() <= 2.0 ** -9 * np.abs(ref).max() + 1e-3


def test_bf16_activations_match_reference(golden):
    """bf16 statistics, scale and fake-quant output (the reference run in bf16 on CPU)."""
    g = golden("act_quant_bf16")
    calib = O.bf16_to_f32(g["calib"])
    mx = max(float(calib.max()), 0.0)
    mn = min(float(calib.min()), 0.0)
    scale = O.minmax_params_bf16(O.f32_to_bf16_bits(np.float32(mn)), O.f32_to_bf16_bits(np.float32(mx)))
    assert np.array_equal(scale.reshape(-1), g["scale"].reshape(-1))
    codes, y = O.act_fake_quant_bf16(g["x"], g["scale"].reshape(()))
    assert np.array_equal(codes, g["codes"])
    assert np.array_equal(y, g["y"])


@pytest.mark.parametrize("tag", ["h4096", "h3584", "h1000"])
def test_rmsn_oracle_matches_reference(golden, tag):
    """oracle.rmsn vs RMSN.forward of the reference (fake_quant/module_util.py:42-61) run on the CPU;
    np.sum and torch.sum add the 4096 squares in different orders: <= 1 fp16 ulp on < 1e-3 of the
    elements.  The codes follow from the reference's own fp16 tensor exactly."""
    g = golden("rmsn")
    x, y = g[tag + "_x"], g[tag + "_y"]
    yo = O.rmsn(x, x.shape[1], float(g[tag + "_eps"]))
    d = np.abs(yo.view(np.uint16).astype(np.int32) - y.view(np.uint16).astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    assert np.array_equal(O.act_quant_codes(y, g[tag + "_scale"]).astype(np.int8), g[tag + "_codes"])


def _np_fasterquant(W0, H, scale, maxq=7, blocksize=128, percdamp=0.01):
    """GPTQ.fasterquant (gptq_utils.py:197-303) with numpy's fp32 LAPACK and oracle.gptq_block."""
    W = np.array(W0, np.float32)
    H = np.array(H, np.float32)
    dead = np.diag(H) == 0
    H[dead, dead] = 1
    W[:, dead] = 0
    H[np.diag_indices_from(H)] += np.float32(percdamp) * np.mean(np.diag(H), dtype=np.float32)
    L = np.linalg.cholesky(H)
    Hi = np.linalg.inv(L).T @ np.linalg.inv(L)  # cholesky_inverse
    Hinv = np.linalg.cholesky(Hi.astype(np.float32)).T.astype(np.float32)  # upper factor
    Q = np.zeros_like(W)
    for i1 in range(0, W.shape[1], blocksize):
        i2 = min(i1 + blocksize, W.shape[1])
        W1, Q1, Err1, _ = O.gptq_block(W[:, i1:i2], Hinv[i1:i2, i1:i2], scale, maxq)
        Q[:, i1:i2] = Q1
        W[:, i2:] -= Err1 @ Hinv[i1:i2, i2:]
    return Q


def test_gptq_oracle_matches_reference(golden):
    """oracle.gptq_hessian / gptq_block vs GPTQ.add_batch / fasterquant of the reference run on the CPU
    (tests/golden/gptq.npz).  The Hessian agrees to fp32 summation noise; the quantised weight is the
    reference's up to the entries a different Cholesky rounding can flip (different LAPACK calls)."""
    g = golden("gptq")
    H, n = np.zeros_like(g["H"], dtype=np.float64), 0
    for b in range(3):
        x = g[f"x{b}"]
        H, n = O.gptq_hessian(H, n, x, x.shape[0])
    assert n == int(g["nsamples"])
    assert np.abs(H - g["H"]).max() <= 2e-6 * np.abs(g["H"]).max()
    Q = _np_fasterquant(g["W0"], g["H"], g["scale"])
    codes = np.rint(Q / g["scale"][:, None])
    assert np.abs(codes * g["scale"][:, None] - Q).max() == 0 and codes.min() >= -8 and codes.max() <= 7
    mism = (Q != g["Wq"]).mean()
    assert mism < 0.02, mism

    def loss(Wq):
        D = (g["W0"] - Wq).astype(np.float64)
        return float(np.einsum("ij,jk,ik->", D, g["H"].astype(np.float64), D))

    assert abs(loss(Q) / loss(g["Wq"]) - 1) < 0.02


def test_silu_mul_oracle_matches_reference_mlp(golden):
    """oracle.silu_mul vs the reference's QWenMLP.forward (model/modeling_qwen_opt.py:382-386) inside
    the QWenBlock fixture: inputs = the outputs of the w1 / w2 wrappers, output = what mlp.c_proj was
    fed (16 token rows x 11008 channels, torch CPU fp16).  fp16 results; exp differs between libms by
    an fp32 ulp or two, which moves silu's fp16 rounding by one ulp for a few inputs (here a2 = -0.714)
    and the product by up to two."""
    g = golden("qwen_block")
    got = O.silu_mul(g["mlp_w1_y"], g["mlp_w2_y"])
    want = g["mlp_c_proj_x"]
    d = np.abs(got.view(np.int16).astype(np.int32) - want.view(np.int16).astype(np.int32))
    assert d.max() <= 2 and (d > 0).mean() < 1e-3, (d.max(), (d > 0).mean())
