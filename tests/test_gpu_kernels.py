"""GPU parity tests: every kernel is called through the C ABI (mquant_b200.ops -> ctypes) and
compared with the CPU oracle on the same seeded inputs.  Integer codes, scales and the integer
GEMM are bit-exact; the Hadamard rotation is compared with a stated tolerance (summation order
differs) and its quantisation stage is bit-exact on our own rotated values."""
import numpy as np
import pytest
import torch

from oracle import mq_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import mquant_b200
    from mquant_b200 import ops as _ops

    mquant_b200.lib()  # raises if the .so is missing: no silent fallback
    return _ops


@pytest.fixture
def force(ops):
    """Test-only dispatcher overrides (mq_debug_force_*), reset after the test."""
    from mquant_b200._lib import lib

    class F:
        def gemm(self, kernel):
            lib().mq_debug_force_gemm_kernel({"default": 0, "sk": 2, "v2": 3, "v4": 4, "v5": 5}[kernel])

        def hadamard_cuda_cores(self, on):
            lib().mq_debug_force_hadamard_kernel(1 if on else 0)

        def hadamard_tc2(self):  # FWHT in registers + hadK on tcgen05 (second generation)
            lib().mq_debug_force_hadamard_kernel(2)

        def hadamard_pow2(self, kernel):  # power-of-two sizes: "mma" = mma.sync, "tc" = tcgen05, for every M
            lib().mq_debug_force_hadamard_kernel({"auto": 0, "mma": 3, "tc": 4}[kernel])

    yield F()
    lib().mq_debug_force_gemm_kernel(0)
    lib().mq_debug_force_hadamard_kernel(0)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def rnd_act(rng, M, C, scale=3.0):
    x = rng.standard_normal((M, C)).astype(np.float32) * scale
    if C >= 64:
        x[:, 7] *= 40
        x[:, C // 2 + 3] *= 25
    return x.astype(np.float16)


def bits16(a):
    return np.asarray(a, dtype=np.float16).view(np.uint16)


# ------------------------------------------------------------------------------- act quant
def test_act_quant_golden(ops, golden):
    g = golden("act_quant")
    for tag in ("f16", "tie"):
        x = g[f"{tag}_x"]
        ref = g[f"{tag}_codes"]
        if tag == "tie":  # 2200 values in rows of 8: regroup into rows of 64 (n_out % 16 == 0)
            x, ref = x.reshape(-1)[:2176].reshape(-1, 64), ref.reshape(-1)[:2176].reshape(-1, 64)
        s = np.float32(g[f"{tag}_scale"])
        q, srow, _ = ops.act_quant_i8(dev(x), ops.make_scales(float(s), device="cuda"))
        ref = ref.reshape(-1, x.shape[-1]).astype(np.int8)
        assert np.array_equal(host(q), ref)
        assert np.all(host(srow) == s)


@pytest.mark.parametrize("M,C", [(1, 64), (37, 4096), (1024, 4096), (513, 11008)])
def test_act_quant_random_msq_pad(ops, M, C):
    rng = np.random.default_rng(M * 7 + C)
    x = rnd_act(rng, M, C)
    s_t, s_v = np.float32(np.float16(0.0731)), np.float32(np.float16(0.214))
    mod = (rng.random(M) < 0.4).astype(np.uint8)
    n_out = (C + 127) // 128 * 128 + 128
    q, srow, _ = ops.act_quant_i8(dev(x), ops.make_scales(float(s_t), float(s_v), "cuda"), dev(mod), n_out=n_out)
    q = host(q)
    ref = O.msq_quant_codes(x, mod, s_t, s_v).astype(np.int8)
    assert np.array_equal(q[:, :C], ref)
    assert not q[:, C:].any()
    assert np.array_equal(host(srow), O.msq_token_scales(mod, s_t, s_v))
    # degenerate MSQ (equal scales / no mask) == the reference's single-scale quantiser
    q1, _, _ = ops.act_quant_i8(dev(x), ops.make_scales(float(s_t), float(s_t), "cuda"), dev(mod))
    q2, _, _ = ops.act_quant_i8(dev(x), ops.make_scales(float(s_t), device="cuda"))
    assert np.array_equal(host(q1), O.act_quant_codes(x, s_t).astype(np.int8))
    assert torch.equal(q1, q2)


def test_act_quant_split_and_strided(ops):
    rng = np.random.default_rng(3)
    x = rnd_act(rng, 65, 512)
    s = np.float32(np.float16(0.05))
    q, _, x0 = ops.act_quant_i8(dev(x), ops.make_scales(float(s), device="cuda"), split=True)
    ref = O.act_quant_codes(x, s).astype(np.int8)
    ref[:, 0] = 0
    assert np.array_equal(host(q), ref)
    assert np.array_equal(bits16(host(x0)), bits16(x[:, 0]))
    # a column slice of a wider tensor (row stride != C)
    wide = dev(rnd_act(rng, 33, 1024))
    view = wide[:, 256:768]
    q, _, _ = ops.act_quant_i8(view, ops.make_scales(float(s), device="cuda"))
    assert np.array_equal(host(q), O.act_quant_codes(host(view), s).astype(np.int8))


def test_act_quant_every_fp16_value(ops):
    """Exhaustive over the input domain: all finite fp16 values x several scales (incl. ties)."""
    vals = np.arange(0, 0x7C00, dtype=np.uint16).view(np.float16)
    x = np.concatenate([vals, -vals]).reshape(-1, 64)
    for s in (0.25, 0.0731, 1.0, 3.0517578125e-05, 0.61865234375, 117.5):
        s = np.float32(np.float16(s))
        q, _, _ = ops.act_quant_i8(dev(x), ops.make_scales(float(s), device="cuda"))
        assert np.array_equal(host(q), O.act_quant_codes(x, s).astype(np.int8)), s


def test_fake_quant_every_fp16_value(ops):
    """INT8 fake quantisation, exhaustive over the input domain (every finite fp16 value of either sign,
    so also +-0 and the tiny negatives whose result is -0): fp16-valued scales take the division-free
    kernel (act_fake_quant_i8_kernel), the others the generic one; with a modality mask both scales of
    a call are exercised.  Bit-exact against the oracle, signs of zero included."""
    vals = np.arange(0, 0x7C00, dtype=np.uint16).view(np.float16)
    x = np.concatenate([vals, -vals]).reshape(-1, 64)
    mod = (np.arange(x.shape[0]) % 3 == 0).astype(np.uint8)
    for s_t, s_v in ((0.25, 0.0731), (1.0, 3.0517578125e-05), (0.61865234375, 117.5), (0.0213, 0.0571)):
        if s_t != 0.0213:  # (the last pair is not fp16-valued: generic kernel)
            s_t, s_v = np.float32(np.float16(s_t)), np.float32(np.float16(s_v))
        else:
            s_t, s_v = np.float32(s_t), np.float32(s_v)
        y = host(ops.act_fake_quant(dev(x), ops.make_scales(float(s_t), float(s_v), "cuda"), dev(mod)))
        ref = np.where(mod[:, None] != 0, O.act_fake_quant(x, s_v), O.act_fake_quant(x, s_t))
        assert np.array_equal(bits16(y), bits16(ref)), (s_t, s_v)
        y1 = host(ops.act_fake_quant(dev(x), ops.make_scales(float(s_t), device="cuda")))
        assert np.array_equal(bits16(y1), bits16(O.act_fake_quant(x, s_t))), s_t


@pytest.mark.parametrize("bits", [8, 16])
def test_fake_quant(ops, golden, bits):
    g = golden("act_quant")
    x = g["f16_x"]
    s = g["f16_scale"]
    y = ops.act_fake_quant(dev(x), ops.make_scales(float(s), device="cuda"), bits=bits)
    assert np.array_equal(bits16(host(y)), bits16(O.act_fake_quant(x, s, bits=bits)))
    if bits == 8:
        assert np.array_equal(bits16(host(y)), bits16(g["f16_y"]))
        # channel-wise scales
        xc, sc = g["f16cw_x"], g["f16cw_scale"]
        y = ops.act_fake_quant(dev(xc), dev(sc.astype(np.float32)), per_channel=True)
        assert np.array_equal(bits16(host(y)), bits16(g["f16cw_y"]))


def test_dynamic_per_token(ops, golden):
    g = golden("dynamic")
    x = g["x"]
    q, s = ops.act_quant_i8_dynamic(dev(x))
    assert np.array_equal(host(s).view(np.uint32), g["scale"].view(np.uint32))
    assert np.array_equal(host(q), O.dynamic_quant_codes(x, g["scale"]).astype(np.int8))
    rng = np.random.default_rng(11)
    x = rnd_act(rng, 301, 4096)
    x[5] = 0
    q, s = ops.act_quant_i8_dynamic(dev(x), clip_ratio=0.9, n_out=4096 + 128)
    sref = O.dynamic_per_token_scales(x, clip_ratio=0.9)
    assert np.array_equal(host(s).view(np.uint32), sref.view(np.uint32))
    assert np.array_equal(host(q)[:, :4096], O.dynamic_quant_codes(x, sref).astype(np.int8))


# ------------------------------------------------------------------------------- observers
@pytest.mark.parametrize("mode", ["layer_wise", "channel_wise"])
def test_minmax_observer(ops, golden, mode):
    from mquant_b200.bit_type import BIT_TYPE_DICT
    from mquant_b200.observer import build_observer

    tag = "f16" if mode == "layer_wise" else "f16cw"
    g = golden("act_quant")
    ob = build_observer("minmax", "activation", BIT_TYPE_DICT["int8"], mode)
    for c in g[f"{tag}_calib"]:
        ob.update(dev(c))
    assert np.array_equal(bits16(host(ob.min_val)), bits16(g[f"{tag}_min"]))
    assert np.array_equal(bits16(host(ob.max_val)), bits16(g[f"{tag}_max"]))
    scale, zp = ob.get_quantization_params()
    assert scale.dtype == torch.float16
    assert np.array_equal(bits16(host(scale)), bits16(g[f"{tag}_scale"]))
    assert not host(zp).any()


def test_minmax_large_and_signs(ops):
    rng = np.random.default_rng(5)
    for M, C in ((4096, 4096), (333, 11008)):
        x = rnd_act(rng, M, C)
        for per_channel in (False, True):
            st = torch.zeros(2 * C if per_channel else 2, dtype=torch.float32, device="cuda")
            ops.observer_minmax(dev(x[: M // 2]), st, per_channel)
            ops.observer_minmax(dev(x[M // 2:]), st, per_channel)
            ob = O.MinmaxObserver(8, "channel_wise" if per_channel else "layer_wise")
            ob.update(x)
            got = host(st)
            half = got.size // 2
            assert np.array_equal(got[:half].astype(np.float16), np.asarray(ob.min_val).reshape(-1))
            assert np.array_equal(got[half:].astype(np.float16), np.asarray(ob.max_val).reshape(-1))
    # all-positive / all-negative tensors keep the folded-in zero
    st = torch.zeros(2, dtype=torch.float32, device="cuda")
    ops.observer_minmax(dev(np.full((8, 64), 2.5, np.float16)), st)
    assert host(st).tolist() == [0.0, 2.5]
    st.zero_()
    ops.observer_minmax(dev(np.full((8, 64), -2.5, np.float16)), st)
    assert host(st).tolist() == [-2.5, 0.0]


def test_minmax_scale_sweep(ops, golden):
    """fp16 scale arithmetic on the device == the reference's (every fp16 magnitude)."""
    from mquant_b200.bit_type import BIT_TYPE_DICT
    from mquant_b200.observer import build_observer

    g = golden("minmax_params")
    ob = build_observer("minmax", "activation", BIT_TYPE_DICT["int8"], "channel_wise")
    ob.max_val, ob.min_val = dev(g["max16"]), dev(g["min16"])
    s, _ = ob.get_quantization_params()
    assert np.array_equal(bits16(host(s)), bits16(g["scale16"]))


def test_percentile_observer(ops, golden):
    from mquant_b200.bit_type import BIT_TYPE_DICT
    from mquant_b200.observer import build_observer

    g = golden("percentile")
    ob = build_observer("percentile", "activation", BIT_TYPE_DICT["int8"], "layer_wise")
    for c, st in zip(g["calib"], g["stats"]):
        ob.update(dev(c))
        assert np.float32(host(ob.min_val)) == st[0]
        assert np.float32(host(ob.max_val)) == st[1]
    s, _ = ob.get_quantization_params()
    assert np.array_equal(host(s).astype(np.float32).view(np.uint32), g["scale"].view(np.uint32))
    rng = np.random.default_rng(9)
    x = rnd_act(rng, 2048, 4096)
    got = host(ops.observer_percentile(dev(x), 0.99999))
    flat = np.sort(x.reshape(-1).astype(np.float32))
    assert got[1] == O.torch_quantile_f32(flat, 0.99999)
    assert got[0] == O.torch_quantile_f32(flat, 1.0 - 0.99999)


# ------------------------------------------------------------------------------- rmsn -> quant
@pytest.mark.parametrize("tag", ["h4096", "h3584", "h1000"])
def test_rmsn_quant_golden(ops, golden, tag):
    """mq_rmsn_quant_i8 vs the reference's RMSN + ActQuantizer (tests/golden/rmsn.npz).  The fp32 sum
    of squares runs in another order than torch's: the fp16 result may move by one ulp on < 1e-3 of
    the elements; the quantiser is bit-exact on OUR normalised values, and the end-to-end codes are
    the reference's up to +-1 on < 2e-3 of the elements."""
    g = golden("rmsn")
    x, yref, scale, eps = g[tag + "_x"], g[tag + "_y"], g[tag + "_scale"], float(g[tag + "_eps"])
    M, C = x.shape
    n_out = (C + 127) // 128 * 128
    scales = ops.make_scales(float(scale), None, "cuda")
    r = ops.rmsn_quant_i8(dev(x), C, eps, scales=scales, n_out=n_out, want_codes=True, want_f16=True)
    y, q = host(r["y"]), host(r["q"])
    d = np.abs(bits16(y).astype(np.int32) - bits16(yref).astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3, (d.max(), (d > 0).mean())
    assert np.array_equal(q[:, :C], O.act_quant_codes(y, scale).astype(np.int8))
    assert not q[:, C:].any()
    dc = np.abs(q[:, :C].astype(np.int32) - g[tag + "_codes"].astype(np.int32))
    assert dc.max() <= 1 and (dc > 0).mean() < 2e-3
    assert np.array_equal(host(r["s_row"]), np.full(M, scale, np.float32))
    # plain RMSN (no codes) gives the same fp16 tensor
    y2 = host(ops.rmsn_quant_i8(dev(x), C, eps, want_codes=False, want_f16=True)["y"])
    assert np.array_equal(bits16(y2), bits16(y))


@pytest.mark.parametrize("C", [4096, 5120, 8192, 2560])
def test_rmsn_quant_fused_equals_unfused(ops, C):
    """At the benchmark's row count: the fused kernel against RMSN followed by mq_act_quant_i8 (codes,
    per-token MSQ scales: bit-identical) and against the oracle on sampled rows."""
    M = 16384
    g = torch.Generator(device="cuda").manual_seed(C)
    x = torch.randn(M, C, device="cuda", generator=g) * 3
    x[:, 11] *= 50
    x = x.half()
    mod = (torch.rand(M, device="cuda", generator=g) < 0.4).to(torch.uint8)
    s, sv = float(torch.tensor(0.0213, dtype=torch.float16)), float(torch.tensor(0.0571, dtype=torch.float16))
    scales = ops.make_scales(s, sv, "cuda")
    fused = ops.rmsn_quant_i8(x, C, 1e-6, scales=scales, modality=mod, want_codes=True, want_f16=True)
    y = ops.rmsn_quant_i8(x, C, 1e-6, want_codes=False, want_f16=True)["y"]
    q, srow, _ = ops.act_quant_i8(y, scales, mod)
    assert torch.equal(fused["y"].view(torch.int16), y.view(torch.int16))
    assert torch.equal(fused["q"], q) and torch.equal(fused["s_row"], srow)
    rows = np.random.default_rng(C).integers(0, M, 48)
    ridx = torch.from_numpy(rows).cuda()
    yo = O.rmsn(host(x[ridx]), C, 1e-6)
    d = np.abs(bits16(host(y[ridx])).astype(np.int32) - bits16(yo).astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3
    # a non-fp16-valued scale takes the quantiser with the division fallback: still bit-exact
    scales2 = ops.make_scales(0.0213, 0.0571, "cuda")
    f2 = ops.rmsn_quant_i8(x[:512], C, 1e-6, scales=scales2, modality=mod[:512], want_codes=True, want_f16=False)
    q2, _, _ = ops.act_quant_i8(y[:512], scales2, mod[:512])
    assert torch.equal(f2["q"], q2)


# ------------------------------------------------------------------------------- hadamard
def had_check(ops, x, n, had_k, s_act, fp32_had=False, split=False, modality=None, s_vis=None):
    from mquant_b200 import hadamard_utils as hu

    hk = hu.hadK_device(had_k, "cuda") if had_k > 1 else None
    scales = ops.make_scales(float(s_act), None if s_vis is None else float(s_vis), "cuda")
    r = ops.hadamard(dev(x), had_k=had_k, hadk=hk, n=n, scales=scales, want_codes=True, want_f16=True,
                     fp32_had=fp32_had, split=split, modality=None if modality is None else dev(modality))
    y, q = host(r["y"]), host(r["q"])
    xp = O.pad_last(x, n)
    ref = O.matmul_hadU_cuda(xp, fp32_had=fp32_had)
    # (i) rotation: fp32 summation order differs -> at most 1 fp16 ulp, or tiny absolute error on
    # entries that cancel to ~0
    d = np.abs(bits16(y).astype(np.int32) - bits16(ref).astype(np.int32))
    a = np.abs(y.astype(np.float32) - ref.astype(np.float32))
    rms = np.sqrt((ref.astype(np.float32) ** 2).mean())
    assert np.all((d <= 1) | (a <= 1e-4 * rms)), (d.max(), a.max(), rms)
    assert (d > 0).mean() < 5e-3, (d > 0).mean()
    # (ii) quantisation stage: bit-exact on OUR rotated values
    if modality is None:
        own = O.act_quant_codes(y, s_act).astype(np.int8)
    else:
        own = O.msq_quant_codes(y, modality, s_act, s_vis).astype(np.int8)
    if split:
        own[:, 0] = 0
        assert np.array_equal(bits16(host(r["x0"])), bits16(y[:, 0]))
    assert np.array_equal(q, own)
    # (iii) codes vs the oracle's end-to-end codes: off-by-one only, rare (SURVEY.md hard part 1)
    if modality is None and not split:
        oc = O.act_quant_codes(ref, s_act).astype(np.int32)
        diff = np.abs(q.astype(np.int32) - oc)
        assert diff.max() <= 1
        assert (diff > 0).mean() < 2e-3, (diff > 0).mean()
    return y, q


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_hadamard_pow2(ops, n):
    rng = np.random.default_rng(n)
    M = 67 if n <= 2048 else 37
    x = rnd_act(rng, M, n)
    s = np.float32(np.float16(np.abs(x.astype(np.float32)).max() / 127 / 4))
    had_check(ops, x, n, 1, s)
    had_check(ops, x, n, 1, s, fp32_had=True)


@pytest.mark.parametrize("n,n_in,M", [(1024, 1024, 67), (1024, 960, 5), (2048, 2048, 33), (2048, 1984, 130), (4096, 4096, 37),
                                      (4096, 3968, 301), (8192, 8192, 19), (8192, 8128, 150), (4096, 4096, 1), (8192, 8192, 2),
                                      (4096, 3968, 1031), (2048, 2048, 1500)])
def test_hadamard_pow2_tcgen05(ops, force, n, n_in, M):
    """mq_hadamard_p2.cu (H_A (x) H_64 on tcgen05, hi / lo split of the intermediate; the default for
    M >= 1024) forced for small and ragged token counts: every tile / job tail (tokens past M are
    zero-filled tiles), zero padding from n_in to n, the split and MSQ variants, and the same
    tolerance as every other power-of-two kernel (fp32-grade: <= 1 fp16 ulp)."""
    force.hadamard_pow2("tc")
    rng = np.random.default_rng(n + M)
    x = rnd_act(rng, M, n_in)
    s = np.float32(np.float16(np.abs(x.astype(np.float32)).max() / 127 / 4))
    y, q = had_check(ops, x, n, 1, s)
    had_check(ops, x, n, 1, s, split=True)
    mod = (rng.random(M) < 0.5).astype(np.uint8)
    had_check(ops, x, n, 1, s, modality=mod, s_vis=np.float32(np.float16(s * 2.5)))
    # small-magnitude rows (the lo' part of the split must not underflow): 2^-12 x the same data
    xs = (x.astype(np.float32) * 2.0 ** -12).astype(np.float16)
    had_check(ops, xs, n, 1, np.float32(np.float16(s * 2.0 ** -12)))
    # against the mma.sync kernel: identical codes except rare +-1
    force.hadamard_pow2("mma")
    _, q2 = had_check(ops, x, n, 1, s)
    d = np.abs(q.astype(np.int32) - q2.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 2e-3


@pytest.mark.parametrize("path", ["default", "tc2", "simt"])
@pytest.mark.parametrize("n,n_in", [(11008, 11008), (13824, 13696), (14336, 14336), (19968, 18944),
                                    (5120, 5120), (3584, 3584), (3072, 3072), (5120 * 2, 5120 * 2)])
def test_hadamard_nonpow2(ops, force, path, n, n_in):
    """default = both Kronecker factors on tcgen05 (mq_hadamard_tc3.cu) for n / had_k = 64 / 128 and
    FWHT in registers + hadK on tcgen05 (mq_hadamard_tc2.cu) for 256 / 512; tc2 = the latter forced
    for every shape; simt = the CUDA-core kernel."""
    if path == "tc2":
        force.hadamard_tc2()
    else:
        force.hadamard_cuda_cores(path == "simt")
    k = O.get_hadK(n)[1]
    assert k > 1
    rng = np.random.default_rng(n)
    x = rnd_act(rng, 9, n_in)
    s = np.float32(np.float16(np.abs(x.astype(np.float32)).max() / 127 / 4))
    had_check(ops, x, n, k, s)
    if n == 11008:
        had_check(ops, x, n, k, s, fp32_had=True)
        had_check(ops, x, n, k, s, split=True)
        mod = (rng.random(9) < 0.5).astype(np.uint8)
        had_check(ops, x, n, k, s, modality=mod, s_vis=np.float32(np.float16(s * 2.5)))


@pytest.mark.parametrize("n,n_in,M", [(14336, 14336, 9), (14336, 14336, 1000), (3072, 3072, 70), (10240, 10240, 33)])
def test_hadamard_fast_had_flag(ops, n, n_in, M):
    """MQ_FLAG_FAST_HAD (opt-in; n / had_k in {256, 512}: InternVL2-8B's 14336 = 28 x 512): the FHT factor
    is split as H_E (x) H_128, the H_E part joins hadK on the tensor core, and the first fp16 rounding
    point moves from after the whole FHT to after its 128-point part.  Stated tolerance against the
    oracle (= the reference's rounding points): every output within 4 fp16 ulp OF ITS ROW'S MAXIMUM
    (the two results differ by the intermediate-rounding noise, which scales with the row, not the
    entry), relative Frobenius error <= 2^-10, INT8 codes within +-1, differing on < 1 % of the entries.
    Without the flag the same call runs the kernel that keeps the reference's rounding points."""
    from mquant_b200 import hadamard_utils as hu

    k = O.get_hadK(n)[1]
    rng = np.random.default_rng(n + M)
    x = rnd_act(rng, M, n_in)
    s = np.float32(np.float16(np.abs(x.astype(np.float32)).max() / 127 / 4))
    hk = hu.hadK_device(k, "cuda")
    scales = ops.make_scales(float(s), float(s), "cuda")
    r = ops.hadamard(dev(x), had_k=k, hadk=hk, n=n, scales=scales, want_codes=True, want_f16=True, fast_had=True)
    exact = ops.hadamard(dev(x), had_k=k, hadk=hk, n=n, scales=scales, want_codes=True, want_f16=True)
    y, q = host(r["y"]).astype(np.float32), host(r["q"]).astype(np.int32)
    rows = slice(0, min(M, 64))
    ref = O.matmul_hadU_cuda(O.pad_last(x[rows], n)).astype(np.float32)
    ulp_rowmax = np.spacing(np.abs(ref).max(axis=1, keepdims=True).astype(np.float16)).astype(np.float32)
    assert np.all(np.abs(y[rows] - ref) <= 4 * ulp_rowmax)
    assert np.linalg.norm(y[rows] - ref) <= 2.0 ** -10 * np.linalg.norm(ref)
    # the quantiser stage is exact on the kernel's own rotated values
    assert np.array_equal(q, O.act_quant_codes(host(r["y"]), s).astype(np.int32))
    d = np.abs(q - host(exact["q"]).astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 1e-2, (int(d.max()), float((d > 0).mean()))


@pytest.mark.parametrize("n,n_in,M", [(11008, 11008, 9), (11008, 11008, 300), (13824, 13696, 5),
                                      (19968, 18944, 7), (5120, 5120, 33), (3584, 3584, 130)])
def test_hadamard_nonpow2_tensor_core_path(ops, force, n, n_in, M):
    """codes-only requests on the tensor-core kernel (FWHT in registers + hadK on tcgen05); compare
    with the CUDA-core kernel and with the oracle."""
    from mquant_b200 import hadamard_utils as hu

    k = O.get_hadK(n)[1]
    rng = np.random.default_rng(n + M)
    x = rnd_act(rng, M, n_in)
    s = np.float32(np.float16(np.abs(x.astype(np.float32)).max() / 127 / 4))
    s_vis = np.float32(np.float16(s * 2.5))
    mod = (rng.random(M) < 0.5).astype(np.uint8)
    hk = hu.hadK_device(k, "cuda")
    scales = ops.make_scales(float(s), float(s_vis), "cuda")
    xd, md = dev(x), dev(mod)
    for split in (False, True):
        force.hadamard_cuda_cores(False)
        tc = ops.hadamard(xd, had_k=k, hadk=hk, n=n, scales=scales, modality=md, want_codes=True,
                          want_f16=False, split=split)
        force.hadamard_cuda_cores(True)
        ref = ops.hadamard(xd, had_k=k, hadk=hk, n=n, scales=scales, modality=md, want_codes=True,
                           want_f16=True, split=split)  # CUDA-core kernel
        force.hadamard_cuda_cores(False)
        q_tc, q_ref = host(tc["q"]).astype(np.int32), host(ref["q"]).astype(np.int32)
        d = np.abs(q_tc - q_ref)
        assert d.max() <= 1, d.max()
        assert (d > 0).mean() < 2e-3, (d > 0).mean()
        assert np.array_equal(host(tc["s_row"]), O.msq_token_scales(mod, s, s_vis))
        if split:
            assert not q_tc[:, 0].any()
            x0_tc, x0_ref = host(tc["x0"]).astype(np.float32), host(ref["x0"]).astype(np.float32)
            assert np.abs(x0_tc - x0_ref).max() <= 2.0 ** -9 * max(1.0, np.abs(x0_ref).max())
        else:
            yo = O.matmul_hadU_cuda(O.pad_last(x, n))
            oc = O.msq_quant_codes(yo, mod, s, s_vis).astype(np.int32)
            d = np.abs(q_tc - oc)
            assert d.max() <= 1
            assert (d > 0).mean() < 2e-3, (d > 0).mean()


def test_matmul_hadU_cuda_api(ops, golden):
    from mquant_b200 import hadamard_utils as hu

    g = golden("hadamard")
    for n in (512, 11008):
        x = g[f"hadcuda_x_{n}"]
        hk, K = hu.get_hadK(n)
        y = host(hu.matmul_hadU_cuda(dev(x), hk, K))
        ref = g[f"hadcuda_y_{n}"]
        d = np.abs(bits16(y).astype(np.int32) - bits16(ref).astype(np.int32))
        a = np.abs(y.astype(np.float32) - ref.astype(np.float32))
        rms = np.sqrt((ref.astype(np.float32) ** 2).mean())
        assert np.all((d <= 1) | (a <= 1e-4 * rms))
    # fp32 weights (offline rotation) run the torch butterfly: x H H^T = x
    w = torch.randn(16, 11008, device="cuda")
    hk, K = hu.get_hadK(11008)
    y = hu.matmul_hadU_cuda(w, hk, K)
    ref = O.matmul_hadU(host(w).astype(np.float64))
    np.testing.assert_allclose(host(y), ref, atol=2e-5)


def test_empty_batches_are_noops(ops):
    """M = 0 (an empty torch tensor has a null data pointer): every entry point returns without
    touching its outputs, like the reference on an empty batch."""
    from mquant_b200 import hadamard_utils as hu

    x = torch.empty((0, 4096), dtype=torch.float16, device="cuda")
    scales = ops.make_scales(0.1, 0.2, "cuda")
    mod = torch.empty((0,), dtype=torch.uint8, device="cuda")
    q, srow, _ = ops.act_quant_i8(x, scales, mod, want_row_scales=True)
    assert q.shape == (0, 4096)
    assert ops.act_fake_quant(x, scales, mod).shape == (0, 4096)
    st = torch.tensor([-1.0, 2.0], device="cuda")
    ops.observer_minmax(x, st)
    assert host(st).tolist() == [-1.0, 2.0]
    st_v = torch.zeros(2, device="cuda")
    seen = torch.zeros(2, dtype=torch.int32, device="cuda")
    ops.observer_minmax_msq(x, mod, st, st_v, seen)
    assert host(seen).tolist() == [0, 0] and host(st).tolist() == [-1.0, 2.0]
    r = ops.hadamard(x, n=4096, scales=scales, want_codes=True, want_f16=True)
    assert r["q"].shape == (0, 4096) and r["y"].shape == (0, 4096)
    x2 = torch.empty((0, 11008), dtype=torch.float16, device="cuda")
    r = ops.hadamard(x2, had_k=172, hadk=hu.hadK_device(172, "cuda"), n=11008, scales=scales, want_codes=True, want_f16=False)
    assert r["q"].shape == (0, 11008)
    w4 = ops.pack_w4(torch.zeros((256, 128), dtype=torch.int8, device="cuda"))
    out = ops.w4a8_gemm(torch.empty((0, 128), dtype=torch.int8, device="cuda"), w4, torch.ones(1, device="cuda"),
                        torch.ones(256, device="cuda"), 256)
    assert out.shape == (0, 256)
    torch.cuda.synchronize()


def test_observer_minmax_msq_fused(ops):
    """One pass for both modalities (mq_observer_minmax_msq) == two MinmaxObservers on x[~mask] and
    x[mask]; row counts tell which modality occurred."""
    rng = np.random.default_rng(11)
    st_t = torch.zeros(2, device="cuda")
    st_v = torch.zeros(2, device="cuda")
    seen = torch.zeros(2, dtype=torch.int32, device="cuda")
    ot, ov = O.MinmaxObserver(), O.MinmaxObserver()
    n_t = n_v = 0
    for M, C, frac in [(300, 4096, 0.3), (17, 11008, 0.0), (1024, 512, 0.9), (5, 64, 1.0)]:
        x = rnd_act(rng, M, C)
        mod = (rng.random(M) < frac).astype(np.uint8)
        ops.observer_minmax_msq(dev(x), dev(mod), st_t, st_v, seen)
        if (mod == 0).any():
            ot.update(x[mod == 0])
        if (mod != 0).any():
            ov.update(x[mod != 0])
        n_t += int((mod == 0).sum())
        n_v += int((mod != 0).sum())
    assert host(seen).tolist() == [n_t, n_v]
    assert np.array_equal(host(st_t), np.array([ot.min_val, ot.max_val], dtype=np.float32))
    assert np.array_equal(host(st_v), np.array([ov.min_val, ov.max_val], dtype=np.float32))


# ------------------------------------------------------------------------------- weights
def test_weight_rtn_and_pack(ops, golden):
    g = golden("weight_quant")
    W = g["f16_W"]
    codes, s_w = ops.weight_quant_rtn(dev(W), 4)
    assert np.array_equal(host(s_w).view(np.uint32), g["f16_mse0_scale"].reshape(-1).view(np.uint32))
    assert np.array_equal(host(codes), g["f16_mse0_codes"])
    # clip-search scales supplied from outside
    codes2, _ = ops.weight_quant_rtn(dev(W), 4, dev(g["f16_mse1_scale"]))
    assert np.array_equal(host(codes2), g["f16_mse1_codes"])
    # codes are exactly recoverable from the reference's fake-quantised fp16 weight
    codes3, _ = ops.weight_quant_rtn(dev(g["f16_mse0_fq"]), 4, dev(g["f16_mse0_scale"]))
    assert np.array_equal(host(codes3), g["f16_mse0_codes"])


@pytest.mark.parametrize("N,K", [(96, 256), (256, 128), (300, 1024), (1024, 4096)])
def test_pack_roundtrip(ops, N, K):
    rng = np.random.default_rng(N + K)
    codes = rng.integers(-8, 8, (N, K)).astype(np.int8)
    w4 = ops.pack_w4(dev(codes))
    assert w4.numel() == ((N + 255) // 256) * (K // 128) * 16384
    assert np.array_equal(host(ops.unpack_w4(w4, N, K)), codes)
    ref_packed = O.pack_i4(codes)
    w4b = ops.repack_w4_from_ref(dev(ref_packed), N, K)
    assert torch.equal(w4, w4b)
    w4c = ops.pack_w4(dev(codes), zero_col0=True)
    c0 = codes.copy()
    c0[:, 0] = 0
    assert np.array_equal(host(ops.unpack_w4(w4c, N, K)), c0)


# ------------------------------------------------------------------------------- GEMM
def gemm_case(rng, M, N, K, per_row, bias, split):
    qa = rng.integers(-128, 128, (M, K)).astype(np.int8)
    cw = rng.integers(-8, 8, (N, K)).astype(np.int8)
    s_a = (rng.random(M if per_row else 1).astype(np.float32) * 0.05 + 0.01)
    s_w = (rng.random(N).astype(np.float32) * 0.01 + 0.001)
    b = (rng.standard_normal(N) * 0.5).astype(np.float16) if bias else None
    x0 = w0 = None
    if split:
        qa[:, 0] = 0
        cw[:, 0] = 0
        x0 = (rng.standard_normal(M) * 30).astype(np.float16)
        w0 = (rng.standard_normal(N) * 0.02).astype(np.float16)
    return qa, cw, s_a, s_w, b, x0, w0


def run_gemm(ops, case, simt):
    qa, cw, s_a, s_w, b, x0, w0 = case
    w4 = ops.pack_w4(dev(cw))
    out = ops.w4a8_gemm(dev(qa), w4, dev(s_a), dev(s_w), cw.shape[0],
                        bias=None if b is None else dev(b), x0=None if x0 is None else dev(x0),
                        w0=None if w0 is None else dev(w0), simt=simt)
    torch.cuda.synchronize()
    return host(out)


def oracle_gemm(case):
    qa, cw, s_a, s_w, b, x0, w0 = case
    # exact integer accumulation via float64 BLAS (|acc| < 2^53)
    acc = (qa.astype(np.float64) @ cw.astype(np.float64).T)
    s = (s_a.reshape(-1, 1) * s_w.reshape(1, -1)).astype(np.float32)
    y = (acc.astype(np.float32) * s).astype(np.float32)
    if x0 is not None:
        y = (y + (x0.astype(np.float32).reshape(-1, 1) * w0.astype(np.float32).reshape(1, -1))).astype(np.float32)
    if b is not None:
        y = (y + b.astype(np.float32).reshape(1, -1)).astype(np.float32)
    return y.astype(np.float16)


@pytest.mark.parametrize("M,N,K,per_row,bias,split", [
    (5, 64, 128, False, False, False),
    (130, 264, 256, True, True, False),
    (64, 512, 1024, True, True, True),
])
def test_gemm_simt_matches_oracle(ops, M, N, K, per_row, bias, split):
    case = gemm_case(np.random.default_rng(M + N + K), M, N, K, per_row, bias, split)
    assert np.array_equal(bits16(run_gemm(ops, case, simt=True)), bits16(oracle_gemm(case)))
    # the two oracle formulations agree (int32 matmul restatement vs float64 BLAS)
    qa, cw, s_a, s_w, b, x0, w0 = case
    assert np.array_equal(bits16(O.qlinear_int(qa, s_a, cw, s_w, b, x0, w0)), bits16(oracle_gemm(case)))


@pytest.mark.parametrize("kernel", ["default", "v2", "v4", "v5"])
@pytest.mark.parametrize("M,N,K,per_row,bias,split", [
    (128, 256, 128, False, False, False),     # one tile, one k-block
    (128, 256, 512, False, False, False),     # pipeline wraps once
    (300, 512, 1024, True, True, False),      # ragged M, per-row scales, bias
    (1, 256, 4096, False, True, False),       # decode-shaped
    (1000, 1000, 1152, True, False, True),    # ragged N (N % 256 != 0), split rank-1 term
    (2048, 4096, 4096, True, True, False),    # several tiles per CTA (both TMEM buffers, phases)
    (600, 768, 640, True, True, True),        # 256-row tiles: last tile's second half is all padding
    (5000, 1024, 2048, True, False, False),   # many tiles per CTA, ragged M
])
def test_gemm_tcgen05_matches_oracle(ops, force, kernel, M, N, K, per_row, bias, split):
    """The dispatcher's choice and each of the two persistent tcgen05 kernels forced onto every
    shape (128 x 256 and 256 x 256 tiles) are bit-exact against the oracle."""
    force.gemm(kernel)
    case = gemm_case(np.random.default_rng(M * 3 + N + K), M, N, K, per_row, bias, split)
    got = run_gemm(ops, case, simt=False)
    assert np.array_equal(bits16(got), bits16(oracle_gemm(case)))


@pytest.mark.parametrize("M,N,K,per_row,bias,split", [
    (1, 4096, 4096, False, True, False),
    (3, 1000, 1152, True, False, True),      # ragged N, split rank-1 term
    (16, 11008, 4096, True, True, False),
    (9, 4096, 11008, True, False, False),
])
def test_gemm_decode_weight_streaming_path(ops, M, N, K, per_row, bias, split):
    """M <= 16 takes the split-K dp4a kernel (int32 atomics in a self-cleaning workspace): bit-exact,
    also when the same workspace is reused by back-to-back calls."""
    rng = np.random.default_rng(M + N + K)
    for rep in range(3):
        case = gemm_case(rng, M, N, K, per_row, bias, split)
        got = run_gemm(ops, case, simt=False)
        assert np.array_equal(bits16(got), bits16(oracle_gemm(case))), rep


@pytest.mark.parametrize("M,N,K,per_row,bias,split", [
    (64, 4096, 4096, True, False, False),
    (200, 1000, 2304, True, True, True),     # ragged M and N, split rank-1 term, uneven K slices
    (256, 4096, 11008, False, True, False),  # uneven K slices (86 k-blocks)
    (500, 768, 2048, True, False, True),
    (17, 4096, 11008, True, True, False),
])
def test_gemm_split_k_path(ops, force, M, N, K, per_row, bias, split):
    """The split-K kernel (int32 partial sums in the self-cleaning workspace, the last CTA of a tile
    applies the epilogue) wherever it is applicable (the dispatcher takes it when its cost model beats
    the one-wave kernels: few tiles, long K): bit-exact, also back to back on one workspace."""
    from mquant_b200._lib import lib

    assert lib().mq_w4a8_gemm_workspace_bytes(M, N, K) > 0
    force.gemm("sk")
    rng = np.random.default_rng(M + N + K)
    for rep in range(3):
        case = gemm_case(rng, M, N, K, per_row, bias, split)
        got = run_gemm(ops, case, simt=False)
        assert np.array_equal(bits16(got), bits16(oracle_gemm(case))), rep


@pytest.mark.parametrize("M,N,K", [(1024, 4096, 4096), (1024, 4096, 11008), (700, 4096 + 8, 4096)])
def test_gemm_single_request_default_dispatch(ops, M, N, K):
    """One 1024-token request (one wave of tiles): the DEFAULT dispatcher must pick the kernel that keeps
    the converted weights in tensor memory (mq_gemm_w4a8_v5.cu: weights = the A operand, accumulator
    transposed), bit-exact vs the exact-integer oracle incl. ragged M / N, bias and the split term."""
    from mquant_b200._lib import lib

    assert lib().mq_w4a8_gemm_plan(M, N, K, lib().mq_w4a8_gemm_workspace_bytes(M, N, K)) == 5
    case = gemm_case(np.random.default_rng(M + N + K), M, N, K, True, True, K == 4096)
    assert np.array_equal(bits16(run_gemm(ops, case, simt=False)), bits16(oracle_gemm(case)))


# The shapes the headline benchmark runs (bench.py: Qwen-VL-Chat prefill, 16384 tokens per step; q|k|v
# and w1|w2 fused) plus the GLM-4V widths of BASELINE config 5 (13696 -> K padded to 13824).
BENCH_SHAPES = [(4096, 4096), (4096, 12288), (4096, 22016), (11008, 4096), (13824, 4096), (4096, 13696)]


@pytest.mark.parametrize("K,N", BENCH_SHAPES)
def test_gemm_default_dispatch_at_benchmark_shapes(ops, K, N):
    """M = 16384 through the DEFAULT dispatcher (must pick the 256 x 256 kernel: 64 m-tiles x up to
    86 n-tiles x up to 108 k-blocks per launch), per-row scales + bias (+ split rank-1 term), checked
    bit for bit against the exact integer oracle on 640 sampled rows x all columns."""
    from mquant_b200._lib import lib

    M = 16384
    split = N == 4096 and K == 11008
    assert lib().mq_w4a8_gemm_plan(M, N, K, lib().mq_w4a8_gemm_workspace_bytes(M, N, K)) == 4
    g = torch.Generator(device="cuda").manual_seed(K * 7 + N)
    qa = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda", generator=g)
    cw = torch.randint(-8, 8, (N, K), dtype=torch.int8, device="cuda", generator=g)
    s_a = torch.rand(M, device="cuda", generator=g) * 0.05 + 0.01
    s_w = torch.rand(N, device="cuda", generator=g) * 0.01 + 0.001
    bias = (torch.randn(N, device="cuda", generator=g) * 0.5).half()
    x0 = w0 = None
    if split:
        qa[:, 0] = 0
        cw[:, 0] = 0
        x0 = (torch.randn(M, device="cuda", generator=g) * 30).half()
        w0 = (torch.randn(N, device="cuda", generator=g) * 0.02).half()
    out = ops.w4a8_gemm(qa, ops.pack_w4(cw), s_a, s_w, N, bias=bias, x0=x0, w0=w0)
    torch.cuda.synchronize()
    rng = np.random.default_rng(N + K)
    rows = np.unique(np.concatenate([np.arange(0, 64), np.arange(M - 64, M), np.arange(8128, 8256),
                                     rng.integers(0, M, 400)]))
    ridx = torch.from_numpy(rows).cuda()
    case = (host(qa[ridx]), host(cw), host(s_a[ridx]), host(s_w), host(bias),
            None if x0 is None else host(x0[ridx]), None if w0 is None else host(w0))
    want = oracle_gemm(case)
    got = host(out[ridx])
    assert np.array_equal(bits16(got), bits16(want)), int((bits16(got) != bits16(want)).sum())
    # every row that was not sampled: a cheap whole-matrix property (finite, and the column sums of
    # the sampled block are reproduced when the GEMM is re-run on the sampled rows alone)
    assert bool(torch.isfinite(out).all())
    again = ops.w4a8_gemm(qa[ridx].contiguous(), ops.pack_w4(cw), s_a[ridx].contiguous(), s_w, N, bias=bias,
                          x0=None if x0 is None else x0[ridx].contiguous(), w0=w0)
    assert torch.equal(again, out[ridx])


def test_gemm_seeded_fuzz(ops):
    """A seeded slice of tools/fuzz.py inside the suite: random shapes (K up to 19968, ragged M and N)
    through the dispatcher against the CUDA-core kernel, bit-exact."""
    rng = np.random.default_rng(2026)
    for it in range(40):
        K = 128 * int(rng.integers(1, 157))
        N = 8 * int(rng.integers(1, 400))
        M = int(rng.choice([1, 3, 17, 64, 129, 300, 777, 1500]))
        if it % 8 == 0:
            K, N, M = 19968, 8 * int(rng.integers(30, 200)), int(rng.choice([5, 130, 520]))
        case = gemm_case(rng, M, N, K, bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))
        got, ref = run_gemm(ops, case, simt=False), run_gemm(ops, case, simt=True)
        assert np.array_equal(bits16(got), bits16(ref)), (M, N, K)


@pytest.mark.parametrize("n,n_in", [(4096, 4096), (8192, 8192), (11008, 11008), (13824, 13696), (19968, 18944)])
def test_hadamard_at_benchmark_rows(ops, force, n, n_in):
    """M = 16384 with the benchmark's outlier distribution: the fused kernel against the CUDA-core
    kernel on every row (codes within +-1, rare) and against the oracle on 96 sampled rows."""
    from mquant_b200 import hadamard_utils as hu

    M = 16384
    k = O.get_hadK(n)[1]
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.randn(M, n_in, device="cuda", generator=g) * 3
    x[:, 7] *= 40
    x[:, n_in // 2 + 3] *= 25
    x = x.half()
    s = float(torch.tensor(float(x.abs().max()) / 127 / 4, dtype=torch.float16))
    s_vis = float(torch.tensor(s * 2.5, dtype=torch.float16))
    mod = (torch.rand(M, device="cuda", generator=g) < 0.4).to(torch.uint8)
    scales = ops.make_scales(s, s_vis, "cuda")
    hk = hu.hadK_device(k, "cuda") if k > 1 else None
    fast = ops.hadamard(x, had_k=k, hadk=hk, n=n, scales=scales, modality=mod, want_codes=True, want_f16=False)
    force.hadamard_cuda_cores(True)
    slow = ops.hadamard(x, had_k=k, hadk=hk, n=n, scales=scales, modality=mod, want_codes=True, want_f16=False)
    force.hadamard_cuda_cores(False)
    d = (fast["q"].int() - slow["q"].int()).abs()
    assert int(d.max()) <= 1
    assert float((d > 0).float().mean()) < 2e-3
    assert torch.equal(fast["s_row"], slow["s_row"])
    rows = np.concatenate([np.arange(0, 32), np.arange(M - 32, M), np.random.default_rng(n).integers(0, M, 32)])
    ridx = torch.from_numpy(rows).cuda()
    ref = O.matmul_hadU_cuda(O.pad_last(host(x[ridx]), n))
    oc = O.msq_quant_codes(ref, host(mod[ridx]), np.float32(s), np.float32(s_vis)).astype(np.int32)
    d = np.abs(host(fast["q"][ridx]).astype(np.int32) - oc)
    assert d.max() <= 1
    assert (d > 0).mean() < 2e-3, (d > 0).mean()


@pytest.mark.parametrize("n", [64, 256, 1024, 4096, 8192])
def test_hadamard_transform_seam(ops, force, n):
    """mq_hadamard_transform_f16 == fast_hadamard_transform.hadamard_transform(x, scale): any scale,
    strided leading dims (the reference calls it on a transposed view, quant_utils.py:351)."""
    from mquant_b200.fht import hadamard_transform

    rng = np.random.default_rng(n)
    x = rnd_act(rng, 50, n)
    for scale in (1.0, 1.0 / np.sqrt(n), 0.37, -0.37):
        if scale < 0:  # the tcgen05 kernel (default from 1024 rows on) on the same 50 rows
            force.hadamard_pow2("tc")
            scale = -scale
        y = host(hadamard_transform(dev(x).reshape(5, 10, n), scale))
        ref = O.fht(x, np.float32(scale)).reshape(5, 10, n)
        d = np.abs(bits16(y).astype(np.int32) - bits16(ref).astype(np.int32))
        a = np.abs(y.astype(np.float32) - ref.astype(np.float32))
        rms = np.sqrt((ref.astype(np.float32) ** 2).mean())
        assert np.all((d <= 1) | (a <= 1e-4 * rms)), (n, scale, d.max())
    # fp32 / odd sizes: the torch butterfly (offline weight rotation), zero padded like the original
    w = torch.randn(7, 96, device="cuda")
    yw = host(hadamard_transform(w, 0.5))
    refw = O.fwht_last(np.pad(host(w).astype(np.float64), ((0, 0), (0, 32))))[:, :96] * 0.5
    assert np.allclose(yw, refw, atol=1e-4)


def test_gemm_extreme_codes(ops):
    """Saturated operands: every activation -128, every weight -8 / +7 (largest |acc|)."""
    M, N, K = 128, 256, 19968
    qa = np.full((M, K), -128, np.int8)
    cw = np.where(np.arange(N)[:, None] % 2 == 0, -8, 7).astype(np.int8) * np.ones((1, K), np.int8)
    case = (qa, cw, np.array([0.01], np.float32), np.full(N, 1e-3, np.float32), None, None, None)
    assert np.array_equal(bits16(run_gemm(ops, case, simt=False)), bits16(oracle_gemm(case)))


# ------------------------------------------------------------------------------- GPTQ (f-2)
def test_gptq_hessian_golden(ops, golden):
    """mq_gptq_hessian_update vs GPTQ.add_batch of the reference (tests/golden/gptq.npz): three
    batches accumulated in place.  fp16 x fp16 products are exact in fp32; only the order of the
    fp32 additions differs from the reference's matmul."""
    g = golden("gptq")
    C = g["H"].shape[0]
    H = torch.zeros((C, C), dtype=torch.float32, device="cuda")
    H64, n = np.zeros((C, C)), 0
    for b in range(3):
        x = g[f"x{b}"]
        ops.gptq_hessian_update(dev(x), H, n, x.shape[0])
        H64, n = O.gptq_hessian(H64, n, x, x.shape[0])
    assert n == int(g["nsamples"])
    top = np.abs(g["H"]).max()
    assert np.abs(host(H) - g["H"]).max() <= 2e-6 * top
    assert np.abs(host(H) - H64).max() <= 2e-6 * top
    assert np.array_equal(host(H), host(H).T)  # tile (i, j) and tile (j, i) add in the same order


@pytest.mark.parametrize("T,C,ld", [(1, 8, 8), (63, 136, 136), (1000, 4096, 4096), (4100, 1032, 1040), (2048, 11008, 11008)])
def test_gptq_hessian_shapes(ops, T, C, ld):
    """Ragged token counts, C not a multiple of the 128 x 256 tile, strided rows, the benchmark's
    widths; against a float64 X^T X on sampled rows of H."""
    rng = np.random.default_rng(T + C)
    xs = torch.zeros((T, ld), dtype=torch.float16, device="cuda")
    xs[:, :C] = dev((rng.standard_normal((T, C)) * (1 + 4 * rng.random(C))).astype(np.float16))
    x = xs[:, :C]
    H0 = dev(rng.standard_normal((C, C)).astype(np.float32))
    H = H0.clone()
    n0, t = 5, 3
    ops.gptq_hessian_update(x, H, n0, t)
    rows = torch.from_numpy(rng.choice(C, size=min(C, 64), replace=False)).cuda()
    x64 = x.double()
    want = H0[rows].double() * (n0 / (n0 + t)) + (2.0 / (n0 + t)) * (x64[:, rows].t() @ x64)
    err = (H[rows].double() - want).abs().max().item()
    # fp32-grade: within a small multiple of the error of the reference's own fp32 matmul
    assert err <= 4e-6 * want.abs().max().item(), err
    ops.gptq_hessian_update(x[:0], H, n0, t)  # no tokens: H untouched (the reference would only rescale by n/(n+0))


@pytest.mark.parametrize("rows,count", [(1, 1), (96, 128), (300, 128), (4096, 128), (257, 40)])
def test_gptq_block_bit_exact(ops, rows, count):
    """mq_gptq_block_f32 vs oracle.gptq_block (the reference's column loop, fp32 operation by fp32
    operation): W1, Q1, Err1, Losses1 bit-identical.  W and Hinv are views into wider matrices."""
    rng = np.random.default_rng(rows * 131 + count)
    cols = count + 72
    W = (rng.standard_normal((rows, cols)) * 0.05).astype(np.float32)
    A = rng.standard_normal((cols, 2 * cols)).astype(np.float32)
    Hinv = np.linalg.cholesky((A @ A.T / (2 * cols) + 0.05 * np.eye(cols)).astype(np.float32)).T.astype(np.float32)
    scale = (np.abs(W).max(axis=1) / 7).astype(np.float32)
    i1, i2 = 8, 8 + count
    Wd = dev(W)
    Q, Err, Losses = ops.gptq_block(Wd, i1, i2, dev(Hinv), dev(scale), 7)
    W1, Q1, Err1, L1 = O.gptq_block(W[:, i1:i2], Hinv[i1:i2, i1:i2], scale, 7)
    for name, got, want in (("Q", Q, Q1), ("Err", Err, Err1), ("Losses", Losses, L1), ("W", Wd[:, i1:i2], W1)):
        assert np.array_equal(host(got).view(np.uint32), want.view(np.uint32)), name
    untouched = np.ones(cols, bool)
    untouched[i1:i2] = False
    assert np.array_equal(host(Wd)[:, untouched], W[:, untouched])


def test_gptq_fasterquant_matches_reference_golden(ops, golden):
    """mquant_b200.gptq.GPTQ (add_batch + fasterquant on the kernels) on the reference's golden case:
    the quantised weight is the reference's up to the entries a differently rounded Cholesky factor
    flips, the GPTQ objective tr(D H D^T) agrees, and with the SAME factor the kernel loop is
    bit-identical to the reference's eager loop."""
    from mquant_b200 import gptq as G, quant_utils as qu

    g = golden("gptq")
    rows, cols = g["W0"].shape
    layer = torch.nn.Linear(cols, rows, bias=False).cuda()
    layer.weight.data = dev(g["W0"])
    gp = G.GPTQ(layer)
    gp.quantizer = qu.WeightQuantizer()
    gp.quantizer.configure(4, perchannel=True, sym=True, mse=False)
    for b in range(3):
        gp.add_batch(dev(g[f"x{b}"]), None)
    assert gp.nsamples == int(g["nsamples"])
    Hd = gp.H.clone()
    gp.fasterquant(percdamp=0.01, groupsize=-1)
    assert np.array_equal(host(gp.quantizer.scale).reshape(-1).view(np.uint32), g["scale"].view(np.uint32))
    Wq = host(layer.weight.data)
    assert (Wq != g["Wq"]).mean() < 0.02

    def loss(Wx):
        D = (g["W0"] - Wx).astype(np.float64)
        return float(np.einsum("ij,jk,ik->", D, g["H"].astype(np.float64), D))

    assert abs(loss(Wq) / loss(g["Wq"]) - 1) < 0.02
    # the reference's loop (gptq_utils.py:230-290) in eager torch on the same device and factor
    W = dev(g["W0"]).clone()
    H = Hd.clone()
    H.diagonal().add_(0.01 * torch.mean(torch.diag(H)))
    Hinv = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(H)), upper=True)
    s = gp.quantizer.scale.reshape(-1, 1)
    Qe = torch.zeros_like(W)
    for i1 in range(0, cols, 128):
        i2 = min(i1 + 128, cols)
        W1, Hinv1 = W[:, i1:i2].clone(), Hinv[i1:i2, i1:i2]
        Err1 = torch.zeros_like(W1)
        for i in range(i2 - i1):
            w, d = W1[:, i], Hinv1[i, i]
            q = (s * torch.clamp(torch.round(w.unsqueeze(1) / s), -8, 7)).flatten()
            Qe[:, i1 + i] = q
            err1 = (w - q) / d
            W1[:, i:] -= err1.unsqueeze(1).matmul(Hinv1[i, i:].unsqueeze(0))
            Err1[:, i] = err1
        W[:, i2:] -= Err1.matmul(Hinv[i1:i2, i2:])
    assert torch.equal(Qe, layer.weight.data)


# ------------------------------------------------------------------------------- gated MLP (f-3)
def _ulp16(a, b):
    return np.abs(np.asarray(a, np.float16).view(np.int16).astype(np.int32) - np.asarray(b, np.float16).view(np.int16).astype(np.int32))


@pytest.mark.parametrize("kernel,M,F,K,per_row,bias", [
    ("v4", 300, 256, 256, True, False),
    ("v4", 1000, 1280, 512, True, True),
    ("v4", 257, 128, 1152, False, True),
    ("default", 40, 256, 256, True, True),     # split-K / 128 x 256 kernels: GEMM into tmp, then the gate kernel
    ("default", 3, 384, 512, False, False),    # decode kernel
])
def test_gemm_gated_matches_oracle(ops, force, kernel, M, F, K, per_row, bias):
    """mq_w4a8_gemm_gated_ws (a1 * silu(a2) in the epilogue of the w1 | w2 GEMM over interleaved rows,
    QWenMLP.forward model/modeling_qwen_opt.py:382-386) vs the oracle: a1 / a2 = the exact integer
    GEMM + epilogue, then oracle.silu_mul (pinned to the reference's MLP in test_oracle_golden.py).
    expf on the device and numpy's exp differ by an fp32 ulp on some inputs: <= 2 fp16 ulp on < 1e-3 of
    the entries; against torch's own CUDA silu (the reference's operator on a GPU) bit for bit."""
    rng = np.random.default_rng(M * 3 + F + K)
    c1 = gemm_case(rng, M, F, K, per_row, bias, False)
    c2 = gemm_case(rng, M, F, K, per_row, bias, False)
    qa, s_a = c1[0], c1[2]
    a1 = oracle_gemm((qa, c1[1], s_a, c1[3], c1[4], None, None))
    a2 = oracle_gemm((qa, c2[1], s_a, c2[3], c2[4], None, None))
    want = O.silu_mul(a1, a2)
    w4 = ops.pack_w4(ops.interleave_gated(dev(c1[1]), dev(c2[1])))
    s_w = ops.interleave_gated(dev(c1[3]), dev(c2[3]))
    b = ops.interleave_gated(dev(c1[4]), dev(c2[4])) if bias else None
    if kernel != "default":
        force.gemm(kernel)
    got = ops.w4a8_gemm_gated(dev(qa), w4, dev(s_a), s_w, F, bias=b)
    torch.cuda.synchronize()
    d = _ulp16(host(got), want)
    assert d.max() <= 2 and (d > 0).mean() < 1e-3, (int(d.max()), float((d > 0).mean()))
    ref = dev(a1) * torch.nn.functional.silu(dev(a2))
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    # the plain GEMM over the same interleaved weights + the stand-alone gate kernel: same bits
    full = ops.w4a8_gemm(dev(qa), w4, dev(s_a), s_w, 2 * F, bias=b)
    assert torch.equal(ops.silu_mul_interleaved(full, F), got)


def test_gemm_gated_at_benchmark_shape(ops):
    """Qwen-VL's MLP at the benchmark's batch: M = 16384, K = 4096, F = 11008 through the default
    dispatcher (the fused 256 x 256 epilogue), against the un-fused chain on the device (plain GEMM,
    bit-exact vs the integer oracle in test_gemm_default_dispatch_at_benchmark_shapes, then torch's
    silu and product) bit for bit, and against the oracle on sampled rows."""
    from mquant_b200._lib import lib

    M, K, F = 16384, 4096, 11008
    assert lib().mq_w4a8_gemm_plan(M, 2 * F, K, lib().mq_w4a8_gemm_workspace_bytes(M, 2 * F, K)) == 4
    g = torch.Generator(device="cuda").manual_seed(385)
    qa = torch.randint(-128, 128, (M, K), dtype=torch.int8, device="cuda", generator=g)
    cw1 = torch.randint(-8, 8, (F, K), dtype=torch.int8, device="cuda", generator=g)
    cw2 = torch.randint(-8, 8, (F, K), dtype=torch.int8, device="cuda", generator=g)
    s_a = torch.rand(M, device="cuda", generator=g) * 0.05 + 0.01
    sw1 = torch.rand(F, device="cuda", generator=g) * 0.01 + 0.001
    sw2 = torch.rand(F, device="cuda", generator=g) * 0.01 + 0.001
    w4 = ops.pack_w4(ops.interleave_gated(cw1, cw2))
    got = ops.w4a8_gemm_gated(qa, w4, s_a, ops.interleave_gated(sw1, sw2), F)
    a1 = ops.w4a8_gemm(qa, ops.pack_w4(cw1), s_a, sw1, F)
    a2 = ops.w4a8_gemm(qa, ops.pack_w4(cw2), s_a, sw2, F)
    ref = a1 * torch.nn.functional.silu(a2)
    assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    rows = torch.from_numpy(np.random.default_rng(1).integers(0, M, 96)).cuda()
    o1 = oracle_gemm((host(qa[rows]), host(cw1), host(s_a[rows]), host(sw1), None, None, None))
    o2 = oracle_gemm((host(qa[rows]), host(cw2), host(s_a[rows]), host(sw2), None, None, None))
    d = _ulp16(host(got[rows]), O.silu_mul(o1, o2))
    assert d.max() <= 2 and (d > 0).mean() < 1e-3
