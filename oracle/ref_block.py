"""BASELINE.json config 1 on the REAL reference: one Qwen-VL decoder layer (``QWenBlock``,
model/modeling_qwen_opt.py:390-420; hidden 4096, ffn 11008, 32 heads), its 7 Linears wrapped by the
reference's ``add_actquant`` (fake_quant/quant_utils.py:626-662), RTN W4 per-channel weights
(gptq/qwenvl_gptq_plus.py:450-481), static INT8 minmax activations, online Hadamard K = 172 on
``mlp.c_proj`` (exam/quant_qwenvl.py:141-148).

TEST / BASELINE INFRASTRUCTURE: used by oracle/gen_golden.py (block fixture) and by bench.py's
reference arm / cpu_baseline leg -- never by the product.  Weights are drawn from per-layer seeded
generators so that a test can regenerate them without the reference."""
import torch

HIDDEN, FFN, HEADS = 4096, 11008, 32
LINEARS = [  # (name, K_in, N_out, bias) in the order the seeds are assigned
    ("attn.q_proj", HIDDEN, HIDDEN, True), ("attn.k_proj", HIDDEN, HIDDEN, True), ("attn.v_proj", HIDDEN, HIDDEN, True),
    ("attn.c_proj", HIDDEN, HIDDEN, False), ("mlp.w1", HIDDEN, FFN, False), ("mlp.w2", HIDDEN, FFN, False),
    ("mlp.c_proj", FFN, HIDDEN, False),
]
SEED0 = 4200


def linear_weights(i, K_in, N_out, bias, dtype=torch.float16):
    """nn.Linear-like init (uniform +-1/sqrt(K)) from generator seed SEED0 + i; CPU, deterministic."""
    g = torch.Generator().manual_seed(SEED0 + i)
    bound = K_in ** -0.5
    W = ((torch.rand(N_out, K_in, generator=g) * 2 - 1) * bound).to(dtype)
    b = ((torch.rand(N_out, generator=g) * 2 - 1) * bound).to(dtype) if bias else None
    return W, b


def hidden_states(seed, tokens, dtype=torch.float16):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, tokens, HIDDEN, generator=g)
    x[..., 7] *= 20  # outlier channels
    x[..., HIDDEN // 2 + 3] *= 12
    return x.to(dtype)


def build_reference_block(qu, hu, mq, dtype=torch.float16, seq_length=2048):
    """-> (block, wrappers {name: ActQuantWrapper}, rtn scales {name: tensor})."""
    cfg = mq.QWenConfig(hidden_size=HIDDEN, intermediate_size=2 * FFN, num_attention_heads=HEADS, kv_channels=128,
                        fp16=dtype == torch.float16, bf16=dtype == torch.bfloat16, seq_length=seq_length,
                        no_bias=True, use_flash_attn=False)
    block = mq.QWenBlock(cfg).to(dtype).eval()
    mods = dict(block.named_modules())
    for i, (name, K_in, N_out, bias) in enumerate(LINEARS):
        W, b = linear_weights(i, K_in, N_out, bias, dtype)
        lin = mods[name]
        assert tuple(lin.weight.shape) == (N_out, K_in) and (lin.bias is not None) == bias, name
        lin.weight.data = W
        if bias:
            lin.bias.data = b
    # offline rotation of mlp.c_proj so that the online Hadamard cancels (rotation_utils.py:447-451
    # -> hadamard_utils.apply_exact_had_to_linear :135-191, input side; its arithmetic without the
    # unconditional .cuda() so that it also runs on a CPU-only box)
    had_K, K = hu.get_hadK(FFN)
    lin = mods["mlp.c_proj"]
    lin.weight.data = hu.matmul_hadU_cuda(lin.weight.data.float(), had_K, K).to(dtype)
    qu.add_actquant(block)
    wrappers = qu.find_qlayers(block, layers=[qu.ActQuantWrapper])
    wrappers["mlp.c_proj"].online_full_had = True
    wrappers["mlp.c_proj"].had_K, wrappers["mlp.c_proj"].K = had_K, K
    scales = {}
    for name, w in wrappers.items():  # RTN W4 per-channel symmetric (qwenvl_llm_rtn)
        wq = qu.WeightQuantizer()
        wq.configure(4, perchannel=True, sym=True, mse=False)
        wq.find_params(w.module.weight.data)
        w.module.weight.data = wq.quantize(w.module.weight.data).to(dtype)
        scales[name] = wq.scale
        w.quantizer.configure(bits=8, groupsize=-1, sym=True, clip_ratio=1.0, static=True, observer_type="minmax",
                              calibration_mode="layer_wise")
    return block, wrappers, scales


def run_block(block, mq, x):
    """One forward of the block on [1, T, HIDDEN]: additive causal mask as QWenModel.forward builds
    it for this attention (modeling_qwen_opt.py:176-210); no rotary embedding (position encoding
    does not touch the quantized-linear path and keeps the harness independent of the model root)."""
    tokens = x.shape[1]
    mask = torch.full((tokens, tokens), torch.finfo(x.dtype).min, dtype=x.dtype).triu(1).view(1, 1, tokens, tokens)
    out = block(x, rotary_pos_emb=None, registered_causal_mask=None, attention_mask=mask)
    return out[0]


def set_flags(wrappers, **flags):
    for w in wrappers.values():
        for k, v in flags.items():
            setattr(w.quantizer, k, v)


def calibrate(block, wrappers, mq, batches):
    set_flags(wrappers, calibrate=True)
    for i, b in enumerate(batches):
        if i == len(batches) - 1:
            set_flags(wrappers, last_calibrate=True)
        run_block(block, mq, b)
    set_flags(wrappers, calibrate=False, quant=True)
