"""The quantized-linear path of an MLLM prefill, assembled through the drop-in API.

BASELINE.json configs[1]: "Qwen-VL-Chat 9.6B full W4A8 prefill with MSQ visual/text scales and
online Hadamard".  The LLM has 32 blocks (model/configuration_qwen.py:15-36); every block calls
seven wrapped Linears (model/modeling_qwen_opt.py:306-308,361,383-386):

    attn.q_proj/k_proj/v_proj  4096 -> 4096 (bias)     attn.c_proj  4096 -> 4096
    mlp.w1 / mlp.w2            4096 -> 11008           mlp.c_proj   11008 -> 4096, online
                                                       Hadamard K = 172 (exam/quant_qwenvl.py:141-148)

This module builds exactly those 7 x n_layers ``QuantLinear`` objects the way the reference driver
does (add wrapper -> set online-Hadamard attributes -> RTN W4 through WeightQuantizer -> configure
static INT8 -> calibrate -> quant) and then ``freeze()``s them.  ``forward_tokens`` runs one pass
of all of them over a token batch: that is the bench's "step".  Attention, RMSNorm and SiLU are
NOT part of the hot path and are not executed; each Linear kind reads its
own resident synthetic activation tensor so no work of the path is skipped or cached.
"""
from __future__ import annotations

import torch

from . import hadamard_utils, msq, quant_utils

SHAPES = {
    # name: (hidden, ffn, n_layers, qkv_bias)   -- SURVEY.md appendix B
    "qwenvl-9.6b": (4096, 11008, 32, True),
}

# (name, in, out, bias, online_hadamard, input buffer kind)
def _linears(hidden, ffn, qkv_bias):
    return [
        ("attn.q_proj", hidden, hidden, qkv_bias, False, "h"),
        ("attn.k_proj", hidden, hidden, qkv_bias, False, "h"),
        ("attn.v_proj", hidden, hidden, qkv_bias, False, "h"),
        ("attn.c_proj", hidden, hidden, False, False, "a"),
        ("mlp.w1", hidden, ffn, False, False, "h2"),
        ("mlp.w2", hidden, ffn, False, False, "h2"),
        ("mlp.c_proj", ffn, hidden, False, True, "f"),
    ]


class _Args:
    skip_names: list = []


def synthetic_tokens(n_tokens, width, device, seed=0, visual_span=256, seq_len=1024):
    """fp16 randn*3 with injected outlier channels, and a modality mask with one contiguous span
    of `visual_span` visual tokens per `seq_len`-token sequence."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.randn(n_tokens, width, device=device, generator=g, dtype=torch.float32) * 3
    idx = torch.arange(7, width, 997, device=device)
    x[:, idx] *= 20
    mask = torch.zeros(n_tokens, dtype=torch.uint8, device=device)
    for s in range(0, n_tokens, seq_len):
        a = s + 1
        mask[a:min(a + visual_span, n_tokens, s + seq_len)] = 1
    return x.half(), mask


class QuantLinearStack(torch.nn.Module):
    def __init__(self, model="qwenvl-9.6b", n_layers=None, device="cuda", seed=42, calib_tokens=2048,
                 split=False, keep_fp16_weights=False):
        super().__init__()
        hidden, ffn, L, qkv_bias = SHAPES[model]
        self.hidden, self.ffn = hidden, ffn
        self.n_layers = L if n_layers is None else n_layers
        self.spec = _linears(hidden, ffn, qkv_bias)
        self.layers = torch.nn.ModuleList()
        torch.manual_seed(seed)
        args = _Args()
        calib = {
            "h": synthetic_tokens(calib_tokens, hidden, device, 1)[0],
            "a": synthetic_tokens(calib_tokens, hidden, device, 2)[0],
            "h2": synthetic_tokens(calib_tokens, hidden, device, 3)[0],
            "f": synthetic_tokens(calib_tokens, ffn, device, 4)[0],
        }
        cmask = synthetic_tokens(calib_tokens, 8, device, 0)[1]
        for _ in range(self.n_layers):
            block = torch.nn.ModuleDict()
            for name, k_in, n_out, bias, had, kind in self.spec:
                lin = torch.nn.Linear(k_in, n_out, bias=bias).half().to(device)
                w = quant_utils.ActQuantWrapper(lin)
                if had:
                    # offline: rotate the weight so that the online transform cancels
                    hadamard_utils.apply_exact_had_to_linear(lin, had_dim=-1, output=False)
                    w.online_full_had = True
                    w.had_K, w.K = hadamard_utils.get_hadK(k_in)
                    if split:
                        w.split = True
                        w.split_weights()
                tgt = w.L2 if (had and split) else w.module
                wq = quant_utils.WeightQuantizer()
                wq.configure(4, perchannel=True, sym=True, mse=False)  # --llm_w_rtn
                wq.find_params(tgt.weight.data)
                tgt.weight.data = wq.quantize(tgt.weight.data).to(torch.float16)
                w.quantizer.configure(bits=8, sym=True, static=True, observer_type="minmax", msq=True)
                with msq.modality_context(cmask):
                    quant_utils.calibrate_static(w, args, [calib[kind]], forward=lambda m, b: m(b))
                w.freeze(wq.scale)
                if not keep_fp16_weights:
                    # the integer path owns the weights now; drop the fp16 copy (13 GB at 32 layers)
                    for lin_ in ([w.module] + ([w.L1, w.L2] if w.split else [])):
                        lin_.weight.data = torch.empty(0, dtype=torch.float16, device=device)
                block[name.replace(".", "_")] = w
            self.layers.append(block)

    def make_inputs(self, n_tokens, device="cuda", seed=0, visual_span=256, seq_len=1024):
        h, mask = synthetic_tokens(n_tokens, self.hidden, device, seed, visual_span, seq_len)
        return {
            "h": h,
            "a": synthetic_tokens(n_tokens, self.hidden, device, seed + 1)[0],
            "h2": synthetic_tokens(n_tokens, self.hidden, device, seed + 2)[0],
            "f": synthetic_tokens(n_tokens, self.ffn, device, seed + 3)[0],
            "mask": mask,
        }

    @torch.no_grad()
    def forward_tokens(self, inputs):
        """One pass of the hot path: 7 x n_layers QuantLinear forwards.  Returns the last output."""
        out = None
        with msq.modality_context(inputs["mask"]):
            for block in self.layers:
                for name, _, _, _, _, kind in self.spec:
                    out = block[name.replace(".", "_")](inputs[kind])
        return out

    def gemm_ops_per_token(self):
        return sum(2 * k * n for _, k, n, _, _, _ in self.spec) * self.n_layers

    def launches_per_step(self):
        # per Linear: one activation kernel (quant or fused Hadamard+quant) + one GEMM
        return 2 * len(self.spec) * self.n_layers
