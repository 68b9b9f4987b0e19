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

# Linear inventories of the LLMs named by BASELINE.json configs 2-4.
# Each entry: (name, in_features, out_features, bias, online_hadamard, input buffer kind); an
# online-Hadamard input narrower than auto_pad_size(in) is zero padded (utils.revise_down_input,
# folded into the fused kernel through QuantLinear.pad_to).
def _qwenvl(hidden=4096, ffn=11008):
    return [
        ("attn.q_proj", hidden, hidden, True, False, "h"),
        ("attn.k_proj", hidden, hidden, True, False, "h"),
        ("attn.v_proj", hidden, hidden, True, False, "h"),
        ("attn.c_proj", hidden, hidden, False, False, "a"),
        ("mlp.w1", hidden, ffn, False, False, "h2"),
        ("mlp.w2", hidden, ffn, False, False, "h2"),
        ("mlp.c_proj", ffn, hidden, False, True, "f"),   # 11008 = 172 x 64
    ]


def _qwen2vl(hidden=3584, ffn=18944, kv=512):
    return [
        ("self_attn.q_proj", hidden, hidden, True, False, "h"),
        ("self_attn.k_proj", hidden, kv, True, False, "h"),
        ("self_attn.v_proj", hidden, kv, True, False, "h"),
        ("self_attn.o_proj", hidden, hidden, False, False, "a"),
        ("mlp.gate_proj", hidden, ffn, False, False, "h2"),
        ("mlp.up_proj", hidden, ffn, False, False, "h2"),
        ("mlp.down_proj", ffn, hidden, False, True, "f"),  # 18944 -> padded 19968 = 156 x 128
    ]


def _internvl2(hidden=4096, ffn=14336, qkv=6144):
    return [
        ("attention.wqkv", hidden, qkv, False, False, "h"),
        ("attention.wo", hidden, hidden, False, False, "a"),
        ("feed_forward.w1", hidden, ffn, False, False, "h2"),
        ("feed_forward.w3", hidden, ffn, False, False, "h2"),
        ("feed_forward.w2", ffn, hidden, False, True, "f"),  # 14336 = 28 x 512
    ]


MODELS = {
    # name: (hidden, ffn, n_layers, linears)
    "qwenvl-9.6b": (4096, 11008, 32, _qwenvl()),
    "qwen2vl-7b": (3584, 18944, 28, _qwen2vl()),
    "internvl2-8b": (4096, 14336, 32, _internvl2()),
}


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
    spans = None
    if visual_span == "dynamic":
        # Qwen2-VL dynamic resolution: (H/28)*(W/28) visual tokens per image, drawn per sequence
        gs = torch.Generator().manual_seed(seed)
        n_seq = (n_tokens + seq_len - 1) // seq_len
        spans = torch.randint(64, max(65, seq_len // 2), (n_seq,), generator=gs).tolist()
    for i, s in enumerate(range(0, n_tokens, seq_len)):
        a = s + 1
        span = spans[i] if spans is not None else visual_span
        mask[a:min(a + span, n_tokens, s + seq_len)] = 1
    return x.half(), mask


class QuantLinearStack(torch.nn.Module):
    def __init__(self, model="qwenvl-9.6b", n_layers=None, device="cuda", seed=42, calib_tokens=2048,
                 split=False, keep_fp16_weights=False, calibrate=True, fuse_siblings=False, skeleton=False):
        """calibrate=False leaves the wrappers configured (static INT8, MSQ, minmax observers) but
        neither calibrated nor frozen -- the starting point of a (sharded) calibration run.
        skeleton=True (with calibrate=False) builds the wrappers around 8-channel Linears: the
        calibration hot path (rotation + observers) never touches the weights.
        fuse_siblings=True lets the Linears of a block that read the same tensor (q/k/v, the two
        up projections) share one activation quantisation and one GEMM launch (bit-identical)."""
        super().__init__()
        hidden, ffn, L, spec = MODELS[model]
        self.model = model
        self.hidden, self.ffn = hidden, ffn
        self.n_layers = L if n_layers is None else n_layers
        self.spec = spec
        self.layers = torch.nn.ModuleList()
        torch.manual_seed(seed)
        args = _Args()
        assert not (skeleton and calibrate), "a skeleton stack has no weights to freeze"
        calib = {} if not calibrate else {
            "h": synthetic_tokens(calib_tokens, hidden, device, 1)[0],
            "a": synthetic_tokens(calib_tokens, hidden, device, 2)[0],
            "h2": synthetic_tokens(calib_tokens, hidden, device, 3)[0],
            "f": synthetic_tokens(calib_tokens, ffn, device, 4)[0],
        }
        cmask = synthetic_tokens(calib_tokens, 8, device, 0)[1]
        for _ in range(self.n_layers):
            block = torch.nn.ModuleDict()
            for name, k_in, n_out, bias, had, kind in self.spec:
                k_pad = hadamard_utils.auto_pad_size(k_in) if had else k_in
                if skeleton:
                    w = quant_utils.ActQuantWrapper(torch.nn.Linear(k_pad, 8, bias=False).half().to(device))
                    if had:
                        w.online_full_had = True
                        w.had_K, w.K = hadamard_utils.get_hadK(k_pad)
                        if k_pad != k_in:
                            w.pad_to = k_pad
                    w.quantizer.configure(bits=8, sym=True, static=True, observer_type="minmax", msq=True)
                    block[name.replace(".", "_")] = w
                    continue
                lin = torch.nn.Linear(k_pad, n_out, bias=bias).half().to(device)
                if k_pad != k_in:
                    lin.weight.data[:, k_in:] = 0  # the padded input columns are zero anyway
                w = quant_utils.ActQuantWrapper(lin)
                if had:
                    # offline: rotate the weight so that the online transform cancels
                    hadamard_utils.apply_exact_had_to_linear(lin, had_dim=-1, output=False)
                    w.online_full_had = True
                    w.had_K, w.K = hadamard_utils.get_hadK(k_pad)
                    if k_pad != k_in:
                        w.pad_to = k_pad
                    if split:
                        w.split = True
                        w.split_weights()
                tgt = w.L2 if (had and split) else w.module
                wq = quant_utils.WeightQuantizer()
                wq.configure(4, perchannel=True, sym=True, mse=False)  # --llm_w_rtn
                wq.find_params(tgt.weight.data)
                tgt.weight.data = wq.quantize(tgt.weight.data).to(torch.float16)
                w.quantizer.configure(bits=8, sym=True, static=True, observer_type="minmax", msq=True)
                w._wq_scale = wq.scale
                if not calibrate:
                    block[name.replace(".", "_")] = w
                    continue
                with msq.modality_context(cmask):
                    quant_utils.calibrate_static(w, args, [calib[kind]], forward=lambda m, b: m(b))
                w.freeze(wq.scale)
                if not keep_fp16_weights:
                    # the integer path owns the weights now; drop the fp16 copy (13 GB at 32 layers)
                    for lin_ in ([w.module] + ([w.L1, w.L2] if w.split else [])):
                        lin_.weight.data = torch.empty(0, dtype=torch.float16, device=device)
                block[name.replace(".", "_")] = w
            if fuse_siblings and calibrate:
                by_input = {}
                for name, _, _, _, had, kind in self.spec:
                    if not had:
                        by_input.setdefault(kind, []).append(block[name.replace(".", "_")])
                for group in by_input.values():
                    quant_utils.fuse_siblings(group)
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

    @torch.no_grad()
    def capture_tokens(self, inputs, warmup=2):
        """The same pass as ``forward_tokens(inputs)`` captured into a CUDA graph (one request's launches --
        8 kernels per layer -- replayed without host work between them; at 1024 tokens the eager loop is
        close to launch bound).  ``inputs`` are static buffers: refill them in place, ``graph.replay()``,
        read ``out``.  Returns (graph, out)."""
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=cur.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # warm-up on the capture stream: workspaces / attributes exist before capture
            for _ in range(warmup):
                self.forward_tokens(inputs)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            out = self.forward_tokens(inputs)
        return graph, out

    @torch.no_grad()
    def observe_tokens(self, inputs):
        """Calibration pass of the hot path only: online rotation (fp16 output) + observer update of
        every wrapper, without the fp16 Linear the reference's calibration forward also runs (that
        GEMM is the caller's model compute, not the path).  Wrappers must be in calibrate mode."""
        with msq.modality_context(inputs["mask"]):
            for block in self.layers:
                for name, _, _, _, _, kind in self.spec:
                    w = block[name.replace(".", "_")]
                    w.quantizer(w._rotate(inputs[kind]))

    def describe(self):
        had = [(n, k) for n, k, _, _, h, _ in self.spec if h][0]
        kp = hadamard_utils.auto_pad_size(had[1])
        return (f"{self.model} LLM prefill, quantized-linear path ({len(self.spec)} QuantLinear x {self.n_layers} layers, "
                f"W4A8, MSQ, online Hadamard on {had[0]}: {had[1]}"
                + (f" padded to {kp}" if kp != had[1] else "") + f", K={hadamard_utils.get_hadK(kp)[1]})")

    def gemm_ops_per_token(self):
        return sum(2 * k * n for _, k, n, _, _, _ in self.spec) * self.n_layers

    def gemm_bytes_per_step(self, n_tokens):
        """Algorithmic bytes of the step's GEMM launches: N*K/2 + M*K + 2*M*N + 4*N."""
        tot = 0
        for _, k, n, _, had, _ in self.spec:
            kp = hadamard_utils.auto_pad_size(k) if had else k
            tot += n * kp // 2 + n_tokens * kp + 2 * n_tokens * n + 4 * n
        return tot * self.n_layers

    def launches_per_step(self):
        # per Linear (or per fused sibling group): one activation kernel (quant or fused
        # Hadamard+quant) + one GEMM
        units = 0
        seen = set()
        for name, *_ in self.spec:
            w = self.layers[0][name.replace(".", "_")]
            g = getattr(w, "_group", None)
            if g is None:
                units += 1
            elif id(g) not in seen:
                seen.add(id(g))
                units += 1
        return 2 * units * self.n_layers
