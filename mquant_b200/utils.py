"""The slice of fake_quant/utils.py the hot path touches."""
import random

import numpy as np
import torch

DEV = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")


def revise_down_input(m, i, new_size):
    """forward-pre-hook: right zero-pad the last dim (reference fake_quant/utils.py:465-471).
    QuantLinear folds the pad into the fused Hadamard kernel when it sees ``pad_to``; the hook is
    kept for drivers that register it themselves."""
    pad = new_size - i[0].shape[-1]
    if pad <= 0:
        return i
    return (torch.nn.functional.pad(i[0], (0, pad)),) + tuple(i[1:])


def seed_everything(seed=42):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    return seed


def cleanup_memory(verbos=True):
    import gc

    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
