"""build_observer (mirror of fake_quant/observer/build.py:8-19).  ema / omse / ptf are never
selected by any reference driver (observer_type="minmax" is hard-coded, exam/quant_qwenvl.py:221)
and are out of scope: asking for them fails loudly."""
from .minmax import MinmaxObserver
from .percentile import PercentileObserver

str2observer = {"minmax": MinmaxObserver, "percentile": PercentileObserver}


def build_observer(observer_str, module_type, bit_type, calibration_mode):
    if observer_str not in str2observer:
        raise NotImplementedError(f"observer '{observer_str}' is outside the mquant_b200 hot path")
    return str2observer[observer_str](module_type, bit_type, calibration_mode)
