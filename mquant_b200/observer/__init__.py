from .base import BaseObserver
from .build import build_observer
from .minmax import MinmaxObserver
from .percentile import PercentileObserver

__all__ = ["BaseObserver", "MinmaxObserver", "PercentileObserver", "build_observer"]
