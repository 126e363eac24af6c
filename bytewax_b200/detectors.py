"""Declared ``stateful_map`` mappers: plain callables with the host semantics of the reference's examples, which also
carry a plan the engine's recogniser can hand to the CUDA path (``run_main(..., gpu=True)`` / ``BYTEWAX_B200_GPU=1``).

``ZScoreDetector`` is the mapper of the reference's ``examples/anomaly_detector.py:16-48`` (BASELINE config C2) with its two
constants as parameters::

    labeled = op.stateful_map("detector", metrics, ZScoreDetector(window=10, threshold_z=2.0))
    # ("metric", (value, mu, sigma, is_anomalous))
"""
from dataclasses import dataclass, field
from typing import List, Optional

from bytewax_b200.operators import GpuSmapPlan


@dataclass
class DetectorState:
    """examples/anomaly_detector.py:16-37."""

    window: int = 10
    last: List[float] = field(default_factory=list)
    mu: Optional[float] = None
    sigma: Optional[float] = None

    def push(self, value):
        self.last.insert(0, value)
        del self.last[self.window:]
        n = len(self.last)
        self.mu = sum(self.last) / n
        self.sigma = (sum((v - self.mu) ** 2 for v in self.last) / n) ** 0.5

    def is_anomalous(self, value, threshold_z):
        if self.mu and self.sigma:
            return abs(value - self.mu) / self.sigma > threshold_z
        return False


class ZScoreDetector:
    """``mapper(state, value) -> (state, (value, mu, sigma, is_anomalous))`` (examples/anomaly_detector.py:40-48)."""

    def __init__(self, window: int = 10, threshold_z: float = 2.0):
        if not 1 <= window <= 32:
            raise ValueError("window must be in 1..32")
        self.window, self.threshold_z = window, threshold_z
        self._gpu_plan = GpuSmapPlan(window, float(threshold_z))

    def __call__(self, state, value):
        if state is None:
            state = DetectorState(self.window)
        is_anomalous = state.is_anomalous(value, self.threshold_z)
        state.push(value)
        # always return the state so that it is never discarded
        return (state, (value, state.mu, state.sigma, is_anomalous))
