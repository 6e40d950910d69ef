"""
Observability of the in-box engine (SURVEY.md §5.1 / §5.5 — the reference has wall-clock prints and a tqdm bar only):

* ``StageTimer``   CUDA-event timing of the stages of a training step (gate, dispatch, expert forward, combine, backward
                   stages, optimizers) on the launching stream, plus NVTX ranges with the same names for ncu / nsys.
                   Disabled by default: ``mark()`` is a no-op, so the hot path pays nothing.
* ``MetricsLog``   structured per-step JSONL (samples/s, per-stage ms, exposed communication wait, tokens-per-expert
                   statistics, shadowed experts, failure-injection drops, loss).
"""
import json
import os
import time
from collections import OrderedDict
from typing import Dict, Optional

import torch


class StageTimer:
    def __init__(self, enabled: bool = False, nvtx: bool = False):
        self.enabled, self.nvtx = enabled, nvtx and torch.cuda.is_available()
        self._events = []      # (name, event) in program order; an interval is attributed to the mark that ENDS it
        self._open_range = False

    def start(self):
        """call at the beginning of a step"""
        if not self.enabled:
            return
        self._events = []
        self.mark("_start")

    def mark(self, name: str):
        if not self.enabled:
            return
        if self.nvtx:
            if self._open_range:
                torch.cuda.nvtx.range_pop()
            torch.cuda.nvtx.range_push(name)
            self._open_range = True
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self._events.append((name, ev))

    def report(self) -> Dict[str, float]:
        """ms per stage of the last step (synchronises); stages that occur several times (one per layer) are summed"""
        if not self.enabled or len(self._events) < 2:
            return {}
        if self.nvtx and self._open_range:
            torch.cuda.nvtx.range_pop()
            self._open_range = False
        self._events[-1][1].synchronize()
        out: "OrderedDict[str, float]" = OrderedDict()
        for (_, prev), (name, ev) in zip(self._events[:-1], self._events[1:]):
            out[name] = out.get(name, 0.0) + prev.elapsed_time(ev)
        out["total"] = self._events[0][1].elapsed_time(self._events[-1][1])
        return out


class MetricsLog:
    """append-only JSONL writer; one record per call"""

    def __init__(self, path: Optional[str]):
        self.path = path
        if path:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            self._f = open(path, "a")
        else:
            self._f = None

    def write(self, **record):
        record.setdefault("time", time.time())
        if self._f is not None:
            self._f.write(json.dumps(record, default=float) + "\n")
            self._f.flush()
        return record

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None
