"""Drop-in for ``planner/utils/timers.py`` (:5-113): ``ExecTimer`` with hierarchical ``/``-separated labels.

Same public surface as the reference's class -- ``start_timer / stop_timer / time_with_cm / time_with_dec / reset /
get_timing_dict``, the ``running`` / ``finished`` dicts, a disabled timer is a no-op -- so ``exec_timing.json`` and the
timing plots of the reference's evaluation tools read the result unchanged.  One addition: ``record(label, seconds)``
files a duration measured elsewhere under a label; the fused path uses it for CUDA-event times of its kernels
(``etp_get_stage_timing``), because host wall clock around an asynchronous launch measures nothing.
"""
from __future__ import annotations

import functools
import time


class _Section:
    def __init__(self, timer, label):
        self._timer, self._label = timer, label

    def __enter__(self):
        self._timer.start_timer(self._label)

    def __exit__(self, exc_type, exc_value, exc_tb):
        self._timer.stop_timer(self._label)


class ExecTimer:
    """Times code sections by label; every label collects a list of durations in seconds."""

    TimerCM = _Section  # the reference exposes its context manager class under this name

    def __init__(self, timing_enabled=True):
        self.running = {}
        self.finished = {}
        self._timing_enabled = timing_enabled

    @property
    def enabled(self):
        return self._timing_enabled

    def reset(self):
        self.running = {}
        self.finished = {}

    def get_timing_dict(self):
        return self.finished

    def start_timer(self, label):
        if self._timing_enabled:
            self.running[label] = time.time()

    def stop_timer(self, label):
        if self._timing_enabled:
            self.record(label, time.time() - self.running.pop(label))

    def record(self, label, seconds):
        """File a duration measured elsewhere (CUDA events) under ``label``."""
        if self._timing_enabled:
            self.finished.setdefault(label, []).append(seconds)

    def time_with_cm(self, label):
        return _Section(self, label)

    def time_with_dec(self, label):
        def decorate(fn):
            @functools.wraps(fn)
            def timed(*args, **kwargs):
                with _Section(self, label):
                    return fn(*args, **kwargs)

            return timed

        return decorate


def timer_or_noop(exec_timer):
    """``ExecTimer(timing_enabled=False) if exec_timer is None else exec_timer`` (frenet_functions.py:236-238)."""
    return ExecTimer(timing_enabled=False) if exec_timer is None else exec_timer


class device_stages:
    """Context manager around fused launches on ``ctx``: switches the CUDA-event timing of the context on for the
    duration and files the stages of the last launch under ``prefix`` + "device/...".  No-op for a disabled timer."""

    def __init__(self, ctx, timer, prefix):
        self.ctx, self.timer, self.prefix = ctx, timer, prefix
        self.on = bool(getattr(timer, "_timing_enabled", False))

    def __enter__(self):
        if self.on:
            self.ctx.enable_timing(True)
        return self

    def __exit__(self, exc_type, exc_value, exc_tb):
        if self.on:
            try:
                if exc_type is None:
                    prof, fused, select = self.ctx.stage_timing()
                    rec = getattr(self.timer, "record", None)
                    if rec is None:  # a reference ExecTimer: same dict, same list-per-label convention
                        rec = lambda k, v: self.timer.finished.setdefault(k, []).append(v)  # noqa: E731
                    rec(self.prefix + "device/profile stage", prof * 1e-3)
                    rec(self.prefix + "device/fused scoring kernel", fused * 1e-3)
                    rec(self.prefix + "device/selection", select * 1e-3)
            finally:
                self.ctx.enable_timing(False)
        return False
