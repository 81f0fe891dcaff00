"""Opt-in device timeline: ``mark(label)`` records a CUDA event on the current stream when tracing is on.

``NUMS_TRACE=1`` (or ``trace.enable()``) turns it on; ``report()`` returns ``[(label, ms since the first mark,
ms since the previous mark)]`` after a device synchronisation.  Used to find where a step's time goes
(exchange barriers, staging copies, each grouped launch) without a profiler; costs nothing when off.
"""
import os

ENABLED = bool(int(os.environ.get("NUMS_TRACE", "0")))
_marks = []


def enable(on=True):
    global ENABLED
    ENABLED = bool(on)
    _marks.clear()


def mark(label, stream=None):
    if not ENABLED:
        return
    import time
    import torch
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(stream if stream is not None else torch.cuda.current_stream())
    _marks.append((label, ev, time.perf_counter()))


def report(clear=True):
    import torch
    torch.cuda.synchronize()
    out = []
    if _marks:
        first, prev = _marks[0][1], _marks[0][1]
        host0 = _marks[0][2]
        for label, ev, host in _marks:
            out.append((label, first.elapsed_time(ev), prev.elapsed_time(ev), (host - host0) * 1e3))
            prev = ev
    if clear:
        _marks.clear()
    return out
