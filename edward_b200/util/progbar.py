"""Console progress line for `Inference.run` / `print_progress`.

The text a user sees is the reference's (edward/util/progbar.py): counter, percentage, a 30-cell bar, ETA while
running and the elapsed time at the end, then ` | name: value` per reported statistic. The implementation is split
into a pure formatter (`format_progress`, unit-testable without a terminal) and a small rate-limited writer.
"""
from __future__ import annotations

import sys
import time
from typing import Mapping, Optional


def _cells(n: int, stream) -> str:
  """n filled cells: a block glyph where the stream can encode it, '*' elsewhere."""
  glyph = "█"
  codec = getattr(stream, "encoding", None) or "utf-8"
  try:
    glyph.encode(codec)
  except (UnicodeEncodeError, LookupError):
    glyph = "*"
  return glyph * n


def format_progress(done: int, total: int, elapsed: float, stats: Mapping[str, float], width: int = 30, stream=None):
  """Returns (bar, tail): `  7/100 [  7%] ██            ` and ` ETA: 12s | Acceptance Rate: 0.912`."""
  digits = len(str(total))
  filled = int(width * float(done) / total)
  bar = "".join([
      str(done).rjust(digits), "/", str(total).rjust(digits),
      " [", str(int(done / total * 100)).rjust(3), "%] ",
      _cells(filled, stream if stream is not None else sys.stdout), " " * (width - filled),
  ])
  if done < total:
    per_unit = elapsed / done if done else 0.0
    tail = " ETA: %ds" % (per_unit * (total - done))
  else:
    tail = " Elapsed: %ds" % elapsed
  tail += "".join(" | %s: %0.3f" % (name, value) for name, value in stats.items())
  return bar, tail


class Progbar(object):
  """`Progbar(target).update(current, values={'Loss': x})`; `verbose=0` silences it."""

  def __init__(self, target, width=30, interval=0.01, verbose=1):
    self.target = target
    self.width = width
    self.interval = interval
    self.verbose = verbose
    self.stored_values = {}
    self.seen_so_far = 0
    self.total_width = 0
    self.start = time.time()
    self.last_update = 0

  def _due(self, now: float, current: int, force: bool) -> bool:
    return force or current >= self.target or (now - self.last_update) >= self.interval

  def update(self, current, values: Optional[Mapping[str, float]] = None, force=False):
    if values:
      self.stored_values.update(values)
    self.seen_so_far = current
    now = time.time()
    if not self._due(now, current, force):
      return
    self.last_update = now
    if not self.verbose:
      return
    stream = sys.stdout
    bar, tail = format_progress(current, self.target, now - self.start, self.stored_values, self.width, stream)
    shown = len(bar) + len(tail)
    pad = " " * max(0, self.total_width - shown)  # blank out the remains of a longer previous line
    stream.write("\b" * self.total_width + "\r" + bar + tail + pad)
    self.total_width = shown
    stream.flush()
    if current >= self.target:
      stream.write("\n")
