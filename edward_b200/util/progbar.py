"""Text progress bar with the output format of edward/util/progbar.py:12-115."""
from __future__ import annotations

import sys
import time


class Progbar(object):
  def __init__(self, target, width=30, interval=0.01, verbose=1):
    self.target = target
    self.width = width
    self.interval = interval
    self.verbose = verbose
    self.stored_values = {}
    self.start = time.time()
    self.last_update = 0
    self.total_width = 0
    self.seen_so_far = 0

  def update(self, current, values=None, force=False):
    """Print `current/target [pct%] bar ETA|Elapsed | name: value` (progbar.py:38-115): at most every
    `interval` seconds unless `force` or the target is reached."""
    for k, v in (values or {}).items():
      self.stored_values[k] = v
    self.seen_so_far = current
    now = time.time()
    if not force and (now - self.last_update) < self.interval and current < self.target:
      return
    self.last_update = now
    if self.verbose == 0:
      return
    prev_total_width = self.total_width
    out = sys.stdout
    out.write("\b" * prev_total_width)
    out.write("\r")
    n_digits = len(str(self.target))
    bar = "%*d/%*d" % (n_digits, current, n_digits, self.target)
    bar += " [{0}%] ".format(str(int(current / self.target * 100)).rjust(3))
    prog_width = int(self.width * float(current) / self.target)
    if prog_width > 0:
      try:
        block = "█" * prog_width
        block.encode(getattr(out, "encoding", None) or "utf-8")
      except (UnicodeEncodeError, LookupError):
        block = "*" * prog_width
      bar += block
    bar += " " * (self.width - prog_width)
    out.write(bar)
    time_per_unit = (now - self.start) / current if current else 0
    eta = time_per_unit * (self.target - current)
    info = " ETA: %ds" % eta if current < self.target else " Elapsed: %ds" % (now - self.start)
    for k, v in self.stored_values.items():
      info += " | {0:s}: {1:0.3f}".format(k, v)
    self.total_width = len(bar) + len(info)
    if prev_total_width > self.total_width:
      info += (prev_total_width - self.total_width) * " "
    out.write(info)
    out.flush()
    if current >= self.target:
      out.write("\n")
