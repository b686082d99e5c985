"""Runs the reference's OWN examples/bayesian_logistic_regression.py — the file as shipped, not a rewrite — on this
package: `edward` -> edward_b200, `edward.models` -> edward_b200.models, `tensorflow` -> edward_b200.tfshim are provided
as module aliases and matplotlib (absent from this image) as a no-op stub; the script text is executed unchanged.

The script is looked up at /root/reference/examples/ (this container) or at baseline/_ref/examples/ (a git-ignored copy
staged by __graft_entry__.build() so that it travels to the GPU box; nothing of it is tracked in this repository).

    python tools/run_reference_example.py [--T 5000]        # prints one JSON line: transitions/s of the update() loop
"""
from __future__ import annotations

import json
import os
import runpy
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
CANDIDATES = ["/root/reference/examples/bayesian_logistic_regression.py",
              os.path.join(ROOT, "baseline", "_ref", "examples", "bayesian_logistic_regression.py")]


def find_script():
  for p in CANDIDATES:
    if os.path.exists(p):
      return p
  return None


def run(argv=(), script=None):
  """Executes the script as __main__ with the module aliases in place. Returns a dict with the wall time of main(),
  the HMC objects it created, and transitions/s."""
  script = script or find_script()
  if script is None:
    raise FileNotFoundError("the reference example is not available (neither /root/reference nor baseline/_ref)")
  import edward_b200 as ed
  import edward_b200.models as ed_models
  from edward_b200 import graph as g
  from edward_b200 import tfshim
  g.reset_default_graph()
  created = []

  class RecordingHMC(ed.HMC):
    def __init__(self, *a, **k):
      super(RecordingHMC, self).__init__(*a, **k)
      created.append(self)

  class _NoOp(object):
    """matplotlib is absent from this image: every attribute is a no-op callable that returns the stub and counts its
    calls (a MagicMock records every call with its arguments and cost 20 % of the example's loop)."""
    def __init__(self):
      self.calls = {}
    def __getattr__(self, name):
      stub, calls = self, self.calls
      def fn(*a, **k):
        calls[name] = calls.get(name, 0) + 1
        return stub
      return fn
    def __iter__(self):
      return iter(())
  plt = _NoOp()
  mpl = types.ModuleType("matplotlib")
  mpl.pyplot = plt
  aliases = {"edward": ed, "edward.models": ed_models, "tensorflow": tfshim, "matplotlib": mpl, "matplotlib.pyplot": plt}
  saved = {k: sys.modules.get(k) for k in aliases}
  saved_argv = sys.argv
  old_hmc = ed.HMC
  sys.modules.update(aliases)
  ed.HMC = RecordingHMC
  sys.argv = [script] + list(argv)
  tfshim.flags.FLAGS.__init__()  # fresh flag registry for every run
  t0 = time.perf_counter()
  try:
    try:
      runpy.run_path(script, run_name="__main__")
    except SystemExit as e:
      if e.code not in (None, 0):
        raise
  finally:
    dt = time.perf_counter() - t0
    ed.HMC = old_hmc
    sys.argv = saved_argv
    for k, v in saved.items():
      if v is None:
        sys.modules.pop(k, None)
      else:
        sys.modules[k] = v
  inf = created[-1] if created else None
  out = {"script": script, "seconds": dt, "inferences": created}
  if inf is not None:
    out.update(t=int(inf.t.eval()), n_iter=int(inf.n_iter), n_accept=int(inf.n_accept.eval()),
               transitions_per_s=int(inf.t.eval()) / dt, plot_calls=plt.calls.get("draw", 0))
  return out


if __name__ == "__main__":
  r = run(sys.argv[1:])
  r.pop("inferences")
  print(json.dumps(r))
