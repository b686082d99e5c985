"""Measures the read throughput of an L2-resident buffer and of an HBM-sized one with edhmc_probe_read (LDG.128 and TMA
bulk-copy paths), best of `reps` launches timed with CUDA events. Writes a JSON record (profiles/r02_read_peaks.json)."""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from edward_b200 import _C


def measure(nbytes, mode, iters, reps=10):
  lib = _C.lib()
  dev = torch.device("cuda:0")
  buf = torch.ones(nbytes // 4, dtype=torch.float32, device=dev)
  sink = torch.zeros(4, dtype=torch.float32, device=dev)
  st = torch.cuda.current_stream(dev).cuda_stream
  _C.check(lib.edhmc_probe_read(buf.data_ptr(), nbytes, 2, mode, sink.data_ptr(), st))  # warm-up (fills L2)
  torch.cuda.synchronize()
  best = None
  for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _C.check(lib.edhmc_probe_read(buf.data_ptr(), nbytes, iters, mode, sink.data_ptr(), st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = ms if best is None else min(best, ms)
  return nbytes * iters / best / 1e6  # GB/s


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--out", default=None)
  a = ap.parse_args()
  rec = {"when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "gpu": torch.cuda.get_device_name(0),
         "how": "edhmc_probe_read: `iters` read sweeps per launch, best of 10 launches, CUDA events; L2 sizes are re-read "
                "in place after a warm-up sweep, the HBM size (4 GiB) is 32x the L2"}
  for mb in (32, 64, 100, 120):
    n = mb * 1000 * 1000 // 32768 * 32768
    rec["l2_%dMB_ldg_gbs" % mb] = measure(n, 0, 200)
    rec["l2_%dMB_tma_gbs" % mb] = measure(n, 1, 200)
  n = 4 * 1024 ** 3
  rec["hbm_read_ldg_gbs"] = measure(n, 0, 5)
  rec["hbm_read_tma_gbs"] = measure(n, 1, 5)
  print(json.dumps(rec, indent=1))
  if a.out:
    with open(a.out, "w") as f:
      json.dump(rec, f, indent=1)


if __name__ == "__main__":
  main()
