"""What would a shared-memory-free GEMV pass cost? (edhmc_probe_read modes 2..5; development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from edward_b200 import _C
lib = _C.lib()
dev = torch.device("cuda:0")
ntiles = (581012 + 31) // 32
buf = torch.randn(ntiles * 6912 // 4, device=dev)
sink = torch.zeros(4, device=dev)
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
for mode, name in ((2, "8 warps, theta in registers"), (5, "8 warps, theta in smem"), (3, "12 warps, theta in smem"), (4, "16 warps, theta in smem")):
  for iters in (200,):
    _C.check(lib.edhmc_probe_read(buf.data_ptr(), buf.numel() * 4, 10, mode, sink.data_ptr(), None))
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(); _C.check(lib.edhmc_probe_read(buf.data_ptr(), buf.numel() * 4, iters, mode, sink.data_ptr(), None)); e1.record()
      torch.cuda.synchronize()
      best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    print("mode %d (%s): %.2f us per pass over 581012 x 54 = %.1f GB/s" % (mode, name, best, buf.numel() * 4 / best / 1e3))
