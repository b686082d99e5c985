"""Measures the dense TF32 GEMM peak of this GPU with the protocol of MEASURED_PEAKS.json (torch.matmul 8192^3,
best of 10 = burst; back to back for 4 s = sustained)."""
import json, time, torch
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
for _ in range(3): a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
  best = min(best, e0.elapsed_time(e1))
t0 = time.time(); cnt = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
  for _ in range(10): a @ b
  cnt += 10
  torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / cnt
fl = 2.0 * n ** 3
print(json.dumps({"tf32_tflops": fl / best / 1e9, "tf32_tflops_sustained": fl / sus / 1e9, "how": "torch.matmul fp32 inputs, allow_tf32, 8192^3"}))
