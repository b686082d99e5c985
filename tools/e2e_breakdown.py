"""Where the end-to-end time of ed.HMC(...).run() at cfg 2 goes (host side): cProfile over 5 calls."""
import cProfile
import pstats
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import edward_b200 as ed
from edward_b200 import graph as g
from edward_b200 import tfshim as tf
from edward_b200.models import Bernoulli, Empirical, Normal

N, D, T, L = 581012, 54, 100, 10
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
X = torch.randn(N, D, device=dev, generator=gen)
y = (torch.rand(N, device=dev, generator=gen) < 0.5).to(torch.int32)
Xh, yh = X.cpu().pin_memory(), y.cpu().pin_memory()


def once():
  g.reset_default_graph()
  xs = tf.placeholder(tf.float32, [N, D])
  beta = Normal(loc=tf.zeros(D), scale=tf.ones(D))
  ys = Bernoulli(logits=ed.dot(xs, beta))
  qbeta = Empirical(params=tf.Variable(tf.zeros([T, D])))
  inference = ed.HMC({beta: qbeta}, data={xs: Xh, ys: yh})
  inference.run(step_size=0.5 / N, n_steps=L, n_print=0, device=dev)
  s = qbeta.params.eval()
  return int(inference.n_accept.eval())


once()
ts = []
for _ in range(8):
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  once()
  ts.append((time.perf_counter() - t0) * 1e3)
print("ms per call:", ["%.1f" % t for t in ts])
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
  once()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
