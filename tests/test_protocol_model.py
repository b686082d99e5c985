"""A host-side model of the two-slot inbox protocol the persistent kernel uses between row shards (csrc/chain.cuh,
peer_allreduce / wide_allreduce): R ranks run passes 1, 2, ... ; in pass s every rank writes its totals into entry
`rank` of slot s & 1 of EVERY rank's inbox, then waits until its own inbox holds pass s from all ranks, then reads.
The model executes random interleavings and checks the property the design relies on: no entry is overwritten
before every reader of the previous use of that slot has consumed it — i.e. two slots suffice, because a rank can
run at most one pass ahead of the slowest one."""
import random

import pytest


def simulate(n_ranks, n_passes, seed, slots=2):
  rng = random.Random(seed)
  # inbox[dst][slot][src] = (pass number written, value)
  inbox = [[[(0, None) for _ in range(n_ranks)] for _ in range(slots)] for _ in range(n_ranks)]
  consumed = [[[0 for _ in range(n_ranks)] for _ in range(slots)] for _ in range(n_ranks)]  # last pass read by dst
  # per-rank program counter: (pass, phase) with phase 0 = writing to destination k, 1 = waiting/reading
  state = [{"s": 1, "phase": 0, "k": 0, "sums": []} for _ in range(n_ranks)]
  done = 0
  steps = 0
  while done < n_ranks:
    steps += 1
    assert steps < 10_000_000, "deadlock"
    r = rng.randrange(n_ranks)
    st = state[r]
    if st["s"] > n_passes:
      continue
    s, slot = st["s"], st["s"] % slots
    if st["phase"] == 0:
      dst = st["k"]
      prev_pass, _ = inbox[dst][slot][r]
      # the entry being overwritten must have been consumed by its reader (or never used)
      assert prev_pass == 0 or consumed[dst][slot][r] == prev_pass, (
          "rank %d overwrites pass %d in rank %d's slot %d before it was read" % (r, prev_pass, dst, slot))
      inbox[dst][slot][r] = (s, (r + 1) * 1000 + s)
      st["k"] += 1
      if st["k"] == n_ranks:
        st["phase"], st["k"] = 1, 0
    else:
      if all(inbox[r][slot][src][0] >= s for src in range(n_ranks)):
        vals = []
        for src in range(n_ranks):
          p, v = inbox[r][slot][src]
          assert p == s, "rank %d reads pass %d where it expects %d" % (r, p, s)
          vals.append(v)
          consumed[r][slot][src] = s
        st["sums"].append(sum(vals))  # rank order: identical on every rank
        st["s"] += 1
        st["phase"] = 0
        if st["s"] > n_passes:
          done += 1
  sums = [tuple(st["sums"]) for st in state]
  assert all(x == sums[0] for x in sums)
  return sums[0]


@pytest.mark.parametrize("n_ranks", [2, 3, 8])
def test_two_slots_suffice_under_random_interleavings(n_ranks):
  for seed in range(40):
    sums = simulate(n_ranks, n_passes=12, seed=seed)
    assert len(sums) == 12


def test_one_slot_is_not_enough():
  """The same model with a single slot must trip the overwrite check for some interleaving: the check has teeth."""
  tripped = False
  for seed in range(200):
    try:
      simulate(3, n_passes=6, seed=seed, slots=1)
    except AssertionError:
      tripped = True
      break
  assert tripped
