"""ed.copy for the tiny symbolic layer (edward/util/random_variables.py:139-445): rebuilds an expression or
random variable with the nodes in `dict_swap` replaced — enough to form posterior predictives such as
`y_post = ed.copy(y, {w: qw, b: qb})` (docs/tex/api/criticism.tex) for the models of this path."""
from __future__ import annotations

from .. import graph as _g
from ..models.random_variable import RandomVariable
from ..models.random_variables import Bernoulli, Normal, Poisson


def copy(org_instance, dict_swap=None, scope="copied", replace_itself=False, copy_q=False):
  dict_swap = dict_swap or {}
  memo = {}

  def lookup(node):
    for k, v in dict_swap.items():
      if node is k:
        return _g.convert_to_tensor(v) if not isinstance(v, RandomVariable) else v
    return None

  def rec(node, top=False):
    if not isinstance(node, _g.Tensor):
      return node
    if not (top and not replace_itself):
      sw = lookup(node)
      if sw is not None:
        return sw
    if id(node) in memo:
      return memo[id(node)]
    if isinstance(node, (_g.Constant, _g.Placeholder, _g.Variable, _g.Lazy)):
      out = node
    elif isinstance(node, _g._Binary):
      out = type(node)(rec(node.a), rec(node.b))
    elif isinstance(node, _g.Dot):
      out = _g.Dot(rec(node.x), rec(node.y))
    elif isinstance(node, _g.Unary):
      out = _g.Unary(rec(node.a), node.fn, node.op_type)
    elif isinstance(node, _g.Stack):
      out = _g.Stack([rec(v) for v in node.values])
    elif isinstance(node, Normal):
      out = Normal(loc=rec(node.loc), scale=rec(node.scale), sample_shape=tuple(node.sample_shape))
    elif isinstance(node, Bernoulli):
      out = Bernoulli(logits=rec(node.logits)) if node.logits is not None else Bernoulli(probs=rec(node._probs))
    elif isinstance(node, Poisson):
      out = Poisson(log_rate=rec(node.log_rate)) if node.log_rate is not None else Poisson(rate=rec(node._rate))
    elif isinstance(node, RandomVariable):
      out = node  # e.g. Empirical: nothing inside to swap
    else:
      raise NotImplementedError("ed.copy: unsupported node %s" % type(node).__name__)
    memo[id(node)] = out
    return out

  return rec(org_instance, top=True)
