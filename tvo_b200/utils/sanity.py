"""Parameter sanity checks (mirror of tvo/utils/sanity.py:14-86).

The reference fixes theta on rank 0 and broadcasts.  Here every rank holds bit-identical
statistics after the packed all-reduce and applies the same deterministic fix, so no broadcast is
needed; the operations stay on the device and never synchronise with the host.
"""
from typing import Dict, List, Union

import torch as to
from torch import Tensor


def fix_infinite(values: Tensor, replacement: Union[float, Tensor]) -> Tensor:
    bad = to.isnan(values) | to.isinf(values)
    if isinstance(replacement, Tensor):
        return to.where(bad, replacement.expand_as(values), values)
    return to.where(bad, to.full_like(values, replacement), values)


def fix_bounds(values: Tensor, lower=None, upper=None) -> Tensor:
    if lower is not None:
        values = to.maximum(values, lower) if isinstance(lower, Tensor) else to.clamp(values, min=lower)
    if upper is not None:
        values = to.minimum(values, upper) if isinstance(upper, Tensor) else to.clamp(values, max=upper)
    return values


def fix_theta(theta: Dict[str, Tensor], policy: Dict[str, List]):
    """Replace non-finite entries by policy[key][0], clamp to [policy[key][1], policy[key][2]]."""
    assert set(theta.keys()) == set(policy.keys()), "theta and policy must have same keys"
    for key, (replacement, low, up) in policy.items():
        theta[key] = fix_bounds(fix_infinite(theta[key], replacement), low, up)
