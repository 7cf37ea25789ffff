"""Binding the accelerated query() implementations onto the REFERENCE's own Strategy base class.

The classes in ``active_learning_b200.query_strategies`` derive from a query-path-only Strategy.
A maintainer of the reference who wants to keep its training loop, checkpointing and Comet
logging and only swap the query tail does:

    import query_strategies.strategy as ref                       # the reference's module
    from active_learning_b200.integration import make_drop_in
    ACCEL = make_drop_in(ref.Strategy)
    # in query_strategies/get_strategy.py:   def get_strategy(name): return ACCEL.get(name) or eval(name)

See INTEGRATION.md.
"""
from __future__ import annotations

from .query_strategies.coreset import (BADGEQuery, CoresetQuery, PartitionedBADGEQuery,
                                       PartitionedCoresetQuery)
from .query_strategies.balancing import BalancingQuery
from .query_strategies.mase import BASEQuery, MASEQuery
from .query_strategies.uncertainty import ConfidenceQuery, EntropyQuery, MarginQuery

_MIXINS = {
    "MarginSampler": MarginQuery,
    "ConfidenceSampler": ConfidenceQuery,
    "EntropySampler": EntropyQuery,
    "CoresetSampler": CoresetQuery,
    "PartitionedCoresetSampler": PartitionedCoresetQuery,
    "BADGESampler": BADGEQuery,
    "PartitionedBADGESampler": PartitionedBADGEQuery,
    "MASESampler": MASEQuery,
    "BASESampler": BASEQuery,
    "BalancingSampler": BalancingQuery,
}


def make_drop_in(strategy_base):
    """{class name: class(mixin, strategy_base)} for every accelerated sampler."""
    return {name: type(name, (mixin, strategy_base), {"__doc__": mixin.__doc__})
            for name, mixin in _MIXINS.items()}
