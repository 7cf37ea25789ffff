from .coreset import PartitionedCoresetQuery
from .strategy import Strategy


class PartitionedCoresetSampler(PartitionedCoresetQuery, Strategy):
    """Drop-in for /root/reference/src/query_strategies/partitioned_coreset_sampler.py."""
