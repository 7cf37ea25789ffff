from .coreset import PartitionedBADGEQuery
from .strategy import Strategy


class PartitionedBADGESampler(PartitionedBADGEQuery, Strategy):
    """Drop-in for /root/reference/src/query_strategies/partitioned_badge_sampler.py (K2p + K3 + K5)."""
