from .coreset import CoresetQuery
from .strategy import Strategy


class CoresetSampler(CoresetQuery, Strategy):
    """Drop-in for /root/reference/src/query_strategies/coreset_sampler.py (K3 + K4)."""
