from .mase import MASEQuery
from .strategy import Strategy


class MASESampler(MASEQuery, Strategy):
    """Drop-in for /root/reference/src/query_strategies/mase_sampler.py (K6 + K1b)."""
