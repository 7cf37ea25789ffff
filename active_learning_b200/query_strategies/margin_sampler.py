from .strategy import Strategy
from .uncertainty import MarginQuery


class MarginSampler(MarginQuery, Strategy):
    """Drop-in for /root/reference/src/query_strategies/margin_sampler.py (K1 margin + K1b)."""
