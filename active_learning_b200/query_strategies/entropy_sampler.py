from .strategy import Strategy
from .uncertainty import EntropyQuery


class EntropySampler(EntropyQuery, Strategy):
    """New sampler named by BASELINE.json (no reference counterpart): K1 entropy + K1b."""
