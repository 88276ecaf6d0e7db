from .bsc import BSC

__all__ = ["BSC"]
