from .Trainer import Trainer

__all__ = ["Trainer"]
