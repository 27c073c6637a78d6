from .ctrl_aviary import BatchedCtrlAviary, CtrlAviary

__all__ = ["BatchedCtrlAviary", "CtrlAviary"]
