from .atst import ATST
__all__ = ['ATST']
