from typing import Any
_PATH = Any
LRSchedulerTypeUnion = Any
