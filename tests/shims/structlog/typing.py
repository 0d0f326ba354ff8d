from typing import Any, Callable

Processor = Callable[..., Any]
