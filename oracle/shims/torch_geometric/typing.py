from typing import Any, Optional, Tuple
Adj = Any
Size = Optional[Tuple[int, int]]
SparseTensor = Any
