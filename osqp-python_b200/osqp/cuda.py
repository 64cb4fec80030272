"""osqp.cuda.OSQP == OSQP(algebra='cuda')  (reference: src/osqp/cuda.py:1-6)."""
from osqp.interface import OSQP as _Base


class OSQP(_Base):
    def __init__(self, *args, **kwargs):
        kwargs["algebra"] = "cuda"
        super().__init__(*args, **kwargs)
