"""CPU oracle for the TensoFlow hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package
(``tensoflow_b200``) never imports this package and fails loudly when its CUDA
library is missing.
"""
