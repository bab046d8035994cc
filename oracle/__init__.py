"""CPU oracles for the SaSPA augmentation-generation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
reported CPU baseline.  The product package (``saspa_aug_b200``) never imports
this package and fails loudly when its CUDA library is missing.
"""
