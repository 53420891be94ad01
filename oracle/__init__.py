"""ORACLE — test infrastructure, NOT part of the product.

CPU (PyTorch fp32) restatement of the reference algorithm for the WVN hot path, used only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs as the checker.
The product package ``wild_visual_navigation_b200`` never imports it and has no CPU fallback.
"""
