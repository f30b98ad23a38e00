"""Parity oracle for the GAPIT hot path -- TEST INFRASTRUCTURE ONLY.

Importable from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs; never from ``gapit_b200`` (the product path).
"""
