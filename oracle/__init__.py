"""CPU oracle for the I2VSGG region-level hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``i2vsgg_b200/``
does; the product path fails loudly when its CUDA library is missing.
"""
