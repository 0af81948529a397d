"""TEST INFRASTRUCTURE: parity oracles for the ANN_Force hot path.

* ``oracle.reference`` – the reference's own CPU kernel, compiled unmodified (``oracle/_ref``).
* ``oracle.port``      – a plain-C restatement of the same algorithm (``oracle/ann_oracle.c``).

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import
this package.  The product (``ann_force_b200``) never does.
"""
