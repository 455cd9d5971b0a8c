"""CPU oracle for the SR sampling hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / reported CPU baseline.  The product
package (``superresolutionhep_b200``) never imports it and raises when its CUDA extension
is missing.

Parity status: the network part (``sr_oracle``, ``pflow_oracle``) is pinned against the
reference's own modules imported from ``/root/reference`` (tests/test_oracle_vs_reference.py,
run in the build container) and against committed golden vectors minted from them
(tests/golden/).  The ODE driver (``odeint``) restates the third-party ``torchdiffeq``
(unpinned, absent from the reference tree and from this image): **parity unpinned** for
that piece -- see the header of ``oracle/odeint.py``.
"""
