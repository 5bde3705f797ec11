"""CPU oracle for the SOAP post-processing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU arm), never as the thing that produces results for
the CUDA path.  ``cpctools_b200`` never imports this package.
"""
