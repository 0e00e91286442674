"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the DiBS SVGD hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU
baseline.  ``dibs_b200`` never imports from here and has no CPU fallback.

Parity pin status (see DESIGN.md "Oracle"):
  * the reference (larslorch/dibs @ 5350d1a) ships no tests and no golden
    vectors, and JAX is not installable in this image, so the reference cannot
    be executed on its own backend here ("parity unpinned" by the reference's
    own artefacts);
  * what pins this oracle instead: (1) Threefry-2x32 / JAX PRNG known-answer
    vectors (Random123 KATs and the values printed in the JAX documentation),
    (2) golden fixtures under ``tests/golden/`` produced by executing the
    UNMODIFIED reference sources from ``/root/reference/dibs`` on top of
    ``oracle/jaxshim`` (a minimal torch-CPU implementation of the ``jax`` API
    subset the reference uses) with ``oracle/gen_golden.py``.
"""
