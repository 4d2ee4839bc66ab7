"""CPU oracle for the EaseVoice stage-2 hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain torch-fp32 / numpy restatement of the reference's
algorithms (each function cites the reference file:line it follows).  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it -- as the checker or the timed
CPU baseline, never as part of the product path.  The product
(``easevoice-trainer_b200``) never imports ``oracle``.

Pinning: the reference's own tests hold no vectors for this path (SURVEY.md
section 4), so the oracle is pinned against the reference itself executed in
the authoring container: ``oracle/pin_against_reference.py`` imports
``/root/reference`` (CPU), checks every oracle function against the
corresponding reference function on seeded inputs, and writes the outputs to
``tests/golden/`` so the pin travels to machines where the reference is absent.
"""
