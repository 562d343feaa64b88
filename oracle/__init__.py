"""CPU oracle for the EMO-Disentanger hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU (torch fp32/fp64 + numpy) restatement of the
reference algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker or the
timed CPU baseline -- never as part of the product path (``emo_disentanger_b200``), which
fails loudly when ``libemo_b200.so`` is missing.

Pinning status (see DESIGN.md "Oracle"):
  * stage-1 ``PlainTransformer`` restatement (``txl_oracle``): PINNED -- checked against the
    reference module imported from /root/reference (``oracle/validate_against_reference.py``).
  * stage-2 ``MusicGPT2`` restatement (``gpt2_oracle``): PINNED against the reference module +
    HF transformers 5.5.0 ``GPT2Block`` (tuple shim); the README-pinned 4.28.0 is not
    installable offline.
  * sampling (``sampling_oracle``): PINNED against the reference ``temperature``/``nucleus``.
  * stage-2 ``MusicPerformer`` (``performer_oracle``): the model glue is pinned against the
    reference ``music_performer.py`` imported over ``oracle/standin/fast_transformers``; the
    third-party ``fast_transformers`` math itself (package absent, un-pinned upstream 0.4.0)
    is restated from its published algorithm -> **parity unpinned** for rows A3-A6.
"""
