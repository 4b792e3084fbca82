"""CPU oracle for the rule-guided SCG sampling hot path -- TEST INFRASTRUCTURE ONLY.

A plain numpy / torch-fp32 restatement of the reference algorithm (yjhuangcd/rule-guided-music), each function citing
the reference file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this package, and only as the checker / reported baseline -- never as the product path, which is
``rule_guided_music_b200`` + ``librgm_b200.so`` and fails loudly without the CUDA library.

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned instead against
outputs of the UNMODIFIED reference imported in the build container (``tests/golden/make_golden.py`` writes
``tests/golden/*.npz``; ``tests/test_oracle_cpu.py`` checks the oracle against them).  Exceptions, stated here as
required: the chord-progression rule (music21 8.3.0 + mido 1.2.10, neither vendored nor installed) is NOT restated --
parity unpinned; ``timm.Mlp`` (0.9.2) and ``rotary_embedding_torch.RotaryEmbedding`` (0.3.2) are restated from their
published behaviour because their sources are not under /root/reference -- parity at those two
boundaries is pinned only through the reference's own call sites (dit.py:253-286, 326).
"""
