"""CPU oracle for the Flamed-TTS hot path - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import anything from this package, and only as the checker
or as the timed CPU baseline.  The product (`flamed/`, `flamed_tts_b200/`) never does.

Parity is PINNED: oracle/make_golden.py ran the unmodified reference
(/root/reference, imported in the build container with stub modules for its missing
non-hot-path dependencies) next to this restatement on identical weights, inputs and
noise, asserted agreement, and committed the reference outputs under tests/golden/.
"""
