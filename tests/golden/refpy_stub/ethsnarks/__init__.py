"""TEST INFRASTRUCTURE ONLY -- stand-in for the third-party `ethsnarks` package (absent from this image), just large
enough to import the reference's own Python prototype (/root/reference/python/{poly,sumcheck,gkr}.py, unmodified) and run
its prover when golden vectors are generated (tests/golden/make_refpy_vectors.py).  Two choices are made here, and
only here, so that the prototype's transcript lines up with the Rust prover's (rust/src/gkr/prover.rs:10, :74-78):
`mimc.mimc_hash` is MiMC7-91 `multi_hash(message, 0)` as in `mimc-rs`, and `field.FQ.random()` returns zero (the Rust
prover starts from z_0 = 0; the prototype draws it at random, python/gkr.py:142-143).  Never imported by the product."""
