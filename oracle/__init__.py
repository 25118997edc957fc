"""CPU oracle for the Stable Audio hot path.  TEST INFRASTRUCTURE ONLY.

A plain torch-fp32 / numpy restatement of the reference algorithm for the DiT
denoiser, the Oobleck autoencoder, the multi-resolution STFT loss and the
samplers.  Every function cites the reference file:line it follows
(paths relative to /root/reference/stable_audio_tools).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` leg may import this package, and only as the checker or the
reported CPU baseline.  The product (`b200sat`) never imports it and fails
loudly when its CUDA library is missing.

Pinning: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference itself, imported in
the authoring container (`oracle/ref_harness.py`, `oracle/gen_golden.py`) and
committed as fixtures under `tests/golden/`.  The k-diffusion pieces
(`oracle/sampling.py`: dpmpp-3m-sde, VDenoiser, polyexponential sigmas) live in
the un-vendored dependency k-diffusion==0.1.1 which is absent here: for those
the status is "parity unpinned" (restated from the published algorithm,
checked by closed-form known-answer tests only).
"""
