"""CPU oracle for the rollout hot path of jeremy-collins/sd-video-gen.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the timed
CPU baseline.  The product path (``sd-video-gen_b200/`` -> ``libsdvg.so``) never
imports this package and raises if the CUDA library is missing.

Parity pin: the reference ships no tests, golden vectors or checkpoints for this
path, so the oracle is pinned against the *reference
module itself*, imported unmodified from /root/reference in the build container by
``oracle/make_golden.py``; its outputs on seeded weights/inputs are committed under
``tests/golden/`` and ``tests/test_oracle_golden.py`` replays them against

  * ``oracle.ref_module.RefTransformer``  - restatement of models/transformer.py on
    top of ``torch.nn.Transformer`` (the reference's own third-party arithmetic,
    pinned pytorch=1.11.0 in environment.yml:93, 2.11.0 in this image), and
  * ``oracle.functional.forward``         - an independent op-by-op restatement
     that does not use ``nn.Transformer`` at all.
"""
