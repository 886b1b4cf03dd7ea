"""CPU oracle for the MLD latent-diffusion sampling path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as the
timed CPU baseline), never as the path that is measured or shipped.  The product
(``mld_b200``) fails loudly when its CUDA library is missing; it never routes here.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the reference's own
``nn.Module``s from ``/root/reference`` (in the build container) and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this restatement
against those fixtures.  The one third-party piece that is absent from the reference
tree is ``diffusers`` (unpinned, ``requirements.txt:23``): the DDIM/DDPM restatement
follows the published update rule and is anchored on the reference's call sites
(``mld/models/modeltype/mld.py:81,310-320,345``) plus known-answer vectors
(``tests/test_scheduler.py``).
"""
