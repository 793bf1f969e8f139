"""CPU oracle for the BPR hot path of 0411tony/Yue.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker or the timed CPU baseline.  Nothing
under ``yue_b200/`` imports it; the product path fails loudly when the CUDA
library is missing instead of falling back to anything in here.

What it restates (file:line relative to the reference tree):

* ``recommender/cf/BPR.py:31-62``      per-triplet SGD epoch (the commented numpy loop)
* ``tool/qmath.py:115-116``            sigmoid
* ``base/IterativeRecommender.py:36-39``   P/Q initialisation (U[0,0.1) float32)
* ``base/IterativeRecommender.py:47-75``   learning-rate schedule / convergence test
* ``base/IterativeRecommender.py:58-60``   predict = Q.dot(P[u])
* ``base/IterativeRecommender.py:77-145``  masking + top-N (exact) and the lossy
                                           selection actually shipped (``ref_quirk``)
* ``evaluation/measure.py:6-101``      hits / precision / recall / F1 / MAP / coverage
* ``data/record.py:138-202``           id assignment and test-set semantics
* ``recommender/advanced/APR.py:25-76,95-137``  adversarial BPR losses and gradients (``apr_ref.py``; its closed forms pinned by the
                                        reference's own graph evaluated over ``tf1_shim.py`` on single triplets: ``make_golden_apr.py``)
* ``recommender/cf/WRMF.py:17-88``      implicit-feedback ALS (``wrmf_ref.py``; pinned by the reference CLASS itself,
                                        ``make_golden_wrmf.py`` -> ``tests/golden/wrmf_small.npz``)
* ``recommender/advanced/CUNE.py:118-178``  two-level BPR training loop (``cune_ref.py``; pinned by the reference's loop
                                        text, ``make_golden_cune.py``)

* ``recommender/advanced/LightGCN.py:15-105``  LightGCN's graph, loss, Adam loop (``lightgcn_ref.py``; pinned by the reference's own
                                        two files run unmodified over ``tf1_shim.py`` -- a stand-in for their TensorFlow-1 calls on
                                        torch autograd -- ``make_golden_lightgcn.py`` -> ``tests/golden/lightgcn_small.npz``; the
                                        meaning of the TF ops themselves is restated from TF's documentation)

Parity pinning.  The reference ships no tests, golden vectors or data, and its
RNG streams (CPython Mersenne Twister, unseeded) cannot be reproduced by a GPU
sampler, so the *sampler stream* is defined by this build (Philox4x32-10, see
``philox.py``) and is "parity unpinned" by the reference.  Everything downstream
of the sampler IS pinned: ``oracle/make_golden.py`` executes the reference's own
code here (``data.record.Record``, ``base.IterativeRecommender`` incl. its
``evalRanking``, ``evaluation.measure.Measure``, ``tool.config``, ``tool.qmath``
and the SGD loop text of ``BPR.py:31-62`` exec'd from the reference file with
``random.choice`` replaced by the Philox stream) and commits the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every oracle function
against them.
"""
