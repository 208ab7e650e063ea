"""CPU oracle for the T2ONet operator / planner hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as
the thing timed for the CPU baseline), never as a fallback for the CUDA path.

What it is: a plain PyTorch-CPU restatement (fp32, autograd-capable) of the
reference's algorithm for the path named by BASELINE.json's north_star:

* ``oracle.hsv``      kornia's RGB<->HSV (third-party, un-vendored and un-pinned in
                      the reference: ``requirements.txt:5`` is a bare ``kornia``)
* ``oracle.ops``      models/operators.py process/execute, utils/operator_utils.py,
                      executors/executor.py dispatch
* ``oracle.planner``  utils/beam_search.py (get_dist, get_param*, execute, beam_search)
                      and the fixed-order / eps-greedy variants

Pinning status
--------------
The reference ships NO tests, golden vectors or fixtures for this path
(SURVEY.md section 4 / 8c).  The restatement is therefore pinned against the
reference ITSELF: ``oracle/ref_shims.py`` imports the unmodified modules from
``/root/reference`` (possible only in the authoring container) and
``oracle/make_golden.py`` records their outputs on seeded inputs into
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks the restatement
against those vectors.  One boundary stays unpinned: **kornia** itself is not
installed and cannot be installed (no network), so the RGB<->HSV arithmetic is
the published kornia 0.4/0.5 algorithm restated from its documentation
(``eps=1e-6`` variant) -- "parity unpinned" for that third-party boundary.
"""
