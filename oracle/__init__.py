"""CPU oracle for the SemiSegECG training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or the reported CPU baseline -- never as the thing shipped.

``segnet_oracle`` is a restatement (written from the reference's behaviour,
not copied from it) of the reference's algorithm for this path in plain
PyTorch CPU arithmetic (fp32 or fp64).  All arithmetic of the reference lives
in PyTorch (pinned ``torch==1.11.0+cu113`` in the reference's
requirements.txt:8; torch 2.11.0 in this image); the oracle uses
``torch.nn.functional.conv1d`` for the contraction and spells every other
piece (BatchNorm, max-pool, linear upsample, soft-max, cross-entropy variants,
AdamW, EMA, LR schedule) out by formula.

Parity pinning: the reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the
reference itself, produced in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference/src``) and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks
the oracle against those fixtures on every run.
"""
