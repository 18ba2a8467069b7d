"""CPU oracle for the UltraRE sharded-retraining hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ultrare_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU arm -- never as the thing shipped.

Every function restates one piece of the reference
(``/root/reference`` = ZhangYizhao/UltraRE) in NumPy / PyTorch-CPU and cites the
file:line it follows.  Pinning status (see DESIGN.md "Oracle"):

* MF step / epoch, ensemble score, RMSE / HR / NDCG, SISA merge + routing,
  group re-ordering: PINNED against the reference itself, executed in the dev
  container by ``oracle/ref_shim.py`` (shims of SURVEY.md Appendix A) on the
  reference's own ``data/toy`` files; fixtures under ``tests/golden/`` were
  written by ``oracle/make_golden.py``.
* Deletion set / centroid seeds: PINNED (pure NumPy legacy-RNG known answers).
* Transport plan: the reference calls POT 0.9.0 ``ot.emd`` (exact network
  simplex), a third-party dependency that is absent from /root/reference and
  not installable offline.  Its published semantics are restated as an exact LP
  (SciPy HiGHS) in ``oracle/ot.py``; there is no golden vector for it in the
  reference => that piece is "parity unpinned".  The Sinkhorn restatement is the
  float64 arithmetic the CUDA kernels are checked against.
"""
