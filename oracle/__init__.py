"""CPU oracle for the RHCCQ encoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the thing timed for the CPU baseline).  The product package
``roibasedimagecompression_b200`` never imports this package and raises when
its CUDA library is missing.

Contents
--------
``rhccq_oracle``   numpy restatement of the reference quantiser
                   (encoder/compression/{clustering,merging,subregions,regions,
                   image}.py) — every function cites the lines it follows.
``kmeans_restated`` the exact-arithmetic restatement of scikit-learn's
                   KMeans(k, random_state=42, n_init='auto') that the CUDA
                   k-means kernel must match bit for bit.
``minibatch_restated`` the same for MiniBatchKMeans (the >= 10 000-colour branch).

Large point sets: ``rhccq_oracle.dbscan_labels`` is a brute-force restatement of
scikit-learn's DBSCAN for up to ~20 000 points; beyond that the tests compare with
scikit-learn itself (262 144 and 1 048 576 points on the GPU box) and check
size-independent properties (determinism, numbering by lowest core index, brute
force inside sampled windows, strips == unsplit).

Parity pinning: the restatement is checked in ``tests/`` against (1) golden
vectors produced by importing the reference's own modules
(``tests/golden/make_golden.py``, run once in the build container where
``/root/reference`` is mounted) and (2) scikit-learn 1.9.0 itself, which is the
third-party code the reference calls (requirements.txt:6, unpinned).
"""
