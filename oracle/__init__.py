"""CPU oracle for the TVO truncated E-step / M-step hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain numpy (+ a few lines of C for
the dedup triple loop) restatement of the reference algorithm
(tvlearn/tvo, files cited function by function).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product (``tvo_b200``) never does.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks it against
 * the reference's own inline known-answer vectors
   (test/variationalutils_test.py:51-102, test/bsc_test.py:23-84,
   test/noisyor_test.py:17-50), and
 * fixtures produced by importing the unmodified reference in the build
   container (``tools/make_golden.py`` -> ``tests/golden/*.npz``).

The one place where the oracle *defines* rather than restates behaviour is the
random stream: the reference draws from torch's / numpy's global generators
(evo.py:263,325,363,414,440), which cannot be reproduced on a GPU.  The oracle
instead specifies a counter-based Philox4x32-7 stream (``oracle.philox``) and
consumes it with exactly the reference's sampling semantics (uniform parent
subsets, one distinct bit per child, independent per-bit flips with p0/p1, ...).
The CUDA kernels implement the same stream, so candidate sets are bit-identical
between oracle and product; agreement with the reference's operators is
distributional and is tested statistically.
"""
