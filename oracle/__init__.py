"""CPU oracle for the qibotn hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The
product (``qibotn_b200``) never does; it fails loudly when its CUDA library is
missing instead of falling back to anything here.

What it restates (numpy, complex128 unless asked otherwise), each function
citing the reference ``file:line`` it follows:

* ``oracle.network``      -- circuit -> interleaved einsum operands
                             (``circuit_convertor.py``) and observable handling
                             (``eval.py:15-150``);
* ``oracle.contract``     -- pairwise ``numpy.tensordot`` contraction of such a
                             network, which is what quimb/cotengra dispatch to
                             for numpy arrays (quimb 1.13.0 / cotengra 0.7.5,
                             pinned in the reference's ``poetry.lock``; neither
                             is vendored nor installable here);
* ``oracle.mps``          -- the MPS gate loop of ``mps_utils.py`` with
                             ``numpy.linalg.svd`` and cuQuantum
                             ``contract_decompose`` truncation semantics
                             (cuquantum-python 25.11.1, likewise absent);
* ``oracle.statevector``  -- an independent dense simulator used to pin the
                             restatements above.

PARITY PINNING.  The reference ships no golden vectors and none of its
numerical dependencies (qibo, quimb, cotengra, cuQuantum) can be imported or
installed in this image, so the oracle cannot be checked against the
reference *running*.  It is pinned instead against every value the
reference's own tests pin, all of which have closed forms:
QFT(n) applied to a state equals ``ifft(state, norm="ortho")``
(``tests/test_quimb_backend.py:23-66``); QFT(n)|0..0> is the uniform vector
(``tests/test_cuquantum_cutensor_backend.py:46-142``); <sum 0.5 X_i Z_{i+1}>
on it is 0 (``:145-199``).  Beyond those, parity with the third-party engines
is UNPINNED; the oracle is additionally cross-checked against the independent
dense simulator on random circuits (``tests/test_oracle.py``).
"""
