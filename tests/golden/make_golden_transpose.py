"""Golden vectors of the reference's host transpose (matrix_transposition, src/utils.cpp:20-46),
the step that forms B = A^T for its AAT mode.  Host code: runs in the build container without
a GPU, through oracle/_ref/libmhref.so (built by `make -C oracle ref` from the sources in place).

    python tests/golden/make_golden_transpose.py        # writes tests/golden/ref_transpose_*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases  # noqa: E402


def main():
    from oracle import Reference
    R = Reference()
    for name, make in cases.TRANSPOSE.items():
        A = make()
        T = R.transpose(A)
        np.savez_compressed(os.path.join(HERE, f"ref_transpose_{name}.npz"), ptr=T.ptr, col=T.col, val=T.val)
        print(name, A.M, A.N, A.nnz)


if __name__ == "__main__":
    main()
