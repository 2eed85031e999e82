"""``mfrec.lib.datasets``: the boolean sparse structures ALS-WRMF walks
(reference: mfrec/lib/datasets.py:13-32).  ``row`` is ``[0, count_0, count_1, ...]`` -- counts, not
offsets -- over the rows up to the last non-empty one; ``col`` lists the neighbour ids row by row."""
import numpy as np


def create_bool_sparse_row(sparse_matrix):
    rows, cols = sparse_matrix.nonzero()
    order = np.lexsort((cols, rows))          # scipy's nonzero() order for lil / csr: by row, then column
    count = np.bincount(rows.astype(np.int64)).astype(np.int32)
    return np.r_[0, count].astype(np.int32), np.ascontiguousarray(cols[order], dtype=np.int32)


def create_bool_sparse_col(sparse_matrix):
    return create_bool_sparse_row(sparse_matrix.T)
