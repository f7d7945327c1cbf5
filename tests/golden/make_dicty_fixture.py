"""Build tests/golden/dicty_matrices.npz from the reference's dicty data files (container only).

The three matrices of BASELINE config C2 (skfusion/datasets/base.py:45-61): gene x GO-term annotations
(binary), gene x condition expression (log of max(x, eps)), gene x gene protein interactions (constraint
matrix, entries in [-0.1, 0]).  Only derived numeric data is stored (float32 / index lists), no source."""
import gzip
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = "/root/reference/skfusion/datasets/data/dicty"


def read_matrix(name):
    with gzip.open(os.path.join(DATA, name)) as fh:
        next(fh)
        next(fh)
        return np.genfromtxt(fh, delimiter=",", missing_values=[""], filling_values="0")


def main():
    ann = read_matrix("dicty.gene_annnotations.csv.gz")
    expr = read_matrix("dicty.gene_expression.csv.gz")
    expr = np.log(np.maximum(expr, np.finfo(float).eps))
    ppi = read_matrix("dicty.ppi.csv.gz")
    assert ann.shape == (1219, 116) and expr.shape == (1219, 282) and ppi.shape == (1219, 1219)
    assert set(np.unique(ann)) <= {0.0, 1.0}
    rows, cols = np.nonzero(ppi)
    out = os.path.join(HERE, "dicty_matrices.npz")
    np.savez_compressed(out, ann_bits=np.packbits(ann.astype(np.uint8), axis=1), ann_shape=np.array(ann.shape),
                        expr=expr.astype(np.float32), ppi_rows=rows.astype(np.int32), ppi_cols=cols.astype(np.int32),
                        ppi_vals=ppi[rows, cols].astype(np.float32), ppi_shape=np.array(ppi.shape))
    print("wrote %s (%.1f KB); ppi nnz=%d range [%.3f, %.3f]" % (out, os.path.getsize(out) / 1024., len(rows), ppi.min(), ppi.max()))


if __name__ == "__main__":
    main()
