"""PETSc binary files ([PETSc] PetscViewerBinary; read by $PETSC_DIR/lib/petsc/bin/PetscBinaryIO.py, which the reference's
c/ch5/plotTS.py:12,44-46 uses): a sequence of big-endian records, each starting with a 32-bit class id --

    Vec   1211214, int32 n, n float64            (VecView;  -ts_monitor_solution binary:u.dat)
    Real  1211213, float64                       ([PETSc] TSMonitorDefault on a binary viewer; -ts_monitor binary:t.dat)
    Mat   1211216, int32 M, N, nz, int32 row lengths [M], int32 column indices [nz], float64 values [nz]
                                                 (MatView of an AIJ matrix; `./tri -ksp_view_mat binary:A.dat`,
                                                  c/ch2/loadsolve.c:55-59 reads it back with MatLoad)

Writer for the hosts of pattern.c (c/ch5/MOVIES.md:44: `-ts_monitor binary:t.dat -ts_monitor_solution binary:u.dat`) and
a reader with PetscBinaryIO.readBinaryFile's result shape, so that plotTS.py-style post-processing works without PETSc.
32-bit indices, real double scalars (PETSc's default configuration)."""
import struct

import numpy as np

VEC_CLASSID, REAL_CLASSID, MAT_CLASSID = 1211214, 1211213, 1211216


def write_real(fh, value):
    fh.write(struct.pack(">id", REAL_CLASSID, float(value)))


def write_vec(fh, array):
    a = np.ascontiguousarray(array, dtype=np.float64).ravel()
    fh.write(struct.pack(">ii", VEC_CLASSID, a.size))
    fh.write(a.astype(">f8").tobytes())


def write_mat(fh, rowptr, colind, vals, ncols=None):
    """A CSR matrix as [PETSc] MatView writes an AIJ matrix (columns sorted within a row, as PETSc stores them)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    colind = np.asarray(colind, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float64)
    m = rowptr.size - 1
    fh.write(struct.pack(">iiii", MAT_CLASSID, m, m if ncols is None else int(ncols), int(rowptr[-1])))
    fh.write(np.diff(rowptr).astype(">i4").tobytes())
    fh.write(colind.astype(">i4").tobytes())
    fh.write(vals.astype(">f8").tobytes())


def read_file(path):
    """The objects of a PETSc binary file, in order: a float for a Real record, a 1-D float64 array for a Vec, and for a
    Mat ((M, N), (rowptr, colind, values)) -- PetscBinaryIO.readMatSparse's shape."""
    out = []
    with open(path, "rb") as fh:
        while True:
            head = fh.read(4)
            if not head:
                return out
            (cid,) = struct.unpack(">i", head)
            if cid == REAL_CLASSID:
                out.append(struct.unpack(">d", fh.read(8))[0])
            elif cid == VEC_CLASSID:
                (n,) = struct.unpack(">i", fh.read(4))
                out.append(np.frombuffer(fh.read(8 * n), dtype=">f8").astype(np.float64))
            elif cid == MAT_CLASSID:
                m, n, nz = struct.unpack(">iii", fh.read(12))
                lens = np.frombuffer(fh.read(4 * m), dtype=">i4").astype(np.int64)
                colind = np.frombuffer(fh.read(4 * nz), dtype=">i4").astype(np.int32)
                vals = np.frombuffer(fh.read(8 * nz), dtype=">f8").astype(np.float64)
                if lens.size != m or colind.size != nz or vals.size != nz or int(lens.sum()) != nz:
                    raise ValueError("%s: truncated or inconsistent Mat record" % path)
                rowptr = np.concatenate(([0], np.cumsum(lens))).astype(np.int32)
                out.append(((m, n), (rowptr, colind, vals)))
            else:
                raise ValueError("%s: class id %d is not a Vec, Real or Mat record" % (path, cid))
