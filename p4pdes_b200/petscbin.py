"""PETSc binary files ([PETSc] PetscViewerBinary; read by $PETSC_DIR/lib/petsc/bin/PetscBinaryIO.py, which the reference's
c/ch5/plotTS.py:12,44-46 uses): a sequence of big-endian records, each starting with a 32-bit class id --

    Vec   1211214, int32 n, n float64            (VecView;  -ts_monitor_solution binary:u.dat)
    Real  1211213, float64                       ([PETSc] TSMonitorDefault on a binary viewer; -ts_monitor binary:t.dat)

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


def read_file(path):
    """The objects of a PETSc binary file, in order: a float for a Real record, a 1-D float64 array for a Vec."""
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
            else:
                raise ValueError("%s: class id %d is not a Vec or Real record" % (path, cid))
