"""Reader of oracle/ref_build/raw_dump.f90's files (test infrastructure)."""
import numpy as np


def read_raw(path):
    """-> (t, {tile_id: {'u': (nY, nX, 13), 'b0': (nvy, nX+1), 'bt': ..., 'maxima': (5, 2, nY, nX), 'tfirst': (nY, nX)}})
    in the array conventions of kestrel_b200.capi.Stepper.download_tile (C order = Fortran order reversed)."""
    raw = open(path, "rb").read()
    n_active, nX, nY, one_d = np.frombuffer(raw, dtype="<i4", count=4, offset=0)
    t = float(np.frombuffer(raw, dtype="<f8", count=1, offset=16)[0])
    off = 24
    nvx, nvy = nX + 1, nY + 1
    tiles = {}
    for _ in range(int(n_active)):
        tid = int(np.frombuffer(raw, dtype="<i4", count=1, offset=off)[0]); off += 4

        def take(count, shape):
            nonlocal off
            a = np.frombuffer(raw, dtype="<f8", count=count, offset=off).reshape(shape).copy()
            off += 8 * count
            return a
        u = take(13 * nX * nY, (nY, nX, 13))
        b0 = take(nvx * nvy, (nvy, nvx))
        bt = take(nvx * nvy, (nvy, nvx))
        maxima = np.stack([take(2 * nX * nY, (2, nY, nX)) for _ in range(5)])
        tfirst = take(nX * nY, (nY, nX))
        if one_d:
            b0, bt = b0[:1], bt[:1]
        tiles[tid] = {"u": u, "b0": b0, "bt": bt, "maxima": maxima, "tfirst": tfirst}
    assert off == len(raw), (off, len(raw))
    return t, tiles
