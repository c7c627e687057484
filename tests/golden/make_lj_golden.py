"""Extract the reference's Lennard-Jones rollout fixture into a small .npy (no h5py here).

Source: ``/root/reference/tests/3D_LJ_3_1214every1/valid.h5`` -- dataset
``00000/position``, float32, shape (405, 3, 3), gzip-chunked with chunk (203, 3, 3), no
shuffle filter (HDF5 superblock v0).  h5py is not installed in this image, so the two
zlib streams are located by their headers, inflated and concatenated; the result is
truncated to the 405 frames ``metadata.json`` declares (``sequence_length_test``).
``particle_type`` is all FLUID (0) per ``metadata.json:52``.

Run in the build container (the reference tree does not exist on the GPU box):
    python tests/golden/make_lj_golden.py
Writes ``tests/golden/lj3d_valid_position.npy`` (T, N, d) and ``lj3d_metadata.json``.
"""

import json
import os
import shutil
import zlib

import numpy as np

SRC = "/root/reference/tests/3D_LJ_3_1214every1"
HERE = os.path.dirname(os.path.abspath(__file__))


def inflate_chunks(raw, min_len=100):
    chunks, i = [], 0
    while i < len(raw) - 2:
        if raw[i] == 0x78 and raw[i + 1] in (0x01, 0x5E, 0x9C, 0xDA):
            try:
                d = zlib.decompressobj()
                out = d.decompress(raw[i:])
                if len(out) > min_len:
                    chunks.append(out)
                    i += len(raw) - i - len(d.unused_data)
                    continue
            except zlib.error:
                pass
        i += 1
    return chunks


def main():
    with open(os.path.join(SRC, "metadata.json")) as f:
        meta = json.load(f)
    raw = open(os.path.join(SRC, "valid.h5"), "rb").read()
    chunks = inflate_chunks(raw)
    n, d, t = meta["num_particles_max"], meta["dim"], meta["sequence_length_test"]
    flat = np.frombuffer(b"".join(chunks), dtype="<f4")
    pos = flat[: t * n * d].reshape(t, n, d).copy()
    assert np.isfinite(pos).all() and pos.min() >= 0.0 and pos.max() <= 5.0
    np.save(os.path.join(HERE, "lj3d_valid_position.npy"), pos)
    shutil.copy(os.path.join(SRC, "metadata.json"), os.path.join(HERE, "lj3d_metadata.json"))
    print("wrote", pos.shape, pos.dtype)


if __name__ == "__main__":
    main()
