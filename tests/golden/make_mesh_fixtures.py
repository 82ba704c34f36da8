"""Regenerate tests/golden/{env,rob}.npz from the reference's OBJ fixtures.

Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_mesh_fixtures.py
Semantics follow the reference's test OBJ loader (test/test_fcl_utility.h:194-280):
records whose first token is 'v' add a vertex, 'f' adds a (1-based) triangle, anything
else -- including the "6540 2180" header line of env.obj -- is ignored.
"""
import os
import sys

import numpy as np

SRC = "/root/reference/test/fcl_resources"
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_obj(path):
    verts, tris = [], []
    with open(path, "rb") as f:
        for raw in f:
            tok = raw.decode("ascii", "replace").split()
            if not tok or tok[0].startswith("#"):
                continue
            if tok[0] == "v":
                verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif tok[0] == "f":
                idx = [int(t.split("/")[0]) - 1 for t in tok[1:]]
                for t in range(len(idx) - 2):
                    tris.append([idx[0], idx[t + 1], idx[t + 2]])
    return np.asarray(verts, dtype=np.float64), np.asarray(tris, dtype=np.int32)


def main():
    for name in ("env", "rob"):
        v, t = parse_obj(os.path.join(SRC, name + ".obj"))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), verts=v, tris=t)
        print(name, v.shape, t.shape, v.min(0), v.max(0))


if __name__ == "__main__":
    sys.exit(main())
