#!/usr/bin/env python3
"""Extract the reference's golden vectors into small JSON fixtures.

Run ONCE in the build container (needs /root/reference, which does not exist on the
GPU box); the JSON it writes is committed and is what the tests read.

The reference stores its regression data as JLD2 (HDF5-flavoured) files written by
`check_jld2(...; OVERWRITE_JLD2=true)` (reference test/test_utils.jl:24-25) and read back at
test/test_utils.jl:27-29.  No HDF5/JLD2 reader is installed here, so the payloads are read as
raw little-endian FP64 blocks at the byte offsets established in SURVEY.md section 4
(stress block: 6 doubles in Tensors.jl storage order 11,21,31,22,32,33, stride 95 bytes;
tangent block: 36 doubles = 6x6 column-major, row = ij, col = kl, stride 336 bytes).

Only the three files the reference's tests actually load are extracted
(test_linear_elastic.jl:42-46, test_plastic.jl:44-49, test_crystal_visco_plastic.jl:44-51);
NeoHook/Yeoh/StVenant.jld2 are orphaned (no test reads them, parameters unknown).
"""
import json
import os
import struct
import sys

REF = os.environ.get("MM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

FILES = {
    # name: (stress offset, tangent offset, number of load steps)
    "LinearElastic1": (5839, 8352, 2),
    "Plastic1": (5919, 9462, 12),
    "CrystalViscoPlastic1": (5855, 8574, 4),
}


def doubles(buf, off, n):
    return list(struct.unpack_from("<%dd" % n, buf, off))


def main():
    out = {}
    for name, (s_off, t_off, nsteps) in FILES.items():
        path = os.path.join(REF, "test", "jld2_files", name + ".jld2")
        buf = open(path, "rb").read()
        stresses = [doubles(buf, s_off + 95 * k, 6) for k in range(nsteps)]
        tangents = [doubles(buf, t_off + 336 * k, 36) for k in range(nsteps)]
        out[name] = {
            "source": "test/jld2_files/%s.jld2" % name,
            "stress_order": "11,21,31,22,32,33",
            "tangent_order": "36 = I + 6*J, I<->(ij), J<->(kl), same pair order; raw tensor components",
            "stresses": stresses,
            "tangents": tangents,
        }
    dst = os.path.join(HERE, "reference_golden.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    sys.exit(main())
