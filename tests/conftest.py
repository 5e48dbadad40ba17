import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"          # only present in the build container; never read by -m gpu tests


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


_bundle = None


def bundle():
    global _bundle
    if _bundle is None:
        _bundle = json.load(open(os.path.join(GOLDEN, "codes.json")))
    return _bundle


def get_code_dict(name):
    """code dict (oracle.pyoracle format) of a bundled code or the H_bit image"""
    from oracle import pyoracle as po
    b = bundle()
    if name in b["codes"]:
        c = b["codes"][name]
        return dict(N=c["N"], M=c["M"], q=c["q"], vdeg=[len(l) for l in c["var"]], cdeg=[len(l) for l in c["chk"]],
                    var=[[tuple(e) for e in l] for l in c["var"]], chk=[[tuple(e) for e in l] for l in c["chk"]])
    return po.hbit_to_code(b["h_bit"][name], 256)


def constellation(name):
    return np.array(bundle()["constellations"][name.replace(".txt", "")], np.float64)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def golden_names():
    """fixtures of the iterative decoders (the OSD fixtures f3_* have their own tests)"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith("f3_"))


def golden_osd_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith("f3_"))


def golden_lch(z, meta, frames=None):
    """channel LLRs of the golden frames, rebuilt with the oracle demapper (bit-exact with the reference)"""
    from oracle import pyoracle as po
    F = len(z["ret"]) if frames is None else frames
    q, N = meta["q"], meta["N"]
    out = np.zeros((F, N, q - 1))
    for f in range(F):
        if meta["nqam"] == 2:
            out[f] = po.demap_bpsk(q, N, z["rx"][f, :, 0], meta["sigma"])
        else:
            out[f] = po.demap_qam(q, N, z["rx"][f], constellation(meta["constellation"]), meta["sigma"])
    return out


@pytest.fixture(scope="session")
def code_files(tmp_path_factory):
    """every bundled code re-emitted in the reference's text format (exercises the product's file loader)"""
    from oracle import pyoracle as po
    d = tmp_path_factory.mktemp("codes")
    paths = {}
    for name in bundle()["codes"]:
        p = d / (name + ".txt")
        p.write_text(po.code_text(get_code_dict(name)))
        paths[name] = str(p)
    rows = bundle()["h_bit"]["H_bit_GF256"]
    p = d / "H_bit_GF256.txt"
    p.write_text("".join("\t".join(r) + "\t\n" for r in rows))
    paths["H_bit_GF256"] = str(p)
    return paths


def load_product_code(code_files, name):
    from nbldpc_b200 import capi
    if name.startswith("H_bit"):
        return capi.Code.from_hbit(code_files[name], 256)
    return capi.Code.from_file(code_files[name])
