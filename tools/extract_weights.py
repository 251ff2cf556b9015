#!/usr/bin/env python3
"""Extract the generated-model weight tables and known-answer vectors of the reference into flat
little-endian float32 blobs.

The reference ships its networks as hex byte tables inside generated C++
(models/generated/modelm_befe75da.cpp:16-1761, modelc_{5c241121,01266c1b,b00bf70c}.cpp:22-1821,
models/expiry/modelc_bf4dd6c8.cpp).  The numbers are *data* the replacement must reproduce bit for
bit; this script converts them, it does not copy any reference code.  Run in the build container
(needs /root/reference); outputs are committed:

  card.io-dmz_b200/weights/<model>.bin   concatenated tensors, order listed in <model>.json
  tests/golden/kat_<model>.bin            the reference's embedded KAT input/output vectors
"""
import json, os, re, struct, sys
import numpy as np

REF = os.environ.get("DMZ_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WDIR = os.path.join(ROOT, "card.io-dmz_b200", "weights")
GDIR = os.path.join(ROOT, "tests", "golden")

TABLE = re.compile(r"static uint8_t (data_\w+)\[(\d+)\][^=]*=\s*\{\s*//\s*(.*?)\n(.*?)\};", re.S)


def tables(path):
    src = open(path).read()
    out = []
    for name, n, label, body in TABLE.findall(src):
        vals = bytes(int(x, 16) for x in re.findall(r"0x([0-9A-Fa-f]{2})", body))
        assert len(vals) == int(n), (name, len(vals), n)
        out.append((name, label.strip(), np.frombuffer(vals, dtype="<f4").copy()))
    return out


def dump(model, relpath):
    tabs = tables(os.path.join(REF, relpath))
    weights = [(n, l, a) for (n, l, a) in tabs if not l.startswith("test")]
    kats = [(n, l, a) for (n, l, a) in tabs if l.startswith("test")]
    meta, off = [], 0
    with open(os.path.join(WDIR, model + ".bin"), "wb") as f:
        for n, l, a in weights:
            f.write(a.astype("<f4").tobytes())
            meta.append({"table": n, "label": l, "offset": off, "count": int(a.size)})
            off += int(a.size)
    json.dump({"source": relpath, "tensors": meta, "total_floats": off},
              open(os.path.join(WDIR, model + ".json"), "w"), indent=1)
    kmeta, off = [], 0
    with open(os.path.join(GDIR, "kat_" + model + ".bin"), "wb") as f:
        for n, l, a in kats:
            f.write(a.astype("<f4").tobytes())
            kmeta.append({"table": n, "label": l, "offset": off, "count": int(a.size)})
            off += int(a.size)
    json.dump({"source": relpath, "vectors": kmeta, "tolerance_abs": 1e-5},
              open(os.path.join(GDIR, "kat_" + model + ".json"), "w"), indent=1)
    print(model, [(m["label"], m["count"]) for m in meta], [(m["label"], m["count"]) for m in kmeta])


if __name__ == "__main__":
    os.makedirs(WDIR, exist_ok=True); os.makedirs(GDIR, exist_ok=True)
    dump("modelm_befe75da", "models/generated/modelm_befe75da.cpp")
    for m in ("5c241121", "01266c1b", "b00bf70c"):
        dump("modelc_" + m, "models/generated/modelc_%s.cpp" % m)
    dump("modelc_bf4dd6c8", "models/expiry/modelc_bf4dd6c8.cpp")
    dump("modelm_730c4cbd", "models/expiry/modelm_730c4cbd.cpp")
