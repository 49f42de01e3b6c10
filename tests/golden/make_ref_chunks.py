"""Golden fixture of the chunk layout: runs the unmodified `chunk_parameters` of /root/reference/bin/submit_mocks.py
(extracted from the script's source with ast, because importing the script needs absent wheels) for every box size it
knows and writes tests/golden/ref_chunks.json.  Run in the build container only (needs /root/reference)."""
import ast
import json
import os

import numpy as np

SRC = "/root/reference/bin/submit_mocks.py"
tree = ast.parse(open(SRC).read())
fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "chunk_parameters"][0]
ns = {"np": np}
exec(compile(ast.Module(body=[fn], type_ignores=[]), SRC, "exec"), ns)
out = {}
for cells in (2560, 256, 512, 1024, 128, 32):
    for stripe in (False, True):
        ra0, dra, dec0, ddec, cid, nslice = ns["chunk_parameters"](cells, stripe)
        out["%d_%d" % (cells, int(stripe))] = {"ra0": list(ra0), "dra": list(dra), "dec0": list(dec0),
                                               "ddec": list(ddec), "chunkid": list(cid), "nslice": int(nslice)}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_chunks.json"), "w"), indent=1)
print("wrote", len(out), "layouts")
