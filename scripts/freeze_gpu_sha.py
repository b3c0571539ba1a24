"""Freeze the SHA-256 of the bytes the CUDA encoder produces for small versions of the five BASELINE configs
(tests/test_real_reference.py:_inputs) at levels 0 / 5 / 8.  Run on a B200:

    python scripts/freeze_gpu_sha.py gpurun_out/gpu_encoded_sha256.json      # then copy into tests/golden/

tests/test_gpu_parity.py::test_encoded_bytes_match_frozen_sha256 compares later builds with it: the encoder is
deterministic, so any change of these hashes is a deliberate change of the encoder's decisions (re-freeze, say
why in the commit) or a bug."""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as g

g.build()
import flacarray_b200 as fa
from oracle.small_configs import small_configs as _inputs

out = {}
for name, x in _inputs().items():
    for level in (0, 5, 8):
        comp, starts, nbytes, _, _ = fa.array_compress(x, level=level)
        out[f"{name}/L{level}"] = {"sha256": hashlib.sha256(np.asarray(comp).tobytes()).hexdigest(), "nbytes": int(np.asarray(comp).size)}
json.dump(out, open(sys.argv[1], "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
