"""CONTOUR2 camera files (K + Euler angles in degrees + translation; image/camera.cpp:116-131, quat2proj :241-261): the library's
pmk_contour2_to_projection -- what the host mirror's PhotoSet::init goes through -- must produce the reference's level-0 projection
bit for bit.  The reference is loaded in a subprocess (it holds one scene per process).  CPU only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys, tempfile
import numpy as np
sys.path.insert(0, %(root)r)
from mvskit_b200 import synth
from oracle import pyoracle
scene = synth.make_scene(1, scale=0.25).render()
prefix = synth.write_scene(scene, tempfile.mkdtemp(prefix="pm_contour2_"))
cams = json.loads(%(cams)r)
for v, (intr, extr) in enumerate(cams):
    with open(prefix + "txt/%%08d.txt" %% v, "w") as fh:
        fh.write("CONTOUR2\n" + " ".join(repr(float(np.float32(x))) for x in intr) + "\n" + " ".join(repr(float(np.float32(x))) for x in extr) + "\n")
ref = pyoracle.RefLib(prefix)
out = [np.asarray(ref.camera(v, 0)["P"], np.float32).reshape(-1).view(np.uint32).tolist() for v in range(len(cams))]
print("RESULT " + json.dumps(out))
'''


def test_contour2_projection_matches_the_reference():
    from oracle import pyoracle
    from mvskit_b200 import pmk
    if not os.path.exists(pyoracle.REF_SO):
        if os.path.isdir("/root/reference/pmmvps"):
            pyoracle.build(ref=True)
        else:
            pytest.skip("oracle/_ref/libpmref.so not built and /root/reference absent")
    rng = np.random.RandomState(3)
    cams = []
    for v in range(5):
        intr = [765.7 + 10 * v, 760.1 - 7 * v, 0.5 * v, 160.0 + 3 * v, 120.0 - 2 * v, 0.0]
        extr = [float(x) for x in rng.uniform(-170, 170, 3)] + [float(x) for x in rng.uniform(-2, 2, 2)] + [3.0 + 0.3 * v]
        cams.append((intr, extr))
    res = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, cams=json.dumps(cams))], capture_output=True, text=True, timeout=600)
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT ")]
    assert line, res.stdout[-2000:] + res.stderr[-2000:]
    want = json.loads(line[0][7:])
    for v, (intr, extr) in enumerate(cams):
        got = pmk.contour2_to_projection(np.float32(intr), np.float32(extr)).reshape(-1).view(np.uint32).tolist()
        assert got == want[v], (v, got, want[v])
